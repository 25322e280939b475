// encoder.cuh -- the post-norm transformer encoder stack shared by MF-MDM G (denoiser.cu) and R (refiner.cu), plus the
// small fp32 helpers of the once-per-sample conditioning path.
//
// Reference: nn.TransformerEncoder of 8 x nn.TransformerEncoderLayer(d, nhead, ff, activation="gelu"), post-norm,
// batch_first=False, no mask -- constructed at src/oakink2_tamf/model/interaction_segment_mdm.py:63-70 (G) and
// src/oakink2_tamf/model/segment_refine_model.py:88-95 (R).
//
// Per layer, 5 kernels over the token matrix (row = b*S + s):
//   QKV = Xb . Win^T + b                    tcgen05 GEMM, bf16 out           [M,3d]
//   ATT = softmax(Q K^T / sqrt(hd)) V       fused attention (attn_tc.cuh)       [M,d]
//   X   = LN1(X + ATT . Wo^T + b)           tcgen05 GEMM, residual+LayerNorm epilogue (X = Xb + Xlo, two bf16 planes)
//   H   = gelu(Xb . W1^T + b)               tcgen05 GEMM, exact-erf GELU epilogue, bf16 out  [M,ff]
//   X   = LN2(X + H . W2^T + b)             tcgen05 GEMM, residual+LayerNorm epilogue
#pragma once
#include <vector>

#include "common.cuh"

namespace tamf {

// Device allocations owned by a handle (weights converted at create time; nothing is allocated in the hot path).
struct DevPool {
  std::vector<void*> owned;
  int alloc(void** p, size_t bytes);
  int upload_f32(float** dst, const float* src_host, size_t n);
  // fp32 host [rows, cols] -> bf16 device [rows, ld] (columns cols..ld-1 zero)
  int upload_bf16(__nv_bfloat16** dst, const float* src_host, int rows, int cols, int ld);
  void free_all();
};

struct LayerDev {
  __nv_bfloat16 *w_in, *w_out, *w1, *w2;
  float *b_in, *b_out, *b1, *b2, *g1, *be1, *g2, *be2;
  CUtensorMap tm_in, tm_out, tm_w1, tm_w2;
};

struct LayerWeightPtrs;  // layer_chain.cuh
struct EncoderStack {
  int d = 0, ff = 0, L = 0, H = 0;
  std::vector<LayerDev> layers;
  // stack form of the layer kernel (one launch for all layers): per-layer weight tensor maps [L][4] (out_proj, linear1,
  // linear2, in_proj of the NEXT layer) and parameter pointers [L] in device memory
  CUtensorMap* d_wmaps = nullptr;
  LayerWeightPtrs* d_lw = nullptr;
  int upload(DevPool& pool, const tamf_layer_weights* w, int d, int ff, int L, int H);
};

// Activation buffers of one bound (B, S) problem; all row-major, row = b*S + s.
struct EncoderBuffers {
  int B = 0, S = 0, M = 0;
  __nv_bfloat16* Xb = nullptr;    // residual stream, high plane bf16(x) [M,d] = the GEMM A operand
  __nv_bfloat16* Xlo = nullptr;   // residual stream, low plane bf16(x - Xb) [M,d]
  __nv_bfloat16* QKV = nullptr;   // [M,3d]
  __nv_bfloat16* ATT = nullptr;   // [M,d]
  __nv_bfloat16* Hb = nullptr;    // [M,ff]
  CUtensorMap tm_Xb, tm_ATT, tm_H;              // GEMM A-operand loads, box {64, 128} (tm_ATT / tm_H feed the LayerNorm
                                                // GEMMs: box {64, 32}, one per TMEM lane quarter)
  int ln_rq = 32;                               // rows per lane quarter of the LayerNorm tiles (gemm_ln_rq(M))
  CUtensorMap tm_QKV_st, tm_H_st;               // bf16 epilogue stores, box {64, 32}
  CUtensorMap tm_Xb_st, tm_Xlo;                 // LayerNorm epilogue: Xb / Xlo residual load + store, box {32, ln_rq} bf16
  CUtensorMap tm_att_kv, tm_att_o;    // attention: [B][S][3d] views of QKV (K/V box, Q box), [B][S][d] view of ATT
  // ---- layer kernel (layer_chain.cuh): out_proj+LN1 -> linear1+GELU -> linear2+LN2 -> next in_proj in one launch ----
  bool chain = false;                 // TAMF_CHAIN=0 keeps the five-kernel layer of round 1 (A/B comparisons)
  void* aux = nullptr;                // caller-owned (workspace): counters, row statistics, schedules
  CUtensorMap tm_ATT128, tm_H128;     // A operands of LN1 / LN2, box {64, 128}
  CUtensorMap tm_Xlo128;              // low residual plane, box {64, 128} (the high plane is tm_Xb)
  CUtensorMap tm_Xh_st, tm_Xl_st;     // residual planes, box {32, 32}: LayerNorm result stores (half-slab staging tiles)
  CUtensorMap tm_H_st32, tm_QKV_st32; // H / QKV stores of the layer kernel, box {32, 32}
  CUtensorMap tm_ident;               // 64 x 64 identity (layer_chain.cuh chain_identity_map)
  unsigned* ctr = nullptr;            // [6][tiles_m] row-tile counters + [B] sequence counters (layer_chain.cuh)
  size_t ctr_bytes = 0;
  unsigned long long* stats = nullptr;  // [2][tiles_m][halves][2][4][128] row statistics words
  int tiles_m = 0, halves = 0;
  int *sched = nullptr, *schedL = nullptr;  // [pairs + 1 offsets | unit codes]: with / without the next in_proj
  int pairs = 0, pairsL = 0;
  // stack form (TAMF_CHAIN=2): ONE persistent layer kernel on `pairsS` CTA pairs for all layers next to ONE persistent
  // attention kernel on `att_ctas` CTAs (a second stream inside the same graph); same counters
  bool stack = false;
  int* schedS = nullptr;
  int pairsS = 0, att_ctas = 0;
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int make_maps(int d, int ff, int layers = 0, int heads = 0);
  void release();  // the side stream / events of the stack form (handle destruction)
};
// bytes of EncoderBuffers::aux for an [M, d] problem (256-byte multiple)
size_t encoder_aux_bytes(int M, int d, int ff);

// Enqueue all L layers on `s`.  `marks` (profiling only): an event is recorded after every kernel.
// `ktime` (profiling only): in-graph timing slots, 4 int64 per kernel in launch order starting at slot *kidx.
int enqueue_encoder(const EncoderStack& enc, const EncoderBuffers& buf, cudaStream_t s,
                    std::vector<cudaEvent_t>* marks, long long* ktime = nullptr, int* kidx = nullptr);
// Clears the dependency counters of the layer kernels: enqueue ahead of the first kernel of every evaluation.
int encoder_begin_evaluation(const EncoderBuffers& buf, cudaStream_t s);
int configure_encoder_kernels();  // cudaFuncSetAttribute for every instantiation used above (once per process)

// ---- small fp32 helpers (conditioning path, once per sample) ----
// out[r,n] = post(sum_k in[r,k] W[n,k] + bias[n]); post: 0 none, 1 silu, 2 nan_to_num(.) + add[r,n]
int linear_f32(const float* in, int ld_in, const float* W, int ld_w, const float* bias, float* out, int ld_out, int R,
               int N, int K, int post, const float* add, int ld_add, cudaStream_t s);
// out[o,i] = mean_r in[o,r,i]
int mean_axis(const float* in, float* out, int outer, int red, int inner, cudaStream_t s);
// obj_traj [B,nobj,T,9] -> mean over the (padded) object axis, frame-major [B*T,9]
int traj_mean(const float* traj, float* out, int B, int nobj, int T, cudaStream_t s);
int fill_int(int* p, int n, int v, cudaStream_t s);

inline void mark_event(std::vector<cudaEvent_t>* marks, cudaStream_t s) {
  if (!marks) return;
  cudaEvent_t e;
  cudaEventCreate(&e);
  cudaEventRecord(e, s);
  marks->push_back(e);
}

}  // namespace tamf
