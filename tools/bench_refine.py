#!/usr/bin/env python
"""BASELINE.json configs[2]: MF-MDM R (arch_refine) sampling pass with ManoLayer FK and the hand->object NN query,
batch 64, T=160, 1 object x 8192 points on one B200 (SURVEY.md 8d "Config 3").

Prints ONE JSON line: sequences/s of the whole `SegmentRefineModel.forward` (the 13-key dict, 3 FK + 3 NN passes + the
transformer) plus the separate FK / normals / NN / transformer times, each with its achieved GB/s on the ALGORITHMIC
bytes of SURVEY.md 8d (FK 10 024 B/frame; NN 18.7 KB + 36 B per frame) against the measured HBM peak, and pairs/s for
the brute-force NN.  Timed with CUDA events on the launching stream after warm-up; L2 flushed between repetitions.

    python tools/bench_refine.py [--batch 64] [--reps 10] [--points 8192] [--nobj 1]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oakink2-tamf_b200"))
sys.path.insert(0, ROOT)


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6500.0, "fallback (B200_PROFILING.md)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--frames", type=int, default=160)
    ap.add_argument("--points", type=int, default=8192)
    ap.add_argument("--nobj", type=int, default=1)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    import tamf_b200
    from tamf_b200 import _lib, synth
    from tamf_b200.chamfer import h2o_dist
    from tamf_b200.refine import vertex_normals

    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    B, T, P, nobj = a.batch, a.frames, a.points, a.nobj
    cfg = synth.ARCH["arch_refine"]
    m = tamf_b200.SegmentRefineModel("unused", **cfg, use_pc=True,
                                     mano_assets={"right": synth.mano_assets("right"), "left": synth.mano_assets("left")})
    m.load_state_dict(synth.r_state_dict(cfg, 0), strict=False)
    m = m.eval().to(dev)
    batch = synth.make_batch(B, T, nobj=nobj, seed=1, npoints=P, with_pointcloud=True)
    batch = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    def timed(fn, reps=a.reps, warm=a.warmup):
        for _ in range(warm):
            fn()
        ms = []
        for _ in range(reps):
            flush.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        return float(np.mean(ms)), float(np.min(ms))

    n0 = _lib.lib().tamf_kernel_launch_count()
    full_ms, full_min = timed(lambda: m(batch))
    launches = (_lib.lib().tamf_kernel_launch_count() - n0) // (a.reps + a.warmup)

    # ---- pieces (one call each, the same C-ABI entries forward() uses) ----
    x_in = batch["sample_pose_repr"].float().contiguous()
    shape = batch["shape"].float().contiguous()
    traj = batch["obj_traj"].float().contiguous()
    emb = batch["obj_embedding"].float().contiguous()
    N = B * T
    layer = m.mano_layer_rh
    pose = x_in.view(N, -1)
    betas = shape.view(N, 10)
    verts = torch.empty((N, 778, 3), device=dev)
    joints = torch.empty((N, 21, 3), device=dev)
    ids = torch.arange(N, dtype=torch.int32, device=dev)
    L = _lib.lib()

    def fk():
        _lib.check(L.tamf_mano_fk_select(layer._handle(dev), _lib.POSE_REPR, _lib.ptr(pose), _lib.ptr(betas), _lib.ptr(ids),
                                         N, _lib.ptr(verts), _lib.ptr(joints), _lib.stream_ptr(dev)), "fk")

    fk_ms, fk_min = timed(fk)
    nrm_ms, nrm_min = timed(lambda: vertex_normals(verts, layer.th_faces))
    hv = verts.view(B, T, 778, 3)
    pts = [np.asarray(o, np.float32)[: len(l)] for o, l in zip(batch["obj_pointcloud"], batch["obj_list"])]
    # the Python wrapper re-uploads the canonical clouds on every call (host list API of the reference); time the
    # C-ABI entry alone too, with the clouds resident
    nn_api_ms, _ = timed(lambda: h2o_dist(hv, traj, pts))
    first = [0]
    for o in pts:
        first.append(first[-1] + int(o.shape[0]))
    pts_d = torch.from_numpy(np.concatenate(pts, 0)).to(dev).contiguous()
    first_t = torch.tensor(first, dtype=torch.int32)
    dist = torch.empty((B, T, 778), device=dev)
    idx = torch.empty((B, T, 778), dtype=torch.int64, device=dev)

    def nn():
        _lib.check(L.tamf_h2o_dist(_lib.ptr(hv), _lib.ptr(traj), _lib.ptr(pts_d), _lib.C.c_void_p(first_t.data_ptr()), B, T,
                                   778, traj.shape[1], P, _lib.ptr(dist), _lib.ptr(idx), _lib.stream_ptr(dev)), "h2o")

    nn_ms, nn_min = timed(nn)
    from tamf_b200.chamfer import H2OIndex
    oix = H2OIndex(pts, dev)
    build_ms, _ = timed(lambda: H2OIndex(pts, dev))  # upload of the canonical clouds + index build (once per forward)

    def nn_q():
        _lib.check(L.tamf_h2o_dist_indexed(_lib.ptr(hv), _lib.ptr(traj), _lib.ptr(oix.index),
                                           _lib.C.c_void_p(first_t.data_ptr()), B, T, 778, traj.shape[1], P,
                                           _lib.ptr(dist), _lib.ptr(idx), _lib.stream_ptr(dev)), "h2o indexed")

    nnq_ms, nnq_min = timed(nn_q) if oix.index is not None else (nn_ms, nn_min)

    def nn_ex():
        _lib.check(L.tamf_h2o_dist_exhaustive(_lib.ptr(hv), _lib.ptr(traj), _lib.ptr(pts_d),
                                              _lib.C.c_void_p(first_t.data_ptr()), B, T, 778, traj.shape[1], P,
                                              _lib.ptr(dist), _lib.ptr(idx), _lib.stream_ptr(dev)), "h2o exhaustive")

    nnx_ms, nnx_min = timed(nn_ex, reps=max(1, a.reps // 3))
    side = torch.tensor(tamf_b200.InterationSegmentMDM.hand_side_ids(batch["hand_side"]), dtype=torch.int32, device=dev)
    out = torch.empty_like(x_in)
    m._ensure_bound(B, T, dev)

    def tr():
        _lib.check(L.tamf_refiner_forward(m._handle, _lib.ptr(x_in), _lib.ptr(dist), _lib.ptr(side), _lib.ptr(shape),
                                          _lib.ptr(traj), _lib.ptr(emb), traj.shape[1], _lib.ptr(out), _lib.stream_ptr(dev)),
                   "refiner")

    tr_ms, tr_min = timed(tr)

    peak, src = hbm_peak()
    fk_bytes = N * 10024.0
    nn_bytes = N * (778 * 12 + 778 * 12 + 36.0) + sum(first[-1:]) * P * 12.0
    nrm_bytes = N * (778 * 12 * 2.0)
    pairs = N * 778.0 * P * (np.mean([len(l) for l in batch["obj_list"]]))
    d, ff, Lr, S = cfg["latent_dim"], cfg["ff_size"], cfg["num_layers"], T + 3
    tr_flops = B * (Lr * (8 * S * d * d + 4 * S * S * d + 4 * S * d * ff) + 2 * T * 960 * d + 2 * T * d * d + 2 * T * d * 99)
    gbs = lambda by, ms: by / (ms * 1e-3) / 1e9
    line = {
        "metric": "refined motion sequences/sec, SegmentRefineModel forward (3 FK + 3 NN + transformer)",
        "value": B / (full_ms * 1e-3), "unit": "sequences/s", "n_gpus": 1, "ms_per_forward": full_ms,
        "ms_per_forward_min": full_min, "reps": a.reps, "warmup": a.warmup, "dtype": "f32 (FK, NN) / bf16 (transformer)",
        "data": "synthetic", "gpu_launches_per_forward": int(launches),
        "config": {"workload": f"MF-MDM R arch_refine, batch {B}, T={T}, {nobj} object(s) x {P} points, use_pc "
                               "(BASELINE.json configs[2])", "l2": "256 MB flush between timed repetitions"},
        "pieces": {
            "mano_fk": {"ms": fk_ms, "ms_min": fk_min, "frames": N, "alg_bytes": fk_bytes, "GBps": gbs(fk_bytes, fk_ms),
                        "hbm_frac": gbs(fk_bytes, fk_ms) / peak, "GFLOPs_fp32": N * 1.2e6 / (fk_ms * 1e-3) / 1e9},
            "vertex_normals": {"ms": nrm_ms, "ms_min": nrm_min, "alg_bytes": nrm_bytes, "GBps": gbs(nrm_bytes, nrm_ms),
                               "hbm_frac": gbs(nrm_bytes, nrm_ms) / peak},
            "h2o_nn": {"ms": nnq_ms, "ms_min": nnq_min, "ms_one_shot_with_index_build": nn_ms,
                       "ms_python_api": nn_api_ms, "ms_upload_and_index_build": build_ms, "alg_bytes": nn_bytes,
                       "GBps": gbs(nn_bytes, nnq_ms), "hbm_frac": gbs(nn_bytes, nnq_ms) / peak,
                       "pairs_per_s_equivalent": pairs / (nnq_ms * 1e-3),
                       "search": "exact, block-pruned (64-point blocks, object-frame boxes)" if P <= 8192 else "exhaustive"},
            "h2o_nn_exhaustive": {"ms": nnx_ms, "ms_min": nnx_min, "pairs_per_s": pairs / (nnx_ms * 1e-3),
                                  "TFLOPs_fp32": pairs * 8 / (nnx_ms * 1e-3) / 1e12},
            "transformer": {"ms": tr_ms, "ms_min": tr_min, "flops": tr_flops, "TFLOPs": tr_flops / (tr_ms * 1e-3) / 1e12},
        },
        "hbm_peak_GBps": peak, "peak_source": src,
    }
    print(json.dumps(line))


if __name__ == "__main__":
    main()
