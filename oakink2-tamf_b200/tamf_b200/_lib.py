"""ctypes binding of libtamf_b200.so (include/tamf_b200.h).  There is no CPU fallback: every product entry point
goes through this library and raises if it is missing or the device is not a B200."""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libtamf_b200.so")

TAMF_OK, TAMF_E_BADARG, TAMF_E_CUDA, TAMF_E_ARCH, TAMF_E_ALIGN, TAMF_E_STATE = 0, -1, -2, -3, -4, -5
POSE_QUAT, POSE_REPR = 0, 1

# every symbol include/tamf_b200.h declares (tests/test_abi.py checks the .so exports them all)
SYMBOLS = [
    "tamf_version", "tamf_last_error", "tamf_nn_query", "tamf_h2o_dist", "tamf_h2o_dist_exhaustive", "tamf_h2o_index_bytes",
    "tamf_h2o_index_build", "tamf_h2o_dist_indexed", "tamf_mano_create", "tamf_mano_destroy",
    "tamf_mano_fk", "tamf_mano_fk_full", "tamf_denoiser_create", "tamf_denoiser_destroy", "tamf_denoiser_workspace_bytes",
    "tamf_denoiser_bind", "tamf_denoiser_set_cond", "tamf_denoiser_forward", "tamf_p_sample_step",
    "tamf_p_sample_chain", "tamf_p_sample_loop_host", "tamf_denoiser_profile_step", "tamf_denoiser_profile_graph", "tamf_kernel_launch_count", "tamf_philox_normal",
    "tamf_gemm_selftest", "tamf_refiner_create", "tamf_refiner_destroy", "tamf_refiner_workspace_bytes",
    "tamf_refiner_bind", "tamf_refiner_forward", "tamf_mano_fk_select", "tamf_vertex_normals", "tamf_gemm_trace",
    "tamf_attn_selftest", "tamf_attn_trace", "tamf_denoiser_set_sampler", "tamf_denoiser_sampler_steps",
    "tamf_layer_aux_bytes", "tamf_layer_run", "tamf_debug_chain_trace", "tamf_layer_schedule", "tamf_stack_schedule",
]


class TamfCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "input_dim", "obj_input_dim", "hand_shape_dim", "obj_embed_dim", "latent_dim", "ff_size", "num_layers",
        "num_heads", "clip_dim", "num_steps")]


_FP = C.c_void_p


class TamfLayerWeights(C.Structure):
    _fields_ = [(n, _FP) for n in (
        "in_proj_w", "in_proj_b", "out_proj_w", "out_proj_b", "lin1_w", "lin1_b", "lin2_w", "lin2_b", "norm1_w",
        "norm1_b", "norm2_w", "norm2_b")]


class TamfGWeights(C.Structure):
    _fields_ = [(n, _FP) for n in (
        "shape_w", "shape_b", "objemb_w", "objemb_b", "pose_w", "pose_b", "objtraj_w", "objtraj_b", "merge0_w",
        "merge0_b", "merge2_w", "merge2_b", "time0_w", "time0_b", "time2_w", "time2_b", "text_w", "text_b", "final_w",
        "final_b", "pe")] + [("pe_rows", C.c_int32), ("layers", C.POINTER(TamfLayerWeights)),
                             ("posterior_mean_coef1", _FP), ("posterior_mean_coef2", _FP),
                             ("posterior_log_variance_clipped", _FP)]


class TamfRWeights(C.Structure):
    _fields_ = [(n, _FP) for n in (
        "shape_w", "shape_b", "objemb_w", "objemb_b", "pose_w", "pose_b", "objtraj_w", "objtraj_b", "dist_w", "dist_b",
        "merge0_w", "merge0_b", "merge2_w", "merge2_b", "final_w", "final_b", "pe")] + [
        ("pe_rows", C.c_int32), ("layers", C.POINTER(TamfLayerWeights))]


_lib = None


def lib() -> C.CDLL:
    """Load the CUDA library; fails loudly (no fallback) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(oakink2-tamf_b200/build.sh).  tamf_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, u64, sz = C.c_void_p, C.c_int, C.c_uint64, C.c_size_t
    L.tamf_version.restype = C.c_int
    L.tamf_last_error.restype = C.c_char_p
    L.tamf_kernel_launch_count.restype = C.c_uint64
    L.tamf_nn_query.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp]
    L.tamf_h2o_dist.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp]
    L.tamf_h2o_dist_exhaustive.argtypes = L.tamf_h2o_dist.argtypes
    L.tamf_h2o_dist_indexed.argtypes = L.tamf_h2o_dist.argtypes
    L.tamf_h2o_index_bytes.argtypes = [i32, i32]
    L.tamf_h2o_index_bytes.restype = sz
    L.tamf_h2o_index_build.argtypes = [vp, i32, i32, vp, sz, vp]
    L.tamf_mano_create.argtypes = [vp, vp, vp, vp, vp, i32, C.POINTER(vp)]
    L.tamf_mano_destroy.argtypes = [vp]
    L.tamf_mano_fk.argtypes = [vp, i32, vp, vp, i32, vp, vp, vp]
    L.tamf_mano_fk_full.argtypes = [vp, i32, vp, vp, i32, vp, vp, vp, vp, vp]
    L.tamf_denoiser_create.argtypes = [C.POINTER(TamfCfg), C.POINTER(TamfGWeights), C.POINTER(vp)]
    L.tamf_denoiser_destroy.argtypes = [vp]
    L.tamf_denoiser_workspace_bytes.argtypes = [vp, i32, i32]
    L.tamf_denoiser_workspace_bytes.restype = sz
    L.tamf_denoiser_bind.argtypes = [vp, i32, i32, vp, sz]
    L.tamf_denoiser_set_cond.argtypes = [vp, vp, vp, vp, vp, vp, i32, vp]
    L.tamf_denoiser_forward.argtypes = [vp, vp, vp, vp, vp]
    L.tamf_p_sample_step.argtypes = [vp, vp, i32, vp, u64, vp, vp]
    L.tamf_p_sample_chain.argtypes = [vp, vp, i32, i32, u64, vp]
    L.tamf_p_sample_loop_host.argtypes = [vp, vp, vp, vp, vp, vp, i32, vp, u64, vp, vp]
    L.tamf_denoiser_profile_step.argtypes = [vp, vp, i32, u64, vp, i32, vp, vp]
    L.tamf_denoiser_profile_graph.argtypes = [vp, vp, i32, i32, u64, vp, vp, vp, i32, vp, vp, vp]
    L.tamf_philox_normal.argtypes = [vp, sz, u64, C.c_uint32, vp]
    L.tamf_gemm_selftest.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]
    L.tamf_refiner_create.argtypes = [C.POINTER(TamfCfg), C.POINTER(TamfRWeights), C.POINTER(vp)]
    L.tamf_refiner_destroy.argtypes = [vp]
    L.tamf_refiner_workspace_bytes.argtypes = [vp, i32, i32]
    L.tamf_refiner_workspace_bytes.restype = sz
    L.tamf_refiner_bind.argtypes = [vp, i32, i32, vp, sz]
    L.tamf_refiner_forward.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, vp, vp]
    L.tamf_mano_fk_select.argtypes = [vp, i32, vp, vp, vp, i32, vp, vp, vp]
    L.tamf_vertex_normals.argtypes = [vp, vp, i32, i32, i32, vp, vp]
    L.tamf_gemm_trace.argtypes = [i32, vp, vp, vp, vp, vp, i32, i32, i32, vp, vp]
    L.tamf_attn_selftest.argtypes = [vp, vp, i32, i32, i32, i32, vp]
    L.tamf_denoiser_set_sampler.argtypes = [vp, i32, vp, vp, vp, vp]
    L.tamf_denoiser_sampler_steps.argtypes = [vp]
    L.tamf_attn_trace.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp]
    L.tamf_layer_aux_bytes.argtypes = [i32, i32, i32]
    L.tamf_layer_aux_bytes.restype = sz
    L.tamf_layer_run.argtypes = [vp] * 12 + [i32, i32, i32, i32, vp, sz, vp, vp]
    L.tamf_debug_chain_trace.argtypes = [vp, i32]
    L.tamf_layer_schedule.argtypes = [i32, i32, i32, i32, i32, vp, vp, i32, vp]
    L.tamf_stack_schedule.argtypes = [i32, i32, i32, i32, i32, i32, i32, i32, C.c_double, vp, vp, i32, vp]
    _lib = L
    return L


def check(rc: int, what: str = ""):
    """Map the C error convention onto the reference's Python exceptions (ValueError for bad input)."""
    if rc == TAMF_OK:
        return
    msg = lib().tamf_last_error().decode("utf-8", "replace")
    if rc in (TAMF_E_BADARG, TAMF_E_ALIGN):
        raise ValueError(f"{what}: {msg}")
    raise RuntimeError(f"{what}: {msg} (code {rc})")


def ptr(t: torch.Tensor | None):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def dev_f32(t: torch.Tensor, device) -> torch.Tensor:
    return t.to(device=device, dtype=torch.float32).contiguous()
