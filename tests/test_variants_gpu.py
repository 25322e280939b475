"""The opt-in forms of the encoder (environment switches read once per process) against the same parity tests as the
default (the per-layer form of csrc/layer_chain.cuh, list order):
  TAMF_CHAIN=0        the five-kernel layer of round 1, with its own switches TAMF_LN_RQ (rows per TMEM lane quarter of the
                      LayerNorm tiles, csrc/gemm.cuh gemm_ln_rq) and TAMF_ATTN_DBG=2 (3 in 8 softmax exponentials as an
                      FMA-pipe polynomial, csrc/attn_tc.cuh)
  TAMF_CHAIN=2        the stack form: one persistent layer kernel for all layers + one persistent attention kernel
  TAMF_CHAIN_OOO=8    run-time unit selection in the layer kernel (also on top of the stack form)
  TAMF_FINE=0         grid-wide waits between the attention and layer kernels instead of per-row-tile / per-sequence ones
Each case re-runs the golden-vector, edge-shape and replay-determinism tests of tests/test_denoiser_gpu.py in a child
process with the switches set."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

VARIANTS = {
    "round1_form_ln_all_sms+poly_exp2": {"TAMF_CHAIN": "0", "TAMF_LN_RQ": "0", "TAMF_ATTN_DBG": "2"},
    "round1_form_ln_96_rows": {"TAMF_CHAIN": "0", "TAMF_LN_RQ": "24"},
    "stack_form": {"TAMF_CHAIN": "2"},
    "stack_form+unit_selection": {"TAMF_CHAIN": "2", "TAMF_STACK_ATT": "24", "TAMF_CHAIN_OOO": "8"},
    "unit_selection": {"TAMF_CHAIN_OOO": "8"},
    "grid_wide_waits": {"TAMF_FINE": "0"},
}


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(VARIANTS))
def test_forward_parity_with_variant(name):
    child_env = dict(os.environ, **VARIANTS[name])
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_denoiser_gpu.py"), "-q", "-x",
                        "-m", "gpu", "-k", "golden or edge_shapes or replays", "-p", "no:cacheprovider"],
                       cwd=ROOT, env=child_env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
