"""The opt-in kernel variants (environment switches read once per process) against the same parity tests as the defaults:
TAMF_LN_RQ (rows per TMEM lane quarter of the LayerNorm tiles, csrc/gemm.cuh gemm_ln_rq) and TAMF_ATTN_DBG=2 (3 in 8
softmax exponentials as an FMA-pipe polynomial, csrc/attn_tc.cuh).  Each case re-runs the golden-vector and edge-shape
forward tests of tests/test_denoiser_gpu.py in a child process with the switch set."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("env", [{"TAMF_LN_RQ": "0", "TAMF_ATTN_DBG": "2"}, {"TAMF_LN_RQ": "24"}],
                         ids=["ln_all_sms+poly_exp2", "ln_96_rows"])
def test_forward_parity_with_variant(env):
    child_env = dict(os.environ, **env)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_denoiser_gpu.py"), "-q", "-x",
                        "-m", "gpu", "-k", "golden or edge_shapes", "-p", "no:cacheprovider"],
                       cwd=ROOT, env=child_env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
