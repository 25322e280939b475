"""The C-ABI library loads and exports every symbol include/tamf_b200.h declares (no compute calls: no GPU here)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def header_symbols():
    src = open(os.path.join(ROOT, "include", "tamf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tamf_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    from tamf_b200 import _lib
    assert sorted(_lib.SYMBOLS) == header_symbols()


def test_library_exports_every_declared_symbol():
    from tamf_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    for s in header_symbols():
        assert hasattr(L, s), f"{s} declared in include/tamf_b200.h but not exported"
    L.tamf_version.restype = ctypes.c_int
    assert L.tamf_version() >= 100
    L.tamf_kernel_launch_count.restype = ctypes.c_uint64
    assert L.tamf_kernel_launch_count() == 0


def test_struct_layouts_match_header():
    """ctypes mirrors of tamf_cfg / tamf_layer_weights / tamf_g_weights: field counts and sizes (LP64)."""
    from tamf_b200 import _lib
    assert ctypes.sizeof(_lib.TamfCfg) == 10 * 4
    assert ctypes.sizeof(_lib.TamfLayerWeights) == 12 * 8
    assert ctypes.sizeof(_lib.TamfGWeights) == 21 * 8 + 8 + 8 + 3 * 8  # pe_rows padded to 8
    assert ctypes.sizeof(_lib.TamfRWeights) == 17 * 8 + 8 + 8


def test_no_gpu_fails_loudly():
    """No CPU fallback: without a CUDA device the product entry points raise instead of computing elsewhere."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import tamf_b200
    from tamf_b200 import synth
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tamf_b200.nn_query(torch.zeros(1, 4, 3), torch.zeros(1, 4, 3))
    m = tamf_b200.InterationSegmentMDM(**synth.ARCH["arch_mdm"], text_encoder=synth.text_features)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 99, 1, 8), torch.zeros(1, dtype=torch.long), synth.make_batch(1, 8, nobj=1))
    layer = tamf_b200.ManoLayer(side="right", assets=synth.mano_assets("right"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        layer(pose_coeffs=torch.zeros(2, 16, 4), betas=torch.zeros(2, 10))


def test_product_path_never_imports_oracle():
    pkg = os.path.join(ROOT, "oakink2-tamf_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                txt = open(os.path.join(dp, fn)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f"{fn} references oracle/"


def test_library_carries_tcgen05_and_tma_sass():
    """The shipped kernels are sm_100a code that issues tensor-core MMAs from shared-memory descriptors (UTCHMMA), reads
    accumulators out of TMEM (LDTM) and moves tiles with TMA loads / stores (UTMALDG / UTMASTG) -- checked on the SASS
    of the in-tree library (mnemonics as listed in the B200 profiling recipe)."""
    import shutil
    import subprocess
    from tamf_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    assert "arch = sm_100a" in sass
    for mnemonic, least in (("UTCHMMA", 50), ("LDTM", 20), ("UTMALDG", 20), ("UTMASTG", 10)):
        assert sass.count(mnemonic) >= least, f"{mnemonic}: {sass.count(mnemonic)} occurrences"
