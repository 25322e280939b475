"""TEST / BASELINE INFRASTRUCTURE ONLY -- recipe for `oracle/_ref/`: a private copy of the reference's own Python modules
for the sampling path, so that `bench.py --impl reference` and `cpu_baseline` can execute the UNMODIFIED reference
(`InterationSegmentMDM.forward` + `GaussianDiffusion.p_sample`) on the GPU box's host cores, where /root/reference does
not exist.

Nothing is copied into the repository's history: `oracle/_ref/` is git-ignored (it is NOT gpurun-ignored, so it travels
to the GPU box like the built .so).  The file set is not listed by hand: the reference is imported here through
`oracle/ref_shims.py` (the four offline shims of SURVEY.md 8c) in a subprocess, and every module file that was loaded
from under the reference root is copied with its relative path, plus the package `__init__.py` files on the way and
CLIP's tokenizer vocabulary.  `python oracle/make_ref.py` (also run by `__graft_entry__.build()` when /root/reference is
present)."""
from __future__ import annotations

import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
DEST = os.path.join(HERE, "_ref")
REF = os.environ.get("TAMF_REFERENCE_SOURCE", "/root/reference")

_PROBE = r"""
import json, os, sys
sys.path.insert(0, {root!r})
os.environ["TAMF_REFERENCE_ROOT"] = {ref!r}
from oracle import ref_shims
ref_shims.install()
ref = os.path.realpath({ref!r}) + os.sep
files = sorted({{os.path.realpath(m.__file__) for m in list(sys.modules.values())
                if getattr(m, "__file__", None) and os.path.realpath(m.__file__).startswith(ref)}})
print("@@" + json.dumps(files))
"""


def build_ref(verbose: bool = True) -> int:
    """Returns the number of files in oracle/_ref (0: reference absent and nothing was built before)."""
    if not os.path.isdir(os.path.join(REF, "src", "oakink2_tamf")):
        n = sum(len(f) for _, _, f in os.walk(DEST)) if os.path.isdir(DEST) else 0
        if verbose:
            print(f"oracle/_ref: reference not present at {REF}; keeping the {n} files already there")
        return n
    out = subprocess.run([sys.executable, "-c", _PROBE.format(root=ROOT, ref=REF)], capture_output=True, text=True,
                         timeout=600)
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("@@")]
    if out.returncode != 0 or not line:
        raise RuntimeError("oracle/_ref: importing the reference failed:\n" + out.stderr[-2000:])
    files = json.loads(line[0][2:])
    ref = os.path.realpath(REF) + os.sep
    extra = []
    for f in files:  # package __init__.py files between the root and every module
        d = os.path.dirname(f)
        while d.startswith(ref):
            init = os.path.join(d, "__init__.py")
            if os.path.isfile(init):
                extra.append(init)
            d = os.path.dirname(d)
    vocab = os.path.join(ref, "thirdparty", "CLIP", "clip", "bpe_simple_vocab_16e6.txt.gz")
    if os.path.isfile(vocab):
        extra.append(vocab)
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    n = 0
    for f in sorted(set(files + extra)):
        dst = os.path.join(DEST, os.path.relpath(f, ref))
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(f, dst)
        n += 1
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": REF, "files": [os.path.relpath(f, ref) for f in sorted(set(files + extra))]}, fh, indent=1)
    if verbose:
        print(f"oracle/_ref: {n} reference files copied from {REF}")
    return n


if __name__ == "__main__":
    build_ref()
