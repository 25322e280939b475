"""TEST INFRASTRUCTURE ONLY -- loader for the *real* reference (read-only at /root/reference).

Imports the reference's own Python modules with the four offline shims SURVEY.md section 8c lists,
so that `oracle/make_golden.py` can generate golden vectors and `tests/test_oracle_pin.py` can pin
the restatement in `oracle/tamf_oracle.py` against the reference itself.  Nothing in the product
path, the `-m gpu` tests, `smoke()` or `bench.py` imports this file: `/root/reference` does not exist
on the GPU box.

Shims (each replaces something that needs the network, a licensed asset or an absent wheel):
  1. `ftfy`              -> stub module (`fix_text` = identity; only used by CLIP's text cleaning).
  2. `clip.load`         -> random-init ViT-B/32 `clip.model.CLIP(512,224,12,768,32,77,49408,512,8,12)`
                            (weights are a network download, thirdparty/CLIP/clip/clip.py:30-40,120).
  3. `pytorch3d`         -> stub package: `ops.knn.knn_points` (K=1 brute force with the arithmetic
                            oracle/tamf_oracle.py:nn_query defines) and `structures.Meshes`
                            (area-weighted vertex normals).  pytorch3d 0.7.2 is neither vendored nor
                            installed => NN / normals parity is "unpinned" (DESIGN.md).
  4. `manotorch.manolayer.ready_arguments` -> synthetic MANO-shaped assets (MANO .pkl is licensed and
                            chumpy does not import on py3.12).
"""
from __future__ import annotations

import os
import sys
import tempfile
import types
from collections import namedtuple

import numpy as np
import torch

REF_ROOT = os.environ.get("TAMF_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "src", "oakink2_tamf"))


_installed = False


def _install_paths():
    for p in (
        "src",
        "thirdparty/CLIP",
        "thirdparty/manotorch",
        "thirdparty/chamfer_distance",
        "thirdparty/config_reg/src",
    ):
        full = os.path.join(REF_ROOT, p)
        if full not in sys.path:
            sys.path.insert(0, full)


def _stub_ftfy():
    if "ftfy" not in sys.modules:
        m = types.ModuleType("ftfy")
        m.fix_text = lambda s: s
        sys.modules["ftfy"] = m


def _stub_pytorch3d():
    """knn_points(K=1) and Meshes.verts_normals_packed with the semantics DESIGN.md documents."""
    if "pytorch3d" in sys.modules and getattr(sys.modules["pytorch3d"], "_tamf_stub", False):
        return
    from . import tamf_oracle as orc

    p3d = types.ModuleType("pytorch3d")
    p3d._tamf_stub = True
    ops = types.ModuleType("pytorch3d.ops")
    knn = types.ModuleType("pytorch3d.ops.knn")
    structures = types.ModuleType("pytorch3d.structures")
    pointclouds = types.ModuleType("pytorch3d.structures.pointclouds")

    KNN = namedtuple("KNN", ["dists", "idx", "knn"])

    def knn_points(p1, p2, lengths1=None, lengths2=None, K=1, **kw):
        assert K == 1
        d2, idx = orc.nn_query(p1.detach().cpu().numpy(), p2.detach().cpu().numpy())
        return KNN(
            dists=torch.from_numpy(d2)[..., None].to(p1.device),
            idx=torch.from_numpy(idx)[..., None].to(p1.device),
            knn=None,
        )

    def knn_gather(x, idx, lengths=None):
        N, P1, K = idx.shape
        D = x.shape[2]
        return x[:, :, None].expand(-1, -1, K, -1).gather(1, idx[..., None].expand(-1, -1, -1, D))

    class Pointclouds:  # only used in isinstance checks
        pass

    class Meshes:
        def __init__(self, verts, faces):
            self._verts = verts  # [N,V,3]
            self._faces = faces  # [1 or N,F,3]

        def verts_normals_packed(self):
            v = self._verts.detach().cpu().numpy()
            f = self._faces.detach().cpu().numpy()
            n = orc.vertex_normals(v, f[0])
            return torch.from_numpy(n.reshape(-1, 3)).to(self._verts.device)

    knn.knn_points = knn_points
    knn.knn_gather = knn_gather
    ops.knn = knn
    ops.knn_points = knn_points
    ops.knn_gather = knn_gather
    pointclouds.Pointclouds = Pointclouds
    structures.Meshes = Meshes
    structures.pointclouds = pointclouds
    structures.Pointclouds = Pointclouds
    p3d.ops = ops
    p3d.structures = structures
    sys.modules["pytorch3d"] = p3d
    sys.modules["pytorch3d.ops"] = ops
    sys.modules["pytorch3d.ops.knn"] = knn
    sys.modules["pytorch3d.structures"] = structures
    sys.modules["pytorch3d.structures.pointclouds"] = pointclouds


def _patch_clip(seed: int = 1234):
    import clip  # vendored thirdparty/CLIP
    import clip.model as clip_model_mod

    def _load(name, device="cpu", jit=False, download_root=None):
        g = torch.random.get_rng_state()
        torch.manual_seed(seed)
        model = clip_model_mod.CLIP(512, 224, 12, 768, 32, 77, 49408, 512, 8, 12)
        torch.random.set_rng_state(g)
        clip_model_mod.convert_weights(model)  # build_model() does this (clip/model.py:435)
        model = model.eval().to(device)
        if str(device) == "cpu":
            model.float()  # clip/clip.py:138-142
        return model, None

    clip.load = _load


class _R:
    """Mimics a chumpy array: `.r` returns the ndarray (manolayer.py:73-79)."""

    def __init__(self, a):
        self.r = np.asarray(a)


def _patch_mano(asset_fn):
    import manotorch.manolayer as ml
    import scipy.sparse as sp

    def ready_arguments(path, *a, **k):
        side = "left" if "LEFT" in os.path.basename(path).upper() else "right"
        A = asset_fn(side)
        return {
            "betas": _R(np.zeros(10)),
            "shapedirs": _R(A["shapedirs"]),
            "posedirs": _R(A["posedirs"]),
            "v_template": _R(A["v_template"]),
            "weights": _R(A["weights"]),
            "J_regressor": sp.csc_matrix(A["J_regressor"]),
            "f": A["faces"],
            "kintree_table": np.array(
                [[4294967295, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14], list(range(16))], dtype=np.int64
            ),
            "hands_components": np.eye(45),
            "hands_mean": np.zeros(45),
        }

    ml.ready_arguments = ready_arguments


_mano_root = None


def mano_assets_root() -> str:
    """A temp dir holding empty models/MANO_{RIGHT,LEFT}.pkl to satisfy the isfile assert (manolayer.py:69-71)."""
    global _mano_root
    if _mano_root is None:
        _mano_root = tempfile.mkdtemp(prefix="tamf_mano_")
        os.makedirs(os.path.join(_mano_root, "models"), exist_ok=True)
        for s in ("RIGHT", "LEFT"):
            open(os.path.join(_mano_root, "models", f"MANO_{s}.pkl"), "wb").close()
    return _mano_root


def install(clip_seed: int = 1234):
    """Install all shims; idempotent.  Returns a namespace of reference modules."""
    global _installed
    if not reference_available():
        raise RuntimeError(f"reference not present at {REF_ROOT}")
    _pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oakink2-tamf_b200")
    if _pkg not in sys.path:
        sys.path.insert(0, _pkg)
    from tamf_b200 import synth

    _install_paths()
    _stub_ftfy()
    _stub_pytorch3d()
    if not _installed:
        _patch_clip(clip_seed)
        _patch_mano(synth.mano_assets)
        _installed = True

    import oakink2_tamf.model.interaction_segment_mdm as mdm
    import oakink2_tamf.model.segment_refine_model as refine
    import oakink2_tamf.model.diffusion_util as diffusion_util
    import oakink2_tamf.model.diffusion.gaussian_diffusion as gd
    import oakink2_tamf.model.diffusion.respace as respace
    import oakink2_tamf.model.loss.chamfer_distance as p2p
    import dev_fn.transform.rotation as rotation
    import dev_fn.transform.transform as transform
    import manotorch.manolayer as manolayer
    import chamfer_distance as chd

    return types.SimpleNamespace(
        mdm=mdm,
        refine=refine,
        diffusion_util=diffusion_util,
        gd=gd,
        respace=respace,
        p2p=p2p,
        rotation=rotation,
        transform=transform,
        manolayer=manolayer,
        chd=chd,
        mano_root=mano_assets_root(),
    )
