"""ORACLE - TEST INFRASTRUCTURE ONLY.  Imports the UNMODIFIED reference from
`/root/reference` (build container only; the GPU box has no such directory).

Used by `tests/golden/make_golden.py` to generate the committed golden vectors
and by `tests/test_oracle_golden.py::test_live_reference_*` (skipped when the
reference tree is absent).  No reference source is copied; three shims make it
importable on CPU:

  1. matplotlib is imported at module top (`losses.py:5`,
     `geometry/transform.py:8`) but is not installed -> stub modules;
  2. `geometry/transform.py:134` calls `.cuda()` -> identity on a CPU-only host;
  3. `geometry/transform.py:110` hard-codes batch 4 -> `k_hom` is replaced by a
     batch-agnostic twin ONLY when asked (`patch_batch=True`); at B=4 the
     reference runs untouched.
"""
import contextlib
import importlib.util
import io
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PL_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "losses.py"))


def load(patch_batch=False):
    """Returns a namespace with the reference's Losses, SSIM, inverse_warp,
    disp_to_depth, Transform, pose functions and PseudoLiDAR class."""
    import torch
    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    for m in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    # The product's drop-in modules use the same top-level names (`losses`,
    # `geometry`, `utils`): make sure the reference's win inside this process.
    for name in list(sys.modules):
        if name in ("losses", "geometry", "utils") or name.startswith(("geometry.", "utils.")):
            mod = sys.modules[name]
            if REFERENCE_ROOT not in (getattr(mod, "__file__", "") or ""):
                del sys.modules[name]
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import geometry.pose_geometry as pg
    import geometry.transform as tr
    import losses as ref_losses
    if patch_batch:
        def _k_hom(self, K):
            Kh = torch.eye(4).reshape(1, 4, 4).repeat(K.shape[0], 1, 1).to(device=K.device)
            Kh[:, :3, :3] = K.clone()
            return Kh
        tr.Transform.k_hom = _k_hom
    spec = importlib.util.spec_from_file_location(
        "ref_pseudolidar", os.path.join(REFERENCE_ROOT, "pseudo-lidar", "utils", "PseudoLiDAR.py"))
    pl = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(pl)
    ns = types.SimpleNamespace(
        Losses=ref_losses.Losses, SSIM=ref_losses.SSIM, inverse_warp=pg.inverse_warp,
        disp_to_depth=pg.disp_to_depth, pose_vec2mat=pg.pose_vec2mat, euler2mat=pg.euler2mat,
        invert_pose=pg.invert_pose, transformation_from_parameters=pg.transformation_from_parameters,
        Transform=tr.Transform, PseudoLiDAR=pl.PseudoLiDAR)
    return ns


@contextlib.contextmanager
def quiet():
    """The reference prints inside the loss (`losses.py:191`)."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def load_velo_transform():
    """The reference's `pseudo-lidar/Transform/Transform.py::Transform` (Velodyne -> image depth map),
    imported unmodified by file path (the directory name has a hyphen)."""
    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    for m in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    spec = importlib.util.spec_from_file_location(
        "ref_velo_transform", os.path.join(REFERENCE_ROOT, "pseudo-lidar", "Transform", "Transform.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.Transform
