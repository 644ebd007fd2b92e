"""The documented integration really resolves the reference's own import statements to
this repository - checked in a clean interpreter per case, against the reference tree when it is present (build
container only; skipped on the GPU box)."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "unsupervised-pseuso-lidar_b200")
REF = os.environ.get("PL_REFERENCE_ROOT", "/root/reference")

needs_ref = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "dataloaders.py")), reason="reference tree absent")


def _run(code):
    env = dict(os.environ, PKG=PKG, REF=REF, PYTHONPATH="")
    p = subprocess.run([sys.executable, "-c", textwrap.dedent(code)], env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    return p.stdout


@needs_ref
def test_reference_dataloader_imports_through_the_dropin():
    """`dataloaders.py:15-17` star-imports geometry.pose_geometry and calls `mat2euler` per sample
    (`dataloaders.py:113`): the drop-in must export it, resolve `geometry.calibration` / `geometry.oxts_parser`
    to the reference (namespace merge), and must NOT load libplb200.so in a loader worker."""
    out = _run("""
        import os, sys, types
        PKG, REF = os.environ["PKG"], os.environ["REF"]
        for m in ("matplotlib", "matplotlib.pyplot"):
            sys.modules.setdefault(m, types.ModuleType(m))
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
        sys.path[:0] = [PKG, REF]
        import dataloaders
        import geometry.pose_geometry as pg, geometry.calibration as cal
        assert pg.__file__.startswith(PKG), pg.__file__
        assert cal.__file__.startswith(REF), cal.__file__
        assert dataloaders.mat2euler is pg.mat2euler
        for name in ("mat2euler", "isRotationMatrix", "invert_pose_np", "rot_from_axisangle", "get_translation_matrix",
                     "invert_pose", "transformation_from_parameters", "pose_vec2mat", "euler2mat", "disp_to_depth",
                     "inverse_warp", "Transform", "torch", "np", "math", "F"):
            assert hasattr(dataloaders, name), name
        assert not any(m.startswith("plb200") for m in sys.modules), "libplb200 loaded by a star-import"
        # host helpers against the reference's own, loaded by file path
        import importlib.util, numpy as np
        ref_geo = types.ModuleType("ref_geo"); ref_geo.__path__ = [os.path.join(REF, "geometry")]
        sys.modules["ref_geo"] = ref_geo
        spec = importlib.util.spec_from_file_location("ref_geo.pose_geometry", os.path.join(REF, "geometry", "pose_geometry.py"))
        rpg = importlib.util.module_from_spec(spec); sys.modules["ref_geo.pose_geometry"] = rpg
        spec.loader.exec_module(rpg)
        rng = np.random.default_rng(0)
        for k in range(20):
            a = rng.normal(size=3) * (0.3 if k else 0.0)
            cx, sx, cy, sy, cz, sz = np.cos(a[0]), np.sin(a[0]), np.cos(a[1]), np.sin(a[1]), np.cos(a[2]), np.sin(a[2])
            R = (np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]) @ np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
                 @ np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]]))
            if k == 5:
                R = np.array([[0.0, 0, 1], [0, 1, 0], [-1, 0, 0]])          # gimbal lock branch
            assert np.array_equal(pg.mat2euler(R), rpg.mat2euler(R)), (k, pg.mat2euler(R), rpg.mat2euler(R))
            assert pg.isRotationMatrix(R) == rpg.isRotationMatrix(R)
            T = np.eye(4); T[:3, :3] = R; T[:3, 3] = rng.normal(size=3)
            assert np.array_equal(pg.invert_pose_np(T), rpg.invert_pose_np(T))
        import torch
        t = torch.randn(3, 1, 3)
        assert torch.equal(pg.get_translation_matrix(t), rpg.get_translation_matrix(t))
        print("ok")
    """)
    assert "ok" in out


@needs_ref
def test_pseudo_lidar_imports_resolve_after_install():
    """`from utils.PseudoLiDAR import PseudoLiDAR` (PseudoLidarPipeline.py:14) inside the reference's `pseudo-lidar/`
    directory: path shadowing alone loses to the reference's regular `utils` package; after
    `plb200.dropin.install()` it resolves here, and `utils.model` still resolves to the reference."""
    out = _run("""
        import importlib.util, os, sys
        PKG, REF = os.environ["PKG"], os.environ["REF"]
        PLD = os.path.join(REF, "pseudo-lidar")
        sys.path[:0] = [PKG, PLD]
        assert importlib.util.find_spec("utils.PseudoLiDAR").origin.startswith(PLD)     # shadowing alone: reference wins
        import plb200.dropin
        names = plb200.dropin.install()
        assert "utils.PseudoLiDAR" in names and "Transform.Transform" in names
        from utils.PseudoLiDAR import PseudoLiDAR
        from Transform.Transform import Transform
        assert PseudoLiDAR.__module__ == "plb200.pseudolidar", PseudoLiDAR.__module__
        assert Transform.__module__ == "plb200.velodyne", Transform.__module__
        assert importlib.util.find_spec("utils.model").origin.startswith(PLD)
        print("ok")
    """)
    assert "ok" in out


def test_dropin_names_without_reference():
    """Without any reference tree on the path the drop-in directory alone serves every shadowed module."""
    out = _run("""
        import os, sys
        sys.path.insert(0, os.environ["PKG"])
        import plb200.dropin
        plb200.dropin.install()
        from utils.PseudoLiDAR import PseudoLiDAR
        from Transform.Transform import Transform
        import geometry.pose_geometry as pg
        import inspect
        sig = inspect.signature(pg.inverse_warp)
        assert list(sig.parameters)[:7] == ["img", "depth", "pose", "K", "pose_inv", "rotation_mode", "padding_mode"]
        assert sig.parameters["rotation_mode"].default == "euler" and sig.parameters["padding_mode"].default == "zeros"
        assert PseudoLiDAR.__module__ == "plb200.pseudolidar" and Transform.__module__ == "plb200.velodyne"
        print("ok")
    """)
    assert "ok" in out
