"""CPU: the C-ABI library loads and exports every symbol include/plb200.h declares;
argument validation returns error codes without touching a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    import sys
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.build()
    from plb200 import _lib
    return _lib


def test_every_declared_symbol_is_exported(L):
    header = open(os.path.join(ROOT, "include", "plb200.h")).read()
    declared = set(re.findall(r"\b(plb_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(L.SYMBOLS), declared ^ set(L.SYMBOLS)
    for name in declared:
        assert getattr(L.lib, name) is not None
    assert "sm_100a" in L.version()


def test_struct_sizes_match_header(L):
    # compile a tiny C program against the header and compare sizeof()
    import subprocess, tempfile
    src = r'''
    #include <stdio.h>
    #include "plb200.h"
    int main(void){ printf("%zu %zu %zu %zu %zu %zu\n", sizeof(plb_photo_job), sizeof(plb_photo_args),
        sizeof(plb_smooth_args), sizeof(plb_warp_args), sizeof(plb_cloud_args), sizeof(plb_velo_args)); return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    got = [ctypes.sizeof(t) for t in (L.PhotoJob, L.PhotoArgs, L.SmoothArgs, L.WarpArgs, L.CloudArgs, L.VeloArgs)]
    assert got == sizes, (got, sizes)


def test_argument_validation_without_gpu(L):
    a = L.PhotoArgs()
    assert L.lib.plb_photo_loss(None, None) == -2
    assert L.lib.plb_photo_loss(a, None) == -1          # B = 0
    a.B, a.H, a.W, a.n_jobs, a.n_pose = 1, 8, 8, 1, 1
    assert L.lib.plb_photo_loss(a, None) == -2          # poses NULL
    s = L.SmoothArgs()
    assert L.lib.plb_smooth_loss(s, None) == -1
    c = L.CloudArgs()
    assert L.lib.plb_cloud_project(c, None) == -1
    c.B, c.H, c.W = 1, 4, 4
    assert L.lib.plb_cloud_project(c, None) == -2
    v = L.VeloArgs()
    assert L.lib.plb_velo_project(v, None) == -1
    v.B, v.H, v.W, v.N, v.point_stride = 1, 4, 4, 0, 4
    assert L.lib.plb_velo_project(v, None) == -2         # no output
    w = L.WarpArgs()
    assert L.lib.plb_warp_forward(w, None) == -1
    pr = L.PrepArgs()
    pr.B, pr.in_h, pr.in_w, pr.H, pr.W, pr.n_K = 2, 8, 8, 4, 4, 3      # more intrinsics matrices than frames
    pr.frames = pr.out_planar = pr.K_in = pr.K_out = pr.workspace = 1   # non-NULL: validation only, nothing is launched
    assert L.lib.plb_prep_frames(pr, None) == -1
    with pytest.raises(L.PlbError):
        L.check(-3, "x")


def test_ops_refuse_cpu_tensors(L):
    import torch
    from plb200 import ops, synth
    inp = synth.make_photo_inputs(1, 8, 16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.fused_losses(inp["tgt"], inp["ref_imgs"], inp["disparity"], inp["poses"], inp["intrinsics"])
    from geometry.pose_geometry import inverse_warp
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        inverse_warp(inp["tgt"], inp["disparity"][0][0], inp["poses"][:, 0], inp["intrinsics"], False)


def test_integration_stub_matches_the_abi(L):
    """The ctypes stub printed in INTEGRATION.md (what a reference maintainer would paste) mirrors the real structs."""
    import ctypes as C
    src = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    code = src[src.index("class Job(C.Structure):"):src.index("lib.plb_photo_workspace_bytes.restype")]
    ns = {"C": C}
    exec(code, ns)
    assert C.sizeof(ns["Job"]) == C.sizeof(L.PhotoJob)
    assert C.sizeof(ns["Args"]) == C.sizeof(L.PhotoArgs)
    assert [f[0] for f in ns["Args"]._fields_] == [f[0] for f in L.PhotoArgs._fields_]


def test_torch_binding_loads_and_refuses_cpu_tensors(L):
    """The thin torch C++ binding (csrc/torch_binding.cpp) is built in-tree, links the same libplb200.so, and has no
    CPU path either (both bindings are checked: the default one and the ctypes one)."""
    from plb200 import ops, synth, _tb
    assert _tb.mod is not None and os.path.isfile(_tb.PATH)
    assert _tb.mod.version() == L.version()
    inp = synth.make_photo_inputs(1, 8, 16)
    for binding in ("torch", "ctypes"):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            ops.fused_losses(inp["tgt"], inp["ref_imgs"], inp["disparity"], inp["poses"], inp["intrinsics"], binding=binding)


def test_ops_reject_mismatched_shapes(L):
    """The kernels see pointers, not extents: shapes the reference's torch ops would reject are rejected on the host
    (checked before the device check, so this runs without a GPU)."""
    import torch
    from plb200 import ops, synth
    inp = synth.make_photo_inputs(2, 8, 16, n_src=2, n_scales=2)
    t, r, d, p, K = inp["tgt"], inp["ref_imgs"], inp["disparity"], inp["poses"], inp["intrinsics"]
    bad = [
        (t, r, d, p, K[:1]),                                   # one intrinsics matrix for two samples
        (t, r, d, p[:, :1], K),                                # fewer poses than sources
        (t, [r[0], r[1][:1]], d, p, K),                        # a reference image of another batch
        (t, r, [[x[:1] for x in d[0]]] + list(d[1:]), p, K),   # a disparity pyramid of another batch
        (t[:, :1], r, d, p, K),                                # not an RGB image
    ]
    for args in bad:
        with pytest.raises(ValueError):
            ops.fused_losses(*args)
