"""CPU: pins oracle/restated.py against the golden vectors produced by the
unmodified reference (tests/golden/make_golden.py)."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import restated as O
from helpers import load_golden, golden_inputs, rel_err, max_rel_err

LIVE = ["live_b4_s1_32x48", "live_b4_s4_32x64", "live_b4_s1_init_24x40",
        "live_b2_s2_32x48_patched", "live_b3_s1_24x40_patched"]


@pytest.mark.parametrize("name", LIVE)
def test_live_loss_and_grads(name):
    g = load_golden(name)
    tgt, refs, disparity, poses, K = golden_inputs(g)
    disparity = [[t.requires_grad_(True) for t in fr] for fr in disparity]
    poses.requires_grad_(True)
    tgt.requires_grad_(True)
    refs = [r.requires_grad_(True) for r in refs]
    loss = O.losses_forward(tgt, refs, disparity, poses, K)
    sum(loss).backward()
    assert abs(float(loss[0]) - float(g["loss_mam"])) <= 1e-6 * abs(float(g["loss_mam"]))
    assert abs(float(loss[1]) - float(g["loss_smooth"])) <= 1e-6 * abs(float(g["loss_smooth"]))
    assert rel_err(poses.grad, g["g_poses"]) < 1e-5
    assert rel_err(tgt.grad, g["g_tgt"]) < 1e-5
    for i, r in enumerate(refs):
        assert rel_err(r.grad, g["g_ref%d" % i]) < 1e-5
    for f, fr in enumerate(disparity):
        for s, t in enumerate(fr):
            assert rel_err(t.grad, g["g_disp_f%d_s%d" % (f, s)]) < 1e-5


def test_warp_and_pose_functions():
    g = load_golden("warp_b4_24x40")
    K = torch.from_numpy(g["K"])
    cot = torch.from_numpy(g["cot"])
    for inv, tag in ((False, "fwd"), (True, "inv")):
        img = torch.from_numpy(g["img"]).requires_grad_(True)
        depth = torch.from_numpy(g["depth"]).requires_grad_(True)
        pose = torch.from_numpy(g["pose"]).requires_grad_(True)
        proj = O.inverse_warp(img, depth, pose, K, inv)
        (proj * cot).sum().backward()
        assert max_rel_err(proj, g["proj_" + tag]) < 1e-6
        assert rel_err(img.grad, g["g_img_" + tag]) < 1e-6
        assert rel_err(depth.grad, g["g_depth_" + tag]) < 1e-5
        assert rel_err(pose.grad, g["g_pose_" + tag]) < 1e-5
    pose = torch.from_numpy(g["pose"])
    M = O.pose_matrix(pose, False)
    assert torch.allclose(M, torch.from_numpy(g["M_axisangle"]), atol=1e-7)
    assert torch.allclose(O.invert_pose(M), torch.from_numpy(g["M_inverted"]), atol=1e-7)
    assert torch.allclose(O.pose_vec2mat(pose, "euler"), torch.from_numpy(g["M_euler"]), atol=1e-7)
    Xc = O.reconstruct(torch.from_numpy(g["depth"])[:, 0], K)
    assert max_rel_err(Xc, g["Xc"]) < 1e-6
    assert max_rel_err(O.project(Xc, K, M), g["grid"]) < 1e-6


def test_dormant_ssim_photometric_min_automask():
    g = load_golden("dormant_b4_s2_32x48")
    tgt, refs, disparity, poses, K = golden_inputs(g)
    assert max_rel_err(O.ssim_standard_loss(refs[0], tgt), g["ssim_ref0_tgt"]) < 1e-6
    assert max_rel_err(O.compute_photometric_loss(refs[0], tgt), g["photo_clip_ref0_tgt"]) < 1e-6
    assert max_rel_err(O.compute_photometric_loss(refs[0], tgt, no_ssim=True),
                       g["photo_clip_nossim_ref0_tgt"]) < 1e-6
    for clip, ctag in ((0.5, "clip"), (None, "noclip")):
        for automask, atag in ((True, "auto"), (False, "noauto")):
            tag = "%s_%s" % (ctag, atag)
            disp = [t.clone().requires_grad_(True) for t in disparity[0]]
            p = poses.clone().requires_grad_(True)
            srcs = [r.clone().requires_grad_(True) for r in refs]
            depths = O.disp_to_depth([disp])[0]
            loss = O.min_reprojection_loss(tgt, srcs, depths, p, K, automask=automask, clip_loss=clip)
            loss.backward()
            assert abs(float(loss) - float(g["loss_" + tag])) <= 2e-6 * abs(float(g["loss_" + tag])), tag
            assert rel_err(p.grad, g["g_poses_" + tag]) < 1e-4, tag
            for s, t in enumerate(disp):
                assert rel_err(t.grad, g["g_disp_s%d_%s" % (s, tag)]) < 1e-4, tag
            for i, r in enumerate(srcs):
                assert rel_err(r.grad, g["g_ref%d_%s" % (i, tag)]) < 1e-4, tag


def test_cloud_small_and_full():
    g = load_golden("cloud_kitti")
    T, P = g["T"], g["P"]
    for sp in (0, 3):
        cloud = O.project_PL(g["small_depth"], T, P, sparsity=sp)
        ref = g["small_cloud_sp%d" % sp]
        assert cloud.dtype == np.float64 and cloud.shape == ref.shape
        assert np.array_equal(cloud, ref)          # bit-exact, fp64
        assert np.all(cloud[:, 3] == 0.0)
    from plb200 import synth
    full = synth.make_depth_images(1, 375, 1242, seed=int(g["full_seed"]))[0].numpy()
    cloud = O.project_PL(full, T, P)
    assert cloud.shape[0] == int(g["full_count"])
    digest = np.frombuffer(hashlib.sha256(np.ascontiguousarray(cloud).tobytes()).digest(), dtype=np.uint8)
    assert np.array_equal(digest, g["full_sha256"])
    assert np.array_equal(cloud[::997], g["full_sample"])
    assert O.project_PL(full, T, P, sparsity=10).shape[0] == int(g["full_count_sp10"])


def test_velo_projection_bit_exact():
    """Velodyne -> image (Transform.py:69-104): the vectorised oracle equals the unmodified reference's
    per-point loop bit for bit, collisions (later point wins) included."""
    from plb200 import synth
    g = load_golden("velo_kitti")
    depth, winner = O.project_velo_to_img(g["small_points"], g["T"], g["small_P"], 124, 37)
    assert depth.dtype == np.float64 and depth.shape == (37, 124)
    assert np.array_equal(depth, g["small_depth"])
    kept = O.project_velo_to_img(g["small_points"], g["T"], g["small_P"], 124, 37)[1]
    assert (kept >= 0).sum() == (g["small_depth"] != 0).sum()
    full = synth.make_velodyne_cloud(int(g["full_n"]), seed=int(g["full_seed"]))
    depth, _ = O.project_velo_to_img(full, g["T"], g["P"], 1242, 375)
    digest = np.frombuffer(hashlib.sha256(np.ascontiguousarray(depth).tobytes()).digest(), dtype=np.uint8)
    assert np.array_equal(digest, g["full_sha256"])
    nz = np.flatnonzero(depth)
    assert np.array_equal(nz, g["full_nz_index"]) and np.array_equal(depth.reshape(-1)[nz], g["full_nz_value"])


def test_disp_head_golden():
    """`alpha * sigmoid(x) + beta` against the disparities the unmodified DispNetS returned for the captured
    pre-activations (models/depth/disp_net.py:121-139)."""
    g = load_golden("head_dispnet")
    for k in ("1", "3"):
        out = O.disp_head(torch.from_numpy(g["x" + k]))
        assert torch.equal(out, torch.from_numpy(g["disp" + k]))


def test_live_reference_if_present():
    """When the reference tree is mounted (build container), run it live against
    the oracle at B=4 once more - guards against a stale fixture."""
    from oracle import reference_shim
    if not reference_shim.available():
        pytest.skip("reference tree not mounted")
    import subprocess, sys, os
    code = (
        "import sys; sys.path[:0]=[%r,%r]\n"
        "import torch\n"
        "from oracle import reference_shim, restated as O\n"
        "from plb200 import synth\n"
        "ref = reference_shim.load()\n"
        "inp = synth.make_photo_inputs(4, 16, 24, n_scales=2, seed=5)\n"
        "with reference_shim.quiet():\n"
        "    a = ref.Losses().forward(inp['tgt'], inp['ref_imgs'], inp['disparity'], inp['poses'], inp['intrinsics'], None)\n"
        "b = O.losses_forward(inp['tgt'], inp['ref_imgs'], inp['disparity'], inp['poses'], inp['intrinsics'])\n"
        "assert abs(float(a[0])-float(b[0])) < 1e-6*abs(float(a[0])), (a, b)\n"
        "assert abs(float(a[1])-float(b[1])) < 1e-6*abs(float(a[1])), (a, b)\n"
        "print('OK')\n") % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                              os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                           "unsupervised-pseuso-lidar_b200"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "OK" in out.stdout, out.stderr[-2000:]


def test_loader_chain_oracle_matches_reference_bit_for_bit():
    """`oracle.load_img_chain` / `scale_intrinsics` against the unmodified `KittiDataset.load_img` + trainer transform
    list (tests/golden/prep_frames.npz): down- and up-sampling, identity size, odd sizes.  Bit-exact."""
    g = load_golden("prep_frames")
    for tag in ("down", "same", "up", "odd"):
        H, W = (int(v) for v in g[tag + "_size"])
        frame = g[tag + "_frame"]
        assert np.array_equal(O.load_img_chain(frame, H, W), g[tag + "_out"]), tag
        assert np.array_equal(O.scale_intrinsics(g[tag + "_K_in"], frame.shape[0], frame.shape[1], H, W), g[tag + "_K_out"])
