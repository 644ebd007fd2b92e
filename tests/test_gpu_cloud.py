"""GPU parity of the pseudo-LiDAR back-projection: bit-exact mask / order /
fp64 xyz against the reference's golden clouds and the oracle."""
import hashlib

import numpy as np
import pytest
import torch

from helpers import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def calib(tmp_path_factory):
    from plb200 import synth
    return synth.write_kitti_calib(str(tmp_path_factory.mktemp("calib")))


def test_small_golden_bit_exact(calib):
    from utils.PseudoLiDAR import PseudoLiDAR
    g = load_golden("cloud_kitti")
    for sp in (0, 3):
        pl = PseudoLiDAR(calib, sp)
        assert np.array_equal(pl.T, g["T"]) and np.array_equal(pl.P, g["P"])
        cloud = pl.project_PL(g["small_depth"])
        ref = g["small_cloud_sp%d" % sp]
        assert cloud.dtype == np.float64 and cloud.shape == ref.shape
        assert np.array_equal(cloud, ref)
        assert np.all(cloud[:, 3] == 0.0)


def test_full_frame_golden_checksum(calib):
    from utils.PseudoLiDAR import PseudoLiDAR
    from plb200 import synth
    g = load_golden("cloud_kitti")
    full = synth.make_depth_images(1, 375, 1242, seed=int(g["full_seed"]))[0]
    cloud = PseudoLiDAR(calib, 0).project_PL(full.cuda())
    assert cloud.shape[0] == int(g["full_count"])
    digest = np.frombuffer(hashlib.sha256(np.ascontiguousarray(cloud).tobytes()).digest(), dtype=np.uint8)
    assert np.array_equal(digest, g["full_sha256"])
    assert PseudoLiDAR(calib, 10).project_PL(full.cuda()).shape[0] == int(g["full_count_sp10"])


# (3, 33, 35): an odd pixel count - the images start on 16-, 4- and 8-byte boundaries (every alignment class of the
# count launch) and the last tile is ragged; (2, 1, 5): narrower than a warp step -> the compacting write launch
@pytest.mark.parametrize("B,H,W,sp", [(3, 37, 124, 0), (2, 1, 5, 0), (4, 375, 1242, 0), (2, 100, 333, 7), (3, 33, 35, 0),
                                      (2, 64, 1030, 0)])
def test_batch_against_oracle(calib, B, H, W, sp):
    from utils.PseudoLiDAR import PseudoLiDAR
    from plb200 import synth
    from oracle import restated as O
    pl = PseudoLiDAR(calib, sp)
    depth = synth.make_depth_images(B, H, W, seed=5, lo=-5.0, hi=70.0)   # negatives: x<0 is masked out
    depth[0, 0, 0] = float("nan")
    if W > 40:
        depth[0, H // 2, 33] = float("inf")      # the write launch's exact-division path for one 128-pixel group
        depth[B - 1, H - 1, W - 1] = 3.0e38
    res = pl.project_batch(depth, want_f64=True, want_f32=True, want_index=True, want_valid=True)
    counts = res["count"].cpu().numpy()
    for b in range(B):
        ref, valid = O.project_PL(depth[b].numpy(), pl.T, pl.P, sparsity=sp, return_valid=True)
        n = int(counts[b])
        assert n == ref.shape[0]
        assert np.array_equal(res["valid"][b].cpu().numpy().astype(bool), valid)       # mask bit-exact
        got = res["cloud_f64"][b, :n].cpu().numpy()
        assert np.array_equal(got, ref, equal_nan=True)                                   # xyz bit-exact (tolerance 1e-6 allowed)
        idx = np.flatnonzero(valid)[::sp] if sp else np.flatnonzero(valid)
        assert np.array_equal(res["index"][b, :n].cpu().numpy(), idx)                   # order-preserving index
        assert np.allclose(res["cloud_f32"][b, :n].cpu().numpy(), ref.astype(np.float32), rtol=1e-6, equal_nan=True)


def test_empty_cloud(calib):
    from utils.PseudoLiDAR import PseudoLiDAR
    depth = -torch.ones(8, 16)
    cloud = PseudoLiDAR(calib, 0).project_PL(depth)
    assert cloud.shape == (0, 4)


@pytest.mark.parametrize("sp", [0, 1, 4])
def test_general_homogeneous_row(calib, sp):
    """`plb_cloud_project` takes any 4x4: with a non-zero last row of T_inv the 4th output column is computed by the
    same chain as x, y, z (the reference's own matrix has an all-zero row, PseudoLiDAR.py:43).  sparsity 0 / 1: the
    direct write launch; 4: the compacting one."""
    from utils.PseudoLiDAR import PseudoLiDAR
    from plb200 import ops, synth
    from oracle import restated as O
    pl = PseudoLiDAR(calib, sp)
    Tinv = np.array(pl.inverse_rigid_trans(pl.T), dtype=np.float64)
    Tinv[3] = [0.25, -0.5, 0.125, 2.0]
    depth = synth.make_depth_images(2, 48, 200, seed=9, lo=-2.0, hi=60.0)
    res = ops.cloud_project(depth.cuda(), pl.P, Tinv, sparsity=sp)
    counts = res["count"].cpu().numpy()
    P = pl.P
    for b in range(2):
        rows, cols = depth[b].shape
        c, r = np.meshgrid(np.arange(cols), np.arange(rows))
        d = depth[b].numpy().astype(np.float64).reshape(-1)
        pts = np.ones((d.size, 4))
        pts[:, 0] = ((c.reshape(-1) - P[0, 2]) * d) / P[0, 0] + P[0, 3] / (-P[0, 0])
        pts[:, 1] = ((r.reshape(-1) - P[1, 2]) * d) / P[1, 1] + P[1, 3] / (-P[1, 1])
        pts[:, 2] = d
        cloud = np.matmul(pts, Tinv.T)
        ref = cloud[(cloud[:, 0] >= 0) & (cloud[:, 2] < 1)]
        if sp:
            ref = ref[0::sp]
        n = int(counts[b])
        assert n == ref.shape[0]
        got = res["cloud_f64"][b, :n].cpu().numpy()
        assert np.array_equal(got[:, :3], ref[:, :3])
        err = np.abs(got[:, 3] - ref[:, 3]).max() if n else 0.0
        print("4th column max abs err %.3e" % err)
        assert err <= 1e-12 * max(1.0, np.abs(ref[:, 3]).max())
