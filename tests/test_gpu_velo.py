"""GPU parity of the Velodyne -> image projection (pseudo-lidar/Transform/Transform.py:69-104):
bit-exact depth and winner index against the reference's golden images and the oracle."""
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
    from Transform.Transform import Transform
    g = load_golden("velo_kitti")
    tr = Transform(calib, 124, 37)
    assert np.array_equal(tr.T, g["T"])
    tr.P = g["small_P"]
    depth = tr.project_velo_to_img(g["small_points"])
    assert depth.dtype == np.float64 and depth.shape == (37, 124)
    assert np.array_equal(depth, g["small_depth"])


def test_full_sweep_golden_checksum(calib):
    from Transform.Transform import Transform
    from plb200 import synth
    g = load_golden("velo_kitti")
    tr = Transform(calib, 1242, 375)
    assert np.array_equal(tr.P, g["P"])
    full = synth.make_velodyne_cloud(int(g["full_n"]), seed=int(g["full_seed"]))
    depth = tr.project_velo_to_img(torch.from_numpy(full).cuda())
    digest = np.frombuffer(hashlib.sha256(np.ascontiguousarray(depth).tobytes()).digest(), dtype=np.uint8)
    assert np.array_equal(digest, g["full_sha256"])
    nz = np.flatnonzero(depth)
    assert np.array_equal(nz, g["full_nz_index"]) and np.array_equal(depth.reshape(-1)[nz], g["full_nz_value"])


@pytest.mark.parametrize("B,N,C,W,H,scale", [(3, 5000, 4, 124, 37, 0.1), (2, 777, 3, 40, 24, 0.03),
                                             (2, 123577, 4, 1242, 375, 1.0), (1, 1, 5, 8, 8, 0.01)])
def test_batch_against_oracle(calib, B, N, C, W, H, scale):
    """Ragged batch (per-sweep counts), 3 / 4 / 5 floats per point, heavy collisions, repeatable, and the
    self-cleaning workspace (second call on the same buffers equals the first)."""
    from Transform.Transform import Transform
    from plb200 import synth
    from oracle import restated as O
    tr = Transform(calib, W, H)
    tr.P = tr.P * np.array([[scale], [scale], [1.0]])
    clouds = np.stack([synth.make_velodyne_cloud(N, seed=300 + b) for b in range(B)])
    if C == 3:
        clouds = clouds[:, :, :3].copy()
    elif C == 5:
        clouds = np.concatenate([clouds, np.zeros((B, N, 1), np.float32)], 2)
    counts = [max(N - 13 * b, 0) for b in range(B)]
    dev = torch.from_numpy(clouds).cuda()
    for rep in range(2):
        res = tr.project_batch(dev, counts=torch.tensor(counts), want_f64=True, want_f32=True, want_winner=True)
        for b in range(B):
            ref, win = O.project_velo_to_img(clouds[b, :counts[b]], tr.T, tr.P, W, H)
            assert np.array_equal(res["depth_f64"][b].cpu().numpy(), ref)                   # bit-exact fp64
            assert np.array_equal(res["winner"][b].cpu().numpy(), win)                      # same point owns each cell
            assert np.array_equal(res["depth_f32"][b].cpu().numpy(), ref.astype(np.float32))


def test_empty_sweep(calib):
    from Transform.Transform import Transform
    tr = Transform(calib, 32, 16)
    depth = tr.project_velo_to_img(np.zeros((0, 4), np.float32))
    assert depth.shape == (16, 32) and not depth.any()
    behind = np.array([[-5.0, 0.0, 0.0, 0.0]], np.float32)           # x <= 0: dropped
    assert not tr.project_velo_to_img(behind).any()
