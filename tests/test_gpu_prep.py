"""GPU parity of the loader-side frame preparation: `plb_prep_frames` against golden
vectors produced by the UNMODIFIED `KittiDataset.load_img` + the transform list of trainer.py:97-103
(tests/golden/make_golden.py::prep_case) and against the oracle restatement at full KITTI size.  Bit-exact."""
import hashlib

import numpy as np
import pytest
import torch

from helpers import load_golden

pytestmark = pytest.mark.gpu


def _run(frames, H, W, K=None, **kw):
    from plb200.frameprep import FramePrep
    dev = torch.device("cuda:0")
    f = torch.from_numpy(np.ascontiguousarray(frames)).to(dev)
    k = None if K is None else torch.from_numpy(K).to(dev)
    res = FramePrep(H, W, **kw)(f, k)
    torch.cuda.synchronize()
    return {n: t.cpu().numpy() for n, t in res.items()}


@pytest.mark.parametrize("tag", ["down", "same", "up", "odd"])
def test_golden_frames_bit_exact(tag):
    g = load_golden("prep_frames")
    H, W = (int(v) for v in g[tag + "_size"])
    res = _run(g[tag + "_frame"][None], H, W, g[tag + "_K_in"][None], nhwc4=True)
    assert np.array_equal(res["planar"][0], g[tag + "_out"]), np.abs(res["planar"][0] - g[tag + "_out"]).max()
    assert np.array_equal(res["nhwc4"][0, :, :, :3], np.transpose(g[tag + "_out"], (1, 2, 0)))
    assert np.all(res["nhwc4"][..., 3] == 0)
    assert np.array_equal(res["K"][0], g[tag + "_K_out"])


def test_full_kitti_frame_checksum_and_batch():
    """375x1242 -> 192x640: sha256 of the reference's output; a batch of different frames against the oracle."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_golden import make_frame
    from oracle import restated as O
    g = load_golden("prep_frames")
    frame = make_frame(375, 1242, seed=int(g["full_seed"]))
    res = _run(frame[None], 192, 640)
    assert np.array_equal(res["planar"][0][:, [0, 95, 191], :], g["full_rows"])
    digest = np.frombuffer(hashlib.sha256(np.ascontiguousarray(res["planar"][0]).tobytes()).digest(), dtype=np.uint8)
    assert np.array_equal(digest, g["full_sha256"])
    frames = np.stack([make_frame(120, 200, seed=20 + k) for k in range(5)])
    res = _run(frames, 64, 96)
    for k in range(5):
        assert np.array_equal(res["planar"][k], O.load_img_chain(frames[k], 64, 96))


def test_round_trip_table_and_errors():
    """Every byte value through the float32 round trip in front of ToPILImage; argument errors are exceptions."""
    from plb200 import ops, _lib
    v = np.arange(256, dtype=np.uint8)
    frame = np.repeat(v[:, None, None], 3, axis=2).reshape(16, 16, 3)
    res = _run(frame[None], 16, 16)
    back = (torch.from_numpy(v.astype(np.float32) / 255.0).mul(255).byte().numpy().astype(np.float32) / np.float32(255.0))
    exp = (back - np.float32(0.485)) / np.float32(0.229)
    assert np.array_equal(res["planar"][0, 0].reshape(-1), exp)
    with pytest.raises(ValueError):
        ops.prep_frames(torch.zeros(1, 4, 4, 3, device="cuda"), 4, 4)
    with pytest.raises(_lib.PlbError):
        ops.prep_frames(torch.zeros(1, 400, 400, 3, dtype=torch.uint8, device="cuda"), 4, 4)      # reduction > 15x


def test_intrinsics_per_sample_not_per_frame():
    """A batch of 3 B frames (target + two sources per sample) with ONE matrix per sample: only those B matrices are
    read and written (plb_prep_args.n_K) - the words behind K_out stay untouched - and more matrices than frames is
    an error, not an out-of-bounds read."""
    from plb200 import ops
    dev = torch.device("cuda:0")
    B = 4
    g = torch.Generator().manual_seed(2)
    frames = torch.randint(0, 256, (3 * B, 40, 64, 3), dtype=torch.uint8, generator=g).to(dev)
    K = torch.rand(B, 3, 3, dtype=torch.float64, generator=g).to(dev)
    res = ops.prep_frames(frames, 20, 32, K=K)
    per_frame = ops.prep_frames(frames, 20, 32, K=K.repeat(3, 1, 1))
    torch.cuda.synchronize()
    assert res["K"].shape == (B, 3, 3)
    assert torch.equal(res["K"], per_frame["K"][:B]) and torch.equal(res["planar"], per_frame["planar"])
    exp = K.clone()
    exp[:, 0] *= 32 / 64
    exp[:, 1] *= 20 / 40
    assert torch.equal(res["K"], exp)
    with pytest.raises(ValueError):
        ops.prep_frames(frames[:2], 20, 32, K=K)
