"""Edge-aware smoothness: ABSENT from the reference, PARITY UNPINNED.
The CUDA kernels are held to our own torch restatement (oracle/restated.py::edge_aware_smooth_loss)."""
import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,H,W,S", [(1, 8, 16, 1), (2, 32, 64, 3), (3, 48, 80, 4), (2, 192, 640, 4)])
@pytest.mark.parametrize("normalize", [True, False])
def test_edge_aware_vs_oracle(B, H, W, S, normalize):
    from losses import Losses
    from plb200 import synth
    from oracle import restated as O
    inp = synth.make_photo_inputs(B, H, W, n_src=1, n_scales=S, seed=500 + H, n_depth_frames=1)
    tgt, disp = inp["tgt"], inp["disparity"][0]
    rd = [d.clone().requires_grad_(True) for d in disp]
    rl = O.edge_aware_smooth_loss(rd, tgt, normalize)
    (1.7 * rl).backward()
    dev = torch.device("cuda:0")
    gd = [d.to(dev).requires_grad_(True) for d in disp]
    loss = Losses().edge_aware_smooth_loss(gd, tgt.to(dev), normalize=normalize)
    (1.7 * loss).backward()
    assert abs(float(loss) - float(rl)) <= 1e-5 * abs(float(rl))
    for a, b in zip(gd, rd):
        assert rel_err(a.grad.cpu(), b.grad) < 1e-4


def test_edge_aware_repeatable_and_single_map():
    from losses import Losses
    from plb200 import synth
    inp = synth.make_photo_inputs(2, 64, 96, n_src=1, n_scales=2, seed=9, n_depth_frames=1)
    dev = torch.device("cuda:0")
    outs = []
    for _ in range(3):
        gd = [d.to(dev).requires_grad_(True) for d in inp["disparity"][0]]
        loss = Losses().edge_aware_smooth_loss(gd, inp["tgt"].to(dev))
        loss.backward()
        outs.append((loss.detach().clone(), [d.grad.clone() for d in gd]))
    for o in outs[1:]:
        assert torch.equal(o[0], outs[0][0])
        for a, b in zip(o[1], outs[0][1]):
            assert torch.equal(a, b)
    one = Losses().edge_aware_smooth_loss(inp["disparity"][0][0].to(dev), inp["tgt"].to(dev), normalize=False)
    assert float(one) > 0.0
