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


def test_edge_inside_the_fused_step():
    """Losses(smoothness="edge"): the edge-aware term evaluated inside the fused call (its gradient accumulated into the
    photometric term's maps) against the two separate calls - weighted losses, so the guarded relaunch of the backward
    pass runs - and through the captured step."""
    from losses import Losses
    from plb200 import ops, synth
    dev = torch.device("cuda:0")
    inp = synth.to_device(synth.make_photo_inputs(3, 64, 128, n_src=2, n_scales=3, seed=21), dev)

    def leaves():
        return [[d.detach().clone().requires_grad_(True) for d in fr] for fr in inp["disparity"]], \
            inp["poses"].detach().clone().requires_grad_(True)

    d1, p1 = leaves()
    mam, sm = Losses(smoothness="edge").forward(inp["tgt"], inp["ref_imgs"], d1, p1, inp["intrinsics"], None)
    (1.3 * mam + 0.7 * sm).backward()
    d2, p2 = leaves()
    mam2, _ = ops.fused_losses(inp["tgt"], inp["ref_imgs"], d2, p2, inp["intrinsics"], do_smooth=False)
    sm2 = Losses().edge_aware_smooth_loss(d2[0], inp["tgt"])
    (1.3 * mam2 + 0.7 * sm2).backward()
    assert torch.equal(mam, mam2) and abs(float(sm) - float(sm2)) <= 1e-6 * abs(float(sm2))
    assert rel_err(p1.grad, p2.grad) < 1e-6
    for fa, fb in zip(d1, d2):
        for a, b in zip(fa, fb):
            assert rel_err(a.grad, b.grad) < 1e-5, rel_err(a.grad, b.grad)
    # unit upstream (the forward launch's gradients stand) and the captured step
    d3, p3 = leaves()
    crit = Losses(smoothness="edge")
    l3 = crit.forward(inp["tgt"], inp["ref_imgs"], d3, p3, inp["intrinsics"], None)
    (l3[0] + l3[1]).backward()
    step = crit.capture(inp["tgt"], inp["ref_imgs"], inp["disparity"], inp["poses"], inp["intrinsics"])
    loss, grads = step()
    assert torch.equal(loss[0], l3[0]) and torch.equal(loss[1], l3[1])
    for fa, fb in zip(grads.disparity, d3):
        for a, b in zip(fa, fb):
            assert torch.equal(a, b.grad)
