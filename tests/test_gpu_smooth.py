"""GPU parity of the stand-alone second-order smoothness (losses.py:242-260): the vector kernel (every width a multiple
of 4, 16-byte aligned maps: four columns per lane) and the scalar kernel (any width / alignment) against the oracle and
against each other, over strip / chunk boundaries and every rows-per-chunk choice of the launch."""
import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu

LOSS_TOL = 1e-5
GRAD_TOL = 1e-4


def _maps(shapes, B, seed, depth=True):
    g = torch.Generator().manual_seed(seed)
    out = []
    for h, w in shapes:
        yy = torch.linspace(0, 3.0, h).view(1, 1, h, 1)
        xx = torch.linspace(0, 5.0, w).view(1, 1, 1, w)
        base = 0.35 + 0.25 * torch.sin(xx + 0.7 * yy) * torch.cos(1.3 * yy) + 0.05 * torch.rand(B, 1, h, w, generator=g)
        out.append((1.0 / (10.0 * base + 0.01)) if depth else base)
    return out


def _misaligned(t):
    """The same values in a contiguous tensor whose storage starts 4 bytes off a 16-byte boundary (-> scalar kernel)."""
    buf = torch.empty(t.numel() + 1, dtype=t.dtype, device=t.device)
    v = buf[1:].view(t.shape)
    v.copy_(t)
    assert v.is_contiguous() and v.data_ptr() % 16 != 0
    return v


def _ours(maps, dev, misalign=False):
    from losses import Losses
    ms = [m.to(dev) for m in maps]
    if misalign:
        ms = [_misaligned(m) for m in ms]
    ms = [m.requires_grad_(True) for m in ms]
    loss = Losses().smooth_loss(ms)
    loss.backward()
    return loss, [m.grad for m in ms]


def _oracle(maps):
    from oracle import restated as O
    ms = [m.clone().requires_grad_(True) for m in maps]
    loss = O.smooth_loss(ms)
    loss.backward()
    return loss, [m.grad for m in ms]


CASES = [
    # (B, shapes, what)
    (2, [(40, 248), (20, 124), (9, 60), (5, 8)], "vector: 3 strips, chunk boundaries, rows 4"),
    (3, [(33, 47), (17, 23), (3, 3)], "scalar: odd widths"),
    (1, [(7, 4), (3, 120), (19, 124)], "vector: one-lane strip, exactly 120 / 124 columns"),
    (32, [(128, 600)], "vector: rows 8"),
    (48, [(128, 600)], "vector: rows 16"),
]


@pytest.mark.parametrize("B,shapes,what", CASES, ids=[c[2] for c in CASES])
def test_smooth_matches_oracle(B, shapes, what):
    dev = torch.device("cuda:0")
    maps = _maps(shapes, B, seed=5 + B)
    lo, go = _oracle(maps)
    l, g = _ours(maps, dev)
    err = abs(float(l) - float(lo)) / abs(float(lo))
    print(what, "loss rel err %.2e" % err, "grad rel err", ["%.2e" % rel_err(a.cpu(), b) for a, b in zip(g, go)])
    assert err <= LOSS_TOL
    for a, b in zip(g, go):
        assert rel_err(a.cpu(), b) < GRAD_TOL


def test_vector_and_scalar_kernels_agree():
    """Same maps through both kernels (the misaligned copy forces the scalar one): same anchors and signs, so the
    gradients are equal bit for bit and the losses differ only by the order of the fp32 partial sums."""
    dev = torch.device("cuda:0")
    maps = _maps([(48, 160), (24, 80), (12, 40), (6, 20)], 5, seed=11)
    lv, gv = _ours(maps, dev)
    ls, gs = _ours(maps, dev, misalign=True)
    assert abs(float(lv) - float(ls)) <= 2e-6 * abs(float(ls))
    for a, b in zip(gv, gs):
        assert torch.equal(a, b)


def test_smooth_repeatable():
    dev = torch.device("cuda:0")
    maps = _maps([(64, 256), (32, 128)], 6, seed=3)
    l0, g0 = _ours(maps, dev)
    for _ in range(3):
        l1, g1 = _ours(maps, dev)
        assert torch.equal(l0, l1)
        for a, b in zip(g0, g1):
            assert torch.equal(a, b)
