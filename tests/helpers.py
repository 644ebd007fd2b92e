"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def golden_inputs(g, n_src=2):
    """Rebuild the reference's argument layout from a live/dormant fixture."""
    tgt = torch.from_numpy(g["tgt"])
    refs = [torch.from_numpy(g["ref%d" % i]) for i in range(n_src)]
    disparity = []
    f = 0
    while "disp_f%d_s0" % f in g:
        frame, s = [], 0
        while "disp_f%d_s%d" % (f, s) in g:
            frame.append(torch.from_numpy(g["disp_f%d_s%d" % (f, s)]))
            s += 1
        disparity.append(frame)
        f += 1
    return tgt, refs, disparity, torch.from_numpy(g["poses"]), torch.from_numpy(g["K"])


def rel_err(a, b):
    """Norm-relative error |a-b| / |b| (the 'relative' of north_star's tolerances)."""
    a = torch.as_tensor(a, dtype=torch.float64).reshape(-1)
    b = torch.as_tensor(b, dtype=torch.float64).reshape(-1)
    denom = float(b.norm())
    if denom == 0.0:
        return float((a - b).norm())
    return float((a - b).norm()) / denom


def max_rel_err(a, b):
    """max|a-b| / max|b|."""
    a = torch.as_tensor(a, dtype=torch.float64).reshape(-1)
    b = torch.as_tensor(b, dtype=torch.float64).reshape(-1)
    denom = float(b.abs().max())
    return float((a - b).abs().max()) / (denom if denom > 0 else 1.0)
