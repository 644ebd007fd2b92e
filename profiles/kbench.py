"""Kernel-only timing of the fused photometric launch on a few workloads (tuning helper).
    python profiles/kbench.py [workload ...]      e.g. headline c2 headline64
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "unsupervised-pseuso-lidar_b200")]
import torch  # noqa: E402
import bench  # noqa: E402
from losses import Losses  # noqa: E402

bench.WORKLOADS["headline64"] = dict(B=64, H=192, W=640, n_src=2, n_scales=1, variant="dir0")
bench.WORKLOADS["headline4"] = dict(B=4, H=192, W=640, n_src=2, n_scales=1, variant="dir0")
# single-direction pieces of the live composition (tuning the cost model of the unit list): d<n_src>s<n_scales>
for _ns in (1, 2, 3):
    for _sc in (1, 2, 3, 4):
        bench.WORKLOADS["d%ds%d" % (_ns, _sc)] = dict(B=12, H=192, W=640, n_src=_ns, n_scales=_sc, variant="dir0")
dev = torch.device("cuda:0")
for wl in (sys.argv[1:] or ["headline", "c2", "headline64"]):
    cfg = bench.WORKLOADS[wl]
    n_sets = int(os.environ.get("KB_SETS", 4 if cfg["B"] <= 16 else 2))
    sets = [bench.synth.to_device(s, dev) for s in bench.make_sets(cfg, n_sets, 1234, dev)] if hasattr(bench, "synth") else None
    if sets is None:
        from plb200 import synth
        sets = [synth.to_device(s, dev) for s in bench.make_sets(cfg, n_sets, 1234, dev)]
    ms = min(bench.time_photo_kernel(Losses(), sets, cfg, dev, 200) for _ in range(3))
    px = cfg["B"] * cfg["H"] * cfg["W"]
    ab = bench.algorithmic_bytes_per_px(cfg) * px
    print("%-11s kernel %.2f us  %.0f Mpix/s  achieved %.0f GB/s  frac %.3f" % (
        wl, ms * 1e3, px / ms / 1e3, ab / ms / 1e6, ab / ms / 1e6 / 6534.1), flush=True)
