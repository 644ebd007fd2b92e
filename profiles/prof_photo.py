"""Profiling driver: a few steps of one workload through the public API
(eager, no graph) so that ncu sees every kernel launch.

    python profiles/prof_photo.py [workload] [steps]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "unsupervised-pseuso-lidar_b200")]
import torch  # noqa: E402
import bench  # noqa: E402
from plb200 import synth  # noqa: E402
from losses import Losses  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "headline"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
cfg = bench.WORKLOADS[wl]
dev = torch.device("cuda:0")
sets = [synth.to_device(s, dev) for s in bench.make_sets(cfg, 3, 1234, dev)]
crit = Losses()
for i in range(steps):
    total, _, _ = bench.step_fn(crit, sets[i % len(sets)], cfg)
torch.cuda.synchronize()
print("ok", wl, float(total))
