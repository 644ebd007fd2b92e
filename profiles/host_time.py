"""Host (Python) time of one eager fwd+bwd step through the public API against its GPU time."""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "unsupervised-pseuso-lidar_b200")]
import torch
import bench
from plb200 import synth
from losses import Losses
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
cfg = bench.WORKLOADS[wl]
dev = torch.device("cuda:0")
g = synth.to_device(bench.make_sets(cfg, 1, 1234, dev)[0], dev)
crit = Losses()
for _ in range(20):
    bench.step_fn(crit, g, cfg)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200):
    bench.step_fn(crit, g, cfg)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("%s: host %.1f us per eager step to issue, %.1f us per step including the drain" % (wl, (t1 - t0) / 200 * 1e6, (t2 - t0) / 200 * 1e6))
if len(sys.argv) > 2 and sys.argv[2] == "profile":
    import cProfile
    import pstats
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(200):
        bench.step_fn(crit, g, cfg)
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("tottime").print_stats(14)
