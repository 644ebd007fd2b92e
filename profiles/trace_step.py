"""Per-kernel durations of one workload's step as it really runs (CUDA-graph replay, warm caches, kernels free
to overlap) from the CUPTI activity records behind torch.profiler - ncu serialises and cold-starts every launch.
    python profiles/trace_step.py [workload] [replays]
"""
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "unsupervised-pseuso-lidar_b200")]
import torch  # noqa: E402
import bench  # noqa: E402
from plb200 import synth  # noqa: E402
from losses import Losses  # noqa: E402
from torch.profiler import profile, ProfilerActivity  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
cfg = bench.WORKLOADS[wl]
dev = torch.device("cuda:0")
sets = [synth.to_device(s, dev) for s in bench.make_sets(cfg, 4, 1234, dev)]
crit = Losses(smoothness="edge", deterministic=True) if cfg["variant"] == "live_edge" else Losses()
side = torch.cuda.Stream(device=dev)
with torch.cuda.stream(side):
    for g in sets:
        for _ in range(2):
            bench.step_fn(crit, g, cfg)
torch.cuda.synchronize()
graphs = []
for g in sets:
    cg = torch.cuda.CUDAGraph()
    with torch.cuda.graph(cg, stream=side):
        bench.step_fn(crit, g, cfg)
    graphs.append(cg)
for i in range(8):
    graphs[i % 4].replay()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(reps):
        graphs[i % 4].replay()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
tot = defaultdict(float); cnt = defaultdict(int)
for e in ev:
    tot[e.name[:60]] += e.device_time if hasattr(e, "device_time") else e.cuda_time
    cnt[e.name[:60]] += 1
ev.sort(key=lambda e: e.time_range.start)
span = (ev[-1].time_range.end - ev[0].time_range.start) / reps
print("workload %s: %.1f us per replay (first kernel start -> last kernel end, %d replays)" % (wl, span, reps))
for k in sorted(tot, key=lambda k: -tot[k]):
    print("  %-60s %8.2f us/step  x%d" % (k, tot[k] / reps, cnt[k] // reps))
# timeline of the last replay
n = len(ev) // reps
t0 = ev[-n].time_range.start
for e in ev[-n:]:
    print("    +%7.1f .. +%7.1f  %s" % (e.time_range.start - t0, e.time_range.end - t0, e.name[:50]))
