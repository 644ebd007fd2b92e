"""Kernel-only timing of the pseudo-LiDAR launches (config C4): project_batch captured in a CUDA graph per input set
(so host time and the allocator are out of the measurement), replays rotating over three 60 MB depth batches, CUDA
events around the replays.  Prints the two layouts and, with `velo`, the Velodyne -> image scatter."""
import json
import os
import sys
import tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "unsupervised-pseuso-lidar_b200")]
import torch
from plb200 import synth
from utils.PseudoLiDAR import PseudoLiDAR

dev = torch.device("cuda:0")
B, H, W = int(os.environ.get("CB", 32)), 375, 1242
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6534.1)
with tempfile.TemporaryDirectory() as d:
    calib = synth.write_kitti_calib(d)
    pl = PseudoLiDAR(calib, 0, device=dev)
sets = [synth.make_depth_images(B, H, W, seed=40 + k).to(dev) for k in range(3)]
px = B * H * W


def timed(graphs, reps=30):
    for g in graphs:
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        graphs[i % len(graphs)].replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for layout, want, bpp in (("f64", dict(want_f64=True), 32.0), ("f32", dict(want_f64=False, want_f32=True), 16.0)):
    pl.project_batch(sets[0], **want)
    torch.cuda.synchronize()
    graphs, keep = [], []
    for s in sets:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            keep.append(pl.project_batch(s, **want))
        graphs.append(g)
    us = timed(graphs)
    kept = int(keep[0]["count"].sum())
    ab = 4.0 * px + bpp * kept
    print("cloud %s  B=%d  %.1f us  %.0f Mpix/s  kept %d  achieved %.0f GB/s  frac %.3f" %
          (layout, B, us, px / us, kept, ab / us / 1e3, ab / us / 1e3 / peak))
    del graphs, keep

if "velo" in sys.argv:
    from Transform.Transform import Transform
    N = 123577
    with tempfile.TemporaryDirectory() as d:
        tr = Transform(synth.write_kitti_calib(d), W, H, device=dev)
    one = [torch.from_numpy(synth.make_velodyne_cloud(N, seed=60 + k)) for k in range(4)]
    vsets = [torch.stack([one[(k + j) % 4] for j in range(B)]).to(dev) for k in range(3)]
    tr.project_batch(vsets[0])
    torch.cuda.synchronize()
    graphs, keep = [], []
    for s in vsets:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            keep.append(tr.project_batch(s))
        graphs.append(g)
    us = timed(graphs)
    ab = 16.0 * B * N + 8.0 * B * H * W
    print("velo f64  B=%d  %.1f us  %.0f Mpoints/s  achieved %.0f GB/s  frac %.3f" % (B, us, B * N / us, ab / us / 1e3, ab / us / 1e3 / peak))
