"""Profiling driver for the pseudo-LiDAR kernels (config C4): a few batches through the public API."""
import os
import sys
import tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "unsupervised-pseuso-lidar_b200")]
import torch
from plb200 import synth
from utils.PseudoLiDAR import PseudoLiDAR
dev = torch.device("cuda:0")
with tempfile.TemporaryDirectory() as d:
    pl = PseudoLiDAR(synth.write_kitti_calib(d), 0, device=dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
sets = [synth.make_depth_images(B, 375, 1242, seed=40 + k).to(dev) for k in range(2)]
for i in range(4):
    out = pl.project_batch(sets[i % 2])
torch.cuda.synchronize()
print("ok", int(out["count"].sum()))
