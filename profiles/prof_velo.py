"""Profiling driver for the Velodyne -> image kernels: a few batches of 32 sweeps through the public API."""
import os
import sys
import tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "unsupervised-pseuso-lidar_b200")]
import torch
from plb200 import synth
from Transform.Transform import Transform
dev = torch.device("cuda:0")
with tempfile.TemporaryDirectory() as d:
    tr = Transform(synth.write_kitti_calib(d), 1242, 375, device=dev)
one = [torch.from_numpy(synth.make_velodyne_cloud(123577, seed=60 + k)) for k in range(4)]
sets = [torch.stack([one[(k + j) % 4] for j in range(32)]).to(dev) for k in range(2)]
for i in range(4):
    out = tr.project_batch(sets[i % 2])
torch.cuda.synchronize()
print("ok", float(out["depth_f64"].sum()))
