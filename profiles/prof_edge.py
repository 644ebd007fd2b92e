"""Profiling driver for the edge-aware smoothness kernels at the C5 shape (B=64, 192x640, 4 scales)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "unsupervised-pseuso-lidar_b200")]
import torch
from plb200 import synth, ops
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
inp = synth.to_device(synth.make_photo_inputs(B, 192, 640, n_src=1, n_scales=4, seed=3, n_depth_frames=1), dev)
for i in range(3):
    disp = [d.detach().requires_grad_(True) for d in inp["disparity"][0]]
    loss = ops.edge_aware_smooth(disp, inp["tgt"], normalize=True)
    loss.backward()
torch.cuda.synchronize()
print("ok", float(loss))
