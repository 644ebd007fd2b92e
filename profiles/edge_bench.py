"""Timing of the edge-aware smoothness (north_star row a17) at the C5 shape: B=64, 192x640, 4 scales, fwd+bwd."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "unsupervised-pseuso-lidar_b200")]
import torch
from plb200 import synth, ops
dev = torch.device("cuda:0")
for B in (12, 64):
    inp = synth.to_device(synth.make_photo_inputs(B, 192, 640, n_src=2, n_scales=4, seed=1, n_depth_frames=1), dev)
    disp = [d.clone().requires_grad_(True) for d in inp["disparity"][0]]
    def step():
        for d in disp:
            d.grad = None
        loss = ops.edge_aware_smooth(disp, inp["tgt"], normalize=True)
        loss.backward()
        return loss
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    torch.cuda.synchronize()
    px = B * 192 * 640
    ms = e0.elapsed_time(e1) / n
    # algorithmic bytes: target 12 B/px read + pyramid (read 4, write 4) * 1.33
    ab = px * (12 + 8 * 1.328125)
    print("edge-aware smoothness B=%d: %.1f us per fwd+bwd (eager), %.0f Mpix/s, %.0f GB/s algorithmic" % (B, ms * 1e3, px / ms / 1e3, ab / ms / 1e6))
