"""Timing of the stand-alone entry points that are NOT on the fused path (inverse_warp, SSIM / photometric maps,
reconstruct / project, disp_to_depth, smoothness alone) at B=12, 192x640: fwd and fwd+bwd, eager, CUDA events."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "unsupervised-pseuso-lidar_b200")]
import torch
from plb200 import synth, ops
from geometry.pose_geometry import inverse_warp, disp_to_depth
from geometry.transform import Transform
from losses import Losses, SSIM

dev = torch.device("cuda:0")
B, H, W = 12, 192, 640
inp = synth.to_device(synth.make_photo_inputs(B, H, W, n_src=2, n_scales=4, seed=1, n_depth_frames=1), dev)
tgt, ref, K, poses = inp["tgt"], inp["ref_imgs"][0], inp["intrinsics"], inp["poses"]
disp = inp["disparity"][0]
depth = 1.0 / (10.0 * disp[0] + 0.01)
px = B * H * W


def timeit(name, fn, n=30, bytes_per_px=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    extra = "  %.0f GB/s algorithmic" % (bytes_per_px * px / ms / 1e6) if bytes_per_px else ""
    print("%-44s %8.1f us  %8.0f Mpix/s%s" % (name, ms * 1e3, px / ms / 1e3, extra), flush=True)


def warp_fb():
    d = depth.clone().requires_grad_(True)
    p = poses[:, 0].clone().requires_grad_(True)
    out = inverse_warp(ref, d, p, K, False)
    out.sum().backward()


def photo_fb(fn):
    x = ref.clone().requires_grad_(True)
    fn(x, tgt).sum().backward()


crit = Losses()
with torch.no_grad():
    timeit("inverse_warp fwd", lambda: inverse_warp(ref, depth, poses[:, 0], K, False), bytes_per_px=12 + 4 + 12)
timeit("inverse_warp fwd+bwd (depth, pose)", warp_fb, bytes_per_px=2 * (12 + 4 + 12) + 12 + 4)
with torch.no_grad():
    timeit("SSIM.standard_loss fwd", lambda: SSIM().standard_loss(ref, tgt), bytes_per_px=36)
    timeit("compute_photometric_loss (clip) fwd", lambda: crit.compute_photometric_loss(ref, tgt), bytes_per_px=36)
timeit("SSIM.standard_loss fwd+bwd", lambda: photo_fb(lambda a, b: SSIM().standard_loss(a, b)), bytes_per_px=36 + 48)
timeit("compute_photometric_loss fwd+bwd", lambda: photo_fb(crit.compute_photometric_loss), bytes_per_px=36 + 48)
with torch.no_grad():
    timeit("Transform.reconstruct", lambda: Transform().reconstruct(depth[:, 0], K), bytes_per_px=16)
    X = Transform().reconstruct(depth[:, 0], K)
    T = torch.eye(4, device=dev).repeat(B, 1, 1)
    timeit("Transform.project", lambda: Transform().project(X, K, T), bytes_per_px=20)
    timeit("disp_to_depth (4-scale pyramid)", lambda: disp_to_depth([disp]), bytes_per_px=8 * 1.33)
timeit("smooth_loss alone fwd+bwd (4 scales)", lambda: crit.smooth_loss([d.clone().requires_grad_(True) for d in
                                                                          [1.0 / (10.0 * t + 0.01) for t in disp]]).backward(),
       bytes_per_px=12 * 1.33)
