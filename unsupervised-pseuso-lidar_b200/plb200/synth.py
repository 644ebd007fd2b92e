"""Seeded synthetic KITTI-shaped inputs for tests and bench.

There is no KITTI in the build container or on the GPU box, so every parity
test and every bench line runs on these frames.  The layouts are the ones the
reference trainer hands to the loss (`trainer.py:290-313`,
`dataloaders.py:74-120`):

  tgt        [B,3,H,W] f32  ImageNet-normalised (`trainer.py:102`)
  ref_imgs   list of n_src x [B,3,H,W] f32
  disparity  list[frame in (tgt, ref0)] of list[scale] of [B,1,H/2^s,W/2^s] f32
  poses      [B,n_src,6] f32   (rot3 | trans3)
  intrinsics [B,3,3] f64       KITTI P_rect_02[:, :3] scaled to (H, W)
"""
import math

import numpy as np
import torch

# KITTI 2011_09_26 P_rect_02 (fx, cx, fy, cy) at 1242x375 and its 4th column.
KITTI_FX, KITTI_FY = 721.5377, 721.5377
KITTI_CX, KITTI_CY = 609.5593, 172.854
KITTI_P_RECT_02 = np.array([
    [7.215377e+02, 0.000000e+00, 6.095593e+02, 4.485728e+01],
    [0.000000e+00, 7.215377e+02, 1.728540e+02, 2.163791e-01],
    [0.000000e+00, 0.000000e+00, 1.000000e+00, 2.745884e-03]], dtype=np.float64)
# velodyne -> camera rigid transform (R | T), the values the reference's
# notebook embeds (`pseudo-lidar/PL_development/fast_matrix_mul.ipynb` cell 1).
KITTI_VELO_TO_CAM_R = np.array([
    [7.533745e-03, -9.999714e-01, -6.166020e-04],
    [1.480249e-02, 7.280733e-04, -9.998902e-01],
    [9.998621e-01, 7.523790e-03, 1.480755e-02]], dtype=np.float64)
KITTI_VELO_TO_CAM_T = np.array([-4.069766e-03, -7.631618e-02, -2.717806e-01], dtype=np.float64)

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def kitti_intrinsics(B, H, W, dtype=torch.float64):
    """[B,3,3] intrinsics: P_rect_02[:, :3] with row 0 x W/1242, row 1 x H/375
    (`dataloaders.py:95-98`)."""
    K = torch.tensor([[KITTI_FX, 0.0, KITTI_CX],
                      [0.0, KITTI_FY, KITTI_CY],
                      [0.0, 0.0, 1.0]], dtype=torch.float64)
    K[0] *= W / 1242.0
    K[1] *= H / 375.0
    return K.to(dtype).unsqueeze(0).repeat(B, 1, 1).contiguous()


def _field(B, H, W, gen, shift=(0, 0), noise=0.1):
    """Smooth colour field + noise, ImageNet-normalised; `shift` moves the
    smooth part so that source frames are shifted copies of the target."""
    v = torch.arange(H, dtype=torch.float32).view(1, 1, H, 1) + shift[1]
    u = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W) + shift[0]
    phase = torch.tensor([0.0, 2.1, 4.2]).view(1, 3, 1, 1)
    smooth = 0.5 + 0.25 * torch.sin(2 * math.pi * u / 97.0 + phase) * torch.cos(2 * math.pi * v / 61.0)
    x = (smooth + noise * torch.randn(B, 3, H, W, generator=gen)).clamp_(0.0, 1.0)
    mean = torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD).view(1, 3, 1, 1)
    return ((x - mean) / std).contiguous()


def make_photo_inputs(B, H, W, n_src=2, n_scales=1, seed=1234, regime="trained",
                      n_depth_frames=2, device="cpu", noise=0.1):
    """One batch in the reference's sample layout.  `regime`:
      "trained"  a road-scene-like disparity: smooth ground-plane ramp (far at the
                 top, near at the bottom) with smooth bumps and 2% multiplicative
                 noise, range ~[0.003, 0.1] (depth 1..30 m) - what a converged
                 depth net emits, and the regime the bench runs;
      "noise"    per-pixel U(0.002, 0.1): no spatial coherence at all (adversarial
                 for gather locality, used by parity tests);
      "init"     sigmoid(N(0,1)) per pixel (`models/depth/disp_net.py:25-29`)."""
    gen = torch.Generator().manual_seed(seed)
    shifts = [(-3, -1), (3, 1), (-6, -2), (6, 2)]
    tgt = _field(B, H, W, gen, noise=noise)
    refs = [_field(B, H, W, gen, shifts[i % 4], noise=noise) for i in range(n_src)]
    disparity = []
    for _ in range(n_depth_frames):
        per_scale = []
        for s in range(n_scales):
            hs, ws = H >> s, W >> s
            if regime == "init":
                d = torch.sigmoid(torch.randn(B, 1, hs, ws, generator=gen))
            elif regime == "noise":
                d = 0.002 + 0.098 * torch.rand(B, 1, hs, ws, generator=gen)
            else:
                v = (torch.arange(hs, dtype=torch.float32).view(1, 1, hs, 1) + 0.5) / hs
                u = (torch.arange(ws, dtype=torch.float32).view(1, 1, 1, ws) + 0.5) / ws
                ph = 6.28318 * torch.rand(B, 1, 1, 1, generator=gen)
                ramp = 0.004 + 0.085 * v.clamp(min=0.35).sub(0.35).div(0.65) ** 1.5
                bumps = 0.004 * torch.sin(9.0 * u + ph) * torch.cos(5.0 * v + 0.5 * ph)
                d = (ramp + bumps + 0.004) * (1.0 + 0.02 * torch.randn(B, 1, hs, ws, generator=gen))
                d = d.clamp(0.002, 0.1)
            per_scale.append(d.contiguous())
        disparity.append(per_scale)
    poses = torch.cat([0.01 * torch.randn(B, n_src, 3, generator=gen),
                       0.05 * torch.randn(B, n_src, 3, generator=gen)], dim=2).contiguous()
    K = kitti_intrinsics(B, H, W)
    out = dict(tgt=tgt, ref_imgs=refs, disparity=disparity, poses=poses, intrinsics=K)
    if device != "cpu":
        out = to_device(out, device)
    return out


def make_frames_u8(B, H, W, n_frames=3, seed=1234, noise=0.1):
    """Decoded camera frames as the loader sees them: n_frames x [B,H,W,3] uint8 RGB (target, then sources) - the same
    smooth colour field + noise as `_field`, quantised to bytes, sources shifted against the target."""
    gen = torch.Generator().manual_seed(seed)
    shifts = [(0, 0), (-3, -1), (3, 1), (-6, -2), (6, 2)]
    frames = []
    for k in range(n_frames):
        sh = shifts[k % 5]
        v = torch.arange(H, dtype=torch.float32).view(1, H, 1, 1) + sh[1]
        u = torch.arange(W, dtype=torch.float32).view(1, 1, W, 1) + sh[0]
        phase = torch.tensor([0.0, 2.1, 4.2]).view(1, 1, 1, 3)
        smooth = 0.5 + 0.25 * torch.sin(2 * math.pi * u / 97.0 + phase) * torch.cos(2 * math.pi * v / 61.0)
        x = (smooth + noise * torch.randn(B, H, W, 3, generator=gen)).clamp_(0.0, 1.0)
        frames.append((x * 255.0).round().to(torch.uint8).contiguous())
    return frames


def to_device(x, device):
    if isinstance(x, torch.Tensor):
        return x.to(device)
    if isinstance(x, dict):
        return {k: to_device(v, device) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(to_device(v, device) for v in x)
    return x


def make_depth_images(B, H=375, W=1242, seed=1234, lo=1.0, hi=80.0):
    """[B,H,W] f32 metric depth U(lo,hi) for the pseudo-LiDAR path (config C4)."""
    gen = torch.Generator().manual_seed(seed)
    return (lo + (hi - lo) * torch.rand(B, H, W, generator=gen)).contiguous()


def write_kitti_calib(calib_dir):
    """Write the two calibration text files `PseudoLiDAR.get_trans_proj` parses
    (`pseudo-lidar/utils/PseudoLiDAR.py:12-29,48-67`); returns the dir with a
    trailing slash, the form the reference concatenates file names onto."""
    import os
    os.makedirs(calib_dir, exist_ok=True)
    with open(os.path.join(calib_dir, "calib_velo_to_cam.txt"), "w") as f:
        f.write("calib_time: 15-Mar-2012 11:37:16\n")
        f.write("R: " + " ".join("%.6e" % v for v in KITTI_VELO_TO_CAM_R.reshape(-1)) + "\n")
        f.write("T: " + " ".join("%.6e" % v for v in KITTI_VELO_TO_CAM_T.reshape(-1)) + "\n")
    with open(os.path.join(calib_dir, "calib_cam_to_cam.txt"), "w") as f:
        f.write("calib_time: 09-Jan-2012 13:57:47\n")
        f.write("P_rect_02: " + " ".join("%.6e" % v for v in KITTI_P_RECT_02.reshape(-1)) + "\n")
        # key read by the Velodyne -> image class (`pseudo-lidar/Transform/Transform.py:66`)
        f.write("P: " + " ".join("%.6e" % v for v in KITTI_P_RECT_02.reshape(-1)) + "\n")
    d = str(calib_dir)
    return d if d.endswith("/") else d + "/"


def make_velodyne_cloud(n=123577, seed=1234):
    """[n,4] float32 x,y,z,reflectance shaped like one KITTI Velodyne sweep (64 beams, 360 degrees, ranges
    3..130 m - some beyond the 120 m cut, half of them behind the car), in azimuth-major order as the
    sensor writes them.  123 577 is the point count the reference quotes for a real frame
    (`pseudo-lidar/test_pipeline.py:72-73`)."""
    gen = np.random.default_rng(seed)
    az = np.sort(gen.uniform(-np.pi, np.pi, n))
    el = np.deg2rad(gen.uniform(-24.8, 2.0, n))
    rng = 3.0 + 127.0 * gen.beta(1.2, 4.0, n)
    x = rng * np.cos(el) * np.cos(az)
    y = rng * np.cos(el) * np.sin(az)
    z = rng * np.sin(el)
    return np.stack([x, y, z, gen.uniform(0, 1, n)], 1).astype(np.float32)
