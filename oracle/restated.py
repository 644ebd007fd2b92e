"""ORACLE - TEST INFRASTRUCTURE ONLY.  Never imported by the product path.

CPU restatement (torch, fp32 like the reference; fp64 numpy for the point
cloud) of the reference's view-synthesis hot path.  Every function cites the
reference file:line it follows.  It is batch-size agnostic (the reference only
runs at B=4 on CUDA, `geometry/transform.py:110,134`), device agnostic (every
tensor it creates follows its inputs: on CUDA tensors it IS the reference's stock
torch-eager op sequence, which `bench.py` times as `gpu_eager_baseline`) and differentiable
through torch autograd, so it gives the gradients the CUDA backward is checked
against.

Pinning: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4), so this restatement is pinned against outputs of the
UNMODIFIED reference executed in the build container:
`tests/golden/make_golden.py` imports `/root/reference` through
`oracle/reference_shim.py` and commits its outputs (loss values, gradients,
warped images, point clouds) under `tests/golden/`;
`tests/test_oracle_golden.py` checks this file against them.  The third-party
arithmetic (`F.grid_sample`, `F.interpolate`, `AvgPool2d`, `ReflectionPad2d`,
`torch.linalg.inv`) lives in PyTorch, which the reference does not pin
(`utils/requirements.txt:1-2`); torch 2.11.0 (installed here and on the GPU
box) is the semantic definition and is called directly.

Rows that the reference does not contain at all (edge-aware smoothness,
SURVEY.md section 8 a17) are marked "parity unpinned" where they are defined.
"""
import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------
# geometry/pose_geometry.py
# --------------------------------------------------------------------------

def disp_to_depth(disps):
    """`geometry/pose_geometry.py:70-95`: D = 1 / (10*d + 0.01), nested lists kept."""
    return [[1 / (10 * d + 0.01) for d in frame] for frame in disps]


def rot_from_axisangle(vec):
    """`geometry/pose_geometry.py:155-199`: Rodrigues with axis = v/(|v|+1e-7).
    vec [B,1,3] -> [B,4,4]."""
    angle = torch.norm(vec, 2, 2, True)
    axis = vec / (angle + 1e-7)
    ca, sa = torch.cos(angle), torch.sin(angle)
    C = 1 - ca
    x, y, z = axis[..., 0:1], axis[..., 1:2], axis[..., 2:3]
    xs, ys, zs = x * sa, y * sa, z * sa
    xC, yC, zC = x * C, y * C, z * C
    xyC, yzC, zxC = x * yC, y * zC, z * xC
    zero = torch.zeros_like(ca)
    one = torch.ones_like(ca)
    rows = [x * xC + ca, xyC - zs, zxC + ys, zero,
            xyC + zs, y * yC + ca, yzC - xs, zero,
            zxC - ys, yzC + xs, z * zC + ca, zero,
            zero, zero, zero, one]
    return torch.cat(rows, dim=2).view(-1, 4, 4)


def get_translation_matrix(t):
    """`geometry/pose_geometry.py:138-153`: [B,*,3] -> [B,4,4] identity with t in column 3."""
    B = t.shape[0]
    T = torch.eye(4, dtype=t.dtype, device=t.device).repeat(B, 1, 1)
    T = T.clone()
    T[:, :3, 3] = t.reshape(B, 3)
    return T


def transformation_from_parameters(axisangle, translation, invert=False):
    """`geometry/pose_geometry.py:124-136`: M = T @ R (or R^T @ T(-t) when invert)."""
    R = rot_from_axisangle(axisangle)
    t = translation.clone()
    if invert:
        R = R.transpose(1, 2)
        t = t * -1
    T = get_translation_matrix(t)
    return torch.matmul(R, T) if invert else torch.matmul(T, R)


def invert_pose(T):
    """`geometry/pose_geometry.py:110-115`: [R|t] -> [R^T | -R^T t]."""
    Rt = T[:, :3, :3].transpose(-2, -1)
    tinv = torch.bmm(-1.0 * Rt, T[:, :3, 3:4])
    top = torch.cat([Rt, tinv], dim=2)
    bottom = torch.tensor([0, 0, 0, 1], dtype=T.dtype, device=T.device).view(1, 1, 4).repeat(len(T), 1, 1)
    return torch.cat([top, bottom], dim=1)


def euler2mat(angle):
    """`geometry/pose_geometry.py:38-68`: R = Rx @ Ry @ Rz."""
    x, y, z = angle[:, 0], angle[:, 1], angle[:, 2]
    zeros = z.detach() * 0
    ones = zeros + 1
    cz, sz = torch.cos(z), torch.sin(z)
    cy, sy = torch.cos(y), torch.sin(y)
    cx, sx = torch.cos(x), torch.sin(x)
    zmat = torch.stack([cz, -sz, zeros, sz, cz, zeros, zeros, zeros, ones], 1).view(-1, 3, 3)
    ymat = torch.stack([cy, zeros, sy, zeros, ones, zeros, -sy, zeros, cy], 1).view(-1, 3, 3)
    xmat = torch.stack([ones, zeros, zeros, zeros, cx, -sx, zeros, sx, cx], 1).view(-1, 3, 3)
    return xmat.bmm(ymat).bmm(zmat)


def pose_vec2mat(vec, mode="euler"):
    """`geometry/pose_geometry.py:97-108`: [B,6] (rot|trans) -> [B,3,4] float."""
    if mode is None:
        return vec
    if mode != "euler":
        raise ValueError("Rotation mode not supported {}".format(mode))
    M = torch.cat([euler2mat(vec[:, :3]), vec[:, 3:].unsqueeze(-1)], dim=2)
    # `.float()` in the reference; fp64 input (the accuracy study in the tests) stays fp64
    return M if vec.dtype == torch.float64 else M.float()


# --------------------------------------------------------------------------
# geometry/transform.py
# --------------------------------------------------------------------------

def image_grid(B, H, W, dtype, device=None):
    """`geometry/transform.py:14-72`: [B,3,H,W] of (x=0..W-1, y=0..H-1, 1)."""
    xs = torch.linspace(0, W - 1, W, dtype=dtype, device=device)
    ys = torch.linspace(0, H - 1, H, dtype=dtype, device=device)
    ys, xs = torch.meshgrid([ys, xs], indexing="ij")
    xs, ys = xs.repeat([B, 1, 1]), ys.repeat([B, 1, 1])
    return torch.stack([xs, ys, torch.ones_like(xs)], dim=1)


def reconstruct(depth, K):
    """`geometry/transform.py:74-105`: Xc = (K^-1 . grid) * depth; depth [B,H,W]."""
    depth = depth.unsqueeze(1)
    B, _, H, W = depth.shape
    Kinv = K.inverse().to(depth.dtype)   # `.float()` in the reference (depth is fp32 there)
    grid = image_grid(B, H, W, depth.dtype, depth.device).view(B, 3, -1)
    return Kinv.bmm(grid).view(B, 3, H, W) * depth


def k_hom(K, dtype=torch.float32):
    """`geometry/transform.py:107-112` with the hard-coded batch 4 replaced by K's
    (fp32 in the reference; `dtype` exists for the fp64 accuracy study in the tests)."""
    Kh = torch.eye(4, dtype=dtype, device=K.device).reshape(1, 4, 4).repeat(K.shape[0], 1, 1)
    Kh[:, :3, :3] = K.clone()
    return Kh


def project(X, K, Tcw):
    """`geometry/transform.py:114-150`: pixel grid in [-1,1] for grid_sample."""
    B, _, H, W = X.shape
    Xc = X.view(B, 3, -1)
    ones = torch.ones(1, Xc.shape[-1], dtype=X.dtype, device=X.device).repeat(B, 1, 1)
    Xh = torch.cat([Xc, ones], 1)
    Tx = (k_hom(K, X.dtype) @ Tcw)[:, :3, :]
    cam = Tx @ Xh
    pix = cam[:, :2, :] / (cam[:, 2, :].unsqueeze(1) + 1e-5)
    pix = pix.view(B, 2, H, W).permute(0, 2, 3, 1)
    px = pix[..., 0] / (W - 1)
    py = pix[..., 1] / (H - 1)
    return (torch.stack([px, py], dim=-1) - 0.5) * 2


def pose_matrix(pose, pose_inv, rotation_mode="axisangle"):
    """The 4x4 (or [R|t;0 0 0 1]) target->source transform `inverse_warp` builds
    (`geometry/pose_geometry.py:218-223`); 'euler' is the dormant variant
    (`notes/toy_problem/geometry/pose_geometry.py:126`)."""
    if rotation_mode == "euler":
        M34 = pose_vec2mat(pose, "euler")
        bottom = torch.tensor([0, 0, 0, 1], dtype=M34.dtype, device=M34.device).view(1, 1, 4).repeat(len(M34), 1, 1)
        Tcw = torch.cat([M34, bottom], dim=1)
    else:
        trans, rot = pose[:, 3:].unsqueeze(1), pose[:, :3].unsqueeze(1)
        Tcw = transformation_from_parameters(rot, trans)
    if pose_inv:
        Tcw = invert_pose(Tcw)
    return Tcw


def inverse_warp(img, depth, pose, K, pose_inv, rotation_mode="axisangle", padding_mode="zeros"):
    """`geometry/pose_geometry.py:201-228`.  depth [B,H,W] (already squeezed)."""
    if depth.dim() == 4:
        depth = depth[:, 0]
    Xc = reconstruct(depth, K)
    Tcw = pose_matrix(pose, pose_inv, rotation_mode)
    grid = project(Xc, K, Tcw)
    return F.grid_sample(img, grid, mode="bilinear", padding_mode=padding_mode, align_corners=True)


# --------------------------------------------------------------------------
# losses.py (live)
# --------------------------------------------------------------------------

def upsample_depth(D, H, W):
    """`losses.py:214-215`."""
    if D.shape[-1] != W:
        D = F.interpolate(D, [H, W], mode="bilinear", align_corners=False)
    return D


def reprojection_loss(tgt, refs, depths, poses, K, rotation_mode="axisangle"):
    """`losses.py:183-240`, mode='min' (which is a mean, `:226-228`), including
    the direction-1 quirks."""
    pose_list = [poses[:, i, :] for i in range(poses.shape[1])]
    loss = []
    for indx in range(len(depths)):
        depth = depths[indx]
        if indx == 0:
            ref_imgs, tgt_img, pose_inv = refs, tgt, False
        else:
            ref_imgs, tgt_img, pose_inv = [tgt], refs[indx], True
            pose_list = [pose_list[indx - 1]]
        H, W = depth[0].shape[-2:]
        for D in depth:
            D = upsample_depth(D, H, W)[:, 0]
            terms = []
            for ref_img, pose in zip(ref_imgs, pose_list):
                proj = inverse_warp(ref_img, D, pose, K, pose_inv, rotation_mode)
                terms.append(F.l1_loss(proj, tgt_img))
            loss.append(torch.mean(torch.stack(terms)))
    return sum(loss) / len(loss)


def smooth_loss(pred_map):
    """`losses.py:242-260`: second-order, not edge-aware, weight /= 2.3 per scale."""
    def gradient(p):
        return p[:, :, :, 1:] - p[:, :, :, :-1], p[:, :, 1:] - p[:, :, :-1]
    if type(pred_map) not in (tuple, list):
        pred_map = [pred_map]
    loss, weight = 0, 1.0
    for m in pred_map:
        dx, dy = gradient(m)
        dx2, dxdy = gradient(dx)
        dydx, dy2 = gradient(dy)
        loss = loss + (dx2.abs().mean() + dxdy.abs().mean() + dydx.abs().mean() + dy2.abs().mean()) * weight
        weight /= 2.3
    return loss


def losses_forward(tgt, ref_imgs, disparity, poses, K, gt=None):
    """`losses.py:262-271`: [loss_mam, loss_smooth]."""
    depths = disp_to_depth(disparity)
    return [reprojection_loss(tgt, ref_imgs, depths, poses, K), smooth_loss(depths[0])]


# --------------------------------------------------------------------------
# losses.py (dormant): SSIM, photometric mix + clip, min-reprojection, automask
# --------------------------------------------------------------------------

def ssim_standard_loss(x, y, C1=1e-4, C2=9e-4):
    """`losses.py:12-54`: ReflectionPad2d(1) + AvgPool2d(3,1); clamp((1-ssim)/2, 0, 1)."""
    x, y = F.pad(x, (1, 1, 1, 1), mode="reflect"), F.pad(y, (1, 1, 1, 1), mode="reflect")
    mu_x, mu_y = F.avg_pool2d(x, 3, 1), F.avg_pool2d(y, 3, 1)
    mu_xy, mu_xx, mu_yy = mu_x * mu_y, mu_x.pow(2), mu_y.pow(2)
    s_x = F.avg_pool2d(x.pow(2), 3, 1) - mu_xx
    s_y = F.avg_pool2d(y.pow(2), 3, 1) - mu_yy
    s_xy = F.avg_pool2d(x * y, 3, 1) - mu_xy
    n = (2 * mu_xy + C1) * (2 * s_xy + C2)
    d = (mu_xx + mu_yy + C1) * (s_x + s_y + C2)
    return torch.clamp((1.0 - n / d) / 2.0, 0.0, 1.0)


def compute_photometric_loss(pred, target, no_ssim=False, clip_loss=0.5):
    """`losses.py:66-84`: 0.85*ssim + 0.15*|t-p| per channel; clamp at mean+clip*std
    with the threshold detached through float().  clip_loss=None skips the clamp."""
    l1 = torch.abs(target - pred)
    photo = l1 if no_ssim else 0.85 * ssim_standard_loss(pred, target) + 0.15 * l1
    if clip_loss is not None:
        mean, std = photo.mean(), photo.std()
        photo = torch.clamp(photo, max=float(mean + clip_loss * std))
    return photo


def min_reprojection_loss(tgt, refs, depth_scales, poses, K, automask=True, no_ssim=False,
                          clip_loss=None, rotation_mode="axisangle", pose_inv=False):
    """The dormant composition: per scale, warp every source with the upsampled
    depth (`losses.py:209-219`), photometric map per source (`:66-84`), per-pixel
    minimum over sources and the binary automask against the un-warped sources
    (`notes/toy_problem/losses.py:107-124`, commented copy at `losses.py:154-162`),
    `reduce_loss` = max over channel then mean (`losses.py:94-96`); scales are
    averaged (`losses.py:181`)."""
    H, W = depth_scales[0].shape[-2:]
    auto = [compute_photometric_loss(r, tgt, no_ssim, clip_loss) for r in refs] if automask else None
    total = 0
    for D in depth_scales:
        D = upsample_depth(D, H, W)[:, 0]
        rp = [compute_photometric_loss(
            inverse_warp(r, D, poses[:, i, :], K, pose_inv, rotation_mode), tgt, no_ssim, clip_loss)
            for i, r in enumerate(refs)]
        m = rp[0]
        for r in rp[1:]:
            m = torch.minimum(m, r)
        if automask:
            a = auto[0]
            for r in auto[1:]:
                a = torch.minimum(a, r)
            mu = (m < a).to(m.dtype)
            m = mu * m
        per_px, _ = torch.max(m, dim=1)
        total = total + per_px.mean(1).mean(-1).mean()
    return total / len(depth_scales)


def edge_aware_smoothness(disp, img, normalize=True):
    """ABSENT from the reference - PARITY UNPINNED.
    Formula of the monodepth2 lineage the reference's model files cite
    (`models/depth/layers.py:1-2`): mean|dx d|*exp(-mean_c|dx I|) + same in y,
    d optionally divided by its per-image mean (+1e-7).  `img` must have the
    resolution of `disp`."""
    if normalize:
        disp = disp / (disp.mean(2, True).mean(3, True) + 1e-7)
    gdx = torch.abs(disp[:, :, :, :-1] - disp[:, :, :, 1:])
    gdy = torch.abs(disp[:, :, :-1, :] - disp[:, :, 1:, :])
    gix = torch.mean(torch.abs(img[:, :, :, :-1] - img[:, :, :, 1:]), 1, keepdim=True)
    giy = torch.mean(torch.abs(img[:, :, :-1, :] - img[:, :, 1:, :]), 1, keepdim=True)
    return (gdx * torch.exp(-gix)).mean() + (gdy * torch.exp(-giy)).mean()


def edge_aware_smooth_loss(disp_scales, tgt, normalize=True):
    """PARITY UNPINNED (see `edge_aware_smoothness`).  Scale s uses the target
    image average-pooled by 2^s and weight 1/2^s (monodepth2's schedule)."""
    loss = 0
    H = tgt.shape[-2]
    for d in disp_scales:
        f = H // d.shape[-2]
        img = tgt if f == 1 else F.avg_pool2d(tgt, f, f)
        loss = loss + edge_aware_smoothness(d, img, normalize) / f
    return loss


# --------------------------------------------------------------------------
# pseudo-lidar/utils/PseudoLiDAR.py
# --------------------------------------------------------------------------

def inverse_rigid_trans(Tr):
    """`pseudo-lidar/utils/PseudoLiDAR.py:39-46`.  zeros_like of the 4x4 input
    leaves the last row all zero, so the 4th output column of the cloud is 0."""
    inv = np.zeros_like(Tr)
    inv[0:3, 0:3] = np.transpose(Tr[0:3, 0:3])
    inv[0:3, 3] = np.dot(-np.transpose(Tr[0:3, 0:3]), Tr[0:3, 3])
    return inv


def project_PL(depth_img, T, P, sparsity=0, return_valid=False):
    """`pseudo-lidar/utils/PseudoLiDAR.py:69-110`, fp64, same operation order.
    T 4x4 velodyne->camera, P 3x4 P_rect_02.  Returns [N,4] f64 (and the
    row-major validity mask when asked)."""
    rows, cols = depth_img.shape
    c, r = np.meshgrid(np.arange(cols), np.arange(rows))
    uvd = np.stack([c, r, depth_img]).reshape((3, -1)).T
    c_u, c_v, f_u, f_v = P[0, 2], P[1, 2], P[0, 0], P[1, 1]
    b_x, b_y = P[0, 3] / (-f_u), P[1, 3] / (-f_v)
    n = uvd.shape[0]
    pts = np.ones((n, 4))
    pts[:, 0] = ((uvd[:, 0] - c_u) * uvd[:, 2]) / f_u + b_x
    pts[:, 1] = ((uvd[:, 1] - c_v) * uvd[:, 2]) / f_v + b_y
    pts[:, 2] = uvd[:, 2]
    cloud = np.matmul(pts, np.transpose(inverse_rigid_trans(T)))
    valid = (cloud[:, 0] >= 0) & (cloud[:, 2] < 1)
    out = cloud[valid]
    if sparsity:
        out = out[0::sparsity]
    return (out, valid) if return_valid else out


def velo_to_cam_matrix(R, t):
    """`pseudo-lidar/utils/PseudoLiDAR.py:48-58`: [[R|t],[0 0 0 1]]."""
    T = np.concatenate((np.asarray(R, dtype=np.float64).reshape(3, 3),
                        np.asarray(t, dtype=np.float64).reshape(3, 1)), axis=1)
    return np.vstack([T, [0, 0, 0, 1]])


# --------------------------------------------------------------------------
# pseudo-lidar/Transform/Transform.py  (the inverse of project_PL)
# --------------------------------------------------------------------------

def _matvec4_like_numpy(M, v):
    """[N,4] vectors through a [R,4] matrix with the rounding order of the per-point `np.matmul(M, pnt)`
    the reference runs (a 4-wide SIMD product followed by the horizontal add (p0 + p2) + (p1 + p3));
    found by matching candidates against the unmodified reference's output bit for bit, and held to it
    by tests/golden/velo_kitti.npz."""
    p = v[:, None, :] * M[None, :, :]                       # [N,R,4] rounded products
    return (p[..., 0] + p[..., 2]) + (p[..., 1] + p[..., 3])


def project_velo_to_img(point_cloud, T, P, width, height):
    """`pseudo-lidar/Transform/Transform.py:69-104`: Velodyne points [N,>=3] (float32 as read from a KITTI
    .bin) -> depth image [height, width] float64.  Per point, in order: dist = sqrt(x^2+y^2+z^2) in the
    cloud's dtype; xyz = T @ [x,y,z,1] and uv = P @ xyz in fp64 (vstack with the int 1 promotes to
    float64); uv /= uv[2]; kept when 0 <= u < width, 0 <= v < height, dist <= 120, x > 0; the cell
    (int(u), int(v)) takes xyz[2] and LATER POINTS OVERWRITE EARLIER ONES.  Vectorised: the winner of a
    cell is the kept point with the largest index.  Returns (depth [height,width], winner index map)."""
    pc = np.asarray(point_cloud)[:, :3]
    x, y, z = pc[:, 0], pc[:, 1], pc[:, 2]
    dist = np.sqrt(x ** 2 + y ** 2 + z ** 2)
    hom = np.concatenate([pc.astype(np.float64), np.ones((pc.shape[0], 1))], axis=1)
    xyz = _matvec4_like_numpy(np.asarray(T, dtype=np.float64), hom)          # [N,4]
    uvw = _matvec4_like_numpy(np.asarray(P, dtype=np.float64), xyz)          # [N,3]
    with np.errstate(divide="ignore", invalid="ignore"):
        u, v = uvw[:, 0] / uvw[:, 2], uvw[:, 1] / uvw[:, 2]
        keep = (u >= 0) & (u < width) & (v >= 0) & (v < height) & (dist <= 120) & (x > 0)
    idx = np.nonzero(keep)[0]
    cell = u[idx].astype(np.int64) * height + v[idx].astype(np.int64)      # depth_array[int(u)][int(v)]
    winner = np.full(width * height, -1, dtype=np.int64)
    np.maximum.at(winner, cell, idx)
    depth = np.zeros(width * height, dtype=np.float64)
    has = winner >= 0
    depth[has] = xyz[winner[has], 2]
    return np.transpose(depth.reshape(width, height)), winner.reshape(width, height).T


# --------------------------------------------------------------------------
# models/depth/disp_net.py:121-139  (the disparity head in front of the loss)
# --------------------------------------------------------------------------

def disp_head(x, alpha=10.0, beta=0.01):
    """`models/depth/disp_net.py:121,127,133,139`: `disp = self.alpha * predict_disp(out) + self.beta`, where
    `predict_disp` ends in `nn.Sigmoid()` (`:24-28`) and `alpha = 10`, `beta = 0.01` (`:53-57`).  `x` is the
    convolution output in front of the sigmoid."""
    return alpha * torch.sigmoid(x) + beta


# --------------------------------------------------------------------------
# dataloaders.py:32-49 + trainer.py:97-103  (the loader's transform chain)
# --------------------------------------------------------------------------

def _pil_bilinear_coeffs(in_size, out_size):
    """Pillow's `precompute_coeffs` + `normalize_coeffs_8bpc` (src/libImaging/Resample.c; Pillow 12.2 installed here,
    unpinned by the reference) for the bilinear filter: per output sample the first input sample, the tap count and
    the taps in 22-bit fixed point.  Double arithmetic in Pillow's operation order; the support of the triangle is
    scaled by the reduction factor (antialiasing)."""
    scale = float(np.float32(in_size) - np.float32(0)) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int64)
    kk = np.zeros((out_size, ksize), np.int64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        k = np.zeros(ksize)
        ww = 0.0
        for x in range(xmax):
            a = (x + xmin - center + 0.5) * ss
            a = -a if a < 0 else a
            k[x] = 1.0 - a if a < 1.0 else 0.0
            ww += k[x]
        if ww != 0.0:
            k[:xmax] /= ww
        bounds[xx] = (xmin, xmax)
        for x in range(ksize):
            kk[xx, x] = int(-0.5 + k[x] * (1 << 22)) if k[x] < 0 else int(0.5 + k[x] * (1 << 22))
    return bounds, kk


def pil_resize_bilinear(img, H, W):
    """`Image.resize((W, H), BILINEAR)` on an RGB uint8 array [h, w, 3] - what `transforms.Resize((H, W))` runs on a
    PIL image (trainer.py:100): horizontal pass, then vertical pass, each only when that size changes, each rounded
    to uint8 through `((1 << 21) + sum) >> 22` and clipped (`ImagingResampleHorizontal_8bpc` / `Vertical_8bpc`)."""
    h0, w0, _ = img.shape
    out = img.astype(np.int64)
    if W != w0:
        b, kk = _pil_bilinear_coeffs(w0, W)
        tmp = np.zeros((h0, W, 3), np.int64)
        for xx in range(W):
            xmin, n = b[xx]
            acc = (1 << 21) + (out[:, xmin:xmin + n, :] * kk[xx, :n][None, :, None]).sum(1)
            tmp[:, xx, :] = np.clip(acc >> 22, 0, 255)
        out = tmp
    if H != h0:
        b, kk = _pil_bilinear_coeffs(h0, H)
        tmp = np.zeros((H, out.shape[1], 3), np.int64)
        for yy in range(H):
            ymin, n = b[yy]
            acc = (1 << 21) + (out[ymin:ymin + n] * kk[yy, :n][:, None, None]).sum(0)
            tmp[yy] = np.clip(acc >> 22, 0, 255)
        out = tmp
    return out.astype(np.uint8)


def load_img_chain(frame_u8, H, W, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
    """`KittiDataset.load_img` (dataloaders.py:32-49) with the transform list of trainer.py:97-103 on a decoded
    uint8 RGB frame [h, w, 3] -> normalised float32 [3, H, W]:
      `np.asarray(img, float32) / 255.0`  (dataloaders.py:33-38)
      ToTensor (float input: transpose only) -> ToPILImage (`mul(255).byte()`: the float32 round trip TRUNCATES)
      -> Resize((H, W)) on the PIL image -> ToTensor (`/ 255`) -> Normalize (`sub_(mean).div_(std)`)."""
    f = frame_u8.astype(np.float32) / np.float32(255.0)
    u8 = (f * np.float32(255.0)).astype(np.uint8)                      # mul(255).byte(): truncation toward zero
    r = pil_resize_bilinear(u8, H, W)
    t = np.transpose(r, (2, 0, 1)).astype(np.float32) / np.float32(255.0)
    m = np.asarray(mean, dtype=np.float32).reshape(3, 1, 1)
    s = np.asarray(std, dtype=np.float32).reshape(3, 1, 1)
    return (t - m) / s


def scale_intrinsics(K, in_h, in_w, H, W):
    """dataloaders.py:95-98: `intrinsics[0] *= W / og_w; intrinsics[1] *= H / og_h` (float64)."""
    K = np.array(K, dtype=np.float64, copy=True)
    K[..., 0, :] *= W / in_w
    K[..., 1, :] *= H / in_h
    return K
