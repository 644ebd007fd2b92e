"""Drop-in for the reference's `geometry/pose_geometry.py` (hot-path functions
only), each a libplb200.so kernel with autograd."""
import torch

from plb200 import ops, _lib
from .transform import Transform  # noqa: F401  (same import as geometry/pose_geometry.py:6)


def disp_to_depth(disps):
    """`geometry/pose_geometry.py:70-95`: D = 1/(10*d + 0.01), nested lists kept."""
    return [[ops.DispToDepthFn.apply(d, 10.0, 0.01) for d in frame] for frame in disps]


def euler2mat(angle):
    """`geometry/pose_geometry.py:38-68`: [B,3] -> [B,3,3], R = Rx Ry Rz."""
    pose = torch.cat([angle, torch.zeros_like(angle)], dim=1)
    return ops.PoseMatrixFn.apply(pose, _lib.ROT_EULER, False)[:, :3, :3]


def pose_vec2mat(vec, mode='euler'):
    """`geometry/pose_geometry.py:97-108`: [B,6] -> [B,3,4] float."""
    if mode is None:
        return vec
    if mode != 'euler':
        raise ValueError('Rotation mode not supported {}'.format(mode))
    return ops.PoseMatrixFn.apply(vec, _lib.ROT_EULER, False)[:, :3, :]


def transformation_from_parameters(axisangle, translation, invert=False):
    """`geometry/pose_geometry.py:124-136`: ([B,1,3], [B,1,3]) -> [B,4,4] = T @ R.
    The reference's `invert=True` branch (R^T @ T(-t)) equals the rigid inverse."""
    B = axisangle.shape[0]
    pose = torch.cat([axisangle.reshape(B, 3), translation.reshape(B, 3)], dim=1)
    return ops.PoseMatrixFn.apply(pose, _lib.ROT_AXISANGLE, bool(invert))


def invert_pose(T):
    """`geometry/pose_geometry.py:110-115`: tiny [B,4,4] op, kept in torch."""
    Rt = T[:, :3, :3].transpose(-2, -1)
    tinv = torch.bmm(-1.0 * Rt, T[:, :3, 3:4])
    bottom = torch.zeros(len(T), 1, 4, device=T.device, dtype=T.dtype)
    bottom[:, :, 3] = 1
    return torch.cat([torch.cat([Rt, tinv], dim=2), bottom], dim=1)


def inverse_warp(img, depth, pose, K, pose_inv, rotation_mode='axisangle', padding_mode='zeros'):
    """`geometry/pose_geometry.py:201-228`.  img [B,3,H,W], depth [B,H,W] (or
    [B,1,H,W]), pose [B,6], K [B,3,3] -> warped [B,3,H,W].  The reference ignores
    its `rotation_mode` argument and always uses axis-angle (`:219-220`);
    'euler' selects the dormant pose_vec2mat variant."""
    if padding_mode != 'zeros':
        raise ValueError("only padding_mode='zeros' (the reference's) is implemented")
    if depth.dim() == 4:
        depth = depth[:, 0]
    rot = _lib.ROT_EULER if rotation_mode == 'euler_dormant' else _lib.ROT_AXISANGLE
    return ops.InverseWarpFn.apply(img, depth, pose, K, bool(pose_inv), rot)
