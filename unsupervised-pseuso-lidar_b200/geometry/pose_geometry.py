"""Drop-in for the reference's `geometry/pose_geometry.py`: the hot-path functions are libplb200.so kernels
with autograd; the host-side helpers the loader uses (`mat2euler`, `isRotationMatrix`, `invert_pose_np`,
`dataloaders.py:17,113`) are plain numpy / torch.

`plb200` (and with it `libplb200.so`) is imported lazily, on the first call of a kernel-backed function:
`from geometry.pose_geometry import *` in a CPU-only DataLoader worker (`dataloaders.py:17`) must not need
the CUDA library.  The module-level names match what the reference's star-import hands to its callers
(`torch`, `F`, `np`, `math`, `Transform` included).
"""
import math  # noqa: F401  (re-exported like the reference's module does)

import numpy as np
import torch
import torch.nn.functional as F  # noqa: F401

from .transform import Transform  # noqa: F401  (same import as geometry/pose_geometry.py:6)


def _plb():
    from plb200 import ops, _lib
    return ops, _lib


# ---- host-side helpers (numpy; used by the data loader, never on the GPU path) ------------------------------
def isRotationMatrix(R):
    """`geometry/pose_geometry.py:9-14`: ||I - R^T R||_F < 1e-6."""
    R = np.asarray(R)
    return np.linalg.norm(np.identity(3, dtype=R.dtype) - R.T @ R) < 1e-6


def mat2euler(R):
    """`geometry/pose_geometry.py:19-36`: rotation matrix -> (x, y, z) euler angles, MATLAB convention with
    x and z swapped; the gimbal-lock branch sets z = 0."""
    assert isRotationMatrix(R)
    sy = math.sqrt(R[0, 0] * R[0, 0] + R[1, 0] * R[1, 0])      # not hypot: same last bit as the reference
    y = math.atan2(-R[2, 0], sy)
    if sy >= 1e-6:
        return np.array([math.atan2(R[2, 1], R[2, 2]), y, math.atan2(R[1, 0], R[0, 0])])
    return np.array([math.atan2(-R[1, 2], R[1, 1]), y, 0])


def invert_pose_np(T):
    """`geometry/pose_geometry.py:117-122`: [4,4] ndarray pose -> its rigid inverse."""
    Tinv = np.array(T, copy=True)
    Rt = np.array(T[:3, :3]).T
    Tinv[:3, :3] = Rt
    Tinv[:3, 3] = -(Rt @ np.asarray(T[:3, 3]))
    return Tinv


# ---- torch helpers the reference exports (tiny [B,4,4] algebra; kept in torch, differentiable) ---------------
def get_translation_matrix(translation_vector):
    """`geometry/pose_geometry.py:141-153`: [B,*,3] -> [B,4,4] identity with t in the last column (fp32)."""
    t = translation_vector.contiguous().view(-1, 3, 1)
    T = torch.eye(4, device=t.device).repeat(t.shape[0], 1, 1)
    return torch.cat([torch.cat([T[:, :3, :3], t.to(T.dtype)], dim=2), T[:, 3:, :]], dim=1)


def rot_from_axisangle(vec):
    """`geometry/pose_geometry.py:155-199`: [B,1,3] axis-angle -> [B,4,4], Rodrigues with axis = v/(|v|+1e-7)."""
    ops, _lib = _plb()
    B = vec.shape[0]
    pose = torch.cat([vec.reshape(B, 3), torch.zeros(B, 3, dtype=vec.dtype, device=vec.device)], dim=1)
    return ops.PoseMatrixFn.apply(pose, _lib.ROT_AXISANGLE, False)


def invert_pose(T):
    """`geometry/pose_geometry.py:110-115`: tiny [B,4,4] op, kept in torch."""
    Rt = T[:, :3, :3].transpose(-2, -1)
    tinv = torch.bmm(-1.0 * Rt, T[:, :3, 3:4])
    bottom = torch.zeros(len(T), 1, 4, device=T.device, dtype=T.dtype)
    bottom[:, :, 3] = 1
    return torch.cat([torch.cat([Rt, tinv], dim=2), bottom], dim=1)


# ---- kernel-backed hot-path functions ------------------------------------------------------------------------
def disp_to_depth(disps):
    """`geometry/pose_geometry.py:70-95`: D = 1/(10*d + 0.01), nested lists kept."""
    ops, _ = _plb()
    return [[ops.DispToDepthFn.apply(d, 10.0, 0.01) for d in frame] for frame in disps]


def euler2mat(angle):
    """`geometry/pose_geometry.py:38-68`: [B,3] -> [B,3,3], R = Rx Ry Rz."""
    ops, _lib = _plb()
    pose = torch.cat([angle, torch.zeros_like(angle)], dim=1)
    return ops.PoseMatrixFn.apply(pose, _lib.ROT_EULER, False)[:, :3, :3]


def pose_vec2mat(vec, mode='euler'):
    """`geometry/pose_geometry.py:97-108`: [B,6] -> [B,3,4] float."""
    if mode is None:
        return vec
    if mode != 'euler':
        raise ValueError('Rotation mode not supported {}'.format(mode))
    ops, _lib = _plb()
    return ops.PoseMatrixFn.apply(vec, _lib.ROT_EULER, False)[:, :3, :]


def transformation_from_parameters(axisangle, translation, invert=False):
    """`geometry/pose_geometry.py:124-136`: ([B,1,3], [B,1,3]) -> [B,4,4] = T @ R.
    The reference's `invert=True` branch (R^T @ T(-t)) equals the rigid inverse."""
    ops, _lib = _plb()
    B = axisangle.shape[0]
    pose = torch.cat([axisangle.reshape(B, 3), translation.reshape(B, 3)], dim=1)
    return ops.PoseMatrixFn.apply(pose, _lib.ROT_AXISANGLE, bool(invert))


def inverse_warp(img, depth, pose, K, pose_inv, rotation_mode='euler', padding_mode='zeros', *, euler_pose=False):
    """`geometry/pose_geometry.py:201-228`, same signature.  img [B,3,H,W], depth [B,H,W] (or [B,1,H,W]),
    pose [B,6], K [B,3,3] -> warped [B,3,H,W].  Like the reference, `rotation_mode` is accepted and IGNORED: the
    pose is always read as axis-angle through `transformation_from_parameters` (`:219-220`).  The keyword-only
    `euler_pose=True` (not in the reference) selects the dormant `pose_vec2mat` reading instead."""
    if padding_mode != 'zeros':
        raise ValueError("only padding_mode='zeros' (the reference's) is implemented")
    ops, _lib = _plb()
    if depth.dim() == 4:
        depth = depth[:, 0]
    rot = _lib.ROT_EULER if euler_pose else _lib.ROT_AXISANGLE
    return ops.InverseWarpFn.apply(img, depth, pose, K, bool(pose_inv), rot)
