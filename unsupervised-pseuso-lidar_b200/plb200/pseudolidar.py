"""`PseudoLiDAR` of `pseudo-lidar/utils/PseudoLiDAR.py` (imported by the reference as `utils.PseudoLiDAR`; see
`plb200.dropin`): same constructor, same
`project_PL(depth_img) -> ndarray[N,4] float64`, computed on the GPU."""
import numpy as np
import torch

from . import ops


class PseudoLiDAR:

    def __init__(self, calib_dir, sparsity=0, device="cuda"):
        # the reference's ROS node forgets `sparsity` (PseudoLidarPipeline.py:25); default it
        self.T, self.P = self.get_trans_proj(calib_dir)
        self.sparsity = sparsity
        self.device = torch.device(device)

    def read_calib_file(self, filepath):
        """`PseudoLiDAR.py:12-29`: 'key: v v v' lines -> dict of float arrays."""
        data = {}
        with open(filepath, 'r') as f:
            for line in f:
                line = line.rstrip()
                if not line:
                    continue
                key, value = line.split(':', 1)
                try:
                    data[key] = np.array([float(x) for x in value.split()])
                except ValueError:
                    pass
        return data

    def cart2hom(self, pts_3d):
        return np.hstack((pts_3d, np.ones((pts_3d.shape[0], 1))))

    def inverse_rigid_trans(self, Tr):
        """`PseudoLiDAR.py:39-46` - host numpy on 16 numbers, kept identical
        (including the all-zero last row that zeros_like of a 4x4 leaves)."""
        inv_Tr = np.zeros_like(Tr)
        inv_Tr[0:3, 0:3] = np.transpose(Tr[0:3, 0:3])
        inv_Tr[0:3, 3] = np.dot(-np.transpose(Tr[0:3, 0:3]), Tr[0:3, 3])
        return inv_Tr

    def get_trans_proj(self, calib_dir):
        """`PseudoLiDAR.py:48-67`."""
        velo = self.read_calib_file(calib_dir + "calib_velo_to_cam.txt")
        cam = self.read_calib_file(calib_dir + "calib_cam_to_cam.txt")
        T = np.concatenate((velo["R"].reshape(3, 3), velo["T"].reshape(3, 1)), axis=1)
        T = np.vstack([T, [0, 0, 0, 1]])
        P = cam["P_rect_02"].reshape(3, 4)
        return T, P

    def project_batch(self, depth, **want):
        """[B,H,W] CUDA/CPU float depth -> device result dict (no sync); see ops.cloud_project."""
        if not isinstance(depth, torch.Tensor):
            depth = torch.from_numpy(np.ascontiguousarray(depth, dtype=np.float32))
        depth = depth.to(self.device, torch.float32)
        return ops.cloud_project(depth, self.P, self.inverse_rigid_trans(self.T), self.sparsity, **want)

    def project_PL(self, depth_img):
        """`PseudoLiDAR.py:69-110`: [H,W] depth -> [N,4] float64 ndarray."""
        depth = depth_img if isinstance(depth_img, torch.Tensor) else torch.from_numpy(
            np.ascontiguousarray(depth_img, dtype=np.float32))
        res = self.project_batch(depth.reshape(1, *depth.shape[-2:]))
        n = int(res["count"][0])                      # the output size is data dependent: one sync
        return res["cloud_f64"][0, :n].cpu().numpy()
