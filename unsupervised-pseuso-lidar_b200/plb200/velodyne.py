"""Drop-in for `pseudo-lidar/Transform/Transform.py` (imported by the reference as
`from Transform.Transform import Transform`, pseudo-lidar/test_pipeline.py:13): same constructor,
same `project_velo_to_img(point_cloud) -> ndarray[height, width] float64`, computed on the GPU."""
import numpy as np
import torch

from . import ops


class Transform:

    def __init__(self, calib_dir, img_width, img_height, device="cuda"):
        self.CALIB_DIR = calib_dir
        self.T, self.P = self.get_trans_proj()
        self.width = img_width
        self.height = img_height
        self.device = torch.device(device)

    def read_calib_file(self, filepath):
        """`Transform.py:20-36`: 'key: v v v' lines -> dict of float arrays (dates are skipped)."""
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

    def get_trans_proj(self):
        """`Transform.py:47-67`: T = [[R|t],[0 0 0 1]] from calib_velo_to_cam.txt, P from key 'P'."""
        velo = self.read_calib_file(self.CALIB_DIR + "calib_velo_to_cam.txt")
        cam = self.read_calib_file(self.CALIB_DIR + "calib_cam_to_cam.txt")
        T = np.concatenate((velo["R"].reshape(3, 3), velo["T"].reshape(3, 1)), axis=1)
        T = np.vstack([T, [0, 0, 0, 1]])
        return T, cam["P"].reshape(3, 4)

    def project_batch(self, clouds, counts=None, **want):
        """[B,N,C>=3] float clouds (tensor or ndarray) -> device result dict (no sync); see ops.velo_project."""
        if not isinstance(clouds, torch.Tensor):
            clouds = torch.from_numpy(np.ascontiguousarray(clouds, dtype=np.float32))
        clouds = clouds.to(self.device, torch.float32)
        return ops.velo_project(clouds, self.T, self.P, self.height, self.width, counts=counts, **want)

    def project_velo_to_img(self, point_cloud):
        """`Transform.py:69-104`: [N,>=3] points -> [height, width] float64 ndarray."""
        pc = point_cloud if isinstance(point_cloud, torch.Tensor) else torch.from_numpy(
            np.ascontiguousarray(point_cloud, dtype=np.float32))
        res = self.project_batch(pc.reshape(1, *pc.shape[-2:]))
        return res["depth_f64"][0].cpu().numpy()
