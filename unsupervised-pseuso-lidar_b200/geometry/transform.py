"""Drop-in for the reference's `geometry/transform.py::Transform` (forward
only; the fused loss does not go through these)."""
import torch


def _ops():
    from plb200 import ops      # lazy: a CPU-only loader worker that star-imports geometry.pose_geometry needs no .so
    return ops


class Transform():

    def meshgrid(self, B, H, W, dtype, device, normalized=False):
        """`geometry/transform.py:14-45`."""
        if normalized:
            xs = torch.linspace(-1, 1, W, device=device, dtype=dtype)
            ys = torch.linspace(-1, 1, H, device=device, dtype=dtype)
        else:
            xs = torch.linspace(0, W - 1, W, device=device, dtype=dtype)
            ys = torch.linspace(0, H - 1, H, device=device, dtype=dtype)
        ys, xs = torch.meshgrid([ys, xs], indexing="ij")
        return xs.repeat([B, 1, 1]), ys.repeat([B, 1, 1])

    def image_grid(self, B, H, W, dtype, device, normalized=False):
        """`geometry/transform.py:47-72`."""
        xs, ys = self.meshgrid(B, H, W, dtype, device, normalized=normalized)
        return torch.stack([xs, ys, torch.ones_like(xs)], dim=1)

    def reconstruct(self, depth, K):
        """`geometry/transform.py:74-105`: depth [B,H,W] -> Xc [B,3,H,W]."""
        return _ops().reconstruct(depth, K)

    def k_hom(self, K):
        """`geometry/transform.py:107-112`, batch-agnostic."""
        Kh = torch.eye(4, device=K.device).reshape(1, 4, 4).repeat(K.shape[0], 1, 1)
        Kh[:, :3, :3] = K.clone()
        return Kh

    def project(self, X, K, Tcw):
        """`geometry/transform.py:114-150`: X [B,3,H,W], Tcw [B,4,4] -> grid [B,H,W,2]."""
        return _ops().project(X, K, Tcw)
