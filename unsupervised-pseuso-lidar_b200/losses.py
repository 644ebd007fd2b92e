"""Drop-in for the reference's `losses.py` (same class and method names, same
argument meaning) backed by the fused sm_100a kernels of libplb200.so.

    from losses import Losses            # trainer.py:24
    criterion = Losses()                 # trainer.py:79
    loss = criterion.forward(tgt, ref_imgs, disps, poses, intrinsics, gt)   # trainer.py:312
    sum(loss).backward()                 # trainer.py:264

Differences from the reference, all deliberate:
  * batch-size agnostic (the reference only runs at B=4, geometry/transform.py:110);
  * nothing is printed and nothing synchronises (the reference prints inside the
    loss, losses.py:191, which forces a device->host copy per step);
  * the dormant SSIM / min-reprojection / automask path (losses.py:12-84,
    94-96,154-162) is callable: `multiview_reprojection_loss`.
"""
import torch

from plb200 import ops, _lib
from geometry.pose_geometry import inverse_warp, disp_to_depth  # noqa: F401  (same import as losses.py:8)


class SSIM:
    """`losses.py:11-54`."""

    def standard_loss(self, x, y, C1=1e-4, C2=9e-4, kernel_size=3, stride=1):
        if kernel_size != 3 or stride != 1:
            raise ValueError("only the 3x3 / stride-1 SSIM the reference uses is implemented")
        return ops.ssim_map(x, y, C1, C2)


class Losses:
    """`losses.py:56-271`.  Constructor takes no arguments in the reference
    (trainer.py:79); the keyword options only select implementation variants."""

    def __init__(self, rotation_mode="axisangle", fused_backward=True, disp_head=None, deterministic=None,
                 smoothness="second_order"):
        self.SSIM = SSIM()
        self.clip_loss = 0.5
        self.rotation_mode = rotation_mode
        self.fused_backward = fused_backward
        # disp_head=(alpha, beta): `forward` then takes the depth network's PRE-ACTIVATION maps and evaluates the
        # head `alpha * sigmoid(x) + beta` (models/depth/disp_net.py:121-139: alpha=10, beta=0.01) inside the
        # kernels, returning gradients with respect to x
        self.disp_head = disp_head
        # deterministic=True: image gradients (when a frame requires grad) are accumulated in order-independent
        # fixed point, so EVERY output is bitwise repeatable; None follows torch.use_deterministic_algorithms()
        self.deterministic = deterministic
        # smoothness="edge": `forward` returns [loss_mam, edge-aware smoothness of the target frame's disparity pyramid]
        # (`edge_aware_smooth_loss`, not in the reference) - evaluated inside the same call, its gradient accumulated into
        # the photometric term's maps; "second_order" is the reference's `smooth_loss` (losses.py:242-260)
        if smoothness not in ("second_order", "edge"):
            raise ValueError("smoothness must be 'second_order' or 'edge'")
        self.smoothness = smoothness

    # ---- live path -------------------------------------------------------
    def forward(self, tgt_img, ref_imgs, disparity, poses, intrinsics, gt=None):
        """`losses.py:262-271` -> [loss_mam, loss_smooth] (0-d CUDA tensors)."""
        pyr = [list(frame) if isinstance(frame, (list, tuple)) else [frame] for frame in disparity]
        mam, smooth = ops.fused_losses(tgt_img, list(ref_imgs), pyr, poses, intrinsics, input_is_depth=False,
                                       rotation_mode=self.rotation_mode, fused_backward=self.fused_backward,
                                       disp_head=self.disp_head, deterministic=self.deterministic,
                                       edge=self.smoothness == "edge")
        return [mam, smooth]

    __call__ = forward

    def capture(self, tgt_img, ref_imgs, disparity, poses, intrinsics, warmup=2):
        """`forward` + `sum(loss).backward()` (`trainer.py:312` + `:264`) captured as ONE CUDA graph over static
        copies of the arguments (plb200/graphed.py): `step = criterion.capture(...)`, then per training step
        `loss, grads = step(tgt, ref_imgs, disparity, poses, intrinsics)`.  Same kernels, bitwise the same results;
        the host cost of a step drops from ~170 us of autograd + launch issue to one graph launch."""
        from plb200.graphed import CapturedLossStep
        return CapturedLossStep(self, tgt_img, ref_imgs, disparity, poses, intrinsics, warmup=warmup)

    def reprojection_loss(self, tgt, refs, depths, poses, intrinsics, mode='min'):
        """`losses.py:183-240`; `depths` are depth (not disparity) pyramids per frame.
        mode 'min' is the mean the reference computes (`:226-228`)."""
        if mode != 'min':
            raise NotImplementedError("mode %r is dead code in the reference (undefined self.L2, losses.py:230-235)" % mode)
        pyr = [list(frame) if isinstance(frame, (list, tuple)) else [frame] for frame in depths]
        mam, _ = ops.fused_losses(tgt, list(refs), pyr, poses, intrinsics, input_is_depth=True, do_smooth=False,
                                  rotation_mode=self.rotation_mode, fused_backward=self.fused_backward,
                                  deterministic=self.deterministic)
        return mam

    def smooth_loss(self, pred_map):
        """`losses.py:242-260` on a depth map or a list of them."""
        if type(pred_map) not in [tuple, list]:
            pred_map = [pred_map]
        return ops.smooth_only(list(pred_map), fused_backward=self.fused_backward)

    def edge_aware_smooth_loss(self, disparity, tgt_img, normalize=True):
        """NOT in the reference (its smoothness is `smooth_loss`): the edge-aware first-order term of the
        monodepth2 lineage the model files cite (`models/depth/layers.py:1-2`), which north_star lists.
        `disparity`: one [B,1,h,w] map or a pyramid of them (scale s weighs 1/2^s)."""
        if type(disparity) not in [tuple, list]:
            disparity = [disparity]
        return ops.edge_aware_smooth(list(disparity), tgt_img, normalize=normalize)

    # ---- dormant path ----------------------------------------------------
    def compute_photometric_loss(self, pred, target, no_ssim=False):
        """`losses.py:66-84`: 0.85*SSIM + 0.15*L1 per channel, clamped at mean+0.5*std."""
        return ops.photometric_map(pred, target, no_ssim=no_ssim, clip=self.clip_loss)

    def multiview_reprojection_loss(self, tgt_img, ref_imgs, depth, poses, intrinsics, mode='min',
                                    automask=True, no_ssim=False, clip='default'):
        """Min-reprojection + automask composition (`losses.py:86-181` as the
        commented lines and `notes/toy_problem/losses.py:107-129` spell it).  Every photometric map goes
        through `compute_photometric_loss`, i.e. is clamped at mean + `self.clip_loss` * std of that map
        (`losses.py:79-82`); `clip=None` leaves the maps unclamped, a number replaces `self.clip_loss`."""
        if mode != 'min':
            raise NotImplementedError(mode)
        pyr = [list(depth) if isinstance(depth, (list, tuple)) else [depth]]
        flags = (_lib.PHOTO_NO_SSIM if no_ssim else 0) | (0 if automask else _lib.PHOTO_NO_AUTOMASK)
        clip_loss = self.clip_loss if clip == 'default' else clip
        mam, _ = ops.fused_losses(tgt_img, list(ref_imgs), pyr, poses, intrinsics, input_is_depth=True,
                                  do_smooth=False, rotation_mode=self.rotation_mode,
                                  fused_backward=self.fused_backward, mode=_lib.PHOTO_MIN_REPROJ, flags=flags,
                                  deterministic=self.deterministic, clip_loss=clip_loss)
        return mam
