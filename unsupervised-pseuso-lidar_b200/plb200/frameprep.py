"""GPU replacement of the reference loader's per-frame transform chain.

The reference decodes a frame with PIL and runs, in every DataLoader worker and per frame,
`ToTensor -> ToPILImage -> Resize((H, W)) -> ToTensor -> Normalize` (trainer.py:97-103 through
KittiDataset.load_img, dataloaders.py:32-49), then scales the intrinsics (dataloaders.py:95-98) and ships float32
frames to the GPU (trainer.py:291-299).  `FramePrep` does the same arithmetic - bit for bit - on the device, on a
whole batch of decoded uint8 frames: the host-to-device copy carries one byte per channel instead of four, and the
loss kernels read what the networks read.

    prep = FramePrep(img_height, img_width)                 # config['datasets']['augmentation']
    batch = prep(frames_u8.cuda(non_blocking=True), K)      # frames_u8: [B, h, w, 3] uint8 as decoded
    tgt, K = batch["planar"], batch["K"]
"""
import torch

from . import ops


class FramePrep:
    def __init__(self, img_height, img_width, mean=ops.IMAGENET_MEAN, std=ops.IMAGENET_STD, nhwc4=False):
        self.height, self.width = int(img_height), int(img_width)
        self.mean, self.std, self.nhwc4 = tuple(mean), tuple(std), bool(nhwc4)

    def __call__(self, frames, intrinsics=None, out=None):
        """frames: uint8 [B,h,w,3] (or one [h,w,3] frame) on the GPU; intrinsics: [B,3,3] at the decoded size."""
        single = frames.dim() == 3
        if single:
            frames = frames.unsqueeze(0)
            if intrinsics is not None:
                intrinsics = intrinsics.reshape(1, 3, 3)
        res = ops.prep_frames(frames, self.height, self.width, K=intrinsics, mean=self.mean, std=self.std,
                              want_nhwc4=self.nhwc4, out=out)
        if single:
            res = {k: v[0] for k, v in res.items()}
        return res
