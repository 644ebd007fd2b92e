"""CUDA-graphed loss step: `Losses.forward` + `sum(loss).backward()` captured once and replayed.

SURVEY.md section 8(f) rank 1, second half ("CUDA-graph the `process_batch` tail", `trainer.py:296-313`): at the
reference's batch sizes the fused loss is a handful of short launches, and an eager step costs more host time
(autograd bookkeeping + launch issue) than device time.  The captured step has the same inputs, the same kernels and
bitwise the same results; only the host work per step shrinks to one `cudaGraphLaunch`.

    step = criterion.capture(tgt, ref_imgs, disparity, poses, intrinsics)   # static buffers + one capture
    loss, grads = step(tgt, ref_imgs, disparity, poses, intrinsics)         # copy-in (device to device) + replay
    grads.disparity[frame][scale], grads.poses                              # static output buffers, valid until the next replay

The buffers a trainer feeds the loss with live in the networks' outputs, which move between steps; the copy-in is a
few device-to-device copies on the capture stream (26 MB at batch 12).  Callers that own their buffers pass nothing
and write into `step.inputs` directly.
"""
import torch

from . import ops


class GradBuffers:
    def __init__(self, disparity, poses):
        self.disparity = disparity
        self.poses = poses


class CapturedLossStep:
    def __init__(self, criterion, tgt, ref_imgs, disparity, poses, intrinsics, warmup=2):
        if not tgt.is_cuda:
            raise ValueError("capture() needs CUDA tensors")
        self.criterion = criterion
        dev = tgt.device
        pyr = [list(fr) if isinstance(fr, (list, tuple)) else [fr] for fr in disparity]
        clone = lambda t: t.detach().clone(memory_format=torch.contiguous_format)
        self.inputs = {
            "tgt": clone(tgt), "ref_imgs": [clone(r) for r in ref_imgs],
            "disparity": [[clone(d).requires_grad_(True) for d in fr] for fr in pyr],
            "poses": clone(poses).requires_grad_(True), "intrinsics": clone(intrinsics),
        }
        self._one = torch.ones((), dtype=torch.float32, device=dev)
        self._stream = torch.cuda.Stream(device=dev)
        self._stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self._stream):
            for _ in range(max(int(warmup), 1)):        # builds workspaces / argument caches outside the capture
                self._eager()
        torch.cuda.current_stream(dev).wait_stream(self._stream)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=self._stream):
            self.loss = self._eager()
        i = self.inputs
        self.grads = GradBuffers([[d.grad for d in fr] for fr in i["disparity"]], i["poses"].grad)

    def _eager(self):
        i = self.inputs
        for fr in i["disparity"]:
            for d in fr:
                d.grad = None
        i["poses"].grad = None
        # the step owns its backward call, so the upstream gradient of both losses is exactly 1: no guarded relaunch
        with ops.unit_upstream():
            loss = self.criterion.forward(i["tgt"], i["ref_imgs"], i["disparity"], i["poses"], i["intrinsics"], None)
            total = loss[0] + loss[1]
            total.backward(self._one)                   # a static upstream 1: no fill kernel inside the captured step
        self.total = total.detach()
        return [l.detach() for l in loss]

    def __call__(self, tgt=None, ref_imgs=None, disparity=None, poses=None, intrinsics=None):
        """Replay; any argument given is first copied into its static buffer (same shapes as at capture) - all of them
        in ONE multi-tensor copy, so the host cost does not grow with the number of pyramid levels."""
        i = self.inputs
        dst, src = [], []
        if tgt is not None:
            dst.append(i["tgt"]); src.append(tgt)
        if ref_imgs is not None:
            dst += list(i["ref_imgs"]); src += list(ref_imgs)
        if disparity is not None:
            for dfr, sfr in zip(i["disparity"], disparity):
                dst += list(dfr); src += list(sfr if isinstance(sfr, (list, tuple)) else [sfr])
        if poses is not None:
            dst.append(i["poses"]); src.append(poses)
        if intrinsics is not None:
            dst.append(i["intrinsics"]); src.append(intrinsics)
        if dst:
            if len(dst) != len(src):
                raise ValueError("arguments do not match the captured step's inputs")
            with torch.no_grad():
                torch._foreach_copy_(dst, [s_.detach() for s_ in src])
        self.graph.replay()
        return self.loss, self.grads
