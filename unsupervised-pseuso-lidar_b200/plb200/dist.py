"""Batch data-parallel plumbing for the loss path.

The hot path needs no data-path collective: every batch item (and pixel) is
independent.  Rank r takes a contiguous slice of the batch; its loss is a mean
over ITS items, so the global loss is sum_r (B_r / B) * loss_r - one tiny
all-reduce of the two logged scalars.  Gradients w.r.t. disparity / pose belong
to the local items; scaled by B_r / B they equal the single-process gradients
(the network-gradient all-reduce itself belongs to the trainer's DDP wrapper
and is outside this path).
"""
import torch
import torch.distributed as dist


def shard_bounds(B, rank, world):
    """Contiguous, balanced slice [lo, hi) of a batch of B items for `rank`."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sample(sample, rank, world):
    """Slice a reference-layout sample dict (synth.make_photo_inputs) along the batch."""
    B = sample["tgt"].shape[0]
    lo, hi = shard_bounds(B, rank, world)

    def cut(x):
        if isinstance(x, torch.Tensor):
            return x[lo:hi].contiguous()
        if isinstance(x, (list, tuple)):
            return type(x)(cut(v) for v in x)
        return x
    return {k: cut(v) for k, v in sample.items()}, (lo, hi)


def local_loss_weight(B_local, B_global):
    """Factor that turns a rank-local batch-mean loss into its share of the global mean."""
    return float(B_local) / float(B_global)


def allreduce_losses(losses, B_local, B_global, group=None):
    """losses: list of 0-d tensors (rank-local means) -> global means on every rank.
    One collective for all scalars (NCCL over NVLink on GPU tensors, gloo on CPU)."""
    stacked = torch.stack([l.detach().reshape(()) for l in losses]) * local_loss_weight(B_local, B_global)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stacked, op=dist.ReduceOp.SUM, group=group)
    return [stacked[i] for i in range(len(losses))]


def max_over_ranks(value, device=None, group=None):
    """Max of a python float over ranks (device timings are reported as the max)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t[0])


class GradBucketReducer:
    """The exchange step of one data-parallel training step: all-reduce (mean) of the depth +
    pose network gradients in buckets, on a side stream so that it overlaps whatever the main stream computes next,
    with the two logged loss scalars riding at the tail of the LAST bucket (no collective of their own).

        red = GradBucketReducer(params_or_flat_buffer, bucket_mb=25)
        ...loss backward on the main stream...
        red.launch(losses=[loss_mam, loss_smooth], B_local=b, B_global=B)   # returns at once
        ...next step's compute...
        means = red.wait()      # main stream now sees averaged gradients; means = global [loss_mam, loss_smooth]

    `grads`: a flat float32 tensor (the gradient arena a DDP-style wrapper keeps) or a list of tensors with `.grad`.
    Works on CPU tensors with gloo (no streams) - that is how tests/test_dist_cpu.py drives it."""

    def __init__(self, grads, bucket_mb=25.0, group=None):
        self.group = group
        self.flat = grads if isinstance(grads, torch.Tensor) else None
        self.params = None if self.flat is not None else [p for p in grads]
        if self.flat is None:
            n = sum(p.numel() for p in self.params)
            self.flat = torch.zeros(n + 2, dtype=torch.float32, device=self.params[0].device)
        self.cuda = self.flat.is_cuda
        self.n = self.flat.numel() - 2                       # the last two elements carry the loss scalars
        if self.n < 0:
            raise ValueError("the flat gradient buffer needs two trailing elements for the loss scalars")
        per = max(1, int(bucket_mb * 1e6 / 4))
        self.bounds = []
        lo = 0
        while lo < self.n:
            hi = min(self.n, lo + per)
            self.bounds.append((lo, hi))
            lo = hi
        if not self.bounds:
            self.bounds = [(0, 0)]
        lo, hi = self.bounds[-1]
        self.bounds[-1] = (lo, hi + 2)                       # scalars fused into the last bucket
        self.stream = torch.cuda.Stream(device=self.flat.device) if self.cuda else None
        self.ready = torch.cuda.Event() if self.cuda else None
        self.done = torch.cuda.Event() if self.cuda else None
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1

    def _pack(self):
        if self.params is not None:
            off = 0
            for p in self.params:
                k = p.numel()
                self.flat[off:off + k].copy_(p.grad.reshape(-1))
                off += k

    def _unpack(self):
        if self.params is not None:
            off = 0
            for p in self.params:
                k = p.numel()
                p.grad.copy_(self.flat[off:off + k].view_as(p.grad))
                off += k

    def launch(self, losses=None, B_local=1, B_global=None):
        """Enqueue the exchange behind everything the main stream has been given so far."""
        w = local_loss_weight(B_local, B_global if B_global else B_local * self.world)
        ctx = torch.cuda.stream(self.stream) if self.cuda else _Null()
        if self.cuda:
            self.ready.record(torch.cuda.current_stream(self.flat.device))
            self.stream.wait_event(self.ready)
        with ctx:
            self._pack()
            tail = self.flat[self.n:]
            if losses is not None:
                tail.copy_(torch.stack([l.detach().reshape(()).to(torch.float32) for l in losses]) * (w * self.world))
            else:
                tail.zero_()
            if self.world > 1:
                for lo, hi in self.bounds:
                    dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group)
                self.flat.mul_(1.0 / self.world)             # DDP semantics: mean over ranks
            if self.cuda:
                self.done.record(self.stream)

    def wait(self):
        """Main stream waits for the exchange; returns the global means of the loss scalars (a 2-element view)."""
        if self.cuda:
            torch.cuda.current_stream(self.flat.device).wait_event(self.done)
        self._unpack()
        return self.flat[self.n:]


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
