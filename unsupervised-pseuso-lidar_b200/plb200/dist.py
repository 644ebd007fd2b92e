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
