"""CPU, world_size 2 over gloo: the batch-sharding arithmetic of plb200/dist.py.
The per-shard loss is evaluated with the oracle (this is a test), all-reduced with the
product's host logic, and must equal the oracle on the whole batch; shard gradients
scaled by B_r/B must equal the whole-batch gradients of the same items."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "unsupervised-pseuso-lidar_b200")]
    from plb200 import synth, dist as pdist
    from oracle import restated as O
    torch.set_num_threads(1)
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    B = 5                                        # odd on purpose: shards of 3 and 2
    full = synth.make_photo_inputs(B, 24, 40, n_src=2, n_scales=2, seed=77)
    shard, (lo, hi) = pdist.shard_sample(full, rank, world)
    disp = [[d.clone().requires_grad_(True) for d in fr] for fr in shard["disparity"]]
    poses = shard["poses"].clone().requires_grad_(True)
    loss = O.losses_forward(shard["tgt"], shard["ref_imgs"], disp, poses, shard["intrinsics"])
    (sum(loss) * pdist.local_loss_weight(hi - lo, B)).backward()
    glob = pdist.allreduce_losses(loss, hi - lo, B)
    tmax = pdist.max_over_ranks(float(rank + 1))
    # reference: the whole batch in one process
    fd = [[d.clone().requires_grad_(True) for d in fr] for fr in full["disparity"]]
    fp = full["poses"].clone().requires_grad_(True)
    fl = O.losses_forward(full["tgt"], full["ref_imgs"], fd, fp, full["intrinsics"])
    sum(fl).backward()
    ok = True
    # loss_mam is a mean over items -> exact under sharding; loss_smooth likewise
    for a, b in zip(glob, fl):
        ok &= abs(float(a) - float(b)) <= 1e-5 * abs(float(b))
    ok &= torch.allclose(poses.grad, fp.grad[lo:hi], rtol=1e-4, atol=1e-9)
    ok &= torch.allclose(disp[0][0].grad, fd[0][0].grad[lo:hi], rtol=1e-4, atol=1e-10)
    ok &= tmax == float(world)
    # the exchange step: bucketed gradient all-reduce (mean) with the loss scalars fused into the last bucket
    params = [torch.nn.Parameter(torch.zeros(n)) for n in (1000, 37, 5000)]
    for k, p_ in enumerate(params):
        p_.grad = torch.full_like(p_, float(rank + 1) * (k + 1))
    red = pdist.GradBucketReducer(params, bucket_mb=0.008)            # 2000 floats per bucket: four buckets
    ok &= len(red.bounds) == 4 and red.bounds[-1][1] == red.n + 2
    red.launch(losses=loss, B_local=hi - lo, B_global=B)
    means = red.wait()
    for k, p_ in enumerate(params):
        ok &= bool(torch.allclose(p_.grad, torch.full_like(p_, 1.5 * (k + 1))))   # mean of (1, 2) * (k + 1)
    for a, b in zip(means, fl):
        ok &= abs(float(a) - float(b)) <= 1e-5 * abs(float(b))
    flat = torch.arange(10, dtype=torch.float32) * (rank + 1)          # a flat arena, two trailing scalar slots
    red2 = pdist.GradBucketReducer(flat, bucket_mb=25)
    red2.launch()
    red2.wait()
    ok &= bool(torch.allclose(flat[:8], torch.arange(8, dtype=torch.float32) * 1.5)) and float(flat[8:].abs().sum()) == 0.0
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_shard_bounds():
    sys.path[:0] = [ROOT, os.path.join(ROOT, "unsupervised-pseuso-lidar_b200")]
    from plb200.dist import shard_bounds
    for B in (1, 4, 5, 12, 64):
        for world in (1, 2, 3, 8):
            cuts = [shard_bounds(B, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == B
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_two_rank_loss_and_grad_equal_single_process():
    world, port = 2, 29000 + os.getpid() % 2000
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}
