"""CPU, world_size 2 over gloo: the N>1 host logic of the hot path -- per-image sharding and the single flat
gradient all-reduce. The CUDA kernels themselves are rank-local and covered by the -m gpu tests."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lgd_b200 import synth
from lgd_b200.dist import FlatGradBucket, shard_images


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        lin = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 3))
        bucket = FlatGradBucket(lin.parameters())
        bi, _, _ = synth.synth_batch(4, 64, 64, seed=11)
        mine = shard_images(bi, rank, world)
        assert len(mine) == 2 and mine[0] is bi[2 * rank]
        # rank-dependent "loss": per-rank mean, like the reference's per-rank MSE mean
        x = torch.full((2, 6), float(rank + 1))
        bucket.zero_()
        lin(x).mean().backward()
        local = bucket.flat.clone()
        # autograd accumulated INTO the flat views (no re-allocation of .grad)
        assert all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in lin.parameters())
        red = bucket.all_reduce_mean().clone()
        gathered = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        assert torch.allclose(red, torch.stack(gathered).mean(0), atol=1e-7)
        assert torch.equal(lin[0].weight.grad.reshape(-1), red[:30])
        # logging reduction of train.py:196 with one collective / one host read (lgd_b200.optim.reduce_loss_dict)
        from lgd_b200.optim import reduce_loss_dict
        logged = reduce_loss_dict({"loss_cls": torch.tensor(float(rank + 1)), "loss_distill": torch.tensor(2.0 * (rank + 1))})
        if rank == 0:
            assert logged == {"loss_cls": 1.5, "loss_distill": 3.0}, logged
            out.put(red.tolist())
    finally:
        dist.destroy_process_group()


def test_flat_gradient_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    red = q.get(timeout=5)
    assert len(red) == 6 * 5 + 5 + 5 * 3 + 3


def test_shard_requires_divisibility():
    import pytest
    with pytest.raises(ValueError):
        shard_images(list(range(5)), 0, 2)
