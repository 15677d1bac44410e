"""Two ranks on two GPUs (NCCL): the data-parallel step of the hot path (SURVEY.md 8(e), section 4 "multi-GPU test").
Each rank runs its own images through the native chains with the ChainGradReducer; the averaged gradients must equal
the average of the two per-rank gradients computed by ONE process (the loss is a per-rank mean, base_distillator.py:64,
so the N-rank result is the mean of the per-rank results). Needs two devices: run with `gpurun --gpus 2`; skipped on a
single GPU. The CPU-side logic of the same classes is covered by tests/test_dist_gloo.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lgd_b200 import synth

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _grads(model, bi, im, feats, dev, reducer=None):
    for p in model.parameters():
        p.grad = None
    f = {k: v.to(dev).requires_grad_(True) for k, v in feats.items()}
    tea, _, _, loss = model.forward(bi, im, f)
    cot = synth.synth_cotangents({k: v.detach().cpu() for k, v in tea.items()})
    keys = list(tea.keys())
    torch.autograd.backward([loss] + [tea[k] for k in keys], [torch.ones_like(loss)] + [cot[k].to(dev) for k in keys])
    n = reducer.finish() if reducer is not None else 0
    torch.cuda.synchronize(dev)
    return float(loss), {name: p.grad.detach().cpu().clone() for name, p in model.named_parameters() if p.grad is not None}, n


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from lgd_b200.dist import ChainGradReducer
        from lgd_b200.step import HotPathDistillator
        model = HotPathDistillator(synth.make_cfg(device="cuda", add_context_box=True))
        model.load_hot_path_state_dict(synth.synth_state_dict(5))
        model = model.to(dev)
        batches = [synth.synth_batch(2, 200, 264, seed=300 + r) for r in range(world)]
        reducer = ChainGradReducer()
        loss, g, n = _grads(model, *batches[rank], dev, reducer)
        assert n == 3, "adapter chain: one collective; teacher chain: early part + student_proj_2D at the end"
        # adopted, not copied: the .grad tensors are the chain's views that NCCL averaged in place
        reducer.close()
        if rank == 0:
            # reference: both per-rank steps in this process, no collective, averaged on the host
            ref = None
            for r in range(world):
                _, gr, _ = _grads(model, *batches[r], dev, None)
                ref = gr if ref is None else {k: ref[k] + gr[k] for k in ref}
            ref = {k: v / world for k, v in ref.items()}
            worst = 0.0
            for k in ref:
                err = float((g[k].double() - ref[k].double()).norm() / ref[k].double().norm().clamp_min(1e-30))
                if k.endswith("adapter.4.bias"):   # analytically zero gradient: absolute round-off only
                    continue
                worst = max(worst, err)
            out.put(worst)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_rank_average_equals_one_rank_average():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    worst = q.get(timeout=10)
    print("2-rank averaged gradients vs 1-rank average of per-rank gradients: worst relative L2", worst)
    assert worst <= 1e-6
