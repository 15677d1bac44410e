"""Data-parallel plumbing of the hot path: the path shards per image with no activation exchange;
the only collective is ONE all-reduce (average) of the hot-path gradients per step -- 10.07 M fp32 = 40.3 MB -- over
NCCL / NVLink (the reference lets DDP bucket them, train.py:277-281). `torch.distributed` is plumbing here."""
from __future__ import annotations

from typing import Iterable, List, Sequence

import torch
import torch.distributed as dist


class FlatGradBucket:
    """All hot-path parameter gradients as views of one flat fp32 buffer, so that a step needs exactly one
    all-reduce. Autograd accumulates into the views in place (the .grad tensors are pre-set)."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatGradBucket needs at least one trainable parameter")
        dev = self.params[0].device
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, device=dev, dtype=torch.float32)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero_(self):
        self.flat.zero_()

    def all_reduce_mean(self, group=None):
        """Average over ranks (what DDP does to the gradients). No-op for a single process."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            if dist.get_backend(group) == "nccl":
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group)
            else:  # gloo (CPU tests) has no AVG
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
                self.flat.div_(dist.get_world_size(group))
        return self.flat


def shard_images(batched_inputs: Sequence, rank: int, world: int):
    """Per-image data parallelism: rank r takes a contiguous block of IMS_PER_BATCH / world images
    (utils/build.py:281-288 of the reference requires divisibility, so do we)."""
    n = len(batched_inputs)
    if n % world != 0:
        raise ValueError("batch of %d images is not divisible by world size %d" % (n, world))
    per = n // world
    return list(batched_inputs[rank * per:(rank + 1) * per])



class ChainGradReducer:
    """Gradient averaging for the native chains (lgd_b200/csrc/chain.cu): every chain writes its parameter gradients
    into ONE flat buffer. The adapter's 7 MB are all-reduced as soon as the distillation backward has enqueued its last
    kernel (the collective then runs on NCCL's stream underneath the whole teacher backward). The teacher's 33 MB go in
    two parts: everything except student_proj_2D (31 MB) as soon as its producers have finished -- on a side stream
    that waits for the chain's own events (lgd_ctx_wait_early_grads), i.e. underneath the student-side end of the
    backward (pooling backward, GroupNorm backward, student_proj dgrad / wgrad) -- and student_proj_2D's 2.4 MB at the
    end of the chain. No gradient bucket is zeroed and no per-parameter accumulation kernel runs: with `.grad = None`
    before the backward (optimizer.zero_grad(set_to_none=True), the reference's loop at train.py:201-202 on current
    PyTorch) autograd adopts the chain's views as the `.grad` tensors, which NCCL then averages in place. finish()
    makes the current stream wait for the collectives; call it before the optimizer step.

    The reference lets DDP bucket these gradients (train.py:277-281); the student's own parameters stay with DDP."""

    def __init__(self, group=None, split_teacher: bool = True):
        from . import engine
        self.group = group
        self.split_teacher = split_teacher
        self.pending = []
        self.engine = engine
        self._side = {}
        self._hook = self._on_ready
        engine.GRAD_READY_HOOKS.append(self._hook)

    def close(self):
        if self._hook in self.engine.GRAD_READY_HOOKS:
            self.engine.GRAD_READY_HOOKS.remove(self._hook)

    def _all_reduce(self, kind, flat):
        if dist.get_backend(self.group) == "nccl":
            work = dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
        else:
            work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self.pending.append((kind, work, flat))

    def _on_ready(self, kind, flat, n_early=None, wait_early=None):
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1):
            return
        if (self.split_teacher and wait_early is not None and n_early and 0 < n_early < flat.numel() and flat.is_cuda
                and dist.get_backend(self.group) == "nccl"):
            # the early part: NCCL's stream synchronises with the stream the collective is issued from -- a side stream that
            # only waits for the producers of flat[:n_early], not for the rest of the chain on the current stream
            side = self._side.get(flat.device)
            if side is None:
                side = self._side[flat.device] = torch.cuda.Stream(flat.device)
            wait_early(side)
            with torch.cuda.stream(side):
                self._all_reduce(kind + ".early", flat[:n_early])
            self._all_reduce(kind + ".late", flat[n_early:])
            return
        self._all_reduce(kind, flat)

    def finish(self):
        """Current stream waits for every collective started during the backward. Returns the number of collectives."""
        n = len(self.pending)
        for kind, work, flat in self.pending:
            work.wait()
            if dist.get_backend(self.group) != "nccl":
                flat.div_(dist.get_world_size(self.group))
        self.pending.clear()
        return n
