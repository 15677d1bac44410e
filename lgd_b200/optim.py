"""Training-loop pieces next to the hot path (reference: train.py:191-215, utils/build.py:492-529).

The reference builds two optimizers -- student + adapter, teacher -- with ONE PARAMETER GROUP PER PARAMETER
(utils/build.py:497-508) and reads every reduced loss with its own `.item()` (train.py:196). On a step that takes
12 ms those are ~160 tiny launches and up to six host synchronisations. Here:

  * FusedSGD / FusedAdamW: `torch.optim.Optimizer` subclasses (same param_groups / state_dict layout, so the reference's
    LR schedulers and checkpointer work unchanged) whose step() updates every tensor that shares hyper-parameters with
    ONE launch of liblgd_b200's multi-tensor kernel (lgd_mt_sgd / lgd_mt_adamw);
  * build_distillator_optimizer(cfg, network): drop-in for utils/build.py:492-529 returning (stu_optim, tea_optim) with
    the same grouping (student + adapter | teacher), the same per-parameter groups and solver keys;
  * reduce_loss_dict(loss_dict): the logging reduction of train.py:196 with ONE collective and ONE device->host copy.

There is no CPU fallback: parameters must live on a CUDA device (fp32)."""
from __future__ import annotations

import ctypes
from typing import Any, Dict, List, Set

import torch
import torch.distributed as dist

from . import _lib
from ._lib import call


class _MtTensor(ctypes.Structure):
    """lgd_mt_tensor_t"""
    _fields_ = [("param", ctypes.c_void_p), ("grad", ctypes.c_void_p), ("state0", ctypes.c_void_p),
                ("state1", ctypes.c_void_p), ("numel", ctypes.c_int64)]


class _FusedBase(torch.optim.Optimizer):
    """Shared machinery: per step the parameters that have a gradient are bucketed by their hyper-parameter tuple;
    each bucket is one launch. The device tables of a bucket (tensor descriptors + chunk list) are cached by the
    identity of (param, grad, state) storage, so a steady-state step uploads nothing."""

    def __init__(self, params, defaults):
        super().__init__(params, defaults)
        self._tables: Dict[tuple, tuple] = {}
        self._chunk = _lib.load().lgd_mt_chunk_elems()

    def _table(self, entries):
        """entries: list of (param, grad, state0, state1 or None). Returns (tensors_dev, chunks_dev, num_chunks, keepalive)."""
        key = tuple((p.data_ptr(), g.data_ptr(), s0.data_ptr(), s1.data_ptr() if s1 is not None else 0) for p, g, s0, s1 in entries)
        hit = self._tables.get(key)
        if hit is not None:
            return hit
        if len(self._tables) > 16:
            self._tables.clear()
        dev = entries[0][0].device
        arr = (_MtTensor * len(entries))()
        chunks: List[int] = []
        for i, (p, g, s0, s1) in enumerate(entries):
            arr[i].param, arr[i].grad, arr[i].state0 = p.data_ptr(), g.data_ptr(), s0.data_ptr()
            arr[i].state1 = s1.data_ptr() if s1 is not None else 0
            arr[i].numel = p.numel()
            for c in range((p.numel() + self._chunk - 1) // self._chunk):
                chunks += [i, c]
        raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).clone()
        t_dev = raw.to(dev)
        c_dev = torch.tensor(chunks, dtype=torch.int32).to(dev)
        hit = self._tables[key] = (t_dev, c_dev, len(chunks) // 2)
        return hit

    @staticmethod
    def _check(p):
        if p.device.type != "cuda" or p.dtype != torch.float32 or not p.is_contiguous():
            raise RuntimeError("lgd_b200 fused optimizers need contiguous fp32 CUDA parameters (no CPU fallback)")
        g = p.grad
        if g.is_sparse:
            raise RuntimeError("sparse gradients are not supported")
        if g.dtype != torch.float32 or not g.is_contiguous():
            g = g.float().contiguous()   # rare (a channels_last view): dense copy, plumbing
        return g


class FusedSGD(_FusedBase):
    """torch.optim.SGD(params, lr, momentum, weight_decay) -- dampening 0, no nesterov, as the reference builds it
    (utils/build.py:512-513) -- with one launch per hyper-parameter bucket."""

    def __init__(self, params, lr, momentum=0.0, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay, dampening=0, nesterov=False))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        buckets: Dict[tuple, list] = {}
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                g = self._check(p)
                st = self.state[p]
                first = "momentum_buffer" not in st or st["momentum_buffer"] is None
                if first:
                    st["momentum_buffer"] = torch.empty_like(p, memory_format=torch.contiguous_format)
                key = (float(group["lr"]), float(group["weight_decay"]), float(group["momentum"]), first, p.device.index)
                buckets.setdefault(key, []).append((p, g, st["momentum_buffer"], None))
        for (lr, wd, mu, first, dev_idx), entries in buckets.items():
            with torch.cuda.device(dev_idx):
                t_dev, c_dev, n = self._table(entries)
                call("lgd_mt_sgd", ctypes.c_void_p(t_dev.data_ptr()), ctypes.c_void_p(c_dev.data_ptr()), n, lr, wd, mu,
                     int(first))
        return loss


class FusedAdamW(_FusedBase):
    """torch.optim.AdamW(params, lr, betas=(0.9, 0.999)) as the reference builds it (utils/build.py:514-515)."""

    def __init__(self, params, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        buckets: Dict[tuple, list] = {}
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                g = self._check(p)
                st = self.state[p]
                if "exp_avg" not in st:
                    st["step"] = torch.tensor(0.0)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] += 1
                key = (float(group["lr"]), float(group["weight_decay"]), float(b1), float(b2), float(group["eps"]),
                       int(st["step"]), p.device.index)
                buckets.setdefault(key, []).append((p, g, st["exp_avg"], st["exp_avg_sq"]))
        for (lr, wd, b1, b2, eps, step, dev_idx), entries in buckets.items():
            with torch.cuda.device(dev_idx):
                t_dev, c_dev, n = self._table(entries)
                call("lgd_mt_adamw", ctypes.c_void_p(t_dev.data_ptr()), ctypes.c_void_p(c_dev.data_ptr()), n, lr, wd, b1,
                     b2, eps, step)
        return loss


def build_distillator_optimizer(cfg, network):
    """Drop-in for utils/build.py:492-529: (stu_optim over student + adapter, tea_optim over teacher), one parameter
    group per parameter with the solver's BASE_LR / WEIGHT_DECAY, optimizer type from SOLVER.OPTIMIZER."""
    solver_stu, solver_tea = cfg.MODEL.DISTILLATOR.STUDENT.SOLVER, cfg.MODEL.DISTILLATOR.TEACHER.SOLVER
    net = network.module if hasattr(network, "module") else network

    def _get_params(model_list, base_lr, wd):
        params: List[Dict[str, Any]] = []
        memo: Set[torch.nn.parameter.Parameter] = set()
        for model in model_list:
            for _, value in model.named_parameters(recurse=True):
                if not value.requires_grad or value in memo:
                    continue
                memo.add(value)
                params += [{"params": [value], "lr": base_lr, "weight_decay": wd}]
        return params

    def _get_optim(optimizer_type, params, base_lr, momentum):
        if optimizer_type == "SGD":
            return FusedSGD(params, base_lr, momentum=momentum)
        if optimizer_type == "ADAMW":
            return FusedAdamW(params, base_lr, betas=(0.9, 0.999))
        raise NotImplementedError(f"no optimizer type {optimizer_type}")

    stu_params = _get_params([net.student, net.adapter], solver_stu.BASE_LR, solver_stu.WEIGHT_DECAY)
    tea_params = _get_params([net.teacher], solver_tea.BASE_LR, solver_tea.WEIGHT_DECAY)
    return (_get_optim(solver_stu.OPTIMIZER, stu_params, solver_stu.BASE_LR, solver_stu.MOMENTUM),
            _get_optim(solver_tea.OPTIMIZER, tea_params, solver_tea.BASE_LR, solver_tea.MOMENTUM))


def publish_scalars(values: torch.Tensor, pinned: torch.Tensor) -> torch.cuda.Event:
    """Asynchronous read-back of a few device scalars (the step's losses) into a pinned host tensor: a kernel stores
    them (lgd_store_to_host), so the read-back does not go through a copy engine (they work in order and may be busy
    with a long transfer). Returns the event to synchronise before reading `pinned` (train.py:196's `.item()`
    without draining the stream)."""
    from ._lib import call, ptr
    v = values.detach().reshape(-1).float()
    if not v.is_cuda or not pinned.is_pinned() or pinned.dtype != torch.float32 or pinned.numel() < v.numel():
        raise ValueError("publish_scalars: CUDA float values and a pinned float32 host tensor of at least that size")
    with torch.cuda.device(v.device):
        call("lgd_store_to_host", ptr(v.contiguous()), pinned.data_ptr(), v.numel())
        ev = torch.cuda.Event()
        ev.record()
    return ev


def reduce_loss_dict(loss_dict: Dict[str, torch.Tensor], average: bool = True, group=None) -> Dict[str, float]:
    """train.py:196 (`{k: v.item() for k, v in comm.reduce_dict(loss_dict).items()}`) with one collective and ONE
    device->host copy: the scalars are stacked in sorted key order (as detectron2's reduce_dict does), reduced to rank 0
    and read back together. Every rank returns a dict; only rank 0's is averaged over the world (as in the reference)."""
    names = sorted(loss_dict.keys())
    if not names:
        return {}
    with torch.no_grad():
        values = torch.stack([loss_dict[k].detach().reshape(()).float() for k in names], dim=0)
        world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        if world > 1:
            dist.reduce(values, dst=0, group=group)
            if dist.get_rank(group) == 0 and average:
                values = values / world
        host = values.tolist()   # the single synchronisation of the logging path
    return dict(zip(names, host))
