"""Helpers for the GPU parity tests (run only with a CUDA device)."""
import ctypes

import torch

from lgd_b200 import _lib, engine, synth
from lgd_b200._lib import call, ptr
from lgd_b200.step import HotPathDistillator


def rel_l2(a, b):
    a = torch.as_tensor(a).detach().double().cpu().reshape(-1)
    b = torch.as_tensor(b).detach().double().cpu().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def round_tf32_cpu(x):
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def pyr_to_nchw_cpu(g, buf):
    """pyramid buffer (cuda) -> list of (B,256,h,w) cpu tensors"""
    return [v.contiguous().cpu() for v in g.level_views(buf)]


def nchw_to_pyr(g, tensors, rnd=False):
    return engine.to_pyramid(g, [t.cuda() for t in tensors], rnd)


def make_model(cfg_kw, sd, flag=1):
    cfg = synth.make_cfg(device="cuda", **cfg_kw)
    m = HotPathDistillator(cfg)
    m.load_hot_path_state_dict(sd)
    m = m.cuda()
    m.distill_flag = flag
    m.teacher.keep_tape = True   # the tests read intermediate stages from teacher._last
    return m


def run_engine(cfg_kw, sd, bi, im, feats, flag=1, backward=True):
    m = make_model(cfg_kw, sd, flag)
    f = {k: v.detach().clone().cuda().requires_grad_(True) for k, v in feats.items()}
    tea, inst_labels, masks, loss = m.forward(bi, im, f)
    out = dict(model=m, tea={k: v.detach().cpu().contiguous() for k, v in tea.items()}, loss=float(loss),
               inst_labels=inst_labels, masks=[[x.cpu() for x in lvl] for lvl in masks], feats=f)
    if backward:
        cot = synth.synth_cotangents({k: v.detach().cpu() for k, v in tea.items()})
        keys = list(tea.keys())
        torch.autograd.backward([loss] + [tea[k] for k in keys],
                                [torch.ones_like(loss)] + [cot[k].cuda() for k in keys])
        out["gfeat"] = {k: (v.grad.cpu() if v.grad is not None else None) for k, v in f.items()}
        out["gparam"] = {n: (p.grad.cpu() if p.grad is not None else None) for n, p in m.named_parameters()}
    torch.cuda.synchronize()
    return out
