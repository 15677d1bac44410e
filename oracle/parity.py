"""TEST INFRASTRUCTURE ONLY -- gradient parity of the CUDA engine against the CPU oracle at any size.

Imported by tests/ (the `-m gpu` parity tests) and by bench.py's `parity` block; never by lgd_b200/.

Why two comparisons. The path has six pyramid-sized ReLU sites (oracle.lgd_oracle._relu_site). Any implementation
whose forward differs from the fp32 reference by eps (here: 10-bit-mantissa tensor-core operands, ~3e-4) decides the
sign of the ~eps-fraction of activations that are zero to within eps differently, and each such decision changes that
entry's backward contribution by 100 % -- a discrete effect that no kernel accuracy removes (the reference itself
shows it between fp32 and fp64, or between two cuDNN algorithms). So the step is judged by
  (1) the FLIP FRACTION: share of activations whose pattern differs from the fp32 oracle's, and how close to zero the
      oracle's pre-activation is at those entries (they must be rounding-level ties, not errors);
  (2) gradients against the oracle evaluated WITH THE ENGINE'S ACTIVATION PATTERN (y = x * pattern): every feature and
      parameter gradient within the 1e-3 bar of SURVEY.md 8(d) -- this is the statement about kernel accuracy;
  (3) for reference, the same gradients against the plain fp32 oracle (dominated by the flips of (1)).
"""
from __future__ import annotations

from typing import Dict

import torch

from lgd_b200 import synth
from oracle import lgd_oracle as O

SITES = ("sp", "y0", "y1", "y2", "a1", "a2")


def _rel(a, b):
    a = torch.as_tensor(a).detach().double().cpu().reshape(-1)
    b = torch.as_tensor(b).detach().double().cpu().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _oracle_grads(sd, bi, im, feats, cfg_kw, flag, ctl, cot=None):
    f = {k: v.clone().requires_grad_(True) for k, v in feats.items()}
    sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    tea, _, masks, loss, _ = O.distill_step(sdo, bi, im, f, distill_flag=flag, relu_ctl=ctl, **cfg_kw)
    if cot is None:
        cot = synth.synth_cotangents(tea)
    total = loss + sum((tea[k] * cot[k]).sum() for k in tea)
    names = sorted(sdo)
    grads = torch.autograd.grad(total, list(f.values()) + [sdo[n] for n in names], allow_unused=True)
    gfeat = {k: grads[i] for i, k in enumerate(f)}
    gparam = {n: grads[len(f) + i] for i, n in enumerate(names)}
    return dict(tea={k: v.detach() for k, v in tea.items()}, loss=float(loss), gfeat=gfeat, gparam=gparam, masks=masks,
                cot=cot)


def engine_step(cfg_kw, sd, bi, im, feats, flag=1):
    """One forward + backward through the plugin classes on cuda:0; returns outputs, gradients and the activation
    patterns the backward used (engine.relu_patterns)."""
    from lgd_b200 import engine
    from lgd_b200.step import HotPathDistillator
    m = HotPathDistillator(synth.make_cfg(device="cuda", **cfg_kw))
    m.load_hot_path_state_dict(sd)
    m = m.cuda()
    m.distill_flag = flag
    m.teacher.keep_tape = True
    f = {k: v.detach().clone().cuda().requires_grad_(True) for k, v in feats.items()}
    tea, inst_labels, masks, loss = m.forward(bi, im, f)
    St = m.teacher._last
    Sd = getattr(loss.grad_fn, "S", None)
    pats = engine.relu_patterns(St, Sd)
    keys = list(tea.keys())
    pats = {s + "/" + k: p[l].cpu() for s, p in pats.items() for l, k in enumerate(keys)}
    cot = synth.synth_cotangents({k: v.detach().cpu() for k, v in tea.items()})
    out = dict(tea={k: v.detach().cpu().contiguous() for k, v in tea.items()}, loss=float(loss), patterns=pats,
               masks=[[x.cpu() for x in lvl] for lvl in masks], cot=cot)
    torch.autograd.backward([loss] + [tea[k] for k in keys], [torch.ones_like(loss)] + [cot[k].cuda() for k in keys])
    torch.cuda.synchronize()
    out["gfeat"] = {k: (v.grad.cpu() if v.grad is not None else None) for k, v in f.items()}
    out["gparam"] = {n: (p.grad.cpu() if p.grad is not None else None) for n, p in m.named_parameters()}
    return out


def _grad_errors(eng, ora):
    """relative-L2 error of every gradient the oracle has, keyed 'feat/<k>' or the parameter name. Parameters whose
    reference gradient is analytically zero (adapter.4.bias: InstanceNorm removes channel constants) are compared in
    absolute terms against the gradient scale of the same layer's weight and reported under 'abs/<name>'."""
    errs: Dict[str, float] = {}
    for k, gr in ora["gfeat"].items():
        got = eng["gfeat"][k]
        if gr is None:
            assert got is None or float(got.abs().max()) == 0.0, k
            continue
        errs["feat/" + k] = _rel(got, gr)
    for n, gr in ora["gparam"].items():
        got = eng["gparam"][n]
        if gr is None:
            assert got is None or float(got.abs().max()) == 0.0, n
            continue
        assert got is not None, n
        if n.endswith("adapter.4.bias"):
            ref_scale = float(ora["gparam"][n.replace(".bias", ".weight")].double().norm())
            errs["abs/" + n] = float((got.double() - gr.double()).norm()) / max(ref_scale, 1e-30)
            continue
        errs[n] = _rel(got, gr)
    return errs


def step_parity(cfg_kw, B, img_hw, seed, flag=1, sd_seed=5, n_boxes=None):
    """step_parity_inputs on one seeded synthetic batch (synth.synth_batch, SURVEY.md 8(d) box statistics)."""
    sd = synth.synth_state_dict(sd_seed)
    kw = {} if n_boxes is None else dict(n_boxes=n_boxes)
    bi, im, feats = synth.synth_batch(B, img_hw[0], img_hw[1], seed=seed, **kw)
    return step_parity_inputs(cfg_kw, sd, bi, im, feats, flag)


def step_parity_inputs(cfg_kw, sd, bi, im, feats, flag=1):
    """Engine vs oracle on given inputs. Returns a dict of measured numbers (no assertions):
    loss_err, fwd_err (worst level), masks_exact, flip_fraction (all sites), flips per site, flip_margin (largest
    |pre-activation| / rms at a flipped entry, from the oracle), grad_err_pattern (worst, oracle evaluated with the
    engine's activation pattern), grad_err_plain (worst, plain fp32 oracle), and the per-tensor tables."""
    eng = engine_step(cfg_kw, sd, bi, im, feats, flag)
    rec = {"record": {}, "record_x": {}}
    plain = _oracle_grads(sd, bi, im, feats, cfg_kw, flag, rec, eng["cot"])
    forced = _oracle_grads(sd, bi, im, feats, cfg_kw, flag, {"force": eng["patterns"]}, eng["cot"])
    keys = list(feats.keys())
    masks_exact = all(torch.equal(torch.cat(eng["masks"][l], 0), torch.cat(plain["masks"][l], 0)) for l in range(len(keys)))
    flips, total, per_site, margin = 0, 0, {}, 0.0
    for site, pe in eng["patterns"].items():
        po = rec["record"].get(site)
        if po is None:
            continue
        d = pe != po
        n = int(d.sum())
        per_site[site] = (n, d.numel())
        flips += n
        total += d.numel()
        if n:
            x = rec["record_x"][site]
            rms = float(x.double().pow(2).mean().sqrt())
            margin = max(margin, float(x[d].abs().max()) / max(rms, 1e-30))
    e_pat = _grad_errors(eng, forced)
    e_plain = _grad_errors(eng, plain)
    return dict(
        loss_err=abs(eng["loss"] - plain["loss"]) / abs(plain["loss"]),
        loss_err_pattern=abs(eng["loss"] - forced["loss"]) / abs(forced["loss"]),
        fwd_err=max(_rel(eng["tea"][k], plain["tea"][k]) for k in keys),
        masks_exact=masks_exact, flips=flips, activations=total, flip_fraction=flips / max(total, 1),
        flips_per_site=per_site, flip_margin=margin,
        grad_err_pattern=max(e_pat.values()), grad_err_pattern_worst=max(e_pat, key=e_pat.get),
        grad_err_plain=max(e_plain.values()), grad_err_plain_worst=max(e_plain, key=e_plain.get),
        table_pattern=e_pat, table_plain=e_plain)
