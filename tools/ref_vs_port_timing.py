#!/usr/bin/env python
"""Build-container check that the CPU baseline of bench.py (the oracle PORT) is not slower than the REAL reference:
times fwd+bwd of one distillation step of (a) the unmodified reference imported from /root/reference through
oracle/refshim.py and (b) oracle/lgd_oracle.py, same inputs, same weights, same thread count.
  python tools/ref_vs_port_timing.py [--batch 2] [--steps 3]  > profiles/r2_ref_vs_port_timing.txt"""
import argparse
import os
import sys
import time
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    import torch
    from lgd_b200 import synth
    from oracle import lgd_oracle as O
    from oracle import refshim
    torch.set_num_threads(os.cpu_count() or 1)
    kw = dict(add_context_box=True, interact_pattern="stuGuided")
    sd = synth.synth_state_dict(0)
    bi, im, feats = synth.synth_batch(args.batch, 800, 1333, seed=1234)
    hws = synth.pyramid_hw(800, 1344)
    cot = synth.synth_cotangents({k: torch.empty(args.batch, 256, h, w) for k, (h, w) in zip(feats, hws)})

    R = refshim.RefDistillator(synth.make_cfg(**kw))
    R.teacher.load_state_dict({k[len("teacher."):]: v for k, v in sd.items() if k.startswith("teacher.")})
    R.D.adapter.load_state_dict({k[len("adapter."):]: v for k, v in sd.items() if k.startswith("adapter.")})
    rparams = list(R.teacher.parameters()) + list(R.D.adapter.parameters())

    def ref_step():
        f = {k: v.detach().clone().requires_grad_(True) for k, v in feats.items()}
        tea, _, _, loss = R.step(bi, im, f, distill_flag=1)
        total = loss + sum((tea[k] * cot[k]).sum() for k in tea)
        torch.autograd.grad(total, list(f.values()) + rparams, allow_unused=True)
        return float(loss)

    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}

    def port_step():
        f = {k: v.detach().clone().requires_grad_(True) for k, v in feats.items()}
        tea, _, _, loss, _ = O.distill_step(params, bi, im, f, **kw)
        total = loss + sum((tea[k] * cot[k]).sum() for k in tea)
        torch.autograd.grad(total, list(f.values()) + list(params.values()), allow_unused=True)
        return float(loss)

    print("# fwd+bwd of one distillation step, B=%d, 800x1333->800x1344, ctx box, stuGuided, torch %s CPU fp32, %d threads"
          % (args.batch, torch.__version__, torch.get_num_threads()))
    for name, fn in (("reference (unmodified, via oracle/refshim.py)", ref_step), ("oracle port (oracle/lgd_oracle.py)", port_step)):
        l = fn()
        ts = []
        for _ in range(args.steps):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        print("%-48s loss %.6f  s/step %s  images/s %.3f" % (name, l, ["%.2f" % t for t in ts], args.batch / min(ts)))


if __name__ == "__main__":
    main()
