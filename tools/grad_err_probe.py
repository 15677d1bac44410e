"""Development tool: per-parameter gradient error of the engine against the fp32 CPU oracle on a golden case.
usage: python tools/grad_err_probe.py [case] [fp16|tf32x3]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lgd_b200 import engine, synth
from oracle import lgd_oracle as O
from tests.golden_util import load_case, rel_l2
from tests.gpu_util import run_engine

name = sys.argv[1] if len(sys.argv) > 1 else "ctx_stu_adv"
engine.FORWARD_PRECISION = sys.argv[2] if len(sys.argv) > 2 else "tf32x3"
g, cfg_kw, batch_kw, flag, sd, bi, im, feats = load_case(name)
out = run_engine(cfg_kw, sd, bi, im, feats, flag)
f = {k: v.clone().requires_grad_(True) for k, v in feats.items()}
sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
tea, _, _, loss, _ = O.distill_step(sdo, bi, im, f, distill_flag=flag, tf32=False, **cfg_kw)
cot = synth.synth_cotangents(tea)
total = loss + sum((tea[k] * cot[k]).sum() for k in tea)
names = list(sdo)
grads = torch.autograd.grad(total, list(f.values()) + [sdo[n] for n in names], allow_unused=True)
print("loss", out["loss"], float(loss))
for l, k in enumerate(f):
    if grads[l] is not None:
        print("%-50s %.3e" % ("gfeat_" + k, rel_l2(out["gfeat"][k], grads[l])))
for n, gr in zip(names, grads[len(f):]):
    got = out["gparam"][n]
    if gr is None or got is None:
        print("%-50s none" % n)
        continue
    print("%-50s %.3e   |ref| %.3e" % (n, rel_l2(got, gr), float(gr.norm())))
