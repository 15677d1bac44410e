#!/usr/bin/env python
"""Prints the gradient-parity table of oracle/parity.py (flip fraction, gradients against the oracle with the engine's
activation pattern, gradients against the plain fp32 oracle) for the BASELINE shapes, in the default fp16-operand mode
and in the split-operand tf32x3 mode. Development probe; the assertions live in tests/test_gpu_parity.py.

  python tools/grad_parity_probe.py [--batch 2] [--hw 800 1333] [--json out.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--hw", type=int, nargs=2, default=[800, 1333])
    ap.add_argument("--json", default=None)
    ap.add_argument("--modes", nargs="+", default=["fp16", "tf32x3"])
    args = ap.parse_args()
    from lgd_b200 import engine
    from oracle import parity
    out = {}
    for mode in args.modes:
        engine.FORWARD_PRECISION = mode
        for name, kw in (("retinanet_ctx", dict(add_context_box=True)), ("fcos_noctx", dict(add_context_box=False))):
            r = parity.step_parity(kw, args.batch, tuple(args.hw), seed=77)
            out[mode + "/" + name] = r
            print("== %s %s: loss %.2e fwd %.2e flips %d/%d = %.2e (margin %.2e) grad|pattern %.2e (%s) grad|plain %.2e (%s)"
                  % (mode, name, r["loss_err"], r["fwd_err"], r["flips"], r["activations"], r["flip_fraction"],
                     r["flip_margin"], r["grad_err_pattern"], r["grad_err_pattern_worst"], r["grad_err_plain"],
                     r["grad_err_plain_worst"]))
            print("   flips per site:", {k: v[0] for k, v in r["flips_per_site"].items() if v[0]})
            tp = sorted(r["table_pattern"].items(), key=lambda kv: -kv[1])
            print("   worst vs pattern-oracle:", [(k, "%.1e" % v) for k, v in tp[:8]])
            tq = sorted(r["table_plain"].items(), key=lambda kv: -kv[1])
            print("   worst vs plain oracle:  ", [(k, "%.1e" % v) for k, v in tq[:8]])
    if args.json:
        with open(args.json, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
