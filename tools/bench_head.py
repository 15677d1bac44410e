"""Measurement for SURVEY.md 8(f) rank 1: the student's detection head on the teacher pyramid, forward + backward, at the
BASELINE shape (B images of 800x1344, P3-P7) -- lgd_b200.heads.{RetinaNetHeadB200, FCOSHeadB200} against the same modules
run by stock PyTorch on the same GPU (cuDNN, TF32 allowed as in the reference's default; NCHW features).
usage: python tools/bench_head.py [--batch 16] [--steps 10]   -> one JSON line per head"""
import argparse
import json
import os
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lgd_b200 import engine, synth  # noqa: E402
from lgd_b200.heads import FCOSHeadB200, RetinaNetHeadB200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()
dev = torch.device("cuda", 0)
B = args.batch
hws = synth.pyramid_hw(800, 1344)
P = sum(h * w for h, w in hws)
g = engine.Geometry.get(B, hws, dev)
gen = torch.Generator().manual_seed(1)
pyr = torch.randn(g.elems, generator=gen).to(dev)
CONV = 2.0 * 256 * 2304 * P * B     # FLOP of one 3x3 256->256 convolution over the batch


def retina():
    def tower():
        return nn.Sequential(*[m for _ in range(4) for m in (nn.Conv2d(256, 256, 3, 1, 1), nn.ReLU())])
    head = nn.Module()
    head.cls_subnet, head.bbox_subnet = tower(), tower()
    head.cls_score, head.bbox_pred = nn.Conv2d(256, 720, 3, 1, 1), nn.Conv2d(256, 36, 3, 1, 1)
    head = head.to(dev)
    fast = RetinaNetHeadB200.from_module(head)

    def ref(feats):
        logits = [head.cls_score(head.cls_subnet(x)) for x in feats]
        deltas = [head.bbox_pred(head.bbox_subnet(x)) for x in feats]
        return logits, deltas
    return head, fast, ref, (8 + 720 / 256 + 36 / 256)


def fcos():
    class Scale(nn.Module):
        def __init__(self):
            super().__init__()
            self.scale = nn.Parameter(torch.ones(1))

        def forward(self, x):
            return x * self.scale

    def tower():
        return nn.Sequential(*[m for _ in range(4) for m in (nn.Conv2d(256, 256, 3, 1, 1), nn.GroupNorm(32, 256), nn.ReLU())])
    head = nn.Module()
    head.cls_subnet, head.bbox_subnet = tower(), tower()
    head.cls_score, head.bbox_pred, head.centerness = nn.Conv2d(256, 80, 3, 1, 1), nn.Conv2d(256, 4, 3, 1, 1), nn.Conv2d(256, 1, 3, 1, 1)
    head.scales = nn.ModuleList([Scale() for _ in hws])
    head.fpn_strides, head.centerness_on_reg, head.norm_reg_targets = [8, 16, 32, 64, 128], True, True
    head = head.to(dev)
    fast = FCOSHeadB200(head)

    def ref(feats):
        lo, bx, ct = [], [], []
        for l, x in enumerate(feats):
            c, b = head.cls_subnet(x), head.bbox_subnet(x)
            lo.append(head.cls_score(c))
            ct.append(head.centerness(b))
            bx.append(torch.relu(head.scales[l](head.bbox_pred(b))) * head.fpn_strides[l])
        return lo, bx, ct
    return head, fast, ref, (8 + 80 / 256 + 5 / 256)


def timed(fn, feats_of, head):
    ms = []
    for it in range(args.steps + 3):
        head.zero_grad(set_to_none=True)
        feats = feats_of()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        outs = fn(feats)
        loss = sum((o.float() ** 2).mean() for grp in outs for o in grp)
        loss.backward()
        e1.record()
        torch.cuda.synchronize()
        if it >= 3:
            ms.append(e0.elapsed_time(e1))
        del outs, loss, feats
    ms.sort()
    return ms[len(ms) // 2]


for name, make in (("RetinaNetHead", retina), ("FCOSHead", fcos)):
    head, fast, ref, convs_fwd = make()
    views = lambda: [v.detach().requires_grad_(True) for v in g.level_views(pyr)]                     # NHWC teacher pyramid
    nchw = lambda: [v.detach().contiguous().requires_grad_(True) for v in g.level_views(pyr)]        # what stock code gets
    t_fast = timed(fast, views, head)
    torch.cuda.empty_cache()
    t_ref = timed(ref, nchw, head)
    torch.cuda.empty_cache()
    flop = 3 * convs_fwd * CONV
    print(json.dumps({"head": name, "batch": B, "image_hw": [800, 1344], "step": "head fwd + mean-square loss + bwd (features and parameters)",
                      "lgd_b200_ms": t_fast, "torch_cudnn_ms": t_ref, "speedup": t_ref / t_fast, "conv_equivalents_fwd": convs_fwd,
                      "algorithmic_tflop": flop / 1e12, "lgd_b200_tflops": flop / (t_fast * 1e-3) / 1e12,
                      "torch_tflops": flop / (t_ref * 1e-3) / 1e12,
                      "note": "median of %d steps, CUDA events; torch = the same nn.Modules on NCHW copies of the features, cudnn "
                              "allow_tf32=%s" % (args.steps, torch.backends.cudnn.allow_tf32)}))
