"""Development tool: correctness and timing of the fp16-operand forward convolution against the TF32 one."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lgd_b200 import engine, synth  # noqa: E402
from lgd_b200._lib import call, ptr  # noqa: E402

dev = torch.device("cuda", 0)


def run(B, hws, check):
    g = engine.Geometry.get(B, hws, dev)
    gen = torch.Generator().manual_seed(3)
    xs = [torch.randn(B, 256, h, w, generator=gen).half().float() for h, w in hws]
    w = (torch.randn(256, 256, 3, 3, generator=gen) / 48).half().float()
    bias = torch.randn(256, generator=gen)
    x_buf = engine.to_pyramid(g, [x.cuda() for x in xs], False)
    x_half = x_buf.half()
    wc, bc = w.cuda(), bias.cuda()
    pw = torch.empty(9 * 256 * 256, device=dev, dtype=torch.float16)
    call("lgd_pack_conv_weight_f16", ptr(wc), ptr(pw))
    out = g.new()
    out_h = torch.empty(g.elems, device=dev, dtype=torch.float16)
    ts = torch.empty(g.num_tiles * 2, device=dev)
    call("lgd_conv3x3_fwd_f16", g.pref, ptr(x_half), ptr(pw), ptr(bc), 0, 0, ptr(out), ptr(out_h), 1, 0, ptr(ts))
    torch.cuda.synchronize()
    if check:
        for l, (x, v, vh) in enumerate(zip(xs, g.level_views(out), g.level_views(out_h.float()))):
            ref = F.conv2d(x.double(), w.double(), bias.double(), padding=1).relu()
            e = float((v.cpu().double() - ref).norm() / ref.norm())
            eh = float((vh.cpu().double() - ref).norm() / ref.norm())
            print("level %d: rel err fp32 out %.2e, fp16 copy %.2e" % (l, e, eh))
    else:
        pk = engine.PackedWeights().get(wc, 0)
        for name, fn in (("f16", lambda: call("lgd_conv3x3_fwd_f16", g.pref, ptr(x_half), ptr(pw), ptr(bc), 0, 0, ptr(out),
                                             ptr(out_h), 1, 0, ptr(ts))),
                         ("f16 no half copy", lambda: call("lgd_conv3x3_fwd_f16", g.pref, ptr(x_half), ptr(pw), ptr(bc), 0, 0,
                                                           ptr(out), None, 1, 0, ptr(ts))),
                         ("tf32", lambda: engine.conv3x3(g, x_buf, pk, bc, out=out, relu=True))):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print("%-18s %.3f ms  %.0f TFLOP/s" % (name, ms, 2 * 256 * 2304 * B * g.P / ms / 1e9))


run(2, [(20, 24), (9, 7), (3, 5), (1, 2)], True)
run(16, synth.pyramid_hw(800, 1344), False)
