"""The student's detection head applied to the teacher pyramid, on the hot path's own convolution kernel
(reference: distillator.py:107-112 -> customized_detectors/retinanet.py:36-45 ->
detectron2 0.3 `RetinaNetHead`, which is not vendored in the reference: two towers of four
conv3x3(256,256)+ReLU shared by all levels, then conv3x3(256, A*K) (cls_score) and conv3x3(256, A*4) (bbox_pred)).

It is the step right after the teacher pyramid and the source of the teacher's cotangents. Here it reads the NHWC
teacher pyramid as it is (one fp16 cast, no NCHW round trip), runs its 8 + 4 convolution launches on the tcgen05
kernel -- the 720 / 36 output channels are written by 256-column launches straight into pixel-major (B*P, A*K)
matrices, which per level ARE the (N, H*W*A, K) tensors the losses consume (permute_to_N_HWA_K, retinanet.py:13-22) --
and its backward hands the gradient w.r.t. the teacher pyramid back as level views of ONE NHWC buffer, which
lgd_teacher_backward reads in place.

    head = RetinaNetHeadB200.from_module(student.head)          # wraps the parameters of a detectron2 RetinaNetHead
    logits, deltas = head([features_tea[f] for f in student.head_in_features])
    # == [permute_to_N_HWA_K(x, K) for x in student.head(features)[0]], [... for x in ...[1]]

No CPU fallback: CUDA tensors only."""
from __future__ import annotations

import ctypes
from types import SimpleNamespace
from typing import List, Sequence

import torch
import torch.nn as nn

from . import engine
from ._lib import call, ptr, query

C = 256
TOWER = (0, 2, 4, 6)     # conv indices inside detectron2's cls_subnet / bbox_subnet Sequential (ReLU in between)


def _param_names():
    names = []
    for tower in ("cls_subnet", "bbox_subnet"):
        for i in TOWER:
            names += ["%s.%d.weight" % (tower, i), "%s.%d.bias" % (tower, i)]
    return names + ["cls_score.weight", "cls_score.bias", "bbox_pred.weight", "bbox_pred.bias"]


PARAM_NAMES = _param_names()


class _Packed:
    """fp16 packings of every head weight for one step: the eight tower convolutions in one launch, the two output
    convolutions as zero-padded 256-row chunks."""

    def __init__(self, g, P):
        dev = g.device
        n9 = 9 * C * C
        tower_w = [P["%s.%d.weight" % (t, i)].detach().contiguous() for t in ("cls_subnet", "bbox_subnet") for i in TOWER]
        self.fwd = torch.empty(8, n9, device=dev, dtype=torch.float16)
        self.dgrad = torch.empty(8, n9, device=dev, dtype=torch.float16)
        self.gains = torch.empty(8, device=dev, dtype=torch.float32)
        ws = g.workspace()
        arr = lambda ts: (ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
        call("lgd_pack_conv_weights_f16_multi", arr(tower_w), 8, arr(list(self.fwd)), arr(list(self.dgrad)),
             ptr(self.gains), ptr(ws), ws.numel())
        self.out = {}
        for name in ("cls_score", "bbox_pred"):
            w = P[name + ".weight"].detach().contiguous()
            b = P[name + ".bias"].detach().contiguous()
            co = w.shape[0]
            nch = (co + C - 1) // C
            o = SimpleNamespace(co=co, nch=nch, fwd=torch.empty(nch, n9, device=dev, dtype=torch.float16),
                                dgrad=torch.empty(nch, n9, device=dev, dtype=torch.float16),
                                bias=torch.empty(nch, C, device=dev, dtype=torch.float32),
                                gains=torch.empty(nch, device=dev, dtype=torch.float32))
            for j in range(nch):
                call("lgd_pack_conv_weight_f16_rows", ptr(w), ptr(b), co, j * C, ptr(o.fwd[j]), ptr(o.dgrad[j]),
                     ptr(o.bias[j]), ptr(o.gains[j:]), ptr(ws), ws.numel())
            self.out[name] = o


def _conv_relu_half(g, x_h, w_h, bias):
    out_h = g.new_half()
    call("lgd_conv3x3_fwd_f16", g.pref, ptr(x_h), ptr(w_h), ptr(bias), 0, 0, None, ptr(out_h), 1, 1, None)
    return out_h


class _RetinaHeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, n_lvl, *tensors):
        feats, params = tensors[:n_lvl], tensors[n_lvl:]
        P = dict(zip(PARAM_NAMES, params))
        dev = feats[0].device
        B = feats[0].shape[0]
        g = engine.Geometry.get(B, [tuple(f.shape[-2:]) for f in feats], dev)
        with torch.cuda.device(dev):
            _, x_h = engine.to_pyramid(g, feats, False, want_half=True, want_fp32=False)
            pk = _Packed(g, P)
            S = SimpleNamespace(g=g, pk=pk, x_h=x_h, acts={}, P=P)
            outs = {}
            for ti, (tower, head) in enumerate((("cls_subnet", "cls_score"), ("bbox_subnet", "bbox_pred"))):
                t = x_h
                acts = [t]
                for k, i in enumerate(TOWER):
                    t = _conv_relu_half(g, t, pk.fwd[4 * ti + k], P["%s.%d.bias" % (tower, i)])
                    acts.append(t)
                S.acts[tower] = acts
                o = pk.out[head]
                out = torch.empty(B * g.P, o.co, device=dev, dtype=torch.float32)
                for j in range(o.nch):
                    call("lgd_conv3x3_fwd_f16_cols", g.pref, ptr(t), ptr(o.fwd[j]), ptr(o.bias[j]), ptr(out), o.co, j * C,
                         min(C, o.co - j * C), 0)
                outs[head] = out
        ctx.S = S
        ctx.needs_x = [f.requires_grad for f in feats]
        A_K, A_4 = pk.out["cls_score"].co, pk.out["bbox_pred"].co
        A = A_4 // 4
        K = A_K // A
        res = []
        for out, last in ((outs["cls_score"], K), (outs["bbox_pred"], 4)):
            row = 0
            for (h, w) in g.hws:
                n = B * h * w
                res.append(out[row:row + n].view(B, h * w * A, last))    # == permute_to_N_HWA_K of the NCHW head output
                row += n
        return tuple(res)

    @staticmethod
    def backward(ctx, *gouts):
        S = ctx.S
        g, pk, P = S.g, S.pk, S.P
        dev = g.device
        F = g.F
        B = g.B
        grads = {}
        with torch.cuda.device(dev):
            wstream = engine.WgradStream(g)
            ws = g.workspace(max(query("lgd_head_grad_workspace", g.pref, pk.out["cls_score"].co), g.ws_bytes))
            d_x = None
            for ti, (tower, head, K_last) in enumerate((("cls_subnet", "cls_score", None), ("bbox_subnet", "bbox_pred", 4))):
                o = pk.out[head]
                gl = list(gouts[ti * F:(ti + 1) * F])
                srcs, strides = [], []
                for go, (h, w) in zip(gl, g.hws):
                    if go is None:
                        go = torch.zeros(B, h * w * o.co, device=dev, dtype=torch.float32)
                    go = go.detach()
                    if go.dtype != torch.float32:
                        go = go.float()
                    # the rows of one image, (h*w*A*K) floats, must be contiguous; the batch stride is free (the loss
                    # concatenates the levels per image, so a level's gradient is a strided slice of a bigger tensor)
                    if go.dim() == 3 and not (go.stride(2) == 1 and go.stride(1) == go.shape[2]):
                        go = go.contiguous()
                    if go.dim() == 2 and go.stride(1) != 1:
                        go = go.contiguous()
                    srcs.append(go)
                    strides.append(go.stride(0))
                gh = torch.empty(o.nch, g.elems, device=dev, dtype=torch.float16)
                sc = torch.empty(3, device=dev, dtype=torch.float32)
                gb = torch.empty(o.co, device=dev, dtype=torch.float32)
                call("lgd_head_grad_prepare", g.pref, (ctypes.c_void_p * F)(*[t.data_ptr() for t in srcs]),
                     (ctypes.c_int64 * F)(*strides), o.co, ptr(gh), ptr(sc), ptr(gb), ptr(ws), ws.numel())
                grads[head + ".bias"] = gb
                acts = S.acts[tower]
                # output convolution: weight gradient chunk by chunk, input gradient = sum of the chunks' dgrads
                gw = torch.empty_like(P[head + ".weight"], memory_format=torch.contiguous_format)
                for j in range(o.nch):
                    wstream.wgrad_rows(acts[4], gh[j], sc, gw, j * C, min(C, o.co - j * C))
                grads[head + ".weight"] = gw
                gain_sum = o.gains.sum().reshape(1)                       # ||dX|| <= (sum_j gain_j) * ||g||
                out_h = g.new_half()
                sc_out = torch.empty(3, device=dev, dtype=torch.float32)
                tile_stats = torch.empty(g.num_tiles * 2, device=dev, dtype=torch.float32)
                sums = torch.empty(g.F * g.B * C, device=dev, dtype=torch.float32)
                total = torch.empty(C, device=dev, dtype=torch.float32)
                call("lgd_grad_scale", None, 0, 1, ptr(gain_sum), ptr(sc[2:]), 1.0, ptr(sc_out))
                acc = None
                for j in range(o.nch):
                    last = j == o.nch - 1
                    if last:   # masked by the tower's last ReLU; channel sums = bias gradient of that tower convolution
                        if acc is None:
                            call("lgd_conv3x3_dgrad_f16", g.pref, ptr(gh[j]), ptr(o.dgrad[j]), ptr(sc[1:]), None, 0, None,
                                 ptr(acts[4]), ptr(out_h), ptr(sc_out), ptr(tile_stats), ptr(sums), ptr(total), ptr(ws),
                                 ws.numel())
                        else:
                            call("lgd_conv3x3_dgrad_f16_addend", g.pref, ptr(gh[j]), ptr(o.dgrad[j]), ptr(sc[1:]), ptr(acc),
                                 None, ptr(acts[4]), ptr(out_h), ptr(sc_out), ptr(tile_stats), ptr(sums), ptr(total),
                                 ptr(ws), ws.numel())
                    elif acc is None:
                        acc = g.new()
                        call("lgd_conv3x3_dgrad_f16", g.pref, ptr(gh[j]), ptr(o.dgrad[j]), ptr(sc[1:]), ptr(acc), 0, None,
                             None, None, None, None, None, None, None, 0)
                    else:
                        call("lgd_conv3x3_dgrad_f16_addend", g.pref, ptr(gh[j]), ptr(o.dgrad[j]), ptr(sc[1:]), ptr(acc),
                             ptr(acc), None, None, None, None, None, None, None, 0)
                meas = torch.empty(3, device=dev, dtype=torch.float32)
                call("lgd_grad_scale", ptr(tile_stats[1:]), g.num_tiles, 2, None, None, 1.0, ptr(meas))
                operand = (out_h, torch.cat([sc_out[:2], meas[2:]]))
                gb_prev = total
                # tower, last convolution first
                for k in (3, 2, 1, 0):
                    name = "%s.%d" % (tower, TOWER[k])
                    grads[name + ".bias"] = gb_prev
                    grads[name + ".weight"] = wstream.wgrad(None, None, P[name + ".weight"].shape, acts[k], operand)
                    w_d, gain = pk.dgrad[4 * ti + k], pk.gains[4 * ti + k:]
                    if k > 0:
                        nh = g.new_half()
                        sc2 = torch.empty(3, device=dev, dtype=torch.float32)
                        ts2 = torch.empty(g.num_tiles * 2, device=dev, dtype=torch.float32)
                        sums = torch.empty(g.F * g.B * C, device=dev, dtype=torch.float32)
                        total = torch.empty(C, device=dev, dtype=torch.float32)
                        call("lgd_grad_scale", None, 0, 1, ptr(gain), ptr(operand[1][2:]), 1.0, ptr(sc2))
                        call("lgd_conv3x3_dgrad_f16", g.pref, ptr(operand[0]), ptr(w_d), ptr(operand[1][1:]), None, 0, None,
                             ptr(acts[k]), ptr(nh), ptr(sc2), ptr(ts2), ptr(sums), ptr(total), ptr(ws), ws.numel())
                        meas = torch.empty(3, device=dev, dtype=torch.float32)
                        call("lgd_grad_scale", ptr(ts2[1:]), g.num_tiles, 2, None, None, 1.0, ptr(meas))
                        operand = (nh, torch.cat([sc2[:2], meas[2:]]))
                        gb_prev = total
                    elif any(ctx.needs_x):
                        # gradient w.r.t. the input pyramid: the two towers' contributions are summed in the epilogue
                        if d_x is None:
                            d_x = g.new()
                            call("lgd_conv3x3_dgrad_f16", g.pref, ptr(operand[0]), ptr(w_d), ptr(operand[1][1:]), ptr(d_x),
                                 0, None, None, None, None, None, None, None, None, 0)
                        else:
                            call("lgd_conv3x3_dgrad_f16_addend", g.pref, ptr(operand[0]), ptr(w_d), ptr(operand[1][1:]),
                                 ptr(d_x), ptr(d_x), None, None, None, None, None, None, None, 0)
            wstream.join()
        gx = [None] * F
        if d_x is not None:
            gx = [v if need else None for v, need in zip(g.level_views(d_x), ctx.needs_x)]
        return (None, *gx, *[grads.get(n) for n in PARAM_NAMES])


class RetinaNetHeadB200(nn.Module):
    """Drop-in for the forward of detectron2's RetinaNetHead + permute_to_N_HWA_K: `forward(features)` returns
    (pred_logits, pred_anchor_deltas) as lists of (N, Hi*Wi*A, K) / (N, Hi*Wi*A, 4) tensors. The module only REFERENCES
    the parameters of the wrapped head (same tensors: the student's optimizer, DDP and checkpoints see nothing new)."""

    def __init__(self, cls_subnet, bbox_subnet, cls_score, bbox_pred):
        super().__init__()
        self.cls_subnet, self.bbox_subnet, self.cls_score, self.bbox_pred = cls_subnet, bbox_subnet, cls_score, bbox_pred
        for tower in (cls_subnet, bbox_subnet):
            convs = [m for m in tower if isinstance(m, nn.Conv2d)]
            if len(tower) != 8 or len(convs) != 4 or any(tuple(c.weight.shape) != (C, C, 3, 3) for c in convs):
                raise ValueError("RetinaNetHeadB200 needs the stock tower: 4 x [conv3x3(256,256), ReLU], no norm layers")
        for m in (cls_score, bbox_pred):
            if tuple(m.weight.shape[1:]) != (C, 3, 3) or m.weight.shape[0] % 4 != 0 or m.weight.shape[0] > 4 * C:
                raise ValueError("unsupported output convolution %s" % (tuple(m.weight.shape),))

    @classmethod
    def from_module(cls, head):
        return cls(head.cls_subnet, head.bbox_subnet, head.cls_score, head.bbox_pred)

    @classmethod
    def supports(cls, head) -> bool:
        try:
            cls.from_module(head)
            return True
        except Exception:  # noqa: BLE001
            return False

    def forward(self, features: Sequence[torch.Tensor]):
        if features[0].device.type != "cuda":
            raise RuntimeError("lgd_b200.RetinaNetHeadB200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        own = dict(self.named_parameters())
        params = [own[n] for n in PARAM_NAMES]
        outs = _RetinaHeadFn.apply(len(features), *features, *params)
        n = len(features)
        return list(outs[:n]), list(outs[n:])
