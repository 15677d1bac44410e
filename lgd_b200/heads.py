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
    """fp16 packings of every head weight for one step: the eight tower convolutions in one launch, the output
    convolutions as zero-padded 256-row chunks."""

    def __init__(self, g, P, tower_idx=TOWER, out_names=("cls_score", "bbox_pred")):
        dev = g.device
        n9 = 9 * C * C
        tower_w = [P["%s.%d.weight" % (t, i)].detach().contiguous() for t in ("cls_subnet", "bbox_subnet") for i in tower_idx]
        self.fwd = torch.empty(8, n9, device=dev, dtype=torch.float16)
        self.dgrad = torch.empty(8, n9, device=dev, dtype=torch.float16)
        self.gains = torch.empty(8, device=dev, dtype=torch.float32)
        ws = g.workspace()
        arr = lambda ts: (ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
        call("lgd_pack_conv_weights_f16_multi", arr(tower_w), 8, arr(list(self.fwd)), arr(list(self.dgrad)),
             ptr(self.gains), ptr(ws), ws.numel())
        self.out = {}
        for name in out_names:
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



# ======================================================================================================================
# FCOS family: FCOSHead (FCOS, ATSS) and POTOHead -- towers of [Conv2d(256,256,3), GroupNorm(32,256), ReLU] x 4
FCOS_TOWER = (0, 3, 6, 9)    # conv indices inside cls_subnet / bbox_subnet; the GroupNorm sits at index + 1


def fcos_param_names(centerness: bool):
    names = []
    for tower in ("cls_subnet", "bbox_subnet"):
        for i in FCOS_TOWER:
            names += ["%s.%d.weight" % (tower, i), "%s.%d.bias" % (tower, i),
                      "%s.%d.weight" % (tower, i + 1), "%s.%d.bias" % (tower, i + 1)]
    names += ["cls_score.weight", "cls_score.bias", "bbox_pred.weight", "bbox_pred.bias"]
    if centerness:
        names += ["centerness.weight", "centerness.bias"]
    return names


def _ld4(co):
    return (co + 3) // 4 * 4     # row pitch of an output matrix: the epilogue stores 16-byte pieces


class _FcosHeadFn(torch.autograd.Function):
    """(features...) -> per level the RAW outputs of cls_score, bbox_pred and (optionally) centerness as NCHW-shaped
    views of pixel-major matrices; the per-level Scale, relu * stride / exp stay ordinary torch ops on those small
    tensors (FCOSHeadB200.forward)."""

    @staticmethod
    def forward(ctx, n_lvl, ctr_tower, *tensors):
        feats, params = tensors[:n_lvl], tensors[n_lvl:]
        has_ctr = ctr_tower is not None
        names = fcos_param_names(has_ctr)
        P = dict(zip(names, params))
        dev = feats[0].device
        B = feats[0].shape[0]
        g = engine.Geometry.get(B, [tuple(f.shape[-2:]) for f in feats], dev)
        outs_of = {"cls_subnet": ["cls_score"], "bbox_subnet": ["bbox_pred"]}
        if has_ctr:
            outs_of[ctr_tower].append("centerness")
        with torch.cuda.device(dev):
            _, x_h = engine.to_pyramid(g, feats, False, want_half=True, want_fp32=False)
            # the output convolutions of a tower run as ONE convolution (their weights stacked along the output channels:
            # 80 [+1] and 4 [+1] columns of a single 256-column launch)
            Pp = dict(P)
            for tower in ("cls_subnet", "bbox_subnet"):
                Pp["out_" + tower + ".weight"] = torch.cat([P[h + ".weight"].detach() for h in outs_of[tower]], 0)
                Pp["out_" + tower + ".bias"] = torch.cat([P[h + ".bias"].detach() for h in outs_of[tower]], 0)
            pk = _Packed(g, Pp, FCOS_TOWER, ["out_cls_subnet", "out_bbox_subnet"])
            ws = torch.empty(query("lgd_gn32_workspace", g.pref), device=dev, dtype=torch.uint8)
            nseg = g.F * B
            S = SimpleNamespace(g=g, pk=pk, P=P, names=names, outs_of=outs_of, layers={}, ws=ws)
            outs = {}
            for ti, tower in enumerate(("cls_subnet", "bbox_subnet")):
                t = x_h
                layers = []
                for k, i in enumerate(FCOS_TOWER):
                    # conv (+bias) -> fp32 x (GroupNorm needs it, forward and backward) -> statistics -> affine + ReLU -> fp16
                    x32 = g.new()
                    call("lgd_conv3x3_fwd_f16", g.pref, ptr(t), ptr(pk.fwd[4 * ti + k]), ptr(P["%s.%d.bias" % (tower, i)]),
                         0, 0, ptr(x32), None, 0, 0, None)
                    stats = torch.empty(nseg * 64, device=dev, dtype=torch.float32)
                    chsum = torch.empty(nseg * C, device=dev, dtype=torch.float32)
                    call("lgd_gn32_stats", g.pref, ptr(x32), ptr(stats), ptr(chsum), ptr(ws), ws.numel())
                    y_h = g.new_half()
                    call("lgd_gn32_apply", g.pref, ptr(x32), ptr(stats), ptr(P["%s.%d.weight" % (tower, i + 1)]),
                         ptr(P["%s.%d.bias" % (tower, i + 1)]), 1, ptr(y_h), None)
                    layers.append(SimpleNamespace(inp=t, x32=x32, stats=stats, chsum=chsum, out=y_h))
                    t = y_h
                S.layers[tower] = layers
                o = pk.out["out_" + tower]
                ld = _ld4(o.co)
                out = torch.empty(B * g.P, ld, device=dev, dtype=torch.float32)
                for j in range(o.nch):
                    # whole 16-byte pieces: the pad columns up to ld receive the zero rows of the packed weights
                    call("lgd_conv3x3_fwd_f16_cols", g.pref, ptr(t), ptr(o.fwd[j]), ptr(o.bias[j]), ptr(out), ld, j * C,
                         min(C, ld - j * C), 0)
                col = 0
                for head in outs_of[tower]:
                    co = P[head + ".weight"].shape[0]
                    outs[head] = (out, col, co)
                    col += co
        ctx.S = S
        ctx.needs_x = [f.requires_grad for f in feats]
        res = []
        order = ["cls_score", "bbox_pred"] + (["centerness"] if has_ctr else [])
        for head in order:
            out, col, co = outs[head]
            row = 0
            for (h, w) in g.hws:
                n = B * h * w
                # (N, co, H, W)-shaped view of the pixel-major rows: the values of the reference's NCHW tensor
                res.append(out[row:row + n].view(B, h, w, -1)[..., col:col + co].permute(0, 3, 1, 2))
                row += n
        ctx.order = order
        return tuple(res)

    @staticmethod
    def backward(ctx, *gouts):
        S = ctx.S
        g, pk, P = S.g, S.pk, S.P
        dev = g.device
        F_, B = g.F, g.B
        grads = {}
        with torch.cuda.device(dev):
            wstream = engine.WgradStream(g)
            conv_ws = g.workspace(max(query("lgd_head_grad_workspace", g.pref, pk.out["out_cls_subnet"].co), g.ws_bytes))
            gout_of = {head: list(gouts[k * F_:(k + 1) * F_]) for k, head in enumerate(ctx.order)}
            d_x = None
            stacked = []
            for ti, tower in enumerate(("cls_subnet", "bbox_subnet")):
                layers = S.layers[tower]
                # gradient of the tower's stacked output convolution: per level (B, h*w*co) rows, channels of its heads
                # side by side
                o = pk.out["out_" + tower]
                heads = S.outs_of[tower]
                cos = [P[h + ".weight"].shape[0] for h in heads]
                srcs, strides = [], []
                for l, (h, w) in enumerate(g.hws):
                    parts = []
                    for head, co in zip(heads, cos):
                        go = gout_of[head][l]
                        if go is None:
                            go = torch.zeros(B, co, h, w, device=dev, dtype=torch.float32)
                        parts.append(go.detach().float())
                    go = parts[0] if len(parts) == 1 else torch.cat(parts, 1)
                    go = go.permute(0, 2, 3, 1).reshape(B, h * w * o.co).contiguous()
                    srcs.append(go)
                    strides.append(go.stride(0))
                gh = torch.empty(o.nch, g.elems, device=dev, dtype=torch.float16)
                sc = torch.empty(3, device=dev, dtype=torch.float32)
                gb = torch.empty(o.co, device=dev, dtype=torch.float32)
                call("lgd_head_grad_prepare", g.pref, (ctypes.c_void_p * F_)(*[t.data_ptr() for t in srcs]),
                     (ctypes.c_int64 * F_)(*strides), o.co, ptr(gh), ptr(sc), ptr(gb), ptr(conv_ws), conv_ws.numel())
                gw = torch.empty(o.co, C, 3, 3, device=dev, dtype=torch.float32)
                acc = None       # gradient w.r.t. the tower's last activation (fp32, above its ReLU)
                for j in range(o.nch):
                    wstream.wgrad_rows(layers[3].out, gh[j], sc, gw, j * C, min(C, o.co - j * C))
                    if acc is None:
                        acc = g.new()
                        call("lgd_conv3x3_dgrad_f16", g.pref, ptr(gh[j]), ptr(o.dgrad[j]), ptr(sc[1:]), ptr(acc), 0, None,
                             None, None, None, None, None, None, None, 0)
                    else:
                        call("lgd_conv3x3_dgrad_f16_addend", g.pref, ptr(gh[j]), ptr(o.dgrad[j]), ptr(sc[1:]), ptr(acc),
                             ptr(acc), None, None, None, None, None, None, None, 0)
                wstream.keep += srcs
                stacked.append((heads, cos, gw, gb))
                for k in (3, 2, 1, 0):
                    L = layers[k]
                    conv, norm = "%s.%d" % (tower, FCOS_TOWER[k]), "%s.%d" % (tower, FCOS_TOWER[k] + 1)
                    gx_h = g.new_half()
                    sc = torch.empty(3, device=dev, dtype=torch.float32)
                    dg, db, dbias = (torch.empty(C, device=dev, dtype=torch.float32) for _ in range(3))
                    call("lgd_gn32_bwd", g.pref, ptr(acc), ptr(L.x32), ptr(L.stats), ptr(L.chsum), ptr(P[norm + ".weight"]),
                         ptr(P[norm + ".bias"]), 1, ptr(gx_h), ptr(sc), None, ptr(dg), ptr(db), ptr(dbias), ptr(S.ws),
                         S.ws.numel())
                    grads[norm + ".weight"], grads[norm + ".bias"], grads[conv + ".bias"] = dg, db, dbias
                    grads[conv + ".weight"] = wstream.wgrad(None, None, P[conv + ".weight"].shape, L.inp, (gx_h, sc))
                    w_d = pk.dgrad[4 * ti + k]
                    if k > 0:
                        acc = g.new()
                        call("lgd_conv3x3_dgrad_f16", g.pref, ptr(gx_h), ptr(w_d), ptr(sc[1:]), ptr(acc), 0, None, None, None,
                             None, None, None, None, None, 0)
                    elif any(ctx.needs_x):
                        if d_x is None:
                            d_x = g.new()
                            call("lgd_conv3x3_dgrad_f16", g.pref, ptr(gx_h), ptr(w_d), ptr(sc[1:]), ptr(d_x), 0, None, None,
                                 None, None, None, None, None, None, 0)
                        else:
                            call("lgd_conv3x3_dgrad_f16_addend", g.pref, ptr(gx_h), ptr(w_d), ptr(sc[1:]), ptr(d_x), ptr(d_x),
                                 None, None, None, None, None, None, None, 0)
            wstream.join()
            for heads, cos, gw, gb in stacked:    # rows of the stacked gradient -> the heads' own parameters
                row = 0
                for head, co in zip(heads, cos):
                    grads[head + ".weight"] = gw[row:row + co]
                    grads[head + ".bias"] = gb[row:row + co]
                    row += co
        gx = [None] * F_
        if d_x is not None:
            gx = [v if need else None for v, need in zip(g.level_views(d_x), ctx.needs_x)]
        return (None, None, *gx, *[grads.get(n) for n in S.names])


class FCOSHeadB200(nn.Module):
    """Drop-in for the forward of the reference's FCOSHead (thirdparty_heads/fcos.py:433-546; FCOS and ATSS) and POTOHead
    (poto.py:523-625): `forward(features)` returns (logits, bbox_reg, centerness) -- or (logits, bbox_reg) for a head
    without a centerness convolution -- as lists of NCHW-shaped tensors with the reference's values. The 8 tower
    convolutions and the output convolutions run on the tcgen05 kernel from the NHWC pyramid, GroupNorm(32)+ReLU on
    the library's own kernels (lgd_gn32_*); the module only REFERENCES the parameters of the wrapped head."""

    def __init__(self, head):
        super().__init__()
        self.head = head
        for tower in (head.cls_subnet, head.bbox_subnet):
            if len(tower) != 12:
                raise ValueError("FCOSHeadB200 needs the stock tower: 4 x [conv3x3(256,256), GroupNorm(32,256), ReLU]")
            for i in FCOS_TOWER:
                conv, norm = tower[i], tower[i + 1]
                if not isinstance(conv, nn.Conv2d) or tuple(conv.weight.shape) != (C, C, 3, 3) or conv.bias is None:
                    raise ValueError("unsupported tower convolution")
                if not isinstance(norm, nn.GroupNorm) or norm.num_groups != 32 or norm.num_channels != C \
                        or not norm.affine or abs(norm.eps - 1e-5) > 1e-12:
                    raise ValueError("unsupported tower norm")
        self.has_ctr = hasattr(head, "centerness")
        for m in [head.cls_score, head.bbox_pred] + ([head.centerness] if self.has_ctr else []):
            if tuple(m.weight.shape[1:]) != (C, 3, 3) or m.weight.shape[0] > 4 * C:
                raise ValueError("unsupported output convolution %s" % (tuple(m.weight.shape),))

    @classmethod
    def supports(cls, head) -> bool:
        try:
            cls(head)
            return True
        except Exception:  # noqa: BLE001
            return False

    def forward(self, features: Sequence[torch.Tensor]):
        if features[0].device.type != "cuda":
            raise RuntimeError("lgd_b200.FCOSHeadB200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        head = self.head
        own = dict(head.named_parameters())
        params = [own[n] for n in fcos_param_names(self.has_ctr)]
        ctr_tower = None
        if self.has_ctr:
            ctr_tower = "bbox_subnet" if head.centerness_on_reg else "cls_subnet"
        outs = _FcosHeadFn.apply(len(features), ctr_tower, *features, *params)
        n = len(features)
        logits, raw_box = list(outs[:n]), list(outs[n:2 * n])
        bbox_reg = []
        for level, pred in enumerate(raw_box):                      # fcos.py:538-542, ordinary autograd ops
            pred = head.scales[level](pred)
            if head.norm_reg_targets:
                bbox_reg.append(torch.nn.functional.relu(pred) * head.fpn_strides[level])
            else:
                bbox_reg.append(torch.exp(pred))
        if self.has_ctr:
            return logits, bbox_reg, list(outs[2 * n:])
        return logits, bbox_reg
