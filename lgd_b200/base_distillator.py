"""BaseDistillator: drop-in for models/base_distillator.py:11-77 (same attributes, same distill()/distill_loss()
signatures). distill() = detach rules -> adapter -> InstanceNorm both sides -> lambda * MSE over all levels, as ONE
autograd node running liblgd_b200 kernels over the whole pyramid."""
from abc import abstractmethod

import torch
import torch.nn as nn

from . import engine
from .adapters import SequentialConvs, build_adapter
from .customized_detectors import build_customized_detector


def _cached(mod, g, stu, tea):
    """Reuse what the teacher forward of THIS step already produced: the transposed + TF32-rounded student pyramid
    and (zero copy) the teacher pyramid buffer its outputs are views of. The cache holds references to the feature
    tensors, so a (data_ptr, version) match cannot be a recycled allocation."""
    stu_pyr = tea_pyr = tea_stats = None
    cache = getattr(getattr(mod, "teacher", None), "_step_cache", None)
    if cache is not None and cache["g"] is g:
        if cache["key"] == tuple((s.data_ptr(), s._version) for s in stu) \
                and cache.get("stu_h") is not None and cache["stu_h"].dtype == engine.companion_dtype() \
                and (cache["stu"] is not None or engine._bwd_f16()):
            stu_pyr = (cache["stu"], cache["stu_h"])
        if engine._is_pyramid_view(g, tea) == cache["tea"].data_ptr():
            tea_pyr, tea_stats = cache["tea"], cache.get("tea_stats")
    return stu_pyr, tea_pyr, tea_stats


_QUIET = [False]


def _quiet_cross_stream_accumulate():
    """The adapter/loss node runs on its own stream on purpose; autograd's once-per-process warning about an
    AccumulateGrad node on another stream (it inserts the synchronisation itself) is noise here."""
    if not _QUIET[0]:
        _QUIET[0] = True
        fn = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
        if fn is not None:
            fn(False)


class _DistillFn(torch.autograd.Function):
    """Fused path for the stock SequentialConvs adapter: adapter convs + InstanceNorm + MSE over the whole pyramid."""

    @staticmethod
    def forward(ctx, mod, coef, names, n_lvl, tea_ready, *tensors):
        stu, tea, params = tensors[:n_lvl], tensors[n_lvl:2 * n_lvl], tensors[2 * n_lvl:]
        P = {"adapter.distill." + n: p for n, p in zip(names, params)}
        g = engine.Geometry.get(stu[0].shape[0], [tuple(s.shape[-2:]) for s in stu], stu[0].device)
        packed = mod._packed
        packed.new_step()
        with torch.cuda.device(g.device):
            stu_pyr, tea_pyr, tea_stats = _cached(mod, g, stu, tea)
            if getattr(mod, "teacher", None) is not None:
                mod.teacher._step_cache = None   # consumed: do not keep the buffers alive beyond the tape
            if stu_pyr is None:
                stu_pyr = engine.student_operands(g, stu)
            if tea_pyr is None:
                tea_pyr = engine.to_pyramid(g, tea, False)
            if tea_ready is not None:   # on the adapter stream: tensors the other stream allocated stay ours until we are done
                cur = torch.cuda.current_stream(g.device)
                for t in (stu_pyr[0], stu_pyr[1], tea_pyr):
                    if t is not None:
                        t.record_stream(cur)
            if engine.chain_applicable("stuGuided") and stu_pyr[1].dtype == torch.float16:
                loss, S = engine.chain_distill_forward(P, stu_pyr[1], tea_pyr, g, coef, tea_ready=tea_ready,
                                                       nhwc=all(engine.memory_layout(x) == "nhwc" for x in stu))
            else:
                loss, S = engine.distill_forward(P, stu_pyr[0], stu_pyr[1], tea_pyr, g, coef, packed,
                                                 tea_stats=tea_stats, tea_ready=tea_ready)
        ctx.S, ctx.P, ctx.names, ctx.n_lvl, ctx.mod = S, P, names, n_lvl, mod
        ctx.stu_needs = [s.requires_grad for s in stu]
        return loss.reshape(())

    @staticmethod
    def backward(ctx, gloss):
        need = any(ctx.stu_needs)
        with torch.cuda.device(ctx.S.g.device):
            gstu = [None] * ctx.n_lvl
            if getattr(ctx.S, "chain", False):
                grads, outs = engine.chain_distill_backward(ctx.S, gloss, need)
                if outs is not None:
                    gstu = [o if n else None for o, n in zip(outs, ctx.stu_needs)]
            else:
                grads, g_stu = engine.distill_backward(ctx.P, ctx.S, gloss, ctx.mod._packed, need)
                if g_stu is not None:
                    outs = engine.from_pyramid_nchw(ctx.S.g, g_stu)
                    gstu = [o if n else None for o, n in zip(outs, ctx.stu_needs)]
        gparams = [grads.get("adapter.distill." + n) for n in ctx.names]
        return (None, None, None, None, None, *gstu, *([None] * ctx.n_lvl), *gparams)


class _InMseFn(torch.autograd.Function):
    """Generic path for user-registered adapters (the adapters/ hook API): the adapter runs as an ordinary module,
    the InstanceNorm + MSE reduction runs here."""

    @staticmethod
    def forward(ctx, mod, coef, n_lvl, *tensors):
        s, tea = tensors[:n_lvl], tensors[n_lvl:]
        g = engine.Geometry.get(s[0].shape[0], [tuple(x.shape[-2:]) for x in s], s[0].device)
        with torch.cuda.device(g.device):
            _, tea_pyr, tea_stats = _cached(mod, g, s, tea)
            if tea_pyr is None:
                tea_pyr = engine.to_pyramid(g, tea, False)
            s_pyr = engine.to_pyramid(g, s, False)
            loss, S = engine.in_mse_forward(g, s_pyr, tea_pyr, coef, tea_stats)
        ctx.S, ctx.n_lvl = S, n_lvl
        return loss.reshape(())

    @staticmethod
    def backward(ctx, gloss):
        with torch.cuda.device(ctx.S.g.device):
            g_s, _, _ = engine.in_mse_backward(ctx.S, gloss, False)
            outs = engine.from_pyramid_nchw(ctx.S.g, g_s)
        return (None, None, None, *outs, *([None] * ctx.n_lvl))


class BaseDistillator(nn.Module):
    def __init__(self, cfg=None):
        super().__init__()
        # kept for state/attribute compatibility; InstanceNorm2d(affine=False) has no parameters or buffers
        self.norm_stu = nn.InstanceNorm2d(256, affine=False)
        self.norm_tea = nn.InstanceNorm2d(256, affine=False)
        self.student, self.teacher = build_customized_detector(cfg)
        self.coef = cfg.MODEL.DISTILLATOR.LAMBDA
        self.add_bg_box = cfg.MODEL.DISTILLATOR.TEACHER.ADD_CONTEXT_BOX
        self.adapter = nn.ModuleDict({'distill': build_adapter(cfg)})
        self._packed = engine.PackedWeights()

    def distill_loss(self, features, images, batched_inputs, batchified_inside_masks, inst_labels):
        losses = dict()
        losses["loss_distill"] = self.distill(features, images, batched_inputs, batchified_inside_masks, inst_labels)
        return losses

    def distill(self, features, images, batched_inputs, batchified_inside_masks, fg_labels):
        """features: {'stu': dict, 'tea': dict} of (B,256,Hi,Wi); the other arguments are accepted and unused exactly
        as in the reference (base_distillator.py:34-64)."""
        keys = sorted(features['stu'].keys() & features['tea'].keys())
        tea_features = [features['tea'][k].detach() for k in keys]      # teacher always detached (:55)
        stu_features = [features['stu'][k] for k in keys]
        if self.distill_flag == 0:                                      # (:52-53)
            stu_features = [f.detach() for f in stu_features]
        if stu_features[0].device.type != 'cuda':
            raise RuntimeError('lgd_b200 distillation runs on CUDA (sm_100a) only; there is no CPU fallback')
        if not hasattr(self, '_packed'):
            self._packed = engine.PackedWeights()
        adapter = self.adapter['distill']
        if type(adapter) is SequentialConvs:
            named = list(adapter.named_parameters())
            names = tuple(n for n, _ in named)
            args = (*stu_features, *tea_features, *[p for _, p in named])
            if not (engine.WGRAD_SIDE_STREAM and engine.DISTILL_SIDE_STREAM and torch.is_grad_enabled()):
                return _DistillFn.apply(self, float(self.coef), names, len(keys), None, *args)
            # The adapter + loss chain (3 convolutions, IN-MSE; backward: 3 dgrads, 3 wgrads) only shares the student
            # operand pair and the finished teacher pyramid with the teacher chain. It runs on its own stream: its
            # tensor-bound kernels fill the teacher chain's HBM- and latency-bound stretches and vice versa, forward
            # and backward (autograd runs this node's backward on the stream its forward ran on and orders the
            # gradients that cross streams).
            dev = stu_features[0].device
            _quiet_cross_stream_accumulate()
            main = torch.cuda.current_stream(dev)
            side = engine.side_stream("distill", dev)
            cache = getattr(self.teacher, "_step_cache", None)
            if cache is not None and cache.get("stu_ready") is not None and cache.get("stu_h") is not None \
                    and cache["key"] == tuple((s.data_ptr(), s._version) for s in stu_features):
                side.wait_event(cache["stu_ready"])      # first input: the operand pair the teacher forward made
            else:
                side.wait_stream(main)
            tea_ready = torch.cuda.Event()
            tea_ready.record(main)                         # the teacher pyramid is complete at this point of main
            with torch.cuda.stream(side):
                loss = _DistillFn.apply(self, float(self.coef), names, len(keys), tea_ready, *args)
            main.wait_stream(side)
            loss.record_stream(main)
            return loss
        stu_features = [adapter(f) for f in stu_features]               # any registered adapter module (:57)
        return _InMseFn.apply(self, float(self.coef), len(keys), *stu_features, *tea_features)

    @abstractmethod
    def forward(self, batched_inputs, **kwargs):
        pass

    @abstractmethod
    def forward_student(self, batched_inputs, **kwargs):
        pass

    @abstractmethod
    def forward_teacher(self, batched_inputs, **kwargs):
        pass
