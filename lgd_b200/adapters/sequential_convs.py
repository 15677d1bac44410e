"""SequentialConvs adapter (models/adapters/sequential_convs.py:7-15): conv3x3-ReLU-conv3x3-ReLU-conv3x3,
256->256 with bias. The nn.Conv2d modules only HOLD the parameters (state_dict names adapter.{0,2,4}.*);
the arithmetic runs in liblgd_b200's tcgen05 convolution."""
import torch
from torch import nn

from .. import engine
from .build import ADAPTERS_REGISTRY


class _AdapterFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, x, *params):
        P = {"a.%d.%s" % (i, n): p for (i, n), p in zip(((0, "weight"), (0, "bias"), (2, "weight"), (2, "bias"),
                                                              (4, "weight"), (4, "bias")), params)}
        g = engine.Geometry.get(x.shape[0], [tuple(x.shape[-2:])], x.device)
        packed = engine.PackedWeights()
        s = engine.to_pyramid(g, [x], True)
        a1 = engine.conv3x3(g, s, packed.get(P["a.0.weight"], 0), P["a.0.bias"], relu=True, round_out=True)
        a2 = engine.conv3x3(g, a1, packed.get(P["a.2.weight"], 0), P["a.2.bias"], relu=True, round_out=True)
        out = engine.conv3x3(g, a2, packed.get(P["a.4.weight"], 0), P["a.4.bias"])
        ctx.saved = (g, P, s, a1, a2)
        ctx.need_dx = x.requires_grad
        return g.level_views(out)[0]

    @staticmethod
    def backward(ctx, gout):
        g, P, s, a1, a2 = ctx.saved
        packed = engine.PackedWeights()
        g_s = engine.to_pyramid(g, [gout], True)
        grads = {}

        def conv_bwd(i, x_in, go, need_dx, mask=None):
            gw, _, gb = engine.conv_wgrad(g, x_in, go, P["a.%d.weight" % i].shape)
            grads[i] = (gw, gb)
            if not need_dx:
                return None
            return engine.conv3x3(g, go, packed.get(P["a.%d.weight" % i], 1), None, relu_mask=mask,
                                  round_out=mask is not None)

        g2 = conv_bwd(4, a2, g_s, True, a2)
        g1 = conv_bwd(2, a1, g2, True, a1)
        g0 = conv_bwd(0, s, g1, ctx.need_dx)
        gx = engine.from_pyramid_nchw(g, g0)[0] if g0 is not None else None
        return (None, gx, grads[0][0], grads[0][1], grads[2][0], grads[2][1], grads[4][0], grads[4][1])


@ADAPTERS_REGISTRY.register()
class SequentialConvs(nn.Module):
    def __init__(self, cfg) -> None:
        super().__init__()
        self.adapter = nn.Sequential(*[nn.Conv2d(256, 256, 3, 1, 1), nn.ReLU(),
                                       nn.Conv2d(256, 256, 3, 1, 1), nn.ReLU(),
                                       nn.Conv2d(256, 256, 3, 1, 1)])

    def forward(self, x):
        """Stand-alone call (the hook API, base_distillator.py:57). BaseDistillator.distill uses the fused
        whole-pyramid path in engine.distill_forward instead of calling this per level."""
        a = self.adapter
        return _AdapterFn.apply(self, x, a[0].weight, a[0].bias, a[2].weight, a[2].bias, a[4].weight, a[4].bias)
