"""GPU parity tests, one kernel family at a time, through the C ABI (ctypes) against plain torch CPU math.
Tolerances: the tcgen05 convolutions consume TF32-rounded operands; the CPU reference is fed the SAME rounded
operands and evaluated in fp64, so the comparison isolates the kernel (fp32 accumulation order only)."""
import ctypes

import numpy as np

import pytest
import torch
import torch.nn.functional as F

from lgd_b200 import engine, synth
from lgd_b200._lib import call, ptr, query
from oracle import lgd_oracle as O
from tests.gpu_util import nchw_to_pyr, pyr_to_nchw_cpu, rel_l2, round_tf32_cpu

pytestmark = pytest.mark.gpu
HWS = [(20, 24), (9, 7), (3, 5), (1, 2)]
CONV_TOL = 2e-5


def _geom(B=2, hws=HWS):
    return engine.Geometry.get(B, hws, torch.device("cuda"))


def _rand_levels(B, hws, seed, scale=1.0):
    gen = torch.Generator().manual_seed(seed)
    return [torch.randn(B, 256, h, w, generator=gen) * scale for h, w in hws]


def test_layout_movers_roundtrip():
    g = _geom()
    xs = _rand_levels(2, HWS, 0)
    buf = nchw_to_pyr(g, xs)
    for v, x in zip(g.level_views(buf), xs):
        assert torch.equal(v.cpu(), x)
    back = engine.from_pyramid_nchw(g, buf)
    for b, x in zip(back, xs):
        assert torch.equal(b.cpu(), x)
    rounded = nchw_to_pyr(g, xs, True)
    for v, x in zip(g.level_views(rounded), xs):
        assert torch.equal(v.cpu(), round_tf32_cpu(x))
    # channels_last inputs take the copy path
    cl = [x.cuda().contiguous(memory_format=torch.channels_last) for x in xs[:3]] + [xs[3].cuda()]
    buf2 = engine.to_pyramid(g, cl, False)
    assert torch.equal(buf2.cpu(), buf.cpu())


@pytest.mark.parametrize("relu,with_mask,per_image_bias", [(False, False, False), (True, False, True), (False, True, False)])
def test_conv3x3_forward(relu, with_mask, per_image_bias):
    B = 2
    g = _geom(B)
    gen = torch.Generator().manual_seed(1)
    xs = [round_tf32_cpu(x) for x in _rand_levels(B, HWS, 2)]
    w = torch.randn(256, 256, 3, 3, generator=gen) / 48.0
    bias = torch.randn(g.F, B, 256, generator=gen) if per_image_bias else torch.randn(256, generator=gen)
    x_buf = nchw_to_pyr(g, xs)
    packed = engine.PackedWeights().get(w.cuda(), 0)
    masks = _rand_levels(B, HWS, 3) if with_mask else None
    m_buf = nchw_to_pyr(g, masks) if with_mask else None
    out, stats = engine.conv3x3(g, x_buf, packed, bias.cuda().contiguous(), relu=relu, relu_mask=m_buf, stats=True,
                                bias_strides=(B * 256, 256) if per_image_bias else (0, 0))
    torch.cuda.synchronize()
    outs = pyr_to_nchw_cpu(g, out)
    wr = round_tf32_cpu(w).double()
    stats = stats.cpu().view(g.F, B, 2)
    for l, (x, o) in enumerate(zip(xs, outs)):
        ref = F.conv2d(x.double(), wr, None, padding=1)
        ref = ref + (bias[l].double()[:, :, None, None] if per_image_bias else bias.double()[None, :, None, None])
        raw = ref.clone()
        if relu:
            ref = ref.relu()
        if with_mask:
            ref = torch.where(masks[l] > 0, ref, torch.zeros_like(ref))
        assert rel_l2(o, ref) < CONV_TOL, (l, rel_l2(o, ref))
        mean = raw.flatten(1).mean(1)
        var = raw.flatten(1).var(1, unbiased=False)
        assert torch.allclose(stats[l, :, 0].double(), mean, atol=1e-5, rtol=1e-4)
        assert torch.allclose(stats[l, :, 1].double(), (var + 1e-5).rsqrt(), rtol=1e-4)
    # epilogue by-product: per-(level,image) channel sums of the stored values (bias gradient of the layer below when
    # this launch is a dgrad through a ReLU mask), odd tile count incl. the CTA pair's dummy tile
    out2, sums, total = engine.conv3x3(g, x_buf, packed, bias.cuda().contiguous(), relu=relu, relu_mask=m_buf, csum=True,
                                       bias_strides=(B * 256, 256) if per_image_bias else (0, 0), round_out=True)
    torch.cuda.synchronize()
    assert rel_l2(out2, round_tf32_cpu(out.cpu())) < 1e-7
    sums = sums.cpu().view(g.F, B, 256)
    tot_ref = torch.zeros(256, dtype=torch.float64)
    for l, o in enumerate(outs):
        ref_s = o.double().sum((2, 3))
        assert rel_l2(sums[l], ref_s) < 1e-5, (l, rel_l2(sums[l], ref_s))
        tot_ref += ref_s.sum(0)
    assert rel_l2(total.cpu(), tot_ref) < 1e-5


def test_conv3x3_forward_fp16_operands():
    """Forward convolution on fp16 operands (fp32 accumulate): exact up to accumulation order on fp16-representable
    inputs, GN statistics and the fp16 copy of the output included; layout mover and GN apply write matching shadows."""
    B = 2
    g = _geom(B)
    gen = torch.Generator().manual_seed(4)
    xs = [x.half().float() for x in _rand_levels(B, HWS, 5)]
    w = (torch.randn(256, 256, 3, 3, generator=gen) / 48.0).half().float()
    bias = torch.randn(256, generator=gen)
    x_buf, x_half = engine.to_pyramid(g, [x.cuda() for x in xs], True, want_half=True)
    assert torch.equal(x_half, x_buf.half())
    pk = engine.PackedWeights().get(w.cuda(), "h")
    out, st, out_h = engine.conv3x3_f16(g, x_half, pk, bias.cuda(), relu=True, round_out=False, stats=True, want_half=True)
    torch.cuda.synchronize()
    st = st.cpu().view(g.F, B, 2)
    for l, (x, o, oh) in enumerate(zip(xs, pyr_to_nchw_cpu(g, out), pyr_to_nchw_cpu(g, out_h.float()))):
        raw = F.conv2d(x.double(), w.double(), bias.double(), padding=1)
        ref = raw.relu()
        assert rel_l2(o, ref) < CONV_TOL, (l, rel_l2(o, ref))
        assert rel_l2(oh, ref) < 4e-4                      # fp16 rounding of the stored value (2^-11 relative)
        assert torch.allclose(st[l, :, 0].double(), raw.flatten(1).mean(1), atol=1e-5, rtol=1e-4)
    # the fp16 copy of a conv+ReLU output doubles as the ReLU mask of the backward: a positive value, however small,
    # never rounds to zero (weights scaled so that the outputs sit far below fp16's smallest subnormal)
    tiny, tiny_h = engine.conv3x3_f16(g, (x_buf * 1e-4).half(), engine.PackedWeights().get((w * 2.0 ** -14).cuda(), "h"),
                                      (bias * 1e-12).cuda(), relu=True, want_half=True)
    assert float(tiny.max()) < 6e-8 and torch.equal(tiny_h != 0, tiny > 0) and int((tiny > 0).sum()) > 1000
    only_h = engine.conv3x3_f16(g, x_half, pk, bias.cuda(), relu=True, want_half=True, want_fp32=False)
    assert only_h[0] is None and torch.equal(only_h[1], out_h)
    # GroupNorm apply writes the TF32-rounded fp32 tensor and the fp16 shadow from the same un-rounded value
    y, y_h = engine.gn_apply(g, out, torch.stack([torch.zeros(g.F * B), torch.ones(g.F * B)], 1).cuda().contiguous(),
                             True, True, want_half=True)
    assert torch.equal(y_h, out.relu().half())


def test_conv3x3_forward_split_operand_tf32x3(monkeypatch):
    """"tf32x3" forward (x = hi + lo, three chained TF32 launches through the addend epilogue): fp32-accurate on inputs
    that are NOT TF32-representable, statistics and the operand pair of the output included."""
    monkeypatch.setattr(engine, "FORWARD_PRECISION", "tf32x3")
    B = 2
    g = _geom(B)
    gen = torch.Generator().manual_seed(6)
    xs = _rand_levels(B, HWS, 7)
    w = torch.randn(256, 256, 3, 3, generator=gen) / 48.0
    bias = torch.randn(256, generator=gen)
    x_hi, x_lo = engine.student_operands(g, [x.cuda() for x in xs])
    torch.cuda.synchronize()
    for x, hi, lo in zip(xs, pyr_to_nchw_cpu(g, x_hi), pyr_to_nchw_cpu(g, x_lo)):
        assert torch.equal(hi, round_tf32_cpu(x))
        assert torch.equal(lo, round_tf32_cpu(x - hi))
        assert float((x - hi - lo).abs().max()) <= 2.0 ** -21 * float(x.abs().max())
    pw = engine.PackedWeights()
    out, st, out_lo = engine.fwd_conv(g, x_hi, x_lo, w.cuda(), pw, bias.cuda(), relu=True, stats=True, want_comp=True)
    torch.cuda.synchronize()
    st = st.cpu().view(g.F, B, 2)
    for l, (x, o, ol) in enumerate(zip(xs, pyr_to_nchw_cpu(g, out), pyr_to_nchw_cpu(g, out_lo))):
        raw = F.conv2d(x.double(), w.double(), bias.double(), padding=1)
        ref = raw.relu()
        assert rel_l2(o + ol, ref) < 1e-5, (l, rel_l2(o + ol, ref))     # measured 3.6e-6; 3e-4 for one TF32 / fp16 pass
        assert torch.equal(o, round_tf32_cpu(o))
        assert torch.allclose(st[l, :, 0].double(), raw.flatten(1).mean(1), atol=1e-5, rtol=1e-4)
    # a single TF32 pass on the same un-rounded data for comparison: the error the split removes
    single = engine.conv3x3(g, x_hi, pw.get(w.cuda(), 0), bias.cuda(), relu=True)
    e1 = max(rel_l2(o, F.conv2d(x.double(), w.double(), bias.double(), padding=1).relu())
             for x, o in zip(xs, pyr_to_nchw_cpu(g, single)))
    assert 5e-5 < e1 < 1e-3, e1


@pytest.mark.parametrize("magnitude", [1.0, 1e-9, 1e6])
def test_conv3x3_dgrad_fp16_operands_scaled(magnitude):
    """dgrad on fp16 operands with a power-of-two gradient scale: same answer as fp64 on the rounded operands for
    gradients of any magnitude (1e-9: what a mean-reduced loss produces; 1e6: far outside fp16's range unscaled), with
    the ReLU mask, the bias-gradient sums and the scaled fp16 operand of the NEXT dgrad from the same launch."""
    B = 2
    g = _geom(B)
    gen = torch.Generator().manual_seed(21)
    gs = [x * magnitude for x in _rand_levels(B, HWS, 22)]
    masks = [(torch.rand(x.shape, generator=gen) > 0.5).float() for x in gs]
    w = torch.randn(256, 256, 3, 3, generator=gen) / 48.0
    g_buf, m_buf = nchw_to_pyr(g, gs), nchw_to_pyr(g, masks)
    # operand pair of g: scale from the exact norm (what the producers derive from their by-products)
    sq = (g_buf.double() ** 2).sum().float().reshape(1)
    sc = torch.empty(3, device="cuda")
    call("lgd_grad_scale", ptr(sq), 1, 1, None, None, 1.0, ptr(sc))
    gh = (g_buf * sc[0]).half()
    assert torch.isfinite(gh.float()).all()
    pw = engine.PackedWeights()
    dx, sums, total, (dx_h, sc_dx) = engine.dgrad_conv_f16(g, (gh, sc), w.cuda(), pw, relu_mask=m_buf, want_half=True)
    torch.cuda.synchronize()
    ph, gain = pw.get(w.cuda(), "hd")
    wh = w.half().double()
    op_norm_bound = float(gain)
    assert op_norm_bound >= float(torch.linalg.matrix_norm(w.double().reshape(256, -1), 2))   # >= one unfolding's norm
    tot_ref = torch.zeros(256, dtype=torch.float64)
    for gl, m, o in zip(pyr_to_nchw_cpu(g, gh.float() / sc[0]), masks, pyr_to_nchw_cpu(g, dx)):
        ref = F.conv_transpose2d(gl.double(), wh, padding=1) * m.double()
        assert rel_l2(o, ref) < CONV_TOL, rel_l2(o, ref)
        tot_ref += ref.sum((0, 2, 3))
    assert rel_l2(total.cpu(), tot_ref) < 1e-4
    _check_scaled_half_bound(dx, dx_h, sc_dx)
    # the ReLU mask given as the fp16 copy of the activation (nonzero = pass) instead of an fp32 tensor; no fp32 output
    dx2, sums2, total2, (dx2_h, _) = engine.dgrad_conv_f16(g, (gh, sc), w.cuda(), pw, relu_mask_half=m_buf.half(),
                                                            want_half=True, want_fp32=False)
    assert dx2 is None and torch.equal(dx2_h, dx_h) and torch.equal(total2, total) and torch.equal(sums2, sums)
    # against the un-rounded gradient: fp16 operand rounding only (same 10-bit mantissa as TF32)
    for gl, m, o in zip(gs, masks, pyr_to_nchw_cpu(g, dx)):
        ref = F.conv_transpose2d(gl.double(), w.double(), padding=1) * m.double()
        assert rel_l2(o, ref) < 1e-3


def _check_scaled_half_bound(g32, g16, sc):
    """operand written with an a-priori bound (gain * U_in): never overflows, keeps its mantissa, and the triple
    carries the measured norm (pre-mask, so >= the real one) for the next bound."""
    s, inv_s, U = [float(v) for v in sc.cpu()]
    norm = float(g32.double().norm())
    assert norm <= U * (1 + 1e-5) and U <= 4.0 * norm, (norm, U)
    assert torch.isfinite(g16.float()).all() and float(g16.float().abs().max()) < 65504
    assert rel_l2(g16.float() / s, g32) < 4e-4


def test_conv3x3_round_out_and_full_size_tiles():
    # one full-size level exercises every tile position incl. ragged right/bottom edges
    B, hws = 1, [(50, 84)]
    g = _geom(B, hws)
    gen = torch.Generator().manual_seed(5)
    x = round_tf32_cpu(torch.randn(B, 256, 50, 84, generator=gen))
    w = torch.randn(256, 256, 3, 3, generator=gen) / 48.0
    out = engine.conv3x3(g, nchw_to_pyr(g, [x]), engine.PackedWeights().get(w.cuda(), 0), None, round_out=True)
    ref = F.conv2d(x.double(), round_tf32_cpu(w).double(), None, padding=1)
    o = pyr_to_nchw_cpu(g, out)[0]
    assert rel_l2(o, ref) < 3e-4            # rounding of the stored value to TF32
    assert torch.equal(o, round_tf32_cpu(o))  # stored values are TF32-representable


@pytest.mark.parametrize("hws", [[(3, 128), (2, 129)], [(5, 200)], [(2, 337), (1, 1)], [(7, 126), (1, 131)]])
def test_conv3x3_input_strips_at_every_width_class(hws):
    """The forward / dgrad kernel reads ONE input strip per tile for all nine taps: a contiguous run of padded positions
    for narrow levels (W + 2 <= 130), three runs of 130 positions for wide ones. Widths on both sides of that boundary, a
    level wider than two tiles, single-pixel and single-row levels, odd tile counts per level (the padding tile of a CTA
    pair) -- fp16 and TF32 operands, with bias + ReLU, against fp64 on the rounded operands."""
    B = 3
    g = _geom(B, hws)
    gen = torch.Generator().manual_seed(sum(h * 1000 + w for h, w in hws))
    xs = [torch.randn(B, 256, h, w, generator=gen) for h, w in hws]
    w = torch.randn(256, 256, 3, 3, generator=gen) / 48.0
    bias = torch.randn(256, generator=gen)
    x_buf = nchw_to_pyr(g, xs)
    pw = engine.PackedWeights()
    # fp16 operands
    x_h = x_buf.half()
    out = g.new()
    call("lgd_conv3x3_fwd_f16", g.pref, ptr(x_h), ptr(pw.get(w.cuda(), "h")), ptr(bias.cuda()), 0, 0, ptr(out), None, 1, 0, None)
    for x, o in zip(xs, pyr_to_nchw_cpu(g, out)):
        ref = F.conv2d(x.half().double(), w.half().double(), bias.double(), padding=1).relu()
        assert rel_l2(o, ref) < CONV_TOL, (tuple(x.shape), rel_l2(o, ref))
    # TF32 operands (32-channel strips, twice as many of them per tile)
    xr = [round_tf32_cpu(x) for x in xs]
    out32 = engine.conv3x3(g, nchw_to_pyr(g, xr), pw.get(w.cuda(), 0), bias.cuda(), relu=True)
    for x, o in zip(xr, pyr_to_nchw_cpu(g, out32)):
        ref = F.conv2d(x.double(), round_tf32_cpu(w).double(), bias.double(), padding=1).relu()
        assert rel_l2(o, ref) < CONV_TOL, (tuple(x.shape), rel_l2(o, ref))


def test_conv3x3_dgrad_and_wgrad():
    B = 2
    g = _geom(B)
    gen = torch.Generator().manual_seed(7)
    xs = [round_tf32_cpu(x) for x in _rand_levels(B, HWS, 8)]
    gos = [round_tf32_cpu(x) for x in _rand_levels(B, HWS, 9, 1e-3)]
    w = torch.randn(256, 256, 3, 3, generator=gen) / 48.0
    wr = round_tf32_cpu(w).double()
    x_buf, go_buf = nchw_to_pyr(g, xs), nchw_to_pyr(g, gos)
    dx = engine.conv3x3(g, go_buf, engine.PackedWeights().get(w.cuda(), 1), None)
    gw, sums, gb = engine.conv_wgrad(g, x_buf, go_buf, w.shape)
    torch.cuda.synchronize()
    dxs = pyr_to_nchw_cpu(g, dx)
    gw_ref = torch.zeros_like(wr)
    gb_ref = torch.zeros(256, dtype=torch.float64)
    sums = sums.cpu().view(g.F, B, 256)
    for l, (x, go) in enumerate(zip(xs, gos)):
        ref_dx = torch.nn.grad.conv2d_input(x.shape, wr, go.double(), padding=1)
        assert rel_l2(dxs[l], ref_dx) < CONV_TOL, (l, rel_l2(dxs[l], ref_dx))
        gw_ref += torch.nn.grad.conv2d_weight(x.double(), w.shape, go.double(), padding=1)
        gb_ref += go.double().sum((0, 2, 3))
        assert rel_l2(sums[l], go.double().sum((2, 3))) < 1e-5
    assert rel_l2(gw.cpu(), gw_ref) < CONV_TOL, rel_l2(gw.cpu(), gw_ref)
    assert rel_l2(gb.cpu(), gb_ref) < 1e-5


@pytest.mark.parametrize("magnitude", [1.0, 1e-8])
def test_conv3x3_wgrad_fp16_operands(magnitude):
    """Weight gradient on fp16 operands (MN-major, plain SWIZZLE_128B): fp16 copy of the input from the forward,
    scaled fp16 copy of the output gradient, scale divided out in the second-stage reduction. Exact up to
    accumulation order against fp64 on the same rounded operands; every filter tap and both CTA halves covered."""
    B = 2
    g = _geom(B)
    gen = torch.Generator().manual_seed(31)
    xs = [x.half().float() for x in _rand_levels(B, HWS, 32)]
    gos = [x * magnitude for x in _rand_levels(B, HWS, 33)]
    x_buf, go_buf = nchw_to_pyr(g, xs), nchw_to_pyr(g, gos)
    sq = (go_buf.double() ** 2).sum().float().reshape(1)
    sc = torch.empty(3, device="cuda")
    call("lgd_grad_scale", ptr(sq), 1, 1, None, None, 1.0, ptr(sc))
    goh = (go_buf * sc[0]).half()
    ws_ = engine.WgradStream(g)
    gw = ws_.wgrad(x_buf, go_buf, (256, 256, 3, 3), x_half=x_buf.half(), operand=(goh, sc))
    ws_.join()
    torch.cuda.synchronize()
    gw_ref = torch.zeros(256, 256, 3, 3, dtype=torch.float64)
    for x, go in zip(xs, pyr_to_nchw_cpu(g, goh.float() / sc[0])):
        gw_ref += torch.nn.grad.conv2d_weight(x.double(), gw_ref.shape, go.double(), padding=1)
    assert rel_l2(gw.cpu(), gw_ref) < CONV_TOL, rel_l2(gw.cpu(), gw_ref)
    for tap in range(9):   # no tap may hide behind the others
        assert rel_l2(gw.cpu()[:, :, tap // 3, tap % 3], gw_ref[:, :, tap // 3, tap % 3]) < CONV_TOL, tap
    # and against the un-rounded gradient: operand rounding only
    gw_full = torch.zeros_like(gw_ref)
    for x, go in zip(xs, gos):
        gw_full += torch.nn.grad.conv2d_weight(x.double(), gw_ref.shape, go.double(), padding=1)
    assert rel_l2(gw.cpu(), gw_full) < 1e-3


def test_groupnorm_apply_and_backward():
    B = 2
    g = _geom(B)
    xs = _rand_levels(B, HWS, 11)
    xs = [x * (1 + l) + 0.3 * l for l, x in enumerate(xs)]
    gys = _rand_levels(B, HWS, 12)
    # statistics via a bias-only "conv": use torch stats directly to test apply/bwd in isolation
    st = torch.stack([torch.stack([x.flatten(1).mean(1), (x.flatten(1).var(1, unbiased=False) + 1e-5).rsqrt()], 1)
                      for x in xs], 0).contiguous()  # (F,B,2)
    x_buf, gy_buf = nchw_to_pyr(g, xs), nchw_to_pyr(g, gys)
    for relu in (False, True):
        y, ist = engine.gn_apply(g, x_buf, st.cuda(), relu, False, in_stats=True)
        # fused by-product: InstanceNorm statistics (per image and channel) of the stored output
        ist = ist.cpu().view(g.F, B, 256, 2)
        gx, gb, (gx_h, sc) = engine.gn_bwd(g, gy_buf, x_buf, st.cuda(), relu, False, want_half=True)
        _check_scaled_half(gx, gx_h, sc)
        ys, gxs = pyr_to_nchw_cpu(g, y), pyr_to_nchw_cpu(g, gx)
        gb_ref = torch.zeros(256, dtype=torch.float64)
        for l, x in enumerate(xs):
            xd = x.double().requires_grad_(True)
            ref = F.group_norm(xd, 1, eps=1e-5)
            if relu:
                ref = ref.relu()
            ref.backward(gys[l].double())
            assert rel_l2(ys[l], ref) < 1e-5
            assert rel_l2(gxs[l], xd.grad) < 2e-5, (l, relu, rel_l2(gxs[l], xd.grad))
            gb_ref += xd.grad.sum((0, 2, 3))
            yd = ys[l].double().flatten(2)
            assert rel_l2(ist[l, :, :, 0], yd.mean(2)) < 1e-5
            assert rel_l2(ist[l, :, :, 1], (yd.var(2, unbiased=False) + 1e-5).rsqrt()) < 1e-5
        assert rel_l2(gb.cpu(), gb_ref) < 2e-5   # fused by-product: bias gradient of the conv in front


@pytest.mark.parametrize("relu", [False, True])
def test_dgrad_epilogue_groupnorm_sums_match_two_pass_backward(relu):
    """dgrad whose epilogue emits the per-tile sums of (g, g*xhat, g^2) of the GroupNorm(1)(+ReLU) backward that consumes
    its output (layers.py:6-7 backward): the dgrad output is bit-identical to the plain dgrad, and the GroupNorm backward
    fed with the tile sums equals the two-pass one up to the summation order of three scalars per (level, image)."""
    B = 2
    g = _geom(B)
    gen = torch.Generator().manual_seed(91)
    gs = _rand_levels(B, HWS, 92)
    xs = [x * (1 + l) + 0.3 * l for l, x in enumerate(_rand_levels(B, HWS, 93))]
    w = torch.randn(256, 256, 3, 3, generator=gen) / 48.0
    st = torch.stack([torch.stack([x.flatten(1).mean(1), (x.flatten(1).var(1, unbiased=False) + 1e-5).rsqrt()], 1)
                      for x in xs], 0).contiguous().cuda()
    g_buf, x_buf = nchw_to_pyr(g, gs), nchw_to_pyr(g, xs)
    sq = (g_buf.double() ** 2).sum().float().reshape(1)
    sc = torch.empty(3, device="cuda")
    call("lgd_grad_scale", ptr(sq), 1, 1, None, None, 1.0, ptr(sc))
    gh = (g_buf * sc[0]).half()
    pw = engine.PackedWeights()
    dx_plain, _, _, _ = engine.dgrad_conv_f16(g, (gh, sc), w.cuda(), pw)
    dx, _, _, tile_gn = engine.dgrad_conv_f16(g, (gh, sc), w.cuda(), pw, gn_site=(x_buf, st, relu))
    assert torch.equal(dx, dx_plain)
    # tile sums against fp64 over the whole pyramid
    xhat = torch.cat([((x.double() - x.double().flatten(1).mean(1)[:, None, None, None]) *
                       (x.double().flatten(1).var(1, unbiased=False) + 1e-5).rsqrt()[:, None, None, None])
                      .permute(0, 2, 3, 1).reshape(-1) for x in xs])
    gd = dx.double().cpu().reshape(-1)
    if relu:
        gd = gd * (xhat > 0)
    tg = tile_gn.view(-1, 4).double().cpu()
    for j, ref in enumerate((gd.sum(), (gd * xhat).sum(), (gd * gd).sum())):
        scale = float((gd.abs() * (xhat.abs() if j == 1 else 1)).sum()) if j < 2 else float(ref)
        assert abs(float(tg[:, j].sum()) - float(ref)) < 1e-5 * scale, (j, float(tg[:, j].sum()), float(ref))
    a = engine.gn_bwd(g, dx, x_buf, st, relu, False, want_half=True)
    b = engine.gn_bwd(g, dx, x_buf, st, relu, False, want_half=True, tile_gn=tile_gn)
    assert rel_l2(b[0].cpu(), a[0].cpu()) < 1e-6 and rel_l2(b[1].cpu(), a[1].cpu()) < 1e-6
    if relu:
        # the same sums from the fp16 copy of y = relu(GroupNorm(x)) (what the teacher backward uses: the copy is the operand
        # of the convolution being differentiated): y != 0 <=> xhat > 0, g * xhat = g * y up to the fp16 rounding of y
        _, y_h = engine.gn_apply(g, x_buf, st, True, False, want_half=True)
        dx2, _, _, tile_y = engine.dgrad_conv_f16(g, (gh, sc), w.cuda(), pw, gn_site=(x_buf, st, True, y_h))
        assert torch.equal(dx2, dx_plain)
        ty = tile_y.view(-1, 4).double().cpu()
        assert torch.equal(ty[:, 0], tg[:, 0])                                          # same activation pattern
        assert rel_l2(ty[:, 2], tg[:, 2]) < 1e-6
        ref1 = float((gd * xhat).sum())
        assert abs(float(ty[:, 1].sum()) - ref1) < 2e-5 * float((gd.abs() * xhat.abs()).sum())
        c = engine.gn_bwd(g, dx, x_buf, st, True, False, want_half=True, tile_gn=tile_y)
        assert rel_l2(c[0].cpu(), a[0].cpu()) < 1e-5 and rel_l2(c[1].cpu(), a[1].cpu()) < 1e-5
    assert torch.equal(a[2][1][:2], b[2][1][:2])   # same power-of-two scale
    assert rel_l2(b[2][0].float().cpu(), a[2][0].float().cpu()) < 1e-3


@pytest.mark.parametrize("relu", [False, True])
def test_groupnorm32_affine_forward_backward(relu):
    """GroupNorm(32, 256) with affine parameters (+ReLU), the tower norm of the FCOS-family heads
    (thirdparty_heads/fcos.py:455-476): statistics, apply (fp32 and the fp16 operand copy), and the backward (input
    gradient as fp32 and as the scaled fp16 operand, dgamma, dbeta, and the bias gradient of the convolution in front)
    against torch fp64."""
    B = 2
    g = _geom(B)
    gen = torch.Generator().manual_seed(40)
    xs = [x * (1 + l) + 0.3 * l for l, x in enumerate(_rand_levels(B, HWS, 41))]
    gys = _rand_levels(B, HWS, 42)
    gamma = (1 + 0.3 * torch.randn(256, generator=gen)).cuda()
    beta = (0.2 * torch.randn(256, generator=gen)).cuda()
    x_buf, gy_buf = nchw_to_pyr(g, xs), nchw_to_pyr(g, gys)
    nseg = g.F * B
    stats = torch.empty(nseg * 64, device="cuda")
    chsum = torch.empty(nseg * 256, device="cuda")
    ws = torch.empty(query("lgd_gn32_workspace", g.pref), dtype=torch.uint8, device="cuda")
    call("lgd_gn32_stats", g.pref, ptr(x_buf), ptr(stats), ptr(chsum), ptr(ws), ws.numel())
    y_h, y32 = g.new_half(), g.new()
    call("lgd_gn32_apply", g.pref, ptr(x_buf), ptr(stats), ptr(gamma), ptr(beta), int(relu), ptr(y_h), ptr(y32))
    gx_h, gx, sc = g.new_half(), g.new(), torch.empty(3, device="cuda")
    dg, db, dbias = (torch.empty(256, device="cuda") for _ in range(3))
    call("lgd_gn32_bwd", g.pref, ptr(gy_buf), ptr(x_buf), ptr(stats), ptr(chsum), ptr(gamma), ptr(beta), int(relu),
         ptr(gx_h), ptr(sc), ptr(gx), ptr(dg), ptr(db), ptr(dbias), ptr(ws), ws.numel())
    torch.cuda.synchronize()
    gd = gamma.double().cpu().requires_grad_(True)
    bd = beta.double().cpu().requires_grad_(True)
    ys, gxs = pyr_to_nchw_cpu(g, y32), pyr_to_nchw_cpu(g, gx)
    yhs = pyr_to_nchw_cpu(g, y_h.float())
    dbias_ref = torch.zeros(256, dtype=torch.float64)
    st = stats.view(g.F, B, 32, 2).cpu()
    for l, x in enumerate(xs):
        xd = x.double().requires_grad_(True)
        ref = F.group_norm(xd, 32, gd, bd, 1e-5)
        if relu:
            ref = ref.relu()
        ref.backward(gys[l].double())
        xg = x.double().view(B, 32, -1)
        assert rel_l2(st[l, :, :, 0], xg.mean(2)) < 1e-5
        assert rel_l2(st[l, :, :, 1], (xg.var(2, unbiased=False) + 1e-5).rsqrt()) < 1e-5
        assert rel_l2(ys[l], ref) < 1e-5, (l, rel_l2(ys[l], ref))
        assert rel_l2(yhs[l], ref) < 5e-4
        if relu:
            assert torch.equal(yhs[l] != 0, ys[l] > 0), "the fp16 copy shows the activation pattern"
        assert rel_l2(gxs[l], xd.grad) < 3e-5, (l, relu, rel_l2(gxs[l], xd.grad))
        dbias_ref += xd.grad.sum((0, 2, 3))
    assert rel_l2(dg.cpu(), gd.grad) < 2e-5 and rel_l2(db.cpu(), bd.grad) < 2e-5
    # channel sums of gx: tiny next to the gradient itself (GroupNorm removes the group mean), compare on its scale
    assert float((dbias.cpu().double() - dbias_ref).norm()) <= 2e-5 * float(gx.double().norm()) + 1e-5 * float(dbias_ref.norm())
    _check_scaled_half(gx, gx_h, sc)


def _check_scaled_half(g32, g16, sc):
    """fp16 gradient operand: g16 = fp16(g32 * s), s a power of two with U*s <= 2^14 for an upper bound U of ||g32||_2
    that is not absurdly loose (so that the values keep their mantissa)."""
    s, inv_s, U = [float(v) for v in sc.cpu()]
    norm = float(g32.double().norm())
    assert s > 0 and abs(s * inv_s - 1.0) < 1e-6 and np.log2(s) == round(np.log2(s))
    assert norm <= U * (1 + 1e-5), (norm, U)
    assert U <= 16.0 * norm + 1e-30, (norm, U)          # measured: 1.0-1.5x
    assert 2.0 ** 13 < U * s <= 2.0 ** 14
    assert torch.isfinite(g16.float()).all()
    assert rel_l2(g16.float() / s, g32) < 4e-4


@pytest.mark.parametrize("moments,corr", [(True, 0.0), (False, 0.0), (True, 0.97)])
def test_instance_norm_mse_forward_backward(moments, corr):
    """moments=True: loss, statistics and the backward's per-channel totals from one pass of five shifted moments;
    moments=False: explicit statistics + squared-difference passes. corr: teacher built as corr*student + noise, the
    regime late in training where the moment form subtracts nearly equal numbers (loss = 2(1-corr) per element)."""
    B = 2
    g = _geom(B)
    ss = [x * 2 + 0.5 for x in _rand_levels(B, HWS, 13)]
    ts = [corr * s + (1 - corr ** 2) ** 0.5 * 2 * n + 3.0 for s, n in zip(ss, _rand_levels(B, HWS, 14))]
    s_buf, t_buf = nchw_to_pyr(g, ss), nchw_to_pyr(g, ts)
    loss, S = engine.in_mse_forward(g, s_buf, t_buf, 1.7, moments=moments)
    assert (S.bwd_sums is not None) == moments
    gl = torch.tensor([0.6], device="cuda")
    gs, gb, op = engine.in_mse_backward(S, gl, False, want_half=moments)
    if moments:
        _check_scaled_half(gs, *op)
    sd = [s.double().requires_grad_(True) for s in ss]
    a = torch.cat([F.instance_norm(s, eps=1e-5).reshape(B, -1) for s in sd], 1)
    b = torch.cat([F.instance_norm(t.double(), eps=1e-5).reshape(B, -1) for t in ts], 1)
    ref = 1.7 * F.mse_loss(b, a)
    (ref * 0.6).backward()
    assert abs(float(loss) - float(ref)) < 1e-5 * float(ref)
    # levels with <= 2 pixels per channel are degenerate for InstanceNorm (output +-1 whatever the input): the exact
    # gradient is ~1e-13 and only its absolute size is meaningful, so errors are measured against the whole pyramid
    scale = sum(float(s.grad.norm()) ** 2 for s in sd) ** 0.5
    for o, s in zip(pyr_to_nchw_cpu(g, gs), sd):
        err = float((o.double() - s.grad).norm())
        assert err < 5e-5 * (scale if s.shape[-1] * s.shape[-2] <= 2 else float(s.grad.norm())), err
    # d loss / d bias of the conv in front of the InstanceNorm is analytically zero; the fused channel sums of the
    # un-rounded gradient must reproduce that to fp32 round-off
    assert float(gb.abs().max()) < 1e-3 * scale


def _box_setup(B=3, img=(120, 150), seed=3):
    bi, im, feats = synth.synth_batch(B, img[0], img[1], seed=seed, adversarial=True, n_boxes=[None, 0, 5][:B])
    H, W = im.tensor.shape[-2:]
    hws = [tuple(f.shape[-2:]) for f in feats.values()]
    g = engine.Geometry.get(B, hws, torch.device("cuda"))
    return bi, (H, W), g, feats


@pytest.mark.parametrize("ctx", [True, False])
def test_box_ranges_masks_and_descriptors_exact(ctx):
    bi, (H, W), g, _ = _box_setup()
    tb = engine.build_box_table(bi, H, W, ctx, torch.device("cuda"))
    ranges = torch.empty(g.F * tb.T * 4, device="cuda", dtype=torch.int32)
    call("lgd_box_ranges", ptr(tb.boxes), tb.T, H, W, g.pref, ptr(ranges))
    masks = torch.empty(tb.T * g.P, device="cuda", dtype=torch.float32)
    call("lgd_masks_from_ranges", ptr(ranges), tb.T, g.pref, ptr(masks))
    desc = torch.empty(tb.T, 84, device="cuda")
    call("lgd_encode_descriptors", ptr(tb.boxes), ptr(tb.labels), tb.T, H, W, ptr(desc))
    per_img = O.prepare_boxes([x["instances"] for x in bi], H, W, ctx)
    ref_boxes = torch.cat([b for b, _, _ in per_img], 0)
    assert torch.equal(tb.boxes.cpu().view(-1, 4), ref_boxes)
    ref_desc = torch.cat([O.encode_descriptors(b, oh, H, W) for b, oh, _ in per_img], 0)
    assert torch.equal(desc.cpu(), ref_desc), "descriptors must be bit exact"
    off = 0
    for (h, w) in g.hws:
        ref = O.inside_mask(ref_boxes, (H, W), (h, w))
        got = masks[off:off + tb.T * h * w].view(tb.T, h * w).cpu()
        assert torch.equal(got, ref), "label->region assignment must be bit exact"
        off += tb.T * h * w


def test_maskpool_and_render_forward_backward():
    bi, (H, W), g, feats = _box_setup()
    ctx = True
    tb = engine.build_box_table(bi, H, W, ctx, torch.device("cuda"))
    T, F_ = tb.T, g.F
    ranges = torch.empty(F_ * T * 4, device="cuda", dtype=torch.int32)
    call("lgd_box_ranges", ptr(tb.boxes), T, H, W, g.pref, ptr(ranges))
    xs = [f * 1.5 + 0.2 for f in feats.values()]
    x_buf = nchw_to_pyr(g, xs)
    st = torch.stack([torch.stack([x.flatten(1).mean(1), (x.flatten(1).var(1, unbiased=False) + 1e-5).rsqrt()], 1)
                      for x in xs], 0).contiguous().cuda()
    pooled = torch.empty(F_ * T, 256, device="cuda")
    ws = g.workspace(query("lgd_maskpool_workspace", g.pref, T))
    call("lgd_maskpool_fwd", g.pref, ptr(x_buf), ptr(st), ptr(ranges), ptr(tb.img_of), T, ptr(pooled), ptr(ws), ws.numel())
    gen = torch.Generator().manual_seed(5)
    gp = torch.randn(F_ * T, 256, generator=gen)
    gy = g.new()
    gp_c = gp.cuda()
    call("lgd_maskpool_bwd", g.pref, ptr(gp_c), ptr(ranges), ptr(tb.img_start), T, ptr(gy))
    emb = torch.randn(F_ * T, 256, generator=gen)
    rend = g.new()
    emb_c = emb.cuda()
    rend_h = g.new_half()
    call("lgd_render_fwd", g.pref, ptr(emb_c), ptr(ranges), ptr(tb.img_start), ptr(tb.n_render), T, ptr(rend), 0,
         ptr(rend_h))
    assert torch.equal(rend_h, rend.half())      # fp16 shadow written by the same pass
    gr_levels = _rand_levels(g.B, g.hws, 6)
    gemb = torch.empty(F_ * T, 256, device="cuda")
    call("lgd_render_bwd", g.pref, ptr(nchw_to_pyr(g, gr_levels)), ptr(ranges), ptr(tb.img_of), ptr(tb.img_start),
         ptr(tb.n_render), T, ptr(gemb), ptr(ws), ws.numel())
    torch.cuda.synchronize()
    per_img = O.prepare_boxes([x["instances"] for x in bi], H, W, ctx)
    pooled, gemb = pooled.cpu().view(F_, T, 256), gemb.cpu().view(F_, T, 256)
    gys, rends = pyr_to_nchw_cpu(g, gy), pyr_to_nchw_cpu(g, rend)
    for l, (h, w) in enumerate(g.hws):
        t0 = 0
        for b, (boxes, _, _) in enumerate(per_img):
            n = boxes.shape[0]
            m = O.inside_mask(boxes, (H, W), (h, w)).double()
            y = F.group_norm(xs[l][b:b + 1].double(), 1, eps=1e-5).relu()[0].flatten(1).requires_grad_(True)
            ref = (m @ y.T) / m.sum(-1).clamp(min=1.0)[:, None]
            assert rel_l2(pooled[l, t0:t0 + n], ref) < 2e-5
            ref.backward(gp.view(F_, T, 256)[l, t0:t0 + n].double())
            assert rel_l2(gys[l][b].flatten(1), y.grad) < 2e-5
            e = emb.view(F_, T, 256)[l, t0:t0 + n - 1].double()
            ref_r = e.T @ m[:-1]
            assert rel_l2(rends[l][b].flatten(1), ref_r) < 2e-5 or float(ref_r.abs().max()) == 0.0
            ref_ge = m[:-1] @ gr_levels[l][b].flatten(1).double().T
            assert rel_l2(gemb[l, t0:t0 + n - 1], ref_ge) < 2e-5 or n == 1
            assert float(gemb[l, t0 + n - 1].abs().max()) == 0.0   # context row is not rendered
            t0 += n


@pytest.mark.parametrize("n_boxes", [[40, 3, 0], [1100, 3]])
def test_paint_kernels_crowded_images(n_boxes):
    """lgd_render_fwd / lgd_maskpool_bwd on crowded images: more rows than the paint kernel stages in shared memory
    (16), and an image with more rows than it keeps coverage masks for (1024: painted by the general kernel of the same
    launch) -- against the dense masks of lgd_masks_from_ranges in fp64."""
    B = len(n_boxes)
    bi, im, feats = synth.synth_batch(B, 96, 128, seed=9, n_boxes=n_boxes)
    H, W = im.tensor.shape[-2:]
    hws = [tuple(f.shape[-2:]) for f in feats.values()]
    g = engine.Geometry.get(B, hws, torch.device("cuda"))
    tb = engine.build_box_table(bi, H, W, True, torch.device("cuda"))
    T, F_ = tb.T, g.F
    ranges = torch.empty(F_ * T * 4, device="cuda", dtype=torch.int32)
    call("lgd_box_ranges", ptr(tb.boxes), T, H, W, g.pref, ptr(ranges))
    masks = torch.empty(T * g.P, device="cuda", dtype=torch.float32)
    call("lgd_masks_from_ranges", ptr(ranges), T, g.pref, ptr(masks))
    gen = torch.Generator().manual_seed(6)
    emb = torch.randn(F_ * T, 256, generator=gen)
    emb_c = emb.cuda()
    rend, rend_h, gy = g.new(), g.new_half(), g.new()
    call("lgd_render_fwd", g.pref, ptr(emb_c), ptr(ranges), ptr(tb.img_start), ptr(tb.n_render), T, ptr(rend), 0,
         ptr(rend_h))
    call("lgd_maskpool_bwd", g.pref, ptr(emb_c), ptr(ranges), ptr(tb.img_start), T, ptr(gy))
    torch.cuda.synchronize()
    assert torch.equal(rend_h, rend.half())
    rends, gys = pyr_to_nchw_cpu(g, rend), pyr_to_nchw_cpu(g, gy)
    masks = masks.cpu()
    off = 0
    for l, (h, w) in enumerate(g.hws):
        m_l = masks[off:off + T * h * w].view(T, h * w).double()
        off += T * h * w
        t0 = 0
        for b, n in enumerate(tb.counts):
            m = m_l[t0:t0 + n]
            e = emb.view(F_, T, 256)[l, t0:t0 + n].double()
            nr = int(tb.n_render[b])
            ref_r = e[:nr].T @ m[:nr]
            assert rel_l2(rends[l][b].flatten(1), ref_r) < 2e-5 or float(ref_r.abs().max()) == 0.0, (l, b)
            ref_g = (e / m.sum(-1).clamp(min=1.0)[:, None]).T @ m
            assert rel_l2(gys[l][b].flatten(1), ref_g) < 2e-5, (l, b)
            t0 += n


@pytest.mark.parametrize("ctx", [True, False])
def test_tap_render_matches_convolution(ctx):
    """lgd_tap_render_fwd / _bwd (local_inst_proj_2D over the piecewise-constant rendered map, evaluated from per-box tap
    vectors instead of a convolution) against F.conv2d of the rendered map + autograd in fp64: adversarial boxes (whole
    image, zero width / height, one pixel, border), an image without GT, per-(level, image) bias table."""
    bi, (H, W), g, _ = _box_setup()
    tb = engine.build_box_table(bi, H, W, ctx, torch.device("cuda"))
    T, F_, B = tb.T, g.F, g.B
    ranges = torch.empty(F_ * T * 4, device="cuda", dtype=torch.int32)
    call("lgd_box_ranges", ptr(tb.boxes), T, H, W, g.pref, ptr(ranges))
    gen = torch.Generator().manual_seed(17)
    emb = torch.randn(F_ * T, 256, generator=gen)
    wt = torch.randn(256, 256, 3, 3, generator=gen) * 0.05
    table = torch.randn(F_, B, 256, generator=gen)
    emb_c, wt_c, table_c = emb.cuda(), wt.cuda(), table.cuda().contiguous()
    out32, out_h = g.new(), g.new_half()
    ws = g.workspace(query("lgd_tap_render_workspace", g.pref, T, 1))
    call("lgd_tap_render_fwd", g.pref, ptr(emb_c), ptr(wt_c), ptr(ranges), ptr(tb.img_start), ptr(tb.n_render), T,
         tb.max_n, ptr(table_c), B * 256, 256, ptr(out_h), ptr(out32), ptr(ws), ws.numel())
    gout_levels = _rand_levels(B, g.hws, 23)
    gout = nchw_to_pyr(g, gout_levels)
    gemb = torch.full((F_ * T, 256), 7.0, device="cuda")
    gw = torch.full((256, 256, 3, 3), 7.0, device="cuda")
    call("lgd_tap_render_bwd", g.pref, ptr(gout), ptr(emb_c), ptr(wt_c), ptr(ranges), ptr(tb.img_of), ptr(tb.img_start),
         ptr(tb.n_render), T, ptr(gemb), ptr(gw), ptr(ws), ws.numel())
    torch.cuda.synchronize()
    # half copy: same values, and it doubles as the ReLU mask (positive never rounds to zero)
    assert torch.equal((out_h != 0), (out32 > 0))
    assert rel_l2(out_h.float().cpu(), out32.cpu()) < 1e-3
    outs = pyr_to_nchw_cpu(g, out32)
    per_img = O.prepare_boxes([x["instances"] for x in bi], H, W, ctx)
    e64 = emb.double().view(F_, T, 256).clone().requires_grad_(True)
    w64 = wt.double().clone().requires_grad_(True)
    total = 0.0
    n_render = tb.n_render.cpu().tolist()
    for l, (h, w) in enumerate(g.hws):
        t0 = 0
        for b, (boxes, _, _) in enumerate(per_img):
            n, nr = boxes.shape[0], n_render[b]
            m = O.inside_mask(boxes, (H, W), (h, w)).double()
            rendered = (e64[l, t0:t0 + nr].T @ m[:nr]).reshape(1, 256, h, w)
            pre = F.conv2d(rendered, w64, table[l, b].double(), padding=1)[0]
            assert rel_l2(outs[l][b], pre.detach().relu()) < 2e-5, (l, b)
            total = total + (pre * gout_levels[l][b].double()).sum()
            t0 += n
    ge_ref, gw_ref = torch.autograd.grad(total, [e64, w64])
    assert rel_l2(gemb.cpu().view(F_, T, 256), ge_ref) < 2e-5
    assert rel_l2(gw.cpu(), gw_ref) < 2e-5


def test_linear_layernorm_rowvec_segmax():
    gen = torch.Generator().manual_seed(21)
    T, K, N = 37, 84, 200
    x = torch.randn(T, K, generator=gen)
    w = torch.randn(N, K, generator=gen) / 9
    b = torch.randn(N, generator=gen)
    y = engine.linear(x.cuda(), w.cuda(), b.cuda())
    assert rel_l2(y, F.linear(x.double(), w.double(), b.double())) < 1e-6
    gy = torch.randn(T, N, generator=gen)
    gx, gw, gb = engine.linear_bwd(gy.cuda(), x.cuda(), w.cuda())
    assert rel_l2(gx, gy.double() @ w.double()) < 1e-6
    assert rel_l2(gw, gy.double().T @ x.double()) < 1e-6
    assert rel_l2(gb, gy.double().sum(0)) < 1e-6
    # the STN fc3 shapes: long contractions / wide outputs take the deterministic split-K path
    T2, K2, N2 = 130, 256, 7056
    x2 = torch.randn(T2, K2, generator=gen)
    w2 = torch.randn(N2, K2, generator=gen) / 16
    b2 = torch.randn(N2, generator=gen)
    gy2 = torch.randn(T2, N2, generator=gen)
    xc, wc, bc, gc2 = x2.cuda(), w2.cuda(), b2.cuda(), gy2.cuda()
    y2 = engine.linear(xc, wc, bc)
    gx2, gw2, gb2 = engine.linear_bwd(gc2, xc, wc)
    assert rel_l2(y2, F.linear(x2.double(), w2.double(), b2.double())) < 1e-6
    assert rel_l2(gx2, gy2.double() @ w2.double()) < 1e-6
    assert rel_l2(gw2, gy2.double().T @ x2.double()) < 1e-6
    assert rel_l2(gb2, gy2.double().sum(0)) < 1e-6
    y2b = engine.linear(xc, wc, bc)
    assert torch.equal(y2, y2b)   # split-K reduction order is fixed: run-to-run bit identical
    # layernorm + relu
    xd = (x * 3 + 1).double().requires_grad_(True)
    ref = F.layer_norm(xd, (K,), eps=1e-5).relu()
    gyl = torch.randn(T, K, generator=gen)
    ref.backward(gyl.double())
    yl, mean, rstd = engine.layernorm((x * 3 + 1).cuda(), True)
    gxl = engine.layernorm_bwd(gyl.cuda(), (x * 3 + 1).cuda(), mean, rstd, True)
    assert rel_l2(yl, ref) < 1e-6 and rel_l2(gxl, xd.grad) < 1e-5
    # row-vector x matrix
    mats = torch.randn(T, K * K, generator=gen)
    xv = x.double().requires_grad_(True)
    mv = mats.double().view(T, K, K).requires_grad_(True)
    refv = torch.bmm(xv.unsqueeze(1), mv).squeeze(1)
    gv = torch.randn(T, K, generator=gen)
    refv.backward(gv.double())
    yv = engine.rowvec_matmul(x.cuda(), mats.cuda(), K)
    gxv, gmv = engine.rowvec_matmul_bwd(gv.cuda(), x.cuda(), mats.cuda(), K)
    assert rel_l2(yv, refv) < 1e-6 and rel_l2(gxv, xv.grad) < 1e-6 and rel_l2(gmv.view(T, K, K), mv.grad) < 1e-6
    # per-image max + concat
    counts = [5, 1, 20, 11]
    start = torch.tensor([0, 5, 6, 26, 37], dtype=torch.int32).cuda()
    loc = torch.randn(T, 64, generator=gen)
    big = torch.randn(T, 1024, generator=gen).relu()
    cat = torch.empty(T, 1088, device="cuda")
    arg = torch.empty(4, 1024, device="cuda", dtype=torch.int32)
    loc_c, big_c = loc.cuda(), big.cuda()
    call("lgd_segmax_concat_fwd", ptr(loc_c), 64, ptr(big_c), 1024, ptr(start), 4, ptr(cat), ptr(arg))
    bd = big.double().requires_grad_(True)
    ld = loc.double().requires_grad_(True)
    pooled = torch.stack([c.max(0)[0] for c in bd.split(counts, 0)], 0)
    refc = torch.cat([ld, torch.cat([pooled[i:i + 1].expand(n, -1) for i, n in enumerate(counts)], 0)], 1)
    assert torch.equal(cat.cpu().double(), refc.detach())
    gc = torch.randn(T, 1088, generator=gen)
    refc.backward(gc.double())
    gl = torch.empty(T, 64, device="cuda")
    gb_ = torch.empty(T, 1024, device="cuda")
    gc_c = gc.cuda()
    call("lgd_segmax_concat_bwd", ptr(gc_c), 64, 1024, ptr(start), 4, ptr(arg), ptr(gl), ptr(gb_))
    assert rel_l2(gl, ld.grad) < 1e-6
    nz = big > 0   # ties only happen at relu zeros, where the downstream relu mask kills the gradient anyway
    assert rel_l2(gb_.cpu()[nz], bd.grad[nz]) < 1e-5


@pytest.mark.parametrize("pattern", ["stuGuided", "labelGuided"])
def test_attention_forward_backward(pattern):
    gen = torch.Generator().manual_seed(31)
    counts, Fl, E, heads = [4, 1, 9], 3, 256, 8
    T = sum(counts)
    img_of = torch.tensor(sum([[i] * n for i, n in enumerate(counts)], []), dtype=torch.int32)
    start = torch.tensor([0, 4, 5, 14], dtype=torch.int32)
    nq, nkv = (Fl, 1) if pattern == "stuGuided" else (1, Fl)
    q = torch.randn(nq * T, E, generator=gen)
    k = torch.randn(nkv * T, E, generator=gen)
    v = torch.randn(nkv * T, E, generator=gen)
    go = torch.randn(Fl * T, E, generator=gen)
    out = torch.empty(Fl * T, E, device="cuda")
    max_n = max(counts)
    probs = torch.empty(Fl * heads * T * max_n, device="cuda")
    # keep the device copies alive: ptr() only carries the address
    qc, kc, vc, goc, img_c, start_c = q.cuda(), k.cuda(), v.cuda(), go.cuda(), img_of.cuda(), start.cuda()
    call("lgd_attention_fwd", ptr(qc), nq, ptr(kc), ptr(vc), nkv, Fl, T, heads, E, ptr(img_c),
         ptr(start_c), max_n, ptr(out), ptr(probs))
    gq = torch.empty(Fl * T, E, device="cuda")
    gk = torch.empty(nkv * T, E, device="cuda")
    gv = torch.empty(nkv * T, E, device="cuda")
    gs = torch.empty_like(probs)
    call("lgd_attention_bwd", ptr(goc), ptr(qc), nq, ptr(kc), ptr(vc), nkv, Fl, T, heads, E,
         ptr(img_c), ptr(start_c), max_n, ptr(probs), ptr(gs), ptr(gq), ptr(gk), ptr(gv))
    qd, kd, vd = (t.double().requires_grad_(True) for t in (q, k, v))
    hd = E // heads
    outs = []
    for l in range(Fl):
        ql = qd[(l if nq > 1 else 0) * T:][:T].view(T, heads, hd).transpose(0, 1) * hd ** -0.5
        kl = kd[(l if nkv > 1 else 0) * T:][:T].view(T, heads, hd).transpose(0, 1)
        vl = vd[(l if nkv > 1 else 0) * T:][:T].view(T, heads, hd).transpose(0, 1)
        s = torch.bmm(ql, kl.transpose(1, 2)).masked_fill((img_of[:, None] != img_of[None, :])[None], float("-inf"))
        outs.append(torch.bmm(torch.softmax(s, -1), vl).transpose(0, 1).reshape(T, E))
    ref = torch.cat(outs, 0)
    ref.backward(go.double())
    assert rel_l2(out, ref) < 1e-5
    gq_got = gq[:T] if nq == 1 else gq
    assert rel_l2(gq_got, qd.grad) < 1e-5 and rel_l2(gk, kd.grad) < 1e-5 and rel_l2(gv, vd.grad) < 1e-5
