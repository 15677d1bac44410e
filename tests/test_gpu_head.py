"""the student's RetinaNet head on the teacher pyramid (lgd_b200/heads.py) against the CPU
restatement of detectron2's RetinaNetHead + permute_to_N_HWA_K (oracle.lgd_oracle.retinanet_head): forward within 1e-3;
gradients of the output convolutions (above every ReLU) within 1e-3, everything else against the oracle evaluated with
the engine's activation pattern (oracle/parity.py explains why) within 2e-3; and the zero-copy hand-over of the input
gradient to the teacher backward."""
import pytest
import torch
import torch.nn as nn

from lgd_b200 import synth
from lgd_b200.heads import PARAM_NAMES, RetinaNetHeadB200
from oracle import lgd_oracle as O

pytestmark = pytest.mark.gpu

A, K = 9, 80


def _make_head(seed=3):
    torch.manual_seed(seed)
    def tower():
        return nn.Sequential(*[m for _ in range(4) for m in (nn.Conv2d(256, 256, 3, 1, 1), nn.ReLU())])
    head = nn.Module()
    head.cls_subnet, head.bbox_subnet = tower(), tower()
    head.cls_score = nn.Conv2d(256, A * K, 3, 1, 1)
    head.bbox_pred = nn.Conv2d(256, A * 4, 3, 1, 1)
    return head


def _rel(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("B,hw", [(2, (96, 128)), (1, (200, 264))])
def test_retinanet_head_matches_oracle(B, hw):
    from lgd_b200 import engine
    head = _make_head().cuda()
    b200 = RetinaNetHeadB200.from_module(head)
    assert RetinaNetHeadB200.supports(head)
    sd = {k: v.detach().cpu() for k, v in head.state_dict().items()}
    hws = synth.pyramid_hw(synth.pad32(hw[0]), synth.pad32(hw[1]))
    g = engine.Geometry.get(B, hws, torch.device("cuda", 0))
    gen = torch.Generator().manual_seed(11)
    pyr = torch.randn(g.elems, generator=gen).cuda().requires_grad_(True)     # an NHWC pyramid buffer like the teacher's
    feats = g.level_views(pyr)
    logits, deltas = b200(feats)
    # the loss side concatenates the levels per image (detectron2 losses): gradients arrive as strided slices
    cl, cd = torch.cat(logits, 1), torch.cat(deltas, 1)
    G1 = (torch.randn(cl.shape, generator=gen) * 1e-3).cuda()
    G2 = (torch.randn(cd.shape, generator=gen) * 1e-3).cuda()
    ((cl * G1).sum() + (cd * G2).sum()).backward()
    torch.cuda.synchronize()
    S = logits[0].grad_fn.S
    # activation patterns of the eight tower ReLUs as the engine took them (fp16 copies: nonzero = pass)
    force = {}
    for tower, tag in (("cls_subnet", "cls"), ("bbox_subnet", "box")):
        for k, i in enumerate((0, 2, 4, 6)):
            for l, v in enumerate(g.level_views(S.acts[tower][k + 1])):
                force["%s%d/%d" % (tag, i, l)] = (v != 0).cpu()
    fo = [f.detach().cpu().contiguous().requires_grad_(True) for f in feats]
    sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    lo, do = O.retinanet_head(sdo, fo, A, K, relu_ctl={"force": force})
    with torch.no_grad():
        lp, dp = O.retinanet_head(sd, [f.detach() for f in fo], A, K)
    for l in range(len(hws)):
        assert logits[l].shape == lp[l].shape and deltas[l].shape == dp[l].shape
        assert _rel(logits[l], lp[l]) < 1e-3, (l, _rel(logits[l], lp[l]))
        assert _rel(deltas[l], dp[l]) < 1e-3, (l, _rel(deltas[l], dp[l]))
    tot = (torch.cat(lo, 1) * G1.cpu()).sum() + (torch.cat(do, 1) * G2.cpu()).sum()
    names = sorted(sdo)
    grads = torch.autograd.grad(tot, fo + [sdo[n] for n in names])
    own = dict(b200.named_parameters())
    worst = {}
    for n, gr in zip(names, grads[len(fo):]):
        worst[n] = _rel(own[n].grad, gr)
    gx = pyr.grad
    for l, (v, gr) in enumerate(zip(g.level_views(gx), grads[:len(fo)])):
        worst["feat%d" % l] = _rel(v, gr)
    print("head gradient errors vs the pattern-evaluated oracle:", sorted(worst.items(), key=lambda kv: -kv[1])[:6])
    for n in ("cls_score.weight", "cls_score.bias", "bbox_pred.weight", "bbox_pred.bias"):
        assert worst[n] < 1e-3, (n, worst[n])
    assert max(worst.values()) < 2e-3, sorted(worst.items(), key=lambda kv: -kv[1])[:5]


def test_head_backward_feeds_teacher_backward_in_place():
    """teacher pyramid -> head -> loss: the head's input gradient arrives at _TeacherFn.backward as level views of one
    NHWC buffer and is read in place (no NCHW round trip); the result equals the same step with the head run by plain
    PyTorch convolutions on the same device (fp32, TF32 off) within the step's gradient bars."""
    from tests.gpu_util import make_model
    sd = synth.synth_state_dict(5)
    bi, im, feats = synth.synth_batch(2, 120, 150, seed=31)
    head = _make_head(5).cuda()
    b200 = RetinaNetHeadB200.from_module(head)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    res = {}
    for mode in ("b200", "torch"):
        m = make_model(dict(add_context_box=True), sd, 1)
        head.zero_grad()
        f = {k: v.cuda().requires_grad_(True) for k, v in feats.items()}
        tea, _, _, loss = m.forward(bi, im, f)
        keys = list(tea.keys())
        if mode == "b200":
            logits, deltas = b200([tea[k] for k in keys])
        else:
            logits, deltas = [], []
            for k in keys:
                x = tea[k]
                n = x.shape[0]
                logits.append(head.cls_score(head.cls_subnet(x)).view(n, A, K, *x.shape[-2:]).permute(0, 3, 4, 1, 2).reshape(n, -1, K))
                deltas.append(head.bbox_pred(head.bbox_subnet(x)).view(n, A, 4, *x.shape[-2:]).permute(0, 3, 4, 1, 2).reshape(n, -1, 4))
        total = loss + (torch.cat(logits, 1) ** 2).mean() + (torch.cat(deltas, 1) ** 2).mean()
        total.backward()
        torch.cuda.synchronize()
        res[mode] = (float(total), {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None},
                     {n: p.grad.clone() for n, p in head.named_parameters()}, {k: v.grad.clone() for k, v in f.items()})
    a, b = res["b200"], res["torch"]
    assert abs(a[0] - b[0]) <= 1e-3 * abs(b[0])
    for n in b[2]:
        assert _rel(a[2][n], b[2][n]) < 5e-2, (n, _rel(a[2][n], b[2][n]))
    for n in ("cls_score.weight", "bbox_pred.weight"):
        assert _rel(a[2][n], b[2][n]) < 2e-3, (n, _rel(a[2][n], b[2][n]))
    for n in b[1]:
        if n.endswith("adapter.4.bias"):
            continue
        assert _rel(a[1][n], b[1][n]) < 8e-2, (n, _rel(a[1][n], b[1][n]))
    for k in b[3]:
        assert _rel(a[3][k], b[3][k]) < 8e-2, k


# ---------------------------------------------------------------------------------------------------- FCOS family
class _Scale(nn.Module):      # thirdparty_heads/scale.py:10-16
    def __init__(self):
        super().__init__()
        self.scale = nn.Parameter(torch.ones(1))

    def forward(self, x):
        return x * self.scale


def _make_fcos_like_head(sd, strides, centerness_on_reg, norm_reg_targets):
    """A module with the attribute / parameter layout of the reference's FCOSHead (thirdparty_heads/fcos.py:438-501) or,
    without 'centerness.*' in sd, POTOHead (poto.py:528-590), holding the given weights."""
    def tower():
        return nn.Sequential(*[m for _ in range(4) for m in (nn.Conv2d(256, 256, 3, 1, 1), nn.GroupNorm(32, 256), nn.ReLU())])
    head = nn.Module()
    head.cls_subnet, head.bbox_subnet = tower(), tower()
    head.cls_score = nn.Conv2d(256, sd["cls_score.weight"].shape[0], 3, 1, 1)
    head.bbox_pred = nn.Conv2d(256, 4, 3, 1, 1)
    if "centerness.weight" in sd:
        head.centerness = nn.Conv2d(256, 1, 3, 1, 1)
    head.scales = nn.ModuleList([_Scale() for _ in strides])
    head.fpn_strides, head.centerness_on_reg, head.norm_reg_targets = list(strides), centerness_on_reg, norm_reg_targets
    head.load_state_dict(sd)
    return head


def _fcos_patterns(S, g):
    """activation patterns of the eight GroupNorm-ReLUs as the engine took them (fp16 copies: nonzero = pass)"""
    force = {}
    for tower, tag in (("cls_subnet", "cls"), ("bbox_subnet", "box")):
        for k, i in enumerate((0, 3, 6, 9)):
            for l, v in enumerate(g.level_views(S.layers[tower][k].out)):
                force["%s%d/%d" % (tag, i, l)] = (v != 0).cpu()
    return force


def _run_fcos_case(sd, feats, cots, strides, ctr_on_reg, norm_reg):
    from lgd_b200 import engine
    from lgd_b200.heads import FCOSHeadB200
    head = _make_fcos_like_head(sd, strides, ctr_on_reg, norm_reg).cuda()
    assert FCOSHeadB200.supports(head)
    b200 = FCOSHeadB200(head)
    fx = [f.clone().cuda().requires_grad_(True) for f in feats]
    outs = b200(fx)
    total = sum((o * c.cuda()).sum() for group, cg in zip(outs, cots) for o, c in zip(group, cg))
    total.backward()
    torch.cuda.synchronize()
    S = outs[0][0].grad_fn.S
    force = _fcos_patterns(S, S.g)
    if norm_reg:
        for l, o in enumerate(outs[1]):
            force["reg/%d" % l] = (o > 0).cpu()     # the ReLU of the box decoding (fcos.py:540)
    # the oracle with the engine's activation pattern (kernel accuracy) and the plain one (forward values)
    fo = [f.clone().requires_grad_(True) for f in feats]
    sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    oo = O.fcos_head(sdo, fo, strides, ctr_on_reg, norm_reg, relu_ctl={"force": force})
    oo = [x for x in oo if x is not None]
    tot = sum((o * c).sum() for group, cg in zip(oo, cots) for o, c in zip(group, cg))
    names = sorted(sdo)
    grads = torch.autograd.grad(tot, fo + [sdo[n] for n in names])
    own = dict(head.named_parameters())
    worst = {n: _rel(own[n].grad, gr) for n, gr in zip(names, grads[len(fo):]) if not n.startswith("scales.")}
    # the per-level Scale gradients are scalars -- cotangent-weighted sums that may cancel to almost nothing on a level --
    # so they are judged together, against the largest of them
    sc = [(own[n].grad.double().cpu(), gr.double()) for n, gr in zip(names, grads[len(fo):]) if n.startswith("scales.")]
    worst["scales"] = float(max((a - b).abs().max() for a, b in sc) / max(b.abs().max() for _, b in sc))
    for l, gr in enumerate(grads[:len(fo)]):
        worst["feat%d" % l] = _rel(fx[l].grad, gr)
    return outs, worst


# Forward bars. Five chained convolutions on 10-bit-mantissa operands put every output at 6-7e-4 of the fp32 reference
# (a CPU emulation of the operand rounding gives 6.5e-4 for the logits). The 80-channel logits are held to 1e-3; the
# box and centerness outputs are the same pre-activations seen through relu * stride / exp over 4 and 1 channels -- a
# ReLU halves the reference norm (x 1.4 in relative terms) and on the small test levels the statistic is over a few
# hundred values -- and are held to 2e-3 (measured 0.9-1.3e-3).
FWD_TOL = {"logits": 1e-3, "bbox_reg": 2e-3, "centerness": 2e-3}


def _check_fcos_grads(worst):
    """every gradient within 2e-3 of the pattern-evaluated oracle; the Scale scalars (sums of a few hundred products on the
    small test levels, through the relu * stride decoding) within 5e-3 of the largest of them"""
    for n, e in worst.items():
        assert e < (5e-3 if n == "scales" else 2e-3), sorted(worst.items(), key=lambda kv: -kv[1])[:5]


@pytest.mark.parametrize("name", ["fcos_head_ctr_on_reg", "fcos_head_ctr_on_cls_exp", "poto_head"])
def test_fcos_family_head_matches_reference_golden(name):
    """lgd_b200.heads.FCOSHeadB200 against the outputs of the reference's own FCOSHead / POTOHead classes
    (tests/golden/<name>.npz, thirdparty_heads/fcos.py:433-546, poto.py:523-625): every output within 1e-3; every
    gradient against the oracle evaluated with the engine's activation pattern within 2e-3 (oracle/parity.py)."""
    from tests.golden_util import load_head_case
    g, (cls, ctr_on_reg, norm_reg, B, hws, strides, _, _), sd, feats, cots = load_head_case(name)
    outs, worst = _run_fcos_case(sd, feats, cots, strides, ctr_on_reg, norm_reg)
    assert len(outs) == (2 if cls == "POTOHead" else 3)
    for gi, gname in enumerate(("logits", "bbox_reg", "centerness")[:len(outs)]):
        for l, o in enumerate(outs[gi]):
            ref = torch.from_numpy(g["%s_%d" % (gname, l)])
            assert tuple(o.shape) == tuple(ref.shape)
            assert _rel(o, ref) < FWD_TOL[gname], (gname, l, _rel(o, ref))
    print(name, "gradient errors vs the pattern-evaluated oracle:", sorted(worst.items(), key=lambda kv: -kv[1])[:6])
    _check_fcos_grads(worst)


def test_fcos_head_at_pyramid_size():
    """the same on a five-level pyramid of a 200x264 image, B = 2 (wide and narrow levels, partial tiles)"""
    strides = [8, 16, 32, 64, 128]
    sd = synth.synth_fcos_head_state_dict(7, 5, 80, True)
    hws = synth.pyramid_hw(synth.pad32(200), synth.pad32(264))
    gen = torch.Generator().manual_seed(17)
    feats = [torch.randn(2, 256, h, w, generator=gen) for (h, w) in hws]
    cots = [[torch.randn(2, c, h, w, generator=gen) * 1e-3 for (h, w) in hws] for c in (80, 4, 1)]
    outs, worst = _run_fcos_case(sd, feats, cots, strides, True, True)
    with torch.no_grad():
        ref = O.fcos_head(sd, feats, strides, True, True)
    for gname, group, rg in zip(("logits", "bbox_reg", "centerness"), outs, ref):
        for l, (o, r) in enumerate(zip(group, rg)):
            assert _rel(o, r) < FWD_TOL[gname], (gname, l, _rel(o, r))
    print("gradient errors vs the pattern-evaluated oracle:", sorted(worst.items(), key=lambda kv: -kv[1])[:6])
    _check_fcos_grads(worst)


def _torch_fcos_forward(head, feats):
    """FCOSHead.forward (fcos.py:528-546) with the module's own layers: the stock PyTorch path of the student"""
    logits, bbox_reg, ctr = [], [], []
    for l, x in enumerate(feats):
        c, b = head.cls_subnet(x), head.bbox_subnet(x)
        logits.append(head.cls_score(c))
        ctr.append(head.centerness(b if head.centerness_on_reg else c))
        pred = head.scales[l](head.bbox_pred(b))
        bbox_reg.append(torch.relu(pred) * head.fpn_strides[l] if head.norm_reg_targets else torch.exp(pred))
    return logits, bbox_reg, ctr


def test_distillator_fcos_runs_the_b200_head_on_the_teacher_features():
    """DistillatorFCOS.forward_teacher (models/distillator.py:270-295) with a student that owns a stock FCOS head: the
    '*.tea' losses come from lgd_b200.heads.FCOSHeadB200 on the teacher pyramid (B200_HEAD) and agree with the student's
    own PyTorch head on the same features; gradients reach the head, the teacher and the student maps."""
    import lgd_b200
    from lgd_b200.customized_detectors.build import CUSTOMIZED_DETECTORS_REGISTRY
    from tests.test_gpu_distillators import B as NB, IMG_H, IMG_W, KEYS, _MockStudent

    class _FcosStudent(_MockStudent):
        KIND = "fcos"

        def __init__(self, cfg):
            super().__init__(cfg)
            self.head = _make_fcos_like_head(synth.synth_fcos_head_state_dict(9, len(KEYS), 80, True), [8, 16, 32, 64, 128],
                                             True, True)
            self.shift_generator = lambda feats: "shifts"

        def predict(self, feats):
            return ("shifts", *_torch_fcos_forward(self.head, feats))

        def losses(self, gt_classes, gt_shifts, gt_centerness, box_cls, box_delta, box_center):
            return {"loss_cls": sum((c ** 2).mean() for c in box_cls), "loss_box_reg": sum((b ** 2).mean() for b in box_delta) * 0.01,
                    "loss_centerness": sum((c ** 2).mean() for c in box_center) * 0.25}

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    if "FcosHeadStudent" not in CUSTOMIZED_DETECTORS_REGISTRY:
        CUSTOMIZED_DETECTORS_REGISTRY.register(type("FcosHeadStudent", (_FcosStudent,), {}))
    res = {}
    for b200 in (True, False):
        torch.manual_seed(1)      # the mock student's own 1x1 heads take PyTorch's default initialisation
        cfg = synth.make_cfg(device="cuda", add_context_box=False)
        cfg.MODEL.DISTILLATOR.STUDENT.META_ARCH = "FcosHeadStudent"
        model = lgd_b200.META_ARCH_REGISTRY.get("DistillatorFCOS")(cfg)
        model.load_state_dict(synth.synth_state_dict(5), strict=False)
        model = model.cuda().train()
        model.distill_flag = 1
        model.B200_HEAD = b200
        bi, _, _ = synth.synth_batch(NB, IMG_H, IMG_W, seed=13)
        losses = model(bi)
        assert {"loss_cls.tea", "loss_box_reg.tea", "loss_centerness.tea", "loss_distill"} <= set(losses)
        if b200:
            assert model.__dict__["_b200_head"]["head"] is not None, "the stock FCOS head must take the B200 path"
        sum(losses.values()).backward()
        torch.cuda.synchronize()
        res[b200] = ({k: float(v) for k, v in losses.items()},
                     {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None})
    a, b = res[True], res[False]
    for k in b[0]:
        assert abs(a[0][k] - b[0][k]) <= 2e-3 * abs(b[0][k]) + 1e-7, (k, a[0][k], b[0][k])
    assert set(a[1]) == set(b[1])
    for n in b[1]:
        if n.endswith("adapter.4.bias"):
            continue
        assert bool(torch.isfinite(a[1][n]).all()), n
        assert _rel(a[1][n], b[1][n]) < 0.15, (n, _rel(a[1][n], b[1][n]))      # ReLU-flip floor of the small test maps
    for n in ("student.head.cls_score.weight", "student.head.bbox_pred.weight", "student.head.centerness.weight"):
        assert _rel(a[1][n], b[1][n]) < 3e-3, (n, _rel(a[1][n], b[1][n]))
