"""The five META_ARCH wrappers (lgd_b200/distillator.py, mirroring models/distillator.py:23-494) executed end to end with
a mock student detector: training forward (student losses + '*.tea' losses of the student head on the teacher features
+ loss_distill), backward into student, teacher and adapter, and the eval path with and without eval_teacher. The
expected values are composed from the CPU oracle: teacher pyramid -> the same mock head -> the same mock loss.

Also here: the stand-alone SequentialConvs call (the adapters/ hook API, sequential_convs.py:13-15) and the generic
adapter path of BaseDistillator.distill (any registered adapter module + the fused InstanceNorm-MSE node) against
plain PyTorch fp32 on the same device."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

import lgd_b200
from lgd_b200 import synth
from lgd_b200.adapters.build import ADAPTERS_REGISTRY
from lgd_b200.customized_detectors.build import CUSTOMIZED_DETECTORS_REGISTRY
from oracle import lgd_oracle as O

pytestmark = pytest.mark.gpu

KEYS = list(synth.LEVEL_KEYS)
B, IMG_H, IMG_W = 2, 120, 150


class _MockStudent(nn.Module):
    """Stands in for RetinaNetCT / RCNNCT / FCOSCT / POTOCT / ATSSCT (customized_detectors/*.py): learnable FPN maps, a
    1x1-conv head, quadratic losses. Return conventions per detector family as the reference's students have them."""
    KIND = "retinanet"

    def __init__(self, cfg):
        super().__init__()
        gen = torch.Generator().manual_seed(7)
        hws = synth.pyramid_hw(synth.pad32(IMG_H), synth.pad32(IMG_W))
        self.maps = nn.ParameterDict({k: nn.Parameter(torch.randn(B, 256, h, w, generator=gen)) for k, (h, w) in zip(KEYS, hws)})
        self.raw_backbone = nn.Linear(2, 2)     # train.py:206 pokes student.raw_backbone.parameters()
        self.head_cls = nn.Conv2d(256, 6, 1)
        self.head_box = nn.Conv2d(256, 4, 1)
        self.head_in_features = KEYS            # RetinaNetCT
        self.in_features = KEYS                 # FCOSCT / POTOCT / ATSSCT
        self.images = synth.ImageList(torch.empty(B, 3, synth.pad32(IMG_H), synth.pad32(IMG_W), device="meta"),
                                      [(IMG_H, IMG_W)] * B)

    # ---- head + losses (what the '*.tea' losses reuse)
    def _heads(self, feats):
        return [self.head_cls(f) for f in feats], [self.head_box(f) for f in feats]

    @staticmethod
    def _quad(cls, box):
        return {"loss_cls": sum((c ** 2).mean() for c in cls), "loss_box_reg": sum((b ** 2).mean() for b in box) * 0.5}

    def predict(self, *args):
        if self.KIND == "rcnn":                 # RCNNCT.predict(features, images, gt_instances, batched_inputs) -> losses
            feats = args[0]
            cls, box = self._heads([feats[k] for k in KEYS])
            return self._quad(cls, box)
        cls, box = self._heads(args[0])
        if self.KIND == "retinanet":
            return "anchors", cls, box
        if self.KIND == "poto":
            return "shifts", cls, box
        return "shifts", cls, box, [b[:, :1] for b in box]      # fcos / atss: + centerness

    def losses(self, *args):
        if self.KIND == "retinanet":            # (anchors, logits, gt_labels, deltas, gt_boxes)
            return self._quad(args[1], args[3])
        if self.KIND == "poto":                 # (gt_classes, gt_shifts_reg_deltas, box_cls, box_delta)
            return self._quad(args[2], args[3])
        d = self._quad(args[3], args[4])        # fcos / atss: (gt_classes, gt_shifts, gt_centerness, cls, delta, center)
        d["loss_centerness"] = sum((c ** 2).mean() for c in args[5]) * 0.25
        return d

    def inference(self, *args, **kw):
        if self.KIND == "rcnn":                 # RCNNCT.inference(batched_inputs, features=...) -> (results, ...)
            cls, _ = self._heads([kw["features"][k] for k in KEYS])
            return [float(sum(c.mean() for c in cls))], None
        cls = args[1] if self.KIND == "retinanet" else args[0]
        return float(sum(c.mean() for c in cls))

    def get_processed_results(self, results, batched_inputs, images):
        return [results]

    def forward(self, batched_inputs):
        feats = {k: v * 1.0 for k, v in self.maps.items()}
        if self.training:
            cls, box = self._heads([feats[k] for k in KEYS])
            losses = self._quad(cls, box)
            if self.KIND in ("fcos", "atss"):
                losses["loss_centerness"] = sum((b[:, :1] ** 2).mean() for b in box) * 0.25
            targets = ("gt_labels", "gt_boxes") if self.KIND in ("retinanet", "poto") else ("a", "b", "c")
            if self.KIND == "rcnn":
                targets = "gt_instances"
            return losses, None, feats, self.images, targets
        cls, _ = self._heads([feats[k] for k in KEYS])
        return [float(sum(c.mean() for c in cls))], None, feats, self.images


def _student_cls(kind):
    name = "MockStudent_" + kind
    if name not in CUSTOMIZED_DETECTORS_REGISTRY:
        CUSTOMIZED_DETECTORS_REGISTRY.register(type(name, (_MockStudent,), {"KIND": kind}))
    return name


ARCHS = [("DistillatorRetinaNet", "retinanet", True), ("DistillatorGeneralizedRCNN", "rcnn", True),
         ("DistillatorFCOS", "fcos", False), ("DistillatorPOTO", "poto", False), ("DistillatorATSS", "atss", False)]


def _build(arch, kind, ctx):
    cfg = synth.make_cfg(device="cuda", add_context_box=ctx)
    cfg.MODEL.DISTILLATOR.STUDENT.META_ARCH = _student_cls(kind)
    model = lgd_b200.META_ARCH_REGISTRY.get(arch)(cfg)
    sd = synth.synth_state_dict(5)
    missing = model.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys
    assert all(k.startswith("student.") for k in missing.missing_keys), missing.missing_keys
    model = model.cuda()
    model.distill_flag = 1
    return model, sd


@pytest.mark.parametrize("arch,kind,ctx", ARCHS)
def test_distillator_training_forward_backward(arch, kind, ctx):
    model, sd = _build(arch, kind, ctx)
    model.train()
    bi, _, _ = synth.synth_batch(B, IMG_H, IMG_W, seed=11)
    losses = model(bi)
    stu = model.student
    base = {"loss_cls", "loss_box_reg"} | ({"loss_centerness"} if kind in ("fcos", "atss") else set())
    assert set(losses) == base | {k + ".tea" for k in base} | {"loss_distill"}
    # expected values: CPU oracle teacher / distill loss, the same mock head on the oracle's teacher pyramid
    feats = {k: v.detach().cpu() for k, v in stu.maps.items()}
    tea_o, _, _, loss_o, _ = O.distill_step(sd, bi, stu.images, feats, add_context_box=ctx)
    with torch.no_grad():
        cls, box = stu._heads([tea_o[k].cuda() for k in KEYS])
        exp = _MockStudent._quad(cls, box)
        if kind in ("fcos", "atss"):
            exp["loss_centerness"] = sum((b[:, :1] ** 2).mean() for b in box) * 0.25
    assert abs(float(losses["loss_distill"]) - float(loss_o)) <= 1e-3 * float(loss_o)
    for k, v in exp.items():
        assert abs(float(losses[k + ".tea"]) - float(v)) <= 2e-3 * abs(float(v)) + 1e-7, (k, float(losses[k + ".tea"]), float(v))
    total = sum(losses.values())
    assert bool(torch.isfinite(total))
    total.backward()
    for n, p in model.named_parameters():
        if n.startswith("student.raw_backbone") or (not ctx and "global_ctx_proj_1D" in n):
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, n
        else:
            assert p.grad is not None and bool(torch.isfinite(p.grad).all()) and float(p.grad.abs().max()) > 0, n


@pytest.mark.parametrize("arch,kind,ctx", ARCHS)
def test_distillator_eval_paths(arch, kind, ctx):
    model, sd = _build(arch, kind, ctx)
    model.eval()
    bi, _, _ = synth.synth_batch(B, IMG_H, IMG_W, seed=12)
    stu = model.student
    with torch.no_grad():
        plain = model(bi)
        with_tea = model(bi, eval_teacher=True)
        cls_s, _ = stu._heads([stu.maps[k] for k in KEYS])
        feats = {k: v.detach().cpu() for k, v in stu.maps.items()}
        tea_o, _, _, _, _ = O.distill_step(sd, bi, stu.images, feats, add_context_box=ctx)
        cls_t, _ = stu._heads([tea_o[k].cuda() for k in KEYS])
    exp_s = float(sum(c.mean() for c in cls_s))
    exp_t = float(sum(c.mean() for c in cls_t))
    assert abs(plain[0] - exp_s) <= 1e-5 * max(1.0, abs(exp_s))
    got = with_tea[0] if not isinstance(with_tea, float) else with_tea
    assert abs(got - exp_t) <= 2e-3 * max(1.0, abs(exp_t)), (got, exp_t)
    assert abs(exp_t - exp_s) > 1e-3      # the two paths are distinguishable


def test_standalone_sequential_convs_matches_torch():
    """adapter(x) as the hook API calls it (base_distillator.py:57): TF32 tensor-core operands, so 1e-3 against fp32."""
    cfg = synth.make_cfg(device="cuda")
    a = lgd_b200.build_adapter(cfg)
    ref = nn.Sequential(nn.Conv2d(256, 256, 3, 1, 1), nn.ReLU(), nn.Conv2d(256, 256, 3, 1, 1), nn.ReLU(),
                        nn.Conv2d(256, 256, 3, 1, 1)).cuda()
    ref.load_state_dict(a.adapter.state_dict())
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(2, 256, 25, 42, generator=gen).cuda().requires_grad_(True)
    xr = x.detach().clone().requires_grad_(True)
    go = torch.randn(2, 256, 25, 42, generator=gen).cuda()
    y = a(x)
    yr = ref(xr)
    assert float((y - yr).norm() / yr.norm()) < 1e-3
    y.backward(go)
    yr.backward(go)
    assert float((x.grad - xr.grad).norm() / xr.grad.norm()) < 5e-2     # two ReLU layers below: flips
    for (n, p), (_, pr) in zip(a.adapter.named_parameters(), ref.named_parameters()):
        assert float((p.grad - pr.grad).norm() / pr.grad.norm()) < 5e-2, n
    # the last layer sits above every ReLU: operand rounding only
    assert float((a.adapter[4].weight.grad - ref[4].weight.grad).norm() / ref[4].weight.grad.norm()) < 1e-3
    assert float((a.adapter[4].bias.grad - ref[4].bias.grad).norm() / ref[4].bias.grad.norm()) < 1e-3


class _Scale1x1(nn.Module):
    """A user-registered adapter (the adapters/ hook API, adapters/build.py:10-17): plain PyTorch module."""

    def __init__(self, cfg):
        super().__init__()
        self.conv = nn.Conv2d(256, 256, 1)

    def forward(self, x):
        return self.conv(x)


def test_generic_adapter_path_matches_torch():
    if "_Scale1x1" not in ADAPTERS_REGISTRY:
        ADAPTERS_REGISTRY.register(_Scale1x1)
    cfg = synth.make_cfg(device="cuda")
    cfg.MODEL.DISTILLATOR.ADAPTER.META_ARCH = "_Scale1x1"
    from lgd_b200.step import HotPathDistillator
    torch.manual_seed(0)
    m = HotPathDistillator(cfg).cuda()
    assert type(m.adapter["distill"]) is _Scale1x1
    gen = torch.Generator().manual_seed(5)
    hws = [(20, 24), (10, 12), (5, 6)]
    stu = {k: torch.randn(2, 256, h, w, generator=gen).cuda().requires_grad_(True) for k, (h, w) in zip(KEYS, hws)}
    tea = {k: torch.randn(2, 256, h, w, generator=gen).cuda() for k, (h, w) in zip(KEYS, hws)}
    loss = m.distill_loss({"stu": stu, "tea": tea}, None, None, None, None)["loss_distill"]
    loss.backward()
    got_w = m.adapter["distill"].conv.weight.grad.clone()
    got_x = {k: v.grad.clone() for k, v in stu.items()}
    m.zero_grad()
    sr = {k: v.detach().clone().requires_grad_(True) for k, v in stu.items()}
    s_list = [F.instance_norm(m.adapter["distill"](sr[k]), eps=1e-5).reshape(2, -1) for k in sorted(sr)]
    t_list = [F.instance_norm(tea[k], eps=1e-5).reshape(2, -1) for k in sorted(sr)]
    ref = m.coef * F.mse_loss(torch.cat(t_list, 1), torch.cat(s_list, 1))
    ref.backward()
    assert abs(float(loss) - float(ref)) <= 1e-5 * float(ref)
    ref_w = m.adapter["distill"].conv.weight.grad
    assert float((got_w - ref_w).norm() / ref_w.norm()) < 1e-3
    for k in sr:
        assert float((got_x[k] - sr[k].grad).norm() / sr[k].grad.norm()) < 1e-3, k
