"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Imports the *unmodified* reference hot-path files from /root/reference through a ~30-line
detectron2 shim (detectron2 / fvcore / yacs / cvpods are not installed and there is no network).
Only usable in the build container (the GPU box has no /root/reference): it is used to
(1) validate oracle/lgd_oracle.py and (2) generate tests/golden/*.npz (oracle/make_golden.py).

Recipe verified in SURVEY.md Appendix A.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import torch
import torch.nn as nn

REF = os.environ.get("LGD_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "models", "customized_detectors", "dynamic_teacher"))


class _Registry:
    def __init__(self, name):
        self._name = name
        self._m = {}

    def register(self, obj=None):
        if obj is None:
            def deco(o):
                self._m[o.__name__] = o
                return o
            return deco
        self._m[obj.__name__] = obj
        return obj

    def get(self, n):
        return self._m[n]


_loaded = None


def load():
    """Returns namespace with DynamicTeacher, build_adapter, BaseDistillator from the reference."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF)

    def mod(name, **kw):
        m = types.ModuleType(name)
        m.__dict__.update(kw)
        sys.modules[name] = m
        return m

    def pkg(name, path):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
        return m

    for name in ("detectron2", "detectron2.utils", "detectron2.structures"):
        if name not in sys.modules:
            mod(name).__path__ = []
    mod("detectron2.utils.registry", Registry=_Registry)
    mod("detectron2.modeling", META_ARCH_REGISTRY=_Registry("META_ARCH"))
    from lgd_b200 import synth as _synth
    # pycocotools / detectron2 are absent: the reference's MASKS.polygons_to_bitmask (dynamic_teacher/utils.py:113) gets
    # the same deterministic stand-in rasteriser the engine under test is handed (lgd_b200.synth.polygons_to_bitmask)
    mod("detectron2.structures.masks", polygons_to_bitmask=_synth.polygons_to_bitmask)
    mod("detectron2.config", configurable=lambda f: f)

    pkg("models", REF + "/models")
    pkg("models.customized_detectors", REF + "/models/customized_detectors")
    build = importlib.import_module("models.customized_detectors.build")
    sys.modules["models.customized_detectors"].build_customized_detector = build.build_customized_detector
    dt = importlib.import_module("models.customized_detectors.dynamic_teacher.dynamic_teacher")
    pkg("models.adapters", REF + "/models/adapters")
    ab = importlib.import_module("models.adapters.build")
    sys.modules["models.adapters"].build_adapter = ab.build_adapter
    importlib.import_module("models.adapters.sequential_convs")
    bd = importlib.import_module("models.base_distillator")
    _loaded = types.SimpleNamespace(DynamicTeacher=dt.DynamicTeacher, build_adapter=ab.build_adapter,
                                    BaseDistillator=bd.BaseDistillator, dt_module=dt)
    return _loaded


class RefDistillator(nn.Module):
    """The reference's teacher + adapter + distill(), assembled without a student detector
    (BaseDistillator.__init__ would build one, base_distillator.py:19)."""

    def __init__(self, cfg, seed: int = 0):
        super().__init__()
        ref = load()
        torch.manual_seed(seed)
        self.teacher = ref.DynamicTeacher(cfg)
        adapter = ref.build_adapter(cfg)
        D = ref.BaseDistillator.__new__(ref.BaseDistillator)
        nn.Module.__init__(D)
        D.norm_stu = nn.InstanceNorm2d(256, affine=False)
        D.norm_tea = nn.InstanceNorm2d(256, affine=False)
        D.coef = cfg.MODEL.DISTILLATOR.LAMBDA
        D.adapter = nn.ModuleDict({"distill": adapter})
        D.distill_flag = 1
        D.teacher = self.teacher
        self.D = D

    def hot_path_state_dict(self):
        """state_dict with the reference's checkpoint names (teacher.*, adapter.distill.*)."""
        sd = {}
        for k, v in self.teacher.state_dict().items():
            sd["teacher." + k] = v.detach().clone()
        for k, v in self.D.adapter.state_dict().items():
            sd["adapter." + k] = v.detach().clone()
        return sd

    def step(self, batched_inputs, images, features, distill_flag=1):
        self.D.distill_flag = distill_flag
        tea, inst_labels, masks = self.teacher((batched_inputs, images, None, features))
        loss = self.D.distill_loss({"stu": features, "tea": tea}, images, batched_inputs, masks,
                                   inst_labels)["loss_distill"]
        return tea, inst_labels, masks, loss


_heads = None


def load_heads():
    """The reference's FCOSHead / POTOHead classes (thirdparty_heads/fcos.py:433-546, poto.py:523-625), imported from the
    unmodified files. cvpods and detectron2 are absent: the names those files import at module level are stubbed; the
    only one the head classes EXECUTE is ShiftGenerator(cfg, input_shape).num_cell_shifts (one shift per cell in every
    shipped FCOS / ATSS / POTO config) -- everything else belongs to the detectors' loss / inference code."""
    global _heads
    if _heads is not None:
        return _heads
    load()

    def mod(name, **kw):
        m = types.ModuleType(name)
        m.__dict__.update(kw)
        m.__path__ = []
        sys.modules[name] = m
        return m

    class ShiftGenerator:
        def __init__(self, cfg, input_shape):
            self.num_cell_shifts = [1] * len(input_shape)

    class ShapeSpec:
        def __init__(self, channels=None, height=None, width=None, stride=None):
            self.channels, self.height, self.width, self.stride = channels, height, width, stride

    def _unused(*a, **k):
        raise RuntimeError("stub of a cvpods / detectron2 function the head classes never call")

    mod("cvpods")
    mod("cvpods.modeling")
    mod("cvpods.modeling.anchor_generator", ShiftGenerator=ShiftGenerator)
    mod("cvpods.layers", ShapeSpec=ShapeSpec, cat=torch.cat, generalized_batched_nms=_unused)
    mod("cvpods.modeling.box_regression", Shift2BoxTransform=_unused)
    mod("cvpods.modeling.losses", iou_loss=_unused, sigmoid_focal_loss_jit=_unused, sigmoid_focal_loss=_unused)
    mod("cvpods.utils", comm=None, log_first_n=_unused)
    mod("cvpods.structures", Boxes=_unused, ImageList=_unused, Instances=_unused, pairwise_iou=_unused)
    sys.modules["detectron2.modeling"].build_backbone = _unused
    sys.modules["detectron2.structures"].ImageList = _unused
    sys.modules["detectron2.structures"].Instances = _unused
    sys.modules["detectron2.structures"].Boxes = _unused
    sys.modules["detectron2.structures"].pairwise_iou = _unused
    pkg = types.ModuleType("models.customized_detectors.thirdparty_heads")
    pkg.__path__ = [REF + "/models/customized_detectors/thirdparty_heads"]    # bypass its __init__ (imports all three)
    sys.modules["models.customized_detectors.thirdparty_heads"] = pkg
    fcos = importlib.import_module("models.customized_detectors.thirdparty_heads.fcos")
    try:
        poto = importlib.import_module("models.customized_detectors.thirdparty_heads.poto")
        poto_head = poto.POTOHead
    except Exception:  # noqa: BLE001  (poto.py imports more of cvpods; the head class is all that is needed)
        poto_head = None
    _heads = types.SimpleNamespace(FCOSHead=fcos.FCOSHead, POTOHead=poto_head, ShapeSpec=ShapeSpec)
    return _heads
