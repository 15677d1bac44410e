"""HotPathDistillator: the distillation step in isolation (teacher forward -> distill loss [-> backward]) driven
through the same plugin classes train.py would build, with a null student standing in for the detectron2
detector. Used by bench.py, __graft_entry__.smoke() and the parity tests; SURVEY.md 8(d) defines the step."""
from __future__ import annotations

import torch
import torch.nn as nn

from .base_distillator import BaseDistillator
from .customized_detectors.build import CUSTOMIZED_DETECTORS_REGISTRY


class NullStudent(nn.Module):
    """Stands in for RetinaNetCT & co. (out of scope): the hot path only consumes the student's FPN maps."""

    def __init__(self, cfg):
        super().__init__()
        self.raw_backbone = nn.Identity()


if "NullStudent" not in getattr(CUSTOMIZED_DETECTORS_REGISTRY, "_obj_map", {}):
    CUSTOMIZED_DETECTORS_REGISTRY.register(NullStudent)


class HotPathDistillator(BaseDistillator):
    def __init__(self, cfg):
        cfg.MODEL.DISTILLATOR.STUDENT.META_ARCH = "NullStudent"
        super().__init__(cfg)
        self.distill_flag = 1

    def load_hot_path_state_dict(self, sd):
        """sd uses the reference's checkpoint names: teacher.*, adapter.distill.*"""
        own = self.state_dict()
        missing = [k for k in sd if k not in own]
        assert not missing, "unexpected keys: %s" % missing[:3]
        self.load_state_dict({k: v for k, v in sd.items()}, strict=False)

    def forward(self, batched_inputs, images, features):
        """returns (features_tea, inst_labels, masks, loss_distill)"""
        tea, inst_labels, masks = self.teacher((batched_inputs, images, None, features))
        loss = self.distill_loss({"stu": features, "tea": tea}, images, batched_inputs, masks, inst_labels)
        return tea, inst_labels, masks, loss["loss_distill"]

    def forward_student(self, batched_inputs, **kwargs):
        raise NotImplementedError

    def forward_teacher(self, batched_inputs, **kwargs):
        raise NotImplementedError

    def step(self, batched_inputs, images, features, cotangents=None):
        """One distillation step. With cotangents (stand-ins for the student-head gradient on the teacher pyramid):
        forward + backward of loss_distill + sum_l <tea_l, G_l>; parameter / feature grads land in .grad."""
        tea, inst_labels, masks, loss = self.forward(batched_inputs, images, features)
        if cotangents is not None:
            keys = list(tea.keys())
            torch.autograd.backward([loss] + [tea[k] for k in keys],
                                    [torch.ones_like(loss)] + [cotangents[k] for k in keys])
        return tea, loss
