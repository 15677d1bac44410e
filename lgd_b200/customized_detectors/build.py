"""CUSTOMIZED_DETECTORS_REGISTRY + build_customized_detector (models/customized_detectors/build.py:11-17,40-43).
The student detectors (RetinaNetCT, FCOSCT, ...) stay with detectron2/cvpods and register themselves into this
registry from the reference's own files; this package registers the hot-path entry `DynamicTeacher`."""
import torch

from ..registry import META_ARCH_REGISTRY, Registry  # noqa: F401

CUSTOMIZED_DETECTORS_REGISTRY = Registry("CUSTOMIZED_DETECTORS")
CUSTOMIZED_DETECTORS_REGISTRY.__doc__ = ""


def get_model(cfg, meta_arch):
    model = CUSTOMIZED_DETECTORS_REGISTRY.get(meta_arch)(cfg)
    model = model.to(torch.device(cfg.MODEL.DEVICE))
    return model


def build_customized_detector(cfg):
    student = get_model(cfg, cfg.MODEL.DISTILLATOR.STUDENT.META_ARCH)
    teacher = get_model(cfg, cfg.MODEL.DISTILLATOR.TEACHER.META_ARCH)
    return student, teacher
