"""Detector registry of the plug-in surface. The student detectors stay with detectron2 / cvpods and register
themselves here from the reference's own files; this package contributes the hot-path entry `DynamicTeacher`."""
from .build import CUSTOMIZED_DETECTORS_REGISTRY, build_customized_detector
from .dynamic_teacher import DynamicTeacher

__all__ = ["CUSTOMIZED_DETECTORS_REGISTRY", "build_customized_detector", "DynamicTeacher"]
