from .build import CUSTOMIZED_DETECTORS_REGISTRY, build_customized_detector
from .dynamic_teacher import DynamicTeacher

__all__ = [k for k in globals().keys() if not k.startswith('_')]
