"""lgd_b200: B200-native engine for the distillation hot path of LGD (megvii-research/LGD).

Importing the package registers the same plugin names as the reference's `models` package
(train.py:72-73): META_ARCH `Distillator{RetinaNet,GeneralizedRCNN,FCOS,POTO,ATSS}`,
CUSTOMIZED_DETECTORS `DynamicTeacher`, ADAPTERS `SequentialConvs`."""
from . import adapters, customized_detectors  # noqa: F401
from .adapters import ADAPTERS_REGISTRY, SequentialConvs, build_adapter  # noqa: F401
from .base_distillator import BaseDistillator  # noqa: F401
from .customized_detectors import CUSTOMIZED_DETECTORS_REGISTRY, DynamicTeacher, build_customized_detector  # noqa: F401
from . import distillator  # noqa: F401
from .registry import META_ARCH_REGISTRY  # noqa: F401

__version__ = "0.1.0"
