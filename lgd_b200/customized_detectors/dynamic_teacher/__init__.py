"""The label-guided dynamic teacher (registry entry `DynamicTeacher`) and its parameter-holder sub-modules."""
from .dynamic_teacher import DynamicTeacher
from .label_encoder import STN, LabelEncoder

__all__ = ["DynamicTeacher", "LabelEncoder", "STN"]
