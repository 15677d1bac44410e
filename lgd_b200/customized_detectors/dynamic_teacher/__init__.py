from .dynamic_teacher import DynamicTeacher

__all__ = [k for k in globals().keys() if not k.startswith('_')]
