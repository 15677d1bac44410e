from .build import ADAPTERS_REGISTRY, build_adapter
from .sequential_convs import SequentialConvs

__all__ = [k for k in globals().keys() if not k.startswith('_')]
