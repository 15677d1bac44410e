"""Adapter plug-ins of the distillation loss: the registry, its builder and the stock 3x conv3x3 adapter."""
from .build import ADAPTERS_REGISTRY, build_adapter
from .sequential_convs import SequentialConvs

__all__ = ["ADAPTERS_REGISTRY", "build_adapter", "SequentialConvs"]
