"""ADAPTERS_REGISTRY + build_adapter: same hook API as the reference (models/adapters/build.py:10-17)."""
import torch

from ..registry import Registry

ADAPTERS_REGISTRY = Registry("ADAPTERS")
ADAPTERS_REGISTRY.__doc__ = ""


def build_adapter(cfg):
    """Instantiate the adapter class named by cfg.MODEL.DISTILLATOR.ADAPTER.META_ARCH on cfg.MODEL.DEVICE."""
    adapter_cls = ADAPTERS_REGISTRY.get(cfg.MODEL.DISTILLATOR.ADAPTER.META_ARCH)
    return adapter_cls(cfg).to(torch.device(cfg.MODEL.DEVICE))
