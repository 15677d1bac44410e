"""ADAPTERS_REGISTRY + build_adapter: same hook API as the reference (models/adapters/build.py:10-17)."""
import torch

from ..registry import Registry

ADAPTERS_REGISTRY = Registry("ADAPTERS")
ADAPTERS_REGISTRY.__doc__ = ""


def build_adapter(cfg):
    meta_arch = cfg.MODEL.DISTILLATOR.ADAPTER.META_ARCH
    model = ADAPTERS_REGISTRY.get(meta_arch)(cfg)
    model = model.to(torch.device(cfg.MODEL.DEVICE))
    return model
