"""Registries of the plugin surface. With detectron2 installed these ARE detectron2's registries, so
`train.py` of the reference finds the classes exactly where it looks for them (train.py:72-73,247-248);
without it (unit tests, the GPU box) a minimal fvcore-compatible Registry stands in."""
from __future__ import annotations

try:  # pragma: no cover - detectron2 is not in the build image
    from detectron2.modeling import META_ARCH_REGISTRY  # type: ignore
    from detectron2.utils.registry import Registry  # type: ignore
    HAVE_DETECTRON2 = True
except Exception:  # noqa: BLE001
    HAVE_DETECTRON2 = False

    class Registry:  # same surface as fvcore.common.registry.Registry
        def __init__(self, name: str):
            self._name = name
            self._obj_map = {}

        def _do_register(self, name, obj):
            assert name not in self._obj_map, "An object named '{}' was already registered in '{}' registry!".format(
                name, self._name)
            self._obj_map[name] = obj

        def register(self, obj=None):
            if obj is None:
                def deco(func_or_class):
                    self._do_register(func_or_class.__name__, func_or_class)
                    return func_or_class
                return deco
            self._do_register(obj.__name__, obj)
            return obj

        def get(self, name):
            ret = self._obj_map.get(name)
            if ret is None:
                raise KeyError("No object named '{}' found in '{}' registry!".format(name, self._name))
            return ret

        def __contains__(self, name):
            return name in self._obj_map

    META_ARCH_REGISTRY = Registry("META_ARCH")
