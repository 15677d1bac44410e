"""CPU: the C-ABI shared library loads and exports exactly what include/lgd_b200.h declares, and the ctypes table in
lgd_b200/_lib.py agrees with the header (names and argument counts). No compute calls (no GPU needed)."""
import ctypes
import os
import re

import pytest

from lgd_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_prototypes():
    src = open(os.path.join(ROOT, "include", "lgd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(lgd_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        n = 0 if args in ("", "void") else args.count(",") + 1
        protos[name] = n
    return protos


def test_library_is_built_and_loads():
    assert os.path.exists(_lib.LIB_PATH), "run `python -c 'import __graft_entry__ as g; g.build()'` first"
    lib = _lib.load()
    assert lib.lgd_version() >= 100
    assert lib.lgd_launch_count() == 0 or lib.lgd_launch_count() > 0


def test_every_declared_symbol_is_exported_and_bound():
    protos = _header_prototypes()
    assert len(protos) >= 40
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name, nargs in protos.items():
        assert hasattr(raw, name), "liblgd_b200.so does not export %s" % name
        assert name in _lib.SIGNATURES, "%s is declared in the header but missing from _lib.SIGNATURES" % name
        assert len(_lib.SIGNATURES[name][1]) == nargs, "%s: header has %d arguments, ctypes table %d" % (
            name, nargs, len(_lib.SIGNATURES[name][1]))
    extra = set(_lib.SIGNATURES) - set(protos)
    assert not extra, "bound but not declared in include/lgd_b200.h: %s" % sorted(extra)


def test_argument_validation_happens_on_the_host():
    """Bad arguments are rejected before any launch, with an error string, never an abort (SURVEY 8(b) 'errors')."""
    lib = _lib.load()
    pyr = _lib.Pyramid.make(1, [(4, 4)])
    rc = lib.lgd_gn_apply(ctypes.byref(pyr), None, None, None, 0, 0, None, None, None, 0, None)
    assert rc == -1 and b"null pointer" in lib.lgd_last_error()
    bad = _lib.Pyramid.make(1, [(4, 4)])
    bad.num_levels = 0
    assert lib.lgd_conv3x3_num_tiles(ctypes.byref(bad)) < 0
    assert lib.lgd_pyramid_elems(ctypes.byref(pyr)) == 4 * 4 * 256
    assert lib.lgd_conv3x3_num_tiles(ctypes.byref(_lib.Pyramid.make(2, [(100, 168), (7, 11)]))) == 2 * (133 + 1)   # tiles of 128 slots of the zero-padded rows: ceil(100*170/128) + ceil(7*13/128) per image
    with pytest.raises(ValueError):
        _lib.Pyramid.make(1, [(1, 1)] * 9)
