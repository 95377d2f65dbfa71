"""The C-ABI library loads without a GPU and exports every symbol that
include/xinv.h declares (no compute calls here)."""
import ctypes
import os
import re

from xinvert_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "xinv.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(xinv_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_loads():
    path = build.build()
    assert os.path.exists(path)
    L = _lib.load()
    assert L.xinv_version() == 102


def test_every_declared_symbol_is_exported_and_bound():
    declared = _declared_symbols()
    assert len(declared) >= 25
    L = ctypes.CDLL(build.build())
    for name in declared:
        assert hasattr(L, name), f"{name} declared in xinv.h but not exported"
    bound = {s[0] for s in _lib.SYMBOLS}
    assert set(declared) == bound, set(declared) ^ bound


def test_opts_struct_layout_matches_header():
    # int32 x6 + int64 x8
    assert ctypes.sizeof(_lib.XinvOpts) == 6 * 4 + 8 * 8 + 2 * 4
    o = _lib.make_opts(ordering="lex", engine="fused", coef_strides=[0, -1, 5])
    assert (o.ordering, o.engine, o.coef_stride[0], o.coef_stride[1], o.coef_stride[2], o.coef_stride[7]) == \
        (1, 2, 0, -1, 5, -1)


def test_no_gpu_means_loud_failure_not_fallback():
    import numpy as np
    import pytest
    import xinvert_b200 as xb
    if xb.device_count() > 0:
        pytest.skip("a GPU is visible")
    S = np.zeros((8, 8))
    with pytest.raises(xb.XinvError):
        xb.solve_standard_2D(S, S + 1, None, S + 1, S, "fixed", "fixed", 1.0, 0.25, 1.0, 1.4)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under xinvert_b200/ may import it."""
    pkg = os.path.join(ROOT, "xinvert_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
                assert "sor_oracle" not in src, fn
