"""The C-ABI library builds, loads and exports every symbol include/ctb.h declares.
No compute calls (no GPU needed)."""
import ctypes
import os
import re

import pytest

import __graft_entry__ as G
from climate_toolbox_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    G.build()
    return _native.lib()


def _declared():
    src = open(os.path.join(ROOT, "include", "ctb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ctb_[a-z_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _declared() == sorted(_native.SYMBOLS)


def test_library_exports_every_declared_symbol(built):
    L = ctypes.CDLL(_native.LIB_PATH)
    for sym in _declared():
        assert hasattr(L, sym), sym
    assert built.ctb_version() == 100


def test_struct_layouts_match_header(built):
    # sizes follow from the field lists in include/ctb.h
    assert ctypes.sizeof(_native.PlanOpts) == 32 + 8            # 8 x int32 + the cell_gate pointer
    assert ctypes.sizeof(_native.AggOpts) == 8 + 8 + 4 + 4 + 8 + 8 + 8   # groups, t_begin, flush, n_peer_out, day_of_year, peer_out, peer_row
    assert _native.gate_word(0, 511, 0) == _native.GATE_ALWAYS
    assert ctypes.sizeof(_native.PlanInfo) == 4 * 8 + 2 * 4 + 2 * 8 + 8 * 4 + 2 * 8


def test_invalid_calls_fail_cleanly_without_gpu(built):
    # argument validation happens before any CUDA call
    rc = built.ctb_plan_get_info(None, None)
    assert rc == _native.ERR_INVALID and "null" in _native.last_error()
    with pytest.raises(ValueError):
        _native.check(rc)
    assert built.ctb_aggregate_workspace_bytes(None, 10, 1) == 0
    # ctb_push_rows: a column block that does not fit the leading dimension, too many peers
    assert built.ctb_push_rows(None, 10, 8, 4, 1, 1, None, 0, None) == _native.ERR_INVALID
    assert built.ctb_push_rows(None, 10, 0, 4, 1, _native.MAX_PEERS + 1, None, 0, None) == _native.ERR_INVALID
    assert built.ctb_push_rows(None, 10, 0, 0, 1, 1, None, 0, None) == 0          # nothing to copy


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(_native, "LIB_PATH", "/nonexistent/libctb.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _native.lib()


def test_no_cuda_no_fallback():
    import torch
    from climate_toolbox_b200 import _engine

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _engine.default_device()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "climate_toolbox_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle", src, flags=re.M), f
