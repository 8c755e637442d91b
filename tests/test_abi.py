"""The C-ABI library loads and exports every symbol include/vitk.h declares; host-only entry points
(layout / workspace queries) agree with the state_dict contract.  No GPU, no compute calls."""
import ctypes
import os
import re

import pytest

import __graft_entry__ as ge
from oracle import vit_oracle as vo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    ge.build()
    from vit_spoof_detection_pda_b200 import _lib
    return _lib


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "vitk.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vitk_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    raw = ctypes.CDLL(lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/vitk.h but not exported"
    assert set(names) == set(lib.PROTOTYPES), "ctypes prototypes out of sync with include/vitk.h"


def test_version_and_error_string(lib):
    l = lib.load()
    assert l.vitk_version() == 101
    assert l.vitk_is_dev_build() == 0
    assert l.vitk_trace_start(None, 0) != 0          # the tracer exists only in libvitk_dev.so
    assert l.vitk_debug_set(7, 1) != 0               # the release build has no result-invalidating knob
    assert isinstance(l.vitk_last_error_string(), bytes)


@pytest.mark.parametrize("depth", [1, 2, 12])
def test_param_layout_matches_state_dict_contract(lib, depth):
    total, offs, sizes = lib.param_layout(depth, 2)
    spec = vo.expected_state_dict_spec(depth)
    assert len(offs) == len(spec)
    cur = 0
    for (name, shape), off, n in zip(spec, offs, sizes):
        numel = 1
        for s in shape:
            numel *= s
        assert n == numel, name
        assert off == cur and off % 64 == 0, name
        cur += (numel + 63) // 64 * 64
    assert total == cur


def test_workspace_bytes_monotone(lib):
    l = lib.load()
    ev = l.vitk_workspace_bytes(8, 12, lib.PREC_BF16, 0)
    tr = l.vitk_workspace_bytes(8, 12, lib.PREC_BF16, 1)
    fz = l.vitk_workspace_bytes(8, 12, lib.PREC_BF16, 3)
    tr32 = l.vitk_workspace_bytes(8, 12, lib.PREC_FP32, 1)
    assert 0 < ev < tr < tr32 and ev <= fz < tr
    # bs 64 training activations must fit comfortably in 180 GB
    assert l.vitk_workspace_bytes(64, 12, lib.PREC_BF16, 1) < 16e9


def test_missing_library_fails_loudly(lib, monkeypatch, tmp_path):
    monkeypatch.setattr(lib, "_lib", None)
    monkeypatch.setattr(lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU/PyTorch fallback"):
        lib.load()
