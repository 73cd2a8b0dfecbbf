"""CPU: the C-ABI library builds, loads, and exports every symbol include/casmtr_b200.h declares
(no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'casmtr_b200.h')).read()
    return sorted(set(re.findall(r'CASMTR_API[^;(]*?\b(casmtr_\w+)\s*\(', text)))


def test_header_declares_the_expected_surface():
    names = declared_symbols()
    for must in ('casmtr_score5d_fwd', 'casmtr_value_agg_fwd', 'casmtr_score3d_fwd', 'casmtr_qtatt_fwd',
                 'casmtr_cascade_qtatt_fwd', 'casmtr_cascade_match_fwd', 'casmtr_match_extract', 'casmtr_fine_match_fwd',
                 'casmtr_version', 'casmtr_last_error_string'):
        assert must in names


def test_library_exports_every_declared_symbol():
    from casmtr_b200 import build, _lib
    path = build.build()
    handle = ctypes.CDLL(path)
    for name in declared_symbols():
        assert hasattr(handle, name), f'{name} declared in include/casmtr_b200.h but not exported'
    assert set(_lib.SIGNATURES) == set(declared_symbols())      # the ctypes table covers the whole header
    assert _lib.lib().casmtr_version() == 100


def test_argument_validation_needs_no_gpu():
    """Validation runs before any CUDA call, so the error convention is testable on CPU."""
    from casmtr_b200 import _lib
    lib = _lib.lib()
    d = _lib.QtattDesc()
    d.B, d.nhead, d.D, d.levels, d.type = 1, 8, 64, 3, 0          # D=64 unsupported
    assert lib.casmtr_qtatt_workspace_bytes(ctypes.byref(d)) == 0
    assert b'head dim 64' in lib.casmtr_last_error_string()
    rc = lib.casmtr_score3d_fwd(None, None, None, None, 1, 4, 4, 6, 2, None)        # C % 4 != 0
    assert rc == -2
    with pytest.raises(_lib.CasmtrError):
        _lib.check(rc, 'casmtr_score3d_fwd')
    rc = lib.casmtr_cascade_match_fwd(None, None, None, None, None, None, 1.0, None, None, None, None, None, None, 1, 4, 4, 8, 4, 0, 0, None, 0, None)
    assert rc == -1                                                                  # null pointers


def test_missing_library_fails_loudly(monkeypatch):
    from casmtr_b200 import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', '/nonexistent/libcasmtr_b200.so')
    with pytest.raises(ImportError, match='no CPU fallback'):
        _lib.lib()
