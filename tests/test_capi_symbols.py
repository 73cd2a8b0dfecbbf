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
    assert _lib.lib().casmtr_version() == 200


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


def test_argument_validation_of_the_widening_entries():
    """Same convention for the SURVEY 8f entry points: status code + message, before any CUDA call."""
    from casmtr_b200 import _lib
    lib = _lib.lib()
    assert lib.casmtr_window_idx_fwd(None, None, 1, 16, 4, 4, 5, None) == -1                 # 5-window does not fit a 4x4 grid
    assert b'window' in lib.casmtr_last_error_string()
    assert lib.casmtr_window_idx_fwd(None, None, 1, 16, 8, 8, 4, None) == -1                 # even window
    rc = lib.casmtr_cascade_qtatt_window_fwd(None, None, None, None, 7, None, None, None, 1, 4, 32, 16, 16, 16, 16, 0, None, 0, None)
    assert rc == -2 and b'window 7' in lib.casmtr_last_error_string()                        # 49 candidates > 32
    rc = lib.casmtr_cascade_qtatt_window_fwd(None, None, None, None, 5, None, None, None, 1, 4, 32, 16, 16, 8, 8, 0, None, 0, None)
    assert rc == -1                                                                          # parent grid 4x4 cannot hold a 5-window
    assert lib.casmtr_fine_window_gather(None, None, None, None, 4, 64, 32, 32, 16, 2, 4, None) == -1     # even W
    assert lib.casmtr_fine_match_dev_fwd(None, None, None, None, None, 1.0, None, None, None, 8, 25, 64, None) == -1    # no count
    assert lib.casmtr_pack_matches_dev(None, None, None, None, None, None, None, 0, 8, None, None) == -1
    assert lib.casmtr_coarse_match_workspace_bytes(1, 0, 16, 256) == 0
    one8 = ctypes.cast(ctypes.pointer(ctypes.c_float(0)), ctypes.c_void_p)
    rc = lib.casmtr_coarse_match_masked_fwd(one8, one8, one8, None, 0.1, one8, one8, one8, one8, 1, 64, 64, 64, one8, 1 << 20, None)
    assert rc == -1 and b'both masks' in lib.casmtr_last_error_string()
    d = _lib.QtattDesc()
    d.B, d.nhead, d.D, d.levels, d.type = 1, 2, 32, 2, 0
    d.qh[0] = d.qw[0] = d.kh[0] = d.kw[0] = 16
    d.qh[1] = d.qw[1] = d.kh[1] = 8
    d.kw[1] = 6                                                                              # not the 2x2 pooling of level 0
    d.topks[0] = d.topks[1] = 4
    one = ctypes.c_float(0)
    p = ctypes.cast(ctypes.pointer(one), ctypes.c_void_p)
    rc = lib.casmtr_qtatt_tokens_fwd(ctypes.byref(d), p, p, p, p, p, None, None, p, 1 << 30, None)
    assert rc == -1 and b'2x level' in lib.casmtr_last_error_string()        # check_qtatt_desc already refuses it
    prev = lib.casmtr_set_pdl(0)
    assert lib.casmtr_set_pdl(prev) == 0 and lib.casmtr_set_pdl(prev) == prev


def test_argument_validation_of_the_relative_pe_entries():
    from casmtr_b200 import _lib
    lib = _lib.lib()
    one = ctypes.c_float(0)
    p = ctypes.cast(ctypes.pointer(one), ctypes.c_void_p)
    d = _lib.RelpeDesc()
    d.w_table, d.h_table, d.tgt_idx = p, p, p
    d.n_emb, d.LB, d.h8, d.w8, d.w8_other = 22, 10, 6, 8, 9
    assert lib.casmtr_relative_pe_fwd(None, p, p, 1, 2, 12, 16, 25, None) == -1                 # no descriptor
    assert lib.casmtr_relative_pe_fwd(ctypes.byref(d), p, p, 1, 2, 12, 18, 25, None) == -1      # w0 != w8 * s
    assert b'1/8 grid' in lib.casmtr_last_error_string()
    assert lib.casmtr_relative_pe_fwd(ctypes.byref(d), p, p, 1, 2, 18, 24, 25, None) == -2      # s = 3
    assert b'must be even' in lib.casmtr_last_error_string()
    args = (1, 2, 32, 12, 16, 14, 20, 25, 0, None, 0, None)                                      # w1 = 20 != w8_other * s = 18
    assert lib.casmtr_cascade_qtatt_relpe_fwd(p, p, p, p, None, 0, ctypes.byref(d), p, None, *args) == -1
    assert b'w8_other' in lib.casmtr_last_error_string()
    args = (1, 2, 32, 12, 16, 14, 18, 25, 0, None, 0, None)
    assert lib.casmtr_cascade_qtatt_relpe_fwd(p, p, p, p, p, 5, ctypes.byref(d), p, None, *args) == -1      # topk_pos AND next_idx
    assert lib.casmtr_cascade_qtatt_relpe_fwd(p, p, p, None, p, 7, ctypes.byref(d), p, None, *args) == -2   # window 7
    assert lib.casmtr_cascade_qtatt_relpe_fwd(p, p, p, p, None, 0, ctypes.byref(d), p, None, *args) == -1   # passes the PE checks, no workspace
    assert b'null pointer' in lib.casmtr_last_error_string()


def test_missing_library_fails_loudly(monkeypatch):
    from casmtr_b200 import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', '/nonexistent/libcasmtr_b200.so')
    with pytest.raises(ImportError, match='no CPU fallback'):
        _lib.lib()


def test_host_mirror_of_the_relative_pe_module_needs_no_gpu():
    """CascadeRelativePE carries the reference's parameter names and LB rule (transformer.py:356-362); constructing it and loading a
    state dict is pure host logic."""
    import torch
    from casmtr_b200.modules.attention_layers import CascadeRelativePE
    m = CascadeRelativePE(4, window_size=5, sr_ratio=2)
    assert m.LB == 10 and m.h_pos_bias.weight.shape == (22, 4) and m.w_pos_bias.weight.shape == (22, 4)
    assert CascadeRelativePE(2, 5, 4).LB == 30 and CascadeRelativePE(2, 5, 4).w_pos_bias.weight.shape == (64, 2)
    m.load_state_dict({'h_pos_bias.weight': torch.zeros(22, 4), 'w_pos_bias.weight': torch.ones(22, 4)})
    with pytest.raises(RuntimeError):                    # CPU tensors are refused, nothing falls back
        m.get_relative_pe({'hw0_8c': (6, 8), 'hw1_8c': (6, 8), 'stage_8c': {'next_idx_c01': torch.zeros(1, 48, dtype=torch.long)}},
                          12, torch.zeros(1, 48, 25, 2, dtype=torch.long), None, 0)


def test_tuning_switches_round_trip():
    from casmtr_b200 import _lib
    lib = _lib.lib()
    prev = lib.casmtr_set_pdl(0)
    assert lib.casmtr_set_pdl(prev) == 0 and lib.casmtr_set_pdl(prev) == prev
    # the switches belong to the calling thread: another thread still sees the default
    import threading
    seen = []
    lib.casmtr_set_pdl(0)
    t = threading.Thread(target=lambda: seen.append(lib.casmtr_set_pdl(1)))
    t.start(); t.join()
    assert seen == [1] and lib.casmtr_set_pdl(prev) == 0


def test_header_is_plain_c_and_links(tmp_path):
    """include/casmtr_b200.h is the contract a non-Python host binds: it must compile as strict C99 and a C program must link
    against the library with nothing but -lcasmtr_b200 (no torch, no C++ runtime symbols leaking into the interface)."""
    import shutil
    import subprocess
    from casmtr_b200 import build
    gcc = shutil.which('gcc')
    if gcc is None:
        pytest.skip('no gcc')
    lib = build.build()
    src = tmp_path / 't.c'
    src.write_text('#include "casmtr_b200.h"\n'
                   'int main(void) {\n'
                   '    casmtr_relpe_desc pe; casmtr_qtatt_desc qd; casmtr_extract_desc ed;\n'
                   '    (void)pe; (void)qd; (void)ed;\n'
                   '    if (casmtr_version() != CASMTR_VERSION) return 1;\n'
                   '    /* argument validation needs no device: a null descriptor is refused with a message */\n'
                   '    if (casmtr_qtatt_workspace_bytes((const casmtr_qtatt_desc *)0) != 0) return 2;\n'
                   '    return casmtr_last_error_string()[0] ? 0 : 3;\n'
                   '}\n')
    exe = tmp_path / 't'
    inc = os.path.join(ROOT, 'include')
    r = subprocess.run([gcc, '-std=c99', '-Wall', '-Wextra', '-pedantic', '-Werror', '-I', inc, str(src), '-o', str(exe),
                        '-L', os.path.dirname(lib), '-lcasmtr_b200', '-Wl,-rpath,' + os.path.dirname(lib)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stderr)
