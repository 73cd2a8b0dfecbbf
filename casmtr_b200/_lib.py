"""ctypes binding of libcasmtr_b200.so (include/casmtr_b200.h).

There is NO CPU fallback: if the shared library is missing or a call fails the
caller gets an exception.  Build it with ``python -m casmtr_b200.build`` (or
``__graft_entry__.build()``).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libcasmtr_b200.so')
MAX_LEVELS = 4
K_COUNT = 11          # CASMTR_K_COUNT

c_float_p = C.c_void_p      # raw device addresses (tensor.data_ptr())
c_i64_p = C.c_void_p
c_u8_p = C.c_void_p


class QtattDesc(C.Structure):
    _fields_ = [('B', C.c_int), ('nhead', C.c_int), ('D', C.c_int), ('levels', C.c_int), ('type', C.c_int),
                ('qh', C.c_int * MAX_LEVELS), ('qw', C.c_int * MAX_LEVELS),
                ('kh', C.c_int * MAX_LEVELS), ('kw', C.c_int * MAX_LEVELS),
                ('topks', C.c_int * MAX_LEVELS), ('flags', C.c_int), ('weight_len', C.c_int), ('concurrent_calls', C.c_int)]


QT_NO_OVERLAP, QT_SIMT_COARSE = 1, 2        # casmtr_qtatt_desc.flags


class ExtractDesc(C.Structure):
    _fields_ = [('B', C.c_int), ('h0', C.c_int), ('w0', C.c_int), ('h1', C.c_int), ('w1', C.c_int),
                ('nms_window', C.c_int), ('test_thr', C.c_float), ('border_rm', C.c_int), ('double_check', C.c_int),
                ('n_pre', C.c_int), ('pre_conf', C.c_void_p * 2), ('pre_h', C.c_int * 2), ('pre_w', C.c_int * 2),
                ('pre_thr', C.c_float * 2), ('pad_mask0', C.c_void_p), ('pad_mask1', C.c_void_p),
                ('scale', C.c_float), ('scale0', C.c_void_p), ('scale1', C.c_void_p), ('coarse_mode', C.c_int)]


class RelpeDesc(C.Structure):
    _fields_ = [('w_table', C.c_void_p), ('h_table', C.c_void_p), ('tgt_idx', C.c_void_p), ('n_emb', C.c_int), ('LB', C.c_int),
                ('h8', C.c_int), ('w8', C.c_int), ('w8_other', C.c_int)]


# name -> (restype, argtypes); every symbol include/casmtr_b200.h declares
SIGNATURES = {
    'casmtr_version': (C.c_int, []),
    'casmtr_last_error_string': (C.c_char_p, []),
    'casmtr_device_info': (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_size_t)]),
    'casmtr_launch_count': (C.c_uint64, []),
    'casmtr_profile_enable': (C.c_int, [C.c_int]),
    'casmtr_profile_collect': (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    'casmtr_kernel_kind_name': (C.c_char_p, [C.c_int]),
    'casmtr_plan_dense_tiles': (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    'casmtr_fastdiv': (C.c_int, [C.c_int, C.POINTER(C.c_uint)]),
    'casmtr_score5d_fwd': (C.c_int, [c_float_p, c_float_p, c_i64_p, c_float_p] + [C.c_int] * 6 + [C.c_void_p]),
    'casmtr_value_agg_fwd': (C.c_int, [c_float_p, c_float_p, c_i64_p, c_float_p] + [C.c_int] * 6 + [C.c_void_p]),
    'casmtr_score3d_fwd': (C.c_int, [c_float_p, c_float_p, c_i64_p, c_float_p] + [C.c_int] * 5 + [C.c_void_p]),
    'casmtr_score5d_bwd': (C.c_int, [c_float_p, c_float_p, c_float_p, c_i64_p, c_float_p, c_float_p] + [C.c_int] * 6 + [C.c_void_p]),
    'casmtr_value_agg_bwd': (C.c_int, [c_float_p, c_float_p, c_float_p, c_i64_p, c_float_p, c_float_p] + [C.c_int] * 6 + [C.c_void_p]),
    'casmtr_score3d_bwd': (C.c_int, [c_float_p, c_float_p, c_float_p, c_i64_p, c_float_p, c_float_p] + [C.c_int] * 5 + [C.c_void_p]),
    'casmtr_nchw_to_tokens': (C.c_int, [c_float_p, c_float_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'casmtr_qtatt_workspace_bytes': (C.c_size_t, [C.POINTER(QtattDesc)]),
    'casmtr_qtatt_fwd': (C.c_int, [C.POINTER(QtattDesc), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                   c_float_p, c_float_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                   C.c_void_p, C.c_size_t, C.c_void_p]),
    'casmtr_qtatt_tokens_fwd': (C.c_int, [C.POINTER(QtattDesc), c_float_p, c_float_p, c_float_p,
                                          c_float_p, c_float_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                          C.c_void_p, C.c_size_t, C.c_void_p]),
    'casmtr_cascade_qtatt_tokens_fwd': (C.c_int, [c_float_p, c_float_p, c_float_p, c_i64_p, c_float_p, c_float_p, c_i64_p]
                                        + [C.c_int] * 9 + [C.c_void_p, C.c_size_t, C.c_void_p]),
    'casmtr_qtatt_guided_workspace_bytes': (C.c_size_t, [C.c_int] * 8),
    'casmtr_qtatt_guided_fwd': (C.c_int, [c_float_p, c_float_p, c_float_p, c_i64_p, c_float_p, C.c_int, c_float_p] + [C.c_int] * 8
                                + [C.c_void_p, C.c_size_t, C.c_void_p]),
    'casmtr_set_pdl': (C.c_int, [C.c_int]),
    'casmtr_set_overlap': (C.c_int, [C.c_int]),
    'casmtr_window_idx_fwd': (C.c_int, [c_i64_p, c_i64_p] + [C.c_int] * 5 + [C.c_void_p]),
    'casmtr_cascade_qtatt_window_fwd': (C.c_int, [c_float_p, c_float_p, c_float_p, c_i64_p, C.c_int, c_float_p, c_float_p, c_i64_p]
                                        + [C.c_int] * 8 + [C.c_void_p, C.c_size_t, C.c_void_p]),
    'casmtr_relative_pe_fwd': (C.c_int, [C.POINTER(RelpeDesc), c_i64_p, c_float_p] + [C.c_int] * 5 + [C.c_void_p]),
    'casmtr_cascade_qtatt_relpe_fwd': (C.c_int, [c_float_p, c_float_p, c_float_p, c_i64_p, c_i64_p, C.c_int, C.POINTER(RelpeDesc), c_float_p, c_i64_p]
                                       + [C.c_int] * 9 + [C.c_void_p, C.c_size_t, C.c_void_p]),
    'casmtr_cascade_qtatt_workspace_bytes': (C.c_size_t, [C.c_int] * 6),
    'casmtr_cascade_qtatt_fwd': (C.c_int, [c_float_p, c_float_p, c_float_p, c_i64_p, c_float_p, c_float_p, c_i64_p]
                                 + [C.c_int] * 9 + [C.c_void_p, C.c_size_t, C.c_void_p]),
    'casmtr_cascade_match_fwd': (C.c_int, [c_float_p, c_float_p, c_i64_p, c_i64_p, c_u8_p, c_u8_p, C.c_float,
                                           c_float_p, c_float_p, c_i64_p, c_float_p, c_float_p, c_i64_p]
                                 + [C.c_int] * 7 + [C.c_void_p, C.c_size_t, C.c_void_p]),
    'casmtr_cascade_match_workspace_bytes': (C.c_size_t, [C.c_int] * 3),
    'casmtr_coarse_match_workspace_bytes': (C.c_size_t, [C.c_int] * 4),
    'casmtr_coarse_match_fwd': (C.c_int, [c_float_p, c_float_p, C.c_float, c_float_p, c_i64_p, c_float_p, c_i64_p]
                                + [C.c_int] * 4 + [C.c_void_p, C.c_size_t, C.c_void_p]),
    'casmtr_coarse_match_masked_fwd': (C.c_int, [c_float_p, c_float_p, c_u8_p, c_u8_p, C.c_float, c_float_p, c_i64_p, c_float_p, c_i64_p]
                                       + [C.c_int] * 4 + [C.c_void_p, C.c_size_t, C.c_void_p]),
    'casmtr_coarse_match_mutual_fwd': (C.c_int, [c_float_p, c_float_p, c_u8_p, c_u8_p, C.c_float, c_float_p, c_i64_p, c_float_p, c_i64_p,
                                                 c_float_p, c_i64_p, c_i64_p] + [C.c_int] * 4 + [C.c_void_p, C.c_size_t, C.c_void_p]),
    'casmtr_match_extract_workspace_bytes': (C.c_size_t, [C.POINTER(ExtractDesc)]),
    'casmtr_match_extract': (C.c_int, [C.POINTER(ExtractDesc), c_float_p, c_i64_p, c_i64_p, c_u8_p, c_i64_p, c_i64_p, c_i64_p,
                                       c_float_p, c_float_p, c_float_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    'casmtr_pack_matches': (C.c_int, [c_i64_p, c_i64_p, c_i64_p, c_float_p, c_float_p, c_float_p, C.c_int, C.c_int64, C.c_int,
                                      C.c_void_p, C.c_void_p]),
    'casmtr_pack_matches_dev': (C.c_int, [c_i64_p, c_i64_p, c_i64_p, c_float_p, c_float_p, c_float_p, C.c_void_p, C.c_int64, C.c_int,
                                          C.c_void_p, C.c_void_p]),
    'casmtr_fine_match_dev_fwd': (C.c_int, [c_float_p, c_float_p, c_float_p, c_float_p, c_i64_p, C.c_float,
                                            c_float_p, c_float_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'casmtr_fine_window_gather': (C.c_int, [c_float_p, c_i64_p, c_i64_p, c_float_p] + [C.c_int] * 7 + [C.c_void_p]),
    'casmtr_fine_match_fwd': (C.c_int, [c_float_p, c_float_p, c_float_p, c_float_p, c_i64_p, C.c_float,
                                        c_float_p, c_float_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
}

_lib = None


class CasmtrError(RuntimeError):
    """A libcasmtr_b200 call returned a negative status (mirrors the RuntimeError the reference's
    TORCH_CHECKs raise, cuda_imp/QuadTreeAttention/QuadtreeAttention/src/score_computation.cpp:6-8)."""


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f'{LIB_PATH} is missing: the CUDA library has not been built. Run '
                '`python -m casmtr_b200.build` (needs nvcc). There is no CPU fallback.')
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError if the symbol is not exported
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(status, what):
    if status != 0:
        msg = lib().casmtr_last_error_string()
        raise CasmtrError(f'{what} failed ({status}): {msg.decode() if msg else "?"}')
