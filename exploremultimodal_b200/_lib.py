"""ctypes binding of libmome.so (the C ABI declared in include/mome.h).

The product path has no fallback: if the library is missing or a call fails, a RuntimeError with
the library's own error text is raised. Nothing here imports `oracle/`.
"""
import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libmome.so')

F32, BF16 = 0, 1
EPI_STORE, EPI_GELU, EPI_RESIDUAL, EPI_DGELU, EPI_ATOMIC = 0, 1, 2, 3, 4
K_MAJOR, MN_MAJOR = 0, 1
MAX_GROUPS = 4
ABI_VERSION = 6


class GemmGroup(C.Structure):
    _fields_ = [('a', C.c_void_p), ('b', C.c_void_p), ('M', C.c_int64), ('K', C.c_int64),
                ('out', C.c_void_p), ('out2', C.c_void_p), ('bias', C.c_void_p), ('res', C.c_void_p),
                ('aux', C.c_void_p), ('row0', C.c_int64), ('colsum', C.c_void_p)]


class GemmArgs(C.Structure):
    _fields_ = [('dtype', C.c_int32), ('a_major', C.c_int32), ('b_major', C.c_int32),
                ('epilogue', C.c_int32), ('out_dtype', C.c_int32), ('num_groups', C.c_int32),
                ('split_k', C.c_int32), ('reserved', C.c_int32), ('N', C.c_int64),
                ('lda', C.c_int64), ('ldb', C.c_int64), ('ldo', C.c_int64), ('ldo2', C.c_int64),
                ('ldres', C.c_int64), ('ldaux', C.c_int64), ('gamma', C.c_void_p),
                ('group', GemmGroup * MAX_GROUPS), ('drop_seed', C.c_void_p), ('row_scale', C.c_void_p),
                ('drop_salt', C.c_uint32), ('drop_p', C.c_float)]


class Dropout(C.Structure):
    """MomeDropout of include/mome.h."""
    _fields_ = [('seed', C.c_void_p), ('row_scale', C.c_void_p), ('row0', C.c_int64), ('salt', C.c_uint32), ('p', C.c_float)]


class BlockGroup(C.Structure):
    _fields_ = [('first_row', C.c_int64), ('rows', C.c_int64), ('w1', C.c_void_p), ('b1', C.c_void_p),
                ('w2', C.c_void_p), ('b2', C.c_void_p), ('dw1', C.c_void_p), ('db1', C.c_void_p),
                ('dw2', C.c_void_p), ('db2', C.c_void_p), ('colsum_part', C.c_void_p)]


class BlockArgs(C.Structure):
    """MomeBlockArgs of include/mome.h (field order is the ABI)."""
    _fields_ = ([('dtype', C.c_int32), ('num_heads', C.c_int32), ('num_groups', C.c_int32), ('num_seqs', C.c_int32),
                 ('max_seq_len', C.c_int32), ('reserved', C.c_int32), ('tokens', C.c_int64), ('d', C.c_int64),
                 ('hid', C.c_int64), ('eps', C.c_float), ('scale', C.c_float), ('seq_desc', C.c_void_p),
                 ('key_mask', C.c_void_p)]
                + [(n, C.c_void_p) for n in ('gamma_1', 'gamma_2', 'n1w', 'n1b', 'n2w', 'n2b', 'qkv_bias', 'proj_b',
                                             'w_qkv', 'w_proj')]
                + [('group', BlockGroup * MAX_GROUPS)]
                + [(n, C.c_void_p) for n in ('x', 'h', 'mean1', 'rstd1', 'qkv', 'o', 'lse', 'br1', 'x1', 'h2', 'mean2',
                                             'rstd2', 'gp', 'u', 'br2', 'x2', 'dx2', 'dx', 'dgamma_1', 'dgamma_2',
                                             'dn1w', 'dn1b', 'dn2w', 'dn2b', 'dq_bias', 'dv_bias', 'dproj_b', 'dw_qkv',
                                             'dw_proj', 's_dbr2', 's_dh2', 's_dbr1', 's_do', 's_dh', 's_dz', 's_dqkv',
                                             's_dx1', 's_delta', 'ws')]
                + [('ws_bytes', C.c_size_t), ('drop_seed', C.c_void_p), ('row_sample', C.c_void_p),
                   ('row_scale1', C.c_void_p), ('row_scale2', C.c_void_p), ('drop_salt', C.c_uint32),
                   ('p_attn', C.c_float), ('p_hidden', C.c_float), ('p_branch', C.c_float), ('p_path', C.c_float)])


_P, _I, _L, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_float
_SIGNATURES = {
    'mome_version': (C.c_int, []),
    'mome_last_error': (C.c_char_p, []),
    'mome_sm_count': (C.c_int, []),
    'mome_ln_fwd': (C.c_int, [_P, _P, _P, _P, C.c_int, _P, _P, _L, _L, _F, _P]),
    'mome_ln_bwd': (C.c_int, [_P, C.c_int, _P, _P, _P, _P, _P, _P, _P, _P, _L, _L, _P, C.c_size_t, _P]),
    'mome_ln_bwd_scale': (C.c_int, [_P, C.c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _L, _P, _P, C.c_size_t, _P]),
    'mome_scale_bwd': (C.c_int, [_P, _P, C.c_int, _P, _P, C.c_int, _P, _P, _L, _L, _P, _P, C.c_size_t, _P]),
    'mome_droppath_scales': (C.c_int, [_P, _L, _P, C.c_uint32, _F, _P, _P]),
    'mome_colsum': (C.c_int, [_P, C.c_int, _L, _L, _L, _P, _P, C.c_size_t, _P]),
    'mome_colreduce': (C.c_int, [_P, _L, _L, _P, _P]),
    'mome_reduce_ws_bytes': (C.c_size_t, [_L]),
    'mome_cast_bf16': (C.c_int, [_P, _P, _L, _P]),
    'mome_gemm': (C.c_int, [C.POINTER(GemmArgs), _P]),
    'mome_attn_fwd': (C.c_int, [_P, C.c_int, _P, _P, _P, _P, _L, _I, _I, _I, _F, _P, C.c_uint32, _F, _P]),
    'mome_attn_bwd_ws_floats': (C.c_int64, [_L, _I, _I, _I]),
    'mome_attn_bwd': (C.c_int, [_P, _P, _P, C.c_int, _P, _P, _P, _P, _P, _L, _I, _I, _I, _F, _P, C.c_uint32, _F, _P]),
    'mome_l2norm_fwd': (C.c_int, [_P, C.c_int, _P, _P, _L, _L, _P]),
    'mome_l2norm_bwd': (C.c_int, [_P, _P, _P, _P, _L, _L, _P]),
    'mome_itc_fwd': (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    'mome_itc_bwd': (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P]),
    'mome_itc_fwd_peer': (C.c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    'mome_itc_bwd_peer': (C.c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P]),
    'mome_itc_gather_peer': (C.c_int, [_P, _I, _I, _I, _P, _P, _P]),
    'mome_ce_fwd': (C.c_int, [_P, _L, _I, _I, _P, _L, _P, _P, _P, _P, _P]),
    'mome_ce_bwd': (C.c_int, [_P, _L, _I, _I, _P, _L, _P, _P, _P]),
    'mome_text_embed_fwd': (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _F, _P, C.c_uint32, _F, _P]),
    'mome_text_embed_ws_bytes': (C.c_size_t, [_I]),
    'mome_text_embed_bwd': (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _L, _P, C.c_uint32, _F, _P, C.c_size_t, _P]),
    'mome_adamw_flat': (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _F, _F, _F, _L, _P]),
    'mome_sumsq': (C.c_int, [_P, _L, _P, _P]),
    'mome_block_fwd': (C.c_int, [C.POINTER(BlockArgs), _P]),
    'mome_block_bwd': (C.c_int, [C.POINTER(BlockArgs), _P]),
    'mome_prof_enable': (C.c_int, [C.c_int]),
    'mome_prof_read': (C.c_int, [C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int]),
    'mome_launch_count': (C.c_int64, []),
}

_lib = None


def exported_symbols():
    return sorted(_SIGNATURES)


def lib():
    """Load libmome.so once. Raises if it is absent: there is no other implementation."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f'{LIB_PATH} not found: build it with `python -m exploremultimodal_b200.build_ext` '
                '(there is no CPU or PyTorch fallback for the MoME kernels)')
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.mome_version() != ABI_VERSION:
            raise RuntimeError(f'libmome ABI {handle.mome_version()} != binding ABI {ABI_VERSION}; rebuild')
        _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        raise RuntimeError(f'{what} failed (status {rc}): {lib().mome_last_error().decode()}')


def dtype_code(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f'unsupported dtype {t.dtype}')


def ptr(t):
    if t is None:
        return None
    assert t.is_cuda, 'libmome operates on CUDA tensors only'
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def launch_count():
    return int(lib().mome_launch_count())
