"""ctypes binding of libemdr2_b200.so — the only way the package reaches the GPU.

There is no CPU fallback: if the shared library is missing and cannot be built, or a call fails,
this module raises.  Signatures mirror include/emdr2_b200.h one to one.
"""
import ctypes
import os

from . import build as _build

_LIB = None

EMDR2_DTYPE_FP16 = 0
EMDR2_DTYPE_BF16 = 1
MAX_K = 64
QUERIES_PER_PASS = 64

_c_void_pp = ctypes.POINTER(ctypes.c_void_p)

_SIGNATURES = {
    "emdr2_last_error": (ctypes.c_char_p, []),
    "emdr2_version": (ctypes.c_char_p, []),
    "emdr2_mips_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, _c_void_pp]),
    "emdr2_mips_set_shard": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_int64, ctypes.c_int64]),
    "emdr2_mips_search": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                         ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p]),
    "emdr2_mips_search_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                              ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                              ctypes.c_void_p]),
    "emdr2_mips_merge": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_void_p]),
    "emdr2_mips_set_option": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int64]),
    "emdr2_mips_get_stat": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p,
                                           ctypes.POINTER(ctypes.c_int64)]),
    "emdr2_mips_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "emdr2_gemm": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                  ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                  ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                  ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "emdr2_attention_fwd": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64,
                                           ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                           ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                           ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_int,
                                           ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p]),
    "emdr2_gemm_ex": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,
                                     ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p,
                                     ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                     ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "emdr2_layernorm_fwd": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_void_p]),
    "emdr2_embedding_fwd": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "emdr2_token_logprob": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "emdr2_attention_bwd": (ctypes.c_int, [ctypes.c_int] + [ctypes.c_void_p, ctypes.c_int64] * 8 +
                            [ctypes.c_int] * 4 + [ctypes.c_void_p] * 4 + [ctypes.c_int, ctypes.c_float,
                                                                          ctypes.c_void_p, ctypes.c_void_p,
                                                                          ctypes.c_void_p]),
    "emdr2_layernorm_bwd": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                           ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_void_p]),
    "emdr2_colsum": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                    ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "emdr2_token_logprob_bwd": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                               ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                               ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "emdr2_embedding_bwd": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                           ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_void_p]),
    "emdr2_format_passages": (ctypes.c_int, [ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32,
                                             ctypes.c_int32, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                             ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_void_p, ctypes.c_void_p]),
    "emdr2_format_passages_flat": (ctypes.c_int, [ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
                                                  ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                  ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                                  ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                                                  ctypes.c_int32, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                                  ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                  ctypes.c_void_p, ctypes.c_void_p]),
    "emdr2_attention_varlen_fwd": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                                  ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                                  ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                                  ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_float,
                                                  ctypes.c_void_p, ctypes.c_void_p]),
    "emdr2_embedding_fwd_pos": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                               ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                               ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                               ctypes.c_void_p]),
    "emdr2_dropout_colhash": (ctypes.c_int, [ctypes.c_uint64, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "emdr2_dropout_mask": (ctypes.c_int, [ctypes.c_float, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p]),
    "emdr2_dropout_add": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                         ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_float, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p,
                                         ctypes.c_void_p]),
    "emdr2_attention_fwd_dropout": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64,
                                                   ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                                   ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                                   ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                   ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_void_p,
                                                   ctypes.c_float, ctypes.c_uint64, ctypes.c_uint64,
                                                   ctypes.c_void_p, ctypes.c_void_p]),
    "emdr2_attention_bwd_dropout": (ctypes.c_int, [ctypes.c_int] + [ctypes.c_void_p, ctypes.c_int64] * 8 +
                                    [ctypes.c_int] * 4 + [ctypes.c_void_p] * 4 +
                                    [ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_float, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p,
                                     ctypes.c_void_p]),
    "emdr2_ops_set_option": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int64]),
    "emdr2_ops_get_option": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(ctypes.c_int64)]),
    "emdr2_ops_timing": (ctypes.c_int, [ctypes.c_int]),
    "emdr2_ops_timing_add_flops": (ctypes.c_int, [ctypes.c_int, ctypes.c_double]),
    "emdr2_ops_timing_read": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_int64),
                                             ctypes.POINTER(ctypes.c_int64),
                                             ctypes.POINTER(ctypes.c_double)]),
}


def exported_symbols():
    """Names include/emdr2_b200.h declares (used by the CPU-side ABI test)."""
    return sorted(_SIGNATURES)


def lib_path():
    return _build.LIB_PATH


def load():
    """Load (building in-tree first if needed) and type the shared library."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.build()        # no-op when the stamp matches the sources; never loads a stale library
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == ABI mismatch: fail loudly
        fn.restype = restype
        fn.argtypes = argtypes
    _LIB = lib
    return lib


class Emdr2Error(RuntimeError):
    pass


def check(rc, what):
    if rc != 0:
        msg = load().emdr2_last_error().decode("utf-8", "replace")
        raise Emdr2Error("%s failed (code %d): %s" % (what, rc, msg))
