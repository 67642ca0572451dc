"""ctypes binding of libshortseq_b200.so (the C ABI declared in include/shortseq_b200.h).

The library is the only compute path: if it is missing, or there is no CUDA
device, every operation raises -- there is no CPU fallback.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SSQ_LIB") or os.path.join(HERE, "libshortseq_b200.so")

OK, ERR_BAD_BASE, ERR_TOO_LONG, ERR_CLASS, ERR_CUDA, ERR_LEN_MISMATCH, ERR_TABLE_FULL, ERR_ARG, ERR_EXCHANGE = range(9)
CLASS_64, CLASS_192, CLASS_VAR = 0, 1, 2


class Report(C.Structure):
    _fields_ = [("code", C.c_int32), ("reserved", C.c_int32), ("first_bad_read", C.c_int64)]


class LibraryError(RuntimeError):
    """The CUDA library is missing, or a CUDA / argument error occurred inside it."""


_p, _i64, _i32, _int, _u64 = C.c_void_p, C.c_int64, C.c_int32, C.c_int, C.c_uint64

# name -> (restype, argtypes); every symbol include/shortseq_b200.h declares
PROTOTYPES = {
    "ssq_abi_version": (_int, []),
    "ssq_last_error": (C.c_char_p, []),
    "ssq_launch_count": (_u64, []),
    "ssq_device_count": (_int, [C.POINTER(_int)]),
    "ssq_ctx_create": (_int, [_int, C.POINTER(_p)]),
    "ssq_ctx_destroy": (_int, [_p]),
    "ssq_ctx_set_stream": (_int, [_p, _p]),
    "ssq_ctx_reset_stream": (_int, [_p]),
    "ssq_ctx_stream": (_p, [_p]),
    "ssq_ctx_sync": (_int, [_p, C.POINTER(Report)]),
    "ssq_malloc": (_int, [_p, C.c_size_t, C.POINTER(_p)]),
    "ssq_free": (_int, [_p, _p]),
    "ssq_host_alloc": (_int, [C.c_size_t, C.POINTER(_p)]),
    "ssq_host_free": (_int, [_p]),
    "ssq_memcpy_h2d": (_int, [_p, _p, _p, C.c_size_t]),
    "ssq_memcpy_d2h": (_int, [_p, _p, _p, C.c_size_t]),
    "ssq_memset": (_int, [_p, _p, _int, C.c_size_t]),
    "ssq_pack64": (_int, [_p, _p, _i64, _p, _i64, _p, _p]),
    "ssq_pack192": (_int, [_p, _p, _i64, _p, _i64, _p, _p]),
    "ssq_packvar_words_bound": (_i64, [_i64, _i64]),
    "ssq_packvar": (_int, [_p, _p, _i64, _p, _i64, _p, _p, _p]),
    "ssq_lens_to_offsets": (_int, [_p, _p, _int, _i64, _p]),
    "ssq_decode64": (_int, [_p, _p, _p, _i64, _p, _p]),
    "ssq_decode192": (_int, [_p, _p, _p, _i64, _p, _p]),
    "ssq_decode_tiles": (_int, [_p, _p, _i64, _int, _p]),
    "ssq_decode64_fused": (_int, [_p, _p, _p, _i64, _p, _p, _p]),
    "ssq_decode192_fused": (_int, [_p, _p, _p, _i64, _p, _p, _p]),
    "ssq_decodevar": (_int, [_p, _p, _p, _p, _i64, _p, _p]),
    "ssq_hamming_pairs64": (_int, [_p, _p, _p, _p, _p, _i64, _p]),
    "ssq_hamming_pairs192": (_int, [_p, _p, _p, _p, _p, _i64, _p]),
    "ssq_hamming_pairsvar": (_int, [_p, _p, _p, _p, _p, _p, _p, _i64, _p]),
    "ssq_hamming_refset": (_int, [_p, _int, _p, _p, _i64, _p, _p, _i32, _i32, _p, _p, _p]),
    "ssq_counter_create": (_int, [_p, _int, _i64, _int, C.POINTER(_p)]),
    "ssq_counter_destroy": (_int, [_p]),
    "ssq_counter_clear": (_int, [_p]),
    "ssq_counter_insert": (_int, [_p, _p, _p, _i64]),
    "ssq_counter_merge": (_int, [_p, _p, _p, _p, _i64]),
    "ssq_counter_pack_count": (_int, [_p, _p, _i64, _p, _i64, _p, _p]),
    "ssq_counter_first_index": (_int, [_p, _p, _p, _i64, _i64]),
    "ssq_counter_lookup": (_int, [_p, _p, _p, _i64, _p]),
    "ssq_counter_size": (_int, [_p, C.POINTER(_i64)]),
    "ssq_counter_capacity": (_int, [_p, C.POINTER(_i64)]),
    "ssq_counter_last_pass_ms": (_int, [_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "ssq_counter_regions": (_int, [_p, C.POINTER(_i64)]),
    "ssq_counter_merge_regions": (_int, [_p, _p, _p, _p, _p, _p, _int, _p, _i64]),
    "ssq_counter_export_region_bases": (_int, [_p, _int, _p]),
    "ssq_counter_export_counts": (_int, [_p, _int, _p]),
    "ssq_counter_export_to": (_int, [_p, _int, _int, _p, _p, _p]),
    "ssq_ipc_get_handle": (_int, [_p, _p, _p]),
    "ssq_ipc_open": (_int, [_p, _p, C.POINTER(_p)]),
    "ssq_ipc_close": (_int, [_p, _p]),
    "ssq_classify": (_int, [_p, _p, _i64, _p]),
    "ssq_host_fastq_count": (_int, [_p, _p, _p, _p, _i64, _i64, _int, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(Report)]),
    "ssq_counter_last_pass_detail": (_int, [_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "ssq_counter_export": (_int, [_p, _int, _p, _p, _p, _p, _p]),
    "ssq_host_pack_count": (_int, [_p, _p, _p, _p, _i64, _p, _p, _i64, C.POINTER(Report)]),
    "ssq_host_pack_count_lens": (_int, [_p, _p, _p, _p, _i64, _p, _i64, C.POINTER(Report)]),
    "ssq_synth_reads": (_int, [_p, _u64, _i64, _i64, _i64, _i32, _i32, _p, _p]),
    "ssq_slice_words": (_int, [_p, _int, _p, _i64, _p, _p, _i64, _i64, _i32, _p]),
    "ssq_slice": (_int, [_p, _int, _p, _p, _p, _i64, _p, _p, _i64, _i64, _i32, _int, _p, _p, _p]),
    "ssq_kmers_count": (_int, [_p, _int, _p, _i64, _i32, _i32, _p]),
    "ssq_kmers64": (_int, [_p, _int, _p, _p, _p, _i64, _i32, _i32, _p, _p, _p]),
    "ssq_normalize": (_int, [_p, _p, _i64, _p]),
    "ssq_umi_cluster": (_int, [_p, _p, _p, _p, _i64, _p, _i64, _i32, _i32, _p, _p]),
    "ssq_pack_one": (_int, [_p, C.c_char_p, _i32, _p, C.POINTER(_i32), C.POINTER(_i32)]),
    "ssq_decode_one": (_int, [_p, _p, _i32, _p]),
    "ssq_hamming_one": (_int, [_p, _p, _p, _i32, C.POINTER(_i32)]),
    "ssq_comm_unique_id": (_int, [_p]),
    "ssq_comm_init": (_int, [_p, _p, _int, _int, C.POINTER(_p)]),
    "ssq_comm_destroy": (_int, [_p]),
    "ssq_comm_uses_peer_stores": (_int, [_p]),
    "ssq_comm_attach": (_int, [_p, _p, _p, C.POINTER(C.c_int)]),
    "ssq_comm_last_merge_streamed": (_int, [_p]),
    "ssq_counter_merge_alltoall": (_int, [_p, _p, _p, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
}

_LIB = None


def lib():
    """Load the shared library (once).  Raises LibraryError when it has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise LibraryError(
                f"{LIB_PATH} not found: build it with `python -m shortseq_b200.build` "
                "(nvcc, sm_100a).  shortseq_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        if L.ssq_abi_version() != 1:
            raise LibraryError("libshortseq_b200.so ABI version mismatch; rebuild it")
        _LIB = L
    return _LIB


def check(rc):
    """Raise for a non-zero C-ABI return code."""
    if rc != OK:
        msg = lib().ssq_last_error().decode(errors="replace")
        raise LibraryError(f"shortseq_b200 C-ABI call failed (status {rc}): {msg}")
