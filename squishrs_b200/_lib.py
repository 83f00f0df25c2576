"""ctypes binding of libsquish_b200.so — the stub a maintainer would write for the C ABI in
include/squish_b200.h.  Loading fails loudly when the CUDA library has not been built; there is
no Python or CPU fallback for any entry point."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
# SQ_B200_LIB points the loader at another build of the same library (A/B runs of kernel variants)
LIB_PATH = Path(os.environ["SQ_B200_LIB"]) if os.environ.get("SQ_B200_LIB") else _HERE / "libsquish_b200.so"

CHUNK_SIZE = 2048 * 1024
DIGEST_BYTES = 16

SQ_OK = 0
SQ_ERR_READER = -5
SQ_ERR_COMPRESSION = -7
SQ_ERR_ARCHIVE = -8
SQ_ERR_FILE_NOT_EXIST = -14
SQ_ERR_MISSING_CHUNK = -16
SQ_ERR_INVALID_CHUNK_SIZE = -17
SQ_ERR_NO_DEVICE = -100
SQ_ERR_CUDA = -101
SQ_ERR_INVALID_ARG = -102
SQ_ERR_CAPACITY = -103
SQ_FLAG_DENSE_SEARCH = 1
SQ_FLAG_STAGE_TIMING = 2
SQ_FLAG_DETERMINISTIC = 4


class SqConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("chunk_size", C.c_uint32), ("dedup_capacity", C.c_uint64),
                ("max_batch_chunks", C.c_uint32), ("flags", C.c_uint32)]


class SqSpan(C.Structure):
    _fields_ = [("off", C.c_uint64), ("len", C.c_uint32), ("reserved", C.c_uint32)]


class SqChunkResult(C.Structure):
    _fields_ = [("digest", C.c_uint8 * 16), ("frame_off", C.c_uint64), ("frame_len", C.c_uint32),
                ("is_new", C.c_uint8), ("reserved", C.c_uint8 * 3)]


class SqFrame(C.Structure):
    _fields_ = [("src_off", C.c_uint64), ("dst_off", C.c_uint64), ("src_len", C.c_uint32), ("capacity", C.c_uint32)]


class SqFrameResult(C.Structure):
    _fields_ = [("out_len", C.c_uint32), ("status", C.c_int32)]


class SqPackReport(C.Structure):
    _fields_ = [("archive_size", C.c_uint64), ("unique_chunks", C.c_uint64), ("total_chunks", C.c_uint64),
                ("total_input_bytes", C.c_uint64), ("payload_bytes", C.c_uint64), ("file_count", C.c_uint32),
                ("seconds_total", C.c_double), ("seconds_device", C.c_double)]


class SqSummary(C.Structure):
    _fields_ = [("unique_chunks", C.c_uint64), ("total_original_size", C.c_uint64), ("archive_size", C.c_uint64),
                ("timestamp", C.c_uint64), ("compression_ratio", C.c_double), ("file_count", C.c_uint32),
                ("version", C.c_char * 16), ("seconds_total", C.c_double), ("seconds_device", C.c_double)]


assert C.sizeof(SqSpan) == 16 and C.sizeof(SqChunkResult) == 32 and C.sizeof(SqFrame) == 24 and C.sizeof(SqFrameResult) == 8

# every symbol include/squish_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "sq_create": (C.c_int32, [C.POINTER(SqConfig), C.POINTER(_P)]),
    "sq_destroy": (None, [_P]),
    "sq_last_error": (C.c_char_p, [_P]),
    "sq_strerror": (C.c_char_p, [C.c_int32]),
    "sq_abi_version": (C.c_int32, []),
    "sq_synchronize": (C.c_int32, [_P, _P]),
    "sq_kernel_launches": (C.c_int32, [_P, C.POINTER(C.c_uint64)]),
    "sq_host_alloc": (C.c_int32, [_P, C.c_size_t, C.POINTER(_P)]),
    "sq_host_free": (C.c_int32, [_P, _P]),
    "sq_digest_device": (C.c_int32, [_P, _P, _P, C.c_uint32, _P, _P]),
    "sq_digest_host": (C.c_int32, [_P, _P, C.c_size_t, _P, C.c_uint32, _P]),
    "sq_dedup_insert_device": (C.c_int32, [_P, _P, _P, C.c_uint64, C.c_uint32, _P, _P]),
    "sq_dedup_len": (C.c_int32, [_P, C.POINTER(C.c_uint64)]),
    "sq_dedup_reset": (C.c_int32, [_P]),
    "sq_route_digests_device": (C.c_int32, [_P, _P, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, _P, _P, _P]),
    "sq_dedup_insert_routed_device": (C.c_int32, [_P, _P, C.c_uint32, _P, _P]),
    "sq_unroute_verdicts_device": (C.c_int32, [_P, _P, _P, C.c_uint32, _P, _P]),
    "sq_encode_bound": (C.c_size_t, [C.c_size_t]),
    "sq_encode_device": (C.c_int32, [_P, _P, _P, _P, C.c_uint32, _P, C.c_uint64, _P, _P, _P, _P]),
    "sq_encode_status": (C.c_int32, [_P]),
    "sq_encode_stage_ms": (C.c_int32, [_P, _P, C.POINTER(C.c_float)]),
    "sq_decode_device": (C.c_int32, [_P, _P, _P, C.c_uint32, _P, _P, _P]),
    "sq_pack_device": (C.c_int32, [_P, _P, _P, C.c_uint32, C.c_uint64, _P, _P, C.c_uint64, C.POINTER(C.c_uint64), _P]),
    "sq_pack_host": (C.c_int32, [_P, _P, C.c_size_t, _P, C.c_uint32, C.c_uint64, _P, _P, C.c_uint64, C.POINTER(C.c_uint64)]),
    "sq_pack_submit": (C.c_int32, [_P, _P, C.c_size_t, _P, C.c_uint32, C.c_uint64, _P, _P, C.c_uint64, C.POINTER(_P)]),
    "sq_pack_wait": (C.c_int32, [_P, _P, C.POINTER(C.c_uint64)]),
    "sq_unpack_host": (C.c_int32, [_P, _P, C.c_size_t, _P, C.c_uint32, _P, C.c_size_t, _P]),
    "sq_unpack_submit": (C.c_int32, [_P, _P, C.c_size_t, _P, C.c_uint32, _P, C.c_size_t, _P, C.POINTER(_P)]),
    "sq_unpack_wait": (C.c_int32, [_P, _P]),
    "sq_archive_pack": (C.c_int32, [_P, C.c_char_p, C.c_char_p, C.c_int32, C.POINTER(SqPackReport)]),
    "sq_archive_unpack": (C.c_int32, [_P, C.c_char_p, C.c_char_p, C.c_int32, C.POINTER(SqSummary)]),
    "sq_archive_pack_multi": (C.c_int32, [C.POINTER(_P), C.c_uint32, C.c_char_p, C.c_char_p, C.c_int32, C.POINTER(SqPackReport)]),
    "sq_archive_unpack_multi": (C.c_int32, [C.POINTER(_P), C.c_uint32, C.c_char_p, C.c_char_p, C.c_int32, C.POINTER(SqSummary)]),
    "sq_share_dedup": (C.c_int32, [_P, _P]),
    "sq_release_scratch": (C.c_int32, [_P]),
    "sq_archive_list": (C.c_int32, [C.c_char_p, C.POINTER(SqSummary), C.POINTER(C.c_char_p)]),
    "sq_free": (None, [_P]),
    "sq_corpus_fill_device": (C.c_int32, [_P, _P, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, _P]),
    "sq_corpus_fill_host": (C.c_int32, [_P, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32]),
    "sq_corpus_fill_slots_device": (C.c_int32, [_P, _P, C.c_uint64, _P, _P, C.c_uint32, C.c_uint64, _P]),
}

_lib = None


def load() -> C.CDLL:
    """dlopen the product library and bind every declared symbol (no compute happens here)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a).  squishrs_b200 has no CPU fallback.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class SquishError(RuntimeError):
    """An AppError (reference src/util/errors.rs) surfaced through the C ABI status code."""

    def __init__(self, status: int, text: str):
        super().__init__(f"[{status}] {text}")
        self.status = status
