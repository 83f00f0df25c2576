"""squishrs_b200 — B200-native pack/unpack data path of squishRS behind the reference's seams.

The product is libsquish_b200.so (hand-written CUDA for sm_100a + host C++, C ABI in
include/squish_b200.h).  This package is the ctypes binding plus a host-side mirror of the
reference interface (same names as squishRS's Rust modules).  There is no CPU fallback.
"""
from ._lib import SquishError, load  # noqa: F401
from .context import CHUNK_SIZE, COMPRESSION_LEVEL, ChunkStore, Context, InsertReturn, hash_chunk  # noqa: F401

__all__ = ["SquishError", "load", "CHUNK_SIZE", "COMPRESSION_LEVEL", "ChunkStore", "Context", "InsertReturn", "hash_chunk"]
