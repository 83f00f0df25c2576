"""Device context + the reference-facing host mirror of the chunk store seam.

Mirrors (same names, argument meaning and error behaviour):
  hash_chunk            reference src/util/chunk.rs:46-49
  ChunkStore / insert   src/util/chunk.rs:19-25,52-56,80-100,116-136
  InsertReturn          src/util/chunk.rs:14-17
Every call goes through the C ABI into the CUDA kernels; nothing here computes a digest or a
frame on the CPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Iterable, List, Optional, Sequence

from . import _lib as L

CHUNK_SIZE = L.CHUNK_SIZE
COMPRESSION_LEVEL = 12


class Context:
    """Owns one sq_ctx (one GPU).  Raises SquishError(SQ_ERR_NO_DEVICE) without a CUDA device."""

    def __init__(self, device: int = 0, dedup_capacity: int = 1 << 20, max_batch_chunks: int = 4096, dense_search: bool = False,
                 stage_timing: bool = False, deterministic: bool = False):
        """dense_search: SQ_FLAG_DENSE_SEARCH, accepted and without effect since round 2 (every position is searched by default).
        stage_timing: SQ_FLAG_STAGE_TIMING, per-kernel durations of the encoder through sq_encode_stage_ms.
        deterministic: SQ_FLAG_DETERMINISTIC, the same chunk always compresses to the same bytes (slower index kernel)."""
        self.lib = L.load()
        cfg = L.SqConfig(device, 0, dedup_capacity, max_batch_chunks,
                         (L.SQ_FLAG_DENSE_SEARCH if dense_search else 0) | (L.SQ_FLAG_STAGE_TIMING if stage_timing else 0) |
                         (L.SQ_FLAG_DETERMINISTIC if deterministic else 0))
        h = C.c_void_p()
        rc = self.lib.sq_create(C.byref(cfg), C.byref(h))
        if rc != L.SQ_OK:
            raise L.SquishError(rc, (self.lib.sq_last_error(None) or b"").decode() or self.lib.sq_strerror(rc).decode())
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.sq_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc: int):
        if rc != L.SQ_OK:
            msg = (self.lib.sq_last_error(self.h) or b"").decode()
            raise L.SquishError(rc, f"{self.lib.sq_strerror(rc).decode()}: {msg}")

    # ---- batch helpers over HOST buffers (what the reference-facing classes use) ----
    @staticmethod
    def _layout(chunks: Sequence[bytes]):
        """Pack chunks into one batch buffer at 16-byte aligned offsets."""
        spans = (L.SqSpan * len(chunks))()
        off = 0
        for i, c in enumerate(chunks):
            spans[i].off = off
            spans[i].len = len(c)
            off += (len(c) + 15) & ~15
        buf = bytearray(off if off else 16)
        for i, c in enumerate(chunks):
            buf[spans[i].off:spans[i].off + len(c)] = c
        return buf, spans

    def digest_batch(self, chunks: Sequence[bytes]) -> List[bytes]:
        n = len(chunks)
        if n == 0:
            return []
        buf, spans = self._layout(chunks)
        cbuf = (C.c_uint8 * len(buf)).from_buffer(buf)
        out = (C.c_uint8 * (16 * n))()
        self.check(self.lib.sq_digest_host(self.h, cbuf, len(buf), spans, n, out))
        raw = bytes(out)
        return [raw[16 * i:16 * i + 16] for i in range(n)]

    def pack_batch(self, chunks: Sequence[bytes], gidx_base: int = 0):
        """ChunkStore::insert for a batch: returns [(digest, frame-or-None)]."""
        n = len(chunks)
        if n == 0:
            return []
        buf, spans = self._layout(chunks)
        cbuf = (C.c_uint8 * len(buf)).from_buffer(buf)
        cap = sum(self.lib.sq_encode_bound(len(c)) for c in chunks)
        out = (C.c_uint8 * cap)()
        res = (L.SqChunkResult * n)()
        used = C.c_uint64()
        self.check(self.lib.sq_pack_host(self.h, cbuf, len(buf), spans, n, gidx_base, res, out, cap, C.byref(used)))
        mv = memoryview(out)
        ret = []
        for i in range(n):
            frame = bytes(mv[res[i].frame_off:res[i].frame_off + res[i].frame_len]) if res[i].is_new else None
            ret.append((bytes(res[i].digest), frame))
        return ret

    def unpack_batch(self, payloads: Sequence[bytes], capacities: Sequence[int], raise_on_error: bool = True):
        """read_chunks for a batch: returns [bytes] or raises SquishError(ReaderError) like the reference; with
        raise_on_error=False a payload that fails to decode yields None instead (errors are per payload, not per call)."""
        n = len(payloads)
        if n == 0:
            return []
        frames = (L.SqFrame * n)()
        so = do = 0
        for i, (p, cap) in enumerate(zip(payloads, capacities)):
            frames[i].src_off, frames[i].dst_off, frames[i].src_len, frames[i].capacity = so, do, len(p), cap
            so += (len(p) + 15) & ~15
            do += (cap + 15) & ~15
        comp = bytearray(so if so else 16)
        for i, p in enumerate(payloads):
            comp[frames[i].src_off:frames[i].src_off + len(p)] = p
        out = (C.c_uint8 * (do if do else 16))()
        res = (L.SqFrameResult * n)()
        self.check(self.lib.sq_unpack_host(self.h, (C.c_uint8 * len(comp)).from_buffer(comp), len(comp), frames, n, out, do, res))
        mv = memoryview(out)
        ret = []
        for i in range(n):
            if res[i].status != L.SQ_OK:
                if raise_on_error:
                    raise L.SquishError(res[i].status, f"Error reading from squish: frame {i} failed to decode")
                ret.append(None)
                continue
            ret.append(bytes(mv[frames[i].dst_off:frames[i].dst_off + res[i].out_len]))
        return ret

    def dedup_len(self) -> int:
        v = C.c_uint64()
        self.check(self.lib.sq_dedup_len(self.h, C.byref(v)))
        return v.value

    def dedup_reset(self):
        self.check(self.lib.sq_dedup_reset(self.h))


_default: Optional[Context] = None


def default_context() -> Context:
    global _default
    if _default is None:
        _default = Context()
    return _default


def hash_chunk(chunk: bytes, ctx: Optional[Context] = None) -> bytes:
    """hash_chunk(chunk) -> ChunkHash ([u8;16]); reference src/util/chunk.rs:46-49."""
    return (ctx or default_context()).digest_batch([bytes(chunk)])[0]


@dataclass
class InsertReturn:
    """reference src/util/chunk.rs:14-17"""
    hash: bytes
    compressed_data: Optional[bytes]


class ChunkStore:
    """reference src/util/chunk.rs:19-25: a set of digests; insert() = dedup then encode."""

    def __init__(self, ctx: Optional[Context] = None):
        self.ctx = ctx or Context()
        self._next_gidx = 0

    def insert(self, chunk: bytes) -> InsertReturn:
        if len(chunk) == 0 or len(chunk) > CHUNK_SIZE:
            raise L.SquishError(L.SQ_ERR_INVALID_CHUNK_SIZE, f"Invalid chunk size: {len(chunk)} bytes")
        return self.insert_batch([chunk])[0]

    def insert_batch(self, chunks: Iterable[bytes]) -> List[InsertReturn]:
        chunks = [bytes(c) for c in chunks]
        res = self.ctx.pack_batch(chunks, self._next_gidx)
        self._next_gidx += len(chunks)
        return [InsertReturn(d, f) for d, f in res]

    def len(self) -> int:
        return self.ctx.dedup_len()

    def __len__(self) -> int:
        return self.len()

    def is_empty(self) -> bool:
        return self.len() == 0
