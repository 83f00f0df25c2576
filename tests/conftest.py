import ctypes as C
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


class _Stats(C.Structure):
    _fields_ = [("archive_size", C.c_uint64), ("unique_chunks", C.c_uint64), ("total_chunks", C.c_uint64),
                ("total_input_bytes", C.c_uint64), ("payload_bytes", C.c_uint64), ("seconds", C.c_double)]


class _Summary(C.Structure):
    _fields_ = [("unique_chunks", C.c_uint64), ("total_original_size", C.c_uint64), ("archive_size", C.c_uint64),
                ("timestamp", C.c_uint64), ("compression_ratio", C.c_double), ("file_count", C.c_uint32),
                ("version", C.c_char * 16), ("decode_seconds", C.c_double), ("rebuild_seconds", C.c_double)]


class _File(C.Structure):
    _fields_ = [("rel_path", C.c_char_p), ("path_on_disk", C.c_char_p), ("data", C.c_void_p), ("size", C.c_uint64)]


class Oracle:
    """ctypes view of oracle/liboracle.so — the CHECKER (tests only)."""
    Stats, Summary, File = _Stats, _Summary, _File

    def __init__(self):
        so = ROOT / "oracle" / "liboracle.so"
        if not so.exists():
            subprocess.run(["make", "-C", str(ROOT / "oracle")], check=True, capture_output=True)
        L = C.CDLL(str(so))
        L.sqo_hash_chunk.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p]
        L.sqo_zstd_bound.restype = C.c_size_t
        L.sqo_zstd_bound.argtypes = [C.c_size_t]
        L.sqo_zstd_compress.restype = C.c_size_t
        L.sqo_zstd_compress.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_int]
        L.sqo_zstd_decompress.restype = C.c_size_t
        L.sqo_zstd_decompress.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
        L.sqo_pack_dir.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(_Stats)]
        L.sqo_pack.argtypes = [C.POINTER(_File), C.c_uint32, C.c_char_p, C.c_int, C.c_uint64, C.c_int, C.POINTER(_Stats)]
        L.sqo_unpack.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.POINTER(_Summary)]
        L.sqo_list.argtypes = [C.c_char_p, C.POINTER(_Summary), C.POINTER(C.c_void_p)]
        L.sqo_digest_map.argtypes = [C.POINTER(_File), C.c_uint32, C.c_char_p, C.c_char_p, C.c_uint64,
                                     C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.sqo_free.argtypes = [C.c_void_p]
        L.sqo_zstd_version.restype = C.c_uint
        self.L = L

    def hash_chunk(self, b: bytes) -> bytes:
        out = C.create_string_buffer(16)
        self.L.sqo_hash_chunk(b, len(b), out)
        return out.raw

    def compress(self, b: bytes, level: int = 12) -> bytes:
        cap = self.L.sqo_zstd_bound(len(b))
        out = C.create_string_buffer(cap)
        n = self.L.sqo_zstd_compress(b, len(b), out, cap, level)
        assert n > 0
        return out.raw[:n]

    def decompress(self, b: bytes, cap: int):
        """stock ZSTD_decompress; returns None on error (capacity too small, garbage...)."""
        out = C.create_string_buffer(max(cap, 1))
        n = self.L.sqo_zstd_decompress(b, len(b), out, cap)
        if n == C.c_size_t(-1).value:
            return None
        return out.raw[:n]

    def pack_dir(self, d, out, threads=4):
        st = _Stats()
        rc = self.L.sqo_pack_dir(str(d).encode(), str(out).encode(), threads, C.byref(st))
        return rc, st

    def unpack(self, a, out, threads=4, parallel_decode=0):
        s = _Summary()
        rc = self.L.sqo_unpack(str(a).encode(), str(out).encode(), threads, parallel_decode, C.byref(s))
        return rc, s

    def list(self, a):
        s = _Summary()
        p = C.c_void_p()
        rc = self.L.sqo_list(str(a).encode(), C.byref(s), C.byref(p))
        text = ""
        if rc == 0:
            text = C.string_at(p).decode()
            self.L.sqo_free(p)
        return rc, s, text

    def digest_map(self, blobs):
        """blobs: list of bytes (files). Returns (digests[list of bytes], is_new[list of int], n_unique)."""
        n = len(blobs)
        files = (_File * max(n, 1))()
        keep = []
        total = 0
        for i, b in enumerate(blobs):
            buf = C.create_string_buffer(b, len(b)) if len(b) else C.create_string_buffer(1)
            keep.append(buf)
            files[i].rel_path = b"f"
            files[i].path_on_disk = None
            files[i].data = C.cast(buf, C.c_void_p)
            files[i].size = len(b)
            total += (len(b) + (1 << 21) - 1) >> 21
        dig = C.create_string_buffer(max(total, 1) * 16)
        isn = C.create_string_buffer(max(total, 1))
        nc, nu = C.c_uint64(), C.c_uint64()
        rc = self.L.sqo_digest_map(files, n, dig, isn, total, C.byref(nc), C.byref(nu))
        assert rc == 0
        return [dig.raw[16 * i:16 * i + 16] for i in range(nc.value)], list(isn.raw[:nc.value]), nu.value


@pytest.fixture(scope="session")
def oracle():
    return Oracle()


@pytest.fixture(scope="session")
def sq():
    """The product package with its CUDA library loaded (no compute)."""
    import squishrs_b200
    squishrs_b200.load()
    return squishrs_b200


@pytest.fixture(scope="session")
def ctx(sq):
    """A device context: only GPU-marked tests may ask for it."""
    c = sq.Context(dedup_capacity=1 << 20, max_batch_chunks=4096)
    yield c
    c.close()


def make_tree(root: Path, spec):
    """spec: {relative path: bytes}"""
    for rel, data in spec.items():
        p = root / rel
        p.parent.mkdir(parents=True, exist_ok=True)
        p.write_bytes(data)


def read_tree(root: Path):
    out = {}
    for p in sorted(root.rglob("*")):
        if p.is_file():
            out[str(p.relative_to(root))] = p.read_bytes()
    return out
