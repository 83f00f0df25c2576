"""CPU tests of the ORACLE (oracle/): pinned against the golden vectors and the reference's own
test cases.  No GPU, no product code on the path."""
import ctypes as C
import json
import os
import random
import struct
from pathlib import Path

import pytest

from conftest import make_tree, read_tree

GOLD = Path(__file__).resolve().parent / "golden"


def pattern(n):
    return bytes(((i * 2654435761) & 0xFFFFFFFF) >> 24 for i in range(n))


def kat_cases():
    k = json.loads((GOLD / "xxh3_kat.json").read_text())
    for e in k["literals"]:
        data = bytes.fromhex(e["hex_input"]) if e["hex_input"] is not None else bytes([e["fill"]]) * e["len"]
        yield e["name"], data, bytes.fromhex(e["digest"])
    for e in k["pattern"]:
        yield f"pattern{e['len']}", pattern(e["len"]), bytes.fromhex(e["digest"])


def test_oracle_xxh3_matches_golden_vectors(oracle):
    for name, data, want in kat_cases():
        assert oracle.hash_chunk(data) == want, name


def test_oracle_xxh3_matches_libxxhash(oracle):
    """second pin: the image's libxxhash.so.0 (0.8.2), XXH3_128bits -> {low64, high64}"""
    lib = C.CDLL("libxxhash.so.0")

    class H(C.Structure):
        _fields_ = [("lo", C.c_uint64), ("hi", C.c_uint64)]
    lib.XXH3_128bits.restype = H
    lib.XXH3_128bits.argtypes = [C.c_char_p, C.c_size_t]
    rng = random.Random(7)
    for n in list(range(0, 300)) + [rng.randrange(300, 70000) for _ in range(60)] + [1 << 20, (2 << 20) - 1, 2 << 20]:
        b = rng.randbytes(n)
        h = lib.XXH3_128bits(b, n)
        assert oracle.hash_chunk(b) == struct.pack("<QQ", h.lo, h.hi), n


def test_hash_chunk_consistent_and_distinct(oracle):
    # reference src/util/tests.rs:88-104
    assert oracle.hash_chunk(b"some test data") == oracle.hash_chunk(b"some test data")
    assert oracle.hash_chunk(b"data 1") != oracle.hash_chunk(b"data 2")


def test_oracle_reads_reference_fixture(oracle, tmp_path):
    # reference src/archive/tests.rs:14-58,117-166: hand-built archive, digest [1;16] that is NOT the
    # hash of the data, frame without content size, orig_size == exact size
    a = tmp_path / "dummy.squish"
    a.write_bytes(bytes.fromhex((GOLD / "dummy_archive.hex").read_text().strip()))
    rc, s, listing = oracle.list(a)
    assert rc == 0
    assert s.unique_chunks == 1 and s.total_original_size == 4 and s.archive_size == a.stat().st_size
    assert s.compression_ratio > 0 and s.file_count == 1 and listing == "4 file1.txt\n"
    assert s.version == b"1.2.0"
    rc, _ = oracle.unpack(a, tmp_path / "output")
    assert rc == 0
    assert (tmp_path / "output" / "file1.txt").read_bytes() == b"test"


def test_oracle_nonexistent_and_corrupt(oracle, tmp_path):
    # src/archive/tests.rs:168-172, tests/cli_tests.rs:82-120
    rc, _, _ = oracle.list(tmp_path / "nonexistent.squish")
    assert rc == -14  # FileNotExist
    bad = tmp_path / "corrupt.squish"
    bad.write_bytes(b"not a valid archive at all, just some bytes")
    rc, _, _ = oracle.list(bad)
    assert rc != 0
    wrong = tmp_path / "wrongver.squish"
    wrong.write_bytes(b"squish9.9.0" + bytes(64))  # src/util/tests.rs:33-42 incompatible version
    rc, _, _ = oracle.list(wrong)
    assert rc == -8


def test_oracle_roundtrip_small(oracle, tmp_path):
    # tests/roundtrip.rs:1-26 and tests/cli_tests.rs:12-53,122-159
    src = tmp_path / "in"
    make_tree(src, {"file.txt": b"hello squish", "a.txt": b"Hello, world!\n", "dir1/dir2/nested.txt": b"nested file",
                    "empty.bin": b""})
    rc, st = oracle.pack_dir(src, tmp_path / "o.squish")
    assert rc == 0 and st.total_chunks == 3 and st.unique_chunks == 3
    rc, _ = oracle.unpack(tmp_path / "o.squish", tmp_path / "out")
    assert rc == 0
    assert read_tree(tmp_path / "out") == read_tree(src)


def test_oracle_empty_dir(oracle, tmp_path):
    # tests/cli_tests.rs:55-80: empty directory -> 31-byte archive, number_of_files: 0
    (tmp_path / "e").mkdir()
    rc, st = oracle.pack_dir(tmp_path / "e", tmp_path / "e.squish")
    assert rc == 0 and st.archive_size == 31
    rc, s, listing = oracle.list(tmp_path / "e.squish")
    assert rc == 0 and s.file_count == 0 and s.unique_chunks == 0 and listing == ""


def test_oracle_chunk_rule_and_dedup(oracle, tmp_path):
    # chunk rule writer.rs:240-246 (2 MiB, exact multiple => no empty tail) + dedup chunk.rs:80-100
    rng = random.Random(3)
    two = rng.randbytes(2 << 20)
    tail = rng.randbytes(12345)
    src = tmp_path / "in"
    make_tree(src, {"big.bin": two + two + tail, "copy.bin": two + two + tail, "exact.bin": two, "zero.bin": bytes(3 << 20)})
    rc, st = oracle.pack_dir(src, tmp_path / "o.squish", threads=8)
    assert rc == 0
    assert st.total_chunks == 3 + 3 + 1 + 2
    assert st.unique_chunks == 2 + 2  # {two, tail} + {2 MiB zeros, 1 MiB zeros}
    a = (tmp_path / "o.squish").read_bytes()
    # every record carries orig_size == CHUNK_SIZE (writer.rs:255 quirk)
    p = 27
    for _ in range(st.unique_chunks):
        orig, comp = struct.unpack_from("<QQ", a, p + 16)
        assert orig == 2 << 20
        frame = a[p + 32:p + 32 + comp]
        assert frame[:4] == bytes.fromhex("28b52ffd")
        p += 32 + comp
    rc, _ = oracle.unpack(tmp_path / "o.squish", tmp_path / "out", threads=8, parallel_decode=1)
    assert rc == 0 and read_tree(tmp_path / "out") == read_tree(src)


def test_oracle_digest_map_winner_rule(oracle):
    a, b = b"x" * 100, b"y" * 100
    dig, new, uniq = oracle.digest_map([a, b, a, b"", a + b])
    assert len(dig) == 4 and new == [1, 1, 0, 1] and uniq == 3
    assert dig[0] == dig[2] == oracle.hash_chunk(a)


def test_oracle_zstd_frame_shapes(oracle):
    # SURVEY Appendix C.1 probes: what the reference's compress(chunk, 12) emits
    assert oracle.compress(b"hello squish").hex() == "28b52ffd200c61000068656c6c6f20737175697368"
    assert oracle.compress(b"").hex() == "28b52ffd2000010000"
    rnd = random.Random(1).randbytes(2 << 20)
    assert len(oracle.compress(rnd)) == 2097209
    f = oracle.compress(bytes([42]) * 2048)
    assert len(f) < 2048 and oracle.decompress(f, 2 << 20) == bytes([42]) * 2048  # src/util/tests.rs:146-166
    assert oracle.decompress(f, 2047) is None        # capacity one byte short
    assert oracle.decompress(f + b"x", 4096) is None  # trailing garbage
