"""CPU tests of the product's boundary: the C-ABI library loads, exports every symbol that
include/squish_b200.h declares, refuses to run without a GPU (no CPU fallback), and the
device-free host entry points (list) follow the reference."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
GOLD = ROOT / "tests" / "golden"


def declared_symbols():
    text = (ROOT / "include" / "squish_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sq_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported(sq):
    from squishrs_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 25
    lib = C.CDLL(str(_lib.LIB_PATH))
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/squish_b200.h but not exported"
    assert set(names) == set(_lib.SYMBOLS), "ctypes binding table and header disagree"
    assert lib.sq_abi_version() == 1


def test_struct_layout(sq):
    from squishrs_b200 import _lib
    assert C.sizeof(_lib.SqSpan) == 16 and C.sizeof(_lib.SqChunkResult) == 32
    assert C.sizeof(_lib.SqFrame) == 24 and C.sizeof(_lib.SqFrameResult) == 8 and C.sizeof(_lib.SqConfig) == 24
    assert _lib.load().sq_encode_bound(2 << 20) == 2097209  # 9 + 16*3 + 2 MiB: the raw-block frame libzstd emits too


def test_no_cpu_fallback(sq):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(sq.SquishError) as e:
        sq.Context()
    assert e.value.status == -100
    with pytest.raises(sq.SquishError):
        sq.hash_chunk(b"some test data")


def test_list_reference_fixture_without_gpu(sq, tmp_path):
    # reference src/archive/tests.rs:117-139 (get_summary on the hand-built archive)
    from squishrs_b200.archive import ArchiveReader
    a = tmp_path / "dummy.squish"
    a.write_bytes(bytes.fromhex((GOLD / "dummy_archive.hex").read_text().strip()))
    s = ArchiveReader(a).get_summary()
    assert s.unique_chunks == 1 and s.total_original_size == 4 and s.archive_size > 0 and s.compression_ratio > 0
    assert len(s.files) == 1 and s.files[0].path == "file1.txt" and s.files[0].original_size == 4
    assert s.squish_version == "1.2.0"


def test_reader_errors_without_gpu(sq, tmp_path):
    from squishrs_b200.archive import ArchiveReader
    with pytest.raises(sq.SquishError) as e:  # src/archive/tests.rs:168-172
        ArchiveReader(tmp_path / "nonexistent.squish")
    assert e.value.status == -14
    bad = tmp_path / "bad.squish"
    bad.write_bytes(b"garbage garbage garbage garbage")
    with pytest.raises(sq.SquishError) as e:  # tests/cli_tests.rs:82-99
        ArchiveReader(bad)
    assert e.value.status == -8
    wrong = tmp_path / "wrong.squish"
    wrong.write_bytes(b"squish1.3.0" + bytes(40))  # src/util/tests.rs:33-42
    with pytest.raises(sq.SquishError) as e:
        ArchiveReader(wrong)
    assert e.value.status == -8 and "Incompatible version" in str(e.value)


def test_cli_list_simple(sq, tmp_path):
    # tests/cli_tests.rs:55-80 output format `number_of_files: N`
    cli = ROOT / "bin" / "squishrs"
    a = tmp_path / "dummy.squish"
    a.write_bytes(bytes.fromhex((GOLD / "dummy_archive.hex").read_text().strip()))
    r = subprocess.run([str(cli), "list", str(a), "--simple"], capture_output=True, text=True)
    assert r.returncode == 0
    assert "number_of_files: 1" in r.stdout and "chunks_count: 1" in r.stdout and "file1.txt" in r.stdout
    r = subprocess.run([str(cli), "list", str(tmp_path / "nope.squish")], capture_output=True, text=True)
    assert r.returncode == 1 and "Error" in r.stderr


def test_corpus_host_generator_is_deterministic(sq):
    lib = sq.load()
    for klass in range(7):
        a = C.create_string_buffer(10000)
        b = C.create_string_buffer(10000)
        assert lib.sq_corpus_fill_host(a, 10000, 42, 7, klass) == 0
        assert lib.sq_corpus_fill_host(b, 10000, 42, 7, klass) == 0
        assert a.raw == b.raw
        c = C.create_string_buffer(10000)
        lib.sq_corpus_fill_host(c, 10000, 42, 8, klass)
        if klass != 5:
            assert c.raw != a.raw
    t = C.create_string_buffer(4096)
    lib.sq_corpus_fill_host(t, 4096, 1, 1, 0)
    assert all(ch in b"abcdefghijklmnopqrstuvwxyz \n" for ch in t.raw)
