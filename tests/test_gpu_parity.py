"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against
the CPU oracle on the same inputs.  Bit-exact for digests, dedup verdicts and decoded bytes."""
import ctypes as C
import json
import random
import struct
from pathlib import Path

import pytest

from conftest import make_tree, read_tree
from test_oracle_cpu import kat_cases

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"
MiB = 1 << 20


# ---------------------------------------------------------------- K1 digest
def test_digest_golden_vectors(ctx):
    cases = list(kat_cases())
    got = ctx.digest_batch([d for _, d, _ in cases])
    for (name, _, want), g in zip(cases, got):
        assert g == want, name


def test_digest_matches_oracle_every_length_class(ctx, oracle):
    rng = random.Random(11)
    lens = list(range(1, 300)) + [rng.randrange(300, 5000) for _ in range(200)] + [rng.randrange(5000, 2 * MiB) for _ in range(40)]
    lens += [1023, 1024, 1025, 2 * MiB - 1, 2 * MiB, 2 * MiB - 63, 2 * MiB - 64, 2 * MiB - 65, MiB + 1]
    chunks = [rng.randbytes(n) for n in lens]
    got = ctx.digest_batch(chunks)
    for c, g in zip(chunks, got):
        assert g == oracle.hash_chunk(c), len(c)


def test_digest_unaligned_spans(ctx, oracle, sq):
    """spans at arbitrary byte offsets (the slow unaligned path) must give the same digests"""
    from squishrs_b200 import _lib as L
    rng = random.Random(5)
    buf = bytearray(rng.randbytes(3 * MiB))
    offs = [(1, 241), (3, 1024), (7, 4097), (13, 2 * MiB), (5, 70000), (9, 100), (2, 1025)]
    spans = (L.SqSpan * len(offs))()
    for i, (o, n) in enumerate(offs):
        spans[i].off, spans[i].len = o, n
    out = (C.c_uint8 * (16 * len(offs)))()
    ctx.check(ctx.lib.sq_digest_host(ctx.h, (C.c_uint8 * len(buf)).from_buffer(buf), len(buf), spans, len(offs), out))
    raw = bytes(out)
    for i, (o, n) in enumerate(offs):
        assert raw[16 * i:16 * i + 16] == oracle.hash_chunk(bytes(buf[o:o + n])), (o, n)


def test_hash_chunk_mirror(sq, ctx, oracle):
    # reference src/util/tests.rs:88-104
    h1, h2 = sq.hash_chunk(b"some test data", ctx), sq.hash_chunk(b"some test data", ctx)
    assert h1 == h2 == bytes.fromhex("8bb490f540ec66cf9f9ea5e575cc3ff2")
    assert sq.hash_chunk(b"data 1", ctx) != sq.hash_chunk(b"data 2", ctx)


# ---------------------------------------------------------------- K2 dedup (ChunkStore)
def test_insert_first_time_returns_compressed_data(sq, oracle):
    # reference src/util/tests.rs:106-116
    store = sq.ChunkStore(sq.Context())
    data = bytes([1]) * 1024
    r = store.insert(data)
    assert r.hash == oracle.hash_chunk(data) and r.compressed_data is not None and store.len() == 1


def test_insert_duplicate_returns_none(sq):
    # reference src/util/tests.rs:118-130
    store = sq.ChunkStore(sq.Context())
    data = bytes([2]) * 1024
    first, second = store.insert(data), store.insert(data)
    assert first.compressed_data is not None and second.compressed_data is None
    assert first.hash == second.hash and store.len() == 1 and not store.is_empty()


def test_multiple_unique_inserts_increase_len(sq):
    # reference src/util/tests.rs:132-144
    store = sq.ChunkStore(sq.Context())
    assert store.is_empty()
    for v in (1, 2, 3):
        store.insert(bytes([v]) * 1024)
    assert store.len() == 3


def test_compressed_data_is_smaller_and_stock_decodable(sq, oracle):
    # reference src/util/tests.rs:146-166
    store = sq.ChunkStore(sq.Context())
    data = bytes([42]) * 2048
    r = store.insert(data)
    assert r.compressed_data is not None
    assert oracle.decompress(r.compressed_data, 2 * MiB) == data
    assert len(r.compressed_data) < len(data)


def test_insert_rejects_bad_sizes(sq):
    store = sq.ChunkStore(sq.Context())
    with pytest.raises(sq.SquishError):
        store.insert(b"")
    with pytest.raises(sq.SquishError):
        store.insert(bytes(2 * MiB + 1))


def test_dedup_map_matches_oracle_in_one_batch(sq, oracle):
    """duplicates INSIDE one batch and across batches: verdicts = lowest global chunk index wins"""
    rng = random.Random(2)
    uniq = [rng.randbytes(rng.randrange(1, 5000)) for _ in range(300)]
    seq = [uniq[rng.randrange(len(uniq))] for _ in range(2000)]
    want_dig, want_new, want_u = oracle.digest_map(seq)  # each blob < 2 MiB => one chunk per blob
    c = sq.Context()
    got = []
    for i in range(0, len(seq), 700):
        got += c.pack_batch(seq[i:i + 700], gidx_base=i)
    assert [d for d, _ in got] == want_dig
    assert [int(f is not None) for _, f in got] == want_new
    assert c.dedup_len() == want_u
    c.dedup_reset()
    assert c.dedup_len() == 0


# ---------------------------------------------------------------- K3 encode
def corpus_samples(sq, sizes=(1, 100, 4096, 70000, 300000, 2 * MiB)):
    lib = sq.load()
    out = []
    for klass in range(7):
        for n in sizes:
            b = C.create_string_buffer(n)
            lib.sq_corpus_fill_host(b, n, 99, klass * 10 + 1, klass)
            out.append((klass, b.raw))
    return out


def test_frames_decode_with_stock_libzstd(sq, oracle):
    """every GPU-written frame must decode byte-identically with stock ZSTD_decompress, with
    capacity == CHUNK_SIZE (what the reference passes, reader.rs:286-303) and exact capacity"""
    c = sq.Context()
    samples = corpus_samples(sq)
    rng = random.Random(4)
    samples += [(-1, rng.randbytes(n)) for n in (1, 255, 256, 65791, 65792, 131072, 131073, 2 * MiB)]
    samples += [(-2, bytes(n)) for n in (1, 131072, 2 * MiB)]
    res = c.pack_batch([s for _, s in samples])
    for (klass, data), (dig, frame) in zip(samples, res):
        assert dig == oracle.hash_chunk(data)
        if frame is None:
            continue  # duplicate payload (e.g. identical zeros)
        assert frame[:4] == bytes.fromhex("28b52ffd")
        assert oracle.decompress(frame, 2 * MiB) == data, (klass, len(data))
        assert oracle.decompress(frame, len(data)) == data
        assert len(frame) <= c.lib.sq_encode_bound(len(data))


def test_device_corpus_matches_host_corpus(sq, ctx):
    import torch
    lib = sq.load()
    for klass in range(7):
        n = 300000 + klass
        t = torch.empty(n + 8, dtype=torch.uint8, device="cuda")
        ctx.check(lib.sq_corpus_fill_device(ctx.h, t.data_ptr(), n, 1234, 55, klass, None))
        ctx.check(lib.sq_synchronize(ctx.h, None))
        h = C.create_string_buffer(n)
        lib.sq_corpus_fill_host(h, n, 1234, 55, klass)
        assert bytes(t[:n].cpu().numpy()) == h.raw, klass


# ---------------------------------------------------------------- archive level
def tree_spec():
    rng = random.Random(8)
    big = rng.randbytes(2 * MiB)
    text = (b"the quick brown fox jumps over the lazy dog\n" * 60000)[:2 * MiB + 4321]
    return {"file.txt": b"hello squish", "a/b/c/nested.txt": b"nested file", "empty.bin": b"", "big.bin": big + big + b"tail",
            "copy/big2.bin": big + big + b"tail", "text.txt": text, "zeros.bin": bytes(3 * MiB), "exact.bin": big}


def test_archive_pack_reference_can_unpack(sq, oracle, tmp_path):
    """GPU-written archive -> the reference reader (oracle) restores the identical tree;
    manifest digests and the unique set equal the oracle's own pack of the same tree."""
    from squishrs_b200.archive import ArchiveReader, ArchiveWriter
    src = tmp_path / "in"
    make_tree(src, tree_spec())
    w = ArchiveWriter(src, tmp_path / "gpu.squish", ctx=sq.Context(), threads=8)
    size = w.pack()
    assert size == (tmp_path / "gpu.squish").stat().st_size
    rc, st = oracle.pack_dir(src, tmp_path / "cpu.squish", threads=8)
    assert rc == 0
    assert w.report.unique_chunks == st.unique_chunks and w.report.total_chunks == st.total_chunks
    rc, _ = oracle.unpack(tmp_path / "gpu.squish", tmp_path / "out")
    assert rc == 0 and read_tree(tmp_path / "out") == read_tree(src)
    # manifests as path-keyed maps (walk order is OS dependent)
    assert parse_manifest(tmp_path / "gpu.squish") == parse_manifest(tmp_path / "cpu.squish")
    s = ArchiveReader(tmp_path / "gpu.squish").get_summary()
    assert s.total_original_size == sum(len(v) for v in tree_spec().values()) and len(s.files) == len(tree_spec())


def parse_manifest(path):
    a = Path(path).read_bytes()
    assert a[:11] == b"squish1.2.0"
    n = struct.unpack_from("<Q", a, 19)[0]
    p = 27
    uniq = set()
    for _ in range(n):
        orig, comp = struct.unpack_from("<QQ", a, p + 16)
        assert orig == 2 * MiB
        uniq.add(a[p:p + 16])
        p += 32 + comp
    nf = struct.unpack_from("<I", a, p)[0]
    p += 4
    files = {}
    for _ in range(nf):
        pl = struct.unpack_from("<I", a, p)[0]
        path_s = a[p + 4:p + 4 + pl].decode()
        p += 4 + pl
        size, cc = struct.unpack_from("<QI", a, p)
        p += 12
        files[path_s] = (size, a[p:p + 16 * cc])
        p += 16 * cc
    assert p == len(a)
    return uniq, files


def test_archive_empty_dir(sq, oracle, tmp_path):
    # tests/cli_tests.rs:55-80
    from squishrs_b200.archive import ArchiveWriter
    (tmp_path / "e").mkdir()
    assert ArchiveWriter(tmp_path / "e", tmp_path / "e.squish", ctx=sq.Context()).pack() == 31
    rc, s, _ = oracle.list(tmp_path / "e.squish")
    assert rc == 0 and s.file_count == 0 and s.unique_chunks == 0


# ---------------------------------------------------------------- K4 decode
def test_decode_reference_written_frames(sq, oracle):
    """frames written by stock libzstd at several levels (what a reference-written archive holds, plus
    shapes only other levels emit) decode byte-identically; capacity = 2 MiB as the reference passes"""
    c = sq.Context()
    samples = [s for _, s in corpus_samples(sq, sizes=(1, 17, 255, 256, 4096, 65536, 131073, 300000, 2 * MiB))]
    rng = random.Random(9)
    samples += [rng.randbytes(70000), bytes(3000), bytes(rng.choices(range(3), k=100000))]
    for lvl in (1, 3, 12, 19):
        chosen = [s for s in samples if lvl != 19 or len(s) <= 300000]
        frames = [oracle.compress(s, lvl) for s in chosen]
        got = c.unpack_batch(frames, [2 * MiB] * len(frames))
        for s, g in zip(chosen, got):
            assert g == s, (lvl, len(s))


def test_decode_own_frames_roundtrip(sq):
    c = sq.Context()
    samples = [s for _, s in corpus_samples(sq)]
    res = c.pack_batch(samples)
    frames = [(s, f) for s, (_, f) in zip(samples, res) if f is not None]
    got = c.unpack_batch([f for _, f in frames], [len(s) for s, _ in frames])  # exact capacity
    assert got == [s for s, _ in frames]


def test_decode_multiframe_skippable_and_errors(sq, oracle):
    # what stock ZSTD_decompress accepts / rejects (SURVEY Appendix D probes; reader.rs:302-303 semantics)
    c = sq.Context()
    two = oracle.compress(b"hello ") + oracle.compress(b"squish")
    skip = bytes.fromhex("502a4d18") + (3).to_bytes(4, "little") + b"abc"
    nofcs = bytes.fromhex("28b52ffd") + bytes([0x00, 0x58, 0x21, 0, 0]) + b"test"  # zstd::encode_all shape (src/archive/tests.rs:31)
    assert c.unpack_batch([two, skip + two, nofcs, oracle.compress(b"")], [64, 64, 4, 0]) == [b"hello squish", b"hello squish", b"test", b""]
    f = oracle.compress(bytes([42]) * 2048)
    for bad, cap in ((f, 2047), (f + b"x", 4096), (f[:-3], 4096), (b"garbage!", 64), (f[:4] + b"\xff" + f[5:], 4096)):
        with pytest.raises(sq.SquishError) as e:
            c.unpack_batch([bad], [cap])
        assert e.value.status == -5  # ReaderError
        assert oracle.decompress(bad, cap) is None  # stock libzstd rejects it too


def test_archive_roundtrip_both_directions(sq, oracle, tmp_path):
    """reference-written archive -> GPU unpack, GPU-written archive -> GPU unpack; trees byte-identical"""
    from squishrs_b200.archive import ArchiveReader, ArchiveWriter
    src = tmp_path / "in"
    make_tree(src, tree_spec())
    rc, _ = oracle.pack_dir(src, tmp_path / "cpu.squish", threads=8)
    assert rc == 0
    c = sq.Context()
    ArchiveReader(tmp_path / "cpu.squish", ctx=c, threads=8).unpack(tmp_path / "out_from_cpu")
    assert read_tree(tmp_path / "out_from_cpu") == read_tree(src)
    ArchiveWriter(src, tmp_path / "gpu.squish", ctx=c, threads=8).pack()
    ArchiveReader(tmp_path / "gpu.squish", ctx=c, threads=8).unpack(tmp_path / "out_from_gpu")
    assert read_tree(tmp_path / "out_from_gpu") == read_tree(src)
    # into a directory that already exists (files are then created and truncated up front), over stale, longer files
    for f in (tmp_path / "out_from_gpu").rglob("*"):
        if f.is_file():
            f.write_bytes(f.read_bytes() + b"stale tail")
    ArchiveReader(tmp_path / "cpu.squish", ctx=c, threads=8).unpack(tmp_path / "out_from_gpu")
    assert read_tree(tmp_path / "out_from_gpu") == read_tree(src)


def record_digests(path):
    a = Path(path).read_bytes()
    n = struct.unpack_from("<Q", a, 19)[0]
    p, out = 27, []
    for _ in range(n):
        out.append(a[p:p + 16])
        p += 32 + struct.unpack_from("<Q", a, p + 24)[0]
    return out


def test_archive_several_contexts_share_one_index(sq, oracle, tmp_path, monkeypatch):
    """sq_archive_pack_multi / sq_archive_unpack_multi: batches dealt to several contexts (two GPUs when the box has them, else
    two contexts on one GPU -- the same code path, the digests then travel device-to-device on one device).  The archive must
    hold exactly the records a single context writes, in the same order (first occurrence wins across devices), the reference
    reader must restore the tree from it, and the multi-context unpack must restore it too."""
    import torch
    from squishrs_b200.archive import ArchiveReader, ArchiveWriter
    src = tmp_path / "in"
    spec = tree_spec()
    rng = random.Random(9)
    for i in range(6):  # more chunks, with duplicates that sit in different batches
        spec[f"more/f{i}.bin"] = rng.randbytes(2 * MiB + 999 * i) + spec["big.bin"][:2 * MiB]
    make_tree(src, spec)
    monkeypatch.setenv("SQ_PACK_BATCH_BYTES", str(5 * MiB))
    ArchiveWriter(src, tmp_path / "one.squish", ctx=sq.Context(), threads=8).pack()
    devs = [0, 1] if torch.cuda.device_count() >= 2 else [0, 0]
    ctxs = [sq.Context(device=d) for d in devs]
    w = ArchiveWriter(src, tmp_path / "multi.squish", ctxs=ctxs, threads=8)
    w.pack()
    assert record_digests(tmp_path / "multi.squish") == record_digests(tmp_path / "one.squish")
    assert parse_manifest(tmp_path / "multi.squish") == parse_manifest(tmp_path / "one.squish")
    rc, _ = oracle.unpack(tmp_path / "multi.squish", tmp_path / "out_ref")
    assert rc == 0 and read_tree(tmp_path / "out_ref") == read_tree(src)
    ArchiveReader(tmp_path / "multi.squish", threads=8).unpack(tmp_path / "out_multi", ctxs=ctxs)
    assert read_tree(tmp_path / "out_multi") == read_tree(src)
    rc, _ = oracle.pack_dir(src, tmp_path / "cpu.squish", threads=8)
    assert rc == 0
    ArchiveReader(tmp_path / "cpu.squish", threads=8).unpack(tmp_path / "out_multi2", ctxs=ctxs)
    assert read_tree(tmp_path / "out_multi2") == read_tree(src)


def test_unpack_reference_fixture(sq, tmp_path):
    # reference src/archive/tests.rs:141-166
    from squishrs_b200.archive import ArchiveReader
    a = tmp_path / "dummy.squish"
    a.write_bytes(bytes.fromhex((GOLD / "dummy_archive.hex").read_text().strip()))
    ArchiveReader(a, ctx=sq.Context()).unpack(tmp_path / "output")
    assert (tmp_path / "output" / "file1.txt").read_bytes() == b"test"


def test_unpack_missing_chunk(sq, tmp_path):
    # reader.rs:397-401: a manifest digest with no chunk record -> MissingChunk
    from squishrs_b200.archive import ArchiveReader
    a = bytearray(bytes.fromhex((GOLD / "dummy_archive.hex").read_text().strip()))
    a[-1] ^= 0xFF
    p = tmp_path / "missing.squish"
    p.write_bytes(bytes(a))
    with pytest.raises(sq.SquishError) as e:
        ArchiveReader(p, ctx=sq.Context()).unpack(tmp_path / "o")
    assert e.value.status == -16


def test_cli_roundtrip(sq, tmp_path):
    # reference tests/cli_tests.rs:12-53,122-159 through the squishrs binary
    import subprocess
    cli = Path(__file__).resolve().parent.parent / "bin" / "squishrs"
    src = tmp_path / "in"
    make_tree(src, {"file1.txt": b"Hello, world!\n", "file2.txt": b"Another file\n", "dir1/dir2/nested.txt": b"nested file"})
    r = subprocess.run([str(cli), "pack", str(src), "-o", str(tmp_path / "a.squish")], capture_output=True, text=True)
    assert r.returncode == 0 and "Packing complete!" in r.stdout, r.stderr
    r = subprocess.run([str(cli), "unpack", str(tmp_path / "a.squish"), "-o", str(tmp_path / "out")], capture_output=True, text=True)
    assert r.returncode == 0 and "Unpacking complete!" in r.stdout, r.stderr
    assert read_tree(tmp_path / "out") == read_tree(src)
    (tmp_path / "bad.squish").write_bytes(b"this is not an archive")
    r = subprocess.run([str(cli), "unpack", str(tmp_path / "bad.squish"), "-o", str(tmp_path / "o2")], capture_output=True, text=True)
    assert r.returncode != 0 and "Error" in r.stderr


def test_routed_dedup_kernels_single_gpu(sq, oracle):
    """route -> (exchange emulated in place: this GPU owns every shard) -> insert_routed -> unroute must give
    the same verdicts as the plain insert; exercises the three multi-GPU kernels without a second GPU"""
    import torch
    rng = random.Random(21)
    uniq = [rng.randbytes(64) for _ in range(100)]
    seq = [uniq[rng.randrange(len(uniq))] for _ in range(600)]
    want_dig, want_new, want_u = oracle.digest_map(seq)
    world, n = 4, len(seq)
    c = sq.Context(max_batch_chunks=world * n)
    lib = c.lib
    dig = torch.frombuffer(bytearray(b"".join(want_dig)), dtype=torch.uint8).cuda()
    send = torch.empty(world * n * 32, dtype=torch.uint8, device="cuda")
    pos = torch.empty(n, dtype=torch.int32, device="cuda")
    verdict = torch.empty(world * n, dtype=torch.uint8, device="cuda")
    is_new = torch.empty(n, dtype=torch.uint8, device="cuda")
    c.check(lib.sq_route_digests_device(c.h, dig.data_ptr(), 1000, n, world, n, send.data_ptr(), pos.data_ptr(), None))
    c.check(lib.sq_dedup_insert_routed_device(c.h, send.data_ptr(), world * n, verdict.data_ptr(), None))
    c.check(lib.sq_unroute_verdicts_device(c.h, verdict.data_ptr(), pos.data_ptr(), n, is_new.data_ptr(), None))
    c.check(lib.sq_synchronize(c.h, None))
    assert is_new.cpu().tolist() == want_new
    assert c.dedup_len() == want_u
    # records landed in the block of their owner: LE64(digest[:8]) % world
    from squishrs_b200.sharded import owner_of
    p = pos.cpu().tolist()
    for i, d in enumerate(want_dig):
        assert p[i] // n == owner_of(d, world)


# ---------------------------------------------------------------- BASELINE.json configs at test scale
def gen(sq, klass, n, pid):
    b = C.create_string_buffer(max(n, 1))
    sq.load().sq_corpus_fill_host(b, n, 0x5151, pid, klass)
    return b.raw[:n]


def test_config4_random_chunks_are_16_raw_blocks(sq, oracle):
    """configs[3]: incompressible input -> every frame is 16 raw blocks, exactly 2 097 209 bytes (what libzstd emits)"""
    c = sq.Context()
    chunks = [gen(sq, 4, 2 * MiB, i) for i in range(4)] + [gen(sq, 4, 300000, 9)]
    res = c.pack_batch(chunks)
    for data, (_, frame) in zip(chunks, res):
        assert frame is not None and len(frame) == len(oracle.compress(data)) == c.lib.sq_encode_bound(len(data))
        assert oracle.decompress(frame, len(data)) == data
    assert len(res[0][1]) == 2097209


def test_config3_vm_like_dedup_map(sq, oracle):
    """configs[2] at test scale: zero slots + popular copies; digest list, verdicts and unique count = oracle"""
    rng = random.Random(33)
    payloads = [bytes(2 * MiB)] + [gen(sq, 6, 2 * MiB, i) for i in range(6)] + [gen(sq, 4, 2 * MiB, 50)]
    slots = [payloads[0]] + [payloads[min(int(rng.random() ** 2 * len(payloads)), len(payloads) - 1)] for _ in range(39)]
    want_dig, want_new, want_u = oracle.digest_map([b"".join(slots[:20]), b"".join(slots[20:])])  # two 40 MiB "files"
    c = sq.Context()
    got = c.pack_batch(slots[:25], 0) + c.pack_batch(slots[25:], 25)
    assert [d for d, _ in got] == want_dig and [int(f is not None) for _, f in got] == want_new and c.dedup_len() == want_u
    zero_frame = next(f for (d, f), s in zip(got, slots) if s == payloads[0] and f is not None)
    assert len(zero_frame) < 512 and oracle.decompress(zero_frame, 2 * MiB) == payloads[0]


def test_config5_unpack_reference_written_small_files(sq, oracle, tmp_path):
    """configs[4] at test scale: an archive of many 4-64 KiB files written by the reference path (libzstd L12
    single-block frames, orig_size = 2 MiB in every record) unpacks byte-identically on the GPU"""
    from squishrs_b200.archive import ArchiveReader
    rng = random.Random(55)
    spec = {}
    for i in range(600):
        n = rng.randrange(4096, 65537)
        spec[f"d{i % 7}/s{i % 13}/f{i}.dat"] = gen(sq, 0 if i % 2 else 2, n, 1000 + i)
    src = tmp_path / "in"
    make_tree(src, spec)
    rc, st = oracle.pack_dir(src, tmp_path / "ref.squish", threads=8)
    assert rc == 0 and st.unique_chunks == 600
    r = ArchiveReader(tmp_path / "ref.squish", ctx=sq.Context(), threads=8)
    r.unpack(tmp_path / "out")
    assert read_tree(tmp_path / "out") == read_tree(src)


def test_config1_text_tree_roundtrip_and_ratio(sq, oracle, tmp_path):
    """configs[0] at test scale: text-like tree with ~30% whole-file copies, nested dirs.  GPU pack -> reference
    unpack and reference pack -> GPU unpack are byte-identical; unique sets equal; payload within 3% of level 12"""
    from squishrs_b200.archive import ArchiveReader, ArchiveWriter
    rng = random.Random(77)
    spec, originals = {}, []
    for i in range(60):
        if originals and rng.random() < 0.3:
            data = originals[rng.randrange(len(originals))]
        else:
            data = gen(sq, 0, int(min(6 * MiB, max(1024, rng.lognormvariate(13.0, 1.0)))), 2000 + i)
            originals.append(data)
        spec[f"a{i % 4}/b{i % 3}/file{i}.txt"] = data
    src = tmp_path / "in"
    make_tree(src, spec)
    c = sq.Context()
    w = ArchiveWriter(src, tmp_path / "gpu.squish", ctx=c, threads=8)
    w.pack()
    rc, st = oracle.pack_dir(src, tmp_path / "cpu.squish", threads=8)
    assert rc == 0 and w.report.unique_chunks == st.unique_chunks and w.report.total_chunks == st.total_chunks
    assert parse_manifest(tmp_path / "gpu.squish") == parse_manifest(tmp_path / "cpu.squish")
    rc, _ = oracle.unpack(tmp_path / "gpu.squish", tmp_path / "o1")
    assert rc == 0 and read_tree(tmp_path / "o1") == read_tree(src)
    ArchiveReader(tmp_path / "cpu.squish", ctx=c, threads=8).unpack(tmp_path / "o2")
    assert read_tree(tmp_path / "o2") == read_tree(src)
    # compression-ratio tolerance stated by north_star: <= 3% worse than the reference at level 12 (libzstd 1.5.5 here)
    assert w.report.payload_bytes <= 1.03 * st.payload_bytes, (w.report.payload_bytes, st.payload_bytes)


def test_full_size_batch_properties(sq, oracle):
    """size-independent properties at full chunk size: encode -> decode identity on the GPU, and the digest of every
    decoded chunk equals the digest computed before encoding (a checksum of checksums)"""
    c = sq.Context()
    chunks = [gen(sq, k % 7, 2 * MiB, 300 + k) for k in range(28)]
    res = c.pack_batch(chunks)
    frames = [(s, d, f) for s, (d, f) in zip(chunks, res) if f is not None]
    back = c.unpack_batch([f for _, _, f in frames], [2 * MiB] * len(frames))
    assert back == [s for s, _, _ in frames]
    assert c.digest_batch(back) == [d for _, d, _ in frames]
    ratio_gpu = sum(len(f) for _, _, f in frames)
    ratio_cpu = sum(len(oracle.compress(s)) for s, _, _ in frames)
    assert ratio_gpu <= 1.03 * ratio_cpu, (ratio_gpu, ratio_cpu)


def test_ragged_sizes_every_class_roundtrip(sq, oracle):
    """ragged chunk lengths around every internal boundary (8-byte hash window, 24-byte search cap, 1 KiB tiles,
    128 KiB blocks) for every corpus class: GPU frames decode with stock libzstd and with K4, and K4 decodes the
    reference's frames of the same inputs"""
    c = sq.Context()
    rng = random.Random(3)
    chunks = []
    for k in range(7):
        for n in (1, 7, 8, 9, 23, 24, 25, 40, 63, 300, 1023, 1025, 4096, 70001, 131072, 131073, 262149):
            chunks.append(gen(sq, k, n, k * 100 + n % 97))
    chunks += [rng.randbytes(n) for n in (1, 100, 5000)]
    res = c.pack_batch(chunks)
    frames = [(s, f) for s, (_, f) in zip(chunks, res) if f is not None]
    for s, f in frames:
        assert oracle.decompress(f, len(s)) == s, len(s)
    assert c.unpack_batch([f for _, f in frames], [len(s) for s, _ in frames]) == [s for s, _ in frames]
    ref = [oracle.compress(s, 12) for s, _ in frames]
    assert c.unpack_batch(ref, [2 * MiB] * len(ref)) == [s for s, _ in frames]


def test_encode_device_on_several_streams(sq, oracle):
    """sq_encode_device called on three streams of one context (two scratch sets; the third stream takes a set over and is
    ordered behind its previous user): every batch's frames still decode to their own input with stock libzstd"""
    import numpy as np
    import torch
    c = sq.Context()
    lib = c.lib
    n, size = 12, 300000
    streams = [torch.cuda.Stream() for _ in range(3)]
    jobs = []
    for r in range(6):  # two rounds over the three streams, all in flight together
        blobs = [gen(sq, (r + k) % 5, size, 7000 + r * 100 + k) for k in range(n)]
        st = streams[r % 3]
        with torch.cuda.stream(st):
            data = torch.frombuffer(bytearray(b"".join(blobs)), dtype=torch.uint8).cuda()
            spans = torch.tensor([[k * size, size] for k in range(n)], dtype=torch.int64).cuda()
            cap = n * lib.sq_encode_bound(size)
            out = torch.empty(cap, dtype=torch.uint8, device="cuda")
            foff = torch.empty(n, dtype=torch.int64, device="cuda")
            flen = torch.empty(n, dtype=torch.int32, device="cuda")
            total = torch.zeros(1, dtype=torch.int64, device="cuda")
            c.check(lib.sq_encode_device(c.h, data.data_ptr(), spans.data_ptr(), None, n, out.data_ptr(), cap, foff.data_ptr(),
                                         flen.data_ptr(), total.data_ptr(), C.c_void_p(st.cuda_stream)))
        jobs.append((blobs, data, spans, out, foff, flen, total))
    torch.cuda.synchronize()
    assert lib.sq_encode_status(c.h) == 0
    for blobs, _, _, out, foff, flen, total in jobs:
        o = out.cpu().numpy()
        fo, fl = foff.cpu().tolist(), flen.cpu().tolist()
        assert int(total.item()) == sum(fl)
        for k, b in enumerate(blobs):
            assert oracle.decompress(o[fo[k]:fo[k] + fl[k]].tobytes(), size) == b, k


def test_unpack_pipeline_two_tickets_in_flight(sq, oracle):
    # sq_unpack_submit / sq_unpack_wait: same bytes as the synchronous sq_unpack_host, two batches in flight, a third refused
    from squishrs_b200 import _lib as L
    c = sq.Context()
    lib = c.lib
    samples = [s for _, s in corpus_samples(sq)]
    batches = [samples[0::2], samples[1::2]]
    staged = []
    for b in batches:
        payloads = [oracle.compress(s) for s in b]
        n = len(payloads)
        frames = (L.SqFrame * n)()
        so = do = 0
        for i, (pl, s) in enumerate(zip(payloads, b)):
            frames[i].src_off, frames[i].dst_off, frames[i].src_len, frames[i].capacity = so, do, len(pl), len(s)
            so += (len(pl) + 15) & ~15
            do += (len(s) + 15) & ~15
        comp = bytearray(so + 16)
        for i, pl in enumerate(payloads):
            comp[frames[i].src_off:frames[i].src_off + len(pl)] = pl
        staged.append(dict(n=n, frames=frames, comp=(C.c_uint8 * len(comp)).from_buffer(comp), so=so, out=(C.c_uint8 * (do + 16))(), do=do,
                           res=(L.SqFrameResult * n)(), ticket=C.c_void_p(), expect=b))
    for st in staged:
        c.check(lib.sq_unpack_submit(c.h, st["comp"], st["so"], st["frames"], st["n"], st["out"], st["do"], st["res"], C.byref(st["ticket"])))
    extra = C.c_void_p()
    st0 = staged[0]
    assert lib.sq_unpack_submit(c.h, st0["comp"], st0["so"], st0["frames"], st0["n"], st0["out"], st0["do"], st0["res"], C.byref(extra)) != 0  # both slots busy
    for st in staged:
        c.check(lib.sq_unpack_wait(c.h, st["ticket"]))
        mv = memoryview(st["out"])
        for i, s in enumerate(st["expect"]):
            assert st["res"][i].status == 0 and st["res"][i].out_len == len(s)
            assert bytes(mv[st["frames"][i].dst_off:st["frames"][i].dst_off + len(s)]) == s
    assert lib.sq_unpack_wait(c.h, staged[0]["ticket"]) != 0  # not in flight any more
    # the slots are free again, and the synchronous call still works on the same context
    assert c.unpack_batch([oracle.compress(samples[0])], [len(samples[0])]) == [samples[0]]


def real_corpora(limit=8 * MiB):
    """real files of this image (the same on the GPU box): Python sources, shared libraries, site-packages sources"""
    import glob
    import sysconfig

    def blob(pattern):
        out = bytearray()
        for f in sorted(glob.glob(pattern, recursive=True)):
            try:
                out += open(f, "rb").read()
            except OSError:
                continue
            if len(out) >= limit:
                break
        return bytes(out[:limit])
    return {"python sources": blob("/usr/lib/python3*/**/*.py"), "shared libraries": blob("/usr/lib/x86_64-linux-gnu/*.so*"),
            "site-packages": blob(sysconfig.get_paths()["purelib"] + "/**/*.py")}


def test_real_data_ratio_within_3pct_of_level12(sq, oracle):
    """the north_star tolerance (<= 3 % worse than the reference's zstd level 12, src/util/chunk.rs:12,89-90) on REAL files,
    per corpus, at the default settings: 2 MiB chunks of each corpus, and 24 KB pieces of all three"""
    c = sq.Context()
    sets = real_corpora()
    cases = {k: [v[i:i + 2 * MiB] for i in range(0, len(v), 2 * MiB)] for k, v in sets.items() if len(v) >= MiB}
    cases["24 KB pieces"] = [v[i:i + 24000] for v in sets.values() for i in range(0, min(len(v), MiB), 24000)]
    assert len(cases) >= 3
    for name, chunks in cases.items():
        c.dedup_reset()
        res = c.pack_batch(chunks)
        gpu = cpu = 0
        for ch, (_, f) in zip(chunks, res):
            if f is None:
                continue
            assert oracle.decompress(f, len(ch)) == ch, name
            gpu += len(f)
            cpu += len(oracle.compress(ch, 12))
        assert gpu <= 1.03 * cpu, (name, gpu, cpu, gpu / cpu)


def test_dense_search_flag_is_accepted(sq, oracle):
    # SQ_FLAG_DENSE_SEARCH (round 1: search every position) is the default behaviour now; the flag stays in the ABI and changes nothing
    chunks = [s for k, s in corpus_samples(sq, sizes=(300000,)) if k in (0, 1, 2)]
    sizes = {}
    for dense in (False, True):
        c = sq.Context(dense_search=dense)
        res = c.pack_batch(chunks)
        for ch, (_, f) in zip(chunks, res):
            assert f is not None and oracle.decompress(f, len(ch)) == ch
        sizes[dense] = sum(len(f) for _, f in res)
        c.close()
    assert abs(sizes[True] - sizes[False]) <= 0.002 * sizes[False], sizes


def test_release_scratch_and_reuse(sq, oracle):
    """sq_release_scratch hands the cached device scratch back; the next calls allocate it again and give the same answers"""
    rng = random.Random(31)
    c = sq.Context()
    chunks = [rng.randbytes(5000) + bytes(200000), (b"squish " * 40000)[:262144 + 77]]
    first = c.pack_batch(chunks)
    c.check(c.lib.sq_release_scratch(c.h))
    c.dedup_reset()
    again = c.pack_batch(chunks)
    assert [d for d, _ in first] == [d for d, _ in again]
    for data, (_, f) in zip(chunks, again):
        assert f is not None and oracle.decompress(f, 2 * MiB) == data
    c.check(c.lib.sq_release_scratch(c.h))
    assert c.unpack_batch([f for _, f in again], [2 * MiB] * 2) == chunks


def test_block_parallel_decode_of_damaged_frames(sq, oracle):
    """multi-block frames (K3-written and libzstd-written, so the block-parallel passes take them) with bits flipped, a few
    hundred per call next to intact ones: a payload the GPU accepts must decode to what stock libzstd makes of it, a damaged
    neighbour must not disturb an intact frame, and nothing may crash"""
    rng = random.Random(77)
    c = sq.Context()
    lib = sq.load()
    b = C.create_string_buffer(600000)
    lib.sq_corpus_fill_host(b, 600000, 9, 4, 2)
    data = b.raw
    own = c.pack_batch([data])[0][1]
    ref = oracle.compress(data, 12)
    payloads, want = [], []
    for base in (own, ref):
        payloads.append(base); want.append(data)
        for _ in range(150):
            g = bytearray(base)
            for _ in range(rng.randrange(1, 4)):
                g[rng.randrange(4, len(g))] ^= 1 << rng.randrange(8)
            payloads.append(bytes(g)); want.append(oracle.decompress(bytes(g), len(data)))
        payloads.append(base[:len(base) // 2]); want.append(None)
    got = c.unpack_batch(payloads, [len(data)] * len(payloads), raise_on_error=False)
    assert got[0] == data and got[152] == data
    for g, w in zip(got, want):
        if g is not None:
            assert g == w


def test_deterministic_flag_gives_reproducible_frames(sq, oracle):
    """SQ_FLAG_DETERMINISTIC: packing the same chunks twice (two contexts) gives byte-identical frames, and they decode"""
    lib = sq.load()
    chunks = []
    for klass, n in ((0, 2 * MiB), (1, 2 * MiB), (2, 2 * MiB), (3, 2 * MiB), (0, 70000), (2, 300000)):
        b = C.create_string_buffer(n)
        lib.sq_corpus_fill_host(b, n, 5, klass + 11, klass)
        chunks.append(b.raw)
    runs = []
    for _ in range(3):
        c = sq.Context(deterministic=True)
        runs.append([f for _, f in c.pack_batch(chunks)])
    assert runs[0] == runs[1] == runs[2]
    for data, f in zip(chunks, runs[0]):
        assert oracle.decompress(f, 2 * MiB) == data
