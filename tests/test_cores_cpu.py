"""CPU tests of the product's zstd cores compiled for the host (tests/harness/*.cc include the very headers the CUDA kernels
use: zstd_core.h, zstd_enc_block.h, zstd_dec_core.h).  The checker is stock libzstd through the oracle.  No GPU, no compute
through the C ABI -- this is host-side coverage of the format logic that the kernels run per lane."""
import ctypes as C
import random
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
H = ROOT / "tests" / "harness"


def _build(name):
    so = H / f"lib{name}.so"
    src = H / f"{name}.cc"
    deps = [src] + list((ROOT / "squishrs_b200" / "csrc").glob("zstd_*.h"))
    if not so.exists() or any(d.stat().st_mtime > so.stat().st_mtime for d in deps):
        subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-I", str(ROOT / "squishrs_b200" / "csrc"), str(src), "-o", str(so)],
                       check=True, capture_output=True)
    return C.CDLL(str(so))


@pytest.fixture(scope="module")
def dec():
    lib = _build("dec_model")
    lib.dec_model_payload.restype = C.c_long
    lib.dec_model_payload.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32]

    lib.dec_model_payload_two_pass.restype = C.c_long
    lib.dec_model_payload_two_pass.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32]

    def run(payload: bytes, cap: int):
        out = C.create_string_buffer(max(cap, 1))
        n = lib.dec_model_payload(payload, len(payload), out, cap)
        return None if n < 0 else out.raw[:n]

    def two_pass(payload: bytes, cap: int):
        """the block-parallel decoder's two passes run in order on the host: bytes, or "n/a" when the payload is not eligible
        (several frames, chained tables, a block that reads repeat offsets from before itself, anything malformed) and the
        product runs the one-pass decoder, which also produces the errors"""
        out = C.create_string_buffer(max(cap, 1))
        n = lib.dec_model_payload_two_pass(payload, len(payload), out, cap)
        assert n >= 0 or n == -100
        return "n/a" if n == -100 else out.raw[:n]
    run.two_pass = two_pass
    return run


@pytest.fixture(scope="module")
def enc():
    """the CPU model of the shipped K3 parse (tests/harness/lz_model2.cc): same table shape and candidate rules as the search
    kernel, and the very parse function (zstd_enc_parse.h) and block writer (zstd_enc_block.h) the GPU compiles"""
    lib = _build("lz_model2")
    lib.lz_model2_shipped.restype = C.c_long
    lib.lz_model2_shipped.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32]

    def run(data: bytes):
        cap = len(data) + 4096
        dst = C.create_string_buffer(cap)
        n = lib.lz_model2_shipped(data, len(data), dst, cap)
        assert n > 0
        return dst.raw[:n]
    return run


def samples(sq, sizes):
    lib = sq.load()
    out = []
    for klass in range(7):
        for n in sizes:
            b = C.create_string_buffer(max(n, 1))
            lib.sq_corpus_fill_host(b, n, 321, klass * 7 + 3, klass)
            out.append(b.raw[:n])
    rng = random.Random(2)
    out += [rng.randbytes(3000), bytes(5000), bytes(rng.choices(range(5), k=40000)), bytes(rng.choices(range(256), weights=[1 / (i + 1) for i in range(256)], k=90000))]
    return out


def test_decoder_core_decodes_stock_libzstd_frames(sq, oracle, dec):
    for data in samples(sq, (0, 1, 17, 255, 256, 4096, 70000, 131072, 131073, 400000)):
        for lvl in (1, 3, 12, 19):
            if lvl == 19 and len(data) > 200000:
                continue
            assert dec(oracle.compress(data, lvl), len(data)) == data, (len(data), lvl)


def test_decoder_core_accepts_and_rejects_like_libzstd(oracle, dec):
    f = oracle.compress(bytes([42]) * 2048)
    two = oracle.compress(b"hello ") + oracle.compress(b"squish")
    skip = bytes.fromhex("502a4d18") + (3).to_bytes(4, "little") + b"abc"
    nofcs = bytes.fromhex("28b52ffd") + bytes([0x00, 0x58, 0x21, 0, 0]) + b"test"
    assert dec(two, 64) == b"hello squish" and dec(skip + two, 64) == b"hello squish" and dec(nofcs, 1 << 21) == b"test"
    for bad, cap in ((f, 2047), (f + b"x", 4096), (f[:-2], 4096), (b"\x00" * 12, 64)):
        assert dec(bad, cap) is None and oracle.decompress(bad, cap) is None


def test_decoder_core_survives_corrupted_frames(sq, oracle, dec):
    """bit flips must end in an error, never in a crash.  The core is deliberately stricter than libzstd in one place: it
    rejects a sequence bitstream that is over-read (libzstd tolerates overflow and decodes implementation-defined bits), so
    the property checked is: whatever the core accepts, stock libzstd accepts too, with identical bytes."""
    rng = random.Random(4)
    data = samples(sq, (30000,))[2]
    f = bytearray(oracle.compress(data, 12))
    for _ in range(1500):
        g = bytearray(f)
        for _ in range(rng.randrange(1, 4)):
            g[rng.randrange(4, len(g))] ^= 1 << rng.randrange(8)
        got = dec(bytes(g), len(data))
        ref = oracle.decompress(bytes(g), len(data))
        if got is not None:
            assert ref == got


def test_encoder_core_frames_decode_with_stock_libzstd_and_ratio(sq, oracle, enc, dec):
    tot_gpu = tot_ref = 0
    for data in samples(sq, (1, 100, 4096, 70000, 300000, 2 << 20)):
        frame = enc(data)
        assert oracle.decompress(frame, len(data)) == data
        assert dec(frame, len(data)) == data
        if len(data) == 2 << 20:
            tot_gpu += len(frame)
            tot_ref += len(oracle.compress(data, 12))
    assert tot_gpu <= 1.03 * tot_ref, (tot_gpu, tot_ref)  # the north_star tolerance on the parse model


def test_two_pass_decoder_core(sq, oracle, enc, dec):
    """zstd_dec_core.h's two-pass path (table snapshots; per block: entropy decoding, symbolic repeat offsets, literal placement;
    then the matches in order): frames of two or more blocks decode to the same bytes whether the encoder core or stock libzstd
    wrote them (Repeat_Mode, Treeless literals and repeat offsets across blocks included); single-block and multi-frame payloads
    are left to the one-pass decoder; too small a capacity and bit flips never crash it and never produce bytes the reference
    decoder would not."""
    rng = random.Random(6)
    two_pass_ref = 0
    for data in samples(sq, (1, 100, 4096, 70000, 131072, 131073, 300000, 2 << 20)):
        frame = enc(data)
        got = dec.two_pass(frame, len(data))
        assert got == data if len(data) > 131072 and got != "n/a" else got in ("n/a", data), len(data)
        if len(data) > 131072 and len(set(data)) > 1:
            assert dec.two_pass(frame, len(data) - 1) == "n/a"               # capacity: left to the one-pass decoder, which reports it
        for lvl in (1, 3, 12, 19):
            if lvl == 19 and len(data) > 400000:
                continue
            ref = oracle.compress(data, lvl)
            got = dec.two_pass(ref, len(data))
            assert got == "n/a" or got == data, (len(data), lvl)
            two_pass_ref += got != "n/a"
    assert two_pass_ref >= 20, two_pass_ref  # stock frames of several blocks do take the two-pass path
    assert dec.two_pass(oracle.compress(b"hello ") + oracle.compress(b"squish"), 64) == "n/a"  # two frames: one-pass decoder
    for maker in (enc, lambda d: oracle.compress(d, 12)):
        data = samples(sq, (400000,))[2]
        f = bytearray(maker(data))
        for _ in range(500):
            g = bytearray(f)
            for _ in range(rng.randrange(1, 4)):
                g[rng.randrange(4, len(g))] ^= 1 << rng.randrange(8)
            got = dec.two_pass(bytes(g), len(data))
            ref = oracle.decompress(bytes(g), len(data))
            if got != "n/a":
                assert ref == got


def test_encoder_core_ratio_on_real_files(oracle, enc):
    """the parse model (the shipped candidate rules, parse function and block writer) on real files of this image -- source code
    and a shared library, 2 MiB chunks and 24 KB pieces -- stays within the north_star's 3 % of libzstd level 12; the GPU
    repeats this on 32 MB per corpus (test_real_data_ratio_within_3pct_of_level12)"""
    import glob
    def blob(pattern, limit):
        out = bytearray()
        for f in sorted(glob.glob(pattern, recursive=True)):
            try:
                out += open(f, "rb").read()
            except OSError:
                continue
            if len(out) >= limit:
                break
        return bytes(out[:limit])
    corpora = {"python sources": blob("/usr/lib/python3*/**/*.py", 4 << 20), "shared libraries": blob("/usr/lib/x86_64-linux-gnu/*.so*", 8 << 20)}
    if any(len(v) < (2 << 20) for v in corpora.values()):
        pytest.skip("this image does not have the files")
    for name, data in corpora.items():
        chunks = [data[i:i + (2 << 20)] for i in range(0, len(data), 2 << 20)] + [data[i:i + 24000] for i in range(0, 480000, 24000)]
        ours = ref = 0
        for c in chunks:
            f = enc(c)
            assert oracle.decompress(f, len(c)) == c
            ours += len(f)
            ref += len(oracle.compress(c, 12))
        assert ours <= 1.03 * ref, (name, ours, ref)
