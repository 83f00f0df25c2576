"""Small pack + unpack over every corpus class and ragged sizes, meant to run under compute-sanitizer memcheck."""
import ctypes as C, sys, random
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import squishrs_b200 as sq
from conftest import Oracle
O = Oracle(); lib = sq.load(); ctx = sq.Context(max_batch_chunks=256)
rng = random.Random(3); chunks = []
for k in range(7):
    for n in (1, 7, 8, 9, 23, 24, 25, 40, 63, 300, 1023, 1025, 4096, 70001, 131072, 131073, 262144 + 5, 2 << 20):
        b = C.create_string_buffer(max(n, 1)); lib.sq_corpus_fill_host(b, n, 11, k * 100 + n % 97, k); chunks.append(b.raw[:n])
chunks += [rng.randbytes(n) for n in (1, 100, 5000, 2 << 20)]
res = ctx.pack_batch(chunks)
frames = [(c, f) for c, (_, f) in zip(chunks, res) if f is not None]
for c, f in frames:
    assert O.decompress(f, len(c)) == c
back = ctx.unpack_batch([f for _, f in frames], [len(c) for c, _ in frames])
assert back == [c for c, _ in frames]
ref_frames = [O.compress(c, 12) for c, _ in frames[:60]]
assert ctx.unpack_batch(ref_frames, [2 << 20] * len(ref_frames)) == [c for c, _ in frames[:60]]
print("sanitize smoke ok:", len(chunks), "chunks")
