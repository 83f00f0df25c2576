"""Dev tool: device-resident decode throughput of K4 on frames written by K3 (and optionally by libzstd L12)."""
import ctypes as C, sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import squishrs_b200 as sq
from squishrs_b200 import _lib as L
from bench import corpus_plan, CHUNK, SEED
lib = sq.load(); n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
use_ref = len(sys.argv) > 2 and sys.argv[2] == "ref"
ctx = sq.Context(max_batch_chunks=max(n, 4096))
ids, kl = corpus_plan(n, dup_frac=0.0)
corpus = torch.empty(n * CHUNK, dtype=torch.uint8, device="cuda")
d_ids = torch.from_numpy(ids.astype(np.int64)).cuda(); d_kl = torch.from_numpy(kl.astype(np.int32)).cuda()
ctx.check(lib.sq_corpus_fill_slots_device(ctx.h, corpus.data_ptr(), CHUNK, d_ids.data_ptr(), d_kl.data_ptr(), n, SEED, None))
ctx.check(lib.sq_synchronize(ctx.h, None))
sp = np.zeros((n, 2), dtype=np.uint64); sp[:, 0] = np.arange(n) * CHUNK; sp[:, 1] = CHUNK
d_sp = torch.from_numpy(sp.view(np.int64)).cuda()
cap = n * int(lib.sq_encode_bound(CHUNK)); out = torch.empty(cap + 64, dtype=torch.uint8, device="cuda")
res = torch.empty(n * 32, dtype=torch.uint8, device="cuda"); used = C.c_uint64()
ctx.check(lib.sq_pack_device(ctx.h, corpus.data_ptr(), d_sp.data_ptr(), n, 0, res.data_ptr(), out.data_ptr(), cap, C.byref(used), None))
r = np.frombuffer(res.cpu().numpy().tobytes(), dtype=np.dtype([("d", "u1", 16), ("off", "<u8"), ("len", "<u4"), ("new", "u1"), ("pad", "u1", 3)]))
if use_ref:
    from conftest import Oracle
    O = Oracle(); host = corpus.cpu().numpy()
    frames = [O.compress(host[i * CHUNK:(i + 1) * CHUNK].tobytes(), 12) for i in range(n)]
    offs = np.cumsum([0] + [(len(f) + 15) & ~15 for f in frames]); blob = bytearray(int(offs[-1]) + 64)
    for f, o in zip(frames, offs): blob[o:o + len(f)] = f
    out = torch.frombuffer(blob, dtype=torch.uint8).cuda(); foff = offs[:-1]; flen = np.array([len(f) for f in frames])
else:
    foff, flen = r["off"], r["len"]
fr = np.zeros(n, dtype=np.dtype([("src", "<u8"), ("dst", "<u8"), ("len", "<u4"), ("cap", "<u4")]))
fr["src"] = foff; fr["dst"] = np.arange(n) * CHUNK; fr["len"] = flen; fr["cap"] = CHUNK
d_fr = torch.frombuffer(bytearray(fr.tobytes()), dtype=torch.uint8).cuda()
dec = torch.empty(n * CHUNK, dtype=torch.uint8, device="cuda"); d_res = torch.empty(n * 8, dtype=torch.uint8, device="cuda")
for rep in range(3):
    torch.cuda.synchronize(); t = time.time()
    ctx.check(lib.sq_decode_device(ctx.h, out.data_ptr(), d_fr.data_ptr(), n, dec.data_ptr(), d_res.data_ptr(), None))
    ctx.check(lib.sq_synchronize(ctx.h, None)); dt = time.time() - t
ok = bool(torch.equal(dec, corpus)); st = np.frombuffer(d_res.cpu().numpy().tobytes(), dtype="<i4").reshape(-1, 2)
print(f"decode {n} frames ({'libzstd L12' if use_ref else 'own'}): {dt*1e3:.1f} ms  {n*CHUNK/dt/1e9:.2f} GB/s out  identical={ok} errors={(st[:,1]!=0).sum()} comp={int(flen.sum())/1e6:.1f} MB")
