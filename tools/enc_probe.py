"""Dev tool: device-resident pack of N chunks of the bench corpus (default 2048), best of 3; prints GB/s and output bytes.
Used for A/B runs of encoder variants (env knobs / rebuilt library) on the GPU box."""
import ctypes as C, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import squishrs_b200 as sq
from bench import corpus_plan, CHUNK, SEED
import os
lib = sq.load(); ctx = sq.Context(max_batch_chunks=2048, deterministic=bool(os.environ.get("SQ_DETERMINISTIC")))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
ids, kl = corpus_plan(n)
corpus = torch.empty(n * CHUNK, dtype=torch.uint8, device="cuda")
d_ids = torch.from_numpy(ids.astype(np.int64)).cuda(); d_kl = torch.from_numpy(kl.astype(np.int32)).cuda()
ctx.check(lib.sq_corpus_fill_slots_device(ctx.h, corpus.data_ptr(), CHUNK, d_ids.data_ptr(), d_kl.data_ptr(), n, SEED, None))
ctx.check(lib.sq_synchronize(ctx.h, None))
sp = np.zeros((n, 2), dtype=np.uint64); sp[:, 0] = np.arange(n) * CHUNK; sp[:, 1] = CHUNK
d_sp = torch.from_numpy(sp.view(np.int64)).cuda()
cap = n * int(lib.sq_encode_bound(CHUNK)); out = torch.empty(cap, dtype=torch.uint8, device="cuda")
res = torch.empty(n * 32, dtype=torch.uint8, device="cuda"); used = C.c_uint64()
best = 1e9
for rep in range(4):
    ctx.dedup_reset()
    torch.cuda.synchronize(); t = time.time()
    ctx.check(lib.sq_pack_device(ctx.h, corpus.data_ptr(), d_sp.data_ptr(), n, 0, res.data_ptr(), out.data_ptr(), cap, C.byref(used), None))
    ctx.check(lib.sq_synchronize(ctx.h, None))
    dt = time.time() - t
    if rep: best = min(best, dt)
print(f"{sys.argv[2] if len(sys.argv) > 2 else ''} pack {n} chunks: {best*1e3:.1f} ms  {n*CHUNK/best/1e9:.2f} GB/s  out {used.value} B  sum {int(out[:used.value].to(torch.int64).sum())}")
