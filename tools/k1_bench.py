"""Dev tool: K1 digest throughput on a batch larger than L2, CUDA-event timed, + parity spot check."""
import ctypes as C, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import squishrs_b200 as sq
from bench import CHUNK
lib = sq.load(); ctx = sq.Context(max_batch_chunks=8192); n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
data = torch.randint(0, 256, (n * CHUNK,), dtype=torch.uint8, device="cuda")
sp = np.zeros((n, 2), dtype=np.uint64); sp[:, 0] = np.arange(n) * CHUNK; sp[:, 1] = CHUNK
d_sp = torch.from_numpy(sp.view(np.int64)).cuda(); dig = torch.empty(n * 16, dtype=torch.uint8, device="cuda")
st = torch.cuda.Stream(); torch.cuda.set_stream(st); spn = C.c_void_p(st.cuda_stream)
ts = []
for i in range(8):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st); ctx.check(lib.sq_digest_device(ctx.h, data.data_ptr(), d_sp.data_ptr(), n, dig.data_ptr(), spn)); e1.record(st)
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
best = min(ts[2:]); med = sorted(ts[2:])[len(ts[2:]) // 2]
from conftest import Oracle
O = Oracle(); h = data[:CHUNK].cpu().numpy().tobytes()
ok = bytes(dig[:16].cpu().numpy()) == O.hash_chunk(h)
print(f"K1 {n} chunks ({n*CHUNK/2**30:.0f} GiB): best {best:.3f} ms {n*CHUNK/best/1e6:.0f} GB/s, median {med:.3f} ms {n*CHUNK/med/1e6:.0f} GB/s, parity={ok}")
