"""world_size-2 gloo test (CPU) of the multi-GPU dedup exchange: the routing / all-to-all / verdict-return
plumbing in squishrs_b200/sharded.py, with the three device steps replaced by a numpy stand-in that follows
the same record layout and winner rule.  The CUDA kernels themselves are covered by -m gpu tests."""
import os
import random
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


class NumpyOps:
    """CPU stand-in with the device ops' semantics (tests only)."""

    def __init__(self):
        self.index = {}  # digest -> lowest gidx seen (this rank's shard)

    def route(self, digests, gidx_base, n, world, cap, send, send_pos):
        from squishrs_b200.sharded import INVALID, REC_BYTES, owner_of
        rec = np.zeros((world * cap, REC_BYTES), dtype=np.uint8)
        rec[:, 16:24] = 0xFF  # padding records: gidx = ~0
        counts = [0] * world
        d = digests.numpy().reshape(-1, 16)
        pos = np.zeros(cap, dtype=np.int32)
        for i in range(n):
            o = owner_of(bytes(d[i]), world)
            k = o * cap + counts[o]
            counts[o] += 1
            rec[k, :16] = d[i]
            rec[k, 16:24] = np.frombuffer(int(gidx_base + i).to_bytes(8, "little"), dtype=np.uint8)
            pos[i] = k
        send.copy_(torch.from_numpy(rec.reshape(-1)))
        send_pos.copy_(torch.from_numpy(pos))

    def insert(self, recv, count, verdict):
        from squishrs_b200.sharded import INVALID
        rec = recv.numpy().reshape(-1, 32)
        recs = []
        for i in range(count):
            g = int.from_bytes(bytes(rec[i, 16:24]), "little")
            if g != INVALID:
                recs.append((bytes(rec[i, :16]), g, i))
        for dig, g, _ in recs:  # pass 1: atomicMin
            self.index[dig] = min(self.index.get(dig, g), g)
        v = np.zeros(count, dtype=np.uint8)
        for dig, g, i in recs:  # pass 2: verdicts
            v[i] = 1 if self.index[dig] == g else 0
        verdict.copy_(torch.from_numpy(v))

    def unroute(self, verdict_back, send_pos, n, is_new):
        is_new[:n] = verdict_back[send_pos[:n].long()]


def _worker(rank, world, port, batches, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from squishrs_b200.sharded import ShardedDedup
    cap = max(len(x) for b in batches for x in b)  # one capacity for every rank: the exchange uses equal splits
    sd = ShardedDedup(NumpyOps(), world, cap, "cpu")
    verdicts = []
    for step, per_rank in enumerate(batches):
        mine = per_rank[rank]
        n = len(mine)
        dig = torch.from_numpy(np.frombuffer(b"".join(mine) if n else b"", dtype=np.uint8).copy()) if n else torch.empty(0, dtype=torch.uint8)
        is_new = torch.zeros(cap, dtype=torch.uint8)
        gidx_base = step * world * cap + rank * cap  # (step, rank, slot) order == processing order
        sd.exchange(dig, gidx_base, n, is_new)
        verdicts.append(is_new[:n].tolist())
    owned = len(sd.ops.index)
    t = torch.tensor([owned])
    dist.all_reduce(t)
    out_q.put((rank, verdicts, int(t.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_dedup_two_ranks_matches_global_winner_rule():
    world = 2
    rng = random.Random(12)
    uniq = [rng.randbytes(16) for _ in range(40)]
    batches = []
    for step in range(3):
        batches.append([[uniq[rng.randrange(len(uniq))] for _ in range(rng.randrange(5, 30))] for _ in range(world)])
    batches.append([[], [uniq[0], uniq[0]]])  # ragged: one rank has nothing this step
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + rng.randrange(2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, batches, q)) for r in range(world)]
    [p.start() for p in procs]
    results = dict()
    for _ in range(world):
        r, v, total = q.get(timeout=120)
        results[r] = (v, total)
    [p.join(timeout=60) for p in procs]
    # expected: walk chunks in (step, rank, slot) order; first occurrence of a digest wins
    seen = set()
    for step, per_rank in enumerate(batches):
        for r in range(world):
            want = []
            for d in per_rank[r]:
                want.append(0 if d in seen else 1)
                seen.add(d)
            assert results[r][0][step] == want, (step, r)
    assert results[0][1] == results[1][1] == len(seen)  # len() of the sharded store = sum of per-rank shards


def test_owner_is_digest_prefix():
    from squishrs_b200.sharded import owner_of
    d = bytes(range(16))
    assert owner_of(d, 8) == int.from_bytes(d[:8], "little") % 8
    assert {owner_of(bytes([i]) + bytes(15), 4) for i in range(8)} == {0, 1, 2, 3}
