#!/usr/bin/env python
"""bench.py — device-resident pack throughput of the squishRS data path on B200.

Workload (BASELINE.json configs[1]): synthetic mixed corpus (40% log lines / 30% JSON / 30% binary
records by 2 MiB slot, 20% duplicate slots), 64 GiB per GPU when it fits, packed in batches of
`--batch-chunks` 2 MiB chunk slots.  One "step" = one pass of the hot path (K1 digest -> K2 dedup
[-> digest all-to-all at N>1] -> K3 zstd encode of the winners) over one batch that is already
resident in HBM.  `value` = input GB/s over all ranks; `e2e` = the same through the host-buffer
C-ABI call sq_pack_host (H2D + kernels + D2H inside the timed region).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One JSON line on stdout (rank 0).  Everything else goes to stderr.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
MiB = 1 << 20
GiB = 1 << 30
CHUNK = 2 * MiB
SEED = 0x51510002


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# Libraries (NCCL prints its version banner on stdout) must not pollute the one-JSON-line contract: fd 1 is pointed at
# stderr for the whole run and the result line is written to the saved original stdout.
_REAL_STDOUT = None


def protect_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


# --------------------------------------------------------------------------- corpus definition
WORKLOADS = {
    "config2": "configs[1]: mixed logs/JSON/binary corpus, 20% duplicate 2 MiB slots, device-resident pack",
    "config3": "configs[2]: VM-image-like corpus (35% zero / 45% fs-like mixed / 20% random slots, ~70% duplicate slots), device-resident pack",
    "config4": "configs[3]: incompressible random corpus, 0% duplicates (raw-block fallback, hash/dedup ceiling)",
}


def corpus_plan(n_slots: int, first_slot: int = 0, dup_frac: float = 0.2, seed: int = SEED, workload: str = "config2"):
    """Slot table of a corpus stream: (payload id, class) per 2 MiB slot.  A duplicate slot copies an EARLIER slot
    of the global stream (so duplicates cross rank shards): uniformly chosen for config2, Zipf-popular for config3."""
    import numpy as np
    total = first_slot + n_slots
    rng = np.random.default_rng(seed + {"config2": 0, "config3": 1, "config4": 2}[workload])
    u = rng.random(total)
    if workload == "config2":
        klass = np.where(u < 0.4, 1, np.where(u < 0.7, 2, 3)).astype(np.uint32)  # 1 log, 2 json, 3 binary
        is_dup = rng.random(total) < dup_frac
        src = (rng.random(total) * np.arange(total)).astype(np.int64)
    elif workload == "config3":
        klass = np.where(u < 0.35, 5, np.where(u < 0.80, 6, 4)).astype(np.uint32)  # 5 zeros, 6 fs-like mix, 4 random
        is_dup = rng.random(total) < 0.54  # + all-zero slots being duplicates of each other gives ~70% duplicate slots
        src = (np.arange(total) * rng.random(total) ** 3).astype(np.int64)  # popular (early) slots are copied more often
    else:
        klass = np.full(total, 4, dtype=np.uint32)
        is_dup = np.zeros(total, dtype=bool)
        src = np.zeros(total, dtype=np.int64)
    ids = np.arange(total, dtype=np.uint64)
    ids[klass == 5] = 0  # every all-zero slot is the same payload
    is_dup[0] = False
    for i in range(total):  # resolve copy chains so ids[i] names the original payload
        if is_dup[i]:
            ids[i] = ids[src[i]]
            klass[i] = klass[src[i]]
    return ids[first_slot:], klass[first_slot:]


def expected_unique(ids) -> int:
    import numpy as np
    return int(len(np.unique(ids)))


# --------------------------------------------------------------------------- clocks sampler
class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception as e:  # nvidia-smi missing: report empty clocks
            log("clocks sampler unavailable:", e)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [int(float(r[2])) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------- reference arm (CPU)
def load_oracle():
    sys.path.insert(0, str(ROOT / "tests"))
    from conftest import Oracle
    return Oracle()


def host_corpus(lib, ids, klass):
    """Generate slots on the host (same generator as the device) into one contiguous buffer."""
    n = len(ids)
    buf = C.create_string_buffer(n * CHUNK)
    base = C.addressof(buf)
    nthreads = min(os.cpu_count() or 1, 32)

    def work(t):
        for i in range(t, n, nthreads):
            lib.sq_corpus_fill_host(C.c_void_p(base + i * CHUNK), CHUNK, SEED, int(ids[i]), int(klass[i]))
    ths = [threading.Thread(target=work, args=(t,)) for t in range(nthreads)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    return buf


def cpu_pack(oracle, buf, n_slots: int, threads: int, files: int = 0):
    """Reference CPU path on an in-memory sample: thread-per-file over `threads` workers, 2 MiB chunks,
    XXH3-128, concurrent digest set, zstd level 12 per unique chunk, one writer thread -> /dev/shm."""
    files = files or max(1, min(n_slots, threads * 4))
    per = (n_slots + files - 1) // files
    arr = (oracle.File * files)()
    base = C.addressof(buf)
    names = []
    k = 0
    for f in range(files):
        cnt = min(per, n_slots - k)
        if cnt <= 0:
            files = f
            break
        names.append(f"f{f}".encode())
        arr[f].rel_path = names[-1]
        arr[f].path_on_disk = None
        arr[f].data = base + k * CHUNK
        arr[f].size = cnt * CHUNK
        k += cnt
    out = f"/dev/shm/sq_ref_{os.getpid()}.squish"
    st = oracle.Stats()
    t0 = time.perf_counter()
    rc = oracle.L.sqo_pack(arr, files, out.encode(), threads, 1760000000, 0, C.byref(st))
    dt = time.perf_counter() - t0
    try:
        os.unlink(out)
    except OSError:
        pass
    assert rc == 0, f"oracle pack failed rc={rc}"
    return dt, st


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import squishrs_b200 as sq
    lib = sq.load()
    oracle = load_oracle()
    cores = os.cpu_count() or 1
    n = args.ref_chunks or max(64, min(2048, cores * 16))
    ids, klass = corpus_plan(n, workload=args.workload)
    buf = host_corpus(lib, ids, klass)
    for _ in range(args.warmup):
        cpu_pack(oracle, buf, min(n, max(8, cores)), cores)
    t = 0.0
    stats = None
    for _ in range(args.steps):
        dt, stats = cpu_pack(oracle, buf, n, cores)
        t += dt
    gbs = args.steps * n * CHUNK / t / 1e9
    line = {"impl": "reference", "metric": "pack_gb_per_s", "value": gbs, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload] + " (CPU sample)", "sample_chunks": n, "chunk_bytes": CHUNK,
                       "zstd_level": 12, "libzstd": oracle.L.sqo_zstd_version()},
            "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": cores, "kind": "port",
                             "sample": f"{n} x 2 MiB slots of the configs[1] stream per step, in memory, archive to /dev/shm"},
            "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "ratio": stats.payload_bytes / (stats.unique_chunks * CHUNK) if stats and stats.unique_chunks else None}
    emit(line)


# --------------------------------------------------------------------------- our arm (GPU)
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import squishrs_b200 as sq
    from squishrs_b200 import _lib as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = sq.load()
    B = args.batch_chunks
    free, total_mem = torch.cuda.mem_get_info()
    # resident corpus: as much of 64 GiB as fits next to the output/scratch buffers
    want_batches = max(1, (args.corpus_gib * GiB) // (B * CHUNK))
    budget = int(free * 0.55)
    n_batches = int(max(1, min(want_batches, budget // (B * CHUNK))))
    n_slots = n_batches * B
    if world > 1 and not args.single_stream:
        # leave two SMs' worth of search-CTA slots free: the digest / exchange kernels of the step after next (and NCCL's copy
        # kernels) then run beside the encoder instead of waiting for its tail
        os.environ.setdefault("SQ_LZ_CTAS_TOTAL", str(3 * (torch.cuda.get_device_properties(local).multi_processor_count - 2)))
    ctx = sq.Context(device=local, dedup_capacity=max(1 << 20, 4 * n_slots * max(world, 1) * (args.steps + args.warmup + 4) // max(n_batches, 1) + n_slots * 4),
                     max_batch_chunks=B * max(world, 1))
    ids, klass = corpus_plan(n_slots, first_slot=rank * n_slots, workload=args.workload)
    corpus = torch.empty(n_slots * CHUNK, dtype=torch.uint8, device="cuda")
    d_ids = torch.from_numpy(ids.astype(np.int64)).cuda()
    d_kl = torch.from_numpy(klass.astype(np.int32)).cuda()
    stream = torch.cuda.Stream()  # a real (non-NULL) stream: the library launches on it and the events time it
    torch.cuda.set_stream(stream)
    sp = C.c_void_p(stream.cuda_stream)
    assert stream.cuda_stream != 0
    t0 = time.perf_counter()
    ctx.check(lib.sq_corpus_fill_slots_device(ctx.h, corpus.data_ptr(), CHUNK, d_ids.data_ptr(), d_kl.data_ptr(), n_slots, SEED, sp))
    torch.cuda.synchronize()
    log(f"[rank {rank}] corpus: {n_slots} slots ({n_slots * CHUNK / GiB:.1f} GiB) generated in {time.perf_counter() - t0:.2f}s; "
        f"unique payloads {expected_unique(ids)}")

    spans_np = np.zeros((B, 2), dtype=np.uint64)
    spans_np[:, 0] = np.arange(B, dtype=np.uint64) * CHUNK
    spans_np[:, 1] = CHUNK  # len in the low 32 bits, reserved = 0
    d_spans = torch.from_numpy(spans_np.view(np.int64)).cuda()
    out_cap = B * int(lib.sq_encode_bound(CHUNK))

    ev = lambda: torch.cuda.Event(enable_timing=True)
    stage_ms = {"digest": 0.0, "dedup": 0.0, "encode": 0.0}
    totals = {"in": 0, "out": 0, "new": 0, "chunks": 0}

    sd = None
    if world > 1:
        from squishrs_b200.sharded import DeviceOps, ShardedDedup
        sd = ShardedDedup(DeviceOps(ctx, sp), world, B, "cuda")

    # Two streams alternate between steps: the long encode of step k+1 starts while the last chunks of step k are still in
    # flight (one chunk occupies one search CTA for ~0.1 s, so a step's tail would otherwise leave most SMs idle).  The
    # library gives each stream its own encoder scratch; digest + dedup of consecutive steps are ordered with events.
    stream_b = torch.cuda.Stream()
    streams = [stream, stream_b] if not args.single_stream else [stream, stream]
    sps = [C.c_void_p(x.cuda_stream) for x in streams]
    bufs = []
    for _ in range(2):
        bufs.append(dict(out=torch.empty(out_cap, dtype=torch.uint8, device="cuda"), dig=torch.empty(B * 16, dtype=torch.uint8, device="cuda"),
                         new=torch.empty(B, dtype=torch.uint8, device="cuda"), foff=torch.empty(B, dtype=torch.int64, device="cuda"),
                         flen=torch.empty(B, dtype=torch.int32, device="cuda"), total=torch.zeros(1, dtype=torch.int64, device="cuda")))
    d_out, d_dig, d_new, d_foff, d_flen, d_total = (bufs[0][k] for k in ("out", "dig", "new", "foff", "flen", "total"))
    last_dedup = [None]
    # Multi-GPU: digest + digest exchange ("front") of step i+2 run on their own stream while steps i and i+1 encode, so a rank
    # never waits for its peers inside a step: the all-to-all it needs next was finished one step ago.  Three small
    # (digest, verdict) sets rotate; the search kernel leaves a few CTA slots free (SQ_LZ_CTAS_TOTAL) so the front kernels and
    # NCCL's copy kernels can run beside it.
    ahead = sd is not None and not args.single_stream
    front_stream = torch.cuda.Stream() if ahead else None
    fsp = C.c_void_p(front_stream.cuda_stream) if ahead else None
    fsets = [dict(dig=torch.empty(B * 16, dtype=torch.uint8, device="cuda"), new=torch.empty(B, dtype=torch.uint8, device="cuda")) for _ in range(3)] if ahead else None
    front_done, enc_done = {}, {}

    def front(i, gidx_base):  # digest + exchange of step i on the front stream
        fs = fsets[i % 3]
        base = corpus.data_ptr() + (i % n_batches) * B * CHUNK
        with torch.cuda.stream(front_stream):
            if i - 3 in enc_done:  # the set's previous reader
                front_stream.wait_event(enc_done[i - 3])
            ctx.check(lib.sq_digest_device(ctx.h, base, d_spans.data_ptr(), B, fs["dig"].data_ptr(), fsp))
            sd.ops.sp = fsp
            sd.exchange(fs["dig"], gidx_base, B, fs["new"])
            e_ = torch.cuda.Event()
            e_.record(front_stream)
            front_done[i] = e_

    def back(i, timed):  # encode of step i on its own stream, once its verdicts are there
        b = i % n_batches
        st, spx, bf, fs = streams[i % 2], sps[i % 2], bufs[i % 2], fsets[i % 3]
        base = corpus.data_ptr() + b * B * CHUNK
        with torch.cuda.stream(st):
            st.wait_event(front_done[i])
            ctx.check(lib.sq_encode_device(ctx.h, base, d_spans.data_ptr(), fs["new"].data_ptr(), B, bf["out"].data_ptr(), out_cap,
                                           bf["foff"].data_ptr(), bf["flen"].data_ptr(), bf["total"].data_ptr(), spx))
            e_ = torch.cuda.Event()
            e_.record(st)
            enc_done[i] = e_
            tot = (bf["total"].clone(), fs["new"].sum(dtype=torch.int64)) if timed else None
        return None, tot

    def run_ahead(n_steps, gidx_of, timed):
        front_done.clear(); enc_done.clear()
        res = []
        for j in range(min(2, n_steps)):
            front(j, gidx_of(j))
        for i in range(n_steps):
            res.append(back(i, timed))
            if i + 2 < n_steps:
                front(i + 2, gidx_of(i + 2))
        return res

    def step(i, gidx_base, timed):
        b = i % n_batches
        st, spx, bf = streams[i % 2], sps[i % 2], bufs[i % 2]
        base = corpus.data_ptr() + b * B * CHUNK
        e = [ev() for _ in range(4)] if timed else None
        with torch.cuda.stream(st):
            if last_dedup[0] is not None:
                st.wait_event(last_dedup[0])
            if timed:
                e[0].record(st)
            ctx.check(lib.sq_digest_device(ctx.h, base, d_spans.data_ptr(), B, bf["dig"].data_ptr(), spx))
            if timed:
                e[1].record(st)
            if sd is None:
                ctx.check(lib.sq_dedup_insert_device(ctx.h, bf["dig"].data_ptr(), None, gidx_base, B, bf["new"].data_ptr(), spx))
            else:  # digest all-to-all over NCCL to the owner ranks, verdicts back
                sd.ops.sp = spx
                sd.exchange(bf["dig"], gidx_base, B, bf["new"])
            dd = torch.cuda.Event()
            dd.record(st)
            last_dedup[0] = dd
            if timed:
                e[2].record(st)
            ctx.check(lib.sq_encode_device(ctx.h, base, d_spans.data_ptr(), bf["new"].data_ptr(), B, bf["out"].data_ptr(), out_cap,
                                           bf["foff"].data_ptr(), bf["flen"].data_ptr(), bf["total"].data_ptr(), spx))
            if timed:
                e[3].record(st)
            tot = (bf["total"].clone(), bf["new"].sum(dtype=torch.int64)) if timed else None
        return e, tot

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up on a throwaway index state
    if ahead:
        run_ahead(args.warmup, lambda w: w * B * world + rank * B, False)
    else:
        for w in range(args.warmup):
            step(w, w * B * world + rank * B, False)
    barrier()
    ctx.dedup_reset()
    launches0 = C.c_uint64()
    lib.sq_kernel_launches(ctx.h, C.byref(launches0))
    clocks = Clocks(local)
    clocks.start()
    time.sleep(0.3)
    barrier()
    start, stop = ev(), ev()
    start.record(stream)
    evs = []
    outs = []
    if ahead:
        for e_k, tot_k in run_ahead(args.steps, lambda k: k * B * world + rank * B, True):
            evs.append(e_k)
            outs.append(tot_k)
        stream.wait_stream(front_stream)
    else:
        for k in range(args.steps):
            e_k, tot_k = step(k, k * B * world + rank * B, True)
            evs.append(e_k)
            outs.append(tot_k)  # tiny device-side reads on the step's stream, no sync
    stream.wait_stream(stream_b)  # the timed region ends when BOTH streams have drained
    stop.record(stream)
    barrier()
    elapsed_ms = start.elapsed_time(stop)
    clk = clocks.stop()
    launches1 = C.c_uint64()
    lib.sq_kernel_launches(ctx.h, C.byref(launches1))
    rc = lib.sq_encode_status(ctx.h)
    if rc != 0:
        raise SystemExit(f"encode overflow: {rc}")
    # Per-stage durations for the roofline: with two streams the launches of consecutive steps overlap on the device, so their
    # event durations double-count.  A short single-stream pass on fresh dedup state (same batches, same kernels, CUDA events on
    # the launching stream) gives clean per-launch times; the headline value above is untouched by it.
    overlapped = not args.single_stream
    if overlapped:
        ctx.dedup_reset()
        last_dedup[0] = None
        save_streams, save_sps = list(streams), list(sps)
        streams[1], sps[1] = streams[0], sps[0]
        prof = [step(k, k * B * world + rank * B, True) for k in range(min(args.steps, 3))]
        barrier()
        streams[:], sps[:] = save_streams, save_sps
        stage_evs, stage_outs, stage_steps = [p[0] for p in prof], [p[1] for p in prof], len(prof)
    else:
        stage_evs, stage_outs, stage_steps = evs, outs, args.steps
    for e in stage_evs:
        stage_ms["digest"] += e[0].elapsed_time(e[1])
        stage_ms["dedup"] += e[1].elapsed_time(e[2])
        stage_ms["encode"] += e[2].elapsed_time(e[3])
    stage_in = stage_steps * B * CHUNK
    stage_out = sum(int(t.item()) for t, _ in stage_outs)
    stage_new = sum(int(nw.item()) for _, nw in stage_outs)
    for tot, new in outs:
        totals["out"] += int(tot.item())
        totals["new"] += int(new.item())
    totals["chunks"] = args.steps * B
    totals["in"] = args.steps * B * CHUNK
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
    agg = torch.tensor([totals["in"], totals["out"], totals["new"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg)
    elapsed_ms = float(t.item())
    in_bytes, out_bytes, n_new = (float(x) for x in agg.tolist())
    value = in_bytes / (elapsed_ms * 1e-3) / 1e9

    # ---- e2e through the host-buffer C-ABI call (rank-local; pinned input, H2D + kernels + D2H) ----
    e2e = None
    if not args.no_e2e:
        # Through the reference-facing host-buffer calls: sq_pack_submit / sq_pack_wait, two batches in flight, inputs in
        # pinned host memory.  Every step uploads its batch (H2D) and downloads results + frames (D2H) inside the timed region.
        eb = min(B, args.e2e_chunks)
        ocap = eb * int(lib.sq_encode_bound(CHUNK))
        ctx.dedup_reset()
        n_e2e = max(4, min(args.steps, 8))
        n_warm = 2
        try:  # every step's input is pinned up front: stay well inside this rank's share of the host memory that is free
            avail = [int(l.split()[1]) * 1024 for l in open("/proc/meminfo") if l.startswith("MemAvailable:")][0]
            fit = int((avail * 0.4 / max(world, 1) - 2 * ocap) // (eb * CHUNK)) - n_warm  # ranks of one box share the host memory
            n_e2e = max(2, min(n_e2e, fit))
        except Exception:
            pass
        if world > 1:  # every rank times the same number of steps
            t_n = torch.tensor([n_e2e], dtype=torch.int64, device="cuda")
            dist.all_reduce(t_n, op=dist.ReduceOp.MIN)
            n_e2e = int(t_n.item())
        total_steps = n_warm + n_e2e
        # Every step's input sits in its own pinned host buffer BEFORE the clock starts (staging it costs PCIe time that is not
        # part of the workload); outputs and results use a ring of two, like the two pipeline slots.
        hp, ho, hres = [], [], []
        for k in range(total_steps):
            a_ = C.c_void_p()
            ctx.check(lib.sq_host_alloc(ctx.h, eb * CHUNK, C.byref(a_)))
            hp.append(a_)
            b = k % n_batches
            src = corpus[b * B * CHUNK: b * B * CHUNK + eb * CHUNK]
            host_view = torch.frombuffer((C.c_uint8 * (eb * CHUNK)).from_address(a_.value), dtype=torch.uint8)
            host_view.copy_(src)
        torch.cuda.synchronize()
        for _ in range(2):
            b_, r_ = C.c_void_p(), C.c_void_p()
            ctx.check(lib.sq_host_alloc(ctx.h, ocap, C.byref(b_)))
            ctx.check(lib.sq_host_alloc(ctx.h, eb * C.sizeof(L.SqChunkResult), C.byref(r_)))
            ho.append(b_); hres.append(r_)
        hspans = (L.SqSpan * eb)()
        for i in range(eb):
            hspans[i].off, hspans[i].len = i * CHUNK, CHUNK

        tickets = [None, None]
        used = C.c_uint64()
        e2e_out = 0

        def submit(k):
            t = C.c_void_p()
            ctx.check(lib.sq_pack_submit(ctx.h, hp[k], eb * CHUNK, hspans, eb, k * eb, hres[k % 2], ho[k % 2], ocap, C.byref(t)))
            tickets[k % 2] = t

        # Rolling two-slot pipeline, the way the archive packer drives it: wait(k) then submit(k+2), so the upload of the next
        # batch and the frame download of the previous one overlap the kernels of the current one.  The clock starts when the
        # last warm-up step has been delivered and stops when the last step's frames are in host memory.
        if world > 1:
            dist.barrier()
        submit(0)
        submit(1)
        t0 = None
        for k in range(total_steps):
            ctx.check(lib.sq_pack_wait(ctx.h, tickets[k % 2], C.byref(used)))
            if k == n_warm - 1:
                t0 = time.perf_counter()
            if k >= n_warm:
                e2e_out += used.value + eb * C.sizeof(L.SqChunkResult)
            if k + 2 < total_steps:
                submit(k + 2)
        times = [(time.perf_counter() - t0, n_e2e)]
        tsum = sum(t for t, _ in times)
        nsteps = sum(c for _, c in times)
        et = torch.tensor([tsum], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(et, op=dist.ReduceOp.MAX)
        e2e = {"value": world * nsteps * eb * CHUNK / float(et.item()) / 1e9, "unit": "GB/s", "h2d_bytes_per_step": eb * CHUNK + eb * 16,
               "d2h_bytes_per_step": e2e_out // max(nsteps, 1), "chunks_per_step": eb, "steps": nsteps,
               "api": "sq_pack_submit/sq_pack_wait, rolling two-slot pipeline (wait k, submit k+2), pinned host buffers"}
        for h_ in hp + ho + hres:
            lib.sq_host_free(ctx.h, h_)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (the stage with the most time) ----
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    dom = max(stage_ms, key=stage_ms.get)
    local_in, local_out, local_new = totals["in"], totals["out"], totals["new"]
    u_stage = stage_new * CHUNK
    alg = {"digest": stage_in, "dedup": stage_steps * B * 48, "encode": u_stage + stage_out}[dom]
    ach = alg / (stage_ms[dom] * 1e-3) / 1e9 if stage_ms[dom] > 0 else 0.0
    roof = {"bound": "hbm", "kernel": {"digest": "xxh3_128_kernel (K1)", "dedup": "dedup_insert_kernel (K2)", "encode": "zstd encode stages (K3), lz_search_kernel ~78% of it"}[dom],
            "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            # dram bytes/launch of the dominant kernel: 142 B per unique input byte measured by ncu --set full on lz_search_kernel
            # (profiles/r1_enc_final_raw.csv: 205.6 GB read + 58.8 GB written for 888 chunks); K1 reads its input once (ncu: 1.00x)
            "traffic": int(142 * u_stage / stage_steps) if dom == "encode" else (alg // stage_steps if dom == "digest" else None),
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
            "algorithmic_bytes_per_step": alg // stage_steps,
            "timing": ("per-launch CUDA-event durations from a %d-step single-stream pass after the timed region (the timed region itself "
                       "overlaps consecutive steps on two streams)" % stage_steps) if overlapped else "CUDA events over the timed region",
            "stage_ms_per_step": {k: v / stage_steps for k, v in stage_ms.items()},
            "stage_achieved_gbs": {"digest": stage_in / (stage_ms["digest"] * 1e-3) / 1e9 if stage_ms["digest"] else None,
                                   "encode": (u_stage + stage_out) / (stage_ms["encode"] * 1e-3) / 1e9 if stage_ms["encode"] else None}}

    # ---- CPU baseline: the oracle timed on this box's host cores on a bounded sample ----
    cpu = None
    cpu_n = 0
    if not args.no_cpu:
        try:
            oracle = load_oracle()
            cores = os.cpu_count() or 1
            n = args.ref_chunks or max(64, min(2048, cores * 16))
            n = min(n, n_slots, B)
            cpu_n = n
            sample = corpus[: n * CHUNK].cpu().numpy()
            buf = (C.c_uint8 * (n * CHUNK)).from_buffer(sample)
            dt, st = cpu_pack(oracle, buf, n, cores)
            cpu = {"value": n * CHUNK / dt / 1e9, "unit": "GB/s", "cores": cores, "kind": "port",
                   "sample": f"first {n} x 2 MiB slots of this rank's corpus, in memory, {cores} threads, archive to /dev/shm",
                   "ratio": st.payload_bytes / (st.unique_chunks * CHUNK) if st.unique_chunks else None,
                   "libzstd": oracle.L.sqo_zstd_version()}
        except Exception as e:  # the baseline is a report, never a reason to lose the bench line
            log("cpu_baseline failed:", repr(e))

    # ---- unpack (K4) on the frames of one packed batch, device-resident, + same-sample ratio vs the oracle ----
    unpack = None
    try:
        uctx = ctx  # reuse the main context (and its scratch) with a fresh dedup index
        uctx.dedup_reset()
        res = torch.empty(B * 32, dtype=torch.uint8, device="cuda")
        used = C.c_uint64()
        uctx.check(lib.sq_pack_device(uctx.h, corpus.data_ptr(), d_spans.data_ptr(), B, 0, res.data_ptr(), d_out.data_ptr(), out_cap - 64, C.byref(used), sp))
        r = np.frombuffer(res.cpu().numpy().tobytes(), dtype=np.dtype([("d", "u1", 16), ("off", "<u8"), ("len", "<u4"), ("new", "u1"), ("pad", "u1", 3)]))
        sel = np.nonzero(r["new"])[0]
        fr = np.zeros(len(sel), dtype=np.dtype([("src", "<u8"), ("dst", "<u8"), ("len", "<u4"), ("cap", "<u4")]))
        fr["src"], fr["dst"], fr["len"], fr["cap"] = r["off"][sel], np.arange(len(sel)) * CHUNK, r["len"][sel], CHUNK
        d_fr = torch.frombuffer(bytearray(fr.tobytes()), dtype=torch.uint8).cuda()
        dec = torch.empty(len(sel) * CHUNK, dtype=torch.uint8, device="cuda")
        d_res = torch.empty(len(sel) * 8, dtype=torch.uint8, device="cuda")
        t_dec = []
        for _ in range(4):
            e0, e1 = ev(), ev()
            e0.record(stream)
            uctx.check(lib.sq_decode_device(uctx.h, d_out.data_ptr(), d_fr.data_ptr(), len(sel), dec.data_ptr(), d_res.data_ptr(), sp))
            e1.record(stream)
            torch.cuda.synchronize()
            t_dec.append(e0.elapsed_time(e1))
        ok = all(bool(torch.equal(dec[k * CHUNK:(k + 1) * CHUNK], corpus[int(i) * CHUNK:(int(i) + 1) * CHUNK])) for k, i in enumerate(sel[:64]))
        st = np.frombuffer(d_res.cpu().numpy().tobytes(), dtype="<i4").reshape(-1, 2)
        unpack = {"value": len(sel) * CHUNK / (min(t_dec[1:]) * 1e-3) / 1e9, "unit": "GB/s", "frames": int(len(sel)), "what": "K4 decode of one packed batch, device-resident, restored bytes/s",
                  "byte_identical": ok and int((st[:, 1] != 0).sum()) == 0, "algorithmic_gbs": (int(r["len"][sel].sum()) + len(sel) * CHUNK) / (min(t_dec[1:]) * 1e-3) / 1e9}
        # e2e unpack through the host-buffer C-ABI calls (H2D payloads + K4 + D2H restored bytes), one packed batch (<= 2048 frames) per call
        try:
            ne = int(min(len(sel), 2048))
            host_out = d_out[: int(used.value)].cpu().numpy()
            payloads = [host_out[int(r["off"][i]): int(r["off"][i]) + int(r["len"][i])] for i in sel[:ne]]
            hf = (L.SqFrame * ne)()
            so = 0
            for k, pl in enumerate(payloads):
                hf[k].src_off, hf[k].dst_off, hf[k].src_len, hf[k].capacity = so, k * CHUNK, len(pl), CHUNK
                so += (len(pl) + 15) & ~15
            hc = C.c_void_p(); hds = [C.c_void_p(), C.c_void_p()]
            ctx.check(lib.sq_host_alloc(ctx.h, so + 64, C.byref(hc)))
            for hd_ in hds:
                ctx.check(lib.sq_host_alloc(ctx.h, ne * CHUNK, C.byref(hd_)))
            stage = np.frombuffer((C.c_uint8 * (so + 64)).from_address(hc.value), dtype=np.uint8)
            for k, pl in enumerate(payloads):
                stage[hf[k].src_off: hf[k].src_off + len(pl)] = pl
            hress = [(L.SqFrameResult * ne)(), (L.SqFrameResult * ne)()]
            # rolling two-slot pipeline (wait k, submit k+2): every step uploads the payloads and downloads the restored bytes
            n_u, n_uwarm = 8, 2
            utk = [None, None]

            def usubmit(k):
                t = C.c_void_p()
                ctx.check(lib.sq_unpack_submit(ctx.h, hc, so + 64, hf, ne, hds[k % 2], ne * CHUNK, hress[k % 2], C.byref(t)))
                utk[k % 2] = t

            usubmit(0); usubmit(1)
            t0 = None
            for k in range(n_u):
                ctx.check(lib.sq_unpack_wait(ctx.h, utk[k % 2]))
                if k == n_uwarm - 1:
                    t0 = time.perf_counter()
                if k + 2 < n_u:
                    usubmit(k + 2)
            dt_u = time.perf_counter() - t0
            ok_u = all(hress[j][k].status == 0 and hress[j][k].out_len == CHUNK for j in range(2) for k in range(ne))
            back = np.frombuffer((C.c_uint8 * CHUNK).from_address(hds[1].value), dtype=np.uint8)
            ok_u = ok_u and bool((torch.from_numpy(back.copy()).cuda() == corpus[int(sel[0]) * CHUNK:(int(sel[0]) + 1) * CHUNK]).all())
            unpack["e2e"] = {"value": (n_u - n_uwarm) * ne * CHUNK / dt_u / 1e9, "unit": "GB/s", "frames": ne, "steps": n_u - n_uwarm, "h2d_bytes_per_step": so,
                             "d2h_bytes_per_step": ne * CHUNK, "byte_identical": ok_u,
                             "api": "sq_unpack_submit/sq_unpack_wait, rolling two-slot pipeline, pinned host buffers"}
            # CPU baseline for unpack: the reference decodes serially on ONE thread (reader.rs:276-311)
            if not args.no_cpu:
                oracle_u = load_oracle()
                nd = min(ne, 64)
                t0 = time.perf_counter()
                for pl in payloads[:nd]:
                    assert oracle_u.decompress(pl.tobytes(), CHUNK) is not None
                unpack["cpu_baseline"] = {"value": nd * CHUNK / (time.perf_counter() - t0) / 1e9, "unit": "GB/s", "cores": 1, "kind": "port",
                                          "sample": f"{nd} GPU-written frames decoded by stock libzstd on one thread, as the reference's read_chunks does"}
            lib.sq_host_free(ctx.h, hc); lib.sq_host_free(ctx.h, hds[0]); lib.sq_host_free(ctx.h, hds[1])
        except Exception as e:
            log("unpack e2e section failed:", repr(e))
        if cpu and cpu.get("ratio"):
            m = sel[sel < cpu_n]
            gpu_ratio = float(r["len"][m].sum()) / (len(m) * CHUNK) if len(m) else None
            cpu["gpu_ratio_same_sample"] = gpu_ratio
            cpu["ratio_delta_pct"] = (gpu_ratio / cpu["ratio"] - 1) * 100 if gpu_ratio else None
    except Exception as e:
        log("unpack section failed:", repr(e))

    line = {"metric": "pack_gb_per_s", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload],
                       "corpus_gib_per_gpu": n_slots * CHUNK / GiB, "batch_chunks": B, "chunk_bytes": CHUNK,
                       "l2_policy": f"inputs larger than L2: each step reads a fresh {B * CHUNK / GiB:.0f} GiB batch",
                       "streams": 1 if args.single_stream else 2,
                       "parallelism": f"dp{world} (chunks sharded by rank" + (", digest all-to-all over NCCL)" if world > 1 else ")")},
            "gpu_launches": int(launches1.value - launches0.value), "clocks": clk, "e2e": e2e, "roofline": roof, "cpu_baseline": cpu, "unpack": unpack,
            "ratio": {"compressed_over_unique": out_bytes / (n_new * CHUNK) if n_new else None, "unique_fraction": n_new * CHUNK / in_bytes}}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch-chunks", type=int, default=2048)
    ap.add_argument("--corpus-gib", type=int, default=64)
    ap.add_argument("--e2e-chunks", type=int, default=2048)
    ap.add_argument("--ref-chunks", type=int, default=0)
    ap.add_argument("--single-stream", action="store_true", help="run every step on one stream (no overlap between consecutive steps)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    protect_stdout()
    if args.warmup < 3 and args.impl == "ours":
        log("note: fewer than 3 warm-up steps requested; the timing rules ask for >= 3")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
