"""BASELINE.json configs[0] and configs[4] through the real command lines, on the GPU box:
   configs[0]: 1 GiB text-like tree, 2 000 files, ~30 % whole-file copies -> `squishrs pack` + `unpack` (GPU) vs `refcpu pack` + `unpack`
   configs[4]: many small files (4-64 KiB), archive written by the reference path -> `squishrs unpack` (GPU) vs `refcpu unpack`
Prints one JSON object with timings and the byte-identity checks."""
import ctypes as C, filecmp, json, math, os, random, shutil, subprocess, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import squishrs_b200 as sq
lib = sq.load()
CLI, REF = str(ROOT / "bin" / "squishrs"), str(ROOT / "oracle" / "refcpu")
DEV = ["--devices", os.environ["SQ_DEVICES"]] if os.environ.get("SQ_DEVICES") else []  # several GPUs of the box through one process
BASE = Path(os.environ.get("SQ_TMP", "/dev/shm")) / f"sq_cfg_{os.getpid()}"


def gen(n, pid, klass):
    b = C.create_string_buffer(max(n, 1)); lib.sq_corpus_fill_host(b, n, 0x51510001, pid, klass); return b.raw[:n]


def same_tree(a, b):
    c = filecmp.dircmp(a, b)
    def ok(d):
        if d.left_only or d.right_only or d.funny_files: return False
        _, mism, err = filecmp.cmpfiles(d.left, d.right, d.common_files, shallow=False)
        return not mism and not err and all(ok(s) for s in d.subdirs.values())
    return ok(c)


def run(cmd):
    t = time.perf_counter(); r = subprocess.run(cmd, capture_output=True, text=True); dt = time.perf_counter() - t
    if r.returncode: raise SystemExit(f"{cmd} failed: {r.stderr[-400:]}")
    return dt


def phases(cmd):
    """the same command once more with SQ_TIMING=1: the library's own phase marks (start-up breakdown)"""
    r = subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ, SQ_TIMING="1"))
    return [l for l in r.stderr.splitlines() if l.startswith("[sq")]


def config1(total=1 << 30, nfiles=2000):
    rng = random.Random(0x51510001); src = BASE / "c1" / "tree"; sizes = [min(16 << 20, max(1024, rng.lognormvariate(math.log(256 << 10), 1.0))) for _ in range(nfiles)]
    k = total / sum(sizes); sizes = [max(1024, int(s * k)) for s in sizes]; originals = []; nbytes = 0
    for i, n in enumerate(sizes):
        if originals and rng.random() < 0.3: data = originals[rng.randrange(len(originals))]
        else: data = gen(n, i, 0); originals.append(data)
        p = src / f"d{i % 20}" / f"s{i % 7}" / f"f{i}.txt"; p.parent.mkdir(parents=True, exist_ok=True); p.write_bytes(data); nbytes += len(data)
    out = {"bytes": nbytes, "files": nfiles}
    threads = str(os.cpu_count())
    out["gpu_pack_s"] = run([CLI, *DEV, "pack", str(src), "-o", str(BASE / "c1" / "gpu.squish")])
    out["gpu_unpack_s"] = run([CLI, *DEV, "unpack", str(BASE / "c1" / "gpu.squish"), "-o", str(BASE / "c1" / "gpu_out")])
    out["cpu_pack_s"] = run([REF, "-j", threads, "pack", str(src), "-o", str(BASE / "c1" / "cpu.squish")])
    out["cpu_unpack_s"] = run([REF, "-j", threads, "unpack", str(BASE / "c1" / "cpu.squish"), "-o", str(BASE / "c1" / "cpu_out")])
    out["cross_gpu_archive_cpu_unpack_s"] = run([REF, "-j", threads, "unpack", str(BASE / "c1" / "gpu.squish"), "-o", str(BASE / "c1" / "x1")])
    out["cross_cpu_archive_gpu_unpack_s"] = run([CLI, *DEV, "unpack", str(BASE / "c1" / "cpu.squish"), "-o", str(BASE / "c1" / "x2")])
    out["identical"] = all(same_tree(src, BASE / "c1" / d) for d in ("gpu_out", "cpu_out", "x1", "x2"))
    for d in ("gpu_out", "cpu_out", "x1", "x2"): shutil.rmtree(BASE / "c1" / d, ignore_errors=True)  # bounded /dev/shm use
    out["gpu_pack_phases"] = phases([CLI, *DEV, "pack", str(src), "-o", str(BASE / "c1" / "gpu2.squish")])
    out["gpu_unpack_phases"] = phases([CLI, *DEV, "unpack", str(BASE / "c1" / "gpu.squish"), "-o", str(BASE / "c1" / "gpu_out2")])
    shutil.rmtree(BASE / "c1" / "gpu_out2", ignore_errors=True)
    out["gpu_pack_second_run_s"] = run([CLI, *DEV, "pack", str(src), "-o", str(BASE / "c1" / "gpu3.squish")])
    out["devices"] = int(os.environ.get("SQ_DEVICES", "1"))
    out["gpu_archive_bytes"] = (BASE / "c1" / "gpu.squish").stat().st_size; out["cpu_archive_bytes"] = (BASE / "c1" / "cpu.squish").stat().st_size
    out["ratio_delta_pct"] = (out["gpu_archive_bytes"] / out["cpu_archive_bytes"] - 1) * 100
    for k2 in ("gpu_pack", "gpu_unpack", "cpu_pack", "cpu_unpack"): out[k2 + "_gbs"] = nbytes / out[k2 + "_s"] / 1e9
    return out


def config5(nfiles=20000):
    rng = random.Random(0x51510005); src = BASE / "c5" / "tree"; nbytes = 0
    for i in range(nfiles):
        n = rng.randrange(4096, 65537); p = src / f"d{i % 100}" / f"f{i}.dat"; p.parent.mkdir(parents=True, exist_ok=True)
        p.write_bytes(gen(n, i, 0 if i % 2 else 2)); nbytes += n
    threads = str(os.cpu_count()); out = {"bytes": nbytes, "files": nfiles}
    out["ref_pack_s"] = run([REF, "-j", threads, "pack", str(src), "-o", str(BASE / "c5" / "ref.squish")])
    out["gpu_unpack_s"] = run([CLI, *DEV, "unpack", str(BASE / "c5" / "ref.squish"), "-o", str(BASE / "c5" / "gpu_out")])
    out["cpu_unpack_s"] = run([REF, "-j", threads, "unpack", str(BASE / "c5" / "ref.squish"), "-o", str(BASE / "c5" / "cpu_out")])
    out["identical"] = same_tree(src, BASE / "c5" / "gpu_out") and same_tree(src, BASE / "c5" / "cpu_out")
    out["gpu_unpack_phases"] = phases([CLI, "unpack", str(BASE / "c5" / "ref.squish"), "-o", str(BASE / "c5" / "gpu_out2")])
    out["gpu_unpack_second_run_s"] = run([CLI, "unpack", str(BASE / "c5" / "ref.squish"), "-o", str(BASE / "c5" / "gpu_out3")])
    out["cpu_unpack_parallel_decode_s"] = run([REF, "-j", threads, "--parallel-decode", "unpack", str(BASE / "c5" / "ref.squish"), "-o", str(BASE / "c5" / "cpu_out2")]) if os.environ.get("SQ_REF_PD") else None
    out["gpu_unpack_gbs"] = nbytes / out["gpu_unpack_s"] / 1e9; out["cpu_unpack_gbs"] = nbytes / out["cpu_unpack_s"] / 1e9
    if os.environ.get("SQ_C5_GPU_PACK"):  # not a BASELINE config: the same tree of small files packed by the GPU path
        for d in ("gpu_out", "cpu_out", "gpu_out2", "gpu_out3", "cpu_out2"): shutil.rmtree(BASE / "c5" / d, ignore_errors=True)
        out["gpu_pack_s"] = run([CLI, *DEV, "pack", str(src), "-o", str(BASE / "c5" / "gpu.squish")])
        out["gpu_pack_second_run_s"] = run([CLI, *DEV, "pack", str(src), "-o", str(BASE / "c5" / "gpu2.squish")])
        out["gpu_pack_phases"] = phases([CLI, *DEV, "pack", str(src), "-o", str(BASE / "c5" / "gpu3.squish")])
        run([REF, "-j", threads, "unpack", str(BASE / "c5" / "gpu.squish"), "-o", str(BASE / "c5" / "x1")])
        out["gpu_pack_identical_after_ref_unpack"] = same_tree(src, BASE / "c5" / "x1")
        out["gpu_archive_bytes"] = (BASE / "c5" / "gpu.squish").stat().st_size; out["ref_archive_bytes"] = (BASE / "c5" / "ref.squish").stat().st_size
    return out


if __name__ == "__main__":
    try:
        res = {"host_cores": os.cpu_count(), "config1": None if os.environ.get("SQ_SKIP_C1") else config1(int(float(os.environ.get("SQ_C1_GIB", "1")) * (1 << 30))), "config5": None if os.environ.get("SQ_SKIP_C5") else config5(int(os.environ.get("SQ_C5_FILES", "20000")))}
        print(json.dumps(res))
    finally:
        shutil.rmtree(BASE, ignore_errors=True)
