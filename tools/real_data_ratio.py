"""Dev tool (GPU box): ratio of the GPU encoder vs libzstd level 12 on REAL files of this image (Python sources, shared
libraries, locale data), packed as 2 MiB chunks of a concatenated stream and as individual small files."""
import ctypes as C, glob, json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import squishrs_b200 as sq
from conftest import Oracle
O = Oracle(); ctx = sq.Context()
GPU_ONLY = "--gpu-only" in sys.argv  # libzstd sizes are then taken from tools/real_data_libzstd.json (same image, same files)
REF = json.loads((ROOT / "tools" / "real_data_libzstd.json").read_text()) if GPU_ONLY else {}
CH = 2 << 20


def blob(pattern, limit):
    out = bytearray()
    for f in sorted(glob.glob(pattern, recursive=True)):
        try: out += open(f, "rb").read()
        except OSError: continue
        if len(out) >= limit: break
    return bytes(out[:limit])


sets = {"python sources": blob("/usr/lib/python3*/**/*.py", 32 << 20), "shared libraries": blob("/usr/lib/x86_64-linux-gnu/*.so*", 32 << 20),
        "site-packages text": blob("/opt/prime-rl/.venv/lib/python3.12/site-packages/**/*.py", 32 << 20)}
res = {}
for name, data in sets.items():
    chunks = [data[i:i + CH] for i in range(0, len(data), CH)]
    frames = ctx.pack_batch(chunks)
    gpu = sum(len(f) for _, f in frames if f is not None)
    uniq = [c for c, (_, f) in zip(chunks, frames) if f is not None]
    for c, (_, f) in zip(chunks, frames):
        if f is not None and not GPU_ONLY: assert O.decompress(f, len(c)) == c
    cpu = REF[name]["libzstd_l12"] if GPU_ONLY else sum(len(O.compress(c, 12)) for c in uniq)
    res[name] = {"bytes": sum(len(c) for c in uniq), "gpu": gpu, "libzstd_l12": cpu, "delta_pct": (gpu / cpu - 1) * 100}
    ctx.dedup_reset()
small = [data[i:i + 24000] for data in sets.values() for i in range(0, 4 << 20, 24000)]
frames = ctx.pack_batch(small)
pairs = [(c, f) for c, (_, f) in zip(small, frames) if f is not None]
gpu = sum(len(f) for _, f in pairs); cpu = REF["24 KB pieces of all three"]["libzstd_l12"] if GPU_ONLY else sum(len(O.compress(c, 12)) for c, _ in pairs)
res["24 KB pieces of all three"] = {"bytes": sum(len(c) for c, _ in pairs), "gpu": gpu, "libzstd_l12": cpu, "delta_pct": (gpu / cpu - 1) * 100}
print(json.dumps(res, indent=1) if not GPU_ONLY else " ".join(f"{k.split()[0]} {v['delta_pct']:+.2f}%" for k, v in res.items()))
