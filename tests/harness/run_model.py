"""Dev tool: run the CPU encoder model over sample corpora, verify with stock libzstd, print ratio vs L12."""
import ctypes as C, sys, glob, os, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from conftest import Oracle
import squishrs_b200 as sq

class Params(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("hash_log", "row_entries", "min_match", "lazy_depth", "rep_mode", "tile", "target_len", "alt_window", "sel_mul", "accept_thr", "skip_stride", "skip_min", "precheck")]

M = C.CDLL(str(ROOT / "tests/harness/libenc_model.so"))
M.enc_model_frame.restype = C.c_long
M.enc_model_frame.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32, C.POINTER(Params), C.POINTER(C.c_uint32)]
O = Oracle(); lib = sq.load()

def samples(size=2 << 20):
    out = {}
    names = {0: "text", 1: "log", 2: "json", 3: "binary", 6: "fsmix"}
    for k, nm in names.items():
        b = C.create_string_buffer(size); lib.sq_corpus_fill_host(b, size, 0x51510002, 1000 + k, k); out[nm] = b.raw
    src = b"".join(open(f, "rb").read() for f in sorted(glob.glob("/usr/lib/python3*/**/*.py", recursive=True))[:400])
    out["pysrc"] = src[:size]
    out["pysrc256k"] = src[size:size + (256 << 10)]
    out["text64k"] = out["text"][:65536]
    out["json16k"] = out["json"][:16384]
    return out

def run(P, data):
    cap = len(data) + 4096
    dst = C.create_string_buffer(cap); st = (C.c_uint32 * 8)()
    t = time.time(); n = M.enc_model_frame(data, len(data), dst, cap, C.byref(P), st); dt = time.time() - t
    assert n > 0
    back = O.decompress(dst.raw[:n], len(data))
    assert back == data, "stock libzstd could not decode the model's frame"
    return n, st[0], st[1], st[2], dt, (st[5], st[6] * 16 / max(len(data), 1)), st[7] * 16 / max(len(data), 1)

if __name__ == "__main__":
    S = samples()
    ref = {k: len(O.compress(v, 12)) for k, v in S.items()}
    configs = {
        "base hl16 k16 mm5 lazy2 rep0": Params(16, 16, 5, 2, 0, 1024, 64, 4, 0),
        "rep1 approx": Params(16, 16, 5, 2, 1, 1024, 64, 4, 0),
        "rep2 exact": Params(16, 16, 5, 2, 2, 1024, 64, 4, 0),
        "rep1 mm4": Params(16, 16, 4, 2, 1, 1024, 64, 4, 0),
    }
    if len(sys.argv) > 1:
        configs = {a: Params(*[int(x) for x in a.split(",")]) for a in sys.argv[1:]}
    print(f"{'config':34s}" + "".join(f"{k:>11s}" for k in S) + "    total")
    print(f"{'libzstd L12 ratio':34s}" + "".join(f"{ref[k] / len(S[k]):11.4f}" for k in S))
    for name, P in configs.items():
        row = []; tot = 0; tref = 0; ver = []; info = []
        for k, v in S.items():
            n, ns, nl, nr, dt, steps, vpb = run(P, v)
            row.append(n / ref[k]); tot += n; tref += ref[k]; ver.append(vpb); info.append(f"{k}: seq {ns} lit {nl} steps {steps}")
        print(f"{name:34s}" + "".join(f"{(x - 1) * 100:+10.1f}%" for x in row) + f"  {(tot / tref - 1) * 100:+6.1f}%" + "  verif/byte " + " ".join(f"{x:.1f}" for x in ver))
        if os.environ.get("ENC_WINDOW"): print("    " + " | ".join(info))
