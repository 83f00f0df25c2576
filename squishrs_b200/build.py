"""Build script: compiles libsquish_b200.so (hand-written CUDA for sm_100a + host C++) in-tree
with nvcc.  No torch extension machinery: the product is a plain C-ABI shared library."""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
LIB = ROOT / "libsquish_b200.so"
CLI = ROOT.parent / "bin" / "squishrs"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
EXTRA = os.environ.get("SQ_NVCC_EXTRA", "").split()
FLAGS = [*EXTRA, "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function", "--expt-relaxed-constexpr"]


def _sources():
    return sorted(CSRC.glob("*.cu")) + sorted((ROOT / "host").glob("*.cpp"))


def _fingerprint() -> str:
    h = hashlib.sha256()
    h.update(os.environ.get("SQ_NVCC_EXTRA", "").encode())
    for p in sorted(list(CSRC.glob("*")) + list((ROOT / "host").glob("*")) + [ROOT.parent / "include" / "squish_b200.h", Path(__file__)]):
        if p.is_file():
            h.update(p.name.encode())
            h.update(p.read_bytes())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    stamp = ROOT / ".build_stamp"
    fp = _fingerprint()
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == fp:
        return LIB
    objdir = ROOT / "build"
    objdir.mkdir(exist_ok=True)
    objs = []
    procs = []
    for src in _sources():
        obj = objdir / (src.stem + ".o")
        cmd = [NVCC, *ARCH, *FLAGS, "-I", str(ROOT.parent / "include"), "-I", str(CSRC), "-c", str(src), "-o", str(obj)]
        if src.suffix == ".cpp":
            cmd.insert(1, "-x")
            cmd.insert(2, "cu")
        if verbose:
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"nvcc failed on {src}:\n{out}\n")
        elif verbose and out.strip():
            print(out)
    if failed:
        raise RuntimeError("libsquish_b200 build failed")
    link = [NVCC, *ARCH, "-shared", "-o", str(LIB), *objs, "-lcudart", "-lpthread"]
    subprocess.run(link, check=True)
    cli_src = ROOT / "host" / "cli_main.cc"
    if cli_src.exists():
        CLI.parent.mkdir(exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-I", str(ROOT.parent / "include"), str(cli_src), "-o", str(CLI),
                        f"-L{ROOT}", "-lsquish_b200", f"-Wl,-rpath,{ROOT}", "-Wl,-rpath,/usr/local/cuda/lib64"], check=True)
    stamp.write_text(fp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
