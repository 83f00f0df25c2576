#!/bin/bash
# Dev tool: phase timing of `squishrs unpack` on an archive of many small files written by the reference-style CPU path.
set -e
cd "$(dirname "$0")/.."
T=/dev/shm/sq_cli_small_$$
mkdir -p $T
python - <<PY
import ctypes as C, sys, random
sys.path.insert(0, ".")
import squishrs_b200 as sq
from pathlib import Path
lib = sq.load(); rng = random.Random(0x51510005); root = Path("$T/tree")
for i in range(20000):
    n = rng.randrange(4096, 65537)
    b = C.create_string_buffer(n); lib.sq_corpus_fill_host(b, n, 0x51510001, i, 0 if i % 2 else 2)
    p = root / f"d{i % 100}" / f"f{i}.dat"; p.parent.mkdir(parents=True, exist_ok=True); p.write_bytes(b.raw[:n])
PY
oracle/refcpu -j 16 pack $T/tree -o $T/ref.squish
for k in 1 2; do echo "== unpack run $k"; ( time env SQ_TIMING=1 bin/squishrs unpack $T/ref.squish -o $T/out$k ) 2>&1; done
rm -rf $T
