#!/bin/bash
# Dev tool: phase timing of the squishrs CLI on a 1 GiB text tree in /dev/shm (SQ_TIMING=1 prints host phases).
set -e
cd "$(dirname "$0")/.."
T=/dev/shm/sq_cli_timing_$$
mkdir -p $T
python - <<PY
import ctypes as C, sys, random, math
sys.path.insert(0, ".")
import squishrs_b200 as sq
from pathlib import Path
lib = sq.load(); rng = random.Random(1); root = Path("$T/tree")
for i in range(2000):
    n = int(min(16 << 20, max(1024, rng.lognormvariate(math.log(256 << 10), 1.0)) * 1.3))
    b = C.create_string_buffer(n); lib.sq_corpus_fill_host(b, n, 7, i, 0)
    p = root / f"d{i % 20}" / f"f{i}.txt"; p.parent.mkdir(parents=True, exist_ok=True); p.write_bytes(b.raw[:n])
PY
du -sh $T/tree
echo "== pack"; time env SQ_TIMING=1 bin/squishrs pack $T/tree -o $T/a.squish
echo "== unpack"; time env SQ_TIMING=1 bin/squishrs unpack $T/a.squish -o $T/out
rm -rf $T
