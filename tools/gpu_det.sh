#!/bin/bash
mkdir -p gpurun_out
python - <<'PY'
import sys, ctypes as C
sys.path.insert(0, "/root/repo")
import squishrs_b200 as sq
lib = sq.load()
chunks = []
for klass, n in ((0, 2 << 20), (1, 2 << 20), (2, 2 << 20), (3, 2 << 20), (4, 2 << 20), (0, 70000), (2, 300000), (1, 131072), (0, 5000)):
    b = C.create_string_buffer(n); lib.sq_corpus_fill_host(b, n, 5, klass + 11, klass); chunks.append(b.raw)
for det in (True, False):
    outs = []
    for rep in range(3):
        c = sq.Context(deterministic=det)
        outs.append([f for _, f in c.pack_batch(chunks)])
        same_ctx = [f for _, f in (c.dedup_reset(), c.pack_batch(chunks))[1]]
        outs.append(same_ctx)
    for i in range(len(chunks)):
        fs = [o[i] for o in outs]
        d = [next((k for k in range(min(len(a), len(fs[0]))) if a[k] != fs[0][k]), -1 if len(a) == len(fs[0]) else -2) for a in fs]
        print("det" if det else "def", "chunk", i, "len", len(chunks[i]), "frame sizes", [len(a) for a in fs], "first diff vs run0", d)
PY
