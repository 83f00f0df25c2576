#!/bin/bash
# dev: quick A/B of an encoder change: real-data ratio, probe, launch list
mkdir -p gpurun_out
timeout 300 python tools/real_data_ratio.py --gpu-only 2>&1 | tail -1 | tee gpurun_out/ratio.log
timeout 200 python tools/enc_probe.py 2048 probe 2>&1 | grep -E "probe|rror" | tail -2 | tee gpurun_out/cfgs.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python tools/enc_probe.py 2048 > gpurun_out/ncu1.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/launches.csv")) if len(r) > 10 and r[0].isdigit()]
d = collections.defaultdict(list)
for r in rows: d[r[4].split("(")[0][-40:]].append(float(r[-1].replace(",", "")))
for k, v in d.items():
    if sum(v) > 1e6: print(f"{k:42s} n={len(v):3d} mean={sum(v)/len(v)/1e6:9.3f} ms  sum={sum(v)/1e6:9.3f}")
PY
python -c "import __graft_entry__ as g; g.smoke()"
