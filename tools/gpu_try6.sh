#!/bin/bash
# dev: A/B of the parse kernels (warp per block vs thread per 8 KiB piece): probe, real-data ratio, launch list
mkdir -p gpurun_out
for v in 0 1; do SQ_CHASE_SUB=$v timeout 200 python tools/enc_probe.py ${NCH:-2048} chase_sub$v 2>&1 | grep -E "chase_sub|rror" | tail -2; done | tee gpurun_out/cfgs.log
SQ_CHASE_SUB=1 timeout 300 python tools/real_data_ratio.py --gpu-only 2>&1 | tail -2 | tee gpurun_out/ratio.log
SQ_CHASE_SUB=1 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log; tail -3 gpurun_out/pytest.log
SQ_CHASE_SUB=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python tools/enc_probe.py ${NCH:-2048} > gpurun_out/ncu1.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/launches.csv")) if len(r) > 10 and r[0].isdigit()]
d = collections.defaultdict(list)
for r in rows: d[r[4].split("(")[0][-40:]].append(float(r[-1].replace(",", "")))
for k, v in d.items(): print(f"{k:42s} n={len(v):3d} last={v[-1]/1e6:9.3f} ms  mean={sum(v)/len(v)/1e6:9.3f} ms  sum={sum(v)/1e6:9.3f}")
PY
