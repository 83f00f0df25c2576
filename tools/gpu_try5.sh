#!/bin/bash
# dev: parity tests, real-data ratio, encoder probe with per-kernel durations, full ncu captures of the index and search kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -8 gpurun_out/pytest.log
timeout 300 python tools/real_data_ratio.py --gpu-only > gpurun_out/ratio.log 2>&1; tail -3 gpurun_out/ratio.log
SQ_TIMING=1 timeout 200 python tools/enc_probe.py ${NCH:-2048} probe 2>&1 | grep -E "probe|rror" | tail -3 | tee gpurun_out/cfgs.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python tools/enc_probe.py ${NCH:-2048} > gpurun_out/ncu1.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/launches.csv")) if len(r) > 10 and r[0].isdigit()]
d = collections.defaultdict(list)
for r in rows: d[r[4].split("(")[0][-40:]].append(float(r[-1].replace(",", "")))
for k, v in d.items(): print(f"{k:42s} n={len(v):3d} last={v[-1]/1e6:9.3f} ms  mean={sum(v)/len(v)/1e6:9.3f} ms  sum={sum(v)/1e6:9.3f}")
PY
if [ -n "$NCU" ]; then
timeout 900 ncu --set full --import-source on --clock-control none -k regex:search_kernel -s 1 -c 1 -o gpurun_out/search_full -f python tools/enc_probe.py 512 > gpurun_out/ncu2.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:index_kernel -s 1 -c 1 -o gpurun_out/index_full -f python tools/enc_probe.py 512 > gpurun_out/ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep
fi
