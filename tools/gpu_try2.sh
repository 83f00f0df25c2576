#!/bin/bash
# dev: parity tests + per-kernel durations (ncu launch list) + one full ncu capture of the search kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -8 gpurun_out/pytest.log
timeout 300 python tools/real_data_ratio.py --gpu-only > gpurun_out/ratio.log 2>&1; tail -3 gpurun_out/ratio.log
for cfg in ${CFGS:-0 2}; do
  SQ_LZ2_CFG=$cfg SQ_TIMING=1 timeout 200 python tools/enc_probe.py ${NCH:-512} cfg$cfg 2>&1 | grep -E "cfg|rror" | tail -3
done | tee gpurun_out/cfgs.log
SQ_LZ2_CFG=${NCFG:-0} timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python tools/enc_probe.py ${NCH:-512} > gpurun_out/ncu1.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/launches.csv")) if len(r) > 10 and r[0].isdigit()]
d = collections.defaultdict(list)
for r in rows: d[r[4].split("(")[0][-40:]].append(float(r[-1].replace(",", "")))
for k, v in d.items(): print(f"{k:42s} n={len(v):3d} last={v[-1]/1e6:9.3f} ms  mean={sum(v)/len(v)/1e6:9.3f} ms")
PY
SQ_LZ2_CFG=${NCFG:-0} timeout 900 ncu --set full --import-source on --clock-control none -k regex:search_kernel -s 1 -c 1 -o gpurun_out/search_full -f python tools/enc_probe.py ${NCH:-512} > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep
