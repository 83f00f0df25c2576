#!/bin/bash
# one GPU: parity tests, the bench line, the reference arm, and the ncu launch list of the same bench command
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log; tail -3 gpurun_out/pytest.log
timeout 1200 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.log; echo "bench rc=$?"; tail -3 gpurun_out/bench_n1.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref.log; echo "ref rc=$?"; head -c 600 gpurun_out/bench_ref_n1.json; echo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/bench_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/bench_launches.csv")) if len(r) > 10 and r[0].isdigit()]
d = collections.defaultdict(list)
for r in rows: d[r[4].split("(")[0][-44:]].append(float(r[-1].replace(",", "")))
tot = sum(sum(v) for v in d.values())
lines = [f"{k:46s} n={len(v):4d} mean={sum(v)/len(v)/1e6:9.3f} ms  sum={sum(v)/1e6:10.3f} ms  share={sum(v)/tot*100:5.1f} %" for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1]))]
open("gpurun_out/bench_launch_shares.txt", "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:12]))
PY
python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print(d['value'], d['e2e']['value'], d['roofline']['kernel_ms_per_step'], d['unpack']['value'], d['unpack']['e2e']['value'], d['cpu_baseline'].get('ratio_delta_pct'), d['cpu_baseline'].get('ratio_delta_pct_real'))"
