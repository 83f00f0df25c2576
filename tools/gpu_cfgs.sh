#!/bin/bash
# one GPU: configs[3] (random, raw-block fallback) and configs[4] (unpack-only, 200 000 small frames) bench lines
mkdir -p gpurun_out
timeout 900 python bench.py --workload config4 --corpus-gib 32 --steps 6 --warmup 3 > gpurun_out/r2_bench_config4_n1.json 2> gpurun_out/cfg4.log; echo "config4 rc=$?"; tail -2 gpurun_out/cfg4.log; head -c 900 gpurun_out/r2_bench_config4_n1.json; echo
timeout 900 python bench.py --workload config5 --steps 4 --warmup 3 > gpurun_out/r2_bench_config5_n1.json 2> gpurun_out/cfg5.log; echo "config5 rc=$?"; tail -2 gpurun_out/cfg5.log; head -c 900 gpurun_out/r2_bench_config5_n1.json; echo
