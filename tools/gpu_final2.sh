#!/bin/bash
# final set: decoder table, then tests + bench line + reference arm + launch list
mkdir -p gpurun_out
for n in 64 512; do timeout 300 python tools/dec_bench.py $n 2>&1 | tail -1; done | tee gpurun_out/dec_bench3.log
for n in 128 512 1536; do timeout 400 python tools/dec_bench.py $n ref 2>&1 | tail -1; done | tee -a gpurun_out/dec_bench3.log
bash tools/gpu_final1.sh
