#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log; tail -3 gpurun_out/pytest.log
for n in 64 512; do timeout 300 python tools/dec_bench.py $n 2>&1 | tail -1; done | tee gpurun_out/dec_bench2.log
for n in 512 1536; do timeout 400 python tools/dec_bench.py $n ref 2>&1 | tail -1; done | tee -a gpurun_out/dec_bench2.log
SQ_NO_BLOCK_PARALLEL=1 timeout 400 python tools/dec_bench.py 1536 ref 2>&1 | tail -1 | sed 's/^/[one-pass] /' | tee -a gpurun_out/dec_bench2.log
