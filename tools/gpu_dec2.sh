#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log; tail -3 gpurun_out/pytest.log
for n in 64 512; do timeout 300 python tools/dec_bench.py $n 2>&1 | tail -1; done | tee gpurun_out/dec_bench2.log
timeout 400 python tools/dec_bench.py 512 ref 2>&1 | tail -1 | tee -a gpurun_out/dec_bench2.log
