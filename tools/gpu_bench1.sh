#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=memory.total,memory.used --format=csv
timeout 1200 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.log; echo "bench rc=$?"; tail -5 gpurun_out/bench_n1.log; cat gpurun_out/bench_n1.json | head -c 3000
