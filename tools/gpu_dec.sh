#!/bin/bash
# dev: block-parallel decode -- parity tests, then decode throughput of K3-written and libzstd-written frames with and without it
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log; tail -3 gpurun_out/pytest.log
for n in 64 256 512; do
  timeout 300 python tools/dec_bench.py $n 2>&1 | tail -1
  SQ_NO_BLOCK_PARALLEL=1 timeout 300 python tools/dec_bench.py $n 2>&1 | tail -1 | sed 's/^/[one-pass] /'
done | tee gpurun_out/dec_bench.log
for n in 128 512 1536; do
  timeout 400 python tools/dec_bench.py $n ref 2>&1 | tail -1
  SQ_NO_BLOCK_PARALLEL=1 timeout 400 python tools/dec_bench.py $n ref 2>&1 | tail -1 | sed 's/^/[one-pass] /'
done | tee -a gpurun_out/dec_bench.log
