#!/bin/bash
# dev: first GPU pass of a changed encoder: parity tests, real-data ratio, speed of each search-kernel configuration
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -15 gpurun_out/pytest.log
timeout 300 python tools/real_data_ratio.py --gpu-only > gpurun_out/ratio.log 2>&1; tail -3 gpurun_out/ratio.log
for cfg in ${CFGS:-0 1 2 3 4 5 6}; do
  SQ_LZ2_CFG=$cfg SQ_TIMING=1 timeout 200 python tools/enc_probe.py ${NCH:-512} cfg$cfg 2>&1 | grep -E "cfg|search kernel|rror" | tail -3
done | tee gpurun_out/cfgs.log
