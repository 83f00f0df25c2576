#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log; tail -4 gpurun_out/pytest.log
bash tools/gpu_det.sh 2>&1 | grep -E "^det|^def" | cut -c1-200
timeout 200 python tools/real_data_ratio.py --gpu-only 2>&1 | tail -1
timeout 200 python tools/enc_probe.py 2048 probe 2>&1 | grep -E "probe|rror" | tail -1
