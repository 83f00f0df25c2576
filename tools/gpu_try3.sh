#!/bin/bash
# dev: ratio on real data for several search configurations + ncu capture of the chase kernel
mkdir -p gpurun_out
for cfg in ${CFGS:-2 12 16}; do echo "cfg=$cfg $(SQ_LZ2_CFG=$cfg timeout 280 python tools/real_data_ratio.py --gpu-only 2>&1 | tail -1)"; done | tee gpurun_out/ratio_cfgs.log
timeout 900 ncu --set full --import-source on --clock-control none -k regex:chase_kernel -s 1 -c 1 -o gpurun_out/chase_full -f python tools/enc_probe.py 512 > gpurun_out/ncu3.log 2>&1
ls -la gpurun_out/chase_full.ncu-rep
