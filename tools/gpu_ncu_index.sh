#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:index_kernel -s 1 -c 1 -o gpurun_out/index_full -f python tools/enc_probe.py 512 > gpurun_out/ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep
