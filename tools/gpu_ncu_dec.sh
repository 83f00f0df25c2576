#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:bp_execute_kernel -s 1 -c 1 -o gpurun_out/bp_exec_full -f python tools/dec_bench.py 256 > gpurun_out/ncu4.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:bp_entropy_kernel -s 1 -c 1 -o gpurun_out/bp_entropy_full -f python tools/dec_bench.py 256 > gpurun_out/ncu5.log 2>&1
ls -la gpurun_out/*.ncu-rep
