#!/bin/bash
# ncu --set full of the block-parallel decoder's two passes on 256 K3-written frames
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:bp_matches_kernel -s 1 -c 1 -o gpurun_out/bp_matches_full -f python tools/dec_bench.py 256 > gpurun_out/ncu4.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:bp_first_pass_kernel -s 1 -c 1 -o gpurun_out/bp_first_pass_full -f python tools/dec_bench.py 256 > gpurun_out/ncu5.log 2>&1
ls -la gpurun_out/bp_*.ncu-rep
