#!/bin/bash
# dev: multi-GPU bench lines.  usage: tools/gpu_multi.sh N workload steps [extra args]
N=$1; W=$2; S=$3; shift 3
mkdir -p gpurun_out
if [ "$N" = 1 ]; then
  timeout 900 python bench.py --gpus 1 --workload $W --steps $S --warmup 3 "$@" > gpurun_out/bench_${W}_n$N.json 2> gpurun_out/bench_${W}_n$N.log
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $W --steps $S --warmup 3 "$@" > gpurun_out/bench_${W}_n$N.json 2> gpurun_out/bench_${W}_n$N.log
fi
echo rc=$?; grep -v "^\[W\|NCCL\|^$" gpurun_out/bench_${W}_n$N.log | tail -4; cut -c1-1500 gpurun_out/bench_${W}_n$N.json
