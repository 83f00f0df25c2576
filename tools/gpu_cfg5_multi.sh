#!/bin/bash
# N GPUs of one box: configs[4] (unpack-only, 200 000 small frames) as bench line (one process per GPU) and through the CLI (--devices N)
N=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload config5 --steps 4 --warmup 3 > gpurun_out/r2_bench_config5_n$N.json 2> gpurun_out/cfg5_n$N.log; echo "bench rc=$?"; grep -v "^\[W\|NCCL\|^$\|OMP_NUM\|\*\*\*" gpurun_out/cfg5_n$N.log | tail -3; cut -c1-700 gpurun_out/r2_bench_config5_n$N.json; echo
if [ -n "$CLI" ]; then
SQ_SKIP_C1=1 SQ_C5_FILES=200000 SQ_REF_PD=1 timeout 1200 python tools/config_cli.py > gpurun_out/r2_cli_config5_dev1.json 2> gpurun_out/cli5_1.log; echo "cli dev1 rc=$?"
SQ_DEVICES=$N SQ_SKIP_C1=1 SQ_C5_FILES=200000 timeout 1200 python tools/config_cli.py > gpurun_out/r2_cli_config5_dev$N.json 2> gpurun_out/cli5_n.log; echo "cli dev$N rc=$?"
python - <<PY
import json
for f in ("r2_cli_config5_dev1", "r2_cli_config5_dev$N"):
    try:
        c = json.load(open(f"gpurun_out/{f}.json"))["config5"]
        print(f, {k: (round(v, 2) if isinstance(v, float) else v) for k, v in c.items() if not k.endswith("phases")}); print(c.get("gpu_unpack_phases"))
    except Exception as e: print(f, "unreadable", e)
PY
fi
