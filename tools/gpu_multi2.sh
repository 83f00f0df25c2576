#!/bin/bash
# two GPUs of one box: the multi-context archive test, then configs[0] at 4 GiB through the CLI with one and with two devices
mkdir -p gpurun_out
nvidia-smi -L; free -g | head -2; df -h /dev/shm | tail -1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "several_contexts or cli_roundtrip or roundtrip_both" > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_multi.log
SQ_SKIP_C5=1 SQ_C1_GIB=4 timeout 900 python tools/config_cli.py > gpurun_out/r2_cli_4gib_dev1.json 2> gpurun_out/cli1.log; echo "cli dev1 rc=$?"; tail -3 gpurun_out/cli1.log
SQ_DEVICES=2 SQ_SKIP_C5=1 SQ_C1_GIB=4 timeout 900 python tools/config_cli.py > gpurun_out/r2_cli_4gib_dev2.json 2> gpurun_out/cli2.log; echo "cli dev2 rc=$?"; tail -3 gpurun_out/cli2.log
python - <<'PY'
import json
for f in ("r2_cli_4gib_dev1", "r2_cli_4gib_dev2"):
    try:
        c = json.load(open(f"gpurun_out/{f}.json"))["config1"]
        print(f, {k: (round(v, 2) if isinstance(v, float) else v) for k, v in c.items() if not k.endswith("phases")})
    except Exception as e: print(f, "unreadable", e)
PY
