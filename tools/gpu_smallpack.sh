#!/bin/bash
# dev: pack of 20 000 small files through the command line on the GPU (scratch is sized per chunk slot: what does that cost?)
mkdir -p gpurun_out
SQ_SKIP_C1=1 SQ_C5_FILES=20000 SQ_C5_GPU_PACK=1 timeout 900 python tools/config_cli.py > gpurun_out/r2_cli_smallpack.json 2> gpurun_out/smallpack.log; echo "rc=$?"; tail -3 gpurun_out/smallpack.log
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_cli_smallpack.json"))["config5"]
print({a: (round(b, 2) if isinstance(b, float) else b) for a, b in d.items() if not a.endswith("phases")}); print(d.get("gpu_pack_phases"))
PY
nvidia-smi --query-gpu=memory.used --format=csv
