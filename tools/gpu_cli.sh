#!/bin/bash
# one GPU: the new tests, then configs[0] (1 GiB) and configs[4] (200 000 files) through the command lines
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "release_scratch or roundtrip_both or several_contexts or cli_roundtrip" > gpurun_out/pytest_cli.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_cli.log
SQ_C5_FILES=200000 SQ_REF_PD=1 timeout 1500 python tools/config_cli.py > gpurun_out/r2_config1_config5_cli.json 2> gpurun_out/cli.log; echo "cli rc=$?"; tail -2 gpurun_out/cli.log
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_config1_config5_cli.json"))
for k in ("config1", "config5"):
    print(k, {a: (round(b, 2) if isinstance(b, float) else b) for a, b in d[k].items() if not a.endswith("phases")})
    for a in d[k]:
        if a.endswith("phases"): print(" ", a, d[k][a])
PY
