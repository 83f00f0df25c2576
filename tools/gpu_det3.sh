#!/bin/bash
SQ_DETERMINISTIC=1 timeout 200 python tools/enc_probe.py 2048 deterministic 2>&1 | grep -E "deterministic|rror" | tail -1
timeout 300 python -m pytest tests -m gpu -x -q -k "deterministic" 2>&1 | tail -2
