#!/bin/bash
# round 2, the last GPU seconds: measured parity + timing of the text encoder (scripts/prof_textenc.py)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 36 python scripts/prof_textenc.py > gpurun_out/r2ar_textenc.log 2>&1
tail -c 3000 gpurun_out/r2ar_textenc.log
