#!/bin/bash
# compute-sanitizer on a small solve + local repair (one tool per gpurun call: $1 = memcheck | racecheck)
TOOL=${1:-memcheck}
python scripts/probe_solve.py --n 512 --reps 1 > gpurun_out/r2_san_plain.log 2>&1 || exit 1
compute-sanitizer --tool $TOOL --print-limit 20 python scripts/probe_solve.py --n 512 --reps 1 > gpurun_out/r2_sanitizer_$TOOL.log 2>&1
tail -15 gpurun_out/r2_sanitizer_$TOOL.log
