#!/bin/bash
out=gpurun_out/exp_fim16.log
: > $out
for R in 0 1 2 4; do
    echo "=== refresh=$R" >> $out
    DYMU_FIM_REFRESH=$R timeout 120 python scripts/probe_solve.py --n 4096 --reps 3 --check 2>&1 | grep "rep 2\|max rel" >> $out
done
