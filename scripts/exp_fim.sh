#!/bin/bash
out=gpurun_out/exp_fim18.log
: > $out
for B in 6 8 12; do
    echo "=== tile64 band=$B" >> $out
    DYMU_FIM_TILE=64 DYMU_FIM_BAND=$B DYMU_FIM_INNER=128 timeout 120 python scripts/probe_solve.py --n 4096 --reps 3 --nopath 2>&1 | grep "rep 2" >> $out
done
