#!/bin/bash
out=gpurun_out/exp_fim14.log
: > $out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 >> $out
for B in 3 4 5 6; do
    echo "=== band=$B" >> $out
    DYMU_FIM_BAND=$B timeout 120 python scripts/probe_solve.py --n 4096 --reps 3 --nopath 2>&1 | grep "rep 2" >> $out
done
echo "=== 16384" >> $out
timeout 300 python scripts/probe_solve.py --n 16384 --reps 2 --nopath --kind smooth 2>&1 | grep "rep" >> $out
