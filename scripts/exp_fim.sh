#!/bin/bash
out=gpurun_out/exp_fim6.log
: > $out
for B in 1 2 3; do
  for CAP in 6 8 12 16 24; do
    echo "=== block band=$B cap=$CAP" >> $out
    DYMU_FIM_KERNEL=0 DYMU_FIM_BAND=$B DYMU_FIM_INNER=$CAP timeout 300 python scripts/probe_solve.py --n 4096 --reps 2 --nopath 2>&1 | grep "rep 1" >> $out
  done
done
