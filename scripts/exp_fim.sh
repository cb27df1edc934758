#!/bin/bash
out=gpurun_out/exp_fim13.log
: > $out
for B in 0.5 1 1.5 2 3 4 8; do
  for CAP in 16 64; do
    echo "=== band=$B cap=$CAP" >> $out
    DYMU_FIM_BAND=$B DYMU_FIM_INNER=$CAP timeout 120 python scripts/probe_solve.py --n 4096 --reps 3 --nopath 2>&1 | grep "rep 2" >> $out
  done
done
