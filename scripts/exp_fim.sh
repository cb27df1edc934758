#!/bin/bash
out=gpurun_out/exp_fim8.log
: > $out
for B in 3 4 5 6; do
  for CAP in 48 64 96 128; do
    echo "=== band=$B cap=$CAP" >> $out
    DYMU_FIM_BAND=$B DYMU_FIM_INNER=$CAP timeout 300 python scripts/probe_solve.py --n 4096 --reps 3 --nopath 2>&1 | grep "rep 2" >> $out
  done
done
