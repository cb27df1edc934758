#!/bin/bash
out=gpurun_out/exp_fim15.log
: > $out
for R in 5 6 8; do
  DYMU_FIM_ROUNDS=$R python planning-path_planning_b200/build.py --force > /dev/null 2>&1
  echo "=== rounds=$R" >> $out
  for k in 1 2; do timeout 120 python scripts/probe_solve.py --n 4096 --reps 4 --nopath 2>&1 | grep "rep 3" >> $out; done
done
python planning-path_planning_b200/build.py --force > /dev/null 2>&1
echo "=== rounds=4 (default)" >> $out
for k in 1 2; do timeout 120 python scripts/probe_solve.py --n 4096 --reps 4 --nopath 2>&1 | grep "rep 3" >> $out; done
