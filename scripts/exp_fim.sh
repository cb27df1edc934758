#!/bin/bash
out=gpurun_out/exp_fim11.log
: > $out
DYMU_FIM_WARPS=8 python planning-path_planning_b200/build.py --force > /dev/null 2>&1
for G in 0 3 2; do
  echo "=== warps=8 rounds=4 grid_per_sm=$G" >> $out
  DYMU_FIM_GRID_PER_SM=$G timeout 300 python scripts/probe_solve.py --n 4096 --reps 3 --nopath 2>&1 | grep "rep 2" >> $out
done
python planning-path_planning_b200/build.py --force > /dev/null 2>&1
echo "=== warps=16 rounds=4 (default)" >> $out
python scripts/probe_solve.py --n 4096 --reps 3 --check 2>&1 | tail -5 >> $out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3 >> $out
