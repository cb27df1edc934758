#!/bin/bash
# Round-2 profile captures (run under gpurun, one GPU): launch list + full capture of the solve kernel.
CMD="python bench.py --steps 2 --warmup 3 --workload plan4096 --no-cpu-baseline"
$CMD > gpurun_out/r2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_launches.log 2>&1
$CMD > gpurun_out/r2_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_fim -s 3 -c 1 -f -o gpurun_out/r2_k_fim $CMD > gpurun_out/r2_ncu_full.log 2>&1
ls -la gpurun_out/r2_*
