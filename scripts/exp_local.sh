#!/bin/bash
out=gpurun_out/exp_local1.log
: > $out
DYMU_LOCAL_PROFILE=1 python planning-path_planning_b200/build.py --force >> $out 2>&1
DYMU_TRACE_CALLS=1 python scripts/probe_repair.py 2>&1 | grep -E "march|cycles|lane 0|SWEEPING|CONSERVATIVE|dymu_local_propagate|dymu_local_extract|dymu_local_expand" >> $out
python planning-path_planning_b200/build.py --force > /dev/null 2>&1
