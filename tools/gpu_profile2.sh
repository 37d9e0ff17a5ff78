#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T="timeout 900"
NCU="ncu --clock-control none"
$T $NCU --metrics gpu__time_duration.sum -c 1200 --csv --log-file gpurun_out/launches_r2.csv python tools/profile_step.py 3 > gpurun_out/p_list.log 2>&1; echo "list rc=$?"
$T $NCU --set full --import-source on -k regex:lbs_skin -s 1 -c 1 -o gpurun_out/prof_lbs_r2 python tools/profile_step.py 2 > gpurun_out/p_lbs.log 2>&1; echo "lbs rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
cat gpurun_out/bench.log | tail -c 1800
du -sh gpurun_out
