#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T="timeout 900"
NCU="ncu --clock-control none"
$T $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/launches_r1.csv python tools/profile_step.py 2 > gpurun_out/p_list.log 2>&1; echo "list rc=$?"
$T $NCU --set full --import-source on -k regex:lbs_skin -s 1 -c 1 -o gpurun_out/prof_lbs_r1 python tools/profile_step.py 2 > gpurun_out/p_lbs.log 2>&1; echo "lbs rc=$?"
$T $NCU --set full --import-source on -k regex:flow_sample -s 1 -c 1 -o gpurun_out/prof_flow_r1 python tools/profile_step.py 2 > gpurun_out/p_flow.log 2>&1; echo "flow rc=$?"
$T $NCU --set full -k regex:conv_tcgen05 -s 53 -c 53 -o /tmp/prof_conv_r1 python tools/profile_step.py 2 > gpurun_out/p_conv.log 2>&1; echo "conv rc=$?"
ncu -i /tmp/prof_conv_r1.ncu-rep --page raw --csv > gpurun_out/prof_conv_r1_raw.csv 2>/dev/null
ls -la gpurun_out /tmp/prof_conv_r1.ncu-rep
du -sh gpurun_out
