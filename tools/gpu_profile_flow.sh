#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 ncu --clock-control none --set full --import-source on -k regex:flow_sample -s 1 -c 1 -o gpurun_out/prof_flow_r5 python tools/profile_step.py 2 > gpurun_out/p_flow.log 2>&1; echo "rc=$?"
