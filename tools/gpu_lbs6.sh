#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
HF_SKIN=body_parts HF_ITERS=2 timeout 600 ncu --clock-control none --set full --import-source on -k regex:"lbs_pose|lbs_extra" -s 6 -c 2 -o gpurun_out/prof_lbs_small python tools/lbs_time.py > gpurun_out/p_small.log 2>&1; echo "rc=$?"
