#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 ncu --clock-control none --cache-control none --metrics gpu__time_duration.sum -k regex:conv_tcgen05 -s 49 -c 49 --csv --log-file gpurun_out/prof_conv_dual.csv python tools/profile_step.py 2 > gpurun_out/p_conv2.log 2>&1; echo "dual rc=$?"
HF_CONV_DBG=16 timeout 900 ncu --clock-control none --cache-control none --metrics gpu__time_duration.sum -k regex:conv_tcgen05 -s 49 -c 49 --csv --log-file gpurun_out/prof_conv_single.csv python tools/profile_step.py 2 > gpurun_out/p_conv2.log 2>&1; echo "single rc=$?"
