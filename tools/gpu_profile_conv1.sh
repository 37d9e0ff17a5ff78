#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 ncu --clock-control none --set full --import-source on -k regex:conv_tcgen05 -s 57 -c 1 -o gpurun_out/prof_conv4_r4 python tools/profile_step.py 2 > gpurun_out/p_conv1.log 2>&1; echo "rc=$?"
