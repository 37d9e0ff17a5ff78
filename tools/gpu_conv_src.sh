#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for idx in 54 68 82; do
timeout 600 ncu --clock-control none --set full --import-source on -k regex:conv_tcgen05 -s $idx -c 1 -o gpurun_out/prof_conv_s$idx -f python tools/profile_step.py 2 > gpurun_out/p_conv_s.log 2>&1; echo "rc=$?"
done
