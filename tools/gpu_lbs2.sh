#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python tools/lbs_time.py 2>&1 | tail -6
HF_SKIN=body_parts HF_ITERS=2 timeout 600 ncu --clock-control none --set full --import-source on -k regex:lbs_skin_tc2 -s 2 -c 1 -o gpurun_out/prof_lbs2_bp python tools/lbs_time.py > gpurun_out/p_lbs2.log 2>&1; echo "ncu bp rc=$?"
HF_SKIN=random HF_ITERS=2 timeout 600 ncu --clock-control none --set full --import-source on -k regex:lbs_skin_tc2 -s 2 -c 1 -o gpurun_out/prof_lbs2_rnd python tools/lbs_time.py > gpurun_out/p_lbs2r.log 2>&1; echo "ncu rnd rc=$?"
ls -la gpurun_out/*.ncu-rep
