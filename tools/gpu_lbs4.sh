#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
HF_SKIN=body_parts HF_ITERS=3 timeout 600 ncu --clock-control none --cache-control none --metrics gpu__time_duration.sum -k regex:lbs -s 9 -c 6 python tools/lbs_time.py 2>&1 | grep -E "lbs_|gpu__time" | paste - - | awk '{print $1, $NF}'
