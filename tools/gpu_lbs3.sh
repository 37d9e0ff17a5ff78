#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_lbs.py -q -x --timeout 300 > gpurun_out/t_lbs.log 2>&1; echo "lbs tests rc=$?"
tail -n 3 gpurun_out/t_lbs.log
timeout 300 python tools/lbs_time.py 2>&1 | tail -6
HF_SKIN=body_parts HF_ITERS=2 timeout 600 ncu --clock-control none --metrics gpu__time_duration.sum -k regex:lbs -s 8 -c 4 python tools/lbs_time.py 2>&1 | grep -E "lbs_|gpu__time" | paste - - | awk '{print $1, $NF}'
