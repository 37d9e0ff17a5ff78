#!/bin/bash
export PYTHONUNBUFFERED=1
for mask in 1 2 4 3 7; do echo "mask $mask"; HF_LBS_STAGES=$mask HF_SKIN=body_parts HF_ITERS=50 timeout 300 python tools/lbs_time.py 2>&1 | tail -1; done
