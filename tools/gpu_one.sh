#!/bin/bash
# usage: tools/gpu_one.sh <pytest args...> : run selected GPU tests
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest "$@" -q --timeout 600 > gpurun_out/t_one.log 2>&1; echo "tests rc=$?"
tail -n 25 gpurun_out/t_one.log
