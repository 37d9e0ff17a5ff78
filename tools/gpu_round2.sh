#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T="timeout 600"
$T python tools/debug_encoder.py 18 64 > gpurun_out/dbg_enc18.log 2>&1; echo "dbg rc=$?"
$T python -m pytest tests/test_gpu_flow.py -q -m gpu --timeout 300 -k "log_prob or density" > gpurun_out/t_flow.log 2>&1; echo "flow rc=$?"
head -c 6000 gpurun_out/dbg_enc18.log
tail -n 30 gpurun_out/t_flow.log
