#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T="timeout 900"
$T python tools/debug_encoder.py 18 64 > gpurun_out/dbg_enc18.log 2>&1; echo "dbg rc=$?"
$T python -m pytest tests/test_gpu_flow.py -q -m gpu --timeout 300 > gpurun_out/t_flow.log 2>&1; echo "flow rc=$?"
$T python -m pytest tests/test_gpu_encoder.py -q -m gpu --timeout 600 > gpurun_out/t_enc.log 2>&1; echo "enc rc=$?"
$T python -m pytest tests/test_gpu_lbs.py -q -m gpu --timeout 300 > gpurun_out/t_lbs.log 2>&1; echo "lbs rc=$?"
$T python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
$T python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
head -c 2500 gpurun_out/dbg_enc18.log
tail -n 4 gpurun_out/t_flow.log gpurun_out/t_enc.log gpurun_out/t_lbs.log gpurun_out/smoke.log
grep -E "^E  |Error" gpurun_out/t_enc.log | head -20
tail -c 2500 gpurun_out/bench.log
