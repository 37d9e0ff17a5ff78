#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T="timeout 900"
$T python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/t_all.log 2>&1; echo "tests rc=$?"
$T python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -n 5 gpurun_out/t_all.log
tail -c 2600 gpurun_out/bench.log
