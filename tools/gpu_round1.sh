#!/bin/bash
# First GPU pass: parity tests per subsystem (separate processes so one failing kernel cannot mask the others),
# smoke, a short bench.  Everything is logged under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
export PYTHONUNBUFFERED=1
T="timeout 600"
$T python -m pytest tests/test_gpu_lbs.py -q -m gpu -x --timeout 300 > gpurun_out/t_lbs.log 2>&1; echo "lbs rc=$?"
$T python -m pytest tests/test_gpu_flow.py -q -m gpu --timeout 300 > gpurun_out/t_flow.log 2>&1; echo "flow rc=$?"
$T python -m pytest tests/test_gpu_encoder.py -q -m gpu --timeout 300 -k single_conv > gpurun_out/t_conv.log 2>&1; echo "conv rc=$?"
$T python -m pytest tests/test_gpu_encoder.py -q -m gpu --timeout 300 -k "not single_conv" > gpurun_out/t_enc.log 2>&1; echo "enc rc=$?"
$T python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
$T python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -n 5 gpurun_out/t_lbs.log gpurun_out/t_flow.log gpurun_out/t_conv.log gpurun_out/t_enc.log gpurun_out/smoke.log
tail -c 3000 gpurun_out/bench.log
