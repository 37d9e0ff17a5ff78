#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
HF_CONV_PAIR=0 timeout 300 python tools/enc_pdl_check.py > gpurun_out/pair0.log
timeout 300 python tools/enc_pdl_check.py > gpurun_out/pair1.log; echo "pair run rc=$?"
if cmp -s gpurun_out/pair0.log gpurun_out/pair1.log; then echo "pair == single: identical digests"; else echo "MISMATCH pair vs single"; diff gpurun_out/pair0.log gpurun_out/pair1.log | head -6; fi
timeout 900 python -m pytest tests/test_gpu_encoder.py -q -x --timeout 600 > gpurun_out/t_enc.log 2>&1; echo "enc tests rc=$?"
tail -n 12 gpurun_out/t_enc.log
