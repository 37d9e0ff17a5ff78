#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
HF_PDL_EARLY=0 timeout 300 python tools/enc_pdl_check.py > gpurun_out/pdl0.log
HF_PDL_EARLY=1 timeout 300 python tools/enc_pdl_check.py > gpurun_out/pdl1.log
HF_NO_PDL=1 timeout 300 python tools/enc_pdl_check.py > gpurun_out/pdln.log
if cmp -s gpurun_out/pdl0.log gpurun_out/pdl1.log; then echo "early == late: identical digests"; else echo "MISMATCH early vs late"; diff gpurun_out/pdl0.log gpurun_out/pdl1.log | head; fi
if cmp -s gpurun_out/pdl0.log gpurun_out/pdln.log; then echo "nopdl == late: identical digests"; else echo "MISMATCH nopdl vs late"; fi
head -3 gpurun_out/pdl1.log
