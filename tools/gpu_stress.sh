#!/bin/bash
# repeat the encoder parity tests and the digest check to catch rare protocol hangs / races in the conv kernel
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
fails=0
for i in 1 2 3 4 5 6; do
  timeout 300 python -m pytest tests/test_gpu_encoder.py -q -x --timeout 200 > gpurun_out/t_stress.log 2>&1 || { fails=$((fails+1)); tail -5 gpurun_out/t_stress.log; }
done
echo "encoder test rounds failed: $fails of 6"
for i in 1 2 3; do timeout 300 python tools/enc_pdl_check.py 2>/dev/null | md5sum; done
timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('200-step bench: value %.0f ms/step %.3f e2e %.0f'%(d['value'],d['ms_per_step'],d['e2e']['value']))"
