#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_encoder.py -q -x --timeout 600 > gpurun_out/t_enc.log 2>&1; echo "enc tests rc=$?"
tail -n 5 gpurun_out/t_enc.log
bash tools/gpu_pdl.sh 2>&1 | tail -7
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
    print('value %.0f e2e %.0f ms/step %.3f'%(d['value'],d['e2e']['value'],d['ms_per_step']))
    print({k:round(v['ms'],3) for k,v in d['stages'].items()}, d['step_breakdown_ms'], d['roofline']['frac'])
except Exception as e:
    print('bench parse failed',e); print(open('gpurun_out/bench.log').read()[-2000:])
PY
