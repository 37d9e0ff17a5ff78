#!/bin/bash
# usage: tools/gpu_quick.sh "<pytest -k expr or empty>" : run gpu tests + short bench
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -q -m gpu --timeout 600 -s > gpurun_out/t_all.log 2>&1; echo "tests rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -n 6 gpurun_out/t_all.log
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
    print('value %.0f e2e %.0f ms/step %.3f'%(d['value'],d['e2e']['value'],d['ms_per_step']))
    print({k:round(v['ms'],3) for k,v in d['stages'].items()}, d['step_breakdown_ms'], d['roofline']['frac'])
except Exception as e:
    print('bench parse failed',e); print(open('gpurun_out/bench.log').read()[-2000:])
PY
