#!/bin/bash
export PYTHONUNBUFFERED=1
fails=0
for i in 1 2 3 4 5 6 7 8; do r=$(timeout 300 python -m pytest tests/test_gpu_encoder.py -q -m gpu 2>&1 | tail -1); echo "$r"; done
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>&1; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print('value %.0f e2e %.0f ms/step %.3f'%(d['value'],d['e2e']['value'],d['ms_per_step']), {k:round(v['ms'],3) for k,v in d['stages'].items()})
PY
