#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for i in 1 2 3; do timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>&1; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print('value %.0f e2e %.0f (%.2f ms) ms/step %.3f'%(d['value'],d['e2e']['value'],d['e2e']['ms_per_step'],d['ms_per_step']), d['e2e']['pcie_measured'])
PY
done
