#!/bin/bash
# LBS parity tests + bench (no CPU baseline) + per-kernel launch list of one LBS call
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_lbs.py -q -x --timeout 300 -s > gpurun_out/t_lbs.log 2>&1; echo "lbs tests rc=$?"
tail -n 8 gpurun_out/t_lbs.log
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
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:lbs --csv --log-file gpurun_out/lbs_launches.csv python tools/profile_step.py 2 > gpurun_out/ncu_lbs.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/lbs_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[1:]: print(r[ki][:60], r[vi])
PY
