#!/bin/bash
# Round-end check: full GPU test suite, smoke(), then the profiling evidence pass (tools/gpu_profile_final.sh <tag>)
TAG=${1:-r8}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/t_all.log 2>&1; echo "tests rc=$?"
tail -n 4 gpurun_out/t_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
bash tools/gpu_profile_final.sh $TAG
