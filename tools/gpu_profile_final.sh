#!/bin/bash
# Round-end evidence: launch list of one step, --set full captures of the three stage kernels, per-layer conv metrics.
TAG=${1:-r6}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T="timeout 900"
NCU="ncu --clock-control none"
$T $NCU --metrics gpu__time_duration.sum -c 1200 --csv --log-file gpurun_out/launches_$TAG.csv python tools/profile_step.py 3 > gpurun_out/p_list.log 2>&1; echo "list rc=$?"
$T $NCU --set full --import-source on -k regex:lbs_skin -s 1 -c 1 -o gpurun_out/prof_lbs_$TAG python tools/profile_step.py 2 > gpurun_out/p_lbs.log 2>&1; echo "lbs rc=$?"
$T $NCU --set full --import-source on -k regex:flow_sample -s 1 -c 1 -o gpurun_out/prof_flow_$TAG python tools/profile_step.py 2 > gpurun_out/p_flow.log 2>&1; echo "flow rc=$?"
$T $NCU --metrics gpu__time_duration.sum -c 1200 --csv --log-file gpurun_out/launches_e2e_$TAG.csv python tools/e2e_launches.py 2 > gpurun_out/p_e2e.log 2>&1; echo "e2e list rc=$?"
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,l1tex__m_xbar2l1tex_read_bytes.sum,sm__cycles_elapsed.max,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active"
$T $NCU --metrics $M -k regex:conv_tcgen05 -s 49 -c 49 --csv --log-file gpurun_out/prof_conv_${TAG}.csv python tools/profile_step.py 2 > gpurun_out/p_conv.log 2>&1; echo "conv rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.log 2>&1; echo "bench rc=$?"
cuobjdump -sass humaniflow_b200/lib/libhumaniflow_b200.so | grep -E "UTCHMMA|UTMALDG|UTMASTG|LDTM|UBLKCP" | awk '{print $2}' | sort | uniq -c > gpurun_out/sass_mnemonics_$TAG.txt
du -sh gpurun_out; tail -c 600 gpurun_out/bench_$TAG.log
