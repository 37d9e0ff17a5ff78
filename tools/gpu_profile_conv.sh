#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T="timeout 900"
NCU="ncu --clock-control none"
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,l1tex__m_xbar2l1tex_read_bytes.sum,sm__cycles_elapsed.max,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"
$T $NCU --metrics $M -k regex:conv_tcgen05 -s 53 -c 53 --csv --log-file gpurun_out/prof_conv_r4.csv python tools/profile_step.py 2 > gpurun_out/p_conv.log 2>&1; echo "conv rc=$?"
du -sh gpurun_out
