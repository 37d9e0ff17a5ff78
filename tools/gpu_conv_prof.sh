#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__m_xbar2l1tex_read_bytes.sum"
timeout 900 ncu --clock-control none --metrics $M -k regex:conv_tcgen05 -s 53 -c 53 --csv --log-file gpurun_out/prof_conv_r5.csv python tools/profile_step.py 2 > gpurun_out/p_conv.log 2>&1; echo "conv rc=$?"
timeout 900 ncu --clock-control none --cache-control none --metrics gpu__time_duration.sum -k regex:conv_tcgen05 -s 53 -c 53 --csv --log-file gpurun_out/prof_conv_r5_warm.csv python tools/profile_step.py 2 > gpurun_out/p_conv2.log 2>&1; echo "conv warm rc=$?"
for mask in 3 7; do echo "mask $mask"; HF_LBS_STAGES=$mask HF_SKIN=body_parts HF_ITERS=50 timeout 300 python tools/lbs_time.py 2>&1 | tail -1; done
