#!/bin/bash
# one --set full capture of a representative conv launch (layer3 3x3, the 24th conv launch of the second step)
TAG=${1:-r11}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 ncu --clock-control none --set full --import-source on -k regex:conv_tcgen05 -s 72 -c 1 -f -o gpurun_out/prof_conv3x3_$TAG python tools/profile_step.py 2 > gpurun_out/p_conv1.log 2>&1; echo "rc=$?"
ncu -i gpurun_out/prof_conv3x3_$TAG.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]; r=rows[2]
for k in ['Kernel Name','launch__grid_size','gpu__time_duration.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum']:
    print(k, r[hdr.index(k)][:70])"
