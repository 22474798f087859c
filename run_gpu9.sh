#!/bin/bash
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on -f"
timeout 300 $N -k regex:"structure_loss|boundary_weight" -s 3 -c 3 -o gpurun_out/prof_loss python profiles/prof_kernels.py loss > gpurun_out/ncu_loss.log 2>&1; tail -1 gpurun_out/ncu_loss.log
timeout 300 $N -k regex:conv_fwd_kernel -s 102 -c 1 -o gpurun_out/prof_conv_1x1 python profiles/prof_kernels.py head > gpurun_out/ncu_conv1.log 2>&1; tail -1 gpurun_out/ncu_conv1.log
timeout 300 $N -k regex:conv_fwd_kernel -s 143 -c 2 -o gpurun_out/prof_conv_5x5 python profiles/prof_kernels.py head > gpurun_out/ncu_conv2.log 2>&1; tail -1 gpurun_out/ncu_conv2.log
timeout 300 $N -k regex:conv_wgrad_kernel -s 51 -c 3 -o gpurun_out/prof_wgrad python profiles/prof_kernels.py head > gpurun_out/ncu_conv3.log 2>&1; tail -1 gpurun_out/ncu_conv3.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pv2|slabs_to" --csv --log-file gpurun_out/launches_pv2_r1.csv python profiles/prof_kernels.py step > gpurun_out/ncu_step.log 2>&1; wc -l gpurun_out/launches_pv2_r1.csv
du -sh gpurun_out
