#!/bin/bash
# lean ncu full captures, summarised on the box (the .ncu-rep files are too big to travel back)
mkdir -p gpurun_out /tmp/ncu
NCU="ncu --clock-control none"
SEC="--section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section WarpStateStats --section ComputeWorkloadAnalysis"
timeout 200 $NCU --set full -k regex:adam_clamp --launch-skip 1 -c 1 -f -o /tmp/ncu/prof_adam python profiles/prof_kernels.py adam > gpurun_out/prof_adam.log 2>&1
timeout 200 $NCU --set full -k regex:structure_loss --launch-skip 2 -c 2 -f -o /tmp/ncu/prof_loss python profiles/prof_kernels.py loss > gpurun_out/prof_loss.log 2>&1
timeout 200 $NCU $SEC -k regex:tail_ -c 3 -f -o /tmp/ncu/prof_tail python profiles/prof_kernels.py tail > gpurun_out/prof_tail.log 2>&1
timeout 300 $NCU $SEC -k regex:'conv_fwd|conv_wgrad' --launch-skip 153 -c 16 -f -o /tmp/ncu/prof_conv python profiles/prof_kernels.py head > gpurun_out/prof_conv.log 2>&1
timeout 200 $NCU $SEC -k regex:'bilinear|bn_bwd_reduce4|weight_pack_multi|wgrad_unpack_multi|act_apply4|bn_bwd_dx4' --launch-skip 260 -c 12 -f -o /tmp/ncu/prof_misc python profiles/prof_kernels.py head > gpurun_out/prof_misc.log 2>&1
python profiles/summarize_ncu.py /tmp/ncu/prof_adam.ncu-rep /tmp/ncu/prof_loss.ncu-rep /tmp/ncu/prof_tail.ncu-rep /tmp/ncu/prof_conv.ncu-rep /tmp/ncu/prof_misc.ncu-rep > gpurun_out/ncu_summary.txt 2>&1
wc -l gpurun_out/ncu_summary.txt; grep -c "^--" gpurun_out/ncu_summary.txt
