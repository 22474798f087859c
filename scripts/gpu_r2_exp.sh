#!/bin/bash
mkdir -p gpurun_out /tmp/ncu
timeout 400 ncu --clock-control none --set full -k regex:'act_apply4_lean|bn_bwd_reduce4_lean|bn_bwd_dx4_lean|up2_nhwc_bwd4_tiled|conv_wgrad|wgrad_unpack_multi' --launch-skip 120 -c 18 -f -o /tmp/ncu/prof_glue python profiles/prof_kernels.py head > gpurun_out/prof_r2_glue.log 2>&1; echo "ncu rc=$?"
python profiles/summarize_ncu.py /tmp/ncu/prof_glue.ncu-rep > gpurun_out/r2_ncu_glue_final.txt 2>&1; grep -c "^--" gpurun_out/r2_ncu_glue_final.txt
