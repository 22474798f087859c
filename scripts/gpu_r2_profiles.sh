#!/bin/bash
# round 2 evidence for profiles/: per-kernel tables at 16x352 and 64x704, CUPTI timeline of the head step, ncu launch list of the
# bench command, ncu full captures (summarised on the box), the bench lines.  One GPU.
mkdir -p gpurun_out /tmp/ncu
NCU="ncu --clock-control none"
rm -f gpurun_out/r2_timeline.txt*
PV2_TRACE=gpurun_out/r2_timeline.txt timeout 300 python bench_head.py --batches 16 --sizes 352 --iters 40 --kernels --kernels-at 16x352 --out gpurun_out/r2_head_kernels_16x352.jsonl > gpurun_out/r2_hk16.log 2>&1; echo "kernels 16x352 rc=$?"
timeout 400 python bench_head.py --batches 64 --sizes 704 --iters 10 --kernels --kernels-at 64x704 --out gpurun_out/r2_head_kernels_64x704.jsonl > gpurun_out/r2_hk64.log 2>&1; echo "kernels 64x704 rc=$?"
# launch list of the bench command (eager: ncu cannot see launches made during stream capture), 1 warm-up + 2 timed steps
timeout 900 $NCU --metrics gpu__time_duration.sum --launch-skip 5000 --launch-count 12000 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-graph --profile-only > gpurun_out/r2_bench_ncu.log 2>&1; echo "launch list rc=$?"
python profiles/launch_summary.py gpurun_out/r2_launches_bench.csv > gpurun_out/r2_launch_summary.txt 2>&1; head -12 gpurun_out/r2_launch_summary.txt
# full captures: the kernels the bench line's roofline names, at the benchmarked sizes and at 64 x 704^2 (the convs)
timeout 400 $NCU --set full -k regex:'boundary_weight|structure_loss|bilinear|mc_loss|conv_fwd2' -c 40 -f -o /tmp/ncu/prof_r2 python profiles/prof_kernels.py r2 > gpurun_out/prof_r2.log 2>&1; echo "ncu r2 rc=$?"
PV2_PROF_BS=64x704 timeout 400 $NCU --set full -k regex:'conv_fwd2' -c 12 -f -o /tmp/ncu/prof_r2_conv64 python profiles/prof_kernels.py r2conv > gpurun_out/prof_r2_conv64.log 2>&1; echo "ncu conv64 rc=$?"
timeout 300 $NCU --set full -k regex:'act_apply4|bn_bwd_reduce4_lean|bn_bwd_dx4_lean|conv_wgrad|wgrad_unpack_multi|up2_nhwc' --launch-skip 150 -c 18 -f -o /tmp/ncu/prof_r2_glue python profiles/prof_kernels.py head > gpurun_out/prof_r2_glue.log 2>&1; echo "ncu glue rc=$?"
python profiles/summarize_ncu.py /tmp/ncu/prof_r2.ncu-rep /tmp/ncu/prof_r2_conv64.ncu-rep /tmp/ncu/prof_r2_glue.ncu-rep > gpurun_out/r2_ncu_summary.txt 2>&1
grep -c "^--" gpurun_out/r2_ncu_summary.txt
