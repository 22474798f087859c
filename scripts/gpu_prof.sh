#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command (eager, same kernels the graphs replay) + full captures
mkdir -p gpurun_out
NCU="ncu --clock-control none"
# launch list: 1 warm-up + 2 timed eager steps; the first ~5200 launches (lazy init + warm-up step) are skipped
timeout 900 $NCU --metrics gpu__time_duration.sum --launch-skip 5200 --launch-count 11000 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-graph --profile-only > gpurun_out/bench_ncu.log 2>&1
python profiles/launch_summary.py gpurun_out/launches_bench.csv | head -45
timeout 300 $NCU --set full --import-source on -k regex:adam_clamp -c 2 -f -o gpurun_out/prof_adam python profiles/prof_kernels.py adam > gpurun_out/prof_adam.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:structure_loss -c 4 -f -o gpurun_out/prof_loss python profiles/prof_kernels.py loss > gpurun_out/prof_loss.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:tail_ -c 3 -f -o gpurun_out/prof_tail python profiles/prof_kernels.py tail > gpurun_out/prof_tail.log 2>&1
timeout 400 $NCU --set full --import-source on -k regex:'conv_fwd|conv_wgrad|bilinear|bn_bwd_reduce4|weight_pack_multi|wgrad_unpack_multi' --launch-skip 260 -c 60 -f -o gpurun_out/prof_head python profiles/prof_kernels.py head > gpurun_out/prof_head.log 2>&1
ls -la gpurun_out/*.ncu-rep
