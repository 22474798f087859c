#!/bin/bash
# round-2 verification on one B200: GPU suite, smoke, bench lines (default / reference arm), compute-sanitizer on a head step,
# ncu full captures of the kernels the bench line names (summarised on the box)
mkdir -p gpurun_out /tmp/ncu
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 --tb=short 2>&1 | tail -6 > gpurun_out/r2_pytest_gpu.log; tail -2 gpurun_out/r2_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench rc=$?"; tail -1 gpurun_out/r2_bench_default.json | cut -c1-900
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "ref rc=$?"; tail -1 gpurun_out/r2_bench_reference.json | cut -c1-600
for tool in memcheck racecheck; do
  SAN_PREC=bf16 timeout 600 compute-sanitizer --tool $tool python scripts/sanitize_head.py > gpurun_out/r2_sanitizer_$tool.txt 2>&1; echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/r2_sanitizer_$tool.txt | tail -1
done
NCU="ncu --clock-control none"
timeout 400 $NCU --set full -k regex:'boundary_weight|structure_loss|bilinear|mc_loss|conv_fwd2' --launch-skip 16 -c 16 -f -o /tmp/ncu/prof_r2 python profiles/prof_kernels.py r2 > gpurun_out/prof_r2.log 2>&1; echo "ncu r2 rc=$?"
python profiles/summarize_ncu.py /tmp/ncu/prof_r2.ncu-rep > gpurun_out/r2_ncu_bench_kernels.txt 2>&1
grep -c "^--" gpurun_out/r2_ncu_bench_kernels.txt
