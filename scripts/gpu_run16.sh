#!/bin/bash
mkdir -p gpurun_out /tmp/ncu
timeout 300 python bench_head.py --batches 16 --sizes 352 --iters 10 --kernels --out gpurun_out/head_kernels3.jsonl > gpurun_out/head_kernels3.log 2>&1
grep '"bound": "hbm"' gpurun_out/head_kernels3.jsonl | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(f\"{r['us']:8.2f} us {r['achieved_gbs']:8.1f} GB/s {r['frac_of_hbm_peak']:.3f}  {r['kernel']}\")"
cat > /tmp/prof_bil.py <<'PY'
import sys, os, ctypes
sys.path.insert(0, os.getcwd())
import torch
import pranet_v2_b200 as P
from pranet_v2_b200.ops import PV2_F32, _ratio
lib = P._lib.load()
B, S, dev = 16, 352, "cuda"
scs = (8, 16, 32, 8, 8, 16, 32, 8)
lows = [torch.randn(B, 1, S // s, S // s, device=dev) for s in scs]
his = [torch.randn(B, 1, S, S, device=dev) for _ in scs]
ihs = (ctypes.c_int * 8)(*[S // s for s in scs])
rr = (ctypes.c_float * 8)(*[_ratio(S // s, S, False, float(s)) for s in scs])
pl, k1 = P._lib.ptr_array(lows); ph, k2 = P._lib.ptr_array(his)
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    P._lib.check(lib.pv2_bilinear_multi_fwd(pl, ph, ihs, ihs, rr, rr, 8, B, S, S, 0, PV2_F32, st), "f")
    P._lib.check(lib.pv2_bilinear_multi_bwd(ph, pl, ihs, ihs, rr, rr, 8, B, S, S, 0, PV2_F32, st), "b")
x = torch.randn(B, 512, 44, 44, device=dev).bfloat16(); y = torch.empty_like(x); crop = torch.randn(B, 1, 44, 44, device=dev)
P._lib.check(lib.pv2_ra_v1_scale_fwd(x.data_ptr(), crop.data_ptr(), y.data_ptr(), B, 512, 44 * 44, 1, st), "ra")
torch.cuda.synchronize()
PY
timeout 200 ncu --clock-control none --set full -k regex:'bilinear|ra_v1' --launch-skip 2 -c 3 -f -o /tmp/ncu/prof_bil python /tmp/prof_bil.py > gpurun_out/prof_bil.log 2>&1
python profiles/summarize_ncu.py /tmp/ncu/prof_bil.ncu-rep > gpurun_out/ncu_bil.txt 2>&1
cat gpurun_out/ncu_bil.txt | cut -c1-120
