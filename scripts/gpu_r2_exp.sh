#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu -k "multiclass or bench_config" 2>&1 | tail -6
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_default.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_default.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'])
r=d.get('roofline',{})
print('roofline', {k:r.get(k) for k in ('achieved','frac','error')})
for sh in r.get('shapes',[]): print('   ', sh['shape'][:60], round(sh['us'],1), round(sh['frac'],3), sh.get('us_without_bn_stats'))
for o in r.get('other_kernels',[]): print('   ', str(o.get('kernel'))[:70], round(o.get('avg_ms',0)*1e3,1), round(o.get('frac',0),3), o.get('error'))
print('cpu_baseline', d.get('cpu_baseline'))
print('eager', json.dumps(d.get('gpu_eager_reference'))[:900])
print('secondary', json.dumps(d.get('secondary'))[:1200])
PY
