"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.
    python profiles/launch_summary.py gpurun_out/launches.csv [--last-fraction 0.5]
Times are cold-cache and serialised (ncu replays each launch alone): compare SHARES, not absolutes."""
import collections
import csv
import re
import sys

path = sys.argv[1]
frac = float(sys.argv[sys.argv.index("--last-fraction") + 1]) if "--last-fraction" in sys.argv else 1.0
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
data = [(r[ki], float(r[vi].replace(",", "")), r[ui]) for r in rows[hi + 1:] if len(r) > vi]
data = data[int(len(data) * (1 - frac)):]
agg = collections.defaultdict(lambda: [0, 0.0])
for k, v, u in data:
    v = v / 1000 if u in ("nsecond", "ns") else v
    m = re.search(r"pv2::(?:\(anonymous namespace\)::|<unnamed>::)?(\w+)", k)
    name = "pv2::" + m.group(1) if m else ("pv2::slabs_to_nhwc_kernel" if "slabs_to" in k else "torch/cudnn")
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
pv2 = sum(v[1] for k, v in agg.items() if k.startswith("pv2"))
print(f"# {path}: {len(data)} launches, {tot:.1f} us total (serialised), pv2 {pv2:.1f} us in {sum(v[0] for k, v in agg.items() if k.startswith('pv2'))} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:10.1f} us {100 * v[1] / tot:5.1f}% n={v[0]:5d} avg {v[1] / v[0]:7.2f} us  {k}")
