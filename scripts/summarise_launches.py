"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares for
the LAST closure evaluation in the capture (delimited by the box_forward kernels that open each closure)."""
import collections, csv, sys
path = sys.argv[1]
lines = [l for l in open(path, errors="ignore") if not l.startswith("==")]
rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
names = [r["Kernel Name"] for r in rows]
durs = [float(r["Metric Value"]) / 1e3 for r in rows]           # us
starts = [i for i, n in enumerate(names) if "box_forward_kernel" in n]
# a closure = from the first of a pair of box_forward launches to the next pair
pairs = [s for k, s in enumerate(starts) if k % 2 == 0]
if len(pairs) >= 2:
    lo, hi = pairs[-2], pairs[-1]
else:
    lo, hi = 0, len(names)
agg = collections.defaultdict(lambda: [0, 0.0])
for n, d in zip(names[lo:hi], durs[lo:hi]):
    agg[n][0] += 1; agg[n][1] += d
tot = sum(v[1] for v in agg.values())
ours = sum(v[1] for k, v in agg.items() if "pcfa::" in k)
print(f"# one closure evaluation: {hi - lo} launches, {tot:.1f} us summed device time (ncu: serialised, cold cache)")
print(f"# pcfa_b200 kernels: {ours:.1f} us = {100 * ours / tot:.1f} % of the step; cuDNN/ATen: {tot - ours:.1f} us")
print(f"{'us':>10} {'share':>7} {'n':>5} {'avg us':>9}  kernel")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{v[1]:10.1f} {100 * v[1] / tot:6.2f}% {v[0]:5d} {v[1] / v[0]:9.2f}  {k[:120]}")
