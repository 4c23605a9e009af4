#!/usr/bin/env python
"""Per-model kernel breakdown of an ncu launch list of benchmarks/model_bench.py (last forward of each model; a forward
ends with head_kernel).  python tools/launch_breakdown.py gpurun_out/launches_models3.csv [--seq MODEL_INDEX]"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]; ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
data = [(re.sub(r"\(.*", "", r[ki]).replace("vip::<unnamed>::", "").replace("void ", ""), r[gi], float(r[vi].replace(",", "")))
        for r in rows[hi + 1:] if len(r) > vi]
heads = [i for i, d in enumerate(data) if d[0].startswith("head_kernel")]
lens = [heads[0] + 1] + [heads[i] - heads[i - 1] for i in range(1, len(heads))]
# forwards of the same model have the same launch count: group consecutive equal lengths
groups, k = [], 0
while k < len(lens):
    j = k
    while j + 1 < len(lens) and lens[j + 1] == lens[k]: j += 1
    groups.append(j); k = j + 1
for gi_, last in enumerate(groups):
    s = data[(heads[last - 1] + 1 if last > 0 else 0): heads[last] + 1]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, g, t in s: agg[n][0] += 1; agg[n][1] += t
    tot = sum(v[1] for v in agg.values())
    print(f"model #{gi_}: {tot / 1e6:.3f} ms over {len(s)} launches")
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"  {t / 1e6:8.3f} ms {c:4d} {100 * t / tot:5.1f}%  {n[:90]}")
    if "--seq" in sys.argv and int(sys.argv[sys.argv.index("--seq") + 1]) == gi_:
        for n, g, t in s: print(f"{t / 1e3:9.1f} us {g:>16s} {n[:70]}")
