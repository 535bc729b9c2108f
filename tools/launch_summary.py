#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hi]
K, V, G, B = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
agg, seq = collections.defaultdict(lambda: [0, 0.0]), []
for r in rows[hi + 1:]:
    if len(r) <= V:
        continue
    name = r[K].replace("nt::b200::", "").replace("<unnamed>::", "").replace("void ", "").split("(")[0][:40]
    us = float(r[V].replace(",", "")) / 1000.0
    agg[f"{name} g{r[G]} b{r[B]}"][0] += 1
    agg[f"{name} g{r[G]} b{r[B]}"][1] += us
    seq.append((name, us))
tot = sum(v[1] for v in agg.values())
print(f"kernels: {len(seq)}  total {tot:.1f} us")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:20]:
    print(f"{n:70s} n={c:4d} total={t:9.1f}us avg={t / c:7.2f}us share={t / tot:.3f}")
if len(sys.argv) > 2:
    start = next(i for i, (n, u) in enumerate(seq) if sys.argv[2] in n)
    print([(n[:14], round(u, 1)) for n, u in seq[start:start + 18]])
