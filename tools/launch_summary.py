"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i
        break
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for r in rows[start + 1:]:
    if len(r) <= iv:
        continue
    name = r[ik].split("(")[0][:60]
    v = float(r[iv].replace(",", ""))
    v = v / 1000 if r[iu] == "ns" else v * 1000 if r[iu] == "ms" else v
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
print(f"total {tot:.1f} us over {sum(n for n, _ in agg.values())} launches")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"{k:62s} n={n:4d} {t:9.1f}us {100 * t / tot:5.1f}%  avg {t / n:7.1f}")
