"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per-kernel totals and the slowest launches.
Usage: python scripts/analyze_launches.py gpurun_out/launches.csv [top_n]"""
import csv, sys, collections, re
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    rows.append((us, name, r.get("Grid Size", ""), r.get("Block Size", ""), int(r["ID"])))
tot = sum(r[0] for r in rows)
print(f"{len(rows)} launches, total {tot/1000:.3f} ms")
agg = collections.defaultdict(lambda: [0.0, 0])
for us, name, g, b, i in rows:
    agg[name][0] += us; agg[name][1] += 1
print("\nper kernel:")
for name, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"  {name:40s} {n:5d} launches  {us/1000:8.3f} ms  {100*us/tot:5.1f}%  avg {us/n:7.1f} us")
print(f"\ntop {topn} launches:")
for us, name, g, b, i in sorted(rows, reverse=True)[:topn]:
    print(f"  #{i:5d} {us:8.1f} us  {name:28s} grid {g:>16s} block {b}")
if "--gemm" in sys.argv:
    print("\nall gemm launches in order:")
    for us, name, g, b, i in rows:
        if "gemm" in name or "splitk" in name:
            print(f"  #{i:5d} {us:8.1f} us  {name:24s} grid {g}")
