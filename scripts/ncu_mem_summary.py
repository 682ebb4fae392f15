"""Summarise an ncu CSV of the memory-bound kernels (GroupNorm, LayerNorm, casts, small convs):
per kernel name the launches, device time, DRAM bytes (read + write), L2 bytes and the achieved GB/s against the
measured HBM peak.  Cold-cache, serialised per-launch numbers (ncu replays each kernel): compare shares and bytes,
not absolute times — the in-graph GB/s of the same kernels is in the bench line (`roofline_norm`).

  ncu --kernel-name regex:'gn_|layernorm_kernel|cast_kernel|conv_small' --clock-control none --csv \
      --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum \
      --log-file gpurun_out/norm_kernels.csv python scripts/profile_step.py mixed 1 vae
  python scripts/ncu_mem_summary.py gpurun_out/norm_kernels.csv [hbm_peak_gbs]
"""
import collections
import csv
import json
import os
import re
import sys

path = sys.argv[1]
peak = float(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].replace(".", "").isdigit() else None
if peak is None:
    mp = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    peak = json.load(open(mp))["hbm_gbs"] if os.path.exists(mp) else 6650.0
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "nsecond": 1e-9, "us": 1e-6, "usecond": 1e-6,
        "ms": 1e-3, "msecond": 1e-3, "second": 1.0}
per_launch = collections.OrderedDict()
for r in csv.DictReader(lines):
    key = int(r["ID"])
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    d = per_launch.setdefault(key, {"name": name, "grid": r.get("Grid Size", "")})
    d[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * UNIT.get(r.get("Metric Unit", ""), 1.0)
agg = collections.OrderedDict()
for d in per_launch.values():
    a = agg.setdefault(d["name"], collections.Counter())
    a["launches"] += 1
    a["time_s"] += d.get("gpu__time_duration.sum", 0.0)
    a["dram_b"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    a["l2_b"] += d.get("lts__t_bytes.sum", 0.0)
if "--json" in sys.argv:  # totals of the selected kernels, for bench.py's roofline.traffic
    sel = sys.argv[sys.argv.index("--json") + 1]
    n = sum(a["launches"] for k, a in agg.items() if sel in k)
    tot = sum(a["dram_b"] for k, a in agg.items() if sel in k)
    print(json.dumps({"kernels_matching": sel, "launches": int(n), "dram_bytes_total": tot,
                      "dram_bytes_per_launch": tot / max(n, 1), "source": os.path.basename(path)}))
    sys.exit(0)
print(f"HBM peak used: {peak:.1f} GB/s (MEASURED_PEAKS.json)")
print(f"{'kernel':34s} {'launches':>8s} {'time us':>9s} {'avg us':>7s} {'DRAM MB':>9s} {'L2 MB':>9s} {'DRAM GB/s':>10s} "
      f"{'of peak':>8s} {'L2 GB/s':>9s}")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["time_s"]):
    t = a["time_s"]
    print(f"{name:34s} {a['launches']:8d} {t * 1e6:9.1f} {t * 1e6 / a['launches']:7.1f} {a['dram_b'] / 1e6:9.1f} "
          f"{a['l2_b'] / 1e6:9.1f} {a['dram_b'] / t / 1e9:10.1f} {a['dram_b'] / t / 1e9 / peak:8.3f} {a['l2_b'] / t / 1e9:9.1f}")
