"""One-paragraph summary of an `ncu --set full` report: duration, clocks, SM / memory throughput, tensor-pipe activity,
DRAM and L2 bytes, issue utilisation.  Usage: python scripts/ncu_rep_summary.py gpurun_out/prof_x.ncu-rep [more.ncu-rep ...]"""
import csv, io, subprocess, sys

KEYS = [("gpu__time_duration.sum", "duration"), ("sm__cycles_elapsed.avg.per_second", "SM clock"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
        ("sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "tensor (HMMA sub-pipe) active cycles / SM"),
        ("sm__cycles_elapsed.avg", "elapsed cycles / SM"),
        ("gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "memory throughput %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"), ("lts__t_bytes.sum", "L2 bytes"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("launch__registers_per_thread", "registers / thread"), ("launch__grid_size", "grid"),
        ("launch__cluster_size", "cluster"), ("launch__shared_mem_per_block_dynamic", "dynamic smem / block")]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        print(rep, ": no data")
        continue
    h, units = rows[0], rows[1]
    for v in rows[2:]:
        d = {n.split(".", 1)[1] if n.split(".")[0] in ("SM_A", "TPC", "SM_B", "FE_A", "HOST") else n: (v[i], units[i])
             for i, n in enumerate(h)}
        name = v[h.index("Kernel Name")] if "Kernel Name" in h else "?"
        print(f"{rep.split('/')[-1]}: {name[:60]}")
        for k, label in KEYS:
            for kk in (k, "TriageCompute." + k):
                if kk in d:
                    print(f"    {label:44s} {d[kk][0]:>16s} {d[kk][1]}")
                    break
