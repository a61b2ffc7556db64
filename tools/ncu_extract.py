"""Condenses an `ncu --page raw --csv` export to the metrics DESIGN.md / bench.py quote (one row per launch)."""
import csv, sys
src, dst, header = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
rows = list(csv.reader(open(src)))
hdr, units, data = rows[0], rows[1], rows[2:]
WANT = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__cycles_elapsed.avg", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active"]
cols = [hdr.index(w) for w in WANT if w in hdr]
with open(dst, "w", newline="") as f:
    if header:
        f.write("# " + header + "\n")
    w = csv.writer(f)
    w.writerow([hdr[i] + (f" [{units[i]}]" if units[i] else "") for i in cols])
    for r in data:
        w.writerow([r[i] for i in cols])
print(open(dst).read()[:1500])
