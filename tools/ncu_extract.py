"""Summarise an ncu report (raw page CSV on stdin): per launch name, duration, DRAM bytes, tensor / issue utilisation."""
import csv, sys, json
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
out = []
for r in rows[2:]:
    d = {}
    for w in want:
        if w in idx:
            v = r[idx[w]]
            try:
                v = float(v)
            except ValueError:
                v = v[:90]
            d[w] = v
            if w in idx and w != "Kernel Name":
                d[w + ".unit"] = rows[1][idx[w]]
    out.append(d)
json.dump(out, sys.stdout, indent=1)
