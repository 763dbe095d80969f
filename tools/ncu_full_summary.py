"""Summarise `ncu --set full` raw-page CSV of the top kernels into per-launch counters + per-family averages.
Usage: python tools/ncu_full_summary.py raw.csv "<command>" out.json [traffic.json]"""
import csv, json, re, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
h = rows[0]; ix = {n: i for i, n in enumerate(h)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"]
scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
out, fam = [], {}
for r in rows[2:]:
    d = {"kernel": re.sub(r"\(.*", "", r[ix["Kernel Name"]])[:90]}
    for w in want:
        if w in ix:
            try:
                d[w] = float(r[ix[w]].replace(",", ""))
            except ValueError:
                d[w] = r[ix[w]]
            d[w + ".unit"] = rows[1][ix[w]]
    out.append(d)
    m = re.search(r"(gemm_tcgen05|mlp_fused|dwconv7_mma|stem_fused|conv3x3_c16)", d["kernel"])
    f = m.group(1) if m else d["kernel"]
    a = fam.setdefault(f, {"n": 0, "bytes": 0.0, "us": 0.0, "tensor": [], "issue": []})
    a["n"] += 1
    a["bytes"] += sum(d[k] * scale.get(d[k + ".unit"], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    t, u = d["gpu__time_duration.sum"], d["gpu__time_duration.sum.unit"]
    a["us"] += t / 1e3 if u.startswith("n") else (t if u.startswith("u") else t * 1e3)
    a["tensor"].append(d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0))
    a["issue"].append(d.get("sm__issue_active.avg.pct_of_peak_sustained_elapsed", 0.0))
summ = {f: {"launches_captured": a["n"], "dram_bytes_per_launch": a["bytes"] / a["n"], "avg_us": a["us"] / a["n"],
            "tensor_pipe_active_pct_avg": sum(a["tensor"]) / a["n"], "issue_active_pct_avg": sum(a["issue"]) / a["n"]} for f, a in fam.items()}
json.dump({"command": sys.argv[2], "per_kernel_avg": summ, "launches": out}, open(sys.argv[3], "w"), indent=1)
if len(sys.argv) > 4:
    json.dump({k: v["dram_bytes_per_launch"] for k, v in summ.items()}, open(sys.argv[4], "w"), indent=1)
print(json.dumps(summ, indent=1))
for d in out:
    print(d["kernel"][28:80].ljust(52), d["gpu__time_duration.sum"], round(d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0), 1))
