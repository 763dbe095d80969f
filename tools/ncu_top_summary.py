"""Zip the raw-page CSV of `ncu --set full ... python tools/kernel_probe.py top` with the launch order that program wrote
(gpurun_out/r2_top_order.json): one record per isolated launch with its shape, DRAM traffic next to the algorithmic bytes,
and the pipe / issue / bank-conflict counters.

    python tools/ncu_top_summary.py raw.csv order.json "<command>" out.json
"""
import csv, json, re, sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
rows = rows[next(i for i, r in enumerate(rows) if "Kernel Name" in r):]      # skip ncu's ==PROF== preamble (--log-file)
h, units = rows[0], rows[1]
ix = {n: i for i, n in enumerate(h)}
order = json.load(open(sys.argv[2]))
scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
keep = ["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size"]


def val(r, k):
    try:
        return float(r[ix[k]].replace(",", ""))
    except (ValueError, KeyError):
        return None


out = []
data = rows[2:]
assert len(data) == len(order), (len(data), len(order))
for r, o in zip(data, order):
    kn = re.sub(r"\(.*", "", r[ix["Kernel Name"]])
    assert o["kernel"].split("_")[0] in kn, (o["kernel"], kn)
    t, u = val(r, "gpu__time_duration.sum"), units[ix["gpu__time_duration.sum"]]
    us = t / 1e3 if u.startswith("n") else (t if u.startswith("u") else t * 1e3)
    rd = val(r, "dram__bytes_read.sum") * scale.get(units[ix["dram__bytes_read.sum"]], 1.0)
    wr = val(r, "dram__bytes_write.sum") * scale.get(units[ix["dram__bytes_write.sum"]], 1.0)
    d = {"kernel": o["kernel"], "shape": o["shape"], "ncu_name": kn[-60:], "us": round(us, 1),
         "dram_read_MB": round(rd / 1e6, 1), "dram_write_MB": round(wr / 1e6, 1), "traffic_MB": round((rd + wr) / 1e6, 1),
         "algorithmic_MB": round(o["algorithmic_bytes"] / 1e6, 1),
         "traffic_over_algorithmic": round((rd + wr) / o["algorithmic_bytes"], 2),
         "hbm_gbs_under_ncu": round(o["algorithmic_bytes"] / us / 1e3, 0),
         "tflops_under_ncu": round(o["flops"] / us / 1e6, 0) if o.get("flops") else None}
    for k in keep:
        d[k] = val(r, k)
    out.append(d)
json.dump({"command": sys.argv[3],
           "note": "one isolated launch of every shape that matters in the bs256 fp16 step; durations under ncu are cold-clock, "
                   "use them for shares; traffic = dram read + write per launch",
           "launches": out}, open(sys.argv[4], "w"), indent=1)
for d in out:
    print(d["kernel"].ljust(14), d["shape"].ljust(26), d["us"], d["traffic_MB"], d["algorithmic_MB"],
          d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"])
