"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv log) into per-kernel launches / time / share.
Usage: python tools/ncu_launch_summary.py launches.csv "<command that produced it>" > summary.json"""
import csv, json, re, sys
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = rows[0]
ix = {n: i for i, n in enumerate(hdr)}
kern = {}
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    m = re.search(r"(\w+?)(_kernel)?<", r[ix["Kernel Name"]]) or re.search(r"(\w+)\(", r[ix["Kernel Name"]])
    name = m.group(1) if m else r[ix["Kernel Name"]][:40]
    v = float(r[ix["Metric Value"]])
    if r[ix["Metric Unit"]] in ("ns", "nsecond"):
        v /= 1e3
    k = kern.setdefault(name, {"launches": 0, "us": 0.0})
    k["launches"] += 1
    k["us"] += v
tot = sum(k["us"] for k in kern.values())
out = {"command": sys.argv[2] if len(sys.argv) > 2 else "", "note": "cold-cache, serialised per-launch times: compare shares, not absolutes",
       "total_us": round(tot, 1),
       "kernels": {n: {"launches": k["launches"], "us": round(k["us"], 1), "share": round(k["us"] / tot, 4)}
                   for n, k in sorted(kern.items(), key=lambda kv: -kv[1]["us"])}}
json.dump(out, sys.stdout, indent=1)
