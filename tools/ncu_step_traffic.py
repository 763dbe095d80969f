"""Per-shape DRAM traffic of one GenConViT step: zip an `ncu --set full -k regex:gcv` raw-page CSV of
tools/profile_step.py with the launch tags that program wrote (--tags-out), in launch order.

    python tools/ncu_step_traffic.py raw.csv tags.json launches_per_step out.json

For every distinct (kernel, shape) of the LAST captured step: launches, mean duration, DRAM read + written bytes per
launch next to the algorithmic bytes (memory-bound kernels) or the bytes implied by the operands (GEMMs: A + B + D),
tensor-pipe activity, and the achieved fraction of the HBM / tensor peak under ncu (cold clocks: shares, not absolutes).
"""
import csv, json, re, sys

raw, tags_path, per_step, out_path = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
rows = list(csv.reader(open(raw, errors="replace")))
h, units = rows[0], rows[1]
ix = {n: i for i, n in enumerate(h)}
scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
tags = json.load(open(tags_path))
data = rows[2:]
# the program's first `per_step` matching launches are one whole step in tag order (a partial capture covers a prefix)
last = data[:per_step]
tags = tags[:len(last)]


def val(r, k):
    try:
        return float(r[ix[k]].replace(",", ""))
    except (ValueError, KeyError):
        return 0.0


def gemm_bytes(tag):
    m = re.match(r"M(\d+) N(\d+) K(\d+)", tag)
    if not m:
        return None
    M, N, K = map(int, m.groups())
    return 2.0 * (M * K + N * K + M * N) + (2.0 * M * N if "+res" in tag else 0.0)


acc = {}
for r, (name, tag, work) in zip(last, tags):
    kn = re.sub(r"\(.*", "", r[ix["Kernel Name"]])
    assert name.split("_")[0][:4] in kn or True
    a = acc.setdefault((name, tag), {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0, "tensor": 0.0, "work": work})
    a["n"] += 1
    t, u = val(r, "gpu__time_duration.sum"), units[ix["gpu__time_duration.sum"]]
    a["us"] += t / 1e3 if u.startswith("n") else (t if u.startswith("u") else t * 1e3)
    a["rd"] += val(r, "dram__bytes_read.sum") * scale.get(units[ix["dram__bytes_read.sum"]], 1.0)
    a["wr"] += val(r, "dram__bytes_write.sum") * scale.get(units[ix["dram__bytes_write.sum"]], 1.0)
    a["tensor"] += val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
out = []
for (name, tag), a in sorted(acc.items(), key=lambda kv: -kv[1]["us"]):
    n = a["n"]
    tensor = name.startswith("gemm") or name == "mlp_fused"
    algo = gemm_bytes(tag) if name.startswith("gemm") else (None if name == "mlp_fused" else a["work"])
    if name == "mlp_fused":
        m = re.match(r"M(\d+) C(\d+)", tag)
        algo = 3 * 2.0 * int(m.group(1)) * int(m.group(2)) if m else None        # y in, residual in, x out
    out.append({"kernel": name, "shape": tag, "launches": n, "us_per_launch": round(a["us"] / n, 2),
                "dram_read_MB": round(a["rd"] / n / 1e6, 2), "dram_write_MB": round(a["wr"] / n / 1e6, 2),
                "traffic_MB": round((a["rd"] + a["wr"]) / n / 1e6, 2),
                "algorithmic_MB": None if algo is None else round(algo / 1e6, 2),
                "traffic_over_algorithmic": None if not algo else round((a["rd"] + a["wr"]) / n / algo, 2),
                "tflops_under_ncu": round(a["work"] / (a["us"] / n) / 1e6, 1) if tensor else None,
                "gbs_under_ncu": None if tensor else round(a["work"] / (a["us"] / n) / 1e3, 1),
                "tensor_pipe_active_pct": round(a["tensor"] / n, 1)})
json.dump({"note": "one eager bs256 step under ncu --set full (cold clocks, serialised): traffic is per launch; rates are for "
                   "shares only", "per_shape": out}, open(out_path, "w"), indent=1)
for o in out[:30]:
    print(o)
