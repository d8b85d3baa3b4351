#!/usr/bin/env python
"""Turn gpurun_out/<tag>/*.ncu-rep + launches.csv into a committed summary under profiles/."""
import csv
import io
import os
import subprocess
import sys
from collections import defaultdict

tag, out = sys.argv[1], sys.argv[2]
src = os.path.join("gpurun_out", tag)
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__cycles_elapsed.max"]
lines = [f"# ncu summary `{tag}` (B200, `--set full --clock-control none`, one launch per kernel from `bench.py --steps 2 --warmup 3`)", ""]
for f in sorted(os.listdir(src)):
    if not f.endswith(".ncu-rep"):
        continue
    raw = subprocess.run(["ncu", "-i", os.path.join(src, f), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        continue
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    lines += [f"## {f[:-8]} — `{d.get('Kernel Name', ('?', ''))[0][:110]}`", "", "| metric | value |", "|---|---|"]
    for w in WANT:
        if w in d:
            lines.append(f"| {w} | {d[w][0]} {d[w][1]} |")
    lines.append("")
# launch list: per-kernel totals and shares
p = os.path.join(src, "launches.csv")
if os.path.exists(p):
    tot = defaultdict(float); cnt = defaultdict(int)
    with open(p) as fh:
        rows = [r for r in csv.reader(fh) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
        name = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        tot[name] += v; cnt[name] += 1
    total = sum(tot.values())
    lines += ["## launch list (all kernels of the bench process; cold-cache serialised times: compare SHARES)", "",
              "| kernel | launches | total us | share |", "|---|---|---|---|"]
    for k in sorted(tot, key=lambda k: -tot[k]):
        lines.append(f"| `{k[:90]}` | {cnt[k]} | {tot[k]:.1f} | {100 * tot[k] / total:.1f}% |")
open(out, "w").write("\n".join(lines) + "\n")
print("wrote", out)

# dram traffic per launch of the dominant kernels -> profiles/r02_traffic.json (bench.py's roofline.traffic reads it)
if len(sys.argv) > 3:
    import json
    traffic = {}
    for f in sorted(os.listdir(src)):
        if not f.endswith(".ncu-rep"):
            continue
        raw = subprocess.run(["ncu", "-i", os.path.join(src, f), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) < 3:
            continue
        d = {h: (v, u) for h, u, v in zip(rows[0], rows[1], rows[2])}

        def to_bytes(key):
            v, u = d.get(key, ("0", "byte"))
            x = float(v.replace(",", ""))
            return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)
        rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
        traffic[f[:-8]] = {"dram_bytes": rd + wr, "read": rd, "write": wr, "report": f"gpurun_out/{tag}/{f} (summary: {out})",
                           "tensor_active_pct": d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", ("", ""))[0],
                           "duration": " ".join(d.get("gpu__time_duration.sum", ("", ""))), "samples": "775 770 (bench.py headline workload)"}
    json.dump(traffic, open(sys.argv[3], "w"), indent=1)
    print("wrote", sys.argv[3])
