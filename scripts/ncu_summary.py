"""Condenses an .ncu-rep (ncu --set full) into a per-kernel markdown table and a
traffic JSON (dram bytes per launch), for profiles/.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_ncu_summary.md [traffic.json]
"""
import csv
import io
import json
import subprocess
import sys

rep, out_md = sys.argv[1], sys.argv[2]
out_json = sys.argv[3] if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--print-units", "base"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
keys = [
    ("gpu__time_duration.sum", "time_us", 1e-3),
    ("dram__bytes_read.sum", "dram_rd_MB", 1e-6),
    ("dram__bytes_write.sum", "dram_wr_MB", 1e-6),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct", 1),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_pct", 1),
    ("launch__registers_per_thread", "regs", 1),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pct", 1),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct", 1),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct", 1),
    ("launch__grid_size", "grid", 1),
]
lines = ["| kernel | " + " | ".join(k[1] for k in keys) + " | GB/s |", "|---|" + "---|" * (len(keys) + 1)]
traffic = {}
for r in rows[2:]:
    name = r[col["Kernel Name"]].replace("void ", "")
    short = name.split("(")[0].replace("tile_kernel<", "").rstrip(">")
    vals = []
    d = {}
    for k, label, scale in keys:
        try:
            v = float(r[col[k]].replace(",", "")) * scale
        except Exception:
            v = float("nan")
        d[label] = v
        vals.append("%.1f" % v)
    bw = (d["dram_rd_MB"] + d["dram_wr_MB"]) / max(d["time_us"], 1e-9) * 1e3
    lines.append("| %s | %s | %.0f |" % (short, " | ".join(vals), bw))
    key = "gram_kernel" if short.startswith("gram") else short.replace("tma_", "").split("<")[0].split(",")[0].replace("ResFT", "ResF")
    traffic.setdefault(key, []).append((d["dram_rd_MB"] + d["dram_wr_MB"]) * 1e6)
open(out_md, "w").write("\n".join(lines) + "\n")
if out_json:
    # bytes per launch (mean over the captured launches of each kernel)
    json.dump({k: sum(v) / len(v) for k, v in traffic.items()}, open(out_json, "w"), indent=1)
print("\n".join(lines))
