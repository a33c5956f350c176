"""Per-kernel share of the step from an ncu launch list (gpu__time_duration.sum),
next to the shares bench.py measured live with CUDA events.

    python scripts/launch_shares.py profiles/r1_launches.csv profiles/r1_bench_n1.json [iters] > profiles/r1_shares.md
"""
import collections
import csv
import json
import sys

path, bench = sys.argv[1], sys.argv[2]
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
lines = [l for l in open(path) if l.startswith('"')]
rows = list(csv.DictReader(lines))


def short(name):
    name = name.split("(")[0].replace("void ", "").replace("tma_tile_kernel<", "").replace("tile_kernel<", "")
    name = name.rstrip(">")
    if name.startswith("gram"):
        name = "gram_kernel"
    return name.split("<")[0].split(",")[0].replace("ResFT", "ResF")


b = json.load(open(bench))
per_it = b["gpu_launches"] // b["steps"]
tail = rows[-per_it * iters:]
ncu = collections.OrderedDict()
for r in tail:
    n = short(r["Kernel Name"])
    ncu[n] = ncu.get(n, 0.0) + float(r["Metric Value"].replace(",", "")) / 1e6 / iters
live = {short(k): v["ms"] / b["steps"] for k, v in b["roofline"]["kernels"].items()}
sn, sl = sum(ncu.values()), sum(live.values())
print("| kernel | ncu ms/it (cold, serialised) | ncu share | live CUDA-event ms/it | live share |")
print("|---|---|---|---|---|")
for k, v in sorted(ncu.items(), key=lambda kv: -kv[1]):
    lv = live.get(k, float("nan"))
    print("| %s | %.3f | %.1f %% | %.3f | %.1f %% |" % (k, v, 100 * v / sn, lv, 100 * lv / sl))
print("| **sum** | %.3f | | %.3f | |" % (sn, sl))
print()
print("launches per iteration: %d; ms/step of the bench run: %.3f" % (per_it, b["ms_per_step"]))
