"""Runs a few interior-point iterations of a named workload (for ncu captures).
The last `--capture` iterations run between cudaProfilerStart/Stop, so
`ncu --profile-from-start off` sees exactly those.

    python scripts/profile_run.py --config C3 --n 67108864 --iters 13 --capture 1
"""
import argparse
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paropt_b200 import configs  # noqa: E402
from paropt_b200.api import Context, InteriorPoint, problem_from_config  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C3")
ap.add_argument("--n", type=int, default=1 << 24)
ap.add_argument("--iters", type=int, default=13)
ap.add_argument("--capture", type=int, default=1)
args = ap.parse_args()
ctx = Context(0)
cfg = configs.get(args.config, args.n)
prob = problem_from_config(ctx, cfg)
ip = InteriorPoint(prob, dict(cfg["options"], max_major_iters=1000000, history_level=1))
ip.begin()
ip.iterate(args.iters)
ctx.sync()
rt = None
for name in ("libcudart.so", "libcudart.so.12"):
    try:
        rt = ctypes.CDLL(name)
        break
    except OSError:
        pass
if rt is not None:
    rt.cudaProfilerStart()
ip.iterate(args.capture)
ctx.sync()
if rt is not None:
    rt.cudaProfilerStop()
print("iterations", ip.counters(), "launches", ctx.kernel_launches())
