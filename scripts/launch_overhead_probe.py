"""Per-launch fixed cost of the fused kernels: runs the C3 iteration at several sizes on one
GPU with per-kernel CUDA events and fits t(n) = a + b n per kernel.  The intercept a is
what an 8-GPU strong-scaling run (n / 8 rows per GPU) pays per launch on top of the
bandwidth term.  Output: one JSON object (sizes, per-kernel times, fits)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from paropt_b200 import configs  # noqa: E402
from paropt_b200.api import Context, InteriorPoint, problem_from_config  # noqa: E402


def run(ctx, n, steps=30, warmup=12, profile=True):
    cfg = configs.get("C3", n)
    prob = problem_from_config(ctx, cfg)
    ip = InteriorPoint(prob, dict(cfg["options"], max_major_iters=1000000, history_level=1))
    ip.begin()
    ip.iterate(warmup)
    if profile:
        ctx.profile(2)
    ctx.timer_start()
    ip.iterate(steps)
    ms = ctx.timer_stop()
    ctx.profile(0)
    prof = ctx.profile_totals() if profile else {}
    ip.free()
    prob.free()
    return ms / steps, {k: (v[0] / steps, v[1] / steps) for k, v in prof.items()}


def main():
    ctx = Context(0)
    sizes = [1 << 20, 1 << 21, 1 << 22, 1 << 23, 1 << 24]
    out = {"sizes": sizes, "ms_per_step": [], "ms_per_step_noprof": [], "kernels": {}}
    for n in sizes:
        ms, prof = run(ctx, n)
        ms2, _ = run(ctx, n, profile=False)
        out["ms_per_step"].append(ms)
        out["ms_per_step_noprof"].append(ms2)
        for k, (t, c) in prof.items():
            out["kernels"].setdefault(k, []).append(t / max(c, 1e-9) if c >= 0.5 else None)
    fits = {}
    x = np.array(sizes, dtype=float)
    for k, ts in out["kernels"].items():
        if len(ts) == len(sizes) and all(t is not None for t in ts):
            b, a = np.polyfit(x[2:], np.array(ts[2:]), 1)
            fits[k] = {"fixed_us": a * 1e3, "ms_per_Mrow": b * (1 << 20)}
    out["fits"] = fits
    b, a = np.polyfit(x[2:], np.array(out["ms_per_step_noprof"][2:]), 1)
    out["step_fit"] = {"fixed_us": a * 1e3, "ms_per_Mrow": b * (1 << 20)}
    print(json.dumps(out))
    for k, f in sorted(fits.items(), key=lambda kv: -kv[1]["fixed_us"]):
        print("%-20s fixed %7.1f us   %8.4f ms per 2^20 rows   %s" % (
            k, f["fixed_us"], f["ms_per_Mrow"],
            " ".join("%.4f" % t for t in out["kernels"][k])), file=sys.stderr)
    print("step (no events): fixed %.1f us, %.4f ms per 2^20 rows; %s" % (
        out["step_fit"]["fixed_us"], out["step_fit"]["ms_per_Mrow"],
        " ".join("%.4f" % t for t in out["ms_per_step_noprof"])), file=sys.stderr)
    ctx.close()


if __name__ == "__main__":
    main()
