"""Multi-GPU parity worker (launched with torchrun, one rank per GPU): runs the
small named workloads partitioned over the ranks and compares the history with
the reference's (tests/golden).  Rank 0 prints a JSON verdict."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from paropt_b200.api import Context, InteriorPoint, problem_from_config  # noqa: E402
from tests.parity import checked, compare_histories, load_golden  # noqa: E402

local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
ctx = Context(local_rank)
ctx.init_distributed()
verdict = {"world": ctx.size, "cases": {}}
for name, iters in (("C2_small", 41), ("C3_small", 50)):
    gold = load_golden(name)
    cfg = gold["config"]
    prob = problem_from_config(ctx, cfg)
    ip = InteriorPoint(prob, dict(cfg["options"], history_level=2, max_major_iters=iters + 1))
    ip.optimize()
    hist = ip.history()
    n, worst, first = compare_histories(gold["history"], hist, max_iters=iters, cfg=gold["config"])
    verdict["cases"][name] = {"compared": n, "first_violation": first,
                              "worst": max(checked(worst).values()), "nvars_local": prob.nvars}
    ip.free()
    prob.free()
# fused / bulk-copy-staged paths at a size where they are active on every rank,
# against the plain path (PCU_NO_* switches) and against the host-array problem
from paropt_b200 import configs  # noqa: E402
from paropt_b200.api import BuiltinProblem  # noqa: E402

SWITCHES = ("PCU_NO_GRAM_TMA", "PCU_NO_RHSGRAM", "PCU_NO_FUSE21", "PCU_NO_FUSE2S")


def run_big(make, plain):
    for k in SWITCHES:
        if plain:
            os.environ[k] = "1"
        else:
            os.environ.pop(k, None)
    # staged fused passes (tma_tile_kernel) forced on at this size, small grid
    ctx.set_param("no_tma_tile", 1 if plain else 0)
    ctx.set_param("tma_min_tiles", 1)
    ctx.set_param("tma_grid", 7)
    prob = make()
    ip = InteriorPoint(prob, dict(cfg_big["options"], history_level=2, max_major_iters=13))
    ip.optimize()
    hist = ip.history()
    ip.free()
    prob.free()
    for k in SWITCHES:
        os.environ.pop(k, None)
    return hist


for label, nbig in (("C3_big", 8 * 5003 * ctx.size), ("C2_big", 50001 * ctx.size)):
    cfg_big = configs.get(label[:2], nbig)
    fused = run_big(lambda: problem_from_config(ctx, cfg_big), False)
    # the device chain of the KKT solve across the ranks (in-stream all-reduce /
    # all-gather feeding the dense kernel): opt-in at world > 1
    os.environ["PCU_CHAIN"] = "1"
    chained = run_big(lambda: problem_from_config(ctx, cfg_big), False)
    os.environ.pop("PCU_CHAIN", None)
    n0, w0, f0 = compare_histories(fused, chained, max_iters=12, cfg=cfg_big, rtol=1e-12)
    verdict["cases"][label + "_chain"] = {"compared": n0, "first_violation": f0,
                                          "worst": max(checked(w0).values())}
    plain = run_big(lambda: problem_from_config(ctx, cfg_big), True)
    host = run_big(lambda: BuiltinProblem(ctx, "sepquad", host=True, nthreads=2,
                                          **cfg_big["problem"]), False)
    n1, w1, f1 = compare_histories(plain, fused, max_iters=12, cfg=cfg_big)
    n2, w2, f2 = compare_histories(fused, host, max_iters=12, cfg=cfg_big)
    verdict["cases"][label] = {"compared": min(n1, n2), "first_violation": f1 or f2,
                               "worst": max(max(checked(w1).values()), max(checked(w2).values()))}
# C4 (100 dense constraints + L-SR1: 120 columns): wide Gram kernel and the column-split
# staged pass 2 (wide_tile_kernel: whole 64-row tiles per rank) against the plain path;
# 9 iterations -- as far as an L-SR1 history is reproducible (tests/test_oracle_golden.py)
cfg_big = configs.get("C4", 625 * 64 * ctx.size)


def run_c4(plain):
    for k in SWITCHES:
        if plain:
            os.environ[k] = "1"
        else:
            os.environ.pop(k, None)
    ctx.set_param("no_tma_tile", 1 if plain else 0)
    ctx.set_param("tma_min_tiles", 1)
    ctx.set_param("tma_grid", 7)
    prob = problem_from_config(ctx, cfg_big)
    ip = InteriorPoint(prob, dict(cfg_big["options"], history_level=2, max_major_iters=9))
    ip.optimize()
    hist = ip.history()
    ip.free()
    prob.free()
    for k in SWITCHES:
        os.environ.pop(k, None)
    return hist


fused = run_c4(False)
plain = run_c4(True)
n1, w1, f1 = compare_histories(plain, fused, max_iters=8, cfg=cfg_big)
verdict["cases"]["C4_big_wide"] = {"compared": n1, "first_violation": f1,
                                   "worst": max(checked(w1).values())}
if ctx.rank == 0:
    print("MGPU_VERDICT " + json.dumps(verdict))
ctx.close()
dist.destroy_process_group()
