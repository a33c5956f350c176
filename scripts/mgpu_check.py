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
from tests.parity import compare_histories, load_golden  # noqa: E402

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
    n, worst, first = compare_histories(gold["history"], hist, max_iters=iters)
    verdict["cases"][name] = {"compared": n, "first_violation": first,
                              "worst": max(worst.values()), "nvars_local": prob.nvars}
    ip.free()
    prob.free()
if ctx.rank == 0:
    print("MGPU_VERDICT " + json.dumps(verdict))
ctx.close()
dist.destroy_process_group()
