"""Debug helper: run a few iterations with the staged kernels enabled up to a tile size and
compare the history against the numpy oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paropt_b200 import configs
from paropt_b200.api import Context, InteriorPoint, problem_from_config
from tests.parity import compare_histories
from oracle.ip_oracle import InteriorPointOracle
from oracle.problems import SepQuad
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8 * 5003
iters = 6
ctx = Context(0)
cfg = configs.get("C3", n)
def run(maxrows, no):
    ctx.set_param("no_tma_tile", no); ctx.set_param("tma_min_tiles", 1); ctx.set_param("tma_max_rows", maxrows)
    ctx.set_param("tma_grid", 7)
    prob = problem_from_config(ctx, cfg)
    ip = InteriorPoint(prob, dict(cfg["options"], history_level=2, max_major_iters=iters))
    ip.optimize(); h = ip.history(); ip.free(); prob.free()
    return h
ref = InteriorPointOracle(SepQuad(**cfg["problem"]), dict(cfg["options"], max_major_iters=iters))
ref.optimize()
for mr, no in ((0, 1), (1, 0), (128, 0), (256, 0), (512, 0), (1024, 0)):
    h = run(mr, no)
    cnt, worst, first = compare_histories(ref.history, h, max_iters=iters - 1)
    print("no_tma", no, "max_rows", mr, "ok" if first is None else "MISMATCH", first if first else "", flush=True)
h0 = run(0, 1)
h1 = run(128, 0)
for k in range(2):
    print("iter", k)
    for key in sorted(h0[k]):
        a, b = h0[k][key], h1[k][key]
        if isinstance(a, (int, float)) and a != b:
            print("   %-12s legacy %.16g  staged %.16g" % (key, a, b))
