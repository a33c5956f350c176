"""Small C3 run (chain active, L-BFGS memory full after 12 iterations) for profiling the
dense kernel of the KKT chain under ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paropt_b200 import configs
from paropt_b200.api import Context, InteriorPoint, problem_from_config
ctx = Context(0)
cfg = configs.get("C3", 8 * 65536)
prob = problem_from_config(ctx, cfg)
ip = InteriorPoint(prob, dict(cfg["options"], max_major_iters=1000))
ip.begin()
ip.iterate(16)
print("iterations", ip.counters()[0], "qn", ip.history()[-1]["qn_size"])
ip.free(); prob.free(); ctx.close()
