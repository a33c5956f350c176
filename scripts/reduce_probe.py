"""Fixed cost of a fused kernel's grid-wide reduction: times pcu_vec dot (VecRedF: 3 sums +
1 max, register-fed tile_kernel) and axpy (no reduction) at sizes where the data term is
negligible, with per-kernel CUDA events.  PCU_NO_ZEROCOPY=1 drops the mapped-host publish."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paropt_b200.api import Context, PVec as Vec  # noqa: E402

ctx = Context(0)
for n in (4096, 1 << 17, 1 << 20, 1 << 23):
    x, y = Vec(ctx, n), Vec(ctx, n)
    x.set(1.0)
    y.set(2.0)
    for _ in range(20):
        x.dot(y)
        y.axpy(0.0, x)
    ctx.sync()
    ctx.profile(2)
    for _ in range(200):
        x.dot(y)
        y.axpy(0.0, x)
    ctx.sync()
    ctx.profile(0)
    prof = ctx.profile_totals()
    print("n=%d zc=%s  " % (n, "off" if os.environ.get("PCU_NO_ZEROCOPY") else "on") +
          "  ".join("%s %.2f us" % (k, 1e3 * v[0] / max(v[1], 1)) for k, v in sorted(prof.items())))
    x.free()
    y.free()
ctx.close()
