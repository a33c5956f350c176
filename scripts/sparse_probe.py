"""Times pcu_sparsemat (the device ParOptQuasiDefSparseMat) on two patterns: a chain of
200k constraints (tridiagonal K) and a 160 x 160 grid (5-point-stencil K, minimum-degree
ordering).  Prints one JSON line per run: symbolic seconds (host), factor / apply
milliseconds (device, wall clock around a synchronised call, best of 5), nnz(K), nnz(L),
levels of the elimination tree, launches per factorisation, residual of the solve.

    python scripts/sparse_probe.py
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paropt_b200.api import Context, PVec, QuasiDefSparseMat  # noqa: E402


def chain(n):
    """S1's pattern without the long-range columns: K is tridiagonal (a path: the worst case
    for level scheduling -- every level holds one column, one serial launch)."""
    W = (n - 1) // 3
    rowp = np.arange(W + 1, dtype=np.int64) * 4
    cols = (3 * np.arange(W)[:, None] + np.arange(4)[None, :]).ravel()
    return n, W, rowp, cols


def grid(g):
    """Constraint (r, c) of a g x g grid: an own variable + the variables of its incident
    edges, each shared with one neighbour: K has the 5-point stencil pattern."""
    W = g * g
    n = W + 2 * g * (g - 1)
    rowp, cols = [0], []
    for r in range(g):
        for c in range(g):
            cols.append(r * g + c)
            if c < g - 1:
                cols.append(W + r * (g - 1) + c)
            if c > 0:
                cols.append(W + r * (g - 1) + c - 1)
            if r < g - 1:
                cols.append(W + g * (g - 1) + r * g + c)
            if r > 0:
                cols.append(W + g * (g - 1) + (r - 1) * g + c)
            rowp.append(len(cols))
    return n, W, rowp, cols


CASES = [("chain_200k", chain(600001), "natural"), ("grid_160x160", grid(160), "minimum_degree")]
rng = np.random.default_rng(0)
ctx = Context(0)
out = {}
for name, (n, W, rowp, cols), ordering in CASES:
    data = rng.uniform(0.5, 1.5, len(cols))
    t0 = time.time()
    mat = QuasiDefSparseMat(ctx, n, W, rowp, cols, ordering)
    t_sym = time.time() - t0
    mat.set_data(data)
    Dinv, Cd = PVec(ctx, n), PVec(ctx, W)
    Dinv.from_numpy(rng.uniform(0.2, 2.0, n))
    Cd.from_numpy(rng.uniform(0.1, 1.0, W))
    bx, bw, yx, yw = PVec(ctx, n), PVec(ctx, W), PVec(ctx, n), PVec(ctx, W)
    bx.from_numpy(rng.standard_normal(n))
    bw.from_numpy(rng.standard_normal(W))
    tf, ta = [], []
    for _ in range(5):
        ctx.sync()
        t0 = time.time()
        rc = mat.factor(None, Dinv, Cd)
        ctx.sync()
        tf.append(time.time() - t0)
        t0 = time.time()
        mat.apply(bx, bw, yx, yw)
        ctx.sync()
        ta.append(time.time() - t0)
    # yx = D^-1 (bx + A^T yw) and K yw = bw - A D^-1 bx  =>  A yx + C yw - bw = 0
    r = PVec(ctx, W)
    r.from_numpy(np.zeros(W))
    mat.mult_add(1.0, yx, r)
    res = r.to_numpy() + Cd.to_numpy() * yw.to_numpy() - bw.to_numpy()
    out[name] = dict(n=n, nwcon=W, nnzA=len(cols), ordering=ordering,
                     symbolic_s=round(t_sym, 3), factor_ms=round(min(tf) * 1e3, 3),
                     apply_ms=round(min(ta) * 1e3, 3), rc=rc,
                     residual=float(np.max(np.abs(res))), **mat.info())
    for o in (Dinv, Cd, bx, bw, yx, yw, r, mat):
        o.free()
print(json.dumps(out))
