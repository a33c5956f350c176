"""Parity protocol shared by the CPU (oracle vs golden) and GPU (CUDA vs
oracle / golden) tests.

Tolerance (stated once, used everywhere): RTOL = 1e-10, fp64.

  * state quantities (fobj, mu, comp, step length alpha, quasi-Newton b0,
    sums/norms of x, zl, zu, zw, sw, tw and the dense z, s, t, zs, zt):
        |a - b| <= RTOL * max(|a|, |b|, S_key),  S_key = max_k |ref_key(k)|
    i.e. relative to the largest magnitude that quantity takes in the
    reference history.
  * quantities recovered from differences of consecutive iterates -- the accepted
    step length alpha = (x_k - x_{k-1}).p / p.p and the quasi-Newton diagonal
    b0 = y.y / y.s -- lose digits as the step shrinks; their tolerance is
    RTOL * (1 + |x| / |p|) (the conditioning of that subtraction).
  * KKT residual norms (max_prime, max_dual, max_infeas): a residual is a small
    difference of O(|g|) terms, so two correct fp64 implementations can only
    agree to eps * |g|; the norms must satisfy
        |a - b| <= RTOL * (max(|a|, |b|) + G),  G = max(1, |g|_inf, |c|_inf)
    at that iteration.
  * evaluation counters (neval, ngeval) and the quasi-Newton subspace size must
    be identical; the iteration count and convergence status must be identical
    for histories compared over their full length.
"""
import json
import os

RTOL = 1e-10

STATE_KEYS = ("fobj", "mu", "comp", "xsum", "xnorm", "zlsum",
              "zusum", "zwsum", "swsum", "twsum")
DIFF_KEYS = ("alpha", "qn_b0")
ARRAY_KEYS = ("z", "s", "t", "zs", "zt", "c")
RES_KEYS = ("max_prime", "max_dual", "max_infeas")
COUNT_KEYS = ("neval", "ngeval", "qn_size")


def load_golden(name):
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "golden", name + ".json")) as fp:
        return json.load(fp)


def compare_histories(ref, got, rtol=RTOL, max_iters=None):
    """Returns (n_compared, worst_error_per_key, first_violation or None)."""
    n = min(len(ref), len(got))
    if max_iters is not None:
        n = min(n, max_iters)
    scale = {}
    for key in STATE_KEYS + DIFF_KEYS:
        scale[key] = max(abs(r[key]) for r in ref[:n])
    for key in ARRAY_KEYS:
        vals = [abs(v) for r in ref[:n] for v in r[key]]
        scale[key] = max(vals) if vals else 0.0
    worst = {}
    first = None

    def note(k, key, a, b, denom):
        nonlocal first
        err = abs(a - b) / denom if denom > 0.0 else 0.0
        worst[key] = max(worst.get(key, 0.0), err)
        if err > rtol and first is None:
            first = {"iter": k, "key": key, "ref": a, "got": b, "err": err}

    for k in range(n):
        a, b = ref[k], got[k]
        for key in COUNT_KEYS:
            if int(a[key]) != int(b[key]) and first is None:
                first = {"iter": k, "key": key, "ref": a[key], "got": b[key], "err": float("inf")}
        for key in STATE_KEYS:
            note(k, key, a[key], b[key], max(abs(a[key]), abs(b[key]), scale[key]))
        pn = a.get("pnorm2", 0.0) ** 0.5
        cond = 1.0 + (a["xnorm"] / pn if pn > 0.0 else 0.0)
        for key in DIFF_KEYS:
            note(k, key, a[key], b[key], cond * max(abs(a[key]), abs(b[key]), scale[key]))
        for key in ARRAY_KEYS:
            for va, vb in zip(a[key], b[key]):
                note(k, key, va, vb, max(abs(va), abs(vb), scale[key]))
        cmax = max([abs(v) for v in a["c"]] + [0.0])
        G = max(1.0, a["gmax"], cmax)
        for key in RES_KEYS:
            note(k, key, a[key], b[key], max(abs(a[key]), abs(b[key])) + G)
    return n, worst, first
