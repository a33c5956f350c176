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
  * KKT residual norms (max_prime, max_dual, max_infeas): RELATIVE 1e-10, plus a
    round-off floor of K_EPS = 64 machine epsilons of the terms the residual is a
    difference of, plus what the measured difference of the STATES at that
    iteration explains:
        |a - b| <= RTOL * max(|a|, |b|) + (K_EPS * eps + e_k) * G,
        G = max(1, |g|_inf, |c|_inf, sqrt(n) |x|_2) at that iteration, eps = 2^-52,
        e_k = the worst relative error of the state quantities above at iteration k,
              each against its OWN magnitude at that iteration, capped at RTOL (so
              the rule is never looser than RTOL * (mag + G); in the runs here e_k is
              1e-16 .. 1e-11).
    A residual is a function of the state with O(G) coefficients (dr = dz A + ...):
    two runs whose multipliers differ by e_k cannot agree better than e_k G in rx,
    whatever the size of rx -- the numpy restatement against the reference reaches
    e_k = 1e-11 and |d max_prime| = 1e-11 at max_prime = 5.8e-5 in iteration 26 of
    C2_var_mehrotra, CPU against CPU.
    A residual is a small difference of O(G) terms (rx = zl - zu - g + A^T z,
    rz = -(c - s + t)), each of which two correct fp64 implementations only share to
    a few eps * G: c_j(x) = beta_j + a_j . x alone is a sum over n terms whose
    round-off is bounded by eps |a_j|_inf |x|_1 <= eps sqrt(n) |x|_2 (the
    coefficients of all workloads are O(1)), whatever the size of the result --
    the reference itself, run on 1 and on 2 ranks, differs by 2e-13 in max_infeas
    = 4e-13 at iteration 9 of C2_small (tests/golden/C2_small{,_np2}.json;
    test_reference_two_rank_history_matches_single_rank holds the reference to this
    same rule).  The floor is 1.4e-14 G: it matters only once the norm itself is
    below ~1e-4 G.  n is taken from the workload (`nvars`); without it the last
    term is left out.  With norm_type = l1 (l2) the norm is a sum over n entries
    whose errors are coherent (a 1e-13 relative difference in a dense multiplier
    z_j shifts every rx_i by dz A_ji), so the floor is n (sqrt(n)) times larger.
  * evaluation counters (neval, ngeval) and the quasi-Newton subspace size must
    be identical; the iteration count and convergence status must be identical
    for histories compared over their full length.

compare_histories returns the worst error per key in units where <= RTOL passes
(`worst`; keys starting with "info:" or "rel:" are reported figures, not checks:
"info:e_k" is the largest e_k that entered a residual tolerance), and -- for the residual norms -- also the worst PURE relative error
|a - b| / max(|a|, |b|) over the rows where the norm is above its round-off floor
(`worst["rel:<key>"]`), so that the achieved agreement is reported, not just
pass / fail.  Every history test records its table through `report()`;
tests/conftest.py prints the tables in the terminal summary and writes them to
gpurun_out/parity_worst_<cpu|gpu>.json.
"""
import json
import os

RTOL = 1e-10
K_EPS = 64
EPS = 2.0 ** -52

STATE_KEYS = ("fobj", "mu", "comp", "xsum", "xnorm", "zlsum",
              "zusum", "zwsum", "swsum", "twsum")
DIFF_KEYS = ("alpha", "qn_b0")
ARRAY_KEYS = ("z", "s", "t", "zs", "zt", "c")
RES_KEYS = ("max_prime", "max_dual", "max_infeas")
COUNT_KEYS = ("neval", "ngeval", "qn_size")

REPORTS = {}  # test label -> {"compared": n, "worst": {...}}


def checked(worst):
    """The entries of `worst` that are held to the tolerance (no info: / rel: rows)."""
    return {k: v for k, v in worst.items() if not k.startswith(("info:", "rel:"))}


def report(label, n, worst, first=None):
    """Records the achieved errors of one history comparison (printed by conftest)."""
    REPORTS[label] = {"compared": n, "first_violation": first,
                      "worst": {k: float("%.3e" % v) for k, v in sorted(worst.items())}}


def load_golden(name):
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "golden", name + ".json")) as fp:
        return json.load(fp)


def nvars_of(cfg):
    """Global number of design variables of a named workload (configs.py)."""
    if cfg is None:
        return None
    p = cfg["problem"]
    return int(p["n"]) - 1 if cfg["kind"] == "rosenbrock" else int(p["ntotal"])


def compare_histories(ref, got, rtol=RTOL, max_iters=None, label=None, nvars=None,
                      cfg=None):
    """Returns (n_compared, worst_error_per_key, first_violation or None)."""
    if nvars is None:
        nvars = nvars_of(cfg)
    norm_type = (cfg or {}).get("options", {}).get("norm_type", "infinity")
    nsum = 1.0
    if nvars and norm_type == "l1":
        nsum = float(nvars)
    elif nvars and norm_type == "l2":
        nsum = float(nvars) ** 0.5
    n = min(len(ref), len(got))
    if max_iters is not None:
        n = min(n, max_iters)
    scale = {}
    for key in STATE_KEYS + DIFF_KEYS:
        scale[key] = max(abs(r[key]) for r in ref[:n]) if n else 0.0
    for key in ARRAY_KEYS:
        vals = [abs(v) for r in ref[:n] for v in r[key]]
        scale[key] = max(vals) if vals else 0.0
    worst = {}
    first = None

    def note(k, key, a, b, denom):
        nonlocal first
        err = abs(a - b) / denom if denom > 0.0 else 0.0
        if err != err:  # NaN on either side is a violation
            err = float("inf")
        worst[key] = max(worst.get(key, 0.0), err)
        if err > rtol and first is None:
            first = {"iter": k, "key": key, "ref": a, "got": b, "err": err}

    def relerr(va, vb, sc):
        d = max(abs(va), abs(vb), sc)
        return abs(va - vb) / d if d > 0.0 else 0.0

    for k in range(n):
        a, b = ref[k], got[k]
        for key in COUNT_KEYS:
            if int(a[key]) != int(b[key]) and first is None:
                first = {"iter": k, "key": key, "ref": a[key], "got": b[key], "err": float("inf")}
        # e_k: measured relative difference of the states at this iteration
        e_k = max([relerr(a[key], b[key], 0.0) for key in STATE_KEYS] +
                  [relerr(va, vb, 0.0) for key in ARRAY_KEYS
                   for va, vb in zip(a[key], b[key])])
        e_k = min(e_k, rtol) if e_k == e_k else rtol
        worst["info:e_k"] = max(worst.get("info:e_k", 0.0), e_k)
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
        if nvars:
            G = max(G, nvars ** 0.5 * a["xnorm"])
        floor = (K_EPS * EPS + e_k) * G * nsum
        for key in RES_KEYS:
            mag = max(abs(a[key]), abs(b[key]))
            # |a - b| <= rtol * mag + floor   <=>   |a - b| / (mag + floor / rtol) <= rtol
            note(k, key, a[key], b[key], mag + floor / rtol)
            if mag * rtol > K_EPS * EPS * G * nsum:  # above the round-off floor
                rel = abs(a[key] - b[key]) / mag
                worst["rel:" + key] = max(worst.get("rel:" + key, 0.0), rel)
    if label is None:  # under pytest: the running test (+ a counter for repeated calls)
        cur = os.environ.get("PYTEST_CURRENT_TEST")
        if cur:
            label = cur.split(" ")[0].replace("tests/", "")
            k = 2
            base = label
            while label in REPORTS:
                label = "%s#%d" % (base, k)
                k += 1
    if label is not None:
        report(label, n, worst, first)
    return n, worst, first
