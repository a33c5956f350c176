"""Trust-region front end on the device: pcu_tr -- ParOptTrustRegion's
SL1QP penalty method with the adaptive penalty update over GPU-resident
ParOptQuadraticSubproblem / ParOptInfeasSubproblem models -- against histories of the
unmodified reference's ParOptOptimizer with algorithm = "tr" (tests/golden/C1_tr.json,
C2_tr_small.json; `python -m oracle.make_golden --tr`).

What is compared, per trust-region iteration: the centre x_k through its checksums
(sum, 2-norm, max-abs; full precision, from the reference's writeOutput hook), the
iteration counts of BOTH interior-point solves of the iteration (the `info` column
"14/52": quadratic subproblem / steering problem), acceptance ("rej"), quasi-Newton tags,
and the 12 numeric columns of the reference's log row to their printed 3 digits.
Tolerance on the centres: 1e-8 relative (stated): every trust-region iteration is two
interior-point solves converged to abs_res_tol = 1e-6, whose last-iteration round-off
enters x_k at ~1e-12 .. 1e-10 and accumulates over the iterations."""
import numpy as np
import pytest

from tests.parity import load_golden

pytestmark = pytest.mark.gpu
CENTRE_RTOL = 1e-8


@pytest.fixture(scope="module")
def ctx():
    from paropt_b200.api import Context
    c = Context(0)
    yield c
    c.close()


def run_tr(ctx, cfg, extra=None):
    from paropt_b200.api import TrustRegion, problem_from_config
    prob = problem_from_config(ctx, cfg)
    opts = dict(cfg["options"], tr_write_output_frequency=1, **(extra or {}))
    tr = TrustRegion(prob, opts)
    tr.optimize()
    out = dict(history=tr.history(), converged=tr.converged(),
               x=tr.getOptimizedPoint()[0].to_numpy())
    tr.free()
    prob.free()
    return out


@pytest.mark.parametrize("name", ["C1_tr", "C2_tr_small"])
def test_trust_region_history_matches_reference(ctx, name):
    gold = load_golden(name)
    out = run_tr(ctx, gold["config"])
    hist, rows, centres = out["history"], gold["log"], gold["centres"]
    assert out["converged"]
    assert len(hist) == len(rows) == len(centres), (len(hist), len(rows))
    worst_centre = 0.0
    for k, (h, row, c) in enumerate(zip(hist, rows, centres)):
        assert h["info"] == row["info"], (k, h["info"], row["info"])
        for key in ("xsum", "xnorm", "xmaxabs"):
            err = abs(h[key] - c[key]) / max(abs(c[key]), c["xnorm"], 1e-300)
            worst_centre = max(worst_centre, err)
            assert err <= CENTRE_RTOL, (k, key, h[key], c[key], err)
        # the log row prints 3 significant digits (%12.5e for fobj)
        assert abs(h["fobj"] - row["fobj"]) <= 2e-5 * max(abs(row["fobj"]), 1e-300), (k, "fobj")
        # rho = actual / predicted reduction and the predicted reduction itself are
        # differences of objective values: below ~1e-11 |f| they are round-off (the
        # last rows of C1_tr: model_red = 1.4e-10 at f = 987) and are not compared
        noise = abs(row["model_red"]) <= 1e-11 * max(1.0, abs(row["fobj"]))
        for key in ("infeas", "l1", "linfty", "dx", "tr", "rho", "model_red", "zav", "zmax",
                    "gav", "gmax"):
            if noise and key in ("rho", "model_red"):
                continue
            ref = row[key]
            assert abs(h[key] - ref) <= 6e-3 * abs(ref) + 1e-12, (k, key, h[key], ref)
    final = gold["final"]
    assert abs(float(np.sum(out["x"])) - final["xsum"]) <= CENTRE_RTOL * max(abs(final["xsum"]), final["xnorm"])
    assert abs(float(np.linalg.norm(out["x"])) - final["xnorm"]) <= CENTRE_RTOL * final["xnorm"]
    print("%s: %d trust-region iterations, worst centre error %.2e" % (name, len(hist), worst_centre))


def test_optimizer_facade_runs_the_trust_region_by_default(ctx):
    """ParOpt.Optimizer without an `algorithm` runs the trust-region front end, like the
    reference (ParOptOptimizer.cpp:41), on a host-callback problem."""
    from paropt_b200 import ParOpt
    n = 64
    rng = np.random.default_rng(3)
    lam = 1.0 + rng.random(n)
    a = 0.5 + rng.random(n)

    class Quad(ParOpt.Problem):
        def __init__(self):
            super().__init__(ctx, nvars=n, ncon=1)

        def getVarsAndBounds(self, x, lb, ub):
            x[:] = 0.3
            lb[:] = -2.0
            ub[:] = 2.0

        def evalObjCon(self, x):
            return 0, float(0.5 * np.dot(lam * x, x) - np.sum(x)), [float(np.dot(a, x)) - 1.0]

        def evalObjConGradient(self, x, g, A):
            g[:] = lam * x - 1.0
            A[0][:] = a
            return 0

    prob = Quad()
    opt = ParOpt.Optimizer(prob, {"tr_max_iterations": 60, "output_level": 0})
    opt.optimize()
    x = np.asarray(opt.getOptimizedPoint()[0])
    # KKT of the equality-free QP: lam x - 1 = z a with z >= 0 and a.x >= 1
    z = opt.getOptimizedPoint()[1]
    assert opt.tr.converged()
    assert np.dot(a, x) - 1.0 >= -1e-5
    assert np.max(np.abs(lam * x - 1.0 - z[0] * a)) <= 1e-4
    opt.tr.free()
    prob.free()
