"""GPU parity, function by function: every hot-path entry point of the C ABI
against the numpy restatement of the reference function of the same name
(oracle/ip_oracle.py), on a state taken from a few real interior-point
iterations.  Tolerances are relative to the largest entry of the reference
result: 1e-12 for single-pass kernels, 1e-9 for the KKT step (the CUDA path
solves the same linear system through the Gram identity instead of q sequential
solves, so rounding differs but not the mathematics)."""
import numpy as np
import pytest

from oracle.ip_oracle import InteriorPointOracle, Vars
from oracle.problems import Rosenbrock, SepQuad
from paropt_b200 import configs

pytestmark = pytest.mark.gpu

N_COMP = ("x", "zl", "zu")
W_COMP = ("zw", "sw", "tw", "zsw", "ztw")
ALL_COMP = N_COMP + W_COMP
VARS, RES, UPD, REF = 0, 1, 2, 3


@pytest.fixture(scope="module")
def ctx():
    from paropt_b200.api import Context
    c = Context(0)
    yield c
    c.close()


def relerr(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    if a.size == 0:
        return 0.0
    scale = max(np.max(np.abs(a)), 1e-300)
    return float(np.max(np.abs(a - b)) / scale)


def make_pair(ctx, cfg, iters):
    """Oracle advanced `iters` iterations + a GPU optimizer loaded with its state."""
    from paropt_b200.api import InteriorPoint, problem_from_config
    if cfg["kind"] == "rosenbrock":
        oprob = Rosenbrock(cfg["problem"]["n"] - 1)
    else:
        oprob = SepQuad(**cfg["problem"])
    ora = InteriorPointOracle(oprob, dict(cfg["options"], max_major_iters=iters))
    ora.optimize()
    prob = problem_from_config(ctx, cfg)
    gpu = InteriorPoint(prob, cfg["options"])
    put_vars(gpu, VARS, ora.variables)
    gpu.state_vec(0).from_numpy(ora.lb)
    gpu.state_vec(1).from_numpy(ora.ub)
    gpu.state_vec(2).from_numpy(ora.g)
    for j in range(ora.ncon):
        gpu.state_vec(100 + j).from_numpy(ora.Ac[j])
    c = np.ascontiguousarray(ora.c, dtype=np.float64)
    from paropt_b200 import _lib
    gpu.lib.pcu_ip_set_obj_con(gpu.h, float(ora.fobj), c.ctypes.data_as(_lib.c_double_p))
    gpu.lib.pcu_ip_set_barrier(gpu.h, float(ora.barrier_param), float(ora.rho_penalty_search))
    # replay the stored quasi-Newton pairs (oldest first)
    from paropt_b200.api import PVec
    import ctypes as C
    gpu.lib.pcu_ip_qn_reset(gpu.h)
    for i in range(ora.qn.msub):
        s, y = PVec(ctx, ora.nvars), PVec(ctx, ora.nvars)
        s.from_numpy(ora.qn.S[i])
        y.from_numpy(ora.qn.Y[i])
        ut = C.c_int()
        assert gpu.lib.pcu_ip_qn_update(gpu.h, s.h, y.h, C.byref(ut)) == 0
        assert ut.value == 0
        s.free()
        y.free()
    return ora, gpu, prob


def put_vars(gpu, which, v):
    for i, name in enumerate(ALL_COMP):
        gpu.vars_vec(which, i).from_numpy(getattr(v, name))
    gpu.dense_set(which, {k: getattr(v, k) for k in ("z", "s", "t", "zs", "zt")})


def get_vars(gpu, which, n, nw, nc):
    v = Vars(n, nw, nc)
    for i, name in enumerate(ALL_COMP):
        getattr(v, name)[:] = gpu.vars_vec(which, i).to_numpy()
    d = gpu.dense_get(which)
    for k in ("z", "s", "t", "zs", "zt"):
        getattr(v, k)[:] = d[k]
    return v


def assert_vars_close(ref, got, tol, label):
    for name in ALL_COMP + ("z", "s", "t", "zs", "zt"):
        err = relerr(getattr(ref, name), getattr(got, name))
        assert err <= tol, "%s.%s: rel err %.3e > %.1e" % (label, name, err, tol)


CASES = {
    "C2": (configs.get("C2", 6001), 6),       # odd n, dense constraints, Householder
    "C3": (configs.get("C3", 4096 + 64), 6),  # aligned weighting blocks (shuffle path)
    "C1": (configs.get("C1"), 6),             # 5-of-6 weighting pattern (generic path)
    "C3nw4": (configs.get("C3", 4000, nw=4), 5),
    "C3nw16": (configs.get("C3", 4096 + 48, nw=16), 5),
}


@pytest.mark.parametrize("case", list(CASES))
def test_hot_path_functions(ctx, case):
    import ctypes as C
    from paropt_b200 import _lib
    cfg, iters = CASES[case]
    ora, gpu, prob = make_pair(ctx, cfg, iters)
    n, nw, nc = ora.nvars, ora.nwcon, ora.ncon
    lib, h = gpu.lib, gpu.h
    mu = ora.barrier_param
    dbl = lambda: C.c_double()  # noqa: E731

    # quasi-Newton compact form (ParOptLBFGS::getCompactMat, QN.cpp:471-487)
    b0, qsz = dbl(), C.c_int()
    q = len(ora.qn.Z)
    d0 = np.zeros(max(q, 1))
    M = np.zeros(max(q * q, 1))
    assert lib.pcu_ip_qn_compact(h, C.byref(b0), C.byref(qsz), d0.ctypes.data_as(_lib.c_double_p),
                                 M.ctypes.data_as(_lib.c_double_p)) == 0
    assert qsz.value == q
    assert abs(b0.value - ora.qn.b0) <= 1e-12 * abs(ora.qn.b0)
    assert relerr(ora.qn.M, M[:q * q].reshape(q, q).T) <= 1e-12

    # R1/R2/R3: computeKKTRes + computeResNorm + computeComp
    ora.computeKKTRes(ora.variables, mu, ora.residual)
    assert lib.pcu_ip_kkt_res(h, VARS, mu, RES) == 0
    assert_vars_close(ora.residual, get_vars(gpu, RES, n, nw, nc), 1e-12, "residual")
    mp, md, mi, rn = dbl(), dbl(), dbl(), dbl()
    lib.pcu_ip_res_norm(h, C.byref(mp), C.byref(md), C.byref(mi), C.byref(rn))
    ref_norms = ora.computeResNorm(ora.residual)
    for a, b in zip(ref_norms, (mp.value, md.value, mi.value, rn.value)):
        assert abs(a - b) <= 1e-12 * max(abs(a), 1.0)
    comp = dbl()
    assert lib.pcu_ip_comp(h, C.byref(comp)) == 0
    assert abs(comp.value - ora.computeComp(ora.variables)) <= 1e-13 * abs(comp.value)

    # R4/R5/R10: diagonal, Ew factor, G and Ce
    ora.setUpKKTDiagSystem(ora.variables, 1)
    ora.setUpKKTSystem(ora.variables, 1)
    assert lib.pcu_ip_setup_kkt_diag(h, 1) == 0
    assert lib.pcu_ip_setup_kkt(h, 1) == 0
    assert relerr(ora.Dinv, gpu.state_vec(3).to_numpy()) <= 1e-13
    if nw:
        assert relerr(ora.mat.Cw, gpu.state_vec(4).to_numpy()) <= 1e-12
    G = np.zeros(max(nc * nc, 1))
    Ce = np.zeros(max(q * q, 1))
    qq = C.c_int()
    lib.pcu_ip_get_gram(h, G.ctypes.data_as(_lib.c_double_p), Ce.ctypes.data_as(_lib.c_double_p),
                        C.byref(qq))
    assert relerr(ora.Graw, G[:nc * nc].reshape(nc, nc).T) <= 1e-11
    assert qq.value == q
    assert relerr(ora.Ce_raw, Ce[:q * q].reshape(q, q).T) <= 1e-9

    # R6/R7/R11: computeKKTStep (consumes the residual)
    assert lib.pcu_ip_kkt_step(h, RES, UPD, 1) == 0
    ora.computeKKTStep(ora.variables, ora.residual, ora.update, 1)
    gstep = get_vars(gpu, UPD, n, nw, nc)
    assert_vars_close(ora.update, gstep, 1e-9, "step")

    # R12: addKKTResStep on top of a fresh residual (the refinement residual)
    ora.computeKKTRes(ora.variables, mu, ora.residual)
    ora.addKKTResStep(ora.variables, ora.update, ora.residual)
    put_vars(gpu, UPD, ora.update)  # identical step on both sides
    assert lib.pcu_ip_add_kkt_res_step(h, UPD, RES) == 0
    gres = get_vars(gpu, RES, n, nw, nc)
    # the refinement residual is ~1e-10 of the original one: compare absolutely
    # against the scale of the step equation terms
    for name in ALL_COMP + ("z", "s", "t", "zs", "zt"):
        a, b = getattr(ora.residual, name), getattr(gres, name)
        if a.size:
            scale = max(1.0, float(np.max(np.abs(getattr(ora.update, name)))),
                        float(np.max(np.abs(ora.g))))
            assert np.max(np.abs(a - b)) <= 1e-10 * scale, name

    # R13: computeMaxStep / computeCompStep
    tau = 0.95
    mx, mz = dbl(), dbl()
    assert lib.pcu_ip_max_step(h, tau, UPD, C.byref(mx), C.byref(mz)) == 0
    rx, rz = ora.computeMaxStep(ora.variables, tau, ora.update)
    assert abs(mx.value - rx) <= 1e-13 * rx and abs(mz.value - rz) <= 1e-13 * rz
    cs = dbl()
    assert lib.pcu_ip_comp_step(h, rx, rz, UPD, C.byref(cs)) == 0
    ref_cs = ora.computeCompStep(ora.variables, rx, rz, ora.update)
    assert abs(cs.value - ref_cs) <= 1e-11 * max(abs(ref_cs), comp.value)

    # R16: evalMeritInitDeriv on the alpha_x-scaled step
    _, ax, az = ora.scaleKKTStep(ora.variables, ora.update, tau, comp.value)
    put_vars(gpu, UPD, ora.update)
    rho0 = ora.rho_penalty_search
    m0, dm0 = ora.evalMeritInitDeriv(ora.variables, ora.update, ax)
    lib.pcu_ip_set_barrier(h, float(mu), float(rho0))
    gm, gdm = dbl(), dbl()
    assert lib.pcu_ip_merit_init_deriv(h, ax, C.byref(gm), C.byref(gdm)) == 0
    assert abs(gm.value - m0) <= 1e-11 * max(abs(m0), 1.0)
    assert abs(gdm.value - dm0) <= 1e-9 * max(abs(dm0), 1.0)

    # Q1: ParOptLBFGS::mult
    from paropt_b200.api import PVec
    xv, yv = PVec(ctx, n), PVec(ctx, n)
    rng = np.random.default_rng(3)
    xa = rng.standard_normal(n)
    xv.from_numpy(xa)
    assert lib.pcu_ip_qn_mult(h, xv.h, yv.h) == 0
    assert relerr(ora.qn.mult(xa), yv.to_numpy()) <= 1e-11
    xv.free()
    yv.free()
    gpu.free()
    prob.free()


def test_host_callback_problem_matches_builtin(ctx):
    """The ParOpt.Problem-style host callback boundary: a numpy Rosenbrock problem
    (examples/rosenbrock) driven through pcu_problem_create must reproduce the
    golden history of the reference."""
    from paropt_b200.api import InteriorPoint, Problem
    from tests.parity import compare_histories, load_golden

    class HostRosen(Problem):
        def __init__(self, ctx, n):
            self.o = Rosenbrock(n)
            super().__init__(ctx, n, 2, weighting=dict(nwcon=5, wstart=1, nw=5, wstride=6,
                                                      coef0=-1.0, coef_rest=-1.0, wconst=1.0))

        def getVarsAndBounds(self, x, lb, ub):
            self.o.getVarsAndBounds(x, lb, ub)

        def evalObjCon(self, x):
            return self.o.evalObjCon(x)

        def evalObjConGradient(self, x, g, A):
            return self.o.evalObjConGradient(x, g, A)

    gold = load_golden("C1_small")
    prob = HostRosen(ctx, 999)
    ip = InteriorPoint(prob, dict(gold["config"]["options"], history_level=2))
    ip.optimize()
    n, worst, first = compare_histories(gold["history"], ip.history(), cfg=gold["config"])
    assert first is None, (first, worst)
    assert ip.counters()[0] == gold["final"]["niter"]
    assert prob.h2d_bytes > 0 and prob.d2h_bytes > 0
    ip.free()
    prob.free()


def test_unknown_option_is_rejected(ctx):
    from paropt_b200.api import InteriorPoint, problem_from_config
    prob = problem_from_config(ctx, configs.get("C1"))
    with pytest.raises(ValueError):
        InteriorPoint(prob, {"no_such_option": 1})
    with pytest.raises(ValueError):
        InteriorPoint(prob, {"qn_type": "not_a_type"})
    prob.free()
