"""GPU parity of the stand-alone boundary objects:
pcu_blockmat <-> ParOptQuasiDefBlockMat (ParOptSparseMat.cpp:41-224) and
pcu_qn <-> ParOptLBFGS / ParOptLSR1 (ParOptQuasiNewton.cpp:162-459, 636-809),
against the oracle classes of the same name.  fp64 tolerances: 1e-13 relative
for the single-pass block solve, 1e-11 for quasi-Newton products (dot products
are summed in a different order)."""
import numpy as np
import pytest

from oracle.ip_oracle import LBFGS, LSR1, BlockMat, SerialComm, Weighting

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from paropt_b200.api import Context
    c = Context(0)
    yield c
    c.close()


def vec(ctx, arr):
    from paropt_b200.api import PVec
    v = PVec(ctx, len(arr))
    v.from_numpy(np.ascontiguousarray(arr, dtype=np.float64))
    return v


def relerr(a, b):
    scale = max(float(np.max(np.abs(a))) if a.size else 0.0, 1e-300)
    return float(np.max(np.abs(a - b)) / scale) if a.size else 0.0


PATTERNS = [
    # (nvars, weighting dict)  -- multi-material blocks, rosenbrock's 5-of-6, ragged, none
    (8 * 4096 + 24, dict(nwcon=4096, wstart=0, nw=8, wstride=8, coef0=1.0, coef_rest=-1.0)),
    (6 * 1000 + 7, dict(nwcon=1000, wstart=1, nw=5, wstride=6, coef0=-1.0, coef_rest=-1.0)),
    (2 * 333 + 1, dict(nwcon=333, wstart=0, nw=2, wstride=2, coef0=2.0, coef_rest=0.5)),
    (1000, dict(nwcon=0, wstart=0, nw=0, wstride=0, coef0=0.0, coef_rest=0.0)),
]


@pytest.mark.parametrize("nvars,w", PATTERNS)
def test_blockmat_factor_and_apply(ctx, nvars, w):
    from paropt_b200.api import PVec, QuasiDefBlockMat
    rng = np.random.default_rng(nvars)
    nwcon = w["nwcon"]
    Dinv = 0.1 + rng.random(nvars)
    Cd = 0.05 + rng.random(nwcon)
    bx = rng.standard_normal(nvars)
    bw = rng.standard_normal(nwcon)
    ow = Weighting(nwcon, start=w["wstart"], nw=max(w["nw"], 1), stride=max(w["wstride"], 1),
                   coef0=w["coef0"], coef_rest=w["coef_rest"])
    ref = BlockMat(ow, nvars, nwcon)
    assert ref.factor(Dinv, Cd) == 0
    mat = QuasiDefBlockMat(ctx, nvars, w)
    dD, dC = vec(ctx, Dinv), vec(ctx, Cd)
    assert mat.factor(None, dD, dC) == 0
    dbx, dbw = vec(ctx, bx), vec(ctx, bw)
    yx, yw = PVec(ctx, nvars), PVec(ctx, nwcon)
    for with_bw in (False, True):
        rx, rw = ref.apply(bx, bw if with_bw else None)
        if with_bw:
            mat.apply(dbx, dbw, yx, yw)
        else:
            mat.apply(dbx, yx, yw)
        gx, gw = yx.to_numpy(), yw.to_numpy()
        assert relerr(rx, gx) < 1e-13 and relerr(rw, gw) < 1e-13
        # the inputs stay unmodified (ParOptSparseMat.cpp:117-121)
        assert np.array_equal(dbx.to_numpy(), bx) and np.array_equal(dbw.to_numpy(), bw)
        # and the pair solves [[D, Aw^T], [Aw, -C]] [yx; -yw] = [bx; bw]
        r1 = gx / Dinv - bx
        ow.add_jac_t(-1.0, gw, r1)
        r2 = Cd * gw - (bw if with_bw else 0.0)
        ow.add_jac(1.0, gx, r2) if nwcon else None
        assert relerr(bx, bx + r1) < 1e-12
        if nwcon:
            assert np.max(np.abs(r2)) < 1e-12 * max(1.0, np.max(np.abs(bw)))
    for o in (dD, dC, dbx, dbw, yx, yw, mat):
        o.free()


def test_blockmat_reports_zero_pivot(ctx):
    from paropt_b200.api import QuasiDefBlockMat
    nvars, nwcon = 80, 10
    w = dict(nwcon=nwcon, wstart=0, nw=8, wstride=8, coef0=1.0, coef_rest=-1.0)
    Dinv = np.ones(nvars)
    Cd = np.ones(nwcon)
    Dinv[24:32] = 0.0
    Cd[3] = 0.0
    mat = QuasiDefBlockMat(ctx, nvars, w)
    dD, dC = vec(ctx, Dinv), vec(ctx, Cd)
    assert mat.factor(None, dD, dC) == 4  # row 3, reported 1-based (0 = success)
    for o in (dD, dC, mat):
        o.free()


@pytest.mark.parametrize("nb,nw,nblocks,start,skip", [(2, 8, 4099, 0, 0), (3, 5, 1001, 2, 1),
                                                      (4, 4, 513, 0, 3), (8, 16, 257, 5, 0),
                                                      (1, 6, 100, 0, 0)])
def test_blockmat_with_nwblock_greater_than_one(ctx, nb, nw, nblocks, start, skip):
    """pcu_blockmat_create_blocks: ParOptQuasiDefBlockMat with nwblock = nb
    (ParOptSparseMat.cpp:72-111 packed-upper dpptrf, :196-224 dpptrs), dense nb x nb blocks
    of Ew = Cdiag + Aw Dinv Aw^T, against the oracle's per-block Cholesky; a block that is
    not positive definite is reported with its row."""
    from oracle.ip_oracle import BlockMatNB, BlockWeighting
    from paropt_b200.api import PVec, QuasiDefBlockMat
    rng = np.random.default_rng(100 * nb + nw)
    stride = nw + skip
    nvars = start + nblocks * stride + 3
    coef = rng.standard_normal((nb, nw))
    w = BlockWeighting(nblocks, coef, start=start, stride=stride)
    Dinv = 0.1 + rng.random(nvars)
    Cd = 0.05 + rng.random(w.nwcon)
    bx, bw = rng.standard_normal(nvars), rng.standard_normal(w.nwcon)
    ref = BlockMatNB(w, nvars)
    assert ref.factor(Dinv, Cd) == 0
    mat = QuasiDefBlockMat(ctx, nvars, blocks=dict(nblocks=nblocks, wstart=start, wstride=stride,
                                                    coef=coef))
    dD, dC = vec(ctx, Dinv), vec(ctx, Cd)
    assert mat.factor(None, dD, dC) == 0
    dbx, dbw = vec(ctx, bx), vec(ctx, bw)
    yx, yw = PVec(ctx, nvars), PVec(ctx, w.nwcon)
    mat.apply(dbx, dbw, yx, yw)
    rx, rw = ref.apply(bx, bw)
    assert relerr(rx, yx.to_numpy()) <= 1e-12 and relerr(rw, yw.to_numpy()) <= 1e-12
    mat.apply(dbx, yx, yw)
    rx, rw = ref.apply(bx)
    assert relerr(rx, yx.to_numpy()) <= 1e-12 and relerr(rw, yw.to_numpy()) <= 1e-12
    # inputs unmodified
    assert np.array_equal(dbx.to_numpy(), bx) and np.array_equal(dbw.to_numpy(), bw)
    # an indefinite block: its first failing row (1-based) comes back
    if nb > 1:
        Cbad = Cd.copy()
        blk = nblocks // 2
        Cbad[blk * nb: (blk + 1) * nb] = -1e6
        dCb = vec(ctx, Cbad)
        rc = mat.factor(None, dD, dCb)
        assert blk * nb < rc <= (blk + 1) * nb, rc
        dCb.free()
    for o in (dD, dC, dbx, dbw, yx, yw, mat):
        o.free()


@pytest.mark.parametrize("kind,n,m,updates", [("bfgs", 5003, 4, 9), ("bfgs", 70001, 10, 14),
                                               ("sr1", 4099, 5, 7)])
def test_quasi_newton_object(ctx, kind, n, m, updates):
    from paropt_b200.api import PVec, QuasiNewton
    rng = np.random.default_rng(n + m)
    comm = SerialComm()
    ref = LBFGS(comm, n, m) if kind == "bfgs" else LSR1(comm, n, m)
    qn = QuasiNewton(ctx, n, kind, m)
    # ParOptLBFGS: 2 m, ParOptLSR1: m (QN.cpp:127, 603)
    assert qn.getMaxLimitedMemorySize() == (2 * m if kind == "bfgs" else m)
    hdiag = 1.0 + 4.0 * rng.random(n)
    x = rng.standard_normal(n)
    dx, dy, dacc = vec(ctx, x), PVec(ctx, n), vec(ctx, np.ones(n))
    for k in range(updates):
        s = rng.standard_normal(n)
        y = hdiag * s + 0.05 * rng.standard_normal(n)
        if k == 3:
            y = -y  # negative curvature: skipped by L-BFGS (QN.cpp:217-227)
        ds, dyv = vec(ctx, s), vec(ctx, y)
        assert qn.update(ds, dyv) == ref.update(s, y)
        ds.free()
        dyv.free()
        qn.mult(dx, dy)
        want = ref.mult(x)
        assert relerr(want, dy.to_numpy()) < 1e-11, k
    b0, d0, M, Z = qn.getCompactMat()
    rb0, rd0, rM, rZ = ref.compact()
    assert len(Z) == len(rZ) == len(d0)
    assert abs(b0 - rb0) <= 1e-12 * abs(rb0)
    assert relerr(np.asarray(rd0), d0) < 1e-12 and relerr(np.asarray(rM), M) < 1e-11
    for zg, zr in zip(Z, rZ):
        assert relerr(zr, zg.to_numpy()) < 1e-12
    acc = np.ones(n)
    ref.mult_add(-0.75, x, acc)
    qn.multAdd(-0.75, dx, dacc)
    assert relerr(acc, dacc.to_numpy()) < 1e-11
    qn.reset()
    qn.mult(dx, dy)
    assert np.array_equal(dy.to_numpy(), x)  # B = I after reset (QN.cpp:132-146)
    for o in (dx, dy, dacc, qn):
        o.free()


def test_interior_point_with_a_supplied_quasi_newton_object(ctx):
    """ParOptInteriorPoint::setQuasiNewton (IP.cpp:1193-1235): with the caller's
    L-BFGS object the optimizer must walk the reference's history exactly as with
    its own, and leave the pairs it gathered in the caller's object."""
    from paropt_b200.api import InteriorPoint, QuasiNewton, problem_from_config
    from tests.parity import compare_histories, load_golden
    gold = load_golden("C2_small")
    cfg = gold["config"]
    prob = problem_from_config(ctx, cfg)
    qn = QuasiNewton(ctx, prob.nvars, "bfgs", cfg["options"]["qn_subspace_size"])
    ip = InteriorPoint(prob, dict(cfg["options"], history_level=2))
    ip.setQuasiNewton(qn)
    ip.optimize()
    n, worst, first = compare_histories(gold["history"], ip.history(), cfg=gold["config"])
    assert first is None and n == len(gold["history"]), (first, worst)
    b0, d0, M, Z = qn.getCompactMat()
    assert len(Z) == 2 * cfg["options"]["qn_subspace_size"] and b0 > 0.0
    ip.setQuasiNewton(None)
    ip.free()
    qn.free()
    prob.free()
