"""GPU parity of pcu_sparsemat <-> ParOptQuasiDefSparseMat (ParOptSparseMat.cpp:231-451):
the factor / apply pair and the two CSR products against dense numpy algebra on the same
inputs.  fp64 tolerance: 1e-11 relative (a sparse Cholesky solve; the reference's own
factorisation uses another elimination order)."""
import numpy as np
import pytest

from tests.test_cpu_sparse import random_csr

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


def dense(nw, nv, rowp, cols, data):
    A = np.zeros((nw, nv))
    for i in range(nw):
        for e in range(rowp[i], rowp[i + 1]):
            A[i, cols[e]] = data[e]
    return A


@pytest.mark.parametrize("kind,nw", [("chain", 700), ("arrow", 300), ("blocks", 5000),
                                     ("random", 2000), ("random", 1)])
@pytest.mark.parametrize("ordering", ["minimum_degree", "natural"])
def test_sparsemat_factor_and_apply(ctx, kind, nw, ordering):
    from paropt_b200.api import PVec, QuasiDefSparseMat
    rng = np.random.default_rng(11 + nw)
    nv = 3 * nw + 3
    rowp, cols, data = random_csr(rng, nw, nv, kind)
    A = dense(nw, nv, rowp, cols, data)
    Dinv = rng.uniform(0.2, 2.0, nv)
    Cd = rng.uniform(0.1, 1.0, nw)
    K = np.diag(Cd) + (A * Dinv) @ A.T
    mat = QuasiDefSparseMat(ctx, nv, nw, rowp, cols, ordering)
    mat.set_data(data)
    dD, dC = vec(ctx, Dinv), vec(ctx, Cd)
    assert mat.factor(None, dD, dC) == 0
    bx, bw = rng.standard_normal(nv), rng.standard_normal(nw)
    dbx, dbw = vec(ctx, bx), vec(ctx, bw)
    yx, yw = PVec(ctx, nv), PVec(ctx, nw)
    for with_bw in (False, True):
        rhs = (bw if with_bw else 0.0) - A @ (Dinv * bx)
        rw = np.linalg.solve(K, rhs)
        rx = Dinv * (bx + A.T @ rw)
        if with_bw:
            mat.apply(dbx, dbw, yx, yw)
        else:
            mat.apply(dbx, yx, yw)
        gx, gw = yx.to_numpy(), yw.to_numpy()
        sx, sw = max(1.0, np.max(np.abs(rx))), max(1.0, np.max(np.abs(rw)))
        assert np.max(np.abs(gx - rx)) <= 1e-11 * sx and np.max(np.abs(gw - rw)) <= 1e-11 * sw
        assert np.array_equal(dbx.to_numpy(), bx) and np.array_equal(dbw.to_numpy(), bw)
    # a second factorisation with other values reuses the symbolic phase
    data2 = data * rng.uniform(0.5, 2.0, data.size)
    A2 = dense(nw, nv, rowp, cols, data2)
    K2 = np.diag(Cd) + (A2 * Dinv) @ A2.T
    mat.set_data(data2)
    assert mat.factor(None, dD, dC) == 0
    mat.apply(dbx, dbw, yx, yw)
    rw = np.linalg.solve(K2, bw - A2 @ (Dinv * bx))
    assert np.max(np.abs(yw.to_numpy() - rw)) <= 1e-11 * max(1.0, np.max(np.abs(rw)))
    # CSR products (addSparseJacobian / addSparseJacobianTranspose)
    out_w, out_x = vec(ctx, bw), vec(ctx, bx)
    mat.mult_add(-0.75, dbx, out_w)
    mat.mult_transpose_add(1.5, dbw, out_x)
    assert np.max(np.abs(out_w.to_numpy() - (bw - 0.75 * A2 @ bx))) <= 1e-12 * max(1.0, np.max(np.abs(bw)) * 4)
    assert np.max(np.abs(out_x.to_numpy() - (bx + 1.5 * A2.T @ bw))) <= 1e-12 * max(1.0, np.max(np.abs(bx)) * 4)
    info = mat.info()
    assert info["nnzL"] >= info["nnzK"] >= nw
    for o in (dD, dC, dbx, dbw, yx, yw, out_w, out_x, mat):
        o.free()


def test_sparsemat_reports_a_non_positive_pivot(ctx):
    from paropt_b200.api import QuasiDefSparseMat
    rng = np.random.default_rng(3)
    nw, nv = 50, 153
    rowp, cols, data = random_csr(rng, nw, nv, "blocks")
    Dinv, Cd = np.ones(nv), np.ones(nw)
    Dinv[3 * 17:3 * 17 + 3] = 0.0
    Cd[17] = 0.0
    mat = QuasiDefSparseMat(ctx, nv, nw, rowp, cols)
    mat.set_data(data)
    dD, dC = vec(ctx, Dinv), vec(ctx, Cd)
    assert mat.factor(None, dD, dC) == 18  # constraint 17, reported 1-based (0 = success)
    for o in (dD, dC, mat):
        o.free()
