"""CPU checks of the oracle classes that the stand-alone boundary objects
(pcu_blockmat, pcu_qn) are tested against on the GPU: the block matrix must
solve the dense quasi-definite system it stands for (ParOptSparseMat.cpp:117-190)
and the compact quasi-Newton product must equal the dense formula the reference
checks in examples/limited_memory_test/limited_memory_test.py:
B = b0 I - Z diag(d0) M^-1 diag(d0) Z^T."""
import numpy as np
import pytest

from oracle.ip_oracle import LBFGS, LSR1, BlockMat, SerialComm, Weighting


@pytest.mark.parametrize("nvars,nwcon,start,nw,stride", [(83, 10, 0, 8, 8), (67, 11, 1, 5, 6),
                                                          (40, 0, 0, 1, 1)])
def test_blockmat_solves_the_quasi_definite_system(nvars, nwcon, start, nw, stride):
    rng = np.random.default_rng(nvars)
    w = Weighting(nwcon, start=start, nw=nw, stride=stride, coef0=1.5, coef_rest=-0.75)
    Dinv = 0.2 + rng.random(nvars)
    C = 0.1 + rng.random(nwcon)
    Aw = np.zeros((nwcon, nvars))
    for i in range(nwcon):
        Aw[i, w.idx[i]] = w.coef
    K = np.block([[np.diag(1.0 / Dinv), Aw.T], [Aw, -np.diag(C)]])
    mat = BlockMat(w, nvars, nwcon)
    assert mat.factor(Dinv, C) == 0
    bx, bw = rng.standard_normal(nvars), rng.standard_normal(nwcon)
    for rhs_w in (None, bw):
        yx, yw = mat.apply(bx, rhs_w)
        sol = np.linalg.solve(K, np.concatenate([bx, np.zeros(nwcon) if rhs_w is None else bw]))
        assert np.allclose(yx, sol[:nvars], rtol=1e-11, atol=1e-13)
        assert np.allclose(-yw, sol[nvars:], rtol=1e-11, atol=1e-13)  # the unknown is -yw


@pytest.mark.parametrize("cls,m", [(LBFGS, 4), (LBFGS, 7), (LSR1, 5)])
def test_compact_quasi_newton_matches_the_dense_formula(cls, m):
    n = 60
    rng = np.random.default_rng(m)
    qn = cls(SerialComm(), n, m)
    H = np.diag(1.0 + 3.0 * rng.random(n))
    for k in range(2 * m + 1):  # more updates than memory: the oldest pairs drop out
        s = rng.standard_normal(n)
        y = H @ s + 0.01 * rng.standard_normal(n)
        qn.update(s, y)
        b0, d0, M, Z = qn.compact()
        Zm = np.array(Z).T
        B = b0 * np.eye(n)
        if len(Z):
            B -= Zm @ np.diag(d0) @ np.linalg.solve(M, np.diag(d0) @ Zm.T)
        x = rng.standard_normal(n)
        assert np.allclose(qn.mult(x), B @ x, rtol=1e-10, atol=1e-12)
        acc = np.ones(n)
        qn.mult_add(0.5, x, acc)
        assert np.allclose(acc, 1.0 + 0.5 * (B @ x), rtol=1e-10, atol=1e-12)
        if cls is LBFGS and len(Z):
            # the pairs satisfy the secant condition of the most recent update
            assert np.allclose(B @ s, y, rtol=1e-8, atol=1e-8)


def test_block_form_of_the_block_matrix_against_dense_formulas():
    """BlockMatNB (nwblock = nb > 1, ParOptSparseMat.cpp:72-111, 196-224): the solution
    of [[D, Aw^T], [Aw, -C]] [yx; -yw] = [bx; bw] against a dense solve."""
    from oracle.ip_oracle import BlockMatNB, BlockWeighting
    rng = np.random.default_rng(5)
    nb, nw, nblocks, extra = 3, 5, 7, 4
    n = nblocks * (nw + 1) + extra
    coef = rng.standard_normal((nb, nw))
    w = BlockWeighting(nblocks, coef, start=2, stride=nw + 1)
    Dinv = 0.2 + rng.random(n)
    C = 0.1 + rng.random(w.nwcon)
    bx, bw = rng.standard_normal(n), rng.standard_normal(w.nwcon)
    mat = BlockMatNB(w, n)
    assert mat.factor(Dinv, C) == 0
    yx, yw = mat.apply(bx, bw)
    Aw = np.zeros((w.nwcon, n))
    for b in range(nblocks):
        for r in range(nb):
            Aw[b * nb + r, w.idx[b]] = coef[r]
    K = np.block([[np.diag(1.0 / Dinv), Aw.T], [Aw, -np.diag(C)]])
    sol = np.linalg.solve(K, np.concatenate([bx, bw]))
    assert np.allclose(yx, sol[:n], rtol=1e-11, atol=1e-12)
    assert np.allclose(yw, -sol[n:], rtol=1e-11, atol=1e-12)
    yx3, yw3 = mat.apply(bx)
    sol3 = np.linalg.solve(K, np.concatenate([bx, np.zeros(w.nwcon)]))
    assert np.allclose(yx3, sol3[:n], rtol=1e-11, atol=1e-12)
    assert np.allclose(yw3, -sol3[n:], rtol=1e-11, atol=1e-12)
