"""GPU parity: ParOptVec kernels through the C ABI vs numpy (oracle side of
ParOptBasicVec, src/ParOptVec.cpp:32-217).  Tolerance: 1e-13 relative for the
reductions (different summation order), bit-exact for the elementwise ops."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from paropt_b200.api import Context
    c = Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("n", [0, 1, 2, 3, 63, 64, 65, 1000, 4097, 1 << 20, (1 << 22) + 5])
def test_vec_ops(ctx, n):
    from paropt_b200.api import PVec
    rng = np.random.default_rng(n)
    xa, ya = rng.standard_normal(n), rng.standard_normal(n)
    x, y = PVec(ctx, n), PVec(ctx, n)
    x.from_numpy(xa)
    y.from_numpy(ya)
    assert len(x) == n
    tol = 1e-13
    ref = float(np.dot(xa, ya))
    assert abs(x.dot(y) - ref) <= tol * max(1.0, float(np.sum(np.abs(xa * ya))))
    assert abs(x.norm() - np.linalg.norm(xa)) <= tol * max(1.0, np.linalg.norm(xa))
    assert x.maxabs() == (np.max(np.abs(xa)) if n else 0.0)
    assert abs(x.l1norm() - np.sum(np.abs(xa))) <= tol * max(1.0, np.sum(np.abs(xa)))
    y.axpy(0.5, x)
    np.testing.assert_array_equal(y.to_numpy(), ya + 0.5 * xa)  # exact product: fma == mul + add
    y.scale(-2.0)
    np.testing.assert_array_equal(y.to_numpy(), -2.0 * (ya + 0.5 * xa))
    y.copyValues(x)
    np.testing.assert_array_equal(y.to_numpy(), xa)
    y.set(3.25)
    np.testing.assert_array_equal(y.to_numpy(), np.full(n, 3.25))
    y.zeroEntries()
    np.testing.assert_array_equal(y.to_numpy(), np.zeros(n))
    x.free()
    y.free()


@pytest.mark.parametrize("n,k", [(1000, 1), (4099, 7), (1 << 18, 8), (100001, 20), (5000, 33)])
def test_mdot(ctx, n, k):
    from paropt_b200.api import PVec
    rng = np.random.default_rng(k)
    xa = rng.standard_normal(n)
    Va = rng.standard_normal((k, n))
    x = PVec(ctx, n)
    x.from_numpy(xa)
    vecs = []
    for j in range(k):
        v = PVec(ctx, n)
        v.from_numpy(Va[j])
        vecs.append(v)
    out = x.mdot(vecs)
    ref = Va @ xa
    scale = np.abs(Va) @ np.abs(xa)
    assert np.all(np.abs(out - ref) <= 1e-13 * scale)
    for v in vecs + [x]:
        v.free()


def test_reduction_is_deterministic(ctx):
    from paropt_b200.api import PVec
    n = (1 << 21) + 3
    rng = np.random.default_rng(5)
    x = PVec(ctx, n)
    x.from_numpy(rng.standard_normal(n))
    vals = {x.dot(x) for _ in range(5)}
    assert len(vals) == 1
    x.free()
