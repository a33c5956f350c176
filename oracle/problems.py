"""
oracle/problems.py -- TEST INFRASTRUCTURE ONLY.

numpy statements of the synthetic problems of DESIGN.md ("Synthetic problems"),
written independently of oracle/ref_driver.cpp (C++, reference side) and of
paropt_b200/csrc/pcu_problems.cu (CUDA, product side) so the three check each
other.  The inputs are produced by a counter-based generator, so any rank can
produce its slice of the global vectors and results do not depend on the
partition.
"""
import numpy as np

from .ip_oracle import Weighting

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(z):
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def stream_key(seed, stream):
    with np.errstate(over="ignore"):
        k = np.uint64(seed) ^ (np.uint64(stream) * np.uint64(0x9E3779B97F4A7C15))
    return splitmix64(k)


def uniform01(key, idx):
    idx = np.asarray(idx, dtype=np.uint64)
    with np.errstate(over="ignore"):
        h = splitmix64(np.uint64(key) + idx)
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


SEPQUAD_DEFAULTS = dict(
    ntotal=1000, ncon=1, nw=0, seed=0, lam_min=1.0, lam_max=1e3, b_lo=0.0,
    b_w=1.0, a_lo=0.0, a_w=1.0, beta_c=0.0, beta_n=0.0, beta_u=1.0,
    x0_lo0=-2.0, x0_lo1=-2.0, x0_w0=1.0, x0_w1=1.0, lb0=-5.0, lb1=-5.0,
    ub0=5.0, ub1=5.0, householder=0,
)


def partition(ntotal, nw, rank, size):
    """Block-row partition in units of one weighting block (IP.cpp:214-229)."""
    unit = nw if nw > 0 else 1
    nunits = ntotal // unit
    u0 = (nunits * rank) // size
    u1 = (nunits * (rank + 1)) // size
    offset = u0 * unit
    n = (u1 - u0) * unit
    if rank == size - 1:
        n = ntotal - offset
    nwc = (u1 - u0) if nw > 0 else 0
    return offset, n, nwc


class SepQuad:
    """f = 1/2 (Px)^T diag(lam) (Px) + b^T x, c_j = beta_j + a_j^T x >= 0,
    cw_i = x[nw i] - sum_{k>=1} x[nw i + k] >= 0."""

    def __init__(self, comm=None, **kw):
        p = dict(SEPQUAD_DEFAULTS)
        for k, v in kw.items():
            if k not in p:
                raise ValueError("unknown sepquad parameter %s" % k)
            p[k] = v
        self.p = p
        self.comm = comm
        rank = comm.rank if comm else 0
        size = comm.size if comm else 1
        self.offset, self.nvars, self.nwcon = partition(p["ntotal"], p["nw"], rank, size)
        self.ncon = p["ncon"]
        self.ninequality = self.ncon
        self.nwinequality = self.nwcon
        n = self.nvars
        self.gi = np.arange(self.offset, self.offset + n, dtype=np.uint64)
        seed = p["seed"]
        self.lam = p["lam_min"] + (p["lam_max"] - p["lam_min"]) * uniform01(stream_key(seed, 1), self.gi)
        self.b = p["b_lo"] + p["b_w"] * uniform01(stream_key(seed, 2), self.gi)
        self.vh = 0.5 + uniform01(stream_key(seed, 7), self.gi)
        self.vtv = self._sum([np.sum(self.vh * self.vh)])[0]
        self.beta = np.array([
            p["beta_c"] + p["beta_n"] * float(p["ntotal"])
            + p["beta_u"] * float(uniform01(stream_key(seed, 5), np.uint64(j)))
            for j in range(self.ncon)])
        self.A = [p["a_lo"] + p["a_w"] * uniform01(stream_key(seed, 100 + j), self.gi)
                  for j in range(self.ncon)]
        nw = p["nw"]
        self.cls = ((np.arange(n) % nw) != 0).astype(np.int64) if nw > 0 else np.zeros(n, np.int64)
        self.weighting = Weighting(self.nwcon, 0, nw if nw > 0 else 1, nw if nw > 0 else 1,
                                   1.0, -1.0, 0.0)

    def _sum(self, vals):
        if self.comm:
            return self.comm.allreduce(vals, "sum")
        return np.asarray(vals, dtype=np.float64)

    def getVarsAndBounds(self, x, lb, ub):
        p = self.p
        u = uniform01(stream_key(p["seed"], 3), self.gi)
        c1 = self.cls == 1
        x[:] = np.where(c1, p["x0_lo1"] + p["x0_w1"] * u, p["x0_lo0"] + p["x0_w0"] * u)
        lb[:] = np.where(c1, p["lb1"], p["lb0"])
        ub[:] = np.where(c1, p["ub1"], p["ub0"])

    def _applyP(self, x):
        if not self.p["householder"]:
            return x
        vx = self._sum([np.dot(self.vh, x)])[0]
        return x - (2.0 * vx / self.vtv) * self.vh

    def evalObjCon(self, x):
        y = self._applyP(x)
        loc = [np.sum(0.5 * self.lam * y * y + self.b * x)]
        loc += [np.dot(a, x) for a in self.A]
        out = self._sum(loc)
        return 0, float(out[0]), self.beta + out[1:]

    def evalObjConGradient(self, x, g, Ac):
        w = self.lam * self._applyP(x)
        if self.p["householder"]:
            vw = self._sum([np.dot(self.vh, w)])[0]
            g[:] = (w - (2.0 * vw / self.vtv) * self.vh) + self.b
        else:
            g[:] = w + self.b
        for j in range(self.ncon):
            Ac[j][:] = self.A[j]
        return 0

    def evalHvecProduct(self, x, z, zw, px, hvec):
        """H = P diag(lam) P (constant): the callback of the inexact-Newton GMRES path."""
        w = self.lam * self._applyP(px)
        if self.p["householder"]:
            vw = self._sum([np.dot(self.vh, w)])[0]
            hvec[:] = w - (2.0 * vw / self.vtv) * self.vh
        else:
            hvec[:] = w
        return 0


class Rosenbrock:
    """examples/rosenbrock/rosenbrock.cpp:9-199 with scale = 1 (single rank)."""

    def __init__(self, nvars=999, nwcon=5, nwstart=1, nw=5, nwskip=1):
        self.nvars = nvars
        self.ncon = 2
        self.nwcon = nwcon
        self.ninequality = 2
        self.nwinequality = nwcon
        self.weighting = Weighting(nwcon, nwstart, nw, nw + nwskip, -1.0, -1.0, 1.0)

    def getVarsAndBounds(self, x, lb, ub):
        x[:] = -1.0
        lb[:] = -2.0
        ub[:] = 1.0

    def evalObjCon(self, x):
        d = x[1:] - x[:-1] ** 2
        fobj = float(np.sum((1.0 - x[:-1]) ** 2 + 100.0 * d * d))
        con = np.array([0.25 - np.sum(x * x), 10.0 + np.sum(x[::2])])
        return 0, fobj, con

    def evalObjConGradient(self, x, g, Ac):
        d = x[1:] - x[:-1] ** 2
        g[:] = 0.0
        g[:-1] += -2.0 * (1.0 - x[:-1]) + 200.0 * d * (-2.0 * x[:-1])
        g[1:] += 200.0 * d
        Ac[0][:] = -2.0 * x
        Ac[1][:] = 0.0
        Ac[1][::2] = 1.0
        return 0


class CSRWeighting:
    """The sparse-constraint callbacks of ParOptSparseProblem (ParOptProblem.cpp:734-816):
    evalSparseCon returns the values cached by the last evalObjCon (whatever x is passed),
    the products use the Jacobian values of the last gradient evaluation."""

    def __init__(self, prob):
        self.prob = prob

    def eval(self, x):
        return self.prob.cw.copy()

    def add_jac(self, alpha, px, out):
        out += alpha * (self.prob.jac() @ px)

    def add_jac_t(self, alpha, pzw, out):
        out += alpha * (self.prob.jac().T @ pzw)


class SparseMatOracle:
    """ParOptQuasiDefSparseMat (SM.cpp:231-451) with a dense Cholesky of
    K = C + A D^-1 A^T (the reference's sparse Cholesky is an exact factorisation too)."""

    def __init__(self, prob):
        self.prob = prob

    def factor(self, Dinv, C):
        import scipy.linalg as sla
        self.Dinv = Dinv
        self.A = self.prob.jac().copy()
        K = np.diag(C) + (self.A * Dinv) @ self.A.T
        try:
            self.chol = sla.cho_factor(K)
        except Exception:
            return 1
        return 0

    def apply(self, bx, bw=None):
        import scipy.linalg as sla
        rhs = -(self.A @ (self.Dinv * bx))
        if bw is not None:
            rhs = bw + rhs
        yw = sla.cho_solve(self.chol, rhs) if rhs.size else rhs
        yx = self.Dinv * (bx + self.A.T @ yw)
        return yx, yw


class SparseQuad:
    """oracle/ref_driver.cpp class SparseQuad: separable quadratic objective, dense linear
    constraints, W = (n - 1) / 3 sparse constraints cw_i = rad - sum_k w_ik (x_j - xc)^2
    over the columns 3i .. 3i+3 (+ (7i + 11) mod n for i % 5 == 0).  Single rank."""

    RAD, XC = 1.0, 0.3

    def __init__(self, comm=None, **p):
        defaults = dict(ntotal=400, ncon=1, nw=0, seed=0, lam_min=1.0, lam_max=10.0,
                        b_lo=-2.0, b_w=0.0, a_lo=0.0, a_w=-1.0, beta_c=0.0, beta_n=0.6,
                        beta_u=0.0, x0_lo0=0.3, x0_w0=0.2, lb0=-1.0, ub0=1.5)
        defaults.update(p)
        self.p = p = defaults
        n = self.nvars = int(p["ntotal"])
        W = self.nwcon = (n - 1) // 3
        self.ncon = int(p["ncon"])
        self.ninequality, self.nwinequality = self.ncon, W
        rowp, cols = [0], []
        for i in range(W):
            cols += [3 * i + k for k in range(4)]
            if i % 5 == 0:
                far = (7 * i + 11) % n
                if far < 3 * i or far > 3 * i + 3:
                    cols.append(far)
            rowp.append(len(cols))
        self.rowp = np.array(rowp, dtype=np.int32)
        self.cols = np.array(cols, dtype=np.int32)
        self.rowid = np.repeat(np.arange(W), np.diff(self.rowp))
        seed = p["seed"]
        self.wk = 0.5 + uniform01(stream_key(seed, 9), np.arange(len(cols)))
        gi = np.arange(n)
        self.lam = p["lam_min"] + (p["lam_max"] - p["lam_min"]) * uniform01(stream_key(seed, 1), gi)
        self.b = p["b_lo"] + p["b_w"] * uniform01(stream_key(seed, 2), gi)
        self.A = [p["a_lo"] + p["a_w"] * uniform01(stream_key(seed, 100 + j), gi)
                  for j in range(self.ncon)]
        self.beta = np.array([p["beta_c"] + p["beta_n"] * float(n)
                              + p["beta_u"] * float(uniform01(stream_key(seed, 5), np.uint64(j)))
                              for j in range(self.ncon)])
        self.x0 = p["x0_lo0"] + p["x0_w0"] * uniform01(stream_key(seed, 3), gi)
        self.data = np.zeros(len(cols))
        self.cw = np.zeros(W)
        self.weighting = CSRWeighting(self)

    def make_mat(self):
        return SparseMatOracle(self)

    def jac(self):
        J = np.zeros((self.nwcon, self.nvars))
        J[self.rowid, self.cols] = self.data
        return J

    def getVarsAndBounds(self, x, lb, ub):
        x[:] = self.x0
        lb[:] = self.p["lb0"]
        ub[:] = self.p["ub0"]

    def evalObjCon(self, x):
        f = float(np.sum(0.5 * self.lam * x * x + self.b * x))
        con = self.beta + np.array([np.dot(a, x) for a in self.A])
        d = x[self.cols] - self.XC
        self.cw = self.RAD - np.bincount(self.rowid, weights=self.wk * d * d,
                                         minlength=self.nwcon)
        return 0, f, con

    def evalObjConGradient(self, x, g, Ac):
        g[:] = self.lam * x + self.b
        for j in range(self.ncon):
            Ac[j][:] = self.A[j]
        self.data = -2.0 * self.wk * (x[self.cols] - self.XC)
        return 0
