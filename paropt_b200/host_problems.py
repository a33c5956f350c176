"""Host-callback problems: user code on the far side of the drop-in boundary.

`HostSepQuad` states the separable/Householder QP workload (DESIGN.md "Synthetic
problems") as an ordinary ParOpt-style problem class with numpy callbacks, the
way a user of the reference writes `ParOpt.Problem` subclasses
(examples/random_quadratic/random_quadratic.py:10-57).  Every callback receives
host arrays; the iterate travels device->host and the gradients host->device on
every call, which is what the end-to-end benchmark measures.
"""
import numpy as np

from .api import Problem


def _splitmix64(z):
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _key(seed, stream):
    with np.errstate(over="ignore"):
        return _splitmix64(np.uint64(seed) ^ (np.uint64(stream) * np.uint64(0x9E3779B97F4A7C15)))


def _uniform(key, idx):
    with np.errstate(over="ignore"):
        h = _splitmix64(np.uint64(key) + idx)
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


class HostSepQuad(Problem):
    def __init__(self, ctx, allreduce=None, **p):
        defaults = dict(ntotal=1000, ncon=1, nw=0, seed=0, lam_min=1.0, lam_max=1e3,
                        b_lo=0.0, b_w=1.0, a_lo=0.0, a_w=1.0, beta_c=0.0, beta_n=0.0,
                        beta_u=1.0, x0_lo0=-2.0, x0_lo1=-2.0, x0_w0=1.0, x0_w1=1.0,
                        lb0=-5.0, lb1=-5.0, ub0=5.0, ub1=5.0, householder=0)
        defaults.update(p)
        self.p = p = defaults
        self.allreduce = allreduce or (lambda a: a)
        rank, size = ctx.rank, ctx.size
        unit = p["nw"] if p["nw"] > 0 else 1
        nunits = p["ntotal"] // unit
        u0, u1 = (nunits * rank) // size, (nunits * (rank + 1)) // size
        offset = u0 * unit
        n = (u1 - u0) * unit
        if rank == size - 1:
            n = p["ntotal"] - offset
        nwcon = (u1 - u0) if p["nw"] > 0 else 0
        gi = np.arange(offset, offset + n, dtype=np.uint64)
        seed = p["seed"]
        self.lam = p["lam_min"] + (p["lam_max"] - p["lam_min"]) * _uniform(_key(seed, 1), gi)
        self.b = p["b_lo"] + p["b_w"] * _uniform(_key(seed, 2), gi)
        self.hh = bool(p["householder"])
        if self.hh:
            self.vh = 0.5 + _uniform(_key(seed, 7), gi)
            self.vtv = float(self.allreduce(np.array([np.dot(self.vh, self.vh)]))[0])
        self.A = [p["a_lo"] + p["a_w"] * _uniform(_key(seed, 100 + j), gi)
                  for j in range(p["ncon"])]
        self.beta = np.array([p["beta_c"] + p["beta_n"] * float(p["ntotal"])
                              + p["beta_u"] * float(_uniform(_key(seed, 5), np.uint64(j)))
                              for j in range(p["ncon"])])
        u = _uniform(_key(seed, 3), gi)
        if p["nw"] > 0:
            c1 = (np.arange(n) % p["nw"]) != 0
        else:
            c1 = np.zeros(n, dtype=bool)
        self._x0 = np.where(c1, p["x0_lo1"] + p["x0_w1"] * u, p["x0_lo0"] + p["x0_w0"] * u)
        self._lb = np.where(c1, p["lb1"], p["lb0"])
        self._ub = np.where(c1, p["ub1"], p["ub0"])
        del gi, u
        self._tmp = np.empty(n)
        weighting = None
        if nwcon > 0:
            weighting = dict(nwcon=nwcon, wstart=0, nw=p["nw"], wstride=p["nw"],
                             coef0=1.0, coef_rest=-1.0, wconst=0.0)
        super().__init__(ctx, n, p["ncon"], weighting=weighting)
        self.nwcon = nwcon

    def transfer_bytes(self):
        return self.h2d_bytes, self.d2h_bytes

    def getVarsAndBounds(self, x, lb, ub):
        x[:] = self._x0
        lb[:] = self._lb
        ub[:] = self._ub

    # The O(n) arithmetic runs on torch CPU tensors sharing memory with the numpy
    # arrays (threaded over the host cores; plain numpy elementwise ops are
    # single-threaded and ~5x slower at n = 64M).
    @staticmethod
    def _t(a):
        import torch
        return torch.from_numpy(a)

    def _tensors(self):
        if not hasattr(self, "_tt"):
            self._tt = dict(lam=self._t(self.lam), b=self._t(self.b), tmp=self._t(self._tmp),
                            A=[self._t(a) for a in self.A],
                            vh=self._t(self.vh) if self.hh else None)
        return self._tt

    def evalObjCon(self, x):
        import torch
        t = self._tensors()
        xt = self._t(x)
        tmp = t["tmp"]
        if self.hh:
            vx = float(self.allreduce(np.array([float(torch.dot(t["vh"], xt))]))[0])
            y = torch.add(xt, t["vh"], alpha=-(2.0 * vx / self.vtv))
            torch.mul(t["lam"], y, out=tmp)
            loc = [0.5 * float(torch.dot(tmp, y)) + float(torch.dot(t["b"], xt))]
        else:
            # f = sum x (lam x / 2 + b)
            torch.addcmul(t["b"], t["lam"], xt, value=0.5, out=tmp)
            loc = [float(torch.dot(tmp, xt))]
        loc += [float(torch.dot(a, xt)) for a in t["A"]]
        out = self.allreduce(np.array(loc))
        return 0, float(out[0]), self.beta + out[1:]

    def evalObjConGradient(self, x, g, A):
        import torch
        t = self._tensors()
        xt, gt = self._t(x), self._t(g)
        if self.hh:
            vx = float(self.allreduce(np.array([float(torch.dot(t["vh"], xt))]))[0])
            y = torch.add(xt, t["vh"], alpha=-(2.0 * vx / self.vtv))
            torch.mul(t["lam"], y, out=gt)
            vw = float(self.allreduce(np.array([float(torch.dot(t["vh"], gt))]))[0])
            gt.add_(t["vh"], alpha=-(2.0 * vw / self.vtv))
            gt.add_(t["b"])
        else:
            torch.addcmul(t["b"], t["lam"], xt, out=gt)
        for j in range(self.ncon):
            self._t(A[j]).copy_(t["A"][j])
        return 0

    def evalHvecProduct(self, x, z, zw, px, hvec):
        """H = P diag(lam) P (constant): for the option use_hvec_product."""
        import torch
        t = self._tensors()
        pt, ht = self._t(px), self._t(hvec)
        if self.hh:
            vx = float(self.allreduce(np.array([float(torch.dot(t["vh"], pt))]))[0])
            y = torch.add(pt, t["vh"], alpha=-(2.0 * vx / self.vtv))
            torch.mul(t["lam"], y, out=ht)
            vw = float(self.allreduce(np.array([float(torch.dot(t["vh"], ht))]))[0])
            ht.add_(t["vh"], alpha=-(2.0 * vw / self.vtv))
        else:
            torch.mul(t["lam"], pt, out=ht)
        return 0
