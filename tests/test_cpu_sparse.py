"""CPU tests of the host (symbolic) phase of pcu_sparsemat -- the device counterpart of
ParOptQuasiDefSparseMat (ParOptSparseMat.cpp:231-451).  The handle is created without a
context (no device call); the numeric phase is EMULATED here in numpy on the symbolic
arrays exactly as the kernels of pcu_sparse.cu walk them (assembly at kpos, left-looking
column Cholesky in level order, row-oriented forward / column-oriented backward
substitution) and compared with a dense solve of K = C + A D^-1 A^T."""
import ctypes as C

import numpy as np
import pytest

from paropt_b200 import _lib


def random_csr(rng, nw, nv, kind):
    rows = []
    for i in range(nw):
        if kind == "chain":      # neighbours share a variable: K is tridiagonal
            cols = [2 * i, 2 * i + 1, 2 * i + 2]
        elif kind == "arrow":    # every row touches variable 0: K is dense
            cols = [0, 1 + i]
        elif kind == "blocks":   # disjoint rows: K is diagonal
            cols = [3 * i, 3 * i + 1, 3 * i + 2]
        else:                    # random rows, some empty, unsorted
            k = int(rng.integers(0, 5))
            cols = list(rng.choice(nv, size=k, replace=False))
        cols = [c for c in cols if c < nv]
        rng.shuffle(cols)
        rows.append(cols)
    rowp = np.zeros(nw + 1, dtype=np.int32)
    for i, r in enumerate(rows):
        rowp[i + 1] = rowp[i] + len(r)
    cols = np.array([c for r in rows for c in r], dtype=np.int32)
    data = rng.uniform(0.5, 1.5, size=cols.size) * rng.choice([-1.0, 1.0], size=cols.size)
    return rowp, cols, data


def symbolic(lib, nv, nw, rowp, cols, ordering):
    ip = lambda a: a.ctypes.data_as(_lib.c_int_p)
    h = lib.pcu_sparsemat_create(None, nv, nw, ip(rowp), ip(cols if cols.size else np.zeros(1, np.int32)),
                                 ordering)
    assert h
    nk, nl, nlev, nlaunch = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    assert lib.pcu_sparsemat_info(h, C.byref(nk), C.byref(nl), C.byref(nlev), C.byref(nlaunch)) == 0
    nk, nl, nlev = nk.value, nl.value, nlev.value
    z = lambda n: np.zeros(max(n, 1), dtype=np.int32)
    s = dict(perm=z(nw), Lp=z(nw + 1), Li=z(nl), Rp=z(nw + 1), Rk=z(nl - nw), Rpos=z(nl - nw),
             kpos=z(nk), ka=z(nk), kb=z(nk), level_ptr=z(nlev + 1), level_cols=z(nw))
    order = ("perm", "Lp", "Li", "Rp", "Rk", "Rpos", "kpos", "ka", "kb", "level_ptr", "level_cols")
    assert lib.pcu_sparsemat_symbolic(h, *[ip(s[k]) for k in order]) == 0
    lib.pcu_sparsemat_destroy(h)
    s.update(nk=nk, nl=nl, nlev=nlev, nlaunch=nlaunch.value)
    return s


def emulate_solve(s, K, b, nw):
    """The numeric phase as the kernels run it, on the symbolic arrays."""
    Lp, Li, Rp, Rk, Rpos = s["Lp"], s["Li"], s["Rp"], s["Rk"], s["Rpos"]
    Lx = np.zeros(max(s["nl"], 1))
    for e in range(s["nk"]):  # sp_assemble_kernel
        Lx[s["kpos"][e]] = K[s["ka"][e], s["kb"][e]]
    done = np.zeros(nw, dtype=bool)
    for l in range(s["nlev"]):  # sp_chol_kernel, one level after the other
        for idx in range(s["level_ptr"][l], s["level_ptr"][l + 1]):
            j = s["level_cols"][idx]
            c0, c1 = Lp[j], Lp[j + 1]
            assert Li[c0] == j
            for r in range(Rp[j], Rp[j + 1]):
                k, pos = Rk[r], Rpos[r]
                assert done[k] and Li[pos] == j  # a finished column of a lower level
                ljk = Lx[pos]
                for t in range(pos, Lp[k + 1]):
                    at = c0 + np.searchsorted(Li[c0:c1], Li[t])
                    assert at < c1 and Li[at] == Li[t]  # fill stays inside the pattern
                    Lx[at] -= Lx[t] * ljk
            assert Lx[c0] > 0.0
            sq = np.sqrt(Lx[c0])
            Lx[c0] = sq
            Lx[c0 + 1:c1] /= sq
        for idx in range(s["level_ptr"][l], s["level_ptr"][l + 1]):
            done[s["level_cols"][idx]] = True
    y = b[s["perm"][:nw]].copy()
    for l in range(s["nlev"]):  # sp_forward_kernel
        for idx in range(s["level_ptr"][l], s["level_ptr"][l + 1]):
            j = s["level_cols"][idx]
            acc = sum(Lx[Rpos[r]] * y[Rk[r]] for r in range(Rp[j], Rp[j + 1]))
            y[j] = (y[j] - acc) / Lx[Lp[j]]
    for l in range(s["nlev"] - 1, -1, -1):  # sp_backward_kernel
        for idx in range(s["level_ptr"][l], s["level_ptr"][l + 1]):
            j = s["level_cols"][idx]
            acc = sum(Lx[t] * y[Li[t]] for t in range(Lp[j] + 1, Lp[j + 1]))
            y[j] = (y[j] - acc) / Lx[Lp[j]]
    out = np.zeros(nw)
    out[s["perm"][:nw]] = y
    return out


@pytest.mark.parametrize("kind,nw", [("chain", 40), ("arrow", 25), ("blocks", 30), ("random", 60),
                                     ("random", 1)])
@pytest.mark.parametrize("ordering", [0, 1])
def test_symbolic_phase_supports_an_exact_solve(kind, nw, ordering):
    lib = _lib.load()
    rng = np.random.default_rng(7 + nw + ordering)
    nv = 3 * nw + 3
    rowp, cols, data = random_csr(rng, nw, nv, kind)
    s = symbolic(lib, nv, nw, rowp, cols, ordering)
    assert sorted(s["perm"][:nw].tolist()) == list(range(nw))
    A = np.zeros((nw, nv))
    for i in range(nw):
        for e in range(rowp[i], rowp[i + 1]):
            A[i, cols[e]] = data[e]
    Dinv = rng.uniform(0.2, 2.0, nv)
    Cd = rng.uniform(0.1, 1.0, nw)
    K = np.diag(Cd) + (A * Dinv) @ A.T
    # the declared pattern of K covers every non-zero of K
    patt = np.zeros((nw, nw), dtype=bool)
    for e in range(s["nk"]):
        patt[s["ka"][e], s["kb"][e]] = patt[s["kb"][e], s["ka"][e]] = True
    assert not np.any((np.abs(K) > 0) & ~patt)
    b = rng.standard_normal(nw)
    got = emulate_solve(s, K, b, nw)
    ref = np.linalg.solve(K, b)
    assert np.max(np.abs(got - ref)) <= 1e-11 * max(1.0, np.max(np.abs(ref)))
    if kind == "blocks":
        assert s["nl"] == nw and s["nlev"] == 1  # diagonal K: no fill, one level
    if kind == "chain" and ordering == 0:
        assert s["nl"] == 2 * nw - 1 and s["nlaunch"] == 1  # a path: one serial launch


def test_bad_patterns_are_rejected():
    lib = _lib.load()
    ip = lambda a: a.ctypes.data_as(_lib.c_int_p)
    rowp = np.array([0, 2], dtype=np.int32)
    assert not lib.pcu_sparsemat_create(None, 4, 1, ip(rowp), ip(np.array([1, 7], dtype=np.int32)), 1)
    assert not lib.pcu_sparsemat_create(None, 4, 1, ip(rowp), ip(np.array([2, 2], dtype=np.int32)), 1)


def test_degenerate_sizes():
    """No sparse constraints at all, and constraints with empty rows (K = C, diagonal)."""
    lib = _lib.load()
    s0 = symbolic(lib, 10, 0, np.zeros(1, dtype=np.int32), np.zeros(0, dtype=np.int32), 1)
    assert (s0["nk"], s0["nl"], s0["nlev"]) == (0, 0, 0)
    s1 = symbolic(lib, 10, 5, np.zeros(6, dtype=np.int32), np.zeros(0, dtype=np.int32), 1)
    assert (s1["nk"], s1["nl"], s1["nlev"], s1["nlaunch"]) == (5, 5, 1, 1)
    K = np.diag(np.arange(1.0, 6.0))
    b = np.ones(5)
    assert np.allclose(emulate_solve(s1, K, b, 5), b / np.arange(1.0, 6.0), rtol=1e-15)
