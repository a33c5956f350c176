"""GPU parity of the fused / bulk-copy-staged KKT paths at sizes where they are
active (n >= 16384 rows, ragged tails, a slab straddling the end of the
weighting blocks): the optimizer history must match (a) the numpy oracle and
(b) the plain path (separate pass 1 kernels, register-fed Gram) selected with the
PCU_NO_* switches.  Tolerances: tests/parity.py (RTOL = 1e-10)."""
import os

import numpy as np
import pytest

from paropt_b200 import configs
from tests.parity import compare_histories

pytestmark = pytest.mark.gpu

SWITCHES = ("PCU_NO_GRAM_TMA", "PCU_NO_RHSGRAM", "PCU_NO_FUSE21", "PCU_NO_FUSE2S")


@pytest.fixture(scope="module")
def ctx():
    from paropt_b200.api import Context
    c = Context(0)
    yield c
    c.close()


def run(ctx, make_problem, options, iters, plain):
    from paropt_b200.api import InteriorPoint
    for k in SWITCHES:
        if plain:
            os.environ[k] = "1"
        else:
            os.environ.pop(k, None)
    # bulk-copy staged fused passes (tma_tile_kernel): forced on at these sizes with
    # a grid small enough that every CTA wraps its stage ring several times
    ctx.set_param("no_tma_tile", 1 if plain else 0)
    ctx.set_param("tma_min_tiles", 1)
    ctx.set_param("tma_grid", 7)
    try:
        prob = make_problem()
        ip = InteriorPoint(prob, dict(options, history_level=2, max_major_iters=iters))
        ip.optimize()
        hist = ip.history()
        ip.free()
        prob.free()
    finally:
        for k in SWITCHES:
            os.environ.pop(k, None)
        ctx.set_param("no_tma_tile", 0)
        ctx.set_param("tma_min_tiles", 0)
        ctx.set_param("tma_grid", 0)
    return hist


# C4 (100 dense constraints + L-SR1: 120 columns) exercises the wide tile-pair
# Gram kernel; its L-SR1 history is only reproducible for 9 iterations (see
# tests/test_oracle_golden.py).  With an even n its two pass-2 kernels run on the
# column-split staged harness (wide_tile_kernel, pcu_wide.cuh; 40008: a ragged last tile
# of 8 rows), with an odd n (40009) on the register-fed kernels.
@pytest.mark.parametrize("name,n,iters", [("C3", 8 * 5003, 14), ("C2", 50001, 14),
                                          ("C3", 8 * 4096, 14), ("C4", 40008, 9),
                                          ("C4", 625 * 64, 9), ("C4", 40009, 9)])
def test_fused_paths_match_oracle_and_plain_path(ctx, name, n, iters):
    from oracle.ip_oracle import InteriorPointOracle
    from oracle.problems import SepQuad
    from paropt_b200.api import problem_from_config
    cfg = configs.get(name, n)
    fused = run(ctx, lambda: problem_from_config(ctx, cfg), cfg["options"], iters, False)
    plain = run(ctx, lambda: problem_from_config(ctx, cfg), cfg["options"], iters, True)
    ref = InteriorPointOracle(SepQuad(**cfg["problem"]),
                              dict(cfg["options"], max_major_iters=iters))
    ref.optimize()
    for other, label in ((ref.history, "oracle"), (plain, "plain path")):
        cnt, worst, first = compare_histories(other, fused, max_iters=iters - 1, cfg=cfg)
        assert cnt == iters - 1 and first is None, (label, first, worst)


def test_widest_column_set_matches_plain_path(ctx):
    """139 dense constraints + L-SR1 m = 20 + the right-hand side = PCU_MAX_COLS columns: the wide
    Gram kernel with 20 tile rows (eight common segments per warp) and both pass-2 kernels
    on the column-split staged kernel with 159 columns (27 running dot products per lane in
    the refinement half-solve)."""
    from paropt_b200.api import problem_from_config
    cfg = configs.get("C4", 625 * 64, ncon=139)
    fused = run(ctx, lambda: problem_from_config(ctx, cfg), cfg["options"], 9, False)
    plain = run(ctx, lambda: problem_from_config(ctx, cfg), cfg["options"], 9, True)
    cnt, worst, first = compare_histories(plain, fused, max_iters=8, cfg=cfg)
    assert cnt == 8 and first is None, (first, worst)


def test_weighting_blocks_ending_inside_a_slab(ctx):
    """Blocks of 8 cover only the first 24008 of 40024 variables: the Gram slab
    that straddles their end and the ragged tail go to the general kernel."""
    from paropt_b200.api import Problem

    n, nwcon = 40024, 3001
    rng = np.random.default_rng(5)
    lam = 1.0 + 9.0 * rng.random(n)
    b = rng.random(n) - 0.3
    a = 0.5 + rng.random(n)

    class Partial(Problem):
        def __init__(self):
            super().__init__(ctx, n, 1, weighting=dict(nwcon=nwcon, wstart=0, nw=8, wstride=8,
                                                      coef0=1.0, coef_rest=-1.0, wconst=0.0))

        def getVarsAndBounds(self, x, lb, ub):
            x[:] = 0.5
            lb[:] = 0.0
            ub[:] = 1.0
            x[0:8 * nwcon:8] = 4.0
            lb[0:8 * nwcon:8] = -1e30
            ub[0:8 * nwcon:8] = 10.0

        def evalObjCon(self, x):
            return 0, float(0.5 * np.dot(lam * x, x) + np.dot(b, x)), [0.3 * n - float(np.dot(a, x))]

        def evalObjConGradient(self, x, g, A):
            g[:] = lam * x + b
            A[0][:] = -a
            return 0

    opts = dict(configs.get("C3", 8192)["options"])
    fused = run(ctx, Partial, opts, 14, False)
    plain = run(ctx, Partial, opts, 14, True)
    cnt, worst, first = compare_histories(plain, fused, max_iters=13, nvars=8192)
    assert cnt == 13 and first is None, (first, worst)


@pytest.mark.parametrize("name,n", [("C3", 8 * 5003), ("C2", 50001)])
def test_device_chain_matches_host_dense_algebra(ctx, name, n):
    """The device chain of the KKT solve (pcu_dense.cu: LU of G and Ce, SMW
    coefficients and dense residuals in a single-CTA kernel between the streaming
    passes) executes the statements of the host path, with reciprocals (<= 1 ulp)
    in place of the divisions on its critical path: with PCU_NO_CHAIN the dense algebra
    runs on the host instead, and the two histories agree to 1e-12."""
    from paropt_b200.api import problem_from_config
    cfg = configs.get(name, n)
    chain = run(ctx, lambda: problem_from_config(ctx, cfg), cfg["options"], 16, False)
    os.environ["PCU_NO_CHAIN"] = "1"
    try:
        host = run(ctx, lambda: problem_from_config(ctx, cfg), cfg["options"], 16, False)
    finally:
        os.environ.pop("PCU_NO_CHAIN", None)
    cnt, worst, first = compare_histories(host, chain, max_iters=16, cfg=cfg, rtol=1e-12)
    assert cnt == 16 and first is None, (first, worst)
    for a, b in zip(chain, host):
        assert a["info"] == b["info"] and a["neval"] == b["neval"]


@pytest.mark.parametrize("name,n", [("C3", 8 * 5003), ("C2", 50001)])
def test_residual_statistics_taken_by_the_update_passes(ctx, name, n):
    """Update1FT<1> / Update2FT<1> take the next iteration's residual statistics at the
    new point with ResF's own expressions: with PCU_NO_UPDSTATS the stand-alone residual
    pass runs instead.  Same terms, another summation order across the tiles: the two
    histories agree to 1e-12 (measured: last-bit differences in comp / fobj)."""
    from paropt_b200.api import problem_from_config
    cfg = configs.get(name, n)
    fused = run(ctx, lambda: problem_from_config(ctx, cfg), cfg["options"], 16, False)
    os.environ["PCU_NO_UPDSTATS"] = "1"
    try:
        plain = run(ctx, lambda: problem_from_config(ctx, cfg), cfg["options"], 16, False)
    finally:
        os.environ.pop("PCU_NO_UPDSTATS", None)
    cnt, worst, first = compare_histories(plain, fused, max_iters=16, cfg=cfg, rtol=1e-12)
    assert cnt == 16 and first is None, (first, worst)
    for a, b in zip(fused, plain):
        assert a["info"] == b["info"] and a["neval"] == b["neval"]
