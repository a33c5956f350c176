"""CPU tests: the numpy restatement (oracle/ip_oracle.py) against histories of
the unmodified reference (tests/golden/*.json, made by oracle/make_golden.py)."""
import numpy as np
import pytest

from oracle.ip_oracle import InteriorPointOracle
from oracle.problems import Rosenbrock, SepQuad, splitmix64, uniform01, stream_key
from tests.parity import RTOL, compare_histories, load_golden


def build_oracle(cfg, comm=None):
    if cfg["kind"] == "rosenbrock":
        prob = Rosenbrock(cfg["problem"]["n"] - 1)
    else:
        prob = SepQuad(comm=comm, **cfg["problem"])
    return InteriorPointOracle(prob, cfg["options"], comm=comm)


def test_generator_known_answers():
    # splitmix64 reference outputs for seed 0 (first outputs of the published
    # generator started from state 0: each call hashes state + k*golden)
    assert int(splitmix64(np.uint64(0))) == 0xE220A8397B1DCDAF
    u = uniform01(stream_key(0, 1), np.arange(4))
    assert np.all((u >= 0.0) & (u < 1.0))
    # 53-bit mantissa: u * 2^53 is an integer
    assert np.all(np.floor(u * 2.0 ** 53) == u * 2.0 ** 53)


# full-length parity: identical iteration count / status / counters, 1e-10 state
@pytest.mark.parametrize("name", ["C1_small", "C2_small"])
def test_oracle_matches_reference_full_history(name):
    gold = load_golden(name)
    ip = build_oracle(gold["config"])
    ip.optimize()
    n, worst, first = compare_histories(gold["history"], ip.history)
    assert first is None, (first, worst)
    assert ip.niter == gold["final"]["niter"]
    assert ip.neval == gold["final"]["neval"]
    assert ip.ngeval == gold["final"]["ngeval"]
    assert ip.converged == gold["status"]
    assert n == len(gold["history"])
    # log tags (skipH, cmpEq, LNoImprv ...) of every iteration
    for row, mine in zip(gold["log"], ip.log):
        assert row["info"] == mine["info"], (row, mine)
        if row["alpha_x"] != "--":
            assert abs(float(row["alpha_x"]) - mine["alpha_x"]) <= 0.06 * mine["alpha_x"]
            assert abs(float(row["alpha_z"]) - mine["alpha_z"]) <= 0.06 * mine["alpha_z"]


# weighting-constraint and L-SR1 workloads: the tail of these runs is decided by
# |merit change| <= function_precision tests that the reference itself does not
# reproduce across BLAS thread counts, so parity is required on the first 50
# iterations (all barrier updates down to mu ~ 1e-5 included).  The reference's
# L-SR1 (ParOptQuasiNewton.cpp:636-747) has no update-skipping safeguard: its M
# matrix turns ill-conditioned and round-off grows ~10x per iteration from
# iteration 8 on (measured: 1e-15 -> 1e-9 by iteration 11, O(1) by 21), so two
# correct fp64 implementations only share the first 9 iterations at 1e-10.
@pytest.mark.parametrize("name,iters", [("C3_small", 50), ("C4_small", 9)])
def test_oracle_matches_reference_prefix(name, iters):
    gold = load_golden(name)
    ip = build_oracle(gold["config"])
    ip.opt["max_major_iters"] = iters + 1
    ip.optimize()
    n, worst, first = compare_histories(gold["history"], ip.history, max_iters=iters)
    assert n == iters
    assert first is None, (first, worst)


def test_reference_two_rank_history_matches_single_rank():
    """The reference partitioned over 2 shim ranks follows the 1-rank history."""
    for name in ("C2_small", "C3_small"):
        g1 = load_golden(name)
        g2 = load_golden(name + "_np2")
        n, worst, first = compare_histories(g1["history"], g2["history"], max_iters=30)
        assert first is None, (name, first)
