"""Named workloads (BASELINE.json `configs`) as plain data.

Each entry gives the synthetic-problem parameters (DESIGN.md "Synthetic
problems") and the ParOptInteriorPoint options (names as in the reference,
ParOptInteriorPoint.cpp:536-727).  `small(...)` returns the same workload at a
size the CPU oracle finishes in seconds; the full sizes are the BASELINE.json
ones.  Pure data: no oracle, no CUDA.
"""
import copy

_C3_PROBLEM = dict(
    ncon=1, nw=8, seed=0, lam_min=1.0, lam_max=10.0, b_lo=-1.0, b_w=0.0,
    a_lo=0.0, a_w=-1.0, beta_c=0.0, beta_n=0.2, beta_u=0.0,
    x0_lo0=0.5, x0_w0=0.4, x0_lo1=0.02, x0_w1=0.1,
    lb0=-1e30, ub0=1.0, lb1=0.0, ub1=1.0, householder=0,
)

CONFIGS = {
    # configs[0]: examples/rosenbrock, IP + L-BFGS, n=1000, 1 rank
    "C1": dict(
        kind="rosenbrock", problem=dict(n=1000),
        options=dict(qn_type="bfgs", qn_subspace_size=10, abs_res_tol=1e-6),
    ),
    # configs[1]: random convex QP, ncon=10, n=16M, L-BFGS m=10
    # (options of examples/random_quadratic/random_quadratic.py:92-102)
    "C2": dict(
        kind="sepquad",
        problem=dict(ntotal=16 * 1024 * 1024, ncon=10, nw=0, seed=0,
                     lam_min=1.0, lam_max=10.0, householder=1),
        options=dict(qn_type="bfgs", qn_subspace_size=10, abs_res_tol=1e-8,
                     start_affine_multiplier_min=0.01, penalty_gamma=1000.0,
                     starting_point_strategy="affine_step",
                     barrier_strategy="monotone"),
    ),
    # configs[2]: multi-material topology-style, n=64M, nwcon=8M, 1 dense con
    "C3": dict(
        kind="sepquad",
        problem=dict(ntotal=64 * 1024 * 1024, **_C3_PROBLEM),
        options=dict(qn_type="bfgs", qn_subspace_size=10),
    ),
    # configs[3]: dense-constraint stress, n=32M, ncon=100, L-SR1 m=20
    "C4": dict(
        kind="sepquad",
        problem=dict(ntotal=32 * 1024 * 1024, ncon=100, nw=0, seed=0,
                     lam_min=1.0, lam_max=10.0, householder=1),
        options=dict(qn_type="sr1", qn_subspace_size=20, abs_res_tol=1e-8,
                     start_affine_multiplier_min=0.01),
    ),
    # general sparse constraints (a ParOptSparseProblem: CSR
    # Jacobian, ParOptQuasiDefSparseMat); oracle/ref_driver.cpp class SparseQuad
    "S1": dict(
        kind="sparsequad",
        problem=dict(ntotal=400, ncon=1, seed=0, lam_min=1.0, lam_max=10.0, b_lo=-2.0,
                     b_w=0.0, a_lo=0.0, a_w=-1.0, beta_c=0.0, beta_n=0.3, beta_u=0.0,
                     x0_lo0=0.3, x0_w0=0.2, lb0=-1.0, ub0=1.5),
        options=dict(qn_type="bfgs", qn_subspace_size=10),
    ),
    # configs[4]: weak scaling, C2 with n = 16M per GPU
    "C5": dict(
        kind="sepquad",
        problem=dict(ntotal=16 * 1024 * 1024, ncon=10, nw=0, seed=0,
                     lam_min=1.0, lam_max=10.0, householder=1),
        options=dict(qn_type="bfgs", qn_subspace_size=10, abs_res_tol=1e-8,
                     start_affine_multiplier_min=0.01),
        per_gpu=True,
    ),
}


def get(name, ntotal=None, **problem_overrides):
    cfg = copy.deepcopy(CONFIGS[name])
    if ntotal is not None:
        key = "n" if cfg["kind"] == "rosenbrock" else "ntotal"
        cfg["problem"][key] = int(ntotal)
    cfg["problem"].update(problem_overrides)
    return cfg


def small(name):
    sizes = {"C1": 1000, "C2": 20000, "C3": 16000, "C4": 12000, "C5": 20000, "S1": 400}
    cfg = get(name, sizes[name])
    if name == "C4":
        cfg["problem"]["ncon"] = 24
    return cfg
