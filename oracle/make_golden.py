"""
oracle/make_golden.py -- TEST INFRASTRUCTURE ONLY.

Generates tests/golden/*.json by running the UNMODIFIED reference (compiled by
oracle/Makefile into oracle/_ref/ref_driver) on the small variants of the named
workloads.  Run in the development container (needs /root/reference to build
oracle/_ref); the fixtures are committed so the GPU box never needs the
reference sources.

    python -m oracle.make_golden
"""
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from paropt_b200 import configs  # noqa: E402

DRIVER = os.path.join(HERE, "_ref", "ref_driver")


def driver_args(cfg):
    args = []
    if cfg["kind"] == "rosenbrock":
        args += ["problem=rosenbrock", "n=%d" % cfg["problem"]["n"]]
    else:
        args += ["problem=" + cfg["kind"]]  # sepquad | sparsequad
        for k, v in cfg["problem"].items():
            key = "n" if k == "ntotal" else k
            args.append("%s=%r" % (key, v))
    for k, v in cfg["options"].items():
        if isinstance(v, bool):
            v = int(v)
        args.append("opt:%s=%s" % (k, v))
    return args


def parse_log(path):
    """Iteration rows of the reference's text log (IP.cpp:4777-4801)."""
    rows = []
    status = None
    for line in open(path):
        parts = line.split()
        if line.startswith("ParOpt: Successfully converged to requested"):
            status = "tolerance"
        elif line.startswith("ParOpt: Successfully converged on relative"):
            status = "rel_function"
        elif line.startswith("ParOpt Warning: Current design point could not"):
            status = "no_improvement"
        if len(parts) >= 15 and parts[0].isdigit() and parts[1].isdigit():
            rows.append({"iter": int(parts[0]), "alpha": parts[4], "alpha_x": parts[5],
                         "alpha_z": parts[6], "info": " ".join(parts[15:])})
    return rows, status


def run_reference(cfg, nranks=1, extra_env=None):
    env = dict(os.environ)
    env["OPENBLAS_NUM_THREADS"] = "1"
    env["PCU_SHIM_NP"] = str(nranks)
    if extra_env:
        env.update(extra_env)
    with tempfile.TemporaryDirectory() as tmp:
        hist = os.path.join(tmp, "hist.jsonl")
        log = os.path.join(tmp, "paropt.out")
        cmd = [DRIVER] + driver_args(cfg) + ["hist=" + hist, "log=" + log]
        subprocess.run(cmd, check=True, env=env, stdout=subprocess.DEVNULL)
        recs = [json.loads(line) for line in open(hist)]
        rows, status = parse_log(log)
    its = [r for r in recs if "iter" in r]
    final = [r for r in recs if "final" in r][0]
    return {"config": cfg, "nranks": nranks, "history": its, "final": final,
            "log": rows, "status": status}


# Option coverage (barrier strategies, starting-point
# strategies, norms, line-search and quasi-Newton flavours): (label, base
# workload, option overrides, problem-size override).  Histories are capped at
# 30 iterations to keep the fixtures small.
VARIANTS = [
    ("mehrotra", "C1", dict(barrier_strategy="mehrotra"), None),
    ("mpc", "C1", dict(barrier_strategy="mehrotra_predictor_corrector"), None),
    ("compfrac", "C1", dict(barrier_strategy="complementarity_fraction"), None),
    ("lsq_start", "C1", dict(starting_point_strategy="least_squares_multipliers"), None),
    ("no_start", "C1", dict(starting_point_strategy="no_start_strategy"), None),
    ("l1norm", "C1", dict(norm_type="l1"), None),
    ("l2norm", "C1", dict(norm_type="l2"), None),
    ("backtrack", "C1", dict(use_backtracking_alpha=True), None),
    ("damped", "C1", dict(qn_update_type="damped_update"), None),
    ("yts_sts", "C1", dict(qn_diag_type="yts_over_sts"), None),
    ("mpc", "C3", dict(barrier_strategy="mehrotra_predictor_corrector"), 4000),
    ("compfrac", "C3", dict(barrier_strategy="complementarity_fraction"), 4000),
    ("lsq_start", "C3", dict(starting_point_strategy="least_squares_multipliers"), 4000),
    ("slp", "C3", dict(sequential_linear_method=True), 4000),
    ("mehrotra", "C2", dict(barrier_strategy="mehrotra"), 4000),
    ("norefine", "C2", dict(iterative_refinement_steps=0), 4000),
    ("refine2", "C3", dict(iterative_refinement_steps=2), 4000),
]


GMRES_VARIANTS = [
    ("C2_var_gmres", "C2", {}),
    ("C3_var_gmres", "C3", {}),
    ("C3_var_gmres_noprecon", "C3", dict(use_qn_gmres_precon=False)),
]


def gmres_config(base, extra):
    cfg = configs.small(base)
    cfg["problem"]["ntotal"] = 4000
    cfg["options"] = dict(cfg["options"], max_major_iters=40, use_hvec_product=True,
                          gmres_subspace_size=15, nk_switch_tol=1e3, max_gmres_rtol=0.5,
                          **extra)
    return cfg


def variant_config(base, overrides, n):
    cfg = configs.small(base)
    if n is not None:
        key = "n" if cfg["kind"] == "rosenbrock" else "ntotal"
        cfg["problem"][key] = n
    cfg["options"] = dict(cfg["options"], max_major_iters=30, **overrides)
    return cfg


def parse_tr_log(path):
    """Iteration rows of the reference's trust-region log (ParOptTrustRegion.cpp:1433-1445:
    %5d %12.5e, eleven %9.2e, time, info)."""
    keys = ["fobj", "infeas", "l1", "linfty", "dx", "tr", "rho", "model_red", "zav", "zmax",
            "gav", "gmax"]
    rows = []
    for line in open(path):
        parts = line.split()
        if len(parts) >= 14 and parts[0].isdigit():
            try:
                vals = [float(v) for v in parts[1:13]]
            except ValueError:
                continue
            row = {"iter": int(parts[0]), "info": " ".join(parts[14:])}
            row.update(dict(zip(keys, vals)))
            rows.append(row)
    return rows


def run_reference_tr(cfg):
    """The reference's ParOptOptimizer with algorithm = tr (sl1qp penalty method, the
    configuration of examples/rosenbrock/rosenbrock.cpp:234-242) on a named workload:
    centre points at full precision from the writeOutput hook + the rows of its log."""
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1", PCU_SHIM_NP="1")
    with tempfile.TemporaryDirectory() as tmp:
        hist = os.path.join(tmp, "hist.jsonl")
        trlog = os.path.join(tmp, "paropt.tr")
        cmd = [DRIVER] + driver_args(cfg) + ["algorithm=tr", "hist=" + hist, "tr_log=" + trlog,
                                             "log=" + os.path.join(tmp, "paropt.out")]
        subprocess.run(cmd, check=True, env=env, stdout=subprocess.DEVNULL)
        recs = [json.loads(line) for line in open(hist)]
        rows = parse_tr_log(trlog)
    return {"config": cfg, "algorithm": "tr", "centres": [r for r in recs if "tr_iter" in r],
            "final": [r for r in recs if "final" in r][0], "log": rows}


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    if "--tr" in sys.argv:
        # SURVEY.md section 8f-1 (next row): histories of the reference's trust-region
        # front end, for the oracle / CUDA port of ParOptTrustRegion to be pinned against
        jobs = [("C1_tr", configs.small("C1"), dict(barrier_strategy="mehrotra")),
                ("C2_tr_small", configs.get("C2", 4000), dict(barrier_strategy="mehrotra"))]
        for fname, cfg, extra in jobs:
            cfg["options"] = dict(cfg["options"], **extra)
            data = run_reference_tr(cfg)
            data["generator"] = "oracle/make_golden.py --tr (oracle/_ref/ref_driver algorithm=tr, unmodified reference)"
            with open(os.path.join(out_dir, fname + ".json"), "w") as fp:
                json.dump(data, fp, separators=(",", ":"))
            print(fname, "tr iterations", len(data["log"]), "centres", len(data["centres"]))
        return
    if "--variants" in sys.argv:
        for label, base, overrides, n in VARIANTS:
            data = run_reference(variant_config(base, overrides, n))
            data["generator"] = "oracle/make_golden.py --variants (oracle/_ref/ref_driver, unmodified reference)"
            fname = "%s_var_%s.json" % (base, label)
            with open(os.path.join(out_dir, fname), "w") as fp:
                json.dump(data, fp, separators=(",", ":"))
            print(fname, "niter", data["final"]["niter"], data["status"])
        return
    if "--gmres" in sys.argv:
        # the inexact-Newton path (computeKKTGMRESStep, IP.cpp:5789-6191)
        # with the problem's exact Hessian-vector products, switched on early (nk_switch_tol)
        # so that most iterations of the history take it
        for fname, base, extra in GMRES_VARIANTS:
            cfg = gmres_config(base, extra)
            data = run_reference(cfg)
            data["generator"] = ("oracle/make_golden.py --gmres (oracle/_ref/ref_driver, unmodified "
                                 "reference, use_hvec_product)")
            with open(os.path.join(out_dir, fname + ".json"), "w") as fp:
                json.dump(data, fp, separators=(",", ":"))
            print(fname, "niter", data["final"]["niter"], data["status"],
                  "nhvec", data["history"][-1]["nhvec"])
        return
    if "--sparse" in sys.argv:
        # a ParOptSparseProblem (general CSR sparse constraints,
        # ParOptQuasiDefSparseMat + the reference's sparse Cholesky)
        data = run_reference(configs.small("S1"))
        data["generator"] = ("oracle/make_golden.py --sparse (oracle/_ref/ref_driver, unmodified "
                             "reference, ParOptSparseProblem / ParOptQuasiDefSparseMat)")
        with open(os.path.join(out_dir, "S1_small.json"), "w") as fp:
            json.dump(data, fp, separators=(",", ":"))
        print("S1_small.json", "niter", data["final"]["niter"], data["status"])
        return
    if "--full-c4" in sys.argv:
        # the dense-constraint stress config at its FULL size (n = 32M, c = 100, L-SR1
        # m = 20): the first iterations of the unmodified reference, 8 shim ranks;
        # ~50 GB of host memory and ~2.5 min per iteration (c (c + 1) / 2 = 5050 dots
        # and c + q sequential solves per set-up)
        cfg = configs.get("C4")
        cfg["options"] = dict(cfg["options"], max_major_iters=11)
        data = run_reference(cfg, 8)
        data["generator"] = "oracle/make_golden.py --full-c4 (oracle/_ref/ref_driver, unmodified reference, 8 ranks)"
        data["log"] = data["log"][:12]
        with open(os.path.join(out_dir, "C4_full.json"), "w") as fp:
            json.dump(data, fp, separators=(",", ":"))
        print("C4_full.json", "niter", data["final"]["niter"], data["status"], flush=True)
        return
    if "--checkpoint" in sys.argv:
        # the reference's binary checkpoint (writeSolutionFile, IP.cpp:883-975) of the final
        # state of a small C3 run: the byte layout the CUDA reader / writer must honour
        cfg = configs.get("C3", 256)
        cfg["options"] = dict(cfg["options"], max_major_iters=20)
        env = dict(os.environ, OPENBLAS_NUM_THREADS="1", PCU_SHIM_NP="1")
        with tempfile.TemporaryDirectory() as tmp:
            hist = os.path.join(tmp, "hist.jsonl")
            out = os.path.join(out_dir, "C3_ckpt.bin")
            cmd = [DRIVER] + driver_args(cfg) + ["hist=" + hist, "log=/dev/null", "checkpoint=" + out]
            subprocess.run(cmd, check=True, env=env, stdout=subprocess.DEVNULL)
            recs = [json.loads(line) for line in open(hist)]
        data = {"config": cfg, "history": [r for r in recs if "iter" in r],
                "final": [r for r in recs if "final" in r][0],
                "generator": "oracle/make_golden.py --checkpoint (unmodified reference, writeSolutionFile)"}
        with open(os.path.join(out_dir, "C3_ckpt.json"), "w") as fp:
            json.dump(data, fp, separators=(",", ":"))
        print("C3_ckpt.bin", os.path.getsize(out), "bytes; niter", data["final"]["niter"])
        return
    if "--nb2" in sys.argv:
        # ParOptQuasiDefBlockMat with nwblock = 2 (dense 2 x 2 blocks of Ew, dpptrf /
        # dpptrs): C3-style workload whose blocks carry two sparse constraints
        cfg = configs.get("C3", 4000)
        cfg["problem"]["nb"] = 2
        cfg["options"] = dict(cfg["options"], max_major_iters=40)
        data = run_reference(cfg)
        data["generator"] = ("oracle/make_golden.py --nb2 (oracle/_ref/ref_driver, unmodified "
                             "reference, ParOptQuasiDefBlockMat nwblock = 2)")
        with open(os.path.join(out_dir, "C3_nb2_small.json"), "w") as fp:
            json.dump(data, fp, separators=(",", ":"))
        print("C3_nb2_small.json", "niter", data["final"]["niter"], data["status"])
        return
    if "--full-c4-np4" in sys.argv:
        # evidence for the stated C4_full tolerance: the SAME unmodified reference on 4
        # instead of 8 ranks (another summation partition of the 5050 Gram dot
        # products), first 6 iterations
        cfg = configs.get("C4")
        cfg["options"] = dict(cfg["options"], max_major_iters=6)
        data = run_reference(cfg, 4)
        data["generator"] = "oracle/make_golden.py --full-c4-np4 (oracle/_ref/ref_driver, unmodified reference, 4 ranks)"
        data["log"] = data["log"][:7]
        with open(os.path.join(out_dir, "C4_full_np4.json"), "w") as fp:
            json.dump(data, fp, separators=(",", ":"))
        print("C4_full_np4.json", "niter", data["final"]["niter"], data["status"], flush=True)
        return
    if "--full" in sys.argv:
        # the first iterations of the reference at the FULL sizes of BASELINE.json
        # (C3: n = 64M, W = 8M; C2: n = 16M, c = 10), 8 shim ranks; ~40 GB, minutes
        for name in ("C2", "C3"):
            cfg = configs.get(name)
            cfg["options"] = dict(cfg["options"], max_major_iters=13)
            data = run_reference(cfg, 8)
            data["generator"] = "oracle/make_golden.py --full (oracle/_ref/ref_driver, unmodified reference, 8 ranks)"
            data["log"] = data["log"][:14]
            fname = "%s_full.json" % name
            with open(os.path.join(out_dir, fname), "w") as fp:
                json.dump(data, fp, separators=(",", ":"))
            print(fname, "niter", data["final"]["niter"], data["status"], flush=True)
        return
    jobs = [("C1", 1), ("C2", 1), ("C3", 1), ("C4", 1), ("C2", 2), ("C3", 2), ("C4", 2)]
    for name, nranks in jobs:
        cfg = configs.small(name)
        data = run_reference(cfg, nranks)
        data["generator"] = "oracle/make_golden.py (oracle/_ref/ref_driver, unmodified reference)"
        fname = "%s_small%s.json" % (name, "" if nranks == 1 else "_np%d" % nranks)
        with open(os.path.join(out_dir, fname), "w") as fp:
            json.dump(data, fp, indent=0)
        print(fname, "niter", data["final"]["niter"], data["status"])


if __name__ == "__main__":
    main()
