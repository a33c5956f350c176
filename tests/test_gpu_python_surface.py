"""The Python drop-in surface (`from paropt_b200 import ParOpt`) used
the way examples/random_quadratic/random_quadratic.py uses the reference: a
Problem subclass with numpy callbacks, ParOpt.Optimizer with the option dict,
the text log parsed by unpack_output; results against the numpy oracle on the
same problem (tests/parity.py tolerance)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_optimizer_facade_matches_oracle(tmp_path):
    from oracle.ip_oracle import InteriorPointOracle
    from oracle.problems import SepQuad
    from paropt_b200 import ParOpt, configs
    from tests.parity import compare_histories

    cfg = configs.get("C2", 3000)
    ref_prob = SepQuad(**cfg["problem"])
    n, ncon = ref_prob.nvars, ref_prob.ncon

    class Quadratic(ParOpt.Problem):
        def __init__(self):
            super().__init__(None, nvars=n, ncon=ncon)

        def getVarsAndBounds(self, x, lb, ub):
            ref_prob.getVarsAndBounds(x, lb, ub)

        def evalObjCon(self, x):
            return ref_prob.evalObjCon(x)

        def evalObjConGradient(self, x, g, A):
            return ref_prob.evalObjConGradient(x, g, A)

    log = str(tmp_path / "paropt.out")
    opts = dict(cfg["options"], algorithm="ip", max_major_iters=25, output_file=log,
                history_level=2)
    prob = Quadratic()
    opt = ParOpt.Optimizer(prob, opts)
    opt.optimize()
    x, z, zw, zl, zu = opt.getOptimizedPoint()
    ora = InteriorPointOracle(ref_prob, dict(cfg["options"], max_major_iters=25))
    ora.optimize()
    cnt, worst, first = compare_histories(ora.history, opt.ip.history(), max_iters=24, cfg=cfg)
    assert cnt == 24 and first is None, (first, worst)
    assert np.allclose(np.asarray(x), ora.variables.x, rtol=1e-9, atol=1e-12)
    assert np.allclose(z, ora.variables.z, rtol=1e-8, atol=1e-12)
    names, cols = ParOpt.unpack_output(log)
    assert names[0] == "iter" and list(cols[0][:5]) == [0, 1, 2, 3, 4]
    hist = opt.ip.history()
    k = min(len(cols[7]), len(hist)) - 1
    assert abs(cols[7][k] - hist[k]["fobj"]) <= 1e-5 * max(1.0, abs(hist[k]["fobj"]))
    with pytest.raises(ValueError):
        ParOpt.Optimizer(prob, {"algorithm": "mma"})


def test_pvec_array_protocols():
    import torch
    from paropt_b200 import ParOpt
    ctx = ParOpt.Context(0)
    v = ParOpt.PVec(ctx, 1000)
    v.set(2.0)
    v[3] = 7.0
    v[10:20] = np.arange(10.0)
    assert v[3] == 7.0 and v[12] == 2.0 and float(np.asarray(v).sum()) == 2.0 * 989 + 7.0 + 45.0
    t = torch.as_tensor(v, device="cuda")  # zero-copy view through __cuda_array_interface__
    assert t.data_ptr() == v.device_ptr() and t.dtype == torch.float64
    t.mul_(0.5)
    torch.cuda.synchronize()
    assert v[3] == 3.5 and abs(v.l1norm() - 0.5 * (2.0 * 989 + 7.0 + 45.0)) < 1e-9
    v.free()
