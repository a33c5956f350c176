import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def repo_root():
    return ROOT


def pytest_terminal_summary(terminalreporter, exitstatus, config):
    """Achieved parity errors of every history comparison of the session
    (tests/parity.py: units where <= 1e-10 passes; `rel:` = pure relative error of
    the residual norms above their round-off floor)."""
    from tests import parity

    if not parity.REPORTS:
        return
    tr = terminalreporter
    tr.section("parity: worst error per history comparison (tolerance 1e-10)")
    for label, rec in sorted(parity.REPORTS.items()):
        w = parity.checked(rec["worst"])
        top = max(w.values()) if w else 0.0
        key = max(w, key=w.get) if w else "-"
        w = rec["worst"]
        rel = max([v for k, v in w.items() if k.startswith("rel:")] + [0.0])
        tr.write_line("%-46s iters %3d  worst %.2e (%s)  residual norms rel %.2e"
                      % (label, rec["compared"], top, key, rel))
    kind = "gpu" if "gpu" in (config.getoption("-m") or "") and "not gpu" not in (
        config.getoption("-m") or "") else "cpu"
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_worst_%s.json" % kind), "w") as fp:
            json.dump(parity.REPORTS, fp, indent=1, sort_keys=True)
    except OSError:
        pass
