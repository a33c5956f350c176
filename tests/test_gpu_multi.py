"""Multi-GPU parity (needs >= 2 GPUs, skipped otherwise): the design vector is
partitioned over the ranks exactly as the reference partitions it over MPI ranks
and the history must still follow the reference's."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_history_matches_reference():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29600 + (os.getpid() % 1000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "scripts", "mgpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("MGPU_VERDICT ")][-1]
    verdict = json.loads(line[len("MGPU_VERDICT "):])
    assert verdict["world"] == 2
    for name, res in verdict["cases"].items():
        assert res["first_violation"] is None, (name, res)
