"""The reference's binary checkpoint (ParOptInteriorPoint::writeSolutionFile /
readSolutionFile, IP.cpp:883-1104): tests/golden/C3_ckpt.bin was written by the UNMODIFIED
reference at the end of a 20-iteration C3 run (n = 256; `make_golden --checkpoint`).
CPU: the layout.  GPU: pcu_ip_read_solution takes it in, pcu_ip_write_solution gives the
same bytes back; a CUDA run of the same workload writes the same file up to round-off."""
import os
import struct

import numpy as np
import pytest

from tests.parity import load_golden

HERE = os.path.dirname(os.path.abspath(__file__))
CKPT = os.path.join(HERE, "golden", "C3_ckpt.bin")


def parse(path):
    raw = open(path, "rb").read()
    n, w, c = struct.unpack("3i", raw[:12])
    off = 12
    vals = np.frombuffer(raw, dtype=np.float64, offset=off, count=1 + 5 * c)
    out = {"n": n, "w": w, "c": c, "mu": vals[0]}
    for k, name in enumerate(("s", "t", "z", "zs", "zt")):
        out[name] = vals[1 + k * c: 1 + (k + 1) * c]
    off += 8 * (1 + 5 * c)
    for name, cnt in (("x", n), ("zl", n), ("zu", n), ("zw", w), ("sw", w)):
        out[name] = np.frombuffer(raw, dtype=np.float64, offset=off, count=cnt)
        off += 8 * cnt
    assert off == len(raw)
    return out


def test_reference_checkpoint_layout():
    gold = load_golden("C3_ckpt")
    ck = parse(CKPT)
    prob = gold["config"]["problem"]
    assert (ck["n"], ck["w"], ck["c"]) == (prob["ntotal"], prob["ntotal"] // prob["nw"], prob["ncon"])
    last = gold["history"][-1]  # the state of the last writeOutput hook is one step behind
    assert abs(ck["mu"] - gold["final"]["mu"]) <= 1e-15 * abs(ck["mu"])
    assert np.all(ck["zl"] >= 0.0) and np.all(ck["zu"] >= 0.0) and np.all(ck["sw"] > 0.0)
    assert abs(np.sum(ck["x"]) - last["xsum"]) <= 0.2 * abs(last["xsum"])


@pytest.mark.gpu
def test_checkpoint_round_trip_and_parity(tmp_path):
    from paropt_b200.api import Context, InteriorPoint, problem_from_config
    gold = load_golden("C3_ckpt")
    cfg = gold["config"]
    ctx = Context(0)
    prob = problem_from_config(ctx, cfg)
    # (a) read the reference's file, write it back: the same bytes
    ip = InteriorPoint(prob, cfg["options"])
    assert ip.readSolutionFile(CKPT) == 0
    ck = parse(CKPT)
    x = ip.getOptimizedPoint()[0].to_numpy()
    assert np.array_equal(x, ck["x"]) and ip.getBarrierParameter() == ck["mu"]
    assert np.array_equal(ip.get_dense()["z"], ck["z"])
    out = str(tmp_path / "back.bin")
    assert ip.writeSolutionFile(out) == 0
    assert open(out, "rb").read() == open(CKPT, "rb").read()
    ip.free()
    # (b) the same run on the CUDA path, checkpoint written by the optimizer itself
    mine = str(tmp_path / "mine.bin")
    ip = InteriorPoint(prob, dict(cfg["options"], ip_checkpoint_file=mine, write_output_frequency=1))
    ip.optimize()
    assert ip.writeSolutionFile(mine) == 0  # final state, as the reference driver does
    a, b = parse(mine), ck
    assert (a["n"], a["w"], a["c"]) == (b["n"], b["w"], b["c"])
    for key in ("mu", "s", "t", "z", "zs", "zt", "x", "zl", "zu", "zw", "sw"):
        va, vb = np.atleast_1d(a[key]), np.atleast_1d(b[key])
        scale = max(float(np.max(np.abs(vb))), 1e-300)
        assert float(np.max(np.abs(va - vb))) <= 1e-9 * scale, key
    # (c) a file of another size is refused with the reference's message
    other = problem_from_config(ctx, dict(cfg, problem=dict(cfg["problem"], ntotal=512)))
    ip2 = InteriorPoint(other, cfg["options"])
    assert ip2.readSolutionFile(CKPT) == 1
    for o in (ip2, other, ip, prob):
        o.free()
    ctx.close()
