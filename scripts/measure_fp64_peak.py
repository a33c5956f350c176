"""FP64 tensor-core peak of this GPU: cuBLAS DGEMM through torch.matmul (the
denominator of the Gram kernel's compute roofline, SURVEY.md section 8d: "fp64 tensor
peak not in MEASURED_PEAKS.json -- measure cuBLAS DGEMM").  Burst = best of 10 single
launches, sustained = back to back for 3 s; CUDA events.  Prints one JSON line."""
import json
import sys
import time

import torch


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    dev = torch.device("cuda", 0)
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    c = torch.empty(n, n, dtype=torch.float64, device=dev)
    flops = 2.0 * n ** 3
    for _ in range(3):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    # sustained
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    cnt = 0
    e0.record()
    while time.time() - t0 < 3.0:
        for _ in range(5):
            torch.matmul(a, b, out=c)
        cnt += 5
        torch.cuda.synchronize()
    e1.record()
    e1.synchronize()
    sus = e0.elapsed_time(e1) / cnt
    # tall-skinny Gram shape of config C4: (121 x N) (N x 121), N = 4M rows per call
    m, rows = 121, 4 * 1024 * 1024
    v = torch.randn(rows, m, dtype=torch.float64, device=dev)
    g = torch.empty(m, m, dtype=torch.float64, device=dev)
    for _ in range(2):
        torch.matmul(v.t(), v, out=g)
    torch.cuda.synchronize()
    gb = 1e30
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(v.t(), v, out=g)
        e1.record()
        e1.synchronize()
        gb = min(gb, e0.elapsed_time(e1))
    print(json.dumps({
        "gpu": torch.cuda.get_device_name(0), "n": n,
        "fp64_dgemm_tflops_burst": flops / (best * 1e-3) / 1e12,
        "fp64_dgemm_tflops_sustained": flops / (sus * 1e-3) / 1e12,
        "cublas_gram_121x4M": {"ms": gb, "tflops_full_square": 2.0 * m * m * rows / (gb * 1e-3) / 1e12,
                               "gbs": (rows * m * 8) / (gb * 1e-3) / 1e9},
        "how": "torch.matmul float64 (cuBLAS DGEMM), CUDA events"}))


if __name__ == "__main__":
    main()
