"""Measure the FP64 denominators MEASURED_PEAKS.json lacks (run on the GPU box):
cuBLAS DGEMM throughput, the DMMA.8x8x4 and DFMA issue ceilings, and an HBM copy."""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from chromevol_enhanced_b200 import _lib  # noqa: E402


def timed(fn, reps=5):
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e-3)
    return best


def main():
    info = _lib.require_device()
    lib = _lib.load()
    out = {"device": torch.cuda.get_device_name(0), **info}
    for n in (4096, 8192):
        a = torch.randn(n, n, dtype=torch.float64, device="cuda")
        b = torch.randn(n, n, dtype=torch.float64, device="cuda")
        torch.matmul(a, b)
        t = timed(lambda: torch.matmul(a, b))
        out[f"dgemm_{n}_tflops"] = 2 * n ** 3 / t / 1e12
    # batched small DGEMM like our workload (n = 104)
    a = torch.randn(4096, 104, 104, dtype=torch.float64, device="cuda")
    torch.bmm(a, a)
    t = timed(lambda: torch.bmm(a, a))
    out["dgemm_batched104_tflops"] = 4096 * 2 * 104 ** 3 / t / 1e12
    del a, b
    sink = torch.zeros(8, dtype=torch.float64, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    blocks = info["sm_count"] * 4
    for kind, name, per_thread in ((0, "dmma", None), (1, "dfma", 16)):
        iters = 20000
        lib.cevb_diag_fp64_peak(kind, blocks, 100, _lib.ptr(sink), st)
        t = timed(lambda: lib.cevb_diag_fp64_peak(kind, blocks, iters, _lib.ptr(sink), st), reps=3)
        if kind == 0:
            flops = blocks * 16 * iters * 8 * (2 * 8 * 8 * 4)
        else:
            flops = blocks * 512 * iters * 16 * 2
        out[f"{name}_issue_tflops"] = flops / t / 1e12
    x = torch.empty(1 << 28, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    y.copy_(x)
    t = timed(lambda: y.copy_(x))
    out["hbm_copy_gbs"] = 2 * x.numel() * 8 / t / 1e9
    print(json.dumps(out))


if __name__ == "__main__":
    main()
