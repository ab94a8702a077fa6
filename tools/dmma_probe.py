"""DMMA.8x8x4 throughput vs resident warps per SM (one CTA per SM, 13 accumulator tiles per warp):
operands from registers, from shared memory (one LDS.64 per DMMA), and register double-buffered."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from chromevol_enhanced_b200 import _lib  # noqa: E402

info = _lib.require_device()
lib = _lib.load()
sink = torch.zeros(8, dtype=torch.float64, device="cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
sms = info["sm_count"]
iters = 4000
for mode, name in ((0, "registers"), (1, "LDS per DMMA"), (2, "LDS double-buffered"),
                   (3, "13 LDS then 13 DMMA"), (4, "LDS.128 per 2 DMMA"),
                   (5, "13 volatile LDS, 13 DMMA"), (6, "volatile, 2 register sets"),
                   (100, "13 DMMA + 0 DADD"), (108, "13 DMMA + 8 DADD"), (128, "13 DMMA + 28 DADD"),
                   (152, "13 DMMA + 52 DADD"), (200, "13 DMMA + 0 LOP3 (loop)"),
                   (252, "13 DMMA + 52 LOP3"), (304, "13 DMMA + 104 LOP3"),
                   (408, "13 DMMA + 208 LOP3"), (616, "13 DMMA + 416 LOP3")):
    line = []
    for warps in (1, 2, 4, 8, 12, 16):
        lib.cevb_diag_dmma_probe(sms, warps * 32, 50, mode, _lib.ptr(sink), st)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        lib.cevb_diag_dmma_probe(sms, warps * 32, iters, mode, _lib.ptr(sink), st)
        b.record()
        torch.cuda.synchronize()
        t = a.elapsed_time(b) * 1e-3
        tf = sms * warps * iters * 13 * 512 / t / 1e12
        line.append(f"{warps}w:{tf:5.1f}")
    print(f"{name:22s}", " ".join(line), "TFLOP/s")
