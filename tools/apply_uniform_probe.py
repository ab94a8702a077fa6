"""Intrinsic efficiency of taylor_apply_kernel: every branch has the same length (same number of
series chunks), leaf-only variant, one launch.  Prints executed DMMA TFLOP/s."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chromevol_enhanced_b200 import _lib, accounting, synthetic
from chromevol_enhanced_b200.engine import ExpmWorkspace

lib = _lib.load()
_lib.require_device()
dev = torch.device("cuda:0")
N, n, B = 100, 101, 256
rows = np.tile(synthetic.default_param_row(), (B, 1))
rows[:, :5] = [1.0, 0.8, 0.4, 0.2, 0.1]
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
ws = ExpmWorkspace(n, 1, B, dev, p_slots=1, n_nodes=1)
ws.params.copy_(torch.as_tensor(rows, device=dev))
lib.cevb_build_q(_lib.ptr(ws.params), B, n, _lib.ptr(ws.pw), st)
lib.cevb_taylor_prepare(_lib.ptr(ws.pw), B, n, 0.0, _lib.ptr(ws.norms), _lib.ptr(ws.csc), _lib.ptr(ws.csc_cnt),
                        _lib.ptr(ws.tc), st)
torch.cuda.synchronize()
nu, mu = float(ws.norms[0, 12]), float(ws.norms[0, 11])
for n_items in (960, 120, 24, 8):
    for x in (0.05, 0.6, 2.0, 5.0):
        t = x / nu
        nc = int(accounting.taylor_ncoef(np.array([x]), np.array([mu * t]))[0])
        t_node = torch.full((n_items,), t, dtype=torch.float64, device=dev)
        code = torch.as_tensor(np.arange(n_items) % n, dtype=torch.int32, device=dev)
        items = torch.arange(n_items, dtype=torch.int32, device=dev)
        node_vec = torch.zeros((B, n_items, ws.ldp), dtype=torch.float64, device=dev)
        w = torch.zeros_like(node_vec)
        wr = torch.ones_like(node_vec)
        mode = torch.zeros((B, n_items), dtype=torch.uint8, device=dev)
        for leaf in (1, 0):
            def run():
                lib.cevb_taylor_apply(_lib.ptr(ws.tc), _lib.ptr(ws.norms), B, n, n_items, _lib.ptr(items),
                                      n_items, leaf, _lib.ptr(t_node), _lib.ptr(code), _lib.ptr(node_vec),
                                      _lib.ptr(w), _lib.ptr(wr), _lib.ptr(mode), st)
            run()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3):
                run()
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 3
            kc = (nc + 3) // 4
            flops = B * ((n_items + 7) // 8) * n * 13 * kc * 512.0
            print(f"items={n_items:4d} x={x:4.2f} terms={nc:2d} chunks={kc:2d} leaf={leaf} {ms:8.3f} ms "
                  f"{flops / ms / 1e9:6.2f} TFLOP/s executed")
