"""Time of the per-vector preparation (cevb_build_q + cevb_taylor_prepare) for the bench batch, alone
on one stream -- for the stub builds of the coefficient kernel (tools/stub_build.py 16 32 64 ...,
selected with CEVB_LIB).  Nothing downstream runs, so the wrong tables of a stub build are never
read.  usage: CEVB_LIB=... python tools/prepare_probe.py [label]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chromevol_enhanced_b200 import _lib, synthetic                # noqa: E402
from chromevol_enhanced_b200.engine import ExpmWorkspace           # noqa: E402
from chromevol_enhanced_b200.flatten import FlatTree               # noqa: E402

B, N = 512, 100
tree = synthetic.random_tree(1000, seed=1)
flat = FlatTree(tree, {}, N)
rows = synthetic.random_param_rows(B, seed=4, linear='positive')
lib = _lib.load()
ws = ExpmWorkspace(N + 1, flat.n_t, B, torch.device("cuda"), p_slots=4, n_nodes=flat.n_nodes)
ws.params.copy_(torch.as_tensor(rows, device="cuda"))
t_max = float(flat.t_unique.max())
st = torch.cuda.current_stream().cuda_stream


def run():
    _lib.check(lib.cevb_build_q(_lib.ptr(ws.params), B, N + 1, _lib.ptr(ws.pw), st), "build_q")
    _lib.check(lib.cevb_taylor_prepare(_lib.ptr(ws.pw), B, N + 1, t_max, _lib.ptr(ws.norms), _lib.ptr(ws.csc),
                                       _lib.ptr(ws.csc_cnt), _lib.ptr(ws.tc), st), "taylor_prepare")


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    run()
e1.record()
torch.cuda.synchronize()
print(f"{sys.argv[1] if len(sys.argv) > 1 else '':10s} prepare {e0.elapsed_time(e1) / 10:.3f} ms per 512 vectors")
