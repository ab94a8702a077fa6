"""How long would the non-leaf branches take if tree levels imposed no launch boundaries?
All 998 internal-child branches of the bench tree in ONE general apply launch (results are not
used: the child vectors of upper levels are stale), next to the real per-level sequence."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chromevol_enhanced_b200 import _lib, synthetic
from chromevol_enhanced_b200.engine import LikelihoodEngine
from chromevol_enhanced_b200.flatten import FlatTree

B = 512
tree = synthetic.random_tree(1000, seed=1)
counts = synthetic.simulate_tip_counts(tree, 100, seed=2)
flat = FlatTree(tree, counts, 100)
rows = torch.as_tensor(synthetic.random_param_rows(B, seed=4), device="cuda")
eng = LikelihoodEngine(flat, max_workspace_bytes=48 << 30, n_streams=1)
eng.want_node_vectors = False
eng.log_likelihood(rows)
torch.cuda.synchronize()
lib, ws = _lib.load(), eng._ws
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
items = torch.as_tensor(np.sort(flat.level_child), dtype=torch.int32, device="cuda")
order = np.argsort(-flat.dist_eff[flat.level_child], kind="stable")
items_sorted = torch.as_tensor(flat.level_child[order], dtype=torch.int32, device="cuda")
t_node = eng.d_t_node


def run(it):
    lib.cevb_taylor_apply(_lib.ptr(ws.tc), _lib.ptr(ws.norms), B, eng.n, flat.n_nodes, _lib.ptr(it),
                          int(it.numel()), 0, _lib.ptr(t_node), _lib.ptr(eng.d_leaf_code),
                          _lib.ptr(eng._node_vec), _lib.ptr(ws.w), _lib.ptr(ws.wr), _lib.ptr(ws.mode), st)


for name, it in (("unsorted", items), ("longest first", items_sorted)):
    run(it)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        run(it)
    b.record()
    torch.cuda.synchronize()
    print(f"all {it.numel()} non-leaf branches in one launch ({name}): {a.elapsed_time(b) / 3:.2f} ms")
