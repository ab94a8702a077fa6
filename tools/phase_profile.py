"""Per-phase SM-clock breakdown of expm_branch_kernel on the bench workload (diagnostics)."""
import os, sys, json
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
import bench
from chromevol_enhanced_b200 import _lib, synthetic
from chromevol_enhanced_b200.engine import LikelihoodEngine
from chromevol_enhanced_b200.flatten import FlatTree
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
mode = sys.argv[2] if len(sys.argv) > 2 else "config3"
tree, counts, rows = bench.build_workload(B, 0)
if mode == "default":
    rows = np.tile(synthetic.default_param_row(), (B, 1))
flat = FlatTree(tree, counts, bench.MAX_CHR)
eng = LikelihoodEngine(flat)
lib = _lib.load()
d = torch.as_tensor(rows, device="cuda")
eng.log_likelihood(d)
buf = torch.zeros(10, dtype=torch.int64, device="cuda")
lib.cevb_diag_set_phase_buffer(_lib.ptr(buf))
eng.stage_events = []
eng.log_likelihood(d)
torch.cuda.synchronize()
lib.cevb_diag_set_phase_buffer(None)
clk = buf.cpu().numpy().astype(float)
names = ["order", "WV build", "table stage", "U sparse", "form Qm/Pm", "GJ staging(+Dinv wait)", "GJ update", "squarings", "post+store", "-"]
M = B * flat.n_t
ms = {k: a.elapsed_time(b) for k, a, b in eng.stage_events}
print(json.dumps({"matrices": M, "expm_ms": ms.get("expm"), "us_per_matrix_per_sm": ms.get("expm") * 1e3 / M * 148}))
for nm, c in zip(names, clk):
    print(f"{nm:26s} {c / M:10.0f} clk/matrix  {c / clk.sum() * 100:5.1f}%")
