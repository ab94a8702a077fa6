"""Small batches through the CUDA-graph path with a growing workspace (1 -> 5 -> 9 -> 3 vectors)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chromevol_enhanced_b200 import synthetic                      # noqa: E402
from chromevol_enhanced_b200.engine import LikelihoodEngine        # noqa: E402
from chromevol_enhanced_b200.flatten import FlatTree               # noqa: E402

tree = synthetic.random_tree(38, seed=3)
counts = synthetic.simulate_tip_counts(tree, 35, seed=4)
flat = FlatTree(tree, counts, 35)
eng = LikelihoodEngine(flat)
for B in (1, 5, 9, 3, 1):
    rows = synthetic.random_param_rows(B, seed=B)
    try:
        out = eng.log_likelihood(rows)
        torch.cuda.synchronize()
        print("B", B, "ok", out.cpu().numpy()[:2])
    except Exception as e:  # noqa: BLE001
        print("B", B, "FAILED", type(e).__name__, str(e)[:300])
        break
