"""One warm-up and one measured evaluation of the bench batch on ONE stream (launch order = leaf
list, then level by level), for `ncu -k regex:taylor_apply --launch-skip 20 --launch-count 20`.
usage: python tools/step_once.py [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chromevol_enhanced_b200 import synthetic                      # noqa: E402
from chromevol_enhanced_b200.engine import LikelihoodEngine        # noqa: E402
from chromevol_enhanced_b200.flatten import FlatTree               # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
tree = synthetic.random_tree(1000, seed=1)
counts = synthetic.simulate_tip_counts(tree, 100, seed=2)
flat = FlatTree(tree, counts, 100)
rows = torch.as_tensor(synthetic.random_param_rows(B, seed=4, linear='positive'), device="cuda")
eng = LikelihoodEngine(flat, n_streams=1)
eng.want_node_vectors = False
for _ in range(2):
    out = eng.log_likelihood(rows)
torch.cuda.synchronize()
print("launches per chunk", eng.launches_per_chunk(), "logL[0]", float(out[0]))
