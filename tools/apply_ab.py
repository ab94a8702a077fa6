"""Apply-stage time of one 512-vector step under the current environment knobs (CEVB_*), one
stream, stage events; prints one line.  usage: python tools/apply_ab.py [label]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chromevol_enhanced_b200 import synthetic                      # noqa: E402
from chromevol_enhanced_b200.engine import LikelihoodEngine        # noqa: E402
from chromevol_enhanced_b200.flatten import FlatTree               # noqa: E402

B = 512
tree = synthetic.random_tree(1000, seed=1)
counts = synthetic.simulate_tip_counts(tree, 100, seed=2)
flat = FlatTree(tree, counts, 100)
rows = torch.as_tensor(synthetic.random_param_rows(B, seed=4, linear='positive'), device="cuda")
eng = LikelihoodEngine(flat, n_streams=1)
eng.want_node_vectors = False
for _ in range(3):
    eng.log_likelihood(rows)
torch.cuda.synchronize()
tot = {}
first_apply = []
per_launch = None
R = 5
for _ in range(R):
    eng.stage_events = []
    eng.log_likelihood(rows)
    torch.cuda.synchronize()
    seen_apply = 0
    cur = []
    for name, e0, e1 in eng.stage_events:
        ms = e0.elapsed_time(e1)
        tot[name] = tot.get(name, 0.0) + ms
        if name == "leaf":
            cur.append(ms)
            first_apply.append(ms)
        if name == "apply":
            cur.append(ms)
    per_launch = cur if per_launch is None else [a + b for a, b in zip(per_launch, cur)]
eng.stage_events = None
if os.environ.get("AB_LEVELS"):
    import numpy as np
    cnt = [int(flat.leaf_child.size)] + np.diff(flat.level_child_off).tolist()
    print("apply per launch (branches: ms): " + "  ".join(f"{c}:{t / R:.3f}" for c, t in zip(cnt, per_launch)))
NS = int(os.environ.get("AB_STREAMS", "4"))
# whole step on 4 streams (the bench configuration)
eng4 = LikelihoodEngine(flat, n_streams=NS)
eng4.want_node_vectors = False
for _ in range(3):
    eng4.log_likelihood(rows)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(R):
    eng4.log_likelihood(rows)
e1.record()
torch.cuda.synchronize()
label = sys.argv[1] if len(sys.argv) > 1 else ""
knobs = {k: v for k, v in os.environ.items() if k.startswith("CEVB_")}
print(f"{label:14s} leaf {sum(first_apply) / R:7.3f} ms  apply {(tot['apply'] + tot.get('leaf', 0.0)) / R:7.3f}  prepare {tot['prepare'] / R:6.3f}  "
      f"expm {tot['expm'] / R:6.3f}  combine {tot['combine'] / R:6.3f}  step({NS} streams) {e0.elapsed_time(e1) / R:7.3f} ms  {knobs}")
