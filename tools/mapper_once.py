"""1e6 stochastic-mapping replicates of the bench tree (BASELINE configs[4]) on one GPU, once warm
and once measured, for `ncu --set full -k regex:stoch_map`.  usage: python tools/mapper_once.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chromevol_enhanced_b200 import mapper, synthetic              # noqa: E402
from chromevol_enhanced_b200.flatten import FlatTree               # noqa: E402

tree = synthetic.random_tree(1000, seed=1)
flat = FlatTree(tree, {}, 100)
root = np.zeros(101)
root[20] = 1.0
row = synthetic.default_param_row()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
mapper.simulate(flat, row, 101, root, reps, 7, rep_begin=0, rep_end=4096, return_device=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
hist, sums, n_events = mapper.simulate(flat, row, 101, root, reps, 7, rep_begin=0, rep_end=reps,
                                       return_device=True)
e1.record()
torch.cuda.synchronize()
print(f"{reps} replicates in {e0.elapsed_time(e1):.2f} ms, {int(n_events.item())} events")
