"""Per-launch timing of the fused path on the bench workload (CUDA events, no profiler).
   python tools/level_probe.py [batch]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chromevol_enhanced_b200 import synthetic
from chromevol_enhanced_b200.engine import LikelihoodEngine
from chromevol_enhanced_b200.flatten import FlatTree

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
tree = synthetic.random_tree(1000, seed=1)
counts = synthetic.simulate_tip_counts(tree, 100, seed=2)
flat = FlatTree(tree, counts, 100)
rows = torch.as_tensor(synthetic.random_param_rows(B, seed=4), device="cuda")
NS = int(sys.argv[2]) if len(sys.argv) > 2 else 4
eng = LikelihoodEngine(flat, max_workspace_bytes=48 << 30, n_streams=NS)
for _ in range(2):
    eng.log_likelihood(rows)
torch.cuda.synchronize()
reps = 3
eng.stage_events = []
for _ in range(reps):
    eng.log_likelihood(rows)
torch.cuda.synchronize()
ev = eng.stage_events
eng.stage_events = None
per = len(ev) // reps
acc = {}
order = []
for r in range(reps):
    idx = {}
    for name, a, b in ev[r * per:(r + 1) * per]:
        k = idx.get(name, 0)
        idx[name] = k + 1
        key = (name, k)
        if key not in acc:
            acc[key] = 0.0
            order.append(key)
        acc[key] += a.elapsed_time(b) / reps
sizes = [int(flat.leaf_child.size)] + np.diff(flat.level_child_off).tolist()
tot = {}
apply_ms = [0.0] * len(sizes)          # a batch larger than one chunk repeats the launch sequence
combine_ms = [0.0] * flat.n_levels
for key in order:
    tot[key[0]] = tot.get(key[0], 0.0) + acc[key]
    if key[0] == "apply":
        apply_ms[key[1] % len(sizes)] += acc[key]
    elif key[0] == "combine":
        combine_ms[key[1] % flat.n_levels] += acc[key]
print("apply per launch (items:ms):", " ".join(f"{c}:{t:.2f}" for c, t in zip(sizes, apply_ms)))
print("combine per level (nodes:ms):", " ".join(f"{c}:{t:.3f}" for c, t in zip(
    np.diff(flat.level_off).tolist(), combine_ms)))
print("totals ms:", {k: round(v, 2) for k, v in tot.items()})
# un-instrumented wall time
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    eng.log_likelihood(rows)
e1.record()
torch.cuda.synchronize()
print("step ms:", round(e0.elapsed_time(e1) / reps, 2), "evals/s:", round(B * reps / e0.elapsed_time(e1) * 1e3, 1))
