"""Wall-clock of the drop-in classes on the shipped-size problem (38 taxa, N=35): one objective
call, one optimisation, the mapper -- what a user of the reference's CLI experiences."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chromevol_enhanced_b200 as cev

meta = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "golden.json")))
tree = cev.Tree(meta["shipped_newick"], format=1)
counts = dict(meta["shipped_counts"])
model = cev.ChromEvolutionModel(35)
opt = cev.ChromEvolOptimizer(tree, counts, model)
opt.calculator.set_root_frequencies(np.ones(36) / 36)
names = opt.DEFAULT_PARAM_NAMES
x0 = np.array([model.params[p] for p in names])
opt.objective_function(x0, names)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    opt.objective_function(x0, names)
t1 = time.perf_counter()
print(f"objective_function: {(t1 - t0) / 50 * 1e3:.3f} ms per call")
for batched in (False, True):
    model.set_parameters(**cev.ChromEvolutionModel(35).params)
    opt.batched_gradient = batched
    t0 = time.perf_counter()
    res = opt.optimize_parameters()
    t1 = time.perf_counter()
    print(f"optimize_parameters(batched_gradient={batched}): {t1 - t0:.2f} s, nfev={res.nfev}, "
          f"logL={-res.fun:.6f}")
opt.calculator.ancestral_state_reconstruction()
t0 = time.perf_counter()
m = cev.StochasticMapper(tree, model, num_mappings=100)
m.perform_stochastic_mapping()
print(f"stochastic mapping (100 replicates): {time.perf_counter() - t0:.3f} s")
