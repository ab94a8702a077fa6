"""Step time of the bench workload against chunk size and stream count: small chunks keep the
series coefficients of a chunk (about 1.8 MB read per vector and level) in the 126 MB L2 across
the level launches.  usage: python tools/chunk_probe.py [B]"""
import os
import sys
import time

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


def timed(eng, reps=4):
    eng.want_node_vectors = False
    for _ in range(2):
        eng.log_likelihood(rows)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = eng.log_likelihood(rows)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3, out


for chunk in (512, 128, 64, 48, 32, 16):
    for streams in (1, 2, 4):
        eng = LikelihoodEngine(flat, n_streams=streams, chunk=chunk)
        ms, out = timed(eng)
        print(f"chunk={chunk} streams={streams}: {ms:.2f} ms", flush=True)
        del eng
