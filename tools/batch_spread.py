"""Step time of the bench workload on ONE GPU for the parameter batches the ranks 0..7 of a
multi-GPU run draw (bench.py: seed PARAM_SEED + 1000 * rank): how much of the gap to perfect weak
scaling is simply the slowest of N different random batches.  usage: python tools/batch_spread.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                        # noqa: E402
from chromevol_enhanced_b200.engine import LikelihoodEngine        # noqa: E402
from chromevol_enhanced_b200.flatten import FlatTree               # noqa: E402

eng = None
times = []
for rank in range(8):
    tree, counts, rows = bench.build_workload(512, rank)
    if eng is None:
        eng = LikelihoodEngine(FlatTree(tree, counts, bench.MAX_CHR))
        eng.want_node_vectors = False
    d = torch.as_tensor(rows, device="cuda")
    for _ in range(3):
        eng.log_likelihood_deferred(d)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        eng.log_likelihood_deferred(d)
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1) / 5)
print("ms per 512-vector step for the batches of ranks 0..7:", " ".join(f"{t:.2f}" for t in times))
for n in (1, 2, 4, 8):
    print(f"  N={n}: slowest of the first {n} batches {max(times[:n]):.2f} ms -> "
          f"weak-scaling efficiency bound {times[0] / max(times[:n]):.3f}")
