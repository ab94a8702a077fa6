"""GPU: the coefficient ring of the apply kernels streams an item (row) in PARTS when two whole
items do not fit (narrow CTA shapes, long series).  Forcing tiny rings through the development
knobs must not change a single bit of the result: the accumulators add the chunks in the same
order either way."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))

SCRIPT = r"""
import json, sys
import numpy as np, torch
sys.path.insert(0, %r)
from chromevol_enhanced_b200 import synthetic
from chromevol_enhanced_b200.engine import LikelihoodEngine
from chromevol_enhanced_b200.flatten import FlatTree
out = {}
for name, taxa, N, B in (("n100", 300, 100, 24), ("n40", 120, 40, 8), ("n256", 60, 256, 4)):
    tree = synthetic.random_tree(taxa, seed=3)
    counts = synthetic.simulate_tip_counts_fast(tree, N, seed=5, root_state=min(20, N // 2))
    flat = FlatTree(tree, counts, N)
    rows = synthetic.random_param_rows(B, seed=7, linear='positive')
    rows[0, :5] = [6.0, 5.0, 3.0, 2.0, 1.0]          # long series on every branch
    rows[1, 7] = -0.004                              # no generator: FP64 leaf kernel
    eng = LikelihoodEngine(flat, n_streams=1)
    ll = eng.log_likelihood(torch.as_tensor(rows, device="cuda")).cpu().numpy()
    out[name] = [float(x).hex() for x in ll]
print("RESULT" + json.dumps(out))
""" % ROOT


def _run(env_extra):
    env = dict(os.environ)
    for k in ("CEVB_RING_S", "CEVB_RING_M", "CEVB_RING_W", "CEVB_RING_L"):
        env.pop(k, None)
    env.update(env_extra)
    proc = subprocess.run([sys.executable, "-c", SCRIPT], env=env, capture_output=True, text=True,
                          timeout=600)
    assert proc.returncode == 0, proc.stderr[-2000:]
    line = [ln for ln in proc.stdout.splitlines() if ln.startswith("RESULT")][-1]
    return json.loads(line[len("RESULT"):])


def test_streaming_an_item_in_parts_is_bit_identical():
    base = _run({})
    for knobs in ({"CEVB_RING_S": "2", "CEVB_RING_M": "2", "CEVB_RING_W": "2", "CEVB_RING_L": "2"},
                  {"CEVB_RING_S": "5", "CEVB_RING_M": "7", "CEVB_RING_W": "9", "CEVB_RING_L": "3"}):
        got = _run(knobs)
        assert got == base, knobs
    vals = np.array([float.fromhex(x) for x in base["n100"]])
    assert np.all(np.isfinite(vals)) and np.all(vals < 0)
