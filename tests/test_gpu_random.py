"""GPU: seeded random problems (tree shape, polytomies, zero / negative / missing branch lengths,
missing and out-of-range counts, base numbers, negative linear rates, small and odd N) through the
drop-in engine against the CPU oracle: logL to 1e-9 relative, rescale counts and MAP states exact."""
import numpy as np
import pytest

from oracle import chromevol_oracle as O

pytestmark = pytest.mark.gpu


def _random_problem(seed):
    from chromevol_enhanced_b200 import synthetic
    rng = np.random.default_rng(seed)
    N = int(rng.choice([3, 6, 9, 17, 33, 40, 64, 103]))
    n_leaves = int(rng.integers(3, 60))
    tree = synthetic.random_tree(n_leaves, seed=seed + 1, mean_len=float(rng.choice([0.02, 0.2, 1.0])))
    nodes = [nd for nd in tree.traverse() if not nd.is_root()]
    # polytomies: dissolve a few internal nodes into their parents
    for nd in [x for x in nodes if x.children and rng.random() < 0.15]:
        parent = nd.up
        k = parent.children.index(nd)
        parent.children[k:k + 1] = nd.children
        for ch in nd.children:
            ch.up = parent
    nodes = [nd for nd in tree.traverse() if not nd.is_root()]
    for nd in nodes:                       # special branch lengths (AR:224-227, AR:143-148)
        u = rng.random()
        if u < 0.05:
            nd.dist = 0.0
        elif u < 0.08:
            nd.dist = -0.3
        elif u < 0.10:
            nd.dist = None
    hi = max(2, min(N, 12))
    counts = {}
    for leaf in tree.get_leaves():
        u = rng.random()
        if u < 0.08:
            continue                       # species missing from the table -> uniform vector
        counts[leaf.name] = int(N + 5) if u < 0.12 else int(rng.integers(1, hi + 1))
    rows = synthetic.random_param_rows(5, seed=seed + 2, linear='mixed' if seed % 2 else 'positive',
                                       corners=1 if seed % 3 == 0 else 0)
    if seed % 4 == 0:
        rows[:, 8] = int(rng.integers(1, max(2, N // 2)))
    if seed % 5 == 0:
        rows[0, :5] *= 30.0                # forces the scaling-and-squaring path
    return tree, counts, N, rows


@pytest.mark.parametrize("seed", range(100, 124))
def test_random_problem_matches_oracle(seed):
    from chromevol_enhanced_b200 import synthetic
    from chromevol_enhanced_b200.engine import LikelihoodEngine
    from chromevol_enhanced_b200.flatten import FlatTree
    tree, counts, N, rows = _random_problem(seed)
    flat = FlatTree(tree, counts, N)
    eng = LikelihoodEngine(flat)
    freq = None
    if seed % 3 == 1:
        freq = np.zeros(N + 1)
        freq[min(N, 2)] = 1.0
    ll = eng.log_likelihood(rows, root_freq=freq).cpu().numpy()
    for b in range(rows.shape[0]):
        ref, vectors, resc = O.log_likelihood(tree, counts, N, synthetic.row_to_dict(rows[b]),
                                              root_freq=freq)
        if np.isfinite(ref):
            assert ll[b] == pytest.approx(ref, rel=1e-9, abs=1e-9), (seed, b, N)
        else:
            assert ll[b] == ref or (np.isnan(ll[b]) and np.isnan(ref)), (seed, b)
        assert int(eng.rescale_counts[b]) == resc
    # MAP states of the last vector (one more single-vector call: graph replay path)
    b = rows.shape[0] - 1
    one = eng.log_likelihood(rows[b:b + 1], root_freq=freq)
    assert float(one[0]) == pytest.approx(float(ll[b]), rel=1e-12, abs=1e-12) or not np.isfinite(ll[b])
    ref, vectors, _ = O.log_likelihood(tree, counts, N, synthetic.row_to_dict(rows[b]), root_freq=freq)
    got = eng.ancestral_states(0)[0].cpu().numpy()
    want = np.array([O.ancestral_states(vectors[id(nd)])[0] for nd in flat.nodes])
    if np.isfinite(ref):
        # argmax ties are broken by rounding only where two states are equally likely to ~1e-12
        differ = np.flatnonzero(got != want)
        for k in differ:
            v = vectors[id(flat.nodes[k])]
            p = v / v.sum() if v.sum() > 0 else np.full(N + 1, 1.0 / (N + 1))
            assert abs(p[got[k]] - p[want[k]]) <= 1e-9 * max(p.max(), 1e-300), (seed, int(k))


@pytest.mark.parametrize("scale", [1e-3, 3e-2, 1.0])
def test_floor_driven_likelihoods_through_the_tf32_leaf_launch(scale):
    """Improbable data: tip counts scattered over 1..100 under slow rates, so most of every leaf
    branch's transition column sits on the 1e-12 floor and the row-sum correction of the leaf
    launch (TF32 tensor cores for generator vectors) matters as much as it ever
    can.  logL and the per-leaf-branch messages against the oracle; the FP64 leaf kernel
    (CEVB_LOWP is a library-wide knob, so: a vector whose Q is no generator) gives the same."""
    from chromevol_enhanced_b200 import synthetic
    from chromevol_enhanced_b200.engine import LikelihoodEngine
    from chromevol_enhanced_b200.flatten import FlatTree
    rng = np.random.default_rng(5)
    N = 100
    tree = synthetic.random_tree(150, seed=41, mean_len=0.05)
    counts = {leaf.name: int(rng.integers(1, N + 1)) for leaf in tree.get_leaves()}
    rows = synthetic.random_param_rows(6, seed=43, linear='positive')
    rows[:, :7] *= scale
    rows[5, 7] = -1e-4                      # tiny negative linear rate: FP64 leaf kernel, same physics
    flat = FlatTree(tree, counts, N)
    eng = LikelihoodEngine(flat)
    ll = eng.log_likelihood(rows).cpu().numpy()
    for b in range(rows.shape[0]):
        ref, vectors, resc = O.log_likelihood(tree, counts, N, synthetic.row_to_dict(rows[b]))
        assert np.isfinite(ref)
        assert ll[b] == pytest.approx(ref, rel=1e-9), (scale, b, ll[b], ref)
        assert int(eng.rescale_counts[b]) == resc
    # node vectors of one generator vector, entry by entry relative to each vector's largest entry
    eng.log_likelihood(rows[2:3])
    got = eng.node_vectors(0).cpu().numpy()
    _, vectors, _ = O.log_likelihood(tree, counts, N, synthetic.row_to_dict(rows[2]))
    for k, nd in enumerate(flat.nodes):
        v = vectors[id(nd)]
        assert np.abs(got[k] - v).max() <= 1e-9 * v.max(), (scale, k)
