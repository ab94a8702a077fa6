"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY, NOT A PRODUCT PATH.

A plain numpy/scipy restatement of the reference's chromosome-number likelihood path
(`/root/reference/src/ancestral_reconstruction.py`, abbreviated AR below).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may import
this module, and only as the checker / the timed CPU baseline.  The product package
(`chromevol_enhanced_b200/`) never imports it and has no CPU fallback.

Third-party arithmetic the reference delegates to and that is NOT under /root/reference:
  * scipy (requirements.txt:4 `scipy>=1.7.0`, unpinned; this image has 1.18.1):
    `scipy.linalg.expm` at AR:151.  The oracle calls the same function, exactly like the
    reference does; `oracle/expm_amh.py` additionally restates its published algorithm
    (Al-Mohy & Higham 2009) so the Pade order m and squaring count s can be pinned.
  * numpy (requirements.txt:3 `numpy>=1.21.0`; image has 2.3.5): `np.dot` AR:237, `np.sum`
    AR:131/241/268/312, `np.argmax` AR:315, `np.random.*` AR:489/510/545.

Parity pinning: the reference ships no tests, so the oracle is pinned
(i) LIVE against the unmodified reference module imported in the build container
(tests/test_oracle_vs_reference.py, skipped where /root/reference is absent) and (ii) against
golden vectors generated from that module by tests/golden/make_golden.py and committed under
tests/golden/ (these travel to the GPU box).

Every function cites the reference lines it follows.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.linalg

# AR:22
MODEL_PARAMETER_NAMES = ['gain', 'loss', 'dupl', 'demidupl_gain', 'demidupl_loss',
                         'base_number_gain', 'base_number_loss', 'linear_rate']

# AR:38-48
DEFAULT_PARAMS = {
    'gain': 0.1, 'loss': 0.1, 'dupl': 0.05, 'demidupl_gain': 0.01, 'demidupl_loss': 0.01,
    'base_number': None, 'base_number_gain': 0.01, 'base_number_loss': 0.01, 'linear_rate': 0.01,
}

# AR:602-604
EVENT_TYPES = ['gain', 'loss', 'duplication', 'base_number_gain', 'base_number_loss',
               'demi_gain_like', 'demi_loss_like', 'other']


def default_params(**overrides):
    p = dict(DEFAULT_PARAMS)
    p.update(overrides)
    return p


# --------------------------------------------------------------------------- a2
def build_q(params, max_chr):
    """Rate matrix Q over states 0..max_chr -- AR:75-134, same loop, same operation order."""
    base_number = params.get('base_number')
    Q = np.zeros((max_chr + 1, max_chr + 1))
    for i in range(1, max_chr + 1):                                   # AR:86
        if i < max_chr:                                               # AR:88-90
            Q[i, i + 1] = params['gain'] * (1 + params['linear_rate'] * i)
        if i > 1:                                                     # AR:93-95
            Q[i, i - 1] = params['loss'] * (1 + params['linear_rate'] * i)
        if 2 * i <= max_chr and i > 0:                                # AR:98-99
            Q[i, 2 * i] += params['dupl']
        if params.get('demidupl_gain', 0) > 0:                        # AR:107-111
            if i + 1 <= max_chr:
                Q[i, i + 1] += params['demidupl_gain'] / 2
            if i + 2 <= max_chr:
                Q[i, i + 2] += params['demidupl_gain'] / 2
        if params.get('demidupl_loss', 0) > 0:                        # AR:113-117
            if i - 1 >= 1:
                Q[i, i - 1] += params['demidupl_loss'] / 2
            if i - 2 >= 1:
                Q[i, i - 2] += params['demidupl_loss'] / 2
        if base_number is not None and base_number > 0:               # AR:120-127
            if i + base_number <= max_chr:
                Q[i, i + base_number] += params.get('base_number_gain', 0)
            if i - base_number >= 1:
                Q[i, i - base_number] += params.get('base_number_loss', 0)
    for i in range(max_chr + 1):                                      # AR:130-131
        Q[i, i] = -np.sum(Q[i, :])
    return Q


# --------------------------------------------------------------------------- a3
def transition_probabilities(Q, branch_length, expm=scipy.linalg.expm):
    """P(t) with the reference's floor / renormalise / fallbacks -- AR:143-166."""
    n = Q.shape[0]
    if branch_length < 0:                                             # AR:143-145
        branch_length = 0.0
    if branch_length == 0:                                            # AR:146-148
        return np.eye(n)
    try:
        P = expm(Q * branch_length)                                   # AR:151
        P = np.maximum(P, 1e-12)                                      # AR:153
        P = P / P.sum(axis=1, keepdims=True)                          # AR:154
        if np.isnan(P).any():                                         # AR:156-159
            return np.eye(n)
        return P
    except Exception:                                                 # AR:161-166
        return np.eye(n)


def transition_probabilities_many(args):
    """(Q, branch lengths) -> list of post-processed matrices; a picklable worker so that tests
    can spread the oracle's per-branch scipy expm calls over host processes."""
    Q, lengths = args
    return [transition_probabilities(Q, float(t)) for t in lengths]


def raw_expm(Q, branch_length):
    """Unprocessed expm(Q t) -- the third-party call at AR:151."""
    return scipy.linalg.expm(Q * branch_length)


# --------------------------------------------------------------------------- a5 / a6
def leaf_vector(observed_count, max_chr):
    """Leaf conditional-likelihood vector -- AR:201-214."""
    v = np.zeros(max_chr + 1)
    if observed_count is not None:
        if 0 <= observed_count <= max_chr:
            v[int(observed_count)] = 1.0
        else:                      # out of range (NaN fails the comparison too) -> state max_chr
            v[max_chr] = 1.0
    else:
        v.fill(1.0 / (max_chr + 1))
    return v


def effective_branch_length(dist):
    """AR:224-227: missing or negative child.dist becomes 1e-6 before P is requested."""
    if dist is None or dist < 0:
        return 1e-6
    return dist


def prune(tree, data, max_chr, params, branch_params=None, faithful_q=False):
    """Felsenstein up-pass -- AR:189-249 -- as an explicit post-order loop (the reference
    recurses; the arithmetic per node is identical).

    Returns (vectors: dict id(node)->ndarray, rescale_events: int).
    ``faithful_q=True`` rebuilds Q for every branch like AR:141 does (same numbers, the
    reference's real cost); otherwise Q is built once per distinct parameter dict.
    """
    branch_params = branch_params or {}
    vectors = {}
    rescales = 0
    q_cache = {}

    def q_for(p):
        if faithful_q:
            return build_q(p, max_chr)
        key = id(p)
        if key not in q_cache:
            q_cache[key] = build_q(p, max_chr)
        return q_cache[key]

    for node in tree.traverse("postorder"):
        if node.is_leaf():
            vectors[id(node)] = leaf_vector(data.get(node.name), max_chr)
            continue
        v = np.ones(max_chr + 1)                                      # AR:219
        for child in node.children:                                   # AR:221
            Lc = vectors[id(child)]
            t = effective_branch_length(child.dist)
            P = transition_probabilities(q_for(branch_params.get(id(child), params)), t)
            v = v * np.dot(P, Lc)                                     # AR:237-238
            s = np.sum(v)                                             # AR:241
            if s > 0 and s < 1e-100:                                  # AR:242-243 (untracked)
                v = v / s
                rescales += 1
        vectors[id(node)] = v
    return vectors, rescales


def normalise_root_frequencies(freq, max_chr):
    """AR:180-186."""
    if freq is not None and len(freq) == max_chr + 1:
        freq = np.asarray(freq, dtype=float)
        return freq / np.sum(freq)
    return np.ones(max_chr + 1) / (max_chr + 1)


def total_log_likelihood(root_vector, root_freq):
    """AR:251-272."""
    total = np.sum(root_vector * root_freq)
    if total <= 0:
        return -np.inf
    return float(np.log(total))


def log_likelihood(tree, data, max_chr, params, root_freq=None, branch_params=None,
                   faithful_q=False):
    vectors, rescales = prune(tree, data, max_chr, params, branch_params, faithful_q)
    root_freq = normalise_root_frequencies(root_freq, max_chr)
    root = tree.get_tree_root()
    return total_log_likelihood(vectors[id(root)], root_freq), vectors, rescales


# --------------------------------------------------------------------------- a7
def ancestral_states(vector):
    """AR:312-316 for one node: (ml_count, ml_probability, probabilities)."""
    s = np.sum(vector)
    p = vector / s if s > 0 else np.ones_like(vector) / len(vector)
    k = int(np.argmax(p))
    return k, float(p[k]), p


# --------------------------------------------------------------------------- a8
def objective(x, names, base_params, tree, data, max_chr, root_freq):
    """Negative log-likelihood with the reference's clamps and penalties -- AR:341-371."""
    try:
        params = dict(base_params)
        for i, name in enumerate(names):
            params[name] = max(x[i], 1e-7) if name != 'linear_rate' else x[i]
        ll, _, _ = log_likelihood(tree, data, max_chr, params, root_freq)
        if np.isinf(ll) or np.isnan(ll):
            return 1e9
        return -ll
    except Exception:
        return 1e10


def aic_bic(log_likelihood_value, n_params, n_datapoints):
    """AR:788-789."""
    aic = 2 * n_params - 2 * log_likelihood_value
    bic = np.log(n_datapoints) * n_params - 2 * log_likelihood_value if n_datapoints > 0 else aic
    return aic, bic


# --------------------------------------------------------------------------- a10
def classify_event(from_state, to_state, base_number=None):
    """AR:441-464."""
    if to_state == from_state + 1:
        return 'gain'
    if to_state == from_state - 1:
        return 'loss'
    if to_state == 2 * from_state and from_state > 0:
        return 'duplication'
    if base_number and base_number > 0:
        if to_state == from_state + base_number:
            return 'base_number_gain'
        if to_state == from_state - base_number:
            return 'base_number_loss'
    delta = to_state - from_state
    if delta > 1 and delta < from_state:
        return 'demi_gain_like'
    if delta < -1 and abs(delta) < from_state / 2:
        return 'demi_loss_like'
    return 'other'


def simulate_branch(Q, start_state, branch_length, rng, base_number=None):
    """Forward Gillespie walk along one branch -- AR:467-521.  Returns (type counts[8], end)."""
    n = Q.shape[0]
    counts = np.zeros(len(EVENT_TYPES), dtype=np.int64)
    state = int(start_state)
    time = 0.0
    while time < branch_length:
        if state < 0 or state >= n:
            break
        rate_out = -Q[state, state]
        if rate_out <= 1e-9:                                           # AR:484-486
            break
        dt = rng.exponential(1.0 / rate_out)                           # AR:489
        if time + dt >= branch_length:                                 # AR:491-493
            break
        time += dt
        w = Q[state, :].copy()                                         # AR:499-503
        w[state] = 0
        w = np.maximum(w, 0)
        tot = np.sum(w)
        if tot <= 1e-9:                                                # AR:505-507
            break
        nxt = int(rng.choice(n, p=w / tot))                            # AR:509-510
        counts[EVENT_TYPES.index(classify_event(state, nxt, base_number))] += 1
        state = nxt
    return counts, state


def stochastic_mapping_counts(tree, Q, root_probs, num_mappings, rng, base_number=None,
                              branch_q=None):
    """Per-replicate event counts -- AR:523-576.

    Returns (branch_nodes in preorder, counts[num_mappings, n_branches, 8]).  Branches with
    dist None/<=0 copy the parent state and record no events (AR:562-566).  branch_q maps
    id(node) -> (Q, base_number) for branches with their own parameters (AR:569).
    """
    branch_q = branch_q or {}
    root = tree.get_tree_root()
    order = [nd for nd in tree.traverse("preorder") if not nd.is_root()]
    index = {id(nd): k for k, nd in enumerate(order)}
    out = np.zeros((num_mappings, len(order), len(EVENT_TYPES)), dtype=np.int64)
    n = Q.shape[0]
    for rep in range(num_mappings):
        states = {id(root): int(rng.choice(n, p=root_probs))}          # AR:544-545
        for nd in order:
            start = states[id(nd.up)]
            t = nd.dist
            if t is None or t <= 0:
                states[id(nd)] = start
                continue
            q_here, bn_here = branch_q.get(id(nd), (Q, base_number))
            c, end = simulate_branch(q_here, start, t, rng, bn_here)
            out[rep, index[id(nd)]] = c
            states[id(nd)] = end
    return order, out


def summarize_counts(counts):
    """AR:627-640 for one (branch, type) series of per-replicate counts."""
    counts = np.asarray(counts)
    if counts.size == 0:
        return {'mean': 0, 'std': 0, 'ci_lower': 0, 'ci_upper': 0, 'total_occurrences': 0}
    return {'mean': np.mean(counts), 'std': np.std(counts),
            'ci_lower': np.percentile(counts, 2.5), 'ci_upper': np.percentile(counts, 97.5),
            'total_occurrences': np.sum(counts)}


# --------------------------------------------------------------------------- f1
def rearrangement_rows(tree, count_attr='count'):
    """Fusion/fission event rows, sorted like the reference -- AR:1155-1195."""
    rows = []
    for node in tree.traverse():
        if node.is_root():
            continue
        c = getattr(node, count_attr, None)
        p = getattr(node.up, count_attr, None)
        if c is None or p is None:
            continue
        change = c - p
        if change == 0:
            continue
        rows.append({
            'branch_to_node': node.name if node.name else f"To_Leaf_{node.get_leaves()[0].name[:10]}",
            'parent_node': node.up.name if node.up.name else "Root_or_Unnamed",
            'event_type': 'fission' if change > 0 else 'fusion',
            'change_magnitude': abs(change),
            'count_from': p,
            'count_to': c,
        })
    # pandas sort_values(by=[type, magnitude], ascending=[True, False]) is a stable lexsort
    rows.sort(key=lambda r: (r['event_type'], -r['change_magnitude']))
    return rows


# --------------------------------------------------------------------------- numpy sum order
def numpy_pairwise_sum(a):
    """numpy's float add-reduce order for a contiguous 1-D array (pairwise summation with
    an 8-way unrolled base case, numpy/_core/src/umath/loops_utils.h.src `pairwise_sum`).
    Restated so device kernels that want bit-identical `np.sum` rows (AR:131) have a spec."""
    a = np.asarray(a, dtype=np.float64)
    n = a.shape[0]
    if n < 8:
        res = 0.0
        for x in a:
            res += float(x)
        return res
    if n <= 128:
        r = [float(a[k]) for k in range(8)]
        i = 8
        while i < n - (n % 8):
            for k in range(8):
                r[k] += float(a[i + k])
            i += 8
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
        while i < n:
            res += float(a[i])
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return numpy_pairwise_sum(a[:n2]) + numpy_pairwise_sum(a[n2:])


def expm_flops(n, s):
    """Algorithmic FLOPs of one Pade-13 scaling-and-squaring expm."""
    return (2.0 * (6 + s) + 8.0 / 3.0) * float(n) ** 3


__all__ = [name for name in globals() if not name.startswith('_') and name not in ('np', 'math', 'scipy', 'annotations')]
