"""Seeded synthetic workloads of the shapes BASELINE.json names.

Workload GENERATION only (host numpy/scipy): random binary trees, tip counts simulated down the
tree, and batches of parameter vectors.  Nothing here is on the measured path.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg

from .newick import TreeNode

DEFAULT_RATES = dict(gain=0.1, loss=0.1, dupl=0.05, demidupl_gain=0.01, demidupl_loss=0.01,
                     base_number_gain=0.01, base_number_loss=0.01, linear_rate=0.01)


def random_tree(n_leaves, seed, mean_len=0.05, min_len=1e-4):
    """Random binary topology (uniform random joins), branch lengths min_len + Exp(mean_len),
    all distinct; leaves are named t0..t{n-1}; the root carries length 0."""
    rng = np.random.default_rng(seed)
    pool = [TreeNode(name=f"t{i}") for i in range(n_leaves)]
    k = 0
    while len(pool) > 1:
        i, j = sorted(rng.choice(len(pool), size=2, replace=False))
        b = pool.pop(j)
        a = pool.pop(i)
        parent = TreeNode(name=f"n{k}")
        k += 1
        parent.add_child(a)
        parent.add_child(b)
        pool.append(parent)
    root = pool[0]
    for node in root.traverse():
        node.dist = float(min_len + rng.exponential(mean_len))
    root.dist = 0.0
    return root


def _rate_matrix(params, max_chr):
    """Plain numpy Q for data generation (same model as AR:75-134; not the parity oracle)."""
    n = max_chr + 1
    Q = np.zeros((n, n))
    bn = params.get('base_number') or 0
    for i in range(1, n):
        f = 1 + params['linear_rate'] * i
        if i < max_chr:
            Q[i, i + 1] = params['gain'] * f
        if i > 1:
            Q[i, i - 1] = params['loss'] * f
        if 2 * i <= max_chr:
            Q[i, 2 * i] += params['dupl']
        if params['demidupl_gain'] > 0:
            if i + 1 <= max_chr:
                Q[i, i + 1] += params['demidupl_gain'] / 2
            if i + 2 <= max_chr:
                Q[i, i + 2] += params['demidupl_gain'] / 2
        if params['demidupl_loss'] > 0:
            if i - 1 >= 1:
                Q[i, i - 1] += params['demidupl_loss'] / 2
            if i - 2 >= 1:
                Q[i, i - 2] += params['demidupl_loss'] / 2
        if bn > 0:
            if i + bn <= max_chr:
                Q[i, i + bn] += params['base_number_gain']
            if i - bn >= 1:
                Q[i, i - bn] += params['base_number_loss']
    Q[np.arange(n), np.arange(n)] = -Q.sum(axis=1)
    return Q


def simulate_tip_counts(tree, max_chr, seed, root_state=20, params=None):
    """Tip counts drawn down the tree from `root_state` under `params` (default rates)."""
    rng = np.random.default_rng(seed)
    params = dict(DEFAULT_RATES, **(params or {}))
    Q = _rate_matrix(params, max_chr)
    state = {id(tree): int(root_state)}
    counts = {}
    for node in tree.traverse("preorder"):
        if node.is_root():
            continue
        P = scipy.linalg.expm(Q * node.dist)
        row = np.maximum(P[state[id(node.up)]], 0)
        s = int(rng.choice(max_chr + 1, p=row / row.sum()))
        state[id(node)] = s
        if node.is_leaf():
            counts[node.name] = s
    return counts


def simulate_tip_counts_fast(tree, max_chr, seed, root_state=20, params=None):
    """Tip counts from a jump-chain (Gillespie) simulation down the tree: no matrix exponential,
    so 10,000-taxon / N=256 workloads are generated in a second.  Negative rates are ignored."""
    rng = np.random.default_rng(seed)
    params = dict(DEFAULT_RATES, **(params or {}))
    Q = np.maximum(_rate_matrix(params, max_chr), 0.0)
    np.fill_diagonal(Q, 0.0)
    hold = Q.sum(axis=1)
    cum = np.cumsum(Q, axis=1)
    state = {id(tree): int(root_state)}
    counts = {}
    for node in tree.traverse("preorder"):
        if node.is_root():
            continue
        s, t = state[id(node.up)], float(node.dist)
        while hold[s] > 0:
            t -= rng.exponential(1.0 / hold[s])
            if t <= 0:
                break
            s = int(np.searchsorted(cum[s], rng.random() * hold[s], side="right"))
            s = min(s, max_chr)
        state[id(node)] = s
        if node.is_leaf():
            counts[node.name] = s
    return counts


def random_param_rows(batch, seed, linear='positive', base_number=0, corners=0):
    """(batch, 9) parameter rows in device order: rates
    LogUniform(1e-3, 10) (base-number rates LogUniform(1e-3, 1)); linear_rate U(0, 0.1) for
    'positive', U(-0.05, 0.05) for 'mixed'; the last `corners` rows are optimiser bound corners."""
    rng = np.random.default_rng(seed)
    rows = np.zeros((batch, 9))
    lu = lambda lo, hi, size: np.exp(rng.uniform(np.log(lo), np.log(hi), size=size))
    rows[:, 0:5] = lu(1e-3, 10.0, (batch, 5))
    rows[:, 5:7] = lu(1e-3, 1.0, (batch, 2))
    if linear == 'positive':
        rows[:, 7] = rng.uniform(0.0, 0.1, size=batch)
    else:
        rows[:, 7] = rng.uniform(-0.05, 0.05, size=batch)
    rows[:, 8] = base_number
    for c in range(min(corners, batch)):
        r = batch - 1 - c
        rows[r, 0:7] = np.where(rng.random(7) < 0.5, 1e-7, 10.0)
        rows[r, 7] = 1.0 if (c % 2 == 0) else -1.0
    return rows


def default_param_row(base_number=0):
    d = DEFAULT_RATES
    return np.array([d['gain'], d['loss'], d['dupl'], d['demidupl_gain'], d['demidupl_loss'],
                     d['base_number_gain'], d['base_number_loss'], d['linear_rate'],
                     float(base_number)])


def row_to_dict(row):
    names = ['gain', 'loss', 'dupl', 'demidupl_gain', 'demidupl_loss', 'base_number_gain',
             'base_number_loss', 'linear_rate']
    d = {nm: float(row[k]) for k, nm in enumerate(names)}
    d['base_number'] = int(row[8]) if row[8] > 0 else None
    return d
