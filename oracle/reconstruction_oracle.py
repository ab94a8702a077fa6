"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY, NOT A PRODUCT PATH.

Marginal (up-pass + down-pass) and joint (Pupko et al. 2000) ancestral-state reconstruction over
the reference's own transition matrices.

PARITY UNPINNED AGAINST THE REFERENCE: `/root/reference/src/ancestral_reconstruction.py` has no
down-pass and no joint backtrack -- its `ancestral_state_reconstruction` (AR:274-322) documents
that a true marginal reconstruction "requires a full upward and downward pass" and then takes the
argmax of the normalised up-pass vector.  BASELINE.json's north star names "marginal and joint
ancestral-state reconstruction", so the device package offers both as ADDITIVE calls
(`marginal_reconstruction`, `joint_reconstruction`); `ml_count` keeps the reference's meaning.
What pins this oracle instead:
  * every transition matrix comes from `chromevol_oracle.transition_probabilities` (AR:136-166,
    live-pinned to the reference), the leaf vectors from AR:201-214, the root prior from
    AR:180-186;
  * `brute_force` enumerates every assignment of states to the internal nodes of a small tree
    with those same matrices; tests/test_oracle.py checks that the two dynamic programmes below
    reproduce its marginals and its best assignment.

Definitions (P'_c = post-processed matrix of the branch above node c, obs_c = leaf vector):
    up_c(j)   = obs_c(j)                             (leaf)
              = prod_{s child of c} sum_k P'_s(j,k) up_s(k)   (internal; AR:219-238 without the
                conditional rescale, whose factor the reference drops -- the posterior is
                invariant to any per-node scale)
    down_root = root prior;  down_c(j) = sum_i [down_p(i) prod_{s sibling of c} msg_s(i)] P'_c(i,j)
    posterior_c(j) = up_c(j) down_c(j) / sum_j(...)
Joint: L_c(i) = max_j P'_c(i,j) M_c(j), C_c(i) = the first j attaining it, M_c = obs_c (leaf) or the
product of the children's L in file order, rescaled by the power of two that brings its largest
entry into [0.5, 1) (exact in binary floating point, so a device implementation that multiplies
in the same order reproduces every comparison bit for bit); the root takes the first i that
maximises prior(i) M_root(i) and the states follow C down the tree.
"""
from __future__ import annotations

import itertools
import math

import numpy as np

from . import chromevol_oracle as O


def _matrices(tree, max_chr, params, P_of=None):
    Q = O.build_q(params, max_chr)
    out = {}
    for node in tree.traverse():
        if node.up is None:
            continue
        if P_of is not None:
            out[id(node)] = np.asarray(P_of(node), dtype=np.float64)
        else:
            out[id(node)] = O.transition_probabilities(Q, O.effective_branch_length(node.dist))
    return out


def _pow2_scale(m):
    """m * 2^-e with e = exponent of max(m) (frexp convention: max in [0.5, 1)); (scaled, e)."""
    mx = float(np.max(m))
    if not (mx > 0.0) or not math.isfinite(mx):
        return m, 0
    e = math.frexp(mx)[1]
    return np.ldexp(m, -e), e


def marginal_posteriors(tree, data, max_chr, params, root_freq=None, P_of=None):
    """dict id(node) -> posterior state probabilities of every node given ALL the data."""
    n = max_chr + 1
    P = _matrices(tree, max_chr, params, P_of)
    prior = O.normalise_root_frequencies(root_freq, max_chr)
    up, msg = {}, {}
    for node in tree.traverse("postorder"):
        if node.is_leaf():
            up[id(node)] = O.leaf_vector(data.get(node.name), max_chr)
        else:
            v = np.ones(n)
            for ch in node.children:
                v = v * msg[id(ch)]
            s = v.sum()
            up[id(node)] = v / s if s > 0 else v          # any positive scale: posterior unchanged
        if node.up is not None:
            msg[id(node)] = P[id(node)] @ up[id(node)]
    down, post = {}, {}
    for node in tree.traverse("levelorder"):
        if node.up is None:
            d = prior
        else:
            tmp = down[id(node.up)].copy()
            for sib in node.up.children:
                if sib is not node:
                    tmp = tmp * msg[id(sib)]
            s = tmp.sum()
            if s > 0:
                tmp = tmp / s
            d = tmp @ P[id(node)]
            s = d.sum()
            if s > 0:
                d = d / s
        down[id(node)] = d
        q = up[id(node)] * d
        s = q.sum()
        post[id(node)] = q / s if s > 0 else np.ones(n) / n
    return post


def joint_reconstruction(tree, data, max_chr, params, root_freq=None, P_of=None):
    """(dict id(node) -> state of the single most probable joint assignment, its log probability
    jointly with the data)."""
    P = _matrices(tree, max_chr, params, P_of)
    prior = O.normalise_root_frequencies(root_freq, max_chr)
    L, C, expo = {}, {}, {}
    M_root = None
    for node in tree.traverse("postorder"):
        if node.is_leaf():
            m, e = O.leaf_vector(data.get(node.name), max_chr), 0
        else:
            m = L[id(node.children[0])]
            e = expo[id(node.children[0])]
            for ch in node.children[1:]:
                m = m * L[id(ch)]
                e += expo[id(ch)]
            m, e2 = _pow2_scale(m)
            e += e2
        if node.up is None:
            M_root, e_root = m, e
            continue
        prod = P[id(node)] * m[None, :]
        C[id(node)] = np.argmax(prod, axis=1)
        L[id(node)] = prod[np.arange(prod.shape[0]), C[id(node)]]
        expo[id(node)] = e
    root = tree.get_tree_root()
    score = prior * M_root
    states = {id(root): int(np.argmax(score))}
    best = float(score[states[id(root)]])
    logp = (math.log(best) if best > 0 else -math.inf) + e_root * math.log(2.0)
    for node in tree.traverse("levelorder"):
        if node.up is not None:
            states[id(node)] = int(C[id(node)][states[id(node.up)]])
    return states, logp


def brute_force(tree, data, max_chr, params, root_freq=None):
    """Enumerate every state assignment of the internal nodes (and of the leaves without data)
    of a SMALL tree: (dict id(node) -> marginal posterior, dict id(node) -> state of the best
    assignment, log probability of the best assignment)."""
    n = max_chr + 1
    P = _matrices(tree, max_chr, params)
    prior = O.normalise_root_frequencies(root_freq, max_chr)
    nodes = list(tree.traverse("levelorder"))
    obs = {id(nd): O.leaf_vector(data.get(nd.name), max_chr) for nd in nodes if nd.is_leaf()}
    marg = {id(nd): np.zeros(n) for nd in nodes}
    best, best_assign = -1.0, None
    for assign in itertools.product(range(n), repeat=len(nodes)):
        st = {id(nd): s for nd, s in zip(nodes, assign)}
        p = prior[st[id(nodes[0])]]
        for nd in nodes:
            if nd.up is not None:
                p *= P[id(nd)][st[id(nd.up)], st[id(nd)]]
            if nd.is_leaf():
                p *= obs[id(nd)][st[id(nd)]]
            if p == 0.0:
                break
        if p == 0.0:
            continue
        for nd in nodes:
            marg[id(nd)][st[id(nd)]] += p
        if p > best:
            best, best_assign = p, st
    tot = marg[id(nodes[0])].sum()
    return ({k: v / tot for k, v in marg.items()}, best_assign,
            math.log(best) if best > 0 else -math.inf)
