"""CPU: the marginal / joint reconstruction oracle (oracle/reconstruction_oracle.py) against an
exhaustive enumeration over the reference's own transition matrices.  The reference has no
down-pass or joint backtrack (AR:283-295), so this enumeration is what pins the two dynamic
programmes; the GPU tests then compare the device kernels with them."""
import numpy as np
import pytest

from conftest import make_tree
from oracle import chromevol_oracle as O
from oracle import reconstruction_oracle as R

CASES = [
    # (newick, data, max_chr, parameter overrides)
    ("((A:0.3,B:0.7):0.4,C:1.1);", {"A": 2, "B": 3, "C": 1}, 4, {}),
    ("((A:0.3,B:0.7):0.4,(C:0.2,D:0.9):0.05);", {"A": 1, "B": 3, "C": 3, "D": 2}, 3,
     {"gain": 0.8, "loss": 0.5, "dupl": 0.3}),
    ("((A:0.3,B:0.7):0.4,C:1.1,D:0.6);", {"A": 2, "B": 0, "C": 3}, 3,           # polytomy, D missing
     {"gain": 1.5, "loss": 0.2, "dupl": 0.6, "demidupl_gain": 0.3}),
    ("((A:0.01,B:0.02):0.03,C:0.01);", {"A": 1, "B": 4, "C": 2}, 4, {"gain": 0.05, "loss": 0.02}),  # floor-driven
    ("((A:0.3,B:0.7):0.4,C:1.1);", {"A": 2, "B": 3, "C": 1}, 4, {"base_number": 2}),
]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_dynamic_programmes_match_enumeration(case):
    nwk, data, max_chr, over = CASES[case]
    tree = make_tree(nwk)
    params = O.default_params(**over)
    rng = np.random.default_rng(case)
    prior = rng.random(max_chr + 1) + 0.1 if case % 2 else None
    marg_bf, best_bf, logp_bf = R.brute_force(tree, data, max_chr, params, prior)
    post = R.marginal_posteriors(tree, data, max_chr, params, prior)
    states, logp = R.joint_reconstruction(tree, data, max_chr, params, prior)
    for nd in tree.traverse():
        assert np.allclose(post[id(nd)], marg_bf[id(nd)], rtol=1e-12, atol=1e-300), nd.name
        assert states[id(nd)] == best_bf[id(nd)], nd.name
        assert abs(post[id(nd)].sum() - 1.0) < 1e-12
    assert abs(logp - logp_bf) <= 1e-12 * abs(logp_bf)


def test_marginal_differs_from_up_pass_argmax_somewhere(shipped):
    """The reference's `ml_count` is the argmax of the UP-pass vector (AR:312-316); the marginal
    posterior also sees the data outside the subtree, so it cannot be a copy of it."""
    nwk, counts = shipped
    tree = make_tree(nwk)
    params = O.default_params()
    post = R.marginal_posteriors(tree, counts, 35, params)
    _, vectors, _ = O.log_likelihood(tree, counts, 35, params)
    root = tree.get_tree_root()
    # at the root the two coincide up to the prior (uniform here)
    pr = vectors[id(root)] / vectors[id(root)].sum()
    assert np.allclose(post[id(root)], pr, rtol=1e-10)
    diff = max(np.abs(post[id(nd)] - vectors[id(nd)] / vectors[id(nd)].sum()).max()
               for nd in tree.traverse() if not nd.is_leaf())
    assert diff > 1e-3
    # observed leaves keep their observation
    for nd in tree.get_leaves():
        if nd.name in counts:
            assert post[id(nd)][int(counts[nd.name])] == 1.0


def test_joint_states_are_a_fixed_point_of_their_own_probability(shipped):
    """No single-node change of the joint assignment increases its probability."""
    nwk, counts = shipped
    tree = make_tree(nwk)
    params = O.default_params()
    states, logp = R.joint_reconstruction(tree, counts, 35, params)
    Q = O.build_q(params, 35)
    P = {id(nd): O.transition_probabilities(Q, O.effective_branch_length(nd.dist))
         for nd in tree.traverse() if nd.up is not None}
    prior = np.ones(36) / 36

    def logprob(st):
        lp = np.log(prior[st[id(tree.get_tree_root())]])
        for nd in tree.traverse():
            if nd.up is not None:
                lp += np.log(P[id(nd)][st[id(nd.up)], st[id(nd)]])
            if nd.is_leaf():
                lp += np.log(O.leaf_vector(counts.get(nd.name), 35)[st[id(nd)]])
        return lp

    base = logprob(states)
    assert abs(base - logp) <= 1e-9 * abs(logp)
    internal = [nd for nd in tree.traverse() if not nd.is_leaf()]
    for nd in internal[:12]:
        for s in range(36):
            if s == states[id(nd)]:
                continue
            alt = dict(states)
            alt[id(nd)] = s
            assert logprob(alt) <= base + 1e-12
