"""GPU: marginal (down-pass) and joint (Pupko) reconstruction kernels against
oracle/reconstruction_oracle.py.  ADDITIVE calls -- the reference has neither (AR:283-295); the
oracle is pinned by exhaustive enumeration (tests/test_reconstruction_oracle.py)."""
import numpy as np
import pytest

from conftest import make_tree
from oracle import chromevol_oracle as O
from oracle import reconstruction_oracle as R

pytestmark = pytest.mark.gpu


def _calc(tree, data, max_chr, **over):
    from chromevol_enhanced_b200 import ChromEvolutionModel, ChromEvolLikelihoodCalculator
    model = ChromEvolutionModel(max_chr)
    model.set_parameters(**O.default_params(**over))
    return ChromEvolLikelihoodCalculator(tree, data, model)


def _device_P(calc):
    """branch matrix of every node from the device (the materialised companion engine)."""
    eng = calc._engine._materialised_companion()
    P = eng.transition_matrices(0).cpu().numpy()
    flat = calc._flat
    return lambda node: P[flat.branch_of[flat.index[id(node)]]]


def test_shipped_data_marginal_and_joint(shipped):
    nwk, counts = shipped
    tree = make_tree(nwk)
    calc = _calc(tree, counts, 35)
    before = None
    calc.ancestral_state_reconstruction()
    before = {id(nd): (int(nd.ml_count), float(nd.ml_probability)) for nd in tree.traverse()}
    calc.marginal_reconstruction()
    _, logp = calc.joint_reconstruction()
    params = O.default_params()
    post = R.marginal_posteriors(tree, counts, 35, params)
    states, logp_ref = R.joint_reconstruction(tree, counts, 35, params)
    for nd in tree.traverse():
        assert np.abs(nd.marginal_probabilities - post[id(nd)]).max() <= 1e-11
        assert int(nd.marginal_count) == int(np.argmax(post[id(nd)]))
        assert int(nd.joint_count) == states[id(nd)]
        # the reference-shaped attributes keep the reference's meaning
        assert (int(nd.ml_count), float(nd.ml_probability)) == before[id(nd)]
    assert abs(logp - logp_ref) <= 1e-9 * abs(logp_ref)
    # bit-exact states and probability against the oracle run on the DEVICE's matrices
    states_dev, logp_dev = R.joint_reconstruction(tree, counts, 35, params, P_of=_device_P(calc))
    assert all(int(nd.joint_count) == states_dev[id(nd)] for nd in tree.traverse())
    assert abs(logp - logp_dev) <= 1e-13 * abs(logp_dev)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_trees_with_polytomies_and_missing_data(seed):
    from chromevol_enhanced_b200 import synthetic
    rng = np.random.default_rng(seed)
    N = [12, 40, 103][seed]
    tree = synthetic.random_tree(60, seed=seed)
    counts = synthetic.simulate_tip_counts(tree, N, seed=seed + 10, root_state=min(20, N // 2))
    leaves = tree.get_leaves()
    for nd in rng.choice(leaves, 5, replace=False):
        counts.pop(nd.name, None)                                   # missing data -> uniform leaf
    # a polytomy: graft the grandchildren of one child onto the root
    root = tree.get_tree_root()
    big = max(root.children, key=lambda c: len(c.children))
    if big.children:
        root.children.remove(big)
        for g in list(big.children):
            g.up = root
            g.dist = (g.dist or 0.0) + (big.dist or 0.0)
            root.children.append(g)
    over = dict(gain=float(rng.uniform(0.1, 2)), loss=float(rng.uniform(0.1, 2)),
                dupl=float(rng.uniform(0.01, 0.5)), linear_rate=float(rng.uniform(0, 0.05)))
    if seed == 2:
        over["base_number"] = 7
    calc = _calc(tree, counts, N, **over)
    prior = rng.random(N + 1) + 0.05
    calc.set_root_frequencies(prior)
    calc.marginal_reconstruction()
    _, logp = calc.joint_reconstruction()
    params = O.default_params(**over)
    post = R.marginal_posteriors(tree, counts, N, params, prior)
    P_of = _device_P(calc)
    states_dev, logp_dev = R.joint_reconstruction(tree, counts, N, params, prior, P_of=P_of)
    states_ref, logp_ref = R.joint_reconstruction(tree, counts, N, params, prior)
    agree = 0
    for nd in tree.traverse():
        assert np.abs(nd.marginal_probabilities - post[id(nd)]).max() <= 1e-10
        assert abs(nd.marginal_probabilities.sum() - 1.0) <= 1e-12
        assert int(nd.joint_count) == states_dev[id(nd)]
        agree += int(nd.joint_count) == states_ref[id(nd)]
    assert abs(logp - logp_dev) <= 1e-13 * abs(logp_dev)
    assert abs(logp - logp_ref) <= 1e-9 * abs(logp_ref)
    assert agree == len(list(tree.traverse()))          # no near ties on these seeds


def test_small_tree_against_enumeration():
    tree = make_tree("((A:0.3,B:0.7):0.4,(C:0.2,D:0.9):0.05);")
    data = {"A": 1, "B": 3, "C": 3, "D": 2}
    over = {"gain": 0.8, "loss": 0.5, "dupl": 0.3}
    calc = _calc(tree, data, 3, **over)
    calc.marginal_reconstruction()
    _, logp = calc.joint_reconstruction()
    marg, best, logp_bf = R.brute_force(tree, data, 3, O.default_params(**over))
    for nd in tree.traverse():
        assert np.abs(nd.marginal_probabilities - marg[id(nd)]).max() <= 1e-12
        assert int(nd.joint_count) == best[id(nd)]
    assert abs(logp - logp_bf) <= 1e-11 * abs(logp_bf)


def test_branch_sets_are_refused():
    tree = make_tree("((A:0.3,B:0.7):0.4,C:1.1);")
    calc = _calc(tree, {"A": 2, "B": 3, "C": 1}, 4)
    node = tree.get_leaves()[0]
    calc.model.set_branch_parameters(id(node), gain=0.5)
    with pytest.raises(NotImplementedError):
        calc.marginal_reconstruction()
