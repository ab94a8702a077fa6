"""GPU: branch-specific parameter sets (reference set_branch_parameters,
AR:67-73, used in the up-pass at AR:230-233 and by the mapper at AR:569) through the drop-in
classes, against the oracle (whose branch_params path is pinned to the live reference in
tests/test_oracle_vs_reference.py)."""
import numpy as np
import pytest

from oracle import chromevol_oracle as O

pytestmark = pytest.mark.gpu


def _hetero_setup(taxa, N, seed):
    from chromevol_enhanced_b200 import synthetic
    from chromevol_enhanced_b200.ancestral_reconstruction import ChromEvolutionModel
    tree = synthetic.random_tree(taxa, seed=seed)
    counts = synthetic.simulate_tip_counts(tree, N, seed=seed + 1)
    del counts[sorted(counts)[3]]                      # one leaf without data
    model = ChromEvolutionModel(N, custom_params={"gain": 0.4})
    overrides = {}
    for k, nd in enumerate(n for n in tree.traverse() if not n.is_root()):
        if k % 3 == 0:
            kw = {"gain": 1.5, "dupl": 0.3} if k % 2 == 0 else {"loss": 2.0, "base_number": 5}
            model.set_branch_parameters(id(nd), **kw)
            overrides[id(nd)] = O.default_params(**{"gain": 0.4, **kw})
    return tree, counts, model, overrides


@pytest.mark.parametrize("taxa,N", [(30, 24), (200, 100)])
def test_branch_specific_likelihood_and_states(taxa, N):
    from chromevol_enhanced_b200.ancestral_reconstruction import ChromEvolLikelihoodCalculator
    tree, counts, model, overrides = _hetero_setup(taxa, N, seed=5)
    n = N + 1
    calc = ChromEvolLikelihoodCalculator(tree, counts, model)
    calc.set_root_frequencies(np.ones(n) / n)
    rows, node_set = calc.branch_parameter_sets()
    assert rows.shape[0] == 3 and set(np.unique(node_set)) == {0, 1, 2}
    ll = calc.calculate_total_log_likelihood()
    ref, vectors, _ = O.log_likelihood(tree, counts, N, O.default_params(gain=0.4), np.ones(n) / n,
                                       branch_params=overrides)
    assert ll == pytest.approx(ref, rel=1e-9)
    calc.ancestral_state_reconstruction()
    for nd in tree.traverse():
        v = calc.calculate_likelihood_at_node(nd)
        assert np.allclose(v, vectors[id(nd)], rtol=1e-8, atol=1e-300)
        ml, prob, _ = O.ancestral_states(vectors[id(nd)])
        assert int(nd.ml_count) == ml
        assert float(nd.ml_probability) == pytest.approx(prob, rel=1e-8)
    # dropping the overrides gives the homogeneous answer through the fused path again
    model.branch_params = {}
    calc.likelihoods = {}
    ll_h = calc.calculate_total_log_likelihood()
    ref_h, _, _ = O.log_likelihood(tree, counts, N, O.default_params(gain=0.4), np.ones(n) / n)
    assert ll_h == pytest.approx(ref_h, rel=1e-9)


def test_branch_specific_objective_is_optimisable():
    """objective_function (AR:341-371) mutates model.params; the branch overrides stay."""
    from chromevol_enhanced_b200.ancestral_reconstruction import ChromEvolOptimizer
    tree, counts, model, overrides = _hetero_setup(30, 24, seed=9)
    opt = ChromEvolOptimizer(tree, counts, model)
    opt.calculator.set_root_frequencies(np.ones(25) / 25)
    f = opt.objective_function(np.array([0.7, 0.2]), ["gain", "loss"])
    base = O.default_params(gain=0.7, loss=0.2)
    ref, _, _ = O.log_likelihood(tree, counts, 24, base, np.ones(25) / 25, branch_params=overrides)
    assert f == pytest.approx(-ref, rel=1e-9)


def test_branch_specific_mapper_matches_oracle_simulator():
    """Per-(branch, type) mean event counts of the device mapper with three parameter sets
    against the CPU oracle simulator with the same per-branch (Q, base_number)."""
    from chromevol_enhanced_b200 import synthetic
    from chromevol_enhanced_b200.ancestral_reconstruction import ChromEvolutionModel, StochasticMapper
    N = 20
    tree = synthetic.random_tree(12, seed=3, mean_len=0.3)
    base = dict(gain=1.2, loss=0.7, dupl=0.4, demidupl_gain=0.3, demidupl_loss=0.2)
    model = ChromEvolutionModel(N, custom_params=base)
    branch_q = {}
    for k, nd in enumerate(n for n in tree.traverse() if not n.is_root()):
        if k % 2 == 0:
            kw = ({"gain": 3.0} if k % 4 == 0 else
                  {"base_number": 3, "base_number_gain": 0.8, "base_number_loss": 0.6})
            model.set_branch_parameters(id(nd), **kw)
            p = O.default_params(**{**base, **kw})
            branch_q[id(nd)] = (O.build_q(p, N), p.get("base_number"))
    root_probs = np.zeros(N + 1)
    root_probs[[6, 7, 9]] = [0.5, 0.3, 0.2]
    tree.add_features(state_probabilities=root_probs)
    R = 300_000
    mapper = StochasticMapper(tree, model, num_mappings=R, seed=17)
    summary = mapper.perform_stochastic_mapping()
    rng = np.random.default_rng(2)
    Rc = 3000
    order, out = O.stochastic_mapping_counts(tree, O.build_q(O.default_params(**base), N), root_probs,
                                             Rc, rng, base_number=None, branch_q=branch_q)
    worst, n_tests, bn_events = 0.0, 0, 0.0
    for b, nd in enumerate(order):
        for e, et in enumerate(O.EVENT_TYPES):
            m_dev, s_dev = summary[id(nd)][et]['mean'], summary[id(nd)][et]['std']
            m_cpu = out[:, b, e].mean()
            if et.startswith('base_number'):
                bn_events += m_dev
                if id(nd) not in branch_q or branch_q[id(nd)][1] is None:
                    assert m_dev == 0          # no base number on this branch: type cannot occur
            if m_dev * Rc < 20:
                continue
            se = max(s_dev, 1e-3) * np.sqrt(1.0 / Rc + 1.0 / R)
            worst = max(worst, abs(m_dev - m_cpu) / se)
            n_tests += 1
    assert n_tests > 15 and bn_events > 0
    assert worst < 5.0, worst


def test_too_many_parameter_sets_for_the_mapper_is_an_error():
    """The mapper keeps one compact rate table per set in shared memory; a model with more sets
    than fit fails loudly (status + message), it does not fall back to anything."""
    import torch
    from chromevol_enhanced_b200 import _lib, mapper, synthetic
    from chromevol_enhanced_b200.flatten import FlatTree
    N = 100
    tree = synthetic.random_tree(40, seed=1)
    flat = FlatTree(tree, {}, N)
    S = 40
    rows = np.tile(np.array([0.1, 0.1, 0.05, 0.01, 0.01, 0.01, 0.01, 0.01, 0.0]), (S, 1))
    rows[:, 0] += np.arange(S) * 0.01
    node_set = np.arange(flat.n_nodes) % S
    with pytest.raises(_lib.CevbError, match="shared memory"):
        mapper.simulate(flat, rows, N + 1, np.ones(N + 1) / (N + 1), 1000, seed=1, node_set=node_set)
    # a number of sets that fits runs
    res = mapper.simulate(flat, rows[:6], N + 1, np.ones(N + 1) / (N + 1), 2000, seed=1,
                          node_set=np.arange(flat.n_nodes) % 6)
    assert res.n_events > 0 and not res.overflowed()
    torch.cuda.synchronize()


def test_one_parameter_set_per_branch_costs_one_matrix_per_branch():
    """Every branch with its own parameter set (S = number of branches): the device path forms
    one transition matrix per branch like the reference (AR:230-237), not S x n_t of them
    (round-1 advisor finding: that was 300 GB at 1,000 taxa)."""
    from chromevol_enhanced_b200 import synthetic
    from chromevol_enhanced_b200.ancestral_reconstruction import (ChromEvolLikelihoodCalculator,
                                                                  ChromEvolutionModel)
    N, taxa = 60, 150
    tree = synthetic.random_tree(taxa, seed=21)
    counts = synthetic.simulate_tip_counts(tree, N, seed=22)
    model = ChromEvolutionModel(N)
    overrides = {}
    rng = np.random.default_rng(3)
    for nd in tree.traverse():
        if nd.is_root():
            continue
        kw = {"gain": float(rng.uniform(0.05, 2.0)), "loss": float(rng.uniform(0.05, 2.0))}
        model.set_branch_parameters(id(nd), **kw)
        overrides[id(nd)] = O.default_params(**kw)
    calc = ChromEvolLikelihoodCalculator(tree, counts, model)
    calc.set_root_frequencies(np.ones(N + 1) / (N + 1))
    rows, node_set = calc.branch_parameter_sets()
    n_branches = 2 * taxa - 2
    assert rows.shape[0] == n_branches + 1
    ll = calc.calculate_total_log_likelihood()
    ref, _, _ = O.log_likelihood(tree, counts, N, O.default_params(), np.ones(N + 1) / (N + 1),
                                 branch_params=overrides)
    assert ll == pytest.approx(ref, rel=1e-9)
    ws = calc._het_engine._het_ws
    assert ws.P.shape[0] == n_branches          # one slot per (set, length) pair in use
