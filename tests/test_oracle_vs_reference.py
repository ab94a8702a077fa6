"""CPU, build container only: the oracle against the UNMODIFIED reference module imported live
from /root/reference (skipped on the GPU box, where that tree does not exist)."""
import numpy as np
import pytest

from refload import load_reference, reference_available
from oracle import chromevol_oracle as O

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def AR():
    return load_reference()


def test_q_and_p_live(AR):
    rng = np.random.default_rng(0)
    for N in (12, 35, 100):
        for trial in range(4):
            kw = dict(gain=float(10 ** rng.uniform(-3, 1)), loss=float(10 ** rng.uniform(-3, 1)),
                      dupl=float(10 ** rng.uniform(-3, 1)), demidupl_gain=float(rng.uniform(0, 1)),
                      linear_rate=float(rng.uniform(-0.1, 0.1)),
                      base_number=[None, 3, 7][trial % 3])
            m = AR.ChromEvolutionModel(N, custom_params=kw)
            Q = O.build_q(O.default_params(**kw), N)
            assert np.array_equal(m.build_transition_rate_matrix(), Q)
            for t in (0.0, 0.01, 0.4):
                assert np.array_equal(m.get_transition_probabilities(t), O.transition_probabilities(Q, t))


def test_tree_likelihood_live(AR):
    from ete3 import Tree
    from chromevol_enhanced_b200 import synthetic
    tree = synthetic.random_tree(40, seed=2)
    counts = synthetic.simulate_tip_counts(tree, 30, seed=3)
    del counts["t1"]
    nw = tree.write(dist_formatter="%.17g")
    for kw in ({}, {"base_number": 4}, {"linear_rate": -0.04, "gain": 2.0}):
        t_ref = Tree(nw, format=1)
        model = AR.ChromEvolutionModel(30, custom_params=kw)
        calc = AR.ChromEvolLikelihoodCalculator(t_ref, counts, model)
        calc.set_root_frequencies(np.ones(31) / 31)
        ll_ref = calc.calculate_total_log_likelihood()
        calc.ancestral_state_reconstruction()
        t_or = Tree(nw, format=1)
        ll, vectors, _ = O.log_likelihood(t_or, counts, 30, O.default_params(**kw), np.ones(31) / 31)
        assert ll == ll_ref
        for a, b in zip(t_ref.traverse(), t_or.traverse()):
            assert np.array_equal(calc.likelihoods[id(a)], vectors[id(b)])
            assert O.ancestral_states(vectors[id(b)])[0] == a.ml_count


def test_objective_and_classify_live(AR):
    from ete3 import Tree
    nw = "((A:0.1,B:0.1):0.05,(C:0.1,D:0.1):0.05);"
    counts = {"A": 22, "B": 24, "C": 20, "D": 18}
    tree = Tree(nw, format=1)
    model = AR.ChromEvolutionModel(34)
    opt = AR.ChromEvolOptimizer(tree, counts, model)
    opt.calculator.set_root_frequencies(np.ones(35) / 35)
    names = AR.ChromEvolOptimizer.DEFAULT_PARAM_NAMES
    x = np.array([0.3, -1.0, 0.2, 0.0, 0.05, 0.01, 0.01, -0.02])
    f_ref = opt.objective_function(x, names)
    f = O.objective(x, names, O.default_params(), Tree(nw, format=1), counts, 34, np.ones(35) / 35)
    assert f == f_ref
    mapper = AR.StochasticMapper(tree, model)
    for a in range(0, 12):
        for b in range(0, 12):
            for bn in (None, 3):
                assert mapper._classify_event(a, b, {"base_number": bn}) == O.classify_event(a, b, bn)


def test_branch_specific_parameters_live(AR):
    """set_branch_parameters (AR:67-73) + branch_params.get(id(child), params) in the up-pass
    (AR:230-233): the oracle's branch_params argument against the reference itself."""
    from ete3 import Tree
    from chromevol_enhanced_b200 import synthetic
    tree = synthetic.random_tree(30, seed=5)
    counts = synthetic.simulate_tip_counts(tree, 24, seed=6)
    nw = tree.write(dist_formatter="%.17g")
    t_ref, t_or = Tree(nw, format=1), Tree(nw, format=1)
    model = AR.ChromEvolutionModel(24, custom_params={"gain": 0.4})
    ref_nodes = [nd for nd in t_ref.traverse() if not nd.is_root()]
    or_nodes = [nd for nd in t_or.traverse() if not nd.is_root()]
    overrides = {}
    for k, (a, b) in enumerate(zip(ref_nodes, or_nodes)):
        if k % 3 == 0:
            kw = {"gain": 1.5, "dupl": 0.3} if k % 2 == 0 else {"loss": 2.0, "base_number": 5}
            model.set_branch_parameters(id(a), **kw)
            overrides[id(b)] = O.default_params(**{"gain": 0.4, **kw})
    calc = AR.ChromEvolLikelihoodCalculator(t_ref, counts, model)
    calc.set_root_frequencies(np.ones(25) / 25)
    ll_ref = calc.calculate_total_log_likelihood()
    ll, vectors, _ = O.log_likelihood(t_or, counts, 24, O.default_params(gain=0.4), np.ones(25) / 25,
                                      branch_params=overrides)
    assert ll == ll_ref
    for a, b in zip(t_ref.traverse(), t_or.traverse()):
        assert np.array_equal(calc.likelihoods[id(a)], vectors[id(b)])
