"""CPU: host-side logic -- Newick reader, tree flattening, parameter marshalling, histogram
statistics, and that the C-ABI library loads and exports every symbol the header declares."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, make_tree


def test_newick_format1_semantics(shipped):
    newick, counts = shipped
    tree = make_tree(newick)
    nodes = list(tree.traverse())
    assert len(nodes) == 74 and len(tree) == 38
    assert tree.name == "Deuterostomia" and [c.name for c in tree.children] == \
        ["Pecten_maximus", "Ambulacraria", "Branchiostoma_floridae"]
    assert tree.children[0].dist == pytest.approx(0.562046)
    # level-order is the default traversal and the reference CSV's row order
    assert [nd.name for nd in nodes][:4] == ["Deuterostomia", "Pecten_maximus", "Ambulacraria",
                                             "Branchiostoma_floridae"]
    t = make_tree("((A,B:0.5)X,(C:1e-3,'D d':2)[&&NHX:x=1]:0.25,E)R:7;")
    by = {nd.name: nd for nd in t.traverse()}
    assert by["A"].dist == 1.0 and by["B"].dist == 0.5 and by["X"].dist == 1.0
    assert by["D d"].dist == 2.0 and by["R"].dist == 7.0 and by["R"].is_root()
    assert [nd.name for nd in t.traverse("preorder")] == ["R", "X", "A", "B", "", "C", "D d", "E"]
    assert [nd.name for nd in t.traverse("postorder")] == ["A", "B", "X", "C", "D d", "", "E", "R"]
    assert [leaf.name for leaf in t] == ["A", "B", "C", "D d", "E"]
    again = make_tree(t.write())
    assert [nd.name for nd in again.traverse()] == [nd.name for nd in t.traverse()]
    with pytest.raises(ValueError):
        make_tree("((A,B);")


def test_flatten_levels_and_codes():
    from chromevol_enhanced_b200.flatten import FlatTree
    t = make_tree("((A:0.1,B:0.1):0.05,(C:0.1,D:0.1,E:0.3)X:0.05,F:0,G:-2,H);")
    f = FlatTree(t, {"A": 22, "B": 24, "C": 20, "D": 99, "F": 3.0, "G": float("nan"), "H": -1}, 34)
    assert f.n_nodes == 11 and f.root == 0 and f.parent[0] == -1
    names = [nd.name for nd in f.nodes]
    code = dict(zip(names, f.leaf_code.tolist()))
    assert code["A"] == 22 and code["D"] == 34 and code["E"] == -1 and code["F"] == 3
    assert code["G"] == 34 and code["H"] == 34 and code["X"] == -2       # AR:204-211
    # children precede parents level by level; file order kept in the CSR lists
    seen = set(np.flatnonzero(f.leaf_code != -2).tolist())
    for lv in range(f.n_levels):
        for node in f.level_nodes[f.level_off[lv]:f.level_off[lv + 1]]:
            kids = f.child_idx[f.child_off[node]:f.child_off[node + 1]]
            assert set(kids.tolist()) <= seen
        seen |= set(f.level_nodes[f.level_off[lv]:f.level_off[lv + 1]].tolist())
    assert f.level_nodes[-1] == 0
    # G has a negative length -> 1e-6 for the likelihood (AR:224-227); H's missing length is 1.0
    g, h = names.index("G"), names.index("H")
    assert f.t_unique[f.branch_of[g]] == 1e-6 and f.t_unique[f.branch_of[h]] == 1.0
    assert f.dist_raw[g] == -2.0
    assert f.preorder[0] == 0 and sorted(f.preorder.tolist()) == list(range(11))
    assert f.n_t == len(set(f.dist_eff[1:].tolist()))


def test_params_to_row_layout():
    from chromevol_enhanced_b200.engine import params_to_row, PARAM_ORDER
    from chromevol_enhanced_b200 import MODEL_PARAMETER_NAMES
    assert PARAM_ORDER[:8] == MODEL_PARAMETER_NAMES
    row = params_to_row({'gain': 1, 'loss': 2, 'dupl': 3, 'demidupl_gain': 4, 'demidupl_loss': 5,
                         'base_number': None, 'base_number_gain': 6, 'base_number_loss': 7,
                         'linear_rate': -0.5})
    assert row.tolist() == [1, 2, 3, 4, 5, 6, 7, -0.5, 0]
    assert params_to_row({'base_number': 7})[8] == 7.0


def test_percentiles_from_histogram_match_numpy():
    from chromevol_enhanced_b200.mapper import _percentile_from_hist
    rng = np.random.default_rng(0)
    for _ in range(300):
        x = rng.poisson(rng.uniform(0.01, 3), size=rng.integers(1, 400))
        h = np.bincount(x, minlength=32)
        for q in (2.5, 50, 97.5):
            assert _percentile_from_hist(h, len(x), q) == np.percentile(x, q)


def test_abi_library_loads_and_exports_header_symbols():
    from chromevol_enhanced_b200 import _build, _lib
    _build.build()
    lib = ctypes.CDLL(_build.LIB_PATH)
    header = open(os.path.join(ROOT, "include", "chromevol_b200.h")).read()
    declared = set(re.findall(r"\b(cevb_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.HEADER_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    typed = _lib.load()
    assert typed.cevb_abi_version() == 2
    assert typed.cevb_ldq(101) == 102 and typed.cevb_ldp(36) == 36


def test_product_path_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import chromevol_enhanced_b200 as cev
    from chromevol_enhanced_b200._lib import CevbError
    with pytest.raises(CevbError):
        cev.ChromEvolutionModel(10).build_transition_rate_matrix()
    t = make_tree("(A:1,B:1);")
    calc = cev.ChromEvolLikelihoodCalculator(t, {"A": 3, "B": 4}, cev.ChromEvolutionModel(10))
    with pytest.raises(CevbError):
        calc.calculate_total_log_likelihood()
    # device failures are not numerical penalties: the optimiser and model selection raise
    # instead of reporting a "converged" fit at -1e10 (round-1 advisor finding)
    opt = cev.ChromEvolOptimizer(t, {"A": 3, "B": 4}, cev.ChromEvolutionModel(10))
    with pytest.raises(CevbError):
        opt.objective_function(np.array([0.1, 0.1]), ["gain", "loss"])
    with pytest.raises(CevbError):
        opt.optimize_parameters()
    with pytest.raises(CevbError):
        cev.calculate_model_selection_criteria(t, {"A": 3, "B": 4}, 10)


def test_objective_keeps_numerical_penalties(monkeypatch):
    """AR:365-371: a numerical exception inside the evaluation is a 1e10 penalty, inf / NaN a 1e9
    one -- only device errors propagate."""
    import chromevol_enhanced_b200 as cev
    t = make_tree("(A:1,B:1);")
    opt = cev.ChromEvolOptimizer(t, {"A": 3, "B": 4}, cev.ChromEvolutionModel(10))
    monkeypatch.setattr(opt.calculator, "calculate_total_log_likelihood",
                        lambda: (_ for _ in ()).throw(FloatingPointError("overflow")))
    assert opt.objective_function(np.array([0.1, 0.1]), ["gain", "loss"]) == 1e10
    monkeypatch.setattr(opt.calculator, "calculate_total_log_likelihood", lambda: float("nan"))
    assert opt.objective_function(np.array([0.1, 0.1]), ["gain", "loss"]) == 1e9
    monkeypatch.setattr(opt.calculator, "calculate_total_log_likelihood", lambda: -np.inf)
    assert opt.objective_function(np.array([-3.0, 0.1]), ["gain", "loss"]) == 1e9
    assert opt.model.params["gain"] == 1e-7                                   # AR:352 clamp


def test_max_chromosome_number_limit_is_reported():
    """The kernels support at most 288 states (CEVB_MAX_N); a larger model says so instead of
    failing inside a launch."""
    import chromevol_enhanced_b200 as cev
    from chromevol_enhanced_b200._lib import CevbError, MAX_STATES
    header = open(os.path.join(ROOT, "include", "chromevol_b200.h")).read()
    assert int(re.search(r"#define CEVB_MAX_N (\d+)", header).group(1)) == MAX_STATES
    t = make_tree("(A:1,B:1);")
    calc = cev.ChromEvolLikelihoodCalculator(t, {"A": 3, "B": 4}, cev.ChromEvolutionModel(MAX_STATES))
    with pytest.raises(CevbError, match="at most"):
        calc.calculate_total_log_likelihood()


def test_lazy_memo_behaves_like_a_dict():
    from chromevol_enhanced_b200.ancestral_reconstruction import _NodeVectors
    calls = []

    def fill():
        calls.append(1)
        return {1: np.arange(3.0), 2: np.ones(3)}

    memo = _NodeVectors(fill, -1.5)
    assert memo.pending and memo.logl == -1.5 and not calls
    assert 1 in memo and not memo.pending and len(calls) == 1
    assert len(memo) == 2 and sorted(memo) == [1, 2] and memo.get(3) is None
    assert np.array_equal(memo[1], np.arange(3.0)) and len(calls) == 1
    assert bool(memo) and list(memo.keys()) == [1, 2] and len(list(memo.items())) == 2
    empty = _NodeVectors(lambda: {}, 0.0)
    assert not empty and not empty.pending


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "chromevol_enhanced_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+\.*oracle", src, flags=re.M), fn
            assert "import oracle" not in src and "from oracle" not in src, fn


def test_synthetic_generators_are_seeded():
    from chromevol_enhanced_b200 import synthetic
    a, b = synthetic.random_tree(50, seed=3), synthetic.random_tree(50, seed=3)
    assert a.write() == b.write() and len(a) == 50
    assert len({nd.dist for nd in a.traverse()}) == 99
    r1, r2 = synthetic.random_param_rows(8, 4, corners=2), synthetic.random_param_rows(8, 4, corners=2)
    assert np.array_equal(r1, r2) and abs(r1[-1, 7]) == 1.0


def test_branch_parameter_sets_groups_identical_overrides():
    """Host logic of SURVEY 8 row f4: distinct parameter sets and the per-node set index follow
    ``branch_params.get(id(node), params)`` (AR:230-233); set 0 is the model-wide set and equal
    override dicts share a set."""
    from chromevol_enhanced_b200.engine import branch_parameter_sets, params_to_row
    from chromevol_enhanced_b200.flatten import FlatTree
    tree = make_tree("((A:0.1,B:0.2)X:0.3,(C:0.1,D:0.4)Y:0.2,E:0.5);")
    flat = FlatTree(tree, {"A": 3, "B": 4, "C": 5, "D": 6, "E": 7}, 12)

    class Model:
        params = {"gain": 0.1, "loss": 0.2, "dupl": 0.0, "demidupl_gain": 0.0, "demidupl_loss": 0.0,
                  "base_number_gain": 0.0, "base_number_loss": 0.0, "linear_rate": 0.0,
                  "base_number": None}
        branch_params = {}

    by_name = {nd.name: nd for nd in flat.nodes}
    m = Model()
    rows, node_set = branch_parameter_sets(m, flat)
    assert rows.shape == (1, 9) and not node_set.any()
    fast = dict(m.params, gain=2.0)
    other = dict(m.params, loss=1.0, base_number=4)
    m.branch_params = {id(by_name["A"]): fast, id(by_name["Y"]): dict(fast), id(by_name["E"]): other,
                       id(by_name["C"]): dict(m.params)}            # an override equal to the default
    rows, node_set = branch_parameter_sets(m, flat)
    assert rows.shape == (3, 9)
    assert np.array_equal(rows[0], params_to_row(m.params))
    assert np.array_equal(rows[1], params_to_row(fast)) and np.array_equal(rows[2], params_to_row(other))
    idx = {nd.name: k for k, nd in enumerate(flat.nodes)}
    assert node_set[idx["A"]] == 1 and node_set[idx["Y"]] == 1 and node_set[idx["E"]] == 2
    assert node_set[idx["C"]] == 0 and node_set[idx["B"]] == 0 and node_set[idx["X"]] == 0
    assert rows[2, 8] == 4.0 and rows[0, 8] == 0.0
