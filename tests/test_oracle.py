"""CPU: the oracle (oracle/) against the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py) and against scipy itself.  No GPU, no /root/reference needed."""
import io

import numpy as np
import pytest
import scipy.linalg

from conftest import make_tree
from oracle import chromevol_oracle as O
from oracle.expm_amh import expm_amh

OPT_FULL = dict(gain=2.1342596548720283, loss=0.2677580853199203, dupl=2.016554428743558,
                demidupl_gain=0.38097654656835955, demidupl_loss=1.1888333123910255,
                base_number_gain=0.01, base_number_loss=0.01, linear_rate=-0.03935220722444711)
OPT_GL = dict(gain=1.9823905216545576, loss=4.424978107995602, dupl=0, demidupl_gain=0,
              demidupl_loss=0, base_number_gain=0, base_number_loss=0, linear_rate=0)
Q_CASES = {"default": {}, "bn7": {"base_number": 7}, "opt_full": OPT_FULL, "gl": OPT_GL,
           "neg_lr": {"linear_rate": -0.5}, "bn1": {"base_number": 1}, "bn2": {"base_number": 2}}


@pytest.mark.parametrize("tag", sorted(Q_CASES))
def test_q_bit_exact_against_reference_golden(golden, tag):
    arrays, _ = golden
    Q = O.build_q(O.default_params(**Q_CASES[tag]), 35)
    assert np.array_equal(Q, arrays[f"Q35_{tag}"])


@pytest.mark.parametrize("N", [100, 256])
def test_q_large_diag_and_rowsums(golden, N):
    arrays, _ = golden
    Q = O.build_q(O.default_params(base_number=5), N)
    assert np.array_equal(np.diag(Q), arrays[f"Q{N}_bn5_diag"])
    assert np.array_equal(np.abs(Q).sum(axis=1), arrays[f"Q{N}_bn5_rowsum_abs"])


def test_q_structure_kat7():
    Q = O.build_q(O.default_params(), 34)
    assert np.allclose(Q[5, 3:8], [0.005, 0.11, -0.28, 0.11, 0.005], atol=1e-15)
    assert not Q[0].any()
    P = O.transition_probabilities(Q, 0.1)
    assert abs(P[0, 0] - 0.999999999966) < 1e-12 and abs(P.min() - 9.99999999966e-13) < 1e-20


@pytest.mark.parametrize("tag", ["default", "opt_full", "gl"])
@pytest.mark.parametrize("t", [0.0, 1e-6, 0.00299, 0.05, 0.3, 0.758285])
def test_p_against_reference_golden(golden, tag, t):
    arrays, _ = golden
    Q = O.build_q(O.default_params(**Q_CASES[tag]), 35)
    P = O.transition_probabilities(Q, t)
    # golden was produced with scipy 1.18.1; allow a different scipy on another box
    assert np.abs(P - arrays[f"P35_{tag}_t{t}"]).max() <= 1e-13


def test_p_n100_rows(golden):
    arrays, _ = golden
    Q = O.build_q(O.default_params(**OPT_FULL), 100)
    for t in (0.01, 0.2):
        P = O.transition_probabilities(Q, t)
        assert np.abs(P[[0, 1, 20, 50, 100]] - arrays[f"P100_opt_full_t{t}_rows"]).max() <= 1e-12


def test_p_edge_semantics():
    Q = O.build_q(O.default_params(), 10)
    assert np.array_equal(O.transition_probabilities(Q, 0.0), np.eye(11))
    assert np.array_equal(O.transition_probabilities(Q, -3.0), np.eye(11))       # AR:143-148
    Qbad = Q.copy()
    Qbad[3, 4] = np.nan
    assert np.array_equal(O.transition_probabilities(Qbad, 0.1), np.eye(11))     # AR:156-166


KATS = ["kat1_default_N35", "kat2_N60_root24", "kat3_opt_full_N35", "kat4_gainloss_N35", "kat5_bn7_N35"]


@pytest.mark.parametrize("name", KATS)
def test_shipped_kats(golden, shipped, name):
    arrays, meta = golden
    newick, counts = shipped
    kat = meta["kat"][name]
    N = kat["max_chr"]
    tree = make_tree(newick)
    freq = None
    if kat["root_state"] is not None:
        freq = np.zeros(N + 1)
        freq[kat["root_state"]] = 1.0
    params = O.default_params(**(kat["params"] or {}))
    ll, vectors, rescales = O.log_likelihood(tree, counts, N, params, freq)
    assert ll == pytest.approx(kat["logL"], rel=1e-12)
    nodes = list(tree.traverse())
    V = np.stack([vectors[id(nd)] for nd in nodes])
    ref = arrays[f"{name}_vectors"]
    assert np.abs(V - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())
    states = [O.ancestral_states(vectors[id(nd)]) for nd in nodes]
    assert [s[0] for s in states] == arrays[f"{name}_ml_count"].tolist()
    assert np.allclose([s[1] for s in states], arrays[f"{name}_ml_probability"], rtol=1e-10)
    assert rescales == 0


def test_four_taxa_kat6(golden):
    arrays, meta = golden
    kat = meta["kat"]["kat6_four_taxa"]
    tree = make_tree(kat["newick"])
    ll, vectors, _ = O.log_likelihood(tree, kat["counts"], 34, O.default_params())
    assert ll == pytest.approx(-26.032378881391995, rel=1e-12)
    nodes = list(tree.traverse())
    assert [O.ancestral_states(vectors[id(nd)])[0] for nd in nodes] == arrays["kat6_ml_count"].tolist()
    assert [O.ancestral_states(vectors[id(nd)])[0] for nd in nodes][:3] == [22, 24, 20]


def test_edge_polytomy_missing_clamped(golden):
    arrays, meta = golden
    kat = meta["kat"]["edge_polytomy"]
    tree = make_tree(kat["newick"])
    ll, vectors, _ = O.log_likelihood(tree, kat["counts"], 12, O.default_params(**kat["params"]))
    assert ll == pytest.approx(kat["logL"], rel=1e-12)
    nodes = list(tree.traverse())
    V = np.stack([vectors[id(nd)] for nd in nodes])
    assert np.abs(V - arrays["edge_vectors"]).max() <= 1e-14
    # C = 99 is clamped to state 12 (AR:206-208); D is missing -> uniform, argmax 0 (AR:209-211)
    name = {nd.name: k for k, nd in enumerate(nodes)}
    assert V[name["C"], 12] == 1.0 and V[name["C"]].sum() == 1.0
    assert np.allclose(V[name["D"]], 1.0 / 13)


@pytest.mark.parametrize("tag", ["default", "opt_full"])
def test_synthetic_200_taxa(golden, tag):
    arrays, meta = golden
    kat = meta["kat"][f"synth200_{tag}"]
    tree = make_tree(kat["newick"])
    params = O.default_params(**(kat["params"] or {}))
    ll, vectors, _ = O.log_likelihood(tree, kat["counts"], 60, params)
    assert ll == pytest.approx(kat["logL"], rel=1e-11)
    nodes = list(tree.traverse())
    assert [O.ancestral_states(vectors[id(nd)])[0] for nd in nodes] == arrays[f"synth200_{tag}_ml_count"].tolist()


def test_rearrangement_rows_match_reference_csv(golden, shipped):
    import pandas as pd
    arrays, meta = golden
    newick, _ = shipped
    tree = make_tree(newick)
    for nd, c in zip(tree.traverse(), arrays["kat3_opt_full_N35_ml_count"]):
        nd.add_features(count=int(c))
    rows = O.rearrangement_rows(tree)
    ref = pd.read_csv(io.StringIO(meta["kat3_rearrangements_csv"]))
    assert len(rows) == len(ref) == 15
    got = pd.DataFrame(rows)
    assert got["branch_to_node"].tolist() == ref["branch_to_node"].tolist()
    assert got["event_type"].tolist() == ref["event_type"].tolist()
    assert got["change_magnitude"].tolist() == ref["change_magnitude"].tolist()
    assert got["count_from"].tolist() == ref["count_from"].tolist()


def test_aic_bic(golden):
    _, meta = golden
    ll = meta["kat"]["kat3_opt_full_N35"]["logL"]
    aic, bic = O.aic_bic(ll, 8, 38)
    assert aic == pytest.approx(126.2409, abs=1e-4) and bic == pytest.approx(139.3415, abs=1e-4)


def test_numpy_pairwise_sum_restatement():
    rng = np.random.default_rng(0)
    for n in list(range(1, 40)) + [101, 128, 129, 257, 300, 1000]:
        x = rng.standard_normal(n) * 10.0 ** rng.integers(-8, 8, n)
        assert O.numpy_pairwise_sum(x) == np.sum(x)


@pytest.mark.parametrize("N", [35, 100])
def test_expm_restatement_pins_scipy(N):
    rng = np.random.default_rng(3)
    seen = set()
    for trial in range(12):
        kw = dict(gain=float(10 ** rng.uniform(-3, 1)), loss=float(10 ** rng.uniform(-3, 1)),
                  dupl=float(10 ** rng.uniform(-3, 1)), linear_rate=float(rng.uniform(-0.02, 0.1)))
        Q = O.build_q(O.default_params(**kw), N)
        for t in (1e-4, 0.02, 0.3):
            X, m, s = expm_amh(Q * t, return_info=True)
            R = scipy.linalg.expm(Q * t)
            seen.add(m)
            assert np.abs(X - R).max() <= 1e-13 * max(1.0, np.abs(R).max())
    assert {3, 13} <= seen


def test_classify_event_table():
    assert O.classify_event(5, 6) == 'gain' and O.classify_event(5, 4) == 'loss'
    assert O.classify_event(5, 10) == 'duplication' and O.classify_event(1, 2) == 'gain'
    assert O.classify_event(5, 12, 7) == 'base_number_gain' and O.classify_event(9, 2, 7) == 'base_number_loss'
    assert O.classify_event(5, 7) == 'demi_gain_like' and O.classify_event(9, 7) == 'demi_loss_like'
    assert O.classify_event(4, 2) == 'other' and O.classify_event(2, 9) == 'other'


def test_simulator_statistics_match_reference_golden(golden, shipped):
    """The oracle simulator reproduces the reference's per-branch event means (z-test)."""
    arrays, meta = golden
    newick, counts = shipped
    tree = make_tree(newick)
    Q = O.build_q(O.default_params(**OPT_FULL), 35)
    rng = np.random.default_rng(5)
    R = 1500
    order, out = O.stochastic_mapping_counts(tree, Q, arrays["map_root_probs"], R, rng)
    nodes = list(tree.traverse())
    idx = {id(nd): k for k, nd in enumerate(nodes)}
    mean = out.mean(axis=0)
    Rref = meta["mapping"]["replicates"]
    worst = 0.0
    for b, nd in enumerate(order):
        k = idx[id(nd)]
        var = arrays["map_std"][k] ** 2
        se = np.sqrt(var / Rref + var / R + 1e-12)
        z = np.abs(mean[b] - arrays["map_mean"][k]) / np.maximum(se, 1e-3)
        worst = max(worst, z.max())
    assert worst < 5.0
