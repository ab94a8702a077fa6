"""GPU parity tests proper: every stage of the CUDA path, called through the C ABI, against the
CPU oracle (oracle/) on the same seeded inputs and against the committed golden vectors.

Tolerances (BASELINE.json north_star): Q bit-exact; post-processed P within 1e-12 absolute;
log-likelihood within 1e-9 relative; MAP counts and event calls bit-exact.
"""
import io

import numpy as np
import pytest

from conftest import make_tree
from oracle import chromevol_oracle as O
from oracle.expm_amh import expm_amh

pytestmark = pytest.mark.gpu

OPT_FULL = dict(gain=2.1342596548720283, loss=0.2677580853199203, dupl=2.016554428743558,
                demidupl_gain=0.38097654656835955, demidupl_loss=1.1888333123910255,
                base_number_gain=0.01, base_number_loss=0.01, linear_rate=-0.03935220722444711)
OPT_GL = dict(gain=1.9823905216545576, loss=4.424978107995602, dupl=0, demidupl_gain=0,
              demidupl_loss=0, base_number_gain=0, base_number_loss=0, linear_rate=0)


def _row(params):
    from chromevol_enhanced_b200.engine import params_to_row
    return params_to_row(params)


# ------------------------------------------------------------------------------ rate matrix
@pytest.mark.parametrize("N", [5, 35, 100, 127, 128, 256])
def test_q_bit_exact(N):
    from chromevol_enhanced_b200.engine import rate_matrix
    cases = [{}, {"base_number": 7}, {"base_number": 1}, {"base_number": 2}, OPT_FULL, OPT_GL,
             {"linear_rate": -0.5}, {"demidupl_gain": 0.0}, {"demidupl_loss": -1.0},
             {"gain": 10.0, "loss": 1e-7, "dupl": 3.3, "linear_rate": 1.0, "base_number": 3}]
    rows = np.stack([_row(O.default_params(**kw)) for kw in cases])
    Q = rate_matrix(rows, N + 1)
    for k, kw in enumerate(cases):
        ref = O.build_q(O.default_params(**kw), N)
        assert np.array_equal(Q[k], ref), (N, kw)


def test_q_golden(golden):
    from chromevol_enhanced_b200.engine import rate_matrix
    arrays, _ = golden
    for tag, kw in (("default", {}), ("bn7", {"base_number": 7}), ("opt_full", OPT_FULL)):
        assert np.array_equal(rate_matrix(_row(O.default_params(**kw)), 36)[0], arrays[f"Q35_{tag}"])


# ------------------------------------------------------------------------------ expm / P(t)
def _check_p(rows, times, N, atol=1e-12, full_pivot=0):
    from chromevol_enhanced_b200.engine import transition_matrices
    from chromevol_enhanced_b200 import synthetic
    P, m, s, flags = transition_matrices(rows, times, N + 1, return_info=True, full_pivot=full_pivot)
    worst = 0.0
    mism = 0
    for b in range(rows.shape[0]):
        Q = O.build_q(synthetic.row_to_dict(rows[b]), N)
        for e, t in enumerate(times):
            ref = O.transition_probabilities(Q, t)
            err = np.abs(P[b, e] - ref).max()
            worst = max(worst, err)
            assert err <= atol, (b, t, err, m[b, e], s[b, e], flags[b, e])
            if t > 0 and not (flags[b, e] & 2):
                _, m_ref, s_ref = expm_amh(Q * t, return_info=True)
                mism += (m_ref, s_ref) != (int(m[b, e]), int(s[b, e]))
    return worst, mism, flags


@pytest.mark.parametrize("N", [35, 100])
def test_p_matches_scipy_path_config3_distribution(N):
    from chromevol_enhanced_b200 import synthetic
    rows = synthetic.random_param_rows(24, seed=4, linear='positive')
    rows[::3, 8] = 7
    times = np.concatenate([[0.0, 1e-6], 1e-4 + np.random.default_rng(1).exponential(0.05, 6), [0.758285]])
    worst, mism, flags = _check_p(rows, times, N)
    assert mism == 0                      # same Pade order and squaring count as Algorithm 5.1
    assert not ((flags >> 0) & 4).any()   # growth monitor quiet


def test_p_non_generator_and_corners():
    """linear_rate < 0 gives negative 'rates' (SURVEY.md 7.2-2); optimiser bound corners give
    ||Qt|| in the thousands and, for lr = -1, overflow -> NaN -> identity (AR:156-159)."""
    from chromevol_enhanced_b200 import synthetic
    from chromevol_enhanced_b200.engine import transition_matrices
    rows = synthetic.random_param_rows(16, seed=9, linear='mixed', corners=6)
    rows[-1, :8] = [10, 10, 10, 10, 10, 10, 10, -1.0]      # overflows to NaN at t = 0.758
    times = np.array([0.003, 0.05, 0.758])
    P, m, s, flags = transition_matrices(rows, times, 101, return_info=True)
    n_identity = 0
    for b in range(rows.shape[0]):
        Q = O.build_q(synthetic.row_to_dict(rows[b]), 100)
        for e, t in enumerate(times):
            raw = O.raw_expm(Q, t)
            ref = O.transition_probabilities(Q, t)
            scale = max(1.0, np.nanmax(np.abs(raw))) if np.isfinite(raw).all() else 1.0
            if np.array_equal(ref, np.eye(101)):
                n_identity += 1
                assert np.array_equal(P[b, e], np.eye(101)) and (flags[b, e] & 2)
            else:
                assert np.abs(P[b, e] - ref).max() <= 1e-12 * scale, (b, t, scale)
    assert n_identity > 0


def test_p_shipped_branch_lengths_golden(golden, shipped):
    from chromevol_enhanced_b200.engine import transition_matrices
    arrays, _ = golden
    for tag, kw in (("default", {}), ("opt_full", OPT_FULL), ("gl", OPT_GL)):
        times = [0.0, 1e-6, 0.00299, 0.05, 0.3, 0.758285]
        P = transition_matrices(_row(O.default_params(**kw)), times, 36)
        for e, t in enumerate(times):
            assert np.abs(P[0, e] - arrays[f"P35_{tag}_t{t}"]).max() <= 1e-12
        assert np.array_equal(P[0, 0], np.eye(36))        # t == 0 -> exact identity


def test_p_full_pivot_path_agrees():
    from chromevol_enhanced_b200 import synthetic
    rows = synthetic.random_param_rows(6, seed=21, linear='mixed')
    times = np.array([0.01, 0.2, 0.7])
    _check_p(rows, times, 60, full_pivot=1)


@pytest.mark.parametrize("N", [7, 8, 63, 103, 104, 120, 256])
def test_p_odd_sizes_and_global_scratch_path(N):
    """N+1 not a multiple of 8, the largest shared-memory size (n=104..105) and the L2-scratch
    path (n > 104, config 4's n = 257)."""
    from chromevol_enhanced_b200 import synthetic
    rows = synthetic.random_param_rows(3, seed=N, linear='positive')
    rows[0] = synthetic.default_param_row()
    times = np.array([0.004, 0.09, 0.4])
    _check_p(rows, times, N)


def test_expm_of_general_dense_matrices():
    """The expm kernels do not rely on ChromEvol structure: random dense (non-sparse) inputs."""
    import scipy.linalg
    from chromevol_enhanced_b200.engine import transition_matrices
    rng = np.random.default_rng(2)
    n = 40
    A = rng.standard_normal((4, n, n)) * np.array([0.01, 0.1, 0.3, 0.6])[:, None, None]
    times = np.array([0.5, 1.0])
    P = transition_matrices(None, times, n, q_matrices=A)
    for b in range(4):
        for e, t in enumerate(times):
            ref = O.transition_probabilities(A[b], t)
            raw = scipy.linalg.expm(A[b] * t)
            assert np.abs(P[b, e] - ref).max() <= 1e-12 * max(1.0, np.abs(raw).max())


# ------------------------------------------------------------------------------ pruning
def _engine(tree, counts, N, **kw):
    from chromevol_enhanced_b200.engine import LikelihoodEngine
    from chromevol_enhanced_b200.flatten import FlatTree
    flat = FlatTree(tree, counts, N)
    return LikelihoodEngine(flat, **kw), flat


KATS = ["kat1_default_N35", "kat2_N60_root24", "kat3_opt_full_N35", "kat4_gainloss_N35", "kat5_bn7_N35"]


PATHS = ["fused", "materialised"]


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("name", KATS)
def test_shipped_kats_logl_vectors_states(golden, shipped, name, path):
    arrays, meta = golden
    newick, counts = shipped
    kat = meta["kat"][name]
    N = kat["max_chr"]
    tree = make_tree(newick)
    eng, flat = _engine(tree, counts, N, path=path)
    freq = None
    if kat["root_state"] is not None:
        freq = np.zeros(N + 1)
        freq[kat["root_state"]] = 1.0
    ll = eng.log_likelihood(_row(O.default_params(**(kat["params"] or {})))[None], root_freq=freq)
    assert float(ll[0]) == pytest.approx(kat["logL"], rel=1e-9)
    V = eng.node_vectors(0).cpu().numpy()
    ref = arrays[f"{name}_vectors"]
    assert np.abs(V - ref).max() <= 1e-11 * max(1.0, np.abs(ref).max())
    ml_count, ml_prob, probs = eng.ancestral_states(0)
    assert ml_count.cpu().numpy().tolist() == arrays[f"{name}_ml_count"].tolist()     # bit-exact
    assert np.allclose(ml_prob.cpu().numpy(), arrays[f"{name}_ml_probability"], rtol=1e-9)
    assert int(eng.rescale_counts[0]) == 0


@pytest.mark.parametrize("path", PATHS)
def test_four_taxa_and_edge_cases(golden, path):
    arrays, meta = golden
    for key, vec_key, cnt_key in (("kat6_four_taxa", "kat6_vectors", "kat6_ml_count"),
                                  ("edge_polytomy", "edge_vectors", "edge_ml_count")):
        kat = meta["kat"][key]
        tree = make_tree(kat["newick"])
        eng, flat = _engine(tree, kat["counts"], kat["max_chr"], path=path)
        ll = eng.log_likelihood(_row(O.default_params(**(kat.get("params") or {})))[None])
        assert float(ll[0]) == pytest.approx(kat["logL"], rel=1e-9)
        V = eng.node_vectors(0).cpu().numpy()
        assert np.abs(V - arrays[vec_key]).max() <= 1e-12
        assert eng.ancestral_states(0)[0].cpu().numpy().tolist() == arrays[cnt_key].tolist()


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("tag", ["default", "opt_full"])
def test_synthetic_200_taxa_golden(golden, tag, path):
    arrays, meta = golden
    kat = meta["kat"][f"synth200_{tag}"]
    tree = make_tree(kat["newick"])
    eng, flat = _engine(tree, kat["counts"], 60, path=path)
    ll = eng.log_likelihood(_row(O.default_params(**(kat["params"] or {})))[None])
    assert float(ll[0]) == pytest.approx(kat["logL"], rel=1e-9)
    assert eng.ancestral_states(0)[0].cpu().numpy().tolist() == arrays[f"synth200_{tag}_ml_count"].tolist()
    # root vector: entries that matter for logL to 1e-9 relative; the tiny ones (1e-9 of the
    # largest) are sums of P entries near the 1e-12 floor, where scipy's own expm carries about
    # 1e-17 absolute = 1e-5 relative error, so they are only compared to 1e-6
    got, ref = eng.node_vectors(0)[0].cpu().numpy(), arrays[f"synth200_{tag}_root_vector"]
    big = ref >= 1e-6 * ref.max()
    assert np.allclose(got[big], ref[big], rtol=1e-9, atol=0)
    assert np.allclose(got, ref, rtol=1e-6, atol=0)


@pytest.mark.parametrize("path", PATHS)
def test_batch_vs_oracle_live_and_chunking(path):
    """A batch of parameter vectors (incl. non-generator ones) on a seeded tree against the
    oracle, evaluated once in one chunk and once in ragged chunks of 3."""
    from chromevol_enhanced_b200 import synthetic
    tree = synthetic.random_tree(60, seed=5)
    counts = synthetic.simulate_tip_counts(tree, 40, seed=6)
    del counts["t3"]                       # one missing species
    counts["t4"] = 77                      # one out-of-range count
    rows = synthetic.random_param_rows(10, seed=8, linear='mixed')
    rows[4, 8] = 5
    eng, flat = _engine(tree, counts, 40, path=path)
    ll = eng.log_likelihood(rows).cpu().numpy()
    eng3, _ = _engine(tree, counts, 40, chunk=3, path=path)
    ll3 = eng3.log_likelihood(rows).cpu().numpy()
    assert np.array_equal(ll, ll3)
    for b in range(rows.shape[0]):
        ref, _, resc = O.log_likelihood(tree, counts, 40, synthetic.row_to_dict(rows[b]))
        assert ll[b] == pytest.approx(ref, rel=1e-9), b
        assert int(eng.rescale_counts[b]) == resc


@pytest.mark.parametrize("path", PATHS)
def test_rescale_branch_is_reproduced(path):
    """Large divergent tree: the AR:242 rescale (0 < sum < 1e-100, factor dropped) must fire on
    the same nodes as in the oracle, otherwise logL differs by hundreds."""
    from chromevol_enhanced_b200 import synthetic
    tree = synthetic.random_tree(400, seed=31, mean_len=0.4)
    counts = synthetic.simulate_tip_counts(tree, 60, seed=32, params=dict(gain=1.0, loss=1.0))
    row = synthetic.default_param_row()
    eng, flat = _engine(tree, counts, 60, path=path)
    ll = float(eng.log_likelihood(row[None])[0])
    ref, _, resc = O.log_likelihood(tree, counts, 60, synthetic.row_to_dict(row))
    assert resc > 0
    assert int(eng.rescale_counts[0]) == resc
    assert ll == pytest.approx(ref, rel=1e-9)


@pytest.mark.parametrize("path", PATHS)
def test_thousand_taxa_n100_single_vector(path):
    """BASELINE config 2 at full size, against the oracle (about 2 s of CPU)."""
    from chromevol_enhanced_b200 import synthetic
    tree = synthetic.random_tree(1000, seed=1)
    counts = synthetic.simulate_tip_counts(tree, 100, seed=2)
    row = synthetic.default_param_row()
    eng, flat = _engine(tree, counts, 100, path=path)
    ll = float(eng.log_likelihood(row[None])[0])
    ref, vectors, resc = O.log_likelihood(tree, counts, 100, synthetic.row_to_dict(row))
    assert ll == pytest.approx(ref, rel=1e-9)
    got = eng.ancestral_states(0)[0].cpu().numpy()
    want = np.array([O.ancestral_states(vectors[id(nd)])[0] for nd in flat.nodes])
    assert np.array_equal(got, want)
    m, s, flags = eng.order_info(0)
    assert not (flags & 4).any()


# ------------------------------------------------------------------------------ drop-in classes
def test_dropin_classes_on_shipped_data(golden, shipped):
    import chromevol_enhanced_b200 as cev
    arrays, meta = golden
    newick, counts = shipped
    tree = make_tree(newick)
    model = cev.ChromEvolutionModel(35, custom_params=OPT_FULL)
    assert np.array_equal(model.build_transition_rate_matrix(), arrays["Q35_opt_full"])
    assert np.abs(model.get_transition_probabilities(0.05) - arrays["P35_opt_full_t0.05"]).max() <= 1e-12
    assert np.array_equal(model.get_transition_probabilities(0.0), np.eye(36))
    calc = cev.ChromEvolLikelihoodCalculator(tree, counts, model)
    calc.set_root_frequencies(np.ones(36) / 36)
    ll = calc.calculate_total_log_likelihood()
    assert ll == pytest.approx(-55.12042715196926, rel=1e-9)
    calc.ancestral_state_reconstruction()
    nodes = list(tree.traverse())
    assert [int(nd.ml_count) for nd in nodes] == arrays["kat3_opt_full_N35_ml_count"].tolist()
    for nd in nodes:
        nd.add_features(count=nd.ml_count)
    rows = O.rearrangement_rows(tree)
    import pandas as pd
    ref = pd.read_csv(io.StringIO(meta["kat3_rearrangements_csv"]))
    assert [r["branch_to_node"] for r in rows] == ref["branch_to_node"].tolist()
    assert [r["change_magnitude"] for r in rows] == ref["change_magnitude"].tolist()
    assert [r["event_type"] for r in rows] == ref["event_type"].tolist()


def test_objective_penalties_and_batch(shipped):
    import chromevol_enhanced_b200 as cev
    newick, counts = shipped
    tree = make_tree(newick)
    model = cev.ChromEvolutionModel(35)
    opt = cev.ChromEvolOptimizer(tree, counts, model)
    opt.calculator.set_root_frequencies(np.ones(36) / 36)
    names = cev.ChromEvolOptimizer.DEFAULT_PARAM_NAMES
    x0 = np.array([model.params[p] for p in names])
    f0 = opt.objective_function(x0, names)
    assert f0 == pytest.approx(84.77969399785866, rel=1e-9)
    x = x0.copy()
    x[0] = -5.0                                   # clamped to 1e-7 (AR:352)
    f1 = opt.objective_function(x, names)
    assert model.params['gain'] == 1e-7
    ref = O.objective(x, names, O.default_params(), tree, counts, 35, np.ones(36) / 36)
    assert f1 == pytest.approx(ref, rel=1e-9)
    fb = opt.objective_batch(np.stack([x0, x]), names)
    assert fb[0] == pytest.approx(f0, rel=1e-12) and fb[1] == pytest.approx(f1, rel=1e-12)


def test_optimizer_reaches_reference_optimum(shipped, golden):
    """L-BFGS-B from the defaults on the shipped data: the optimiser trajectory is not
    reproducible across implementations, the optimum's logL and the MAP counts are."""
    import chromevol_enhanced_b200 as cev
    arrays, _ = golden
    newick, counts = shipped
    tree = make_tree(newick)
    model, tree = cev.perform_chromevol_analysis(tree, counts, 35, optimize_params=True,
                                                 stochastic_mapping=False)
    calc = cev.ChromEvolLikelihoodCalculator(tree, counts, model)
    calc.set_root_frequencies(np.ones(36) / 36)
    ll = calc.calculate_total_log_likelihood()
    assert ll >= -55.12042715196926 - 0.05
    got = [int(nd.count) for nd in tree.traverse()]
    assert got == arrays["kat3_opt_full_N35_ml_count"].tolist()
