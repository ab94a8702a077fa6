"""GPU parity tests of the fused path (series coefficients + tensor-core apply kernel), called
through the C ABI, against the CPU oracle.

The apply kernel never writes P, so P parity is checked through its action on one-hot child
vectors: w for a child observed in state j is column j of the post-processed P (AR:153-154),
which is compared entry by entry (1e-12 absolute, north_star) with the oracle's P.
"""
import ctypes

import numpy as np
import pytest

from oracle import chromevol_oracle as O

pytestmark = pytest.mark.gpu


def _apply(rows, times, N, child_vectors=None, codes=None):
    """w[b, e, k, :] for every (vector b, time e, child k); children are one-hot columns 0..n-1
    unless `child_vectors` (K, n) / `codes` (K,) are given.  Returns (w, mode)."""
    import torch
    from chromevol_enhanced_b200 import _lib
    from chromevol_enhanced_b200.engine import ExpmWorkspace, _stream_ptr
    _lib.require_device()
    lib = _lib.load()
    dev = torch.device("cuda:0")
    n = N + 1
    rows = np.atleast_2d(np.asarray(rows, dtype=np.float64))
    times = np.asarray(times, dtype=np.float64)
    B, E = rows.shape[0], times.shape[0]
    if codes is None:
        codes = np.arange(n) if child_vectors is None else np.full(child_vectors.shape[0], -2)
    codes = np.asarray(codes, dtype=np.int32)
    K = codes.shape[0]
    n_nodes = E * K
    ws = ExpmWorkspace(n, E, B, dev)
    ws.params.copy_(torch.as_tensor(rows, device=dev))
    st = _stream_ptr()
    _lib.check(lib.cevb_build_q(_lib.ptr(ws.params), B, n, _lib.ptr(ws.pw), st), "build_q")
    tc = torch.empty((B, lib.cevb_taylor_doubles(n)), dtype=torch.float64, device=dev)
    _lib.check(lib.cevb_taylor_prepare(_lib.ptr(ws.pw), B, n, 0.0, _lib.ptr(ws.norms), _lib.ptr(ws.csc),
                                       _lib.ptr(ws.csc_cnt), _lib.ptr(tc), st), "taylor_prepare")
    ldp = ws.ldp
    t_node = torch.as_tensor(np.repeat(times, K), device=dev)
    leaf_code = torch.as_tensor(np.tile(codes, E), device=dev)
    node_vec = torch.zeros((B, n_nodes, ldp), dtype=torch.float64, device=dev)
    if child_vectors is not None:
        cv = torch.as_tensor(np.asarray(child_vectors, dtype=np.float64), device=dev)
        node_vec[:, :, :n] = cv.repeat(E, 1)[None]
    items = torch.arange(n_nodes, dtype=torch.int32, device=dev)
    w = torch.full((B, n_nodes, ldp), float("nan"), dtype=torch.float64, device=dev)
    wr = torch.ones((B, n_nodes, ldp), dtype=torch.float64, device=dev)
    mode = torch.zeros((B, n_nodes), dtype=torch.uint8, device=dev)
    leaf_only = int(bool((codes >= 0).all()))      # exercises the one-hot variant where it applies
    _lib.check(lib.cevb_taylor_apply(_lib.ptr(tc), _lib.ptr(ws.norms), B, n, n_nodes, _lib.ptr(items),
                                     n_nodes, leaf_only, _lib.ptr(t_node), _lib.ptr(leaf_code),
                                     _lib.ptr(node_vec), _lib.ptr(w), _lib.ptr(wr), _lib.ptr(mode), st),
               "taylor_apply")
    torch.cuda.synchronize()
    w = (w / wr)[:, :, :n].cpu().numpy().reshape(B, E, K, n)          # AR:154
    return w, mode.cpu().numpy().reshape(B, E, K), ws.norms.cpu().numpy()


def _host_plan(Q, times):
    """Host mirror of the device's per-pair decision (accounting.plan_pairs) from the oracle Q."""
    from chromevol_enhanced_b200 import accounting
    mu = max(0.0, -Q.diagonal().min())
    Qs = Q + mu * np.eye(Q.shape[0])
    nu = min(np.abs(Qs).sum(axis=0).max(), np.abs(Qs).sum(axis=1).max())
    kind, nc, s = accounting.plan_pairs(times, [nu], [mu])
    return kind[0], nc[0], s[0], nu, mu


@pytest.mark.parametrize("N", [35, 100])
def test_fused_p_columns_match_oracle(N):
    """Column j of post(expm(Q t)) through the apply kernel, config-3 parameter distribution."""
    from chromevol_enhanced_b200 import synthetic
    rows = synthetic.random_param_rows(12, seed=4, linear='positive')
    rows[::3, 8] = 7
    times = np.concatenate([[0.0, 1e-6], 1e-4 + np.random.default_rng(1).exponential(0.05, 5), [0.758285]])
    w, mode, norms = _apply(rows, times, N)
    n_fused = 0
    for b in range(rows.shape[0]):
        Q = O.build_q(synthetic.row_to_dict(rows[b]), N)
        kind, nc, sq, nu, mu = _host_plan(Q, times)
        assert norms[b, 0] == pytest.approx(np.abs(Q).sum(axis=0).max(), rel=1e-14)
        assert norms[b, 11] == pytest.approx(mu, rel=1e-14)
        assert norms[b, 12] == pytest.approx(nu, rel=1e-13)
        for e, t in enumerate(times):
            # 1: identity (t == 0, AR:146-148); 2: routed to the scaling-and-squaring kernel
            assert (mode[b, e] == kind[e]).all(), (b, t, kind[e])
            if kind[e] == 0:
                P = w[b, e].T                           # w[k] = column k of P
                ref = O.transition_probabilities(Q, t)
                assert np.abs(P - ref).max() <= 1e-12, (b, t, nu * t)
                n_fused += 1
    assert n_fused > 40


def test_fused_general_variant_on_one_hot_children():
    """The same columns through the variant that stages child vectors (leaf_only = 0)."""
    from chromevol_enhanced_b200 import synthetic
    rows = synthetic.random_param_rows(3, seed=4, linear='positive')
    rows[:, :5] *= 0.3
    times = np.array([0.002, 0.03, 0.09])
    N = 100
    eye = np.eye(N + 1)
    w, mode, norms = _apply(rows, times, N, child_vectors=eye, codes=np.full(N + 1, -2))
    for b in range(rows.shape[0]):
        Q = O.build_q(synthetic.row_to_dict(rows[b]), N)
        for e, t in enumerate(times):
            if mode[b, e, 0] != 0:
                continue
            assert np.abs(w[b, e].T - O.transition_probabilities(Q, t)).max() <= 1e-12


def test_fused_p_sweep_up_to_the_switch_over():
    """Series length selection: x = nu t swept from 1e-6 up to where 64 terms stop sufficing."""
    from chromevol_enhanced_b200 import synthetic
    row = synthetic.default_param_row()
    row[:5] = [2.0, 1.5, 0.7, 0.4, 0.3]
    Q = O.build_q(synthetic.row_to_dict(row), 100)
    _, _, _, nu, mu = _host_plan(Q, np.array([1.0]))
    xs = np.array([1e-6, 1e-3, 0.05, 0.3, 0.9, 1.7, 2.5, 3.3, 4.1, 4.9, 5.7, 6.5, 7.5, 8.5, 9.3, 10.5, 13.0,
                   14.5, 15.4, 17.0, 24.0])
    times = xs / nu
    kind, nc, sq, _, _ = _host_plan(Q, times)
    assert kind[0] == 0 and kind[-1] == 2 and nc[kind == 0].max() >= 60
    w, mode, _ = _apply(row, times, 100)
    for e, t in enumerate(times):
        assert (mode[0, e] == kind[e]).all()
        if kind[e] == 0:
            ref = O.transition_probabilities(Q, t)
            assert np.abs(w[0, e].T - ref).max() <= 1e-12, xs[e]


def test_fused_non_generator_q():
    """linear_rate < 0 gives negative off-diagonal 'rates' (SURVEY.md 7.2-2): entries of the raw
    exponential exceed 1, so the tolerance scales with max|raw| as for the Pade path."""
    from chromevol_enhanced_b200 import synthetic
    rows = synthetic.random_param_rows(10, seed=9, linear='mixed')
    times = np.array([0.003, 0.02, 0.05])
    w, mode, norms = _apply(rows, times, 100)
    checked = 0
    for b in range(rows.shape[0]):
        Q = O.build_q(synthetic.row_to_dict(rows[b]), 100)
        for e, t in enumerate(times):
            if mode[b, e, 0] != 0:
                continue
            raw = O.raw_expm(Q, t)
            scale = max(1.0, np.abs(raw).max())
            ref = O.transition_probabilities(Q, t)
            assert np.abs(w[b, e].T - ref).max() <= 1e-12 * scale
            checked += 1
    assert checked > 10


def test_fused_child_vector_kinds_and_ragged_groups():
    """Internal (dense vector), missing-data (uniform) and observed children in one launch, with
    a branch count that is not a multiple of the 8-branch MMA group."""
    from chromevol_enhanced_b200 import synthetic
    rng = np.random.default_rng(3)
    N = 100
    rows = synthetic.random_param_rows(3, seed=11, linear='positive')
    rows[:, :5] *= 0.2
    times = np.array([0.01, 0.07, 0.2])
    K = 11
    vecs = rng.random((K, N + 1)) * 10.0 ** rng.integers(-30, 1, size=(K, 1))
    codes = np.array([-2, -2, -1, 5, -2, 100, -2, -1, 0, -2, 37], dtype=np.int32)
    w, mode, norms = _apply(rows, times, N, child_vectors=vecs, codes=codes)
    for b in range(rows.shape[0]):
        Q = O.build_q(synthetic.row_to_dict(rows[b]), N)
        for e, t in enumerate(times):
            if mode[b, e, 0] != 0:
                continue
            P = O.transition_probabilities(Q, t)
            for k in range(K):
                L = vecs[k] if codes[k] == -2 else (np.full(N + 1, 1.0 / (N + 1)) if codes[k] == -1
                                                    else O.leaf_vector(int(codes[k]), N))
                ref = P @ L
                assert np.allclose(w[b, e, k], ref, rtol=1e-11, atol=1e-12 * np.abs(L).max())


@pytest.mark.parametrize("N", [7, 40, 103, 104, 120, 256])
def test_fused_odd_sizes(N):
    """Padding columns, multi-pass rows (n > 104) and the one-group-per-warp variant (n = 257)."""
    from chromevol_enhanced_b200 import synthetic
    rows = np.stack([synthetic.default_param_row(), synthetic.default_param_row(base_number=3)])
    rows[1, :5] = [0.8, 0.6, 0.3, 0.1, 0.05]
    times = np.array([0.004, 0.11, 0.6])
    w, mode, norms = _apply(rows, times, N)
    for b in range(2):
        Q = O.build_q(synthetic.row_to_dict(rows[b]), N)
        for e, t in enumerate(times):
            if mode[b, e, 0] == 2:
                continue
            ref = O.transition_probabilities(Q, t)
            assert np.abs(w[b, e].T - ref).max() <= 1e-12, (N, b, t)


def _series_matrices(rows, times, N):
    """post(expm(Q_b t_e)) for the pairs that need squarings through cevb_expm_plan +
    cevb_taylor_matrices; returns (P[slot], slot[b, e], squarings[slot], nan_count)."""
    import torch
    from chromevol_enhanced_b200 import _lib
    from chromevol_enhanced_b200.engine import ExpmWorkspace, _stream_ptr
    _lib.require_device()
    lib = _lib.load()
    dev = torch.device("cuda:0")
    n = N + 1
    rows = np.atleast_2d(np.asarray(rows, dtype=np.float64))
    times = np.asarray(times, dtype=np.float64)
    B, E = rows.shape[0], times.shape[0]
    ws = ExpmWorkspace(n, E, B, dev, p_slots=B * E, n_nodes=1)
    ws.params.copy_(torch.as_tensor(rows, device=dev))
    st = _stream_ptr()
    _lib.check(lib.cevb_build_q(_lib.ptr(ws.params), B, n, _lib.ptr(ws.pw), st), "build_q")
    _lib.check(lib.cevb_taylor_prepare(_lib.ptr(ws.pw), B, n, 0.0, _lib.ptr(ws.norms), _lib.ptr(ws.csc),
                                       _lib.ptr(ws.csc_cnt), _lib.ptr(ws.tc), st), "taylor_prepare")
    d_t = torch.as_tensor(times, device=dev)
    order = torch.as_tensor(np.argsort(-times, kind="stable").astype(np.int32), device=dev)
    _lib.check(lib.cevb_expm_plan(_lib.ptr(ws.norms), _lib.ptr(d_t), B, E, B * E, _lib.ptr(ws.slot),
                                  _lib.ptr(ws.worklist), _lib.ptr(ws.work_s), _lib.ptr(ws.work_count),
                                  _lib.ptr(ws.plan_nc), _lib.ptr(ws.plan_tau), _lib.ptr(ws.plan_a0),
                                  st), "plan")
    ws.P.fill_(float("nan"))
    _lib.check(lib.cevb_taylor_matrices(_lib.ptr(ws.tc), _lib.ptr(ws.norms), _lib.ptr(order), B, E, n,
                                        _lib.ptr(ws.slot), _lib.ptr(ws.work_s), _lib.ptr(ws.work_count),
                                        B * E, _lib.ptr(ws.plan_nc), _lib.ptr(ws.plan_tau),
                                        _lib.ptr(ws.plan_a0), _lib.ptr(ws.P), _lib.ptr(ws.counter),
                                        _lib.ptr(ws.nan_count),
                                        _lib.ptr(ws.sq_scratch[0]) if ws.sq_scratch is not None else _lib.ptr(None),
                                        st), "matrices")
    torch.cuda.synchronize()
    cnt = int(ws.work_count[0])
    return (ws.P[:cnt, :, :n].cpu().numpy(), ws.slot.cpu().numpy(), ws.work_s[:cnt].cpu().numpy(),
            int(ws.nan_count[0]))


@pytest.mark.parametrize("N", [35, 100, 111, 140, 256])
def test_series_squaring_matrices_match_oracle(N):
    """Pairs beyond the switch-over: series at t 2^-s, s squarings, post-processing.  N <= 111
    squares in shared memory, the larger orders through global memory."""
    from chromevol_enhanced_b200 import synthetic
    rows = synthetic.random_param_rows(8, seed=4, linear='positive')
    rows[:, :5] *= 4.0
    rows[1, 8] = 5
    times = np.array([0.0, 1e-4, 0.011, 0.06, 0.2, 0.45, 0.758285, 1.4, 3.0])
    P, slot, sq, n_nan = _series_matrices(rows, times, N)
    assert n_nan == 0
    n_checked = 0
    for b in range(rows.shape[0]):
        Q = O.build_q(synthetic.row_to_dict(rows[b]), N)
        kind, nc, s_host, nu, mu = _host_plan(Q, times)
        for e, t in enumerate(times):
            if kind[e] == 2:
                k = slot[b, e]
                assert k >= 0
                assert sq[k] == s_host[e] and sq[k] >= 1
                ref = O.transition_probabilities(Q, t)
                assert np.abs(P[k] - ref).max() <= 1e-12, (b, t, nu * t, sq[k])
                n_checked += 1
            else:
                assert slot[b, e] == -1
    assert n_checked >= 20


@pytest.mark.parametrize("squaring", ["series", "pade"])
def test_large_order_squaring_paths(squaring):
    """n = 257: two matrices do not fit in shared memory; the series path squares through global
    memory (the default), the Pade kernels remain selectable.  Both against the oracle."""
    from chromevol_enhanced_b200 import _lib, synthetic
    from chromevol_enhanced_b200.engine import LikelihoodEngine
    from chromevol_enhanced_b200.flatten import FlatTree
    assert _lib.load().cevb_taylor_matrices_supported(101) == 1
    assert _lib.load().cevb_taylor_matrices_supported(257) == 1
    assert _lib.load().cevb_taylor_scratch_doubles(101) == 0
    assert _lib.load().cevb_taylor_scratch_doubles(257) > 0
    tree = synthetic.random_tree(30, seed=41, mean_len=0.15)
    counts = synthetic.simulate_tip_counts(tree, 256, seed=42, root_state=40)
    rows = synthetic.random_param_rows(3, seed=43, linear='positive')
    rows[:, 7] *= 0.1
    flat = FlatTree(tree, counts, 256)
    eng = LikelihoodEngine(flat, path="fused", squaring=squaring)
    assert eng.squaring == squaring
    a = eng.log_likelihood(rows).cpu().numpy()
    assert int(eng._materialised) > 0, "the case is meant to exercise the squaring path"
    for b in range(rows.shape[0]):
        ref, _, _ = O.log_likelihood(tree, counts, 256, synthetic.row_to_dict(rows[b]))
        assert a[b] == pytest.approx(ref, rel=1e-9), b


def test_series_squaring_overflow_falls_back_to_identity():
    """Optimiser bound corner with linear_rate = -1: expm overflows, the reference finds NaN
    after renormalising and returns the identity (AR:156-159)."""
    from chromevol_enhanced_b200 import synthetic
    row = synthetic.default_param_row()
    row[:8] = [10, 10, 10, 10, 10, 10, 10, -1.0]
    times = np.array([0.05, 0.758])
    P, slot, sq, n_nan = _series_matrices(row, times, 100)
    Q = O.build_q(synthetic.row_to_dict(row), 100)
    for e, t in enumerate(times):
        ref = O.transition_probabilities(Q, t)
        k = slot[0, e]
        assert k >= 0
        if np.array_equal(ref, np.eye(101)):
            assert np.array_equal(P[k], np.eye(101))
        else:
            scale = max(1.0, np.abs(O.raw_expm(Q, t)).max())
            assert np.abs(P[k] - ref).max() <= 1e-12 * scale
    assert n_nan >= 1


@pytest.mark.parametrize("squaring", ["series", "pade"])
def test_fused_engine_squaring_variants_agree(squaring):
    from chromevol_enhanced_b200 import synthetic
    from chromevol_enhanced_b200.engine import LikelihoodEngine
    from chromevol_enhanced_b200.flatten import FlatTree
    tree = synthetic.random_tree(120, seed=21, mean_len=0.2)
    counts = synthetic.simulate_tip_counts(tree, 50, seed=22)
    rows = synthetic.random_param_rows(10, seed=23, linear='mixed', corners=2)
    rows[:, :5] *= 2.0
    flat = FlatTree(tree, counts, 50)
    eng = LikelihoodEngine(flat, path="fused", squaring=squaring)
    a = eng.log_likelihood(rows).cpu().numpy()
    for b in range(rows.shape[0]):
        ref, _, resc = O.log_likelihood(tree, counts, 50, synthetic.row_to_dict(rows[b]))
        if np.isfinite(ref):
            assert a[b] == pytest.approx(ref, rel=1e-9), b
        else:
            assert a[b] == ref or (np.isnan(a[b]) and np.isnan(ref))
        assert int(eng.rescale_counts[b]) == resc


def test_fused_and_materialised_paths_agree_on_config3_batch():
    """1,000 taxa / N=100, 16 vectors of the bench distribution: both device paths against each
    other (1e-11 relative) and three of them against the oracle (1e-9 relative)."""
    from chromevol_enhanced_b200 import synthetic
    from chromevol_enhanced_b200.engine import LikelihoodEngine
    from chromevol_enhanced_b200.flatten import FlatTree
    tree = synthetic.random_tree(1000, seed=1)
    counts = synthetic.simulate_tip_counts(tree, 100, seed=2)
    rows = synthetic.random_param_rows(16, seed=4, linear='positive')
    flat = FlatTree(tree, counts, 100)
    fused = LikelihoodEngine(flat, path="fused")
    mat = LikelihoodEngine(flat, path="materialised", max_workspace_bytes=8 << 30)
    a = fused.log_likelihood(rows).cpu().numpy()
    b = mat.log_likelihood(rows).cpu().numpy()
    assert np.allclose(a, b, rtol=1e-11, atol=0)
    assert np.array_equal(fused.rescale_counts.cpu().numpy(), mat.rescale_counts.cpu().numpy())
    for k in (0, 5, 11):
        ref, _, resc = O.log_likelihood(tree, counts, 100, synthetic.row_to_dict(rows[k]))
        assert a[k] == pytest.approx(ref, rel=1e-9)


def test_fused_worklist_overflow_is_redone():
    """More (vector, branch) pairs need squarings than the P buffer has slots: the batch is
    re-run in smaller chunks and still matches the materialised path."""
    from chromevol_enhanced_b200 import synthetic
    from chromevol_enhanced_b200.engine import LikelihoodEngine
    from chromevol_enhanced_b200.flatten import FlatTree
    tree = synthetic.random_tree(80, seed=12, mean_len=0.3)
    counts = synthetic.simulate_tip_counts(tree, 40, seed=13)
    rows = synthetic.random_param_rows(12, seed=14, linear='positive')
    rows[:, :5] *= 10.0                     # nu t beyond the 64-term series on most branches
    flat = FlatTree(tree, counts, 40)
    # workspace so small that only one vector's worth of matrices fits
    fused = LikelihoodEngine(flat, path="fused", chunk=6, max_workspace_bytes=1 << 20)
    mat = LikelihoodEngine(flat, path="materialised")
    a = fused.log_likelihood(rows).cpu().numpy()
    b = mat.log_likelihood(rows).cpu().numpy()
    assert fused.overflow_events >= 1
    assert np.allclose(a, b, rtol=1e-11, atol=0)


def test_results_do_not_depend_on_batch_composition():
    """A vector's logL is bit-identical alone (graph replay), inside a small batch, inside a batch
    that is split over stream slices, and inside a chunked batch: groups are formed over the
    branches of one vector, never across vectors."""
    import torch
    from chromevol_enhanced_b200 import synthetic
    from chromevol_enhanced_b200.engine import LikelihoodEngine
    from chromevol_enhanced_b200.flatten import FlatTree
    tree = synthetic.random_tree(150, seed=51)
    counts = synthetic.simulate_tip_counts(tree, 60, seed=52)
    flat = FlatTree(tree, counts, 60)
    rows = synthetic.random_param_rows(70, seed=53, linear='mixed')
    eng = LikelihoodEngine(flat)
    full = eng.log_likelihood(rows).cpu().numpy()                     # 4 stream slices
    again = eng.log_likelihood(rows).cpu().numpy()
    assert np.array_equal(full, again)
    small = eng.log_likelihood(rows[:9]).cpu().numpy()                # CUDA-graph path
    assert np.array_equal(small, full[:9])
    for k in (0, 33, 69):
        one = eng.log_likelihood(rows[k:k + 1]).cpu().numpy()
        assert np.array_equal(one, full[k:k + 1])
    chunked = LikelihoodEngine(flat, chunk=16).log_likelihood(rows).cpu().numpy()
    assert np.array_equal(chunked, full)
    single_stream = LikelihoodEngine(flat, n_streams=1).log_likelihood(rows).cpu().numpy()
    assert np.array_equal(single_stream, full)
