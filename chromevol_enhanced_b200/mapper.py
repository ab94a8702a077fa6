"""Host side of the Philox stochastic-mapping simulator (reference: StochasticMapper, AR:428-644).

The device kernel accumulates, per (branch, event type), a histogram of per-replicate event
counts plus exact integer sums; mean / std / percentiles / totals (AR:633-637) are derived from
those without ever materialising the per-replicate event lists the reference keeps in RAM.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from .engine import branch_parameter_sets, params_to_row, _stream_ptr
from .flatten import FlatTree

EVENT_TYPES = ['gain', 'loss', 'duplication', 'base_number_gain', 'base_number_loss',
               'demi_gain_like', 'demi_loss_like', 'other']


def _percentile_from_hist(h, n_total, q):
    """np.percentile(counts, q) (default 'linear' method) for integer counts given their
    histogram h[v] = #replicates with v events."""
    cum = np.cumsum(h)
    pos = (q / 100.0) * (n_total - 1)
    lo = int(np.floor(pos))
    hi = min(lo + 1, n_total - 1)
    frac = pos - lo
    a = float(np.searchsorted(cum, lo, side='right'))
    b = float(np.searchsorted(cum, hi, side='right'))
    # numpy's _lerp
    diff = b - a
    out = a + diff * frac
    if frac >= 0.5:
        out = b - diff * (1 - frac)
    return out


class MappingResult:
    """Raw device accumulators of one stochastic-mapping run."""

    def __init__(self, flat, hist, sums, n_reps, n_events, n_bins):
        self.flat = flat
        self.hist = hist          # (n_nodes, 8, n_bins) uint32 ndarray; bin 0 filled in
        self.sums = sums          # (n_nodes, 8, 2) uint64 ndarray
        self.n_reps = n_reps
        self.n_events = n_events
        self.n_bins = n_bins

    def overflowed(self):
        return bool(self.hist[:, :, -1].any())

    def summary_for(self, k):
        out = {}
        R = self.n_reps
        for e, et in enumerate(EVENT_TYPES):
            tot = int(self.sums[k, e, 0])
            sq = int(self.sums[k, e, 1])
            if R == 0:
                out[et] = {'mean': 0, 'std': 0, 'ci_lower': 0, 'ci_upper': 0, 'total_occurrences': 0}
                continue
            mean = tot / R
            var = max(sq / R - mean * mean, 0.0)
            h = self.hist[k, e]
            out[et] = {'mean': np.float64(mean), 'std': np.float64(np.sqrt(var)),
                       'ci_lower': np.float64(_percentile_from_hist(h, R, 2.5)),
                       'ci_upper': np.float64(_percentile_from_hist(h, R, 97.5)),
                       'total_occurrences': np.int64(tot)}
        return out


def simulate(flat: FlatTree, q_row_params, n, root_probs, n_reps, seed, base_number=0,
             rep_begin=0, rep_end=None, n_bins=256, rep_chunk=1 << 17, device=None, q_matrix=None,
             return_device=False, node_set=None):
    """Run replicates [rep_begin, rep_end) of an `n_reps`-replicate mapping on the current GPU.

    Returns a MappingResult (host arrays), or the raw device tensors when ``return_device``.
    Branch-specific models: ``q_row_params`` is (S, 9) and ``node_set`` (n_nodes,) names the set
    of the branch above each node (AR:569); base numbers are taken from the rows.
    """
    if node_set is not None:
        return _simulate_sets(flat, q_row_params, node_set, n, root_probs, n_reps, seed, rep_begin,
                              rep_end, n_bins, rep_chunk, device, return_device)
    _lib.require_device()
    lib = _lib.load()
    device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    rep_end = n_reps if rep_end is None else rep_end
    ldq = lib.cevb_ldq(n)
    st = _stream_ptr()
    pw = torch.zeros((1, 7, n, ldq), dtype=torch.float64, device=device)
    if q_matrix is not None:
        pw[0, 0, :, :n].copy_(torch.as_tensor(np.asarray(q_matrix, dtype=np.float64), device=device))
    else:
        d_params = torch.as_tensor(np.asarray(q_row_params, dtype=np.float64)[None, :], device=device)
        _lib.check(lib.cevb_build_q(_lib.ptr(d_params), 1, n, _lib.ptr(pw), st), "cevb_build_q")
    d_root = torch.as_tensor(np.ascontiguousarray(root_probs, dtype=np.float64), device=device)
    i32 = dict(dtype=torch.int32, device=device)
    d_parent = torch.as_tensor(flat.parent, **i32)
    d_pre = torch.as_tensor(flat.preorder, **i32)
    d_len = torch.as_tensor(flat.dist_raw, dtype=torch.float64, device=device)
    hist = torch.zeros((flat.n_nodes, 8, n_bins), dtype=torch.int32, device=device)
    sums = torch.zeros((flat.n_nodes, 8, 2), dtype=torch.int64, device=device)
    n_events = torch.zeros(1, dtype=torch.int64, device=device)
    count = max(0, rep_end - rep_begin)
    chunk = int(min(rep_chunk, max(count, 1)))
    chunk = (chunk + 255) // 256 * 256
    state = torch.empty((flat.n_nodes, chunk), dtype=torch.int16, device=device)
    _lib.check(lib.cevb_stoch_map(_lib.ptr(pw), n, _lib.ptr(d_root), flat.n_nodes, _lib.ptr(d_parent),
                                  _lib.ptr(d_pre), _lib.ptr(d_len), int(base_number),
                                  ctypes.c_ulonglong(seed & (2 ** 64 - 1)), rep_begin, rep_end, n_bins,
                                  _lib.ptr(hist), _lib.ptr(sums), _lib.ptr(state), chunk,
                                  _lib.ptr(n_events), st), "cevb_stoch_map")
    if return_device:
        return hist, sums, n_events
    return finalize(flat, hist, sums, n_events, n_reps, n_bins)


def _simulate_sets(flat, rows, node_set, n, root_probs, n_reps, seed, rep_begin, rep_end, n_bins,
                   rep_chunk, device, return_device):
    _lib.require_device()
    lib = _lib.load()
    device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    rep_end = n_reps if rep_end is None else rep_end
    rows = np.ascontiguousarray(np.atleast_2d(rows), dtype=np.float64)
    S = rows.shape[0]
    ldq = lib.cevb_ldq(n)
    st = _stream_ptr()
    pw = torch.zeros((S, 7, n, ldq), dtype=torch.float64, device=device)
    d_params = torch.as_tensor(rows, device=device)
    _lib.check(lib.cevb_build_q(_lib.ptr(d_params), S, n, _lib.ptr(pw), st), "cevb_build_q")
    i32 = dict(dtype=torch.int32, device=device)
    d_set = torch.as_tensor(np.asarray(node_set, dtype=np.int32), **i32)
    d_bn = torch.as_tensor(rows[:, 8].astype(np.int32), **i32)
    d_root = torch.as_tensor(np.ascontiguousarray(root_probs, dtype=np.float64), device=device)
    d_parent = torch.as_tensor(flat.parent, **i32)
    d_pre = torch.as_tensor(flat.preorder, **i32)
    d_len = torch.as_tensor(flat.dist_raw, dtype=torch.float64, device=device)
    hist = torch.zeros((flat.n_nodes, 8, n_bins), dtype=torch.int32, device=device)
    sums = torch.zeros((flat.n_nodes, 8, 2), dtype=torch.int64, device=device)
    n_events = torch.zeros(1, dtype=torch.int64, device=device)
    count = max(0, rep_end - rep_begin)
    chunk = (int(min(rep_chunk, max(count, 1))) + 255) // 256 * 256
    state = torch.empty((flat.n_nodes, chunk), dtype=torch.int16, device=device)
    _lib.check(lib.cevb_stoch_map_sets(
        _lib.ptr(pw), S, 7 * n * ldq, _lib.ptr(d_set), _lib.ptr(d_bn), n, _lib.ptr(d_root),
        flat.n_nodes, _lib.ptr(d_parent), _lib.ptr(d_pre), _lib.ptr(d_len),
        ctypes.c_ulonglong(seed & (2 ** 64 - 1)), rep_begin, rep_end, n_bins, _lib.ptr(hist),
        _lib.ptr(sums), _lib.ptr(state), chunk, _lib.ptr(n_events), st), "cevb_stoch_map_sets")
    if return_device:
        return hist, sums, n_events
    return finalize(flat, hist, sums, n_events, n_reps, n_bins)


def finalize(flat, hist, sums, n_events, n_reps, n_bins):
    h = hist.cpu().numpy().astype(np.int64)
    h[:, :, 0] = n_reps - h[:, :, 1:].sum(axis=2)
    s = sums.cpu().numpy().astype(np.int64)
    return MappingResult(flat, h, s, n_reps, int(n_events.item()), n_bins)


def _root_distribution(mapper, n):
    root = mapper.tree.get_tree_root()
    if hasattr(root, 'state_probabilities'):                              # AR:544-547
        return np.asarray(root.state_probabilities, dtype=np.float64)
    p = np.zeros(n)
    p[int(root.ml_count)] = 1.0
    return p


def run_stochastic_mapping(mapper):
    """Backs StochasticMapper.perform_stochastic_mapping (AR:523-590)."""
    model = mapper.model
    n = model.max_chr + 1
    flat = FlatTree(mapper.tree, {}, model.max_chr)
    bn = model.params.get('base_number')
    bn = int(bn) if (bn and bn > 0) else 0
    rows, node_set = params_to_row(model.params), None
    if model.branch_params:      # AR:569: the branch above a node may have its own parameters
        rows, node_set = branch_parameter_sets(model, flat)
    res = simulate(flat, rows, n, _root_distribution(mapper, n), mapper.num_mappings,
                   mapper._draw_seed(), base_number=bn, n_bins=mapper.n_bins, node_set=node_set)
    if res.overflowed():   # a branch with >= n_bins-1 events of one type in some replicate
        res = simulate(flat, rows, n, _root_distribution(mapper, n), mapper.num_mappings,
                       mapper._draw_seed(), base_number=bn, n_bins=mapper.n_bins * 16,
                       node_set=node_set)
    mapper.last_result = res
    summary = {}
    for k, node in enumerate(flat.nodes):
        if k == flat.root:
            continue
        summary[id(node)] = res.summary_for(k)
    return summary


def simulate_single_branch(mapper, start_state, branch_length, params, max_events=4096):
    """Backs StochasticMapper.simulate_events_on_branch (AR:467-521): (events, end_state)."""
    _lib.require_device()
    lib = _lib.load()
    model = mapper.model
    n = model.max_chr + 1
    device = torch.device(f"cuda:{torch.cuda.current_device()}")
    st = _stream_ptr()
    pw = torch.zeros((1, 7, n, lib.cevb_ldq(n)), dtype=torch.float64, device=device)
    d_params = torch.as_tensor(params_to_row(params)[None, :], device=device)
    _lib.check(lib.cevb_build_q(_lib.ptr(d_params), 1, n, _lib.ptr(pw), st), "cevb_build_q")
    bn = params.get('base_number')
    bn = int(bn) if (bn and bn > 0) else 0
    ev_time = torch.zeros(max_events, dtype=torch.float64, device=device)
    ev_from = torch.zeros(max_events, dtype=torch.int32, device=device)
    ev_to = torch.zeros(max_events, dtype=torch.int32, device=device)
    ev_type = torch.zeros(max_events, dtype=torch.int32, device=device)
    out = torch.zeros(2, dtype=torch.int32, device=device)
    mapper._trace_calls = getattr(mapper, '_trace_calls', 0) + 1
    if not hasattr(mapper, '_trace_seed'):
        mapper._trace_seed = mapper._draw_seed()
    t = float(branch_length) if branch_length is not None else 0.0
    _lib.check(lib.cevb_stoch_branch_trace(_lib.ptr(pw), n, int(start_state), t, bn,
                                           ctypes.c_ulonglong(mapper._trace_seed & (2 ** 64 - 1)),
                                           ctypes.c_ulonglong(mapper._trace_calls), max_events,
                                           _lib.ptr(ev_time), _lib.ptr(ev_from), _lib.ptr(ev_to),
                                           _lib.ptr(ev_type), _lib.ptr(out), st),
               "cevb_stoch_branch_trace")
    ne, end = (int(v) for v in out.cpu().numpy())
    k = min(ne, max_events)
    tm, fr, to, ty = (x[:k].cpu().numpy() for x in (ev_time, ev_from, ev_to, ev_type))
    events = [{'time_on_branch': float(tm[i]), 'from_state': int(fr[i]), 'to_state': int(to[i]),
               'type': EVENT_TYPES[int(ty[i])]} for i in range(k)]
    return events, end
