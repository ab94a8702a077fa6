"""Batched device engine: parameter vectors -> tree log-likelihoods on one B200.

This is the additive, batched entry point that sits beside the reference-shaped classes in
``ancestral_reconstruction.py``.  Buffers are torch CUDA tensors (device
memory and streams only); all arithmetic happens in the hand-written kernels behind the C ABI.

Pipeline per chunk of parameter vectors (one stream, no host synchronisation inside):
  path="fused" (default): P(t) is never written to HBM for pairs whose series converges within
  64 terms (nu t below about 15.6; the rest is scaled and squared into a few P slots)
    cevb_build_q -> cevb_taylor_prepare -> cevb_expm_plan ->
    cevb_taylor_matrices (only the pairs that need squarings) -> cevb_prune_fused
  path="materialised": every P(t) by Pade + scaling-and-squaring, then streamed by the pruning
    cevb_build_q -> cevb_expm_prepare -> cevb_expm_branches -> cevb_prune
[-> cevb_argmax_states]
"""
from __future__ import annotations

import ctypes
import gc
import os

import numpy as np
import torch

from . import _lib
from .flatten import FlatTree

PARAM_ORDER = ['gain', 'loss', 'dupl', 'demidupl_gain', 'demidupl_loss',
               'base_number_gain', 'base_number_loss', 'linear_rate', 'base_number']

FUSED_INIT, FUSED_APPLY, FUSED_COMBINE, FUSED_LEAVES, FUSED_LEAFVEC, FUSED_ALL = 1, 2, 4, 8, 16, 31
INFO_IDENTITY_T0 = 1
INFO_NAN_FALLBACK = 2
INFO_GROWTH = 4
INFO_REPIVOTED = 8


def params_to_row(params):
    """dict (reference layout, AR:38-48) -> 9 doubles in the device order."""
    row = np.zeros(9, dtype=np.float64)
    for k, name in enumerate(PARAM_ORDER[:8]):
        v = params.get(name, 0)
        row[k] = float(v) if v is not None else 0.0
    bn = params.get('base_number')
    row[8] = float(bn) if (bn is not None and bn > 0) else 0.0
    return row


def branch_parameter_sets(model, flat):
    """Distinct parameter sets of a branch-heterogeneous model as (S, 9) rows and, per flattened
    node, the set used on the branch above it: ``branch_params.get(id(node), params)``
    (reference AR:230-233 and AR:569).  Set 0 is always ``model.params``."""
    rows = [params_to_row(model.params)]
    index = {rows[0].tobytes(): 0}
    node_set = np.zeros(flat.n_nodes, dtype=np.int64)
    for k, nd in enumerate(flat.nodes):
        bp = model.branch_params.get(id(nd))
        if bp is None:
            continue
        row = params_to_row(bp)
        key = row.tobytes()
        if key not in index:
            index[key] = len(rows)
            rows.append(row)
        node_set[k] = index[key]
    return np.stack(rows), node_set


def _stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class ExpmWorkspace:
    """Device buffers for `cap` parameter vectors x `n_t` branch lengths of order n."""

    def __init__(self, n, n_t, cap, device, p_slots=None, n_nodes=0):
        """p_slots: number of materialised matrices P can hold (default cap * n_t, i.e. every
        pair); n_nodes > 0 also allocates the fused path's tables."""
        lib = _lib.load()
        self.n, self.n_t, self.cap, self.device = n, n_t, cap, device
        self.p_slots = int(cap * n_t if p_slots is None else p_slots)
        self.ldq = lib.cevb_ldq(n)
        self.ldp = lib.cevb_ldp(n)
        f64 = dict(dtype=torch.float64, device=device)
        self.params = torch.zeros((cap, 9), **f64)
        self.pw = torch.zeros((cap, 7, n, self.ldq), **f64)
        self.norms = torch.zeros((cap, 16), **f64)
        self.csc = torch.zeros((cap, 2, 12, n), dtype=torch.int16, device=device)
        self.csc_cnt = torch.zeros((cap, 2, n), dtype=torch.uint8, device=device)
        if p_slots is None:
            self.P = torch.empty((cap, n_t, n, self.ldp), **f64)
        else:
            self.P = torch.empty((max(self.p_slots, 1), n, self.ldp), **f64)
        self.info = torch.zeros((cap, n_t), dtype=torch.int32, device=device)
        self.counter = torch.zeros(16, dtype=torch.int32, device=device)
        if n_nodes:
            self.tc = torch.empty((cap, lib.cevb_taylor_doubles(n)), **f64)
            self.slot = torch.empty((cap, n_t), dtype=torch.int32, device=device)
            self.plan_nc = torch.empty((cap, n_t), dtype=torch.int32, device=device)
            self.plan_tau = torch.empty((cap, n_t), **f64)
            self.plan_a0 = torch.empty((cap, n_t), **f64)
            self.worklist = torch.empty(max(self.p_slots, 1), dtype=torch.int32, device=device)
            self.work_s = torch.zeros(max(self.p_slots, 1), dtype=torch.int32, device=device)
            self.nan_count = torch.zeros(16, dtype=torch.int32, device=device)
            self.work_count = torch.zeros(16, dtype=torch.int32, device=device)   # one per stream
            # squarings of orders that do not fit in shared memory go through global memory:
            # one scratch matrix per persistent CTA, per stream slice
            sq = lib.cevb_taylor_scratch_doubles(n)
            self.sq_scratch = torch.empty((16, sq), **f64) if sq else None
            self.w = torch.zeros((cap, n_nodes, self.ldp), **f64)
            self.wr = torch.ones((cap, n_nodes, self.ldp), **f64)
            self.mode = torch.zeros((cap, n_nodes), dtype=torch.uint8, device=device)
        scratch_bytes = lib.cevb_expm_scratch_bytes(n)
        grid = lib.cevb_expm_grid(n, cap * n_t)
        self.scratch = (torch.empty(scratch_bytes * grid // 8 + 8, **f64)
                        if scratch_bytes else None)

    @staticmethod
    def bytes_per_vector(n, n_t):
        ld = n + (n & 1)
        return 8 * (7 * n * ld + n_t * n * ld) + 4 * n_t + 64 * n


class LikelihoodEngine:
    """Tree + data resident on the device; evaluates batches of parameter vectors."""

    def __init__(self, flat: FlatTree, device=None, max_workspace_bytes=32 << 30, chunk=None,
                 path="fused", squaring="series", n_streams=4):
        """path: "fused" (series on the tensor cores, P never written for nu t <= 15.6) or
        "materialised" (every P by the Pade kernel).  squaring: how the fused path forms the
        matrices that need scaling and squaring -- "series" (series at t 2^-s + squarings in
        shared memory) or "pade" (the scipy-algorithm kernel on a work list)."""
        _lib.require_device()
        if path not in ("fused", "materialised"):
            raise ValueError("path must be 'fused' or 'materialised'")
        if squaring not in ("series", "pade"):
            raise ValueError("squaring must be 'series' or 'pade'")
        self.path = path
        self.n_streams = max(1, min(16, int(n_streams)))
        self._streams = None
        if squaring == "series" and not _lib.load().cevb_taylor_matrices_supported(flat.n_states):
            squaring = "pade"     # two n x n matrices do not fit in shared memory
        self.squaring = squaring
        self.lib = _lib.load()
        self.flat = flat
        self.device = torch.device(device if device is not None else
                                   f"cuda:{torch.cuda.current_device()}")
        self.n = flat.n_states
        self.ldp = self.lib.cevb_ldp(self.n)
        i32 = dict(dtype=torch.int32, device=self.device)
        self.d_level_nodes = torch.as_tensor(flat.level_nodes, **i32)
        self.d_child_off = torch.as_tensor(flat.child_off, **i32)
        self.d_child_idx = torch.as_tensor(flat.child_idx, **i32)
        self.d_branch_of = torch.as_tensor(flat.branch_of, **i32)
        self.d_leaf_code = torch.as_tensor(flat.leaf_code, **i32)
        self.d_t = torch.as_tensor(flat.t_unique, dtype=torch.float64, device=self.device)
        self.h_level_off = np.ascontiguousarray(flat.level_off, dtype=np.int32)
        self.d_level_child = torch.as_tensor(flat.level_child, **i32)
        self.h_level_child_off = np.ascontiguousarray(flat.level_child_off, dtype=np.int32)
        self.d_leaf_child = torch.as_tensor(flat.leaf_child, **i32)
        self.d_t_node = torch.as_tensor(flat.dist_eff, dtype=torch.float64, device=self.device)
        tm = float(np.nanmax(flat.t_unique)) if flat.t_unique.size else 0.0
        self.t_max = tm if np.isfinite(tm) and not np.isnan(flat.t_unique).any() else 0.0
        self.d_t_order = torch.as_tensor(np.argsort(-np.nan_to_num(flat.t_unique), kind="stable")
                                         .astype(np.int32), device=self.device)
        mat_bytes = 8 * self.n * self.ldp
        if path == "fused":
            # per vector: powers + series coefficients + node vectors + w; the rest of the budget
            # holds materialised matrices (only pairs with nu t > 15.6 need one)
            per_vec = (8 * 7 * self.n * self.lib.cevb_ldq(self.n)
                       + 8 * self.lib.cevb_taylor_doubles(self.n)
                       + 24 * flat.n_nodes * self.ldp + 16 * flat.n_t + 64 * self.n)
            if chunk is None:
                chunk = max(1, min(4096, int((max_workspace_bytes // 2) // max(per_vec, 1))))
            self.p_budget = max(int(max_workspace_bytes) - int(chunk) * per_vec,
                                flat.n_t * mat_bytes)
        else:
            per_vec = (ExpmWorkspace.bytes_per_vector(self.n, flat.n_t)
                       + 8 * flat.n_nodes * self.ldp)
            if chunk is None:
                chunk = max(1, min(4096, int(max_workspace_bytes // max(per_vec, 1))))
            self.p_budget = None
        self.mat_bytes = mat_bytes
        # leaf rows of node_vec are outputs only; batched logL callers (bench, optimiser) switch
        # them off, the drop-in classes read node vectors and keep them on
        self.want_node_vectors = True
        # small batches (one evaluation, the 9 points of a finite-difference gradient) are launch-
        # bound: about 50 kernels of a few microseconds each.  Their launch sequence is captured
        # once per batch size into a CUDA graph and replayed.
        self.graph_max_batch = int(os.environ.get("CEVB_GRAPH_MAX", "32"))
        self._graphs = {}
        self.chunk = int(chunk)
        self.overflow_events = 0
        self.eval_token = 0          # bumped by every evaluation that rewrites the resident vectors
        self._ws = None
        self._node_vec = None
        self._cap = 0
        self.last_info = None
        self.repivot_events = 0
        self.stage_events = None     # set to a list to record (stage, start, end) CUDA events
        self.keep_info = False       # accumulate (m, s, flags) of every chunk on the host
        self.info_log = []

    # ------------------------------------------------------------------ buffers
    def _ensure(self, cap):
        if self._ws is not None and self._cap >= cap:
            return
        self._ws = None
        self._node_vec = None
        self._graphs = {}            # captured graphs hold pointers into the old workspace
        if self.path == "fused":
            slots = int(min(max(cap, self.n_streams) * self.flat.n_t,
                            max(self.p_budget // self.mat_bytes, self.n_streams * self.flat.n_t)))
            self._ws = ExpmWorkspace(self.n, self.flat.n_t, cap, self.device, p_slots=slots,
                                     n_nodes=self.flat.n_nodes)
        else:
            self._ws = ExpmWorkspace(self.n, self.flat.n_t, cap, self.device)
        f64 = dict(dtype=torch.float64, device=self.device)
        self._node_vec = torch.empty((cap, self.flat.n_nodes, self.ldp), **f64)
        self._logl = torch.empty(cap, **f64)
        self._rescales = torch.zeros(cap, dtype=torch.int32, device=self.device)
        self._cap = cap

    def root_freq_tensor(self, root_freq):
        if root_freq is None:
            root_freq = np.ones(self.n) / self.n
        return torch.as_tensor(np.ascontiguousarray(root_freq, dtype=np.float64),
                               device=self.device)

    # ------------------------------------------------------------------ stages (async)
    def _mark(self, stage, start=None):
        """Record a CUDA event on the launching stream when stage timing is on."""
        if self.stage_events is None:
            return None
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(torch.cuda.current_stream())
        if start is not None:
            self.stage_events.append((stage, start, ev))
        return ev

    def _transition_matrices(self, ws, nb, full_pivot=0):
        lib, st = self.lib, _stream_ptr()
        e0 = self._mark("prepare")
        _lib.check(lib.cevb_build_q(_lib.ptr(ws.params), nb, self.n, _lib.ptr(ws.pw), st),
                   "cevb_build_q")
        _lib.check(lib.cevb_expm_prepare(_lib.ptr(ws.pw), nb, self.n, _lib.ptr(ws.norms),
                                         _lib.ptr(ws.csc), _lib.ptr(ws.csc_cnt), st),
                   "cevb_expm_prepare")
        e1 = self._mark("prepare", e0)
        _lib.check(lib.cevb_expm_branches(_lib.ptr(ws.pw), _lib.ptr(ws.norms), _lib.ptr(ws.csc),
                                          _lib.ptr(ws.csc_cnt), _lib.ptr(self.d_t), nb,
                                          self.flat.n_t, self.n, _lib.ptr(ws.P), _lib.ptr(ws.info),
                                          _lib.ptr(ws.scratch), _lib.ptr(ws.counter), full_pivot,
                                          st), "cevb_expm_branches")
        self._mark("expm", e1)

    def _prune(self, ws, nb, d_root_freq):
        f = self.flat
        e0 = self._mark("prune")
        _lib.check(self.lib.cevb_prune(
            _lib.ptr(ws.P), nb, f.n_t, self.n, f.n_nodes, f.root, _lib.ptr(self.d_level_nodes),
            self.h_level_off.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), f.n_levels,
            _lib.ptr(self.d_child_off), _lib.ptr(self.d_child_idx), _lib.ptr(self.d_branch_of),
            _lib.ptr(self.d_leaf_code), _lib.ptr(d_root_freq), _lib.ptr(self._node_vec),
            _lib.ptr(self._logl), _lib.ptr(self._rescales), _stream_ptr()), "cevb_prune")
        self._mark("prune", e0)

    def _fused_slice(self, ws, lo, hi, k, d_root_freq, full_pivot, st):
        """Vectors [lo, hi) of the chunk, work-list partition k, all launches on stream `st`."""
        lib, f, nb = self.lib, self.flat, hi - lo
        P = lambda t: _lib.ptr(t[lo:])                       # per-vector buffers
        pps = ws.p_slots // self.n_streams                   # P slots of one partition
        Q = lambda t: _lib.ptr(t[k * pps:])                  # per-partition work-list buffers
        one = lambda t: _lib.ptr(t[k:])
        e0 = self._mark("prepare")
        _lib.check(lib.cevb_build_q(P(ws.params), nb, self.n, P(ws.pw), st), "cevb_build_q")
        _lib.check(lib.cevb_taylor_prepare(P(ws.pw), nb, self.n, self.t_max, P(ws.norms), P(ws.csc),
                                           P(ws.csc_cnt), P(ws.tc), st), "cevb_taylor_prepare")
        _lib.check(lib.cevb_expm_plan(P(ws.norms), _lib.ptr(self.d_t), nb, f.n_t, pps, P(ws.slot),
                                      Q(ws.worklist), Q(ws.work_s), one(ws.work_count), P(ws.plan_nc),
                                      P(ws.plan_tau), P(ws.plan_a0), st), "cevb_expm_plan")
        e1 = self._mark("prepare", e0)
        if self.squaring == "series":
            _lib.check(lib.cevb_taylor_matrices(
                P(ws.tc), P(ws.norms), _lib.ptr(self.d_t_order), nb, f.n_t, self.n, P(ws.slot),
                Q(ws.work_s), one(ws.work_count), pps, P(ws.plan_nc), P(ws.plan_tau),
                P(ws.plan_a0), Q(ws.P), one(ws.counter), one(ws.nan_count),
                _lib.ptr(ws.sq_scratch[k]) if ws.sq_scratch is not None else _lib.ptr(None), st),
                "cevb_taylor_matrices")
        else:
            # Pade-13 + scaling and squaring (the scipy algorithm) for the listed pairs
            _lib.check(lib.cevb_expm_prepare(P(ws.pw), nb, self.n, P(ws.norms), P(ws.csc),
                                             P(ws.csc_cnt), st), "cevb_expm_prepare")
            _lib.check(lib.cevb_expm_branches_list(
                P(ws.pw), P(ws.norms), P(ws.csc), P(ws.csc_cnt), _lib.ptr(self.d_t), nb, f.n_t,
                self.n, Q(ws.P), P(ws.info), _lib.ptr(ws.scratch), one(ws.counter), full_pivot,
                Q(ws.worklist), one(ws.work_count), pps, st), "cevb_expm_branches_list")
        e2 = self._mark("expm", e1)

        def prune(l0, l1, stages):
            _lib.check(lib.cevb_prune_fused(
                P(ws.tc), P(ws.norms), Q(ws.P), P(ws.slot), nb, f.n_t, self.n, f.n_nodes, f.root,
                _lib.ptr(self.d_level_nodes),
                self.h_level_off.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), f.n_levels,
                _lib.ptr(self.d_child_off), _lib.ptr(self.d_child_idx), _lib.ptr(self.d_branch_of),
                _lib.ptr(self.d_leaf_code), _lib.ptr(self.d_level_child),
                self.h_level_child_off.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                _lib.ptr(self.d_leaf_child), int(f.leaf_child.size), P(ws.plan_nc),
                P(ws.plan_tau), P(ws.plan_a0), _lib.ptr(d_root_freq), P(self._node_vec), P(ws.w), P(ws.wr), P(ws.mode),
                P(self._logl), P(self._rescales), l0, l1, stages, st), "cevb_prune_fused")

        leafvec = FUSED_LEAFVEC if self.want_node_vectors else 0
        if self.stage_events is None:
            prune(0, f.n_levels, FUSED_ALL & ~FUSED_LEAFVEC | leafvec)
        else:
            # same launches, issued level by level so the kernels can be timed separately
            prune(0, 0, FUSED_INIT | leafvec)
            ev = self._mark("init", e2)
            prune(0, 0, FUSED_LEAVES)
            ev = self._mark("leaf", ev)
            for lv in range(f.n_levels):
                prune(lv, lv + 1, FUSED_APPLY)
                ev = self._mark("apply", ev)
                prune(lv, lv + 1, FUSED_COMBINE)
                ev = self._mark("combine", ev)

    def _fused_chunk(self, ws, nb, d_root_freq, full_pivot=0):
        """One chunk: its vectors are split over `n_streams` CUDA streams so that the narrow
        upper tree levels of one slice (latency-bound: few branches per coefficient tile)
        overlap the wide lower levels of another.  Stage timing runs on the caller's stream."""
        S = self.n_streams if (self.stage_events is None and nb >= 2 * self.n_streams
                               and self.squaring == "series") else 1   # the Pade scratch is shared
        cur = torch.cuda.current_stream()
        if S == 1:
            self._fused_slice(ws, 0, nb, 0, d_root_freq, full_pivot,
                              ctypes.c_void_p(cur.cuda_stream))
            return
        if self._streams is None:
            self._streams = [torch.cuda.Stream(device=self.device) for _ in range(self.n_streams)]
        start = torch.cuda.Event()
        start.record(cur)
        per = (nb + S - 1) // S
        for k in range(S):
            lo, hi = k * per, min(nb, (k + 1) * per)
            if hi <= lo:
                continue
            stream = self._streams[k]
            stream.wait_event(start)
            self._fused_slice(ws, lo, hi, k, d_root_freq, full_pivot,
                              ctypes.c_void_p(stream.cuda_stream))
            done = torch.cuda.Event()
            done.record(stream)
            cur.wait_event(done)

    def _run_chunk(self, d_params_chunk, d_root_freq, full_pivot=0):
        nb = d_params_chunk.shape[0]
        ws = self._ws
        ws.params[:nb].copy_(d_params_chunk)
        if self.path == "fused":
            ws.work_count.zero_()
            self._fused_chunk(ws, nb, d_root_freq, full_pivot)
            pps = ws.p_slots // self.n_streams
            if self.squaring == "pade":
                self._growth_flag |= ((ws.info[:nb] >> 16) & INFO_GROWTH).any()
            self._overflow_flag |= (ws.work_count > pps).any()
            self._materialised += ws.work_count.sum().to(torch.int64)
            if self.keep_info:
                cnts = torch.clamp(ws.work_count, max=pps).cpu().numpy()
                s_list = np.concatenate([ws.work_s[k * pps:k * pps + int(c)].cpu().numpy()
                                         for k, c in enumerate(cnts)])
                self.info_log.append((s_list, ws.slot[:nb].cpu().numpy().copy(),
                                      ws.norms[:nb].cpu().numpy().copy()))
            return
        self._transition_matrices(ws, nb, full_pivot)
        # element-growth monitor of the tile-pivoted solve: OR into a device flag, no host sync
        self._growth_flag |= ((ws.info[:nb] >> 16) & INFO_GROWTH).any()
        if self.keep_info:
            self.info_log.append(ws.info[:nb].cpu().numpy().copy())
        self._prune(ws, nb, d_root_freq)

    def launches_per_chunk(self):
        """Kernel launches of ours per chunk.  materialised: build_q, 4 power GEMMs, norms, order,
        expm, leaf init and one pruning kernel per non-empty level.  fused: build_q, Q structure,
        series coefficients, plan, series matrices + squaring (or, squaring="pade": 4 power
        GEMMs, norms, order, expm on the work list), leaf init, one combine kernel per non-empty
        level and one apply kernel per non-empty (level, leaf / non-leaf) part."""
        levels = int(np.count_nonzero(np.diff(self.flat.level_off)))
        if self.path != "fused":
            return 9 + levels
        applies = int(self.flat.leaf_child.size > 0) + \
            int(np.count_nonzero(np.diff(self.flat.level_child_off)))
        base = (6 if self.squaring == "series" else 11) + (1 if self.want_node_vectors else 0)
        per_slice = base + levels + applies
        # a chunk of at least 2 * n_streams vectors is issued as n_streams slices, each with its
        # own launch sequence
        nb = min(getattr(self, "last_batch", self.chunk), self.chunk)
        slices = self.n_streams if (self.squaring == "series" and nb >= 2 * self.n_streams) else 1
        return per_slice * slices

    # ------------------------------------------------------------------ public
    def log_likelihood(self, params, root_freq=None, check_growth=True):
        """params: (B, 9) array-like or CUDA tensor in PARAM_ORDER -> CUDA tensor (B,) of logL."""
        if torch.is_tensor(params):
            d_params = params.to(device=self.device, dtype=torch.float64)
        else:
            d_params = torch.as_tensor(np.ascontiguousarray(params, dtype=np.float64),
                                       device=self.device)
        if d_params.dim() == 1:
            d_params = d_params[None, :]
        B = d_params.shape[0]
        self.eval_token += 1
        d_root = root_freq if torch.is_tensor(root_freq) else self.root_freq_tensor(root_freq)
        if (check_growth and 0 < B <= min(self.graph_max_batch, self.chunk) and self.path == "fused"
                and self.squaring == "series" and self.stage_events is None and not self.keep_info):
            out = self._graph_log_likelihood(d_params, d_root)
            if out is not None:
                return out
        out = torch.empty(B, dtype=torch.float64, device=self.device)
        self.rescale_counts = torch.zeros(B, dtype=torch.int32, device=self.device)
        self._ensure(min(self.chunk, B))
        full_pivot = 0
        step = self._cap
        while True:
            self._growth_flag = torch.zeros((), dtype=torch.bool, device=self.device)
            self._overflow_flag = torch.zeros((), dtype=torch.bool, device=self.device)
            self._materialised = torch.zeros((), dtype=torch.int64, device=self.device)
            for lo in range(0, B, step):
                hi = min(B, lo + step)
                self._run_chunk(d_params[lo:hi], d_root, full_pivot)
                out[lo:hi].copy_(self._logl[:hi - lo])
                self.rescale_counts[lo:hi].copy_(self._rescales[:hi - lo])
            if not check_growth:
                break
            # one host read at the end of the batch
            flags = torch.stack([self._growth_flag, self._overflow_flag]).cpu().numpy()
            if flags[1]:
                # more pairs needed a materialised matrix than P has slots: redo the batch in
                # chunks small enough that every pair of a chunk fits
                self.overflow_events += 1
                step = max(1, (self._ws.p_slots // self.n_streams) // max(self.flat.n_t, 1))
                continue
            if flags[0] and not full_pivot:
                # a tripped element-growth monitor (not seen in practice) redoes the batch with
                # full partial pivoting in every panel
                self.repivot_events += 1
                full_pivot = 1
                continue
            break
        self.n_chunks = (B + step - 1) // step
        self.last_batch = B
        return out

    # ------------------------------------------------------------------ pipelined callers
    def log_likelihood_deferred(self, params, root_freq=None):
        """Enqueue one batch WITHOUT reading the overflow / growth flags back: returns
        (logL tensor, ticket).  The flags travel to pinned host memory behind the kernels;
        ``resolve(ticket)`` -- typically called after the NEXT batch has been enqueued -- waits
        for that copy only, and redoes the batch synchronously in the (rare) flagged case.  A
        caller that pipelines like this never drains the GPU between optimiser steps; the
        multi-GPU bench gathers the log-likelihoods of step k while step k+1 computes."""
        d_params = params if torch.is_tensor(params) else torch.as_tensor(
            np.ascontiguousarray(params, dtype=np.float64), device=self.device)
        out = self.log_likelihood(d_params, root_freq=root_freq, check_growth=False)
        flags = torch.stack([self._growth_flag, self._overflow_flag])
        host = torch.empty(2, dtype=torch.bool).pin_memory()
        host.copy_(flags, non_blocking=True)
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream())
        return out, (host, done, d_params, root_freq, out)

    def resolve(self, ticket):
        """Finish a deferred batch: returns True when the enqueued result stands."""
        host, done, d_params, root_freq, out = ticket
        done.synchronize()
        if not bool(host[0]) and not bool(host[1]):
            return True
        out.copy_(self.log_likelihood(d_params, root_freq=root_freq))     # checked, synchronous
        return False

    def _graph_log_likelihood(self, d_params, d_root):
        """Small batch through a captured CUDA graph; None when the captured run has to be redone
        eagerly (work-list overflow)."""
        B = d_params.shape[0]
        self._ensure(min(self.chunk, max(B, 1)))
        key = (B, bool(self.want_node_vectors))
        entry = self._graphs.get(key)
        cur = torch.cuda.current_stream()
        if entry is None:
            f64 = dict(dtype=torch.float64, device=self.device)
            sp = torch.zeros((B, 9), **f64)
            sr = torch.zeros(self.n, **f64)
            flags = torch.zeros(2, dtype=torch.bool, device=self.device)
            sp.copy_(d_params)
            sr.copy_(d_root)

            def body():
                self._growth_flag = torch.zeros((), dtype=torch.bool, device=self.device)
                self._overflow_flag = torch.zeros((), dtype=torch.bool, device=self.device)
                self._materialised = torch.zeros((), dtype=torch.int64, device=self.device)
                self._run_chunk(sp, sr, 0)
                flags.copy_(torch.stack([self._growth_flag, self._overflow_flag]))

            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):          # warm-up outside capture (function attributes,
                body()                             # allocator) as CUDA graph capture requires
            cur.wait_stream(side)
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            # no cyclic-GC pass while the stream is capturing: collecting an old CUDAGraph (or
            # any object whose finaliser touches the device) inside a capture invalidates it
            gc_was_on = gc.isenabled()
            gc.disable()
            try:
                with torch.cuda.graph(graph):
                    body()
            finally:
                if gc_was_on:
                    gc.enable()
            entry = (graph, sp, sr, flags)
            self._graphs[key] = entry
        graph, sp, sr, flags = entry
        sp.copy_(d_params)
        sr.copy_(d_root)
        graph.replay()
        if bool(flags[1].item()):
            self.overflow_events += 1
            return None
        self.rescale_counts = self._rescales[:B].clone()
        self.n_chunks = 1
        self.last_batch = B
        return self._logl[:B].clone()

    def log_likelihood_branch_sets(self, param_rows, node_set, root_freq=None):
        """Branch-heterogeneous model (reference AR:230-233, where the branch
        above a child uses ``model.branch_params.get(id(child), model.params)``).

        param_rows: (S, 9) distinct parameter sets; node_set[k]: the set used on the branch above
        node k (the root's entry is ignored).  The transition matrix of every (set, length) pair
        in use is formed by the scipy-algorithm kernels (work-list form: at most one matrix per
        branch, as in the reference) and ONE pruning pass runs over them.  Materialised path only.  Returns the log-likelihood as a 0-d CUDA tensor;
        node vectors / ancestral states of the evaluation are read with which=0."""
        if self.path != "materialised":
            raise _lib.CevbError("branch-specific parameter sets need path='materialised'")
        f = self.flat
        rows = np.ascontiguousarray(np.atleast_2d(param_rows), dtype=np.float64)
        S = rows.shape[0]
        node_set = np.asarray(node_set, dtype=np.int64)
        if node_set.shape != (f.n_nodes,) or node_set.min() < 0 or node_set.max() >= S:
            raise ValueError("node_set must hold one parameter-set index per node")
        d_root = root_freq if torch.is_tensor(root_freq) else self.root_freq_tensor(root_freq)
        self.eval_token += 1
        non_root = np.arange(f.n_nodes) != f.root
        pair = node_set * f.n_t + f.branch_of.astype(np.int64)       # (set, branch-length) index
        uniq, inverse = np.unique(pair[non_root], return_inverse=True)
        branch_of = np.zeros(f.n_nodes, dtype=np.int32)
        branch_of[non_root] = inverse.astype(np.int32)
        U = int(uniq.size)
        # Only the (set, length) pairs in use are computed -- at most one matrix per branch, like
        # the reference (AR:230-237) -- through the work-list form of the expm kernels: position q
        # of the list computes pair uniq[q] into P slot q, so P is already in pruning order.
        ws = self._het_workspace(S, U)
        ws.params[:S].copy_(torch.as_tensor(rows, device=self.device))
        d_branch_of = torch.as_tensor(branch_of, device=self.device)
        ws.worklist[:U].copy_(torch.as_tensor(uniq.astype(np.int32), device=self.device))
        ws.work_count.zero_()
        ws.work_count[0] = U
        lib, st = self.lib, _stream_ptr()
        _lib.check(lib.cevb_build_q(_lib.ptr(ws.params), S, self.n, _lib.ptr(ws.pw), st), "cevb_build_q")
        _lib.check(lib.cevb_expm_prepare(_lib.ptr(ws.pw), S, self.n, _lib.ptr(ws.norms),
                                         _lib.ptr(ws.csc), _lib.ptr(ws.csc_cnt), st), "cevb_expm_prepare")
        full_pivot = 0
        while True:
            _lib.check(lib.cevb_expm_branches_list(
                _lib.ptr(ws.pw), _lib.ptr(ws.norms), _lib.ptr(ws.csc), _lib.ptr(ws.csc_cnt),
                _lib.ptr(self.d_t), S, f.n_t, self.n, _lib.ptr(ws.P), _lib.ptr(ws.info),
                _lib.ptr(ws.scratch), _lib.ptr(ws.counter), full_pivot, _lib.ptr(ws.worklist),
                _lib.ptr(ws.work_count), U, st), "cevb_expm_branches_list")
            flagged = ws.info.reshape(-1).index_select(0, ws.worklist[:U].to(torch.int64))
            if full_pivot or not bool((((flagged >> 16) & INFO_GROWTH) != 0).any().item()):
                break
            self.repivot_events += 1     # element-growth monitor tripped: full pivoting
            full_pivot = 1
        if self._node_vec is None:
            # only the vectors of one evaluation: not the S x n_t matrices _ensure() would size
            f64 = dict(dtype=torch.float64, device=self.device)
            self._node_vec = torch.empty((1, f.n_nodes, self.ldp), **f64)
            self._logl = torch.empty(1, **f64)
            self._rescales = torch.zeros(1, dtype=torch.int32, device=self.device)
        _lib.check(self.lib.cevb_prune(
            _lib.ptr(ws.P), 1, U, self.n, f.n_nodes, f.root,
            _lib.ptr(self.d_level_nodes),
            self.h_level_off.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), f.n_levels,
            _lib.ptr(self.d_child_off), _lib.ptr(self.d_child_idx), _lib.ptr(d_branch_of),
            _lib.ptr(self.d_leaf_code), _lib.ptr(d_root), _lib.ptr(self._node_vec),
            _lib.ptr(self._logl), _lib.ptr(self._rescales), _stream_ptr()), "cevb_prune")
        self.rescale_counts = self._rescales[:1].clone()
        self.last_batch = 1
        return self._logl[0].clone()

    def _het_workspace(self, S, U):
        """Buffers for S parameter sets and U (set, length) pairs in use: powers and norms per
        set, one P slot per pair (the materialised engine's own workspace is sized for every
        set x length combination and is not touched)."""
        ws = getattr(self, "_het_ws", None)
        if ws is None or ws.cap < S or ws.p_slots < U:
            ws = ExpmWorkspace(self.n, self.flat.n_t, max(S, 1), self.device, p_slots=max(U, 1))
            ws.worklist = torch.empty(max(U, 1), dtype=torch.int32, device=self.device)
            ws.work_count = torch.zeros(16, dtype=torch.int32, device=self.device)
            self._het_ws = ws
        return ws

    # ------------------------------------------------------------------ marginal / joint
    # ADDITIVE calls (the reference has no down-pass and no joint backtrack, AR:283-295; its
    # ml_count stays the arg max of the up-pass vector).  Both need every P(t), so they run on a
    # materialised companion engine; one parameter vector per call.
    def _materialised_companion(self):
        if self.path == "materialised":
            return self
        eng = getattr(self, "_mat_engine", None)
        if eng is None:
            eng = LikelihoodEngine(self.flat, device=self.device, path="materialised", chunk=1)
            self._mat_engine = eng
        return eng

    def _recon_tree_tensors(self):
        if getattr(self, "d_parent", None) is None:
            i32 = dict(dtype=torch.int32, device=self.device)
            self.d_parent = torch.as_tensor(self.flat.parent, **i32)
            self.d_depth_off = torch.as_tensor(self.flat.depth_off, **i32)
            self.h_depth_off = np.ascontiguousarray(self.flat.depth_off, dtype=np.int32)

    def marginal_reconstruction(self, params_row, root_freq=None):
        """Posterior state probabilities of every node given ALL the data (up-pass + down-pass)
        for one parameter vector -> (post (n_nodes, n) CUDA tensor, logL float)."""
        eng = self._materialised_companion()
        eng._recon_tree_tensors()
        f = eng.flat
        row = np.ascontiguousarray(np.asarray(params_row, dtype=np.float64).reshape(1, 9))
        d_root = root_freq if torch.is_tensor(root_freq) else eng.root_freq_tensor(root_freq)
        logl = eng.log_likelihood(row, root_freq=d_root)
        f64 = dict(dtype=torch.float64, device=eng.device)
        down = torch.empty((1, f.n_nodes, eng.ldp), **f64)
        post = torch.empty((1, f.n_nodes, eng.ldp), **f64)
        _lib.check(eng.lib.cevb_marginal_down(
            _lib.ptr(eng._ws.P), 1, f.n_t, eng.n, f.n_nodes, _lib.ptr(eng.d_parent),
            _lib.ptr(eng.d_child_off), _lib.ptr(eng.d_child_idx), _lib.ptr(eng.d_branch_of),
            _lib.ptr(eng.d_leaf_code),
            eng.h_depth_off.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), f.n_depths,
            _lib.ptr(d_root), _lib.ptr(eng._node_vec), _lib.ptr(down), _lib.ptr(post),
            _stream_ptr()), "cevb_marginal_down")
        return post[0, :, :eng.n], float(logl[0])

    def joint_reconstruction(self, params_row, root_freq=None):
        """The single most probable joint assignment of states to all nodes (Pupko et al. 2000)
        for one parameter vector -> (states (n_nodes,) int32 CUDA tensor, log P(states, data))."""
        eng = self._materialised_companion()
        eng._recon_tree_tensors()
        f = eng.flat
        row = np.ascontiguousarray(np.asarray(params_row, dtype=np.float64).reshape(1, 9))
        d_root = root_freq if torch.is_tensor(root_freq) else eng.root_freq_tensor(root_freq)
        eng.log_likelihood(row, root_freq=d_root)                   # fills P (and the up vectors)
        dev = eng.device
        lj = torch.zeros((1, f.n_nodes, eng.ldp), dtype=torch.float64, device=dev)
        cj = torch.zeros((1, f.n_nodes, eng.n), dtype=torch.int32, device=dev)
        expo = torch.zeros((1, f.n_nodes), dtype=torch.int64, device=dev)
        states = torch.zeros((1, f.n_nodes), dtype=torch.int32, device=dev)
        logp = torch.zeros(1, dtype=torch.float64, device=dev)
        _lib.check(eng.lib.cevb_joint_reconstruct(
            _lib.ptr(eng._ws.P), 1, f.n_t, eng.n, f.n_nodes, f.root, _lib.ptr(eng.d_parent),
            _lib.ptr(eng.d_level_nodes),
            eng.h_level_off.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), f.n_levels,
            _lib.ptr(eng.d_child_off), _lib.ptr(eng.d_child_idx), _lib.ptr(eng.d_branch_of),
            _lib.ptr(eng.d_leaf_code), _lib.ptr(eng.d_depth_off), f.n_depths, _lib.ptr(d_root),
            _lib.ptr(lj), _lib.ptr(cj), _lib.ptr(expo), _lib.ptr(states), _lib.ptr(logp),
            _stream_ptr()), "cevb_joint_reconstruct")
        return states[0], float(logp[0])

    def node_vectors(self, which=0):
        """Conditional-likelihood vectors (n_nodes, n) of vector `which` of the last chunk."""
        return self._node_vec[which, :, :self.n]

    def transition_matrices(self, which=0):
        """(n_t, n, n) view of the P matrices of vector `which` of the last chunk
        (materialised path only: the fused path does not form them)."""
        if self.path == "fused":
            raise _lib.CevbError("the fused path does not materialise P; use path='materialised'")
        return self._ws.P[which, :, :, :self.n]

    def order_info(self, which=0):
        info = self._ws.info[which].cpu().numpy()
        return info & 0xff, (info >> 8) & 0xff, info >> 16

    def ancestral_states(self, which=0, want_probs=True):
        """argmax states for vector `which` of the last chunk -> (ml_count, ml_prob, probs)."""
        return self.ancestral_states_of(self._node_vec[which], want_probs)

    def ancestral_states_of(self, vec, want_probs=True):
        """argmax states (AR:298-320) of conditional-likelihood vectors (n_nodes, ldp) or
        (n_nodes, n) on the device."""
        f = self.flat
        if vec.shape[-1] != self.ldp or not vec.is_contiguous():
            full = torch.zeros((f.n_nodes, self.ldp), dtype=torch.float64, device=self.device)
            full[:, :self.n] = vec[:, :self.n]
            vec = full
        ml_count = torch.empty(f.n_nodes, dtype=torch.int32, device=self.device)
        ml_prob = torch.empty(f.n_nodes, dtype=torch.float64, device=self.device)
        probs = (torch.zeros((f.n_nodes, self.ldp), dtype=torch.float64, device=self.device)
                 if want_probs else None)
        _lib.check(self.lib.cevb_argmax_states(_lib.ptr(vec), 1, f.n_nodes, self.n,
                                               _lib.ptr(ml_count), _lib.ptr(ml_prob),
                                               _lib.ptr(probs), _stream_ptr()),
                   "cevb_argmax_states")
        return ml_count, ml_prob, (probs[:, :self.n] if want_probs else None)


# ---------------------------------------------------------------------- single-matrix helpers
def rate_matrix(params_row, n, device=None):
    """One Q (n x n, host ndarray) from the device kernel -- AR:75-134."""
    _lib.require_device()
    lib = _lib.load()
    device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    rows = np.atleast_2d(np.asarray(params_row, dtype=np.float64))
    B = rows.shape[0]
    ldq = lib.cevb_ldq(n)
    d_params = torch.as_tensor(rows, device=device)
    pw = torch.zeros((B, 7, n, ldq), dtype=torch.float64, device=device)
    _lib.check(lib.cevb_build_q(_lib.ptr(d_params), B, n, _lib.ptr(pw), _stream_ptr()),
               "cevb_build_q")
    return pw[:, 0, :, :n].cpu().numpy()


def transition_matrices(params_rows, times, n, device=None, full_pivot=0, return_info=False,
                        q_matrices=None):
    """P[b, e] = post(expm(Q_b t_e)) as a host ndarray (B, E, n, n) -- AR:136-166.

    ``q_matrices`` (B, n, n), when given, replaces the rate matrices built from ``params_rows``
    (any real square matrices: the expm kernels do not assume ChromEvol structure)."""
    _lib.require_device()
    device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    times = np.atleast_1d(np.asarray(times, dtype=np.float64))
    if q_matrices is not None:
        q_matrices = np.asarray(q_matrices, dtype=np.float64).reshape(-1, n, n)
        rows = np.zeros((q_matrices.shape[0], 9))
    else:
        rows = np.atleast_2d(np.asarray(params_rows, dtype=np.float64))
    B, E = rows.shape[0], times.shape[0]
    ws = ExpmWorkspace(n, E, B, device)
    ws.params.copy_(torch.as_tensor(rows, device=device))
    d_t = torch.as_tensor(times, device=device)
    lib, st = _lib.load(), _stream_ptr()
    if q_matrices is not None:
        ws.pw[:, 0, :, :n].copy_(torch.as_tensor(q_matrices, device=device))
    else:
        _lib.check(lib.cevb_build_q(_lib.ptr(ws.params), B, n, _lib.ptr(ws.pw), st),
                   "cevb_build_q")
    _lib.check(lib.cevb_expm_prepare(_lib.ptr(ws.pw), B, n, _lib.ptr(ws.norms), _lib.ptr(ws.csc),
                                     _lib.ptr(ws.csc_cnt), st), "cevb_expm_prepare")
    _lib.check(lib.cevb_expm_branches(_lib.ptr(ws.pw), _lib.ptr(ws.norms), _lib.ptr(ws.csc),
                                      _lib.ptr(ws.csc_cnt), _lib.ptr(d_t), B, E, n, _lib.ptr(ws.P),
                                      _lib.ptr(ws.info), _lib.ptr(ws.scratch),
                                      _lib.ptr(ws.counter), full_pivot, st), "cevb_expm_branches")
    P = ws.P[:, :, :, :n].cpu().numpy()
    if return_info:
        info = ws.info.cpu().numpy()
        return P, info & 0xff, (info >> 8) & 0xff, info >> 16
    return P
