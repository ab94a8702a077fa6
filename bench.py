#!/usr/bin/env python
"""Benchmark of the chromosome-number likelihood hot path (BASELINE.json metric:
tree logL evals/sec, 1k taxa, N=100; expm FP64 TFLOP/s).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm

A "step" is one pass of the hot path (rate matrices -> P(t) for every branch -> pruning ->
logL) over one batch of synthetic parameter vectors on the 1,000-taxon / N=100 tree.  The
workload at N=1 is the per-GPU share of BASELINE configs[2] (4,096 vectors over 8 GPUs = 512
per GPU per step, weak scaling); the single-vector case (configs[1]) is reported beside it.
One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_TAXA = 1000
MAX_CHR = 100
TREE_SEED, COUNT_SEED, PARAM_SEED = 1, 2, 4
METRIC = "tree_logL_evals_per_sec_1k_taxa_N100"
UNIT = "evals/s"


def build_workload(per_gpu_batch, rank):
    from chromevol_enhanced_b200 import synthetic
    tree = synthetic.random_tree(N_TAXA, seed=TREE_SEED)
    counts = synthetic.simulate_tip_counts(tree, MAX_CHR, seed=COUNT_SEED)
    rows = synthetic.random_param_rows(per_gpu_batch, seed=PARAM_SEED + 1000 * rank, linear='positive')
    return tree, counts, rows


# ------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def begin(self):
        """Host time at which the timed region starts (the sampler itself is started before the
        warm-up steps: nvidia-smi needs ~0.1 s to deliver its first line and the timed region of
        five 16 ms steps is shorter than that)."""
        self.t_begin = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t_end = time.perf_counter()
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t_begin = getattr(self, "t_begin", 0.0)
        inside = [ln for (ts, ln) in self.lines if t_begin <= ts <= t_end + 0.03]
        # a line reports the state a few ms before it arrives; if the region was too short for any
        # line to land inside it, the lines of the warm-up steps right before it (same load) stand in
        chosen = inside if inside else [ln for (ts, ln) in self.lines if ts >= t_begin - 0.25]
        for line in chosen:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "power_w_max": float(max(power)) if power else None,
                "samples": len(sm), "samples_inside_timed_region": len(inside) if sm else 0,
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """One host process: faithful oracle evaluations (Q rebuilt per branch like AR:141)."""
    newick, counts, rows = args
    from oracle import chromevol_oracle as O
    from chromevol_enhanced_b200 import synthetic
    from chromevol_enhanced_b200.newick import Tree
    tree = Tree(newick, format=1)
    t0 = time.perf_counter()
    out = []
    for row in rows:
        ll, _, _ = O.log_likelihood(tree, counts, MAX_CHR, synthetic.row_to_dict(row), faithful_q=True)
        out.append(ll)
    return time.perf_counter() - t0, out


def cpu_reference_throughput(tree, counts, rows, evals_per_core=1, cores=None):
    """The reference's numpy/scipy algorithm (oracle port) on the host cores: one process per
    core, each evaluating a disjoint slice of the batch.  Returns (evals/s, cores, n_evals, logLs)."""
    import multiprocessing as mp
    cores = cores or max(1, min(os.cpu_count() or 1, 32))
    n_eval = cores * evals_per_core
    take = np.resize(np.arange(rows.shape[0]), n_eval)
    chunks = [rows[take[c * evals_per_core:(c + 1) * evals_per_core]] for c in range(cores)]
    newick = tree.write(dist_formatter="%.17g")
    ctx = mp.get_context("spawn")
    # one BLAS thread per worker process.  The variables are set HERE, before the workers are
    # spawned: a spawned interpreter imports numpy (and starts its BLAS pool) while it imports
    # this module, i.e. before any code of the worker function runs.
    blas_env = ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "MKL_NUM_THREADS")
    saved = {k: os.environ.get(k) for k in blas_env}
    os.environ.update({k: "1" for k in blas_env})
    t0 = time.perf_counter()
    try:
        with ctx.Pool(cores) as pool:
            res = pool.map(_cpu_worker, [(newick, counts, ch) for ch in chunks])
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    wall = time.perf_counter() - t0
    busy = max(r[0] for r in res)
    lls = [x for r in res for x in r[1]]
    return n_eval / busy, cores, n_eval, lls, wall


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    tree, counts, rows = build_workload(max(64, args.batch), 0)
    per_step = []
    cores = None
    for step in range(args.warmup + args.steps):
        eps, cores, n_eval, _, wall = cpu_reference_throughput(
            tree, counts, rows[step % 8::8] if rows.shape[0] >= 64 else rows, evals_per_core=1)
        if step >= args.warmup:
            per_step.append((eps, n_eval, wall))
    value = float(np.mean([p[0] for p in per_step]))
    ms = float(np.mean([p[1] / p[0] for p in per_step]) * 1e3)
    sample = (f"{per_step[0][1]} evaluations per step (1 per host process, {cores} processes), "
              f"faithful per-branch Q rebuild + scipy.linalg.expm + numpy pruning")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def ncu_traffic(path):
    """Mean dram__bytes_read.sum + dram__bytes_write.sum per launch, per kernel family, from the
    committed ncu CSV of one 512-vector step of the CURRENT kernels (profiles/README.md has the
    command).  The squaring path's series launches (template MODE 2) are kept apart from the
    fused apply launches the roofline block describes."""
    import csv
    import gzip
    import re
    if not os.path.exists(path):
        return {}
    with gzip.open(path, "rt") as fh:
        rows = list(csv.reader(fh))
    hdr = next((i for i, r in enumerate(rows) if "Kernel Name" in r and "Metric Name" in r), None)
    if hdr is None:
        return {}
    h = rows[hdr]
    ki, mi, vi, ui, idi = (h.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per = {}
    for r in rows[hdr + 1:]:
        if len(r) <= vi or not r[mi].startswith("dram__bytes_"):
            continue
        name = r[ki]
        fam = name.split("<")[0].split("(")[0].replace("void ", "").replace("cevb::", "")
        if fam == "taylor_apply_kernel":
            m = re.search(r"taylor_apply_kernel<[^,]*,[^,]*,\s*\(?(?:int\))?\s*(\d+)", name)
            if m and m.group(1) == "2":
                fam = "taylor_apply_kernel<EVAL>"
        try:
            val = float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
        except ValueError:
            continue
        d = per.setdefault(fam, {})
        d[r[idi]] = d.get(r[idi], 0.0) + val
    return {fam: int(sum(d.values()) / len(d)) for fam, d in per.items() if d}


def workload_config(args, world):
    return {"workload": f"BASELINE configs[2] per-GPU share: synthetic {N_TAXA}-taxon tree, max count "
                        f"N={MAX_CHR} (n={MAX_CHR + 1} states, 1998 branches), {args.batch} parameter "
                        f"vectors per GPU per step (rates LogUniform(1e-3,10), linear_rate U(0,0.1))",
            "per_gpu_batch": args.batch, "global_batch": args.batch * world,
            "tree_seed": TREE_SEED, "param_seed": PARAM_SEED, "parallelism": f"dp{world} (vectors sharded)",
            "l2": "per-step working set (series coefficients 4.0 MB + node vectors 4.8 MB per vector, "
                  "x batch) far exceeds the 126 MB L2; no flush needed"}


# ------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    from chromevol_enhanced_b200 import distributed as D
    from chromevol_enhanced_b200.engine import LikelihoodEngine
    from chromevol_enhanced_b200.flatten import FlatTree

    rank, world, local = D.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    tree, counts, rows = build_workload(args.batch, rank)
    flat = FlatTree(tree, counts, MAX_CHR)
    eng = LikelihoodEngine(flat, device=dev, max_workspace_bytes=args.workspace_gb << 30)
    eng.want_node_vectors = False      # logL only; switched on for the single-vector + states case
    B = rows.shape[0]
    d_rows = torch.as_tensor(rows, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # One step = enqueue this rank's batch, start the all-gather of its log-likelihoods, then
    # settle the PREVIOUS step (its flag copy and its gather).  Nothing drains the GPU between
    # steps; the last step is settled inside the timed region by `drain`.
    pipe = D.GatherPipeline(B * world, world)
    state = {"ticket": None, "ll": None}

    def settle():
        if state["ticket"] is not None:
            eng.resolve(state["ticket"])
            state["ticket"] = None
        got = pipe.collect()
        if got is not None:
            state["ll"] = got

    def step_resident():
        out, ticket = eng.log_likelihood_deferred(d_rows)
        prev = state["ticket"]
        if prev is not None:
            eng.resolve(prev)
        got = pipe.collect()
        if got is not None:
            state["ll"] = got
        pipe.submit(out)
        state["ticket"] = ticket
        return state["ll"]

    def drain():
        settle()
        return state["ll"]

    # the call a user of the drop-in makes for a batch: host array in, host array out
    # (ChromEvolLikelihoodCalculator.log_likelihood_batch: H2D of the parameter rows, the device
    # evaluation, D2H of the log-likelihoods, all inside the call)
    import chromevol_enhanced_b200 as cev
    model = cev.ChromEvolutionModel(MAX_CHR)
    calc = cev.ChromEvolLikelihoodCalculator(tree, counts, model)
    calc.set_root_frequencies(np.ones(MAX_CHR + 1) / (MAX_CHR + 1))
    calc._device_engine()
    calc._engine = eng                       # same resident tree and workspace as the timed engine
    names8 = cev.MODEL_PARAMETER_NAMES
    rows8 = np.ascontiguousarray(rows[:, :8])

    def step_e2e():
        out = calc.log_likelihood_batch(rows8, names8)
        if world > 1:
            full = D.all_gather_ragged(torch.as_tensor(out, device=dev), B * world, world)
            return full.cpu().numpy()
        return out

    # ---- device-resident timing ("value")
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # before the warm-up steps, see ClockSampler.begin
    for _ in range(args.warmup):
        step_resident()
    drain()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if rank == 0:
        sampler.begin()
    ev0.record()
    for _ in range(args.steps):
        step_resident()
    ll = drain()
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t_own = ev0.elapsed_time(ev1) * 1e-3
    t_res = D.max_across_ranks(t_own, dev)
    # every rank's own device time: the ranks evaluate different random batches (different series
    # lengths), so part of the gap to perfect weak scaling is the slowest batch, not communication
    per_rank_ms = [t_own / args.steps * 1e3]
    if world > 1:
        own = torch.tensor([t_own / args.steps * 1e3], dtype=torch.float64, device=dev)
        per_rank_ms = D.all_gather_ragged(own, world, world).cpu().tolist()
    n_chunks = eng.n_chunks
    launches_per_step = n_chunks * eng.launches_per_chunk()      # of our kernels, all stream slices
    # Pade orders / squaring counts the device chose, from one extra untimed step (deterministic)
    eng.keep_info, eng.info_log = True, []
    step_resident()
    drain()
    barrier()
    eng.keep_info = False
    info_log = eng.info_log

    # ---- per-stage timing with CUDA events on the launching stream (roofline numerators)
    eng.stage_events = []
    barrier()
    for _ in range(args.steps):
        step_resident()
    drain()
    barrier()
    stage_ms = {}
    for name, a, b in eng.stage_events:
        stage_ms[name] = stage_ms.get(name, 0.0) + a.elapsed_time(b)
    eng.stage_events = None

    # ---- end-to-end through the public call with host buffers ("e2e")
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        step_e2e()
    ev1.record()
    barrier()
    t_e2e = D.max_across_ranks(max(ev0.elapsed_time(ev1) * 1e-3, 0.0), dev)
    del t0

    # ---- single-vector latency, BASELINE configs[1] (logL + ancestral states)
    one = d_rows[:1].clone()
    eng.want_node_vectors = True
    from chromevol_enhanced_b200 import synthetic
    one.copy_(torch.as_tensor(synthetic.default_param_row()[None], device=dev))
    for _ in range(3):
        eng.log_likelihood(one)
        eng.ancestral_states(0, want_probs=False)
    barrier()
    ev0.record()
    reps1 = 10
    for _ in range(reps1):
        eng.log_likelihood(one)
        eng.ancestral_states(0, want_probs=False)
    ev1.record()
    barrier()
    t_single = ev0.elapsed_time(ev1) * 1e-3 / reps1

    eng.want_node_vectors = False
    extras = None if args.no_configs else run_other_configs(dev, rank, world, eng, tree, counts)
    if rank != 0:
        D.shutdown()
        return 0

    # ---- algorithmic work of the timed region
    from chromevol_enhanced_b200 import accounting
    n = MAX_CHR + 1
    steps = max(args.steps, 1)
    s_list = np.concatenate([x[0].reshape(-1) for x in info_log])    # squarings of the listed pairs
    norms = np.concatenate([x[2] for x in info_log])
    nu, mu = norms[:, 12], norms[:, 11]                   # CEVB_NORM_SHIFTED, CEVB_NORM_SHIFT
    # generator vectors (all of this workload): their observed-leaf branches run in
    # leaf_lowp_kernel (FP64 column + TF32 floor correction), everything else in the FP64 kernel
    lowp = accounting.is_generator_row(rows, n) & (os.environ.get("CEVB_LOWP", "1") != "0")
    work = accounting.fused_work(flat, nu, mu, lowp_leaf=lowp)
    t_stage = {k: v * 1e-3 / steps for k, v in stage_ms.items()}
    leaf_s = t_stage.get("leaf", 0.0)
    # the FP64 tensor kernel: every level launch, plus the leaf launch when no vector takes the
    # TF32 path
    apply_s = t_stage.get("apply", 0.0) + (0.0 if lowp.any() else leaf_s)
    combine_s = t_stage.get("combine", 0.0)
    expm_s, prep_s, init_s = t_stage.get("expm", 0.0), t_stage.get("prepare", 0.0), t_stage.get("init", 0.0)
    apply_launches = max(n_chunks * (int(flat.leaf_child.size > 0 and not lowp.all()) +
                                     int(np.count_nonzero(np.diff(flat.level_child_off)))), 1)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    # FP64 tensor peak: MEASURED_PEAKS.json has no FP64 figure; measure cuBLAS DGEMM here
    a = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
    torch.matmul(a, a)
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, a)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    fp64_peak = 2 * 8192 ** 3 / best / 1e12
    del a
    traffic = ncu_traffic(os.path.join(ROOT, "profiles", "r02_apply_traffic.csv.gz"))
    peak_src = "cuBLAS DGEMM 8192^3 measured in this run (MEASURED_PEAKS.json has no FP64 figure)"
    apply_tf = work["algorithmic_flops"] / apply_s / 1e12 if apply_s > 0 else None
    apply_exec_tf = work["dmma_flops"] / apply_s / 1e12 if apply_s > 0 else None
    # pairs with nu t > 15.6: series at t 2^-s (2 n^2 per term) + s squarings (2 n^3 each)
    list_pairs = int(s_list.size)
    list_flops = float(np.sum(2.0 * n ** 3 * s_list)) + 2.0 * n * n * 4 * accounting.KC * list_pairs
    list_tf = list_flops / expm_s / 1e12 if (expm_s > 0 and list_pairs) else None
    # what the reference's algorithm (one Pade-13 expm per branch) would have cost,
    # counted without any squarings (a lower bound)
    pade_equiv = float(accounting.pade_flops(n, 0)) * B * flat.n_branches
    pade_equiv_tf = pade_equiv / (apply_s + leaf_s + expm_s + prep_s) / 1e12 if apply_s > 0 else None
    # combine kernel: reads w (or streams a materialised P) and writes the node vectors
    combine_bytes = 8.0 * (work["fused_pairs"] * n + B * flat.n_nodes * n) + \
        8.0 * float(list_pairs) * (n * n + n)
    combine_gbs = combine_bytes / combine_s / 1e9 if combine_s > 0 else None
    s_hist = {int(k): int(v) for k, v in zip(*np.unique(s_list, return_counts=True))}

    # ---- CPU baseline (oracle port) on the host cores: bounded sample
    cpu = None
    if world == 1 and not args.skip_cpu:
        eps, cores, n_eval, cpu_ll, wall = cpu_reference_throughput(tree, counts, rows, evals_per_core=1)
        gpu_ll = ll.cpu().numpy()
        take = np.resize(np.arange(rows.shape[0]), n_eval)
        rel = float(np.max(np.abs(gpu_ll[take] - np.array(cpu_ll)) / np.abs(np.array(cpu_ll))))
        cpu = {"value": eps, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n_eval} of the step's {B} vectors, one per host process "
                         f"(faithful per-branch Q rebuild + scipy expm + numpy pruning), {wall:.1f} s wall",
               "max_rel_logL_diff_vs_gpu": rel}

    total = B * world * args.steps
    line = {
        "metric": METRIC, "value": total / t_res, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t_res / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, world),
        "e2e": {"value": total / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(B * 9 * 8),
                "d2h_bytes_per_step": int(B * 8), "ms_per_step": t_e2e / args.steps * 1e3},
        "gpu_launches": int(args.steps * launches_per_step),
        "roofline": {"kernel": "taylor_apply_kernel", "bound": "tensor", "achieved": apply_tf,
                     "peak": fp64_peak, "unit": "TFLOP/s",
                     "frac": (apply_tf / fp64_peak) if apply_tf else None,
                     "traffic": (traffic or {}).get("taylor_apply_kernel"),
                     "peak_source": peak_src,
                     "algorithmic_flops_per_launch": work["algorithmic_flops"] / apply_launches,
                     "launch_ms": apply_s / apply_launches * 1e3,
                     "executed_dmma_tflops": apply_exec_tf,
                     "executed_frac": (apply_exec_tf / fp64_peak) if apply_exec_tf else None,
                     "share_of_step": apply_s / max(sum(t_stage.values()), 1e-12),
                     "fused_pairs": work["fused_pairs"], "mean_series_terms": work["mean_ncoef"],
                     "note": "the FP64 tensor kernel of the level launches (branches above internal nodes); "
                             "algorithmic = 2 n^2 terms + 3 n^2 per (vector, branch) pair it processes; "
                             "executed_dmma = what a dense coefficient table would execute incl. 8-branch "
                             "group and 104-column padding (the slot-prefix table skips about 38 % of it); "
                             "observed-leaf branches of generator vectors are in roofline_leaf"},
        "roofline_leaf": {"kernel": "leaf_lowp_kernel", "bound": "tensor (tf32, mma.sync) / issue",
                          "pairs": work["lowp_pairs"],
                          "achieved": (work["lowp_tf32_flops"] / leaf_s / 1e12) if (leaf_s > 0 and work["lowp_pairs"]) else None,
                          "unit": "TFLOP/s (tf32)",
                          "peak": peaks.get("bf16_tflops", 1590.0) / 2.0,
                          "peak_source": "half of MEASURED_PEAKS.json bf16_tflops (dense TF32 runs at half the bf16 rate)",
                          "frac": (work["lowp_tf32_flops"] / leaf_s / 1e12 / (peaks.get("bf16_tflops", 1590.0) / 2.0))
                          if (leaf_s > 0 and work["lowp_pairs"]) else None,
                          "fp64_column_gflops": (work["lowp_fp64_flops"] / leaf_s / 1e9) if leaf_s > 0 else None,
                          "launch_ms": leaf_s / max(n_chunks, 1) * 1e3,
                          "note": "observed-leaf branches of generator vectors: one column of P in FP64 plus the "
                                  "floor correction of the row sum from the full matrix in TF32 (2 n * 104 * 8 ceil(terms/8) "
                                  "flops per pair); the kernel is issue / latency bound, the fraction of the TF32 "
                                  "tensor peak is reported for completeness"},
        "roofline_squaring": {"kernel": "taylor_apply_kernel<EVAL> + square_post_kernel", "bound": "tensor",
                              "achieved": list_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                              "frac": (list_tf / fp64_peak) if list_tf else None,
                              "pairs": list_pairs, "squarings_hist": s_hist,
                              "note": "pairs whose series needs more than 64 terms; algorithmic = 128 n^2 (series) + 2 s n^3 (squarings)"},
        "roofline_combine": {"kernel": "combine_level_kernel", "bound": "hbm", "achieved": combine_gbs,
                             "peak": hbm_peak, "unit": "GB/s",
                             "frac": (combine_gbs / hbm_peak) if combine_gbs else None,
                             "traffic": (traffic or {}).get("combine_level_kernel"), "peak_source": hbm_src,
                             "algorithmic_bytes_per_step": combine_bytes},
        "pade13_equivalent_tflops": pade_equiv_tf,
        "stage_ms_per_step": {k: v * 1e3 for k, v in t_stage.items()},
        "single_vector": {"workload": "BASELINE configs[1]: one vector logL + argmax states",
                          "ms": t_single * 1e3, "evals_per_s": 1.0 / t_single},
        "repivot_events": eng.repivot_events,
        "clocks": clocks,
        "per_rank_ms_per_step": per_rank_ms,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if extras is not None:
        mu = extras.get("model_selection_unit", {}).get("roofline")
        if mu and mu.get("achieved"):
            mu["peak"] = fp64_peak
            mu["frac"] = mu["achieved"] / fp64_peak
        if "dropin_objective" in extras:
            extras["dropin_objective"]["vs_single_vector_ms"] = extras["dropin_objective"]["ms"] / (t_single * 1e3)
        line["other_configs"] = extras
    print(json.dumps(line))
    D.shutdown()
    return 0


def _shipped_data():
    """The reference's shipped tree + counts (data/pruned_phylogeny.nwk, data/chromosome_counts.csv)
    as committed with the golden fixtures (tests/golden/golden.json)."""
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as fh:
        meta = json.load(fh)
    return meta["shipped_newick"], dict(meta["shipped_counts"])


def run_other_configs(dev, rank, world, eng2, tree2, counts2):
    """The other BASELINE configs, every run (they add a few seconds):
    configs[0] shipped data, --use_chromevol --optimize_params through the drop-in;
    configs[1/2] the reference-shaped objective call, the 9-point finite-difference step sharded
    over the ranks, and the full 4,096-vector batch on one GPU;
    configs[3] one evaluation on a 10,000-taxon tree with N=256, with its own roofline block;
    configs[4] 1e6 stochastic-mapping replicates sharded over the ranks."""
    import torch
    import chromevol_enhanced_b200 as cev
    from chromevol_enhanced_b200 import accounting, mapper, synthetic
    from chromevol_enhanced_b200 import distributed as D
    from chromevol_enhanced_b200.engine import LikelihoodEngine
    from chromevol_enhanced_b200.flatten import FlatTree
    from chromevol_enhanced_b200.newick import Tree
    out = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()

    # ---- configs[0]: shipped data through the drop-in classes (rank 0; 73 branches, N=35)
    if rank == 0:
        newick, counts = _shipped_data()
        res0 = {}
        for label, batched in (("sequential_gradient", False), ("batched_gradient", True)):
            tree = Tree(newick, format=1)
            model = cev.ChromEvolutionModel(35)
            calc = cev.ChromEvolLikelihoodCalculator(tree, counts, model)
            calc.set_root_frequencies(np.ones(36) / 36)
            opt = cev.ChromEvolOptimizer(tree, counts, model)
            opt.calculator = calc
            opt.batched_gradient = batched
            calc.calculate_total_log_likelihood()          # flatten + first graph capture: untimed
            calc.likelihoods = {}
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = opt.optimize_parameters()
            calc.likelihoods = {}
            ll = calc.calculate_total_log_likelihood()
            calc.ancestral_state_reconstruction()
            torch.cuda.synchronize()
            res0[label] = {"seconds": time.perf_counter() - t0, "nfev": int(r.nfev) if r else None,
                           "nit": int(r.nit) if r else None, "logL": float(ll)}
        res0["workload"] = ("BASELINE configs[0]: shipped pruned_phylogeny.nwk + chromosome_counts.csv, N=35, "
                            "optimize_parameters (L-BFGS-B from the defaults) + final logL + ancestral states "
                            "through ChromEvolOptimizer / ChromEvolLikelihoodCalculator")
        res0["reference_seconds"] = 14.4
        res0["reference_note"] = ("the unmodified reference on the survey container's CPU (SURVEY.md section "
                                  "8(d): 504 evaluations, logL -55.1204); not re-measured on this box")
        out["shipped_optimise"] = res0

    # ---- the reference-shaped objective call at configs[1] (1,000 taxa, N=100), rank 0
    if rank == 0:
        model = cev.ChromEvolutionModel(MAX_CHR)
        opt = cev.ChromEvolOptimizer(tree2, counts2, model)
        opt.calculator.set_root_frequencies(np.ones(MAX_CHR + 1) / (MAX_CHR + 1))
        names = cev.ChromEvolOptimizer.DEFAULT_PARAM_NAMES
        x = np.array([model.params[nm] for nm in names])
        for _ in range(3):
            opt.objective_function(x, names)
        torch.cuda.synchronize()
        reps = 20
        t0 = time.perf_counter()
        for k in range(reps):
            x[0] = 0.1 + 1e-4 * k
            f = opt.objective_function(x, names)
        wall = (time.perf_counter() - t0) / reps
        out["dropin_objective"] = {"workload": "ChromEvolOptimizer.objective_function (AR:341-371), 1,000 taxa, N=100: "
                                               "one scalar back per call, node vectors stay on the device",
                                   "ms": wall * 1e3, "neg_logL": float(f)}

    # ---- configs[1]: ancestral reconstruction of one parameter vector (1,000 taxa, N=100), rank 0:
    # the reference's arg max of the up-pass vectors, and the additive marginal (down-pass) and
    # joint (Pupko) reconstructions over materialised P
    if rank == 0:
        model = cev.ChromEvolutionModel(MAX_CHR)
        calc = cev.ChromEvolLikelihoodCalculator(tree2, counts2, model)
        calc.set_root_frequencies(np.ones(MAX_CHR + 1) / (MAX_CHR + 1))
        rec = {}
        for label, fn in (("argmax_up_pass", lambda: (calc.__setattr__("likelihoods", {}),
                                                       calc.ancestral_state_reconstruction())),
                          ("marginal_down_pass", calc.marginal_reconstruction),
                          ("joint_pupko", calc.joint_reconstruction)):
            fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            rec[label + "_ms"] = (time.perf_counter() - t0) / 3 * 1e3
        nodes = list(tree2.traverse())
        rec["nodes_where_marginal_differs_from_up_pass_argmax"] = int(sum(
            int(nd.marginal_count) != int(nd.ml_count) for nd in nodes if not nd.is_leaf()))
        rec["nodes_where_joint_differs_from_marginal"] = int(sum(
            int(nd.joint_count) != int(nd.marginal_count) for nd in nodes if not nd.is_leaf()))
        rec["workload"] = ("BASELINE configs[1]: ancestral reconstruction of one parameter vector, 1,000 taxa, "
                           "N=100, host wall clock incl. annotating the 1,999 nodes (marginal / joint: every "
                           "P(t) materialised by the Pade kernel, then one launch per depth)")
        out["reconstruction"] = rec

    # ---- the 9-point finite-difference step of one L-BFGS-B gradient, sharded over the ranks
    rows9 = np.tile(synthetic.default_param_row(), (9, 1))
    for k in range(8):
        rows9[k + 1, k] += 1e-8
    d9 = torch.as_tensor(rows9, device=dev)
    lo, hi = D.shard_bounds(9, rank, world)

    def fd_step():
        local = eng2.log_likelihood(d9[lo:hi]) if hi > lo else torch.zeros(0, dtype=torch.float64, device=dev)
        return D.all_gather_ragged(local, 9, world)

    for _ in range(3):
        fd_step()
    sync()
    e0.record()
    for _ in range(10):
        g = fd_step()
    e1.record()
    sync()
    t_fd = D.max_across_ranks(e0.elapsed_time(e1) * 1e-3 / 10, dev)
    out["fd_gradient_step"] = {"workload": "BASELINE configs[1/2]: the 9 evaluations of one forward-difference "
                                           "gradient (1,000 taxa, N=100) split over the ranks + all-gather",
                               "ms": t_fd * 1e3, "evals_per_s": 9 / t_fd, "logL0": float(g[0])}

    # ---- the full configs[2] batch (4,096 vectors) on ONE GPU (rank 0 only: a strong-scaling anchor)
    if rank == 0:
        rows4k = synthetic.random_param_rows(4096, seed=PARAM_SEED, linear='positive')
        d4k = torch.as_tensor(rows4k, device=dev)
        eng2.log_likelihood(d4k)
        torch.cuda.synchronize()
        e0.record()
        eng2.log_likelihood(d4k)
        e1.record()
        torch.cuda.synchronize()
        t4k = e0.elapsed_time(e1) * 1e-3
        out["full_batch_one_gpu"] = {"workload": "BASELINE configs[2]: all 4,096 parameter vectors of one optimiser "
                                                 "step on a single GPU", "ms": t4k * 1e3, "evals_per_s": 4096 / t4k,
                                     "chunks": eng2.n_chunks}
        del d4k
    sync()

    # ---- configs[4]: 1e6 stochastic-mapping replicates, sharded over the ranks
    flat = FlatTree(tree2, {}, MAX_CHR)
    root = np.zeros(MAX_CHR + 1)
    root[20] = 1.0
    reps = 1_000_000
    lo, hi = D.shard_bounds(reps, rank, world)
    row = synthetic.default_param_row()
    warm = mapper.simulate(flat, row, MAX_CHR + 1, root, reps, 7, rep_begin=lo, rep_end=min(hi, lo + 4096),
                           return_device=True)
    if world > 1:
        # the first all-reduce of this size sets up NCCL's channels (90 ms at N=8): untimed, like
        # the warm-up steps of the main loop
        D.sum_across_ranks(warm[0])
        D.sum_across_ranks(warm[2].to(torch.float64))
    del warm
    sync()
    e0.record()
    hist, sums, n_events = mapper.simulate(flat, row, MAX_CHR + 1, root, reps, 7, rep_begin=lo, rep_end=hi,
                                           return_device=True)
    if world > 1:
        D.sum_across_ranks(hist)
    e1.record()
    sync()
    t = D.max_across_ranks(e0.elapsed_time(e1) * 1e-3, dev)
    ev = D.sum_across_ranks(n_events.to(torch.float64)).item() if world > 1 else float(n_events.item())
    out["stochastic_mapping"] = {"workload": "BASELINE configs[4]: 1e6 replicates of the 1,000-taxon tree (N=100, default "
                                             "rates), replicates sharded over the ranks, histograms summed (NCCL)",
                                 "replicates_per_s": reps / t, "events_per_s": ev / t,
                                 "branch_visits_per_s": reps * flat.n_branches / t, "seconds": t}

    # ---- configs[3] building block: one evaluation (and the 9-point gradient batch) on a
    # 10,000-taxon tree with N=256, rank 0, with the roofline of its own apply launches
    if rank == 0:
        big = synthetic.random_tree(10_000, seed=11)
        bcounts = synthetic.simulate_tip_counts_fast(big, 256, seed=12, root_state=40)
        fl = FlatTree(big, bcounts, 256)
        eng = LikelihoodEngine(fl, device=dev, max_workspace_bytes=64 << 30)
        eng.want_node_vectors = False
        res = {}
        for nb in (1, 9):
            rows = np.tile(synthetic.default_param_row(), (nb, 1))
            rows[:, 0] *= 1.0 + 1e-3 * np.arange(nb)
            d = torch.as_tensor(rows, device=dev)
            eng.log_likelihood(d)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(3):
                ll = eng.log_likelihood(d)
            e1.record()
            torch.cuda.synchronize()
            res[f"batch_{nb}_ms"] = e0.elapsed_time(e1) / 3
        # stage times + series accounting of the 9-vector batch
        eng.keep_info, eng.info_log = True, []
        eng.graph_max_batch = 0                     # eager launches so that stage events can be placed
        eng.log_likelihood(d)
        torch.cuda.synchronize()
        eng.keep_info = False
        norms = np.concatenate([x[2] for x in eng.info_log])
        work = accounting.fused_work(fl, norms[:, 12], norms[:, 11])
        eng.stage_events = []
        eng.log_likelihood(d)
        torch.cuda.synchronize()
        st = {}
        for name, a, b in eng.stage_events:
            st[name] = st.get(name, 0.0) + a.elapsed_time(b)
        eng.stage_events = None
        apply_s = st.get("apply", 0.0) * 1e-3
        res["stage_ms"] = st
        res["logL"] = float(ll[0])
        res["rescales"] = int(eng.rescale_counts[0])
        res["squaring_path"] = eng.squaring
        res["workload"] = ("BASELINE configs[3] building block: 10,000-taxon tree, N=256 (n=257, 19,998 branches), "
                           "default rates; one evaluation and the 9-point gradient batch of one L-BFGS-B step "
                           "(model-selection candidates are independent optimiser runs, sharded by candidate)")
        res["roofline"] = {"kernel": "taylor_apply_kernel (n=257: three passes of 13 tiles per row)",
                           "bound": "tensor", "unit": "TFLOP/s",
                           "achieved": work["algorithmic_flops"] / apply_s / 1e12 if apply_s > 0 else None,
                           "executed_dmma_tflops": work["dmma_flops"] / apply_s / 1e12 if apply_s > 0 else None,
                           "fused_pairs": work["fused_pairs"], "mean_series_terms": work["mean_ncoef"],
                           "note": "9-vector batch, single stream, CUDA events per stage; peak = the DGEMM "
                                   "figure of the main roofline block"}
        out["model_selection_unit"] = res
    sync()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=512, help="parameter vectors per GPU per step")
    ap.add_argument("--workspace-gb", type=int, default=48)
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--extras", action="store_true", help="(default now; kept for old command lines)")
    ap.add_argument("--no-configs", action="store_true",
                    help="skip the other BASELINE configs (shipped-data optimise, N=256 / 10,000 taxa, "
                         "stochastic mapping, finite-difference step, 4,096-vector batch)")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: libraries that print to file descriptor 1 (NCCL's
    # version banner on rank 0) are sent to stderr, and the JSON line goes to the real stdout
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, "w", buffering=1)
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
