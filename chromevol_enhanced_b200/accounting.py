"""Host-side accounting of the work one evaluation does on the device (bench.py / DESIGN.md
roofline numerators).  Nothing here is on the measured path."""
from __future__ import annotations

import numpy as np

KC = 16              # CEVB_TAYLOR_KC
TRUNC = 1e-18        # CEVB_TAYLOR_TRUNC


def taylor_ncoef(x, mut):
    """Vectorised copy of the device rule (csrc/fused.cu: taylor_ncoef): series coefficients
    (degree + 1) for x = nu t, mut = mu t; 4*KC + 1 when the table is too short."""
    x = np.asarray(x, dtype=np.float64)
    mut = np.broadcast_to(np.asarray(mut, dtype=np.float64), x.shape)
    out = np.full(x.shape, 4 * KC + 1, dtype=np.int64)
    done = ~(x > 0.0)
    out[done] = 1
    term = x.copy()
    with np.errstate(all='ignore'):
        sc = np.exp(-mut)
        for k in range(1, 4 * KC):
            term = term * (x / (k + 1))
            ok = (~done) & ((k + 2) > x) & (term / (1.0 - x / (k + 2)) * sc <= TRUNC)
            out[ok] = k + 1
            done |= ok
    return out


def plan_pairs(t, nu, mu):
    """(kind, ncoef, squarings) per (vector, branch): kind 0 fused, 1 identity, 2 materialised."""
    t = np.asarray(t, dtype=np.float64)[None, :]
    x = np.asarray(nu, dtype=np.float64)[:, None] * t
    mut = np.asarray(mu, dtype=np.float64)[:, None] * t
    ident = np.broadcast_to(~(t > 0.0), x.shape)
    nc = taylor_ncoef(x, mut)
    s = np.zeros(x.shape, dtype=np.int64)
    need = (nc > 4 * KC) & ~ident
    while need.any():
        s[need] += 1
        nc[need] = taylor_ncoef(x[need] / 2.0 ** s[need], mut[need] / 2.0 ** s[need])
        need = (nc > 4 * KC) & ~ident & (s < 1000)
    kind = np.where(ident, 1, np.where(s > 0, 2, 0))
    return kind, nc, s


def pade_flops(n, s):
    """Canonical Pade-13 + scaling-and-squaring cost SURVEY.md section 8(d) defines per matrix."""
    return (2.0 * (6 + np.asarray(s, dtype=np.float64)) + 8.0 / 3.0) * float(n) ** 3


def is_generator_row(rows, n_states):
    """Per parameter row (device order): True when every off-diagonal rate of Q is >= 0, i.e.
    the vector takes the TF32 leaf path (n <= 104).  Mirrors CEVB_NORM_GENERATOR for ChromEvol
    rate matrices: rates are r * (1 + linear_rate * i), i = 1..N."""
    rows = np.atleast_2d(np.asarray(rows, dtype=np.float64))
    N = n_states - 1
    lin_ok = (1.0 + rows[:, 7] * 1.0 >= 0.0) & (1.0 + rows[:, 7] * N >= 0.0)
    rates_ok = np.all(rows[:, :7] >= 0.0, axis=1) & np.all(np.isfinite(rows[:, :8]), axis=1)
    some = np.any(rows[:, :7] > 0.0, axis=1)
    return lin_ok & rates_ok & some & (n_states <= 104)


def fused_work(flat, nu, mu, groups_per_warp=1, lowp_leaf=None):
    """Work of the fused path for parameter vectors with shifted norm nu[b] and shift mu[b].

    ``lowp_leaf[b]`` (default: none): vectors whose observed-leaf branches go through
    leaf_lowp_kernel (one FP64 column + the TF32 floor correction) instead of the FP64 tensor
    kernel; their leaf pairs are accounted under "lowp_*", not under the FP64 figures.

    algorithmic flops per fused (vector, branch): 2 n^2 ncoef (the series contraction) + 3 n^2
    (floor/row sum/dot product); executed DMMA flops follow the kernel's grouping (a warp owns
    16 consecutive branches of a level's list and runs max ceil(ncoef/4) chunks for all of
    them, on rows padded to 8 columns)."""
    n = flat.n_states
    npad = (n + 7) // 8 * 8
    nu = np.asarray(nu, dtype=np.float64)
    mu = np.asarray(mu, dtype=np.float64)
    lists = [flat.leaf_child] + [flat.level_child[flat.level_child_off[lv]:flat.level_child_off[lv + 1]]
                                 for lv in range(len(flat.level_child_off) - 1)]
    per_warp = 8 * groups_per_warp
    alg = dmma = 0.0
    n_fused = n_mat = n_ident = 0
    nc_sum = 0.0
    lowp_pairs = 0
    lowp_tf32 = lowp_fp64 = 0.0
    lowp_leaf = (np.zeros(nu.shape, dtype=bool) if lowp_leaf is None
                 else np.asarray(lowp_leaf, dtype=bool))
    for li, items in enumerate(lists):                    # one launch each, in this order
        if len(items) == 0:
            continue
        kind, nc, _ = plan_pairs(flat.dist_eff[items], nu, mu)
        fused, mat, ident = kind == 0, kind == 2, kind == 1
        if li == 0 and lowp_leaf.any():
            # leaf launch, generator vectors: 2 n terms for the FP64 column, 2 n * 104 * 8 ceil(terms/8)
            # TF32 flops for the floor correction, per pair
            lp = fused & lowp_leaf[:, None]
            ncl = np.where(lp, nc, 0)
            lowp_pairs += int(lp.sum())
            lowp_fp64 += float(np.sum(2.0 * n * ncl))
            lowp_tf32 += float(np.sum(lp * (2.0 * n * 104 * 8 * ((ncl + 7) // 8))))
            n_mat += int((mat & lowp_leaf[:, None]).sum())
            n_ident += int((ident & lowp_leaf[:, None]).sum())
            keep = ~lowp_leaf
            fused, mat, ident = fused & keep[:, None], mat & keep[:, None], ident & keep[:, None]
        nc = np.where(fused, nc, 0)
        kc = (nc + 3) // 4
        alg += float(np.sum(fused * (2.0 * n * n * nc + 3.0 * n * n)))
        pad = (-kc.shape[1]) % per_warp
        if pad:
            kc = np.concatenate([kc, np.zeros((kc.shape[0], pad), dtype=kc.dtype)], axis=1)
        wk = kc.reshape(kc.shape[0], -1, per_warp).max(axis=2)
        dmma += float(wk.sum()) * groups_per_warp * n * (npad // 8) * 512.0
        n_fused += int(fused.sum())
        n_mat += int(mat.sum())
        n_ident += int(ident.sum())
        nc_sum += float(nc[fused].sum())
    return {"algorithmic_flops": alg, "dmma_flops": dmma, "fused_pairs": n_fused,
            "materialised_pairs": n_mat, "identity_pairs": n_ident,
            "mean_ncoef": nc_sum / max(n_fused, 1),
            "lowp_pairs": lowp_pairs, "lowp_tf32_flops": lowp_tf32, "lowp_fp64_flops": lowp_fp64}
