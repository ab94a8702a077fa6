"""Host-side accounting of the work one evaluation does on the device (bench.py / DESIGN.md
roofline numerators).  Nothing here is on the measured path."""
from __future__ import annotations

import numpy as np

THETA = 5.4          # CEVB_TAYLOR_THETA
KC = 10              # CEVB_TAYLOR_KC


def taylor_ncoef(eta):
    """Vectorised copy of the device rule (csrc/fused.cu: taylor_ncoef): series coefficients
    (degree + 1) for eta = t ||Q||_1."""
    eta = np.asarray(eta, dtype=np.float64)
    out = np.full(eta.shape, 4 * KC, dtype=np.int64)
    done = ~(eta > 0.0)
    out[done] = 1
    term = eta.copy()
    with np.errstate(all='ignore'):
        for k in range(1, 4 * KC - 1):
            term = term * (eta / (k + 1))
            ok = (~done) & ((k + 2) > eta) & (term / (1.0 - eta / (k + 2)) <= 1e-18)
            out[ok] = k + 1
            done |= ok
    return out


def pade_flops(n, s):
    """Canonical Pade-13 + scaling-and-squaring cost SURVEY.md section 8(d) defines per matrix."""
    return (2.0 * (6 + np.asarray(s, dtype=np.float64)) + 8.0 / 3.0) * float(n) ** 3


def fused_work(flat, norm1, groups_per_warp=1):
    """Work of the fused path for parameter vectors with ||Q||_1 = norm1[b].

    algorithmic flops per fused (vector, branch): 2 n^2 ncoef (the series contraction) + 3 n^2
    (floor/row sum/dot product); executed DMMA flops follow the kernel's grouping (a warp owns
    16 consecutive branches of a level's list and runs max ceil(ncoef/4) chunks for all of
    them, on rows padded to 8 columns)."""
    n = flat.n_states
    npad = (n + 7) // 8 * 8
    norm1 = np.asarray(norm1, dtype=np.float64)
    lists = [flat.leaf_child] + [flat.level_child[flat.level_child_off[lv]:flat.level_child_off[lv + 1]]
                                 for lv in range(len(flat.level_child_off) - 1)]
    per_warp = 8 * groups_per_warp
    alg = dmma = 0.0
    n_fused = n_mat = n_ident = 0
    nc_sum = 0.0
    for items in lists:                                   # one launch each, in this order
        if len(items) == 0:
            continue
        t = flat.dist_eff[items]
        eta = norm1[:, None] * t[None, :]
        ident = ~(t > 0.0)
        mat = (~ident)[None, :] & ~(eta <= THETA)
        fused = (~ident)[None, :] & ~mat
        nc = np.where(fused, taylor_ncoef(np.where(fused, eta, 0.0)), 0)
        kc = (nc + 3) // 4
        alg += float(np.sum(fused * (2.0 * n * n * nc + 3.0 * n * n)))
        pad = (-kc.shape[1]) % per_warp
        if pad:
            kc = np.concatenate([kc, np.zeros((kc.shape[0], pad), dtype=kc.dtype)], axis=1)
        wk = kc.reshape(kc.shape[0], -1, per_warp).max(axis=2)
        dmma += float(wk.sum()) * groups_per_warp * n * (npad // 8) * 512.0
        n_fused += int(fused.sum())
        n_mat += int(mat.sum())
        n_ident += int(ident.sum()) * norm1.shape[0]
        nc_sum += float(nc[fused].sum())
    return {"algorithmic_flops": alg, "dmma_flops": dmma, "fused_pairs": n_fused,
            "materialised_pairs": n_mat, "identity_pairs": n_ident,
            "mean_ncoef": nc_sum / max(n_fused, 1)}
