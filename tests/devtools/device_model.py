"""Development aid: a numpy model of the arithmetic the CUDA expm kernels perform
(shared even powers of Q per parameter vector, per-branch (m, s) from scalar norms, U = W.A,
blocked Gauss-Jordan with partial pivoting, repeated squaring).  Used on the CPU-only build
box to check that this formulation stays within the parity tolerance of scipy.linalg.expm
before spending GPU time.  Not imported by the package, the tests' GPU path or bench.py.
"""
from __future__ import annotations

import math

import numpy as np

from oracle.expm_amh import ELL_C, PADE_B, THETA

ELL_P = (7, 11, 15, 19, 27)


def prepare(Q):
    """Per-parameter-vector tables: even powers Q^2..Q^12 and the scalar norms."""
    Q2 = Q @ Q
    Q4 = Q2 @ Q2
    Q6 = Q4 @ Q2
    Q8 = Q4 @ Q4
    Q10 = Q6 @ Q4
    Q12 = Q6 @ Q6
    one = lambda M: float(np.max(np.sum(np.abs(M), axis=0)))
    norms = {'n1': one(Q), 'd4': one(Q4) ** 0.25, 'd6': one(Q6) ** (1 / 6.), 'd8': one(Q8) ** 0.125,
             'd10': one(Q10) ** 0.1}
    v = np.ones(Q.shape[0])
    M = np.abs(Q).T
    for p in range(1, 28):
        v = M @ v
        if p in ELL_P:
            norms[f'N{p}'] = float(np.max(v))
    return {2: Q2, 4: Q4, 6: Q6, 8: Q8, 10: Q10, 12: Q12}, norms


def ell_scaled(norms, m, ts):
    """ell(ts*Q, m) from the per-Q scalars, evaluated in log2 space."""
    num = norms[f'N{2 * m + 1}']
    if not num or not norms['n1'] or ts <= 0:
        return 0
    lg = (2 * m) * math.log2(ts) + math.log2(num) - math.log2(norms['n1']) - math.log2(ELL_C[m]) + 53.0
    return max(int(math.ceil(lg / (2 * m))), 0)


def select(norms, t):
    d4, d6, d8, d10 = (t * norms[k] for k in ('d4', 'd6', 'd8', 'd10'))
    eta1 = max(d4, d6)
    if eta1 < THETA[3] and ell_scaled(norms, 3, t) == 0:
        return 3, 0
    if eta1 < THETA[5] and ell_scaled(norms, 5, t) == 0:
        return 5, 0
    eta3 = max(d6, d8)
    if eta3 < THETA[7] and ell_scaled(norms, 7, t) == 0:
        return 7, 0
    if eta3 < THETA[9] and ell_scaled(norms, 9, t) == 0:
        return 9, 0
    eta5 = min(eta3, max(d8, d10))
    s = 0 if eta5 == 0 else max(int(math.ceil(math.log2(eta5 / THETA[13]))), 0)
    s += ell_scaled(norms, 13, t * 2.0 ** -s)
    return 13, s


def gauss_jordan(Am, Bm, nb=8):
    """Solve Am X = Bm by blocked Gauss-Jordan, partial pivoting inside each panel."""
    n = Am.shape[0]
    M = np.concatenate([Am, Bm], axis=1).copy()
    for k0 in range(0, n, nb):
        k1 = min(k0 + nb, n)
        # panel: unblocked elimination on columns k0..k1 (pivot rows searched below k)
        for k in range(k0, k1):
            p = k + int(np.argmax(np.abs(M[k:, k])))
            if p != k:
                M[[k, p]] = M[[p, k]]
            M[k, k0:] = M[k, k0:] / M[k, k]        # (device: scales the whole remaining row)
            rows = np.r_[0:k, k + 1:n]
            f = M[rows, k].copy()
            # inside the panel only the panel columns are updated eagerly ...
            M[np.ix_(rows, range(k0, k1))] -= np.outer(f, M[k, k0:k1])
            # ... the trailing columns are updated as one rank-nb product in the kernel; the
            # model applies it eagerly, which is the same arithmetic in a different order.
            M[np.ix_(rows, range(k1, 2 * n))] -= np.outer(f, M[k, k1:])
    return M[:, n:]


def expm_model(Q, t, tables=None, solver='gj'):
    n = Q.shape[0]
    powers, norms = tables if tables is not None else prepare(Q)
    m, s = select(norms, t)
    ts = t * 2.0 ** -s
    b = PADE_B[m]
    I = np.eye(n)
    W = b[1] * I
    V = b[0] * I
    for k in range(1, m // 2 + 1):
        c = ts ** (2 * k)
        W = W + (b[2 * k + 1] * c) * powers[2 * k]
        V = V + (b[2 * k] * c) * powers[2 * k]
    U = W @ (ts * Q)
    if solver == 'gj':
        X = gauss_jordan(V - U, V + U)
    else:
        X = np.linalg.solve(V - U, V + U)
    for _ in range(s):
        X = X @ X
    return X, m, s


def postprocess(P):
    P = np.maximum(P, 1e-12)
    return P / P.sum(axis=1, keepdims=True)
