"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

Restatement of the matrix-exponential algorithm behind `scipy.linalg.expm` (the reference's
call at /root/reference/src/ancestral_reconstruction.py:151).

scipy is a third-party dependency (requirements.txt:4, unpinned; image: scipy 1.18.1) whose
dense `expm` is compiled C with no source in this image.  Its published algorithm is
Al-Mohy & Higham, "A New Scaling and Squaring Algorithm for the Matrix Exponential", SIAM J.
Matrix Anal. Appl. 31(3), 2009 (Algorithm 5.1 with exact 1-norms for n < 400), the same
algorithm that `scipy/sparse/linalg/_matfuncs.py::_expm(A, use_exact_onenorm=True)` spells out
in Python.  This file restates that algorithm in plain numpy so that
  * the Pade order m in {3,5,7,9,13} and squaring count s chosen per matrix can be observed
    (the device kernels report theirs in `info`), and
  * the result can be pinned against `scipy.linalg.expm` itself (tests/test_oracle.py).
It does not reproduce scipy's strictly-triangular special case (Code Fragment 2.1), which
only changes rounding.
"""
from __future__ import annotations

import math

import numpy as np

THETA = {3: 1.495585217958292e-002, 5: 2.539398330063230e-001, 7: 9.504178996162932e-001,
         9: 2.097847961257068e+000, 13: 4.25}

# reciprocal of the leading coefficient of the backward-error series (2005 paper eq. 2.2/2.6)
ELL_C = {3: 100800., 5: 10059033600., 7: 4487938430976000., 9: 5914384781877411840000.,
         13: 113250775606021113483283660800000000.}

PADE_B = {
    3: (120., 60., 12., 1.),
    5: (30240., 15120., 3360., 420., 30., 1.),
    7: (17297280., 8648640., 1995840., 277200., 25200., 1512., 56., 1.),
    9: (17643225600., 8821612800., 2075673600., 302702400., 30270240., 2162160., 110880.,
        3960., 90., 1.),
    13: (64764752532480000., 32382376266240000., 7771770303897600., 1187353796428800.,
         129060195264000., 10559470521600., 670442572800., 33522128640., 1323241920.,
         40840800., 960960., 16380., 182., 1.),
}


def onenorm(A):
    return float(np.max(np.sum(np.abs(A), axis=0))) if A.size else 0.0


def onenorm_abs_power(A, p):
    """|| |A|^p ||_1 through p transposed mat-vecs with the all-ones vector."""
    v = np.ones(A.shape[0])
    M = np.abs(A).T
    for _ in range(p):
        v = M @ v
    return float(np.max(v))


def ell(A, m):
    """Extra scaling demanded by the backward-error bound (paper eq. 5.1 / Alg. 5.1 `ell`)."""
    num = onenorm_abs_power(A, 2 * m + 1)
    if not num:
        return 0
    alpha = num / (onenorm(A) * ELL_C[m])
    return max(int(math.ceil(math.log2(alpha / 2.0 ** -53) / (2 * m))), 0)


def select_order(A):
    """(m, s) picked by Algorithm 5.1 with exact 1-norms, plus the powers computed so far."""
    A2 = A @ A
    A4 = A2 @ A2
    A6 = A4 @ A2
    d4 = onenorm(A4) ** (1 / 4.)
    d6 = onenorm(A6) ** (1 / 6.)
    powers = {2: A2, 4: A4, 6: A6}
    eta_1 = max(d4, d6)
    if eta_1 < THETA[3] and ell(A, 3) == 0:
        return 3, 0, powers
    eta_2 = max(d4, d6)
    if eta_2 < THETA[5] and ell(A, 5) == 0:
        return 5, 0, powers
    A8 = A6 @ A2
    powers[8] = A8
    d8 = onenorm(A8) ** (1 / 8.)
    eta_3 = max(d6, d8)
    if eta_3 < THETA[7] and ell(A, 7) == 0:
        return 7, 0, powers
    if eta_3 < THETA[9] and ell(A, 9) == 0:
        return 9, 0, powers
    d10 = onenorm(A4 @ A6) ** (1 / 10.)
    eta_4 = max(d8, d10)
    eta_5 = min(eta_3, eta_4)
    if eta_5 == 0:
        s = 0
    else:
        s = max(int(math.ceil(math.log2(eta_5 / THETA[13]))), 0)
    s = s + ell(2.0 ** -s * A, 13)
    return 13, s, powers


def pade_uv(A, m, s, powers):
    n = A.shape[0]
    I = np.eye(n)
    b = PADE_B[m]
    A2, A4, A6 = powers[2], powers[4], powers[6]
    if m == 3:
        U = A @ (b[3] * A2 + b[1] * I)
        V = b[2] * A2 + b[0] * I
    elif m == 5:
        U = A @ (b[5] * A4 + b[3] * A2 + b[1] * I)
        V = b[4] * A4 + b[2] * A2 + b[0] * I
    elif m == 7:
        U = A @ (b[7] * A6 + b[5] * A4 + b[3] * A2 + b[1] * I)
        V = b[6] * A6 + b[4] * A4 + b[2] * A2 + b[0] * I
    elif m == 9:
        A8 = powers[8]
        U = A @ (b[9] * A8 + b[7] * A6 + b[5] * A4 + b[3] * A2 + b[1] * I)
        V = b[8] * A8 + b[6] * A6 + b[4] * A4 + b[2] * A2 + b[0] * I
    else:
        B = A * 2.0 ** -s
        B2 = A2 * 2.0 ** (-2 * s)
        B4 = A4 * 2.0 ** (-4 * s)
        B6 = A6 * 2.0 ** (-6 * s)
        U2 = B6 @ (b[13] * B6 + b[11] * B4 + b[9] * B2)
        U = B @ (U2 + b[7] * B6 + b[5] * B4 + b[3] * B2 + b[1] * I)
        V2 = B6 @ (b[12] * B6 + b[10] * B4 + b[8] * B2)
        V = V2 + b[6] * B6 + b[4] * B4 + b[2] * B2 + b[0] * I
    return U, V


def expm_amh(A, return_info=False):
    A = np.asarray(A, dtype=np.float64)
    n = A.shape[0]
    if n == 0:
        out = np.zeros((0, 0))
        return (out, 0, 0) if return_info else out
    if n == 1:
        out = np.array([[math.exp(A[0, 0])]])
        return (out, 0, 0) if return_info else out
    m, s, powers = select_order(A)
    U, V = pade_uv(A, m, s, powers)
    X = np.linalg.solve(V - U, V + U)
    for _ in range(s):
        X = X @ X
    return (X, m, s) if return_info else X
