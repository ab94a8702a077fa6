// Batched FP64 matrix exponential for (parameter vector x branch) pairs:  P = post(expm(Q_b t_e)).
//
// Reference behaviour: ChromEvolutionModel.get_transition_probabilities, AR:136-166, i.e.
// scipy.linalg.expm (Al-Mohy & Higham 2009, Algorithm 5.1: Pade order m in {3,5,7,9,13} chosen
// from exact 1-norms, 2^s scaling + s squarings for m = 13), then max(P, 1e-12), row
// renormalisation, NaN -> identity.
//
// One persistent CTA per SM; a CTA owns one matrix at a time and keeps its whole working set
// (two n_pad x ldw FP64 matrices + staging) in shared memory when it fits (n <= 104), otherwise
// in a per-CTA global scratch slot that lives in L2.  Per matrix:
//   1. (m, s) from the per-vector norms scaled by t              (scalar work)
//   2. W = sum b_{2k+1} ts^{2k} Q^{2k},  V = sum b_{2k} ts^{2k} Q^{2k}   from the shared powers
//   3. U = W (ts Q)       -- Q has <= 9 non-zeros per column, so this is a sparse product
//   4. solve (V - U) X = (V + U) by blocked Gauss-Jordan on [V-U | V+U]: rank-8 updates on the
//      FP64 tensor cores (DMMA.8x8x4); pivoting inside the 8x8 diagonal tile with an
//      element-growth monitor; optional full partial pivoting (row swaps) per panel
//   5. X <- X^2, s times (DMMA)
//   6. floor, renormalise, NaN check, coalesced store of P
#include "common.cuh"

namespace cevb {

constexpr int NT = 512;
constexpr int NW = NT / 32;
constexpr int GLD = 12;      // row stride of the G staging tile (12 mod 16 -> conflict-free A frags)
constexpr int MAXJJ = 9;     // ceil(n_pad / 32) for n_pad <= 288

__constant__ double c_b3[4] = {120., 60., 12., 1.};
__constant__ double c_b5[6] = {30240., 15120., 3360., 420., 30., 1.};
__constant__ double c_b7[8] = {17297280., 8648640., 1995840., 277200., 25200., 1512., 56., 1.};
__constant__ double c_b9[10] = {17643225600., 8821612800., 2075673600., 302702400., 30270240.,
                                2162160., 110880., 3960., 90., 1.};
__constant__ double c_b13[14] = {64764752532480000., 32382376266240000., 7771770303897600.,
                                 1187353796428800., 129060195264000., 10559470521600.,
                                 670442572800., 33522128640., 1323241920., 40840800., 960960.,
                                 16380., 182., 1.};

struct BranchArgs {
    const double* pw;
    const double* norms;
    const uint16_t* csc;
    const uint8_t* csc_cnt;
    const double* t;
    double* P;
    int32_t* info;
    double* scratch;
    unsigned int* counter;
    long long M;
    int B, E, n, ldq, ldp, n8, npad, ldw, ldr, use_smem, full_pivot;
    size_t scratch_doubles;  // per CTA
    unsigned long long* phase_clk;  // optional diagnostics: SM clocks per phase (thread 0 of each CTA)
    // optional compact work list (cevb_expm_plan): queue position q handles matrix worklist[q]
    // and writes P slot q; the list length is read from the device
    const int32_t* worklist;
    const unsigned int* work_count;
    long long work_cap;
};

struct Order {
    int m, s, fail;
};

// ell(ts * Q, m) of Alg. 5.1 from the per-vector scalars, in log2 space.
__device__ int ell_scaled(const double* nr, int which, int m, double ts, int* fail) {
    const double num = nr[CEVB_NORM_ABS7 + which];
    if (num != num) {
        *fail = 1;
        return 0;
    }
    if (num == 0.0) return 0;  // "if not A_abs_onenorm: return 0"
    static const double c[5] = {100800., 10059033600., 4487938430976000.,
                                5914384781877411840000., 113250775606021113483283660800000000.};
    const double lg = (double)(2 * m) * log2(ts) + log2(num) - log2(nr[CEVB_NORM_ONE]) -
                      log2(c[which]) + 53.0;
    if (lg != lg || lg == INFINITY) {  // int(ceil(inf/nan)) raises in the reference -> identity
        *fail = 1;
        return 0;
    }
    if (lg == -INFINITY) return 0;
    const double v = ceil(lg / (double)(2 * m));
    if (v > 1000.0) {
        *fail = 1;
        return 0;
    }
    return v > 0.0 ? (int)v : 0;
}

__device__ Order select_order(const double* nr, double t) {
    Order o = {13, 0, 0};
    const double d4 = t * nr[CEVB_NORM_D4], d6 = t * nr[CEVB_NORM_D6];
    const double d8 = t * nr[CEVB_NORM_D8], d10 = t * nr[CEVB_NORM_D10];
    const double eta1 = fmax(d4, d6);
    if (d4 != d4 || d6 != d6 || d8 != d8 || d10 != d10) {
        o.fail = 1;
        return o;
    }
    if (eta1 < 1.495585217958292e-002 && ell_scaled(nr, 0, 3, t, &o.fail) == 0 && !o.fail) {
        o.m = 3;
        return o;
    }
    if (eta1 < 2.539398330063230e-001 && ell_scaled(nr, 1, 5, t, &o.fail) == 0 && !o.fail) {
        o.m = 5;
        return o;
    }
    const double eta3 = fmax(d6, d8);
    if (eta3 < 9.504178996162932e-001 && ell_scaled(nr, 2, 7, t, &o.fail) == 0 && !o.fail) {
        o.m = 7;
        return o;
    }
    if (eta3 < 2.097847961257068e+000 && ell_scaled(nr, 3, 9, t, &o.fail) == 0 && !o.fail) {
        o.m = 9;
        return o;
    }
    if (o.fail) return o;
    const double eta4 = fmax(d8, d10);
    const double eta5 = fmin(eta3, eta4);
    int s = 0;
    if (eta5 != 0.0) {
        const double v = ceil(log2(eta5 / 4.25));
        if (v != v || v > 1000.0) {
            o.fail = 1;
            return o;
        }
        s = v > 0.0 ? (int)v : 0;
    }
    s += ell_scaled(nr, 4, 13, ldexp(t, -s), &o.fail);
    o.m = 13;
    o.s = s;
    if (s > 1000) o.fail = 1;
    return o;
}

// Inverse of the 8x8 diagonal tile D by Gauss-Jordan with partial pivoting inside the tile.
// One warp; lane 4r+q holds columns 4q..4q+3 of row r of [D | I].
__device__ void tile_inverse(const double* D, int ld, double* dinv, int lane) {
    const int r = lane >> 2, q = lane & 3;
    double x[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
        x[j] = (q < 2) ? D[(size_t)r * ld + 4 * q + j] : ((4 * (q - 2) + j) == r ? 1.0 : 0.0);
    unsigned used = 0u;
    int mycol = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const double myc = __shfl_sync(0xffffffffu, x[c & 3], (lane & ~3) | (c >> 2));
        double key = ((used >> r) & 1u) ? -1.0 : ((myc != myc) ? -0.5 : fabs(myc));
        int kr = r;
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
            const double ok = __shfl_xor_sync(0xffffffffu, key, o);
            const int orr = __shfl_xor_sync(0xffffffffu, kr, o);
            if (ok > key || (ok == key && orr < kr)) {
                key = ok;
                kr = orr;
            }
        }
        const int prow = kr;
        const double pv = __shfl_sync(0xffffffffu, myc, prow << 2);
        const double inv = 1.0 / pv;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double pr = __shfl_sync(0xffffffffu, x[j], (prow << 2) | q) * inv;
            x[j] = (r == prow) ? pr : fma(-myc, pr, x[j]);
        }
        if (r == prow) mycol = c;
        used |= 1u << prow;
    }
    if (q >= 2) {
#pragma unroll
        for (int j = 0; j < 4; ++j) dinv[mycol * 8 + 4 * (q - 2) + j] = x[j];
    }
}

// C(rt, ct) += G(rt) * R(:, ct) for one 8x8 tile of the augmented matrix [Aq | Bp].
__device__ __forceinline__ void gj_update_tile(double* Aq, double* Bp, const double* rst, int ldw,
                                               int ldr, int n8, int rt, int ct, double a0,
                                               double a1, int g, int t, int sg, bool pivot_rows,
                                               double& gmax) {
    double* Cp = ((ct < n8) ? (Aq + ct * 8) : (Bp + (ct - n8) * 8)) + (size_t)(rt * 8 + g) * ldw;
    // pivot rows become Dinv * R outright (G holds Dinv there); forming them as
    // R + (Dinv - I) R would cancel two O(b0 ~ 6e16) numbers into an O(1) result.
    double c0 = pivot_rows ? 0.0 : Cp[t], c1 = pivot_rows ? 0.0 : Cp[t + 4];
    const double b0 = rst[t * ldr + ct * 8 + sg];
    const double b1 = rst[(4 + t) * ldr + ct * 8 + sg];
    dmma884(c0, c1, a0, b0);
    dmma884(c0, c1, a1, b1);
    Cp[t] = c0;
    Cp[t + 4] = c1;
    gmax = fmax(gmax, fmax(fabs(c0), fabs(c1)));
}

__device__ double block_max(double v, double* red) {
    v = warp_max(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = red[0];
#pragma unroll
    for (int i = 1; i < NW; ++i) r = fmax(r, red[i]);
    __syncthreads();
    return r;
}

#define CEVB_PHASE(k)                                                          \
    do {                                                                       \
        if (a.phase_clk != nullptr && tid == 0) {                              \
            const long long _now = clock64();                                  \
            ph[k] += (unsigned long long)(_now - ph_t);                        \
            ph_t = _now;                                                       \
        }                                                                      \
    } while (0)

// Pade order and squaring count of every (vector, branch) pair, one thread each.
__global__ void __launch_bounds__(256) order_kernel(const double* __restrict__ norms,
                                                    const double* __restrict__ t, int E, long long M,
                                                    int32_t* __restrict__ info) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M) return;
    const int b = (int)(idx / E), e = (int)(idx - (long long)b * E);
    const double tb = t[e];
    Order o = {0, 0, 0};
    if (tb > 0.0) o = select_order(norms + (size_t)b * CEVB_NNORM, tb);
    if (o.s > 200) o.fail = 1;
    info[idx] = o.m | ((o.s & 0xff) << 8) | (o.fail << 16);
}

__global__ void __launch_bounds__(NT, 1) expm_branch_kernel(BranchArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3, sg = sigma8(g);
    const int n = a.n, npad = a.npad, n8 = a.n8, ldw = a.ldw, ldr = a.ldr, ldq = a.ldq;
    const size_t matw = (size_t)npad * ldw;

    double* sm = reinterpret_cast<double*>(smem_raw);
    double *bufA, *bufB;
    if (a.use_smem) {
        bufA = sm;
        bufB = sm + matw;
        sm += 2 * matw;
    } else {
        bufA = a.scratch + (size_t)blockIdx.x * a.scratch_doubles;
        bufB = bufA + matw;
    }
    double* gst = sm;                       // npad x GLD
    double* rst = gst + (size_t)npad * GLD; // 8 x ldr
    double* dinv = rst + 8 * ldr;           // 2 x 64
    double* red = dinv + 128;               // 32
    int* ish = reinterpret_cast<int*>(red + 32);  // [0]=work idx lo, [1]=hi, [2]=m, [3]=s, [4]=fail, [5..12]=pivots
    // the sparse-column table of step 3 reuses the gst/rst region
    double* tvals = gst;
    uint16_t* trows = reinterpret_cast<uint16_t*>(tvals + (size_t)CEVB_MAX_SPARSE * n);
    uint8_t* tcnt = reinterpret_cast<uint8_t*>(trows + (size_t)CEVB_MAX_SPARSE * n);

    const size_t matq = (size_t)n * ldq;
    unsigned long long ph[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long ph_t = clock64();

    for (;;) {
        if (tid == 0) {
            const unsigned long long w = atomicAdd(a.counter, 1u);
            ish[0] = (int)(w & 0xffffffffu);
        }
        __syncthreads();
        const long long qpos = (long long)(unsigned)ish[0];
        long long idx = qpos;
        if (a.worklist != nullptr) {
            long long lim = (long long)*a.work_count;
            if (lim > a.work_cap) lim = a.work_cap;
            if (qpos >= lim) break;
            idx = a.worklist[qpos];
        } else if (qpos >= a.M) {
            break;
        }
        const int b = (int)(idx / a.E), e = (int)(idx - (long long)b * a.E);
        const double tb = a.t[e];
        const double* nr = a.norms + (size_t)b * CEVB_NNORM;
        const double* pwb = a.pw + (size_t)b * CEVB_NPOW * matq;
        double* Pout = a.P + (size_t)qpos * n * a.ldp;

        const int code = a.info[idx];   // written by order_kernel: m | s << 8 | fail << 16
        const int m = code & 0xff, s = (code >> 8) & 0xff, fail = code >> 16;
        CEVB_PHASE(0);

        if (!(tb > 0.0) || fail) {
            // t == 0 -> identity (AR:146-148); selection failure mirrors the reference's
            // exception path (AR:161-166), also identity.  (t < 0 / NaN never reach here.)
            for (int i2 = tid; i2 < n * a.ldp; i2 += NT) {
                const int i = i2 / a.ldp, j = i2 - i * a.ldp;
                Pout[i2] = (i == j) ? 1.0 : 0.0;
            }
            if (tid == 0)
                a.info[idx] = fail ? (13 | (CEVB_INFO_NAN_FALLBACK << 16))
                                   : (CEVB_INFO_IDENTITY_T0 << 16);
            __syncthreads();
            continue;
        }

        const double ts = ldexp(tb, -s);
        const int kmax = (m == 13) ? 6 : (m >> 1);
        const double* bc = (m == 3) ? c_b3 : (m == 5) ? c_b5 : (m == 7) ? c_b7 : (m == 9) ? c_b9 : c_b13;

        // ---- 2. W -> bufA, V -> bufB ------------------------------------------------------
        {
            double cw[7], cv[7];
            double pk = 1.0;
            const double ts2 = ts * ts;
#pragma unroll
            for (int k = 1; k <= 6; ++k) {
                pk *= ts2;
                cw[k] = (k <= kmax) ? bc[2 * k + 1] * pk : 0.0;
                cv[k] = (k <= kmax) ? bc[2 * k] * pk : 0.0;
            }
            const double b1 = bc[1], b0 = bc[0];
            for (int i2 = tid; i2 < npad * npad; i2 += NT) {
                const int i = i2 / npad, j = i2 - i * npad;
                double w = 0.0, v = 0.0;
                if (i < n && j < n) {
                    const double* src = pwb + (size_t)i * ldq + j;
#pragma unroll
                    for (int k = 6; k >= 1; --k) {
                        if (k <= kmax) {
                            const double x = src[(size_t)k * matq];
                            w = fma(cw[k], x, w);
                            v = fma(cv[k], x, v);
                        }
                    }
                }
                if (i == j) {
                    w += b1;
                    v += b0;
                }
                bufA[(size_t)i * ldw + j] = w;
                bufB[(size_t)i * ldw + j] = v;
            }
        }
        CEVB_PHASE(1);
        // stage the sparse structure of A = ts * Q (values pre-scaled)
        const bool dense = nr[CEVB_NORM_DENSE] != 0.0;
        if (!dense) {
            for (int j = tid; j < n; j += NT) tcnt[j] = a.csc_cnt[(size_t)b * 2 * n + j];
            for (int i2 = tid; i2 < CEVB_MAX_SPARSE * n; i2 += NT) {
                const int c = i2 / n, j = i2 - c * n;
                const int cnt = a.csc_cnt[(size_t)b * 2 * n + j];
                uint16_t r = 0;
                double v = 0.0;
                if (c < cnt) {
                    r = a.csc[((size_t)b * 2 * CEVB_MAX_SPARSE + c) * n + j];
                    v = ts * pwb[(size_t)r * ldq + j];
                }
                trows[i2] = r;
                tvals[i2] = v;
            }
        }
        __syncthreads();
        CEVB_PHASE(2);

        // ---- 3. U = W * (ts Q), in place in bufA, one warp per row -------------------------
        for (int i = warp; i < npad; i += NW) {
            double* wrow = bufA + (size_t)i * ldw;
            double out[MAXJJ];
#pragma unroll
            for (int jj = 0; jj < MAXJJ; ++jj) {
                const int j = lane + 32 * jj;
                double acc = 0.0;
                if (i < n && j < n) {
                    if (!dense) {
                        const int cnt = tcnt[j];
                        for (int c = 0; c < cnt; ++c)
                            acc = fma(wrow[trows[c * n + j]], tvals[c * n + j], acc);
                    } else {
                        for (int k = 0; k < n; ++k)
                            acc = fma(wrow[k], ts * pwb[(size_t)k * ldq + j], acc);
                    }
                }
                out[jj] = acc;
            }
            __syncwarp();
#pragma unroll
            for (int jj = 0; jj < MAXJJ; ++jj) {
                const int j = lane + 32 * jj;
                if (j < npad) wrow[j] = out[jj];
            }
        }
        __syncthreads();
        CEVB_PHASE(3);

        // ---- 4. Aq = V - U (bufA), Bp = V + U (bufB) --------------------------------------
        double* Aq = bufA;
        double* Bp = bufB;
        double gmax0 = 0.0;
        for (int i2 = tid; i2 < npad * npad; i2 += NT) {
            const int i = i2 / npad, j = i2 - i * npad;
            const size_t o = (size_t)i * ldw + j;
            const double u = Aq[o], v = Bp[o];
            const double qm = v - u, pm = v + u;
            Aq[o] = qm;
            Bp[o] = pm;
            gmax0 = fmax(gmax0, fmax(fabs(qm), fabs(pm)));
        }
        __syncthreads();

        CEVB_PHASE(4);
        // ---- 5. blocked Gauss-Jordan ------------------------------------------------------
        double gmax = 0.0;
        if (!a.full_pivot) {
            if (warp == 0) tile_inverse(Aq, ldw, dinv, lane);
            __syncthreads();
        }
        for (int kt = 0; kt < n8; ++kt) {
            const int k0 = kt * 8;
            double* dcur = dinv + (kt & 1) * 64;
            if (a.full_pivot) {
                // choose the 8 pivot rows of this panel by partial pivoting over all rows >= k0
                // (LU on a copy of the panel), swap them into the diagonal tile, then invert it.
                dcur = dinv;
                for (int i2 = tid; i2 < (npad - k0) * 8; i2 += NT) {
                    const int i = k0 + (i2 >> 3), c = i2 & 7;
                    gst[i * GLD + c] = Aq[(size_t)i * ldw + k0 + c];
                }
                __syncthreads();
                if (warp == 0) {
                    for (int c = 0; c < 8; ++c) {
                        double best = -1.0;
                        int bi = k0 + c;
                        for (int i = k0 + c + lane; i < npad; i += 32) {
                            const double v = fabs(gst[i * GLD + c]);
                            if (v > best) {
                                best = v;
                                bi = i;
                            }
                        }
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                            if (ob > best || (ob == best && oi < bi)) {
                                best = ob;
                                bi = oi;
                            }
                        }
                        if (lane == 0) ish[5 + c] = bi;
                        if (bi != k0 + c && lane < 8) {
                            const double tmp = gst[bi * GLD + lane];
                            gst[bi * GLD + lane] = gst[(k0 + c) * GLD + lane];
                            gst[(k0 + c) * GLD + lane] = tmp;
                        }
                        __syncwarp();
                        const double pv = gst[(k0 + c) * GLD + c];
                        for (int i = k0 + c + 1 + lane; i < npad; i += 32) {
                            const double l = gst[i * GLD + c] / pv;
                            for (int c2 = c + 1; c2 < 8; ++c2)
                                gst[i * GLD + c2] = fma(-l, gst[(k0 + c) * GLD + c2], gst[i * GLD + c2]);
                        }
                        __syncwarp();
                    }
                }
                __syncthreads();
                for (int c = 0; c < 8; ++c) {
                    const int p = ish[5 + c];
                    if (p != k0 + c) {
                        for (int cc = k0 + tid; cc < 2 * npad; cc += NT) {
                            double* r1 = (cc < npad) ? (Aq + cc) : (Bp + (cc - npad));
                            const double tmp = r1[(size_t)p * ldw];
                            r1[(size_t)p * ldw] = r1[(size_t)(k0 + c) * ldw];
                            r1[(size_t)(k0 + c) * ldw] = tmp;
                        }
                    }
                    __syncthreads();
                }
                if (warp == 0) tile_inverse(Aq + (size_t)k0 * ldw + k0, ldw, dcur, lane);
                __syncthreads();
            }
            // G = -(column panel) * Dinv, with Dinv itself in the pivot rows; R = pivot rows
            for (int i2 = tid; i2 < npad * 8; i2 += NT) {
                const int i = i2 >> 3, c = i2 & 7;
                double gv;
                if (i >= k0 && i < k0 + 8) {
                    gv = dcur[(i - k0) * 8 + c];
                } else {
                    const double* ar = Aq + (size_t)i * ldw + k0;
                    double acc = 0.0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc = fma(ar[j], dcur[j * 8 + c], acc);
                    gv = -acc;
                }
                gst[i * GLD + c] = gv;
            }
            {
                const int c_lo = k0 + 8, wcols = 2 * npad - c_lo;
                for (int i2 = tid; i2 < 8 * wcols; i2 += NT) {
                    const int r = i2 / wcols, cc = c_lo + (i2 - r * wcols);
                    rst[r * ldr + cc] = (cc < npad) ? Aq[(size_t)(k0 + r) * ldw + cc]
                                                    : Bp[(size_t)(k0 + r) * ldw + cc - npad];
                }
            }
            __syncthreads();
            CEVB_PHASE(5);

            // trailing update: tiles (rt, ct), ct in (kt, 2*n8)
            const int ct_lo = kt + 1;
            const bool lookahead = (!a.full_pivot) && (kt + 1 < n8);
            if (warp == 0 && lookahead) {
                const int rt = kt + 1;
                const double a0 = gst[(rt * 8 + g) * GLD + t4], a1 = gst[(rt * 8 + g) * GLD + 4 + t4];
                gj_update_tile(Aq, Bp, rst, ldw, ldr, n8, rt, kt + 1, a0, a1, g, t4, sg, false, gmax);
                __syncwarp();
                tile_inverse(Aq + (size_t)(k0 + 8) * ldw + k0 + 8, ldw, dinv + ((kt + 1) & 1) * 64, lane);
            }
            {
                const int nct = 2 * n8 - ct_lo;
                const int nch = (nct + 3) >> 2;
                const int ntask = n8 * nch;
                const int w0 = lookahead ? 1 : 0;  // warp 0 is busy with the next diagonal tile
                if (warp >= w0) {
                    for (int task = warp - w0; task < ntask; task += NW - w0) {
                        const int rt = task % n8, ch = task / n8;
                        const double a0 = gst[(rt * 8 + g) * GLD + t4];
                        const double a1 = gst[(rt * 8 + g) * GLD + 4 + t4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int ct = ct_lo + ch * 4 + q;
                            if (ct < 2 * n8 && !(lookahead && rt == kt + 1 && ct == kt + 1))
                                gj_update_tile(Aq, Bp, rst, ldw, ldr, n8, rt, ct, a0, a1, g, t4, sg,
                                               rt == kt, gmax);
                        }
                    }
                }
            }
            __syncthreads();
            CEVB_PHASE(6);
        }

        // ---- 6. squarings -----------------------------------------------------------------
        double* src = Bp;
        double* dst = Aq;
        for (int it = 0; it < s; ++it) {
            const int nch = (n8 + 3) >> 2;
            const int ntask = n8 * nch;
            for (int task = warp; task < ntask; task += NW) {
                const int rt = task % n8, ch = task / n8;
                const int ct0 = ch * 4;
                double acc[4][2] = {};
                const double* arow = src + (size_t)(rt * 8 + g) * ldw + t4;
                const double* bcol = src + (size_t)t4 * ldw + ct0 * 8 + sg;
                for (int k4 = 0; k4 < npad; k4 += 4) {
                    const double av = arow[k4];
                    const double* bp = bcol + (size_t)k4 * ldw;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (ct0 + q < n8) dmma884(acc[q][0], acc[q][1], av, bp[q * 8]);
                    }
                }
                double* drow = dst + (size_t)(rt * 8 + g) * ldw + ct0 * 8 + t4;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (ct0 + q < n8) {
                        drow[q * 8] = acc[q][0];
                        drow[q * 8 + 4] = acc[q][1];
                    }
                }
            }
            __syncthreads();
            double* tmp = src;
            src = dst;
            dst = tmp;
        }

        CEVB_PHASE(7);
        // ---- 7. floor, renormalise, NaN check, store (AR:153-159) --------------------------
        int nanflag = 0;
        for (int i = warp; i < n; i += NW) {
            const double* xr = src + (size_t)i * ldw;
            double p[MAXJJ];
            double rs = 0.0;
#pragma unroll
            for (int jj = 0; jj < MAXJJ; ++jj) {
                const int j = lane + 32 * jj;
                double v = 0.0;
                if (j < n) {
                    const double x = xr[j];
                    v = (x != x) ? x : fmax(x, 1e-12);
                }
                p[jj] = v;
                rs += v;
            }
            rs = warp_sum(rs);
#pragma unroll
            for (int jj = 0; jj < MAXJJ; ++jj) {
                const int j = lane + 32 * jj;
                if (j < n) {
                    const double o = p[jj] / rs;
                    if (o != o) nanflag = 1;
                    Pout[(size_t)i * a.ldp + j] = o;
                } else if (j < a.ldp) {
                    Pout[(size_t)i * a.ldp + j] = 0.0;
                }
            }
        }
        nanflag = __syncthreads_or(nanflag);
        if (nanflag) {
            for (int i2 = tid; i2 < n * a.ldp; i2 += NT) {
                const int i = i2 / a.ldp, j = i2 - i * a.ldp;
                Pout[i2] = (i == j) ? 1.0 : 0.0;
            }
        }
        const double g0 = block_max(gmax0, red);
        const double g1 = block_max(gmax, red);
        if (tid == 0) {
            int flags = nanflag ? CEVB_INFO_NAN_FALLBACK : 0;
            if (g1 > 1.0e3 * g0 && g1 < INFINITY) flags |= CEVB_INFO_GROWTH;
            if (a.full_pivot) flags |= CEVB_INFO_REPIVOTED;
            a.info[idx] = m | (s << 8) | (flags << 16);
        }
        __syncthreads();
        CEVB_PHASE(8);
    }
    if (a.phase_clk != nullptr && tid == 0) {
        for (int k = 0; k < 10; ++k) atomicAdd(&a.phase_clk[k], ph[k]);
    }
}


// =========================================================================================
// Fast path, n <= 104: everything in shared memory, shared-space pointers (LDS/STS), no integer
// divisions in the loops, 16-byte accesses where rows allow it.
//   - working matrices have row stride ldw = 8 (mod 16) doubles, so the 16-byte C-fragment
//     accesses of the Gauss-Jordan update (row g, columns 2t..2t+1) are bank-conflict free;
//     G (stride 12) and R (stride 4 mod 16) staging tiles serve the A / B fragments.
//   - U = (ts Q) W is formed row-wise from the compressed rows of Q into REGISTERS (7 rows x 4
//     column slots per warp), then V -+ U are written straight over W and V.
//   - the diagonal tile is inverted without a pivot search (one warp, shuffles only); the growth
//     monitor and a relative pivot-size check guard that choice (full_pivot launches use the
//     generic kernel above).
__device__ __forceinline__ int hi_abs(double v) { return __double2hiint(v) & 0x7fffffff; }

// 1/x to full double precision for normal, finite x: hardware seed (about 20 bits) + Newton.
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}

__device__ int block_max_int(int v, int* red) {
    v = __reduce_max_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    int r = red[0];
#pragma unroll
    for (int i = 1; i < NW; ++i) r = max(r, red[i]);
    __syncthreads();
    return r;
}

constexpr int FR = 7;   // rows per warp kept in registers (npad <= 112)
constexpr int FJ = 4;   // 32-wide column slots per row (npad <= 128)

__device__ __forceinline__ int tile_inverse_fast(const double* D, int ld, double* dinv, int lane) {
    const int r = lane >> 2, q = lane & 3;
    double x[4];
    double dmax = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        x[j] = (q < 2) ? D[r * ld + 4 * q + j] : ((4 * (q - 2) + j) == r ? 1.0 : 0.0);
        if (q < 2) dmax = fmax(dmax, fabs(x[j]));
    }
    dmax = warp_max(dmax);
    int bad = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const double myc = __shfl_sync(0xffffffffu, x[c & 3], (lane & ~3) | (c >> 2));
        const double pv = __shfl_sync(0xffffffffu, x[c & 3], (c << 2) | (c >> 2));
        if (!(fabs(pv) >= 1e-6 * dmax)) bad = 1;
        const double inv = fast_rcp(pv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double pr = __shfl_sync(0xffffffffu, x[j], (c << 2) | q) * inv;
            x[j] = (r == c) ? pr : fma(-myc, pr, x[j]);
        }
    }
    if (q >= 2) {
#pragma unroll
        for (int j = 0; j < 4; ++j) dinv[r * 8 + 4 * (q - 2) + j] = x[j];
    }
    return bad;
}

__global__ void __launch_bounds__(NT, 1) expm_branch_fast_kernel(BranchArgs a) {
    extern __shared__ __align__(16) double smf[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int n = a.n, npad = a.npad, n8 = a.n8, ldw = a.ldw, ldr = a.ldr, ldq = a.ldq, ldp = a.ldp;
    // augmented matrix [A | B]: npad rows, row stride ldw = 2*npad + 8 (= 8 mod 16 doubles)
    double* const bufA = smf;
    double* const bufB = smf + npad;
    double* const gst = smf + npad * ldw;     // npad x GLD
    double* const rst = gst + npad * GLD;     // 8 x ldr
    double* const dinv = rst + 8 * ldr;       // 2 x 64
    double* const red = dinv + 128;           // 32
    int* const ish = reinterpret_cast<int*>(red + 32);
    // compressed-row table of ts*Q, staged in the (not yet used) gst/rst region
    double* const tval = gst;
    uint16_t* const tidx = reinterpret_cast<uint16_t*>(tval + CEVB_MAX_SPARSE * n);
    uint8_t* const tcnt = reinterpret_cast<uint8_t*>(tidx + CEVB_MAX_SPARSE * n);

    const size_t matq = (size_t)n * ldq;
    __shared__ unsigned long long ph[10];
    __shared__ long long ph_t;
    if (tid == 0) {
        for (int k = 0; k < 10; ++k) ph[k] = 0;
        ph_t = clock64();
    }

    for (;;) {
        if (tid == 0) ish[0] = (int)atomicAdd(a.counter, 1u);
        __syncthreads();
        const long long qpos = (long long)(unsigned)ish[0];
        long long idx = qpos;
        if (a.worklist != nullptr) {
            long long lim = (long long)*a.work_count;
            if (lim > a.work_cap) lim = a.work_cap;
            if (qpos >= lim) break;
            idx = a.worklist[qpos];
        } else if (qpos >= a.M) {
            break;
        }
        const int b = (int)(idx / a.E), e = (int)(idx - (long long)b * a.E);
        const double tb = a.t[e];
        const double* __restrict__ pwb = a.pw + (size_t)b * CEVB_NPOW * matq;
        double* __restrict__ Pout = a.P + (size_t)qpos * n * ldp;
        const int code = a.info[idx];
        const int m = code & 0xff, s = (code >> 8) & 0xff, fail = code >> 16;
        CEVB_PHASE(0);

        if (!(tb > 0.0) || fail) {
            for (int i = warp; i < n; i += NW)
                for (int j = lane; j < ldp; j += 32) Pout[(size_t)i * ldp + j] = (i == j) ? 1.0 : 0.0;
            if (tid == 0)
                a.info[idx] = fail ? (13 | (CEVB_INFO_NAN_FALLBACK << 16)) : (CEVB_INFO_IDENTITY_T0 << 16);
            __syncthreads();
            continue;
        }

        const double ts = ldexp(tb, -s);
        const int kmax = (m == 13) ? 6 : (m >> 1);
        const double* bc = (m == 3) ? c_b3 : (m == 5) ? c_b5 : (m == 7) ? c_b7 : (m == 9) ? c_b9 : c_b13;

        // ---- W -> bufA, V -> bufB: 16-byte loads of the shared powers, all in flight -------
        {
            double cw[7], cv[7];
            double pk = 1.0;
            const double ts2 = ts * ts;
#pragma unroll
            for (int k = 1; k <= 6; ++k) {
                pk *= ts2;
                cw[k] = (k <= kmax) ? bc[2 * k + 1] * pk : 0.0;
                cv[k] = (k <= kmax) ? bc[2 * k] * pk : 0.0;
            }
            const double b1 = bc[1], b0 = bc[0];
            for (int item = warp; item < 2 * npad; item += NW) {
                const int i = item >> 1, pass = item & 1;
                const int j = 2 * lane + 64 * pass;
                if (j >= npad) continue;
                double2 xx[6];
                const bool in = (i < n) && (j < ldq);
                const double2* src = reinterpret_cast<const double2*>(pwb + (size_t)i * ldq + j);
#pragma unroll
                for (int k = 1; k <= 6; ++k)
                    xx[k - 1] = (in && k <= kmax) ? __ldg(src + (size_t)k * (matq >> 1)) : make_double2(0.0, 0.0);
                double w0 = 0.0, w1 = 0.0, v0 = 0.0, v1 = 0.0;
#pragma unroll
                for (int k = 6; k >= 1; --k) {
                    w0 = fma(cw[k], xx[k - 1].x, w0);
                    w1 = fma(cw[k], xx[k - 1].y, w1);
                    v0 = fma(cv[k], xx[k - 1].x, v0);
                    v1 = fma(cv[k], xx[k - 1].y, v1);
                }
                if (i == j) { w0 += b1; v0 += b0; }
                if (i == j + 1) { w1 += b1; v1 += b0; }
                *reinterpret_cast<double2*>(bufA + i * ldw + j) = make_double2(w0, w1);
                *reinterpret_cast<double2*>(bufB + i * ldw + j) = make_double2(v0, v1);
            }
        }
        CEVB_PHASE(1);
        // ---- stage the compressed rows of ts*Q ------------------------------------------------
        const bool dense = a.norms[(size_t)b * CEVB_NNORM + CEVB_NORM_DENSE] != 0.0;
        if (!dense) {
            const uint16_t* gidx = a.csc + ((size_t)b * 2 + 1) * CEVB_MAX_SPARSE * n;
            const uint8_t* gcnt = a.csc_cnt + ((size_t)b * 2 + 1) * n;
            for (int i = tid; i < n; i += NT) tcnt[i] = gcnt[i];
            for (int c = warp; c < CEVB_MAX_SPARSE; c += NW) {
                for (int i = lane; i < n; i += 32) {
                    const int cnt = gcnt[i];
                    uint16_t k = 0;
                    double v = 0.0;
                    if (c < cnt) {
                        k = gidx[c * n + i];
                        v = ts * pwb[(size_t)i * ldq + k];
                    }
                    tidx[c * n + i] = k;
                    tval[c * n + i] = v;
                }
            }
        }
        __syncthreads();
        CEVB_PHASE(2);

        // ---- U = (ts Q) W, rows held in registers ---------------------------------------------
        double u[FR][FJ];
#pragma unroll
        for (int r = 0; r < FR; ++r) {
            const int i = warp + NW * r;
#pragma unroll
            for (int jj = 0; jj < FJ; ++jj) u[r][jj] = 0.0;
            if (i < n) {
                if (!dense) {
                    const int cnt = tcnt[i];
                    for (int c = 0; c < cnt; ++c) {
                        const double av = tval[c * n + i];
                        const double* wr = bufA + (int)tidx[c * n + i] * ldw + lane;
#pragma unroll
                        for (int jj = 0; jj < FJ; ++jj)
                            if (lane + 32 * jj < npad) u[r][jj] = fma(av, wr[32 * jj], u[r][jj]);
                    }
                } else {
                    for (int k = 0; k < n; ++k) {
                        const double av = ts * pwb[(size_t)i * ldq + k];
                        const double* wr = bufA + k * ldw + lane;
#pragma unroll
                        for (int jj = 0; jj < FJ; ++jj)
                            if (lane + 32 * jj < npad) u[r][jj] = fma(av, wr[32 * jj], u[r][jj]);
                    }
                }
            }
        }
        __syncthreads();
        CEVB_PHASE(3);
        // ---- Aq = V - U -> bufA, Bp = V + U -> bufB -------------------------------------------
        double* const Aq = bufA;
        double* const Bp = bufB;
        int gmax0 = 0;   // high words of |x|: a monotone integer proxy for magnitude (no FP64 pipe)
#pragma unroll
        for (int r = 0; r < FR; ++r) {
            const int i = warp + NW * r;
            if (i < npad) {
#pragma unroll
                for (int jj = 0; jj < FJ; ++jj) {
                    const int j = lane + 32 * jj;
                    if (j < npad) {
                        const double v = Bp[i * ldw + j];
                        const double qm = v - u[r][jj], pm = v + u[r][jj];
                        Aq[i * ldw + j] = qm;
                        Bp[i * ldw + j] = pm;
                        gmax0 = max(gmax0, max(hi_abs(qm), hi_abs(pm)));
                    }
                }
            }
        }
        __syncthreads();
        CEVB_PHASE(4);

        // ---- blocked Gauss-Jordan -------------------------------------------------------------
        int gmax = 0;
        int pivbad = 0;
        if (warp == 0) pivbad |= tile_inverse_fast(Aq, ldw, dinv, lane);
        __syncthreads();
        for (int kt = 0; kt < n8; ++kt) {
            const int k0 = kt * 8;
            const double* dcur = dinv + (kt & 1) * 64;
            // G = -(column panel) * Dinv (Dinv itself in the pivot rows)
            for (int i2 = tid; i2 < npad * 8; i2 += NT) {
                const int i = i2 >> 3, c = i2 & 7;
                double gv;
                if ((i >> 3) == kt) {
                    gv = dcur[(i - k0) * 8 + c];
                } else {
                    const double2* ar = reinterpret_cast<const double2*>(Aq + i * ldw + k0);
                    const double2 x0 = ar[0], x1 = ar[1], x2 = ar[2], x3 = ar[3];
                    double acc = x0.x * dcur[c];
                    acc = fma(x0.y, dcur[8 + c], acc);
                    acc = fma(x1.x, dcur[16 + c], acc);
                    acc = fma(x1.y, dcur[24 + c], acc);
                    acc = fma(x2.x, dcur[32 + c], acc);
                    acc = fma(x2.y, dcur[40 + c], acc);
                    acc = fma(x3.x, dcur[48 + c], acc);
                    acc = fma(x3.y, dcur[56 + c], acc);
                    gv = -acc;
                }
                gst[i * GLD + c] = gv;
            }
            // R = pivot rows, columns (k0 + 8) .. 2 npad of the augmented matrix
            {
                const int r = tid >> 6, c0 = tid & 63;
                const double* ra = Aq + (k0 + r) * ldw;
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const int cc = k0 + 8 + c0 + 64 * p;
                    if (cc < 2 * npad) rst[r * ldr + cc] = ra[cc];
                }
            }
            __syncthreads();
            CEVB_PHASE(5);

            const int ct_lo = kt + 1;
            const bool lookahead = (kt + 1 < n8);
            const double* const gA = gst + g * GLD + t4;        // A fragments: + rt*8*GLD (+4)
            const double* const rB = rst + t4 * ldr + g;        // B fragments: + ct*8 (+4*ldr)
            double* const cB = Aq + g * ldw + 2 * t4;           // C fragments: + rt*8*ldw + ct*8
            if (warp == 0 && lookahead) {
                const int rt = kt + 1;
                const double a0 = gA[rt * 8 * GLD], a1 = gA[rt * 8 * GLD + 4];
                double2* cp = reinterpret_cast<double2*>(cB + rt * 8 * ldw + rt * 8);
                double2 c = *cp;
                dmma884(c.x, c.y, a0, rB[rt * 8]);
                dmma884(c.x, c.y, a1, rB[4 * ldr + rt * 8]);
                *cp = c;
                gmax = max(gmax, max(hi_abs(c.x), hi_abs(c.y)));
                __syncwarp();
                pivbad |= tile_inverse_fast(Aq + (k0 + 8) * ldw + k0 + 8, ldw, dinv + ((kt + 1) & 1) * 64, lane);
            }
            {
                const int nct = 2 * n8 - ct_lo;
                const int nch = (nct + 3) >> 2;
                const int w0 = lookahead ? 1 : 0;
                const int step = NW - w0;
                if (warp >= w0) {
                    int task = warp - w0;
                    int ch = task / n8, rt = task - ch * n8;
                    while (ch < nch) {
                        const int ct0 = ct_lo + ch * 4;
                        const double a0 = gA[rt * 8 * GLD], a1 = gA[rt * 8 * GLD + 4];
                        double* const cp = cB + rt * 8 * ldw + ct0 * 8;
                        const double* const bp = rB + ct0 * 8;
                        const bool piv = (rt == kt);
                        int q_lo = (lookahead && rt == kt + 1 && ch == 0) ? 1 : 0;
                        const int q_hi = min(4, 2 * n8 - ct0);
                        if (q_lo == 0 && q_hi == 4) {
                            double2 c[4];
                            double b0[4], b1[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                c[q] = piv ? make_double2(0.0, 0.0) : *reinterpret_cast<double2*>(cp + q * 8);
                                b0[q] = bp[q * 8];
                                b1[q] = bp[4 * ldr + q * 8];
                            }
#pragma unroll
                            for (int q = 0; q < 4; ++q) dmma884(c[q].x, c[q].y, a0, b0[q]);
#pragma unroll
                            for (int q = 0; q < 4; ++q) dmma884(c[q].x, c[q].y, a1, b1[q]);
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                *reinterpret_cast<double2*>(cp + q * 8) = c[q];
                                gmax = max(gmax, max(hi_abs(c[q].x), hi_abs(c[q].y)));
                            }
                        } else {
                            for (int q = q_lo; q < q_hi; ++q) {
                                double2 c = piv ? make_double2(0.0, 0.0) : *reinterpret_cast<double2*>(cp + q * 8);
                                dmma884(c.x, c.y, a0, bp[q * 8]);
                                dmma884(c.x, c.y, a1, bp[4 * ldr + q * 8]);
                                *reinterpret_cast<double2*>(cp + q * 8) = c;
                                gmax = max(gmax, max(hi_abs(c.x), hi_abs(c.y)));
                            }
                        }
                        rt += step;
                        while (rt >= n8) {
                            rt -= n8;
                            ++ch;
                        }
                    }
                }
            }
            __syncthreads();
            CEVB_PHASE(6);
        }

        // ---- squarings ------------------------------------------------------------------------
        double* src = Bp;
        double* dst = Aq;
        for (int it = 0; it < s; ++it) {
            const int nch = (n8 + 3) >> 2;
            const int ntask = n8 * nch;
            for (int task = warp; task < ntask; task += NW) {
                const int ch = task / n8, rt = task - ch * n8;
                const int ct0 = ch * 4;
                double acc[4][2] = {};
                const double* arow = src + (rt * 8 + g) * ldw + t4;
                const double* bcol = src + t4 * ldw + ct0 * 8 + g;
                for (int k4 = 0; k4 < npad; k4 += 4) {
                    const double av = arow[k4];
                    const double* bp = bcol + k4 * ldw;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (ct0 + q < n8) dmma884(acc[q][0], acc[q][1], av, bp[q * 8]);
                }
                double* drow = dst + (rt * 8 + g) * ldw + ct0 * 8 + 2 * t4;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (ct0 + q < n8)
                        *reinterpret_cast<double2*>(drow + q * 8) = make_double2(acc[q][0], acc[q][1]);
            }
            __syncthreads();
            double* tmp = src;
            src = dst;
            dst = tmp;
        }
        CEVB_PHASE(7);

        // ---- floor, renormalise, NaN check, store (AR:153-159) ----------------------------------
        int nanflag = 0;
        for (int i = warp; i < n; i += NW) {
            const double* xr = src + i * ldw;
            double2 p[2];
            double rs = 0.0;
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                const int j = 2 * lane + 64 * pass;
                double2 v = make_double2(0.0, 0.0);
                if (j < ldp) {
                    const double2 x = *reinterpret_cast<const double2*>(xr + j);
                    v.x = (x.x != x.x) ? x.x : fmax(x.x, 1e-12);
                    if (j + 1 < n) v.y = (x.y != x.y) ? x.y : fmax(x.y, 1e-12);
                }
                p[pass] = v;
                rs += v.x + v.y;
            }
            rs = warp_sum(rs);
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                const int j = 2 * lane + 64 * pass;
                if (j < ldp) {
                    double2 o;
                    o.x = p[pass].x / rs;
                    o.y = (j + 1 < n) ? p[pass].y / rs : 0.0;
                    if (o.x != o.x || o.y != o.y) nanflag = 1;
                    *reinterpret_cast<double2*>(Pout + (size_t)i * ldp + j) = o;
                }
            }
        }
        nanflag = __syncthreads_or(nanflag);
        if (nanflag) {
            for (int i = warp; i < n; i += NW)
                for (int j = lane; j < ldp; j += 32) Pout[(size_t)i * ldp + j] = (i == j) ? 1.0 : 0.0;
        }
        const int g0i = block_max_int(gmax0, ish + 16);
        const int g1i = block_max_int(gmax, ish + 16);
        if (warp == 0 && lane == 0) {
            int flags = nanflag ? CEVB_INFO_NAN_FALLBACK : 0;
            // upper bound of the largest |entry| seen vs lower bound of the initial maximum
            const double g1 = __hiloint2double(g1i, (int)0xffffffff), g0 = __hiloint2double(g0i, 0);
            if ((g1i < 0x7ff00000 && g1 > 1.0e3 * g0) || pivbad) flags |= CEVB_INFO_GROWTH;
            a.info[idx] = m | (s << 8) | (flags << 16);
        }
        __syncthreads();
        CEVB_PHASE(8);
    }
    if (a.phase_clk != nullptr && tid == 0) {
        for (int k = 0; k < 10; ++k) atomicAdd(&a.phase_clk[k], ph[k]);
    }
}

static int fast_ldw(int npad) { return 2 * npad + 8; }

static size_t fast_smem_bytes(int n) {
    const int n8 = (n + 7) / 8, npad = n8 * 8, ldw = fast_ldw(npad), ldr = 2 * npad + 4;
    const size_t d = (size_t)npad * ldw + (size_t)npad * GLD + 8 * (size_t)ldr + 128 + 32;
    return d * 8 + 128;
}

static size_t branch_smem_bytes(int n, bool with_matrices) {
    const int n8 = (n + 7) / 8, npad = n8 * 8, ldw = npad + 4, ldr = 2 * npad + 4;
    size_t d = (size_t)npad * GLD + 8 * (size_t)ldr + 128 + 32;
    // the sparse table (values + u16 rows + u8 counts) must fit in the gst+rst region
    const size_t tab = (size_t)CEVB_MAX_SPARSE * n * 10 + n + 16;
    size_t stage = ((size_t)npad * GLD + 8 * (size_t)ldr) * 8;
    if (tab > stage) d += (tab - stage + 7) / 8;
    if (with_matrices) d += 2 * (size_t)npad * ldw;
    return d * 8 + 64;
}

static int g_sm_count = 0;
static unsigned long long* g_phase_clk = nullptr;
static size_t g_smem_optin = 0;

static int query_device() {
    if (g_sm_count) return 0;
    int dev = 0;
    CEVB_CHECK_CUDA(cudaGetDevice(&dev));
    int sm = 0, optin = 0;
    CEVB_CHECK_CUDA(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev));
    CEVB_CHECK_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    g_sm_count = sm;
    g_smem_optin = (size_t)optin;
    return 0;
}

}  // namespace cevb

using namespace cevb;

extern "C" size_t cevb_expm_scratch_bytes(int n) {
    if (query_device() != 0) return 0;
    if (branch_smem_bytes(n, true) <= g_smem_optin) return 0;
    const int n8 = (n + 7) / 8, npad = n8 * 8, ldw = npad + 4;
    return 2 * (size_t)npad * ldw * sizeof(double);
}

extern "C" int cevb_expm_grid(int n, long long n_matrices) {
    if (query_device() != 0) return 0;
    const bool in_smem = branch_smem_bytes(n, true) <= g_smem_optin;
    long long ctas = (long long)g_sm_count * (in_smem ? 1 : 2);
    if (n_matrices < ctas) ctas = n_matrices;
    return (int)(ctas < 1 ? 1 : ctas);
}

static int expm_branches_impl(const double* pw, const double* norms, const uint16_t* csc,
                              const uint8_t* csc_cnt, const double* t, int B, int E, int n,
                              double* P, int32_t* info, void* scratch, unsigned int* counter,
                              int full_pivot, const int32_t* worklist,
                              const unsigned int* work_count, long long work_cap, void* stream) {
    if (B <= 0 || E <= 0) return 0;
    if (n < 2 || n > 288) {
        set_error("cevb_expm_branches: n=%d outside the supported range 2..288", n);
        return -1;
    }
    if (query_device() != 0) return -2;
    cudaStream_t st = (cudaStream_t)stream;
    BranchArgs a;
    a.pw = pw; a.norms = norms; a.csc = csc; a.csc_cnt = csc_cnt; a.t = t; a.P = P; a.info = info;
    a.scratch = (double*)scratch; a.counter = counter;
    a.M = (long long)B * E; a.B = B; a.E = E; a.n = n;
    a.ldq = ldq_of(n); a.ldp = ldp_of(n);
    a.n8 = (n + 7) / 8; a.npad = a.n8 * 8; a.ldw = a.npad + 4; a.ldr = 2 * a.npad + 4;
    a.full_pivot = full_pivot ? 1 : 0;
    a.use_smem = branch_smem_bytes(n, true) <= g_smem_optin ? 1 : 0;
    a.scratch_doubles = 2 * (size_t)a.npad * a.ldw;
    a.phase_clk = g_phase_clk;
    a.worklist = worklist; a.work_count = work_count; a.work_cap = work_cap;
    CEVB_CHECK_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned int), st));
    {
        const long long blocks = (a.M + 255) / 256;
        order_kernel<<<(unsigned)blocks, 256, 0, st>>>(norms, t, E, a.M, info);
        CEVB_CHECK_LAUNCH("order_kernel");
    }
    const int grid = cevb_expm_grid(n, worklist ? (work_cap < a.M ? work_cap : a.M) : a.M);
    const bool fast = !a.full_pivot && a.npad <= 104 && fast_smem_bytes(n) <= g_smem_optin;
    if (fast) {
        a.ldw = fast_ldw(a.npad);
        const size_t smem = fast_smem_bytes(n);
        CEVB_CHECK_CUDA(cudaFuncSetAttribute(expm_branch_fast_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        expm_branch_fast_kernel<<<grid, NT, smem, st>>>(a);
        CEVB_CHECK_LAUNCH("expm_branch_fast_kernel");
        return 0;
    }
    if (!a.use_smem && scratch == nullptr) {
        set_error("cevb_expm_branches: n=%d needs a global scratch buffer", n);
        return -1;
    }
    const size_t smem = branch_smem_bytes(n, a.use_smem != 0);
    if (smem > g_smem_optin) {
        set_error("cevb_expm_branches: staging for n=%d needs %zu B of shared memory", n, smem);
        return -1;
    }
    CEVB_CHECK_CUDA(cudaFuncSetAttribute(expm_branch_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    expm_branch_kernel<<<grid, NT, smem, st>>>(a);
    CEVB_CHECK_LAUNCH("expm_branch_kernel");
    return 0;
}

extern "C" int cevb_expm_branches(const double* pw, const double* norms, const uint16_t* csc,
                                  const uint8_t* csc_cnt, const double* t, int B, int E, int n,
                                  double* P, int32_t* info, void* scratch, unsigned int* counter,
                                  int full_pivot, void* stream) {
    return expm_branches_impl(pw, norms, csc, csc_cnt, t, B, E, n, P, info, scratch, counter,
                              full_pivot, nullptr, nullptr, 0, stream);
}

extern "C" int cevb_expm_branches_list(const double* pw, const double* norms, const uint16_t* csc,
                                       const uint8_t* csc_cnt, const double* t, int B, int E, int n,
                                       double* P, int32_t* info, void* scratch,
                                       unsigned int* counter, int full_pivot,
                                       const int32_t* worklist, const unsigned int* work_count,
                                       long long work_cap, void* stream) {
    if (worklist == nullptr || work_count == nullptr || work_cap <= 0) return 0;
    return expm_branches_impl(pw, norms, csc, csc_cnt, t, B, E, n, P, info, scratch, counter,
                              full_pivot, worklist, work_count, work_cap, stream);
}

// Diagnostics: when set to a device buffer of 10 uint64, expm_branch_kernel accumulates SM clocks
// spent per phase (0 order, 1 W/V build, 2 table staging, 3 U, 4 form, 5 GJ staging, 6 GJ update,
// 7 squarings, 8 post-process + store).  nullptr switches it off.
extern "C" int cevb_diag_set_phase_buffer(unsigned long long* buf) {
    g_phase_clk = buf;
    return 0;
}
