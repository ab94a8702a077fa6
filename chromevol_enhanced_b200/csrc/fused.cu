// Fused transition-probability + pruning path: P(t) is never written to HBM.
//
// Reference behaviour reproduced (per branch, per parameter vector):
//     P  = scipy.linalg.expm(Q t)                       AR:151
//     P  = max(P, 1e-12);  P /= rowsum(P);  NaN -> I    AR:153-159   (t == 0 -> I, AR:146-148)
//     w  = P @ L_child                                  AR:237
// followed, per parent node, by the product over children in file order with the reference's
// conditional rescale (AR:238-243) and the root log-likelihood (AR:251-272).
//
// Every branch of one parameter vector shares Q (the CLI never sets branch_params), so with the
// uniformisation shift mu = max(0, -min_i Q_ii)
//     expm(Q t) = e^{-mu t} sum_k t^k C_k,   C_k = (Q + mu I)^k / k!,
// is a polynomial in the SCALAR t with matrix coefficients that are formed once per vector
// (taylor_coeff_kernel; Q has <= 9 non-zeros per column, so each C_k costs O(9 n^2)) and, for a
// generator, without cancellation.  Evaluating it for many branches at once is the dense
// contraction
//     P[e, (i,j)] = sum_k T[e,k] C[k,(i,j)],   T[e,k] = e^{-mu t_e} t_e^k,
// which taylor_apply_kernel runs on the FP64 tensor cores (DMMA.8x8x4: M = 8 branches, N = 8
// columns j of one row i, K = 4 series terms) with the coefficient rows staged in shared memory
// by bulk async copies (cp.async.bulk + mbarrier ring).  The floor, the row sum and the dot
// product with L_child are applied to the accumulator fragments in registers, so only w (n
// doubles per branch) leaves the SM.  Branches that end in an observed leaf need one column of
// P and, for the row sum, only the floor correction of the rest: leaf_lowp_kernel (FP64 column +
// TF32 tensor cores).  The number of terms is chosen per (vector, branch) from x = nu t,
// nu = min(||Q + mu I||_1, ||Q + mu I||_inf), so that the truncation bound
// x^(K+1)/(K+1)! / (1 - x/(K+2)) e^{-mu t} is below 1e-18.  Pairs that would need more than the 64
// tabulated terms (x above about 15.6) are scaled and squared and therefore materialised: the
// series at t 2^-s by the same kernel (MODE 2), the squarings by square_post_kernel, and the
// combine kernel streams their P.
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

// Timing experiments only (tools/stub_build.py): -DCEVB_STUB=<mask> removes parts of the apply
// kernel -- 1 epilogue arithmetic, 2 ring waits and copies, 4 the DMMAs of chunks >= 1, 8 the
// per-row reduction and stores -- and of the coefficient kernel -- 16 the TF32 copy, 32 the
// activity scan / slot bookkeeping / write-out of a chunk, 64 the sparse product of a step.
// Results are wrong with any bit set; the product build has 0.
#ifndef CEVB_STUB
#define CEVB_STUB 0
#endif

namespace cevb {

constexpr int F_KC = CEVB_TAYLOR_KC;       // chunks of 4 series coefficients (degrees 0..63)
static_assert(F_KC <= 16 && F_KC % 2 == 0, "slot counts are 16 nibbles of one 64-bit word; TF32 k-blocks pair chunks");
constexpr int F_JT = 13;                   // 8-column tiles per pass (104 columns)
constexpr double F_TRUNC = CEVB_TAYLOR_TRUNC;  // truncation bound of the series (absolute, entries <= 1)

// Layout of the coefficient table of one vector.  An ITEM is (row i, pass p): the up-to-13
// 8-column tiles [13 p, 13 p + 13) of row i.  tc[item][chunk c][slot s][column 8][degree 4]:
// 13 * 32 doubles per chunk, F_KC chunks per item, contiguous.  The tiles of an item are stored in
// SLOT order: slots are handed out in the order in which tiles first hold a non-zero coefficient
// (C_k of a sparse Q fills in slowly: a third of all (row, chunk, tile) triples of the N = 100
// workload are structurally zero), so the tiles that matter at chunk c are exactly the slot prefix
// [0, m(c)).  The kernels neither copy nor multiply the rest.  Per item four 64-bit words of
// meta data follow the coefficient data of the vector:
//   [0] perm:   nibble s = tile (within the pass) held by slot s
//   [1] counts: nibble c = m(c), the active slot prefix of chunk c (0..13)
//   [2] inv:    nibble t = slot of tile t
//   [3] unused
constexpr int F_CHUNK = F_JT * 32;         // doubles per (item, chunk)
constexpr int F_ITEM = F_KC * F_CHUNK;     // doubles per item
constexpr int F_META = 4;                  // 64-bit words per item
__host__ __device__ inline int taylor_npass(int n) { return ((n + 7) / 8 + F_JT - 1) / F_JT; }
__host__ __device__ inline size_t taylor_data_doubles(int n) {
    return (size_t)n * taylor_npass(n) * F_ITEM;
}
// Low-precision copy of the coefficients for the leaf launch (n <= 104, generator Q only): the
// floor correction sum_j max(1e-12 - P_ij, 0) of a leaf branch is at most 1e-10 and is needed to
// about 1 % (see leaf_lowp_kernel), so it is formed on the TF32 tensor cores from
//     P_ij = sum_k That[e][k] Chat_k[i][j],  That = e^{-mu t} (nu t)^k / k!,  Chat_k = ((Q + mu I) / nu)^k,
// both factors in [0, 1].  tf[row][k-block of 8 degrees][tile 13][lane 32][2]: the B fragment of
// mma.m16n8k8.tf32 (lane 4g+t holds degrees t and t+4 of column g of the tile).
constexpr int L_KB = F_KC / 2;                  // k-blocks of 8 degrees
constexpr int L_TILE = 64;                      // floats per (k-block, tile)
constexpr int L_ROW = L_KB * F_JT * L_TILE;     // floats per row
__host__ __device__ inline size_t taylor_lowp_doubles(int n) {
    return taylor_npass(n) == 1 ? (size_t)n * L_ROW / 2 : 0;
}
__host__ __device__ inline size_t taylor_meta_end_doubles(int n) {
    return taylor_data_doubles(n) + (size_t)n * taylor_npass(n) * F_META;
}
__host__ __device__ inline size_t taylor_vec_doubles(int n) {
    return taylor_meta_end_doubles(n) + taylor_lowp_doubles(n);
}

// The series is summed for the SHIFTED matrix (uniformisation): with mu = max(0, -min_i Q_ii),
//     expm(Q t) = e^{-mu t} sum_k t^k C_k,   C_k = (Q + mu I)^k / k!.
// For a generator Q + mu I is entrywise non-negative, so the sum has no cancellation and its
// norm nu = min(||Q + mu I||_1, ||Q + mu I||_inf) is about half of ||Q||_1.
// Number of series coefficients (degree + 1) so that the truncation bound
//     x^(K+1)/(K+1)! / (1 - x/(K+2)) * e^{-mu t} <= 1e-18 (CEVB_TAYLOR_TRUNC),   x = nu t;
// a result above 4 * F_KC means the coefficient table is too short (scale and square).
__host__ __device__ inline int taylor_ncoef(double x, double mut) {
    if (!(x > 0.0)) return 1;
    const double sc = exp(-mut);
    double term = x;  // x^k / k!
    for (int k = 1; k < 4 * F_KC; ++k) {
        term *= x / (double)(k + 1);  // x^(k+1) / (k+1)!
        if ((double)(k + 2) > x && term / (1.0 - x / (double)(k + 2)) * sc <= F_TRUNC) return k + 1;
    }
    return 4 * F_KC + 1;
}

// what to do with one (vector, branch length) pair
struct PairPlan {
    int nc;      // series coefficients (0: nothing to sum)
    int s;       // squarings (> 0 or non-finite rates: the matrix has to be materialised)
    int kind;    // 0 fused series, 1 identity (t == 0), 2 materialised
    double tau;  // t 2^-s
    double a0;   // e^{-mu tau}
};

__host__ __device__ inline PairPlan plan_pair(double t, double nu, double mu) {
    PairPlan p = {0, 0, 0, 0.0, 1.0};
    if (!(t > 0.0)) {  // t == 0 -> identity (AR:146-148); t < 0 / NaN never reach here
        p.kind = 1;
        return p;
    }
    const double x = t * nu, mut = t * mu;
    if (!(x < INFINITY) || !(mut < INFINITY)) {
        // NaN / inf rates: the series is evaluated at NaN and the post-processing falls back to
        // the identity like the reference does (AR:156-166)
        p.kind = 2;
        p.nc = 4;
        p.tau = NAN;
        p.a0 = NAN;
        return p;
    }
    int s = 0;
    int nc = taylor_ncoef(x, mut);
    while (nc > 4 * F_KC && s < 1000) {
        ++s;
        nc = taylor_ncoef(ldexp(x, -s), ldexp(mut, -s));
    }
    p.nc = nc;
    p.s = s;
    p.kind = s > 0 ? 2 : 0;
    p.tau = ldexp(t, -s);
    p.a0 = exp(-ldexp(mut, -s));
    return p;
}

// ---------------------------------------------------------------------------------------------
// ||Q||_1, the shift mu = max(0, -min_i Q_ii), nu = min(||Q + mu I||_1, ||Q + mu I||_inf) and the
// compressed column structure of Q (row indices of the non-zeros of every column); one CTA per
// parameter vector, one thread per column / row.
__global__ void __launch_bounds__(128) q_structure_kernel(const double* __restrict__ pw, int n,
                                                          int ldq, double* __restrict__ norms,
                                                          uint16_t* __restrict__ csc,
                                                          uint8_t* __restrict__ csc_cnt, double t_max) {
    __shared__ double red[4];
    const int b = blockIdx.x, tid = threadIdx.x;
    const double* Q = pw + (size_t)b * CEVB_NPOW * n * ldq;
    uint16_t* idx_c = csc + (size_t)b * 2 * CEVB_MAX_SPARSE * n;
    uint8_t* cnt_c = csc_cnt + (size_t)b * 2 * n;
    auto block_max = [&](double v) {
        v = warp_max(v);
        __syncthreads();
        if ((tid & 31) == 0) red[tid >> 5] = v;
        __syncthreads();
        return fmax(fmax(red[0], red[1]), fmax(red[2], red[3]));
    };
    // shift
    double dmax = 0.0;
    bool nan = false;
    for (int j = tid; j < n; j += 128) {
        const double d = Q[(size_t)j * ldq + j];
        if (d != d) nan = true;
        dmax = fmax(dmax, -d);
    }
    const double mu = block_max(dmax);
    // column pass: structure, ||Q||_1, ||Q + mu I||_1
    double best = 0.0, best_s = 0.0;
    bool negoff = false;   // a negative (or non-finite) off-diagonal entry: Q is no generator
    for (int j = tid; j < n; j += 128) {
        int cc = 0;
        double sum = 0.0, sum_s = 0.0;
        for (int i = 0; i < n; ++i) {
            const double q = Q[(size_t)i * ldq + j];
            if (i != j && !(q >= 0.0 && q < INFINITY)) negoff = true;
            if (q != 0.0) {  // NaN != 0 is true: kept as "non-zero"
                if (cc < CEVB_MAX_SPARSE) idx_c[(size_t)cc * n + j] = (uint16_t)i;
                ++cc;
            }
            sum += fabs(q);
            sum_s += fabs(i == j ? q + mu : q);
        }
        cnt_c[j] = (cc > CEVB_MAX_SPARSE) ? 255 : (uint8_t)cc;
        if (sum != sum) nan = true;
        best = fmax(best, sum);
        best_s = fmax(best_s, sum_s);
    }
    // row pass: ||Q + mu I||_inf
    double best_r = 0.0;
    for (int i = tid; i < n; i += 128) {
        double sum_s = 0.0;
        for (int j = 0; j < n; ++j) {
            const double q = Q[(size_t)i * ldq + j];
            sum_s += fabs(i == j ? q + mu : q);
        }
        best_r = fmax(best_r, sum_s);
    }
    const double one = block_max(best);
    const double nu1 = block_max(best_s);
    const double nui = block_max(best_r);
    const bool anynan = __syncthreads_or(nan);
    const bool anyneg = __syncthreads_or(negoff);
    if (tid == 0) {
        double* out = norms + (size_t)b * CEVB_NNORM;
        const double nu = fmin(nu1, nui);
        out[CEVB_NORM_ONE] = anynan ? NAN : one;   // numpy max propagates NaN
        out[CEVB_NORM_SHIFT] = anynan ? NAN : mu;
        out[CEVB_NORM_SHIFTED] = anynan ? NAN : nu;
        out[CEVB_NORM_GENERATOR] = (!anynan && !anyneg && nu > 0.0 && nu < INFINITY) ? 1.0 : 0.0;
        // chunks any branch of this vector can need: the longest branch decides (all of them when
        // it has to be scaled and squared, or when the caller gives no bound).  Decided once per
        // vector here -- the series-length rule is a 64-step scalar loop with a division per step
        int kc_need = F_KC;
        if (t_max > 0.0) {
            const PairPlan pp = plan_pair(t_max, anynan ? NAN : nu, anynan ? NAN : mu);
            if (pp.kind == 0) kc_need = (pp.nc + 3) >> 2;
        }
        out[CEVB_NORM_CHUNKS] = (double)kc_need;
    }
}

// ---------------------------------------------------------------------------------------------
// C_0 = I, C_k = C_{k-1} (Q + mu I) / k, TCR rows per CTA; row i of C_k only needs row i of
// C_{k-1}.  Output: the slot-ordered item layout described at the top of this file (inside a
// tile: [column 8][degree 4], exactly the B-operand fragment order of DMMA, lane 4g+t <- column
// g, degree t).  `dense` != 0 hands every tile its slot at chunk 0 in natural order (identity
// permutation, m(c) = all tiles): the layout without sparsity, kept for A/B measurements.
constexpr int TCR = 4;                                   // rows of C_k per CTA
// 1 / k for k = 1 .. 4 F_KC as the IEEE quotient 1.0 / k (filled by the host on first use): the
// recurrence divides by k once per step and a double division costs ~20 instructions per thread
__constant__ double c_inv_int[4 * CEVB_TAYLOR_KC + 1];
constexpr int TC_MAXPASS = ((CEVB_MAX_N + 7) / 8 + F_JT - 1) / F_JT;

__global__ void __launch_bounds__(128) taylor_coeff_kernel(const double* __restrict__ pw,
                                                           const uint16_t* __restrict__ csc,
                                                           const uint8_t* __restrict__ csc_cnt,
                                                           const double* __restrict__ norms,
                                                           int n, int ldq, int npad, double t_max,
                                                           int dense, int lowp_on,
                                                           double* __restrict__ tc) {
    extern __shared__ double shc[];
    double* cur = shc;                    // [TCR][npad] (row-major per row: the gather of a step reads
    double* nxt = shc + TCR * npad;       // neighbouring columns, i.e. conflict-free 8-byte loads)
    // staging of one chunk, DEGREE-major: [TCR][4 degrees][sst] with a row stride of npad + 4
    // doubles, so that the per-degree store of a step (lane = column) and the TF32 read-out are
    // conflict-free and the two degree rows a write-out lane pairs up sit 8 banks apart (the
    // column-major [npad][4] layout cost 8 wavefronts per warp store: 126 M bank conflicts and
    // 67 % LSU utilisation per 512-vector launch in the round-2 ncu capture)
    double* stg = shc + 2 * TCR * npad;
    const int sst = npad + 4;
    uint32_t* tfs = reinterpret_cast<uint32_t*>(stg + 4 * TCR * sst);   // [TCR][4][104] parked TF32 words
    __shared__ uint8_t s_perm[TCR][TC_MAXPASS][16];       // slot -> tile of the pass
    __shared__ int8_t s_slot[TCR][TC_MAXPASS][16];        // tile of the pass -> slot (-1: none yet)
    __shared__ int s_nslot[TCR][TC_MAXPASS];
    __shared__ unsigned long long s_counts[TCR][TC_MAXPASS];
    const int i0 = blockIdx.x * TCR, b = blockIdx.y, tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int nj = npad >> 3, npass = taylor_npass(n);
    const double* Q = pw + (size_t)b * CEVB_NPOW * n * ldq;
    const uint16_t* idx_c = csc + (size_t)b * 2 * CEVB_MAX_SPARSE * n;
    const uint8_t* cnt_c = csc_cnt + (size_t)b * 2 * n;
    double* tcb = tc + (size_t)b * taylor_vec_doubles(n);
    const double mu = norms[(size_t)b * CEVB_NNORM + CEVB_NORM_SHIFT];
    // low-precision copy (see taylor_lowp_doubles): generator vectors with one pass per row
    const double nu_l = norms[(size_t)b * CEVB_NNORM + CEVB_NORM_SHIFTED];
    const bool lowp = lowp_on && npass == 1 && norms[(size_t)b * CEVB_NNORM + CEVB_NORM_GENERATOR] != 0.0;
    float* tfb = reinterpret_cast<float*>(tcb + taylor_meta_end_doubles(n));
    const double inv_nu = 1.0 / nu_l;
    double scl = 1.0;          // k! / nu^k of the current degree
    double sc4[4] = {1.0, 1.0, 1.0, 1.0};
    // chunks any branch of this vector can need (q_structure_kernel decides from the longest branch)
    const int kc_need = (int)norms[(size_t)b * CEVB_NNORM + CEVB_NORM_CHUNKS];
    const int k_end = 4 * kc_need;
    // The sparse column of (Q + mu I) a thread multiplies with is the same for every row and
    // all steps: it is kept in registers (<= 12 entries), so a step is shared-memory loads
    // and FMAs only, for TCR rows at once.  Columns beyond 128 (n > 128) and columns with more
    // than 12 non-zeros re-read the structure from global memory.
    const int j0 = tid;
    int cnt0 = 0;
    int ridx[CEVB_MAX_SPARSE];
    double qval[CEVB_MAX_SPARSE];
#pragma unroll
    for (int q = 0; q < CEVB_MAX_SPARSE; ++q) {
        ridx[q] = 0;
        qval[q] = 0.0;
    }
    if (j0 < n) {
        cnt0 = cnt_c[j0];
        if (cnt0 != 255) {
#pragma unroll
            for (int q = 0; q < CEVB_MAX_SPARSE; ++q) {
                if (q < cnt0) {
                    ridx[q] = idx_c[(size_t)q * n + j0];
                    qval[q] = Q[(size_t)ridx[q] * ldq + j0];
                }
            }
        }
    }
    for (int j = tid; j < npad; j += 128) {
#pragma unroll
        for (int r = 0; r < TCR; ++r) cur[r * npad + j] = (j == i0 + r) ? 1.0 : 0.0;
    }
    if (tid < TCR * TC_MAXPASS) {
        const int r = tid / TC_MAXPASS, p = tid - r * TC_MAXPASS;
        s_nslot[r][p] = 0;
        s_counts[r][p] = 0ull;
        for (int q = 0; q < 16; ++q) {
            s_slot[r][p][q] = -1;
            s_perm[r][p][q] = 0;
        }
    }
    // the longest sparse column among this warp's threads bounds the unrolled product below
    // (typically 6-8 of the 12 slots are in use)
    const int cmax = __reduce_max_sync(0xffffffffu, (j0 < n && cnt0 != 255) ? cnt0 : 0);
    __syncthreads();
    for (int k = 0; k < k_end; ++k) {
        sc4[k & 3] = scl;
        scl *= (double)(k + 1) * inv_nu;       // (TF32 copy only: the rounding of k!/nu^k is immaterial)
        for (int j = tid; j < npad; j += 128) {
#pragma unroll
            for (int r = 0; r < TCR; ++r)
                stg[(size_t)(r * 4 + (k & 3)) * sst + j] = (j < n) ? cur[r * npad + j] : 0.0;
        }
        if (k + 1 < k_end) {
            const double inv = c_inv_int[k + 1];       // 1 / (k + 1), correctly rounded (host table)
            if (j0 < n) {
                double acc[TCR];
#pragma unroll
                for (int r = 0; r < TCR; ++r) acc[r] = cur[r * npad + j0] * mu;
                if (cnt0 == 255) {
                    for (int rr = 0; rr < n; ++rr) {
                        const double q = Q[(size_t)rr * ldq + j0];
#pragma unroll
                        for (int r = 0; r < TCR; ++r) acc[r] = fma(cur[r * npad + rr], q, acc[r]);
                    }
                } else if (!(CEVB_STUB & 64)) {
#pragma unroll
                    for (int q = 0; q < CEVB_MAX_SPARSE; ++q) {
                        if (q < cmax) {   // warp-uniform
#pragma unroll
                            for (int r = 0; r < TCR; ++r)
                                acc[r] = fma(cur[r * npad + ridx[q]], qval[q], acc[r]);
                        }
                    }
                }
#pragma unroll
                for (int r = 0; r < TCR; ++r) nxt[r * npad + j0] = acc[r] * inv;
            }
            for (int j = tid + 128; j < n; j += 128) {   // n > 128
                const int cnt = cnt_c[j];
                double acc[TCR];
#pragma unroll
                for (int r = 0; r < TCR; ++r) acc[r] = cur[r * npad + j] * mu;
                const int lim = (cnt == 255) ? n : cnt;
                for (int q = 0; q < lim; ++q) {
                    const int rr = (cnt == 255) ? q : idx_c[(size_t)q * n + j];
                    const double qv = Q[(size_t)rr * ldq + j];
#pragma unroll
                    for (int r = 0; r < TCR; ++r) acc[r] = fma(cur[r * npad + rr], qv, acc[r]);
                }
#pragma unroll
                for (int r = 0; r < TCR; ++r) nxt[r * npad + j] = acc[r] * inv;
            }
        }
        __syncthreads();
        if ((k & 3) == 3) {
            const int c = k >> 2;
            if (!(CEVB_STUB & 16) && lowp && tid < F_JT * 8) {
                // scaled TF32 copy in mma.m16n8k8 B-fragment order: lane 4 g + d of a tile holds
                // degrees d (even chunk of the k-block) and d + 4 (odd chunk) of column g, so a
                // thread's four lanes are 32 contiguous bytes.  The even chunk is parked in shared
                // memory and the pair goes out as two 16-byte stores with the odd chunk (or at
                // once, with zeros for the other half, if the table ends with the even chunk):
                // full sectors instead of sixteen scattered 4-byte stores per row and chunk.
                const int tile = tid >> 3, g8 = tid & 7, kb = c >> 1, half = c & 1;
                const bool tail = (c == kc_need - 1) && half == 0;
                for (int r = 0; r < TCR; ++r) {
                    if (i0 + r >= n) break;
                    uint32_t u[4];
#pragma unroll
                    for (int d = 0; d < 4; ++d) {
                        // all 13 tiles are written: columns beyond npad are zeros
                        const float v = tid < npad ? (float)(stg[(size_t)(r * 4 + d) * sst + tid] * sc4[d]) : 0.0f;
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u[d]) : "f"(v));
                    }
                    uint32_t* park = tfs + (size_t)(r * 4) * (F_JT * 8) + tid;
                    if (half == 0 && !tail) {
#pragma unroll
                        for (int d = 0; d < 4; ++d) park[d * (F_JT * 8)] = u[d];
                    } else {
                        uint32_t e[4] = {0u, 0u, 0u, 0u};     // the even chunk's words (zeros: none)
                        if (half == 1) {
#pragma unroll
                            for (int d = 0; d < 4; ++d) e[d] = park[d * (F_JT * 8)];
                        }
                        uint4* dst = reinterpret_cast<uint4*>(
                            tfb + (((size_t)(i0 + r) * L_KB + kb) * F_JT + tile) * L_TILE + g8 * 8);
                        if (half == 1) {
                            dst[0] = make_uint4(e[0], u[0], e[1], u[1]);
                            dst[1] = make_uint4(e[2], u[2], e[3], u[3]);
                        } else {                              // tail: this (even) chunk + zeros
                            dst[0] = make_uint4(u[0], 0u, u[1], 0u);
                            dst[1] = make_uint4(u[2], 0u, u[3], 0u);
                        }
                    }
                }
            }
            if (CEVB_STUB & 32) {
                double* sw2 = cur;
                cur = nxt;
                nxt = sw2;
                continue;
            }
            // Warp r takes row r: which tiles of this chunk hold a non-zero (NaN counts as
            // non-zero), then the newly active tiles take the next slots of their item -- in tile
            // order, by a ballot instead of a serial loop -- and the chunk's slot count is noted.
            static_assert(TCR == 4, "one warp of the 128-thread CTA per row");
            {
                const int r = warp;
                unsigned long long actmask = 0ull;              // bit jt: tile jt of row r is active
                for (int jt = 0; jt < nj; ++jt) {
                    const double v = stg[(size_t)(r * 4 + (lane >> 3)) * sst + jt * 8 + (lane & 7)];
                    if (__any_sync(0xffffffffu, !(v == 0.0))) actmask |= 1ull << jt;
                }
                for (int p = 0; p < npass; ++p) {
                    int ns = s_nslot[r][p];
                    const int jt = p * F_JT + lane;
                    const bool take = lane < F_JT && jt < nj && s_slot[r][p][lane] < 0 &&
                                      (dense || ((actmask >> jt) & 1ull));
                    const unsigned mask = __ballot_sync(0xffffffffu, take);
                    if (take) {
                        const int sl = ns + __popc(mask & ((1u << lane) - 1u));
                        s_slot[r][p][lane] = (int8_t)sl;
                        s_perm[r][p][sl] = (uint8_t)lane;
                    }
                    ns += __popc(mask);
                    if (lane == 0) {
                        s_nslot[r][p] = ns;
                        // chunk 0 is multiplied by straight-line code for 0, 3, 6, 9 or 13 slots
                        // (the apply kernel interleaves it with an epilogue): its count is rounded
                        // up; the extra slots hold zeros
                        int cnt = ns;
                        if (c == 0 && ns > 0) cnt = ns <= 3 ? 3 : (ns <= 6 ? 6 : (ns <= 9 ? 9 : F_JT));
                        s_counts[r][p] |= (unsigned long long)cnt << (4 * c);
                    }
                }
            }
            __syncthreads();
            // write-out in slot order.  Only the slot prefix the kernels copy is written: the
            // active slots and, for chunk 0, the zero slots its count is rounded up to.
            for (int r = 0; r < TCR; ++r) {
                if (i0 + r >= n) break;
                for (int p = 0; p < npass; ++p) {
                    const int ns = s_nslot[r][p];
                    const int nw = (int)((s_counts[r][p] >> (4 * c)) & 15ull);
                    double2* dst = reinterpret_cast<double2*>(
                        tcb + ((size_t)((i0 + r) * npass + p) * F_KC + c) * F_CHUNK);
                    for (int x = tid; x < nw * 16; x += 128) {
                        const int sl = x >> 4, rem = x & 15;   // rem = column * 2 + half
                        double2 v = make_double2(0.0, 0.0);
                        if (sl < ns) {
                            const int jt = p * F_JT + s_perm[r][p][sl];
                            // rem = column * 2 + half: degrees 2 half and 2 half + 1 of the column
                            const double* sp = stg + (size_t)(r * 4 + 2 * (rem & 1)) * sst + jt * 8 + (rem >> 1);
                            v = make_double2(sp[0], sp[sst]);
                        }
                        dst[x] = v;
                    }
                }
            }
            __syncthreads();
        }
        double* sw = cur;
        cur = nxt;
        nxt = sw;
    }
    // meta data: tiles that never became active (and the tiles beyond npad) take the last slots
    if (tid < TCR * npass) {
        const int r = tid / npass, p = tid - r * npass;
        if (i0 + r < n) {
            int ns = s_nslot[r][p];
            for (int tl = 0; tl < F_JT; ++tl) {
                if (s_slot[r][p][tl] < 0) {
                    s_slot[r][p][tl] = (int8_t)ns;
                    s_perm[r][p][ns] = (uint8_t)tl;
                    ++ns;
                }
            }
            unsigned long long perm = 0ull, inv = 0ull;
            for (int q = 0; q < F_JT; ++q) {
                perm |= (unsigned long long)s_perm[r][p][q] << (4 * q);
                inv |= (unsigned long long)(uint8_t)s_slot[r][p][q] << (4 * q);
            }
            unsigned long long* meta = reinterpret_cast<unsigned long long*>(
                                           tcb + taylor_data_doubles(n)) +
                                       (size_t)((i0 + r) * npass + p) * F_META;
            meta[0] = perm;
            meta[1] = s_counts[r][p];
            meta[2] = inv;
            meta[3] = 0ull;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Which (vector, distinct branch length) pairs need a materialised matrix (eta > theta), as a
// compact work list for the Pade kernel; slot[b*E+e] = position in the list (-1 = fused path).
__global__ void __launch_bounds__(256) expm_plan_kernel(const double* __restrict__ norms,
                                                        const double* __restrict__ t, int E,
                                                        long long M, long long cap,
                                                        int32_t* __restrict__ slot,
                                                        int32_t* __restrict__ worklist,
                                                        int32_t* __restrict__ work_s,
                                                        unsigned int* __restrict__ count,
                                                        int32_t* __restrict__ plan_nc,
                                                        double* __restrict__ plan_tau,
                                                        double* __restrict__ plan_a0) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M) return;
    const int b = (int)(idx / E), e = (int)(idx - (long long)b * E);
    const double* nr = norms + (size_t)b * CEVB_NNORM;
    const PairPlan pp = plan_pair(t[e], nr[CEVB_NORM_SHIFTED], nr[CEVB_NORM_SHIFT]);
    if (plan_nc != nullptr) {
        // the per-pair decision is scalar FP64 work with long dependency chains (exp, divisions):
        // done here once, one thread per pair, instead of serially at the start of every CTA
        plan_nc[idx] = pp.nc | (pp.kind << 8);
        plan_tau[idx] = pp.tau;
        plan_a0[idx] = pp.a0;
    }
    if (slot == nullptr) return;
    int s = -1;
    if (pp.kind == 2) {
        const unsigned int q = atomicAdd(count, 1u);
        if ((long long)q < cap) {
            worklist[q] = (int32_t)idx;
            if (work_s != nullptr) work_s[q] = pp.s;
            s = (int)q;
        } else {
            s = -2;  // overflow: the host sees count > cap and re-runs with smaller chunks
        }
    }
    slot[idx] = s;
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
// try_wait with a suspend-time hint: a waiting warp sleeps in hardware until the phase completes
// (or the hint expires) instead of spinning.  The kernels are issue-bound -- a DMMA.8x8x4 holds
// its scheduler for 16 cycles and every other warp instruction costs one -- so polling loops
// of waiting warps are paid for by the warps that have work.
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        " selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
        : "memory");
    return ok != 0;
}
// A warp that finds its barrier not ready sleeps between polls (64 ns, then 256 ns): the warps
// that wait are the ones AHEAD of the ring, so waking up late costs nothing, while every poll
// takes an issue slot from a warp that has work.
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    __nanosleep(64);
    while (!mbar_try_wait(bar, parity)) __nanosleep(256);
}

struct ApplyArgs {
    const double* tc;
    const double* norms;
    const int32_t* items;     // child node of every branch of this launch
    const double* t_node;     // [n_nodes] effective branch length above a node
    const int32_t* leaf_code; // [n_nodes]
    const double* node_vec;   // [B][n_nodes][ldp]
    double* w;                // [B][n_nodes][ldp]  numerators  sum_j P'[i][j] L[j]
    double* wr;               // [B][n_nodes][ldp]  row sums of the floored P (AR:154 divides)
    uint8_t* mode;            // [B][n_nodes]: only non-zero values are ever written
    int n_items, n, npad, ldp, ldl, n_nodes, rows_per_cta;
    int npass;                // passes of 13 tiles per row (1 for n <= 104)
    int meta_bytes;           // shared memory for the per-item meta data of a CTA's rows
    int stage_doubles_cap;    // doubles available for the coefficient ring
    int skew;                 // experiment: clocks by which every other warp of a scheduler starts late
    int skip_generators;      // leaf launch: generator vectors are handled by leaf_lowp_kernel
    // EVAL mode: items index the branch-length array t_node [E]; the raw series value at
    // t 2^-s goes to X[slot[b * E + item]] (n rows of ldp doubles) for the squaring kernel
    const int32_t* slot;
    double* X;
    int E;
    // per-pair plan [B][E] from expm_plan_kernel: series length | kind << 8, (scaled) time and
    // e^{-mu tau}; branch_of maps a child node to its row of the plan (NULL: the node itself)
    const int32_t* plan_nc;
    const double* plan_tau;
    const double* plan_a0;
    const int32_t* branch_of;
};

constexpr int APPLY_GENERAL = 0, APPLY_LEAF = 1, APPLY_EVAL = 2;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

constexpr int F_MAXSTAGE = 4;

__host__ __device__ inline size_t apply_header_bytes(int eb) {
    // barriers, child / code / terms (int), tau / e^{-mu tau} (double) per branch
    return (size_t)((16 * F_MAXSTAGE + 4 * (3 * eb + 4) + 16 * eb + 127) / 128) * 128;
}
// shared memory of the apply kernel: header (barriers, per-branch ints), the meta data of the
// CTA's items (16 bytes each), the t^k table A [eb/8][F_KC][32], the child vectors L [eb][ldl]
// and the ring of coefficient item tiles
__host__ __device__ inline size_t apply_meta_bytes(int items) {
    return (size_t)((16 * items + 127) / 128) * 128;
}
__host__ __device__ inline size_t apply_fixed_bytes(int eb, int ldl, int items) {  // ldl = 0: no child vectors
    return apply_header_bytes(eb) + apply_meta_bytes(items) +
           sizeof(double) * ((size_t)(eb / 8) * F_KC * 32 + (size_t)eb * ldl);
}

// nibble q (compile-time) of the 64-bit word (lo, hi)
template <int Q>
__device__ __forceinline__ int nib(uint32_t lo, uint32_t hi) {
    return (int)(((Q < 8 ? lo : hi) >> (4 * (Q & 7))) & 15u);
}

// DMMA switched off by a (warp-uniform) predicate: @!p leaves the accumulator untouched
__device__ __forceinline__ void dmma884_if(double& c0, double& c1, double a, double b, int on) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.s32 p, %4, 0;\n"
        " @p mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n}\n"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b), "r"(on));
}

// One chunk of an item for the M active slots: the M B-fragment loads are issued back to back
// and only then the M DMMAs (an LDS placed right before its DMMA, or right after a DMMA that
// still has to read the same register, serialises the warp on the shared-memory latency -- see
// tools/dmma_probe.py; the compiler barriers keep the two groups apart).
template <int M, bool DIRECT>
__device__ __forceinline__ void chunk_slots(double (&acc)[F_JT][2], const double* Bc, const double av) {
    double bv[M > 0 ? M : 1];
#pragma unroll
    for (int sl = 0; sl < M; ++sl) bv[sl] = DIRECT ? __ldg(Bc + sl * 32) : Bc[sl * 32];
#pragma unroll
    for (int sl = 0; sl < M; ++sl) dmma884(acc[sl][0], acc[sl][1], av, bv[sl]);
}
// the same for a run-time slot count m (warp-uniform): a compare tree over the 13 straight-line
// variants -- four compare-and-branch pairs per chunk, no per-slot predicates
template <bool DIRECT>
__device__ __forceinline__ void chunk_dispatch(const int m, double (&acc)[F_JT][2], const double* Bc,
                                               const double av) {
    if (m == F_JT) {   // the most frequent count from the third chunk on
        chunk_slots<13, DIRECT>(acc, Bc, av);
        return;
    }
    if (m > 6) {
        if (m > 9) {
            if (m > 11) {
                chunk_slots<12, DIRECT>(acc, Bc, av);
            } else {
                if (m > 10) chunk_slots<11, DIRECT>(acc, Bc, av);
                else chunk_slots<10, DIRECT>(acc, Bc, av);
            }
        } else {
            if (m > 8) chunk_slots<9, DIRECT>(acc, Bc, av);
            else if (m > 7) chunk_slots<8, DIRECT>(acc, Bc, av);
            else chunk_slots<7, DIRECT>(acc, Bc, av);
        }
    } else {
        if (m > 3) {
            if (m > 5) chunk_slots<6, DIRECT>(acc, Bc, av);
            else if (m > 4) chunk_slots<5, DIRECT>(acc, Bc, av);
            else chunk_slots<4, DIRECT>(acc, Bc, av);
        } else {
            if (m > 2) chunk_slots<3, DIRECT>(acc, Bc, av);
            else if (m > 1) chunk_slots<2, DIRECT>(acc, Bc, av);
            else if (m > 0) chunk_slots<1, DIRECT>(acc, Bc, av);
        }
    }
}

// One CTA: vector b = blockIdx.z, branches [blockIdx.x * EB, +EB) of the launch's item list
// (EB = NW * G * 8), rows [blockIdx.y * rows_per_cta, +rows_per_cta) of every P.  NW consumer
// warps each own a group of 8 branches and keep the 13 accumulator tiles of the current item
// (row, pass) in registers; one extra producer warp streams the coefficient items (the same
// tile for all consumers) through a ring of shared-memory stages with full/empty mbarriers,
// so the consumers drift apart and one warp's epilogue overlaps another warp's DMMAs.
// Only the ACTIVE slot prefix of every (item, chunk) is copied and multiplied (see the layout
// notes at the top of the file): slot s of an item holds tile perm[s].
// MODE APPLY_LEAF: every child of the launch is an observed leaf (one-hot L): the dot product
// is a select of one accumulator and no child vectors are staged.  MODE APPLY_EVAL: no
// epilogue arithmetic at all -- the series value at the scaled time t 2^-s of every pair with
// t ||Q||_1 > theta is written out for the squaring kernel.
// ONEPASS: n <= 104, a row is a single item (compile-time trip count).
// DIRECT: the small launches of the upper tree levels (a few 8-branch groups per vector).  With
// one or three consumer warps per CTA and a 40 KB ring per CTA, an SM holds about one consumer
// warp per scheduler and every latency of that warp is exposed (measured: 2.5 us per row
// against 0.5 us of issue time).  The DIRECT variant has no ring, no producer and no mbarriers:
// a warp loads the B fragments of a chunk straight from the table in global memory (the slot
// layout is lane-contiguous, so a fragment load is one coalesced 256-byte request), which
// leaves room for 16 consumer warps per SM that hide each other's load latency.
template <int G, int NW, int MODE, bool ONEPASS, bool DIRECT = false>
__global__ void __launch_bounds__(32 * (NW + (DIRECT ? 0 : 1)),
                                  DIRECT ? 16 / NW : (NW >= 15 ? 1 : (NW == 7 ? 2 : 4)))
    taylor_apply_kernel(ApplyArgs a) {
    constexpr bool LEAF = MODE != APPLY_GENERAL;   // no child vectors in shared memory
    constexpr bool EVAL = MODE == APPLY_EVAL;
    extern __shared__ __align__(128) unsigned char smraw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    constexpr int EB = NW * G * 8;
    const int b = blockIdx.z;
    const int e0 = blockIdx.x * EB;
    const int n = a.n, ldl = a.ldl;
    const int npass = ONEPASS ? 1 : a.npass;
    const int i_begin = blockIdx.y * a.rows_per_cta;
    const int i_end = min(n, i_begin + a.rows_per_cta);

    uint64_t* full = reinterpret_cast<uint64_t*>(smraw);
    uint64_t* empty = full + F_MAXSTAGE;
    int* s_child = reinterpret_cast<int*>(smraw + 16 * F_MAXSTAGE);
    int* s_code = s_child + EB;
    int* s_nc = s_code + EB;
    double* s_tau = reinterpret_cast<double*>(s_nc + EB + (EB & 1) + 2);   // 8-byte aligned: EB is a multiple of 8
    double* s_a0 = s_tau + EB;
    uint4* s_meta = reinterpret_cast<uint4*>(smraw + apply_header_bytes(EB));   // [rows][npass]
    // [EB/8][ck][32]: only the ck chunks this CTA's longest branch needs; the child vectors and
    // the ring follow (the shared memory is sized for ck = F_KC, the ring takes what is left)
    double* As = reinterpret_cast<double*>(smraw + apply_header_bytes(EB) + a.meta_bytes);

    if (MODE == APPLY_LEAF && a.skip_generators &&
        a.norms[(size_t)b * CEVB_NNORM + CEVB_NORM_GENERATOR] != 0.0)
        return;
    if (EVAL) {
        // The launch covers every (vector, branch length) pair, but only the few pairs whose
        // series needs squarings have work here.  The list is ordered by decreasing length and
        // "needs squarings" is monotone in the length, so a CTA whose FIRST item does not need
        // them has nothing to do: it leaves before the prologue (one broadcast load).
        if (e0 >= a.n_items) return;
        const int pn0 = a.plan_nc[(size_t)b * a.E + a.items[e0]];
        if ((pn0 >> 8) != 2) return;
    }
    const double* tcb = a.tc + (size_t)b * taylor_vec_doubles(n);
    {
        // meta data of this CTA's items: (perm | inv, counts).  The leaf variant selects one
        // accumulator by SLOT, so it wants the inverse permutation; the others map slot -> tile.
        const unsigned long long* mg =
            reinterpret_cast<const unsigned long long*>(tcb + taylor_data_doubles(n)) +
            (size_t)i_begin * npass * F_META;
        const int n_it = (i_end - i_begin) * npass;
        for (int x = tid; x < n_it; x += blockDim.x) {
            const unsigned long long w0 = mg[(size_t)x * F_META + (MODE == APPLY_LEAF ? 2 : 0)];
            const unsigned long long w1 = mg[(size_t)x * F_META + 1];
            s_meta[x] = make_uint4((uint32_t)w0, (uint32_t)(w0 >> 32), (uint32_t)w1, (uint32_t)(w1 >> 32));
        }
    }
    for (int x = tid; x < EB; x += blockDim.x) {
        const int e = e0 + x;
        int nc = 0, child = -1, code = -2;
        double tau = 0.0, a0 = 0.0;
        if (e < a.n_items && EVAL) {
            const int item = a.items[e];
            const size_t pi = (size_t)b * a.E + item;
            const int pn = a.plan_nc[pi];
            if ((pn >> 8) == 2) {
                child = a.slot[pi];                         // P slot of this pair (< 0: overflow)
                if (child >= 0) {
                    nc = pn & 0xff;
                    tau = a.plan_tau[pi];
                    a0 = a.plan_a0[pi];
                }
            }
        } else if (e < a.n_items) {
            child = a.items[e];
            code = a.leaf_code[child];
            const size_t pi = (size_t)b * a.E + (a.branch_of ? a.branch_of[child] : child);
            const int pn = a.plan_nc[pi];
            if ((pn >> 8) != 0) {
                // 1: t == 0 -> identity (AR:146-148); 2: scaling-and-squaring path
                a.mode[(size_t)b * a.n_nodes + child] = (uint8_t)(pn >> 8);
            } else {
                nc = pn & 0xff;
                tau = a.plan_tau[pi];
                a0 = a.plan_a0[pi];
            }
        }
        s_child[x] = child;
        s_code[x] = code;
        s_nc[x] = nc;
        s_tau[x] = tau;
        s_a0[x] = a0;
    }
    __syncthreads();

    // chunks needed by this warp / by the CTA, and which consumer warps have any work
    int wk = 0, ck = 0;
    unsigned active = 0;
    for (int x = lane; x < EB; x += 32) {
        const int kc = (s_nc[x] + 3) >> 2;
        ck = max(ck, kc);
        if (kc > 0) active |= 1u << (x / (8 * G));
        if (x / (8 * G) == warp) wk = max(wk, kc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        wk = max(wk, __shfl_xor_sync(0xffffffffu, wk, o));
        ck = max(ck, __shfl_xor_sync(0xffffffffu, ck, o));
        active |= __shfl_xor_sync(0xffffffffu, active, o);
    }
    if (ck == 0) return;  // nothing on the fused path in this CTA (uniform across the CTA)

    double* Ls = As + (size_t)(EB / 8) * ck * 32;                             // [EB][ldl] (not LEAF)
    double* Bs = Ls + (LEAF ? 0 : (size_t)EB * ldl);                          // ring of item tiles
    // A operand e^{-mu tau} tau^k of every branch, chunk-major per 8-branch group
    for (int x = tid; x < EB; x += blockDim.x) {
        const int nc = s_nc[x];
        const double tau = s_tau[x];
        double* arow = As + (size_t)(x >> 3) * ck * 32 + (x & 7) * 4;
        double p = s_a0[x];  // e^{-mu tau} tau^k
        for (int k = 0; k < 4 * ck; ++k) {
            arow[(k >> 2) * 32 + (k & 3)] = (k < nc) ? p : 0.0;
            p *= tau;
        }
    }

    // Ring geometry (CTA-uniform).  A stage holds `cps` consecutive chunks of one item.  When the
    // ring has room for at least two whole items a stage is one item (cps = ck); otherwise the
    // item is streamed in parts of cps < ck chunks through two stages, so that the ring can be
    // smaller than an item of 16 chunks (53 KB) and the narrow CTA shapes keep four (3 warps) or
    // two (7 warps) CTAs per SM.  The accumulators simply keep adding across the parts.
    const int cap_chunks = (a.stage_doubles_cap + (EB / 8) * (F_KC - ck) * 32) / F_CHUNK;
    int cps = ck, ns = cap_chunks / ck;
    if (DIRECT) {
        ns = 1;
    } else if (ns >= 2) {
        if (ns > F_MAXSTAGE) ns = F_MAXSTAGE;
    } else {
        ns = 2;
        cps = cap_chunks >> 1;      // >= 1: the launcher never hands out less than two chunks
    }
    const int stage_doubles = cps * F_CHUNK;

    if (!DIRECT && tid == 0) {
        for (int s = 0; s < ns; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], __popc(active));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();   // barriers initialised, A operand in place

    if (warp != NW && !LEAF) {
        // child vectors (one-hot / uniform / conditional-likelihood vector of an internal child),
        // staged by the consumer warps while the producer already fills the ring.  Eight
        // independent loads per thread are in flight before the first store: a load-store loop
        // per element serialised ~25 HBM latencies per warp, 8-10 % of a level CTA's lifetime.
        // Columns >= n (up to the 104 * npass the slots cover) hold zeros.
        const double unif = 1.0 / (double)n;  // AR:211
        const int total = EB * ldl;
        constexpr int NT = 32 * NW;
        for (int q0 = tid; q0 < total; q0 += NT * 8) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int q = q0 + u * NT;
                v[u] = 0.0;
                if (q < total) {
                    const int x = q / ldl, j = q - x * ldl;
                    const int child = s_child[x], code = s_code[x];
                    if (child >= 0 && j < n)
                        v[u] = (code >= 0) ? (j == code ? 1.0 : 0.0)
                               : (code == -1 ? unif
                                             : a.node_vec[((size_t)b * a.n_nodes + child) * a.ldp + j]);
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int q = q0 + u * NT;
                if (q < total) Ls[q] = v[u];
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * NW) : "memory");   // consumer warps only
    }

    if (!DIRECT && warp == NW) {
        // ---- producer: one lane streams the items of the coefficient table; per chunk only the
        // active slot prefix m(c) * 256 bytes is copied (runs of full chunks are one copy)
        if (lane != 0) return;
        if (CEVB_STUB & 2) return;
        int s = 0, use = 0;
        for (int i = i_begin; i < i_end; ++i) {
            for (int pass = 0; pass < npass; ++pass) {
                const uint4 mt = s_meta[(i - i_begin) * npass + pass];
                const unsigned long long cnts = ((unsigned long long)mt.w << 32) | mt.z;
                const double* src = tcb + (size_t)(i * npass + pass) * F_ITEM;
                for (int cb = 0; cb < ck; cb += cps) {          // one stage per part of the item
                    const int ce = min(ck, cb + cps);
                    uint32_t total = 0;
                    for (int c = cb; c < ce; ++c) total += (uint32_t)((cnts >> (4 * c)) & 15ull);
                    if (use > 0) {
                        const uint32_t par = (uint32_t)((use - 1) & 1);
                        mbar_wait_backoff(&empty[s], par);
                    }
                    mbar_expect_tx(&full[s], total * 256u);
                    double* dst = Bs + (size_t)s * stage_doubles;
                    int c = cb;
                    while (c < ce) {
                        const int c_start = c;
                        uint32_t bytes = 0;
                        while (c < ce) {
                            const uint32_t m = (uint32_t)((cnts >> (4 * c)) & 15ull);
                            bytes += m * 256u;
                            ++c;
                            if (m != (uint32_t)F_JT) break;
                        }
                        if (bytes)
                            bulk_g2s(dst + (size_t)(c_start - cb) * F_CHUNK, src + (size_t)c_start * F_CHUNK,
                                     bytes, &full[s]);
                    }
                    if (++s == ns) {
                        s = 0;
                        ++use;
                    }
                }
            }
        }
        return;
    }
    if (wk == 0) return;  // consumer warp without fused branches (not counted in `empty`)

    if (a.skew > 0 && ((warp >> 2) & 1)) {
        const long long t0 = clock64();
        while (clock64() - t0 < (long long)a.skew) {
        }
    }
    // ---- consumers (one group of 8 branches per warp)
    static_assert(G == 1, "the pipelined consumer loop keeps one group per warp");
    // np.maximum(P, 1e-12) (AR:153) as a signed compare of the HIGH word only (one ISETP instead
    // of two per value): negative values and everything below 1e-12 are replaced, NaN (positive
    // high word) is kept like np.maximum keeps it.  Values that share the high word of 1e-12
    // (0x3D719799) but lie below it -- within 4.4e-19 of 1e-12 -- are kept as they are: an
    // absolute difference seven orders below the 1e-12 tolerance on P.
    constexpr int floor_hi = 0x3D719799;
    const int xw = warp * 8 + g;                       // this lane's branch within the CTA
    // LEAF: the pass / tile / half of the row that holds column `code` of this lane's branch
    int my_pass = -1, my_tl = 0, my_half = 0;
    if (MODE == APPLY_LEAF) {
        const int code = s_code[xw];
        if (code >= 0 && ((code >> 1) & 3) == t4) {
            my_pass = (code >> 3) / F_JT;
            my_tl = (code >> 3) - my_pass * F_JT;
            my_half = code & 1;
        }
    }
    // Every slot is floored at 1e-12, including the columns beyond n that the 13 * npass tiles
    // cover (their accumulators are exact zeros): what they add to the row sum is taken out
    // again at the end of the row (padfix = this lane's count of such columns times 1e-12).
    double padfix;
    {
        int cnt = 0;
        for (int j = 2 * t4; j < npass * F_JT * 8; j += 8) cnt += (j >= n) + (j + 1 >= n);
        padfix = (double)cnt * 1e-12;
    }
    const bool mine = s_nc[xw] > 0;
    const int my_child = s_child[xw];
    const unsigned int row_off = (unsigned)my_child * (unsigned)a.ldp;   // n_nodes * ldp < 2^31
    const double* Aw = As + (size_t)warp * ck * 32 + lane;
    const double* Lbase = Ls + (size_t)xw * ldl + 2 * t4;
    double rs0 = -padfix, rs1 = 0.0, dt0 = 0.0, dt1 = 0.0;
    bool have_prev = false;   // false only while the first item of the CTA is in flight
    double acc[F_JT][2];
#pragma unroll
    for (int sl = 0; sl < F_JT; ++sl) acc[sl][0] = acc[sl][1] = 0.0;

    // epilogue of accumulator slot SL of the item (row i_e, pass pass_e) whose slot -> tile map
    // (or, LEAF, selected slot) is (pm_lo, pm_hi): floor (AR:153), row sum, dot product with
    // L_child -- or the raw store for the squaring kernel
    auto epi_tile = [&](auto SLc, double p0, double p1, const int i_e, const int pass_e,
                        const uint32_t pm_lo, const uint32_t pm_hi, const bool floored) {
        constexpr int SL = decltype(SLc)::value;
        if (CEVB_STUB & 1) {
            rs0 += p0;
            return;
        }
        if (EVAL) {
            const int jt = nib<SL>(pm_lo, pm_hi);
            const int j = (pass_e * F_JT + jt) * 8 + 2 * t4;
            if (mine && have_prev && j < a.ldp)
                *reinterpret_cast<double2*>(a.X + ((size_t)my_child * n + i_e) * a.ldp + j) =
                    make_double2(p0, p1);
            return;
        }
        if (!floored) {
            p0 = (__double2hiint(p0) < floor_hi) ? 1e-12 : p0;
            p1 = (__double2hiint(p1) < floor_hi) ? 1e-12 : p1;
        }
        rs0 += p0;
        rs1 += p1;
        if (MODE == APPLY_LEAF) {
            // pm_lo = the slot of this lane's column in this item (-1: none)
            dt0 = ((int)pm_lo == SL) ? (my_half ? p1 : p0) : dt0;
        } else {
            const int jt = nib<SL>(pm_lo, pm_hi);
            const double2 l = *reinterpret_cast<const double2*>(Lbase + (pass_e * F_JT + jt) * 8);
            dt0 = fma(p0, l.x, dt0);
            dt1 = fma(p1, l.y, dt1);
        }
    };
    // end of a row: quad reduction and the two stores (the combine kernel forms d / r, AR:154).
    // The sum over the four lanes of a quad is one DMMA per quantity (partials as the 8x4 A
    // operand times a matrix of ones) instead of a chain of dependent shuffles: the result is only
    // read by the stores, so the warp keeps issuing.
    auto finish_row = [&](const int i_e, const bool store) {
        if (EVAL) return;
        if (CEVB_STUB & 8) {
            rs0 = -padfix;
            rs1 = dt0 = dt1 = 0.0;
            return;
        }
        double r0 = 0.0, r1 = 0.0, d0 = 0.0, d1 = 0.0;
        dmma884(r0, r1, rs0 + rs1, 1.0);
        dmma884(d0, d1, MODE == APPLY_LEAF ? dt0 : dt0 + dt1, 1.0);
        if (t4 == 0 && mine && store) {
            // vector base: CTA-uniform 64-bit; row offset inside the vector: 32-bit per lane
            const size_t vb = (size_t)b * a.n_nodes * a.ldp;
            const unsigned int o = row_off + (unsigned)i_e;
            (a.w + vb)[o] = d0;
            (a.wr + vb)[o] = r0;
        }
        rs0 = -padfix;   // takes the floor of the columns beyond n out of the row sum again
        rs1 = dt0 = dt1 = 0.0;
    };

    // Software pipeline over the items (row, pass): while the first chunk of item q is issued
    // to the tensor pipe, the finished accumulator slots of item q-1 are read out one by one
    // and post-processed, so a warp keeps the FP64 pipe fed during its own epilogue.
    int s = 0;
    uint32_t par = 0;
    // pass_prev starts at the last pass: the epilogue the first item runs on the zero-initialised
    // accumulators is then closed by a finish_row that stores nothing and resets the sums
    int i_prev = 0, pass_prev = npass - 1;
    uint32_t pm_prev_lo = 0, pm_prev_hi = 0;

    for (int i = i_begin; i < i_end; ++i) {
        for (int pass = 0; pass < npass; ++pass) {
            const uint4 mt = s_meta[(i - i_begin) * npass + pass];
            uint32_t c_lo = mt.z, c_hi = mt.w;
            if (!DIRECT && !(CEVB_STUB & 2)) mbar_wait_backoff(&full[s], par);
            const double* Bq = DIRECT ? tcb + (size_t)(i * npass + pass) * F_ITEM + lane
                                      : Bs + (size_t)s * stage_doubles + lane;   // part 0 of the item
            // chunk 0 is interleaved with the epilogue of the previous item, so its DMMAs are
            // predicated per slot (SL < m(0)); the later chunks branch to the straight-line
            // variant of their slot count
            const int m0 = (int)(c_lo & 15u);
            {
                // chunk 0 with the epilogue of the previous item between its DMMAs.  Straight-
                // line variants for 0, 3, 6, 9 or 13 slots (the table rounds m(0) up to one of
                // them), so no per-slot predicate is needed.  The first item of a CTA runs the
                // epilogue on its zero-initialised accumulators; nothing of that is stored.
                auto chunk0 = [&](auto Mc) {
                    constexpr int M = decltype(Mc)::value;
                    const double av = Aw[0];
                    double bv[M > 0 ? M : 1];
#pragma unroll
                    for (int sl = 0; sl < M; ++sl) bv[sl] = DIRECT ? __ldg(Bq + sl * 32) : Bq[sl * 32];
                    asm volatile("" ::: "memory");
                    auto step = [&](auto SLc) {
                        constexpr int SL = decltype(SLc)::value;
                        double p0 = acc[SL][0], p1 = acc[SL][1];
                        if (!EVAL) {
                            // the floor consumes the old accumulator values, so no register copy
                            // is needed before the DMMA overwrites the slot (AR:153)
                            p0 = (__double2hiint(p0) < floor_hi) ? 1e-12 : p0;
                            p1 = (__double2hiint(p1) < floor_hi) ? 1e-12 : p1;
                        }
                        acc[SL][0] = 0.0;
                        acc[SL][1] = 0.0;
                        if (SL < M) dmma884(acc[SL][0], acc[SL][1], av, bv[SL < M ? SL : 0]);
                        epi_tile(SLc, p0, p1, i_prev, pass_prev, pm_prev_lo, pm_prev_hi, true);
                    };
                    step(std::integral_constant<int, 0>{});
                    step(std::integral_constant<int, 1>{});
                    step(std::integral_constant<int, 2>{});
                    step(std::integral_constant<int, 3>{});
                    step(std::integral_constant<int, 4>{});
                    step(std::integral_constant<int, 5>{});
                    step(std::integral_constant<int, 6>{});
                    step(std::integral_constant<int, 7>{});
                    step(std::integral_constant<int, 8>{});
                    step(std::integral_constant<int, 9>{});
                    step(std::integral_constant<int, 10>{});
                    step(std::integral_constant<int, 11>{});
                    step(std::integral_constant<int, 12>{});
                    asm volatile("" ::: "memory");
                };
                if (m0 <= 6) {
                    if (m0 <= 3) {
                        if (m0 == 0) chunk0(std::integral_constant<int, 0>{});
                        else chunk0(std::integral_constant<int, 3>{});
                    } else {
                        chunk0(std::integral_constant<int, 6>{});
                    }
                } else {
                    if (m0 <= 9) chunk0(std::integral_constant<int, 9>{});
                    else chunk0(std::integral_constant<int, 13>{});
                }
            }
            // the remaining chunks, part by part (one part = one ring stage; a single part when
            // the whole item fits a stage).  Every active consumer warp takes every stage of the
            // item -- also the ones beyond its own wk chunks -- so the barrier counts stay fixed.
            {
                int c = 1, cbase = 0;
                for (;;) {
                    const int c_stop = min(wk, cbase + cps);
                    for (; c < c_stop; ++c) {
                        c_lo = __funnelshift_r(c_lo, c_hi, 4);
                        c_hi >>= 4;
                        const int m = (int)(c_lo & 15u);
                        const double av = Aw[(size_t)c * 32];
                        const double* Bc = Bq + (size_t)(c - cbase) * F_CHUNK;
                        if (!(CEVB_STUB & 4)) chunk_dispatch<DIRECT>(m, acc, Bc, av);
                    }
                    if (DIRECT) break;
                    // every shared-memory read of this stage has been consumed by an issued DMMA
                    // (the DMMAs are warp-wide, so all lanes' loads are complete)
                    if (lane == 0 && !(CEVB_STUB & 2)) mbar_arrive(&empty[s]);
                    if (++s == ns) {
                        s = 0;
                        par ^= 1u;
                    }
                    cbase += cps;
                    if (cbase >= ck) break;          // CTA-uniform
                    if (!(CEVB_STUB & 2)) mbar_wait_backoff(&full[s], par);
                    Bq = Bs + (size_t)s * stage_doubles + lane;
                }
            }
            if (ONEPASS || pass_prev == npass - 1) finish_row(i_prev, have_prev);
            have_prev = true;
            i_prev = i;
            pass_prev = pass;
            if (MODE == APPLY_LEAF) {
                // slot that holds this lane's observed column in this item
                const uint32_t wsel = (my_tl < 8) ? mt.x : mt.y;
                pm_prev_lo = (pass == my_pass) ? ((wsel >> (4 * (my_tl & 7))) & 15u) : 0xffffffffu;
            } else {
                pm_prev_lo = mt.x;
                pm_prev_hi = mt.y;
            }
        }
    }
    if (have_prev) {
        auto last = [&](auto SLc) {
            constexpr int SL = decltype(SLc)::value;
            epi_tile(SLc, acc[SL][0], acc[SL][1], i_prev, pass_prev, pm_prev_lo, pm_prev_hi, false);
        };
        last(std::integral_constant<int, 0>{});
        last(std::integral_constant<int, 1>{});
        last(std::integral_constant<int, 2>{});
        last(std::integral_constant<int, 3>{});
        last(std::integral_constant<int, 4>{});
        last(std::integral_constant<int, 5>{});
        last(std::integral_constant<int, 6>{});
        last(std::integral_constant<int, 7>{});
        last(std::integral_constant<int, 8>{});
        last(std::integral_constant<int, 9>{});
        last(std::integral_constant<int, 10>{});
        last(std::integral_constant<int, 11>{});
        last(std::integral_constant<int, 12>{});
        finish_row(i_prev, true);
    }
}

// ---------------------------------------------------------------------------------------------
// Leaf launch for generator vectors (n <= 104): w = max(P[i][code], 1e-12) and
// wr = sum_j max(P[i][j], 1e-12) for every observed-leaf branch, without the FP64 matrix.
//
// For a generator Q the rows of P = expm(Q t) sum to one, so
//     wr_i = 1 + corr_i,   corr_i = sum_j max(1e-12 - P_ij, 0)  <=  n * 1e-12.
// The reference's result P'_ic = max(P_ic, 1e-12) / wr_i therefore needs P_ic in full precision
// (one column: a dot product of length `terms` per row, summed here from the FP64 table), but
// corr_i only to about 1 % for an absolute error of 1e-12 in P' -- the tolerance of the parity
// tests; the log-likelihood tolerance is another three orders looser.  corr_i comes from the full
// matrix in TF32 on the tensor cores (mma.sync.m16n8k8, FP32 accumulation of non-negative terms
// scaled into [0, 1]: relative error about 1e-3 on the entries near the floor, 1e-13 absolute on
// wr), at an eighth of the issue slots of the FP64 DMMA path and without its epilogue.
// Vectors whose Q is no generator (negative linear_rate) keep the FP64 kernel: there the rows do
// not sum to one and negative entries of any size are floored.
struct LowpArgs {
    const double* tc;
    const double* norms;
    const int32_t* items;
    const int32_t* leaf_code;
    double* w;
    double* wr;
    uint8_t* mode;
    int n_items, n, ldp, n_nodes, E;
    const int32_t* plan_nc;
    const double* plan_tau;
    const double* plan_a0;
    const int32_t* branch_of;
    int ring_floats;          // floats available for the ring of TF32 row tiles
};

__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const float4& a, const float2& b) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
        "{%0,%1,%2,%3};\n"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(__float_as_uint(a.x)), "r"(__float_as_uint(a.y)), "r"(__float_as_uint(a.z)),
          "r"(__float_as_uint(a.w)), "r"(__float_as_uint(b.x)), "r"(__float_as_uint(b.y)));
}

// the first k-block of a row: D = A * B (no accumulator to clear beforehand)
__device__ __forceinline__ void mma_tf32_16x8x8_first(float (&c)[4], const float4& a, const float2& b) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
        "{%10,%10,%10,%10};\n"
        : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3])
        : "r"(__float_as_uint(a.x)), "r"(__float_as_uint(a.y)), "r"(__float_as_uint(a.z)),
          "r"(__float_as_uint(a.w)), "r"(__float_as_uint(b.x)), "r"(__float_as_uint(b.y)), "f"(0.0f));
}

constexpr int LP_MAXSTAGE = 4;

__host__ __device__ inline size_t lowp_fixed_bytes(int nw, int n) {
    const int eb = nw * 16;
    size_t b = 16 * LP_MAXSTAGE + 4 * (3 * eb + 4) + 16 * eb;  // barriers, child / code / terms, tau / a0
    b = (b + 127) / 128 * 128;
    b += (size_t)((16 * n + 127) / 128) * 128;                 // per-row meta (inv, counts)
    b += sizeof(double) * (size_t)(eb / 8) * F_KC * 32;        // FP64 t^k table (column dot product)
    b += sizeof(float) * (size_t)nw * L_KB * 32 * 4;           // TF32 A fragments
    return b;
}

// One CTA: vector blockIdx.z, branches [blockIdx.x * 16 NW, +16 NW) of the leaf list, rows
// [blockIdx.y * rows_per_cta, ...).  NW consumer warps of 16 branches each + one producer warp
// that streams the TF32 row tiles through a ring of shared-memory stages.
template <int NW>
__global__ void __launch_bounds__(32 * (NW + 1), NW >= 15 ? 1 : 2)
    leaf_lowp_kernel(LowpArgs a, int rows_per_cta) {
    extern __shared__ __align__(128) unsigned char smraw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    constexpr int EB = NW * 16;
    const int b = blockIdx.z, n = a.n;
    if (a.norms[(size_t)b * CEVB_NNORM + CEVB_NORM_GENERATOR] == 0.0) return;   // FP64 kernel's vector
    const int e0 = blockIdx.x * EB;
    const int i_begin = blockIdx.y * rows_per_cta;
    const int i_end = min(n, i_begin + rows_per_cta);

    uint64_t* full = reinterpret_cast<uint64_t*>(smraw);
    uint64_t* empty = full + LP_MAXSTAGE;
    int* s_child = reinterpret_cast<int*>(smraw + 16 * LP_MAXSTAGE);
    int* s_code = s_child + EB;
    int* s_nc = s_code + EB;
    double* s_tau = reinterpret_cast<double*>(s_nc + EB + 2);                   // 8-byte aligned
    double* s_a0 = s_tau + EB;
    size_t off = (16 * LP_MAXSTAGE + 4 * (3 * EB + 4) + 16 * EB + 127) / 128 * 128;
    uint4* s_meta = reinterpret_cast<uint4*>(smraw + off);
    off += (size_t)((16 * n + 127) / 128) * 128;
    // the FP64 t^k table [EB/8][2 ck][32], the TF32 A fragments [NW][ck][32][4] and the ring follow;
    // they are laid out once ck (the k-blocks this CTA's longest branch needs) is known
    unsigned char* dyn = smraw + off;

    const double* tcb = a.tc + (size_t)b * taylor_vec_doubles(n);
    const float* tfb = reinterpret_cast<const float*>(tcb + taylor_meta_end_doubles(n));
    const double nu = a.norms[(size_t)b * CEVB_NNORM + CEVB_NORM_SHIFTED];
    {
        const unsigned long long* mg =
            reinterpret_cast<const unsigned long long*>(tcb + taylor_data_doubles(n)) + (size_t)i_begin * F_META;
        for (int x = tid; x < i_end - i_begin; x += blockDim.x) {
            const unsigned long long w0 = mg[(size_t)x * F_META + 2];   // inv: tile -> slot
            const unsigned long long w1 = mg[(size_t)x * F_META + 1];   // counts
            s_meta[x] = make_uint4((uint32_t)w0, (uint32_t)(w0 >> 32), (uint32_t)w1, (uint32_t)(w1 >> 32));
        }
    }
    for (int x = tid; x < EB; x += blockDim.x) {
        const int e = e0 + x;
        int nc = 0, child = -1, code = -2;
        double tau = 0.0, a0 = 0.0;
        if (e < a.n_items) {
            child = a.items[e];
            code = a.leaf_code[child];
            const size_t pi = (size_t)b * a.E + (a.branch_of ? a.branch_of[child] : child);
            const int pn = a.plan_nc[pi];
            if ((pn >> 8) != 0) {
                a.mode[(size_t)b * a.n_nodes + child] = (uint8_t)(pn >> 8);   // identity / materialised
            } else {
                nc = pn & 0xff;
                tau = a.plan_tau[pi];
                a0 = a.plan_a0[pi];
            }
        }
        s_child[x] = child;
        s_code[x] = code;
        s_nc[x] = nc;
        s_tau[x] = tau;
        s_a0[x] = a0;
    }
    __syncthreads();

    int wk = 0, ck = 0;      // 8-degree blocks needed by this warp / by the CTA
    unsigned active = 0;
    for (int x = lane; x < EB; x += 32) {
        const int kb = (s_nc[x] + 7) >> 3;
        ck = max(ck, kb);
        if (kb > 0) active |= 1u << (x >> 4);
        if ((x >> 4) == warp) wk = max(wk, kb);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        wk = max(wk, __shfl_xor_sync(0xffffffffu, wk, o));
        ck = max(ck, __shfl_xor_sync(0xffffffffu, ck, o));
        active |= __shfl_xor_sync(0xffffffffu, active, o);
    }
    if (ck == 0) return;
    const int ck4 = 2 * ck;                                                     // 4-degree chunks
    double* As = reinterpret_cast<double*>(dyn);                                // [EB/8][ck4][32]
    float* Tf = reinterpret_cast<float*>(As + (size_t)(EB / 8) * ck4 * 32);     // [NW][ck][32][4]
    float* Bs = Tf + (size_t)NW * ck * 128;                                     // ring
    for (int x = tid; x < NW * ck * 128; x += blockDim.x) Tf[x] = 0.0f;
    __syncthreads();
    for (int x = tid; x < EB; x += blockDim.x) {
        const int nc = s_nc[x];
        const double tau = s_tau[x];
        double* arow = As + (size_t)(x >> 3) * ck4 * 32 + (x & 7) * 4;
        // A fragment of mma.m16n8k8 (row-major 16 x 8): lane 4g+t holds (row g, k t), (row g+8, k t),
        // (row g, k t+4), (row g+8, k t+4)
        float* trow = Tf + (size_t)(x >> 4) * ck * 128;
        const int r16 = x & 15, gg = r16 & 7, hi = r16 >> 3;
        double p = s_a0[x];   // e^{-mu tau} tau^k
        double fk = 1.0;      // nu^k / k!
        for (int k = 0; k < 4 * ck4; ++k) {
            arow[(k >> 2) * 32 + (k & 3)] = (k < nc) ? p : 0.0;
            if (k < nc) {
                const float v = (float)(p * fk);
                uint32_t u;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
                const int kk = k & 7;
                trow[((k >> 3) * 32 + gg * 4 + (kk & 3)) * 4 + hi + 2 * (kk >> 2)] = __uint_as_float(u);
            }
            p *= tau;
            fk *= nu / (double)(k + 1);
        }
    }
    // ring geometry as in taylor_apply_kernel: whole rows when two of them fit, otherwise a row is
    // streamed in parts of cps k-blocks through two stages.  The ring also takes the room the
    // two tables above leave (the shared memory is sized for ck = L_KB).
    const int cap_kb = (a.ring_floats + (L_KB - ck) * ((EB / 8) * 2 * 32 * 2 + NW * 128)) / (F_JT * L_TILE);
    int cps = ck, ns = cap_kb / ck;
    if (ns >= 2) {
        if (ns > LP_MAXSTAGE) ns = LP_MAXSTAGE;
    } else {
        ns = 2;
        cps = cap_kb >> 1;
    }
    const int stage_floats = cps * F_JT * L_TILE;
    if (tid == 0) {
        for (int s = 0; s < ns; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], __popc(active));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == NW) {
        if (lane != 0) return;
        int s = 0, use = 0;
        for (int i = i_begin; i < i_end; ++i) {
            for (int cb = 0; cb < ck; cb += cps) {
                const uint32_t bytes = (uint32_t)(min(ck, cb + cps) - cb) * F_JT * L_TILE * sizeof(float);
                if (use > 0) mbar_wait_backoff(&empty[s], (uint32_t)((use - 1) & 1));
                mbar_expect_tx(&full[s], bytes);
                bulk_g2s(Bs + (size_t)s * stage_floats, tfb + (size_t)i * L_ROW + (size_t)cb * F_JT * L_TILE,
                         bytes, &full[s]);
                if (++s == ns) {
                    s = 0;
                    ++use;
                }
            }
        }
        return;
    }
    if (wk == 0) return;

    // ---- consumers: 16 branches per warp
    const int xq = warp * 16 + (lane >> 1);            // the branch whose column this lane pair sums
    const int part = lane & 1;
    const int q_code = s_code[xq], q_child = s_child[xq];
    const int q_kc = (s_nc[xq] + 3) >> 2;              // its 4-degree chunks
    const bool q_ok = q_child >= 0 && q_kc > 0 && q_code >= 0;
    const int q_tile = q_ok ? (q_code >> 3) : 0, q_col = q_code & 7;
    const double* q_T = As + (size_t)(xq >> 3) * ck4 * 32 + (xq & 7) * 4;
    const size_t vb = (size_t)b * a.n_nodes * a.ldp;
    // columns >= n of the last tile hold zeros: they count 1e-12 each and are taken out again
    float padfix;
    {
        int cnt = 0;
        for (int j = 2 * t4; j < F_JT * 8; j += 8) cnt += (j >= n) + (j + 1 >= n);
        padfix = (float)cnt * 1e-12f;
    }
    const float4* Tw = reinterpret_cast<const float4*>(Tf) + (size_t)warp * ck * 32 + lane;
    int s = 0;
    uint32_t par = 0;
    for (int i = i_begin; i < i_end; ++i) {
        const uint4 mt = s_meta[i - i_begin];
        // (1) this lane pair's column of P in FP64 from the coefficient table (chunks of the
        // branch alternate between the two lanes); the loads are in flight during the MMAs
        double col = 0.0;
        constexpr int PF = 3;                 // chunks per lane whose loads fly during the MMAs
        double2 cf[PF][2];
        bool has[PF];
        int slot = 0;
        unsigned long long cnts = 0ull;
        const double* base = tcb;
        if (q_ok) {
            const uint32_t invw = (q_tile < 8) ? mt.x : mt.y;
            slot = (int)((invw >> (4 * (q_tile & 7))) & 15u);
            cnts = ((unsigned long long)mt.w << 32) | mt.z;
            base = tcb + (size_t)i * F_ITEM + slot * 32 + q_col * 4;
        }
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            const int c = part + 2 * u;
            has[u] = q_ok && c < q_kc && slot < (int)((cnts >> (4 * c)) & 15ull);
            if (has[u]) {
                cf[u][0] = __ldg(reinterpret_cast<const double2*>(base + (size_t)c * F_CHUNK));
                cf[u][1] = __ldg(reinterpret_cast<const double2*>(base + (size_t)c * F_CHUNK) + 1);
            }
        }
        // (2) the whole row in TF32
        mbar_wait_backoff(&full[s], par);
        const float2* Bq = reinterpret_cast<const float2*>(Bs + (size_t)s * stage_floats) + lane;
        float acc[F_JT][4];
        {
            const float4 af = Tw[0];
#pragma unroll
            for (int jt = 0; jt < F_JT; ++jt) mma_tf32_16x8x8_first(acc[jt], af, Bq[jt * 32]);
        }
        {
            int kb = 1, kbase = 0;
            for (;;) {
                const int k_stop = min(wk, kbase + cps);
                for (; kb < k_stop; ++kb) {
                    const float4 af = Tw[(size_t)kb * 32];
                    const float2* Bk = Bq + (size_t)(kb - kbase) * F_JT * 32;
#pragma unroll
                    for (int jt = 0; jt < F_JT; ++jt) mma_tf32_16x8x8(acc[jt], af, Bk[jt * 32]);
                }
                if (lane == 0) mbar_arrive(&empty[s]);
                if (++s == ns) {
                    s = 0;
                    par ^= 1u;
                }
                kbase += cps;
                if (kbase >= ck) break;           // CTA-uniform: every warp takes every stage
                mbar_wait_backoff(&full[s], par);
                Bq = reinterpret_cast<const float2*>(Bs + (size_t)s * stage_floats) + lane;
            }
        }
        // (1b) the column dot product from the loads issued above (and the rare later chunks)
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            if (has[u]) {
                const int c = part + 2 * u;
                const double2 t01 = *reinterpret_cast<const double2*>(q_T + c * 32);
                const double2 t23 = *reinterpret_cast<const double2*>(q_T + c * 32 + 2);
                col = fma(t01.x, cf[u][0].x, col);
                col = fma(t01.y, cf[u][0].y, col);
                col = fma(t23.x, cf[u][1].x, col);
                col = fma(t23.y, cf[u][1].y, col);
            }
        }
        if (q_ok) {
            for (int c = part + 2 * PF; c < q_kc; c += 2) {
                if (slot < (int)((cnts >> (4 * c)) & 15ull)) {
                    const double2 c01 = __ldg(reinterpret_cast<const double2*>(base + (size_t)c * F_CHUNK));
                    const double2 c23 = __ldg(reinterpret_cast<const double2*>(base + (size_t)c * F_CHUNK) + 1);
                    const double2 t01 = *reinterpret_cast<const double2*>(q_T + c * 32);
                    const double2 t23 = *reinterpret_cast<const double2*>(q_T + c * 32 + 2);
                    col = fma(t01.x, c01.x, col);
                    col = fma(t01.y, c01.y, col);
                    col = fma(t23.x, c23.x, col);
                    col = fma(t23.y, c23.y, col);
                }
            }
        }
        // (3) corr = sum_j max(1e-12 - P_ij, 0) for rows g and g + 8 of the group
        // as 26 * 1e-12 - sum_j min(P_ij, 1e-12) per lane: one min and one add per entry (FP32 keeps
        // 1e-17 absolute on sums of this size)
        float m0 = 0.0f, m1 = 0.0f;
#pragma unroll
        for (int jt = 0; jt < F_JT; ++jt) {
            m0 += fminf(acc[jt][0], 1e-12f) + fminf(acc[jt][1], 1e-12f);
            m1 += fminf(acc[jt][2], 1e-12f) + fminf(acc[jt][3], 1e-12f);
        }
        float c0 = (26.0f * 1e-12f - padfix) - m0, c1 = (26.0f * 1e-12f - padfix) - m1;
        c0 += __shfl_xor_sync(0xffffffffu, c0, 1);
        c0 += __shfl_xor_sync(0xffffffffu, c0, 2);
        c1 += __shfl_xor_sync(0xffffffffu, c1, 1);
        c1 += __shfl_xor_sync(0xffffffffu, c1, 2);
        // branch (lane >> 1) of the group: rows 0..7 sit in c0 of quad g, rows 8..15 in c1
        const int r16 = lane >> 1;
        const float cq0 = __shfl_sync(0xffffffffu, c0, (r16 & 7) * 4);
        const float cq1 = __shfl_sync(0xffffffffu, c1, (r16 & 7) * 4);
        col += __shfl_xor_sync(0xffffffffu, col, 1);
        if (part == 0 && q_ok) {
            const double corr = (double)((r16 < 8) ? cq0 : cq1);
            const unsigned int o = (unsigned)q_child * (unsigned)a.ldp + (unsigned)i;
            (a.w + vb)[o] = (col < 1e-12) ? 1e-12 : col;          // AR:153 on the one entry that is read
            (a.wr + vb)[o] = 1.0 + corr;                           // AR:154: rows of a generator sum to one
        }
    }
}

// ---------------------------------------------------------------------------------------------
constexpr int CT = 128;

__device__ double block_sum_ct(double v, double* red) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    const double r = (red[0] + red[1]) + (red[2] + red[3]);
    __syncthreads();
    return r;
}

// One CTA per (parent node of the level, vector): product over the children in file order of
// the per-branch vectors w = P @ L_child (fused path), L_child itself (P = I) or a streamed
// materialised P (Pade path), with the reference's conditional rescale after every child.
__global__ void __launch_bounds__(CT) combine_level_kernel(
    const double* __restrict__ w, const double* __restrict__ wr, const uint8_t* __restrict__ mode,
    const double* __restrict__ P,
    const int32_t* __restrict__ slot, int E, int n, int ldp, int n_nodes, int root,
    const int32_t* __restrict__ level_nodes, const int32_t* __restrict__ child_off,
    const int32_t* __restrict__ child_idx, const int32_t* __restrict__ branch_of,
    const int32_t* __restrict__ leaf_code, const double* __restrict__ root_freq,
    double* __restrict__ node_vec, double* __restrict__ logl, int32_t* __restrict__ rescales) {
    extern __shared__ double shv[];
    double* v = shv;
    double* Lc = shv + n;
    __shared__ double red[4];
    const int node = level_nodes[blockIdx.x], b = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int j = tid; j < n; j += CT) v[j] = 1.0;  // AR:219
    int nres = 0;
    __syncthreads();
    const double unif = 1.0 / (double)n;
    const int c_begin = child_off[node], c_end = child_off[node + 1];
    for (int ci = c_begin; ci < c_end; ++ci) {
        const int child = child_idx[ci];
        const int code = leaf_code[child];
        int md = mode[(size_t)b * n_nodes + child];
        const double* cv = node_vec + ((size_t)b * n_nodes + child) * ldp;
        double fac[(CEVB_MAX_N + CT - 1) / CT];
        if (md == 0) {
            // a NaN or +inf entry of P makes its row sum non-finite: the reference then finds
            // NaN in P / rowsum and returns the identity for the whole matrix (AR:156-159)
            const double* wv = w + ((size_t)b * n_nodes + child) * ldp;
            const double* rv = wr + ((size_t)b * n_nodes + child) * ldp;
            int bad = 0;
#pragma unroll
            for (int k = 0; k < (CEVB_MAX_N + CT - 1) / CT; ++k) {
                const int i = tid + k * CT;
                fac[k] = 1.0;
                if (i < n) {
                    const double r = rv[i];
                    bad |= !(fabs(r) < INFINITY);
                    fac[k] = wv[i] / r;
                }
            }
            if (__syncthreads_or(bad)) md = 1;
        }
        if (md == 0) {
#pragma unroll
            for (int k = 0; k < (CEVB_MAX_N + CT - 1) / CT; ++k) {
                const int i = tid + k * CT;
                if (i < n) v[i] *= fac[k];
            }
        } else if (md == 1) {
            for (int i = tid; i < n; i += CT)
                v[i] *= (code >= 0) ? (i == code ? 1.0 : 0.0) : (code == -1 ? unif : cv[i]);
        } else {
            long long sl = slot[(size_t)b * E + branch_of[child]];
            if (sl < 0) sl = 0;  // work-list overflow: the host discards this pass and re-runs
            const double* Pm = P + (size_t)sl * n * ldp;
            if (code >= 0) {
                for (int i = tid; i < n; i += CT) v[i] *= Pm[(size_t)i * ldp + code];
            } else {
                for (int j = tid; j < n; j += CT) Lc[j] = (code == -1) ? unif : cv[j];
                __syncthreads();
                // four rows per warp in flight: this path is a single CTA streaming 80 KB, so it
                // is bound by how many loads it keeps outstanding
                for (int i = warp * 4; i < n; i += 4 * (CT / 32)) {
                    const double* r0 = Pm + (size_t)i * ldp;
                    const double* r1 = (i + 1 < n) ? r0 + ldp : r0;
                    const double* r2 = (i + 2 < n) ? r0 + 2 * ldp : r0;
                    const double* r3 = (i + 3 < n) ? r0 + 3 * ldp : r0;
                    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                    for (int j = lane; j < n; j += 32) {
                        const double l = Lc[j];
                        a0 = fma(r0[j], l, a0);
                        a1 = fma(r1[j], l, a1);
                        a2 = fma(r2[j], l, a2);
                        a3 = fma(r3[j], l, a3);
                    }
                    a0 = warp_sum(a0);
                    a1 = warp_sum(a1);
                    a2 = warp_sum(a2);
                    a3 = warp_sum(a3);
                    if (lane == 0) {
                        v[i] *= a0;
                        if (i + 1 < n) v[i + 1] *= a1;
                        if (i + 2 < n) v[i + 2] *= a2;
                        if (i + 3 < n) v[i + 3] *= a3;
                    }
                }
            }
        }
        __syncthreads();
        double part = 0.0;
        for (int j = tid; j < n; j += CT) part += v[j];
        const double s = block_sum_ct(part, red);  // AR:241
        if (s > 0.0 && s < 1e-100) {               // AR:242-243
            for (int j = tid; j < n; j += CT) v[j] = v[j] / s;
            ++nres;
        }
        __syncthreads();
    }
    double* out = node_vec + ((size_t)b * n_nodes + node) * ldp;
    for (int j = tid; j < ldp; j += CT) out[j] = (j < n) ? v[j] : 0.0;
    if (tid == 0 && nres) atomicAdd(&rescales[b], nres);
    if (node == root) {
        double part = 0.0;
        for (int j = tid; j < n; j += CT) part += v[j] * root_freq[j];  // AR:268
        const double L = block_sum_ct(part, red);
        if (tid == 0) logl[b] = (L <= 0.0) ? -INFINITY : log(L);        // AR:270-272
    }
}

// Same, one WARP per (parent node, vector) for n <= 128: four entries per lane in registers, warp
// shuffles instead of block barriers, a quarter of the CTAs.  The level kernels are tiny, so their
// cost is CTA scheduling and barrier latency rather than bytes.
__global__ void __launch_bounds__(128) combine_level_warp_kernel(
    const double* __restrict__ w, const double* __restrict__ wr, const uint8_t* __restrict__ mode,
    const double* __restrict__ P, const int32_t* __restrict__ slot, int E, int n, int ldp,
    int n_nodes, int root, int n_level, int B, const int32_t* __restrict__ level_nodes,
    const int32_t* __restrict__ child_off, const int32_t* __restrict__ child_idx,
    const int32_t* __restrict__ branch_of, const int32_t* __restrict__ leaf_code,
    const double* __restrict__ root_freq, double* __restrict__ node_vec, double* __restrict__ logl,
    int32_t* __restrict__ rescales) {
    const int lane = threadIdx.x & 31;
    const long long task = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (task >= (long long)n_level * B) return;
    const int b = (int)(task / n_level);
    const int node = level_nodes[task - (long long)b * n_level];
    double v[4] = {1.0, 1.0, 1.0, 1.0};  // AR:219; entry lane + 32 k
    int nres = 0;
    const double unif = 1.0 / (double)n;
    const int c_begin = child_off[node], c_end = child_off[node + 1];
    // one child: v *= P_child @ L_child and the AR:241 rescale; (wp, rp) are this child's rows of
    // w / wr, loaded by the caller for two children at a time so that their latencies overlap
    auto apply_child = [&](const int child, const int code, int md, const double (&wp)[4],
                           const double (&rp)[4]) {
        const double* cv = node_vec + ((size_t)b * n_nodes + child) * ldp;
        double x[4];
        if (md == 0) {
            int bad = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = lane + 32 * k;
                x[k] = 1.0;
                if (i < n) {
                    bad |= !(fabs(rp[k]) < INFINITY);   // NaN / inf in P -> identity (AR:156-159)
                    x[k] = wp[k] / rp[k];
                }
            }
            if (__any_sync(0xffffffffu, bad)) md = 1;
        }
        if (md == 1) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = lane + 32 * k;
                x[k] = (i < n) ? ((code >= 0) ? (i == code ? 1.0 : 0.0) : (code == -1 ? unif : cv[i])) : 1.0;
            }
        } else if (md == 2) {
            long long sl = slot[(size_t)b * E + branch_of[child]];
            if (sl < 0) sl = 0;  // work-list overflow: the host discards this pass and re-runs
            const double* Pm = P + (size_t)sl * n * ldp;
            if (code >= 0) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int i = lane + 32 * k;
                    x[k] = (i < n) ? Pm[(size_t)i * ldp + code] : 1.0;
                }
            } else {
                double l[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int j = lane + 32 * k;
                    l[k] = (j < n) ? ((code == -1) ? unif : cv[j]) : 0.0;
                    x[k] = 1.0;
                }
                for (int i0 = 0; i0 < n; i0 += 4) {   // four rows in flight
                    double a[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const double* row = Pm + (size_t)min(i0 + r, n - 1) * ldp;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int j = lane + 32 * k;
                            if (j < n) a[r] = fma(row[j], l[k], a[r]);
                        }
                    }
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const double sum = warp_sum(a[r]);
                        const int i = i0 + r;
                        if (i < n && (i & 31) == lane) {
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                if ((i >> 5) == k) x[k] = sum;
                        }
                    }
                }
            }
        }
        double part = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            v[k] *= x[k];                                   // AR:237-238
            if (lane + 32 * k < n) part += v[k];
        }
        const double s = warp_sum(part);                    // AR:241
        if (s > 0.0 && s < 1e-100) {                        // AR:242-243
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = v[k] / s;
            ++nres;
        }
    };
    for (int ci = c_begin; ci < c_end; ci += 2) {
        const bool two = ci + 1 < c_end;
        const int child0 = child_idx[ci], child1 = two ? child_idx[ci + 1] : child0;
        const int code0 = leaf_code[child0], code1 = leaf_code[child1];
        const int md0 = mode[(size_t)b * n_nodes + child0], md1 = mode[(size_t)b * n_nodes + child1];
        // rows of w / wr exist for every node; they are only used where mode == 0
        const double* w0 = w + ((size_t)b * n_nodes + child0) * ldp;
        const double* r0 = wr + ((size_t)b * n_nodes + child0) * ldp;
        const double* w1 = w + ((size_t)b * n_nodes + child1) * ldp;
        const double* r1 = wr + ((size_t)b * n_nodes + child1) * ldp;
        double wp0[4], rp0[4], wp1[4], rp1[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = lane + 32 * k;
            const bool in = i < n;
            wp0[k] = in ? w0[i] : 1.0;      // not predicated on mode: the loads must not wait for it
            rp0[k] = in ? r0[i] : 1.0;
            wp1[k] = in ? w1[i] : 1.0;
            rp1[k] = in ? r1[i] : 1.0;
        }
        apply_child(child0, code0, md0, wp0, rp0);
        if (two) apply_child(child1, code1, md1, wp1, rp1);
    }
    double* out = node_vec + ((size_t)b * n_nodes + node) * ldp;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int j = lane + 32 * k;
        if (j < ldp) out[j] = (j < n) ? v[k] : 0.0;
    }
    if (lane == 0 && nres) atomicAdd(&rescales[b], nres);
    if (node == root) {
        double part = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int j = lane + 32 * k;
            if (j < n) part += v[k] * root_freq[j];         // AR:268
        }
        const double L = warp_sum(part);
        if (lane == 0) logl[b] = (L <= 0.0) ? -INFINITY : log(L);   // AR:270-272
    }
}

// leaf vectors: one-hot at the observed count, uniform 1/n for missing data (AR:201-214)
__global__ void __launch_bounds__(128) leaf_init_fused_kernel(int n, int ldp, int n_nodes,
                                                              const int32_t* __restrict__ leaf_code,
                                                              double* __restrict__ node_vec) {
    const int node = blockIdx.x, b = blockIdx.y;
    const int code = leaf_code[node];
    if (code < -1) return;  // internal node
    double* v = node_vec + ((size_t)b * n_nodes + node) * ldp;
    const double u = 1.0 / (double)n;
    for (int j = threadIdx.x; j < ldp; j += 128) {
        double x = 0.0;
        if (j < n) x = (code >= 0) ? (j == code ? 1.0 : 0.0) : u;
        v[j] = x;
    }
}

static int g_f_sm = 0;
static size_t g_f_optin = 0;
static int fused_query() {
    if (g_f_sm) return 0;
    int dev = 0, sm = 0, optin = 0;
    CEVB_CHECK_CUDA(cudaGetDevice(&dev));
    CEVB_CHECK_CUDA(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev));
    CEVB_CHECK_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    g_f_sm = sm;
    g_f_optin = (size_t)optin;
    return 0;
}

// Development knobs, read from the environment ONCE (first use), never on the launch path:
//   CEVB_TW / CEVB_TM / CEVB_TS   branch-count thresholds of the CTA shapes (launch_apply_shapes)
//   CEVB_DENSE=1                  coefficient table without the sparse slot order (A/B runs)
struct Tuning {
    int tw, tm, ts, td, dense, skew, ring_small, ring_medium, ring_wide, ring_lowp, lowp, eval_ry;
};
static const Tuning& tuning() {
    static const Tuning t = [] {
        auto env = [](const char* name, int dflt) {
            const char* v = getenv(name);
            return v ? atoi(v) : dflt;
        };
        Tuning x;
        x.tw = env("CEVB_TW", 200);   // wide: 15 consumer warps (one 512-thread CTA per SM: slower per launch from 200 branches down, and it keeps the other stream slices off the SM)
        x.tm = env("CEVB_TM", 100);   // medium: 7
        x.ts = env("CEVB_TS", 8);     // small: 3, else 1
        x.td = env("CEVB_TD", 16);    // at most this many branches: the ring-less variant
        x.dense = env("CEVB_DENSE", 0);
        x.skew = env("CEVB_SKEW", 0);
        x.ring_small = env("CEVB_RING_S", 16);   // most chunk slots in the ring of the 1- and 3-warp shapes
        x.ring_medium = env("CEVB_RING_M", 16);  // ... of the 7-warp shape
        x.ring_wide = env("CEVB_RING_W", F_MAXSTAGE * F_KC);   // ... of the 15-warp shape
        x.ring_lowp = env("CEVB_RING_L", LP_MAXSTAGE * L_KB);  // k-block slots of leaf_lowp_kernel
        // (small values force the multi-part streaming of an item: tests/test_gpu_ring_parts.py)
        x.lowp = env("CEVB_LOWP", 1);            // TF32 floor correction for the leaf launch
        x.eval_ry = env("CEVB_EVAL_RY", 2);      // row split of the squaring path's series launch (0.69 -> 0.47 ms per 512 vectors)
        return x;
    }();
    return t;
}

template <int G, int NW, int MODE, bool ONEPASS>
static int launch_apply_one(ApplyArgs& a, int B, cudaStream_t st) {
    constexpr int EB = NW * G * 8;
    // rows of every P are split over blockIdx.y when the launch has too few CTAs to fill the GPU
    const int gx = (a.n_items + EB - 1) / EB;
    const long long want = 4LL * g_f_sm;
    int ry = 1;
    if ((long long)gx * B < want) {
        ry = (int)((want + (long long)gx * B - 1) / ((long long)gx * B));
        const int ry_max = (a.n + 7) / 8;
        if (ry > ry_max) ry = ry_max;
    }
    // The series launch of the squaring path: a handful of CTAs (the longest branches of the few
    // vectors that need squarings) do all the work, 64 terms on every row, while the rest of the
    // grid leaves at once -- its duration is the life of ONE working CTA, so the rows are split
    if (MODE == APPLY_EVAL && ry < tuning().eval_ry) ry = tuning().eval_ry;
    a.rows_per_cta = (a.n + ry - 1) / ry;
    ry = (a.n + a.rows_per_cta - 1) / a.rows_per_cta;
    a.meta_bytes = (int)apply_meta_bytes(a.rows_per_cta * a.npass);
    const size_t fixed = apply_fixed_bytes(EB, MODE != APPLY_GENERAL ? 0 : a.ldl,
                                           a.rows_per_cta * a.npass);
    // Ring: as many chunk slots as fit beside the fixed part when the SM holds the CTAs its shape
    // is meant for (15 warps: 1, 7: 2, 3 and 1: 4); fewer CTAs per SM when the child vectors of a
    // large n leave no room.  At least two chunks (the kernel streams an item in parts).
    const size_t one_chunk = sizeof(double) * (size_t)F_CHUNK;
    const int ring_cap = NW >= 15 ? tuning().ring_wide : (NW == 7 ? tuning().ring_medium : tuning().ring_small);
    size_t ring = 0;
    for (int per_sm = (NW >= 15 ? 1 : (NW == 7 ? 2 : 4)); per_sm >= 1 && ring == 0; --per_sm) {
        const size_t budget = g_f_optin / per_sm - (per_sm > 1 ? 1024 : 0);
        if (budget < fixed + 2 * one_chunk) continue;
        size_t chunks = (budget - fixed) / one_chunk;
        if (chunks > (size_t)ring_cap) chunks = (size_t)ring_cap;
        if (per_sm > 1 && chunks < 4) continue;      // too shallow a ring: fewer CTAs per SM
        ring = chunks * one_chunk;
    }
    if (ring == 0) return 1;  // does not fit: caller tries a smaller variant
    a.stage_doubles_cap = (int)(ring / sizeof(double));
    a.skew = tuning().skew;
    const size_t smem = fixed + ring;
    dim3 grid(gx, ry, B);
    // the opt-in shared-memory size of an instantiation is raised once (the largest request so far)
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        CEVB_CHECK_CUDA(cudaFuncSetAttribute(taylor_apply_kernel<G, NW, MODE, ONEPASS>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem = smem;
    }
    taylor_apply_kernel<G, NW, MODE, ONEPASS><<<grid, 32 * (NW + 1), smem, st>>>(a);
    CEVB_CHECK_LAUNCH("taylor_apply_kernel");
    return 0;
}

template <int G, int NW, int MODE>
static int launch_apply_variant(ApplyArgs& a, int B, cudaStream_t st) {
    if (a.npass == 1) return launch_apply_one<G, NW, MODE, true>(a, B, st);
    return launch_apply_one<G, NW, MODE, false>(a, B, st);
}

// The ring-less variant for small launches: rows are split finely over blockIdx.y so that every
// SM holds its 16 consumer warps (16 / NW CTAs).
template <int NW, int MODE, bool ONEPASS>
static int launch_apply_direct(ApplyArgs& a, int B, cudaStream_t st) {
    constexpr int EB = NW * 8;
    const int gx = (a.n_items + EB - 1) / EB;
    const long long want = 3LL * (16 / NW) * g_f_sm;       // three waves of resident CTAs
    int ry = 1;
    if ((long long)gx * B < want) {
        ry = (int)((want + (long long)gx * B - 1) / ((long long)gx * B));
        const int ry_max = (a.n + 7) / 8;
        if (ry > ry_max) ry = ry_max;
    }
    a.rows_per_cta = (a.n + ry - 1) / ry;
    ry = (a.n + a.rows_per_cta - 1) / a.rows_per_cta;
    a.meta_bytes = (int)apply_meta_bytes(a.rows_per_cta * a.npass);
    const size_t smem = apply_fixed_bytes(EB, MODE != APPLY_GENERAL ? 0 : a.ldl, a.rows_per_cta * a.npass);
    if (smem * (16 / NW) > g_f_optin) return 1;             // large n: the ring variants take over
    a.stage_doubles_cap = F_ITEM;
    a.skew = 0;
    dim3 grid(gx, ry, B);
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        CEVB_CHECK_CUDA(cudaFuncSetAttribute(taylor_apply_kernel<1, NW, MODE, ONEPASS, true>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem = smem;
    }
    taylor_apply_kernel<1, NW, MODE, ONEPASS, true><<<grid, 32 * NW, smem, st>>>(a);
    CEVB_CHECK_LAUNCH("taylor_apply_kernel (direct)");
    return 0;
}

template <int MODE>
static int launch_apply_shapes(ApplyArgs& a, int B, cudaStream_t st) {
    // CTA shapes.  The kernel is latency-bound per warp (every warp alternates a DMMA phase and
    // an epilogue phase), so many narrow CTAs per SM beat few wide ones except for the long
    // all-leaves list, where wide CTAs share each coefficient tile among 128 branches.
    // (thresholds measured on B200 with tools/level_probe.py)
    int rc = 1;
    const Tuning& tn = tuning();
    if (MODE != APPLY_EVAL && a.n_items <= tn.td) {
        // small launches: no ring, one 8-branch group per CTA, 16 consumer warps per SM
        rc = a.npass == 1 ? launch_apply_direct<1, MODE, true>(a, B, st)
                          : launch_apply_direct<1, MODE, false>(a, B, st);
        if (rc != 1) return rc;
    }
    if (a.n_items > tn.tw) rc = launch_apply_variant<1, 15, MODE>(a, B, st);
    if (rc == 1 && a.n_items > tn.tm) rc = launch_apply_variant<1, 7, MODE>(a, B, st);
    if (rc == 1 && a.n_items > tn.ts) rc = launch_apply_variant<1, 3, MODE>(a, B, st);
    if (rc == 1 && a.n_items > tn.ts) rc = launch_apply_variant<1, 7, MODE>(a, B, st);  // large n
    if (rc == 1) rc = launch_apply_variant<1, 1, MODE>(a, B, st);
    return rc;
}

// ---------------------------------------------------------------------------------------------
// Squaring + post-processing of the pairs with t ||Q||_1 > theta: X <- X^(2^s) on the FP64
// tensor cores with both matrices resident in shared memory, then max(X, 1e-12), row
// renormalisation, NaN -> identity (AR:153-159), written back over X.  Persistent CTAs pull
// work-list positions from an atomic counter.
constexpr int SQ_NT = 512;
constexpr int SQ_NW = SQ_NT / 32;

__host__ __device__ inline size_t square_smem_bytes(int npad) {
    return sizeof(double) * 2 * (size_t)npad * (npad + 4) + 64;
}

__global__ void __launch_bounds__(SQ_NT, 1) square_post_kernel(double* __restrict__ X,
                                                               const int32_t* __restrict__ work_s,
                                                               const unsigned int* __restrict__ count,
                                                               long long cap, unsigned int* counter,
                                                               int n, int ldp, int npad,
                                                               int32_t* __restrict__ nan_count) {
    extern __shared__ __align__(16) double smq[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int ldw = npad + 4, n8 = npad >> 3;
    double* buf0 = smq;
    double* buf1 = smq + (size_t)npad * ldw;
    int* sh_q = reinterpret_cast<int*>(buf1 + (size_t)npad * ldw);
    long long lim = (long long)*count;
    if (lim > cap) lim = cap;
    // column tiles are split into nch nearly equal chunks of at most 4 tiles
    const int nch = (n8 + 3) >> 2, cbase = n8 / nch, cextra = n8 % nch;
    for (;;) {
        if (tid == 0) sh_q[0] = (int)atomicAdd(counter, 1u);
        __syncthreads();
        const long long q = (long long)(unsigned)sh_q[0];
        if (q >= lim) break;
        double* Xm = X + (size_t)q * n * ldp;
        const int s = work_s[q];
        // load (zero padding rows / columns)
        for (int i = warp; i < npad; i += SQ_NW) {
            for (int j = 2 * lane; j < ldw; j += 64) {
                double2 v = make_double2(0.0, 0.0);
                if (i < n && j < ldp) {
                    v = *reinterpret_cast<const double2*>(Xm + (size_t)i * ldp + j);
                    if (j + 1 >= n) v.y = 0.0;
                }
                *reinterpret_cast<double2*>(buf0 + (size_t)i * ldw + j) = v;
            }
        }
        __syncthreads();
        double* src = buf0;
        double* dst = buf1;
        for (int it = 0; it < s; ++it) {
            const int ntask = n8 * nch;
            for (int task = warp; task < ntask; task += SQ_NW) {
                const int ch = task / n8, rt = task - ch * n8;
                const int ct0 = ch * cbase + min(ch, cextra);
                const int nct = cbase + (ch < cextra ? 1 : 0);
                double acc[4][2] = {};
                const double* arow = src + (size_t)(rt * 8 + g) * ldw + t4;
                const double* bcol = src + (size_t)t4 * ldw + ct0 * 8 + g;
                for (int k4 = 0; k4 < npad; k4 += 4) {
                    const double av = arow[k4];
                    const double* bp = bcol + (size_t)k4 * ldw;
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (c < nct) dmma884(acc[c][0], acc[c][1], av, bp[c * 8]);
                }
                double* drow = dst + (size_t)(rt * 8 + g) * ldw + ct0 * 8 + 2 * t4;
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (c < nct)
                        *reinterpret_cast<double2*>(drow + c * 8) = make_double2(acc[c][0], acc[c][1]);
            }
            __syncthreads();
            double* tmp = src;
            src = dst;
            dst = tmp;
        }
        // floor, renormalise, NaN check, store (AR:153-159)
        int nanflag = 0;
        for (int i = warp; i < n; i += SQ_NW) {
            const double* xr = src + (size_t)i * ldw;
            double rs = 0.0;
            for (int j = 2 * lane; j < ldp; j += 64) {
                const double2 x = *reinterpret_cast<const double2*>(xr + j);
                const double a0 = (x.x != x.x) ? x.x : fmax(x.x, 1e-12);
                const double a1 = (j + 1 < n) ? ((x.y != x.y) ? x.y : fmax(x.y, 1e-12)) : 0.0;
                rs += a0 + a1;
            }
            rs = warp_sum(rs);
            for (int j = 2 * lane; j < ldp; j += 64) {
                const double2 x = *reinterpret_cast<const double2*>(xr + j);
                double2 o;
                o.x = ((x.x != x.x) ? x.x : fmax(x.x, 1e-12)) / rs;
                o.y = (j + 1 < n) ? ((x.y != x.y) ? x.y : fmax(x.y, 1e-12)) / rs : 0.0;
                if (o.x != o.x || o.y != o.y) nanflag = 1;
                *reinterpret_cast<double2*>(Xm + (size_t)i * ldp + j) = o;
            }
        }
        nanflag = __syncthreads_or(nanflag);
        if (nanflag) {
            for (int i = warp; i < n; i += SQ_NW)
                for (int j = lane; j < ldp; j += 32) Xm[(size_t)i * ldp + j] = (i == j) ? 1.0 : 0.0;
            if (tid == 0 && nan_count != nullptr) atomicAdd(nan_count, 1);
        }
        __syncthreads();
    }
}


// The same for matrix orders whose two buffers do not fit in shared memory (n > 112, e.g. the
// N = 256 shape): operands are read from global memory / L2 (a 0.5 MB matrix stays in L2), the
// square goes to a per-CTA scratch matrix, the roles alternate.  Slower per flop than the
// shared-memory kernel, but it keeps the series + squaring path -- and with it the fused engine
// -- independent of the Pade kernels at every supported order.
__global__ void __launch_bounds__(SQ_NT, 1) square_post_global_kernel(
    double* __restrict__ X, const int32_t* __restrict__ work_s, const unsigned int* __restrict__ count,
    long long cap, unsigned int* counter, int n, int ldp, int npad, double* __restrict__ scratch,
    int32_t* __restrict__ nan_count) {
    __shared__ int sh_q[2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int n8 = npad >> 3;
    long long lim = (long long)*count;
    if (lim > cap) lim = cap;
    double* scr = scratch + (size_t)blockIdx.x * n * ldp;
    const int nch = (n8 + 3) >> 2, cbase = n8 / nch, cextra = n8 % nch;
    for (;;) {
        if (tid == 0) sh_q[0] = (int)atomicAdd(counter, 1u);
        __syncthreads();
        const long long q = (long long)(unsigned)sh_q[0];
        if (q >= lim) break;
        double* Xm = X + (size_t)q * n * ldp;
        const int s = work_s[q];
        double* src = Xm;
        double* dst = scr;
        for (int it = 0; it < s; ++it) {
            const int ntask = n8 * nch;
            for (int task = warp; task < ntask; task += SQ_NW) {
                const int ch = task / n8, rt = task - ch * n8;
                const int ct0 = ch * cbase + min(ch, cextra);
                const int nct = cbase + (ch < cextra ? 1 : 0);
                double acc[4][2] = {};
                const int arow_i = rt * 8 + g;                 // row of the A fragment
                const bool arow_ok = arow_i < n;
                const double* arow = src + (size_t)(arow_ok ? arow_i : 0) * ldp + t4;
                for (int k4 = 0; k4 < npad; k4 += 4) {
                    const int kk = k4 + t4;                    // inner index of this lane
                    const bool k_ok = kk < n;
                    const double av = (arow_ok && k_ok) ? arow[k4] : 0.0;
                    const double* brow = src + (size_t)(k_ok ? kk : 0) * ldp;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (c < nct) {
                            const int col = (ct0 + c) * 8 + g;
                            const double bv = (k_ok && col < n) ? brow[col] : 0.0;
                            dmma884(acc[c][0], acc[c][1], av, bv);
                        }
                    }
                }
                if (arow_ok) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int col = (ct0 + c) * 8 + 2 * t4;
                        if (c < nct && col < ldp)
                            *reinterpret_cast<double2*>(dst + (size_t)arow_i * ldp + col) =
                                make_double2(acc[c][0], col + 1 < n ? acc[c][1] : 0.0);
                    }
                }
            }
            __syncthreads();          // global writes of this CTA are visible to all its threads
            double* tmp = src;
            src = dst;
            dst = tmp;
        }
        // floor, renormalise, NaN check (AR:153-159); the result goes to X
        int nanflag = 0;
        for (int i = warp; i < n; i += SQ_NW) {
            const double* xr = src + (size_t)i * ldp;
            double rs = 0.0;
            for (int j = 2 * lane; j < ldp; j += 64) {
                const double2 x = *reinterpret_cast<const double2*>(xr + j);
                const double a0 = (x.x != x.x) ? x.x : fmax(x.x, 1e-12);
                const double a1 = (j + 1 < n) ? ((x.y != x.y) ? x.y : fmax(x.y, 1e-12)) : 0.0;
                rs += a0 + a1;
            }
            rs = warp_sum(rs);
            for (int j = 2 * lane; j < ldp; j += 64) {
                const double2 x = *reinterpret_cast<const double2*>(xr + j);
                double2 o;
                o.x = ((x.x != x.x) ? x.x : fmax(x.x, 1e-12)) / rs;
                o.y = (j + 1 < n) ? ((x.y != x.y) ? x.y : fmax(x.y, 1e-12)) / rs : 0.0;
                if (o.x != o.x || o.y != o.y) nanflag = 1;
                *reinterpret_cast<double2*>(Xm + (size_t)i * ldp + j) = o;
            }
        }
        nanflag = __syncthreads_or(nanflag);
        if (nanflag) {
            for (int i = warp; i < n; i += SQ_NW)
                for (int j = lane; j < ldp; j += 32) Xm[(size_t)i * ldp + j] = (i == j) ? 1.0 : 0.0;
            if (tid == 0 && nan_count != nullptr) atomicAdd(nan_count, 1);
        }
        __syncthreads();
    }
}

template <int NW>
static int launch_leaf_lowp(const ApplyArgs& src, int B, cudaStream_t st) {
    LowpArgs a;
    a.tc = src.tc; a.norms = src.norms; a.items = src.items; a.leaf_code = src.leaf_code;
    a.w = src.w; a.wr = src.wr; a.mode = src.mode;
    a.n_items = src.n_items; a.n = src.n; a.ldp = src.ldp; a.n_nodes = src.n_nodes; a.E = src.E;
    a.plan_nc = src.plan_nc; a.plan_tau = src.plan_tau; a.plan_a0 = src.plan_a0; a.branch_of = src.branch_of;
    constexpr int EB = NW * 16;
    const int gx = (a.n_items + EB - 1) / EB;
    const long long want = 4LL * g_f_sm;
    int ry = 1;
    if ((long long)gx * B < want) {
        ry = (int)((want + (long long)gx * B - 1) / ((long long)gx * B));
        const int ry_max = (a.n + 7) / 8;
        if (ry > ry_max) ry = ry_max;
    }
    const int rows_per_cta = (a.n + ry - 1) / ry;
    ry = (a.n + rows_per_cta - 1) / rows_per_cta;
    const size_t fixed = lowp_fixed_bytes(NW, a.n);
    const size_t one_stage = sizeof(float) * (size_t)L_ROW;
    const size_t one_kb = sizeof(float) * (size_t)F_JT * L_TILE;
    const size_t budget = g_f_optin / 2 - 2048;               // two CTAs per SM (1 KB reserved each)
    if (fixed + 2 * one_kb > budget) return 1;
    size_t ring = (budget - fixed) / one_kb * one_kb;
    if (ring > LP_MAXSTAGE * one_stage) ring = LP_MAXSTAGE * one_stage;
    if (ring > (size_t)tuning().ring_lowp * one_kb) ring = (size_t)tuning().ring_lowp * one_kb;
    if (ring < 2 * one_kb) ring = 2 * one_kb;
    a.ring_floats = (int)(ring / sizeof(float));
    const size_t smem = fixed + ring;
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        CEVB_CHECK_CUDA(cudaFuncSetAttribute(leaf_lowp_kernel<NW>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem = smem;
    }
    dim3 grid(gx, ry, B);
    leaf_lowp_kernel<NW><<<grid, 32 * (NW + 1), smem, st>>>(a, rows_per_cta);
    CEVB_CHECK_LAUNCH("leaf_lowp_kernel");
    return 0;
}

struct PlanRef {
    const int32_t* nc;
    const double* tau;
    const double* a0;
    const int32_t* branch_of;
    int E;
};

static int launch_apply(const double* tc, const double* norms, int B, int n, int n_nodes,
                        const int32_t* items, int n_items, int leaf_only, const PlanRef& plan,
                        const int32_t* leaf_code, const double* node_vec, double* w, double* wr,
                        uint8_t* mode, cudaStream_t st) {
    if (n_items <= 0 || B <= 0) return 0;
    if (fused_query() != 0) return -2;
    if (n > CEVB_MAX_N) {
        set_error("taylor_apply: n=%d exceeds CEVB_MAX_N=%d", n, CEVB_MAX_N);
        return -1;
    }
    ApplyArgs a;
    a.tc = tc; a.norms = norms; a.items = items; a.t_node = nullptr; a.leaf_code = leaf_code;
    a.node_vec = node_vec; a.w = w; a.wr = wr; a.mode = mode;
    a.n_items = n_items; a.n = n; a.npad = round_up(n, 8); a.ldp = ldp_of(n); a.n_nodes = n_nodes;
    a.npass = taylor_npass(n);
    // child vectors cover every column the 13 * npass tiles of a row address (zeros beyond n);
    // a stride of 8 (mod 16) doubles keeps the epilogue's 16-byte loads conflict-free
    a.ldl = a.npass * F_JT * 8;
    if ((a.ldl % 16) == 0) a.ldl += 8;
    a.slot = nullptr; a.X = nullptr; a.E = plan.E;
    a.plan_nc = plan.nc; a.plan_tau = plan.tau; a.plan_a0 = plan.a0; a.branch_of = plan.branch_of;
    a.skip_generators = 0;
    if (leaf_only && tuning().lowp && a.npass == 1) {
        // generator vectors: column in FP64 + floor correction on the TF32 tensor cores; the
        // FP64 launch below then only works on the vectors whose Q is no generator
        const int rl = launch_leaf_lowp<7>(a, B, st);
        if (rl < 0) return rl;
        if (rl == 0) a.skip_generators = 1;
    }
    const int rc = leaf_only ? launch_apply_shapes<APPLY_LEAF>(a, B, st)
                             : launch_apply_shapes<APPLY_GENERAL>(a, B, st);
    if (rc == 1) {
        set_error("taylor_apply: n=%d does not fit in shared memory", n);
        return -1;
    }
    return rc;
}

}  // namespace cevb

using namespace cevb;

extern "C" size_t cevb_taylor_doubles(int n) { return taylor_vec_doubles(n); }

extern "C" int cevb_taylor_prepare(const double* pw, int B, int n, double t_max, double* norms,
                                   uint16_t* csc, uint8_t* csc_cnt, double* tc, void* stream) {
    if (B <= 0) return 0;
    if (B > 65535) {
        set_error("cevb_taylor_prepare: B=%d exceeds one launch; chunk the batch", B);
        return -1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int npad = round_up(n, 8);
    q_structure_kernel<<<B, 128, 0, st>>>(pw, n, ldq_of(n), norms, csc, csc_cnt, t_max);
    CEVB_CHECK_LAUNCH("q_structure_kernel");
    static bool inv_ready = false;
    if (!inv_ready) {
        double h[4 * CEVB_TAYLOR_KC + 1];
        h[0] = 0.0;
        for (int k = 1; k <= 4 * CEVB_TAYLOR_KC; ++k) h[k] = 1.0 / (double)k;
        CEVB_CHECK_CUDA(cudaMemcpyToSymbol(c_inv_int, h, sizeof(h)));
        inv_ready = true;
    }
    dim3 grid((n + TCR - 1) / TCR, B);
    const size_t csm = sizeof(double) * (2 * TCR * (size_t)npad + 4 * TCR * (size_t)(npad + 4)) +
                       sizeof(uint32_t) * 4 * TCR * (size_t)(F_JT * 8);
    static size_t coeff_attr = 0;
    if (csm > coeff_attr) {
        CEVB_CHECK_CUDA(cudaFuncSetAttribute(taylor_coeff_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csm));
        coeff_attr = csm;
    }
    taylor_coeff_kernel<<<grid, 128, csm, st>>>(
        pw, csc, csc_cnt, norms, n, ldq_of(n), npad, t_max, tuning().dense, tuning().lowp, tc);
    CEVB_CHECK_LAUNCH("taylor_coeff_kernel");
    return 0;
}

extern "C" int cevb_expm_plan(const double* norms, const double* t, int B, int E, long long cap,
                              int32_t* slot, int32_t* worklist, int32_t* work_s, unsigned int* count,
                              int32_t* plan_nc, double* plan_tau, double* plan_a0, void* stream) {
    if (B <= 0 || E <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (count != nullptr) CEVB_CHECK_CUDA(cudaMemsetAsync(count, 0, sizeof(unsigned int), st));
    const long long M = (long long)B * E;
    expm_plan_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(norms, t, E, M, cap, slot, worklist,
                                                                work_s, count, plan_nc, plan_tau,
                                                                plan_a0);
    CEVB_CHECK_LAUNCH("expm_plan_kernel");
    return 0;
}

extern "C" int cevb_taylor_matrices_supported(int n) {
    if (fused_query() != 0) return 0;
    return (n >= 2 && n <= CEVB_MAX_N) ? 1 : 0;
}

extern "C" size_t cevb_taylor_scratch_doubles(int n) {
    if (fused_query() != 0) return 0;
    if (square_smem_bytes(round_up(n, 8)) <= g_f_optin) return 0;   // both buffers in shared memory
    return (size_t)g_f_sm * n * ldp_of(n);                           // one matrix per persistent CTA
}

extern "C" int cevb_taylor_matrices(const double* tc, const double* norms, const int32_t* t_order,
                                    int B, int E, int n, const int32_t* slot,
                                    const int32_t* work_s, const unsigned int* count, long long cap,
                                    const int32_t* plan_nc, const double* plan_tau,
                                    const double* plan_a0, double* P, unsigned int* counter,
                                    int32_t* nan_count, double* scratch, void* stream) {
    if (B <= 0 || E <= 0 || cap <= 0) return 0;
    if (B > 65535) {
        set_error("cevb_taylor_matrices: B=%d exceeds one launch; chunk the batch", B);
        return -1;
    }
    if (fused_query() != 0) return -2;
    if (n > CEVB_MAX_N) {
        set_error("cevb_taylor_matrices: n=%d exceeds CEVB_MAX_N=%d", n, CEVB_MAX_N);
        return -1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    ApplyArgs a;
    a.tc = tc; a.norms = norms; a.items = t_order; a.t_node = nullptr; a.leaf_code = nullptr;
    a.node_vec = nullptr; a.w = nullptr; a.wr = nullptr; a.mode = nullptr;
    a.n_items = E; a.n = n; a.npad = round_up(n, 8); a.ldp = ldp_of(n); a.n_nodes = 0;
    a.ldl = 0;
    a.npass = taylor_npass(n);
    a.slot = slot; a.X = P; a.E = E;
    a.skip_generators = 0;
    a.plan_nc = plan_nc; a.plan_tau = plan_tau; a.plan_a0 = plan_a0; a.branch_of = nullptr;
    int rc = launch_apply_shapes<APPLY_EVAL>(a, B, st);
    if (rc == 1) {
        set_error("cevb_taylor_matrices: n=%d does not fit in shared memory", n);
        return -1;
    }
    if (rc != 0) return rc;
    const size_t smem = square_smem_bytes(a.npad);
    CEVB_CHECK_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned int), st));
    if (smem > g_f_optin) {
        if (scratch == nullptr) {
            set_error("cevb_taylor_matrices: n=%d needs cevb_taylor_scratch_doubles(n) doubles of scratch", n);
            return -1;
        }
        long long grid = g_f_sm;
        if (grid > cap) grid = cap;
        square_post_global_kernel<<<(unsigned)grid, SQ_NT, 0, st>>>(P, work_s, count, cap, counter, n,
                                                                    a.ldp, a.npad, scratch, nan_count);
        CEVB_CHECK_LAUNCH("square_post_global_kernel");
        return 0;
    }
    static size_t sq_attr = 0;
    if (smem > sq_attr) {
        CEVB_CHECK_CUDA(cudaFuncSetAttribute(square_post_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sq_attr = smem;
    }
    long long grid = g_f_sm;
    if (grid > cap) grid = cap;
    square_post_kernel<<<(unsigned)grid, SQ_NT, smem, st>>>(P, work_s, count, cap, counter, n, a.ldp,
                                                            a.npad, nan_count);
    CEVB_CHECK_LAUNCH("square_post_kernel");
    return 0;
}

extern "C" int cevb_taylor_apply(const double* tc, const double* norms, int B, int n, int n_nodes,
                                 const int32_t* items, int n_items, int leaf_only,
                                 const double* t_node, const int32_t* leaf_code,
                                 const double* node_vec, double* w, double* wr, uint8_t* mode,
                                 void* stream) {
    if (B > 65535) {
        set_error("cevb_taylor_apply: B=%d exceeds one launch; chunk the batch", B);
        return -1;
    }
    if (B <= 0 || n_items <= 0) return 0;
    // stand-alone call: the per-pair plan is indexed by node and lives in stream-ordered scratch
    cudaStream_t st = (cudaStream_t)stream;
    const size_t M = (size_t)B * n_nodes;
    int32_t* p_nc = nullptr;
    double* p_tau = nullptr;
    double* p_a0 = nullptr;
    CEVB_CHECK_CUDA(cudaMallocAsync((void**)&p_nc, M * sizeof(int32_t), st));
    CEVB_CHECK_CUDA(cudaMallocAsync((void**)&p_tau, M * sizeof(double), st));
    CEVB_CHECK_CUDA(cudaMallocAsync((void**)&p_a0, M * sizeof(double), st));
    int rc = cevb_expm_plan(norms, t_node, B, n_nodes, 0, nullptr, nullptr, nullptr, nullptr, p_nc,
                            p_tau, p_a0, stream);
    if (rc == 0) {
        const PlanRef plan = {p_nc, p_tau, p_a0, nullptr, n_nodes};
        rc = launch_apply(tc, norms, B, n, n_nodes, items, n_items, leaf_only, plan, leaf_code,
                          node_vec, w, wr, mode, st);
    }
    cudaFreeAsync(p_nc, st);
    cudaFreeAsync(p_tau, st);
    cudaFreeAsync(p_a0, st);
    return rc;
}

extern "C" int cevb_prune_fused(const double* tc, const double* norms, const double* P,
                                const int32_t* slot, int B, int E, int n, int n_nodes, int root,
                                const int32_t* level_nodes, const int32_t* level_off_host,
                                int n_levels, const int32_t* child_off, const int32_t* child_idx,
                                const int32_t* branch_of, const int32_t* leaf_code,
                                const int32_t* level_child, const int32_t* level_child_off_host,
                                const int32_t* leaf_child, int n_leaf_child,
                                const int32_t* plan_nc, const double* plan_tau,
                                const double* plan_a0, const double* root_freq, double* node_vec,
                                double* w, double* wr,
                                uint8_t* mode, double* logl, int32_t* rescales,
                                int level_begin, int level_end, int stages, void* stream) {
    if (B <= 0) return 0;
    if (B > 65535) {
        set_error("cevb_prune_fused: B=%d exceeds one launch; chunk the batch", B);
        return -1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int ldp = ldp_of(n);
    const PlanRef plan = {plan_nc, plan_tau, plan_a0, branch_of, E};
    if (stages & CEVB_FUSED_INIT) {
        CEVB_CHECK_CUDA(cudaMemsetAsync(rescales, 0, sizeof(int32_t) * (size_t)B, st));
        CEVB_CHECK_CUDA(cudaMemsetAsync(mode, 0, (size_t)B * n_nodes, st));
    }
    if (stages & CEVB_FUSED_LEAFVEC) {
        // conditional-likelihood vectors of the leaves: outputs only (the kernels build the
        // one-hot / uniform child vectors from leaf_code), so callers that want only logL skip it
        dim3 lgrid(n_nodes, B);
        leaf_init_fused_kernel<<<lgrid, 128, 0, st>>>(n, ldp, n_nodes, leaf_code, node_vec);
        CEVB_CHECK_LAUNCH("leaf_init_kernel");
    }
    if (stages & CEVB_FUSED_LEAVES) {
        // branches above observed leaves depend on nothing below them: all levels in one launch
        const int rc = launch_apply(tc, norms, B, n, n_nodes, leaf_child, n_leaf_child, 1, plan,
                                    leaf_code, node_vec, w, wr, mode, st);
        if (rc != 0) return rc;
    }
    const size_t csmem = 2 * (size_t)n * sizeof(double);
    if (level_begin < 0) level_begin = 0;
    if (level_end > n_levels) level_end = n_levels;
    for (int lv = level_begin; lv < level_end; ++lv) {
        const int cnt = level_off_host[lv + 1] - level_off_host[lv];
        if (cnt <= 0) continue;
        if (stages & CEVB_FUSED_APPLY) {
            const int c0 = level_child_off_host[lv], c1 = level_child_off_host[lv + 1];
            const int rc = launch_apply(tc, norms, B, n, n_nodes, level_child + c0, c1 - c0, 0, plan,
                                        leaf_code, node_vec, w, wr, mode, st);
            if (rc != 0) return rc;
        }
        // wide levels: one warp per (node, vector) -- a quarter of the CTAs and no block
        // barriers; narrow levels keep the block kernel, whose four warps share the (rare but
        // long) mat-vec of a materialised P
        if ((stages & CEVB_FUSED_COMBINE) && n <= 128 && cnt >= 30) {
            const long long tasks = (long long)cnt * B;
            combine_level_warp_kernel<<<(unsigned)((tasks + 3) / 4), 128, 0, st>>>(
                w, wr, mode, P, slot, E, n, ldp, n_nodes, root, cnt, B,
                level_nodes + level_off_host[lv], child_off, child_idx, branch_of, leaf_code,
                root_freq, node_vec, logl, rescales);
            CEVB_CHECK_LAUNCH("combine_level_warp_kernel");
        } else if (stages & CEVB_FUSED_COMBINE) {
            dim3 grid(cnt, B);
            combine_level_kernel<<<grid, CT, csmem, st>>>(
                w, wr, mode, P, slot, E, n, ldp, n_nodes, root, level_nodes + level_off_host[lv],
                child_off, child_idx, branch_of, leaf_code, root_freq, node_vec, logl, rescales);
            CEVB_CHECK_LAUNCH("combine_level_kernel");
        }
    }
    return 0;
}
