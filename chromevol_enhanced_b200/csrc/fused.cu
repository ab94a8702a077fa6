// Fused transition-probability + pruning path: P(t) is never written to HBM.
//
// Reference behaviour reproduced (per branch, per parameter vector):
//     P  = scipy.linalg.expm(Q t)                       AR:151
//     P  = max(P, 1e-12);  P /= rowsum(P);  NaN -> I    AR:153-159   (t == 0 -> I, AR:146-148)
//     w  = P @ L_child                                  AR:237
// followed, per parent node, by the product over children in file order with the reference's
// conditional rescale (AR:238-243) and the root log-likelihood (AR:251-272).
//
// Every branch of one parameter vector shares Q (the CLI never sets branch_params), so
//     expm(Q t) = sum_k t^k C_k,   C_k = Q^k / k!,
// is a polynomial in the SCALAR t with matrix coefficients that are formed once per vector
// (taylor_coeff_kernel; Q has <= 9 non-zeros per column, so each C_k costs O(9 n^2)).  For
// eta = t ||Q||_1 <= 5.4 (the same bound scipy's Pade-13 uses unscaled) the series is summed
// directly: evaluating it for many branches at once is the dense contraction
//     P[e, (i,j)] = sum_k T[e,k] C[k,(i,j)],   T[e,k] = t_e^k,
// which taylor_apply_kernel runs on the FP64 tensor cores (DMMA.8x8x4: M = 8 branches, N = 8
// columns j of one row i, K = 4 series terms) with the coefficient rows staged in shared memory
// by bulk async copies (cp.async.bulk + mbarrier, double buffered).  The floor, the row sum and
// the dot product with L_child are applied to the accumulator fragments in registers, so only
// w (n doubles per branch) leaves the SM.  The number of terms is chosen per branch from eta so
// that the truncation bound eta^(K+1)/(K+1)! / (1 - eta/(K+2)) is below 1e-18.
// Branches with eta > 5.4 need squarings and therefore a materialised matrix: they are routed
// (cevb_expm_plan) to the Pade kernel of expm.cu, whose P the combine kernel reads.
#include "common.cuh"

namespace cevb {

constexpr int F_KC = CEVB_TAYLOR_KC;       // chunks of 4 series coefficients (degrees 0..39)
constexpr int F_JT = 13;                   // 8-column tiles per pass (104 columns)
constexpr double F_THETA = CEVB_TAYLOR_THETA;
constexpr int F_MAXW = 8;                  // warps per CTA of the apply kernel

// number of series coefficients (degree + 1) needed for eta = t ||Q||_1
__host__ __device__ inline int taylor_ncoef(double eta) {
    if (!(eta > 0.0)) return 1;
    double term = eta;  // eta^k / k!
    int k = 1;
    for (; k < 4 * F_KC - 1; ++k) {
        term *= eta / (double)(k + 1);  // eta^(k+1) / (k+1)!
        if ((double)(k + 2) > eta && term / (1.0 - eta / (double)(k + 2)) <= 1e-18) break;
    }
    return k + 1;
}

// ---------------------------------------------------------------------------------------------
// C_0 = I, C_k = C_{k-1} Q / k, one CTA per (row i, vector b); row i of C_k only needs row i of
// C_{k-1}.  Output layout tc[b][i][c][j][4]: chunk c holds degrees 4c..4c+3 interleaved per
// column, exactly the B-operand fragment order of DMMA (lane 4g+t <- column g, degree t).
__global__ void __launch_bounds__(128) taylor_coeff_kernel(const double* __restrict__ pw,
                                                           const uint16_t* __restrict__ csc,
                                                           const uint8_t* __restrict__ csc_cnt,
                                                           int n, int ldq, int npad,
                                                           double* __restrict__ tc) {
    extern __shared__ double shc[];
    double* cur = shc;
    double* nxt = shc + npad;
    double* stg = shc + 2 * npad;  // [npad][4]
    const int i = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const double* Q = pw + (size_t)b * CEVB_NPOW * n * ldq;
    const uint16_t* idx_c = csc + (size_t)b * 2 * CEVB_MAX_SPARSE * n;
    const uint8_t* cnt_c = csc_cnt + (size_t)b * 2 * n;
    double* out = tc + ((size_t)b * n + i) * F_KC * npad * 4;
    for (int j = tid; j < npad; j += 128) cur[j] = (j == i) ? 1.0 : 0.0;
    __syncthreads();
    for (int k = 0; k < 4 * F_KC; ++k) {
        for (int j = tid; j < npad; j += 128) stg[j * 4 + (k & 3)] = (j < n) ? cur[j] : 0.0;
        if (k + 1 < 4 * F_KC) {
            const double inv_den = (double)(k + 1);
            for (int j = tid; j < n; j += 128) {
                const int cnt = cnt_c[j];
                double acc = 0.0;
                if (cnt == 255) {
                    for (int r = 0; r < n; ++r) acc = fma(cur[r], Q[(size_t)r * ldq + j], acc);
                } else {
                    for (int q = 0; q < cnt; ++q) {
                        const int r = idx_c[(size_t)q * n + j];
                        acc = fma(cur[r], Q[(size_t)r * ldq + j], acc);
                    }
                }
                nxt[j] = acc / inv_den;
            }
        }
        __syncthreads();
        if ((k & 3) == 3) {
            double* dst = out + (size_t)(k >> 2) * npad * 4;
            for (int x = tid; x < npad * 4; x += 128) dst[x] = stg[x];
            __syncthreads();
        }
        double* sw = cur;
        cur = nxt;
        nxt = sw;
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
                                                        unsigned int* __restrict__ count) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M) return;
    const int b = (int)(idx / E), e = (int)(idx - (long long)b * E);
    const double tb = t[e];
    const double eta = tb * norms[(size_t)b * CEVB_NNORM + CEVB_NORM_ONE];
    int s = -1;
    if (tb > 0.0 && !(eta <= F_THETA)) {
        const unsigned int q = atomicAdd(count, 1u);
        if ((long long)q < cap) {
            worklist[q] = (int32_t)idx;
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
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        " selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

struct ApplyArgs {
    const double* tc;
    const double* norms;
    const int32_t* items;     // child node of every branch of this launch
    const double* t_node;     // [n_nodes] effective branch length above a node
    const int32_t* leaf_code; // [n_nodes]
    const double* node_vec;   // [B][n_nodes][ldp]
    double* w;                // [B][n_nodes][ldp]
    uint8_t* mode;            // [B][n_nodes]: only non-zero values are ever written
    int n_items, n, npad, ldp, ldl, n_nodes, rows_per_cta;
};

__host__ __device__ inline size_t apply_header_bytes(int eb) {
    return (size_t)((16 + 4 * (3 * eb + 4) + 127) / 128) * 128;
}
__host__ __device__ inline size_t apply_smem_bytes(int nwarps, int G, int ldl) {
    const int eb = nwarps * G * 8;
    return apply_header_bytes(eb) +
           sizeof(double) * ((size_t)(eb / 8) * F_KC * 32 + (size_t)eb * ldl +
                             2 * (size_t)F_KC * F_JT * 32);
}

// One CTA: vector b = blockIdx.z, branches [blockIdx.x * EB, +EB) of the launch's item list
// (EB = warps * G * 8), rows [blockIdx.y * rows_per_cta, +rows_per_cta) of every P.  All warps
// work on the same row at a time and share its coefficient tile; a warp owns G groups of 8
// branches and keeps the G x 13 accumulator tiles of the row in registers.
template <int G>
__global__ void __launch_bounds__(32 * F_MAXW) taylor_apply_kernel(ApplyArgs a) {
    extern __shared__ __align__(128) unsigned char smraw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int nwarps = blockDim.x >> 5;
    const int EB = nwarps * G * 8;
    const int b = blockIdx.z;
    const int e0 = blockIdx.x * EB;
    const int n = a.n, npad = a.npad, ldl = a.ldl;
    const int i_begin = blockIdx.y * a.rows_per_cta;
    const int i_end = min(n, i_begin + a.rows_per_cta);

    uint64_t* full = reinterpret_cast<uint64_t*>(smraw);
    int* s_child = reinterpret_cast<int*>(smraw + 16);
    int* s_code = s_child + EB;
    int* s_nc = s_code + EB;
    double* As = reinterpret_cast<double*>(smraw + apply_header_bytes(EB));  // [EB/8][F_KC][32]
    double* Ls = As + (size_t)(EB / 8) * F_KC * 32;                           // [EB][ldl]
    double* Bs = Ls + (size_t)EB * ldl;                                       // [2][F_KC][F_JT*32]

    const double norm1 = a.norms[(size_t)b * CEVB_NNORM + CEVB_NORM_ONE];
    for (int x = tid; x < EB; x += blockDim.x) {
        const int e = e0 + x;
        int nc = 0, child = -1, code = -2;
        double tau = 0.0;
        if (e < a.n_items) {
            child = a.items[e];
            code = a.leaf_code[child];
            const double t = a.t_node[child];
            const double eta = t * norm1;
            if (!(t > 0.0)) {
                a.mode[(size_t)b * a.n_nodes + child] = 1;  // t == 0 -> identity (AR:146-148)
            } else if (!(eta <= F_THETA)) {
                a.mode[(size_t)b * a.n_nodes + child] = 2;  // materialised by the Pade kernel
            } else {
                nc = taylor_ncoef(eta);
                tau = t;
            }
        }
        s_child[x] = child;
        s_code[x] = code;
        s_nc[x] = nc;
        double* arow = As + (size_t)(x >> 3) * F_KC * 32 + (x & 7) * 4;
        double p = 1.0;
        for (int k = 0; k < 4 * F_KC; ++k) {
            arow[(k >> 2) * 32 + (k & 3)] = (k < nc) ? p : 0.0;
            p *= tau;
        }
    }
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // child vectors (one-hot / uniform / conditional-likelihood vector of an internal child)
    const double unif = 1.0 / (double)n;  // AR:211
    for (int x = warp; x < EB; x += nwarps) {
        const int child = s_child[x], code = s_code[x];
        const double* src =
            (child >= 0 && code == -2) ? a.node_vec + ((size_t)b * a.n_nodes + child) * a.ldp : nullptr;
        for (int j = lane; j < ldl; j += 32) {
            double v = 0.0;
            if (child >= 0 && j < n) v = (code >= 0) ? (j == code ? 1.0 : 0.0) : (code == -1 ? unif : src[j]);
            Ls[(size_t)x * ldl + j] = v;
        }
    }
    // chunks needed by this warp / by the CTA
    int wk = 0, ck = 0;
    for (int x = lane; x < EB; x += 32) {
        const int kc = (s_nc[x] + 3) >> 2;
        ck = max(ck, kc);
        if (x >= warp * G * 8 && x < (warp + 1) * G * 8) wk = max(wk, kc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        wk = max(wk, __shfl_xor_sync(0xffffffffu, wk, o));
        ck = max(ck, __shfl_xor_sync(0xffffffffu, ck, o));
    }
    __syncthreads();
    if (ck == 0) return;  // nothing on the fused path in this CTA (uniform across the CTA)

    const int nj = npad >> 3;
    const int npass = (nj + F_JT - 1) / F_JT;
    const int nitems = (i_end - i_begin) * npass;
    const double* tcb = a.tc + (size_t)b * n * F_KC * npad * 4;

    auto issue = [&](int q) {
        const int i = i_begin + q / npass, pass = q % npass;
        const int j0 = pass * F_JT * 8;
        const int ncols = min(F_JT * 8, npad - j0);
        const uint32_t bytes = (uint32_t)ncols * 4 * sizeof(double);
        uint64_t* bar = &full[q & 1];
        mbar_expect_tx(bar, bytes * (uint32_t)ck);
        double* dst = Bs + (size_t)(q & 1) * F_KC * F_JT * 32;
        const double* src = tcb + ((size_t)i * F_KC * npad + j0) * 4;
        for (int c = 0; c < ck; ++c)
            bulk_g2s(dst + (size_t)c * F_JT * 32, src + (size_t)c * npad * 4, bytes, bar);
    };
    if (tid == 0 && nitems > 0) issue(0);

    const long long floor_bits = __double_as_longlong(1e-12);
    double rs[G], dt[G];
    int bad[G];
#pragma unroll
    for (int gi = 0; gi < G; ++gi) {
        rs[gi] = 0.0;
        dt[gi] = 0.0;
        bad[gi] = 0;
    }

    for (int q = 0; q < nitems; ++q) {
        if (tid == 0 && q + 1 < nitems) issue(q + 1);
        const int i = i_begin + q / npass, pass = q % npass;
        const int j0 = pass * F_JT * 8;
        const int njt = min(F_JT, nj - pass * F_JT);
        if (wk > 0) {
            while (!mbar_try_wait(&full[q & 1], (uint32_t)((q >> 1) & 1))) {
            }
            double acc[G][F_JT][2];
#pragma unroll
            for (int gi = 0; gi < G; ++gi)
#pragma unroll
                for (int jt = 0; jt < F_JT; ++jt) {
                    acc[gi][jt][0] = 0.0;
                    acc[gi][jt][1] = 0.0;
                }
            const double* Bq = Bs + (size_t)(q & 1) * F_KC * F_JT * 32 + lane;
            const double* Aw = As + (size_t)(warp * G) * F_KC * 32 + lane;
            for (int c = 0; c < wk; ++c) {
                double av[G];
#pragma unroll
                for (int gi = 0; gi < G; ++gi) av[gi] = Aw[(size_t)(gi * F_KC + c) * 32];
                const double* Bc = Bq + (size_t)c * F_JT * 32;
#pragma unroll
                for (int jt = 0; jt < F_JT; ++jt) {
                    if (jt < njt) {
                        const double bv = Bc[jt * 32];
#pragma unroll
                        for (int gi = 0; gi < G; ++gi) dmma884(acc[gi][jt][0], acc[gi][jt][1], av[gi], bv);
                    }
                }
            }
            // floor, row sum and dot product with L_child on the accumulator fragments
#pragma unroll
            for (int gi = 0; gi < G; ++gi) {
                const double* Lrow = Ls + (size_t)((warp * G + gi) * 8 + g) * ldl + j0 + 2 * t4;
#pragma unroll
                for (int jt = 0; jt < F_JT; ++jt) {
                    if (jt < njt) {
                        const int j = j0 + jt * 8 + 2 * t4;
                        double p0 = acc[gi][jt][0], p1 = acc[gi][jt][1];
                        const long long b0 = __double_as_longlong(p0), b1 = __double_as_longlong(p1);
                        if (b0 < floor_bits) p0 = 1e-12;  // np.maximum(P, 1e-12), AR:153
                        if (b1 < floor_bits) p1 = 1e-12;
                        if (j >= n) p0 = 0.0;             // padding columns
                        if (j + 1 >= n) p1 = 0.0;
                        const int h0 = __double2hiint(p0), h1 = __double2hiint(p1);
                        bad[gi] |= ((h0 & 0x7ff00000) == 0x7ff00000) | ((h1 & 0x7ff00000) == 0x7ff00000);
                        const double2 l = *reinterpret_cast<const double2*>(Lrow + jt * 8);
                        rs[gi] += p0;
                        rs[gi] += p1;
                        dt[gi] = fma(p0, l.x, dt[gi]);
                        dt[gi] = fma(p1, l.y, dt[gi]);
                    }
                }
            }
            if (pass == npass - 1) {
#pragma unroll
                for (int gi = 0; gi < G; ++gi) {
                    double r = rs[gi], d = dt[gi];
                    int bd = bad[gi];
                    r += __shfl_xor_sync(0xffffffffu, r, 1);
                    d += __shfl_xor_sync(0xffffffffu, d, 1);
                    bd |= __shfl_xor_sync(0xffffffffu, bd, 1);
                    r += __shfl_xor_sync(0xffffffffu, r, 2);
                    d += __shfl_xor_sync(0xffffffffu, d, 2);
                    bd |= __shfl_xor_sync(0xffffffffu, bd, 2);
                    const int x = (warp * G + gi) * 8 + g;
                    if (t4 == 0 && s_nc[x] > 0) {
                        const int child = s_child[x];
                        a.w[((size_t)b * a.n_nodes + child) * a.ldp + i] = d / r;   // AR:154, AR:237
                        if (bd) a.mode[(size_t)b * a.n_nodes + child] = 1;          // NaN -> I, AR:156-159
                    }
                    rs[gi] = 0.0;
                    dt[gi] = 0.0;
                    bad[gi] = 0;
                }
            }
        }
        __syncthreads();
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
    const double* __restrict__ w, const uint8_t* __restrict__ mode, const double* __restrict__ P,
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
        const int md = mode[(size_t)b * n_nodes + child];
        const double* cv = node_vec + ((size_t)b * n_nodes + child) * ldp;
        if (md == 0) {
            const double* wv = w + ((size_t)b * n_nodes + child) * ldp;
            for (int i = tid; i < n; i += CT) v[i] *= wv[i];
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
                for (int i = warp * 2; i < n; i += 2 * (CT / 32)) {
                    const double* r0 = Pm + (size_t)i * ldp;
                    const bool two = (i + 1) < n;
                    const double* r1 = two ? r0 + ldp : r0;
                    double a0 = 0.0, a1 = 0.0;
                    for (int j = lane; j < n; j += 32) {
                        const double l = Lc[j];
                        a0 = fma(r0[j], l, a0);
                        a1 = fma(r1[j], l, a1);
                    }
                    a0 = warp_sum(a0);
                    a1 = warp_sum(a1);
                    if (lane == 0) {
                        v[i] *= a0;
                        if (two) v[i + 1] *= a1;
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

static int launch_apply(const double* tc, const double* norms, int B, int n, int n_nodes,
                        const int32_t* items, int n_items, const double* t_node,
                        const int32_t* leaf_code, const double* node_vec, double* w, uint8_t* mode,
                        cudaStream_t st) {
    if (n_items <= 0 || B <= 0) return 0;
    if (fused_query() != 0) return -2;
    ApplyArgs a;
    a.tc = tc; a.norms = norms; a.items = items; a.t_node = t_node; a.leaf_code = leaf_code;
    a.node_vec = node_vec; a.w = w; a.mode = mode;
    a.n_items = n_items; a.n = n; a.npad = round_up(n, 8); a.ldp = ldp_of(n); a.n_nodes = n_nodes;
    a.ldl = a.npad + ((a.npad % 16) == 0 ? 8 : 0);
    // groups per warp: 2 while the child vectors of 128 branches fit beside the staging buffers
    int G = (apply_smem_bytes(F_MAXW, 2, a.ldl) <= g_f_optin) ? 2 : 1;
    int nw = F_MAXW;
    while (nw > 1 && apply_smem_bytes(nw, G, a.ldl) > g_f_optin) nw >>= 1;
    if (apply_smem_bytes(nw, G, a.ldl) > g_f_optin) {
        set_error("taylor_apply: n=%d does not fit in shared memory", n);
        return -1;
    }
    const int need = (n_items + 8 * G - 1) / (8 * G);
    if (need < nw) nw = need;
    const int EB = nw * G * 8;
    const int gx = (n_items + EB - 1) / EB;
    const long long want = 4LL * g_f_sm;
    int ry = 1;
    if ((long long)gx * B < want) {
        ry = (int)((want + (long long)gx * B - 1) / ((long long)gx * B));
        const int ry_max = (n + 7) / 8;
        if (ry > ry_max) ry = ry_max;
    }
    a.rows_per_cta = (n + ry - 1) / ry;
    ry = (n + a.rows_per_cta - 1) / a.rows_per_cta;
    const size_t smem = apply_smem_bytes(nw, G, a.ldl);
    dim3 grid(gx, ry, B);
    if (G == 2) {
        CEVB_CHECK_CUDA(cudaFuncSetAttribute(taylor_apply_kernel<2>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        taylor_apply_kernel<2><<<grid, 32 * nw, smem, st>>>(a);
    } else {
        CEVB_CHECK_CUDA(cudaFuncSetAttribute(taylor_apply_kernel<1>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        taylor_apply_kernel<1><<<grid, 32 * nw, smem, st>>>(a);
    }
    CEVB_CHECK_LAUNCH("taylor_apply_kernel");
    return 0;
}

}  // namespace cevb

using namespace cevb;

extern "C" size_t cevb_taylor_doubles(int n) {
    return (size_t)n * F_KC * round_up(n, 8) * 4;
}

extern "C" int cevb_taylor_prepare(const double* pw, const uint16_t* csc, const uint8_t* csc_cnt,
                                   int B, int n, double* tc, void* stream) {
    if (B <= 0) return 0;
    if (B > 65535) {
        set_error("cevb_taylor_prepare: B=%d exceeds one launch; chunk the batch", B);
        return -1;
    }
    const int npad = round_up(n, 8);
    dim3 grid(n, B);
    taylor_coeff_kernel<<<grid, 128, sizeof(double) * 6 * (size_t)npad, (cudaStream_t)stream>>>(
        pw, csc, csc_cnt, n, ldq_of(n), npad, tc);
    CEVB_CHECK_LAUNCH("taylor_coeff_kernel");
    return 0;
}

extern "C" int cevb_expm_plan(const double* norms, const double* t, int B, int E, long long cap,
                              int32_t* slot, int32_t* worklist, unsigned int* count, void* stream) {
    if (B <= 0 || E <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    CEVB_CHECK_CUDA(cudaMemsetAsync(count, 0, sizeof(unsigned int), st));
    const long long M = (long long)B * E;
    expm_plan_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(norms, t, E, M, cap, slot, worklist,
                                                                count);
    CEVB_CHECK_LAUNCH("expm_plan_kernel");
    return 0;
}

extern "C" int cevb_taylor_apply(const double* tc, const double* norms, int B, int n, int n_nodes,
                                 const int32_t* items, int n_items, const double* t_node,
                                 const int32_t* leaf_code, const double* node_vec, double* w,
                                 uint8_t* mode, void* stream) {
    if (B > 65535) {
        set_error("cevb_taylor_apply: B=%d exceeds one launch; chunk the batch", B);
        return -1;
    }
    return launch_apply(tc, norms, B, n, n_nodes, items, n_items, t_node, leaf_code, node_vec, w,
                        mode, (cudaStream_t)stream);
}

extern "C" int cevb_prune_fused(const double* tc, const double* norms, const double* P,
                                const int32_t* slot, int B, int E, int n, int n_nodes, int root,
                                const int32_t* level_nodes, const int32_t* level_off_host,
                                int n_levels, const int32_t* child_off, const int32_t* child_idx,
                                const int32_t* branch_of, const int32_t* leaf_code,
                                const int32_t* level_child, const int32_t* level_child_off_host,
                                const double* t_node, const double* root_freq, double* node_vec,
                                double* w, uint8_t* mode, double* logl, int32_t* rescales,
                                int level_begin, int level_end, int stages, void* stream) {
    if (B <= 0) return 0;
    if (B > 65535) {
        set_error("cevb_prune_fused: B=%d exceeds one launch; chunk the batch", B);
        return -1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int ldp = ldp_of(n);
    if (stages & CEVB_FUSED_INIT) {
        CEVB_CHECK_CUDA(cudaMemsetAsync(rescales, 0, sizeof(int32_t) * (size_t)B, st));
        CEVB_CHECK_CUDA(cudaMemsetAsync(mode, 0, (size_t)B * n_nodes, st));
        dim3 lgrid(n_nodes, B);
        leaf_init_fused_kernel<<<lgrid, 128, 0, st>>>(n, ldp, n_nodes, leaf_code, node_vec);
        CEVB_CHECK_LAUNCH("leaf_init_kernel");
    }
    const size_t csmem = 2 * (size_t)n * sizeof(double);
    if (level_begin < 0) level_begin = 0;
    if (level_end > n_levels) level_end = n_levels;
    for (int lv = level_begin; lv < level_end; ++lv) {
        const int cnt = level_off_host[lv + 1] - level_off_host[lv];
        if (cnt <= 0) continue;
        if (stages & CEVB_FUSED_APPLY) {
            const int c0 = level_child_off_host[lv], c1 = level_child_off_host[lv + 1];
            const int rc = launch_apply(tc, norms, B, n, n_nodes, level_child + c0, c1 - c0, t_node,
                                        leaf_code, node_vec, w, mode, st);
            if (rc != 0) return rc;
        }
        if (stages & CEVB_FUSED_COMBINE) {
            dim3 grid(cnt, B);
            combine_level_kernel<<<grid, CT, csmem, st>>>(
                w, mode, P, slot, E, n, ldp, n_nodes, root, level_nodes + level_off_host[lv],
                child_off, child_idx, branch_of, leaf_code, root_freq, node_vec, logl, rescales);
            CEVB_CHECK_LAUNCH("combine_level_kernel");
        }
    }
    return 0;
}
