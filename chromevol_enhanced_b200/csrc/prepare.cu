// Rate-matrix assembly and the per-parameter-vector tables the expm kernel consumes.
//
//   cevb_build_q        reference: ChromEvolutionModel.build_transition_rate_matrix, AR:75-134
//   cevb_expm_prepare   the quantities scipy.linalg.expm (AR:151) derives from one matrix:
//                       even powers and the 1-norms of Al-Mohy & Higham 2009, Alg. 5.1.
//                       All branches of one parameter vector share Q, so these are computed
//                       once per vector instead of once per branch.
#include "common.cuh"

namespace cevb {

// ---------------------------------------------------------------------------------------
// numpy's float64 add.reduce order over a contiguous row (pairwise summation with an 8-way
// unrolled base case); AR:131 takes `-np.sum(Q[i, :])` so the diagonal is bit-identical only
// if the additions associate the same way.  __dadd_rn keeps the compiler from re-associating
// or fusing.
__device__ double np_pairwise_sum(const double* a, int n) {
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; ++i) r = __dadd_rn(r, a[i]);
        return r;
    }
    if (n <= 128) {
        double r[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = a[k];
        int i = 8;
        const int full = n - (n % 8);
        for (; i < full; i += 8) {
#pragma unroll
            for (int k = 0; k < 8; ++k) r[k] = __dadd_rn(r[k], a[i + k]);
        }
        double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                               __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
        for (; i < n; ++i) res = __dadd_rn(res, a[i]);
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __dadd_rn(np_pairwise_sum(a, n2), np_pairwise_sum(a + n2, n - n2));
}

// One CTA per parameter vector.  Each row is written by exactly one thread with the same
// sequence of `=` and `+=` the Python loop performs (AR:86-127), every product and sum rounded
// separately (no FMA contraction), so Q matches the reference to the last bit.
__global__ void __launch_bounds__(128) build_q_kernel(const double* __restrict__ params, int n,
                                                      int ldq, double* __restrict__ pw) {
    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const int N = n - 1;  // max_chr
    double* Q = pw + (size_t)b * CEVB_NPOW * n * ldq;
    const double* p = params + (size_t)b * CEVB_NPARAM;
    const double gain = p[0], loss = p[1], dupl = p[2], dgain = p[3], dloss = p[4];
    const double bngain = p[5], bnloss = p[6], lin = p[7];
    const int bn = (p[8] > 0.0) ? (int)p[8] : 0;  // 0 encodes base_number=None (AR:120)

    for (int i = tid; i < n * ldq; i += blockDim.x) Q[i] = 0.0;
    __syncthreads();

    for (int i = 1 + tid; i <= N; i += blockDim.x) {
        double* row = Q + (size_t)i * ldq;
        const double f = __dadd_rn(1.0, __dmul_rn(lin, (double)i));
        if (i < N) row[i + 1] = __dmul_rn(gain, f);                                  // AR:88-90
        if (i > 1) row[i - 1] = __dmul_rn(loss, f);                                  // AR:93-95
        if (2 * i <= N) row[2 * i] = __dadd_rn(row[2 * i], dupl);                    // AR:98-99
        if (dgain > 0.0) {                                                           // AR:107-111
            const double h = __dmul_rn(dgain, 0.5);
            if (i + 1 <= N) row[i + 1] = __dadd_rn(row[i + 1], h);
            if (i + 2 <= N) row[i + 2] = __dadd_rn(row[i + 2], h);
        }
        if (dloss > 0.0) {                                                           // AR:113-117
            const double h = __dmul_rn(dloss, 0.5);
            if (i - 1 >= 1) row[i - 1] = __dadd_rn(row[i - 1], h);
            if (i - 2 >= 1) row[i - 2] = __dadd_rn(row[i - 2], h);
        }
        if (bn > 0) {                                                                // AR:120-127
            if (i + bn <= N) row[i + bn] = __dadd_rn(row[i + bn], bngain);
            if (i - bn >= 1) row[i - bn] = __dadd_rn(row[i - bn], bnloss);
        }
    }
    __syncthreads();
    for (int i = tid; i <= N; i += blockDim.x) {                                     // AR:130-131
        double* row = Q + (size_t)i * ldq;
        row[i] = -np_pairwise_sum(row, n);
    }
}

// ---------------------------------------------------------------------------------------
// Batched C = A * B over the power slots of pw, FP64 tensor cores (DMMA.8x8x4).
// CTA tile 32x32, four warps as 2x2, each warp 16x16 = 2x2 MMA tiles, K staged 16 at a time.
struct GemmOp {
    int a, b, c;
};
struct GemmOps {
    GemmOp op[3];
    int nops;
};

__global__ void __launch_bounds__(128) pow_gemm_kernel(double* __restrict__ pw, int n, int ldq,
                                                       GemmOps ops) {
    __shared__ double As[32][20];  // stride 20 = 4 (mod 16): conflict-free A fragments
    __shared__ double Bs[16][36];  // stride 36 = 4 (mod 16): conflict-free B fragments
    const int z = blockIdx.z;
    const int b = z / ops.nops;
    const GemmOp op = ops.op[z % ops.nops];
    const size_t mat = (size_t)n * ldq;
    const double* A = pw + ((size_t)b * CEVB_NPOW + op.a) * mat;
    const double* Bm = pw + ((size_t)b * CEVB_NPOW + op.b) * mat;
    double* C = pw + ((size_t)b * CEVB_NPOW + op.c) * mat;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int row0 = blockIdx.y * 32, col0 = blockIdx.x * 32;
    const int wr = (warp >> 1) * 16, wc = (warp & 1) * 16;
    double acc[2][2][2] = {};

    for (int k0 = 0; k0 < n; k0 += 16) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int idx = tid + q * 128;
            const int ar = idx >> 4, ac = idx & 15;
            const int gr = row0 + ar, gc = k0 + ac;
            As[ar][ac] = (gr < n && gc < n) ? A[(size_t)gr * ldq + gc] : 0.0;
            const int br = idx >> 5, bc = idx & 31;
            const int hr = k0 + br, hc = col0 + bc;
            Bs[br][bc] = (hr < n && hc < n) ? Bm[(size_t)hr * ldq + hc] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; kk += 4) {
            const double a0 = As[wr + g][kk + t], a1 = As[wr + 8 + g][kk + t];
            const double b0 = Bs[kk + t][wc + g], b1 = Bs[kk + t][wc + 8 + g];
            dmma884(acc[0][0][0], acc[0][0][1], a0, b0);
            dmma884(acc[0][1][0], acc[0][1][1], a0, b1);
            dmma884(acc[1][0][0], acc[1][0][1], a1, b0);
            dmma884(acc[1][1][0], acc[1][1][1], a1, b1);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int r = row0 + wr + i * 8 + g;
            const int c = col0 + wc + j * 8 + 2 * t;
            if (r < n) {
                if (c < n) C[(size_t)r * ldq + c] = acc[i][j][0];
                if (c + 1 < n) C[(size_t)r * ldq + c + 1] = acc[i][j][1];
            }
        }
}

// ---------------------------------------------------------------------------------------
// Norms and sparsity, grid (B, 7):
//   y = 0..4  ||M||_1^(1/k) for M = Q, Q^4, Q^6, Q^8, Q^10 (k = 1, 4, 6, 8, 10)
//   y = 5     || |Q|^p ||_1 for p = 7, 11, 15, 19, 27 (the `ell` bound of Alg. 5.1)
//   y = 6     compressed-column structure of Q (row indices of the non-zeros per column)
__device__ double block_max_128(double v, double* red) {
    v = warp_max(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = fmax(fmax(red[0], red[1]), fmax(red[2], red[3]));
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(128) norms_kernel(const double* __restrict__ pw, int n, int ldq,
                                                    double* __restrict__ norms,
                                                    uint16_t* __restrict__ csc,
                                                    uint8_t* __restrict__ csc_cnt) {
    extern __shared__ double sh[];  // 2 * n doubles for y == 5
    __shared__ double red[4];
    __shared__ int dense_flag;
    const int b = blockIdx.x, which = blockIdx.y, tid = threadIdx.x;
    const size_t mat = (size_t)n * ldq;
    const double* base = pw + (size_t)b * CEVB_NPOW * mat;
    double* out = norms + (size_t)b * CEVB_NNORM;

    if (which < 5) {
        const int slot = (which == 0) ? 0 : which + 1;  // Q, Q^4 (slot 2), Q^6, Q^8, Q^10
        const double* M = base + (size_t)slot * mat;
        double best = 0.0;
        bool bad = false;
        for (int j = tid; j < n; j += blockDim.x) {
            double s = 0.0;
            for (int i = 0; i < n; ++i) s += fabs(M[(size_t)i * ldq + j]);
            if (s != s) bad = true;
            best = fmax(best, s);
        }
        double m = block_max_128(best, red);
        const int anybad = __syncthreads_or(bad ? 1 : 0);
        if (tid == 0) {
            if (anybad) m = nan("");  // np.max propagates NaN; fmax would hide it
            const double k = (which == 0) ? 1.0 : (double)(2 * which + 2);
            out[which] = (which == 0) ? m : pow(m, 1.0 / k);
        }
    } else if (which == 5) {
        const double* Q = base;
        double* v = sh;
        double* w = sh + n;
        for (int j = tid; j < n; j += blockDim.x) v[j] = 1.0;
        __syncthreads();
        int slot = 0;
        for (int p = 1; p <= 27; ++p) {
            for (int j = tid; j < n; j += blockDim.x) {
                double s = 0.0;
                for (int i = 0; i < n; ++i) s = fma(fabs(Q[(size_t)i * ldq + j]), v[i], s);
                w[j] = s;
            }
            __syncthreads();
            double* tmp = v;
            v = w;
            w = tmp;
            if (p == 7 || p == 11 || p == 15 || p == 19 || p == 27) {
                double best = 0.0;
                bool bad = false;
                for (int j = tid; j < n; j += blockDim.x) {
                    if (v[j] != v[j]) bad = true;
                    best = fmax(best, v[j]);
                }
                double m = block_max_128(best, red);
                const int anybad = __syncthreads_or(bad ? 1 : 0);
                if (tid == 0) out[CEVB_NORM_ABS7 + slot] = anybad ? nan("") : m;
                ++slot;
            }
        }
    } else {
        // sparsity of Q: slot 0 = compressed columns (row indices of each column's non-zeros),
        // slot 1 = compressed rows (column indices of each row's non-zeros); both [c][n] layouts
        const double* Q = base;
        if (tid == 0) dense_flag = 0;
        __syncthreads();
        uint16_t* idx_c = csc + (size_t)b * 2 * CEVB_MAX_SPARSE * n;
        uint16_t* idx_r = idx_c + (size_t)CEVB_MAX_SPARSE * n;
        uint8_t* cnt_c = csc_cnt + (size_t)b * 2 * n;
        uint8_t* cnt_r = cnt_c + n;
        for (int j = tid; j < n; j += blockDim.x) {
            int cc = 0, cr = 0;
            for (int i = 0; i < n; ++i) {
                if (Q[(size_t)i * ldq + j] != 0.0) {  // NaN != 0 is true: kept as "non-zero"
                    if (cc < CEVB_MAX_SPARSE) idx_c[(size_t)cc * n + j] = (uint16_t)i;
                    ++cc;
                }
                if (Q[(size_t)j * ldq + i] != 0.0) {
                    if (cr < CEVB_MAX_SPARSE) idx_r[(size_t)cr * n + j] = (uint16_t)i;
                    ++cr;
                }
            }
            cnt_c[j] = (cc > CEVB_MAX_SPARSE) ? 255 : (uint8_t)cc;
            cnt_r[j] = (cr > CEVB_MAX_SPARSE) ? 255 : (uint8_t)cr;
            if (cc > CEVB_MAX_SPARSE || cr > CEVB_MAX_SPARSE) dense_flag = 1;
        }
        __syncthreads();
        if (tid == 0) out[CEVB_NORM_DENSE] = dense_flag ? 1.0 : 0.0;
    }
}

}  // namespace cevb

using namespace cevb;

extern "C" int cevb_build_q(const double* params, int B, int n, double* pw, void* stream) {
    if (B <= 0) return 0;
    if (n < 2 || n > 4096) {
        set_error("cevb_build_q: unsupported n=%d", n);
        return -1;
    }
    build_q_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(params, n, ldq_of(n), pw);
    CEVB_CHECK_LAUNCH("build_q_kernel");
    return 0;
}

extern "C" int cevb_expm_prepare(double* pw, int B, int n, double* norms, uint16_t* csc,
                                 uint8_t* csc_cnt, void* stream) {
    if (B <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int ldq = ldq_of(n);
    const int tiles = (n + 31) / 32;
    // slot k holds Q^(2k): Q2 = Q.Q ; Q4 = Q2.Q2 ; Q6 = Q4.Q2 ; then Q8 = Q6.Q2, Q10 = Q4.Q6,
    // Q12 = Q6.Q6 together (the same products scipy's helper forms: A6=A4.A2, A8=A6.A2, A4.A6)
    const GemmOps stages[4] = {
        {{{0, 0, 1}, {0, 0, 0}, {0, 0, 0}}, 1},
        {{{1, 1, 2}, {0, 0, 0}, {0, 0, 0}}, 1},
        {{{2, 1, 3}, {0, 0, 0}, {0, 0, 0}}, 1},
        {{{3, 1, 4}, {2, 3, 5}, {3, 3, 6}}, 3},
    };
    for (int s = 0; s < 4; ++s) {
        const long long gz = (long long)B * stages[s].nops;
        if (gz > 65535) {
            set_error("cevb_expm_prepare: batch too large for one launch (B=%d); chunk it", B);
            return -1;
        }
        dim3 grid(tiles, tiles, (unsigned)gz);
        pow_gemm_kernel<<<grid, 128, 0, st>>>(pw, n, ldq, stages[s]);
        CEVB_CHECK_LAUNCH("pow_gemm_kernel");
    }
    dim3 ngrid(B, 7);
    norms_kernel<<<ngrid, 128, 2 * n * sizeof(double), st>>>(pw, n, ldq, norms, csc, csc_cnt);
    CEVB_CHECK_LAUNCH("norms_kernel");
    return 0;
}
