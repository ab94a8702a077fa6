// Level-synchronous Felsenstein pruning and argmax ancestral states.
//
//   cevb_prune          reference: ChromEvolLikelihoodCalculator.calculate_likelihood_at_node
//                       (AR:189-249) and calculate_total_log_likelihood (AR:251-272)
//   cevb_argmax_states  reference: ancestral_state_reconstruction (AR:298-320)
//
// One CTA per (parent node, parameter vector).  For every child, in file order (AR:221), the
// CTA streams that branch's P (row-major, rows of ldp doubles) once from HBM with coalesced
// loads, forms P @ L_child with warp-shuffle reductions, multiplies it into the running
// product, and applies the reference's conditional rescale (0 < sum < 1e-100, AR:241-243; the
// scale factor is dropped exactly like the reference drops it).  A child that is an observed
// leaf contributes column `code` of P, so only that column is read.
#include "common.cuh"

namespace cevb {

constexpr int PT = 128;  // threads per pruning CTA

__device__ double block_sum_128(double v, double* red) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    const double r = (red[0] + red[1]) + (red[2] + red[3]);
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(PT) leaf_init_kernel(int n, int ldp, int n_nodes,
                                                       const int32_t* __restrict__ leaf_code,
                                                       double* __restrict__ node_vec) {
    const int node = blockIdx.x, b = blockIdx.y;
    const int code = leaf_code[node];
    if (code < -1) return;  // internal node
    double* v = node_vec + ((size_t)b * n_nodes + node) * ldp;
    const double u = 1.0 / (double)n;  // AR:211
    for (int j = threadIdx.x; j < ldp; j += PT) {
        double x = 0.0;
        if (j < n) x = (code >= 0) ? (j == code ? 1.0 : 0.0) : u;
        v[j] = x;
    }
}

__global__ void __launch_bounds__(PT) prune_level_kernel(
    const double* __restrict__ P, int E, int n, int ldp, int n_nodes, int root,
    const int32_t* __restrict__ level_nodes, const int32_t* __restrict__ child_off,
    const int32_t* __restrict__ child_idx, const int32_t* __restrict__ branch_of,
    const int32_t* __restrict__ leaf_code, const double* __restrict__ root_freq,
    double* __restrict__ node_vec, double* __restrict__ logl, int32_t* __restrict__ rescales) {
    extern __shared__ double sh[];
    double* v = sh;        // running product for this parent
    double* Lc = sh + n;   // child vector
    __shared__ double red[4];
    const int node = level_nodes[blockIdx.x], b = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int j = tid; j < n; j += PT) v[j] = 1.0;  // AR:219
    int nres = 0;
    __syncthreads();

    const int c_begin = child_off[node], c_end = child_off[node + 1];
    for (int ci = c_begin; ci < c_end; ++ci) {
        const int child = child_idx[ci];
        const int code = leaf_code[child];
        const double* Pm = P + ((size_t)b * E + branch_of[child]) * (size_t)n * ldp;
        if (code >= 0) {
            // one-hot child: (P @ L)[i] = P[i][code]
            for (int i = tid; i < n; i += PT) v[i] *= Pm[(size_t)i * ldp + code];
        } else {
            const double* src = node_vec + ((size_t)b * n_nodes + child) * ldp;
            const double u = 1.0 / (double)n;
            for (int j = tid; j < n; j += PT) Lc[j] = (code == -1) ? u : src[j];
            __syncthreads();
            // rows two at a time per warp to keep more loads in flight
            for (int i = warp * 2; i < n; i += 2 * (PT / 32)) {
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
                    v[i] *= a0;                                   // AR:237-238
                    if (two) v[i + 1] *= a1;
                }
            }
        }
        __syncthreads();
        double part = 0.0;
        for (int j = tid; j < n; j += PT) part += v[j];
        const double s = block_sum_128(part, red);                // AR:241
        if (s > 0.0 && s < 1e-100) {                              // AR:242-243
            for (int j = tid; j < n; j += PT) v[j] = v[j] / s;
            ++nres;
        }
        __syncthreads();
    }

    double* out = node_vec + ((size_t)b * n_nodes + node) * ldp;
    for (int j = tid; j < ldp; j += PT) out[j] = (j < n) ? v[j] : 0.0;
    if (tid == 0 && nres) atomicAdd(&rescales[b], nres);
    if (node == root) {
        double part = 0.0;
        for (int j = tid; j < n; j += PT) part += v[j] * root_freq[j];   // AR:268
        const double L = block_sum_128(part, red);
        if (tid == 0) logl[b] = (L <= 0.0) ? -INFINITY : log(L);         // AR:270-272
    }
}

// best-of-two for np.argmax semantics: NaN counts as the maximum, first index wins ties.
__device__ __forceinline__ bool arg_better(double av, int ai, double bv, int bi) {
    const bool an = av != av, bn = bv != bv;
    if (an || bn) {
        if (an && bn) return ai < bi;
        return an;
    }
    if (av != bv) return av > bv;
    return ai < bi;
}

__global__ void __launch_bounds__(128) argmax_kernel(const double* __restrict__ node_vec,
                                                     long long n_vec, int n, int ldp,
                                                     int32_t* __restrict__ ml_count,
                                                     double* __restrict__ ml_prob,
                                                     double* __restrict__ probs) {
    const long long vec = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (vec >= n_vec) return;
    const double* v = node_vec + (size_t)vec * ldp;
    double part = 0.0;
    for (int j = lane; j < n; j += 32) part += v[j];
    const double s = warp_sum(part);
    const bool ok = s > 0.0;                                              // AR:312
    const double u = 1.0 / (double)n;
    double bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int j = lane; j < n; j += 32) {
        const double p = ok ? v[j] / s : u;
        if (probs) probs[(size_t)vec * ldp + j] = p;
        if (arg_better(p, j, bv, bi)) {
            bv = p;
            bi = j;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (arg_better(ov, oi, bv, bi)) {
            bv = ov;
            bi = oi;
        }
    }
    if (lane == 0) {
        ml_count[vec] = bi;                                               // AR:315
        ml_prob[vec] = bv;                                                // AR:316
    }
}

}  // namespace cevb

using namespace cevb;

extern "C" int cevb_prune(const double* P, int B, int E, int n, int n_nodes, int root,
                          const int32_t* level_nodes, const int32_t* level_off_host, int n_levels,
                          const int32_t* child_off, const int32_t* child_idx,
                          const int32_t* branch_of, const int32_t* leaf_code,
                          const double* root_freq, double* node_vec, double* logl,
                          int32_t* rescales, void* stream) {
    if (B <= 0) return 0;
    if (B > 65535) {
        set_error("cevb_prune: B=%d exceeds one launch; chunk the batch", B);
        return -1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int ldp = ldp_of(n);
    CEVB_CHECK_CUDA(cudaMemsetAsync(rescales, 0, sizeof(int32_t) * (size_t)B, st));
    dim3 lgrid(n_nodes, B);
    leaf_init_kernel<<<lgrid, PT, 0, st>>>(n, ldp, n_nodes, leaf_code, node_vec);
    CEVB_CHECK_LAUNCH("leaf_init_kernel");
    const size_t smem = 2 * (size_t)n * sizeof(double);
    for (int lv = 0; lv < n_levels; ++lv) {
        const int cnt = level_off_host[lv + 1] - level_off_host[lv];
        if (cnt <= 0) continue;
        dim3 grid(cnt, B);
        prune_level_kernel<<<grid, PT, smem, st>>>(P, E, n, ldp, n_nodes, root,
                                                   level_nodes + level_off_host[lv], child_off,
                                                   child_idx, branch_of, leaf_code, root_freq,
                                                   node_vec, logl, rescales);
        CEVB_CHECK_LAUNCH("prune_level_kernel");
    }
    return 0;
}

extern "C" int cevb_argmax_states(const double* node_vec, int B, int n_nodes, int n,
                                  int32_t* ml_count, double* ml_prob, double* probs,
                                  void* stream) {
    const long long n_vec = (long long)B * n_nodes;
    if (n_vec <= 0) return 0;
    const long long blocks = (n_vec + 3) / 4;
    argmax_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(
        node_vec, n_vec, n, ldp_of(n), ml_count, ml_prob, probs);
    CEVB_CHECK_LAUNCH("argmax_kernel");
    return 0;
}
