// Marginal (down-pass) and joint (Pupko) ancestral-state reconstruction over materialised P.
//
// ADDITIVE to the reference's API: /root/reference/src/ancestral_reconstruction.py has no
// down-pass and no joint backtrack (its ancestral_state_reconstruction, AR:274-322, says a true
// marginal reconstruction "requires a full upward and downward pass" and then takes the argmax
// of the up-pass vector, which cevb_argmax_states reproduces).  BASELINE.json's north star names
// "marginal and joint ancestral-state reconstruction" and "an argmax/backtrack kernel", so both
// are offered here; the CPU restatement they are tested against is
// oracle/reconstruction_oracle.py, itself pinned by exhaustive enumeration over the reference's
// transition matrices (AR:136-166).
//
//   cevb_marginal_down      posterior of every node given all the data:
//       down_root = prior;  down_c(j) = sum_i [down_p(i) prod_{s sibling of c} msg_s(i)] P_c(i,j)
//       post_c(j) = up_c(j) down_c(j) / sum,   msg_s(i) = sum_k P_s(i,k) up_s(k)  (AR:237)
//       up = the node vectors of the up-pass (cevb_prune / cevb_prune_fused); any per-node scale
//       of them (the reference's dropped rescale factor, AR:241-243) cancels in the posterior.
//   cevb_joint_reconstruct  the single most probable assignment of states to all nodes:
//       L_c(i) = max_j P_c(i,j) M_c(j),  C_c(i) = first j attaining it,
//       M_c = leaf vector (AR:201-214) or the product of the children's L in file order, scaled
//       by the power of two that brings its largest entry into [0.5, 1) (exact);
//       root state = first i maximising prior(i) M_root(i); then state_c = C_c(state_parent).
//       Every product is a single IEEE multiplication in a fixed order, so the comparisons --
//       and therefore the states -- are bit-identical to the numpy restatement on the same P.
//
// Nodes are numbered in level order (root = 0, parents before children), so the nodes of one
// depth are a contiguous index range and one launch per depth walks the tree downwards.
#include "common.cuh"

namespace cevb {

constexpr int RT = 128;  // threads per CTA

__device__ double block_sum_rt(double v, double* red) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    const double r = (red[0] + red[1]) + (red[2] + red[3]);
    __syncthreads();
    return r;
}

// child vector of `node` as the up-pass sees it: one-hot / uniform for leaves (AR:201-214),
// the stored conditional-likelihood vector otherwise
__device__ __forceinline__ double up_entry(const double* __restrict__ node_vec, int code, int j,
                                           double unif) {
    return (code >= 0) ? (j == code ? 1.0 : 0.0) : (code == -1 ? unif : node_vec[j]);
}

// One CTA per (node of this depth, vector).
__global__ void __launch_bounds__(RT) marginal_down_kernel(
    const double* __restrict__ P, int E, int n, int ldp, int n_nodes, int node0,
    const int32_t* __restrict__ parent, const int32_t* __restrict__ child_off,
    const int32_t* __restrict__ child_idx, const int32_t* __restrict__ branch_of,
    const int32_t* __restrict__ leaf_code, const double* __restrict__ root_freq,
    const double* __restrict__ node_vec, double* __restrict__ down, double* __restrict__ post) {
    extern __shared__ double sh[];
    double* tmp = sh;        // down_p * sibling messages
    double* Ls = sh + n;     // a sibling's up vector
    __shared__ double red[4];
    const int c = node0 + blockIdx.x, b = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const double unif = 1.0 / (double)n;
    const size_t vb = (size_t)b * n_nodes;
    double* dc = down + (vb + c) * ldp;
    const int p = parent[c];
    if (p < 0) {
        for (int j = tid; j < n; j += RT) dc[j] = root_freq[j];
    } else {
        const double* dp = down + (vb + p) * ldp;
        for (int i = tid; i < n; i += RT) tmp[i] = dp[i];
        __syncthreads();
        for (int ci = child_off[p]; ci < child_off[p + 1]; ++ci) {
            const int s = child_idx[ci];
            if (s == c) continue;
            const int code = leaf_code[s];
            const double* Pm = P + ((size_t)b * E + branch_of[s]) * (size_t)n * ldp;
            if (code >= 0) {
                for (int i = tid; i < n; i += RT) tmp[i] *= Pm[(size_t)i * ldp + code];
            } else {
                const double* src = node_vec + (vb + s) * ldp;
                for (int j = tid; j < n; j += RT) Ls[j] = (code == -1) ? unif : src[j];
                __syncthreads();
                for (int i = warp; i < n; i += RT / 32) {
                    const double* r = Pm + (size_t)i * ldp;
                    double a = 0.0;
                    for (int j = lane; j < n; j += 32) a = fma(r[j], Ls[j], a);
                    a = warp_sum(a);
                    if (lane == 0) tmp[i] *= a;
                }
            }
            __syncthreads();
        }
        double part = 0.0;
        for (int i = tid; i < n; i += RT) part += tmp[i];
        const double s = block_sum_rt(part, red);
        if (s > 0.0)
            for (int i = tid; i < n; i += RT) tmp[i] /= s;
        __syncthreads();
        // down_c(j) = sum_i tmp(i) P_c(i,j): thread j walks a column, coalesced across threads
        const double* Pc = P + ((size_t)b * E + branch_of[c]) * (size_t)n * ldp;
        part = 0.0;
        for (int j = tid; j < n; j += RT) {
            double a0 = 0.0, a1 = 0.0;
            int i = 0;
            for (; i + 1 < n; i += 2) {
                a0 = fma(tmp[i], Pc[(size_t)i * ldp + j], a0);
                a1 = fma(tmp[i + 1], Pc[(size_t)(i + 1) * ldp + j], a1);
            }
            if (i < n) a0 = fma(tmp[i], Pc[(size_t)i * ldp + j], a0);
            Ls[j] = a0 + a1;
            part += a0 + a1;
        }
        const double sd = block_sum_rt(part, red);
        for (int j = tid; j < n; j += RT) dc[j] = (sd > 0.0) ? Ls[j] / sd : Ls[j];
        __syncthreads();
    }
    // posterior
    const int code = leaf_code[c];
    const double* up = node_vec + (vb + c) * ldp;
    double part = 0.0;
    for (int j = tid; j < n; j += RT) {
        const double q = up_entry(up, code, j, unif) * dc[j];
        Ls[j] = q;
        part += q;
    }
    const double sq = block_sum_rt(part, red);
    double* out = post + (vb + c) * ldp;
    for (int j = tid; j < ldp; j += RT) out[j] = (j < n) ? (sq > 0.0 ? Ls[j] / sq : unif) : 0.0;
}

// first-index arg max (numpy semantics without NaN: values here are products of finite numbers)
__device__ __forceinline__ void best_of(double& bv, int& bi, double ov, int oi) {
    if (ov > bv || (ov == bv && oi < bi)) {
        bv = ov;
        bi = oi;
    }
}

// One CTA per (node, vector): M = leaf vector or scaled product of the children's L; then, for a
// non-root node, L(i) = max_j P(i,j) M(j) and C(i); for the root the arg max of prior * M.
__global__ void __launch_bounds__(RT) joint_up_kernel(
    const double* __restrict__ P, int E, int n, int ldp, int n_nodes, int root,
    const int32_t* __restrict__ nodes, const int32_t* __restrict__ child_off,
    const int32_t* __restrict__ child_idx, const int32_t* __restrict__ branch_of,
    const int32_t* __restrict__ leaf_code, const double* __restrict__ root_freq,
    double* __restrict__ lj, int32_t* __restrict__ cj, long long* __restrict__ expo,
    int32_t* __restrict__ states, double* __restrict__ joint_logp) {
    extern __shared__ double sh[];
    double* M = sh;
    __shared__ double red[4];
    __shared__ int redi[4];
    const int c = nodes ? nodes[blockIdx.x] : (int)blockIdx.x, b = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const size_t vb = (size_t)b * n_nodes;
    const int code = leaf_code[c];
    if (nodes == nullptr && code < -1) return;   // the leaf launch covers every node id
    const double unif = 1.0 / (double)n;
    long long e_sum = 0;
    if (code >= -1) {
        for (int j = tid; j < n; j += RT) M[j] = (code >= 0) ? (j == code ? 1.0 : 0.0) : unif;
    } else {
        const int c0 = child_off[c], c1 = child_off[c + 1];
        double mx = 0.0;
        for (int j = tid; j < n; j += RT) {
            double m = lj[(vb + child_idx[c0]) * ldp + j];
            for (int ci = c0 + 1; ci < c1; ++ci) m = m * lj[(vb + child_idx[ci]) * ldp + j];
            M[j] = m;
            mx = fmax(mx, m);
        }
        for (int ci = c0; ci < c1; ++ci) e_sum += expo[vb + child_idx[ci]];
        mx = warp_max(mx);
        if (lane == 0) red[warp] = mx;
        __syncthreads();
        mx = fmax(fmax(red[0], red[1]), fmax(red[2], red[3]));
        __syncthreads();
        if (mx > 0.0 && mx < INFINITY) {
            int e;
            frexp(mx, &e);
            for (int j = tid; j < n; j += RT) M[j] = ldexp(M[j], -e);
            e_sum += e;
        }
    }
    __syncthreads();
    if (c == root) {
        double bv = -1.0;
        int bi = 0x7fffffff;
        for (int i = tid; i < n; i += RT) best_of(bv, bi, root_freq[i] * M[i], i);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            best_of(bv, bi, ov, oi);
        }
        if (lane == 0) {
            red[warp] = bv;
            redi[warp] = bi;
        }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < RT / 32; ++w) best_of(bv, bi, red[w], redi[w]);
            states[vb + c] = bi;
            joint_logp[b] = (bv > 0.0 ? log(bv) : -INFINITY) + (double)e_sum * 0.69314718055994530942;
        }
        return;
    }
    if (tid == 0) expo[vb + c] = e_sum;
    const double* Pm = P + ((size_t)b * E + branch_of[c]) * (size_t)n * ldp;
    for (int i = warp; i < n; i += RT / 32) {
        const double* r = Pm + (size_t)i * ldp;
        double bv = -1.0;
        int bi = 0x7fffffff;
        for (int j = lane; j < n; j += 32) best_of(bv, bi, r[j] * M[j], j);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            best_of(bv, bi, ov, oi);
        }
        if (lane == 0) {
            lj[(vb + c) * ldp + i] = bv;
            cj[(vb + c) * (size_t)n + i] = bi;
        }
    }
}

// One CTA per vector: the states follow C down the tree, depth by depth (node ids of a depth are
// contiguous and every parent has a smaller id than its children).
__global__ void __launch_bounds__(256) joint_backtrack_kernel(
    int n, int n_nodes, const int32_t* __restrict__ parent, const int32_t* __restrict__ depth_off,
    int n_depths, const int32_t* __restrict__ cj, int32_t* __restrict__ states) {
    const int b = blockIdx.x;
    const size_t vb = (size_t)b * n_nodes;
    for (int d = 1; d < n_depths; ++d) {
        __syncthreads();
        for (int c = depth_off[d] + threadIdx.x; c < depth_off[d + 1]; c += blockDim.x)
            states[vb + c] = cj[(vb + c) * (size_t)n + states[vb + parent[c]]];
    }
}

}  // namespace cevb

using namespace cevb;

extern "C" int cevb_marginal_down(const double* P, int B, int E, int n, int n_nodes,
                                  const int32_t* parent, const int32_t* child_off,
                                  const int32_t* child_idx, const int32_t* branch_of,
                                  const int32_t* leaf_code, const int32_t* depth_off_host,
                                  int n_depths, const double* root_freq, const double* node_vec,
                                  double* down, double* post, void* stream) {
    if (B <= 0 || n_nodes <= 0) return 0;
    if (B > 65535) {
        set_error("cevb_marginal_down: B=%d exceeds one launch; chunk the batch", B);
        return -1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int ldp = ldp_of(n);
    const size_t smem = 2 * (size_t)n * sizeof(double);
    for (int d = 0; d < n_depths; ++d) {
        const int cnt = depth_off_host[d + 1] - depth_off_host[d];
        if (cnt <= 0) continue;
        dim3 grid(cnt, B);
        marginal_down_kernel<<<grid, RT, smem, st>>>(P, E, n, ldp, n_nodes, depth_off_host[d], parent,
                                                     child_off, child_idx, branch_of, leaf_code,
                                                     root_freq, node_vec, down, post);
        CEVB_CHECK_LAUNCH("marginal_down_kernel");
    }
    return 0;
}

extern "C" int cevb_joint_reconstruct(const double* P, int B, int E, int n, int n_nodes, int root,
                                      const int32_t* parent, const int32_t* level_nodes,
                                      const int32_t* level_off_host, int n_levels,
                                      const int32_t* child_off, const int32_t* child_idx,
                                      const int32_t* branch_of, const int32_t* leaf_code,
                                      const int32_t* depth_off, int n_depths,
                                      const double* root_freq, double* lj, int32_t* cj,
                                      long long* expo, int32_t* states, double* joint_logp,
                                      void* stream) {
    if (B <= 0 || n_nodes <= 0) return 0;
    if (B > 65535) {
        set_error("cevb_joint_reconstruct: B=%d exceeds one launch; chunk the batch", B);
        return -1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int ldp = ldp_of(n);
    const size_t smem = (size_t)n * sizeof(double);
    {
        // every leaf (any height 0 node) in one launch; internal ids leave at once
        dim3 grid(n_nodes, B);
        joint_up_kernel<<<grid, RT, smem, st>>>(P, E, n, ldp, n_nodes, root, nullptr, child_off,
                                                child_idx, branch_of, leaf_code, root_freq, lj, cj,
                                                expo, states, joint_logp);
        CEVB_CHECK_LAUNCH("joint_up_kernel (leaves)");
    }
    for (int lv = 0; lv < n_levels; ++lv) {
        const int cnt = level_off_host[lv + 1] - level_off_host[lv];
        if (cnt <= 0) continue;
        dim3 grid(cnt, B);
        joint_up_kernel<<<grid, RT, smem, st>>>(P, E, n, ldp, n_nodes, root,
                                                level_nodes + level_off_host[lv], child_off,
                                                child_idx, branch_of, leaf_code, root_freq, lj, cj,
                                                expo, states, joint_logp);
        CEVB_CHECK_LAUNCH("joint_up_kernel");
    }
    joint_backtrack_kernel<<<B, 256, 0, st>>>(n, n_nodes, parent, depth_off, n_depths, cj, states);
    CEVB_CHECK_LAUNCH("joint_backtrack_kernel");
    return 0;
}
