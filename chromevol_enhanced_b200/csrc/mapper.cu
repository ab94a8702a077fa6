// Forward stochastic mapping (Gillespie walks down the tree), Philox4x32-10 on the device.
//
// Reference: StochasticMapper.perform_stochastic_mapping / simulate_events_on_branch /
// _classify_event, AR:441-576.  The reference draws from numpy's global MT19937 stream
// (unseeded), so only the DISTRIBUTION of per-branch event counts can be matched; the kernel
// uses a counter-based generator keyed by (seed; replicate, node, draw) so that results do not
// depend on how replicates are sharded across launches or GPUs.
//
// One thread per replicate.  All threads of a warp visit the same node at the same time, so the
// per-node state arrays ([node][replicate], uint16) are read and written fully coalesced.  The
// rate matrix is compacted per CTA into shared memory: hold rate -Q[s][s] (AR:483), and for
// every state the list of positive off-diagonal targets with cumulative weights (AR:499-510).
#include "common.cuh"

namespace cevb {

constexpr int MT = 256;   // threads per CTA
constexpr int KT = 12;    // compact targets per state; denser rows fall back to a dense scan

struct Philox {
    uint32_t key0, key1;
};

__device__ __forceinline__ void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3,
                                             uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}

// two uniforms in [0, 1) with 53 random bits each
__device__ __forceinline__ void philox_uniform2(unsigned long long seed, unsigned long long rep,
                                                uint32_t node, uint32_t draw, double& u0, double& u1) {
    uint32_t c0 = (uint32_t)rep, c1 = (uint32_t)(rep >> 32), c2 = node, c3 = draw;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c0, c1, c2, c3, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    const unsigned long long a = ((unsigned long long)c1 << 32) | c0;
    const unsigned long long b = ((unsigned long long)c3 << 32) | c2;
    u0 = (double)(a >> 11) * (1.0 / 9007199254740992.0);
    u1 = (double)(b >> 11) * (1.0 / 9007199254740992.0);
}

// AR:441-464; returns the index into EVENT_TYPES (AR:602-604)
__device__ __forceinline__ int classify_event(int from, int to, int bn) {
    if (to == from + 1) return 0;                              // gain
    if (to == from - 1) return 1;                              // loss
    if (to == 2 * from && from > 0) return 2;                  // duplication
    if (bn > 0) {
        if (to == from + bn) return 3;                         // base_number_gain
        if (to == from - bn) return 4;                         // base_number_loss
    }
    const int delta = to - from;
    if (delta > 1 && delta < from) return 5;                   // demi_gain_like
    if (delta < -1 && 2 * (-delta) < from) return 6;           // demi_loss_like: |d| < from/2
    return 7;                                                  // other
}

struct QTable {
    double* hold;     // [n]
    double* tot;      // [n]
    double* cum;      // [n][KT]
    uint16_t* tgt;    // [n][KT]
    uint8_t* cnt;     // [n], 255 = dense row
};

__device__ QTable carve_table(unsigned char* raw, int n) {
    QTable tb;
    tb.hold = reinterpret_cast<double*>(raw);
    tb.tot = tb.hold + n;
    tb.cum = tb.tot + n;
    tb.tgt = reinterpret_cast<uint16_t*>(tb.cum + (size_t)n * KT);
    tb.cnt = reinterpret_cast<uint8_t*>(tb.tgt + (size_t)n * KT);
    return tb;
}

static size_t table_bytes(int n) { return (size_t)n * (8 + 8 + 8 * KT + 2 * KT + 1) + 16; }

__device__ void build_table(const QTable& tb, const double* __restrict__ q, int n, int ldq) {
    for (int s = threadIdx.x; s < n; s += blockDim.x) {
        const double* row = q + (size_t)s * ldq;
        tb.hold[s] = -row[s];
        double acc = 0.0;
        int c = 0;
        for (int j = 0; j < n; ++j) {
            const double w = (j == s) ? 0.0 : fmax(row[j], 0.0);   // AR:500-503
            if (w > 0.0) {
                acc += w;
                if (c < KT) {
                    tb.tgt[s * KT + c] = (uint16_t)j;
                    tb.cum[s * KT + c] = acc;
                }
                ++c;
            }
        }
        tb.tot[s] = acc;
        tb.cnt[s] = (c > KT) ? 255 : (uint8_t)c;
    }
    __syncthreads();
}

// next state given the second uniform (AR:509-510: cdf search, side='right')
__device__ __forceinline__ int pick_next(const QTable& tb, const double* __restrict__ q, int n, int ldq,
                                         int s, double u) {
    const double x = u * tb.tot[s];
    const int cnt = tb.cnt[s];
    if (cnt != 255) {
        int c = 0;
        while (c < cnt - 1 && !(tb.cum[s * KT + c] > x)) ++c;
        return tb.tgt[s * KT + c];
    }
    const double* row = q + (size_t)s * ldq;
    double acc = 0.0;
    int last = s;
    for (int j = 0; j < n; ++j) {
        const double w = (j == s) ? 0.0 : fmax(row[j], 0.0);
        if (w > 0.0) {
            acc += w;
            last = j;
            if (acc > x) return j;
        }
    }
    return last;
}

__global__ void __launch_bounds__(MT) stoch_map_kernel(
    const double* __restrict__ q, int n, int ldq, const double* __restrict__ root_probs, int n_nodes,
    const int32_t* __restrict__ parent, const int32_t* __restrict__ preorder,
    const double* __restrict__ branch_len, int bn, unsigned long long seed, long long rep0,
    long long count, int n_bins, uint32_t* __restrict__ hist, unsigned long long* __restrict__ sums,
    uint16_t* __restrict__ state, long long stride, unsigned long long* __restrict__ n_events,
    int n_sets, long long q_stride, const int32_t* __restrict__ node_set,
    const int32_t* __restrict__ bn_set, unsigned int tbytes) {
    // one compact table per parameter set (branch-specific models, AR:569: the branch above a
    // node uses branch_params.get(id(node), params)); n_sets == 1 is the homogeneous model
    extern __shared__ __align__(16) unsigned char raw[];
    for (int qs = 0; qs < n_sets; ++qs)
        build_table(carve_table(raw + (size_t)qs * tbytes, n), q + (size_t)qs * q_stride, n, ldq);

    const long long slot = (long long)blockIdx.x * MT + threadIdx.x;
    const bool active = slot < count;
    const unsigned long long rep = (unsigned long long)(rep0 + slot);
    const long long sl = active ? slot : 0;
    unsigned long long my_events = 0;

    // root state ~ Categorical(root_probs)   (AR:544-545)
    const int root = preorder[0];
    {
        double u0, u1;
        philox_uniform2(seed, rep, (uint32_t)root, 0u, u0, u1);
        double acc = 0.0, total = 0.0;
        for (int j = 0; j < n; ++j) total += root_probs[j];
        const double x = u0 * total;
        int s = n - 1;
        for (int j = 0; j < n; ++j) {
            acc += root_probs[j];
            if (acc > x) {
                s = j;
                break;
            }
        }
        if (active) state[(size_t)root * stride + sl] = (uint16_t)s;
    }

    for (int k = 1; k < n_nodes; ++k) {
        const int node = preorder[k];
        const double t = branch_len[node];
        const int qs = node_set ? node_set[node] : 0;
        const QTable tb = carve_table(raw + (size_t)qs * tbytes, n);
        const double* qn = q + (size_t)qs * q_stride;
        const int bnn = bn_set ? bn_set[qs] : bn;
        int s = state[(size_t)parent[node] * stride + sl];
        int c[CEVB_NEVENT];
#pragma unroll
        for (int e = 0; e < CEVB_NEVENT; ++e) c[e] = 0;
        if (active && t > 0.0) {                                       // AR:562-566
            double time = 0.0;
            uint32_t draw = 0;
            for (;;) {
                const double rate = tb.hold[s];
                if (rate <= 1e-9) break;                               // AR:483-486
                double u0, u1;
                philox_uniform2(seed, rep, (uint32_t)node, draw++, u0, u1);
                const double dt = -log(1.0 - u0) / rate;               // AR:489
                if (time + dt >= t) break;                             // AR:491-493
                time += dt;
                if (tb.tot[s] <= 1e-9) break;                          // AR:505-507
                const int nxt = pick_next(tb, qn, n, ldq, s, u1);
                const int ty = classify_event(s, nxt, bnn);
#pragma unroll
                for (int e = 0; e < CEVB_NEVENT; ++e) c[e] += (e == ty);
                s = nxt;
                ++my_events;
            }
        }
        if (active) state[(size_t)node * stride + sl] = (uint16_t)s;  // AR:575
#pragma unroll
        for (int e = 0; e < CEVB_NEVENT; ++e) {
            const unsigned v = active ? (unsigned)c[e] : 0u;
            if (v) {
                const unsigned bin = v < (unsigned)n_bins ? v : (unsigned)(n_bins - 1);
                atomicAdd(&hist[((size_t)node * CEVB_NEVENT + e) * n_bins + bin], 1u);
            }
            const unsigned tot = __reduce_add_sync(0xffffffffu, v);
            if (tot) {
                const unsigned sq = __reduce_add_sync(0xffffffffu, v * v);
                if ((threadIdx.x & 31) == 0) {
                    atomicAdd(&sums[((size_t)node * CEVB_NEVENT + e) * 2], (unsigned long long)tot);
                    atomicAdd(&sums[((size_t)node * CEVB_NEVENT + e) * 2 + 1], (unsigned long long)sq);
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my_events += __shfl_xor_sync(0xffffffffu, my_events, o);
    if ((threadIdx.x & 31) == 0 && my_events) atomicAdd(n_events, my_events);
}

// One branch, one replicate, with the individual events recorded (host-facing helper behind
// StochasticMapper.simulate_events_on_branch, AR:467-521).
__global__ void stoch_branch_trace_kernel(const double* __restrict__ q, int n, int ldq, int start,
                                          double t, int bn, unsigned long long seed,
                                          unsigned long long rep, int max_events,
                                          double* __restrict__ ev_time, int32_t* __restrict__ ev_from,
                                          int32_t* __restrict__ ev_to, int32_t* __restrict__ ev_type,
                                          int32_t* __restrict__ out /* [0]=n_events, [1]=end */) {
    extern __shared__ __align__(16) unsigned char raw[];
    const QTable tb = carve_table(raw, n);
    build_table(tb, q, n, ldq);
    if (threadIdx.x != 0) return;
    int s = start, ne = 0;
    double time = 0.0;
    uint32_t draw = 0;
    while (time < t) {
        if (s < 0 || s >= n) break;
        const double rate = tb.hold[s];
        if (rate <= 1e-9) break;
        double u0, u1;
        philox_uniform2(seed, rep, 0u, draw++, u0, u1);
        const double dt = -log(1.0 - u0) / rate;
        if (time + dt >= t) break;
        time += dt;
        if (tb.tot[s] <= 1e-9) break;
        const int nxt = pick_next(tb, q, n, ldq, s, u1);
        if (ne < max_events) {
            ev_time[ne] = time;
            ev_from[ne] = s;
            ev_to[ne] = nxt;
            ev_type[ne] = classify_event(s, nxt, bn);
        }
        ++ne;
        s = nxt;
    }
    out[0] = ne;
    out[1] = s;
}

}  // namespace cevb

using namespace cevb;

static int stoch_map_impl(const double* q, int n_sets, long long q_stride, const int32_t* node_set,
                          const int32_t* bn_set, int n, const double* root_probs, int n_nodes,
                          const int32_t* parent, const int32_t* preorder, const double* branch_len,
                          int base_number, unsigned long long seed, long long rep_begin,
                          long long rep_end, int n_bins, uint32_t* hist, unsigned long long* sums,
                          uint16_t* state_scratch, long long rep_chunk, unsigned long long* n_events,
                          void* stream) {
    if (rep_end <= rep_begin) return 0;
    if (n < 2 || n > 65535 || n_bins < 2 || rep_chunk < 1 || n_sets < 1) {
        set_error("cevb_stoch_map: bad arguments (n=%d, n_bins=%d, rep_chunk=%lld, n_sets=%d)", n,
                  n_bins, rep_chunk, n_sets);
        return -1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t tbytes = (table_bytes(n) + 15) / 16 * 16;
    const size_t smem = tbytes * (size_t)n_sets;
    int dev = 0, optin = 0;
    CEVB_CHECK_CUDA(cudaGetDevice(&dev));
    CEVB_CHECK_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (smem > (size_t)optin) {
        set_error("cevb_stoch_map: %d parameter sets of order %d need %zu B of shared memory", n_sets,
                  n, smem);
        return -1;
    }
    CEVB_CHECK_CUDA(cudaFuncSetAttribute(stoch_map_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
    for (long long r0 = rep_begin; r0 < rep_end; r0 += rep_chunk) {
        const long long count = (rep_end - r0 < rep_chunk) ? (rep_end - r0) : rep_chunk;
        const long long blocks = (count + MT - 1) / MT;
        stoch_map_kernel<<<(unsigned)blocks, MT, smem, st>>>(
            q, n, ldq_of(n), root_probs, n_nodes, parent, preorder, branch_len, base_number, seed, r0,
            count, n_bins, hist, sums, state_scratch, rep_chunk, n_events, n_sets, q_stride, node_set,
            bn_set, (unsigned int)tbytes);
        CEVB_CHECK_LAUNCH("stoch_map_kernel");
    }
    return 0;
}

extern "C" int cevb_stoch_map(const double* q, int n, const double* root_probs, int n_nodes,
                              const int32_t* parent, const int32_t* preorder,
                              const double* branch_len, int base_number, unsigned long long seed,
                              long long rep_begin, long long rep_end, int n_bins, uint32_t* hist,
                              unsigned long long* sums, uint16_t* state_scratch, long long rep_chunk,
                              unsigned long long* n_events, void* stream) {
    return stoch_map_impl(q, 1, 0, nullptr, nullptr, n, root_probs, n_nodes, parent, preorder,
                          branch_len, base_number, seed, rep_begin, rep_end, n_bins, hist, sums,
                          state_scratch, rep_chunk, n_events, stream);
}

extern "C" int cevb_stoch_map_sets(const double* q, int n_sets, long long q_stride,
                                   const int32_t* node_set, const int32_t* base_numbers, int n,
                                   const double* root_probs, int n_nodes, const int32_t* parent,
                                   const int32_t* preorder, const double* branch_len,
                                   unsigned long long seed, long long rep_begin, long long rep_end,
                                   int n_bins, uint32_t* hist, unsigned long long* sums,
                                   uint16_t* state_scratch, long long rep_chunk,
                                   unsigned long long* n_events, void* stream) {
    if (node_set == nullptr || base_numbers == nullptr) {
        set_error("cevb_stoch_map_sets: node_set and base_numbers are required");
        return -1;
    }
    return stoch_map_impl(q, n_sets, q_stride, node_set, base_numbers, n, root_probs, n_nodes, parent,
                          preorder, branch_len, 0, seed, rep_begin, rep_end, n_bins, hist, sums,
                          state_scratch, rep_chunk, n_events, stream);
}

extern "C" int cevb_stoch_branch_trace(const double* q, int n, int start_state, double t,
                                       int base_number, unsigned long long seed,
                                       unsigned long long replicate, int max_events, double* ev_time,
                                       int32_t* ev_from, int32_t* ev_to, int32_t* ev_type,
                                       int32_t* out, void* stream) {
    const size_t smem = table_bytes(n);
    CEVB_CHECK_CUDA(cudaFuncSetAttribute(stoch_branch_trace_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    stoch_branch_trace_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(
        q, n, ldq_of(n), start_state, t, base_number, seed, replicate, max_events, ev_time, ev_from,
        ev_to, ev_type, out);
    CEVB_CHECK_LAUNCH("stoch_branch_trace_kernel");
    return 0;
}
