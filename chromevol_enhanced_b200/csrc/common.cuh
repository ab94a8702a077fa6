// Shared helpers for the sm_100a kernels of the chromosome-number likelihood path.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/chromevol_b200.h"

namespace cevb {

// last error text, one buffer per process (defined in abi.cu)
void set_error(const char* fmt, ...);

#define CEVB_CHECK_CUDA(expr)                                                              \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            cevb::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,          \
                            cudaGetErrorString(_e));                                       \
            return -2;                                                                     \
        }                                                                                  \
    } while (0)

#define CEVB_CHECK_LAUNCH(name)                                                            \
    do {                                                                                   \
        cudaError_t _e = cudaGetLastError();                                               \
        if (_e != cudaSuccess) {                                                           \
            cevb::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));     \
            return -3;                                                                     \
        }                                                                                  \
    } while (0)

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// row strides: Q/powers and P rows are padded to an even number of doubles (16-byte rows)
__host__ __device__ inline int ldq_of(int n) { return round_up(n, 2); }
__host__ __device__ inline int ldp_of(int n) { return round_up(n, 2); }

// ---- FP64 tensor-core MMA: D(8x8) += A(8x4) * B(4x8), SASS DMMA.8x8x4 -------------------
// lane = 4*g + t.  A fragment: A[g][t].  B fragment: B[t][n=g].  C fragment: C[g][2t], C[g][2t+1].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile(
        "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

// Column permutation used for every B/C fragment held in shared memory: MMA column index j is
// stored at physical column sigma(j) of its 8-wide tile, sigma(2t) = t, sigma(2t+1) = t + 4.
// A lane's two accumulators then sit at physical columns t and t+4, so fragment loads/stores
// are plain 8-byte accesses at (row g, column t) -- bank-conflict free when the row stride is
// 4 or 12 (mod 16) doubles.
__device__ __forceinline__ int sigma8(int g) { return (g >> 1) + ((g & 1) << 2); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace cevb
