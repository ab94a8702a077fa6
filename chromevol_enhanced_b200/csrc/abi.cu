// Probes, error text and a small DMMA throughput diagnostic for the C ABI.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace cevb {

static char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

// Diagnostic: issue-rate ceiling of DMMA.8x8x4 from registers (no memory traffic), used once by
// bench.py to state the FP64 tensor peak this pool's B200 actually delivers.
__global__ void __launch_bounds__(512) dmma_peak_kernel(int iters, double* out) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        c[i][0] = 0.0;
        c[i][1] = 0.0;
    }
    double av = 1.0 + threadIdx.x * 1e-9, bv = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], av, bv);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;  // keep the chain alive without a real store
}

__global__ void __launch_bounds__(512) dfma_peak_kernel(int iters, double* out) {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 1e-3 + i;
    const double av = 1.0 + threadIdx.x * 1e-9, bv = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fma(c[i], av, bv);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    if (s == 123.456) out[0] = s;
}

// Diagnostic: DMMA throughput as a function of resident warps, with 13 independent accumulator
// tiles per warp (the apply kernel's shape); smem != 0 loads every B fragment from shared
// memory right before its DMMA (one LDS.64 per DMMA), smem == 2 double-buffers them in registers.
__global__ void __launch_bounds__(512) dmma_probe_kernel(int iters, int smem_mode, double* out) {
    __shared__ double sb[13 * 32 * 8];
    for (int x = threadIdx.x; x < 13 * 32 * 8; x += blockDim.x) sb[x] = 1.0 + 1e-9 * x;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    double c[13][2];
#pragma unroll
    for (int i = 0; i < 13; ++i) c[i][0] = c[i][1] = 0.0;
    const double av = 1.0 + threadIdx.x * 1e-9;
    double bv = 1.0 - threadIdx.x * 1e-9;
    if (smem_mode == 0) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 13; ++i) dmma884(c[i][0], c[i][1], av, bv);
        }
    } else if (smem_mode == 1) {
        for (int it = 0; it < iters; ++it) {
            const double* bp = sb + (it & 7) * 13 * 32 + lane;
#pragma unroll
            for (int i = 0; i < 13; ++i) dmma884(c[i][0], c[i][1], av, bp[i * 32]);
        }
    } else if (smem_mode >= 200) {
        // 13 DMMAs + (smem_mode - 200) independent 32-bit integer ops per iteration: do
        // instructions of other pipes overlap with DMMA issue, or does a DMMA hold the scheduler?
        const int nint = smem_mode - 200;
        unsigned x0 = threadIdx.x, x1 = threadIdx.x * 3u, x2 = 7u, x3 = 11u;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 13; ++i) dmma884(c[i][0], c[i][1], av, bv);
#pragma unroll 1
            for (int q = 0; q < nint; q += 4) {
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x0) : "r"(x1), "r"(x2));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x1) : "r"(x2), "r"(x3));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x2) : "r"(x3), "r"(x0));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x3) : "r"(x0), "r"(x1));
            }
        }
        if ((x0 ^ x1 ^ x2 ^ x3) == 0x12345u) bv = 2.0;
    } else if (smem_mode >= 100) {
        // 13 DMMAs + (smem_mode - 100) independent DADDs per iteration, register operands: what
        // FP64 ALU work in the epilogue costs the tensor pipe (about 2.5 clocks per DADD)
        const int nadd = smem_mode - 100;
        double e0 = av, e1 = bv, e2 = av + 1.0, e3 = bv + 1.0;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 13; ++i) dmma884(c[i][0], c[i][1], av, bv);
            for (int q = 0; q < nadd; q += 4) {
                e0 += av;
                e1 += av;
                e2 += av;
                e3 += av;
            }
        }
        bv = e0 + e1 + e2 + e3;
    } else if (smem_mode == 5) {
        // 13 volatile loads (kept in program order by ptxas) and then the 13 DMMAs
        for (int it = 0; it < iters; ++it) {
            const volatile double* bp = sb + (it & 7) * 13 * 32 + lane;
            double bc[13];
#pragma unroll
            for (int i = 0; i < 13; ++i) bc[i] = bp[i * 32];
#pragma unroll
            for (int i = 0; i < 13; ++i) dmma884(c[i][0], c[i][1], av, bc[i]);
        }
    } else if (smem_mode == 6) {
        // two register sets: the volatile loads of iteration it+1 are issued before the DMMAs of it
        double b0[13], b1[13];
        {
            const volatile double* bp = sb + lane;
#pragma unroll
            for (int i = 0; i < 13; ++i) b0[i] = bp[i * 32];
        }
        for (int it = 0; it < iters; it += 2) {
            const volatile double* bp1 = sb + ((it + 1) & 7) * 13 * 32 + lane;
#pragma unroll
            for (int i = 0; i < 13; ++i) b1[i] = bp1[i * 32];
#pragma unroll
            for (int i = 0; i < 13; ++i) dmma884(c[i][0], c[i][1], av, b0[i]);
            const volatile double* bp0 = sb + ((it + 2) & 7) * 13 * 32 + lane;
#pragma unroll
            for (int i = 0; i < 13; ++i) b0[i] = bp0[i * 32];
#pragma unroll
            for (int i = 0; i < 13; ++i) dmma884(c[i][0], c[i][1], av, b1[i]);
        }
    } else if (smem_mode == 4) {
        // six LDS.128 (two tiles each) + one LDS.64 per 13 DMMAs
        for (int it = 0; it < iters; ++it) {
            const double* bp = sb + (it & 7) * 13 * 32;
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const double2 b2 = *reinterpret_cast<const double2*>(bp + (i * 32 + lane) * 2);
                dmma884(c[2 * i][0], c[2 * i][1], av, b2.x);
                dmma884(c[2 * i + 1][0], c[2 * i + 1][1], av, b2.y);
            }
            dmma884(c[12][0], c[12][1], av, bp[12 * 32 + lane]);
        }
    } else if (smem_mode == 3) {
        for (int it = 0; it < iters; ++it) {
            const double* bp = sb + (it & 7) * 13 * 32 + lane;
            double bc[13];
#pragma unroll
            for (int i = 0; i < 13; ++i) bc[i] = bp[i * 32];
            asm volatile("" ::: "memory");
#pragma unroll
            for (int i = 0; i < 13; ++i) dmma884(c[i][0], c[i][1], av, bc[i]);
            asm volatile("" ::: "memory");
        }
    } else {
        double bc[13];
#pragma unroll
        for (int i = 0; i < 13; ++i) bc[i] = sb[i * 32 + lane];
        for (int it = 0; it < iters; ++it) {
            const double* bp = sb + ((it + 1) & 7) * 13 * 32 + lane;
#pragma unroll
            for (int i = 0; i < 13; ++i) {
                const double b0 = bc[i];
                bc[i] = bp[i * 32];
                dmma884(c[i][0], c[i][1], av, b0);
            }
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 13; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s + bv;
}

}  // namespace cevb

using namespace cevb;

extern "C" int cevb_abi_version(void) { return CEVB_ABI_VERSION; }

extern "C" const char* cevb_last_error(void) { return g_error; }

extern "C" int cevb_ldq(int n) { return ldq_of(n); }
extern "C" int cevb_ldp(int n) { return ldp_of(n); }

extern "C" int cevb_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* smem_optin) {
    int dev = 0;
    CEVB_CHECK_CUDA(cudaGetDevice(&dev));
    int sm = 0, maj = 0, min = 0, optin = 0;
    CEVB_CHECK_CUDA(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev));
    CEVB_CHECK_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
    CEVB_CHECK_CUDA(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
    CEVB_CHECK_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (sm_count) *sm_count = sm;
    if (cc_major) *cc_major = maj;
    if (cc_minor) *cc_minor = min;
    if (smem_optin) *smem_optin = (size_t)optin;
    if (maj != 10) {
        set_error("device compute capability %d.%d is not sm_100", maj, min);
        return -4;
    }
    return 0;
}

// Diagnostics (not part of the reference-facing surface): launches `blocks` CTAs of 512 threads
// that each issue iters*8 DMMA.8x8x4 per warp (kind 0) or iters*16 DFMA per thread (kind 1).
extern "C" int cevb_diag_dmma_probe(int blocks, int threads, int iters, int smem_mode, double* out,
                                    void* stream) {
    dmma_probe_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(iters, smem_mode, out);
    CEVB_CHECK_LAUNCH("dmma probe");
    return 0;
}

extern "C" int cevb_diag_fp64_peak(int kind, int blocks, int iters, double* out, void* stream) {
    if (kind == 0)
        dmma_peak_kernel<<<blocks, 512, 0, (cudaStream_t)stream>>>(iters, out);
    else
        dfma_peak_kernel<<<blocks, 512, 0, (cudaStream_t)stream>>>(iters, out);
    CEVB_CHECK_LAUNCH("fp64 peak diagnostic");
    return 0;
}
