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
extern "C" int cevb_diag_fp64_peak(int kind, int blocks, int iters, double* out, void* stream) {
    if (kind == 0)
        dmma_peak_kernel<<<blocks, 512, 0, (cudaStream_t)stream>>>(iters, out);
    else
        dfma_peak_kernel<<<blocks, 512, 0, (cudaStream_t)stream>>>(iters, out);
    CEVB_CHECK_LAUNCH("fp64 peak diagnostic");
    return 0;
}
