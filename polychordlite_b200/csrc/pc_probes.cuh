// pc_probes.cuh -- small non-templated probe kernels (parity tests drive them through the C ABI).
#pragma once
#include "pc_kernels.cuh"

namespace pc {

// directions of one chain in use order, before whitening: out[i*D + r]
__global__ void pc_directions_kernel(int D, int R, int LD, unsigned seed, unsigned long long uid, double* nh_global,
                                     double* out) {
    extern __shared__ __align__(16) unsigned char smem[];
    const ChainScratch cs = chain_scratch(smem, D, R, LD, false, LIKE_GAUSSIAN, 1, nh_global);
    prep_chain(D, R, LD, seed, uid, cs);
    for (int i = 0; i < R; ++i)
        for (int r = threadIdx.x; r < D; r += 32) out[(size_t)i * D + r] = nh_global[(size_t)cs.deck[i] * LD + r];
}

__global__ void pc_philox_kernel(const unsigned* ctr, const unsigned* key, unsigned* out) {
    u4 o = philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
    out[0] = o.x; out[1] = o.y; out[2] = o.z; out[3] = o.w;
}
__global__ void pc_uniforms_kernel(unsigned seed, unsigned tag, unsigned long long uid, unsigned a0, unsigned b, int n,
                                   double* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = uniform(seed, tag, uid, a0 + (unsigned)i, b);
}
__global__ void pc_inv_normal_kernel(const double* pin, int n, double* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = inv_normal_cdf(pin[i]);
}
// evidence recurrences for an explicit death sequence (parity probe for evidence_deaths)
__global__ void pc_evidence_kernel(DevRun* st, const double* logLs, int count, int n_start, double* logw_out) {
    extern __shared__ __align__(16) unsigned char smem[];
    double* sc = (double*)smem;
    double* skey = sc + 64;
    for (int i = threadIdx.x; i < count; i += blockDim.x) skey[i] = logLs[i];
    __syncthreads();
    evidence_deaths(st, skey, count, n_start, logw_out, sc);
}
// FP64 fused-multiply-add throughput of the device (the arithmetic roofline bench.py quotes the run against):
// eight independent chains per thread, 2 flops per FMA
__global__ void pc_fp64_peak_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}
// which = 0: the one-warp factorisation (clustering: one factor per cluster), launched with 32 threads;
// which = 1: the CTA-wide one of the run kernel's update (a is overwritten), launched with 256 threads
__global__ void pc_cholesky_kernel(double* a, double* L, int D, int* fb, int which) {
    int f = which ? block_cholesky(a, L, D) : warp_cholesky(a, L, D);
    if (threadIdx.x == 0) *fb = f;
}

}  // namespace pc
