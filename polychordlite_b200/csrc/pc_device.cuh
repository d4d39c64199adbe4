// pc_device.cuh -- device-side numeric building blocks of the nested-sampling engine.
//
// RNG stream spec: Philox4x32-10, key = (seed, tag), counter = (a, b, uid_lo, uid_hi).
// Every random number of a run is addressed by (tag, uid, a, b), so the sampler is independent of
// the order warps execute in and bit-reproducible for a fixed seed (the contract the reference's
// tests/test_run_pypolychord.py:77-90 pins for seed >= 0).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace pc {

enum Tag : uint32_t { TAG_INIT = 1, TAG_SEED = 2, TAG_DIR = 3, TAG_SHUF = 4, TAG_SLICE = 5, TAG_POST = 6, TAG_LIKE = 7, TAG_BOOST = 8 };

constexpr unsigned FULL = 0xffffffffu;

struct u4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ u4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                     uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return u4{c0, c1, c2, c3};
}

// 52 random bits + half-ulp offset: strictly inside (0,1).
__host__ __device__ __forceinline__ double bits_to_unit(uint32_t lo, uint32_t hi) {
    uint64_t x = ((uint64_t)hi << 32) | lo;
    return ((double)(x >> 12) + 0.5) * (1.0 / 4503599627370496.0);
}

__host__ __device__ __forceinline__ double uniform(uint32_t seed, uint32_t tag, uint64_t uid, uint32_t a, uint32_t b) {
    u4 o = philox4x32_10(a, b, (uint32_t)uid, (uint32_t)(uid >> 32), seed, tag);
    return bits_to_unit(o.x, o.y);
}
__host__ __device__ __forceinline__ void uniform2(uint32_t seed, uint32_t tag, uint64_t uid, uint32_t a, uint32_t b,
                                                  double& u0, double& u1) {
    u4 o = philox4x32_10(a, b, (uint32_t)uid, (uint32_t)(uid >> 32), seed, tag);
    u0 = bits_to_unit(o.x, o.y);
    u1 = bits_to_unit(o.z, o.w);
}

// Inverse normal CDF, algorithm AS241 PPND16 (Wichura 1988) -- the algorithm the reference's
// utils.F90:777-966 (inv_normal_cdf -> r8_normal_01_cdf_inverse) implements.  p in (0,1).
__host__ __device__ inline double inv_normal_cdf(double p);

// The two halves of inv_normal_cdf for warp code that wants them apart: in a warp of 32 deviates almost always SOME
// lane is in a tail (15 % of the draws are), so the combined function pays for the log / sqrt / second rational on
// every call.  prep_chain evaluates the central form for everybody and queues the tail arguments, which are then
// worked off 32 at a time.  Same expressions, same results, as inv_normal_cdf above.
__host__ __device__ __forceinline__ bool inv_normal_cdf_central(double p, double& out) {
    const double q = p - 0.5;
    const double r = 0.180625 - q * q;
    double num = 2.5090809287301226727e+3;
    num = num * r + 3.3430575583588128105e+4;
    num = num * r + 6.7265770927008700853e+4;
    num = num * r + 4.5921953931549871457e+4;
    num = num * r + 1.3731693765509461125e+4;
    num = num * r + 1.9715909503065514427e+3;
    num = num * r + 1.3314166789178437745e+2;
    num = num * r + 3.3871328727963666080;
    double den = 5.2264952788528545610e+3;
    den = den * r + 2.8729085735721942674e+4;
    den = den * r + 3.9307895800092710610e+4;
    den = den * r + 2.1213794301586595867e+4;
    den = den * r + 5.3941960214247511077e+3;
    den = den * r + 6.8718700749205790830e+2;
    den = den * r + 4.2313330701600911252e+1;
    den = den * r + 1.0;
    out = q * num / den;
    return fabs(q) <= 0.425;
}
__host__ __device__ inline double inv_normal_cdf_tail(double p) {   // |p - 0.5| > 0.425
    const double q = p - 0.5;
    double r = (q < 0.0) ? p : 1.0 - p;
    r = sqrt(-log(r));
    double val;
    if (r <= 5.0) {
        r -= 1.6;
        double num = 7.74545014278341407640e-4;
        num = num * r + 2.27238449892691845833e-2;
        num = num * r + 2.41780725177450611770e-1;
        num = num * r + 1.27045825245236838258;
        num = num * r + 3.64784832476320460504;
        num = num * r + 5.76949722146069140550;
        num = num * r + 4.63033784615654529590;
        num = num * r + 1.42343711074968357734;
        double den = 1.05075007164441684324e-9;
        den = den * r + 5.47593808499534494600e-4;
        den = den * r + 1.51986665636164571966e-2;
        den = den * r + 1.48103976427480074590e-1;
        den = den * r + 6.89767334985100004550e-1;
        den = den * r + 1.67638483018380384940;
        den = den * r + 2.05319162663775882187;
        den = den * r + 1.0;
        val = num / den;
    } else {
        r -= 5.0;
        double num = 2.01033439929228813265e-7;
        num = num * r + 2.71155556874348757815e-5;
        num = num * r + 1.24266094738807843860e-3;
        num = num * r + 2.65321895265761230930e-2;
        num = num * r + 2.96560571828504891230e-1;
        num = num * r + 1.78482653991729133580;
        num = num * r + 5.46378491116411436990;
        num = num * r + 6.65790464350110377720;
        double den = 2.04426310338993978564e-15;
        den = den * r + 1.42151175831644588870e-7;
        den = den * r + 1.84631831751005468180e-5;
        den = den * r + 7.86869131145613259100e-4;
        den = den * r + 1.48753612908506148525e-2;
        den = den * r + 1.36929880922735805310e-1;
        den = den * r + 5.99832206555887937690e-1;
        den = den * r + 1.0;
        val = num / den;
    }
    return (q < 0.0) ? -val : val;
}

__host__ __device__ inline double inv_normal_cdf(double p) {
    double v;
    return inv_normal_cdf_central(p, v) ? v : inv_normal_cdf_tail(p);
}

// utils.F90:376-388
__host__ __device__ __forceinline__ double logaddexp(double a, double b) {
    return (a > b) ? a + log(exp(b - a) + 1.0) : b + log(exp(a - b) + 1.0);
}

#ifdef __CUDACC__
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
// L2-coherent loads for data another CTA produced before a run barrier.
__device__ __forceinline__ double ldcg(const double* p) { return __ldcg(p); }
__device__ __forceinline__ int ldcg(const int* p) { return __ldcg(p); }
__device__ __forceinline__ long long ldcg(const long long* p) { return __ldcg(p); }
// D(8x8) += A(8x4) * B(4x8) on the FP64 tensor cores (DMMA).  Lane l holds A[l/4][l%4], B[l%4][l/4] and
// C[l/4][2*(l%4) + {0,1}].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
#endif

}  // namespace pc
