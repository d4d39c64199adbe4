// pc_cluster.cuh -- clustering of the live points at the update cadence.
//
// Reference: src/polychord/clustering.f90 (NN_clustering :15-97, compute_knn :134-174, do_clustering_k :100-130),
// calculate.f90:94-109 (similarity matrix), run_time_info.f90:913-949 (identify_cluster), :601-641 (per-cluster
// calculate_covmats).  Split of the work:
//   device  everything that is O(n^2 D) or O(n_phantom n D): the k-nearest-neighbour lists of the live points
//           (restricted to a given partition, so that the recursion of NN_clustering is served by the same kernel),
//           the nearest-live-point labels of the phantoms, and the per-cluster covariance (FP64 tensor-core
//           moments, as in phase U) with its Cholesky factor;
//   host    the irregular part on n x 10 integers: connected components of the mutual-neighbour graph for
//           n = 2..10 (union-find), the work list that replaces the recursion, canonical labels (pc_engine.cu).
// In the batched schedule the clusters steer the proposals only (the evidence is accumulated globally, which is
// exact for any posterior shape): a chain whitens its directions with the factor of its seed's cluster.
#pragma once
#include "pc_kernels.cuh"

namespace pc {


// squared distance in the accumulation order the oracle uses (dimension by dimension, fused multiply-add):
// neighbour ORDER must not depend on who computed the distances
__device__ __forceinline__ double cube_dist2(const double* a, const double* b, int D) {
    double s = 0.0;
    for (int k = 0; k < D; ++k) { const double d = a[k] - b[k]; s = fma(d, d, s); }
    return s;
}

// lexicographic (distance, index) minimum over the warp
__device__ __forceinline__ void warp_argmin(double& d, int& j) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double od = __shfl_xor_sync(FULL, d, o);
        const int oj = __shfl_xor_sync(FULL, j, o);
        if (od < d || (od == d && oj < j)) { d = od; j = oj; }
    }
}

// compute_knn restricted to a partition: for live slot i the KNN_K nearest slots j with part[j] == part[i]
// (itself included), ordered by (distance, slot) -- the order compute_knn's insertion rule produces.
// knn[i*KNN_K + t] = slot or -1.  One warp per point, one scan of the live points per neighbour.
__global__ void pc_knn_kernel(const double* live, int T, int D, int n, const int* part, int* knn) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int i = blockIdx.x * W + warp;
    double* xi = (double*)smem + (size_t)warp * D;
    if (i < n)
        for (int k = lane; k < D; k += 32) xi[k] = live[(size_t)i * T + k];
    __syncwarp();
    if (i >= n) return;
    const int pi = part[i];
    double pd = -1.0;
    int pj = -1;
    for (int t = 0; t < KNN_K; ++t) {
        double bd = INFINITY;
        int bj = 0x7fffffff;
        for (int j = lane; j < n; j += 32) {
            if (part[j] != pi) continue;
            const double d = cube_dist2(xi, live + (size_t)j * T, D);
            const bool after = d > pd || (d == pd && j > pj);           // not yet listed
            if (after && (d < bd || (d == bd && j < bj))) { bd = d; bj = j; }
        }
        warp_argmin(bd, bj);
        if (lane == 0) knn[(size_t)i * KNN_K + t] = (bj == 0x7fffffff) ? -1 : bj;
        if (bj == 0x7fffffff) {
            if (lane == 0) for (int u = t + 1; u < KNN_K; ++u) knn[(size_t)i * KNN_K + u] = -1;
            break;
        }
        pd = bd; pj = bj;
    }
}

// identify_cluster (run_time_info.f90:913-949) for the phantoms: label of the nearest live point, ties to the
// lowest slot.  One warp per phantom record.
__global__ void pc_identify_kernel(const double* live, int T, int D, int n, const int* lab, const double* ph,
                                   long long nph, int* phl) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    double* x = (double*)smem + (size_t)warp * D;
    for (long long r = (long long)blockIdx.x * W + warp; r < nph; r += (long long)gridDim.x * W) {
        __syncwarp();
        for (int k = lane; k < D; k += 32) x[k] = ph[(size_t)r * T + k];
        __syncwarp();
        double bd = INFINITY;
        int bj = 0x7fffffff;
        for (int j = lane; j < n; j += 32) {
            const double d = cube_dist2(x, live + (size_t)j * T, D);
            if (d < bd) { bd = d; bj = j; }   // j ascends within a lane: the first minimum is the lowest slot
        }
        warp_argmin(bd, bj);
        if (lane == 0) phl[r] = lab[bj];
    }
}

// calculate_covmats + calc_cholesky for ONE cluster per CTA: first and second moments of the cube coordinates of the
// live points and phantoms labelled blockIdx.x about the pivot (the global mean), formed on the FP64 tensor cores
// exactly as in phase U (pc_run_kernel.cuh), warps combined in warp order, then cov = S2/N - d d^T and its factor.
// Clusters with N <= D points keep the global factor (no covariance can be formed from them).
__global__ void __launch_bounds__(256) pc_cluster_cov_kernel(const double* live, const int* lab, int n, const double* ph,
                                                             const int* phl, long long nph, int T, int D,
                                                             const double* pivot, const double* chol_glob, double* cchol,
                                                             int* ccount) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = blockDim.x >> 5;
    const int p = blockIdx.x;
    const int Dp8 = (D + 1 + 7) & ~7, SX = Dp8 + 4, nt = Dp8 >> 3, ntl = nt * (nt + 1) / 2;
    const int Dpad = (D + 1) & ~1;
    double* s_piv = (double*)smem;                                  // Dpad
    double* s_M = s_piv + Dpad;                                     // Dp8 x Dp8, accumulated over the passes
    double* s_x = s_M + (size_t)Dp8 * Dp8 + (size_t)warp * U_BATCH * SX;  // per warp: U_BATCH x SX staged rows
    double* s_cov = s_M + (size_t)Dp8 * Dp8 + (size_t)W * U_BATCH * SX;   // D x D (+ D): the matrix, then its factor behind it
    double* s_L = s_cov + (size_t)D * D + Dpad;
    for (int e = tid; e < D; e += blockDim.x) s_piv[e] = pivot[e];
    for (int e = tid; e < Dp8 * Dp8; e += blockDim.x) s_M[e] = 0.0;
    __syncthreads();
    const int fr = lane >> 2, fk = lane & 3;
    const int JD = (D + 31) >> 5;
    const long long total = (long long)n + nph;                     // records: live slots, then the phantom pool
    const int passes = (ntl + COV_TPP - 1) / COV_TPP;
    for (int pass = 0; pass < passes; ++pass) {
        double c0[COV_TPP], c1[COV_TPP];
        int tl[COV_TPP];
#pragma unroll
        for (int q = 0; q < COV_TPP; ++q) {
            c0[q] = c1[q] = 0.0;
            const int idx = pass * COV_TPP + q;
            int tj = 0;
            while ((tj + 1) * (tj + 2) / 2 <= idx) ++tj;
            const int ti = idx - tj * (tj + 1) / 2;
            tl[q] = idx < ntl ? ((ti << 8) | tj) : -1;
        }
        for (int e = lane; e < U_BATCH * SX; e += 32) s_x[e] = 0.0;
        __syncwarp();
        for (long long base = (long long)warp * 32; base < total; base += (long long)W * 32) {
            const long long r = base + lane;
            bool mine = false;
            if (r < total) mine = (r < n ? lab[r] : phl[r - n]) == p;
            unsigned rem = __ballot_sync(FULL, mine);
            while (rem) {
                const double* rp[U_BATCH];
                int nb = 0;
#pragma unroll
                for (int b2 = 0; b2 < U_BATCH; ++b2) {
                    rp[b2] = live;
                    if (rem) {
                        const int bit = __ffs(rem) - 1;
                        rem &= rem - 1;
                        const long long rr = base + bit;
                        rp[b2] = rr < n ? live + (size_t)rr * T : ph + (size_t)(rr - n) * T;
                        ++nb;
                    }
                }
                __syncwarp();
                for (int j = 0; j < JD; ++j) {
                    const int e = lane + 32 * j;
                    double v[U_BATCH];
#pragma unroll
                    for (int b2 = 0; b2 < U_BATCH; ++b2) v[b2] = (b2 < nb && e < D) ? __ldcg(rp[b2] + e) : 0.0;
                    const double pv = e < D ? s_piv[e] : 0.0;
#pragma unroll
                    for (int b2 = 0; b2 < U_BATCH; ++b2)
                        if (e <= D) s_x[b2 * SX + e] = b2 < nb ? (e < D ? v[b2] - pv : 1.0) : 0.0;
                }
                if ((D & 31) == 0 && lane < U_BATCH) s_x[lane * SX + D] = lane < nb ? 1.0 : 0.0;
                __syncwarp();
#pragma unroll
                for (int ks = 0; ks < U_BATCH / 4; ++ks) {
                    const double* row = s_x + (4 * ks + fk) * SX + fr;
#pragma unroll
                    for (int q = 0; q < COV_TPP; ++q)
                        if (tl[q] >= 0) dmma884(c0[q], c1[q], row[(tl[q] >> 8) << 3], row[(tl[q] & 0xff) << 3]);
                }
            }
        }
        __syncthreads();
        for (int w = 0; w < W; ++w) {   // warps add their tiles in warp order (deterministic)
            if (warp == w) {
#pragma unroll
                for (int q = 0; q < COV_TPP; ++q)
                    if (tl[q] >= 0) {
                        const int a2 = ((tl[q] >> 8) << 3) + fr, b2 = ((tl[q] & 0xff) << 3) + 2 * fk;
                        s_M[a2 * Dp8 + b2] += c0[q];
                        s_M[a2 * Dp8 + b2 + 1] += c1[q];
                    }
            }
            __syncthreads();
        }
    }
    // M = sum z z^T of z = [x - pivot, 1]: S2 in the leading D x D block (upper triangle), S1 in column D, N at (D, D)
    const double N = s_M[D * Dp8 + D];
    if (tid == 0) ccount[p] = (int)N;
    if (!(N > (double)D)) {
        for (int e = tid; e < D * D; e += blockDim.x) cchol[(size_t)p * D * D + e] = chol_glob[e];
        return;
    }
    for (int idx = tid; idx < D * D; idx += blockDim.x) {
        const int a = idx % D, b = idx / D;
        const int lo = min(a, b), hi = max(a, b);
        s_cov[idx] = s_M[lo * Dp8 + hi] / N - (s_M[a * Dp8 + D] / N) * (s_M[b * Dp8 + D] / N);
    }
    __syncthreads();
    if (tid < 32) warp_cholesky(s_cov, s_L, D);   // utils.F90:621-649, with the sqrt(trace) * I fallback
    __syncthreads();
    for (int e = tid; e < D * D; e += blockDim.x) cchol[(size_t)p * D * D + e] = s_L[e];
}

__host__ __device__ inline size_t cluster_cov_smem(int D, int W) {
    const int Dp8 = (D + 1 + 7) & ~7, SX = Dp8 + 4, Dpad = (D + 1) & ~1;
    return (size_t)(Dpad + Dp8 * Dp8 + W * U_BATCH * SX + D * D + Dpad + D * D) * 8;
}

}  // namespace pc
