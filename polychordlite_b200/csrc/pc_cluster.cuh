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

// The cube coordinates of the live points, staged in shared memory with an odd row stride (conflict-free when the
// lanes of a warp read the same coordinate of 32 consecutive points).  Falls back to global memory when the table
// does not fit (tab == nullptr).
__device__ __forceinline__ const double* stage_live_table(const double* live, int T, int D, int n, double* tab, int TS) {
    if (!tab) return nullptr;
    for (int e = threadIdx.x; e < n * D; e += blockDim.x) {
        const int j = e / D, k = e - j * D;
        tab[(size_t)j * TS + k] = live[(size_t)j * T + k];
    }
    __syncthreads();
    return tab;
}
__host__ __device__ inline size_t live_table_bytes(int n, int D) { return (size_t)n * (D | 1) * 8; }

// compute_knn restricted to a partition: for live slot i the KNN_K nearest slots j with part[j] == part[i]
// (itself included), ordered by (distance, slot) -- the order compute_knn's insertion rule produces.
// knn[i*KNN_K + t] = slot or -1.  One warp per point: every lane keeps the KNN_K best of its share of the live
// points in a sorted register list (one scan), then the 32 lists are merged head by head.
// Dynamic shared memory: W x D doubles (the warps' own points), then the live table when use_tab.
__global__ void pc_knn_kernel(const double* live, int T, int D, int n, const int* part, int* knn, int use_tab) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int TS = D | 1;
    double* xi = (double*)smem + (size_t)warp * D;
    {   // a negative part marks a point whose part is final (its list is not wanted): a CTA without work leaves before it
        // stages the live table -- the later rounds of a clustering pass only look at the pieces of the round before
        bool any = false;
        for (int i = blockIdx.x * W + warp; i < n; i += gridDim.x * W) any = any || part[i] >= 0;
        if (!__syncthreads_or(any ? 1 : 0)) return;
    }
    const double* tab = stage_live_table(live, T, D, n, use_tab ? (double*)smem + (size_t)W * D : nullptr, TS);
    const double* base = tab ? tab : live;
    const int stride = tab ? TS : T;
    for (int i = blockIdx.x * W + warp; i < n; i += gridDim.x * W) {
        __syncwarp();
        for (int k = lane; k < D; k += 32) xi[k] = live[(size_t)i * T + k];
        __syncwarp();
        const int pi = part[i];
        if (pi < 0) continue;
        double bd[KNN_K];
        int bj[KNN_K];
#pragma unroll
        for (int t = 0; t < KNN_K; ++t) { bd[t] = INFINITY; bj[t] = 0x7fffffff; }
        for (int j = lane; j < n; j += 32) {
            if (part[j] != pi) continue;
            double cd = cube_dist2(xi, base + (size_t)j * stride, D);
            int cj = j;
            if (cd < bd[KNN_K - 1] || (cd == bd[KNN_K - 1] && cj < bj[KNN_K - 1])) {
#pragma unroll
                for (int t = 0; t < KNN_K; ++t) {   // insertion: the displaced entry is carried down the list
                    const bool lt = cd < bd[t] || (cd == bd[t] && cj < bj[t]);
                    const double td = bd[t];
                    const int tj = bj[t];
                    bd[t] = lt ? cd : td; bj[t] = lt ? cj : tj;
                    cd = lt ? td : cd;   cj = lt ? tj : cj;
                }
            }
        }
        for (int t = 0; t < KNN_K; ++t) {
            double hd = bd[0];
            int hj = bj[0];
            const int mine = hj;
            warp_argmin(hd, hj);
            if (lane == 0) knn[(size_t)i * KNN_K + t] = (hj == 0x7fffffff) ? -1 : hj;
            if (mine == hj && hj != 0x7fffffff) {   // slots are unique: exactly one lane holds the winner
#pragma unroll
                for (int u = 0; u + 1 < KNN_K; ++u) { bd[u] = bd[u + 1]; bj[u] = bj[u + 1]; }
                bd[KNN_K - 1] = INFINITY; bj[KNN_K - 1] = 0x7fffffff;
            }
        }
    }
}

// identify_cluster (run_time_info.f90:913-949) for the phantoms: label of the nearest live point, ties to the
// lowest slot.  One warp per phantom record, four live points per lane in flight (independent accumulation chains;
// each distance still adds its dimensions in order).  Dynamic shared memory as for pc_knn_kernel.
__global__ void pc_identify_kernel(const double* live, int T, int D, int n, const int* lab, const double* ph,
                                   long long nph, int* phl, int use_tab) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int TS = D | 1;
    double* x = (double*)smem + (size_t)warp * D;
    const double* tab = stage_live_table(live, T, D, n, use_tab ? (double*)smem + (size_t)W * D : nullptr, TS);
    const double* base = tab ? tab : live;
    const int stride = tab ? TS : T;
    for (long long r = (long long)blockIdx.x * W + warp; r < nph; r += (long long)gridDim.x * W) {
        __syncwarp();
        for (int k = lane; k < D; k += 32) x[k] = ph[(size_t)r * T + k];
        __syncwarp();
        double bd = INFINITY;
        int bj = 0x7fffffff;
        for (int j0 = lane; j0 < n; j0 += 128) {
            const double* q0 = base + (size_t)j0 * stride;
            const double* q1 = base + (size_t)min(j0 + 32, n - 1) * stride;
            const double* q2 = base + (size_t)min(j0 + 64, n - 1) * stride;
            const double* q3 = base + (size_t)min(j0 + 96, n - 1) * stride;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            for (int k = 0; k < D; ++k) {
                const double xk = x[k];
                const double d0 = xk - q0[k], d1 = xk - q1[k], d2 = xk - q2[k], d3 = xk - q3[k];
                s0 = fma(d0, d0, s0); s1 = fma(d1, d1, s1); s2 = fma(d2, d2, s2); s3 = fma(d3, d3, s3);
            }
            // j ascends within a lane: the first minimum is the lowest slot
            if (s0 < bd) { bd = s0; bj = j0; }
            if (j0 + 32 < n && s1 < bd) { bd = s1; bj = j0 + 32; }
            if (j0 + 64 < n && s2 < bd) { bd = s2; bj = j0 + 64; }
            if (j0 + 96 < n && s3 < bd) { bd = s3; bj = j0 + 96; }
        }
        warp_argmin(bd, bj);
        if (lane == 0) phl[r] = lab[bj];
    }
}

// The same for nDims <= DCAP <= 32 with the roles swapped: a LANE owns a phantom (its coordinates in registers, zero
// beyond nDims) and the warp walks the live table together, so one broadcast read of a live coordinate serves 32
// distances -- 2.5x less shared-memory traffic than the warp-per-phantom form, which is what bounds that kernel.
// The table rows are padded to DCAP with zeros (a zero difference leaves the fused multiply-add chain unchanged, so
// every distance is still the same number).
//
// The scan is bound by the FP64 pipe (a subtraction and a fused multiply-add per dimension and live point), so it is
// run behind a filter that costs half of that: with |x - q|^2 = |x|^2 + |q|^2 - 2 x.q, the value s = |q|^2 - 2 x.q (one
// fused multiply-add per dimension, |q|^2 kept in the table) orders the live points as the distance does, up to its
// rounding error.  Only a live point whose s is within FILTER_TOL of the best distance found so far (minus |x|^2) can
// be the nearest one or tie with it; for those -- a few per phantom -- the distance itself is formed in the reference
// order and compared exactly, lowest slot first.  The label is therefore the one of the unfiltered scan, bit for bit.
// FILTER_TOL: cube coordinates lie in [0, 1], so |s|'s terms are below 2 and its rounding error below
// (DCAP + 2) * 3 DCAP * 2^-53 < 4e-13 for DCAP = 32.  Dynamic shared memory: n x TS doubles, TS = (DCAP + 1) | 1.
constexpr double FILTER_TOL = 1e-11;
template <int DCAP>
__global__ void pc_identify_lanes_kernel(const double* live, int T, int D, int n, const int* lab, const double* ph,
                                         long long nph, int* phl) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int TS = (DCAP + 1) | 1;
    double* tab = (double*)smem;
    for (int e = threadIdx.x; e < n * DCAP; e += blockDim.x) {
        const int j = e / DCAP, k = e - j * DCAP;
        tab[(size_t)j * TS + k] = k < D ? live[(size_t)j * T + k] : 0.0;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += blockDim.x) {   // |q|^2 behind the coordinates
        const double* q = tab + (size_t)j * TS;
        double qn = 0.0;
#pragma unroll
        for (int k = 0; k < DCAP; ++k) qn = fma(q[k], q[k], qn);
        tab[(size_t)j * TS + DCAP] = qn;
    }
    __syncthreads();
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < nph; r += (long long)gridDim.x * blockDim.x) {
        double x[DCAP], m2x[DCAP];
        double xn = 0.0;
#pragma unroll
        for (int k = 0; k < DCAP; ++k) {
            x[k] = k < D ? ph[(size_t)r * T + k] : 0.0;
            m2x[k] = -2.0 * x[k];
            xn = fma(x[k], x[k], xn);
        }
        double bd = INFINITY, lim = INFINITY;   // best distance so far; lim = bd - |x|^2 + FILTER_TOL in the filter's terms
        int bj = 0;
        auto exact = [&](int j) {   // the distance in the reference order (calculate.f90:94-109 / identify_cluster); strict <: lowest slot wins ties
            const double* q = tab + (size_t)j * TS;
            double sd = 0.0;
#pragma unroll
            for (int k = 0; k < DCAP; ++k) { const double d0 = x[k] - q[k]; sd = fma(d0, d0, sd); }
            if (sd < bd) { bd = sd; bj = j; lim = bd - xn + FILTER_TOL; }
        };
        int j = 0;
        for (; j + 4 <= n; j += 4) {   // four live points in flight
            const double* q = tab + (size_t)j * TS;
            double s0 = q[DCAP], s1 = q[TS + DCAP], s2 = q[2 * TS + DCAP], s3 = q[3 * TS + DCAP];
#pragma unroll
            for (int k = 0; k < DCAP; ++k) {
                s0 = fma(m2x[k], q[k], s0); s1 = fma(m2x[k], q[TS + k], s1);
                s2 = fma(m2x[k], q[2 * TS + k], s2); s3 = fma(m2x[k], q[3 * TS + k], s3);
            }
            if (s0 <= lim) exact(j);          // slots ascend: the first minimum is the lowest slot
            if (s1 <= lim) exact(j + 1);
            if (s2 <= lim) exact(j + 2);
            if (s3 <= lim) exact(j + 3);
        }
        for (; j < n; ++j) exact(j);
        phl[r] = lab[bj];
    }
}

// calculate_covmats (run_time_info.f90:601-641) per cluster, first half: CTA (p, c) forms the moment matrix
// M = sum z z^T, z = [x - pivot, 1], of the records of chunk c (live slots, then the phantom pool) labelled p, on
// the FP64 tensor cores exactly as phase U does (pc_run_kernel.cuh), warps combined in warp order, and writes it
// to cpart[(p * gridDim.y + c) * Dp8^2].
__global__ void __launch_bounds__(256) pc_cluster_moments_kernel(const double* live, const int* lab, int n, const double* ph,
                                                                 const int* phl, long long nph, int T, int D,
                                                                 const double* pivot, double* cpart) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = blockDim.x >> 5;
    const int p = blockIdx.x, c = blockIdx.y;
    const int Dp8 = (D + 1 + 7) & ~7, SX = Dp8 + 4, nt = Dp8 >> 3, ntl = nt * (nt + 1) / 2;
    const int Dpad = (D + 1) & ~1;
    double* s_piv = (double*)smem;                                  // Dpad
    double* s_M = s_piv + Dpad;                                     // Dp8 x Dp8, accumulated over the passes
    double* s_x = s_M + (size_t)Dp8 * Dp8 + (size_t)warp * U_BATCH * SX;  // per warp: U_BATCH x SX staged rows
    for (int e = tid; e < D; e += blockDim.x) s_piv[e] = pivot[e];
    for (int e = tid; e < Dp8 * Dp8; e += blockDim.x) s_M[e] = 0.0;
    __syncthreads();
    const int fr = lane >> 2, fk = lane & 3;
    const int JD = (D + 31) >> 5;
    const long long total = (long long)n + nph;                     // records: live slots, then the phantom pool
    long long chunk = (total + gridDim.y - 1) / gridDim.y;
    chunk = (chunk + W * 32 - 1) / (W * 32) * (W * 32);
    const long long r_begin = (long long)c * chunk, r_end = min(total, r_begin + chunk);
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
        for (long long base = r_begin + (long long)warp * 32; base < r_end; base += (long long)W * 32) {
            const long long r = base + lane;
            bool mine = false;
            if (r < r_end) mine = (r < n ? lab[r] : phl[r - n]) == p;
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
    double* out = cpart + ((size_t)p * gridDim.y + c) * Dp8 * Dp8;
    for (int e = tid; e < Dp8 * Dp8; e += blockDim.x) out[e] = s_M[e];
}

// second half: one CTA per cluster adds the chunks' matrices in chunk order, cov = S2/N - d d^T (S2 in the leading
// D x D block, S1 in column D, N at (D, D)), calc_cholesky (utils.F90:621-649).  Clusters with N <= D points keep the
// global factor (no covariance can be formed from them).
__global__ void pc_cluster_factor_kernel(const double* cpart, int nchunks, int D, const double* chol_glob, double* cchol,
                                         int* ccount) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, p = blockIdx.x;
    const int Dp8 = (D + 1 + 7) & ~7;
    double* s_M = (double*)smem;
    double* s_cov = s_M + (size_t)Dp8 * Dp8;
    double* s_L = s_cov + (size_t)D * D;
    for (int e = tid; e < Dp8 * Dp8; e += blockDim.x) {
        double s = 0.0;
        for (int c = 0; c < nchunks; ++c) s += cpart[((size_t)p * nchunks + c) * Dp8 * Dp8 + e];
        s_M[e] = s;
    }
    __syncthreads();
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
    if (tid < 32) warp_cholesky(s_cov, s_L, D);
    __syncthreads();
    for (int e = tid; e < D * D; e += blockDim.x) cchol[(size_t)p * D * D + e] = s_L[e];
}

constexpr int CLUSTER_CHUNKS = 16;
__host__ __device__ inline size_t cluster_moments_smem(int D, int W) {
    const int Dp8 = (D + 1 + 7) & ~7, SX = Dp8 + 4, Dpad = (D + 1) & ~1;
    return (size_t)(Dpad + Dp8 * Dp8 + W * U_BATCH * SX) * 8;
}
__host__ __device__ inline size_t cluster_factor_smem(int D) {
    const int Dp8 = (D + 1 + 7) & ~7;
    return (size_t)(Dp8 * Dp8 + 2 * D * D) * 8;
}

}  // namespace pc
