// pc_kernels.cuh -- the sm_100a kernels of the nested-sampling engine.
//
// One PERSISTENT kernel (pc_run_kernel) advances whole nested-sampling runs on the device.
// Each run owns a group of CTAs; inside a group every warp owns one slice-sampling chain at
// a time (reference: src/polychord/chordal_sampling.f90:7-92 SliceSampling), the nDims-long
// point lives in registers (lane r <-> dimension r, r+32, ...), directions are staged in
// shared memory, and every likelihood evaluation ends in a warp-shuffle reduction whose
// result drives the step-out / shrink decisions (chordal_sampling.f90:163-273).
//
// A generation (the batched form of one iteration of nested_sampling.F90:239-374):
//   phase S (CTA 0 of the group)  termination test (run_time_info.f90:683-709), bitonic sort of
//                                 (logL, slot), K lowest die: evidence recurrences
//                                 (run_time_info.f90:211-296) as block-wide log-space scans
//   phase C (all warps)           chain k: GenerateSeed (generate.F90:19-55) from the survivors,
//                                 directions (random_utils.F90:381-437 + chordal_sampling.f90:94-145),
//                                 Cholesky whitening (:73), R slice steps; babies stream to the
//                                 phantom pool, the last baby replaces the dead slot
//   phase U (all warps, at the    clean_phantoms (run_time_info.f90:820-877) as a stable
//            update cadence)      compaction fused with calculate_covmats (:601-641); CTA 0
//                                 finishes with calc_cholesky (utils.F90:621-649)
// Groups synchronise with their own global-memory barrier, so many independent runs
// (an ensemble) advance concurrently inside one launch.
#pragma once
#include "pc_chain.cuh"
#include "pc_dense.cuh"

namespace pc {

enum Status : int { ST_RUNNING = 0, ST_DONE = 1, ST_NEED_DEAD = 2, ST_NEED_PHANTOM = 3, ST_DUMP = 4, ST_HOSTCHAINS = 5, ST_CLUSTER = 6, ST_NEED_BOOST = 7, ST_ERROR = -1 };

constexpr double NEG_BIG = -1e300;  // log(0) stand-in that survives additions without NaN
constexpr int COV_TPP = 10;         // 8x8 tiles of the moment matrix a warp accumulates per pass of phase U (2 registers each)
constexpr int U_TILE = 32;          // sizes the per-tile arrays of phase U: a tile is at least one warp's 32 records
constexpr int KNN_K = 10;           // clustering.f90:44: "10 degrees of separation"
constexpr int MAX_CLUSTERS = 256;   // clusters with a factor of their own; further labels share the last one, which keeps the global factor
constexpr int U_BATCH = 8;          // records a warp of phase U keeps in flight
constexpr int MAX_DYN = 16;         // entries of a dynamic-nlive schedule

// Mutable per-run scalars (device global memory; the host reads them between launches).
struct DevRun {
    double logZ, logZ2, logX, logZX, logXX;  // single-cluster evidence state (log space)
    double logX_last_update;
    double Lstar;     // contour of the generation in flight
    double cov_N;     // points that entered the last covariance
    long long ndead, nlike, nchains, ngen, nupdates, nfail, nslices;
    long long nphantom;       // records in the current phantom pool
    long long ph_kept;        // survivors counted by the last phase U (this rank's pool)
    long long nph_glob;       // sharded run: phantoms of all ranks (replicated arithmetic, used for the capacity test)
    unsigned long long xepoch; // sharded run: cross-GPU barriers passed
    long long ndead_base;     // ndead before the generation in flight
    long long nph_base;       // nphantom before the generation in flight
    long long nchains_base;
    long long init_attempts;
    int cur_pool;
    int K;
    int do_update;
    int update_pending;  // phase U ran, CTA 0 still has to finish covariance/Cholesky
    int status;
    int initialised;
    int init_need;
    int chol_fallback;   // number of calc_cholesky identity fallbacks
    int order_valid;     // rb.order + order_off holds the live slots sorted by (logL, slot) as of the last phase S
    int order_off;       // 0 or n: which half of rb.order is current
    int ncl;             // clusters found at the last update (1: the global factor is used)
    int dump_pub;        // asynchronous dumper hand-over: dumps published so far (the snapshot of dump s lies in half s & 1 of live_snap)
    int host_resume;     // host-callback runs: the chains of the generation in flight were run by the host loop
    int n;               // live points now (dynamic nlive, nprior, failed births: it moves; KParams::n is the target)
    // the generation in flight: n_gen live points at its start, K of them die, B chains are born
    int n_gen, B;
    int holes_due;       // it leaves empty live slots (B != K): settle_generation closes them; failed births add to that
    int trimmed;         // the nprior > nlive trim (nested_sampling.F90:201-203) has been done
    int fail_run;        // failed births in a row (nested_sampling.F90:315-319)
    int stop_nfail;      // ... exceeded nfail: the run ends with the reference's warning (:407-409)
    unsigned int nfail_gen;  // failed births of the generation in flight (chain warps add, settle_generation clears)
    // SM-clock cycle counters of the phases (thread 0 of CTA 0; chain phases: warp 0 of the first chain CTA)
    long long cyc_wait, cyc_S, cyc_fin, cyc_U, cyc_prep, cyc_white, cyc_slice, cyc_total;
    // finer cycle counters, printed (in ms) when PC_DEBUG is set: [0] slice loop and [1] derived parameters of the
    // representative chain warp; phase U of CTA 0: [2] pass A, [3] barrier, [4] pass B ([12] its set-up, [13] its tile
    // loop, [15] the warps' combination), [5] closing barrier; phase S1: [6] keys, [7] termination test, [8] sort +
    // merge, [9] publication; the first chain CTA: [10] wait at the generation barrier, [11] release -> first slice
    long long dbg[24];
    // what every warp needs at the start of a generation, in one 64-byte line (written by phase S1, read with one
    // coalesced load per warp): [0] Lstar (bits), [1] ndead_base, [2] nph_base, [3] nchains_base, [4] ngen,
    // [5] K | do_update << 32, [6] order_off | cur_pool << 32, [7] ncl | nupdates << 32, [8] n_gen | B << 32,
    // [9] nfail at the start of the generation (a generation whose chains leave it unchanged had no failed birth)
    unsigned long long pub[12];
    unsigned int bar;    // group barrier, one arrival per CTA (monotonic)
    unsigned int wbar;   // chains-done barrier, one arrival per warp (monotonic)
    unsigned int dbar;   // phase D done, one arrival per ranking CTA (monotonic); only CTA 0 waits for it
    unsigned int snap_arr; // ... CTAs that have written their share of the next dump's live snapshot (monotonic)
    // boost_posterior (clean_phantoms, run_time_info.f90:820-877): phantoms promoted to posterior samples so far
    // (rb.boost rows), and ndead at the last update -- the deaths after it are the posterior stack a removed
    // phantom takes its weight from
    unsigned long long nboost;
    long long ndead_upd;
};

// Control block in mapped pinned host memory: the run kernel publishes a dump (run state + a snapshot of the
// live points) at every update and keeps sampling; the host thread that called the engine picks it up, copies
// the new dead rows on a second stream and calls the user's dumper.  One dump may be outstanding.
struct HostCtl {
    unsigned long long dump_seq;   // device -> host: dumps published
    unsigned long long ack_seq;    // host -> device: dumps consumed (the live snapshot may be overwritten)
    long long ndead;               // state at the published dump
    long long nlike;
    double logZ, logZ2;
    int abort;                     // host -> device: stop waiting (the dumper threw)
    int nlive;                     // live points in the snapshot
};

struct RunBuf {
    DevRun* st;
    double* live;      // n x T records
    double* live_snap; // 2 x nmax x T: the live points at the published dumps, by dump parity (every CTA writes its share)
    HostCtl* ctl;      // mapped host memory, or null (no dumper)
    int* order;        // 2 x n: live slots sorted by (logL, slot), ping-pong (DevRun::order_off)
    double* okey;      // 2 x n: the logL of those slots, same order and ping-pong
    double* dead;      // cap_dead x T
    double* logw;      // cap_dead
    double* ph[2];     // phantom pools (ping-pong), cap_ph x T each
    double* chol;      // D x D column-major
    double* cov;       // D x D column-major
    double* partial;   // per CTA: [0]=count, [1..D]=sum x, then ntri covariance partials
    double* gsum;      // [0] surviving phantoms of all ranks, [2..2+D) mean of live + phantom cube coordinates, [2+D..2+2D) pivot of the next update
    long long* pcount; // survivor count of each phantom tile (phase U; a tile is one record per thread of a CTA)
    unsigned int* pmask; // keep mask of each 32-record segment of the phantom pool (phase U, pass A -> pass B)
    double* nh;        // global direction scratch (used when the directions do not fit in smem)
    int* lab;          // clustering: label of every live slot
    int* phl[2];       // clustering: label of every phantom record (compacted with the pools)
    double* cchol;     // clustering: Cholesky factor per label, MAX_CLUSTERS x D x D
    double* boost;     // boost_posterior: cap_boost x (D + P + 2) rows [theta, phi, birth, logL] of promoted phantoms
    unsigned long long* boost_win;  // ... and the window of dead indices each was removed against: first << 32 | end
    long long cap_boost;
    long long cap_dead, cap_ph;
    int* deadlab;      // clustering: cluster label every dead point carried when it died (attributed local evidences); else null
    int* deadn;        // clustering: live points before each death (the n of its update_evidence); else null
    int* cfail;        // per chain of the generation in flight: 1 when its last baby is not above the contour (a failed birth)
    double* bkey;      // per chain of the generation in flight: logL of its last baby (phase D ranks the babies from here)
    double* dpart;     // per CTA: (max, sum of exp(logL - max)) over the live points the CTA ranked in phase D (termination test)
    unsigned int seed;
    int pad;
};

// Sharded run: one process per GPU, every rank keeps the whole run state and does
// the (deterministic) bookkeeping redundantly, the chains of a generation are dealt k % world, and the ranks
// exchange through peer-mapped memory over NVLink inside the persistent kernel:
//   xin    the last baby of every chain is stored into every rank's incoming buffer (2 x batch_K x T, by
//          generation parity) and scattered into the replicated live array after a cross-GPU barrier;
//   xpart  the covariance statistics (count, sum x | centred outer products) of each rank's phantoms are
//          all-reduced by storing them into every rank's slot and summing in rank order;
//   xbar   one monotonic counter per rank: a barrier adds `world` to each of them.
constexpr int MAX_RANKS = 8;
//   xrun   the logL of the last baby of every chain, stored by the chain into every rank's key buffer (2 x batch_K
//          doubles, by generation parity): phase D ranks the babies from there on every rank.
struct Shard {
    int rank, world;
    long long xstride;                 // doubles per rank slot in xpart
    int kr, pad;                       // ceil(batch_K / world) (sizes the key buffers)
    unsigned int* xbar[MAX_RANKS];
    double* xin[MAX_RANKS];
    double* xpart[MAX_RANKS];
    double* xrun[MAX_RANKS];
};

struct KParams {
    ChainParams cp;              // D, P, T, R, LD, likelihood constants
    Shard sh;                    // sh.world <= 1: a run on one GPU
    int n, batch_K;              // target number of live points (settings%nlive); deaths per generation
    int nmax, n0;                // capacity of the live arrays; live points the run starts with (nprior, cube_samples)
    int nfail;                   // failed births in a row the run tolerates (settings%nfail; <= 0: nlive)
    // dynamic nlive (settings%loglikes / settings%nlives, sorted by loglike: settings.f90:234-235): above the contour
    // dyn_loglikes[i] the target is dyn_nlives[i] (run_time_info.f90:766-771)
    int dyn_m;
    int dyn_nlives[MAX_DYN];
    double dyn_loglikes[MAX_DYN];
    int use_prec, max_ndead;
    int ctas_per_run, warps_per_cta;
    int chain_cta0;              // first CTA of a run's group that runs chains (1: CTA 0 only keeps the books)
    int paired;                  // 1: warps w >= W/2 prepare the chains of warp w - W/2 (a run alone on the device)
    int nh_in_smem, want_dump;
    int dense;                   // 1: the dense chain phase (pc_dense.cuh): one chain per point group, rb.nh holds the slice records
    int backoff;                 // ns a spinning thread sleeps between polls (ensembles: the SM is shared with other runs' warps); 0: spin
    int clustering;              // 1: do_clustering -- the kernel leaves at every update for the clustering pass (pc_cluster.cuh)
    int host_like;               // 1: likelihood/prior are host callbacks -- the kernel leaves before the chain phase (pc_hostchain.cuh)
    int live_given;              // 1: the host uploaded the initial live points (host callbacks, or the caller's cube_samples)
    int ntri, cov_passes, partial_stride;
    int off_like, off_warp, warp_bytes;  // shared-memory byte offsets
    int off_dkeys;               // phase D's key area behind the per-warp areas (nmax + 2 * warps + 2 doubles); 0: phase S orders the live points on CTA 0
    int u_bulk;                  // phase U: the per-warp areas hold the staging ring of the bulk-copy record stream (2 * U_BATCH records behind the staged rows)
    double log_prec, log_comp;
    double boost_thin;           // RTI%thin_posterior (generate.F90:311-316) when posterior files are written, else 0
    const double* like_params;   // gaussian: mu[D], 1/sigma[D]; corr: mu[D], invcov[D*D]
    const double* prior_params;  // lo[D], hi-lo[D]
    RunBuf* runs;
};

// shared-memory layout of phase S (CTA 0), over nmax live points and batches of up to kb births
struct SmemS {
    double* sc;     // 64 doubles of reduction scratch
    int* aval;      // nmax: slots of the survivors in order (merge path)
    double* akey;   // nmax: keys of the survivors in order (merge path) / np2: all keys (full sort)
    double* bkey;   // npB: keys of the new babies
    int* bval;      // npB (merge) / np2 (full sort: slot of every key)
};
__host__ __device__ inline int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }
__host__ __device__ inline size_t smem_S_bytes(int nmax, int batch_K) {
    const int np2 = next_pow2(nmax), npB = next_pow2(batch_K);
    size_t full = (size_t)np2 * 12;
    size_t merge = (size_t)nmax * 8 + (size_t)npB * 12 + (size_t)((nmax + 1) & ~1) * 4;
    size_t settle = (size_t)nmax * 9 + 16 + (size_t)(2 * batch_K + 4) * 4;   // settle_generation: flags, fail flags, two slot lists
    size_t most = full > merge ? full : merge;
    return 64 * 8 + (most > settle ? most : settle) + 16;
}
__device__ inline SmemS smem_S(unsigned char* base, int nmax, int batch_K) {
    SmemS m;
    const int npB = next_pow2(batch_K);
    m.sc = (double*)base;
    m.akey = m.sc + 64;
    m.bkey = m.akey + nmax;        // merge layout
    m.bval = (int*)(m.bkey + npB);
    m.aval = m.bval + npB;
    return m;
}

// ------------------------------------------------------------------------------------------
// group barrier (same fence / atomic / fence pattern cooperative groups uses for grid.sync)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void group_sync(unsigned int* bar, unsigned int G, int backoff = 0) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned int t = atomicAdd(bar, 1u);
        unsigned int target = (t / G + 1u) * G;
        unsigned int v;
        for (;;) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
            if ((int)(v - target) >= 0) break;
            if (backoff) { __nanosleep(backoff); if (backoff < 4096) backoff <<= 1; }   // an ensemble shares the SM: do not burn issue slots polling
        }
        __threadfence();
    }
    __syncthreads();
}

// Cross-GPU barrier of a sharded run, called by ONE thread per rank: publish everything written so far
// (system scope), add one to every rank's counter, wait until this rank's counter shows `world` arrivals for
// the new epoch.  Returns false on time-out (a peer died): the caller stops the run.
__device__ inline bool xgpu_barrier(const Shard& sh, DevRun* st) {
    const unsigned long long epoch = ++st->xepoch;
    __threadfence_system();
    for (int q = 0; q < sh.world; ++q) atomicAdd_system(sh.xbar[q], 1u);
    const unsigned int target = (unsigned int)(epoch * (unsigned long long)sh.world);
    const long long t0 = clock64();
    unsigned int v;
    do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(sh.xbar[sh.rank]) : "memory");
        if (clock64() - t0 > 60000000000LL) return false;  // ~30 s
    } while ((int)(v - target) < 0);
    __threadfence_system();
    return true;
}

// Per-warp split barrier: a warp arrives as soon as its own chains are written, does other work
// (prep of its next chain), and only the warps that need everybody's results wait.
__device__ __forceinline__ unsigned int warp_arrive(unsigned int* wbar, unsigned int GW) {
    __syncwarp();
    unsigned int target = 0;
    if ((threadIdx.x & 31) == 0) {
        __threadfence();
        unsigned int t = atomicAdd(wbar, 1u);
        target = (t / GW + 1u) * GW;
    }
    return __shfl_sync(FULL, target, 0);
}
__device__ __forceinline__ void warp_wait(unsigned int* wbar, unsigned int target, int backoff = 0) {
    if ((threadIdx.x & 31) == 0) {
        unsigned int v;
        for (;;) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(wbar) : "memory");
            if ((int)(v - target) >= 0) break;
            if (backoff) { __nanosleep(backoff); if (backoff < 4096) backoff <<= 1; }
        }
        __threadfence();
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------
// block-wide helpers for phase S (deterministic: fixed combination order)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum(double v, double* sc) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_sum(v);
    if (lane == 0) sc[w] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < nw; ++i) t += sc[i];
    __syncthreads();
    return t;
}
__device__ __forceinline__ double block_max(double v, double* sc) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_max(v);
    if (lane == 0) sc[w] = v;
    __syncthreads();
    double t = sc[0];
    for (int i = 1; i < nw; ++i) t = fmax(t, sc[i]);
    __syncthreads();
    return t;
}
// exclusive prefix sum of ints; *total receives the block total
__device__ __forceinline__ int block_exscan_int(int v, int* total, double* sc) {
    int* si = (int*)sc;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) si[w] = inc;
    __syncthreads();
    int pre = 0, tot = 0;
    for (int i = 0; i < nw; ++i) {
        if (i < w) pre += si[i];
        tot += si[i];
    }
    __syncthreads();
    *total = tot;
    return pre + inc - v;
}

// log( exp(init) + sum_threads exp(v) )
__device__ __forceinline__ double block_lse(double v, double init, double* sc) {
    double m = fmax(block_max(v, sc), init);
    double s = block_sum(exp(v - m), sc) + exp(init - m);
    return m + log(s);
}
// exclusive prefix sum; *total receives the block total
__device__ __forceinline__ double block_exscan_sum(double v, double* total, double* sc) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double t = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) sc[w] = inc;
    double ex = __shfl_up_sync(FULL, inc, 1);
    if (lane == 0) ex = 0.0;
    __syncthreads();
    double pre = 0.0, tot = 0.0;
    for (int i = 0; i < nw; ++i) {
        if (i < w) pre += sc[i];
        tot += sc[i];
    }
    __syncthreads();
    *total = tot;
    return pre + ex;
}
// exclusive scan of affine maps x -> logaddexp(x + a, b), applied in thread order.
__device__ __forceinline__ void affine_combine(double& a, double& b, double ea, double eb) {
    // (ea,eb) earlier, (a,b) later:  x -> logaddexp(logaddexp(x+ea, eb) + a, b)
    b = logaddexp(eb + a, b);
    a = ea + a;
}
__device__ __forceinline__ void block_exscan_affine(double a, double b, double& exa, double& exb, double& tota,
                                                    double& totb, double* sc) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double ia = a, ib = b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double pa = __shfl_up_sync(FULL, ia, o), pb = __shfl_up_sync(FULL, ib, o);
        if (lane >= o) affine_combine(ia, ib, pa, pb);
    }
    if (lane == 31) { sc[2 * w] = ia; sc[2 * w + 1] = ib; }
    double pa = __shfl_up_sync(FULL, ia, 1), pb = __shfl_up_sync(FULL, ib, 1);
    if (lane == 0) { pa = 0.0; pb = NEG_BIG; }
    __syncthreads();
    double wa = 0.0, wb = NEG_BIG, ta = 0.0, tb = NEG_BIG;
    for (int i = 0; i < nw; ++i) {
        double ca = sc[2 * i], cb = sc[2 * i + 1];
        if (i < w) { double xa = ca, xb = cb; affine_combine(xa, xb, wa, wb); wa = xa; wb = xb; }
        { double xa = ca, xb = cb; affine_combine(xa, xb, ta, tb); ta = xa; tb = xb; }
    }
    __syncthreads();
    // exclusive = (warps before) then (lanes before)
    affine_combine(pa, pb, wa, wb);
    exa = pa; exb = pb; tota = ta; totb = tb;
}

// The same sort when every thread holds one element (np2 <= blockDim.x): the exchanges at distance < 32 are warp
// shuffles, only the ones across warps go through shared memory (6 instead of 36 barrier-separated steps at 256).
__device__ inline void block_sort_small(double* key, int* val, int np2) {
    const int i = threadIdx.x;
    double mk = i < np2 ? key[i] : INFINITY;
    int mv = i < np2 ? val[i] : 0x7fffffff;
    __syncthreads();
    for (int k = 2; k <= np2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            double ok;
            int ov;
            if (j < 32) {
                ok = __shfl_xor_sync(FULL, mk, j);
                ov = __shfl_xor_sync(FULL, mv, j);
            } else {
                if (i < np2) { key[i] = mk; val[i] = mv; }
                __syncthreads();
                ok = i < np2 ? key[i ^ j] : INFINITY;
                ov = i < np2 ? val[i ^ j] : 0x7fffffff;
                __syncthreads();
            }
            const bool keep_min = ((i & k) == 0) == ((i & j) == 0);
            const bool other_less = ok < mk || (ok == mk && ov < mv);
            if (keep_min == other_less) { mk = ok; mv = ov; }
        }
    }
    if (i < np2) { key[i] = mk; val[i] = mv; }
    __syncthreads();
}

// The same network with TWO elements per thread (np2 = 2 * blockDim.x): thread i holds elements i and i + blockDim.x, so
// the exchange at distance blockDim.x is a compare-exchange of its own two registers, distances below 32 are warp
// shuffles and only distances 32 .. blockDim.x / 2 go through shared memory.
__device__ inline void block_sort_pair(double* key, int* val, int np2) {
    const int i0 = threadIdx.x, H = blockDim.x;   // np2 == 2 * H
    double mk[2] = {key[i0], key[i0 + H]};
    int mv[2] = {val[i0], val[i0 + H]};
    __syncthreads();
    auto less = [](double ka, int va, double kb, int vb) { return ka < kb || (ka == kb && va < vb); };
    for (int k = 2; k <= np2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j == H) {   // partner = my other element (only at k == np2: ascending)
                if (less(mk[1], mv[1], mk[0], mv[0])) {
                    const double tk = mk[0]; mk[0] = mk[1]; mk[1] = tk;
                    const int tv = mv[0]; mv[0] = mv[1]; mv[1] = tv;
                }
                continue;
            }
            double ok[2];
            int ov[2];
            if (j < 32) {
#pragma unroll
                for (int h = 0; h < 2; ++h) { ok[h] = __shfl_xor_sync(FULL, mk[h], j); ov[h] = __shfl_xor_sync(FULL, mv[h], j); }
            } else {
                key[i0] = mk[0]; key[i0 + H] = mk[1]; val[i0] = mv[0]; val[i0 + H] = mv[1];
                __syncthreads();
#pragma unroll
                for (int h = 0; h < 2; ++h) { ok[h] = key[(i0 + h * H) ^ j]; ov[h] = val[(i0 + h * H) ^ j]; }
                __syncthreads();
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = i0 + h * H;
                const bool keep_min = ((i & k) == 0) == ((i & j) == 0);
                const bool other_less = less(ok[h], ov[h], mk[h], mv[h]);
                if (keep_min == other_less) { mk[h] = ok[h]; mv[h] = ov[h]; }
            }
        }
    }
    key[i0] = mk[0]; key[i0 + H] = mk[1]; val[i0] = mv[0]; val[i0 + H] = mv[1];
    __syncthreads();
}

// bitonic sort of (key, val) ascending by key then val; np2 a power of two
__device__ inline void block_sort(double* key, int* val, int np2) {
    for (int k = 2; k <= np2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < np2; i += blockDim.x) {
                int ixj = i ^ j;
                if (ixj > i) {
                    double ka = key[i], kb = key[ixj];
                    int va = val[i], vb = val[ixj];
                    bool gt = (ka > kb) || (ka == kb && va > vb);
                    bool up = ((i & k) == 0);
                    if (gt == up) { key[i] = kb; key[ixj] = ka; val[i] = vb; val[ixj] = va; }
                }
            }
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------------
// Evidence recurrences for `count` consecutive deaths with live counts n_start, n_start-1, ...
// (run_time_info.f90:211-296 applied as in the final kill-off nested_sampling.F90:381-384).
// With n_j = n_start - j the death j does
//     Z  <- Z  (+) X L / (n+1)                          Z2 <- Z2 (+) 2 ZX L / (n+1) (+) 2 XX L^2 / ((n+1)(n+2))
//     ZX <- ZX n/(n+1) (+) XX L n / ((n+1)(n+2))        X  <- X n/(n+1)             XX <- XX n/(n+2)
// in log space ((+) = logaddexp).  X and XX are products that telescope, ZX is a first-order linear recurrence
// (a composition of affine maps x -> logaddexp(x + a, b)), Z and Z2 are log-sum-exp reductions.  Two levels: a thread
// owns a CONTIGUOUS run of deaths and walks it sequentially (one new logarithm per death: log(n_j + 1) and
// log(n_j + 2) are the previous death's log(n_j) and log(n_j + 1)); its start values of X and XX come in closed form,
// its start value of ZX from ONE block-wide scan of the threads' composed maps.  The work of a death no longer
// scales with block-wide barriers (16 rounds of five scans each for 4096 deaths before; one scan now).
// skey: ascending logL of the dying points.  KEYS_GLOBAL: skey is global memory other CTAs wrote (phase D): read past L1.
// ------------------------------------------------------------------------------------------
template <bool KEYS_GLOBAL = false>
__device__ inline void evidence_deaths(DevRun* st, const double* skey, int count, int n_start, double* logw_out,
                                       double* sc, int* n_out = nullptr) {
    const double LOG2 = 0.69314718055994530942;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const double lX0 = st->logX, lXX0 = st->logXX, lZX0 = st->logZX, lZ0 = st->logZ, lZ20 = st->logZ2;
    __syncthreads();
    const int c = (count + nthr - 1) / nthr;
    const int j0 = min(count, tid * c), j1 = min(count, j0 + c);
    auto keyat = [&](int j) -> double { return KEYS_GLOBAL ? __ldcg(skey + j) : skey[j]; };
    // X and XX before death j0 (products of n/(n+1), n/(n+2) over the deaths before it, telescoped)
    const double nA = (double)(n_start - j0), nB = (double)n_start;
    const double lgA1 = log(nA + 1.0), lgA2 = log(nA + 2.0), lgB1 = log(nB + 1.0), lgB2 = log(nB + 2.0);
    const double lXs = lX0 + (lgA1 - lgB1), lXXs = lXX0 + ((lgA1 + lgA2) - (lgB1 + lgB2));
    // pass 1: the thread's composed ZX map
    double ma = 0.0, mb = NEG_BIG;
    {
        double l1 = lgA1, l2 = lgA2, lXXb = lXXs;
        for (int j = j0; j < j1; ++j) {
            const double l0n = log((double)(n_start - j));
            const double a = l0n - l1, b = lXXb + keyat(j) + l0n - l1 - l2;
            // (ma, mb) earlier, (a, b) later
            mb = logaddexp(mb + a, b);
            ma = ma + a;
            lXXb += l0n - l2;
            l2 = l1; l1 = l0n;
        }
    }
    double exa, exb, tota, totb;
    block_exscan_affine(ma, mb, exa, exb, tota, totb, sc);
    // pass 2: the thread's deaths from its start state
    double zm = NEG_BIG, zs = 0.0, z2m = NEG_BIG, z2s = 0.0;   // running (max, sum of exp(. - max)) of the Z and Z2 terms
    {
        double l1 = lgA1, l2 = lgA2, lXb = lXs, lXXb = lXXs, ZXb = logaddexp(lZX0 + exa, exb);
        auto push = [](double& m, double& s2, double t) {
            if (t > m) { s2 = s2 * exp(m - t) + 1.0; m = t; } else s2 += exp(t - m);
        };
        for (int j = j0; j < j1; ++j) {
            const double l0n = log((double)(n_start - j));
            const double L = keyat(j);
            logw_out[j] = lXb - l1;
            if (n_out) n_out[j] = n_start - j;
            push(zm, zs, lXb + L - l1);
            push(z2m, z2s, logaddexp(LOG2 + ZXb + L - l1, LOG2 + lXXb + 2.0 * L - l1 - l2));
            ZXb = logaddexp(ZXb + (l0n - l1), lXXb + L + l0n - l1 - l2);
            lXb += l0n - l1;
            lXXb += l0n - l2;
            l2 = l1; l1 = l0n;
        }
    }
    // Z and Z2: log-sum-exp over the threads' partials and the start values
    const double gm = fmax(block_max(zm, sc), lZ0), gm2 = fmax(block_max(z2m, sc), lZ20);
    const double sZ = block_sum(zs > 0.0 ? zs * exp(zm - gm) : 0.0, sc) + exp(lZ0 - gm);
    const double sZ2 = block_sum(z2s > 0.0 ? z2s * exp(z2m - gm2) : 0.0, sc) + exp(lZ20 - gm2);
    if (tid == 0) {
        const double nE = (double)(n_start - count);
        st->logX = lX0 + (log(nE + 1.0) - lgB1);
        st->logXX = lXX0 + ((log(nE + 1.0) + log(nE + 2.0)) - (lgB1 + lgB2));
        st->logZX = logaddexp(lZX0 + tota, totb);
        st->logZ = gm + log(sZ);
        st->logZ2 = gm2 + log(sZ2);
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------
// Cholesky with the reference's fallback (utils.F90:621-649), one warp, column-major in smem/global.
// ------------------------------------------------------------------------------------------
__device__ inline int warp_cholesky(const double* a, double* L, int D) {
    const int lane = threadIdx.x & 31;
    for (int e = lane; e < D * D; e += 32) L[e] = 0.0;
    __syncwarp();
    int fallback = 0;
    // Column i in one parallel step: lane l takes row j = i + l (+32, ...) and forms sum_k<i L[i,k] L[j,k] in the order
    // of the reference's loops; lane 0's row is the diagonal itself, whose square root is handed round by a shuffle.
    for (int i = 0; i < D && !fallback; ++i) {
        double dii = 0.0;
        for (int j0 = i; j0 < D; j0 += 32) {
            const int j = j0 + lane;
            double t = 0.0;
            if (j < D)
                for (int k = 0; k < i; ++k) t += L[i + k * D] * L[j + k * D];
            if (j0 == i) {
                const double d = __shfl_sync(FULL, a[i + i * D] - t, 0);
                if (d <= 0.0) { fallback = 1; break; }
                dii = sqrt(d);
                if (lane == 0) L[i + i * D] = dii;
            }
            if (j < D && j > i) L[j + i * D] = (a[i + j * D] - t) / dii;
        }
        __syncwarp();
    }
    if (fallback) {
        double tr = 0.0;
        for (int k = 0; k < D; ++k) tr += a[k + k * D];
        __syncwarp();
        for (int e = lane; e < D * D; e += 32) L[e] = 0.0;
        __syncwarp();
        for (int k = lane; k < D; k += 32) L[k + k * D] = sqrt(tr);
        __syncwarp();
    }
    return fallback;
}

// The same factorisation by a whole CTA, right-looking: with r = 1/sqrt(a_ii) column i of L is a(:, i) * r and the trailing
// lower triangle takes a(j, k) -= (a(j, i) r)(a(k, i) r), both read from the matrix as the previous column left it, so a
// column costs ONE barrier (and no dependent dot product per entry).  Every thread keeps its entries' (row, column) pairs.
// a: symmetric, column-major, DESTROYED (the caller keeps its own copy of the covariance); L: zeros above the diagonal.
// Same fallback as warp_cholesky (utils.F90:621-649: a non-positive pivot -> sqrt(trace) * identity).
__device__ inline int block_cholesky(double* a, double* L, int D) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    double tr = 0.0;
    for (int k = 0; k < D; ++k) tr += a[k + k * D];   // (every thread: the fallback's trace, before a is touched)
    __syncthreads();
    int fallback = 0;
    for (int i = 0; i < D; ++i) {
        const double d = a[i + i * D];   // the same value on every thread
        if (d <= 0.0) { fallback = 1; break; }
        const double r = rsqrt(d);
        for (int e = tid; e < D * D; e += nthr) {
            const int k = e / D, j = e - k * D;   // entry (j, k), column-major
            if (k == i) L[e] = j > i ? a[j + i * D] * r : (j == i ? d * r : 0.0);
            else if (k > i && j >= k) a[e] = fma(-(a[j + i * D] * r), a[k + i * D] * r, a[e]);
        }
        __syncthreads();
    }
    if (fallback) {
        __syncthreads();
        for (int e = tid; e < D * D; e += nthr) L[e] = 0.0;
        __syncthreads();
        for (int k = tid; k < D; k += nthr) L[k + k * D] = sqrt(tr);
        __syncthreads();
    }
    return fallback;
}


// ---------------------------------------------------------------- bulk copies (TMA, 1-D) and their mbarriers
// Phase U streams whole records global -> shared -> global.  A record is contiguous, so it moves as ONE bulk copy
// (cp.async.bulk, the non-tensor form of the TMA engine; SASS UBLKCP): no registers hold data in flight, a warp keeps
// two batches of records on the way, and completion is counted in bytes on an mbarrier.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_test(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    while (!mbar_test(bar, parity)) {}
}
// global -> shared, completion as bytes on `bar`; 16-byte aligned addresses, size a multiple of 16
__device__ __forceinline__ void bulk_g2s(void* sdst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// shared -> global (bulk group of the issuing thread)
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* ssrc, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

}  // namespace pc
