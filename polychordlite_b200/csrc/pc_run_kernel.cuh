// pc_run_kernel.cuh -- the persistent run kernel and the kernel-level probes.
#pragma once
#include "pc_kernels.cuh"

namespace pc {

template <class V>
__device__ __forceinline__ V vload(const V* p) { return *(const volatile V*)p; }

// ---------------------------------------------------------------- initial live points (K1)
// GenerateLivePoints, generate.F90:153-183: attempt a draws cube = U(TAG_INIT, a, dim), accepted
// (in attempt order) when logL > logzero.  One attempt per point group.
template <int G, int DPL>
__device__ inline void init_phase(const KParams& p, const RunBuf& rb, DevRun* st, const Model<G, DPL>& M, int cta,
                                  int NG, double* sc) {
    constexpr int NPT = 32 / G;
    const int tid = threadIdx.x, W = blockDim.x >> 5, gw = cta * W + (tid >> 5), GW = NG * W;
    const int D = p.cp.D, T = p.cp.T, n = p.n;
    if (cta == 0) {
        for (int e = tid; e < D * D; e += blockDim.x) {
            double v = (e % D == e / D) ? 1.0 : 0.0;  // run_time_info.f90:193-194
            rb.chol[e] = v;
            rb.cov[e] = v;
        }
        if (tid == 0) {
            st->logZ = st->logZ2 = st->logZX = p.cp.logzero;  // run_time_info.f90:165-175
            st->logX = st->logXX = 0.0;
            st->logX_last_update = 0.0;
            st->init_need = n;
            st->init_attempts = 0;
        }
    }
    group_sync(&st->bar, NG);
    double* staging = rb.ph[1];
    for (;;) {
        const int need = vload(&st->init_need);
        if (need == 0) break;
        const long long a0 = vload(&st->init_attempts);
        for (int j0 = gw * NPT; j0 < need; j0 += GW * NPT) {
            const int j = j0 + M.grp;
            double x[DPL], th[DPL];
#pragma unroll
            for (int k = 0; k < DPL; ++k)
                x[k] = M.valid(k) ? uniform(rb.seed, TAG_INIT, (unsigned long long)(a0 + j), (unsigned)M.dim(k), 0u) : 0.0;
            bool inc;
            const double l = M.eval(x, th, inc);
            double* rec = staging + (size_t)min(j, need - 1) * T;
            M.write_record(rec, (j < need) ? M.grp : -1, x, th, p.cp.logzero, l, inc);
            __syncwarp();
            if (j < need && M.sub == 0) M.finish_derived(rec, true);
        }
        group_sync(&st->bar, NG);
        if (cta == 0) {
            const int have = n - need;
            int run = 0;
            for (int base = 0; base < need; base += blockDim.x) {
                int j = base + tid;
                bool ok = j < need && __ldcg(staging + (size_t)j * T + T - 1) > p.cp.logzero;
                double tot;
                int pos = (int)block_exscan_sum(ok ? 1.0 : 0.0, &tot, sc);
                if (ok) {
                    const double* s = staging + (size_t)j * T;
                    double* d = rb.live + (size_t)(have + run + pos) * T;
                    for (int e = 0; e < T; ++e) d[e] = __ldcg(s + e);
                }
                run += (int)tot;
            }
            if (tid == 0) {
                st->nlike += run;
                st->init_need = need - run;
                st->init_attempts = a0 + need;
                if (a0 > 1000LL * n + 1000000LL) { st->init_need = 0; st->status = ST_ERROR; }
            }
        }
        group_sync(&st->bar, NG);
    }
    if (cta == 0 && tid == 0) st->initialised = 1;
}

// ---------------------------------------------------------------- phase S (CTA 0)
__device__ inline void phase_S(const KParams& p, const RunBuf& rb, DevRun* st, double* sc, double* skey, int* sval,
                               int np2) {
    const int tid = threadIdx.x, n = p.n, T = p.cp.T;
    for (int i = tid; i < np2; i += blockDim.x) {
        skey[i] = (i < n) ? __ldcg(rb.live + (size_t)i * T + T - 1) : INFINITY;
        sval[i] = i;
    }
    __syncthreads();
    const long long ndead = st->ndead;
    bool more = true;
    if (p.max_ndead == 0) more = false;
    else if (p.max_ndead > 0 && ndead >= p.max_ndead) more = false;
    else if (p.use_prec) {  // live_logZ (run_time_info.f90:683-709) vs precision_criterion (nested_sampling.F90:538)
        double m = -INFINITY;
        for (int i = tid; i < n; i += blockDim.x) m = fmax(m, skey[i]);
        m = block_max(m, sc);
        double s = 0.0;
        for (int i = tid; i < n; i += blockDim.x) s += exp(skey[i] - m);
        s = block_sum(s, sc);
        double lz = m + log(s) - log((double)n) + st->logX;
        if (lz < p.log_prec + st->logZ) more = false;
    }
    int K = min(p.batch_K, n - 1);
    if (p.max_ndead > 0) K = (int)min((long long)K, (long long)p.max_ndead - ndead);
    if (K < 1) more = false;
    if (more) {
        if (ndead + K + n > rb.cap_dead) { if (tid == 0) st->status = ST_NEED_DEAD; return; }
        if (st->nphantom + (long long)K * (p.cp.R - 1) > rb.cap_ph) { if (tid == 0) st->status = ST_NEED_PHANTOM; return; }
    } else if (ndead + n > rb.cap_dead) {
        if (tid == 0) st->status = ST_NEED_DEAD;
        return;
    }
    block_sort(skey, sval, np2);
    for (int i = tid; i < n; i += blockDim.x) rb.order[i] = sval[i];
    if (!more) {  // final kill-off, nested_sampling.F90:381-384
        evidence_deaths(st, skey, n, n, rb.logw + ndead, sc);
        for (size_t e = tid; e < (size_t)n * T; e += blockDim.x) {
            size_t i = e / T, c = e % T;
            rb.dead[(size_t)(ndead + i) * T + c] = __ldcg(rb.live + (size_t)sval[i] * T + c);
        }
        if (tid == 0) { st->ndead = ndead + n; st->K = 0; st->status = ST_DONE; }
        return;
    }
    evidence_deaths(st, skey, K, n, rb.logw + ndead, sc);
    if (tid == 0) {
        st->K = K;
        st->Lstar = skey[K - 1];
        st->ndead_base = ndead;
        st->ndead = ndead + K;
        st->nph_base = st->nphantom;
        st->nphantom += (long long)K * (p.cp.R - 1);
        st->nchains_base = st->nchains;
        st->nchains += K;
        st->ngen += 1;
        st->nslices += (long long)K * p.cp.R;
        st->do_update = (st->logX <= st->logX_last_update + p.log_comp) ? 1 : 0;  // nested_sampling.F90:321
        st->status = ST_RUNNING;
    }
}

// ---------------------------------------------------------------- phase U, pass 1: survivor counts + sum x
// clean_phantoms (run_time_info.f90:820-877) keeps a phantom unless the contour has passed it.  Each warp
// takes 32 consecutive records per step: the lanes test the 32 logL values, then the kept records are added
// four at a time (loads issued together), lane r holding the running sum of dimension r, r+32, ...
__device__ inline void phase_U1(const KParams& p, const RunBuf& rb, DevRun* st, int cta, int NG, unsigned char* smem_warp0,
                                int warp_bytes) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = blockDim.x >> 5;
    const int D = p.cp.D, T = p.cp.T, n = p.n;
    const long long total = vload(&st->nphantom);
    const double Lstar = vload(&st->Lstar);
    const double* src = rb.ph[vload(&st->cur_pool)];
    const long long chunk = (total + NG - 1) / NG, c0 = min(total, cta * chunk), c1 = min(total, c0 + chunk);
    const int lchunk = (n + NG - 1) / NG, l0 = min(n, cta * lchunk), l1 = min(n, l0 + lchunk);
    double* mine = (double*)(smem_warp0 + (size_t)warp * warp_bytes);  // [0]=count, [1..D]=sum x
    __syncthreads();
    double sx[4] = {0.0, 0.0, 0.0, 0.0};  // D <= 128
    double cnt = 0.0;
    auto add4 = [&](const double* r0, const double* r1, const double* r2, const double* r3) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int e = lane + 32 * j;
            if (e < D) {
                const double v0 = r0 ? __ldcg(r0 + e) : 0.0, v1 = r1 ? __ldcg(r1 + e) : 0.0;
                const double v2 = r2 ? __ldcg(r2 + e) : 0.0, v3 = r3 ? __ldcg(r3 + e) : 0.0;
                if (r0) sx[j] += v0;
                if (r1) sx[j] += v1;
                if (r2) sx[j] += v2;
                if (r3) sx[j] += v3;
            }
        }
    };
    for (long long tile = c0 + (long long)warp * 32; tile < c1; tile += (long long)W * 32) {
        const long long rec = tile + lane;
        const bool keep = rec < c1 && !(Lstar > __ldcg(src + (size_t)rec * T + T - 1));
        unsigned rem = __ballot_sync(FULL, keep);
        cnt += (double)__popc(rem);
        while (rem) {
            const double* r[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (rem) { const int b = __ffs(rem) - 1; rem &= rem - 1; r[j] = src + (size_t)(tile + b) * T; }
                else r[j] = nullptr;
            }
            add4(r[0], r[1], r[2], r[3]);
        }
    }
    for (int rec = l0 + warp * 4; rec < l1; rec += W * 4) {
        const double* b = rb.live + (size_t)rec * T;
        add4(b, rec + 1 < l1 ? b + T : nullptr, rec + 2 < l1 ? b + 2 * T : nullptr, rec + 3 < l1 ? b + 3 * T : nullptr);
    }
    if (lane == 0) mine[0] = cnt;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (lane + 32 * j < D) mine[1 + lane + 32 * j] = sx[j];
    __syncthreads();
    double* out = rb.partial + (size_t)cta * p.partial_stride;
    for (int e = tid; e < D + 1; e += blockDim.x) {
        double s = 0.0;
        for (int w = 0; w < W; ++w) s += ((double*)(smem_warp0 + (size_t)w * warp_bytes))[e];
        out[e] = s;
        if (e == 0) rb.pcount[cta] = (long long)s;
    }
    __syncthreads();
}

// ---------------------------------------------------------------- phase U, pass 2: stable compaction + centred outer products
__device__ inline void phase_U2(const KParams& p, const RunBuf& rb, DevRun* st, int cta, int NG, unsigned char* smem_warp0,
                                int warp_bytes, int* s_cnt) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = blockDim.x >> 5;
    const int D = p.cp.D, T = p.cp.T, n = p.n, ntri = p.ntri;
    const int Dpad = (D + 1) & ~1;
    const long long total = vload(&st->nphantom);
    const double Lstar = vload(&st->Lstar);
    const int pool = vload(&st->cur_pool);
    const double* src = rb.ph[pool];
    double* dst = rb.ph[pool ^ 1];
    const long long chunk = (total + NG - 1) / NG, c0 = min(total, cta * chunk), c1 = min(total, c0 + chunk);
    const int lchunk = (n + NG - 1) / NG, l0 = min(n, cta * lchunk), l1 = min(n, l0 + lchunk);
    double* mine = (double*)(smem_warp0 + (size_t)warp * warp_bytes);  // [0..Dpad) mean (warp 0's copy is used), [Dpad..2Dpad) dv, then COV_ACC*32 partials
    double* s_mean = (double*)smem_warp0;
    double* s_dv = mine + Dpad;
    long long* s_base = (long long*)(s_cnt + 16);  // [0] survivors in CTAs before this one, [1] all survivors
    __syncthreads();
    if (warp == 0) {
        long long before = 0, all = 0;
        for (int g = lane; g < NG; g += 32) {
            const long long c = __ldcg(rb.pcount + g);
            all += c;
            if (g < cta) before += c;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            before += __shfl_xor_sync(FULL, before, o);
            all += __shfl_xor_sync(FULL, all, o);
        }
        if (lane == 0) { s_base[0] = before; s_base[1] = all; }
    }
    __syncthreads();
    const long long base = s_base[0], tot = s_base[1];
    const double N = (double)(n + tot);
    for (int e = tid; e < D; e += blockDim.x) {  // the mean, summed over the CTAs in CTA order
        double s = 0.0;
#pragma unroll 8
        for (int g = 0; g < NG; ++g) s += __ldcg(rb.partial + (size_t)g * p.partial_stride + 1 + e);
        s_mean[e] = s / N;
    }
    __syncthreads();
    for (int pass = 0; pass < p.cov_passes; ++pass) {
        double acc[COV_ACC];
        int ab[COV_ACC];
#pragma unroll
        for (int a = 0; a < COV_ACC; ++a) {
            acc[a] = 0.0;
            int idx = (pass * COV_ACC + a) * 32 + lane;
            int ai = 0, bi = 0;
            if (idx < ntri) {
                ai = (int)((sqrt(8.0 * idx + 1.0) - 1.0) * 0.5);
                while (ai * (ai + 1) / 2 > idx) --ai;
                while ((ai + 1) * (ai + 2) / 2 <= idx) ++ai;
                bi = idx - ai * (ai + 1) / 2;
            }
            ab[a] = (idx < ntri) ? ((ai << 16) | bi) : -1;
        }
        auto accumulate = [&](const double* r) {
            __syncwarp();
            for (int e = lane; e < D; e += 32) s_dv[e] = __ldcg(r + e) - s_mean[e];
            __syncwarp();
#pragma unroll
            for (int a = 0; a < COV_ACC; ++a)
                if (ab[a] >= 0) acc[a] += s_dv[ab[a] >> 16] * s_dv[ab[a] & 0xffff];
        };
        long long run = 0;
        for (long long tile = c0; tile < c1; tile += blockDim.x) {
            long long rec = tile + tid;
            bool keep = rec < c1 && !(Lstar > __ldcg(src + (size_t)rec * T + T - 1));
            unsigned bal = __ballot_sync(FULL, keep);
            __syncthreads();
            if (lane == 0) s_cnt[warp] = __popc(bal);
            __syncthreads();
            int woff = 0, ttot = 0;
            for (int w = 0; w < W; ++w) {
                int c = s_cnt[w];
                if (w < warp) woff += c;
                ttot += c;
            }
            unsigned rem = bal;
            int kk = 0;
            while (rem) {
                int b = __ffs(rem) - 1;
                rem &= rem - 1;
                const double* r = src + (size_t)(tile + warp * 32 + b) * T;
                if (pass == 0) {
                    double* d = dst + (size_t)(base + run + woff + kk) * T;
                    for (int e = lane; e < T; e += 32) d[e] = __ldcg(r + e);
                }
                accumulate(r);
                ++kk;
            }
            run += ttot;
        }
        for (int rec = l0 + warp; rec < l1; rec += W) accumulate(rb.live + (size_t)rec * T);
        // combine the warps of this CTA in warp order
        __syncthreads();
        double* pacc = mine + 2 * Dpad;
#pragma unroll
        for (int a = 0; a < COV_ACC; ++a) pacc[a * 32 + lane] = acc[a];
        __syncthreads();
        double* out = rb.partial + (size_t)cta * p.partial_stride + 1 + D + (size_t)pass * COV_ACC * 32;
        for (int e = tid; e < COV_ACC * 32; e += blockDim.x) {
            if (pass * COV_ACC * 32 + e < ntri) {
                double s = 0.0;
                for (int w = 0; w < W; ++w)
                    s += ((double*)(smem_warp0 + (size_t)w * warp_bytes))[2 * Dpad + e];
                out[e] = s;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------- update finalisation (CTA 0)
__device__ inline void finish_update(const KParams& p, const RunBuf& rb, DevRun* st, int NG) {
    const int tid = threadIdx.x, D = p.cp.D, ntri = p.ntri;
    long long tot = 0;
#pragma unroll 8
    for (int g = 0; g < NG; ++g) tot += __ldcg(rb.pcount + g);
    const double N = (double)(p.n + tot);
    for (int idx = tid; idx < ntri; idx += blockDim.x) {
        int ai = (int)((sqrt(8.0 * idx + 1.0) - 1.0) * 0.5);
        while (ai * (ai + 1) / 2 > idx) --ai;
        while ((ai + 1) * (ai + 2) / 2 <= idx) ++ai;
        int bi = idx - ai * (ai + 1) / 2;
        double s = 0.0;
#pragma unroll 8
        for (int g = 0; g < NG; ++g) s += __ldcg(rb.partial + (size_t)g * p.partial_stride + 1 + D + idx);
        s /= N;  // calculate_covmats divides by N, not N-1 (run_time_info.f90:601-641)
        rb.cov[ai + bi * D] = s;
        rb.cov[bi + ai * D] = s;
    }
    __syncthreads();
    if (tid < 32) {
        int fb = warp_cholesky(rb.cov, rb.chol, D);
        if (tid == 0) {
            st->chol_fallback += fb;
            st->cov_N = N;
            st->nphantom = tot;
            st->cur_pool ^= 1;
            st->nupdates += 1;
            st->logX_last_update = st->logX;
            st->update_pending = 0;
        }
    }
    __syncthreads();
}

// ---------------------------------------------------------------- the persistent run kernel
//
// One generation, seen from a warp (NG = CTAs of the run's group, GW = NG * warps):
//   [CTA 0]   wait until every warp arrived (wbar) -> finish a pending update -> phase S -> group barrier B
//   [others]  chains -> arrive (wbar) -> prepare the directions of the NEXT generation's chain -> barrier B
// so the counter-addressed direction/uniform preparation of a chain overlaps the bookkeeping of CTA 0
// and the wait for the slowest chain.  Generations at the update cadence insert phase U (all CTAs)
// between the arrival and the bookkeeping.
template <int G, int DPL>
__global__ void __launch_bounds__(256, 1) pc_run_kernel(const __grid_constant__ KParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int NPT = 32 / G;
    const int NG = p.ctas_per_run;
    const int run = blockIdx.x / NG, cta = blockIdx.x % NG;
    const RunBuf rb = p.runs[run];
    DevRun* st = rb.st;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = blockDim.x >> 5;
    const int gw = cta * W + warp, GW = NG * W;
    const int D = p.cp.D, T = p.cp.T, n = p.n, R = p.cp.R, LD = p.cp.LD;
    const int c0 = p.chain_cta0, Gc = NG - c0;

    double* s_chol = (double*)smem;
    double* s_like = (double*)(smem + p.off_like);
    unsigned char* s_warp0 = smem + p.off_warp;
    unsigned char* s_warp = s_warp0 + (size_t)warp * p.warp_bytes;
    // CTA-wide scratch of phase S overlays the per-warp area
    double* sc = (double*)s_warp0;           // 64 doubles
    int* s_cnt = (int*)(smem + p.off_warp - 64 * (int)sizeof(int));
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    double* skey = sc + 64;
    int* sval = (int*)(skey + np2);

    const int nlp = (p.cp.like_kind == LIKE_GAUSSIAN) ? 2 * D : (p.cp.like_kind == LIKE_CORR ? D + D * D : 0);
    for (int e = tid; e < nlp; e += blockDim.x) s_like[e] = p.like_params[e];
    __syncthreads();

    const ChainScratch cs = chain_scratch(s_warp, D, R, LD, p.nh_in_smem != 0, p.cp.like_kind, NPT,
                                          rb.nh ? rb.nh + (size_t)gw * R * LD : nullptr);
    Model<G, DPL> M;
    M.init(p.cp, s_like, p.prior_params, cs.dvec);

    if (!vload(&st->initialised)) init_phase<G, DPL>(p, rb, st, M, cta, NG, sc);

    // chain whose directions currently sit in this warp's scratch (~0 = none), and whether they are whitened
    unsigned long long prep_uid = ~0ull;
    bool prep_white = false;
    unsigned int wtarget = 0;
    bool have_wtarget = false;

    const bool timer = (tid == 0) && (cta == 0);          // bookkeeping phases
    const bool ctimer = (tid == 0) && (cta == c0);         // chain phases of one representative warp
    const long long t_start = clock64();
    for (;;) {
        if (cta == 0) {
            long long t0 = clock64();
            if (have_wtarget) warp_wait(&st->wbar, wtarget);  // every chain of the previous generation is written
            __syncthreads();
            long long t1 = clock64();
            bool dump_exit = false;
            if (st->update_pending) {
                finish_update(p, rb, st, NG);
                dump_exit = p.want_dump != 0;
            }
            long long t2 = clock64();
            if (dump_exit) {
                if (tid == 0) st->status = ST_DUMP;
            } else if (vload(&st->status) != ST_ERROR) {
                phase_S(p, rb, st, sc, skey, sval, np2);
            }
            __syncthreads();
            if (timer) {
                long long t3 = clock64();
                st->cyc_wait += t1 - t0; st->cyc_fin += t2 - t1; st->cyc_S += t3 - t2; 
            }
            prep_uid = ~0ull;  // phase S overlays this CTA's chain scratch
        }
        group_sync(&st->bar, NG);
        if (vload(&st->status) != ST_RUNNING) {
            if (timer) st->cyc_total += clock64() - t_start;
            return;
        }

        // ---------------- phase C: chains, one warp each ----------------
        const int K = vload(&st->K);
        const double Lstar = vload(&st->Lstar);
        const long long ndead_base = vload(&st->ndead_base), nph_base = vload(&st->nph_base);
        const long long nchains_base = vload(&st->nchains_base);
        const int do_update = vload(&st->do_update);
        double* pool = rb.ph[vload(&st->cur_pool)];
        for (int e = tid; e < D * D; e += blockDim.x) s_chol[e] = __ldcg(rb.chol + e);
        __syncthreads();
        unsigned long long nlike = 0, nfail = 0;
        const int m = n - K;
        if (cta >= c0) {
            for (int li = warp;; li += W) {
                const int k = (cta - c0) + Gc * li;
                if (k >= K) break;
                const unsigned long long uid = (unsigned long long)(nchains_base + k);
                double u = uniform(rb.seed, TAG_SEED, uid, 0u, 0u);  // GenerateSeed, generate.F90:19-55
                int choice = (int)ceil(u * (double)m);
                choice = max(1, min(m, choice));
                const int src = __ldcg(rb.order + K + choice - 1);
                const int dslot = __ldcg(rb.order + k);
                double x[DPL];
#pragma unroll
                for (int j = 0; j < DPL; ++j) x[j] = M.valid(j) ? __ldcg(rb.live + (size_t)src * T + M.dim(j)) : 0.0;
                // the dying point moves to the dead list before its slot is reused (run_time_info.f90:789-817)
                for (int e = lane; e < T; e += 32)
                    rb.dead[(size_t)(ndead_base + k) * T + e] = __ldcg(rb.live + (size_t)dslot * T + e);
                long long tc0 = clock64();
                if (prep_uid != uid) {
                    prep_chain(D, R, LD, rb.seed, uid, cs);
                    prep_white = false;
                }
                long long tc1 = clock64();
                if (!prep_white) whiten_chain(D, R, LD, s_chol, cs);
                prep_uid = ~0ull;
                long long tc2 = clock64();
                double lfin = slice_chain<G, DPL>(p.cp, M, rb.seed, uid, x, Lstar, cs,
                                                  pool + (size_t)(nph_base + (long long)k * (R - 1)) * T,
                                                  rb.live + (size_t)dslot * T, nlike);
                if (ctimer) {
                    long long tc3 = clock64();
                    st->cyc_prep += tc1 - tc0; st->cyc_white += tc2 - tc1; st->cyc_slice += tc3 - tc2;
                }
                if (!(lfin > Lstar)) ++nfail;
            }
        }
        if (lane == 0) {
            if (nlike) atomicAdd((unsigned long long*)&st->nlike, nlike);
            if (nfail) atomicAdd((unsigned long long*)&st->nfail, nfail);
        }
        wtarget = warp_arrive(&st->wbar, GW);
        have_wtarget = true;
        // the first chain this warp will run in the next generation (if the run goes on with the same K)
        const int knext = (cta - c0) + Gc * warp;
        const bool will_chain = cta >= c0 && cta != 0 && knext < p.batch_K;
        if (do_update) {
            long long tu0 = clock64();
            warp_wait(&st->wbar, wtarget);
            long long tu1 = clock64();
            phase_U1(p, rb, st, cta, NG, s_warp0, p.warp_bytes);
            group_sync(&st->bar, NG);
            phase_U2(p, rb, st, cta, NG, s_warp0, p.warp_bytes, s_cnt);
            if (cta == 0 && tid == 0) st->update_pending = 1;
            group_sync(&st->bar, NG);
            if (timer) { st->cyc_wait += tu1 - tu0; st->cyc_U += clock64() - tu1; }
            have_wtarget = false;
            if (will_chain) {  // the Cholesky factor is about to change: whiten after the barrier
                prep_uid = (unsigned long long)(nchains_base + K + knext);
                prep_chain(D, R, LD, rb.seed, prep_uid, cs);
                prep_white = false;
            }
        } else if (will_chain) {
            prep_uid = (unsigned long long)(nchains_base + K + knext);
            prep_chain(D, R, LD, rb.seed, prep_uid, cs);
            whiten_chain(D, R, LD, s_chol, cs);
            prep_white = true;
        }
    }
}

// ---------------------------------------------------------------- probes
// SliceSampling for explicit (seed point, contour, uid) triples; one warp per chain.
template <int G, int DPL>
__global__ void __launch_bounds__(256, 1) pc_slice_chains_kernel(const __grid_constant__ KParams p, int nchains,
                                                                 const double* seed_points, const double* chol,
                                                                 const double* logL, const unsigned long long* uid,
                                                                 unsigned seed, double* babies, long long* nlike_out,
                                                                 double* nh_global) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int NPT = 32 / G;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = blockDim.x >> 5;
    const int D = p.cp.D, T = p.cp.T, R = p.cp.R, LD = p.cp.LD;
    double* s_chol = (double*)smem;
    double* s_like = (double*)(smem + p.off_like);
    unsigned char* s_warp = smem + p.off_warp + (size_t)warp * p.warp_bytes;
    const int nlp = (p.cp.like_kind == LIKE_GAUSSIAN) ? 2 * D : (p.cp.like_kind == LIKE_CORR ? D + D * D : 0);
    for (int e = tid; e < nlp; e += blockDim.x) s_like[e] = p.like_params[e];
    for (int e = tid; e < D * D; e += blockDim.x) s_chol[e] = chol[e];
    __syncthreads();
    const int gw = blockIdx.x * W + warp;
    const ChainScratch cs = chain_scratch(s_warp, D, R, LD, p.nh_in_smem != 0, p.cp.like_kind, NPT,
                                          nh_global ? nh_global + (size_t)gw * R * LD : nullptr);
    Model<G, DPL> M;
    M.init(p.cp, s_like, p.prior_params, cs.dvec);
    for (int c = gw; c < nchains; c += gridDim.x * W) {
        double x[DPL];
#pragma unroll
        for (int j = 0; j < DPL; ++j) x[j] = M.valid(j) ? seed_points[(size_t)c * T + M.dim(j)] : 0.0;
        unsigned long long nl = 0;
        double* out = babies + (size_t)c * R * T;
        prep_chain(D, R, LD, seed, uid[c], cs);
        whiten_chain(D, R, LD, s_chol, cs);
        slice_chain<G, DPL>(p.cp, M, seed, uid[c], x, logL[c], cs, out, out + (size_t)(R - 1) * T, nl);
        if (lane == 0) nlike_out[c] = (long long)nl;
    }
}

template <int G, int DPL>
__global__ void pc_calculate_points_kernel(const __grid_constant__ KParams p, double* records, int npts, int* nlike) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int NPT = 32 / G;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = blockDim.x >> 5;
    const int D = p.cp.D, T = p.cp.T, R = p.cp.R, LD = p.cp.LD;
    double* s_like = (double*)(smem + p.off_like);
    unsigned char* s_warp = smem + p.off_warp + (size_t)warp * p.warp_bytes;
    const int nlp = (p.cp.like_kind == LIKE_GAUSSIAN) ? 2 * D : (p.cp.like_kind == LIKE_CORR ? D + D * D : 0);
    for (int e = tid; e < nlp; e += blockDim.x) s_like[e] = p.like_params[e];
    __syncthreads();
    const ChainScratch cs = chain_scratch(s_warp, D, R, LD, p.nh_in_smem != 0, p.cp.like_kind, NPT, nullptr);
    Model<G, DPL> M;
    M.init(p.cp, s_like, p.prior_params, cs.dvec);
    int cnt = 0;
    for (int c0 = (blockIdx.x * W + warp) * NPT; c0 < npts; c0 += gridDim.x * W * NPT) {
        const int c = c0 + M.grp;
        double* rec = records + (size_t)min(c, npts - 1) * T;
        double x[DPL], th[DPL];
#pragma unroll
        for (int j = 0; j < DPL; ++j) x[j] = M.valid(j) ? rec[M.dim(j)] : 0.0;
        bool inc;
        const double l = M.eval(x, th, inc);
        const double birth = rec[T - 2];
        __syncwarp();
        M.write_record(rec, (c < npts) ? M.grp : -1, x, th, birth, l, inc);
        __syncwarp();
        if (c < npts && M.sub == 0) {
            M.finish_derived(rec, true);
            if (l > p.cp.logzero) ++cnt;
        }
    }
    cnt = warp_sum_int(cnt);
    if (lane == 0 && cnt) atomicAdd(nlike, cnt);
}

}  // namespace pc
