// pc_run_kernel.cuh -- the persistent run kernel and the kernel-level probes.
#pragma once
#include "pc_kernels.cuh"

namespace pc {

template <class V>
__device__ __forceinline__ V vload(const V* p) { return *(const volatile V*)p; }

struct WarpScratch {
    double* nh;    // R x LD (smem or global)
    double* dots;  // Dpad
    double* dvec;  // Dpad
    int* deck;     // R
    int* jd;       // R
};

__device__ __forceinline__ WarpScratch warp_scratch(const KParams& p, unsigned char* s_warp, double* nh_global) {
    WarpScratch ws;
    const int Dpad = (p.D + 1) & ~1;
    double* d = (double*)s_warp;
    ws.dots = d; d += Dpad;
    ws.dvec = d; d += Dpad;
    ws.deck = (int*)d; ws.jd = ws.deck + p.R;
    d += (p.R + 1) & ~1;  // 2*R ints = R doubles, rounded to even
    ws.nh = p.nh_in_smem ? d : nh_global;
    return ws;
}

// ---------------------------------------------------------------- initial live points (K1)
// GenerateLivePoints, generate.F90:153-183: attempt a draws cube = U(TAG_INIT, a, dim), accepted
// (in attempt order) when logL > logzero.
template <int NPL>
__device__ inline void init_phase(const KParams& p, const RunBuf& rb, DevRun* st, const Model<NPL>& M, int cta, int G,
                                  double* sc) {
    const int tid = threadIdx.x, lane = tid & 31, W = blockDim.x >> 5, gw = cta * W + (tid >> 5), GW = G * W;
    const int D = p.D, T = p.T, n = p.n;
    if (cta == 0) {
        for (int e = tid; e < D * D; e += blockDim.x) {
            double v = (e % D == e / D) ? 1.0 : 0.0;  // run_time_info.f90:193-194
            rb.chol[e] = v;
            rb.cov[e] = v;
        }
        if (tid == 0) {
            st->logZ = st->logZ2 = st->logZX = p.logzero;  // run_time_info.f90:165-175
            st->logX = st->logXX = 0.0;
            st->logX_last_update = 0.0;
            st->init_need = n;
            st->init_attempts = 0;
        }
    }
    group_sync(&st->bar, G);
    double* staging = rb.ph[1];
    for (;;) {
        const int need = vload(&st->init_need);
        if (need == 0) break;
        const long long a0 = vload(&st->init_attempts);
        for (int j = gw; j < need; j += GW) {
            double x[NPL], th[NPL];
#pragma unroll
            for (int q = 0; q < NPL; ++q) {
                int r = lane + 32 * q;
                x[q] = (r < D) ? uniform(rb.seed, TAG_INIT, (unsigned long long)(a0 + j), (unsigned)r, 0u) : 0.0;
            }
            double l = M.eval(x, th);
            M.write_record(staging + (size_t)j * T, x, th, p.logzero, l);
        }
        group_sync(&st->bar, G);
        if (cta == 0) {
            const int have = n - need;
            int run = 0;
            for (int base = 0; base < need; base += blockDim.x) {
                int j = base + tid;
                bool ok = j < need && __ldcg(staging + (size_t)j * T + T - 1) > p.logzero;
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
        group_sync(&st->bar, G);
    }
    if (cta == 0 && tid == 0) st->initialised = 1;
}

// ---------------------------------------------------------------- phase S (CTA 0)
__device__ inline void phase_S(const KParams& p, const RunBuf& rb, DevRun* st, double* sc, double* skey, int* sval,
                               int np2) {
    const int tid = threadIdx.x, n = p.n, T = p.T;
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
        if (st->nphantom + (long long)K * (p.R - 1) > rb.cap_ph) { if (tid == 0) st->status = ST_NEED_PHANTOM; return; }
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
        st->nphantom += (long long)K * (p.R - 1);
        st->nchains_base = st->nchains;
        st->nchains += K;
        st->ngen += 1;
        st->nslices += (long long)K * p.R;
        st->do_update = (st->logX <= st->logX_last_update + p.log_comp) ? 1 : 0;  // nested_sampling.F90:321
        st->status = ST_RUNNING;
    }
}

// ---------------------------------------------------------------- phase U, pass 1: survivor counts + sum x
template <int NPL>
__device__ inline void phase_U1(const KParams& p, const RunBuf& rb, DevRun* st, int cta, int G, unsigned char* smem_warp0,
                                int warp_bytes) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = blockDim.x >> 5;
    const int D = p.D, T = p.T, n = p.n;
    const long long total = vload(&st->nphantom);
    const double Lstar = vload(&st->Lstar);
    const double* src = rb.ph[vload(&st->cur_pool)];
    const long long chunk = (total + G - 1) / G, c0 = min(total, cta * chunk), c1 = min(total, c0 + chunk);
    const int lchunk = (n + G - 1) / G, l0 = min(n, cta * lchunk), l1 = min(n, l0 + lchunk);
    double sx[NPL];
#pragma unroll
    for (int j = 0; j < NPL; ++j) sx[j] = 0.0;
    double cnt = 0.0;
    for (long long rec = c0 + warp; rec < c1; rec += W) {
        const double* r = src + (size_t)rec * T;
        double l = __ldcg(r + T - 1);
        if (!(Lstar > l)) {  // clean_phantoms keeps it (run_time_info.f90:842-875)
            cnt += 1.0;
#pragma unroll
            for (int j = 0; j < NPL; ++j)
                if (lane + 32 * j < D) sx[j] += __ldcg(r + lane + 32 * j);
        }
    }
    for (int rec = l0 + warp; rec < l1; rec += W) {
        const double* r = rb.live + (size_t)rec * T;
#pragma unroll
        for (int j = 0; j < NPL; ++j)
            if (lane + 32 * j < D) sx[j] += __ldcg(r + lane + 32 * j);
    }
    __syncthreads();
    double* mine = (double*)(smem_warp0 + (size_t)warp * warp_bytes);
    if (lane == 0) mine[0] = cnt;
#pragma unroll
    for (int j = 0; j < NPL; ++j)
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
template <int NPL>
__device__ inline void phase_U2(const KParams& p, const RunBuf& rb, DevRun* st, int cta, int G, unsigned char* smem_warp0,
                                int warp_bytes, int* s_cnt) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = blockDim.x >> 5;
    const int D = p.D, T = p.T, n = p.n, ntri = p.ntri;
    const long long total = vload(&st->nphantom);
    const double Lstar = vload(&st->Lstar);
    const int pool = vload(&st->cur_pool);
    const double* src = rb.ph[pool];
    double* dst = rb.ph[pool ^ 1];
    const long long chunk = (total + G - 1) / G, c0 = min(total, cta * chunk), c1 = min(total, c0 + chunk);
    const int lchunk = (n + G - 1) / G, l0 = min(n, cta * lchunk), l1 = min(n, l0 + lchunk);
    long long base = 0, tot = 0;
    for (int g = 0; g < G; ++g) {
        long long c = vload(&rb.pcount[g]);
        if (g < cta) base += c;
        tot += c;
    }
    const double N = (double)(n + tot);
    double mean[NPL];
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
        double s = 0.0;
        if (lane + 32 * j < D)
            for (int g = 0; g < G; ++g) s += vload(&rb.partial[(size_t)g * p.partial_stride + 1 + lane + 32 * j]);
        mean[j] = s / N;
    }
    double* mine = (double*)(smem_warp0 + (size_t)warp * warp_bytes);  // [0..D) dv, then COV_ACC*32 partials
    double* s_dv = mine;
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
#pragma unroll
            for (int j = 0; j < NPL; ++j)
                if (lane + 32 * j < D) s_dv[lane + 32 * j] = __ldcg(r + lane + 32 * j) - mean[j];
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
        double* pacc = mine + ((D + 1) & ~1);
#pragma unroll
        for (int a = 0; a < COV_ACC; ++a) pacc[a * 32 + lane] = acc[a];
        __syncthreads();
        double* out = rb.partial + (size_t)cta * p.partial_stride + 1 + D + (size_t)pass * COV_ACC * 32;
        for (int e = tid; e < COV_ACC * 32; e += blockDim.x) {
            if (pass * COV_ACC * 32 + e < ntri) {
                double s = 0.0;
                for (int w = 0; w < W; ++w)
                    s += ((double*)(smem_warp0 + (size_t)w * warp_bytes))[((D + 1) & ~1) + e];
                out[e] = s;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------- update finalisation (CTA 0)
__device__ inline void finish_update(const KParams& p, const RunBuf& rb, DevRun* st, int G) {
    const int tid = threadIdx.x, D = p.D, ntri = p.ntri;
    long long tot = 0;
    for (int g = 0; g < G; ++g) tot += vload(&rb.pcount[g]);
    const double N = (double)(p.n + tot);
    for (int idx = tid; idx < ntri; idx += blockDim.x) {
        int ai = (int)((sqrt(8.0 * idx + 1.0) - 1.0) * 0.5);
        while (ai * (ai + 1) / 2 > idx) --ai;
        while ((ai + 1) * (ai + 2) / 2 <= idx) ++ai;
        int bi = idx - ai * (ai + 1) / 2;
        double s = 0.0;
        for (int g = 0; g < G; ++g) s += vload(&rb.partial[(size_t)g * p.partial_stride + 1 + D + idx]);
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
template <int NPL>
__global__ void __launch_bounds__(256, 1) pc_run_kernel(const __grid_constant__ KParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int G = p.ctas_per_run;
    const int run = blockIdx.x / G, cta = blockIdx.x % G;
    const RunBuf rb = p.runs[run];
    DevRun* st = rb.st;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = blockDim.x >> 5;
    const int gw = cta * W + warp, GW = G * W;
    const int D = p.D, T = p.T, n = p.n, R = p.R;

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

    const int nlp = (p.like_kind == LIKE_GAUSSIAN) ? 2 * D : (p.like_kind == LIKE_CORR ? D + D * D : 0);
    for (int e = tid; e < nlp; e += blockDim.x) s_like[e] = p.like_params[e];
    __syncthreads();

    WarpScratch ws = warp_scratch(p, s_warp, rb.nh ? rb.nh + (size_t)gw * R * p.LD : nullptr);
    Model<NPL> M;
    M.init(p, s_like, ws.dvec);

    if (!vload(&st->initialised)) init_phase<NPL>(p, rb, st, M, cta, G, sc);

    for (;;) {
        if (cta == 0) {
            bool dump_exit = false;
            if (st->update_pending) {
                finish_update(p, rb, st, G);
                dump_exit = p.want_dump != 0;
            }
            if (dump_exit) {
                if (tid == 0) st->status = ST_DUMP;
            } else if (vload(&st->status) != ST_ERROR) {
                phase_S(p, rb, st, sc, skey, sval, np2);
            }
        }
        group_sync(&st->bar, G);
        if (vload(&st->status) != ST_RUNNING) return;

        // ---------------- phase C: one chain per warp ----------------
        const int K = vload(&st->K);
        const double Lstar = vload(&st->Lstar);
        const long long ndead_base = vload(&st->ndead_base), nph_base = vload(&st->nph_base);
        const long long nchains_base = vload(&st->nchains_base);
        double* pool = rb.ph[vload(&st->cur_pool)];
        for (int e = tid; e < D * D; e += blockDim.x) s_chol[e] = __ldcg(rb.chol + e);
        __syncthreads();
        unsigned long long nlike = 0, nfail = 0;
        const int m = n - K;
        for (int k = gw; k < K; k += GW) {
            const unsigned long long uid = (unsigned long long)(nchains_base + k);
            double u = uniform(rb.seed, TAG_SEED, uid, 0u, 0u);  // GenerateSeed, generate.F90:19-55
            int choice = (int)ceil(u * (double)m);
            choice = max(1, min(m, choice));
            const int src = __ldcg(rb.order + K + choice - 1);
            const int dslot = __ldcg(rb.order + k);
            double x[NPL];
#pragma unroll
            for (int j = 0; j < NPL; ++j) x[j] = (lane + 32 * j < D) ? __ldcg(rb.live + (size_t)src * T + lane + 32 * j) : 0.0;
            // the dying point moves to the dead list before its slot is reused (run_time_info.f90:789-817)
            for (int e = lane; e < T; e += 32)
                rb.dead[(size_t)(ndead_base + k) * T + e] = __ldcg(rb.live + (size_t)dslot * T + e);
            __syncwarp();
            double lfin = run_chain<NPL>(p, M, rb.seed, uid, x, Lstar, s_chol, ws.nh, ws.deck, ws.jd, ws.dots,
                                         pool + (size_t)(nph_base + (long long)k * (R - 1)) * T,
                                         rb.live + (size_t)dslot * T, nlike);
            if (!(lfin > Lstar)) ++nfail;
        }
        if (lane == 0) {
            if (nlike) atomicAdd((unsigned long long*)&st->nlike, nlike);
            if (nfail) atomicAdd((unsigned long long*)&st->nfail, nfail);
        }
        if (vload(&st->do_update)) {
            group_sync(&st->bar, G);
            phase_U1<NPL>(p, rb, st, cta, G, s_warp0, p.warp_bytes);
            group_sync(&st->bar, G);
            phase_U2<NPL>(p, rb, st, cta, G, s_warp0, p.warp_bytes, s_cnt);
            if (cta == 0 && tid == 0) st->update_pending = 1;
        }
        group_sync(&st->bar, G);
    }
}

// ---------------------------------------------------------------- probes
// SliceSampling for explicit (seed point, contour, uid) triples; one warp per chain.
template <int NPL>
__global__ void __launch_bounds__(256, 1) pc_slice_chains_kernel(const __grid_constant__ KParams p, int nchains,
                                                                 const double* seed_points, const double* chol,
                                                                 const double* logL, const unsigned long long* uid,
                                                                 unsigned seed, double* babies, long long* nlike_out,
                                                                 double* nh_global) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = blockDim.x >> 5;
    const int D = p.D, T = p.T, R = p.R;
    double* s_chol = (double*)smem;
    double* s_like = (double*)(smem + p.off_like);
    unsigned char* s_warp = smem + p.off_warp + (size_t)warp * p.warp_bytes;
    const int nlp = (p.like_kind == LIKE_GAUSSIAN) ? 2 * D : (p.like_kind == LIKE_CORR ? D + D * D : 0);
    for (int e = tid; e < nlp; e += blockDim.x) s_like[e] = p.like_params[e];
    for (int e = tid; e < D * D; e += blockDim.x) s_chol[e] = chol[e];
    __syncthreads();
    const int gw = blockIdx.x * W + warp;
    WarpScratch ws = warp_scratch(p, s_warp, nh_global ? nh_global + (size_t)gw * R * p.LD : nullptr);
    Model<NPL> M;
    M.init(p, s_like, ws.dvec);
    for (int c = gw; c < nchains; c += gridDim.x * W) {
        double x[NPL];
#pragma unroll
        for (int j = 0; j < NPL; ++j) x[j] = (lane + 32 * j < D) ? seed_points[(size_t)c * T + lane + 32 * j] : 0.0;
        unsigned long long nl = 0;
        double* out = babies + (size_t)c * R * T;
        run_chain<NPL>(p, M, seed, uid[c], x, logL[c], s_chol, ws.nh, ws.deck, ws.jd, ws.dots, out,
                       out + (size_t)(R - 1) * T, nl);
        if (lane == 0) nlike_out[c] = (long long)nl;
    }
}

template <int NPL>
__global__ void pc_calculate_points_kernel(const __grid_constant__ KParams p, double* records, int npts, int* nlike) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = blockDim.x >> 5;
    const int D = p.D, T = p.T;
    double* s_like = (double*)(smem + p.off_like);
    unsigned char* s_warp = smem + p.off_warp + (size_t)warp * p.warp_bytes;
    const int nlp = (p.like_kind == LIKE_GAUSSIAN) ? 2 * D : (p.like_kind == LIKE_CORR ? D + D * D : 0);
    for (int e = tid; e < nlp; e += blockDim.x) s_like[e] = p.like_params[e];
    __syncthreads();
    WarpScratch ws = warp_scratch(p, s_warp, nullptr);
    Model<NPL> M;
    M.init(p, s_like, ws.dvec);
    int cnt = 0;
    for (int c = blockIdx.x * W + warp; c < npts; c += gridDim.x * W) {
        double* rec = records + (size_t)c * T;
        double x[NPL], th[NPL];
#pragma unroll
        for (int j = 0; j < NPL; ++j) x[j] = (lane + 32 * j < D) ? rec[lane + 32 * j] : 0.0;
        double l = M.eval(x, th);
        double birth = rec[T - 2];
        __syncwarp();
        M.write_record(rec, x, th, birth, l);
        if (l > p.logzero) ++cnt;
    }
    if (lane == 0 && cnt) atomicAdd(nlike, cnt);
}

// directions of one chain, de-shuffled into use order: out[i*D + r]
template <int NPL>
__global__ void pc_directions_kernel(int D, int R, int LD, unsigned seed, unsigned long long uid, double* nh_global,
                                     double* out) {
    extern __shared__ __align__(16) unsigned char smem[];
    double* dots = (double*)smem;
    int* deck = (int*)(dots + ((D + 1) & ~1));
    int* jd = deck + R;
    gen_directions<NPL>(D, R, LD, seed, uid, nh_global, deck, jd, dots);
    for (int i = 0; i < R; ++i)
        for (int r = threadIdx.x; r < D; r += 32) out[(size_t)i * D + r] = nh_global[(size_t)deck[i] * LD + r];
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
__global__ void pc_cholesky_kernel(const double* a, double* L, int D, int* fb) {
    int f = warp_cholesky(a, L, D);
    if (threadIdx.x == 0) *fb = f;
}

}  // namespace pc
