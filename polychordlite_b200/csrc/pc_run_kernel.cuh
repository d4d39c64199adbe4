// pc_run_kernel.cuh -- the persistent run kernel and the kernel-level probes.
#pragma once
#include "pc_kernels.cuh"

namespace pc {

template <class V>
__device__ __forceinline__ V vload(const V* p) { return *(const volatile V*)p; }

// ---------------------------------------------------------------- initial live points (K1)
// GenerateLivePoints, generate.F90:153-183: attempt a draws cube = U(TAG_INIT, a, dim), accepted
// (in attempt order) when logL > logzero.  One attempt per point group.
template <int G, int DPL, int KIND>
__device__ inline void init_phase(const KParams& p, const RunBuf& rb, DevRun* st, const Model<G, DPL, KIND>& M, int cta,
                                  int NG, double* sc) {
    constexpr int NPT = 32 / G;
    const int tid = threadIdx.x, W = blockDim.x >> 5, gw = cta * W + (tid >> 5), GW = NG * W;
    const int D = p.cp.D, T = p.cp.T, n = p.n0;   // nprior points when nprior > nlive (generate.F90:142-153), or the caller's cube_samples
    if (cta == 0) {
        for (int e = tid; e < D * D; e += blockDim.x) {
            double v = (e % D == e / D) ? 1.0 : 0.0;  // run_time_info.f90:193-194
            rb.chol[e] = v;
            rb.cov[e] = v;
        }
        for (int e = tid; e < D; e += blockDim.x) rb.gsum[2 + D + e] = 0.5;  // first pivot of the covariance moments: the cube centre
        if (tid == 0) {
            st->logZ = st->logZ2 = st->logZX = p.cp.logzero;  // run_time_info.f90:165-175
            st->logX = st->logXX = 0.0;
            st->logX_last_update = 0.0;
            st->ncl = 1;
            st->n = n;
            st->init_need = p.live_given ? 0 : n;   // host-callback runs, cube_samples: the host evaluated and uploaded the live points
            if (!p.live_given) st->init_attempts = 0;
        }
    }
    group_sync(&st->bar, NG, p.backoff);
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
        group_sync(&st->bar, NG, p.backoff);
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
                if (p.sh.rank == 0) st->nlike += run;  // a sharded run generates the live points on every rank, counts them once
                st->init_need = need - run;
                st->init_attempts = a0 + need;
                if (a0 > 1000LL * n + 1000000LL) { st->init_need = 0; st->status = ST_ERROR; }
            }
        }
        group_sync(&st->bar, NG, p.backoff);
    }
    if (cta == 0 && tid == 0) st->initialised = 1;
}

// ---------------------------------------------------------------- phase S (CTA 0)
// S1 (on the critical path of a generation): termination test (live_logZ, run_time_info.f90:683-709, against
// precision_criterion, nested_sampling.F90:538), the order of the n live points by (logL, slot), the K lowest die:
// contour, number of births, bases, update decision.  The order is maintained incrementally: the n-K survivors of a
// regular generation (K deaths, K successful births into the vacated slots) are still sorted, so only the K new babies
// are sorted (bitonic) and the two lists are merged by rank (binary searches).  The first generation, and one that follows a
// generation that moved live points (settle_generation), sorts everything.
// S2 (off the critical path when CTA 0 only keeps the books): the evidence recurrences of the K deaths.
// log X after `count` deaths from n_start live points: prod (n_j / (n_j + 1)) telescopes; the same expression evidence_deaths
// leaves in DevRun::logX
__device__ inline double logX_after(double lX, int count, int n_start) {
    return lX + (log((double)(n_start - count) + 1.0) - log((double)n_start + 1.0));
}

// target number of live points above the contour (run_time_info.f90:766-771: the threshold with the largest loglike
// below it; none: settings%nlive)
__device__ inline int target_nlive(const KParams& p, double contour) {
    int nlive = p.n, best = -1;
    for (int q = 0; q < p.dyn_m; ++q)
        if (contour > p.dyn_loglikes[q] && (best < 0 || p.dyn_loglikes[q] > p.dyn_loglikes[best])) best = q;
    if (best >= 0) nlive = p.dyn_nlives[best];
    return nlive;
}

// The generation just finished left empty live slots: vacated slots without a birth (B < K), or births that FAILED
// (last baby not above the contour: a non-deterministic or plateau likelihood).  As replace_point (run_time_info.f90:
// 781-785) a failed baby does not become a live point: it goes to the dead list with log-weight logzero, in chain
// order, and counts towards nfail (nested_sampling.F90:315-319).  The empty slots are then closed: with n' = extent -
// holes live points left, the occupied slots >= n' fill the holes < n', the j-th lowest into the j-th lowest (the rule
// the oracle's batched_generation applies).  CTA 0; smem: the phase-S area.
__device__ inline void settle_generation(const KParams& p, const RunBuf& rb, DevRun* st, unsigned char* smem_raw, bool clear_flags) {
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, W = nthr >> 5, T = p.cp.T;
    const int K = st->K, B = st->B, n0 = st->n_gen;
    const int ext = n0 + max(0, B - K);
    const int* ord = rb.order + st->order_off;
    double* sc = (double*)smem_raw;                             // 64 doubles
    unsigned char* hole = smem_raw + 64 * 8;                    // ext flags
    int* fail = (int*)(smem_raw + 64 * 8 + ((ext + 15) & ~15)); // B flags
    int* list_h = fail + ((B + 3) & ~3);                        // holes below n', ascending
    int* list_m = list_h + ext;                                 // occupied slots >= n', ascending
    __shared__ int s_nf, s_run;
    for (int k = tid; k < B; k += nthr) fail[k] = __ldcg(rb.cfail + k);
    for (int e = tid; e < ext; e += nthr) hole[e] = 0;
    if (tid == 0) { s_nf = 0; s_run = st->fail_run; }
    __syncthreads();
    // (a) failed births -> dead list, in chain order
    const long long nd0 = st->ndead;
    for (int k = 0; k < B; ++k) {
        if (!fail[k]) { if (tid == 0) s_run = 0; continue; }   // uniform over the CTA
        const int slot = k < K ? __ldcg(ord + k) : n0 + (k - K);
        const int nf = s_nf;
        __syncthreads();
        for (int e = tid; e < T; e += nthr) rb.dead[(size_t)(nd0 + nf) * T + e] = __ldcg(rb.live + (size_t)slot * T + e);
        if (tid == 0) {
            rb.logw[nd0 + nf] = p.cp.logzero; s_nf = nf + 1; s_run += 1;
            if (rb.deadlab) { rb.deadlab[nd0 + nf] = __ldcg(rb.lab + slot); rb.deadn[nd0 + nf] = 0; }   // (a failed birth carries no weight)
        }
        __syncthreads();
    }
    // (b) the empty slots
    for (int k = tid; k < K; k += nthr)
        if (k >= B || fail[k]) hole[__ldcg(ord + k)] = 1;
    for (int k = K + tid; k < B; k += nthr)
        if (fail[k]) hole[n0 + (k - K)] = 1;
    __syncthreads();
    int cnt = 0;
    for (int e = tid; e < ext; e += nthr) cnt += hole[e];
    int nh;
    block_exscan_int(cnt, &nh, sc);
    const int n1 = ext - nh;
    // (c) ascending lists by contiguous chunks
    {
        const int chunk = (ext + nthr - 1) / nthr, e0 = tid * chunk, e1 = min(ext, e0 + chunk);
        int ch = 0, cm = 0;
        for (int e = e0; e < e1; ++e) { ch += (e < n1 && hole[e]); cm += (e >= n1 && !hole[e]); }
        int th, tm;
        int ph = block_exscan_int(ch, &th, sc);
        int pm = block_exscan_int(cm, &tm, sc);
        for (int e = e0; e < e1; ++e) {
            if (e < n1 && hole[e]) list_h[ph++] = e;
            if (e >= n1 && !hole[e]) list_m[pm++] = e;
        }
        __syncthreads();
        // (d) the moves: distinct sources and destinations, one record per warp
        for (int j = warp; j < th; j += W) {
            const int src = list_m[j], dst = list_h[j];
            for (int e = lane; e < T; e += 32) rb.live[(size_t)dst * T + e] = __ldcg(rb.live + (size_t)src * T + e);
            if (p.clustering && lane == 0) rb.lab[dst] = __ldcg(rb.lab + src);
        }
    }
    __syncthreads();
    if (tid == 0) {
        st->n = n1;
        st->ndead = nd0 + s_nf;
        st->fail_run = s_run;
        const int nfail = p.nfail > 0 ? p.nfail : p.n;
        if (s_run > nfail) st->stop_nfail = 1;
        st->order_valid = 0;
        if (clear_flags) { st->holes_due = 0; st->nfail_gen = 0u; }
    }
    __syncthreads();
}

// What phase S1 decides for the next generation, written by one thread: contour, deaths and births, bases of the dead /
// phantom / chain counters, the update decision, and the 64-byte line every warp reads at the release.
__device__ inline void publish_generation(const KParams& p, const RunBuf& rb, DevRun* st, int n, int K, bool trim,
                                          long long ndead, const int* newo, const double* newk, double lX_new) {
    const double Lstar = __ldcg(newk + K - 1);   // (phase D's CTAs may have written it)
    // births: the live count moves towards its target above the contour, at most 2 batch_K a generation
    // (constant target: B = K; the batched form of run_time_info.f90:766-777)
    int B = trim ? 0 : max(0, min(max(target_nlive(p, Lstar), 1) - (n - K), 2 * p.batch_K));
    if (p.sh.world > 1) B = K;   // a sharded run keeps the live count fixed
    const long long nph = st->nphantom, nch = st->nchains, ngen = st->ngen + (trim ? 0 : 1);
    const int order_off = (int)(newo - rb.order);
    const int do_update = (!trim && lX_new <= st->logX_last_update + p.log_comp) ? 1 : 0;  // nested_sampling.F90:321
    st->K = K;
    st->B = B;
    st->n_gen = n;
    st->trimmed = 1;
    st->holes_due = (B != K) ? 1 : 0;
    st->nfail_gen = 0u;
    st->Lstar = Lstar;
    st->order_off = order_off;
    st->order_valid = 1;
    st->ndead_base = ndead;
    st->ndead = ndead + K;
    st->nph_base = nph;
    const int Bloc = p.sh.world > 1 ? (B - p.sh.rank + p.sh.world - 1) / p.sh.world : B;  // chains of this rank
    st->nphantom = nph + (long long)Bloc * (p.cp.R - 1);
    st->nph_glob += (long long)B * (p.cp.R - 1);
    st->nchains_base = nch;
    st->nchains = nch + B;
    st->ngen = ngen;
    st->nslices += (long long)B * p.cp.R;
    st->do_update = do_update;
    st->status = ST_RUNNING;
    st->pub[0] = (unsigned long long)__double_as_longlong(Lstar);
    st->pub[1] = (unsigned long long)ndead;
    st->pub[2] = (unsigned long long)nph;
    st->pub[3] = (unsigned long long)nch;
    st->pub[4] = (unsigned long long)ngen;
    st->pub[5] = (unsigned long long)(unsigned)K | ((unsigned long long)(unsigned)do_update << 32);
    st->pub[6] = (unsigned long long)(unsigned)order_off | ((unsigned long long)(unsigned)st->cur_pool << 32);
    st->pub[7] = (unsigned long long)(unsigned)st->ncl | ((unsigned long long)(unsigned)st->nupdates << 32);
    st->pub[8] = (unsigned long long)(unsigned)n | ((unsigned long long)(unsigned)B << 32);
    st->pub[9] = (unsigned long long)st->nfail;
}

// Phase S1 after phase D in the common case -- the run goes on, nothing has to grow -- on ONE warp of CTA 0, without a
// block-wide barrier: the termination test from the CTAs' partial sums (they are taken about M0, the largest survivor key),
// the capacity tests, the publication.  Returns false when anything else is due (termination and the final kill-off, a
// pool that has to grow, the nprior trim, nfail): phase_S1 then decides from the same inputs.
__device__ inline bool phase_S1_fast(const KParams& p, const RunBuf& rb, DevRun* st, int nparts) {
    const int lane = threadIdx.x & 31;
    const int n = st->n;
    const long long ndead = st->ndead;
    if (st->stop_nfail || (!st->trimmed && n > p.n) || p.max_ndead == 0 || (p.max_ndead > 0 && ndead >= p.max_ndead)) return false;
    int K = min(p.batch_K, n - 1);
    if (p.max_ndead > 0) K = (int)min((long long)K, (long long)p.max_ndead - ndead);
    if (K < 1) return false;
    int* newo = rb.order + (st->order_off ? 0 : p.nmax);
    const double* oldk = rb.okey + st->order_off;
    double* newk = rb.okey + (st->order_off ? 0 : p.nmax);
    if (p.use_prec) {
        const double M0 = __ldcg(oldk + n - 1);   // phase D's reference point: the largest survivor key
        double s = 0.0;
        for (int c = lane; c < nparts; c += 32) s += __ldcg(rb.dpart + c);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
        const double lz = M0 + log(s) - log((double)n) + st->logX;   // live_logZ, run_time_info.f90:683-709
        if (!(lz >= p.log_prec + st->logZ)) return false;              // precision reached (or not a number): the full path ends the run
    }
    const int Bmax = 2 * p.batch_K;
    if (ndead + K + Bmax + p.nmax > rb.cap_dead) return false;
    const long long nph_test = p.sh.world > 1 ? st->nph_glob : st->nphantom;
    if (nph_test + (long long)Bmax * (p.cp.R - 1) > rb.cap_ph) return false;
    if (p.boost_thin > 0.0 && (long long)st->nboost + st->nphantom + (long long)Bmax * (p.cp.R - 1) > rb.cap_boost) return false;
    // log X after the K deaths: sum_j log((n - j) / (n - j + 1)) telescopes
    const double lX_new = logX_after(st->logX, K, n);
    if (lane == 0) publish_generation(p, rb, st, n, K, false, ndead, newo, newk, lX_new);
    return true;
}

// returns true when the evidence of the K deaths is still to be accumulated (S2).
// merged: phase D (all CTAs) already wrote the new order into the other half of rb.order / rb.okey and left the CTAs'
// partial sums of the termination test in rb.dpart[0 .. nparts).
__device__ inline bool phase_S1(const KParams& p, const RunBuf& rb, DevRun* st, const SmemS& sm, bool merged, int nparts) {
    const int tid = threadIdx.x, T = p.cp.T, nthr = blockDim.x;
    const int n = st->n;
    double* sc = sm.sc;
    long long q0 = clock64();
    const long long ndead = st->ndead;
    const int Kp = st->K;
    const bool merge = st->order_valid != 0 && Kp > 0 && Kp < n;
    const int* oldo = rb.order + st->order_off;
    int* newo = rb.order + (st->order_off ? 0 : p.nmax);
    const double* oldk = rb.okey + st->order_off;
    double* newk = rb.okey + (st->order_off ? 0 : p.nmax);
    const int m = n - Kp, npB = next_pow2(max(Kp, 1)), np2 = next_pow2(n);
    if (merged) {
        // nothing to load: the order is in place
    } else
    if (merge) {
        for (int i = tid; i < m; i += nthr) {  // the survivors: keys and slots as the previous phase S ordered them
            sm.akey[i] = __ldcg(oldk + Kp + i);
            sm.aval[i] = __ldcg(oldo + Kp + i);
        }
        for (int j = tid; j < npB; j += nthr) {
            const int slot = j < Kp ? __ldcg(oldo + j) : 0x7fffffff;
            sm.bval[j] = slot;
            sm.bkey[j] = j < Kp ? __ldcg(rb.live + (size_t)slot * T + T - 1) : INFINITY;
        }
    } else {
        for (int i = tid; i < np2; i += nthr) {
            sm.akey[i] = (i < n) ? __ldcg(rb.live + (size_t)i * T + T - 1) : INFINITY;
            ((int*)(sm.akey + np2))[i] = i;
        }
    }
    __syncthreads();
    long long q1 = clock64();
    // every live key once, whichever layout: index i < m in akey, the rest in bkey (merge) / all in akey
    auto key_at = [&](int i) -> double { return merge ? (i < m ? sm.akey[i] : sm.bkey[i - m]) : sm.akey[i]; };
    // nprior > nlive: the excess points die first, one after the other, without births (nested_sampling.F90:201-203)
    const bool trim = !st->trimmed && n > p.n;
    bool more = true;
    if (trim) more = true;
    else if (st->stop_nfail) more = false;
    else if (p.max_ndead == 0) more = false;
    else if (p.max_ndead > 0 && ndead >= p.max_ndead) more = false;
    else if (p.use_prec) {
        double mx = -INFINITY, s = 0.0;
        if (merged) {   // the CTAs' partial sums of exp(logL - M0), M0 the largest survivor key (phase D)
            mx = __ldcg(oldk + n - 1);
            for (int c = tid; c < nparts; c += nthr) s += __ldcg(rb.dpart + c);
            s = block_sum(s, sc);
        } else {
            for (int i = tid; i < n; i += nthr) mx = fmax(mx, key_at(i));
            mx = block_max(mx, sc);
            for (int i = tid; i < n; i += nthr) s += exp(key_at(i) - mx);
            s = block_sum(s, sc);
        }
        double lz = mx + log(s) - log((double)n) + st->logX;
        if (lz < p.log_prec + st->logZ) more = false;
    }
    int K = min(p.batch_K, n - 1);
    if (trim) K = n - p.n;
    else if (p.max_ndead > 0) K = (int)min((long long)K, (long long)p.max_ndead - ndead);
    if (K < 1) more = false;
    const int Bmax = trim ? 0 : 2 * p.batch_K;   // births of this generation cannot exceed this
    if (more) {
        // room for the K deaths, the births that may fail, and the final kill-off
        if (ndead + K + Bmax + p.nmax > rb.cap_dead) { if (tid == 0) st->status = ST_NEED_DEAD; return false; }
        // sharded run: the decision must be the same on every rank, so it is taken on the phantoms of all ranks
        const long long nph_test = p.sh.world > 1 ? st->nph_glob : st->nphantom;
        if (nph_test + (long long)Bmax * (p.cp.R - 1) > rb.cap_ph) { if (tid == 0) st->status = ST_NEED_PHANTOM; return false; }
        // boost_posterior: room for every phantom the next update could promote
        if (p.boost_thin > 0.0 && (long long)st->nboost + st->nphantom + (long long)Bmax * (p.cp.R - 1) > rb.cap_boost) {
            if (tid == 0) st->status = ST_NEED_BOOST;
            return false;
        }
    } else if (ndead + n > rb.cap_dead) {
        if (tid == 0) st->status = ST_NEED_DEAD;
        return false;
    }
    long long q2 = clock64();
    if (merged && more) {
        // phase D ordered the live points
    } else if (merge && more) {
        // the babies in order: bitonic, one element per thread with warp shuffles below distance 32 when they fit
        if (npB <= (int)blockDim.x) block_sort_small(sm.bkey, sm.bval, npB);
        else if (npB == 2 * (int)blockDim.x) block_sort_pair(sm.bkey, sm.bval, npB);
        else block_sort(sm.bkey, sm.bval, npB);
        // rank of a survivor = its index + number of babies before it; rank of a baby = its index + number of
        // survivors before it ((key, slot) pairs are distinct, so the merged order is the sorted order)
        for (int i = tid; i < m; i += nthr) {
            const double ka = sm.akey[i];
            const int va = sm.aval[i];
            int lo = 0, hi = Kp;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                const double kb = sm.bkey[mid];
                if (kb < ka || (kb == ka && sm.bval[mid] < va)) lo = mid + 1; else hi = mid;
            }
            const int rank = i + lo;
            newo[rank] = va;
            newk[rank] = ka;
        }
        for (int j = tid; j < Kp; j += nthr) {
            const double kb = sm.bkey[j];
            const int vb = sm.bval[j];
            int lo = 0, hi = m;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                const double ka = sm.akey[mid];
                bool less = ka < kb;
                if (ka == kb) less = sm.aval[mid] < vb;
                if (less) lo = mid + 1; else hi = mid;
            }
            const int rank = j + lo;
            newo[rank] = vb;
            newk[rank] = kb;
        }
        __syncthreads();
    } else {
        // full sort: first generation, after live points moved, and the final kill-off (which needs every key in order)
        double* skey = sm.akey;
        int* sval = (int*)(sm.akey + np2);
        if (merge) {  // !more on the merge layout: rebuild the flat layout
            __syncthreads();
            for (int i = tid; i < np2; i += nthr) {
                skey[i] = (i < n) ? __ldcg(rb.live + (size_t)i * T + T - 1) : INFINITY;
                sval[i] = i;
            }
            __syncthreads();
        }
        block_sort(skey, sval, np2);
        for (int i = tid; i < n; i += nthr) { newo[i] = sval[i]; newk[i] = skey[i]; }
        __syncthreads();
        if (!more) {  // final kill-off, nested_sampling.F90:381-384
            evidence_deaths(st, skey, n, n, rb.logw + ndead, sc, rb.deadn ? rb.deadn + ndead : nullptr);
            for (size_t e = tid; e < (size_t)n * T; e += nthr) {
                size_t i = e / T, c = e % T;
                rb.dead[(size_t)(ndead + i) * T + c] = __ldcg(rb.live + (size_t)sval[i] * T + c);
            }
            if (rb.deadlab)
                for (int i = tid; i < n; i += nthr) rb.deadlab[ndead + i] = __ldcg(rb.lab + sval[i]);
            if (tid == 0) { st->ndead = ndead + n; st->K = 0; st->B = 0; st->status = ST_DONE; }
            return false;
        }
    }
    long long q3 = clock64();
    const double lX_new = logX_after(st->logX, K, n);
    if (tid == 0) {
        publish_generation(p, rb, st, n, K, trim, ndead, newo, newk, lX_new);
        long long q4 = clock64();
        st->dbg[6] += q1 - q0; st->dbg[7] += q2 - q1; st->dbg[8] += q3 - q2; st->dbg[9] += q4 - q3;
    }
    return true;
}

// S2: update_evidence (run_time_info.f90:211-296) for the K deaths of the generation just published; their keys are the
// head of the order phase S1 wrote
__device__ inline void phase_S2(const KParams& p, const RunBuf& rb, DevRun* st, const SmemS& sm) {
    const int K = st->K;
    evidence_deaths<true>(st, rb.okey + st->order_off, K, st->n_gen, rb.logw + st->ndead_base, sm.sc,
                          rb.deadn ? rb.deadn + st->ndead_base : nullptr);
}

// ---------------------------------------------------------------- phase D (every CTA that runs chains)
// The order of the live points after a REGULAR generation -- K deaths, K successful births into the vacated slots -- built by
// all warps of the run instead of CTA 0 (find_min / the sorted live set of nested_sampling.F90:262-303, kept incrementally):
// the n-K survivors keep their relative order, so the rank of a survivor is its index plus the number of babies below it,
// and the rank of a baby is the number of babies below it plus the number of survivors below it.  Every CTA stages the K baby
// keys (bkeys: written by the chains, logL of their last babies; baby j sits in slot oldo[j]) and the n-K survivor keys in
// shared memory; a warp takes one live point at a time, its lanes count the babies below it (ties by slot, as everywhere),
// a binary search counts the survivors below a baby, and lane 0 writes (slot, key) to its place in the other half of
// rb.order / rb.okey.  Nothing is sorted and nothing depends on another point's rank: the phase is one pass, spread over
// cta_rel = 0 .. nctas-1.  Beside it every CTA leaves the sum of exp(logL - M0) over the points it ranked (M0: the largest
// survivor key) in rb.dpart for the termination test (live_logZ, run_time_info.f90:683-709).
__device__ inline void phase_D(const KParams& p, const RunBuf& rb, DevRun* st, const double* bkeys, int cta_rel, int nctas,
                               double* s_keys) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = blockDim.x >> 5, nthr = blockDim.x;
    const int K = vload(&st->K), n = vload(&st->n_gen), m = n - K;
    const int off = vload(&st->order_off);
    const int* oldo = rb.order + off;
    const double* oldk = rb.okey + off;
    int* newo = rb.order + (off ? 0 : p.nmax);
    double* newk = rb.okey + (off ? 0 : p.nmax);
    double* sb = s_keys;          // K baby keys
    double* ss = s_keys + K;      // m survivor keys, ascending
    double* sp = s_keys + n;      // 2 W: the warps' partials
    // (bkeys[0 .. K) and oldk[K .. n) land in sb[0 .. K), ss[0 .. m) = s_keys[0 .. n): one sweep, eight loads in flight per thread)
    for (int j0 = tid; j0 < n; j0 += 8 * nthr) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int j = j0 + u * nthr;
            v[u] = j < K ? __ldcg(bkeys + j) : (j < n ? __ldcg(oldk + j) : 0.0);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int j = j0 + u * nthr;
            if (j < n) s_keys[j] = v[u];
        }
    }
    __syncthreads();
    const double M0 = ss[m - 1];   // the largest survivor key: reference point of the termination sums (the same on every CTA)
    double ps = 0.0;               // lane 0: sum of exp(key - M0) over the warp's points
    // a warp ranks up to NQ points per pass over the baby keys (one shared-memory read serves NQ comparisons)
    constexpr int NQ = 4;
    const int first = cta_rel * W + warp, stride = nctas * W;
    for (int e0 = first; e0 < n; e0 += NQ * stride) {
        double key[NQ];
        int idx[NQ], slot[NQ], cnt[NQ];
        bool baby[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int e = e0 + q * stride;
            const bool act = e < n;
            baby[q] = act && e >= m;
            idx[q] = baby[q] ? e - m : (act ? e : -1);
            key[q] = !act ? -INFINITY : (baby[q] ? sb[idx[q]] : ss[idx[q]]);   // nothing is below -inf: an idle entry counts nothing
            slot[q] = act ? __ldcg(oldo + (baby[q] ? idx[q] : K + idx[q])) : 0;
            cnt[q] = 0;
        }
        int eqc[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) eqc[q] = 0;
#pragma unroll 4
        for (int j = lane; j < K; j += 32) {
            const double kb = sb[j];
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                cnt[q] += kb < key[q] ? 1 : 0;
                eqc[q] += kb == key[q] ? 1 : 0;
            }
        }
#pragma unroll
        for (int q = 0; q < NQ; ++q) {   // equal keys (rare; a baby always meets itself): the slot decides
            if (idx[q] < 0) continue;
            if (__reduce_add_sync(FULL, eqc[q]) > (baby[q] ? 1 : 0))
                for (int j = lane; j < K; j += 32)
                    if (sb[j] == key[q] && !(baby[q] && j == idx[q])) cnt[q] += __ldcg(oldo + j) < slot[q] ? 1 : 0;
        }
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            if (idx[q] < 0) continue;   // uniform over the warp
            const int below = __reduce_add_sync(FULL, cnt[q]);
            int rank = below + idx[q];
            if (baby[q]) {   // survivors below it: first survivor that is not
                int lo = 0, hi = m;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (ss[mid] < key[q]) lo = mid + 1; else hi = mid;
                }
                while (lo < m && ss[lo] == key[q] && __ldcg(oldo + K + lo) < slot[q]) ++lo;
                rank = below + lo;
            }
            if (lane == 0) {
                newo[rank] = slot[q];
                newk[rank] = key[q];
                ps += exp(key[q] - M0);
            }
        }
    }
    if (lane == 0) sp[warp] = ps;
    __syncthreads();
    if (tid == 0) {   // the warps' partials in warp order
        double cs = 0.0;
        for (int w = 0; w < W; ++w) cs += sp[w];
        rb.dpart[cta_rel] = cs;
        // arrive (CTA 0 waits for every ranking CTA before phase S1 reads the order; nobody else waits: the CTAs go
        // on to prepare their next chains)
        __threadfence();
        atomicAdd(&st->dbar, 1u);
    }
    __syncthreads();
}

// ---------------------------------------------------------------- phase U
// clean_phantoms (run_time_info.f90:820-877) + calculate_covmats (:601-641) at the update cadence.
// The phantom pool is cut into tiles of blockDim.x records dealt round-robin to the CTAs of the run (old
// phantoms are mostly dead, new ones mostly alive: contiguous chunks would be badly balanced).
//   pass A  per tile: survivor count (rb.pcount[tile]); only the logL column is read
//   pass B  per tile: stable compaction into the other pool at the tile's prefix offset, fused with the first
//           and second moments of the survivors' (and this CTA's share of the live points') cube coordinates
//           about a pivot c -- the mean of the previous update (the cube centre at first) -- so the data is
//           read once:  cov = S2/N - d d^T with d = S1/N, mean = c + d.  |d| is a small fraction of the
//           spread (the mean moves little between updates), so nothing cancels.
//   CTA 0 finishes (finish_update): partials in CTA order, all-reduce over the ranks of a sharded run,
//   calc_cholesky.
// Lane r of a warp owns dimension r, r+32, ... of S1 and COV_ACC entries of the packed triangle of S2.

// boost_posterior (clean_phantoms, run_time_info.f90:846-868): a phantom that phase U removes becomes, with
// probability thin_posterior, a posterior sample carrying the weight of the death since the last update with the
// smallest logL above its own; the host looks that death up in the window stored with the row (the dead points are
// in ascending logL).  The trial is addressed by the bits of the phantom's own logL, so it does not depend on the
// order the phantoms are visited in (the reference draws it from its sequential stream; any fixed addressing gives
// the same distribution, and the same run promotes the same phantoms every time).
static __device__ __noinline__ void boost_harvest(const KParams& p, const RunBuf& rb, DevRun* st, const double* rec, unsigned long long win) {
    const int D = p.cp.D, T = p.cp.T, np = T - D;
    const double l = __ldcg(rec + T - 1);
    if (!(l > __ldcg(rec + T - 2))) return;  // a baby at or below its birth contour (a failed slice) never was a phantom (run_time_info.f90:746-757)
    if (!(uniform(rb.seed, TAG_BOOST, (uint64_t)__double_as_longlong(l), 0u, 0u) < p.boost_thin)) return;
    const unsigned long long slot = atomicAdd(&st->nboost, 1ULL);
    if ((long long)slot >= rb.cap_boost) return;  // phase S1 reserved room for every phantom: not reached
    double* o = rb.boost + (size_t)slot * np;
    for (int k = 0; k < np; ++k) o[k] = __ldcg(rec + D + k);  // theta, phi, birth contour, logL
    // Device likelihoods leave a phantom's derived parameters unset (slice_chain only finishes the last baby);
    // a posterior sample needs them: gaussian.f90:37-40, as Model::finish_derived.  Host-callback runs stored
    // what the user's likelihood returned.
    const int P = p.cp.P;
    if (P > 0 && !p.host_like) {
        const bool gauss = p.cp.like_kind == LIKE_GAUSSIAN;
        double r = 0.0;
        if (gauss) {
            double r2 = 0.0;
            for (int d = 0; d < D; ++d) {
                const double dl = o[d] - __ldg(p.like_params + d);
                r2 += dl * dl;
            }
            r = sqrt(r2);
        }
        o[D] = r;
        if (P >= 2) o[D + 1] = gauss ? log(pow(r, (double)D) * p.cp.Vn) : 0.0;
        for (int i = 2; i < P; ++i) o[D + i] = 0.0;
    }
    rb.boost_win[slot] = win;
}

// A pool record survives an update when no death since the last one lies above it (clean_phantoms, run_time_info.f90:
// 842-851: the deaths' largest logL is the contour) -- and when it was a phantom in the first place: replace_point only
// keeps babies strictly above their birth contour (run_time_info.f90:746-757); on a likelihood with plateaus a baby can
// sit on the contour itself.  The chains write every baby into the pool; the others are dropped here.
__device__ __forceinline__ bool phantom_kept(const double* rec, int T, double Lstar) {
    const double l = __ldcg(rec + T - 1);
    return !(Lstar > l) && l > __ldcg(rec + T - 2);
}

__device__ inline void phase_UA(const KParams& p, const RunBuf& rb, DevRun* st, int cta, int NG) {
    const int tid = threadIdx.x, lane = tid & 31, T = p.cp.T, UT = blockDim.x, W = UT >> 5;   // a tile: one record per thread
    const long long total = vload(&st->nphantom);
    const double Lstar = vload(&st->Lstar);
    const double* src = rb.ph[vload(&st->cur_pool)];
    const long long ntiles = (total + UT - 1) / UT;
    const bool boosting = p.boost_thin > 0.0;
    const unsigned long long bwin = boosting ? ((unsigned long long)vload(&st->ndead_upd) << 32) | (unsigned long long)vload(&st->ndead) : 0ull;
    // keys of up to four tiles in flight before the first count.  Beside the tile's count, every 32-record segment
    // leaves its keep mask: pass B takes its records and its place inside the tile from the masks, without a barrier.
    for (long long t0 = cta; t0 < ntiles; t0 += 4LL * NG) {
        bool keep[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long t = t0 + (long long)j * NG, rec = t * UT + tid;
            keep[j] = t < ntiles && rec < total && phantom_kept(src + (size_t)rec * T, T, Lstar);
            if (boosting && t < ntiles && rec < total && !keep[j]) boost_harvest(p, rb, st, src + (size_t)rec * T, bwin);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long t = t0 + (long long)j * NG;
            if (t < ntiles) {  // uniform over the CTA
                const unsigned bal = __ballot_sync(FULL, keep[j]);
                if (lane == 0) rb.pmask[t * W + (tid >> 5)] = bal;
                const int c = __syncthreads_count(keep[j]);
                if (tid == 0) rb.pcount[t] = c;
            }
        }
    }
}

// Pass B.  The moments are one matrix product: with the augmented, pivot-shifted coordinate rows
// z = [x - c, 1, 0...] (Dp8 = 8*ceil((D+1)/8) entries) of the kept records, M = sum z z^T holds S2 in its leading
// D x D block, S1 in column D and the record count at (D, D).  A warp stages U_BATCH = 8 records (two k-steps of
// the m8n8k4 FP64 tensor-core MMA) in shared memory and multiplies the 8x8 tiles of the upper triangle of M,
// COV_TPP tiles per pass over the data (one pass up to D = 31).
// Record stream.  With an even record length (16-byte aligned records) a kept record moves global -> shared as one
// bulk copy (TMA, cp.async.bulk) into the warp's staging ring -- two batches of U_BATCH records, the next batch on its
// way while this one is multiplied -- and, in pass 0, from there to its place in the other pool as one bulk store; the
// data never sits in registers, and the warp's mbarriers (mbar: two per warp, parity bits in `mpar`) count the bytes.
// Odd record lengths (an odd number of derived parameters) take the register path.
__device__ inline void phase_UB(const KParams& p, const RunBuf& rb, DevRun* st, int cta, int NG, unsigned char* smem_warp0,
                                int warp_bytes, int* s_cnt, long long* tim, unsigned long long* mbar, unsigned& mpar) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = blockDim.x >> 5;
    const int D = p.cp.D, T = p.cp.T, n = vload(&st->n);
    const int Dpad = (D + 1) & ~1;
    const int Dp8 = (D + 1 + 7) & ~7, SX = Dp8 + 4;   // row stride of the staged batch: SX mod 16 in {4, 12}, conflict-free fragments
    const int nt = Dp8 >> 3, ntl = nt * (nt + 1) / 2;  // 8-wide dimension tiles, tiles of the upper triangle
    const long long total = vload(&st->nphantom);
    const double Lstar = vload(&st->Lstar);
    const int pool = vload(&st->cur_pool);
    const double* __restrict__ src = rb.ph[pool];
    double* __restrict__ dst = rb.ph[pool ^ 1];
    const long long ntiles = (total + blockDim.x - 1) / blockDim.x;   // tiles of one record per thread (phase_UA)
    const int lchunk = (n + NG - 1) / NG, l0 = min(n, cta * lchunk), l1 = min(n, l0 + lchunk);
    // per warp: [0..Dpad) pivot (warp 0's copy is used) | U_BATCH x SX staged rows.  After a pass the whole per-warp
    // area (behind the pivot) is reused as the CTA's Dp8 x Dp8 matrix the warps add their tiles to, in warp order.
    double* s_piv = (double*)smem_warp0;
    double* s_x = (double*)(smem_warp0 + (size_t)warp * warp_bytes) + Dpad;
    double* s_M = (double*)smem_warp0 + Dpad;
    long long* s_base = (long long*)(s_cnt + 16);  // [0] survivors in the tiles before this CTA's first tile, [1] all survivors
    const int UT = blockDim.x;   // records per tile (phase_UA)
    const bool tmr = (cta == 0 && tid == 0);
    const long long z0 = clock64();
    __syncthreads();
    // survivors before tile `upto` (exclusive) starting from tile `from`: one warp, coalesced; every lane gets the sum
    auto count_range = [&](long long from, long long upto) -> long long {
        long long c = 0;
        for (long long g = from + lane; g < upto; g += 32) c += __ldcg(rb.pcount + g);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
        return c;
    };
    if (warp == 0) {
        const long long before = count_range(0, min((long long)cta, ntiles));
        const long long all = before + count_range(min((long long)cta, ntiles), ntiles);
        if (lane == 0) { s_base[0] = before; s_base[1] = all; }
    }
    for (int e = tid; e < D; e += blockDim.x) s_piv[e] = __ldcg(rb.gsum + 2 + D + e);
    __syncthreads();
    const long long base = s_base[0];
    if (cta == 0 && tid == 0) st->ph_kept = s_base[1];
    const int JT = (T + 31) >> 5, JD = (D + 31) >> 5;  // 32-wide column chunks of a record / of its cube coordinates
    const bool bulk = (T & 1) == 0 && p.u_bulk != 0;   // 16-byte aligned records and room for the staging ring
    const int fr = lane >> 2, fk = lane & 3;           // fragment row (dimension within a tile) and k index (record within a k-step)
    double* outp = rb.partial + (size_t)cta * p.partial_stride;
    for (int pass = 0; pass < p.cov_passes; ++pass) {
        double c0[COV_TPP], c1[COV_TPP];
        int tl[COV_TPP];  // (ti << 8) | tj of the pass's tiles, -1 beyond the last
#pragma unroll
        for (int q = 0; q < COV_TPP; ++q) {
            c0[q] = c1[q] = 0.0;
            const int idx = pass * COV_TPP + q;
            int tj = 0;
            while ((tj + 1) * (tj + 2) / 2 <= idx) ++tj;  // tiles ordered (0,0) (0,1) (1,1) (0,2) ...
            const int ti = idx - tj * (tj + 1) / 2;
            tl[q] = idx < ntl ? ((ti << 8) | tj) : -1;
        }
        for (int e = lane; e < U_BATCH * SX; e += 32) s_x[e] = 0.0;  // the padding columns stay zero
        __syncwarp();
        // The CTA's work items: its phantom tiles (t = cta, cta + NG, ...), then its share of the live points cut
        // into pseudo-tiles of UT records (never copied).  A warp takes the kept records of its 32-record segment
        // (the mask pass A left) in batches of U_BATCH: all loads of a batch are issued before the first use.  Pass 0
        // copies the phantom records to their place in the other pool (stable compaction): the tile's offset is the
        // prefix of the tile counts, the segment's place inside the tile the population of the masks before it.  The
        // warps of a CTA run through their items independently; the next item's masks and counts are fetched while the
        // current one is worked on.
        long long tbase = base;
        const long long z1 = clock64();
        if (tmr) tim[12] += z1 - z0;
        const long long my_tiles = (ntiles > cta) ? (ntiles - cta + NG - 1) / NG : 0;
        const int my_live = (p.sh.rank == 0) ? (l1 - l0) : 0;  // the live points are replicated: rank 0 counts them
        const long long my_items = my_tiles + (my_live + UT - 1) / UT;
        // masks of the tile's segments (lane w holds segment w's) and, per lane, its share of the counts of the tiles
        // between this tile and the CTA's next one
        auto fetch = [&](long long t, unsigned& mk, long long& cr) {
            mk = (lane < W && t < ntiles) ? __ldcg(rb.pmask + t * W + lane) : 0u;
            cr = 0;
            if (pass == 0)
                for (long long g = t + lane; g < min(t + NG, ntiles); g += 32) cr += __ldcg(rb.pcount + g);
        };
        unsigned mk_n = 0;
        long long cr_n = 0;
        if (my_tiles > 0) fetch(cta, mk_n, cr_n);
        // The batches of this warp, in order: an item's kept records (the mask pass A left, or the live pseudo-tile's
        // extent) in groups of U_BATCH.  next_batch() walks items and masks; a batch is (first record of the segment,
        // mask of its records, count, copy?, place in the other pool).
        struct UBatch { const double* rbase; unsigned bm; int nb; bool copy; long long place; };
        long long it = 0, tnext = 0;
        unsigned rem = 0;
        int kk = 0, woff = 0;
        bool it_copy = false, it_open = false;
        const double* it_rbase = src;
        auto next_batch = [&](UBatch& b) -> bool {
            while (rem == 0) {
                if (it_open && it_copy) tbase = tnext;   // offset of this CTA's next phantom tile
                it_open = false;
                if (it >= my_items) return false;
                const bool is_ph = it < my_tiles;
                const long long t = cta + it * NG;
                it_copy = is_ph && pass == 0;
                woff = 0;
                if (is_ph) {
                    const unsigned mk = mk_n;
                    long long cr = cr_n;
                    if (it + 1 < my_tiles) fetch(t + NG, mk_n, cr_n);   // in flight while this tile is worked on
                    rem = __shfl_sync(FULL, mk, warp);
                    woff = __reduce_add_sync(FULL, lane < warp ? __popc(mk) : 0);
                    if (it_copy) {
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) cr += __shfl_xor_sync(FULL, cr, o);
                        tnext = tbase + cr;
                    }
                    it_rbase = src + (size_t)(t * UT + warp * 32) * T;
                } else {
                    const int lrec = l0 + (int)(it - my_tiles) * UT + tid;
                    rem = __ballot_sync(FULL, lrec < l1);
                    it_rbase = rb.live + (size_t)(l0 + (int)(it - my_tiles) * UT + warp * 32) * T;
                }
                kk = 0;
                it_open = true;
                ++it;
            }
            unsigned bm = 0;
            int nb = 0;
#pragma unroll
            for (int b2 = 0; b2 < U_BATCH; ++b2)
                if (rem) { bm |= rem & (0u - rem); rem &= rem - 1; ++nb; }
            b.rbase = it_rbase; b.bm = bm; b.nb = nb; b.copy = it_copy; b.place = tbase + woff + kk;
            kk += nb;
            return true;
        };
        // the DMMAs of a staged batch
        auto multiply = [&]() {
#pragma unroll
            for (int ks = 0; ks < U_BATCH / 4; ++ks) {
                const double* row = s_x + (4 * ks + fk) * SX + fr;
#pragma unroll
                for (int q = 0; q < COV_TPP; ++q)
                    if (tl[q] >= 0) dmma884(c0[q], c1[q], row[(tl[q] >> 8) << 3], row[(tl[q] & 0xff) << 3]);
            }
        };
        auto move_labels = [&](const UBatch& b) {   // the phantom's cluster label moves with it
            if (b.copy && p.clustering && lane < b.nb) {
                const int bit = __fns(b.bm, 0, lane + 1);
                rb.phl[pool ^ 1][b.place + lane] = __ldcg(rb.phl[pool] + ((b.rbase - src) / T + bit));
            }
        };
        if (bulk) {
            double* s_stage = s_x + U_BATCH * SX;   // 2 x U_BATCH x T doubles behind the staged rows
            unsigned long long* bar = mbar + 2 * warp;
            fence_proxy_async();   // the area was last written through the generic proxy (moment matrix of the pass before)
            // what moves: the cube coordinates (rounded up to 16 bytes); in the compaction pass the whole record, or -- narrow
            // phantoms (ChainParams::ph_narrow) -- the cube coordinates and the record's last 16 bytes, [birth, logL]
            const unsigned cube_bytes = (unsigned)(((D + 1) & ~1) * 8);
            const bool narrow = p.cp.ph_narrow != 0 && cube_bytes + 16u < (unsigned)(T * 8);
            auto issue = [&](const UBatch& b, int buf) {
                const unsigned bytes = (b.copy && !narrow) ? (unsigned)(T * 8) : cube_bytes;
                bulk_wait_read0();   // this lane's store out of the slot it is about to refill (issued a batch ago) has read it
                if (lane == 0) mbar_expect_tx(bar + buf, (bytes + ((b.copy && narrow) ? 16u : 0u)) * (unsigned)b.nb);
                __syncwarp();
                if (lane < b.nb) {
                    const int bit = __fns(b.bm, 0, lane + 1);
                    double* slot = s_stage + ((size_t)buf * U_BATCH + lane) * T;
                    const double* rec = b.rbase + (size_t)bit * T;
                    bulk_g2s(slot, rec, bytes, bar + buf);
                    if (b.copy && narrow) bulk_g2s(slot + (T - 2), rec + (T - 2), 16u, bar + buf);
                }
            };
            UBatch cur, nxt;
            bool have = next_batch(cur);
            int buf = 0;
            if (have) issue(cur, 0);
            while (have) {
                const bool have_n = next_batch(nxt);
                if (have_n) issue(nxt, buf ^ 1);
                mbar_wait(bar + buf, (mpar >> buf) & 1u);
                mpar ^= 1u << buf;
                const double* sb = s_stage + (size_t)buf * U_BATCH * T;
                if (cur.copy) {   // stable compaction: the record goes to its place in the other pool as it lies in the ring
                    if (lane < cur.nb) {
                        double* out = dst + (size_t)(cur.place + lane) * T;
                        const double* slot = sb + (size_t)lane * T;
                        if (narrow) {
                            bulk_s2g(out, slot, cube_bytes);
                            bulk_s2g(out + (T - 2), slot + (T - 2), 16u);
                        } else {
                            bulk_s2g(out, slot, (unsigned)(T * 8));
                        }
                    }
                    bulk_commit();
                    move_labels(cur);
                }
                for (int j = 0; j < JD; ++j) {
                    const int e = lane + 32 * j;
                    const double pv = e < D ? s_piv[e] : 0.0;
#pragma unroll
                    for (int b2 = 0; b2 < U_BATCH; ++b2)   // rows beyond the batch are zero; column D carries the 1 of the augmented row
                        if (e <= D) s_x[b2 * SX + e] = b2 < cur.nb ? (e < D ? sb[b2 * T + e] - pv : 1.0) : 0.0;
                }
                if ((D & 31) == 0 && lane < U_BATCH) s_x[lane * SX + D] = lane < cur.nb ? 1.0 : 0.0;
                __syncwarp();
                multiply();
                __syncwarp();
                cur = nxt;
                have = have_n;
                buf ^= 1;
            }
            bulk_wait0();          // this lane's stores are complete ...
            fence_proxy_async();   // ... and ordered before what follows in the generic proxy (the barrier's release)
        } else {
            UBatch b;
            while (next_batch(b)) {
                const double* rp[U_BATCH];
                unsigned m2 = b.bm;
#pragma unroll
                for (int b2 = 0; b2 < U_BATCH; ++b2) {
                    rp[b2] = b.rbase;
                    if (m2) { rp[b2] = b.rbase + (size_t)(__ffs(m2) - 1) * T; m2 &= m2 - 1; }
                }
                double* out = dst + (size_t)b.place * T;
                move_labels(b);
                __syncwarp();
                const int jmax = b.copy ? JT : JD;
                for (int j = 0; j < jmax; ++j) {
                    const int e = lane + 32 * j;
                    double v[U_BATCH];
#pragma unroll
                    for (int b2 = 0; b2 < U_BATCH; ++b2) v[b2] = (b2 < b.nb && e < T) ? __ldcg(rp[b2] + e) : 0.0;
                    const double pv = e < D ? s_piv[e] : 0.0;
#pragma unroll
                    for (int b2 = 0; b2 < U_BATCH; ++b2) {
                        if (b.copy && b2 < b.nb && e < T) out[(size_t)b2 * T + e] = v[b2];
                        // rows beyond the batch are zero; column D carries the 1 of the augmented row
                        if (e <= D) s_x[b2 * SX + e] = b2 < b.nb ? (e < D ? v[b2] - pv : 1.0) : 0.0;
                    }
                }
                if ((D & 31) == 0 && lane < U_BATCH) s_x[lane * SX + D] = lane < b.nb ? 1.0 : 0.0;  // column D starts a chunk the loop above did not reach
                __syncwarp();
                multiply();
            }
        }
        const long long z3 = clock64();
        if (tmr) tim[13] += z3 - z1;
        // the warps add their tiles to the CTA's matrix in warp order (deterministic), then the pass's entries go
        // to this CTA's partial: [0] count, [1..1+D) S1, then the packed triangle of S2 (entry (a <= b) at b(b+1)/2 + a)
        __syncthreads();
        for (int e = tid; e < Dp8 * Dp8; e += blockDim.x) s_M[e] = 0.0;
        __syncthreads();
        for (int w = 0; w < W; ++w) {
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
        for (int e = tid; e < COV_TPP * 64; e += blockDim.x) {
            const int idx = pass * COV_TPP + (e >> 6);
            if (idx >= ntl) continue;
            int tj = 0;
            while ((tj + 1) * (tj + 2) / 2 <= idx) ++tj;
            const int ti = idx - tj * (tj + 1) / 2;
            const int a2 = (ti << 3) + ((e >> 3) & 7), b2 = (tj << 3) + (e & 7);
            if (a2 > b2 || b2 > D) continue;
            const double v = s_M[a2 * Dp8 + b2];
            if (b2 == D) outp[a2 == D ? 0 : 1 + a2] = v;
            else outp[1 + D + b2 * (b2 + 1) / 2 + a2] = v;
        }
        __syncthreads();
        if (tmr) tim[15] += clock64() - z3;
    }
}

// ---------------------------------------------------------------- update finalisation (CTA 0)
__device__ inline bool finish_update(const KParams& p, const RunBuf& rb, DevRun* st, int NG, double* s_cov, double* s_L, long long* tim) {
    const int tid = threadIdx.x, D = p.cp.D, ntri = p.ntri;
    const long long tot = st->ph_kept;
    const long long f0 = clock64();
    __shared__ int s_ok;
    __shared__ double s_N;
    auto unpack = [](int idx, int& ai, int& bi) {
        ai = (int)((sqrt(8.0 * idx + 1.0) - 1.0) * 0.5);
        while (ai * (ai + 1) / 2 > idx) --ai;
        while ((ai + 1) * (ai + 2) / 2 <= idx) ++ai;
        bi = idx - ai * (ai + 1) / 2;
    };
    // scratch until the matrix is formed: the packed triangle of S2 in s_L, d = S1/N behind the matrix in s_cov
    double* s_S2 = s_L;             // ntri <= D*D
    double* s_d = s_cov + D * D;    // D (the layout reserves it, make_layout_w)
    // this rank's moments: S1 and S2 partials summed in CTA order (coalesced over the entries, 8 CTAs in flight)
    for (int e = tid; e < D + ntri; e += blockDim.x) {
        double s = 0.0;
        for (int g0 = 0; g0 < NG; g0 += 32) {  // 32 partials in flight, added in CTA order
            double v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = (g0 + j < NG) ? __ldcg(rb.partial + (size_t)(g0 + j) * p.partial_stride + 1 + e) : 0.0;
#pragma unroll
            for (int j = 0; j < 32; ++j) s += v[j];
        }
        if (p.sh.world > 1) {
            for (int q = 0; q < p.sh.world; ++q) p.sh.xpart[q][(size_t)p.sh.rank * p.sh.xstride + 2 + e] = s;
        } else if (e < D) s_d[e] = s; else s_S2[e - D] = s;
    }
    const long long f1 = clock64();
    double Nglob = (double)tot;  // surviving phantoms of all ranks
    if (p.sh.world > 1) {
        if (tid == 0)
            for (int q = 0; q < p.sh.world; ++q) p.sh.xpart[q][(size_t)p.sh.rank * p.sh.xstride] = (double)tot;
        __syncthreads();
        if (tid == 0) s_ok = xgpu_barrier(p.sh, st) ? 1 : 0;
        __syncthreads();
        if (!s_ok) return false;
        const double* mine = p.sh.xpart[p.sh.rank];
        for (int e = tid; e < D + ntri; e += blockDim.x) {
            double s = 0.0;
            for (int r = 0; r < p.sh.world; ++r) s += __ldcg(mine + (size_t)r * p.sh.xstride + 2 + e);
            if (e < D) s_d[e] = s; else s_S2[e - D] = s;
        }
        if (tid == 0) {
            double c = 0.0;
            for (int r = 0; r < p.sh.world; ++r) c += __ldcg(mine + (size_t)r * p.sh.xstride);
            s_N = c;
        }
        __syncthreads();
        Nglob = s_N;
    }
    __syncthreads();
    const double N = (double)st->n + Nglob;
    for (int e = tid; e < D; e += blockDim.x) {
        const double d = s_d[e] / N;
        s_d[e] = d;
        const double mean = __ldcg(rb.gsum + 2 + D + e) + d;
        rb.gsum[2 + e] = mean;       // the mean of live + phantom cube coordinates
        rb.gsum[2 + D + e] = mean;   // ... is the pivot of the next update
    }
    __syncthreads();
    // cov = S2/N - d d^T (calculate_covmats divides by N, not N-1: run_time_info.f90:601-641)
    for (int idx = tid; idx < ntri; idx += blockDim.x) {
        int ai, bi;
        unpack(idx, ai, bi);
        const double v = s_S2[idx] / N - s_d[ai] * s_d[bi];
        s_cov[ai + bi * D] = v;
        s_cov[bi + ai * D] = v;
    }
    __syncthreads();
    const long long f2 = clock64();
    for (int e = tid; e < D * D; e += blockDim.x) rb.cov[e] = s_cov[e];
    __syncthreads();
    const int fb = block_cholesky(s_cov, s_L, D);  // calc_cholesky (utils.F90:621-649) in shared memory, by the whole CTA
    for (int e = tid; e < D * D; e += blockDim.x) rb.chol[e] = s_L[e];
    if (tid == 0) {
        const long long f3 = clock64();
        tim[16] += f1 - f0; tim[17] += f2 - f1; tim[18] += f3 - f2;
        rb.gsum[0] = Nglob;
        st->chol_fallback += fb;
        st->cov_N = N;
        st->nphantom = tot;
        st->nph_glob = (long long)Nglob;
        st->cur_pool ^= 1;
        st->nupdates += 1;
        st->logX_last_update = st->logX;
        st->ndead_upd = st->ndead;
        st->update_pending = 0;
    }
    __syncthreads();
    return true;
}

// ---------------------------------------------------------------- sharded run: last-baby exchange
// A chain warp stores its last baby (already in this rank's incoming buffer) into every peer's incoming buffer, and the
// baby's logL into every rank's key buffer (phase D ranks the babies of all ranks from there, on every rank).
__device__ inline void shard_publish(const KParams& p, const double* rec, int k, int parity, double lfin) {
    const int lane = threadIdx.x & 31, T = p.cp.T;
    __syncwarp();
    for (int q = 0; q < p.sh.world; ++q) {
        if (q != p.sh.rank) {
            double* dst = p.sh.xin[q] + ((size_t)parity * p.batch_K + k) * T;
            for (int e = lane; e < T; e += 32) dst[e] = __ldcg(rec + e);
        }
        if (lane == 0) p.sh.xrun[q][(size_t)parity * p.batch_K + k] = lfin;
    }
    __threadfence_system();
}
// CTA 0, after this rank's chains are done: the cross-GPU barrier that closes the generation (every rank's last babies and
// keys have arrived here when it returns).  Returns false when a peer did not arrive.
__device__ inline bool shard_close(const KParams& p, DevRun* st) {
    __shared__ int s_ok;
    const int tid = threadIdx.x;
    __syncthreads();
    if (tid == 0) s_ok = xgpu_barrier(p.sh, st) ? 1 : 0;   // (system-scope fence inside)
    __syncthreads();
    if (!s_ok && tid == 0) st->status = ST_ERROR;
    __syncthreads();
    return s_ok != 0;
}
// Every warp of the run, after shard_close + a group barrier: the K last babies of all ranks are copied from this
// rank's incoming buffer into the vacated live slots (the same on every rank), one record per warp.
__device__ inline void shard_scatter(const KParams& p, const RunBuf& rb, DevRun* st, int gw, int GW) {
    const int lane = threadIdx.x & 31, T = p.cp.T;
    const int K = vload(&st->K);
    const int* ord = rb.order + vload(&st->order_off);
    const double* in = p.sh.xin[p.sh.rank] + (size_t)(vload(&st->ngen) & 1) * p.batch_K * T;
    for (int k = gw; k < K; k += GW) {
        const int slot = __ldcg(ord + k);
        for (int e = lane; e < T; e += 32) rb.live[(size_t)slot * T + e] = __ldcg(in + (size_t)k * T + e);
    }
}

// ---------------------------------------------------------------- dump hand-over (CTA 0)
// dump (nested_sampling.F90:546-590) is called at every update (:335).  The kernel does not stop for it: it
// waits until the host has consumed the previous dump, snapshots the live points, publishes the state through
// the mapped control block and goes on.  Returns true when the host asked to abort.
// The live snapshot of the next dump: every CTA of the run copies its share after the update's last barrier, while CTA 0
// finishes the covariance -- the copy (n x T doubles) used to sit on CTA 0 alone, on the critical path of every update.
// Two halves by dump parity: the half being written belongs to the dump before the last, which the host has acknowledged
// (publish_dump waits for the acknowledgement of dump s-1 before it publishes dump s).
__device__ inline void snapshot_share(const KParams& p, const RunBuf& rb, DevRun* st, int cta, int NG) {
    // (only the columns the dumper reports, [theta, phi, birth, logL], packed: the host copies them in one piece)
    const int T = p.cp.T, D = p.cp.D, np = T - D;
    const size_t nd = (size_t)vload(&st->n) * np;
    double* snap = rb.live_snap + (size_t)((vload(&st->dump_pub) + 1) & 1) * p.nmax * np;
    const size_t chunk = (nd + NG - 1) / NG, e0 = min(nd, (size_t)cta * chunk), e1 = min(nd, e0 + chunk);
    for (size_t e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
        const size_t row = e / np;
        snap[e] = __ldcg(rb.live + row * T + D + (e - row * np));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        atomicAdd(&st->snap_arr, 1u);
    }
}
__device__ inline bool publish_dump(const KParams& p, const RunBuf& rb, DevRun* st, int NG) {
    __shared__ int s_abort;
    volatile HostCtl* ctl = rb.ctl;
    if (threadIdx.x == 0) {
        const unsigned long long seq = ctl->dump_seq;
        int ab = 0;
        while (ctl->ack_seq < seq && !(ab = ctl->abort)) __nanosleep(200);
        s_abort = ab;
    }
    __syncthreads();
    if (s_abort) return true;
    if (threadIdx.x == 0) {   // every CTA of the run has written its share of the snapshot (snapshot_share)
        const unsigned want = (unsigned)(st->dump_pub + 1) * (unsigned)NG;
        unsigned v;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(&st->snap_arr) : "memory"); } while ((int)(v - want) < 0);
        st->dump_pub = st->dump_pub + 1;
        ctl->ndead = st->ndead;
        ctl->nlive = st->n;
        ctl->nlike = st->nlike;
        ctl->logZ = st->logZ;
        ctl->logZ2 = st->logZ2;
        __threadfence_system();
        ctl->dump_seq = ctl->dump_seq + 1;
        __threadfence_system();
    }
    __syncthreads();
    return false;
}

// ---------------------------------------------------------------- the persistent run kernel
//
// One generation, seen from a warp (NG = CTAs of the run's group, GW = NG * warps):
//   [CTA 0]   wait until every warp arrived (wbar) -> finish a pending update -> phase S -> group barrier B
//   [others]  chains -> arrive (wbar) -> prepare the directions of the NEXT generation's chain -> barrier B
// so the counter-addressed direction/uniform preparation of a chain overlaps the bookkeeping of CTA 0
// and the wait for the slowest chain.  Generations at the update cadence insert phase U (all CTAs)
// between the arrival and the bookkeeping.
// MODE 0: a chain per warp (speculative rounds, helper warps; pc_chain.cuh) -- a run alone on the device.
// MODE 1: a chain per point group (pc_dense.cuh) -- an ensemble that fills the device; two CTAs per SM (128 registers).
template <int G, int DPL, int KIND, int MODE>
__global__ void __launch_bounds__(256, MODE == 1 ? 2 : 1) pc_run_kernel(const __grid_constant__ KParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int NPT = 32 / G;
    constexpr int SLB = MODE == 1 ? dense_slb(G * DPL) : 0;
    const int NG = p.ctas_per_run;
    const int run = blockIdx.x / NG, cta = blockIdx.x % NG;
    const RunBuf rb = p.runs[run];
    DevRun* st = rb.st;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = blockDim.x >> 5;
    const int gw = cta * W + warp, GW = NG * W;
    const int D = p.cp.D, T = p.cp.T, R = p.cp.R, LD = p.cp.LD;
    const int c0 = p.chain_cta0, Gc = NG - c0;

    double* s_chol = (double*)smem;
    double* s_like = (double*)(smem + p.off_like);
    unsigned char* s_warp0 = smem + p.off_warp;
    unsigned char* s_warp = s_warp0 + (size_t)warp * p.warp_bytes;
    // CTA-wide scratch of phase S overlays the per-warp area
    const SmemS smS = smem_S(s_warp0, p.nmax, p.batch_K);
    double* sc = smS.sc;
    int* s_cnt = (int*)(smem + p.off_warp - 64 * (int)sizeof(int));

    const int nlp = (p.cp.like_kind == LIKE_GAUSSIAN) ? 2 * D : (p.cp.like_kind == LIKE_CORR ? D + D * D : 0);
    for (int e = tid; e < nlp; e += blockDim.x) s_like[e] = p.like_params[e];
    // Cycle counters of the phases live in shared memory while the kernel runs (a counter in global memory costs the
    // counting thread an L2 round trip per update -- on CTA 0 that is the critical path of every generation) and are
    // added to DevRun when the kernel leaves: [0 .. 24) = DevRun::dbg, then cyc_wait, cyc_S, cyc_fin, cyc_U, cyc_prep,
    // cyc_white, cyc_slice, cyc_total.  Single writer per counter and CTA.
    __shared__ long long tim[32];
    if (tid < 32) tim[tid] = 0;
    // phase U's record stream: two mbarriers per warp (one per batch of the staging ring), their parities on the warp
    __shared__ __align__(8) unsigned long long s_mbar[16];
    unsigned mbar_parity = 0;
    if (tid < 16) mbar_init(&s_mbar[tid], 1);
    mbar_fence_init();
    __syncthreads();
    auto flush_timers = [&]() {
        __syncthreads();
        if (tid < 32 && (cta == 0 || cta == c0) && tim[tid] != 0) {
            long long* dst = tid < 24 ? &st->dbg[tid] : (&st->cyc_wait + (tid - 24));
            atomicAdd((unsigned long long*)dst, (unsigned long long)tim[tid]);
        }
    };

    const ChainScratch cs = chain_scratch(s_warp, D, R, LD, p.nh_in_smem != 0, p.cp.like_kind, NPT,
                                          (MODE == 0 && rb.nh) ? rb.nh + (size_t)gw * R * LD : nullptr, SLB);
    Model<G, DPL, KIND> M;
    M.init(p.cp, s_like, p.prior_params, cs.dvec);
    __shared__ double s_tab[MODE == 1 ? 4 * G * DPL : 1];   // dense chain phase: per-dimension constants (pc_dense.cuh)
    if constexpr (MODE == 1) {
        __syncthreads();
        dense_table_fill<G * DPL>(s_tab, D, p.cp.like_kind, s_like, p.prior_params);
        __syncthreads();
    }

    if (!vload(&st->initialised)) init_phase<G, DPL, KIND>(p, rb, st, M, cta, NG, sc);

    // chain whose directions currently sit in this warp's scratch (~0 = none), and whether they are whitened
    unsigned long long prep_uid = ~0ull;
    bool prep_white = false;
    unsigned pair_seq = 0;  // paired mode: chains this warp pair has run since the launch (buffer parity)
    bool s2_due = false;  // CTA 0: the evidence of the generation in flight is still to be accumulated
    long long chol_epoch = -1;  // st->nupdates when this CTA last loaded the Cholesky factor into shared memory

    if (p.host_like && vload(&st->host_resume)) {
        // host-callback run: the host loop ran the chains of the generation in flight; finish the generation
        // (phase U at the update cadence) exactly where the chain phase would have left it
        // (failed births / empty slots of that generation first: the covariance reads a contiguous live set)
        if (vload(&st->holes_due) || vload(&st->nfail_gen)) {
            if (cta == 0) { __syncthreads(); settle_generation(p, rb, st, s_warp0, false); }
            group_sync(&st->bar, NG, p.backoff);
            if (cta == 0 && tid == 0) { st->holes_due = 0; st->nfail_gen = 0u; }
            group_sync(&st->bar, NG, p.backoff);
        }
        if (vload(&st->do_update)) {
            phase_UA(p, rb, st, cta, NG);
            group_sync(&st->bar, NG, p.backoff);
            phase_UB(p, rb, st, cta, NG, s_warp0, p.warp_bytes, s_cnt, tim, s_mbar, mbar_parity);
            if (cta == 0 && tid == 0) st->update_pending = 1;
        }
        group_sync(&st->bar, NG, p.backoff);
        if (cta == 0 && tid == 0) st->host_resume = 0;
        group_sync(&st->bar, NG, p.backoff);
    }

    const bool timer = (tid == 0) && (cta == 0);          // bookkeeping phases
    const bool ctimer = (tid == 0) && (cta == c0);         // chain phases of one representative warp
    const long long t_start = clock64();
    unsigned int dtarget = vload(&st->dbar);   // CTA 0: arrivals phase D has to show (no phase D is in flight at a launch)
    bool d_done = false;   // phase D ordered the live points of the generation just finished (the same on every CTA of the run)
    for (;;) {
        bool dump_exit = false, cluster_exit = false;
        if (cta == 0) {
            // (every chain of the previous generation is written: all warps waited for that at the end of the last pass)
            long long t1 = clock64();
            __syncthreads();
            // births that failed, slots left empty (B != K): the live set is made contiguous again before anything reads it
            if (st->holes_due || vload(&st->nfail_gen)) { settle_generation(p, rb, st, s_warp0, true); __syncthreads(); }
            if (st->update_pending) {
                if (!finish_update(p, rb, st, NG, smS.akey, s_chol, tim) && tid == 0) st->status = ST_ERROR;
                __syncthreads();
                if (p.clustering) cluster_exit = true;             // leave: the host runs the clustering pass (pc_cluster.cuh), dumps, relaunches
                else if (rb.ctl) dump_exit = publish_dump(p, rb, st, NG);   // the kernel keeps running
                else if (p.want_dump) dump_exit = true;              // "sync_dump": leave, the host dumps and relaunches
            }
            if (timer) tim[26] += clock64() - t1;
        }
        if (cta == 0) {
            if (d_done) {   // phase D complete on every ranking CTA: phase S1 reads the order they left
                const long long tb0 = clock64();
                dtarget += (unsigned)(NG - c0);
                if (tid == 0) {
                    unsigned int v;
                    int backoff = p.backoff;
                    for (;;) {
                        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(&st->dbar) : "memory");
                        if ((int)(v - dtarget) >= 0) break;
                        if (backoff) { __nanosleep(backoff); if (backoff < 4096) backoff <<= 1; }
                    }
                    __threadfence();
                }
                __syncthreads();
                if (timer) tim[23] += clock64() - tb0;
            }
            long long t2 = clock64();
            bool evidence_due = false;
            if (cluster_exit) {
                if (tid == 0) st->status = ST_CLUSTER;
            } else if (dump_exit) {
                if (tid == 0) st->status = ST_DUMP;
            } else if (vload(&st->status) != ST_ERROR) {
                // after phase D the common case is decided by one warp; everything else by the full phase S1
                bool fast = false;
                if (d_done) {
                    __shared__ int s_fast;
                    if (warp == 0) {
                        const long long tf0 = clock64();
                        const bool f = phase_S1_fast(p, rb, st, NG - c0);
                        if (lane == 0) { s_fast = f ? 1 : 0; tim[6] += clock64() - tf0; tim[7] += tf0 - t2; tim[8] -= clock64(); tim[9] -= clock64(); }
                    }
                    if (tid == 32) tim[14] += clock64() - t2;   // (probe) warp 1 reaches the barrier
                    __syncthreads();
                    if (timer) tim[9] += clock64();              // (probe) with tim[8]: S1-fast end -> past the barrier
                    fast = s_fast != 0;
                }
                evidence_due = fast ? true : phase_S1(p, rb, st, smS, d_done, NG - c0);
                if (evidence_due && c0 == 0) { __syncthreads(); phase_S2(p, rb, st, smS); evidence_due = false; }
            }
            s2_due = evidence_due;
            __syncthreads();
            if (timer) { tim[25] += clock64() - t2; if (d_done) tim[8] += clock64(); }
            prep_uid = ~0ull;  // phase S overlays this CTA's chain scratch
        }
        d_done = false;
        const long long tg0 = clock64();
        group_sync(&st->bar, NG, p.backoff);
        const long long tg1 = clock64();
        if (ctimer) tim[10] += tg1 - tg0;
        // the run's status and the generation's parameters: one load per lane, one latency
        unsigned long long pw = 0;
        if (lane < 10) pw = __ldcg(&st->pub[lane]);
        else if (lane == 10) pw = (unsigned long long)(unsigned)vload(&st->status);
        if ((int)__shfl_sync(FULL, pw, 10) != ST_RUNNING) {
            if (timer) tim[31] += clock64() - t_start;
            flush_timers();
            return;
        }
        if (cta == 0 && s2_due) {  // the chains are running: the evidence bookkeeping is off their critical path
            phase_S2(p, rb, st, smS);
            s2_due = false;
        }
        if (p.host_like) {  // the chains of this generation belong to the host loop (pc_hostchain.cuh)
            if (cta == 0) {
                __syncthreads();
                if (tid == 0) st->status = ST_HOSTCHAINS;
            }
            flush_timers();
            return;
        }
        const bool sharded = p.sh.world > 1;
        const int xw = sharded ? p.sh.world : 1, xr = sharded ? p.sh.rank : 0;
        {   // The dying points move to the dead list (run_time_info.f90:789-817).  A chain copies the point whose slot its
            // baby takes (before it writes there); the deaths without a chain of their own -- all of them in a sharded
            // run, whose last babies arrive through the incoming buffers, and those beyond the B births when the live
            // count shrinks -- are copied here, one record per warp.
            const int K_ = (int)(unsigned)__shfl_sync(FULL, pw, 5), B_ = (int)(__shfl_sync(FULL, pw, 8) >> 32);
            const long long nb = (long long)__shfl_sync(FULL, pw, 1);
            const int* ord = rb.order + (int)(unsigned)__shfl_sync(FULL, pw, 6);
            for (int k = (sharded ? 0 : B_) + gw; k < K_; k += GW) {
                const int slot = __ldcg(ord + k);
                for (int e = lane; e < T; e += 32) rb.dead[(size_t)(nb + k) * T + e] = __ldcg(rb.live + (size_t)slot * T + e);
                if (rb.deadlab && lane == 0) rb.deadlab[nb + k] = __ldcg(rb.lab + slot);
            }
        }

        // ---------------- phase C: chains, one warp each ----------------
        const unsigned long long w5 = __shfl_sync(FULL, pw, 5), w6 = __shfl_sync(FULL, pw, 6), w7 = __shfl_sync(FULL, pw, 7);
        const int K = (int)(unsigned)w5;                       // deaths of the generation
        const unsigned long long w8 = __shfl_sync(FULL, pw, 8);
        const int n = (int)(unsigned)w8, B = (int)(w8 >> 32);  // live points at its start, births (chains)
        const double Lstar = __longlong_as_double((long long)__shfl_sync(FULL, pw, 0));
        const long long ndead_base = (long long)__shfl_sync(FULL, pw, 1), nph_base = (long long)__shfl_sync(FULL, pw, 2);
        const long long nchains_base = (long long)__shfl_sync(FULL, pw, 3);
        const long long ngen_now = (long long)__shfl_sync(FULL, pw, 4);
        const int do_update = (int)(w5 >> 32);
        const bool clustered = p.clustering && (int)(unsigned)w7 > 1;
        const int cur_pool_now = (int)(w6 >> 32);
        const int* order = rb.order + (int)(unsigned)w6;
        double* pool = rb.ph[cur_pool_now];
        {   // the Cholesky factor only changes at an update: reload it then (CTA 0 also uses the area as scratch)
            const long long nup = (long long)(w7 >> 32);
            if (cta == 0 || nup != chol_epoch) {
                for (int e = tid; e < D * D; e += blockDim.x) s_chol[e] = __ldcg(rb.chol + e);
                __syncthreads();
                chol_epoch = nup;
            }
        }
        const long long ts_a = clock64();
        if (ctimer) tim[19] += ts_a - tg1;   // release -> generation parameters and Cholesky factor in place
        unsigned long long nlike = 0, nfail = 0;
        const int m = n - K;
        // sharded run: the last babies go to the incoming buffers (by generation parity) instead of the live slots
        const int xpar = (int)(ngen_now & 1);
        double* xin_mine = sharded ? p.sh.xin[xr] + (size_t)xpar * p.batch_K * T : nullptr;
        int knext = -1;  // first chain this warp prepares for the next generation
        if constexpr (MODE == 1) {
            // ---- dense chain phase (pc_dense.cuh): this warp's 32/G point groups run one chain each ----
            if constexpr (KIND != LIKE_CORR) {
                constexpr int GD = G * DPL;
                const int grp = lane / G, sub = lane % G;
                double* gblocks = rb.nh + (size_t)gw * NPT * R * SLB;   // the slice records of this warp's chains
                for (int base = gw * NPT; base < B; base += GW * NPT) {
                    // directions of the pass's chains, one chain at a time through the warp's shared-memory scratch
                    for (int g = 0; g < NPT && base + g < B; ++g) {
                        const unsigned long long uidg = (unsigned long long)(nchains_base + base + g);
                        prep_chain<GD>(D, R, LD, rb.seed, uidg, cs, &p.cp);
                        const double* Lf = s_chol;
                        if (clustered) {  // the factor of the seed's cluster
                            const double ug = uniform(rb.seed, TAG_SEED, uidg, 0u, 0u);
                            const int cg = max(1, min(m, (int)ceil(ug * (double)m)));
                            Lf = rb.cchol + (size_t)min(__ldcg(rb.lab + __ldcg(order + K + cg - 1)), MAX_CLUSTERS - 1) * D * D;
                        }
                        whiten_chain<(G * DPL + 7) / 8>(D, R, LD, Lf, cs);
                        dense_store<GD>(R, LD, rb.seed, uidg, cs, gblocks + (size_t)g * R * SLB);
                    }
                    __syncwarp();
                    const int k = base + grp;
                    const bool active = k < B;
                    const int kc = min(k, B - 1);
                    const unsigned long long uid = (unsigned long long)(nchains_base + kc);
                    const double u = uniform(rb.seed, TAG_SEED, uid, 0u, 0u);  // GenerateSeed, generate.F90:19-55
                    const int choice = max(1, min(m, (int)ceil(u * (double)m)));
                    const int src = __ldcg(order + K + choice - 1);
                    const int dslot = kc < K ? __ldcg(order + kc) : n + (kc - K);   // a vacated slot, or one appended (the live count grows)
                    const int plab = clustered ? min(__ldcg(rb.lab + src), MAX_CLUSTERS - 1) : 0;
                    double x[DPL];
#pragma unroll
                    for (int j = 0; j < DPL; ++j) x[j] = M.valid(j) ? __ldcg(rb.live + (size_t)src * T + M.dim(j)) : 0.0;
                    // the dying point moves to the dead list before its slot is reused (run_time_info.f90:789-817)
                    if (active && k < K) {
                        for (int e = sub; e < T; e += G)
                            rb.dead[(size_t)(ndead_base + k) * T + e] = __ldcg(rb.live + (size_t)dslot * T + e);
                        if (rb.deadlab && sub == 0) rb.deadlab[ndead_base + k] = __ldcg(rb.lab + dslot);
                    }
                    __syncwarp();
                    double lfin = 0.0;
                    slice_chains_dense<G, DPL, KIND>(p.cp, M, rb.seed, uid, active, x, Lstar, gblocks + (size_t)grp * R * SLB,
                                                     cs.stage + (size_t)grp * 2 * SLB,
                                                     pool + (size_t)(nph_base + (long long)kc * (R - 1)) * T,
                                                     rb.live + (size_t)dslot * T, nlike, lfin, s_tab);
                    if (p.clustering && active) {  // the babies carry their seed's label until the next update
                        for (int e = sub; e < R - 1; e += G) rb.phl[cur_pool_now][nph_base + (long long)k * (R - 1) + e] = plab;
                        if (sub == 0) rb.lab[dslot] = plab;
                    }
                    if (active && sub == 0) {   // a failed birth: settle_generation takes it out of the live set
                        const int failed = !(lfin > Lstar);
                        rb.cfail[k] = failed;
                        rb.bkey[k] = lfin;   // phase D ranks the babies from here
                        if (failed) { ++nfail; atomicAdd(&st->nfail_gen, 1u); }
                    }
                }
                // the per-group counts (on the groups' first lanes) -> lane 0
                nlike = (unsigned long long)warp_sum_int((int)nlike);
                nfail = (unsigned long long)warp_sum_int((int)nfail);
            }
        } else if (p.paired) {
            // Warps w < W/2 run chains, warp w + W/2 is the helper of warp w: it prepares (directions, deck,
            // uniforms, whitening) the pair's next chain into the other of the pair's two scratch buffers
            // while the chain warp is slicing, so preparation leaves the critical path.
            const int HW = W >> 1, pair = warp % HW;
            const bool helper = warp >= HW;
            int nmine = 0;
            if (cta >= c0)
                for (int li = pair; ((cta - c0) + Gc * li) * xw + xr < B; li += HW) ++nmine;
            auto buf = [&](unsigned sq) -> ChainScratch {
                const int region = pair + HW * (int)(sq & 1u);
                return chain_scratch(s_warp0 + (size_t)region * p.warp_bytes, D, R, LD, p.nh_in_smem != 0, p.cp.like_kind,
                                     NPT, rb.nh ? rb.nh + ((size_t)cta * W + region) * R * LD : nullptr);
            };
            for (int j = 0; j < nmine; ++j) {
                const int cl = (cta - c0) + Gc * (pair + HW * j);  // chain ordinal on this rank
                const int k = cl * xw + xr;                        // chain of the generation (dealt k % world)
                const unsigned long long uid = (unsigned long long)(nchains_base + k);
                const ChainScratch b = buf(pair_seq + j);
                double u = uniform(rb.seed, TAG_SEED, uid, 0u, 0u);  // GenerateSeed, generate.F90:19-55
                int choice = (int)ceil(u * (double)m);
                choice = max(1, min(m, choice));
                const int src = __ldcg(order + K + choice - 1);
                const int plab = clustered ? min(__ldcg(rb.lab + src), MAX_CLUSTERS - 1) : 0;  // the seed's cluster
                const long long ts_b = clock64();
                if (ctimer && j == 0) tim[20] += ts_b - ts_a;   // seed choice (Philox + order look-up)
                if (helper) {
                    long long th0 = clock64();
                    if (prep_uid != uid) { prep_chain<G * DPL>(D, R, LD, rb.seed, uid, b, &p.cp); prep_white = false; }
                    if (clustered) whiten_chain<(G * DPL + 7) / 8>(D, R, LD, rb.cchol + (size_t)plab * D * D, b);
                    else if (!prep_white) whiten_chain<(G * DPL + 7) / 8>(D, R, LD, s_chol, b);
                    prep_uid = ~0ull;
                    if (lane == 0 && cta == c0 && pair == 0) tim[28] += clock64() - th0;
                }
                asm volatile("bar.sync %0, 64;" ::"r"(1 + pair) : "memory");  // hand-over of the buffer
                if (ctimer && j == 0) tim[21] += clock64() - ts_b;   // wait for the helper's hand-over
                if (!helper) {
                    const int dslot = k < K ? __ldcg(order + k) : n + (k - K);   // a vacated slot, or one appended (the live count grows)
                    double x[DPL];
#pragma unroll
                    for (int jj = 0; jj < DPL; ++jj) x[jj] = M.valid(jj) ? __ldcg(rb.live + (size_t)src * T + M.dim(jj)) : 0.0;
                    // the dying point moves to the dead list before its slot is reused (run_time_info.f90:789-817)
                    if (!sharded && k < K) {
                        for (int e = lane; e < T; e += 32)
                            rb.dead[(size_t)(ndead_base + k) * T + e] = __ldcg(rb.live + (size_t)dslot * T + e);
                        if (rb.deadlab && lane == 0) rb.deadlab[ndead_base + k] = __ldcg(rb.lab + dslot);
                    }
                    double* last = sharded ? xin_mine + (size_t)k * T : rb.live + (size_t)dslot * T;
                    long long tc2 = clock64();
                    if (ctimer && j == 0) tim[11] += tc2 - tg1;
                    double lfin = slice_chain<G, DPL, KIND>(p.cp, M, rb.seed, uid, x, Lstar, b,
                                                      pool + (size_t)(nph_base + (long long)cl * (R - 1)) * T, last, nlike,
                                                      (cta == c0 && warp == 0) ? tim : nullptr, false);
                    if (sharded) shard_publish(p, last, k, xpar, lfin);
                    if (p.clustering) {  // the babies carry their seed's label until the next update
                        for (int e = lane; e < R - 1; e += 32) rb.phl[cur_pool_now][nph_base + (long long)cl * (R - 1) + e] = plab;
                        if (lane == 0) rb.lab[dslot] = plab;
                    }
                    if (ctimer) tim[30] += clock64() - tc2;
                    if (lane == 0) {   // a failed birth: settle_generation takes it out of the live set
                        const int failed = !(lfin > Lstar);
                        if (!sharded) { rb.cfail[k] = failed; rb.bkey[k] = lfin; }
                        if (failed) { ++nfail; if (!sharded) atomicAdd(&st->nfail_gen, 1u); }
                    }
                }
            }
            pair_seq += (unsigned)nmine;
            if (helper && cta >= c0 && ((cta - c0) + Gc * pair) * xw + xr < p.batch_K) knext = ((cta - c0) + Gc * pair) * xw + xr;
        } else if (cta >= c0) {
            for (int li = warp;; li += W) {
                const int cl = (cta - c0) + Gc * li;
                const int k = cl * xw + xr;
                if (k >= B) break;
                const unsigned long long uid = (unsigned long long)(nchains_base + k);
                double u = uniform(rb.seed, TAG_SEED, uid, 0u, 0u);  // GenerateSeed, generate.F90:19-55
                int choice = (int)ceil(u * (double)m);
                choice = max(1, min(m, choice));
                const int src = __ldcg(order + K + choice - 1);
                const int dslot = k < K ? __ldcg(order + k) : n + (k - K);   // a vacated slot, or one appended (the live count grows)
                double x[DPL];
#pragma unroll
                for (int j = 0; j < DPL; ++j) x[j] = M.valid(j) ? __ldcg(rb.live + (size_t)src * T + M.dim(j)) : 0.0;
                // the dying point moves to the dead list before its slot is reused (run_time_info.f90:789-817)
                if (!sharded && k < K) {
                    for (int e = lane; e < T; e += 32)
                        rb.dead[(size_t)(ndead_base + k) * T + e] = __ldcg(rb.live + (size_t)dslot * T + e);
                    if (rb.deadlab && lane == 0) rb.deadlab[ndead_base + k] = __ldcg(rb.lab + dslot);
                }
                double* last = sharded ? xin_mine + (size_t)k * T : rb.live + (size_t)dslot * T;
                long long tc0 = clock64();
                if (prep_uid != uid) {
                    prep_chain<G * DPL>(D, R, LD, rb.seed, uid, cs, &p.cp);
                    prep_white = false;
                }
                long long tc1 = clock64();
                const int plab = clustered ? min(__ldcg(rb.lab + src), MAX_CLUSTERS - 1) : 0;  // the seed's cluster
                if (clustered) whiten_chain<(G * DPL + 7) / 8>(D, R, LD, rb.cchol + (size_t)plab * D * D, cs);
                else if (!prep_white) whiten_chain<(G * DPL + 7) / 8>(D, R, LD, s_chol, cs);
                prep_uid = ~0ull;
                long long tc2 = clock64();
                double lfin = slice_chain<G, DPL, KIND>(p.cp, M, rb.seed, uid, x, Lstar, cs,
                                                  pool + (size_t)(nph_base + (long long)cl * (R - 1)) * T, last, nlike, nullptr, false);
                if (sharded) shard_publish(p, last, k, xpar, lfin);
                if (p.clustering) {
                    for (int e = lane; e < R - 1; e += 32) rb.phl[cur_pool_now][nph_base + (long long)cl * (R - 1) + e] = plab;
                    if (lane == 0) rb.lab[dslot] = plab;
                }
                if (ctimer) {
                    long long tc3 = clock64();
                    tim[28] += tc1 - tc0; tim[29] += tc2 - tc1; tim[30] += tc3 - tc2;
                }
                if (lane == 0) {   // a failed birth: settle_generation takes it out of the live set
                    const int failed = !(lfin > Lstar);
                    if (!sharded) { rb.cfail[k] = failed; rb.bkey[k] = lfin; }
                    if (failed) { ++nfail; if (!sharded) atomicAdd(&st->nfail_gen, 1u); }
                }
            }
            if (cta != 0 && ((cta - c0) + Gc * warp) * xw + xr < p.batch_K) knext = ((cta - c0) + Gc * warp) * xw + xr;
        }
        if (lane == 0) {
            if (nlike) atomicAdd((unsigned long long*)&st->nlike, nlike);
            if (nfail) atomicAdd((unsigned long long*)&st->nfail, nfail);
        }
        const unsigned int wtarget = warp_arrive(&st->wbar, GW);
        // the scratch the next generation's first chain of this warp (pair) will use
        const bool will_chain = knext >= 0;
        const ChainScratch csn = p.paired ? chain_scratch(s_warp0 + (size_t)((warp % (W >> 1)) + (W >> 1) * (int)(pair_seq & 1u)) * p.warp_bytes,
                                                          D, R, LD, p.nh_in_smem != 0, p.cp.like_kind, NPT,
                                                          rb.nh ? rb.nh + ((size_t)cta * W + (warp % (W >> 1)) + (W >> 1) * (int)(pair_seq & 1u)) * R * LD : nullptr)
                                          : cs;
        // Directions of this warp's first chain of the NEXT generation (counter-addressed: they do not depend on its
        // outcome).  A helper warp prepares them while the chains of this generation are still slicing; a warp that ran
        // chains itself does it behind phase D, beside CTA 0's bookkeeping.  Phase U stages its batches in the per-warp
        // scratch, so an update generation prepares afterwards, and whitens once the new factor is there.
        auto prep_next = [&]() {
            prep_uid = (unsigned long long)(nchains_base + B + knext);
            prep_chain<G * DPL>(D, R, LD, rb.seed, prep_uid, csn, &p.cp);
            prep_white = false;
            if (!do_update && !p.clustering) {  // with clusters the factor depends on the chain's seed, which the next phase S decides
                whiten_chain<(G * DPL + 7) / 8>(D, R, LD, s_chol, csn);
                prep_white = true;
            }
        };
        const bool prep_early = will_chain && p.paired && !do_update;
        if (prep_early) prep_next();
        // every chain of the generation is written
        const long long tw0 = clock64();
        warp_wait(&st->wbar, wtarget, p.backoff);
        const long long tw1 = clock64();
        if (timer) tim[24] += tw1 - tw0;
        if (sharded) {
            // Close the generation across the GPUs: this rank's chains are done -> CTA 0 passes the cross-GPU barrier (every
            // rank's last babies and keys have arrived) -> every warp copies its share of the K last babies from the incoming
            // buffer into the live array.
            if (cta == 0) shard_close(p, st);
            group_sync(&st->bar, NG, p.backoff);
            if (vload(&st->status) == ST_ERROR) return;
            shard_scatter(p, rb, st, gw, GW);
            if (timer) tim[14] += clock64() - tw1;
        }
        if (do_update) {
            long long tu1 = clock64();
            if (sharded) group_sync(&st->bar, NG, p.backoff);   // the covariance reads the scattered live points
            // failed births / empty slots: the live set is made contiguous before the covariance reads it (every warp
            // sees the same flags after the barrier)
            if (vload(&st->holes_due) || vload(&st->nfail_gen)) {
                if (cta == 0) { __syncthreads(); settle_generation(p, rb, st, s_warp0, false); }
                group_sync(&st->bar, NG, p.backoff);
                // every warp of the run has read the flags by now: clear them (the next reader is CTA 0's next phase S)
                if (cta == 0 && tid == 0) { st->holes_due = 0; st->nfail_gen = 0u; }
            }
            long long ua0 = clock64();
            phase_UA(p, rb, st, cta, NG);
            long long ua1 = clock64();
            group_sync(&st->bar, NG, p.backoff);
            long long ua2 = clock64();
            phase_UB(p, rb, st, cta, NG, s_warp0, p.warp_bytes, s_cnt, tim, s_mbar, mbar_parity);
            long long ua3 = clock64();
            if (timer) { tim[2] += ua1 - ua0; tim[3] += ua2 - ua1; tim[4] += ua3 - ua2; tim[5] -= ua3; }
            if (cta == 0 && tid == 0) st->update_pending = 1;
            group_sync(&st->bar, NG, p.backoff);
            if (timer) { tim[27] += clock64() - tu1; tim[5] += clock64(); }
            if (rb.ctl && !p.clustering) snapshot_share(p, rb, st, cta, NG);   // (the same condition as publish_dump's)
        }
        // Phase D: a regular generation (as many births as deaths, none failed -- the failure count is where phase S1
        // published it) leaves the survivors in order: every chain CTA ranks its share of the live points.  The
        // decision reads only what every CTA of the run sees alike.
        d_done = p.off_dkeys != 0 && B == K && K > 0 && K < n &&
                 (sharded || (unsigned long long)vload(&st->nfail) == __shfl_sync(FULL, pw, 9));
        if (d_done && cta >= c0) {
            const long long td0 = clock64();
            phase_D(p, rb, st, sharded ? p.sh.xrun[xr] + (size_t)xpar * p.batch_K : rb.bkey, cta - c0, NG - c0,
                    (double*)(smem + p.off_dkeys));
            if (ctimer) tim[22] += clock64() - td0;
        }
        if (will_chain && !prep_early) prep_next();
    }
}

// ---------------------------------------------------------------- probes
// SliceSampling for explicit (seed point, contour, uid) triples; one warp per chain.
template <int G, int DPL, int KIND>
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
    Model<G, DPL, KIND> M;
    M.init(p.cp, s_like, p.prior_params, cs.dvec);
    for (int c = gw; c < nchains; c += gridDim.x * W) {
        double x[DPL];
#pragma unroll
        for (int j = 0; j < DPL; ++j) x[j] = M.valid(j) ? seed_points[(size_t)c * T + M.dim(j)] : 0.0;
        unsigned long long nl = 0;
        double* out = babies + (size_t)c * R * T;
        prep_chain<G * DPL>(D, R, LD, seed, uid[c], cs, &p.cp);
        whiten_chain<(G * DPL + 7) / 8>(D, R, LD, s_chol, cs);
        slice_chain<G, DPL, KIND>(p.cp, M, seed, uid[c], x, logL[c], cs, out, out + (size_t)(R - 1) * T, nl);
        if (lane == 0) nlike_out[c] = (long long)nl;
    }
}

// The same probe through the dense chain phase (pc_dense.cuh): a warp takes 32/G chains at a time, one per point group.
// records: nchains x R x SLB doubles of global scratch for the slice records.
template <int G, int DPL, int KIND>
__global__ void __launch_bounds__(256, 2) pc_slice_chains_dense_kernel(const __grid_constant__ KParams p, int nchains,
                                                                       const double* seed_points, const double* chol,
                                                                       const double* logL, const unsigned long long* uid,
                                                                       unsigned seed, double* babies, long long* nlike_out,
                                                                       double* records) {
    if constexpr (KIND != LIKE_CORR) {
        extern __shared__ __align__(16) unsigned char smem[];
        constexpr int NPT = 32 / G, GD = G * DPL, SLB = dense_slb(GD);
        const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = blockDim.x >> 5;
        const int D = p.cp.D, T = p.cp.T, R = p.cp.R, LD = p.cp.LD;
        double* s_chol = (double*)smem;
        double* s_like = (double*)(smem + p.off_like);
        unsigned char* s_warp = smem + p.off_warp + (size_t)warp * p.warp_bytes;
        const int nlp = (p.cp.like_kind == LIKE_GAUSSIAN) ? 2 * D : 0;
        for (int e = tid; e < nlp; e += blockDim.x) s_like[e] = p.like_params[e];
        for (int e = tid; e < D * D; e += blockDim.x) s_chol[e] = chol[e];
        __syncthreads();
        const int gw = blockIdx.x * W + warp, grp = lane / G, sub = lane % G;
        const ChainScratch cs = chain_scratch(s_warp, D, R, LD, true, p.cp.like_kind, NPT, nullptr, SLB);
        Model<G, DPL, KIND> M;
        M.init(p.cp, s_like, p.prior_params, cs.dvec);
        __shared__ double s_tab[4 * G * DPL];
        dense_table_fill<GD>(s_tab, D, p.cp.like_kind, s_like, p.prior_params);
        __syncthreads();
        for (int base = gw * NPT; base < nchains; base += gridDim.x * W * NPT) {
            for (int g = 0; g < NPT && base + g < nchains; ++g) {
                prep_chain<GD>(D, R, LD, seed, uid[base + g], cs, &p.cp);
                whiten_chain<(G * DPL + 7) / 8>(D, R, LD, s_chol, cs);
                dense_store<GD>(R, LD, seed, uid[base + g], cs, records + (size_t)(base + g) * R * SLB);
            }
            __syncwarp();
            const int c = base + grp, cc = min(c, nchains - 1);
            const bool active = c < nchains;
            double x[DPL];
#pragma unroll
            for (int j = 0; j < DPL; ++j) x[j] = M.valid(j) ? seed_points[(size_t)cc * T + M.dim(j)] : 0.0;
            unsigned long long nl = 0;
            double lfin = 0.0;
            double* out = babies + (size_t)cc * R * T;
            slice_chains_dense<G, DPL, KIND>(p.cp, M, seed, uid[cc], active, x, logL[cc], records + (size_t)cc * R * SLB,
                                             cs.stage + (size_t)grp * 2 * SLB, out, out + (size_t)(R - 1) * T, nl, lfin, s_tab);
            if (active) {   // the chain probe returns every baby's derived parameters, as SliceSampling does
                if (p.cp.P > 0)
                    for (int i = sub; i < R - 1; i += G) M.finish_derived(out + (size_t)i * T, false);
                if (sub == 0) nlike_out[c] = (long long)nl;
            }
            __syncwarp();
        }
    }
}

template <int G, int DPL, int KIND>
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
    Model<G, DPL, KIND> M;
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
