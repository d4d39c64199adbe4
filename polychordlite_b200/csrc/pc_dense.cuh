// pc_dense.cuh -- the dense form of the chain phase: one chain per POINT GROUP, 32/G chains per warp.
//
// pc_chain.cuh spends a warp's 32/G point groups on ONE chain (speculative bracket and shrink rounds): the right
// trade when a run is alone on the device and 250 chains face 9472 warp slots.  When the device is filled with
// independent runs (pc_run_ensemble) the currency is issued instructions, and speculation wastes two thirds of the
// evaluated points.  Here every point group runs its OWN chain through the sequential state machine of slice_sample
// (chordal_sampling.f90:163-273: right end, left end, step out right, step out left, shrink), one likelihood
// evaluation per round, all groups of the warp in lock step: every evaluation is one the algorithm needs.  Same
// random numbers, same arithmetic, same decisions as pc_chain.cuh -- the chain-by-chain parity tests run both.
//
// Directions.  A chain's R whitened directions (with their widths and slice uniforms) do not fit 32/G times into a
// warp's share of shared memory, so after prep_chain + whiten_chain (shared memory, one chain at a time) they are
// written to a global-memory block in SLICE order -- slice i of a chain is one contiguous record
// [nhat (G*DPL, zero padded) | w | u0 .. u8] -- and the slice loop streams them back: while slice i runs, the record
// of slice i+1 arrives in the group's other staging buffer by cp.async (L2 -> shared memory, no registers held).
#pragma once
#include "pc_chain.cuh"

namespace pc {

// doubles per slice record (a multiple of two: records move as 16-byte chunks)
__host__ __device__ constexpr int dense_slb(int GD) { return (GD + 1 + NU + 1) & ~1; }

// phases of slice_sample's state machine (shared with the host-callback path, pc_hostchain.cuh)
enum SlicePhase : int { SP_R0 = 0, SP_L0 = 1, SP_OUT_R = 2, SP_OUT_L = 3, SP_SHRINK = 4, SP_DONE = 5 };

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Per-dimension constants of the prior and the likelihood for the slice loop, one CTA-wide table in shared memory:
// [lo | wid | mu | isig], GD entries each, zero beyond D (Model::init's padding rule).  The loop needs them once per
// slice (chord form) and once per baby (theta); held in registers they pushed the loop over the 128 registers of two
// CTAs per SM, and the spills went to local memory behind an L1 that the shared-memory carve-out leaves almost empty
// (profiles/r02a: 10 % of the ensemble's warp samples sat on the reload in front of the baby's store).
template <int GD>
__device__ inline void dense_table_fill(double* tab, int D, int like_kind, const double* s_like, const double* prior_params) {
    for (int e = threadIdx.x; e < GD; e += blockDim.x) {
        const bool v = e < D;
        tab[e] = v ? prior_params[e] : 0.0;
        tab[GD + e] = v ? prior_params[D + e] : 0.0;
        tab[2 * GD + e] = (v && like_kind != LIKE_RASTRIGIN) ? s_like[e] : 0.0;
        tab[3 * GD + e] = (v && like_kind == LIKE_GAUSSIAN) ? s_like[D + e] : 0.0;
    }
}

// The prepared chain in cs (whitened unit directions, widths, deck) -> its global block, in slice order; the slice
// uniforms are drawn straight into the records (prep_chain was called with cs.uni == nullptr).  One warp.
template <int GD>
__device__ inline void dense_store(int R, int LD, unsigned seed, unsigned long long uid, const ChainScratch& cs,
                                   double* gblock) {
    constexpr int SLB = dense_slb(GD), HEAD = GD + 1;
    const int lane = threadIdx.x & 31;
    for (int e = lane; e < R * HEAD; e += 32) {
        const int i = e / HEAD, o = e - i * HEAD;
        const int c = cs.deck[i];
        gblock[i * SLB + o] = o < GD ? cs.nh[c * LD + o] : cs.wts[c];
    }
    for (int e = lane; e < R * NU; e += 32) {
        const int i = e / NU, s = e - i * NU;
        gblock[i * SLB + HEAD + s] = uniform(seed, TAG_SLICE, uid, (unsigned)i, (unsigned)s);
    }
    if ((HEAD + NU) < SLB)
        for (int i = lane; i < R; i += 32) gblock[i * SLB + HEAD + NU] = 0.0;
}

// R slice steps for the chain of every point group of the warp.  Per group (the lanes of a group hold the same
// values): uid, active (the group has a chain), x (start point, cube coordinates, lane `sub` owns dimensions sub,
// sub+G, ...), gblock (the chain's slice records), stage (2 * SLB doubles of shared memory), ph_base (babies 0..R-2),
// last_dst (baby R-1).  nlike counts the evaluations with logL > logzero (calculate.f90:44) of this lane's group on
// its lane 0; lfin receives the last baby's logL.
template <int G, int DPL, int KIND>
__device__ inline void slice_chains_dense(const ChainParams& p, const Model<G, DPL, KIND>& M, unsigned seed,
                                          unsigned long long uid, bool active, double (&x)[DPL], double Lstar,
                                          const double* gblock, double* stage, double* ph_base, double* last_dst,
                                          unsigned long long& nlike, double& lfin, const double* tab) {
    static_assert(KIND != LIKE_CORR, "the dense chain phase has no correlated-Gaussian form (one matrix-vector product per slice and group)");
    constexpr int GD = G * DPL, SLB = dense_slb(GD);
    constexpr int LOG2G = (G == 1) ? 0 : (G == 2) ? 1 : (G == 4) ? 2 : (G == 8) ? 3 : (G == 16) ? 4 : 5;
    const int R = p.R, T = p.T, D = p.D;
    const int lane = threadIdx.x & 31, grp = lane >> LOG2G, sub = lane & (G - 1);
    const unsigned gmask = Model<G, DPL, KIND>::GMASK << (grp << LOG2G);
    const double logzero = p.logzero;
    const double* t_lo = tab + sub;   // this lane's entries: lo at [k*G], wid at [GD + k*G], mu at [2GD + k*G], 1/sigma at [3GD + k*G]

    int slice = 0, phase = active ? SP_R0 : SP_DONE, istep = 0, s_done = 0;
    double w = 0.0, dL = 0.0, dR = 0.0, a = 0.0, b = 0.0, wd = 0.0, t = 0.0, lR = 0.0, lL = 0.0;
    double nh[DPL], nW[DPL], xs[DPL];
#pragma unroll
    for (int k = 0; k < DPL; ++k) nh[k] = nW[k] = xs[k] = 0.0;
    double lcur = logzero;

    auto fetch = [&](int i) {   // record of slice i -> staging buffer i & 1, 16-byte chunks dealt over the group's lanes
        const double* src = gblock + (size_t)i * SLB;
        double* dst = stage + (i & 1) * SLB;
        for (int c = sub; c < SLB / 2; c += G) cp_async16(dst + 2 * c, src + 2 * c);
    };
    auto start_slice = [&]() {  // chordal_sampling.f90:75-85, :213-215
        const double* sb = stage + (slice & 1) * SLB;
#pragma unroll
        for (int k = 0; k < DPL; ++k) nh[k] = sb[sub + k * G];
        w = sb[GD];
        const double u0 = sb[GD + 1];
        dL = u0 * w;
        dR = (1.0 - u0) * w;
        if constexpr (KIND == LIKE_GAUSSIAN) {
#pragma unroll
            for (int k = 0; k < DPL; ++k) {  // z = (theta - mu)/sigma = t*nW + xs
                const double wk = t_lo[GD + k * G], ik = t_lo[3 * GD + k * G];
                nW[k] = nh[k] * wk * ik;
                xs[k] = fma(x[k], wk, t_lo[k * G] - t_lo[2 * GD + k * G]) * ik;
            }
        } else {
#pragma unroll
            for (int k = 0; k < DPL; ++k) {  // theta - mu = t*nW + xs
                const double wk = t_lo[GD + k * G];
                nW[k] = nh[k] * wk;
                xs[k] = fma(x[k], wk, t_lo[k * G] - t_lo[2 * GD + k * G]);
            }
        }
        phase = SP_R0;
        t = dR;
        istep = 0;
        s_done = 0;
    };

    if (active) fetch(0);
    cp_async_commit();
    cp_async_wait_all();
    __syncwarp();
    if (active) {
        start_slice();
        if (R > 1) fetch(1);
    }
    cp_async_commit();

    // rounds of the loop: a chain needs about five per slice; the cap only keeps a corrupted state from hanging the
    // persistent kernel (and with it the device)
    for (long long rounds = 0, cap = 16384LL * R;; ++rounds) {
        if (__all_sync(FULL, phase == SP_DONE) || rounds > cap) break;
        // calculate_point (calculate.f90:6-50) at x + t*nhat
        double y[DPL];
        bool ok = true;
#pragma unroll
        for (int k = 0; k < DPL; ++k) {
            y[k] = fma(t, nh[k], x[k]);
            ok = ok && (y[k] >= 0.0) && (y[k] <= 1.0);
        }
        const unsigned bal = __ballot_sync(FULL, ok);
        double l;
        if constexpr (KIND == LIKE_GAUSSIAN) {  // gaussian.f90:12-41
            double z[DPL];
#pragma unroll
            for (int k = 0; k < DPL; ++k) z[k] = fma(t, nW[k], xs[k]);
            double a0 = z[0] * z[0], a1 = 0.0;
            if (DPL > 1) a1 = z[1] * z[1];
#pragma unroll
            for (int k = 2; k < DPL; ++k) { if (k & 1) a1 = fma(z[k], z[k], a1); else a0 = fma(z[k], z[k], a0); }
            l = fma(M.group_sum(a0 + a1), -0.5, -M.gauss_norm);
        } else {  // rastrigin.f90:20-35 (mu = 0)
            const double TwoPi = 6.283185307179586476925286766559;
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < DPL; ++k) {
                const double th = fma(t, nW[k], xs[k]);
                acc += M.valid(k) ? M.log_rast + th * th - 10.0 * cos(TwoPi * th) : 0.0;
            }
            l = -M.group_sum(acc);
        }
        const bool incube = (bal & gmask) == gmask;
        if (!incube) l = logzero;  // outside the cube the likelihood is not called
        const bool run = phase != SP_DONE;
        if (run && sub == 0 && l > logzero) ++nlike;
        const bool inside = l >= Lstar && l > logzero;
        // slice_sample's state machine (:213-266), one step, written with selects: the groups of a warp sit in
        // different phases, and divergent branches would make every round pay for all of them.
        //   R0 -> L0 -> [step out right]* -> [step out left]* -> shrink ... -> accepted / given up
        const bool in_R0 = phase == SP_R0, in_L0 = phase == SP_L0, in_OR = phase == SP_OUT_R, in_OL = phase == SP_OUT_L,
                   in_SH = phase == SP_SHRINK;
        lR = in_R0 ? l : lR;
        lL = in_L0 ? l : lL;
        const bool R_in = lR >= Lstar && lR > logzero, L_in = lL >= Lstar && lL > logzero;
        // the right side keeps stepping out: entered from L0 when the right end was inside, continued while inside (:223-227)
        const bool stepR = (in_L0 && R_in) || (in_OR && inside);
        // the left side: entered when the right side is closed and the left end was inside, continued while inside (:232-236)
        const bool stepL = !stepR && (((in_L0 || in_OR) && L_in) || (in_OL && inside));
        const bool to_shrink = (in_L0 || in_OR || in_OL) && !stepR && !stepL;
        const bool rejected = in_SH && !inside;
        const bool fin_acc = in_SH && inside;
        const int istep_n = ((in_OR && stepR) || (in_OL && stepL)) ? istep + 1 : 1;
        const double wi = w * (double)istep_n;
        if (stepR) { dR = wi; istep = istep_n; }
        if (stepL) { dL = wi; istep = istep_n; }
        // a rejected draw becomes the bound on its side: the sign of (baby - x0).nhat picks it (:254-262)
        const bool pos = __double2hiint(t) > 0;
        const double wpos = t - a, wneg = b - t;
        if (rejected) {
            wd = pos ? wpos : wneg;
            a = pos ? a : t;
            b = pos ? t : b;
            s_done += 1;
        }
        if (to_shrink) { a = -dL; b = dR; wd = dR + dL; s_done = 0; }
        const bool gave_up = rejected && s_done >= 101;   // "Non deterministic loglikelihood" (:268-271): kept with logL = logzero
        const bool draw = to_shrink || (rejected && !gave_up);
        double u = 0.0;
        if (draw) {  // baby = x0 + (u*(x0Rd + x0Ld) - x0Ld)*nhat (:247)
            const int sidx = 1 + s_done;
            u = sidx < NU ? stage[(slice & 1) * SLB + GD + 1 + sidx] : slow_uniform(seed, uid, (unsigned)slice, (unsigned)sidx);
        }
        t = in_R0 ? -dL : (stepR ? dR : (stepL ? -dL : (draw ? fma(u, wd, a) : t)));
        phase = !run ? SP_DONE : in_R0 ? SP_L0 : stepR ? SP_OUT_R : stepL ? SP_OUT_L : SP_SHRINK;
        const bool fin = run && (fin_acc || gave_up);
        const double lnew = fin_acc ? l : logzero;
        if (__ballot_sync(FULL, fin)) {
            cp_async_wait_all();   // the next slice's record (issued at least three rounds ago)
            __syncwarp();
            if (fin) {
                // y holds the baby's cube coordinates; record [cube | theta | phi | birth | logL] (settings.f90:163-182)
                double* dst = (slice == R - 1) ? last_dst : ph_base + (size_t)slice * T;
#pragma unroll
                for (int k = 0; k < DPL; ++k)
                    if (M.valid(k)) {
                        dst[M.dim(k)] = y[k];
                        if (slice == R - 1 || !p.ph_narrow) dst[D + M.dim(k)] = incube ? fma(t_lo[GD + k * G], y[k], t_lo[k * G]) : 0.0;
                    }
                if (sub == 0) {
                    dst[2 * D + p.P] = Lstar;
                    dst[2 * D + p.P + 1] = lnew;
                }
#pragma unroll
                for (int k = 0; k < DPL; ++k) x[k] = y[k];  // next start = this baby even if it failed (:88)
                lcur = lnew;
                slice += 1;
                if (slice == R) phase = SP_DONE;
                else {
                    start_slice();
                    if (slice + 1 < R) fetch(slice + 1);
                }
            }
            cp_async_commit();
        }
    }
    __syncwarp();
    if (p.P > 0 && active && sub == 0) M.finish_derived(last_dst, false);  // only the baby that becomes a live point needs phi
    __syncwarp();
    lfin = lcur;
}

}  // namespace pc
