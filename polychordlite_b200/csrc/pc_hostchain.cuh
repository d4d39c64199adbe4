// pc_hostchain.cuh -- slice-sampling chains whose likelihood (and prior) are HOST callbacks
// (arbitrary Python / C callables behind polychord_c_interface).
//
// The ABI only carries host function pointers, and the reference calls them on the calling thread
// (interfaces.F90:438-474, GIL held): the engine therefore runs the K chains of a generation in LOCK STEP.  The slice
// state machine of every chain (chordal_sampling.f90:163-273: bracket ends, step-out, shrink) lives on the device;
// one round = one kernel launch that consumes the previous trial's log-likelihood and emits every chain's next trial
// point into mapped host memory, after which the host thread makes one prior + likelihood call per active chain.
// Directions, whitening, seeds, the phantom pool and the live/dead bookkeeping are the same device code the
// built-in likelihoods use, so a host-callback run makes exactly the evaluations (nlike) the sequential algorithm
// makes and consumes the same random numbers.
#pragma once
#include "pc_kernels.cuh"

namespace pc {

enum HcPhase : int { HC_R0 = 0, HC_L0 = 1, HC_OUT_R = 2, HC_OUT_L = 3, HC_SHRINK = 4, HC_DONE = 5 };

// per-chain state between rounds (device global memory)
struct HcChain {
    double w, dL, dR, a, b, wd, t, lR, lL;
    unsigned long long uid;
    int slice, phase, istep, s_done, dslot, pad;
};

// what the host sees of a chain's pending trial: [cube D | incube flag | active flag]
// what the host returns: [theta D | phi P | logL]
struct HcParams {
    int D, P, T, R, LD, n;
    ChainParams cp;        // for the direction preparation (grades)
    int K;                 // chains (births) of this generation
    int Kdead;             // deaths of this generation (the live count moves when they differ)
    double logzero;
    unsigned seed;
    RunBuf rb;
    unsigned char* scratch;   // K x scratch_bytes (ChainScratch areas, global memory)
    size_t scratch_bytes;
    double* x;                // K x LD current points (cube coordinates)
    HcChain* ch;              // K
    double* out;              // mapped host memory: K x (D + 2)
    const double* in;         // mapped host memory: K x (D + P + 1)
};

__device__ __forceinline__ ChainScratch hc_scratch(const HcParams& p, int k) {
    return chain_scratch(p.scratch + (size_t)k * p.scratch_bytes, p.D, p.R, p.LD, true, LIKE_GAUSSIAN, 1, nullptr);
}

// emit the trial point x + t*nhat of chain k (one warp)
__device__ inline void hc_emit(const HcParams& p, int k, const ChainScratch& cs, const HcChain& c) {
    const int lane = threadIdx.x & 31, D = p.D;
    const double* q = cs.nh + (size_t)cs.deck[c.slice] * p.LD;
    const double* x = p.x + (size_t)k * p.LD;
    double* o = p.out + (size_t)k * (D + 2);
    bool ok = true;
    for (int r = lane; r < D; r += 32) {
        const double y = fma(c.t, q[r], x[r]);
        o[r] = y;
        ok = ok && (y >= 0.0) && (y <= 1.0);
    }
    ok = __all_sync(FULL, ok);
    if (lane == 0) { o[D] = ok ? 1.0 : 0.0; o[D + 1] = 1.0; }
}

// first trial of slice c.slice: the right end of the initial bracket (:213-215)
__device__ inline void hc_start_slice(const HcParams& p, const ChainScratch& cs, HcChain& c) {
    const int col = cs.deck[c.slice];
    c.w = cs.wts[col];
    const double u0 = cs.uni[(size_t)c.slice * NU];
    c.dL = u0 * c.w;
    c.dR = (1.0 - u0) * c.w;
    c.phase = HC_R0;
    c.t = c.dR;
    c.istep = 0;
    c.s_done = 0;
}

// Start of a generation: GenerateSeed (generate.F90:19-55), the dying point moves to the dead list
// (run_time_info.f90:789-817), directions + whitening, first trial.  One warp per chain.
__global__ void hc_begin_kernel(const HcParams p) {
    const int lane = threadIdx.x & 31, W = blockDim.x >> 5;
    const int k = blockIdx.x * W + (threadIdx.x >> 5);
    const DevRun* st = p.rb.st;
    const int T = p.T, D = p.D, n = st->n_gen, K = p.Kdead, B = p.K;
    const int* order = p.rb.order + st->order_off;
    if (k >= B) {   // a death without a birth (the live count shrinks): only the move to the dead list
        if (k < K)
            for (int e = lane; e < T; e += 32) p.rb.dead[(size_t)(st->ndead_base + k) * T + e] = p.rb.live[(size_t)order[k] * T + e];
        return;
    }
    const unsigned long long uid = (unsigned long long)(st->nchains_base + k);
    const int m = n - K;
    const double u = uniform(p.seed, TAG_SEED, uid, 0u, 0u);
    int choice = (int)ceil(u * (double)m);
    choice = max(1, min(m, choice));
    const int src = order[K + choice - 1], dslot = k < K ? order[k] : n + (k - K);
    if (k < K)
        for (int e = lane; e < T; e += 32) p.rb.dead[(size_t)(st->ndead_base + k) * T + e] = p.rb.live[(size_t)dslot * T + e];
    double* x = p.x + (size_t)k * p.LD;
    for (int r = lane; r < p.LD; r += 32) x[r] = r < D ? p.rb.live[(size_t)src * T + r] : 0.0;
    const ChainScratch cs = hc_scratch(p, k);
    prep_chain(D, p.R, p.LD, p.seed, uid, cs, &p.cp);
    whiten_chain<16>(D, p.R, p.LD, p.rb.chol, cs);
    __syncwarp();
    HcChain c;
    c.uid = uid; c.slice = 0; c.dslot = dslot; c.lR = c.lL = 0.0; c.a = c.b = c.wd = 0.0; c.pad = 0;
    hc_start_slice(p, cs, c);
    hc_emit(p, k, cs, c);
    if (lane == 0) p.ch[k] = c;
}

// One round: consume the log-likelihood of the pending trial, advance slice_sample's state machine, emit the next
// trial.  One warp per chain; the scalars are replicated over the lanes.
__global__ void hc_step_kernel(const HcParams p) {
    const int lane = threadIdx.x & 31, W = blockDim.x >> 5;
    const int k = blockIdx.x * W + (threadIdx.x >> 5);
    if (k >= p.K) return;
    HcChain c = p.ch[k];
    if (c.phase == HC_DONE) return;
    const DevRun* st = p.rb.st;
    const int T = p.T, D = p.D, P = p.P, R = p.R;
    const double Lstar = st->Lstar, logzero = p.logzero;
    const ChainScratch cs = hc_scratch(p, k);
    const double* in = p.in + (size_t)k * (D + P + 1);
    const double l = in[D + P];
    const bool inside = l >= Lstar && l > logzero;
    bool accepted = false, gave_up = false;
    auto begin_shrink = [&]() {
        c.a = -c.dL; c.b = c.dR; c.wd = c.dR + c.dL;
        c.phase = HC_SHRINK;
        c.s_done = 0;
    };
    auto next_draw = [&]() {  // baby = x0 + (u*(x0Rd + x0Ld) - x0Ld)*nhat (:247)
        const int sidx = 1 + c.s_done;
        const double u = sidx < NU ? cs.uni[(size_t)c.slice * NU + sidx] : uniform(p.seed, TAG_SLICE, c.uid, (unsigned)c.slice, (unsigned)sidx);
        c.t = fma(u, c.wd, c.a);
    };
    switch (c.phase) {
        case HC_R0:
            c.lR = l;
            c.phase = HC_L0;
            c.t = -c.dL;
            break;
        case HC_L0:
            c.lL = l;
            if (c.lR >= Lstar && c.lR > logzero) { c.phase = HC_OUT_R; c.istep = 1; c.dR = c.w * 1.0; c.t = c.dR; }
            else if (inside) { c.phase = HC_OUT_L; c.istep = 1; c.dL = c.w * 1.0; c.t = -c.dL; }
            else { begin_shrink(); next_draw(); }
            break;
        case HC_OUT_R:  // R = x0 + nhat*w*i while inside (:223-227)
            if (inside) { c.istep += 1; c.dR = c.w * (double)c.istep; c.t = c.dR; }
            else if (c.lL >= Lstar && c.lL > logzero) { c.phase = HC_OUT_L; c.istep = 1; c.dL = c.w * 1.0; c.t = -c.dL; }
            else { begin_shrink(); next_draw(); }
            break;
        case HC_OUT_L:  // (:232-236)
            if (inside) { c.istep += 1; c.dL = c.w * (double)c.istep; c.t = -c.dL; }
            else { begin_shrink(); next_draw(); }
            break;
        default:  // HC_SHRINK (:240-266)
            if (inside) accepted = true;
            else {
                const bool pos = c.t > 0.0;   // sign of (baby - x0).nhat picks the bound to move (:254-262)
                const double wpos = c.t - c.a, wneg = c.b - c.t;
                c.wd = pos ? wpos : wneg;
                if (pos) c.b = c.t; else c.a = c.t;
                c.s_done += 1;
                if (c.s_done >= 101) gave_up = true;   // "Non deterministic loglikelihood" (:268-271)
                else next_draw();
            }
            break;
    }
    if (accepted || gave_up) {
        // the pending trial point is the baby: record [cube | theta | phi | birth | logL] (settings.f90:163-182)
        const double* q = cs.nh + (size_t)cs.deck[c.slice] * p.LD;
        double* x = p.x + (size_t)k * p.LD;
        const long long ph0 = st->nph_base + (long long)k * (R - 1);
        double* dst = (c.slice == R - 1) ? p.rb.live + (size_t)c.dslot * T
                                         : p.rb.ph[st->cur_pool] + (size_t)(ph0 + c.slice) * T;
        for (int r = lane; r < D; r += 32) {
            const double y = fma(c.t, q[r], x[r]);
            dst[r] = y;
            dst[D + r] = in[r];
            x[r] = y;                      // next start = this baby even if it failed (:88)
        }
        for (int r = lane; r < P; r += 32) dst[2 * D + r] = in[D + r];
        if (lane == 0) { dst[2 * D + P] = Lstar; dst[2 * D + P + 1] = accepted ? l : logzero; }
        __syncwarp();
        if (c.slice == R - 1 && lane == 0) {   // a failed spawn (nested_sampling.F90:315-319): settle_generation takes it out
            const int failed = !(accepted && l > Lstar);
            p.rb.cfail[k] = failed;
            if (failed) { atomicAdd((unsigned long long*)&p.rb.st->nfail, 1ull); atomicAdd(&p.rb.st->nfail_gen, 1u); }
        }
        c.slice += 1;
        if (c.slice == R) c.phase = HC_DONE;
        else hc_start_slice(p, cs, c);
    }
    if (c.phase != HC_DONE) hc_emit(p, k, cs, c);
    else if (lane == 0) p.out[(size_t)k * (D + 2) + D + 1] = 0.0;  // the host stops calling back for this chain
    if (lane == 0) p.ch[k] = c;
}

// GenerateLivePoints (generate.F90:153-183) for host callbacks: the cube coordinates of attempts a0 .. a0+count-1,
// cube[a][d] = U(TAG_INIT, a, d) -- the numbers the device path draws in init_phase -- for the host to evaluate
__global__ void hc_init_cubes_kernel(unsigned seed, long long a0, int count, int D, double* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count * D) return;
    const int a = i / D, d = i - a * D;
    out[i] = uniform(seed, TAG_INIT, (unsigned long long)(a0 + a), (unsigned)d, 0u);
}

// end of the generation: the evaluation count of the host callbacks joins the run's counters
__global__ void hc_finish_kernel(DevRun* st, long long nlike_add, int resume) {
    st->nlike += nlike_add;
    st->host_resume = resume;
}

}  // namespace pc
