// pc_chain.cuh -- one slice-sampling chain on one warp (the hot path).
//
// Reference: src/polychord/chordal_sampling.f90:7-92 (SliceSampling), :94-145 (generate_nhats),
// :163-273 (slice_sample); random_utils.F90:381-437 (orthonormal bases), :505-532 (shuffle_deck);
// calculate.f90:6-50 (calculate_point); likelihoods/examples/{gaussian,rastrigin,random_gaussian}.f90.
//
// Warp organisation.  The 32 lanes of a warp form NPT = 32/G "point groups" of G lanes.  Every
// group holds the chain's current point x and the slice direction nhat in registers (lane `sub`
// of a group owns dimensions sub, sub+G, ...; DPL = ceil(D/G) values per lane), so one warp
// instruction stream evaluates NPT different trial points x + t_p * nhat at once and a
// log2(G)-stage shuffle reduction finishes each group's likelihood.  The trial points of a slice
// step are known before their likelihoods are: the bracket ends, the step-out multiples and --
// because a rejected shrink candidate moves a bound to its own position -- the whole sequence of
// shrink candidates under the hypothesis "all earlier ones were rejected".  A round therefore
// evaluates [R0, L0, candidate 1..NPT-2] speculatively, a ballot collects the accept bits, and
// the sequential decisions of slice_sample are replayed on the bit mask.  The result (accepted
// point, bounds, and the nlike count of the evaluations the sequential algorithm would have
// made) is identical to the sequential algorithm's; speculative evaluations it would not have
// made are simply not counted.
//
// Everything random is counter-addressed (pc_device.cuh), so the directions, the shuffle deck and
// the slice uniforms of a chain depend only on (seed, uid): prep_chain() can run before the
// chain's generation starts (the run kernel overlaps it with the generation barrier).
#pragma once
#include "pc_device.cuh"

namespace pc {

enum LikeKind : int { LIKE_GAUSSIAN = 0, LIKE_RASTRIGIN = 1, LIKE_CORR = 2 };

constexpr int NU = 9;  // uniforms staged per slice step: u0 and the first 8 shrink draws (two shrink rounds)

constexpr int MAX_GRADES = 8;

// What the chain code needs from the run configuration.
struct ChainParams {
    int D, P, T, R, LD;
    int like_kind;
    // fast/slow parameter grades (settings%grade_dims; RTI%num_repeats per grade, generate.F90:303-309).
    // ngrade <= 1: one grade of all D dimensions with R repeats.
    int ngrade;
    int gdims[MAX_GRADES], greps[MAX_GRADES];
    // Phantoms (babies 0..R-2) only ever contribute their cube coordinates, birth contour and logL; their theta is read by
    // nothing but boost_posterior's harvest and the resume writer.  1: the chains do not store a phantom's theta and phase U
    // does not carry it along (the columns stay in the record, unwritten): 45 % less traffic on both streams.  Set for
    // ensembles (the dense chain phase), whose phantom pools live in DRAM.
    int ph_narrow;
    double logzero;
    double gauss_norm, Vn, log_rast, corr_const;
};

// Per-warp scratch (shared memory; nh may live in global memory when R*D is large).
struct ChainScratch {
    double* nh;    // R columns, leading dimension LD (odd): raw orthonormal basis, then whitened unit directions
    double* wts;   // R: initial bracket width w = 3*|L q| of each column
    double* uni;   // R x NU slice uniforms (null in the dense layout: they go straight into the slice records)
    double* dots;  // 4 x Dpad Gram-Schmidt projections; before that, the tail queue of the Gaussian deviates
    double* dvec;  // NPT x Dpad (correlated Gaussian only)
    double* stage; // dense layout only: NPT x 2 x slb doubles, the point groups' slice-record buffers (pc_dense.cuh)
    int* deck;     // R: column used by slice i
    int* jd;       // R: Fisher-Yates picks
    bool nh_smem;  // nh lives in shared memory (else in global memory: bases are orthogonalised in a staging area)
};

// Gram-Schmidt projections (4 x Dpad) share their area with the tail queue of the Gaussian deviates (64 arguments + 64
// element indices, prep_chain); the block form of Gram-Schmidt (more than 32 dimensions) keeps the 8-column panel's
// projections on all earlier vectors there: 8 x Dpad
__host__ __device__ constexpr int dots_doubles(int Dpad) { return Dpad > 32 ? 8 * Dpad : (4 * Dpad > 96 ? 4 * Dpad : 96); }

// slb > 0 selects the dense layout (pc_dense.cuh): slb doubles per slice record, no uniforms in shared memory
__host__ __device__ inline size_t chain_scratch_bytes(int D, int R, int LD, bool nh_in_smem, int like_kind, int npt, int slb = 0) {
    const int Dpad = (D + 1) & ~1;
    size_t b = 0;
    if (slb > 0) b += (size_t)npt * 2 * slb * 8;  // stage (first: 16-byte aligned for cp.async)
    if (nh_in_smem) b += (size_t)R * LD * 8;
    b += (size_t)R * 8;                 // wts
    if (slb == 0) b += (size_t)R * NU * 8;            // uni
    b += (size_t)dots_doubles(Dpad) * 8;  // dots / tail queue
    if (like_kind == LIKE_CORR) b += (size_t)npt * Dpad * 8;
    b += (size_t)((2 * R + 1) & ~1) * 4;  // deck + jd
    return (b + 15) & ~(size_t)15;
}

__device__ __forceinline__ ChainScratch chain_scratch(unsigned char* base, int D, int R, int LD, bool nh_in_smem,
                                                      int like_kind, int npt, double* nh_global, int slb = 0) {
    ChainScratch cs;
    const int Dpad = (D + 1) & ~1;
    double* d = (double*)base;
    cs.stage = d;
    if (slb > 0) d += (size_t)npt * 2 * slb;
    if (nh_in_smem) { cs.nh = d; d += (size_t)R * LD; } else cs.nh = nh_global;
    cs.wts = d; d += R;
    cs.uni = nullptr;
    if (slb == 0) { cs.uni = d; d += (size_t)R * NU; }
    cs.dots = d; d += dots_doubles(Dpad);
    cs.dvec = d;
    if (like_kind == LIKE_CORR) d += (size_t)npt * Dpad;
    cs.deck = (int*)d;
    cs.jd = cs.deck + R;
    cs.nh_smem = nh_in_smem;
    return cs;
}

// ------------------------------------------------------------------------------------------
// Model: prior + likelihood as seen by one lane of a point group.
// Padding lanes (dimension index >= D) carry x = nhat = 0, lo = wid = mu = isig = 0, so they sit inside
// the cube and add exact zeros to the Gaussian sums: the hot path needs no validity predicates.
// ------------------------------------------------------------------------------------------
template <int G, int DPL, int KIND>
struct Model {
    static constexpr int NPT = 32 / G;
    static constexpr unsigned GMASK = (G == 32) ? 0xffffffffu : ((1u << G) - 1u);
    double mu[DPL], isig[DPL], lo[DPL], wid[DPL];
    static constexpr int kind = KIND;
    int D, P, lane, grp, sub, Dpad;
    double logzero, gauss_norm, Vn, log_rast, corr_const;
    const double* s_mu;    // shared memory: mu[D] (Gaussian kinds)
    const double* invcov;  // shared memory, D x D column-major (corr only)
    double* dvec;          // this group's D-vector scratch (corr only)

    __device__ __forceinline__ int dim(int k) const { return sub + k * G; }
    __device__ __forceinline__ bool valid(int k) const { return sub + k * G < D; }

    __device__ void init(const ChainParams& p, const double* s_like, const double* prior_params, double* warp_dvec) {
        D = p.D; P = p.P; lane = threadIdx.x & 31; grp = lane / G; sub = lane % G;
        Dpad = (D + 1) & ~1;
        logzero = p.logzero; gauss_norm = p.gauss_norm; Vn = p.Vn; log_rast = p.log_rast; corr_const = p.corr_const;
        s_mu = s_like;
        invcov = s_like + D;
        dvec = warp_dvec + (size_t)grp * Dpad;
#pragma unroll
        for (int k = 0; k < DPL; ++k) {
            const int r = dim(k);
            const bool v = r < D;
            lo[k] = v ? prior_params[r] : 0.0;
            wid[k] = v ? prior_params[D + r] : 0.0;
            mu[k] = 0.0; isig[k] = 0.0;
            if (v && kind != LIKE_RASTRIGIN) mu[k] = s_like[r];
            if (v && kind == LIKE_GAUSSIAN) isig[k] = s_like[D + r];
        }
    }

    // sum over the G lanes of a point group; every lane of the group receives the total
    __device__ __forceinline__ double group_sum(double v) const {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
        return v;
    }

    // calculate_point (calculate.f90:6-50) for this group's trial point y: in-cube test, uniform prior
    // (priors.f90:40-55), log-likelihood.  Warp-collective; each group gets its own result.
    __device__ __forceinline__ double eval(const double (&y)[DPL], double (&theta)[DPL], bool& incube) const {
        bool ok = true;
#pragma unroll
        for (int k = 0; k < DPL; ++k) ok = ok && (y[k] >= 0.0) && (y[k] <= 1.0);
        const unsigned bal = __ballot_sync(FULL, ok);
        incube = ((bal >> (grp * G)) & GMASK) == GMASK;
#pragma unroll
        for (int k = 0; k < DPL; ++k) theta[k] = fma(wid[k], y[k], lo[k]);
        double logL;
        if constexpr (KIND == LIKE_GAUSSIAN) {  // likelihoods/examples/gaussian.f90:12-41
            double a0 = 0.0, a1 = 0.0;
#pragma unroll
            for (int k = 0; k < DPL; ++k) {
                const double z = (theta[k] - mu[k]) * isig[k];
                if (k & 1) a1 = fma(z, z, a1); else a0 = fma(z, z, a0);
            }
            logL = -gauss_norm - group_sum(a0 + a1) / 2.0;
        } else if constexpr (KIND == LIKE_RASTRIGIN) {  // likelihoods/examples/rastrigin.f90:20-35
            const double TwoPi = 6.283185307179586476925286766559;
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < DPL; ++k)
                acc += valid(k) ? log_rast + theta[k] * theta[k] - 10.0 * cos(TwoPi * theta[k]) : 0.0;
            logL = -group_sum(acc);
        } else {  // utils.F90:1028-1048 log_gauss with a dense inverse covariance
            __syncwarp();
#pragma unroll
            for (int k = 0; k < DPL; ++k)
                if (valid(k)) dvec[dim(k)] = theta[k] - mu[k];
            __syncwarp();
            double yk[DPL];
#pragma unroll
            for (int k = 0; k < DPL; ++k) yk[k] = 0.0;
            for (int c = 0; c < D; ++c) {
                const double dc = dvec[c];
                const double* col = invcov + (size_t)c * D;
#pragma unroll
                for (int k = 0; k < DPL; ++k)
                    if (valid(k)) yk[k] += col[dim(k)] * dc;
            }
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < DPL; ++k) acc += valid(k) ? (theta[k] - mu[k]) * yk[k] : 0.0;
            logL = corr_const - group_sum(acc) / 2.0;
        }
        // calculate.f90:36-39: outside the cube logL = logzero and the likelihood is not called (callers
        // that keep such a record zero its theta, see write_record)
        return incube ? logL : logzero;
    }

    // Record [cube | theta | phi | birth | logL] (settings.f90:163-182), written by point group `g`.
    // The derived parameters phi are filled in afterwards by finish_derived.
    __device__ __forceinline__ void write_record(double* rec, int g, const double (&y)[DPL], const double (&theta)[DPL],
                                                 double birth, double logL, bool incube = true) const {
        if (grp == g) {
#pragma unroll
            for (int k = 0; k < DPL; ++k)
                if (valid(k)) {
                    rec[dim(k)] = y[k];
                    rec[D + dim(k)] = incube ? theta[k] : 0.0;
                }
            if (sub == 0) {
                rec[2 * D + P] = birth;
                rec[2 * D + P + 1] = logL;
            }
        }
    }

    // Derived parameters of a finished record, one lane per record.  gaussian.f90:37-40:
    // phi1 = |theta-mu|, phi2 = log(phi1^D * Vn(D)); other likelihoods have none (zeros).  Points outside the
    // cube never reach the likelihood (calculate.f90:36-39): zeros (check_cube is set where such records occur).
    __device__ __forceinline__ void finish_derived(double* rec, bool check_cube) const {
        if (P <= 0) return;
        bool gauss = (kind == LIKE_GAUSSIAN);
        if (gauss && check_cube)
            for (int d = 0; d < D; ++d) {
                const double c = __ldcg(rec + d);
                if (!(c >= 0.0 && c <= 1.0)) gauss = false;
            }
        double r = 0.0;
        if (gauss) {
            double r2 = 0.0;
            for (int d = 0; d < D; ++d) {
                const double dl = __ldcg(rec + D + d) - s_mu[d];
                r2 += dl * dl;
            }
            r = sqrt(r2);
        }
        rec[2 * D] = r;
        if (P >= 2) rec[2 * D + 1] = gauss ? log(pow(r, (double)D) * Vn) : 0.0;
        for (int i = 2; i < P; ++i) rec[2 * D + i] = 0.0;
    }
};

// ------------------------------------------------------------------------------------------
// gram_schmidt_block: Gram-Schmidt of ONE basis (random_utils.F90:381-403) on the FP64 tensor cores, for bases of more
// than 32 dimensions.  q0: column 0 of the basis (column c at q0 + c*LDQ, Dg rows), m <= Dg vectors are wanted.
//
// The columns are taken in panels of eight.  Against the vectors of the earlier panels a panel is orthogonalised with
// two small matrix products, W = Qprev^T V (all projections taken from the raw panel, as the classical form does) and
// V -= Qprev W: m8n8k4 DMMAs whose fragments come straight from the columns; inside the panel the eight vectors follow
// the classical form one after the other with the rows dealt over the lanes and the finished panel vectors in
// registers.  The vector-at-a-time form above costs two shared-memory loads per multiply-add and, with one basis left
// over (R = 5 D: four bases side by side, then one on a quarter of the lanes), most of its lanes: measured 54 % of
// all warp samples of the 50-dimensional BASELINE run (profiles/r02b_summary.md).
// wbuf: 8 * (Dg rounded up to 8) doubles.  MAXMT >= ceil(Dg / 8).
// ------------------------------------------------------------------------------------------
template <int MAXMT>
__device__ __noinline__ void gram_schmidt_block(double* q0, int LDQ, int Dg, int m, double* wbuf) {
    constexpr int MAXRPL = (MAXMT * 8 + 31) / 32;
    const int lane = threadIdx.x & 31, fr = lane >> 2, fk = lane & 3;
    const int MT = (Dg + 7) >> 3, KT = (Dg + 3) >> 2;
    for (int p0 = 0; p0 < m; p0 += 8) {
        const int pw = min(8, m - p0);
        if (p0 > 0) {
            const int NT = p0 >> 3;  // 8-vector tiles of the earlier panels
            {   // W[j][c] = q_j . v_{p0+c}: A = Qprev^T (row = earlier vector, k = dimension), B = the panel (k = dimension, col = vector)
                double a0[MAXMT], a1[MAXMT];
#pragma unroll
                for (int t = 0; t < MAXMT; ++t) a0[t] = a1[t] = 0.0;
                const bool bcol = fr < pw;
                const double* vb = q0 + (size_t)(p0 + (bcol ? fr : 0)) * LDQ;
                for (int kt = 0; kt < KT; ++kt) {
                    const int k = 4 * kt + fk;
                    const bool kin = k < Dg;
                    const double b = (kin && bcol) ? vb[k] : 0.0;
#pragma unroll
                    for (int t = 0; t < MAXMT; ++t)
                        if (t < NT) {
                            const double a = kin ? q0[(size_t)(8 * t + fr) * LDQ + k] : 0.0;
                            dmma884(a0[t], a1[t], a, b);
                        }
                }
#pragma unroll
                for (int t = 0; t < MAXMT; ++t)
                    if (t < NT) {  // this lane holds W[8t + fr][2 fk + {0, 1}]
                        wbuf[(8 * t + fr) * 8 + 2 * fk] = a0[t];
                        wbuf[(8 * t + fr) * 8 + 2 * fk + 1] = a1[t];
                    }
            }
            __syncwarp();
            {   // V -= Qprev W: A = Qprev (row = dimension, k = earlier vector), B = -W, C = the panel (row = dimension, col = vector)
                double c0[MAXMT], c1[MAXMT];
                const int ca = p0 + 2 * fk;
                const bool h0 = 2 * fk < pw, h1 = 2 * fk + 1 < pw;
#pragma unroll
                for (int mt = 0; mt < MAXMT; ++mt) {
                    const int r = 8 * mt + fr;
                    c0[mt] = (mt < MT && r < Dg && h0) ? q0[(size_t)ca * LDQ + r] : 0.0;
                    c1[mt] = (mt < MT && r < Dg && h1) ? q0[(size_t)(ca + 1) * LDQ + r] : 0.0;
                }
                for (int kt = 0; kt < 2 * NT; ++kt) {
                    const int j = 4 * kt + fk;
                    const double b = -wbuf[j * 8 + fr];
                    const double* qj = q0 + (size_t)j * LDQ;
#pragma unroll
                    for (int mt = 0; mt < MAXMT; ++mt)
                        if (mt < MT) {
                            const int r = 8 * mt + fr;
                            const double a = r < Dg ? qj[r] : 0.0;
                            dmma884(c0[mt], c1[mt], a, b);
                        }
                }
#pragma unroll
                for (int mt = 0; mt < MAXMT; ++mt) {
                    const int r = 8 * mt + fr;
                    if (mt < MT && r < Dg) {
                        if (h0) q0[(size_t)ca * LDQ + r] = c0[mt];
                        if (h1) q0[(size_t)(ca + 1) * LDQ + r] = c1[mt];
                    }
                }
            }
            __syncwarp();
        }
        // inside the panel: rows lane, lane + 32, ... of a vector on this lane, the finished vectors of the panel in registers
        double P[8][MAXRPL];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i < pw) {  // warp-uniform
                double* col = q0 + (size_t)(p0 + i) * LDQ;
                double v[MAXRPL];
#pragma unroll
                for (int r = 0; r < MAXRPL; ++r) v[r] = (lane + 32 * r < Dg) ? col[lane + 32 * r] : 0.0;
                double d[8];
#pragma unroll
                for (int j = 0; j < i; ++j) {
                    double sdot = 0.0;
#pragma unroll
                    for (int r = 0; r < MAXRPL; ++r) sdot = fma(P[j][r], v[r], sdot);
                    d[j] = sdot;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                    for (int j = 0; j < i; ++j) d[j] += __shfl_xor_sync(FULL, d[j], o);
                }
#pragma unroll
                for (int j = 0; j < i; ++j) {
#pragma unroll
                    for (int r = 0; r < MAXRPL; ++r) v[r] = fma(-d[j], P[j][r], v[r]);
                }
                double nrm = 0.0;
#pragma unroll
                for (int r = 0; r < MAXRPL; ++r) nrm = fma(v[r], v[r], nrm);
                nrm = warp_sum(nrm);
                const double inv = 1.0 / sqrt(nrm);
#pragma unroll
                for (int r = 0; r < MAXRPL; ++r) {
                    v[r] *= inv;
                    P[i][r] = v[r];
                    if (lane + 32 * r < Dg) col[lane + 32 * r] = v[r];
                }
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// prep_chain: everything of a chain that depends only on (seed, uid).
//   nh   <- ceil(R/D) Haar-random orthonormal bases (Gram-Schmidt on Gaussian vectors,
//           random_utils.F90:381-437), column c at nh + c*LD
//   deck <- Fisher-Yates shuffle of columns 1..R-1 (chordal_sampling.f90:133-136, random_utils.F90:505-532)
//   uni  <- slice uniforms: uni[i*NU + 0] = u0 of slice i, uni[i*NU + 1 + s] = shrink draw s
//           (skipped when cs.uni is null: the dense mode writes them straight into its global slice blocks)
// Preparation is most of the warp instructions of a run (ncu, profiles/r01e: 2.3e4 per chain against 1.5e4 for
// the 40 slice steps), so its two heavy parts are written for instruction count:
//   * the Gaussian deviates evaluate AS241's central rational for every draw and queue the 15 % that fall in the
//     tails; the queue is worked off 32 at a time, so the log / sqrt / second rational are paid per tail draw and not
//     per warp-wide call;
//   * Gram-Schmidt walks columns with pointers (no 64-bit index arithmetic per element) and, when the instantiation
//     knows the padded dimension GD (rows D..GD-1 of every column are zero), runs its dot products fully unrolled.
//     The order of every floating-point sum is the one of the generic loop: both give the same bits.
// ------------------------------------------------------------------------------------------
template <int GD = 0>
__device__ inline void prep_chain(int D, int R, int LD, unsigned seed, unsigned long long uid, const ChainScratch& cs,
                                  const ChainParams* gp = nullptr) {
    const int lane = threadIdx.x & 31;
    const unsigned below = (1u << lane) - 1u;
    double* nh = cs.nh;
    const int ngrade = (gp && gp->ngrade > 1) ? gp->ngrade : 1;
    // (c) shuffle picks and (d) slice uniforms are independent of the directions: issue them first so their
    //     integer work overlaps the FP64 work below
    for (int i = lane; i < R; i += 32) {
        int j = 0;
        if (i >= 1) {
            const double u = uniform(seed, TAG_SHUF, uid, (unsigned)i, 0u);
            j = (int)ceil(u * (double)i);
            if (j < 1) j = 1;
        }
        cs.jd[i] = j;
    }
    // bases of more than 32 dimensions take the block form of Gram-Schmidt (gram_schmidt_block); when the directions
    // live in global memory a basis is drawn and orthogonalised in a staging area of shared memory -- the room of the
    // widths and the slice uniforms, which are only written afterwards -- and copied out once
    constexpr bool HAS_BLOCK = (GD == 0 || GD > 32);
    constexpr int BLOCK_MT = GD > 0 ? (GD + 7) / 8 : 16;
    const bool can_stage = HAS_BLOCK && D > 32 && !cs.nh_smem && cs.uni != nullptr;
    auto draw_uniforms = [&]() {
        for (int e = lane; e < R * NU; e += 32) {
            const int i = e / NU, s = e - i * NU;
            cs.uni[e] = uniform(seed, TAG_SLICE, uid, (unsigned)i, (unsigned)s);
        }
    };
    if (cs.uni && !can_stage) draw_uniforms();
    const int Dpad = (D + 1) & ~1;
    // generate_nhats (chordal_sampling.f90:94-145): grade g contributes Rg columns (from column cbase) drawn from
    // orthonormal bases of the sub-space of the dimensions roff..D-1 (the dimensions of grades >= g); the rows above
    // roff and the padding rows D..LD-1 are zero
    int cbase = 0, roff = 0;
    for (int g = 0; g < ngrade; ++g) {
        const int Dg = D - roff;
        const int Rg = ngrade > 1 ? gp->greps[g] : R;
        // (a) Gaussian deviates, two per Philox block (inv_normal_cdf = AS241, utils.F90:777-966): columns
        //     cfirst .. cfirst + ncols - 1 of this grade, element (column c, row r of the sub-space) at dst[c*ldd + r]
        auto fill_gauss = [&](int cfirst, int ncols, double* dst, int ldd) {
            const int H = (Dg + 1) >> 1, total = ncols * H;
            double* qp = cs.dots;                    // tail queue: up to 64 arguments ...
            int* qi = (int*)(cs.dots + 64);          // ... and the elements of dst they belong to
            int qn = 0;
            for (int e0 = 0; e0 < total; e0 += 32) {
                const int e = e0 + lane;
                const bool act = e < total;
                const int col = act ? e / H : 0, hp = e - col * H;
                double u0 = 0.5, u1 = 0.5;
                if (act) uniform2(seed, TAG_DIR, uid, (unsigned)(cbase + cfirst + col), (unsigned)hp, u0, u1);
                const int slot = col * ldd + 2 * hp;
                const bool has1 = act && (2 * hp + 1 < Dg);
                double z0, z1;
                const bool tl0 = !inv_normal_cdf_central(u0, z0) && act;
                const bool tl1 = !inv_normal_cdf_central(u1, z1) && has1;
                if (act && !tl0) dst[slot] = z0;
                if (has1 && !tl1) dst[slot + 1] = z1;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const bool tl = half ? tl1 : tl0;
                    const unsigned m = __ballot_sync(FULL, tl);
                    if (m == 0u) continue;
                    if (tl) {
                        const int pos = qn + __popc(m & below);
                        qp[pos] = half ? u1 : u0;
                        qi[pos] = slot + half;
                    }
                    qn += __popc(m);
                    __syncwarp();
                    if (qn >= 32) {   // a full warp of tail draws
                        dst[qi[lane]] = inv_normal_cdf_tail(qp[lane]);
                        const int rest = qn - 32;
                        double tp = 0.0;
                        int ti = 0;
                        if (lane < rest) { tp = qp[32 + lane]; ti = qi[32 + lane]; }
                        __syncwarp();
                        if (lane < rest) { qp[lane] = tp; qi[lane] = ti; }
                        qn = rest;
                        __syncwarp();
                    }
                }
            }
            if (lane < qn) dst[qi[lane]] = inv_normal_cdf_tail(qp[lane]);
        };
        const bool block = HAS_BLOCK && Dg > 32;
        const bool staged = block && can_stage && (long long)Dg * Dg <= (long long)R * (1 + NU);
        if (!staged) fill_gauss(0, Rg, nh + (size_t)cbase * LD + roff, LD);
        for (int e = lane; e < Rg * (LD - Dg); e += 32) {  // zeros: rows 0..roff-1 and D..LD-1 of every column
            const int col = e / (LD - Dg), z = e - col * (LD - Dg);
            nh[(cbase + col) * LD + (z < roff ? z : D + z - roff)] = 0.0;
        }
        __syncwarp();
        // (b) Gram-Schmidt, B bases at a time on LB = 32/B lanes each.  Classical form: all projections of
        //     vector i on q_0..q_{i-1} are taken from the raw vector, then subtracted together.
        const int nb = (Rg + Dg - 1) / Dg;
        if constexpr (HAS_BLOCK) {
            if (block) {   // one basis at a time on the whole warp
                for (int basis = 0; basis < nb; ++basis) {
                    const int m = min(Dg, Rg - basis * Dg);
                    double* qcols = nh + (size_t)(cbase + basis * Dg) * LD + roff;   // where the basis belongs
                    if (staged) {
                        double* sq = cs.wts;   // [wts | uni): R * (1 + NU) doubles
                        fill_gauss(basis * Dg, m, sq, Dg);
                        __syncwarp();
                        gram_schmidt_block<BLOCK_MT>(sq, Dg, Dg, m, cs.dots);
                        for (int e = lane; e < m * Dg; e += 32) {
                            const int c = e / Dg, r = e - c * Dg;
                            qcols[(size_t)c * LD + r] = sq[e];
                        }
                        __syncwarp();
                    } else {
                        gram_schmidt_block<BLOCK_MT>(qcols, LD, Dg, m, cs.dots);
                    }
                }
                cbase += Rg;
                roff += ngrade > 1 ? gp->gdims[g] : D;
                continue;
            }
        }
        const int B = nb >= 4 ? 4 : (nb >= 2 ? 2 : 1);
        const int lbs = B >= 4 ? 3 : (B >= 2 ? 4 : 5), LB = 1 << lbs;   // lanes per basis (a power of two: no integer division)
        const int sl = lane & (LB - 1), bslot = lane >> lbs;
        double* dots = cs.dots + bslot * Dpad;
        const bool unrolled = GD > 0 && roff == 0 && D <= GD && LD >= GD;   // padding rows are zero: sum over GD rows
        for (int b0 = 0; b0 < nb; b0 += B) {
            const int basis = b0 + bslot;
            const int col0 = cbase + basis * Dg;
            const int m = (basis < nb) ? min(Dg, Rg - basis * Dg) : 0;   // vectors of my basis that are used
            const int mmax = min(Dg, Rg - b0 * Dg);                      // trip count of the round (first basis is the longest)
            double* const q0 = nh + col0 * LD + roff;                    // column 0 of my basis
            for (int i = 0; i < mmax; ++i) {
                const bool act = i < m;
                double* vp = q0 + i * LD;
                if (act) {
                    for (int jj = sl; jj < i; jj += LB) {
                        const double* q = q0 + jj * LD;
                        double d0 = 0.0, d1 = 0.0;
                        if (unrolled) {
#pragma unroll
                            for (int r = 0; r < (GD > 0 ? GD : 1); ++r) {
                                if (r & 1) d1 += vp[r] * q[r]; else d0 += vp[r] * q[r];
                            }
                        } else {
                            int r = 0;
                            for (; r + 1 < Dg; r += 2) { d0 += vp[r] * q[r]; d1 += vp[r + 1] * q[r + 1]; }
                            if (r < Dg) d0 += vp[r] * q[r];
                        }
                        dots[jj] = d0 + d1;
                    }
                }
                __syncwarp();
                double acc = 0.0;
                if (act) {
                    for (int r = sl; r < Dg; r += LB) {
                        double t0 = vp[r], t1 = 0.0;
                        const double* qc = q0 + r;
                        const double* dj = dots;
                        int jj = 0;
                        for (; jj + 1 < i; jj += 2, qc += 2 * LD, dj += 2) {
                            t0 -= dj[0] * qc[0];
                            t1 -= dj[1] * qc[LD];
                        }
                        if (jj < i) t0 -= dj[0] * qc[0];
                        const double t = t0 + t1;
                        vp[r] = t;
                        acc += t * t;
                    }
                }
                for (int o = LB >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
                if (act) {
                    const double inv = 1.0 / sqrt(acc);
                    for (int r = sl; r < Dg; r += LB) vp[r] *= inv;
                }
                __syncwarp();
            }
        }
        cbase += Rg;
        roff += ngrade > 1 ? gp->gdims[g] : D;
    }
    if (cs.uni && can_stage) draw_uniforms();   // (the staging area is free again)
    // (c) the shuffled deck.  The swaps run i = R-1 .. 1 (swap deck[i], deck[jd[i]]); the element that ends
    //     at position p is found by walking the swaps backwards from p.
    for (int p = lane; p < R; p += 32) {
        int pos = p;
        for (int i = 1; i < R; ++i) {
            const int j = cs.jd[i];
            pos = (pos == i) ? j : ((pos == j) ? i : pos);
        }
        cs.deck[p] = pos;
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------
// whiten_chain: nhats = matmul(cholesky, nhats) (chordal_sampling.f90:73), w = 3*|nhat|, nhat /= |nhat|
// (:80-82) for all R columns, in place.  chol: D x D column-major LOWER-triangular factor (zeros above).
//
// The product runs on the FP64 tensor cores: for a block of 8 columns the warp accumulates the 8-row tiles of
// L * Q with m8n8k4 DMMAs (tiles of L above the diagonal are skipped), so each lane ends up with two columns'
// entries of rows fr, fr+8, ...; the column norms are finished with three shuffles and the normalised columns
// are written back over the raw ones (every read of a column block precedes its writes).
// MAXMT = 8-row tiles the instantiation can hold (>= ceil(D/8)).
// ------------------------------------------------------------------------------------------
template <int MAXMT>
__device__ inline void whiten_chain(int D, int R, int LD, const double* chol, const ChainScratch& cs) {
    const int lane = threadIdx.x & 31, fr = lane >> 2, fk = lane & 3;
    double* nh = cs.nh;
    const int MT = (D + 7) >> 3, KT = (D + 3) >> 2;
    for (int c0 = 0; c0 < R; c0 += 8) {
        double acc0[MAXMT], acc1[MAXMT];
#pragma unroll
        for (int mt = 0; mt < MAXMT; ++mt) acc0[mt] = acc1[mt] = 0.0;
        const int cb = c0 + fr;  // the column this lane feeds into the B fragments
        const double* qcol = nh + (size_t)min(cb, R - 1) * LD;
        for (int kt = 0; kt < KT; ++kt) {
            const int k = 4 * kt + fk;
            const double bq = (k < D && cb < R) ? qcol[k] : 0.0;
#pragma unroll
            for (int mt = 0; mt < MAXMT; ++mt) {
                if (mt < MT && 8 * mt + 7 >= 4 * kt) {  // warp-uniform: the tile touches the lower triangle
                    const int r = 8 * mt + fr;
                    const double al = (r < D && k <= r) ? chol[r + (size_t)k * D] : 0.0;
                    dmma884(acc0[mt], acc1[mt], al, bq);
                }
            }
        }
        // this lane holds rows 8*mt + fr of columns c0 + 2*fk + {0, 1}
        double n0 = 0.0, n1 = 0.0;
#pragma unroll
        for (int mt = 0; mt < MAXMT; ++mt) { n0 = fma(acc0[mt], acc0[mt], n0); n1 = fma(acc1[mt], acc1[mt], n1); }
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) { n0 += __shfl_xor_sync(FULL, n0, o); n1 += __shfl_xor_sync(FULL, n1, o); }
        const double w0 = sqrt(n0), w1 = sqrt(n1);
        const double i0 = 1.0 / w0, i1 = 1.0 / w1;
        const int ca = c0 + 2 * fk;
        __syncwarp();  // every lane has read its B fragments of this column block
#pragma unroll
        for (int mt = 0; mt < MAXMT; ++mt) {
            const int r = 8 * mt + fr;
            if (mt < MT && r < D) {
                if (ca < R) nh[(size_t)ca * LD + r] = acc0[mt] * i0;
                if (ca + 1 < R) nh[(size_t)(ca + 1) * LD + r] = acc1[mt] * i1;
            }
        }
        if (fr == 0) {
            if (ca < R) cs.wts[ca] = 3.0 * w0;
            if (ca + 1 < R) cs.wts[ca + 1] = 3.0 * w1;
        }
    }
    __syncwarp();
}

// shrink draws beyond the staged ones (rare): kept out of line so the slice loop stays small in the I-cache
static __device__ __noinline__ double slow_uniform(unsigned seed, unsigned long long uid, unsigned i, unsigned sidx) {
    return uniform(seed, TAG_SLICE, uid, i, sidx);
}

// ------------------------------------------------------------------------------------------
// slice_chain: R slice steps from x (chordal_sampling.f90:75-90 calling slice_sample :163-273).
// Babies 0..R-2 go to ph_base + i*T, the last one to last_dst.  Returns the final logL.
//
// A slice step is two speculative rounds in the common case:
//   bracket round   the NPT point groups evaluate the right end R0 and the step-out multiples w, 2w, ...
//                   (groups 0..NS-1) and the left end L0 and -w, -2w, ... (groups NS..2NS-1) at once; the
//                   first point of each side that is outside the contour closes the bracket (:213-236)
//   shrink round    groups 0..NC-1 evaluate the next NC shrink draws, each placed under the hypothesis that
//                   the earlier ones were rejected (a rejected draw becomes the bound on its side, :254-262,
//                   so the positions do not depend on the likelihoods); the first accepted one is the baby
// nlike counts exactly the evaluations the sequential algorithm makes (calculate.f90:44).
//
// The chain is bound by the latency of dependent FP64 operations (~20 cycles each on sm_100), so along a chord
// the likelihood argument is kept in chord form: with theta - mu = t*nW + xs (nW = nhat*width,
// xs = x*width + lo - mu, both formed once per slice) an evaluation at offset t starts with ONE fma per
// dimension; the cube test y = t*nhat + x runs beside it, and theta itself is only formed for the accepted point.
// ------------------------------------------------------------------------------------------
template <int G, int DPL, int KIND>
__device__ inline double slice_chain(const ChainParams& p, const Model<G, DPL, KIND>& M, unsigned seed, unsigned long long uid,
                                     double (&x)[DPL], double Lstar, const ChainScratch& cs, double* ph_base,
                                     double* last_dst, unsigned long long& nlike, long long* tim = nullptr,
                                     bool derive_all = true) {
    constexpr int NPT = 32 / G;
    constexpr int LOG2G = (G == 1) ? 0 : (G == 2) ? 1 : (G == 4) ? 2 : (G == 8) ? 3 : (G == 16) ? 4 : 5;
    constexpr int NS = NPT / 2;                 // bracket points per side and round
    constexpr int SIDE = 16;                    // ballot bits per side (NS * G)
    constexpr int NC = NPT < 4 ? NPT : 4;       // shrink candidates per round
    const int R = p.R, LD = p.LD, T = p.T, D = p.D;
    const int lane = threadIdx.x & 31, grp = lane >> LOG2G;
    const int myc = grp < NC ? grp : NC - 1;    // shrink candidate evaluated by this group
    const int side = NS > 0 ? grp / (NS > 0 ? NS : 1) : 0, idx = grp - side * NS;
    const double fidx = (double)idx;
    const double logzero = p.logzero;
    const unsigned gmask = Model<G, DPL, KIND>::GMASK << (grp << LOG2G);
    double logL_cur = logzero;
    long long tq0 = clock64();

    // ballot bits of groups 0..g-1 (the G lanes of a group always vote alike)
    auto lanes_below = [](int g) -> unsigned { return g > 0 ? (0xffffffffu >> (32 - (g << LOG2G))) : 0u; };

    // per-dimension constants of the chord form
    double wdt[DPL], shf[DPL];  // width, lo - mu
#pragma unroll
    for (int k = 0; k < DPL; ++k) { wdt[k] = M.wid[k]; shf[k] = M.lo[k] - M.mu[k]; }

    // Correlated Gaussian (utils.F90:1028-1048): along a chord the quadratic form is a quadratic in t,
    //   (t nW + xs)^T A (t nW + xs) = c0 + 2 t c1 + t^2 c2,  c2 = nW.v, c1 = xs.v, c0 = xs.u  with v = A nW, u = A xs,
    // so a slice step costs ONE matrix-vector product (v; u follows the accepted point: u += t v) and every trial
    // point three fused multiply-adds.  The product is spread over the 32 lanes (rows lane, lane+32, ...) through
    // the warp's scratch vectors.
    double uA[DPL], vA[DPL];
    double cq0 = 0.0, cq1 = 0.0, cq2 = 0.0;
    double* sv0 = M.dvec - (size_t)grp * M.Dpad;   // the warp's scratch: input vector, then (behind it) the product
    double* sv1 = sv0 + M.Dpad;
    auto matvec = [&](const double (&in)[DPL], double (&out)[DPL]) {
        __syncwarp();
        if (grp == 0) {
#pragma unroll
            for (int k = 0; k < DPL; ++k)
                if (M.valid(k)) sv0[M.dim(k)] = in[k];
        }
        __syncwarp();
        for (int r = lane; r < D; r += 32) {
            double a0 = 0.0, a1 = 0.0;
            int cc = 0;
            for (; cc + 1 < D; cc += 2) {
                a0 = fma(M.invcov[r + (size_t)cc * D], sv0[cc], a0);
                a1 = fma(M.invcov[r + (size_t)(cc + 1) * D], sv0[cc + 1], a1);
            }
            if (cc < D) a0 = fma(M.invcov[r + (size_t)cc * D], sv0[cc], a0);
            sv1[r] = a0 + a1;
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < DPL; ++k) out[k] = M.valid(k) ? sv1[M.dim(k)] : 0.0;
    };
    if constexpr (KIND == LIKE_CORR) {
        double xs0[DPL];
#pragma unroll
        for (int k = 0; k < DPL; ++k) xs0[k] = fma(x[k], wdt[k], shf[k]);
        matvec(xs0, uA);
    }

    for (int i = 0; i < R; ++i) {
        const int c = cs.deck[i];
        const double* q = cs.nh + (size_t)c * LD;
        double nh[DPL], nW[DPL], xs[DPL];
#pragma unroll
        for (int k = 0; k < DPL; ++k) nh[k] = q[M.dim(k)];  // columns are zero-padded to G*DPL entries
        const double w = cs.wts[c];
        const double* ui = cs.uni + (size_t)i * NU;
        const double u0 = ui[0];
        double dL = u0 * w, dR = (1.0 - u0) * w;  // bracket [x - dL*nhat, x + dR*nhat] (:213-215)
        if constexpr (KIND == LIKE_GAUSSIAN) {
#pragma unroll
            for (int k = 0; k < DPL; ++k) {  // z = (theta - mu)/sigma = t*nW + xs
                nW[k] = nh[k] * wdt[k] * M.isig[k];
                xs[k] = fma(x[k], wdt[k], shf[k]) * M.isig[k];
            }
        } else {
#pragma unroll
            for (int k = 0; k < DPL; ++k) {  // theta - mu = t*nW + xs
                nW[k] = nh[k] * wdt[k];
                xs[k] = fma(x[k], wdt[k], shf[k]);
            }
        }

        if constexpr (KIND == LIKE_CORR) {
            matvec(nW, vA);
            double p2 = 0.0, p1 = 0.0, p0 = 0.0;
#pragma unroll
            for (int k = 0; k < DPL; ++k) { p2 = fma(nW[k], vA[k], p2); p1 = fma(xs[k], vA[k], p1); p0 = fma(xs[k], uA[k], p0); }
            cq2 = M.group_sum(p2); cq1 = M.group_sum(p1); cq0 = M.group_sum(p0);
        }

        double y[DPL];
        // log-likelihood of this group's point x + tt*nhat (calculate_point, calculate.f90:6-50)
        auto eval_t = [&](double tt) -> double {
            bool ok = true;
#pragma unroll
            for (int k = 0; k < DPL; ++k) {
                y[k] = fma(tt, nh[k], x[k]);
                ok = ok && (y[k] >= 0.0) && (y[k] <= 1.0);
            }
            const unsigned bal = __ballot_sync(FULL, ok);
            double logL;
            if constexpr (KIND == LIKE_GAUSSIAN) {  // gaussian.f90:12-41
                double z[DPL];
#pragma unroll
                for (int k = 0; k < DPL; ++k) z[k] = fma(tt, nW[k], xs[k]);
                double a0 = z[0] * z[0], a1 = 0.0;
                if (DPL > 1) a1 = z[1] * z[1];
#pragma unroll
                for (int k = 2; k < DPL; ++k) { if (k & 1) a1 = fma(z[k], z[k], a1); else a0 = fma(z[k], z[k], a0); }
                logL = fma(M.group_sum(a0 + a1), -0.5, -M.gauss_norm);
            } else if constexpr (KIND == LIKE_RASTRIGIN) {  // rastrigin.f90:20-35 (mu = 0)
                const double TwoPi = 6.283185307179586476925286766559;
                double acc = 0.0;
#pragma unroll
                for (int k = 0; k < DPL; ++k) {
                    const double th = fma(tt, nW[k], xs[k]);
                    acc += M.valid(k) ? M.log_rast + th * th - 10.0 * cos(TwoPi * th) : 0.0;
                }
                logL = -M.group_sum(acc);
            } else {  // utils.F90:1028-1048 log_gauss with a dense inverse covariance: the chord quadratic
                logL = fma(fma(tt, fma(tt, cq2, 2.0 * cq1), cq0), -0.5, M.corr_const);
            }
            return ((bal & gmask) == gmask) ? logL : logzero;  // outside the cube the likelihood is not called
        };

        double l;
        // ---------------- bracket (:213-236) ----------------
        if (NPT >= 2) {
            const double wm = w * fidx;  // this group's step-out multiple
            {
                const double mult = (idx == 0) ? (side ? dL : dR) : wm;
                l = eval_t(side ? -mult : mult);
            }
            const unsigned in_raw = __ballot_sync(FULL, l >= Lstar && l > logzero);
            const unsigned cnt_raw = __ballot_sync(FULL, l > logzero);
            const unsigned smask = (NS * G >= 32) ? 0xffffffffu : ((1u << (NS * G)) - 1u);
#pragma unroll
            for (int sd = 0; sd < 2; ++sd) {
                const unsigned inb = (in_raw >> (sd * SIDE)) & smask, cnb = (cnt_raw >> (sd * SIDE)) & smask;
                const unsigned out = ~inb & smask;
                if (out) {  // the first point outside the contour closes this side
                    const int g = (__ffs(out) - 1) >> LOG2G;
                    nlike += __popc(cnb & lanes_below(g + 1)) >> LOG2G;
                    if (g > 0) {  // the bound is the multiple group g of this side evaluated
                        const double wg = __shfl_sync(FULL, wm, (sd * NS + g) << LOG2G);
                        if (sd) dL = wg; else dR = wg;
                    }
                } else [[unlikely]] {    // all NS points inside: keep stepping this side, NPT multiples at a time (:223-227)
                    nlike += NS;
                    for (int base = NS;; base += NPT) {
                        const double mult = w * (double)(base + grp);
                        l = eval_t(sd ? -mult : mult);
                        const unsigned ib = __ballot_sync(FULL, l >= Lstar && l > logzero);
                        const unsigned cb = __ballot_sync(FULL, l > logzero);
                        if (~ib) {
                            const int g = (__ffs(~ib) - 1) >> LOG2G;
                            nlike += __popc(cb & lanes_below(g + 1)) >> LOG2G;
                            if (sd) dL = w * (double)(base + g); else dR = w * (double)(base + g);
                            break;
                        }
                        nlike += NPT;
                    }
                }
            }
        } else {  // a single point group: the sequential form
            double lR = eval_t(dR);
            if (lR > logzero) ++nlike;
            double lL = eval_t(-dL);
            if (lL > logzero) ++nlike;
            for (int is = 1; lR >= Lstar && lR > logzero; ++is) {
                dR = w * (double)is;
                lR = eval_t(dR);
                if (lR > logzero) ++nlike;
            }
            for (int is = 1; lL >= Lstar && lL > logzero; ++is) {
                dL = w * (double)is;
                lL = eval_t(-dL);
                if (lL > logzero) ++nlike;
            }
        }
        // ---------------- shrink (:240-266) ----------------
        // bounds as signed positions a < 0 < b and their distance wd = dR + dL, exactly the quantities of
        // baby = x0 + (u*(x0Rd + x0Ld) - x0Ld)*nhat (:247)
        double a = -dL, b = dR, wd = dR + dL;
        double t_acc = 0.0, l_acc = logzero;
        int g_acc = 0, s_done = 0;
        bool accepted = false;
        while (!accepted && s_done < 101) {
            double tmine = 0.0;
#pragma unroll
            for (int cnd = 0; cnd < NC; ++cnd) {
                const int sidx = 1 + s_done + cnd;
                double u;
                if (__builtin_expect(sidx < NU, 1)) u = ui[sidx];
                else u = slow_uniform(seed, uid, (unsigned)i, (unsigned)sidx);
                const double tc = fma(u, wd, a);
                if (cnd == myc) tmine = tc;
                // sign of (baby - x0).nhat picks the bound to move (:254); the sign is read from the high word
                // (tc > 0 for every positive normal double), both new widths are formed beside the test
                const bool pos = __double2hiint(tc) > 0;
                const double wpos = tc - a, wneg = b - tc;
                wd = pos ? wpos : wneg;
                a = pos ? a : tc;
                b = pos ? tc : b;
            }
            l = eval_t(tmine);
            const unsigned in_raw = __ballot_sync(FULL, l >= Lstar && l > logzero);
            const unsigned cnt_raw = __ballot_sync(FULL, l > logzero);
            const int ncand = min(NC, 101 - s_done);  // at most 101 draws per slice (:240)
            const unsigned cmask = lanes_below(ncand);
            const unsigned hit = in_raw & cmask;
            if (hit) {
                g_acc = (__ffs(hit) - 1) >> LOG2G;
                nlike += __popc(cnt_raw & lanes_below(g_acc + 1)) >> LOG2G;
                l_acc = __shfl_sync(FULL, l, g_acc << LOG2G);
                accepted = true;
            } else {
                nlike += __popc(cnt_raw & cmask) >> LOG2G;
                g_acc = ncand - 1;  // the last draw: kept if the slice gives up (:268-271)
                s_done += ncand;
            }
            t_acc = __shfl_sync(FULL, tmine, g_acc << LOG2G);
        }
        // "Non deterministic loglikelihood" (:268-271): after 101 rejected draws the last trial point is kept
        // with logL = logzero.
        const double lnew = accepted ? l_acc : logzero;
        // The accepting group's registers hold the baby's cube coordinates: it writes the record (theta is
        // formed here, priors.f90:40-55).  Every group moves to the same point with the same fma, so the chain
        // state stays replicated bit for bit.
        double* dst = (i == R - 1) ? last_dst : ph_base + (size_t)i * T;
        if (grp == g_acc) {
            bool ok = true;
#pragma unroll
            for (int k = 0; k < DPL; ++k) ok = ok && (y[k] >= 0.0) && (y[k] <= 1.0);
            const bool inc = __all_sync(gmask, ok);
#pragma unroll
            for (int k = 0; k < DPL; ++k)
                if (M.valid(k)) {
                    dst[M.dim(k)] = y[k];
                    dst[D + M.dim(k)] = inc ? fma(wdt[k], y[k], M.lo[k]) : 0.0;
                }
            if (M.sub == 0) {
                dst[2 * D + p.P] = Lstar;
                dst[2 * D + p.P + 1] = lnew;
            }
        }
#pragma unroll
        for (int k = 0; k < DPL; ++k) x[k] = fma(t_acc, nh[k], x[k]);  // next start = this baby even if it failed (:88)
        if constexpr (KIND == LIKE_CORR) {
#pragma unroll
            for (int k = 0; k < DPL; ++k) uA[k] = fma(t_acc, vA[k], uA[k]);
        }
        logL_cur = lnew;
    }
    __syncwarp();
    long long tq1 = clock64();
    // derived parameters, one lane per record.  Inside a run only the last baby needs them: it becomes a live
    // point and later a dead one (what the dumper and the output files report); babies 0..R-2 are phantoms, which
    // only ever contribute their cube coordinates (covariance) and logL (cleaning) -- their phi slots are left
    // unset.  The chain probe (derive_all) fills them for all R babies, as the reference's SliceSampling returns them.
    if (p.P > 0) {
        if (derive_all) {
            for (int i = lane; i < R; i += 32) M.finish_derived((i == R - 1) ? last_dst : ph_base + (size_t)i * T, false);
        } else if (lane == 0) {
            M.finish_derived(last_dst, false);
        }
        __syncwarp();
    }
    if (tim && lane == 0) { tim[0] += tq1 - tq0; tim[1] += clock64() - tq1; }
    return logL_cur;
}

}  // namespace pc
