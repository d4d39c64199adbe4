// pc_oracle.cpp -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
//
// A plain single-threaded C++ restatement of PolyChordLite's linear-mode nested
// sampling algorithm, used ONLY as (a) the parity checker for the CUDA engine in
// tests/, __graft_entry__.smoke() and (b) the timed CPU baseline in bench.py.
// Nothing under polychordlite_b200/ may include, link or call this file.
//
// PARITY STATUS: "parity unpinned by the reference" -- the reference
// (/root/reference, Fortran) cannot be compiled in this image (no gfortran) and its
// own tests hold no numeric golden vectors (tests/test_run_pypolychord.py asserts
// behaviour only).  This oracle is therefore pinned to (i) the analytic evidences the
// built-in likelihoods are normalised to (likelihoods/examples/gaussian.f90:8-9,
// rastrigin.f90:33), (ii) scipy/numpy for the numeric helpers (AS241, Cholesky,
// logsumexp), (iii) a brute-force Monte-Carlo check of the evidence recurrences and
// (iv) the published Philox4x32-10 known-answer vectors.  See tests/test_oracle_*.py.
//
// Reference sites restated here (all paths relative to /root/reference/src/polychord):
//   nested_sampling.F90:140-405   driver, linear branch          -> Run::run()
//   generate.F90:19-55            GenerateSeed                   -> Run::generate_seed()
//   generate.F90:153-183,283-320  GenerateLivePoints (linear)    -> Run::generate_live_points()
//   chordal_sampling.f90:7-92     SliceSampling                  -> Run::slice_sampling()
//   chordal_sampling.f90:94-145   generate_nhats                 -> generate_nhats()
//   chordal_sampling.f90:163-273  slice_sample                   -> Run::slice_sample()
//   calculate.f90:6-50            calculate_point                -> Run::calculate_point()
//   run_time_info.f90:211-296     update_evidence                -> Run::update_evidence()
//   run_time_info.f90:601-641     calculate_covmats              -> Run::calculate_covmats()
//   run_time_info.f90:652-678     calculate_logZ_estimate        -> Run::logZ_estimate()
//   run_time_info.f90:683-709     live_logZ                      -> Run::live_logZ()
//   run_time_info.f90:716-787     replace_point                  -> Run::replace_point()
//   run_time_info.f90:789-817     delete_outermost_point         -> Run::delete_outermost_point()
//   run_time_info.f90:820-877     clean_phantoms                 -> Run::clean_phantoms(), clean_phantoms_stable()
//   run_time_info.f90:846-868       its posterior conversion       -> Run::boost() (boost_posterior; oracle_last_boosted)
//   run_time_info.f90:766-777     dynamic nlive in replace_point -> Run::replace_point(), Run::target_nlive() (oracle_set_nlives)
//   generate.F90:311-316          thin_posterior                 -> Run::thin_posterior()
//   pypolychord/polychord.py:650-789  cube_samples start         -> Run::generate_live_points() (oracle_set_initial_cubes)
//   run_time_info.f90:883-909     find_min_loglikelihoods        -> Run::find_min()
//   run_time_info.f90:913-949     identify_cluster               -> Run::identify_cluster()
//   random_utils.F90:381-437      random_orthonormal_basis/bases -> random_orthonormal_basis()
//   random_utils.F90:505-532      shuffle_deck                   -> shuffle_deck()
//   random_utils.F90:581-614      random_inverse_covmat          -> random_inverse_covmat()
//   utils.F90:362-439             logsumexp/logaddexp/logincexp  -> same names
//   utils.F90:621-649             calc_cholesky                  -> calc_cholesky()
//   utils.F90:777-966             inv_normal_cdf (AS241 PPND16)  -> inv_normal_cdf()
//   utils.F90:1028-1048           log_gauss                      -> loglike_corr_gaussian()
//   array_utils.f90:396-458       add_point / delete_point       -> push_back / swap-with-last
//   likelihoods/examples/{gaussian,rastrigin,random_gaussian}.f90
//   priors.f90:40-55              uniform_htp                    -> Run::prior()
//
// The one thing that cannot be restated is the random stream: the reference draws
// from libgfortran's random_number in program order (random_utils.F90:128).  The
// oracle instead uses the counter-based stream spec shared with the CUDA engine
// (DESIGN.md "RNG stream spec"): Philox4x32-10, key=(seed,tag), counter=(a,b,uid).
//
// Two scheduling modes:
//   batch_K == 0 : REFERENCE MODE.  One death + one birth per iteration, seed drawn
//                  from all live points, swap-with-last deletion: the faithful
//                  restatement of nested_sampling.F90:239-374.  This is the CPU baseline.
//   batch_K >= 1 : BATCHED-GENERATION MODE.  The K lowest points die together (evidence
//                  updated sequentially with n, n-1, ... exactly as the reference's own
//                  final kill-off nested_sampling.F90:381-384 does), K chains are seeded
//                  from the n-K survivors at the contour of the K-th lowest, babies are
//                  written into the vacated slots.  This is the schedule the GPU engine
//                  runs; the oracle in this mode is the bit-for-bit (up to FP
//                  re-association) checker for it.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#include <algorithm>
#include <limits>
#include <numeric>
#include <chrono>

namespace {

// ----------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11).  Written independently of the engine's copy;
// both are pinned to the published known-answer vectors in tests.
// ----------------------------------------------------------------------------------
struct U4 { uint32_t v[4]; };

inline U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c0;
        uint64_t p1 = (uint64_t)M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    return U4{{c0, c1, c2, c3}};
}

enum Tag : uint32_t { TAG_INIT = 1, TAG_SEED = 2, TAG_DIR = 3, TAG_SHUF = 4, TAG_SLICE = 5, TAG_POST = 6, TAG_LIKE = 7, TAG_BOOST = 8 };

inline double u64_to_unit(uint64_t x) {
    // 52 random bits + 1/2 ulp offset: strictly inside (0,1), exactly representable.
    return ((double)(x >> 12) + 0.5) * (1.0 / 4503599627370496.0);
}

struct Rng {
    uint32_t seed;
    // first 64-bit lane of the Philox block
    double uniform(uint32_t tag, uint64_t uid, uint32_t a, uint32_t b) const {
        U4 o = philox4x32_10(a, b, (uint32_t)uid, (uint32_t)(uid >> 32), seed, tag);
        return u64_to_unit((uint64_t)o.v[0] | ((uint64_t)o.v[1] << 32));
    }
    // both 64-bit lanes
    void uniform2(uint32_t tag, uint64_t uid, uint32_t a, uint32_t b, double& u0, double& u1) const {
        U4 o = philox4x32_10(a, b, (uint32_t)uid, (uint32_t)(uid >> 32), seed, tag);
        u0 = u64_to_unit((uint64_t)o.v[0] | ((uint64_t)o.v[1] << 32));
        u1 = u64_to_unit((uint64_t)o.v[2] | ((uint64_t)o.v[3] << 32));
    }
};

// ----------------------------------------------------------------------------------
// numeric helpers (utils.F90)
// ----------------------------------------------------------------------------------
const double LOG_TWO_PI = 1.8378770664093454835606594728112;
const double HUGE_D = std::numeric_limits<double>::max();

// AS241 PPND16 (Wichura 1988); utils.F90:806-966 uses the same published algorithm.
double inv_normal_cdf(double p) {
    static const double a[8] = {3.3871328727963666080, 1.3314166789178437745e+2, 1.9715909503065514427e+3,
                                1.3731693765509461125e+4, 4.5921953931549871457e+4, 6.7265770927008700853e+4,
                                3.3430575583588128105e+4, 2.5090809287301226727e+3};
    static const double b[8] = {1.0, 4.2313330701600911252e+1, 6.8718700749205790830e+2, 5.3941960214247511077e+3,
                                2.1213794301586595867e+4, 3.9307895800092710610e+4, 2.8729085735721942674e+4,
                                5.2264952788528545610e+3};
    static const double c[8] = {1.42343711074968357734, 4.63033784615654529590, 5.76949722146069140550,
                                3.64784832476320460504, 1.27045825245236838258, 2.41780725177450611770e-1,
                                2.27238449892691845833e-2, 7.74545014278341407640e-4};
    static const double d[8] = {1.0, 2.05319162663775882187, 1.67638483018380384940, 6.89767334985100004550e-1,
                                1.48103976427480074590e-1, 1.51986665636164571966e-2, 5.47593808499534494600e-4,
                                1.05075007164441684324e-9};
    static const double e[8] = {6.65790464350110377720, 5.46378491116411436990, 1.78482653991729133580,
                                2.96560571828504891230e-1, 2.65321895265761230930e-2, 1.24266094738807843860e-3,
                                2.71155556874348757815e-5, 2.01033439929228813265e-7};
    static const double f[8] = {1.0, 5.99832206555887937690e-1, 1.36929880922735805310e-1, 1.48753612908506148525e-2,
                                7.86869131145613259100e-4, 1.84631831751005468180e-5, 1.42151175831644588870e-7,
                                2.04426310338993978564e-15};
    auto poly = [](const double* co, double x) {
        double v = 0.0;
        for (int i = 7; i >= 0; --i) v = v * x + co[i];
        return v;
    };
    if (p <= 0.0) return -HUGE_D;
    if (p >= 1.0) return HUGE_D;
    double q = p - 0.5;
    if (std::fabs(q) <= 0.425) {
        double r = 0.180625 - q * q;
        return q * poly(a, r) / poly(b, r);
    }
    double r = (q < 0.0) ? p : 1.0 - p;
    r = std::sqrt(-std::log(r));
    double val;
    if (r <= 5.0) {
        r -= 1.6;
        val = poly(c, r) / poly(d, r);
    } else {
        r -= 5.0;
        val = poly(e, r) / poly(f, r);
    }
    return (q < 0.0) ? -val : val;
}

inline double logaddexp(double la, double lb) {
    if (la > lb) return la + std::log(std::exp(lb - la) + 1.0);
    return lb + std::log(std::exp(la - lb) + 1.0);
}
inline void logincexp(double& la, double lb) { la = logaddexp(la, lb); }
inline void logincexp(double& la, double lb, double lc) {
    la = logaddexp(la, lb);
    la = logaddexp(la, lc);
}
double logsumexp(const double* v, size_t n) {
    double m = v[0];
    for (size_t i = 1; i < n; ++i) m = std::max(m, v[i]);
    double s = 0.0;
    for (size_t i = 0; i < n; ++i) s += std::exp(v[i] - m);
    return m + std::log(s);
}

// utils.F90:621-649.  a, L column-major D x D.  Falls back to sqrt(trace)*I.
void calc_cholesky(const double* a, double* L, int D) {
    std::fill(L, L + (size_t)D * D, 0.0);
    for (int i = 0; i < D; ++i) {
        double s = 0.0;
        for (int k = 0; k < i; ++k) s += L[i + (size_t)k * D] * L[i + (size_t)k * D];
        double dii = a[i + (size_t)i * D] - s;
        if (dii <= 0.0) {
            double tr = 0.0;
            for (int k = 0; k < D; ++k) tr += a[k + (size_t)k * D];
            std::fill(L, L + (size_t)D * D, 0.0);
            for (int k = 0; k < D; ++k) L[k + (size_t)k * D] = std::sqrt(tr);
            return;
        }
        dii = std::sqrt(dii);
        L[i + (size_t)i * D] = dii;
        for (int j = i + 1; j < D; ++j) {
            double t = 0.0;
            for (int k = 0; k < i; ++k) t += L[i + (size_t)k * D] * L[j + (size_t)k * D];
            L[j + (size_t)i * D] = (a[i + (size_t)j * D] - t) / dii;
        }
    }
}

// ----------------------------------------------------------------------------------
// clustering.f90 (KNN_clustering module) restated.  Indices are zero-based.
// ----------------------------------------------------------------------------------
// Squared distance of two cube points.  calculate_similarity_matrix (calculate.f90:94-109) forms
// |vi|^2 + |vj|^2 - 2 vi.vj with matmul; the difference form below is the same number up to rounding and is
// the form the engine's kernel uses, accumulated with fused multiply-adds in dimension order so that both sides
// produce bit-identical distances (the k-nearest-neighbour ORDER must not depend on who computed it).
inline double dist2(const double* a, const double* b, int D) {
    double s = 0.0;
    for (int k = 0; k < D; ++k) { double d = a[k] - b[k]; s = std::fma(d, d, s); }
    return s;
}
std::vector<double> similarity_matrix(const std::vector<const double*>& pts, int D) {
    const int m = (int)pts.size();
    std::vector<double> sim((size_t)m * m);
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) sim[(size_t)i * m + j] = dist2(pts[i], pts[j], D);
    return sim;
}
// compute_knn (clustering.f90:134-174): for each point the k nearest (itself included), by insertion: a new
// distance goes in front of the first stored distance that is strictly larger, so equal distances keep index order.
std::vector<int> compute_knn(const std::vector<double>& sim, int m, int k) {
    std::vector<int> knn((size_t)m * k, -1);
    std::vector<double> d2(k);
    for (int i = 0; i < m; ++i) {
        std::fill(d2.begin(), d2.end(), HUGE_D);
        int* row = &knn[(size_t)i * k];
        for (int j = 0; j < m; ++j) {
            const double v = sim[(size_t)i * m + j];
            int pos = -1;
            for (int t = 0; t < k; ++t) if (d2[t] > v) { pos = t; break; }
            if (pos < 0) continue;
            for (int t = k - 1; t > pos; --t) { d2[t] = d2[t - 1]; row[t] = row[t - 1]; }
            d2[pos] = v; row[pos] = j;
        }
    }
    return knn;
}
// relabel (utils.F90:713-749): labels in order of first appearance; returns their number
int relabel(std::vector<int>& c) {
    std::vector<int> mapping;
    for (int v : c) if (std::find(mapping.begin(), mapping.end(), v) == mapping.end()) mapping.push_back(v);
    for (int& v : c) v = (int)(std::find(mapping.begin(), mapping.end(), v) - mapping.begin());
    return (int)mapping.size();
}
// do_clustering_k (clustering.f90:100-130) on the first n entries of every neighbour list, with
// neighbours (:178-188): i ~ j when the head of one list appears in the other
std::vector<int> do_clustering_k(const std::vector<int>& knn, int m, int k, int n) {
    std::vector<int> c(m);
    std::iota(c.begin(), c.end(), 0);
    auto has = [&](int i, int v) { for (int t = 0; t < n; ++t) if (knn[(size_t)i * k + t] == v) return true; return false; };
    for (int i = 0; i < m; ++i)
        for (int j = i + 1; j < m; ++j)
            if (c[i] != c[j] && (has(i, knn[(size_t)j * k]) || has(j, knn[(size_t)i * k]))) {
                const int ci = c[i], cj = c[j], lo = std::min(ci, cj);
                for (int& v : c) if (v == ci || v == cj) v = lo;
            }
    return c;
}
// NN_clustering (clustering.f90:15-97).  `do n=2,k` fixes its trip count at entry, so the "expand k" branch
// (:64-67) never extends the search: k = min(m, 10) throughout.
std::vector<int> NN_clustering(const std::vector<double>& sim, int m, int& num_clusters) {
    std::vector<int> cl(m, 0);
    num_clusters = 1;
    if (m < 2) return cl;
    const int k = std::min(m, 10);
    const std::vector<int> knn = compute_knn(sim, m, k);
    std::vector<int> old(m);
    std::iota(old.begin(), old.end(), 0);
    for (int n = 2; n <= k; ++n) {
        cl = do_clustering_k(knn, m, k, n);
        num_clusters = relabel(cl);
        if (num_clusters == 1) return cl;
        if (cl == old) break;
        old = cl;
    }
    if (num_clusters > 1) {  // search within the clusters found (:78-95)
        int ic = 0;
        while (ic < num_clusters) {
            std::vector<int> pts;
            for (int j = 0; j < m; ++j) if (cl[j] == ic) pts.push_back(j);
            const int mm = (int)pts.size();
            std::vector<double> sub((size_t)mm * mm);
            for (int a = 0; a < mm; ++a)
                for (int b = 0; b < mm; ++b) sub[(size_t)a * mm + b] = sim[(size_t)pts[a] * m + pts[b]];
            int nnew = 1;
            const std::vector<int> subl = NN_clustering(sub, mm, nnew);
            for (int a = 0; a < mm; ++a) cl[pts[a]] = num_clusters + subl[a];
            if (nnew == 1) ++ic;
            num_clusters = relabel(cl);
        }
    }
    return cl;
}

// random_utils.F90:381-403 with random_direction (:276-298) and random_gaussian (:251-266).
// basis is column-major D x D; Gaussian element (row r, column c) of basis number
// `col0/D` comes from stream (tag, uid, a = col0 + c, b = r/2), lane r%2.
void random_orthonormal_basis(const Rng& rng, uint32_t tag, uint64_t uid, int col0, int D, double* basis) {
    for (int i = 0; i < D; ++i) {
        double* v = basis + (size_t)i * D;
        for (int r = 0; r < D; r += 2) {
            double u0, u1;
            rng.uniform2(tag, uid, (uint32_t)(col0 + i), (uint32_t)(r / 2), u0, u1);
            v[r] = inv_normal_cdf(u0);
            if (r + 1 < D) v[r + 1] = inv_normal_cdf(u1);
        }
        // random_direction: normalise the Gaussian vector first
        double n2 = 0.0;
        for (int r = 0; r < D; ++r) n2 += v[r] * v[r];
        double inv = std::sqrt(n2);
        for (int r = 0; r < D; ++r) v[r] /= inv;
        // Gram-Schmidt against the earlier vectors, each projection on the updated vector
        for (int j = 0; j < i; ++j) {
            const double* q = basis + (size_t)j * D;
            double dot = 0.0;
            for (int r = 0; r < D; ++r) dot += v[r] * q[r];
            for (int r = 0; r < D; ++r) v[r] -= dot * q[r];
        }
        double m2 = 0.0;
        for (int r = 0; r < D; ++r) m2 += v[r] * v[r];
        double m = std::sqrt(m2);
        for (int r = 0; r < D; ++r) v[r] /= m;
    }
}

// random_utils.F90:505-532 on deck[1..n-1] (deck[0] stays), draws from TAG_SHUF.
void shuffle_deck_tail(const Rng& rng, uint64_t uid, std::vector<int>& deck) {
    int n = (int)deck.size() - 1;  // size of deck(2:)
    for (int i = n; i >= 1; --i) {
        double u = rng.uniform(TAG_SHUF, uid, (uint32_t)i, 0);
        int j = (int)std::ceil(u * i);
        if (j < 1) j = 1;
        std::swap(deck[i], deck[j]);  // deck(2:)(i) == deck[i] zero-based
    }
}

struct Likelihood {
    int kind = 0;  // 0 gaussian, 1 rastrigin, 2 correlated gaussian, 3 host callback
    std::vector<double> mu, sigma;       // kind 0
    std::vector<double> invcov;          // kind 2, column-major D x D (symmetric)
    double logdet = 0.0;                 // kind 2
    double (*cb)(double*, int, double*, int) = nullptr;  // kind 3
    // loop invariants of gaussian.f90 hoisted out of the call (the values are identical to
    // recomputing them per call as the Fortran source is written; this only keeps the timed
    // CPU baseline from being slower than an optimising Fortran compiler would make it)
    double gauss_norm = 0.0, Vn = 0.0, log_rast = 0.0;
};

struct Prior {
    int kind = 0;  // 0 uniform box (priors.f90:40-55), 1 host callback
    std::vector<double> lo, hi;
    void (*cb)(double*, double*, int) = nullptr;
};

}  // namespace

extern "C" {

struct oracle_settings {
    int nDims, nDerived, nlive, num_repeats, nprior, nfail;
    int do_clustering;
    double precision_criterion, logzero;
    int max_ndead;
    double boost_posterior;
    int posteriors, equals, cluster_posteriors;
    double compression_factor;
    int seed;
    int batch_K;  // 0 = reference mode, >=1 batched generations
};

struct oracle_result {
    double logZ, logZerr;     // calculate_logZ_estimate
    double logZ_raw, logZ2_raw;  // RTI%logZ, RTI%logZ2
    long long ndead, nlike, nchains, ngenerations, nupdates, nfailures;
    long long nslices, nphantoms_final;
    double seconds;
    long long ncluster, nsplits;  // clusters at the end of the sampling loop; splits (reference mode) / labels (batched mode)
};

typedef void (*oracle_dumper_t)(int ndead, int nlive, int npars, double* live, double* dead, double* logweights,
                                double logZ, double logZerr);
}

namespace {

struct Cluster {
    std::vector<double> live;     // T x nlive, point-major
    std::vector<double> phantom;  // T x nphantom
    std::vector<double> stack_l;  // pos_l of the posterior stack since the last update (deaths of this cluster)
    std::vector<long long> stack_i;  // ... and the index of each of those deaths in Run::dead
    int nlive = 0, nphantom = 0;
    double logZp, logXp, logZXp, logZp2, logZpXp;
    double logLp;
    int imin = -1;
    std::vector<double> covmat, cholesky;  // column-major D x D
};

std::vector<int> g_grade_dims, g_grade_repeats;   // set by oracle_set_grades, read by every Run until cleared
// phantoms promoted to posterior samples by the last run (clean_phantoms, run_time_info.f90:820-877), read by
// oracle_last_boosted: rows [theta, phi, birth, logL], the dead point whose weight each one carries, log-weight + logL
// dynamic nlive (settings%loglikes / settings%nlives, sorted by loglike: settings.f90:234-235), oracle_set_nlives;
// read by the reference schedule of every Run until cleared
std::vector<double> g_dyn_loglikes;
std::vector<int> g_dyn_nlives;
std::vector<double> g_init_cubes;   // oracle_set_initial_cubes: the next run's live points (cube_samples), one-shot
std::vector<double> g_boost_rows, g_boost_logw;
// evidence clusters of the last batched run with clustering: per cluster {logZp, logZp2, logXp at the end of sampling}, the
// clusters alive at the end first (in label order), then the deleted ones in order of deletion (oracle_last_clusters)
std::vector<double> g_last_clusters;
int g_last_nactive = 0;
std::vector<int> g_last_dead_cluster, g_last_uid_parent, g_last_cluster_uid;   // identities: per dead point, per identity its parent, per listed cluster
std::vector<long long> g_boost_dead;

struct Run {
    oracle_settings S;
    Likelihood like;
    Prior pri;
    Rng rng;
    int D, P, T, R;
    int h0, p0, d0, b0, l0;
    std::vector<int> grade_dims, grade_repeats;

    std::vector<double> sc_R, sc_L, sc_cube, sc_theta;  // scratch (no heap traffic in the hot loop)
    std::vector<Cluster> cl;
    std::vector<double> logXpXq;  // ncluster x ncluster
    double logZ, logZ2, logX_last_update;
    std::vector<double> dead;  // T x ndead
    std::vector<double> logweights;
    long long ndead = 0, nlike = 0, nchains = 0, ngen = 0, nupdates = 0, nfail_total = 0, nslices = 0;
    oracle_dumper_t dumper = nullptr;

    void init_layout() {
        D = S.nDims; P = S.nDerived; T = 2 * D + P + 2; R = S.num_repeats;
        // fast/slow grades (oracle_set_grades): R is the total over the grades
        grade_dims = g_grade_dims; grade_repeats = g_grade_repeats;
        if (grade_dims.empty() || std::accumulate(grade_dims.begin(), grade_dims.end(), 0) != D) {
            grade_dims.assign(1, D); grade_repeats.assign(1, R);
        }
        R = std::accumulate(grade_repeats.begin(), grade_repeats.end(), 0);
        h0 = 0; p0 = D; d0 = 2 * D; b0 = 2 * D + P; l0 = b0 + 1;  // settings.f90:163-182 (zero-based)
        rng.seed = (uint32_t)S.seed;
    }

    // ---- prior + likelihood (calculate.f90:6-50) ----
    void prior(const double* cube, double* theta) {
        if (pri.kind == 0) {
            for (int i = 0; i < D; ++i) theta[i] = pri.lo[i] + (pri.hi[i] - pri.lo[i]) * cube[i];
        } else {
            std::vector<double> c(cube, cube + D);
            pri.cb(c.data(), theta, D);
        }
    }
    double loglikelihood(const double* theta, double* phi) {
        switch (like.kind) {
            case 0: {  // gaussian.f90:12-41
                double norm = like.gauss_norm, chi = 0.0, r2 = 0.0;
                for (int i = 0; i < D; ++i) {
                    double z = (theta[i] - like.mu[i]) / like.sigma[i];
                    chi += z * z;
                    r2 += (theta[i] - like.mu[i]) * (theta[i] - like.mu[i]);
                }
                if (P >= 1) phi[0] = std::sqrt(r2);
                if (P >= 2) {
                    // log( r^D * Vn(D) ), Vn = pi^(D/2) / Gamma(1 + D/2)   (utils.F90:754-760)
                    phi[1] = std::log(std::pow(phi[0], (double)D) * like.Vn);
                }
                for (int i = 2; i < P; ++i) phi[i] = 0.0;
                return -norm - chi / 2.0;
            }
            case 1: {  // rastrigin.f90:20-35
                const double TwoPi = 8.0 * std::atan(1.0);
                double s = 0.0;
                for (int i = 0; i < D; ++i)
                    s += like.log_rast + theta[i] * theta[i] - 10.0 * std::cos(TwoPi * theta[i]);
                for (int i = 0; i < P; ++i) phi[i] = 0.0;
                return -s;
            }
            case 2: {  // utils.F90:1028-1048
                double q = 0.0;
                for (int r = 0; r < D; ++r) {
                    double y = 0.0;
                    for (int c = 0; c < D; ++c) y += like.invcov[r + (size_t)c * D] * (theta[c] - like.mu[c]);
                    q += (theta[r] - like.mu[r]) * y;
                }
                for (int i = 0; i < P; ++i) phi[i] = 0.0;
                return -(D * LOG_TWO_PI + like.logdet) / 2.0 - q / 2.0;
            }
            default: {
                std::vector<double> th(theta, theta + D);
                return like.cb(th.data(), D, phi, P);
            }
        }
    }
    void calculate_point(double* point, long long& n) {
        const double* cube = point + h0;
        bool out = false;
        for (int i = 0; i < D; ++i) out = out || cube[i] < 0.0 || cube[i] > 1.0;
        double logL;
        double* theta = point + p0;
        double* phi = point + d0;
        if (out) {
            for (int i = 0; i < D; ++i) theta[i] = 0.0;
            logL = S.logzero;
        } else {
            prior(cube, theta);
            logL = loglikelihood(theta, phi);
        }
        if (logL > S.logzero) n += 1;
        point[l0] = logL;
    }

    // ---- chordal_sampling.f90 ----
    // generate_nhats, chordal_sampling.f90:94-145.  For every grade g: num_repeats(g) directions from random
    // orthonormal bases of the sub-space of the dimensions of grades >= g (zero in the slower dimensions), then one
    // shuffle of all columns but the first.  The Gaussian element (row r of the sub-space, column c) comes from
    // stream (TAG_DIR, uid, a = c, b = r/2), lane r%2.
    void generate_nhats(uint64_t uid, std::vector<double>& nhats) {
        nhats.assign((size_t)D * R, 0.0);
        std::vector<double> raw((size_t)D * R, 0.0);
        int col0 = 0, off = 0;
        for (size_t g = 0; g < grade_dims.size(); ++g) {
            const int Dg = D - off, Rg = grade_repeats[g];
            std::vector<double> basis((size_t)Dg * Dg);
            auto put = [&](int col, int count) {
                for (int i = 0; i < count; ++i)
                    std::copy(basis.begin() + (size_t)i * Dg, basis.begin() + (size_t)(i + 1) * Dg,
                              raw.begin() + (size_t)(col + i) * D + off);
            };
            int lower = 0;
            // random_orthonormal_bases: full bases while upper_index < num_nhats, then one more, truncated
            while (lower + Dg < Rg) {
                random_orthonormal_basis(rng, TAG_DIR, uid, col0 + lower, Dg, basis.data());
                put(col0 + lower, Dg);
                lower += Dg;
            }
            if (Rg > 0) {
                random_orthonormal_basis(rng, TAG_DIR, uid, col0 + lower, Dg, basis.data());
                put(col0 + lower, Rg - lower);
            }
            col0 += Rg;
            off += grade_dims[g];
        }
        std::vector<int> deck(R);
        std::iota(deck.begin(), deck.end(), 0);
        shuffle_deck_tail(rng, uid, deck);
        for (int i = 0; i < R; ++i)
            std::copy(raw.begin() + (size_t)deck[i] * D, raw.begin() + (size_t)(deck[i] + 1) * D, nhats.begin() + (size_t)i * D);
    }

    // slice_sample, chordal_sampling.f90:163-273
    void slice_sample(double logL, const double* nhat, const double* x0, double w, uint64_t uid, int step,
                      double* baby, long long& n) {
        sc_R.assign(T, 0.0); sc_L.assign(T, 0.0);
        std::vector<double>&Rp = sc_R, &Lp = sc_L;
        std::fill(baby, baby + T, 0.0);
        double u0 = rng.uniform(TAG_SLICE, uid, (uint32_t)step, 0);
        for (int i = 0; i < D; ++i) {
            Lp[h0 + i] = x0[h0 + i] - u0 * w * nhat[i];
            Rp[h0 + i] = x0[h0 + i] + (1.0 - u0) * w * nhat[i];
        }
        calculate_point(Rp.data(), n);
        calculate_point(Lp.data(), n);
        int i_step = 0;
        while (Rp[l0] >= logL && Rp[l0] > S.logzero) {
            i_step++;
            for (int i = 0; i < D; ++i) Rp[h0 + i] = x0[h0 + i] + nhat[i] * w * i_step;
            calculate_point(Rp.data(), n);
        }
        i_step = 0;
        while (Lp[l0] >= logL && Lp[l0] > S.logzero) {
            i_step++;
            for (int i = 0; i < D; ++i) Lp[h0 + i] = x0[h0 + i] - nhat[i] * w * i_step;
            calculate_point(Lp.data(), n);
        }
        for (i_step = 0; i_step <= 100; ++i_step) {
            double dL2 = 0.0, dR2 = 0.0;
            for (int i = 0; i < D; ++i) {
                dL2 += (x0[h0 + i] - Lp[h0 + i]) * (x0[h0 + i] - Lp[h0 + i]);
                dR2 += (x0[h0 + i] - Rp[h0 + i]) * (x0[h0 + i] - Rp[h0 + i]);
            }
            double x0Ld = std::sqrt(dL2), x0Rd = std::sqrt(dR2);
            double u = rng.uniform(TAG_SLICE, uid, (uint32_t)step, (uint32_t)(1 + i_step));
            double t = u * (x0Rd + x0Ld) - x0Ld;
            for (int i = 0; i < D; ++i) baby[h0 + i] = x0[h0 + i] + t * nhat[i];
            calculate_point(baby, n);
            if (baby[l0] < logL || baby[l0] <= S.logzero) {
                double dot = 0.0;
                for (int i = 0; i < D; ++i) dot += (baby[h0 + i] - x0[h0 + i]) * nhat[i];
                if (dot > 0.0) std::copy(baby, baby + T, Rp.begin());
                else std::copy(baby, baby + T, Lp.begin());
            } else {
                return;
            }
        }
        baby[l0] = S.logzero;  // "Non deterministic loglikelihood"
    }

    // SliceSampling, chordal_sampling.f90:7-92.  babies: T x R point-major.
    void slice_sampling(double logL, const double* seed_point, const double* cholesky, uint64_t uid,
                        std::vector<double>& babies, long long& n) {
        std::vector<double> nhats;
        generate_nhats(uid, nhats);
        // nhats = matmul(cholesky, nhats)
        std::vector<double> wh((size_t)D * R);
        for (int c = 0; c < R; ++c)
            for (int r = 0; r < D; ++r) {
                double s = 0.0;
                for (int k = 0; k < D; ++k) s += cholesky[r + (size_t)k * D] * nhats[k + (size_t)c * D];
                wh[r + (size_t)c * D] = s;
            }
        babies.assign((size_t)T * R, 0.0);
        std::vector<double> prev(seed_point, seed_point + T), nhat(D);
        for (int i = 0; i < R; ++i) {
            double w2 = 0.0;
            for (int r = 0; r < D; ++r) w2 += wh[r + (size_t)i * D] * wh[r + (size_t)i * D];
            double w = std::sqrt(w2);
            for (int r = 0; r < D; ++r) nhat[r] = wh[r + (size_t)i * D] / w;
            w *= 3.0;
            double* baby = babies.data() + (size_t)i * T;
            slice_sample(logL, nhat.data(), prev.data(), w, uid, i, baby, n);
            std::copy(baby, baby + T, prev.begin());
            nslices++;
        }
    }

    // ---- run_time_info.f90 ----
    void initialise() {
        cl.assign(1, Cluster());
        Cluster& c = cl[0];
        c.logZp = c.logZXp = c.logZp2 = c.logZpXp = S.logzero;
        c.logXp = 0.0;
        c.logLp = S.logzero;
        c.covmat.assign((size_t)D * D, 0.0);
        c.cholesky.assign((size_t)D * D, 0.0);
        for (int i = 0; i < D; ++i) c.covmat[i + (size_t)i * D] = c.cholesky[i + (size_t)i * D] = 1.0;
        logXpXq.assign(1, 0.0);
        logZ = logZ2 = S.logzero;
        logX_last_update = 0.0;
    }
    double& XX(int p, int q) { return logXpXq[(size_t)p * cl.size() + q]; }

    // update_evidence, run_time_info.f90:211-296, on a set of clusters C with the cross-moment matrix M (the reference
    // schedule: the clusters that hold the points; the batched schedule with clustering: the evidence clusters `ev`)
    double update_evidence_in(std::vector<Cluster>& C, std::vector<double>& M, int p) {
        Cluster& c = C[p];
        const int nc = (int)C.size();
        auto X2 = [&](int a, int b) -> double& { return M[(size_t)a * nc + b]; };
        const double log2 = std::log(2.0);
        double logL = c.logLp;
        double lognp = std::log(c.nlive + 0.0), lognp1 = std::log(c.nlive + 1.0), lognp2 = std::log(c.nlive + 2.0);
        double logweight = c.logXp - lognp1;
        logincexp(logZ, c.logXp + logL - lognp1);
        logincexp(c.logZp, c.logXp + logL - lognp1);
        c.logXp = c.logXp + lognp - lognp1;
        logincexp(logZ2, log2 + c.logZXp + logL - lognp1, log2 + X2(p, p) + 2 * logL - lognp1 - lognp2);
        c.logZXp = c.logZXp + lognp - lognp1;
        logincexp(c.logZXp, X2(p, p) + logL + lognp - lognp1 - lognp2);
        for (int q = 0; q < nc; ++q)
            if (q != p) logincexp(C[q].logZXp, X2(p, q) + logL - lognp1);
        logincexp(c.logZp2, log2 + c.logZpXp + logL - lognp1, log2 + X2(p, p) + 2 * logL - lognp1 - lognp2);
        c.logZpXp = c.logZpXp + lognp - lognp1;
        logincexp(c.logZpXp, X2(p, p) + logL + lognp - lognp1 - lognp2);
        X2(p, p) = X2(p, p) + lognp - lognp2;
        for (int q = 0; q < nc; ++q)
            if (q != p) {
                X2(p, q) += lognp - lognp1;
                X2(q, p) += lognp - lognp1;
            }
        return logweight;
    }
    double update_evidence(int p) { return update_evidence_in(cl, logXpXq, p); }

    void find_min() {
        for (auto& c : cl) {
            c.imin = -1;
            double best = 0.0;
            for (int i = 0; i < c.nlive; ++i) {
                double l = c.live[(size_t)i * T + l0];
                if (c.imin < 0 || l < best) { best = l; c.imin = i; }  // minloc: first minimum
            }
            c.logLp = (c.imin < 0) ? HUGE_D : best;
        }
    }
    int min_cluster() {
        int p = 0;
        for (int q = 1; q < (int)cl.size(); ++q)
            if (cl[q].logLp < cl[p].logLp) p = q;
        return p;
    }
    double sum_logX() {
        std::vector<double> v;
        if (pce()) { for (auto& c : ev) v.push_back(c.logXp); return logsumexp(v.data(), v.size()); }
        for (auto& c : cl) v.push_back(c.logXp);
        return logsumexp(v.data(), v.size());
    }
    int total_live() {
        int n = 0;
        for (auto& c : cl) n += c.nlive;
        return n;
    }
    void push_dead(const double* rec, double logw) {
        dead.insert(dead.end(), rec, rec + T);
        logweights.push_back(logw);
        ndead++;
    }

    // delete_outermost_point, run_time_info.f90:789-817 (swap-with-last deletion)
    void delete_outermost_point() {
        int p = min_cluster();
        Cluster& c = cl[p];
        double logw = update_evidence(p);
        std::vector<double> rec(c.live.begin() + (size_t)c.imin * T, c.live.begin() + (size_t)(c.imin + 1) * T);
        std::copy(c.live.begin() + (size_t)(c.nlive - 1) * T, c.live.begin() + (size_t)c.nlive * T,
                  c.live.begin() + (size_t)c.imin * T);
        c.nlive--;
        c.live.resize((size_t)c.nlive * T);
        find_min();
        push_dead(rec.data(), logw);
        c.stack_l.push_back(rec[l0]);
        c.stack_i.push_back(ndead - 1);
    }

    int identify_cluster(const double* point) {
        if (cl.size() == 1) return 0;
        double best = HUGE_D;
        int which = 0;
        for (int p = 0; p < (int)cl.size(); ++p)
            for (int i = 0; i < cl[p].nlive; ++i) {
                double d2 = 0.0;
                const double* q = &cl[p].live[(size_t)i * T + h0];
                for (int k = 0; k < D; ++k) d2 += (point[h0 + k] - q[k]) * (point[h0 + k] - q[k]);
                if (d2 < best) { best = d2; which = p; }
            }
        return which;
    }

    // replace_point, run_time_info.f90:716-787 (constant nlive schedule)
    bool replace_point(const std::vector<double>& babies, int cluster_add) {
        double logL = cl[0].logLp;
        for (auto& c : cl) logL = std::min(logL, c.logLp);
        for (int i = 0; i < R - 1; ++i) {
            const double* pt = babies.data() + (size_t)i * T;
            if (pt[l0] > logL && identify_cluster(pt) == cluster_add) {
                Cluster& c = cl[cluster_add];
                c.phantom.insert(c.phantom.end(), pt, pt + T);
                c.nphantom++;
            }
        }
        const double* pt = babies.data() + (size_t)(R - 1) * T;
        bool replaced = false;
        if (pt[l0] > logL) {
            if (identify_cluster(pt) == cluster_add) {
                // run_time_info.f90:766-771: the target is that of the largest threshold below the contour
                int nlive = S.nlive;
                {
                    int best = -1;
                    for (size_t q = 0; q < g_dyn_loglikes.size(); ++q)
                        if (logL > g_dyn_loglikes[q] && (best < 0 || g_dyn_loglikes[q] > g_dyn_loglikes[(size_t)best])) best = (int)q;
                    if (best >= 0) nlive = g_dyn_nlives[(size_t)best];
                }
                if (total_live() >= std::max(nlive, 1)) {
                    delete_outermost_point();
                    replaced = true;
                }
                if (total_live() < nlive) {
                    Cluster& c = cl[cluster_add];
                    c.live.insert(c.live.end(), pt, pt + T);
                    c.nlive++;
                    find_min();
                }
            }
        } else {
            push_dead(pt, S.logzero);
        }
        return replaced;
    }

    // RTI%thin_posterior, generate.F90:311-316
    double thin_posterior() const {
        if (!(S.posteriors || S.equals)) return 0.0;   // run_time_info.f90:858: only when posterior files are asked for
        return S.boost_posterior < 0.0 ? 1.0 : S.boost_posterior / (double)R;
    }
    // The posterior conversion of a phantom that clean_phantoms removes (run_time_info.f90:846-868): with
    // probability thin_posterior it becomes a posterior sample carrying the weight of the death of its cluster,
    // since the last update, with the smallest logL above its own (minloc over the mask, :846-848).  The
    // reference draws the Bernoulli trial from its sequential stream; here it is addressed by the bits of the
    // phantom's logL, so that it does not depend on the order in which the phantoms are visited.  (The engine
    // addresses it by the bits of ITS logL, equal to this one only to rounding: with thin_posterior < 1 the two
    // promote different subsets of the same phantoms; tests compare the full lists, boost_posterior < 0.)
    void boost(const Cluster& c, const double* pt) {
        const double thin = thin_posterior();
        if (!(thin > 0.0)) return;
        uint64_t bits;
        std::memcpy(&bits, &pt[l0], 8);
        if (!(rng.uniform(TAG_BOOST, bits, 0, 0) < thin)) return;
        long long best = -1;
        double lbest = HUGE_D;
        for (size_t j = 0; j < c.stack_l.size(); ++j)
            if (c.stack_l[j] > pt[l0] && c.stack_l[j] < lbest) { lbest = c.stack_l[j]; best = c.stack_i[j]; }
        if (best < 0) return;   // cannot happen: the caller found a death above this phantom
        g_boost_rows.insert(g_boost_rows.end(), pt + p0, pt + T);
        g_boost_dead.push_back(best);
        g_boost_logw.push_back(logweights[(size_t)best] + pt[l0]);
    }

    // clean_phantoms, run_time_info.f90:820-877: a phantom is dropped as soon as some death
    // of its cluster since the last update has a larger logL.
    void clean_phantoms() {
        for (auto& c : cl) {
            if (c.stack_l.empty()) continue;
            double lmax = *std::max_element(c.stack_l.begin(), c.stack_l.end());
            int i = 0;
            while (i < c.nphantom) {
                if (lmax > c.phantom[(size_t)i * T + l0]) {
                    boost(c, &c.phantom[(size_t)i * T]);
                    std::copy(c.phantom.begin() + (size_t)(c.nphantom - 1) * T, c.phantom.begin() + (size_t)c.nphantom * T,
                              c.phantom.begin() + (size_t)i * T);
                    c.nphantom--;
                } else {
                    ++i;
                }
            }
            c.phantom.resize((size_t)c.nphantom * T);
            c.stack_l.clear();
            c.stack_i.clear();
        }
    }
    // batched mode keeps the phantom pool in birth order (stable compaction) so that the
    // covariance sums run over the same sequence as the GPU engine's pool.
    void clean_phantoms_stable() {
        for (auto& c : cl) {
            if (c.stack_l.empty()) continue;
            double lmax = *std::max_element(c.stack_l.begin(), c.stack_l.end());
            int w = 0;
            const bool lbl = (int)phlab.size() == c.nphantom && &c == &cl[0];
            for (int i = 0; i < c.nphantom; ++i) {
                if (!(lmax > c.phantom[(size_t)i * T + l0])) {
                    if (w != i)
                        std::copy(c.phantom.begin() + (size_t)i * T, c.phantom.begin() + (size_t)(i + 1) * T,
                                  c.phantom.begin() + (size_t)w * T);
                    if (lbl) phlab[w] = phlab[i];
                    ++w;
                } else {
                    boost(c, &c.phantom[(size_t)i * T]);
                }
            }
            if (lbl) phlab.resize(w);
            c.nphantom = w;
            c.phantom.resize((size_t)w * T);
            c.stack_l.clear();
            c.stack_i.clear();
        }
    }

    // calculate_covmats, run_time_info.f90:601-641
    void calculate_covmats() {
        for (auto& c : cl) {
            int N = c.nlive + c.nphantom;
            if (N == 0) continue;
            std::vector<double> mean(D, 0.0);
            for (int i = 0; i < c.nlive; ++i)
                for (int k = 0; k < D; ++k) mean[k] += c.live[(size_t)i * T + h0 + k];
            std::vector<double> mp(D, 0.0);
            for (int i = 0; i < c.nphantom; ++i)
                for (int k = 0; k < D; ++k) mp[k] += c.phantom[(size_t)i * T + h0 + k];
            for (int k = 0; k < D; ++k) mean[k] = (mean[k] + mp[k]) / N;
            std::vector<double> cov((size_t)D * D, 0.0), covp((size_t)D * D, 0.0), dv(D);
            for (int i = 0; i < c.nlive; ++i) {
                for (int k = 0; k < D; ++k) dv[k] = c.live[(size_t)i * T + h0 + k] - mean[k];
                for (int b = 0; b < D; ++b)
                    for (int a = 0; a < D; ++a) cov[a + (size_t)b * D] += dv[a] * dv[b];
            }
            for (int i = 0; i < c.nphantom; ++i) {
                for (int k = 0; k < D; ++k) dv[k] = c.phantom[(size_t)i * T + h0 + k] - mean[k];
                for (int b = 0; b < D; ++b)
                    for (int a = 0; a < D; ++a) covp[a + (size_t)b * D] += dv[a] * dv[b];
            }
            for (size_t k = 0; k < cov.size(); ++k) c.covmat[k] = (cov[k] + covp[k]) / N;
            calc_cholesky(c.covmat.data(), c.cholesky.data(), D);
        }
    }

    // ---- clusters, reference mode: add_cluster / delete_cluster / do_clustering -----------------
    int ncluster_dead = 0;
    std::vector<double> logZp_dead, logZp2_dead;

    // delete_cluster, run_time_info.f90:507-598: the first cluster without live points leaves the active arrays
    bool delete_cluster() {
        int p = -1;
        for (int q = 0; q < (int)cl.size(); ++q) if (cl[q].nlive == 0) { p = q; break; }
        if (p < 0) return false;
        const int nc = (int)cl.size();
        logZp_dead.push_back(cl[p].logZp);
        logZp2_dead.push_back(cl[p].logZp2);
        ++ncluster_dead;
        std::vector<double> xx((size_t)(nc - 1) * (nc - 1));
        for (int a = 0, ia = 0; a < nc; ++a) {
            if (a == p) continue;
            for (int b = 0, ib = 0; b < nc; ++b) {
                if (b == p) continue;
                xx[(size_t)ia * (nc - 1) + ib] = logXpXq[(size_t)a * nc + b];
                ++ib;
            }
            ++ia;
        }
        cl.erase(cl.begin() + p);
        logXpXq.swap(xx);
        return true;
    }

    // add_cluster, run_time_info.f90:303-505: cluster p splits into num pieces appended after the surviving ones
    void add_cluster(int p, const std::vector<int>& labels, int num) {
        const int nc_old = (int)cl.size();
        const Cluster old = cl[p];
        std::vector<Cluster> before = cl;                       // old_phantom: every cluster's phantoms, old order
        std::vector<int> keep;                                  // old_save
        for (int q = 0; q < nc_old; ++q) if (q != p) keep.push_back(q);
        const int nold = (int)keep.size(), nc = nold + num;
        std::vector<double> xpq(nold);
        for (int a = 0; a < nold; ++a) xpq[a] = logXpXq[(size_t)p * nc_old + keep[a]];
        const double xpp = logXpXq[(size_t)p * nc_old + p];
        std::vector<double> xx((size_t)nc * nc, 0.0);
        for (int a = 0; a < nold; ++a)
            for (int b = 0; b < nold; ++b) xx[(size_t)a * nc + b] = logXpXq[(size_t)keep[a] * nc_old + keep[b]];
        std::vector<Cluster> ncl;
        for (int q : keep) ncl.push_back(cl[q]);
        for (int i = 0; i < num; ++i) {
            Cluster c;
            c.logLp = S.logzero;
            c.covmat.assign((size_t)D * D, 0.0);
            c.cholesky.assign((size_t)D * D, 0.0);
            for (int d = 0; d < D; ++d) c.covmat[d + (size_t)d * D] = c.cholesky[d + (size_t)d * D] = 1.0;
            ncl.push_back(c);
        }
        for (int i = 0; i < old.nlive; ++i) {                   // 3) live points to their new clusters, in order
            Cluster& c = ncl[nold + labels[i]];
            c.live.insert(c.live.end(), old.live.begin() + (size_t)i * T, old.live.begin() + (size_t)(i + 1) * T);
            c.nlive++;
        }
        cl.swap(ncl);
        logXpXq.swap(xx);
        find_min();
        for (auto& c : cl) { c.phantom.clear(); c.nphantom = 0; }   // 4) every phantom is re-assigned
        for (int q = 0; q < nc_old; ++q)
            for (int i = 0; i < before[q].nphantom; ++i) {
                const double* pt = &before[q].phantom[(size_t)i * T];
                const int j = identify_cluster(pt);
                if (pt[l0] > cl[j].logLp) {
                    cl[j].phantom.insert(cl[j].phantom.end(), pt, pt + T);
                    cl[j].nphantom++;
                }
            }
        // 5) volumes and evidences apportioned by the live + phantom counts
        std::vector<double> logni(num), logni1(num);
        for (int i = 0; i < num; ++i) {
            logni[i] = std::log(cl[nold + i].nlive + cl[nold + i].nphantom + 0.0);
            logni1[i] = std::log(cl[nold + i].nlive + cl[nold + i].nphantom + 1.0);
        }
        const double logn = logsumexp(logni.data(), logni.size()), logn1 = logaddexp(logn, 0.0);
        for (int i = 0; i < num; ++i) {
            Cluster& c = cl[nold + i];
            c.logXp = old.logXp + logni[i] - logn;
            c.logZXp = old.logZXp + logni[i] - logn;
            c.logZp = old.logZp + logni[i] - logn;
            c.logZp2 = old.logZp2 + logni[i] + logni1[i] - logn - logn1;
            c.logZpXp = old.logZpXp + logni[i] + logni1[i] - logn - logn1;
            for (int a = 0; a < nold; ++a) XX(nold + i, a) = XX(a, nold + i) = xpq[a] + logni[i] - logn;
            for (int j = 0; j < num; ++j)
                XX(nold + i, nold + j) = (i == j) ? xpp + logni[i] + logni1[i] - logn - logn1
                                                  : xpp + logni[i] + logni[j] - logn - logn1;
        }
    }

    // do_clustering, clustering.f90:253-324
    bool do_clustering() {
        bool found = false;
        const int num_old = (int)cl.size();
        int i = 0;
        while (i < num_old) {
            const int nl = cl[i].nlive;
            if (nl > 2) {
                std::vector<const double*> pts(nl);
                for (int j = 0; j < nl; ++j) pts[j] = &cl[i].live[(size_t)j * T + h0];
                int num = 1;
                const std::vector<int> labels = NN_clustering(similarity_matrix(pts, D), nl, num);
                if (num > 1) { found = true; add_cluster(i, labels, num); nsplits++; }
                else ++i;
            } else ++i;
        }
        return found;
    }
    long long nsplits = 0;

    // ---- clusters, batched mode (the engine's schedule) ------------------------------------------
    // The points stay in cl[0] (one live array, one phantom pool), lab[slot] / phlab[i] carry the cluster of every live
    // point / phantom, and the evidence is kept PER CLUSTER in `ev` exactly as the reference keeps it
    // (run_time_info.f90:211-296 update_evidence, :303-505 add_cluster, :507-598 delete_cluster):
    //   * a generation's K deaths go through update_evidence in death order, each in the cluster of the dying point with
    //     that cluster's live count (which falls as its points die);
    //   * a cluster left without live points is deleted (its local evidence joins the dead clusters');
    //   * the seed of a chain is drawn as GenerateSeed draws it (generate.F90:19-55): a cluster in proportion to its
    //     volume, then one of its surviving points (in rank order); the babies join the seed's cluster;
    //   * at every update every cluster is searched for sub-clusters (do_clustering, clustering.f90:253-324:
    //     NN_clustering on its own live points) and split by add_cluster -- volumes and evidences apportioned by the
    //     live + phantom counts of the pieces -- where a phantom belongs to the cluster of its nearest live point
    //     (identify_cluster); every cluster with more than nDims points gets the covariance / Cholesky factor of its own
    //     live + phantom points (the others keep the global one).
    // One difference of bookkeeping, none of substance: the phantoms' labels are recomputed from the nearest live point
    // at every update (the reference re-assigns them when a cluster splits and drops the ones of a deleted cluster; in
    // between it never reads them).
    std::vector<int> lab, phlab;                  // label of every live slot / phantom record (all in cl[0])
    std::vector<Cluster> ev;                      // the evidence clusters (scalar fields, nlive, cholesky)
    std::vector<double> evXX;                     // their cross moments log<X_p X_q>, ev.size() squared
    int ncl_b = 1, ncl_max_b = 1;
    std::vector<int> ev_uid;                      // persistent identity of every evidence cluster (a split creates new ones)
    std::vector<int> uid_parent;                  // ... and the cluster it was split from (-1: the initial cluster)
    std::vector<int> dead_cluster;                // for every dead point: the identity of its cluster at its death
    std::vector<int> dead_uid;                    // identities of the deleted clusters, in order of deletion
    // do_clustering = 1: the evidence stays GLOBAL (one nested-sampling run over the whole live set: exact whatever the
    //   shape of the posterior) and every death is ATTRIBUTED to the cluster of the dying point -- local evidence
    //   Z_p = sum over its deaths of w L with the global weight w = X/(n+1), and its second moment by the same
    //   recurrences as the global one (run_time_info.f90:211-296 with the global X): this is what the engine does.
    // do_clustering = 2: every cluster keeps its own volume as well, as the reference does (local volumes shrink with the
    //   cluster's own deaths, seeds are drawn by volume).  Kept for the record: the estimator is biased high -- also in
    //   the reference schedule -- see scripts/r02_cluster_bias.py and DESIGN.md section 5.4.
    bool pce() const { return S.batch_K > 0 && S.do_clustering == 2; }
    bool attr() const { return S.batch_K > 0 && S.do_clustering == 1; }
    bool evc() const { return S.batch_K > 0 && S.do_clustering != 0; }
    // a death in evidence cluster p under the global accounting (attr): the global update_evidence gave log-weight logw
    // with n live points before the death; Z_p, <Z_p^2> and <Z_p X> follow the global recurrences restricted to p
    void attribute_death(int p, double logL, int n_before, double XX_before) {
        const double log2 = std::log(2.0);
        const double lognp = std::log(n_before + 0.0), lognp1 = std::log(n_before + 1.0), lognp2 = std::log(n_before + 2.0);
        Cluster& c = ev[p];
        const double X_before = cl[0].logXp - (lognp - lognp1);   // (the global update already shrank it)
        logincexp(c.logZp, X_before + logL - lognp1);
        logincexp(c.logZp2, log2 + c.logZpXp + logL - lognp1, log2 + XX_before + 2 * logL - lognp1 - lognp2);
        for (auto& q : ev) q.logZpXp += lognp - lognp1;          // X shrinks for every cluster's <Z_q X>
        logincexp(c.logZpXp, XX_before + logL + lognp - lognp1 - lognp2);
    }
    void init_ev() {
        ev.assign(1, Cluster());
        Cluster& c = ev[0];
        c.logZp = c.logZXp = c.logZp2 = c.logZpXp = S.logzero;
        c.logXp = 0.0;
        c.logLp = S.logzero;
        c.nlive = cl[0].nlive;
        c.cholesky = cl[0].cholesky;
        c.covmat = cl[0].covmat;
        evXX.assign(1, 0.0);
        lab.assign(cl[0].nlive, 0);
        phlab.clear();
        ev_uid.assign(1, 0);
        uid_parent.assign(1, -1);
        dead_cluster.clear();
        dead_uid.clear();
    }
    // delete_cluster for every evidence cluster without live points (lowest index first, as minloc picks them)
    void delete_empty_ev() {
        for (;;) {
            int p = -1;
            for (int q = 0; q < (int)ev.size(); ++q) if (ev[q].nlive == 0) { p = q; break; }
            if (p < 0 || ev.size() == 1) return;
            const int nc = (int)ev.size();
            logZp_dead.push_back(ev[p].logZp);
            logZp2_dead.push_back(ev[p].logZp2);
            dead_uid.push_back(ev_uid[p]);
            ++ncluster_dead;
            std::vector<double> xx((size_t)(nc - 1) * (nc - 1));
            for (int a2 = 0, ia = 0; a2 < nc; ++a2) {
                if (a2 == p) continue;
                for (int b2 = 0, ib = 0; b2 < nc; ++b2) {
                    if (b2 == p) continue;
                    xx[(size_t)ia * (nc - 1) + ib] = evXX[(size_t)a2 * nc + b2];
                    ++ib;
                }
                ++ia;
            }
            ev.erase(ev.begin() + p);
            ev_uid.erase(ev_uid.begin() + p);
            evXX.swap(xx);
            for (int& v : lab) if (v > p) --v;     // (no live point carries p any more)
            for (int& v : phlab) if (v > p) --v; else if (v == p) v = 0;   // recomputed at the next update anyway
        }
    }
    // add_cluster (run_time_info.f90:303-505) on the evidence clusters: cluster p splits into `num` pieces, appended
    // behind the surviving clusters; piece[j] is the piece of the j-th live point of p (in slot order), nn[i] the live
    // slot nearest to phantom i
    void add_cluster_ev(int p, const std::vector<int>& members, const std::vector<int>& piece, int num, const std::vector<int>& nn) {
        const int nc_old = (int)ev.size(), nold = nc_old - 1, nc = nold + num;
        const Cluster old = ev[p];
        std::vector<int> keep;
        for (int q = 0; q < nc_old; ++q) if (q != p) keep.push_back(q);
        std::vector<double> xpq(nold);
        for (int a2 = 0; a2 < nold; ++a2) xpq[a2] = evXX[(size_t)p * nc_old + keep[a2]];
        const double xpp = evXX[(size_t)p * nc_old + p];
        std::vector<double> xx((size_t)nc * nc, 0.0);
        for (int a2 = 0; a2 < nold; ++a2)
            for (int b2 = 0; b2 < nold; ++b2) xx[(size_t)a2 * nc + b2] = evXX[(size_t)keep[a2] * nc_old + keep[b2]];
        std::vector<Cluster> nev;
        std::vector<int> nuid;
        for (int q : keep) { nev.push_back(ev[q]); nuid.push_back(ev_uid[q]); }
        for (int i = 0; i < num; ++i) { nuid.push_back((int)uid_parent.size()); uid_parent.push_back(ev_uid[p]); }
        ev_uid.swap(nuid);
        for (int i = 0; i < num; ++i) {
            Cluster c = old;           // (cholesky / covmat: the parent's until the update recomputes them)
            c.nlive = 0;
            c.logLp = S.logzero;
            nev.push_back(c);
        }
        // the labels: clusters behind p move up, the points of p go to their pieces
        for (int& v : lab) if (v > p) --v;
        for (size_t j = 0; j < members.size(); ++j) {
            lab[members[j]] = nold + piece[j];
            nev[nold + piece[j]].nlive++;
        }
        ev.swap(nev);
        evXX.swap(xx);
        auto X2 = [&](int a2, int b2) -> double& { return evXX[(size_t)a2 * nc + b2]; };
        // live + phantom counts of the pieces (every phantom identified afresh: the cluster of its nearest live point)
        std::vector<int> nph(num, 0);
        for (int s2 : nn) if (lab[s2] >= nold) nph[lab[s2] - nold]++;
        std::vector<double> logni(num), logni1(num);
        for (int i = 0; i < num; ++i) {
            logni[i] = std::log(ev[nold + i].nlive + nph[i] + 0.0);
            logni1[i] = std::log(ev[nold + i].nlive + nph[i] + 1.0);
        }
        const double logn = logsumexp(logni.data(), logni.size()), logn1 = logaddexp(logn, 0.0);
        if (!pce()) {   // attributed evidences: Z_p splits by the counts, the volume is global (nothing of it splits)
            for (int i = 0; i < num; ++i) {
                Cluster& c = ev[nold + i];
                c.logZp = old.logZp + logni[i] - logn;
                c.logZp2 = old.logZp2 + logni[i] + logni1[i] - logn - logn1;
                c.logZpXp = old.logZpXp + logni[i] - logn;
            }
            return;
        }
        for (int i = 0; i < num; ++i) {
            Cluster& c = ev[nold + i];
            c.logXp = old.logXp + logni[i] - logn;
            c.logZXp = old.logZXp + logni[i] - logn;
            c.logZp = old.logZp + logni[i] - logn;
            c.logZp2 = old.logZp2 + logni[i] + logni1[i] - logn - logn1;
            c.logZpXp = old.logZpXp + logni[i] + logni1[i] - logn - logn1;
            for (int a2 = 0; a2 < nold; ++a2) X2(nold + i, a2) = X2(a2, nold + i) = xpq[a2] + logni[i] - logn;
            for (int j = 0; j < num; ++j)
                X2(nold + i, nold + j) = (i == j) ? xpp + logni[i] + logni1[i] - logn - logn1
                                                  : xpp + logni[i] + logni[j] - logn - logn1;
        }
    }
    void cluster_update_batched() {
        Cluster& c = cl[0];
        const int n = c.nlive;
        calculate_covmats();                      // the global covariance / factor (clusters of few points use it)
        std::vector<const double*> pts(n);
        for (int i = 0; i < n; ++i) pts[i] = &c.live[(size_t)i * T + h0];
        // identify_cluster for every phantom: the nearest live point (ties: the lowest slot)
        std::vector<int> nn(c.nphantom, 0);
        for (int i = 0; i < c.nphantom; ++i) {
            const double* x = &c.phantom[(size_t)i * T + h0];
            double best = HUGE_D;
            for (int j = 0; j < n; ++j) {
                const double d2 = dist2(x, pts[j], D);
                if (d2 < best) { best = d2; nn[i] = j; }
            }
        }
        // do_clustering (clustering.f90:253-324): every cluster is searched for sub-clusters; a cluster that splits
        // leaves its place to the next one (the loop index stays), its pieces go to the end.  At most 256 clusters
        // (the engine keeps that many factors): a split that would exceed them is not made.
        // (The reference's loop index also reaches the first of the appended pieces; NN_clustering returns a piece only
        // after searching it again and finding one cluster, so that visit never splits anything: the old clusters are
        // the ones examined here.)
        bool found = false;
        int i = 0;
        for (int remaining = (int)ev.size(); remaining > 0; --remaining) {
            std::vector<int> members;
            for (int s2 = 0; s2 < n; ++s2) if (lab[s2] == i) members.push_back(s2);
            const int nl = (int)members.size();
            int num = 1;
            std::vector<int> piece;
            if (nl > 2) {
                std::vector<const double*> mp(nl);
                for (int j = 0; j < nl; ++j) mp[j] = pts[members[j]];
                piece = NN_clustering(similarity_matrix(mp, D), nl, num);
            }
            if (num > 1 && (int)ev.size() - 1 + num <= 256) { found = true; add_cluster_ev(i, members, piece, num, nn); }
            else ++i;
        }
        if (found) nsplits++;                     // updates that split a cluster
        ncl_b = (int)ev.size();
        ncl_max_b = std::max(ncl_max_b, ncl_b);
        phlab.resize(c.nphantom);
        for (int k = 0; k < c.nphantom; ++k) phlab[k] = lab[nn[k]];
        // per-cluster covariance and factor
        for (int p = 0; p < ncl_b; ++p) {
            ev[p].cholesky = c.cholesky;
            ev[p].covmat = c.covmat;
            if (ncl_b == 1) continue;
            std::vector<const double*> mem;
            for (int s2 = 0; s2 < n; ++s2) if (lab[s2] == p) mem.push_back(pts[s2]);
            for (int k = 0; k < c.nphantom; ++k) if (phlab[k] == p) mem.push_back(&c.phantom[(size_t)k * T + h0]);
            const int N = (int)mem.size();
            if (N <= D) continue;                 // too few points for a covariance: the global factor stays
            std::vector<double> mean(D, 0.0), cov((size_t)D * D, 0.0), dv(D);
            for (const double* x : mem) for (int k = 0; k < D; ++k) mean[k] += x[k];
            for (int k = 0; k < D; ++k) mean[k] /= N;
            for (const double* x : mem) {
                for (int k = 0; k < D; ++k) dv[k] = x[k] - mean[k];
                for (int b2 = 0; b2 < D; ++b2) for (int a2 = 0; a2 < D; ++a2) cov[a2 + (size_t)b2 * D] += dv[a2] * dv[b2];
            }
            for (double& v : cov) v /= N;
            ev[p].covmat = cov;
            calc_cholesky(cov.data(), ev[p].cholesky.data(), D);
        }
    }

    double live_logZ() {
        double r = S.logzero;
        std::vector<double> ll;
        if (pce()) {   // per evidence cluster: mean likelihood of its live points times its volume
            for (int p = 0; p < (int)ev.size(); ++p) {
                ll.clear();
                for (int s2 = 0; s2 < cl[0].nlive; ++s2) if (lab[s2] == p) ll.push_back(cl[0].live[(size_t)s2 * T + l0]);
                if (!ll.empty()) logincexp(r, logsumexp(ll.data(), ll.size()) - std::log((double)ll.size()) + ev[p].logXp);
            }
            return r;
        }
        for (auto& c : cl) {
            if (c.nlive > 0) {
                ll.resize(c.nlive);
                for (int i = 0; i < c.nlive; ++i) ll[i] = c.live[(size_t)i * T + l0];
                logincexp(r, logsumexp(ll.data(), ll.size()) - std::log(c.nlive + 0.0) + c.logXp);
            }
        }
        return r;
    }
    bool more_samples_needed() {
        if (S.max_ndead == 0) return false;
        if (S.max_ndead > 0 && ndead >= S.max_ndead) return false;
        if (S.precision_criterion > 0 && live_logZ() < std::log(S.precision_criterion) + logZ) return false;
        return true;
    }
    void logZ_estimate(double& lz, double& var) {
        lz = std::max(-HUGE_D, 2 * logZ - 0.5 * logZ2);
        var = logZ2 - 2 * logZ;
    }

    // dump, nested_sampling.F90:546-590
    void dump() {
        if (!dumper) return;
        int npars = D + P + 2;
        int nl = total_live();
        std::vector<double> live_o((size_t)npars * std::max(nl, 1)), dead_o((size_t)npars * std::max<long long>(ndead, 1)),
            lw(std::max<long long>(ndead, 1));
        auto pack = [&](const double* rec, double* out) {
            for (int k = 0; k < D; ++k) out[k] = rec[p0 + k];
            for (int k = 0; k < P; ++k) out[D + k] = rec[d0 + k];
            out[D + P] = rec[b0];
            out[D + P + 1] = rec[l0];
        };
        for (long long i = 0; i < ndead; ++i) {
            pack(&dead[(size_t)i * T], &dead_o[(size_t)i * npars]);
            lw[i] = logweights[i] + dead[(size_t)i * T + l0];
        }
        if (ndead > 0) {
            double lse = logsumexp(lw.data(), (size_t)ndead);
            for (long long i = 0; i < ndead; ++i) lw[i] -= lse;
        }
        int o = 0;
        for (auto& c : cl)
            for (int i = 0; i < c.nlive; ++i) pack(&c.live[(size_t)i * T], &live_o[(size_t)(o++) * npars]);
        double lz, var;
        logZ_estimate(lz, var);
        dumper((int)ndead, nl, npars, live_o.data(), dead_o.data(), lw.data(), lz, std::sqrt(var));
    }

    // GenerateLivePoints linear, generate.F90:153-183
    void generate_live_points() {
        initialise();
        int nprior = S.nprior <= 0 ? S.nlive : S.nprior;
        Cluster& c = cl[0];
        std::vector<double> pt(T);
        uint64_t attempt = 0;
        if (!g_init_cubes.empty()) {   // cube_samples (polychord.py:650-789): the caller's live points, born from the prior
            const int npts = (int)(g_init_cubes.size() / D);
            for (int j = 0; j < npts; ++j) {
                std::fill(pt.begin(), pt.end(), 0.0);
                std::copy(g_init_cubes.begin() + (size_t)j * D, g_init_cubes.begin() + (size_t)(j + 1) * D, pt.begin() + h0);
                calculate_point(pt.data(), nlike);
                pt[b0] = S.logzero;
                c.live.insert(c.live.end(), pt.begin(), pt.end());
                c.nlive++;
            }
            g_init_cubes.clear();
            find_min();
            return;
        }
        while (c.nlive < nprior) {
            std::fill(pt.begin(), pt.end(), 0.0);
            for (int k = 0; k < D; ++k) pt[h0 + k] = rng.uniform(TAG_INIT, attempt, (uint32_t)k, 0);
            attempt++;
            calculate_point(pt.data(), nlike);
            pt[b0] = S.logzero;
            if (pt[l0] > S.logzero) {
                c.live.insert(c.live.end(), pt.begin(), pt.end());
                c.nlive++;
            }
        }
        find_min();
    }

    void do_update() {   // batched mode (nested_sampling.F90:321-368 in one piece: no cluster can die between the steps)
        logX_last_update = sum_logX();
        clean_phantoms_stable();
        dump();
        nupdates++;
        if (S.do_clustering) cluster_update_batched(); else calculate_covmats();
    }

    // ---- reference-mode main loop, nested_sampling.F90:239-374 ----
    void run_reference() {
        int nfail = S.nfail <= 0 ? S.nlive : S.nfail;
        int failures = 0;
        std::vector<double> babies;
        while (more_samples_needed() && failures <= nfail) {
            // GenerateSeed (generate.F90:19-55), single cluster or volume-weighted choice
            uint64_t uid = (uint64_t)nchains;
            int p = 0;
            if (cl.size() > 1) {
                std::vector<double> probs(cl.size());
                double lse = sum_logX();
                for (size_t q = 0; q < cl.size(); ++q) probs[q] = std::exp(cl[q].logXp - lse);
                double norm = 0.0;
                for (double v : probs) norm += v;
                double rnd = rng.uniform(TAG_SEED, uid, 1, 0), cdf = 0.0;
                p = (int)cl.size() - 1;
                for (size_t q = 0; q < cl.size(); ++q) {
                    cdf += probs[q] / norm;
                    if (rnd < cdf) { p = (int)q; break; }
                }
            }
            double u = rng.uniform(TAG_SEED, uid, 0, 0);
            int choice = (int)std::ceil(u * cl[p].nlive);
            if (choice < 1) choice = 1;
            std::vector<double> seed(cl[p].live.begin() + (size_t)(choice - 1) * T, cl[p].live.begin() + (size_t)choice * T);
            double logL = cl[p].logLp;
            slice_sampling(logL, seed.data(), cl[p].cholesky.data(), uid, babies, nlike);
            for (int i = 0; i < R; ++i) babies[(size_t)i * T + b0] = logL;
            nchains++;
            ngen++;
            if (replace_point(babies, p)) failures = 0;
            else { failures++; nfail_total++; }
            const bool update = sum_logX() <= logX_last_update + std::log(S.compression_factor);   // :321
            if (update) {
                logX_last_update = sum_logX();
                clean_phantoms();
                dump();
                nupdates++;
            }
            delete_cluster();                     // :339
            if (cl.empty()) break;                // :346
            if (update) {
                if (S.do_clustering) do_clustering();   // :351-367
                calculate_covmats();              // :368
            }
        }
    }

    // ---- batched-generation main loop (the GPU engine's schedule) ----
    // Dynamic nlive in the batched schedule (the target for the engine's row a11): a generation kills the K lowest of
    // its n live points with counts n, n-1, ... as always, then births B = target(L*) - (n - K) chains (none when the
    // survivors already exceed the target, at most 2 batch_K), so the live count moves to max(n - K, target) --
    // replace_point's rule (run_time_info.f90:766-777) taken a generation at a time.  Births k < min(B, K) take the
    // vacated slots in death order, further births are appended, vacated slots left over are closed from the top down
    // by moving the last record in.  With a constant target B = K: the schedule above, unchanged.
    int target_nlive(double contour) const {
        int nlive = S.nlive, best = -1;
        for (size_t q = 0; q < g_dyn_loglikes.size(); ++q)
            if (contour > g_dyn_loglikes[q] && (best < 0 || g_dyn_loglikes[q] > g_dyn_loglikes[(size_t)best])) best = (int)q;
        if (best >= 0) nlive = g_dyn_nlives[(size_t)best];
        return nlive;
    }

    // One generation of the batched schedule: the K lowest of the n live points die (evidence with counts n, n-1, ...
    // as the final kill-off nested_sampling.F90:381-384 applies it), B chains are seeded from the n-K survivors at the
    // contour of the K-th lowest.  Births k < min(B, K) take the vacated slots in death order, further births are
    // appended; a FAILED birth (last baby not above the contour: stale or non-deterministic likelihood) does not become
    // a live point -- it goes to the dead list with log-weight logzero, as replace_point does (run_time_info.f90:
    // 781-785) -- and counts towards nfail (nested_sampling.F90:315-319).  Slots left empty (vacated and not refilled,
    // or reserved for a failed birth) are closed by moving the records beyond the new count into them.
    // Returns the number of failed births.
    int failures_in_a_row = 0;
    int batched_generation(int K, int B, std::vector<double>& babies) {
        Cluster& c = cl[0];
        const int n = c.nlive;
        std::vector<int> order(n);
        // rank by (logL, slot)
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
            return c.live[(size_t)a * T + l0] < c.live[(size_t)b * T + l0];
        });
        const double Lstar = c.live[(size_t)order[K - 1] * T + l0];
        // K sequential deaths with n, n-1, ... live points (cf. nested_sampling.F90:381-384)
        const int nlive_save = c.nlive;
        const bool per_cluster = pce(), labelled_ev = evc();
        for (int k = 0; k < K; ++k) {
            const double* r = &c.live[(size_t)order[k] * T];
            double logw;
            if (per_cluster) {   // the death belongs to the cluster of the dying point
                const int pc = lab[order[k]];
                ev[pc].logLp = r[l0];
                logw = update_evidence_in(ev, evXX, pc);
                ev[pc].nlive--;
            } else {
                c.logLp = r[l0];
                const double xx0 = XX(0, 0);
                logw = update_evidence(0);
                if (labelled_ev) {
                    attribute_death(lab[order[k]], r[l0], c.nlive, xx0);
                    ev[lab[order[k]]].nlive--;
                }
            }
            if (labelled_ev) dead_cluster.push_back(ev_uid[lab[order[k]]]);
            push_dead(r, logw);
            c.stack_l.push_back(r[l0]);
            c.stack_i.push_back(ndead - 1);
            c.nlive--;
        }
        c.nlive = nlive_save;
        const int m = n - K;
        // survivors of every evidence cluster in rank order (GenerateSeed picks among them), and the clusters' shares
        std::vector<std::vector<int>> members;
        std::vector<double> cdf;
        if (labelled_ev) delete_empty_ev();   // a cluster that lost its last live point is deleted (delete_cluster); no survivor carries its label
        if (per_cluster) {
            members.assign(ev.size(), {});
            for (int i = K; i < n; ++i) members[lab[order[i]]].push_back(order[i]);
            if (ev.size() > 1) {
                const double lse = sum_logX();
                std::vector<double> probs(ev.size());
                double norm = 0.0;
                for (size_t q = 0; q < ev.size(); ++q) { probs[q] = std::exp(ev[q].logXp - lse); norm += probs[q]; }
                double acc = 0.0;
                for (size_t q = 0; q < ev.size(); ++q) { acc += probs[q] / norm; cdf.push_back(acc); }
            }
        }
        std::vector<double> newpts((size_t)std::max(B, 1) * T);
        std::vector<int> newlab(std::max(B, 1), 0);
        std::vector<char> ok(std::max(B, 1), 0);
        // (a chain whitens with its seed's cluster's factor -- also when deletions have left a single cluster: that cluster
        // keeps the factor of its own points until the next update)
        const bool clustered = per_cluster && ev.size() > 1, factors = labelled_ev;
        int nfailed = 0;
        for (int k = 0; k < B; ++k) {
            uint64_t uid = (uint64_t)nchains + k;
            double u = rng.uniform(TAG_SEED, uid, 0, 0);
            int sslot, plab = 0;
            if (clustered) {   // GenerateSeed, generate.F90:19-55: a cluster by volume, then one of its live points
                const double rnd = rng.uniform(TAG_SEED, uid, 1, 0);
                plab = (int)ev.size() - 1;
                for (size_t q = 0; q < ev.size(); ++q) if (rnd < cdf[q]) { plab = (int)q; break; }
                const int mp = (int)members[plab].size();
                int choice = (int)std::ceil(u * mp);
                if (choice < 1) choice = 1;
                if (choice > mp) choice = mp;
                sslot = members[plab][choice - 1];
            } else {
                int choice = (int)std::ceil(u * m);
                if (choice < 1) choice = 1;
                sslot = order[K + choice - 1];
                if (labelled_ev) plab = lab[sslot];
            }
            const double* seed = &c.live[(size_t)sslot * T];
            newlab[k] = plab;
            slice_sampling(Lstar, seed, factors ? ev[plab].cholesky.data() : c.cholesky.data(), uid, babies, nlike);
            for (int i = 0; i < R; ++i) babies[(size_t)i * T + b0] = Lstar;
            for (int i = 0; i < R - 1; ++i) {
                const double* pt = babies.data() + (size_t)i * T;
                if (pt[l0] > Lstar) {
                    c.phantom.insert(c.phantom.end(), pt, pt + T);
                    c.nphantom++;
                    if (S.do_clustering) phlab.push_back(plab);
                }
            }
            const double* last = babies.data() + (size_t)(R - 1) * T;
            ok[k] = last[l0] > Lstar;
            std::copy(last, last + T, newpts.begin() + (size_t)k * T);
        }
        // failed births, in chain order
        for (int k = 0; k < B; ++k) {
            if (ok[k]) { failures_in_a_row = 0; continue; }
            ++nfailed; ++nfail_total; ++failures_in_a_row;
            push_dead(&newpts[(size_t)k * T], S.logzero);
        }
        // placement
        const bool labelled = S.do_clustering && (int)lab.size() == n;
        std::vector<int> holes;
        for (int k = 0; k < K; ++k) {
            if (k < B && ok[k]) {
                std::copy(newpts.begin() + (size_t)k * T, newpts.begin() + (size_t)(k + 1) * T, c.live.begin() + (size_t)order[k] * T);
                if (labelled) lab[order[k]] = newlab[k];
                if (labelled_ev) ev[newlab[k]].nlive++;
            } else holes.push_back(order[k]);
        }
        // the live count grows: birth k >= K takes slot n + (k - K); a failed one leaves that slot empty
        const int ext = n + std::max(0, B - K);
        c.live.resize((size_t)ext * T);
        if (labelled) lab.resize(ext);
        for (int k = K; k < B; ++k) {
            if (!ok[k]) { holes.push_back(n + (k - K)); continue; }
            std::copy(newpts.begin() + (size_t)k * T, newpts.begin() + (size_t)(k + 1) * T, c.live.begin() + (size_t)(n + k - K) * T);
            if (labelled) lab[n + k - K] = newlab[k];
            if (labelled_ev) ev[newlab[k]].nlive++;
        }
        // close the empty slots: with n' = extent - holes live points left, the records at or above n' move into the
        // empty slots below it, lowest into lowest (a rule every slot can apply by itself)
        {
            const int n1 = ext - (int)holes.size();
            std::vector<char> empty(ext + 1, 0);
            for (int sl : holes) empty[sl] = 1;
            int src = n1;
            for (int dst = 0; dst < n1; ++dst) {
                if (!empty[dst]) continue;
                while (empty[src]) ++src;
                std::copy(c.live.begin() + (size_t)src * T, c.live.begin() + (size_t)(src + 1) * T, c.live.begin() + (size_t)dst * T);
                if (labelled) lab[dst] = lab[src];
                ++src;
            }
            c.nlive = n1;
        }
        c.live.resize((size_t)c.nlive * T);
        if (labelled) lab.resize(c.nlive);
        nchains += B;
        ngen++;
        find_min();
        return nfailed;
    }

    void run_batched() {
        Cluster& c = cl[0];
        std::vector<double> babies;
        const int nfail = S.nfail <= 0 ? S.nlive : S.nfail;
        // nprior > nlive: the trim of nested_sampling.F90:201-203 -- the excess points die one after the other, no
        // births -- is a generation of its own
        if (c.nlive > S.nlive) {
            batched_generation(c.nlive - S.nlive, 0, babies);
            ngen--;   // not a sampling generation
        }
        while (more_samples_needed() && failures_in_a_row <= nfail) {
            const int n = c.nlive;
            int K = std::min(S.batch_K, n - 1);
            if (S.max_ndead > 0) K = (int)std::min<long long>(K, S.max_ndead - ndead);
            if (K < 1) break;
            // the live count moves towards its target above the contour (dynamic nlive, run_time_info.f90:766-777;
            // constant: settings%nlive), at most 2 batch_K births a generation
            std::vector<int> order(n);
            std::iota(order.begin(), order.end(), 0);
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
                return c.live[(size_t)a * T + l0] < c.live[(size_t)b * T + l0];
            });
            const double Lstar = c.live[(size_t)order[K - 1] * T + l0];
            const int B = std::max(0, std::min(std::max(target_nlive(Lstar), 1) - (n - K), 2 * S.batch_K));
            batched_generation(K, B, babies);
            if (sum_logX() <= logX_last_update + std::log(S.compression_factor)) do_update();
        }
    }

    void final_killoff() {
        if (S.batch_K > 0) {
            // same order as the engine: ascending (logL, slot)
            Cluster& c = cl[0];
            int n = c.nlive;
            std::vector<int> order(n);
            std::iota(order.begin(), order.end(), 0);
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
                return c.live[(size_t)a * T + l0] < c.live[(size_t)b * T + l0];
            });
            for (int k = 0; k < n; ++k) {
                const double* r = &c.live[(size_t)order[k] * T];
                double logw;
                if (pce()) {
                    const int pc = lab[order[k]];
                    ev[pc].logLp = r[l0];
                    logw = update_evidence_in(ev, evXX, pc);
                    ev[pc].nlive--;
                } else {
                    c.logLp = r[l0];
                    const double xx0 = XX(0, 0);
                    logw = update_evidence(0);
                    if (evc()) { attribute_death(lab[order[k]], r[l0], c.nlive, xx0); ev[lab[order[k]]].nlive--; }
                }
                if (evc()) dead_cluster.push_back(ev_uid[lab[order[k]]]);
                push_dead(r, logw);
                c.stack_l.push_back(r[l0]);
                c.stack_i.push_back(ndead - 1);
                c.nlive--;
            }
            c.live.clear();
        } else {
            while (total_live() > 0) delete_outermost_point();
        }
    }

    void run(oracle_result* out) {
        auto t0 = std::chrono::steady_clock::now();
        g_boost_rows.clear(); g_boost_dead.clear(); g_boost_logw.clear();
        init_layout();
        generate_live_points();
        if (evc()) init_ev();
        if (S.batch_K > 0) run_batched();
        else {
            // nprior > nlive: trim (nested_sampling.F90:201-203)
            while (cl[0].nlive > S.nlive) delete_outermost_point();
            run_reference();
        }
        out->ncluster = S.batch_K > 0 ? ncl_max_b : (long long)cl.size();
        out->nsplits = nsplits;
        std::vector<double> logX_end;
        if (evc()) for (auto& c : ev) logX_end.push_back(pce() ? c.logXp : cl[0].logXp + std::log(std::max(c.nlive, 1) / (double)std::max(cl[0].nlive, 1)));
        final_killoff();
        g_last_clusters.clear();
        g_last_nactive = 0;
        g_last_dead_cluster.clear(); g_last_uid_parent.clear(); g_last_cluster_uid.clear();
        if (evc()) {
            g_last_dead_cluster = dead_cluster;
            g_last_uid_parent = uid_parent;
            g_last_cluster_uid = ev_uid;
            g_last_cluster_uid.insert(g_last_cluster_uid.end(), dead_uid.begin(), dead_uid.end());
            g_last_nactive = (int)ev.size();
            for (size_t q = 0; q < ev.size(); ++q) { g_last_clusters.push_back(ev[q].logZp); g_last_clusters.push_back(ev[q].logZp2); g_last_clusters.push_back(logX_end[q]); }
            for (size_t q = 0; q < logZp_dead.size(); ++q) { g_last_clusters.push_back(logZp_dead[q]); g_last_clusters.push_back(logZp2_dead[q]); g_last_clusters.push_back(S.logzero); }
        }
        long long nph_end = 0;   // phantoms at the end of the sampling loop (what the engine reports too)
        for (auto& c : cl) nph_end += c.nphantom;
        if (thin_posterior() > 0.0) {   // update_posteriors, nested_sampling.F90:386: the last phantoms' conversion
            if (S.batch_K > 0) clean_phantoms_stable(); else clean_phantoms();
        }
        dump();
        auto t1 = std::chrono::steady_clock::now();
        double lz, var;
        logZ_estimate(lz, var);
        out->logZ = lz; out->logZerr = std::sqrt(var);
        out->logZ_raw = logZ; out->logZ2_raw = logZ2;
        out->ndead = ndead; out->nlike = nlike; out->nchains = nchains; out->ngenerations = ngen;
        out->nupdates = nupdates; out->nfailures = nfail_total; out->nslices = nslices;
        out->nphantoms_final = nph_end;
        out->seconds = std::chrono::duration<double>(t1 - t0).count();
    }
};

void setup_like_prior(Run& run, int like_kind, const double* like_params, int n_like_params, const double* prior_lo,
                      const double* prior_hi, double (*ll_cb)(double*, int, double*, int),
                      void (*prior_cb)(double*, double*, int)) {
    int D = run.S.nDims;
    run.like.kind = like_kind;
    run.like.log_rast = std::log(4991.21750);
    run.like.Vn = std::pow(std::sqrt(M_PI), (double)D) / std::tgamma(1.0 + D / 2.0);
    if (like_kind == 0) {
        run.like.mu.assign(D, 0.5);
        run.like.sigma.assign(D, 0.1);
        if (like_params && n_like_params >= 2 * D) {
            run.like.mu.assign(like_params, like_params + D);
            run.like.sigma.assign(like_params + D, like_params + 2 * D);
        } else if (like_params && n_like_params == 2) {
            run.like.mu.assign(D, like_params[0]);
            run.like.sigma.assign(D, like_params[1]);
        }
        run.like.gauss_norm = 0.0;
        for (int i = 0; i < D; ++i) run.like.gauss_norm += std::log(run.like.sigma[i]) + LOG_TWO_PI / 2.0;
    } else if (like_kind == 2) {
        // params: mu[D], invcov[D*D] column-major, logdet
        run.like.mu.assign(like_params, like_params + D);
        run.like.invcov.assign(like_params + D, like_params + D + (size_t)D * D);
        run.like.logdet = like_params[D + (size_t)D * D];
    } else if (like_kind == 3) {
        run.like.cb = ll_cb;
    }
    if (prior_cb) {
        run.pri.kind = 1;
        run.pri.cb = prior_cb;
    } else {
        run.pri.kind = 0;
        run.pri.lo.assign(D, 0.0);
        run.pri.hi.assign(D, 1.0);
        if (prior_lo && prior_hi) {
            run.pri.lo.assign(prior_lo, prior_lo + D);
            run.pri.hi.assign(prior_hi, prior_hi + D);
        }
    }
}

}  // namespace

extern "C" {

// Full nested-sampling run.  like_kind: 0 gaussian (params: mu[D],sigma[D] or {mu,sigma} or none),
// 1 rastrigin, 2 correlated gaussian (params: mu[D], invcov[D*D], logdet), 3 host callback.
int oracle_run(const oracle_settings* s, int like_kind, const double* like_params, int n_like_params,
               const double* prior_lo, const double* prior_hi, double (*ll_cb)(double*, int, double*, int),
               void (*prior_cb)(double*, double*, int), oracle_dumper_t dumper, oracle_result* out) {
    Run run;
    run.S = *s;
    run.dumper = dumper;
    setup_like_prior(run, like_kind, like_params, n_like_params, prior_lo, prior_hi, ll_cb, prior_cb);
    run.run(out);
    return 0;
}

// One chain (SliceSampling) from a given seed record: the per-chain parity probe.
// seed_point: T doubles; cholesky: D*D column-major; babies out: R*T doubles (b0 set to logL).
int oracle_slice_chain(const oracle_settings* s, int like_kind, const double* like_params, int n_like_params,
                       const double* prior_lo, const double* prior_hi, const double* seed_point,
                       const double* cholesky, double logL, unsigned long long uid, double* babies_out,
                       long long* nlike_out) {
    Run run;
    run.S = *s;
    setup_like_prior(run, like_kind, like_params, n_like_params, prior_lo, prior_hi, nullptr, nullptr);
    run.init_layout();
    std::vector<double> babies;
    long long n = 0;
    run.slice_sampling(logL, seed_point, cholesky, uid, babies, n);
    for (int i = 0; i < run.R; ++i) babies[(size_t)i * run.T + run.b0] = logL;
    std::copy(babies.begin(), babies.end(), babies_out);
    *nlike_out = n;
    return 0;
}

// Whitening-free direction set of one chain (D x R column-major), after the shuffle.
int oracle_generate_nhats(const oracle_settings* s, unsigned long long uid, double* nhats_out) {
    Run run;
    run.S = *s;
    run.init_layout();
    std::vector<double> nh;
    run.generate_nhats(uid, nh);
    std::copy(nh.begin(), nh.end(), nhats_out);
    return 0;
}

// Evaluate prior+likelihood for npts cube points (records of T doubles, cube in [0,D)).
int oracle_calculate_points(const oracle_settings* s, int like_kind, const double* like_params, int n_like_params,
                            const double* prior_lo, const double* prior_hi, double* records, int npts) {
    Run run;
    run.S = *s;
    setup_like_prior(run, like_kind, like_params, n_like_params, prior_lo, prior_hi, nullptr, nullptr);
    run.init_layout();
    long long n = 0;
    for (int i = 0; i < npts; ++i) run.calculate_point(records + (size_t)i * run.T, n);
    return (int)n;
}

// Dynamic nlive for the following runs in the reference schedule (batch_K = 0): above the contour loglikes[i] the
// target number of live points is nlives[i] (run_time_info.f90:766-777); m = 0 clears.
// evidence clusters of the last batched run with clustering: returns their number (active first, then deleted),
// *nactive the active ones; out (when not null, 3 doubles per cluster): logZp, logZp2, logXp at the end of sampling
// identity of the cluster every dead point of the last batched run with clustering died in (0: the initial cluster; a
// split creates new identities), and for every identity the one it was split from (-1: none).  Return the counts.
long long oracle_last_dead_clusters(int* out, long long cap) {
    if (out) std::copy(g_last_dead_cluster.begin(), g_last_dead_cluster.begin() + std::min<long long>(cap, (long long)g_last_dead_cluster.size()), out);
    return (long long)g_last_dead_cluster.size();
}
int oracle_last_cluster_tree(int* parent_out, int* uid_of_cluster_out) {
    if (parent_out) std::copy(g_last_uid_parent.begin(), g_last_uid_parent.end(), parent_out);
    if (uid_of_cluster_out) std::copy(g_last_cluster_uid.begin(), g_last_cluster_uid.end(), uid_of_cluster_out);
    return (int)g_last_uid_parent.size();
}
int oracle_last_clusters(int* nactive, double* out) {
    if (nactive) *nactive = g_last_nactive;
    if (out) std::copy(g_last_clusters.begin(), g_last_clusters.end(), out);
    return (int)(g_last_clusters.size() / 3);
}

void oracle_set_nlives(const double* loglikes, const int* nlives, int m) {
    g_dyn_loglikes.assign(loglikes, loglikes + m);
    g_dyn_nlives.assign(nlives, nlives + m);
}

// cube_samples: the live points the next oracle_run starts from (npoints x nDims cube coordinates), one-shot.
void oracle_set_initial_cubes(const double* cubes, int npoints, int nDims) {
    g_init_cubes.assign(cubes, cubes + (size_t)npoints * nDims);
}

// The phantoms the last oracle_run promoted to posterior samples (boost_posterior with posteriors or equals set):
// rows[npars] = [theta, phi, birth, logL], dead_index = the dead point whose weight the sample carries,
// logw = that weight + the sample's logL (unnormalised posterior log-weight).  Returns the count; fills up to cap.
long long oracle_last_boosted(double* rows, long long* dead_index, double* logw, long long cap, int npars) {
    const long long nb = (long long)g_boost_dead.size();
    for (long long i = 0; i < std::min(nb, cap); ++i) {
        std::copy(g_boost_rows.begin() + (size_t)i * npars, g_boost_rows.begin() + (size_t)(i + 1) * npars, rows + (size_t)i * npars);
        dead_index[i] = g_boost_dead[(size_t)i];
        logw[i] = g_boost_logw[(size_t)i];
    }
    return nb;
}

// Fast/slow parameter grades for the following runs (settings%grade_dims, RTI%num_repeats per grade,
// generate.F90:303-309); nGrade = 0 clears them.
void oracle_set_grades(int nGrade, const int* grade_dims, const int* grade_repeats) {
    g_grade_dims.assign(grade_dims, grade_dims + nGrade);
    g_grade_repeats.assign(grade_repeats, grade_repeats + nGrade);
}

// NN_clustering (clustering.f90:15-97) of m points (row-major m x D cube coordinates); returns the number of clusters
int oracle_nn_clustering(const double* points, int m, int D, int* labels_out) {
    std::vector<const double*> pts(m);
    for (int i = 0; i < m; ++i) pts[i] = points + (size_t)i * D;
    int num = 1;
    const std::vector<int> l = NN_clustering(similarity_matrix(pts, D), m, num);
    std::copy(l.begin(), l.end(), labels_out);
    return num;
}

void oracle_philox4x32_10(const unsigned* ctr, const unsigned* key, unsigned* out) {
    U4 o = philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
    for (int i = 0; i < 4; ++i) out[i] = o.v[i];
}
double oracle_uniform(unsigned seed, unsigned tag, unsigned long long uid, unsigned a, unsigned b) {
    Rng r{seed};
    return r.uniform(tag, uid, a, b);
}
double oracle_inv_normal_cdf(double p) { return inv_normal_cdf(p); }
double oracle_logaddexp(double a, double b) { return logaddexp(a, b); }
double oracle_logsumexp(const double* v, int n) { return logsumexp(v, (size_t)n); }
void oracle_calc_cholesky(const double* a, double* L, int D) { calc_cholesky(a, L, D); }

// Apply `ndeaths` consecutive single-cluster evidence updates with the given logL values
// and starting live count n0 (decreasing by `dec` per death: 0 = constant nlive, 1 = kill-off).
// state = {logZ, logZ2, logXp, logZXp, logZp, logZp2, logZpXp, logXpXp}; logweights_out[ndeaths].
void oracle_evidence_sequence(double* state, const double* logLs, int ndeaths, int n0, int dec, double logzero,
                              double* logweights_out) {
    Run run;
    std::memset(&run.S, 0, sizeof(run.S));
    run.S.logzero = logzero;
    run.S.nDims = 1;
    run.init_layout();
    run.initialise();
    Cluster& c = run.cl[0];
    run.logZ = state[0]; run.logZ2 = state[1]; c.logXp = state[2]; c.logZXp = state[3];
    c.logZp = state[4]; c.logZp2 = state[5]; c.logZpXp = state[6]; run.XX(0, 0) = state[7];
    int n = n0;
    for (int i = 0; i < ndeaths; ++i) {
        c.nlive = n;
        c.logLp = logLs[i];
        logweights_out[i] = run.update_evidence(0);
        n -= dec;
    }
    state[0] = run.logZ; state[1] = run.logZ2; state[2] = c.logXp; state[3] = c.logZXp;
    state[4] = c.logZp; state[5] = c.logZp2; state[6] = c.logZpXp; state[7] = run.XX(0, 0);
}

// random_inverse_covmat (random_utils.F90:581-614): Haar basis from stream TAG_LIKE.
// out: invcov[D*D] column-major, then logdet.  sigma is passed by the caller (the reference
// uses the single-precision literal 0.1, random_gaussian.f90:5).
void oracle_random_inverse_covmat(unsigned seed, int D, double sigma, double* invcov, double* logdet) {
    Rng rng{seed};
    std::vector<double> Q((size_t)D * D), ev(D);
    random_orthonormal_basis(rng, TAG_LIKE, 0, 0, D, Q.data());
    double ld = 0.0;
    for (int j = 0; j < D; ++j) {
        ev[j] = sigma * std::pow(1e-2, (double)j / (D - 1.0));
        ld += std::log(ev[j]);
    }
    for (int r = 0; r < D; ++r)
        for (int c = 0; c < D; ++c) {
            double s = 0.0;
            for (int j = 0; j < D; ++j) s += Q[r + (size_t)j * D] * (1.0 / (ev[j] * ev[j])) * Q[c + (size_t)j * D];
            invcov[r + (size_t)c * D] = s;
        }
    *logdet = 2.0 * ld;
}

}  // extern "C"
