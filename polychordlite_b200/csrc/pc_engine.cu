// pc_engine.cu -- host side of the B200 nested-sampling engine and its C ABI (libchord.so).
//
// Boundary replaced: src/polychord/interfaces.F90:285-436 (polychord_c_interface) and the
// Fortran core behind it (nested_sampling.F90:15-510).  See include/polychord_b200.h.
// There is NO CPU fallback: without a CUDA device every compute entry point fails loudly.
#include <cuda_runtime.h>
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <mutex>
#include <thread>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/polychord_b200.h"
#include "pc_errors.h"
#include "pc_probes.cuh"
#include "pc_files.h"
#include "pc_resume_text.h"
#include "pc_ini.h"
#include "pc_maximise.h"
#include "pc_hostchain.cuh"
#include "pc_cluster.cuh"
#include "pc_shapes.h"

namespace pc {

#define PC_CUDA(call)                                                                                   \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            throw pc::RunError(std::string("polychord_b200: CUDA error: ") + cudaGetErrorString(e_) + \
                                     " at " + __FILE__ + ":" + std::to_string(__LINE__));              \
    } while (0)

// Caching device allocator: cudaMalloc/cudaFree cost milliseconds and synchronise the device, and a
// sampler is typically called many times with the same shapes, so freed blocks are kept (per device,
// per exact size) and handed out again.  pc_release_memory() returns them to the driver.
struct DevicePool {
    std::mutex mu;
    std::multimap<std::pair<int, size_t>, void*> free_blocks;
    std::map<void*, std::pair<int, size_t>> live_blocks;
    void* get(size_t bytes) {
        int dev = 0;
        cudaGetDevice(&dev);
        bytes = (bytes + 255) & ~(size_t)255;
        std::lock_guard<std::mutex> lk(mu);
        auto it = free_blocks.find({dev, bytes});
        void* p = nullptr;
        if (it != free_blocks.end()) { p = it->second; free_blocks.erase(it); }
        else {
            cudaError_t e = cudaMalloc(&p, bytes);
            if (e != cudaSuccess) {  // give cached blocks back and retry once
                for (auto& kv : free_blocks) cudaFree(kv.second);
                free_blocks.clear();
                e = cudaMalloc(&p, bytes);
            }
            if (e != cudaSuccess)
                throw pc::RunError(std::string("polychord_b200: cudaMalloc failed: ") + cudaGetErrorString(e));
        }
        live_blocks[p] = {dev, bytes};
        return p;
    }
    void put(void* p) {
        std::lock_guard<std::mutex> lk(mu);
        auto it = live_blocks.find(p);
        if (it == live_blocks.end()) return;
        free_blocks.insert({it->second, p});
        live_blocks.erase(it);
    }
    void trim() {
        std::lock_guard<std::mutex> lk(mu);
        for (auto& kv : free_blocks) cudaFree(kv.second);
        free_blocks.clear();
    }
};
static DevicePool& pool() { static DevicePool* p = new DevicePool; return *p; }  // leaked on purpose: outlives the CUDA context teardown

// Pinned host staging buffer that grows on demand and is reused between runs.
struct PinnedBuf {
    void* p = nullptr;
    size_t cap = 0;
    void* need(size_t bytes) {
        if (bytes > cap) {
            if (p) cudaFreeHost(p);
            cap = std::max(bytes, cap * 2);
            PC_CUDA(cudaHostAlloc(&p, cap, cudaHostAllocDefault));
        }
        return p;
    }
};
static PinnedBuf g_pin_dead, g_pin_live;
// Host mirror of the dead points handed to the dumper (rows [theta, phi, birth, logL], posterior log-weights).
// Kept between runs: a sampler is called again and again with the same shapes, and fresh pages cost more
// than the copies.
struct DumpMirror {
    std::vector<double> rows, logw, lw, live_rows;
    // boost_posterior: phantoms promoted to posterior samples (clean_phantoms, run_time_info.f90:846-868), in the
    // order (update, logL): rows [theta, phi, birth, logL], posterior log-weight (log w of its death + own logL), the
    // dead point whose weight it carries, and the number of dead points at the update that removed it
    std::vector<double> boost_rows, boost_logw;
    std::vector<long long> boost_dead, boost_after;
};
static DumpMirror g_mirror;
// The mirror's rows, page-locked in place (cudaHostRegister) while an asynchronous dumper hand-over is in use: the new dead
// rows of a dump then land in the mirror by DMA instead of in a staging buffer that the host copies from (at nlive 8000
// the copy was a third of the host's work per dump, and the host's work per dump is what bounds the sharded runs end to
// end).  Sized before the launch (registering, like allocating pinned memory, synchronises with a running kernel).
struct MirrorPin {
    void* p = nullptr;
    size_t bytes = 0;
    void drop() { if (p) cudaHostUnregister(p); p = nullptr; bytes = 0; }
    // rows: capacity wanted (doubles); keeps the contents
    void reserve(std::vector<double>& v, size_t doubles) {
        if (v.size() >= doubles && p == (void*)v.data() && bytes >= doubles * sizeof(double)) return;
        drop();
        if (v.size() < doubles) v.resize(std::max(doubles, v.size() * 2));
        if (cudaHostRegister(v.data(), v.size() * sizeof(double), cudaHostRegisterDefault) == cudaSuccess) { p = v.data(); bytes = v.size() * sizeof(double); }
        else cudaGetLastError();   // not fatal: the staged path stays
    }
    bool covers(const std::vector<double>& v, size_t doubles) const { return p && p == (const void*)v.data() && v.size() >= doubles && bytes >= doubles * sizeof(double); }
};
static MirrorPin g_mirror_pin;
// `maximise`: the live points the sampling loop ended with (full records, cube coordinates included) -- the simplex
// of the maximiser is built from them (maximiser.F90:117-135).  Filled by Engine::run when want is set.
struct FinalLive { bool want = false; int n = 0; std::vector<double> recs; };
static FinalLive g_final_live;
static int sm_clock_khz() {  // cudaDevAttrClockRate is a slow driver query (milliseconds): ask once per device
    static std::map<int, int> cache;
    int dev = 0;
    cudaGetDevice(&dev);
    auto it = cache.find(dev);
    if (it != cache.end()) return it->second;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    cache[dev] = khz;
    return khz;
}
static HostCtl* g_ctl = nullptr;          // mapped pinned control block of the asynchronous dumper hand-over
static cudaStream_t g_copy_stream = nullptr;
static HostCtl* host_ctl() {
    if (!g_ctl) {
        PC_CUDA(cudaHostAlloc((void**)&g_ctl, sizeof(HostCtl), cudaHostAllocMapped | cudaHostAllocPortable));
        PC_CUDA(cudaStreamCreateWithFlags(&g_copy_stream, cudaStreamNonBlocking));
    }
    return g_ctl;
}

// Sharded run across the GPUs of one box (one process per GPU): the exchange block of every rank, mapped
// into this process through CUDA IPC.  Layout of a block: [barrier counter, 256 B][incoming last babies,
// 2 x batch_K x T doubles][statistics slots, world x xstride doubles].
struct MgpuCtx {
    int rank = 0, world = 0;
    void* local = nullptr;
    void* peer[MAX_RANKS] = {nullptr};
    size_t bytes = 0;
    int batch_K = 0, T = 0, D = 0;
    unsigned long long epoch = 0;  // cross-GPU barriers passed so far (the counters are never reset)
};
static MgpuCtx g_mgpu;
// clusters of the last run with do_clustering (pc_last_clusters): per cluster {log<Z_p>, log<Z_p^2>} and its identity, the
// clusters alive at the end of sampling first, then the deleted ones; per identity its parent; per dead point its cluster
struct ClusterReport {
    int nactive = 0;
    std::vector<double> rows;
    std::vector<int> uid, parent, point_uid;
    std::vector<double> frac;   // per identity: log share of its parent's evidence it received at the split
    std::vector<double> zp, zp2; // rows, one column each (what the file writer takes)
};
static ClusterReport g_cluster_report;
static size_t mgpu_xstride(int D) { return (size_t)((2 + D + D * (D + 1) / 2 + 1) & ~1); }
static size_t mgpu_kr(int batch_K, int world) { return (size_t)(batch_K + world - 1) / world; }
// [barrier][incoming last babies][statistics slots][sorted (logL, slot) runs of every rank's babies, 2 x world x kr pairs]
static size_t mgpu_block_bytes(int batch_K, int T, int D, int world) {
    return 256 + (size_t)2 * batch_K * T * 8 + (size_t)world * mgpu_xstride(D) * 8 + (size_t)2 * world * mgpu_kr(batch_K, world) * 16;
}

template <class V>
struct DevArr {  // RAII device buffer (exception-transparent: callbacks may throw through the engine)
    V* p = nullptr;
    size_t n = 0;
    DevArr() {}
    explicit DevArr(size_t n_) { alloc(n_); }
    DevArr(const DevArr&) = delete;
    DevArr& operator=(const DevArr&) = delete;
    DevArr(DevArr&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevArr& operator=(DevArr&& o) noexcept { if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; } return *this; }
    ~DevArr() { release(); }
    void release() { if (p) pool().put(p); p = nullptr; n = 0; }
    void alloc(size_t n_) { release(); n = n_; if (n) p = (V*)pool().get(n * sizeof(V)); }
    void zero(cudaStream_t s) { if (n) PC_CUDA(cudaMemsetAsync(p, 0, n * sizeof(V), s)); }
    void upload(const V* h, size_t cnt, cudaStream_t s) { PC_CUDA(cudaMemcpyAsync(p, h, cnt * sizeof(V), cudaMemcpyHostToDevice, s)); }
    void download(V* h, size_t cnt, cudaStream_t s, size_t off = 0) const { PC_CUDA(cudaMemcpyAsync(h, p + off, cnt * sizeof(V), cudaMemcpyDeviceToHost, s)); }
    // grow keeping the first `keep` elements
    void grow(size_t n_, size_t keep, cudaStream_t s) {
        V* q = (V*)pool().get(n_ * sizeof(V));
        if (keep) PC_CUDA(cudaMemcpyAsync(q, p, keep * sizeof(V), cudaMemcpyDeviceToDevice, s));
        PC_CUDA(cudaStreamSynchronize(s));
        if (p) pool().put(p);
        p = q; n = n_;
    }
};

struct Options {
    // K lowest points die per generation: K = round(nlive * batch_fraction).  0 = automatic (batch_size()): about 1/2 for
    // a run that has the device to itself (its wall time is the number of generations: per unit of compression K = n/2
    // needs 2.4 times fewer than K = n/4 for 1.25 times the variance), 1/4 for the runs of an ensemble (the device is
    // full either way, so the smaller variance per likelihood evaluation wins).  DESIGN.md section 2.
    double batch_fraction = 0.0;
    int dense = 0;      // the dense chain phase (pc_dense.cuh): 0 = for ensembles, 1 = always (testing), -1 = never
    int batch_K = 0;
    int device = 0;
    int warps_per_cta = 8;
    int max_ctas = 0;
    int errors_return = 0;
    int nh_global = 0;  // force the direction scratch into global memory (testing)
    int no_pairing = 0; // keep the helper-warp preparation off (testing)
    int no_phase_d = 0; // order the live points on CTA 0 only (testing: phase D off)
    int no_bulk = 0;    // phase U streams its records through registers (testing: the bulk-copy ring off)
    int no_narrow = 0;  // the chains store and phase U carries a phantom's theta even when nothing will read it (testing)
    int no_wave_batch = 0;  // automatic batch size: plain nlive/2 for a run alone (the wave rule of batch_size() off)
    int resume_text = 0; // write_resume writes the reference's text layout (read_write.F90:219-288) instead of the engine's binary one
    int sync_dump = 0;  // the kernel exits at every update for the dumper instead of handing dumps over while running
    long long cap_dead0 = 0, cap_ph0 = 0;  // initial pool capacities in records (0 = automatic)
};
static Options g_opt;
static volatile int g_abort = 0;                  // pc_request_abort(): a host callback asks the run in flight to stop
static std::vector<int> g_grade_dims, g_grade_reps;   // fast/slow grades of the following runs (pc_set_grades)
// dynamic nlive of the following runs (pc_set_nlives): above the contour loglikes[i] the run keeps nlives[i] live points
static std::vector<double> g_dyn_loglikes;
static std::vector<int> g_dyn_nlives;
// cube_samples (polychord.py:576-579, _make_resume_file :650-789): the caller's initial live points, cube coordinates,
// consumed by the next run through polychord_c_interface (one-shot, pc_set_initial_live)
static std::vector<double> g_init_cubes;
static int g_init_n = 0, g_init_D = 0;
static pc_loglikelihood_t g_init_ll = nullptr;    // the run's callbacks, to evaluate those points on the calling thread
static pc_prior_t g_init_prior = nullptr;
static pc_loglikelihood_t g_host_ll = nullptr;   // host-callback run in flight (PC_LIKE_HOST)
static pc_prior_t g_host_prior = nullptr;
struct ResumeOpts {   // <base_dir>/<file_root>.resume
    bool write = false, read = false;
    std::string path;
    double min_interval_s = 1.0;   // rewrites at updates are rate-limited; the one after the final kill-off always happens
};
static ResumeOpts g_resume;
static FileOpts g_files;    // output files of the run in flight (set by polychord_c_interface / pc_set_output)
static FileState g_fstate;
static double g_dbg_dump[4];  // PC_DEBUG: ms spent in copies, row packing, weight normalisation, the user's dumper
static cudaStream_t g_stream = nullptr;
static pc_run_info g_last;
static std::mutex g_mu;

struct ModelSpec {
    int like_kind = 0;
    std::vector<double> like_params;   // raw API params
    std::vector<double> prior_params;  // lo[D], hi[D] or empty
};

struct DevModel {
    DevArr<double> like, prior;
    double gauss_norm = 0, Vn = 0, log_rast = 0, corr_const = 0;
};

static const double LOG_TWO_PI = 1.8378770664093454835606594728112;

static void build_dev_model(const pc_settings& s, const ModelSpec& ms, DevModel& dm, cudaStream_t st) {
    const int D = s.nDims;
    std::vector<double> lp;
    dm.log_rast = std::log(4991.21750);                                                    // rastrigin.f90:33
    dm.Vn = std::pow(std::sqrt(M_PI), (double)D) / std::tgamma(1.0 + D / 2.0);           // utils.F90:754-760
    if (ms.like_kind == PC_LIKE_HOST) {
        lp.assign(2 * D, 1.0);  // never evaluated: the run kernel leaves before its chain phase (pc_hostchain.cuh)
    } else if (ms.like_kind == PC_LIKE_GAUSSIAN) {
        std::vector<double> mu(D, 0.5), sg(D, 0.1);                                        // gaussian.f90:20-21
        const auto& q = ms.like_params;
        if ((int)q.size() >= 2 * D) { mu.assign(q.begin(), q.begin() + D); sg.assign(q.begin() + D, q.begin() + 2 * D); }
        else if (q.size() == 2) { mu.assign(D, q[0]); sg.assign(D, q[1]); }
        else if (!q.empty()) throw pc::ArgError("polychord_b200: gaussian likelihood wants 0, 2 or 2*nDims params");
        dm.gauss_norm = 0.0;
        for (int i = 0; i < D; ++i) dm.gauss_norm += std::log(sg[i]) + LOG_TWO_PI / 2.0;
        lp = mu;
        for (int i = 0; i < D; ++i) lp.push_back(1.0 / sg[i]);
    } else if (ms.like_kind == PC_LIKE_CORR_GAUSSIAN) {
        if ((int)ms.like_params.size() != D + D * D + 1)
            throw pc::ArgError("polychord_b200: correlated gaussian wants mu[D], invcov[D*D], logdet");
        lp.assign(ms.like_params.begin(), ms.like_params.begin() + D + D * D);
        dm.corr_const = -(D * LOG_TWO_PI + ms.like_params[D + D * D]) / 2.0;              // utils.F90:1040-1046
    } else if (ms.like_kind != PC_LIKE_RASTRIGIN) {
        throw pc::ArgError("polychord_b200: unknown device likelihood kind");
    }
    if (lp.empty()) lp.push_back(0.0);
    dm.like.alloc(lp.size());
    dm.like.upload(lp.data(), lp.size(), st);
    std::vector<double> pr(2 * D);
    for (int i = 0; i < D; ++i) { pr[i] = 0.0; pr[D + i] = 1.0; }
    if ((int)ms.prior_params.size() == 2 * D)
        for (int i = 0; i < D; ++i) { pr[i] = ms.prior_params[i]; pr[D + i] = ms.prior_params[D + i] - ms.prior_params[i]; }
    else if (!ms.prior_params.empty())
        throw pc::ArgError("polychord_b200: uniform prior wants lo[D], hi[D]");
    dm.prior.alloc(pr.size());
    dm.prior.upload(pr.data(), pr.size(), st);
    PC_CUDA(cudaStreamSynchronize(st));
}

struct Layout {
    KParams kp;
    size_t smem = 0;
    ShapeFns fn;
    int W = 8;
};

static int device_check() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
        throw pc::RunError("polychord_b200: no CUDA device available -- this engine has no CPU fallback");
    PC_CUDA(cudaSetDevice(g_opt.device));
    return n;
}

// (G, DPL): G lanes share one trial point, DPL dimensions per lane; G*DPL >= nDims
static ShapeFns pick_shape(int D, int kind) {
#define PC_SHAPE(G, DPL)                                                                                     \
    return kind == PC_LIKE_GAUSSIAN ? shape_fns_##G##_##DPL##_0()                                           \
                                    : (kind == PC_LIKE_RASTRIGIN ? shape_fns_##G##_##DPL##_1() : shape_fns_##G##_##DPL##_2())
    if (D <= 8) { PC_SHAPE(4, 2); }
    if (D <= 16) { PC_SHAPE(4, 4); }
    if (D <= 20) { PC_SHAPE(4, 5); }
    if (D <= 32) { PC_SHAPE(4, 8); }
    if (D <= 64) { PC_SHAPE(8, 8); }
    PC_SHAPE(16, 8);
#undef PC_SHAPE
}

static Layout make_layout_w(const pc_settings& s, const ModelSpec& ms, const DevModel& dm, int W, bool alone, bool dense);

// points that die per generation (see Options::batch_fraction).  A run alone on the device (or sharded over `world`
// devices) takes about half the live points per generation -- the critical path of a run is its generation count -- and,
// where that fits, a whole number of WAVES of chains: one wave is a chain warp and its helper on every SM sub-partition
// of every chain CTA ((SMs - 1) * 4 chains per device), and a generation of 500 chains costs the device as much time as
// one of 588.  So K = m * wave for the smallest m with K >= nlive/2, if that keeps K <= 0.6 nlive (variance per unit of
// log-compression 1.61 against 1.44 at nlive/2; the error bar carries it), else nlive/2.  `half_only`: the launch turned
// out not to have the paired geometry (fewer than 8 warps per CTA fit), plain nlive/2.
static bool g_half_only = false;
static int wave_chains(int world) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return std::max(1, (sms - 1) * 4) * std::max(1, world);
}
static int batch_size(int nlive, bool alone, int world = 1) {
    if (g_opt.batch_K > 0) return std::max(1, std::min(g_opt.batch_K, nlive - 1));
    if (g_opt.batch_fraction > 0.0) return std::max(1, std::min((int)std::lround(nlive * g_opt.batch_fraction), nlive - 1));
    int K = (int)std::lround(nlive * (alone ? 0.5 : 0.25));
    if (alone && !g_half_only && !g_opt.no_wave_batch) {
        const long long wave = wave_chains(world);
        const long long m = (K + wave - 1) / wave;
        if (m * wave <= (long long)(0.6 * nlive)) K = (int)(m * wave);
    }
    return std::max(1, std::min(K, nlive - 1));
}

// the largest warp count per CTA (<= the requested one) whose scratch fits in shared memory; the dense chain phase
// wants two CTAs per SM (half the shared memory each) and falls back to the warp-per-chain phase when even four warps
// do not fit
static Layout make_layout(const pc_settings& s, const ModelSpec& ms, const DevModel& dm, int W, bool alone = true, bool dense = false) {
    if (dense) {
        for (int w = W; w >= 4; --w) {
            Layout L = make_layout_w(s, ms, dm, w, alone, true);
            if (L.kp.nh_in_smem && L.smem <= 112 * 1024) return L;
        }
    }
    struct HalfGuard { ~HalfGuard() { g_half_only = false; } } half_guard;
    for (;; --W) {
        Layout L = make_layout_w(s, ms, dm, W, alone, false);
        if (L.smem <= 227 * 1024) {
            // the wave rule of batch_size() assumes the paired geometry (an even number of warps per CTA): without it, nlive/2
            if (alone && !g_half_only && !g_mgpu.local && (L.W % 2) != 0) { g_half_only = true; L = make_layout_w(s, ms, dm, W, alone, false); }
            return L;
        }
        if (W == 1) throw pc::ArgError("polychord_b200: the run does not fit in shared memory (nlive too large for the in-kernel sort, or nDims*nDims too large)");
    }
}

static Layout make_layout_w(const pc_settings& s, const ModelSpec& ms, const DevModel& dm, int W, bool alone, bool dense) {
    Layout L;
    KParams& k = L.kp;
    std::memset(&k, 0, sizeof(k));
    const int D = s.nDims, P = s.nDerived, R = s.num_repeats;
    if (D < 1 || D > 128) throw pc::ArgError("polychord_b200: nDims must be in 1..128");
    if (R < 1) throw pc::ArgError("polychord_b200: num_repeats must be >= 1");  // settings.f90:216
    if (s.nlive < 2) throw pc::ArgError("polychord_b200: nlive must be >= 2");
    L.fn = pick_shape(D, ms.like_kind == PC_LIKE_HOST ? PC_LIKE_GAUSSIAN : ms.like_kind);
    k.host_like = ms.like_kind == PC_LIKE_HOST ? 1 : 0;
    k.live_given = k.host_like;
    const int npt = 32 / L.fn.G;
    k.cp.D = D; k.cp.P = P; k.cp.T = 2 * D + P + 2; k.cp.R = R;
    k.cp.ngrade = 1;
    if (g_grade_dims.size() > 1) {  // pc_set_grades: the dimensions and the repeats must add up to this run's
        int sd = 0, sr = 0;
        for (int v : g_grade_dims) sd += v;
        for (int v : g_grade_reps) sr += v;
        if (sd != D || sr != R || g_grade_dims.size() > MAX_GRADES)
            throw pc::ArgError("polychord_b200: grade_dims must sum to nDims, the repeats per grade to num_repeats (at most 8 grades)");
        k.cp.ngrade = (int)g_grade_dims.size();
        for (int g = 0; g < k.cp.ngrade; ++g) { k.cp.gdims[g] = g_grade_dims[g]; k.cp.greps[g] = g_grade_reps[g]; }
    }
    k.cp.LD = (L.fn.G * L.fn.DPL) | 1;  // odd (bank-conflict free), zero-padded to the lanes' G*DPL dimensions
    k.cp.like_kind = ms.like_kind == PC_LIKE_HOST ? PC_LIKE_GAUSSIAN : ms.like_kind;
    k.cp.logzero = s.logzero;
    k.cp.gauss_norm = dm.gauss_norm; k.cp.Vn = dm.Vn; k.cp.log_rast = dm.log_rast; k.cp.corr_const = dm.corr_const;
    k.n = s.nlive;
    k.nfail = s.nfail;
    // live points the run starts with: nprior draws when nprior > nlive (generate.F90:142-153; the excess dies first,
    // nested_sampling.F90:201-203), the caller's cube_samples when given (any number: the count then moves to nlive)
    k.n0 = g_init_n > 0 ? g_init_n : std::max(s.nlive, s.nprior);
    if (g_dyn_loglikes.size() > MAX_DYN) throw pc::ArgError("polychord_b200: at most 16 entries in the nlives schedule");
    k.dyn_m = (int)g_dyn_loglikes.size();
    k.nmax = std::max(k.n, k.n0);
    for (int i = 0; i < k.dyn_m; ++i) {
        if (g_dyn_nlives[i] < 1) throw pc::ArgError("polychord_b200: the nlives schedule must hold positive counts");
        k.dyn_loglikes[i] = g_dyn_loglikes[i]; k.dyn_nlives[i] = g_dyn_nlives[i];
        k.nmax = std::max(k.nmax, g_dyn_nlives[i]);
    }
    if (k.n0 < 2) throw pc::ArgError("polychord_b200: a run needs at least two live points");
    k.use_prec = s.precision_criterion > 0.0;
    k.max_ndead = s.max_ndead;
    k.log_prec = k.use_prec ? std::log(s.precision_criterion) : 0.0;
    k.log_comp = std::log(s.compression_factor);
    k.like_params = dm.like.p;
    k.prior_params = dm.prior.p;
    k.warps_per_cta = W;
    k.ntri = D * (D + 1) / 2;
    const int Dp8 = (D + 1 + 7) & ~7, nt8 = Dp8 / 8;  // phase U: augmented coordinate rows in 8-wide tiles (pc_run_kernel.cuh)
    k.cov_passes = (nt8 * (nt8 + 1) / 2 + COV_TPP - 1) / COV_TPP;
    k.partial_stride = 1 + D + k.ntri;
    L.W = W;
    const int nlp = (ms.like_kind == PC_LIKE_GAUSSIAN || ms.like_kind == PC_LIKE_HOST) ? 2 * D : (ms.like_kind == PC_LIKE_CORR_GAUSSIAN ? D + D * D : 0);
    const int Dpad = (D + 1) & ~1;
    size_t off = (size_t)D * D * 8;
    k.off_like = (int)off;
    off += (size_t)((nlp + 1) & ~1) * 8;
    off += 64 * sizeof(int);  // s_cnt
    off = (off + 15) & ~(size_t)15;   // the per-warp areas start 16-byte aligned (cp.async staging of the dense chain phase)
    k.off_warp = (int)off;
    // phase U: pivot + a staged batch of augmented rows per warp; all warps' areas together hold the CTA's moment matrix
    const size_t cov_bytes = std::max((size_t)(Dpad + U_BATCH * (Dp8 + 4)) * 8,
                                      ((size_t)(Dpad + Dp8 * Dp8) * 8 + W - 1) / W);
    const int K = (g_mgpu.local && alone) ? g_mgpu.batch_K : batch_size(s.nlive, alone);   // a sharded run: what pc_mgpu_create fixed
    k.batch_K = K;
    const size_t sort_bytes = std::max(smem_S_bytes(k.nmax, K), (size_t)64 * 8 + (size_t)(D * D + Dpad) * 8);  // phase S, or the covariance (+ mean shift) in finish_update
    const size_t budget = dense ? 108 * 1024 : 200 * 1024;
    const int slb = dense ? dense_slb(L.fn.G * L.fn.DPL) : 0;   // dense chain phase: slice records staged per point group
    k.dense = dense ? 1 : 0;
    const size_t with_nh = chain_scratch_bytes(D, R, k.cp.LD, true, ms.like_kind, npt, slb);
    const bool in_smem = (!g_opt.nh_global || dense) && (off + (size_t)W * std::max(with_nh, cov_bytes) <= budget);
    k.nh_in_smem = in_smem ? 1 : 0;
    size_t wb = std::max(chain_scratch_bytes(D, R, k.cp.LD, in_smem, ms.like_kind, npt, slb), cov_bytes);
    wb = (wb + 15) & ~(size_t)15;
    // phase U's record stream (bulk copies into a per-warp ring of 2 * U_BATCH records behind the staged rows): taken when
    // the ring fits into the per-warp area as it is
    k.u_bulk = (!g_opt.no_bulk && (size_t)(Dpad + U_BATCH * (Dp8 + 4) + 2 * U_BATCH * k.cp.T) * 8 <= wb) ? 1 : 0;
    k.warp_bytes = (int)wb;
    L.smem = off + std::max((size_t)W * wb, sort_bytes);
    // phase D (pc_run_kernel.cuh): every CTA stages the nmax live keys behind its per-warp areas; when they do not fit,
    // CTA 0's phase S keeps ordering the live points
    k.off_dkeys = 0;
    {
        const size_t d_off = (off + (size_t)W * wb + 15) & ~(size_t)15;
        const size_t d_end = d_off + ((size_t)k.nmax + 2 * W + 2) * 8;
        if (!g_opt.no_phase_d && d_end <= (dense ? (size_t)112 * 1024 : (size_t)227 * 1024)) {
            k.off_dkeys = (int)d_off;
            L.smem = std::max(L.smem, d_end);
        }
    }
    return L;
}

static void set_smem(const ShapeFns& fn, size_t smem) {
    PC_CUDA(cudaFuncSetAttribute(fn.run, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (fn.run_dense) {
        PC_CUDA(cudaFuncSetAttribute(fn.run_dense, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PC_CUDA(cudaFuncSetAttribute(fn.slice_dense, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    PC_CUDA(cudaFuncSetAttribute(fn.slice, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PC_CUDA(cudaFuncSetAttribute(fn.calc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
}

// ------------------------------------------------------------------------------------------
// One ensemble of runs advanced by the persistent kernel.
// ------------------------------------------------------------------------------------------
struct HostRun {
    DevArr<DevRun> st;
    DevArr<double> live, live_snap, dead, logw, ph0, ph1, chol, cov, partial, nh, gsum, okey, bkey, dpart;
    DevArr<int> order, lab, phl0, phl1, cfail, deadlab, deadn;
    DevArr<double> cchol;
    DevArr<long long> pcount;
    DevArr<unsigned int> pmask;
    DevArr<double> boost;
    DevArr<unsigned long long> boost_win;
    long long boost_mirrored = 0;   // rows of the device list already in the mirror
    RunBuf buf;
    DevRun host_st;
    // dumper mirror
    long long mirrored = 0;
    double lse_max = -std::numeric_limits<double>::infinity(), lse_sum = 0.0;  // running logsumexp of logw + logL
};

// Resume file.  The reference dumps its run_time_info as text (read_write.F90:219-476) so that the same program can
// pick the run up again; nothing else reads that file, so this engine writes ITS OWN state in its own (binary) layout:
// a header that pins the run's shape (the reference checks nDims / nDerived / grades the same way, :402-417), the
// DevRun scalars, and every array a generation reads.  Written to <root>_temp.resume and renamed (read_write.F90:107).
struct ResumeHeader {
    char magic[8];
    int version, sizeof_devrun;
    int D, P, n, nmax, R, batch_K, like_kind, clustering, ngrade;
    int gdims[MAX_GRADES], greps[MAX_GRADES];
    unsigned seed;
    int pad;
    double logzero;
};
struct ResumeData {
    ResumeHeader h;
    DevRun st;
    std::vector<double> live, okey, dead, logw, ph, chol, cov, gsum, cchol, boost;
    std::vector<int> order, lab, phl;
    std::vector<unsigned long long> boost_win;
};
static const char RESUME_MAGIC[8] = {'P', 'C', 'B', '2', '0', '0', 'R', '2'};
template <class V>
static void put_vec(FILE* f, const std::vector<V>& v) {
    const unsigned long long n = v.size();
    std::fwrite(&n, sizeof(n), 1, f);
    if (n) std::fwrite(v.data(), sizeof(V), n, f);
}
template <class V>
static bool get_vec(FILE* f, std::vector<V>& v) {
    unsigned long long n = 0;
    if (std::fread(&n, sizeof(n), 1, f) != 1 || n > (1ull << 34)) return false;
    v.resize(n);
    return n == 0 || std::fread(v.data(), sizeof(V), n, f) == n;
}
static bool read_resume_file(const std::string& path, ResumeData& rd) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    bool ok = std::fread(&rd.h, sizeof(rd.h), 1, f) == 1 && std::memcmp(rd.h.magic, RESUME_MAGIC, 8) == 0 &&
              rd.h.version == 1 && rd.h.sizeof_devrun == (int)sizeof(DevRun) && std::fread(&rd.st, sizeof(DevRun), 1, f) == 1;
    ok = ok && get_vec(f, rd.live) && get_vec(f, rd.order) && get_vec(f, rd.okey) && get_vec(f, rd.dead) && get_vec(f, rd.logw) &&
         get_vec(f, rd.ph) && get_vec(f, rd.chol) && get_vec(f, rd.cov) && get_vec(f, rd.gsum) && get_vec(f, rd.lab) &&
         get_vec(f, rd.phl) && get_vec(f, rd.cchol) && get_vec(f, rd.boost) && get_vec(f, rd.boost_win);
    std::fclose(f);
    if (!ok) throw pc::ArgError("polychord_b200: " + path + " is not a resume file of this engine version");
    return true;
}

struct Engine {
    pc_settings S;
    ModelSpec ms;
    DevModel dm;
    Layout L;
    cudaStream_t stream;
    int nruns = 0, G = 1;
    std::vector<HostRun> runs;
    DevArr<RunBuf> d_bufs;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // resume (row f3)
    bool resumed = false, resumed_finished = false;
    ResumeData rd;
    std::chrono::steady_clock::time_point resume_last;
    bool resume_written = false;
    // host-callback runs (pc_hostchain.cuh)
    bool host_like = false;
    bool given_live = false;   // the initial live points are the caller's cube_samples
    DevArr<unsigned char> hc_scratch;
    DevArr<double> hc_x;
    DevArr<HcChain> hc_ch;
    double *hc_out = nullptr, *hc_in = nullptr;   // mapped pinned host memory
    long long hc_rounds = 0;
    double device_ms = 0;
    int launches = 0;
    long long h2d = 0, d2h = 0;

    ~Engine() {
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (hc_out) cudaFreeHost(hc_out);
        if (hc_in) cudaFreeHost(hc_in);
    }

    // GenerateLivePoints (generate.F90:153-183) with host callbacks: attempt a draws cube = U(TAG_INIT, a, dim) on the
    // device -- the same counter-addressed numbers the device path uses -- and is kept when logL > logzero.
    // cube_samples: the initial live points are the caller's (polychord.py:650-789 writes them into a resume file the
    // Fortran then reads; here they are evaluated through the run's callbacks and uploaded).  Born from the prior
    // (birth contour logzero); a point the likelihood excludes is fatal.
    void host_given_live_points() {
        const KParams& k = L.kp;
        const int D = k.cp.D, P = k.cp.P, T = k.cp.T, n = k.n0;
        std::vector<double> live((size_t)n * T, 0.0), cube(D), theta(D), phi(std::max(P, 1));
        for (int j = 0; j < n; ++j) {
            double* rec = &live[(size_t)j * T];
            std::copy(g_init_cubes.begin() + (size_t)j * D, g_init_cubes.begin() + (size_t)(j + 1) * D, rec);
            for (int i = 0; i < D; ++i)
                if (!(rec[i] >= 0.0 && rec[i] <= 1.0)) throw pc::ArgError("polychord_b200: cube_samples must lie in the unit hypercube");
            std::copy(rec, rec + D, cube.begin());
            std::fill(phi.begin(), phi.end(), 0.0);
            g_init_prior(cube.data(), theta.data(), D);
            const double logL = g_init_ll(theta.data(), D, phi.data(), P);
            if (!(logL > S.logzero)) throw pc::ArgError("polychord_b200: a cube_samples point has loglikelihood <= logzero");
            std::copy(theta.begin(), theta.end(), rec + D);
            for (int i = 0; i < P; ++i) rec[2 * D + i] = phi[i];
            rec[2 * D + P] = S.logzero;
            rec[2 * D + P + 1] = logL;
        }
        DevRun h0;
        std::memset(&h0, 0, sizeof(h0));
        h0.nlike = n;
        h0.init_attempts = n;
        h0.n = n;
        runs[0].live.upload(live.data(), live.size(), stream);
        runs[0].st.upload(&h0, 1, stream);
        h2d += (long long)live.size() * 8;
        PC_CUDA(cudaStreamSynchronize(stream));
    }

    void host_generate_live_points() {
        const KParams& k = L.kp;
        const int D = k.cp.D, P = k.cp.P, T = k.cp.T, n = k.n0;
        std::vector<double> live((size_t)n * T, 0.0), cube(D), theta(D), phi(std::max(P, 1));
        DevRun h0;
        std::memset(&h0, 0, sizeof(h0));
        int have = 0;
        long long a = 0, nl = 0;
        const unsigned seed = runs[0].buf.seed;
        DevArr<double> d_cubes((size_t)n * D);
        std::vector<double> cubes((size_t)n * D);
        while (have < n) {
            if (a > 1000LL * n + 1000000LL) throw pc::RunError("polychord_b200: could not generate live points (likelihood <= logzero everywhere?)");
            const int count = n - have;   // as many attempts as points are still missing, drawn on the device
            hc_init_cubes_kernel<<<(count * D + 255) / 256, 256, 0, stream>>>(seed, a, count, D, d_cubes.p);
            PC_CUDA(cudaGetLastError());
            d_cubes.download(cubes.data(), (size_t)count * D, stream);
            PC_CUDA(cudaStreamSynchronize(stream));
            launches += 1;
            for (int t = 0; t < count; ++t) {
                std::copy(cubes.begin() + (size_t)t * D, cubes.begin() + (size_t)(t + 1) * D, cube.begin());
                std::vector<double> c2(cube);
                std::fill(phi.begin(), phi.end(), 0.0);
                g_host_prior(c2.data(), theta.data(), D);
                const double logL = g_host_ll(theta.data(), D, phi.data(), P);
                if (!(logL > S.logzero)) continue;
                ++nl;
                double* rec = &live[(size_t)have * T];
                std::copy(cube.begin(), cube.end(), rec);
                std::copy(theta.begin(), theta.end(), rec + D);
                for (int i = 0; i < P; ++i) rec[2 * D + i] = phi[i];
                rec[2 * D + P] = S.logzero;  // born from the prior
                rec[2 * D + P + 1] = logL;
                ++have;
            }
            a += count;
        }
        h0.nlike = nl;
        h0.init_attempts = a;
        h0.n = n;
        runs[0].live.upload(live.data(), live.size(), stream);
        runs[0].st.upload(&h0, 1, stream);
        h2d += (long long)live.size() * 8;
        PC_CUDA(cudaStreamSynchronize(stream));
    }

    // The chains of one generation in lock step: every round the device emits one trial point per running chain
    // and the calling thread makes the prior + likelihood calls (calculate_point, calculate.f90:6-50).
    void host_chains() {
        const KParams& k = L.kp;
        const int D = k.cp.D, P = k.cp.P, K = runs[0].host_st.B, Kdead = runs[0].host_st.K;   // births (chains), deaths
        HcParams hp;
        std::memset(&hp, 0, sizeof(hp));
        hp.D = D; hp.P = P; hp.T = k.cp.T; hp.R = k.cp.R; hp.LD = k.cp.LD; hp.n = runs[0].host_st.n_gen; hp.K = K; hp.Kdead = Kdead; hp.cp = k.cp;
        hp.logzero = S.logzero; hp.seed = runs[0].buf.seed; hp.rb = runs[0].buf;
        hp.scratch_bytes = chain_scratch_bytes(D, k.cp.R, k.cp.LD, true, LIKE_GAUSSIAN, 1);
        if (!hc_out) {
            const size_t Kmax = (size_t)2 * k.batch_K;   // births of a generation (phase S1)
            hc_scratch.alloc(Kmax * hp.scratch_bytes);
            hc_x.alloc(Kmax * k.cp.LD);
            hc_ch.alloc(Kmax);
            PC_CUDA(cudaHostAlloc((void**)&hc_out, Kmax * (D + 2) * 8, cudaHostAllocMapped));
            PC_CUDA(cudaHostAlloc((void**)&hc_in, Kmax * (D + P + 1) * 8, cudaHostAllocMapped));
        }
        hp.scratch = hc_scratch.p; hp.x = hc_x.p; hp.ch = hc_ch.p;
        double *d_out = nullptr, *d_in = nullptr;
        PC_CUDA(cudaHostGetDevicePointer((void**)&d_out, hc_out, 0));
        PC_CUDA(cudaHostGetDevicePointer((void**)&d_in, hc_in, 0));
        hp.out = d_out; hp.in = d_in;
        const int W = 4, blocks = std::max(1, (K + W - 1) / W);
        hc_begin_kernel<<<std::max(1, (std::max(K, Kdead) + W - 1) / W), W * 32, 0, stream>>>(hp);
        PC_CUDA(cudaGetLastError());
        std::vector<double> cube(D), theta(D), phi(std::max(P, 1));
        long long nl = 0;
        for (;;) {
            PC_CUDA(cudaStreamSynchronize(stream));
            int active = 0;
            for (int c = 0; c < K; ++c) {
                const double* o = hc_out + (size_t)c * (D + 2);
                if (o[D + 1] == 0.0) continue;
                ++active;
                double* in = hc_in + (size_t)c * (D + P + 1);
                if (o[D] != 0.0) {
                    std::copy(o, o + D, cube.begin());
                    std::fill(phi.begin(), phi.end(), 0.0);
                    g_host_prior(cube.data(), theta.data(), D);
                    const double logL = g_host_ll(theta.data(), D, phi.data(), P);
                    if (logL > S.logzero) ++nl;  // calculate.f90:44
                    std::copy(theta.begin(), theta.end(), in);
                    for (int i = 0; i < P; ++i) in[D + i] = phi[i];
                    in[D + P] = logL;
                } else {  // outside the cube the callbacks are not called (calculate.f90:36-39)
                    std::fill(in, in + D + P, 0.0);
                    in[D + P] = S.logzero;
                }
            }
            if (active == 0) break;
            if (g_abort) throw pc::RunError("polychord_b200: run aborted by a host callback (pc_request_abort)");
            hc_step_kernel<<<blocks, W * 32, 0, stream>>>(hp);
            PC_CUDA(cudaGetLastError());
            ++hc_rounds;
            launches += 1;
        }
        hc_finish_kernel<<<1, 1, 0, stream>>>(runs[0].st.p, nl, 1);
        PC_CUDA(cudaGetLastError());
        launches += 2;
    }

    void setup(const pc_settings& s, const ModelSpec& m, int nruns_, const int* seeds) {
        const bool dbg = std::getenv("PC_DEBUG") != nullptr;
        auto tm0 = std::chrono::steady_clock::now();
        auto mark = [&](const char* what) {
            if (!dbg) return;
            auto t = std::chrono::steady_clock::now();
            std::fprintf(stderr, "[pc dbg setup] %s %.3f ms\n", what, std::chrono::duration<double, std::milli>(t - tm0).count());
            tm0 = t;
        };
        device_check();
        mark("device_check");
        S = s; ms = m; nruns = nruns_;
        stream = g_stream;
        host_like = ms.like_kind == PC_LIKE_HOST;
        if (host_like && (nruns != 1 || g_mgpu.world > 1))
            throw pc::ArgError("polychord_b200: host-callback likelihoods run one run on one GPU");
        if (host_like && (!g_host_ll || !g_host_prior)) throw pc::ArgError("polychord_b200: host callbacks missing");
        build_dev_model(S, ms, dm, stream);
        mark("build_dev_model");
        int W = std::max(1, std::min(8, g_opt.warps_per_cta));
        const bool alone = nruns == 1;
        // the dense chain phase (a chain per point group) for runs that share the device; pc_dense.cuh
        const bool want_dense = g_opt.dense >= 0 && (g_opt.dense > 0 || nruns > 1) && !host_like && g_mgpu.world <= 1 &&
                                ms.like_kind != PC_LIKE_CORR_GAUSSIAN && g_grade_dims.size() <= 1;
        L = make_layout(S, ms, dm, W, alone, want_dense);
        W = L.W;
        set_smem(L.fn, L.smem);
        mark("layout+set_smem");
        KParams& k = L.kp;
        const int K = k.batch_K;
        // do_clustering: one run on one GPU with a device likelihood; elsewhere the run stays one cluster (always valid)
        k.clustering = (S.do_clustering && nruns == 1 && g_mgpu.world <= 1 && !host_like) ? 1 : 0;
        given_live = false;
        if (g_init_n > 0) {
            if (nruns != 1 || g_mgpu.world > 1) throw pc::ArgError("polychord_b200: cube_samples start one run on one GPU");
            if (g_init_D != k.cp.D) throw pc::ArgError("polychord_b200: cube_samples must hold points of nDims coordinates");
            if (!g_init_ll || !g_init_prior) throw pc::ArgError("polychord_b200: cube_samples need the run's callbacks");
            given_live = true;
            k.live_given = 1;
        }
        // boost_posterior: RTI%thin_posterior (generate.F90:311-316), in force when posterior samples are asked for
        // (run_time_info.f90:858); one run on one GPU (the phantoms of a sharded run stay with their ranks)
        k.boost_thin = 0.0;
        if ((S.posteriors || S.equals) && S.boost_posterior != 0.0 && nruns == 1 && g_mgpu.world <= 1)
            k.boost_thin = S.boost_posterior < 0.0 ? 1.0 : std::min(1.0, S.boost_posterior / (double)k.cp.R);
        // narrow phantoms (ChainParams::ph_narrow): nothing reads a phantom's theta unless phantoms are promoted to posterior
        // samples or the state is written out
        // -- and only where the streams are bound by DRAM traffic: the ensembles (dense chain phase).  For a run alone on
        // the device the pools sit in the L2, and two bulk copies per record cost phase U more than the bytes they save
        // (measured: G20 phase U 0.72 -> 0.80 ms, C50 183 -> 190 ms; ensemble 2.87 -> 2.93e9 evals/s)
        k.cp.ph_narrow = (k.dense && k.boost_thin == 0.0 && !g_resume.write && !g_opt.no_narrow) ? 1 : 0;
        g_mirror.boost_logw.clear(); g_mirror.boost_rows.clear(); g_mirror.boost_dead.clear(); g_mirror.boost_after.clear();
        // read_resume: a file of this run's shape continues the run (nested_sampling.F90:175-183)
        // (the caller's cube_samples win over an existing file, as in the reference: polychord.py:576-579 overwrites the
        // resume file with them)
        if (g_resume.read && g_init_n == 0 && nruns == 1 && g_mgpu.world <= 1 && pc::is_reference_resume(g_resume.path)) {
            import_reference_resume(g_resume.path, (unsigned)seeds[0]);   // the reference's text layout (or pypolychord's)
            resumed = true;
        } else if (g_resume.read && g_init_n == 0 && nruns == 1 && g_mgpu.world <= 1 && read_resume_file(g_resume.path, rd)) {
            const ResumeHeader& rh = rd.h;
            bool same = rh.D == k.cp.D && rh.P == k.cp.P && rh.n == k.n && rh.nmax == k.nmax && rh.R == k.cp.R && rh.batch_K == K &&
                        rh.like_kind == ms.like_kind && rh.clustering == k.clustering && rh.ngrade == k.cp.ngrade;
            for (int g = 0; same && g < k.cp.ngrade && k.cp.ngrade > 1; ++g)
                same = rh.gdims[g] == k.cp.gdims[g] && rh.greps[g] == k.cp.greps[g];
            if (!same)  // read_write.F90:402-417: a resume file of a different run is fatal
                throw pc::ArgError("polychord_b200: the resume file " + g_resume.path + " belongs to a run of a different shape "
                                            "(nDims, nDerived, nlive, num_repeats, grades, likelihood kind, clustering or batch size)");
            resumed = true;
            g_files.seed = rh.seed;   // the continuation is the file's run: the equal-weights thinning follows its seed too
        }
        // CTAs per run: one warp per chain unless capped by residency
        int dev = 0, sms = 0, per_sm = 0;
        PC_CUDA(cudaGetDevice(&dev));
        PC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        PC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessorWithFlags(&per_sm, k.dense ? L.fn.run_dense : L.fn.run, W * 32, L.smem, 0));
        if (per_sm < 1) throw pc::RunError("polychord_b200: run kernel does not fit on an SM");
        long long capacity = (long long)sms * per_sm;
        // A run alone on the device spreads its chains one per CTA (a chain warp then has an SM sub-partition
        // to itself) and leaves CTA 0 to the bookkeeping; an ensemble packs W chains per CTA.
        const bool sharded = g_mgpu.world > 1;
        if (sharded) {
            if (nruns != 1) throw pc::ArgError("polychord_b200: a sharded run cannot be part of an ensemble");
            if (k.n0 != k.n || k.dyn_m > 0)
                throw pc::ArgError("polychord_b200: a sharded run keeps its live count fixed (no nprior > nlive, nlives schedule or cube_samples of another length)");
            if (g_mgpu.batch_K != K || g_mgpu.T != k.cp.T || g_mgpu.D != k.cp.D)
                throw pc::ArgError("polychord_b200: pc_mgpu_create was called with different settings than this run");
            k.sh.rank = g_mgpu.rank; k.sh.world = g_mgpu.world; k.sh.xstride = (long long)mgpu_xstride(k.cp.D);
            k.sh.kr = (int)mgpu_kr(K, g_mgpu.world);
            for (int q = 0; q < g_mgpu.world; ++q) {
                unsigned char* b = (unsigned char*)g_mgpu.peer[q];
                k.sh.xbar[q] = (unsigned int*)b;
                k.sh.xin[q] = (double*)(b + 256);
                k.sh.xpart[q] = (double*)(b + 256 + (size_t)2 * K * k.cp.T * 8);
                k.sh.xrun[q] = k.sh.xpart[q] + (size_t)g_mgpu.world * mgpu_xstride(k.cp.D);
            }
        }
        G = (sharded ? (K + g_mgpu.world - 1) / g_mgpu.world : K) + 1;
        if (k.dense) G = (K + W * (32 / L.fn.G) - 1) / (W * (32 / L.fn.G));   // W * 32/G chains per CTA and pass
        if (g_opt.max_ctas > 0) G = std::min(G, g_opt.max_ctas);
        G = (int)std::max(1LL, std::min<long long>(G, capacity / nruns));
        if ((long long)G * nruns > capacity) throw pc::ArgError("polychord_b200: too many concurrent runs for one launch");
        k.ctas_per_run = G;
        k.chain_cta0 = (G >= 8 && !k.dense) ? 1 : 0;
        k.paired = (k.chain_cta0 == 1 && W >= 2 && (W % 2) == 0 && !g_opt.no_pairing) ? 1 : 0;
        k.backoff = nruns > 1 ? 128 : 0;
        PC_CUDA(cudaEventCreate(&ev0));
        PC_CUDA(cudaEventCreate(&ev1));
        mark("occupancy+events");

        const int T = k.cp.T, D = k.cp.D, R = k.cp.R, n = k.nmax;   // capacities follow the largest live count the run can reach
        runs.resize(nruns);
        std::vector<RunBuf> hb(nruns);
        for (int r = 0; r < nruns; ++r) {
            HostRun& h = runs[r];
            // a generation needs room for its K deaths, up to 2K births (failed ones go to the dead list) and the final kill-off
            long long cap_dead = S.max_ndead > 0 ? (long long)S.max_ndead + 2LL * n + 3LL * K : 40LL * n + 3LL * K;
            long long cap_ph = std::max<long long>(3LL * n * (R - 1) + 2LL * K * (R - 1), n);
            if (g_opt.cap_dead0 > 0) cap_dead = std::max<long long>(g_opt.cap_dead0, 2LL * n + 3LL * K);
            if (g_opt.cap_ph0 > 0) cap_ph = std::max<long long>(g_opt.cap_ph0, std::max<long long>(2LL * K * (R - 1), n));
            if (resumed) {  // room for what the file holds plus the next generation
                cap_dead = std::max<long long>(cap_dead, rd.st.ndead + 2LL * n + 3LL * K);
                cap_ph = std::max<long long>(cap_ph, rd.st.nphantom + 2LL * K * (R - 1) + n);
            }
            long long cap_boost = 0;
            if (k.boost_thin > 0.0) {
                cap_boost = 2 * cap_ph + (resumed ? (long long)rd.st.nboost : 0);
                h.boost.alloc((size_t)cap_boost * (T - D));
                h.boost_win.alloc((size_t)cap_boost);
            }
            h.st.alloc(1); h.st.zero(stream);
            h.live.alloc((size_t)n * T); h.live.zero(stream);
            h.order.alloc(2 * (size_t)n);
            h.okey.alloc(2 * (size_t)n);
            h.dead.alloc((size_t)cap_dead * T);
            h.logw.alloc(cap_dead);
            h.ph0.alloc((size_t)cap_ph * T);
            h.ph1.alloc((size_t)cap_ph * T);
            h.chol.alloc((size_t)D * D); h.cov.alloc((size_t)D * D);
            h.gsum.alloc((size_t)2 * D + 4);
            h.cfail.alloc((size_t)2 * K + 8); h.cfail.zero(stream);
            h.bkey.alloc((size_t)2 * K + 8); h.bkey.zero(stream);
            h.dpart.alloc((size_t)2 * G + 8); h.dpart.zero(stream);
            h.partial.alloc((size_t)G * k.partial_stride);
            h.pcount.alloc((size_t)cap_ph / U_TILE + 2); h.pcount.zero(stream);
            h.pmask.alloc((size_t)cap_ph / 32 + 8); h.pmask.zero(stream);
            if (k.dense) h.nh.alloc((size_t)G * W * (32 / L.fn.G) * R * dense_slb(L.fn.G * L.fn.DPL));   // slice records of the chains in flight
            else if (!k.nh_in_smem) h.nh.alloc((size_t)G * W * R * k.cp.LD);
            if (k.clustering) {
                h.deadlab.alloc(cap_dead); h.deadlab.zero(stream);
                h.deadn.alloc(cap_dead); h.deadn.zero(stream);
                h.lab.alloc(n); h.lab.zero(stream);
                h.phl0.alloc(cap_ph); h.phl0.zero(stream);
                h.phl1.alloc(cap_ph); h.phl1.zero(stream);
                h.cchol.alloc((size_t)MAX_CLUSTERS * D * D);
            }
            RunBuf& b = h.buf;
            std::memset(&b, 0, sizeof(b));
            b.st = h.st.p; b.live = h.live.p; b.live_snap = nullptr; b.ctl = nullptr; b.order = h.order.p; b.okey = h.okey.p; b.dead = h.dead.p; b.logw = h.logw.p;
            b.ph[0] = h.ph0.p; b.ph[1] = h.ph1.p; b.chol = h.chol.p; b.cov = h.cov.p; b.partial = h.partial.p;
            b.pcount = h.pcount.p; b.pmask = h.pmask.p; b.nh = h.nh.p; b.cap_dead = cap_dead; b.cap_ph = cap_ph; b.gsum = h.gsum.p;
            b.lab = h.lab.p; b.phl[0] = h.phl0.p; b.phl[1] = h.phl1.p; b.cchol = h.cchol.p;
            b.boost = h.boost.p; b.boost_win = h.boost_win.p; b.cap_boost = cap_boost;
            b.cfail = h.cfail.p; b.bkey = h.bkey.p; b.dpart = h.dpart.p; b.deadlab = h.deadlab.p; b.deadn = h.deadn.p;
            if (sharded)  // continue the cross-GPU barrier count of earlier runs (the counters are monotonic)
                PC_CUDA(cudaMemcpyAsync(&h.st.p->xepoch, &g_mgpu.epoch, sizeof(g_mgpu.epoch), cudaMemcpyHostToDevice, stream));
            b.seed = resumed ? rd.h.seed : (unsigned)seeds[r];   // a resumed run continues its own random stream
            hb[r] = b;
        }
        mark("allocs");
        if (resumed) {  // the run continues from the state the file holds
            HostRun& h = runs[0];
            DevRun st0 = rd.st;
            {   // the file's counters index device arrays sized from THIS run's settings: a damaged or foreign file that
                // passed the shape header must not reach the kernel (the reference's reader trusts its text file; here
                // the state is binary and drives pointers)
                const long long nl = st0.n, nd = st0.ndead, np = st0.nphantom, nmaxl = k.nmax;
                const bool counts = nl >= 1 && nl <= nmaxl && nd >= 0 && np >= 0 && (st0.cur_pool == 0 || st0.cur_pool == 1) &&
                                    (st0.order_off == 0 || st0.order_off == k.nmax) && st0.K >= 0 && st0.K <= k.nmax &&
                                    st0.ncl >= 0 && st0.ncl <= MAX_CLUSTERS && st0.nchains >= 0 && st0.ngen >= 0;
                const bool sizes = counts && (long long)rd.live.size() <= nmaxl * T && (long long)rd.live.size() >= nl * T &&
                                   rd.order.size() == 2 * (size_t)k.nmax && rd.okey.size() == 2 * (size_t)k.nmax &&
                                   (long long)rd.dead.size() == nd * T && (rd.logw.empty() || (long long)rd.logw.size() == nd) &&
                                   (long long)rd.ph.size() == np * T && rd.chol.size() == (size_t)D * D && rd.cov.size() == (size_t)D * D &&
                                   rd.gsum.size() == (size_t)2 * D + 4 && (rd.lab.empty() || (long long)rd.lab.size() <= nmaxl) &&
                                   (rd.phl.empty() || (long long)rd.phl.size() == np) &&
                                   (rd.cchol.empty() || rd.cchol.size() <= (size_t)MAX_CLUSTERS * D * D) &&
                                   nd <= (long long)h.dead.n / T && np <= (long long)(st0.cur_pool == 0 ? h.ph0.n : h.ph1.n) / T;
                bool order_ok = sizes;
                for (long long e = 0; order_ok && st0.order_valid && e < nl; ++e) { const int o = rd.order[(size_t)st0.order_off + e]; order_ok = o >= 0 && o < k.nmax; }   // (the other half is scratch)
                for (size_t e = 0; order_ok && e < rd.lab.size(); ++e) order_ok = rd.lab[e] >= 0 && rd.lab[e] < MAX_CLUSTERS;
                if (!order_ok)
                    throw pc::ArgError("polychord_b200: the resume file " + g_resume.path + " is damaged (its counters or array sizes do not fit the run it describes)");
            }
            st0.status = ST_RUNNING;
            st0.bar = 0; st0.wbar = 0;   // the barrier counters restart with the launch geometry of this process
            st0.dump_pub = 0; st0.snap_arr = 0u;   // ... and the dumper hand-over with this process's control block
            if (rd.st.status == ST_DONE) resumed_finished = true;
            h.st.upload(&st0, 1, stream);
            h.host_st = rd.st;
            h.live.upload(rd.live.data(), rd.live.size(), stream);
            h.order.upload(rd.order.data(), rd.order.size(), stream);
            h.okey.upload(rd.okey.data(), rd.okey.size(), stream);
            if (!rd.dead.empty()) h.dead.upload(rd.dead.data(), rd.dead.size(), stream);
            if (!rd.logw.empty()) h.logw.upload(rd.logw.data(), rd.logw.size(), stream);
            DevArr<double>& pool = rd.st.cur_pool == 0 ? h.ph0 : h.ph1;
            if (!rd.ph.empty()) pool.upload(rd.ph.data(), rd.ph.size(), stream);
            h.chol.upload(rd.chol.data(), rd.chol.size(), stream);
            h.cov.upload(rd.cov.data(), rd.cov.size(), stream);
            h.gsum.upload(rd.gsum.data(), rd.gsum.size(), stream);
            if (k.boost_thin > 0.0 && !rd.boost_win.empty() && rd.boost_win.size() == rd.st.nboost) {
                h.boost.upload(rd.boost.data(), rd.boost.size(), stream);
                h.boost_win.upload(rd.boost_win.data(), rd.boost_win.size(), stream);
            } else if (rd.st.nboost) {   // the file was written without the samples (or the run no longer boosts)
                const unsigned long long zero = 0;
                PC_CUDA(cudaMemcpyAsync(&h.st.p->nboost, &zero, sizeof(zero), cudaMemcpyHostToDevice, stream));
                h.host_st.nboost = 0;
            }
            if (k.clustering) {
                h.lab.upload(rd.lab.data(), rd.lab.size(), stream);
                DevArr<int>& pl = rd.st.cur_pool == 0 ? h.phl0 : h.phl1;
                if (!rd.phl.empty()) pl.upload(rd.phl.data(), rd.phl.size(), stream);
                if (!rd.cchol.empty()) h.cchol.upload(rd.cchol.data(), rd.cchol.size(), stream);
            }
            h2d += (long long)(rd.live.size() + rd.dead.size() + rd.ph.size()) * 8;
        }
        d_bufs.alloc(nruns);
        d_bufs.upload(hb.data(), nruns, stream);
        h2d += (long long)nruns * sizeof(RunBuf);
        k.runs = d_bufs.p;
        PC_CUDA(cudaStreamSynchronize(stream));
        mark("upload+sync");
    }

    void upload_bufs() {
        std::vector<RunBuf> hb(nruns);
        for (int r = 0; r < nruns; ++r) hb[r] = runs[r].buf;
        d_bufs.upload(hb.data(), nruns, stream);
        h2d += (long long)nruns * sizeof(RunBuf);
        PC_CUDA(cudaStreamSynchronize(stream));
    }

    void launch_async() {
        KParams kp = L.kp;
        void* args[] = {&kp};
        PC_CUDA(cudaEventRecord(ev0, stream));
        PC_CUDA(cudaLaunchCooperativeKernel(kp.dense ? L.fn.run_dense : L.fn.run, dim3(G * nruns), dim3(L.W * 32), args, L.smem, stream));
        PC_CUDA(cudaEventRecord(ev1, stream));
    }
    void launch_finish() {
        PC_CUDA(cudaEventSynchronize(ev1));
        float ms = 0;
        PC_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
        device_ms += ms;
        ++launches;
        for (int r = 0; r < nruns; ++r) {
            runs[r].st.download(&runs[r].host_st, 1, stream);
            d2h += sizeof(DevRun);
        }
        PC_CUDA(cudaStreamSynchronize(stream));
    }

    // dump, nested_sampling.F90:546-590: rows [theta, phi, birth, logL]; normalised posterior log-weights
    // state: {ndead, logZ, logZ2}; live_src: device records of the live points to report (null: none);
    // cs: the stream the copies are enqueued on (the copy stream while the run kernel is still sampling)
    void dump(int r, pc_dumper_t dumper, long long ndead, double logZ_raw, double logZ2_raw, const double* live_src,
              int nlive_now, cudaStream_t cs, long long nlike_now = 0, bool final_call = false, bool live_packed = false) {
        HostRun& h = runs[r];
        const KParams& k = L.kp;
        const int T = k.cp.T, D = k.cp.D, P = k.cp.P, n = nlive_now, npars = D + P + 2;
        const long long fresh = ndead - h.mirrored;
        const int nl = live_src ? n : 0;
        auto td0 = std::chrono::steady_clock::now();
        auto lap = [&](int slot) {
            auto t = std::chrono::steady_clock::now();
            g_dbg_dump[slot] += std::chrono::duration<double, std::milli>(t - td0).count();
            td0 = t;
        };
        // one batch of async copies into pinned staging, one synchronisation.  The dumper's rows are [theta, phi, birth,
        // logL] = columns D .. T-1 of a record (nested_sampling.F90:569-584): only those cross the bus (a strided copy;
        // the cube coordinates stay on the device)
        DumpMirror& mr = g_mirror;
        const bool direct = fresh > 0 && g_mirror_pin.covers(mr.rows, (size_t)ndead * npars);   // the rows go straight into the (page-locked) mirror
        double* sd = fresh > 0 ? (double*)g_pin_dead.need((size_t)fresh * (npars + 1) * 8) : nullptr;
        double* sl = nl > 0 ? (double*)g_pin_live.need((size_t)k.nmax * npars * 8) : nullptr;
        if (fresh > 0) {
            double* rows_dst = direct ? &mr.rows[(size_t)h.mirrored * npars] : sd;
            PC_CUDA(cudaMemcpy2DAsync(rows_dst, (size_t)npars * 8, h.dead.p + (size_t)h.mirrored * T + D, (size_t)T * 8, (size_t)npars * 8,
                                      (size_t)fresh, cudaMemcpyDeviceToHost, cs));
            h.logw.download(sd + (size_t)fresh * npars, fresh, cs, h.mirrored);
            d2h += fresh * (npars + 1) * 8;
        }
        if (nl > 0 && live_packed) {   // the kernel's snapshot: the reported columns only, one piece
            PC_CUDA(cudaMemcpyAsync(sl, live_src, (size_t)n * npars * 8, cudaMemcpyDeviceToHost, cs));
            d2h += (long long)n * npars * 8;
        } else if (nl > 0) {
            PC_CUDA(cudaMemcpy2DAsync(sl, (size_t)npars * 8, live_src + D, (size_t)T * 8, (size_t)npars * 8, (size_t)n,
                                      cudaMemcpyDeviceToHost, cs));
            d2h += (long long)n * npars * 8;
        }
        PC_CUDA(cudaStreamSynchronize(cs));
        lap(0);
        if (mr.rows.size() < (size_t)ndead * npars) {   // (not reached with a page-locked mirror: that one was sized before the launch)
            g_mirror_pin.drop();
            mr.rows.resize(std::max((size_t)ndead * npars, mr.rows.size() * 2));
        }
        if (mr.logw.size() < (size_t)ndead) { mr.logw.resize(std::max((size_t)ndead, mr.logw.size() * 2)); mr.lw.resize(mr.logw.size()); }
        if (fresh > 0) {
            const double* lwp = sd + (size_t)fresh * npars;
            if (!direct) std::memcpy(&mr.rows[(size_t)h.mirrored * npars], sd, (size_t)fresh * npars * sizeof(double));
            const double* fr = &mr.rows[(size_t)h.mirrored * npars];
            double mx = h.lse_max;
            for (long long i = 0; i < fresh; ++i) {
                const double v = lwp[i] + fr[(size_t)i * npars + npars - 1];
                mr.logw[h.mirrored + i] = v;
                mx = std::max(mx, v);
            }
            // running logsumexp of the posterior log-weights: only the fresh terms are exponentiated
            double sum = (h.lse_sum > 0.0) ? h.lse_sum * std::exp(h.lse_max - mx) : 0.0;
            if (mx > -std::numeric_limits<double>::infinity())
                for (long long i = 0; i < fresh; ++i) sum += std::exp(mr.logw[h.mirrored + i] - mx);
            h.lse_max = mx; h.lse_sum = sum;
            h.mirrored = ndead;
        }
        if (mr.live_rows.size() < (size_t)npars) mr.live_rows.resize((size_t)npars);
        const double* live_rows = nl > 0 ? sl : mr.live_rows.data();   // the dumper reads the live rows where the copy engine left them
        lap(1);
        std::vector<double>& lw = mr.lw;
        if (lw.empty()) lw.resize(1);
        if (ndead > 0) {  // normalised posterior log-weights (nested_sampling.F90:574-575)
            const double lse = h.lse_max + std::log(h.lse_sum);
            const double* src = mr.logw.data();
            double* dst = lw.data();
            for (long long i = 0; i < ndead; ++i) dst[i] = src[i] - lse;
        }
        double lz = std::max(-std::numeric_limits<double>::max(), 2 * logZ_raw - 0.5 * logZ2_raw);
        double var = logZ2_raw - 2 * logZ_raw;
        std::vector<double> dummy(npars, 0.0);
        lap(2);
        if (dumper)
            dumper((int)ndead, nl, npars, const_cast<double*>(live_rows), ndead > 0 ? mr.rows.data() : dummy.data(), lw.data(), lz,
                   std::sqrt(var));
        lap(3);
        if (k.boost_thin > 0.0 && r == 0) collect_boosted(h, cs, ndead, final_call);
        if (g_files.enabled && r == 0) {  // read_write.F90: the files are rewritten at every update and at the end
            BoostedRows br;
            br.n = (long long)mr.boost_logw.size(); br.rows = mr.boost_rows.data(); br.logw = mr.boost_logw.data();
            br.after = mr.boost_after.data();
            ClusterRows cr;
            const ClusterReport& rep = g_cluster_report;
            if (k.clustering && !rep.uid.empty() && (long long)rep.point_uid.size() >= (final_call ? ndead : 0)) {
                cr.n = (int)rep.uid.size(); cr.nactive = rep.nactive; cr.logZp = rep.zp.data(); cr.logZp2 = rep.zp2.data();
                cr.uid = rep.uid.data(); cr.nuid = (int)rep.parent.size(); cr.parent = rep.parent.data(); cr.frac = rep.frac.data();
                // the cluster files need every dead point's cluster: they are written with the final files
                cr.point_uid = (long long)rep.point_uid.size() >= ndead ? rep.point_uid.data() : nullptr;
                cr.cluster_posteriors = S.cluster_posteriors != 0;
            }
            write_run_files(g_files, g_fstate, D, P, ndead, mr.rows.data(), mr.logw.data(), nl, live_rows, lz,
                            std::sqrt(std::fabs(var)), nlike_now, final_call, br.n ? &br : nullptr, cr.n ? &cr : nullptr);
        }
    }

    // boost_posterior, host side.  The device list (rb.boost / rb.boost_win, filled by phase U) holds the promoted
    // phantoms of the updates so far; at the final call the phantoms still in the pool go through the same rule
    // against the deaths of the final kill-off (update_posteriors after the kill-off, nested_sampling.F90:381-386).
    // Each sample takes the weight of the death, within its window, with the smallest logL above its own
    // (run_time_info.f90:846-848; the dead points are in ascending logL, so that is a binary search).
    // derived parameters of a phantom record (gaussian.f90:37-40; Model::finish_derived leaves them unset for phantoms)
    void derived_of_phantom(double* rec) const {
        const KParams& k = L.kp;
        const int D = k.cp.D, P = k.cp.P;
        const bool gauss = k.cp.like_kind == LIKE_GAUSSIAN && (int)h_mu.size() == D;
        double r = 0.0;
        if (gauss) {
            double r2 = 0.0;
            for (int d = 0; d < D; ++d) { const double dl = rec[D + d] - h_mu[(size_t)d]; r2 += dl * dl; }
            r = std::sqrt(r2);
        }
        rec[2 * D] = r;
        if (P >= 2) rec[2 * D + 1] = gauss ? std::log(std::pow(r, (double)D) * k.cp.Vn) : 0.0;
        for (int i = 2; i < P; ++i) rec[2 * D + i] = 0.0;
    }
    std::vector<double> h_mu;   // host copy of the Gaussian's mean (the head of the device likelihood parameters)

    void collect_boosted(HostRun& h, cudaStream_t cs, long long ndead, bool final_call) {
        const KParams& k = L.kp;
        const int T = k.cp.T, D = k.cp.D, np = T - D;
        DumpMirror& mr = g_mirror;
        struct Item { long long end, dead; double l, lw; std::vector<double> row; };
        std::vector<Item> items;
        auto dead_logL = [&](long long i) { return mr.rows[(size_t)i * np + np - 1]; };
        auto place = [&](const double* row, long long first, long long end) {
            first = std::max(0LL, std::min(first, ndead)); end = std::max(first, std::min(end, ndead));
            const double l = row[np - 1];
            long long lo = first, hi = end;   // first death of the window with logL > l
            while (lo < hi) { const long long mid = (lo + hi) / 2; if (dead_logL(mid) > l) hi = mid; else lo = mid + 1; }
            if (lo >= end) return;            // no death above it: the reference keeps such a phantom (:850-851)
            Item it; it.end = end; it.dead = lo; it.l = l; it.lw = mr.logw[(size_t)lo] - dead_logL(lo) + l;
            it.row.assign(row, row + np);
            items.push_back(std::move(it));
        };
        const long long nb = std::min<long long>((long long)h.host_st.nboost, h.buf.cap_boost);
        const long long fresh = nb - h.boost_mirrored;
        if (fresh > 0) {
            std::vector<double> rows((size_t)fresh * np);
            std::vector<unsigned long long> win((size_t)fresh);
            h.boost.download(rows.data(), rows.size(), cs, (size_t)h.boost_mirrored * np);
            h.boost_win.download(win.data(), win.size(), cs, (size_t)h.boost_mirrored);
            PC_CUDA(cudaStreamSynchronize(cs));
            d2h += fresh * (np + 1) * 8;
            for (long long i = 0; i < fresh; ++i)
                place(&rows[(size_t)i * np], (long long)(win[(size_t)i] >> 32), (long long)(win[(size_t)i] & 0xffffffffull));
            h.boost_mirrored = nb;
        }
        if (final_call && h.host_st.nphantom > 0 && ndead > 0) {
            const long long nph = h.host_st.nphantom;
            std::vector<double> pool((size_t)nph * T);
            (h.host_st.cur_pool == 0 ? h.ph0 : h.ph1).download(pool.data(), pool.size(), cs);
            PC_CUDA(cudaStreamSynchronize(cs));
            d2h += nph * T * 8;
            if (k.cp.P > 0 && !host_like && k.cp.like_kind == LIKE_GAUSSIAN && h_mu.empty()) {
                h_mu.resize((size_t)D);
                dm.like.download(h_mu.data(), (size_t)D, cs);
                PC_CUDA(cudaStreamSynchronize(cs));
            }
            const double lmax = dead_logL(ndead - 1);
            for (long long i = 0; i < nph; ++i) {
                const double* rec = &pool[(size_t)i * T];
                const double l = rec[T - 1];
                if (!(lmax > l) || !(l > rec[T - 2])) continue;
                uint64_t bits;
                std::memcpy(&bits, &l, 8);
                if (!(uniform(h.buf.seed, TAG_BOOST, bits, 0u, 0u) < k.boost_thin)) continue;
                if (k.cp.P > 0 && !host_like) derived_of_phantom(&pool[(size_t)i * T]);
                place(rec + D, h.host_st.ndead_upd, ndead);
            }
        }
        // the device appends in no particular order: (update, logL, row) makes the list reproducible
        std::sort(items.begin(), items.end(), [](const Item& a, const Item& b) {
            if (a.end != b.end) return a.end < b.end;
            if (a.l != b.l) return a.l < b.l;
            return a.row < b.row;
        });
        for (const Item& it : items) {
            mr.boost_rows.insert(mr.boost_rows.end(), it.row.begin(), it.row.end());
            mr.boost_logw.push_back(it.lw);
            mr.boost_dead.push_back(it.dead);
            mr.boost_after.push_back(it.end);
        }
    }

    // ---- clustering at the update cadence (pc_cluster.cuh) -------------------------------------------------
    long long ncluster_updates = 0, ncluster_max = 0;
    double cluster_ms = 0, cluster_label_ms = 0;   // wall time of the clustering passes / of their labelling part
    std::vector<int> h_lab, h_knn;
    DevArr<int> d_part, d_knn, d_ccount;
    DevArr<double> d_cpart;

    // NN_clustering (clustering.f90:15-97) over all live points, from scratch.  The recursion into the clusters found
    // becomes a work list: every round the device computes the 10 nearest neighbours of every point WITHIN its
    // current part, and each part that is not final yet is split into the connected components of its
    // mutual-neighbour graph for n = 2, 3, ... (stopping at one component or at the first n that changes nothing).
    // Parts that come back whole are final.  Labels are canonical: in order of first appearance over the slots
    // (utils.F90:713-749 relabel), which is what the recursive formulation ends with.
    int cluster_labels(std::vector<int>& lab_out) {
        const KParams& k = L.kp;
        return cluster_labels_of(runs[0].live.p, runs[0].host_st.n, k.cp.D, k.cp.T, lab_out);
    }
    int cluster_labels_of(const double* d_live, int n, int D, int T, std::vector<int>& lab_out) {
        std::vector<int> part(n, 0), origin;
        const int nparts = refine_partition(d_live, n, D, T, part, 1, origin);
        std::vector<int> first(nparts, -1);
        int num = 0;
        lab_out.resize(n);
        for (int i = 0; i < n; ++i) {
            if (first[part[i]] < 0) first[part[i]] = num++;
            lab_out[i] = first[part[i]];
        }
        return num;
    }
    // The recursion itself, from ANY starting partition (do_clustering, clustering.f90:253-324, searches every existing
    // cluster for sub-clusters): part[i] in [0, nparts0) on entry; on return the parts are final, their number is
    // returned and origin[part] names the starting part each one descends from (a part only ever splits).
    int refine_partition(const double* d_live, int n, int D, int T, std::vector<int>& part, int nparts0, std::vector<int>& origin) {
        std::vector<char> final_part((size_t)nparts0, 0);
        int nparts = nparts0;
        origin.resize(nparts0);
        for (int c = 0; c < nparts0; ++c) origin[c] = c;
        if (d_part.n < (size_t)n) { d_part.alloc(n); d_knn.alloc((size_t)n * KNN_K); }
        h_knn.resize((size_t)n * KNN_K);
        std::vector<int> uf(n), loc(n), canon(n), old(n), head(n), twin_next(n), firstseen(n), members, bucket, bucket_off, part_dev;
        members.reserve(n);
        auto find = [&](int a) { while (uf[a] != a) { uf[a] = uf[uf[a]]; a = uf[a]; } return a; };
        for (int round = 0; round < 64; ++round) {
            bool any = false;
            for (int c = 0; c < nparts; ++c) any = any || !final_part[c];
            if (!any) break;
            part_dev.resize(n);   // final parts are marked: the kernel skips their points
            for (int i = 0; i < n; ++i) part_dev[i] = final_part[part[i]] ? -1 : part[i];
            d_part.upload(part_dev.data(), n, stream);
            const int W = 16;
            const bool tab = live_table_bytes(n, D) + (size_t)W * D * 8 <= 200 * 1024;   // the live table fits in shared memory
            const size_t ksm = (size_t)W * D * 8 + (tab ? live_table_bytes(n, D) : 0);
            PC_CUDA(cudaFuncSetAttribute(pc_knn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ksm));
            pc_knn_kernel<<<std::min((n + W - 1) / W, 148), W * 32, ksm, stream>>>(d_live, T, D, n, d_part.p, d_knn.p, tab ? 1 : 0);
            PC_CUDA(cudaGetLastError());
            d_knn.download(h_knn.data(), (size_t)n * KNN_K, stream);
            PC_CUDA(cudaStreamSynchronize(stream));
            launches += 1;
            h2d += (long long)n * 4; d2h += (long long)n * KNN_K * 4;
            const int nparts_now = nparts;
            // the points of every part, in slot order (one counting pass instead of a scan of all points per part)
            bucket_off.assign((size_t)nparts_now + 1, 0);
            for (int i = 0; i < n; ++i) bucket_off[part[i] + 1]++;
            for (int c = 0; c < nparts_now; ++c) bucket_off[c + 1] += bucket_off[c];
            bucket.resize(n);
            {
                std::vector<int> fill(bucket_off.begin(), bucket_off.end() - 1);
                for (int i = 0; i < n; ++i) bucket[fill[part[i]]++] = i;
            }
            for (int c = 0; c < nparts_now; ++c) {
                if (final_part[c]) continue;
                members.assign(bucket.begin() + bucket_off[c], bucket.begin() + bucket_off[c + 1]);
                const int m = (int)members.size();
                if (m <= 2) { final_part[c] = 1; continue; }   // do_clustering: nlive > 2 (clustering.f90:289); two points are each other's neighbours
                const int kk = std::min(m, KNN_K);
                for (int a = 0; a < m; ++a) { loc[members[a]] = a; uf[a] = a; old[a] = a; }
                // neighbours(): a ~ b when the HEAD of b's list appears in a's.  A point's nearest neighbour is itself
                // unless it has an exact twin with a lower slot, so heads are almost always the points themselves;
                // twins are chained behind their head (twin_next) and visited through it.
                bool twins = false;
                for (int a = 0; a < m; ++a) {
                    head[a] = loc[h_knn[(size_t)members[a] * KNN_K]];
                    twin_next[a] = -1;
                    if (head[a] != a) twins = true;
                }
                if (twins)
                    for (int a = m - 1; a >= 0; --a)
                        if (head[a] != a) { twin_next[a] = twin_next[head[a]]; twin_next[head[a]] = a; }
                auto join = [&](int a, int b) {
                    const int ra = find(a), rb = find(b);
                    if (ra != rb) uf[std::max(ra, rb)] = std::min(ra, rb);
                };
                int num = m;
                bool split = false;
                for (int nn = 2; nn <= kk; ++nn) {
                    // the graph only gains edges with n: add those of list position nn-1 (and, the first time, position 0)
                    for (int t = (nn == 2 ? 0 : nn - 1); t < nn; ++t)
                        for (int a = 0; a < m; ++a) {
                            const int v = h_knn[(size_t)members[a] * KNN_K + t];
                            if (v < 0) continue;
                            const int lv = loc[v];
                            if (head[lv] == lv) join(a, lv);
                            if (twins) for (int b = twin_next[lv]; b >= 0; b = twin_next[b]) join(a, b);
                        }
                    // canonical labels of the components, in order of first appearance
                    num = 0;
                    for (int a = 0; a < m; ++a) firstseen[a] = -1;
                    for (int a = 0; a < m; ++a) {
                        const int r0 = find(a);
                        if (firstseen[r0] < 0) firstseen[r0] = num++;
                        canon[a] = firstseen[r0];
                    }
                    if (num == 1) break;
                    bool same = true;
                    for (int a = 0; a < m && same; ++a) same = canon[a] == old[a];
                    if (same) { split = true; break; }
                    for (int a = 0; a < m; ++a) old[a] = canon[a];
                    if (nn == kk) split = true;
                }
                if (!split || num == 1) { final_part[c] = 1; continue; }
                // component 0 keeps the part's number, the others become new parts; all of them are searched again
                std::vector<int> newid(num, c);
                for (int s2 = 1; s2 < num; ++s2) { newid[s2] = nparts++; final_part.push_back(0); origin.push_back(origin[c]); }
                for (int a = 0; a < m; ++a) part[members[a]] = newid[canon[a]];
            }
        }
        return nparts;
    }

    // ---- clusters with an identity ----------------------------------------
    // The evidence of a run is GLOBAL (one nested-sampling run over the whole live set is exact whatever the shape of
    // the posterior; keeping a volume per cluster as run_time_info.f90:211-296 does is measurably biased, DESIGN.md
    // section 5.4), and the clusters are the reference's in every other respect: they persist -- at an update every
    // cluster is searched for sub-clusters and SPLIT (do_clustering, clustering.f90:253-324; add_cluster,
    // run_time_info.f90:303-505), a cluster without live points is DELETED (delete_cluster, :507-598) -- and every death
    // is attributed to the cluster of the dying point: local evidence Z_p = sum of w L over its deaths with the global
    // weight w = X / (n + 1), second moment <Z_p^2> by the global recurrences restricted to the cluster; a split hands
    // the pieces the parent's evidence in proportion to their live + phantom counts (as add_cluster does).
    struct ClusterBook {
        std::vector<int> uid_of_index;     // device label -> identity (the device labels only change in a clustering pass)
        std::vector<int> parent;           // per identity: the cluster it was split from (-1: the initial one)
        std::vector<double> Zp, Zp2, ZpXn; // per identity: log<Z_p>, log<Z_p^2>, log<Z_p X> - log<X>
        std::vector<double> frac;          // per identity: log(n_i / n), its share of the parent's evidence at the split (0: none)
        std::vector<int> dead_order;       // identities of the deleted clusters, in order of deletion
        std::vector<int> point_uid;        // per dead point: the identity of its cluster at its death
        long long processed = 0;           // dead points attributed so far
        double logXX = 0.0;                // global log<X^2> before the next death to attribute
        int new_uid(int par, double lz) {
            parent.push_back(par); Zp.push_back(lz); Zp2.push_back(lz); ZpXn.push_back(lz); frac.push_back(0.0);
            return (int)parent.size() - 1;
        }
    };
    ClusterBook book;
    void book_init(long long ndead0, int ncl0, double logXX0) {
        book = ClusterBook();
        for (int c = 0; c < std::max(1, ncl0); ++c) book.uid_of_index.push_back(book.new_uid(-1, S.logzero));
        book.processed = ndead0;
        book.point_uid.assign((size_t)ndead0, 0);
        book.logXX = logXX0;
    }
    static double lae(double a, double b) { return a > b ? a + std::log1p(std::exp(b - a)) : b + std::log1p(std::exp(a - b)); }
    // the deaths [book.processed, ndead): update_evidence's local part for the cluster of each (run_time_info.f90:211-296)
    void attribute_deaths(long long ndead) {
        HostRun& h = runs[0];
        const int T = L.kp.cp.T;
        const long long d0 = book.processed, cnt = ndead - d0;
        if (cnt <= 0) return;
        std::vector<int> lab((size_t)cnt), nn((size_t)cnt);
        std::vector<double> lw((size_t)cnt), ll((size_t)cnt);
        h.deadlab.download(lab.data(), (size_t)cnt, stream, (size_t)d0);
        h.deadn.download(nn.data(), (size_t)cnt, stream, (size_t)d0);
        h.logw.download(lw.data(), (size_t)cnt, stream, (size_t)d0);
        PC_CUDA(cudaMemcpy2DAsync(ll.data(), 8, h.dead.p + (size_t)d0 * T + T - 1, (size_t)T * 8, 8, (size_t)cnt, cudaMemcpyDeviceToHost, stream));
        PC_CUDA(cudaStreamSynchronize(stream));
        d2h += cnt * 24;
        const double log2 = std::log(2.0);
        book.point_uid.resize((size_t)ndead);
        for (long long i = 0; i < cnt; ++i) {
            const int lbl = std::max(0, std::min(lab[i], (int)book.uid_of_index.size() - 1));
            const int u = book.uid_of_index[lbl];
            book.point_uid[(size_t)(d0 + i)] = u;
            const int n = nn[i];
            if (n <= 0) continue;   // a failed birth: no weight
            const double L0 = ll[i], lognp = std::log((double)n), lognp1 = std::log(n + 1.0), lognp2 = std::log(n + 2.0);
            const double X_before = lw[i] + lognp1, X_after = X_before + lognp - lognp1, XX = book.logXX;
            const double ZpX_before = book.ZpXn[u] + X_before;
            book.Zp[u] = lae(book.Zp[u], X_before + L0 - lognp1);
            book.Zp2[u] = lae(book.Zp2[u], lae(log2 + ZpX_before + L0 - lognp1, log2 + XX + 2 * L0 - lognp1 - lognp2));
            // <Z_p X> shrinks with X for every cluster (that is the normalisation); the dying point's cluster gains a term
            book.ZpXn[u] = lae(book.ZpXn[u], XX + L0 + lognp - lognp1 - lognp2 - X_after);
            book.logXX = XX + lognp - lognp2;
        }
        book.processed = ndead;
    }
    // what pc_last_clusters reports: the clusters alive at the end (label order) then the deleted ones (order of deletion)
    void publish_clusters() {
        g_cluster_report.rows.clear(); g_cluster_report.uid.clear();
        g_cluster_report.nactive = (int)book.uid_of_index.size();
        g_cluster_report.parent = book.parent;
        g_cluster_report.point_uid = book.point_uid;
        g_cluster_report.frac = book.frac;
        g_cluster_report.zp.clear(); g_cluster_report.zp2.clear();
        auto put = [&](int u) {
            g_cluster_report.rows.push_back(book.Zp[u]); g_cluster_report.rows.push_back(book.Zp2[u]);
            g_cluster_report.zp.push_back(book.Zp[u]); g_cluster_report.zp2.push_back(book.Zp2[u]);
            g_cluster_report.uid.push_back(u);
        };
        for (int u : book.uid_of_index) put(u);
        for (int u : book.dead_order) put(u);
    }

    // the update's clustering pass: the new deaths are attributed, empty clusters deleted, every cluster searched for
    // sub-clusters and split; then the labels of the phantoms (identify_cluster) and one covariance + Cholesky factor per
    // cluster (calculate_covmats, run_time_info.f90:601-641)
    void cluster_pass() {
        const KParams& k = L.kp;
        const int n = runs[0].host_st.n, D = k.cp.D, T = k.cp.T;
        HostRun& h = runs[0];
        const auto tc0 = std::chrono::steady_clock::now();
        h_lab.resize(n);
        h.lab.download(h_lab.data(), n, stream);
        PC_CUDA(cudaStreamSynchronize(stream));
        d2h += (long long)n * 4;
        attribute_deaths(h.host_st.ndead);
        // delete_cluster: labels without a live point leave, the others keep their order
        int ncl0 = (int)book.uid_of_index.size();
        {
            std::vector<int> cnt(ncl0, 0), remap(ncl0, -1);
            for (int i = 0; i < n; ++i) { h_lab[i] = std::max(0, std::min(h_lab[i], ncl0 - 1)); cnt[h_lab[i]]++; }
            std::vector<int> keep;
            for (int c = 0; c < ncl0; ++c) {
                if (cnt[c] > 0 || (keep.empty() && c == ncl0 - 1)) { remap[c] = (int)keep.size(); keep.push_back(book.uid_of_index[c]); }
                else book.dead_order.push_back(book.uid_of_index[c]);
            }
            for (int i = 0; i < n; ++i) h_lab[i] = remap[h_lab[i]];
            book.uid_of_index.swap(keep);
            ncl0 = (int)book.uid_of_index.size();
        }
        // do_clustering: the recursion from the current clusters; then the new labels -- the clusters that stay whole first,
        // in their order, the pieces of the split ones behind them (add_cluster appends), pieces in order of first
        // appearance over the slots (relabel).  At most MAX_CLUSTERS clusters: a split that would exceed them is not made.
        std::vector<int> part = h_lab, origin;
        const int nparts = refine_partition(h.live.p, n, D, T, part, ncl0, origin);
        cluster_label_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tc0).count();
        struct Split { int parent_uid; std::vector<int> child_label; };
        std::vector<Split> splits;
        {
            std::vector<int> npieces(ncl0, 0), piece_no(nparts, -1);
            for (int i = 0; i < n; ++i)
                if (piece_no[part[i]] < 0) piece_no[part[i]] = npieces[origin[part[i]]]++;
            std::vector<char> split(ncl0, 0);
            int total = ncl0;
            for (int c = 0; c < ncl0; ++c)
                if (npieces[c] > 1 && total - 1 + npieces[c] <= MAX_CLUSTERS) { split[c] = 1; total += npieces[c] - 1; }
            std::vector<int> newlabel(ncl0, -1), base(ncl0, -1), uid_new;
            int next = 0;
            for (int c = 0; c < ncl0; ++c) if (!split[c]) { newlabel[c] = next++; uid_new.push_back(book.uid_of_index[c]); }
            for (int c = 0; c < ncl0; ++c)
                if (split[c]) {
                    base[c] = next;
                    Split sp; sp.parent_uid = book.uid_of_index[c];
                    for (int j = 0; j < npieces[c]; ++j) { sp.child_label.push_back(next++); uid_new.push_back(book.new_uid(sp.parent_uid, S.logzero)); }
                    splits.push_back(sp);
                }
            for (int i = 0; i < n; ++i) {
                const int c = origin[part[i]];
                h_lab[i] = split[c] ? base[c] + piece_no[part[i]] : newlabel[c];
            }
            book.uid_of_index.swap(uid_new);
        }
        const int num = (int)book.uid_of_index.size();
        const int ncl = num;
        ++ncluster_updates;
        ncluster_max = std::max<long long>(ncluster_max, num);
        h.lab.upload(h_lab.data(), n, stream);
        h2d += (long long)n * 4;
        if (ncl > 1) {
            const int cur = h.host_st.cur_pool;
            const long long nph = h.host_st.nphantom;
            const double* ph = cur == 0 ? h.ph0.p : h.ph1.p;
            int* phl = cur == 0 ? h.phl0.p : h.phl1.p;
            const int W = 8;
            if (nph > 0) {
                const int WI = 16;
                // register tile of the lane-per-phantom kernel: nDims rounded up to 2 (to 4 above 16); 0 = too wide
                const int dcap = D <= 16 ? ((D + 1) & ~1) : (D <= 32 ? ((D + 3) & ~3) : 0);
                const size_t lsm = dcap ? (size_t)n * ((dcap + 1) | 1) * 8 : 0;   // coordinates + |q|^2, odd row stride
                if (dcap && lsm <= 200 * 1024) {  // a lane per phantom, the live table in shared memory
                    const int blocks = (int)std::min<long long>((nph + 255) / 256, 148LL);
                    auto launch = [&](auto kern) {
                        PC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lsm));
                        kern<<<blocks, 256, lsm, stream>>>(h.live.p, T, D, n, h.lab.p, ph, nph, phl);
                    };
                    switch (dcap) {
                        case 2: launch(pc_identify_lanes_kernel<2>); break;
                        case 4: launch(pc_identify_lanes_kernel<4>); break;
                        case 6: launch(pc_identify_lanes_kernel<6>); break;
                        case 8: launch(pc_identify_lanes_kernel<8>); break;
                        case 10: launch(pc_identify_lanes_kernel<10>); break;
                        case 12: launch(pc_identify_lanes_kernel<12>); break;
                        case 14: launch(pc_identify_lanes_kernel<14>); break;
                        case 16: launch(pc_identify_lanes_kernel<16>); break;
                        case 20: launch(pc_identify_lanes_kernel<20>); break;
                        case 24: launch(pc_identify_lanes_kernel<24>); break;
                        case 28: launch(pc_identify_lanes_kernel<28>); break;
                        default: launch(pc_identify_lanes_kernel<32>); break;
                    }
                } else {
                    const bool tab = live_table_bytes(n, D) + (size_t)WI * D * 8 <= 200 * 1024;
                    const size_t ism = (size_t)WI * D * 8 + (tab ? live_table_bytes(n, D) : 0);
                    const int blocks = (int)std::min<long long>((nph + WI - 1) / WI, tab ? 148LL : 148LL * 8);
                    PC_CUDA(cudaFuncSetAttribute(pc_identify_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ism));
                    pc_identify_kernel<<<blocks, WI * 32, ism, stream>>>(h.live.p, T, D, n, h.lab.p, ph, nph, phl, tab ? 1 : 0);
                }
                PC_CUDA(cudaGetLastError());
            }
            if (!d_ccount.p) d_ccount.alloc(MAX_CLUSTERS);
            const int Dp8 = (D + 1 + 7) & ~7;
            if (d_cpart.n < (size_t)ncl * CLUSTER_CHUNKS * Dp8 * Dp8) d_cpart.alloc((size_t)MAX_CLUSTERS * CLUSTER_CHUNKS * Dp8 * Dp8);
            const size_t msm = cluster_moments_smem(D, W), fsm = cluster_factor_smem(D);
            PC_CUDA(cudaFuncSetAttribute(pc_cluster_moments_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msm));
            PC_CUDA(cudaFuncSetAttribute(pc_cluster_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm));
            pc_cluster_moments_kernel<<<dim3(ncl, CLUSTER_CHUNKS), W * 32, msm, stream>>>(h.live.p, h.lab.p, n, ph, phl, nph, T, D,
                                                                                        h.gsum.p + 2, d_cpart.p);
            PC_CUDA(cudaGetLastError());
            pc_cluster_factor_kernel<<<ncl, 128, fsm, stream>>>(d_cpart.p, CLUSTER_CHUNKS, D, h.chol.p, h.cchol.p, d_ccount.p);
            PC_CUDA(cudaGetLastError());
            launches += 2;
            if (!splits.empty()) {   // add_cluster step 5: the pieces share the parent's evidence by their live + phantom counts
                std::vector<int> cc(ncl);
                d_ccount.download(cc.data(), ncl, stream);
                PC_CUDA(cudaStreamSynchronize(stream));
                for (const Split& sp : splits) {
                    const int m = (int)sp.child_label.size();
                    std::vector<double> logni(m), logni1(m);
                    double tot = 0.0;
                    for (int j = 0; j < m; ++j) { const double c = (double)cc[sp.child_label[j]]; logni[j] = std::log(c); logni1[j] = std::log(c + 1.0); tot += c; }
                    const double logn = std::log(tot), logn1 = std::log(tot + 1.0);
                    const int pu = sp.parent_uid;
                    for (int j = 0; j < m; ++j) {
                        const int u = book.uid_of_index[sp.child_label[j]];
                        book.Zp[u] = book.Zp[pu] + logni[j] - logn;
                        book.Zp2[u] = book.Zp2[pu] + logni[j] + logni1[j] - logn - logn1;
                        book.ZpXn[u] = book.ZpXn[pu] + logni[j] - logn;
                        book.frac[u] = logni[j] - logn;
                    }
                }
            }
        }
        PC_CUDA(cudaMemcpyAsync(&h.st.p->ncl, &ncl, sizeof(int), cudaMemcpyHostToDevice, stream));
        PC_CUDA(cudaStreamSynchronize(stream));
        publish_clusters();
        cluster_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tc0).count();
    }

    // read_resume_file (read_write.F90:384-476) for a file in the reference's text layout -- written by the reference, by
    // pypolychord's _make_resume_file (polychord.py:650-789: a run started from the caller's live points) or by this
    // engine with the option "resume_text".  The file's state becomes the ResumeData the binary reader would have
    // produced, and the run continues in this engine's schedule.  nDims, nDerived and the grades must match (fatal
    // otherwise, :402-417).  Several clusters are taken as one (pc_resume_text.h).  The file carries no chain counter:
    // the chains are numbered from the dead count on, so a continuation is a valid run but not the bits of the
    // uninterrupted one (the binary layout gives that).
    void import_reference_resume(const std::string& path, unsigned seed) {
        const KParams& k = L.kp;
        pc::RefResume r;
        pc::read_reference_resume(path, r);
        const int D = k.cp.D, P = k.cp.P, T = k.cp.T;
        if (r.nDims != D) throw pc::ArgError("polychord_b200: resume error: nDims does not match");
        if (r.nDerived != P) throw pc::ArgError("polychord_b200: resume error: nDerived does not match");
        const int ngrade = std::max(1, k.cp.ngrade);
        if ((int)r.grade_dims.size() != ngrade) throw pc::ArgError("polychord_b200: resume error: number of grades does not match");
        for (int g = 0; g < ngrade; ++g)
            if (r.grade_dims[g] != (k.cp.ngrade > 1 ? k.cp.gdims[g] : D)) throw pc::ArgError("polychord_b200: resume error: Grades do not match");
        long long n = 0, nph = 0;
        for (int c = 0; c < r.ncluster; ++c) { n += r.nlive[c]; nph += r.nphantom[c]; }
        if (n < 2) throw pc::ArgError("polychord_b200: the resume file holds fewer than two live points (the run it describes has finished)");
        if (n > k.nmax) throw pc::ArgError("polychord_b200: the resume file holds more live points than this run's nlive / nprior allow");
        auto lse = [](const std::vector<double>& v) {
            double m = -INFINITY;
            for (double x : v) m = std::max(m, x);
            if (!std::isfinite(m)) return m;
            double s2 = 0.0;
            for (double x : v) s2 += std::exp(x - m);
            return m + std::log(s2);
        };
        std::memset(&rd.h, 0, sizeof(rd.h));
        std::memcpy(rd.h.magic, RESUME_MAGIC, 8);
        rd.h.version = 1; rd.h.sizeof_devrun = (int)sizeof(DevRun);
        rd.h.D = D; rd.h.P = P; rd.h.n = k.n; rd.h.nmax = k.nmax; rd.h.R = k.cp.R; rd.h.batch_K = k.batch_K; rd.h.like_kind = ms.like_kind;
        rd.h.clustering = k.clustering; rd.h.ngrade = k.cp.ngrade;
        for (int g = 0; g < MAX_GRADES; ++g) { rd.h.gdims[g] = k.cp.gdims[g]; rd.h.greps[g] = k.cp.greps[g]; }
        rd.h.seed = seed; rd.h.logzero = S.logzero;
        DevRun& st = rd.st;
        std::memset(&st, 0, sizeof(st));
        st.logZ = r.logZ; st.logZ2 = r.logZ2;
        st.logX = lse(r.logXp); st.logZX = lse(r.logZXp); st.logXX = lse(r.logXpXq);
        st.logX_last_update = r.logX_last_update;
        st.ndead = r.ndead; st.ndead_upd = r.ndead;
        st.nlike = 0;
        for (long long v : r.nlike) st.nlike += v;
        st.nchains = r.ndead;
        st.nphantom = nph; st.nph_glob = nph; st.ph_kept = nph;
        st.status = ST_RUNNING; st.initialised = 1; st.order_valid = 0; st.ncl = 1; st.n = (int)n; st.trimmed = 1;
        rd.live = r.live;
        rd.order.assign(2 * (size_t)k.nmax, 0);
        rd.okey.assign(2 * (size_t)k.nmax, 0.0);
        rd.dead = r.dead; rd.logw = r.logweights; rd.ph = r.phantom;
        rd.chol.assign(r.cholesky.begin(), r.cholesky.begin() + (size_t)D * D);   // the first cluster's factor until the next update
        rd.cov.assign(r.covmat.begin(), r.covmat.begin() + (size_t)D * D);
        rd.gsum.assign((size_t)2 * D + 4, 0.0);
        rd.gsum[0] = (double)nph;
        for (int e = 0; e < D; ++e) rd.gsum[2 + e] = rd.gsum[2 + D + e] = 0.5;   // pivot of the first covariance: the cube centre (init_phase)
        rd.lab.clear(); rd.phl.clear(); rd.cchol.clear(); rd.boost.clear(); rd.boost_win.clear();
        if (k.clustering) {
            rd.lab.assign((size_t)n, 0); rd.phl.assign((size_t)nph, 0);
            rd.cchol = rd.chol;
        }
        if ((long long)rd.live.size() != n * T || (long long)rd.dead.size() != r.ndead * T || (long long)rd.ph.size() != nph * T)
            throw pc::ArgError("polychord_b200: the resume file's point arrays do not have the run's record length");
    }

    // write_resume_file (read_write.F90:219-288) in the reference's text layout: this engine's state as ONE cluster (its
    // evidence is global), every dead point in the global posterior list
    void export_reference_resume(const std::string& path, const ResumeData& w) {
        const KParams& k = L.kp;
        const DevRun& st = w.st;
        const int D = k.cp.D, P = k.cp.P, T = k.cp.T;
        const int n = st.status == ST_DONE ? 0 : st.n;   // after the final kill-off every live point is in the dead list (nested_sampling.F90:381-384)
        pc::RefResume r;
        r.nDims = D; r.nDerived = P; r.ndead = st.ndead; r.ncluster = 1; r.ncluster_dead = 0;
        if (k.cp.ngrade > 1) {
            for (int g = 0; g < k.cp.ngrade; ++g) { r.grade_dims.push_back(k.cp.gdims[g]); r.num_repeats.push_back(k.cp.greps[g]); r.nlike.push_back(g == 0 ? st.nlike : 0); }
        } else { r.grade_dims = {D}; r.num_repeats = {k.cp.R}; r.nlike = {st.nlike}; }
        r.nlive = {n}; r.nphantom = {(int)st.nphantom};
        r.logZ = st.logZ; r.logZ2 = st.logZ2; r.thin_posterior = k.boost_thin;
        double lmin = INFINITY;
        for (int i = 0; i < n; ++i) lmin = std::min(lmin, w.live[(size_t)i * T + T - 1]);
        r.logLp = {lmin}; r.logXp = {st.logX}; r.logX_last_update = st.logX_last_update;
        r.logZXp = {st.logZX}; r.logZp = {st.logZ}; r.logZp2 = {st.logZ2}; r.logZpXp = {st.logZX}; r.logXpXq = {st.logXX};
        r.covmat = w.cov; r.cholesky = w.chol;
        r.live.assign(w.live.begin(), w.live.begin() + (size_t)n * T); r.dead = w.dead; r.logweights = w.logw; r.phantom = w.ph;
        // the dead points as global posterior points [logX, logL, log-weight, logZ, theta, derived] (settings.f90:189-204;
        // the reference's writers use the weight, logL and the parameters: read_write.F90:556-560)
        const int npost = 4 + D + P;
        r.nposterior_global = st.ndead;
        r.posterior_global.resize((size_t)st.ndead * npost);
        double mlw = S.logzero;
        for (long long i = 0; i < st.ndead; ++i) {
            const double* rec = &w.dead[(size_t)i * T];
            double* o = &r.posterior_global[(size_t)i * npost];
            o[0] = w.logw[i] + std::log((double)k.n + 1.0);
            o[1] = rec[T - 1];
            o[2] = w.logw[i];
            o[3] = st.logZ;
            std::memcpy(o + 4, rec + D, (size_t)(D + P) * sizeof(double));
            mlw = std::max(mlw, w.logw[i] + rec[T - 1]);
        }
        r.maxlogweight = {mlw};
        pc::write_reference_resume(path, r);
    }

    // write_resume_file (read_write.F90:219-288): the state between two generations
    void write_resume(bool final_call) {
        if (!g_resume.write || nruns != 1 || g_mgpu.world > 1) return;
        const auto now = std::chrono::steady_clock::now();
        if (!final_call && resume_written && std::chrono::duration<double>(now - resume_last).count() < g_resume.min_interval_s) return;
        const KParams& k = L.kp;
        HostRun& h = runs[0];
        const DevRun& st = h.host_st;
        const int n = st.n, T = k.cp.T, D = k.cp.D;
        ResumeData w;
        std::memset(&w.h, 0, sizeof(w.h));
        std::memcpy(w.h.magic, RESUME_MAGIC, 8);
        w.h.version = 1; w.h.sizeof_devrun = (int)sizeof(DevRun);
        w.h.D = D; w.h.P = k.cp.P; w.h.n = k.n; w.h.nmax = k.nmax; w.h.R = k.cp.R; w.h.batch_K = k.batch_K; w.h.like_kind = ms.like_kind;
        w.h.clustering = k.clustering; w.h.ngrade = k.cp.ngrade;
        for (int g = 0; g < MAX_GRADES; ++g) { w.h.gdims[g] = k.cp.gdims[g]; w.h.greps[g] = k.cp.greps[g]; }
        w.h.seed = h.buf.seed; w.h.logzero = S.logzero;
        w.st = st;
        auto pull = [&](auto& vec, const auto& arr, size_t cnt) { vec.resize(cnt); if (cnt) arr.download(vec.data(), cnt, stream); };
        pull(w.live, h.live, (size_t)n * T);
        pull(w.order, h.order, 2 * (size_t)k.nmax);
        pull(w.okey, h.okey, 2 * (size_t)k.nmax);
        pull(w.dead, h.dead, (size_t)st.ndead * T);
        pull(w.logw, h.logw, (size_t)st.ndead);
        pull(w.ph, st.cur_pool == 0 ? h.ph0 : h.ph1, (size_t)st.nphantom * T);
        pull(w.chol, h.chol, (size_t)D * D);
        pull(w.cov, h.cov, (size_t)D * D);
        pull(w.gsum, h.gsum, (size_t)2 * D + 4);
        if (k.boost_thin > 0.0) {
            pull(w.boost, h.boost, (size_t)st.nboost * (T - D));
            pull(w.boost_win, h.boost_win, (size_t)st.nboost);
        }
        if (k.clustering) {
            pull(w.lab, h.lab, (size_t)n);
            pull(w.phl, st.cur_pool == 0 ? h.phl0 : h.phl1, (size_t)st.nphantom);
            pull(w.cchol, h.cchol, (size_t)std::max(st.ncl, 1) * D * D);
        }
        PC_CUDA(cudaStreamSynchronize(stream));
        d2h += (long long)(w.live.size() + w.dead.size() + w.ph.size()) * 8;
        const std::string tmp = g_resume.path.substr(0, g_resume.path.size() - 7) + "_temp.resume";
        if (g_opt.resume_text) {
            export_reference_resume(tmp, w);
            std::rename(tmp.c_str(), g_resume.path.c_str());
            resume_written = true;
            resume_last = std::chrono::steady_clock::now();
            return;
        }
        FILE* f = std::fopen(tmp.c_str(), "wb");
        if (!f) throw pc::RunError("polychord_b200: cannot write " + tmp);
        std::fwrite(&w.h, sizeof(w.h), 1, f);
        std::fwrite(&w.st, sizeof(DevRun), 1, f);
        put_vec(f, w.live); put_vec(f, w.order); put_vec(f, w.okey); put_vec(f, w.dead); put_vec(f, w.logw); put_vec(f, w.ph);
        put_vec(f, w.chol); put_vec(f, w.cov); put_vec(f, w.gsum); put_vec(f, w.lab); put_vec(f, w.phl); put_vec(f, w.cchol);
        put_vec(f, w.boost); put_vec(f, w.boost_win);
        const bool wrote = std::ferror(f) == 0;
        if (std::fclose(f) != 0 || !wrote) {   // a short write (disk full): the last good file stays in place
            std::remove(tmp.c_str());
            throw pc::RunError("polychord_b200: writing " + tmp + " failed");
        }
        std::rename(tmp.c_str(), g_resume.path.c_str());
        resume_written = true;
        resume_last = std::chrono::steady_clock::now();
    }

    void grow(int r, int status) {
        HostRun& h = runs[r];
        const KParams& k = L.kp;
        if (status == ST_NEED_BOOST) {
            const long long nc = h.buf.cap_boost * 2 + h.buf.cap_ph, np = k.cp.T - k.cp.D;
            h.boost.grow((size_t)nc * np, (size_t)h.host_st.nboost * np, stream);
            h.boost_win.grow((size_t)nc, (size_t)h.host_st.nboost, stream);
            h.buf.boost = h.boost.p; h.buf.boost_win = h.boost_win.p; h.buf.cap_boost = nc;
        } else if (status == ST_NEED_DEAD) {
            long long nc = h.buf.cap_dead * 2 + k.nmax + 3LL * k.batch_K;
            h.dead.grow((size_t)nc * k.cp.T, (size_t)h.host_st.ndead * k.cp.T, stream);
            h.logw.grow(nc, h.host_st.ndead, stream);
            if (k.clustering) {
                h.deadlab.grow(nc, h.host_st.ndead, stream);
                h.deadn.grow(nc, h.host_st.ndead, stream);
                h.buf.deadlab = h.deadlab.p; h.buf.deadn = h.deadn.p;
            }
            h.buf.dead = h.dead.p; h.buf.logw = h.logw.p; h.buf.cap_dead = nc;
        } else {
            long long nc = h.buf.cap_ph * 2;
            int cur = h.host_st.cur_pool;
            DevArr<double>& a = cur == 0 ? h.ph0 : h.ph1;
            DevArr<double>& b = cur == 0 ? h.ph1 : h.ph0;
            a.grow((size_t)nc * k.cp.T, (size_t)h.host_st.nphantom * k.cp.T, stream);
            b.alloc((size_t)nc * k.cp.T);
            if (k.clustering) {
                DevArr<int>& la = cur == 0 ? h.phl0 : h.phl1;
                DevArr<int>& lb = cur == 0 ? h.phl1 : h.phl0;
                la.grow((size_t)nc, (size_t)h.host_st.nphantom, stream);
                lb.alloc((size_t)nc);
                h.buf.phl[0] = h.phl0.p; h.buf.phl[1] = h.phl1.p;
            }
            h.pcount.alloc((size_t)nc / U_TILE + 2);
            h.buf.pcount = h.pcount.p;
            h.pmask.alloc((size_t)nc / 32 + 8);
            h.buf.pmask = h.pmask.p;
            h.buf.ph[0] = h.ph0.p; h.buf.ph[1] = h.ph1.p; h.buf.cap_ph = nc;
        }
    }

    void run(pc_dumper_t dumper, pc_run_info* out) {
        auto t0 = std::chrono::steady_clock::now();
        volatile HostCtl* ctl = nullptr;
        unsigned long long handled = 0;
        const bool boosting = L.kp.boost_thin > 0.0;   // the promoted phantoms are collected at every update
        const bool sync_dump = g_opt.sync_dump || std::getenv("PC_SYNC_DUMP") || host_like || L.kp.clustering || g_resume.write || boosting;  // these runs return to the host at every update anyway
        if (given_live && !resumed) host_given_live_points();
        else if (host_like && !resumed) host_generate_live_points();
        const bool want_files = g_files.enabled && nruns == 1;
        const bool dumping = (dumper != nullptr || want_files || boosting) && nruns == 1;
        if (dumping && !sync_dump) {  // asynchronous dumper hand-over
            HostCtl* c = host_ctl();
            std::memset(c, 0, sizeof(*c));
            ctl = c;
            HostRun& h = runs[0];
            h.live_snap.alloc((size_t)2 * L.kp.nmax * L.kp.cp.T);   // two halves, by dump parity
            h.buf.live_snap = h.live_snap.p;
            HostCtl* dctl = nullptr;
            PC_CUDA(cudaHostGetDevicePointer((void**)&dctl, c, 0));
            h.buf.ctl = dctl;
            upload_bufs();
        }
        L.kp.want_dump = (dumping || (g_resume.write && nruns == 1)) ? 1 : 0;
        double dbg_service_ms = 0, dbg_launch_ms = 0, dbg_finish_ms = 0, dbg_final_ms = 0;
        auto now = [] { return std::chrono::steady_clock::now(); };
        auto ms_since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(now() - t).count(); };
        auto service = [&]() {  // hand one published dump to the user's dumper
            auto ts = now();
            const long long nd = ctl->ndead;
            const double lz = ctl->logZ, lz2 = ctl->logZ2;
            dump(0, dumper, nd, lz, lz2, runs[0].live_snap.p + (size_t)((handled + 1) & 1) * L.kp.nmax * (L.kp.cp.T - L.kp.cp.D), ctl->nlive,
                 g_copy_stream, ctl->nlike, false, true);
            ++handled;
            ctl->ack_seq = handled;
            dbg_service_ms += ms_since(ts);
        };
        if (L.kp.clustering) book_init(resumed ? runs[0].host_st.ndead : 0, resumed ? runs[0].host_st.ncl : 1, resumed ? runs[0].host_st.logXX : 0.0);
        else g_cluster_report = ClusterReport();
        for (; !resumed_finished;) {
            if (ctl) {  // size the pinned staging now: (re)allocating pinned memory synchronises with the running kernel
                g_pin_dead.need((size_t)runs[0].buf.cap_dead * (L.kp.cp.T + 1) * 8);
                g_pin_live.need((size_t)L.kp.nmax * L.kp.cp.T * 8);
                g_mirror_pin.reserve(g_mirror.rows, (size_t)runs[0].buf.cap_dead * (L.kp.cp.T - L.kp.cp.D));
            }
            { auto tl = now(); launch_async(); dbg_launch_ms += ms_since(tl); }
            if (ctl) {
                try {
                    while (cudaEventQuery(ev1) == cudaErrorNotReady) {
                        if (ctl->dump_seq > handled) service();
                        else std::this_thread::yield();
                    }
                    while (ctl->dump_seq > handled) service();
                } catch (...) {  // the dumper threw (e.g. a Python exception): release the kernel, then unwind
                    ctl->abort = 1;
                    cudaEventSynchronize(ev1);
                    throw;
                }
            }
            { auto tf = now(); launch_finish(); dbg_finish_ms += ms_since(tf); }
            bool all_done = true, regrow = false;
            for (int r = 0; r < nruns; ++r) {
                int stt = runs[r].host_st.status;
                if (stt == ST_ERROR) throw pc::RunError("polychord_b200: could not generate live points (likelihood <= logzero everywhere?)");
                if (stt == ST_DUMP && ctl) throw pc::RunError("polychord_b200: run aborted");
                if (stt == ST_DUMP) {  // sync_dump: the kernel left at the update, dump and relaunch
                    if (dumping)
                        dump(r, dumper, runs[r].host_st.ndead, runs[r].host_st.logZ, runs[r].host_st.logZ2, runs[r].live.p,
                             runs[r].host_st.n, stream, runs[r].host_st.nlike);
                    write_resume(false);
                }
                if (g_abort) throw pc::RunError("polychord_b200: run aborted by a host callback (pc_request_abort)");
                if (stt == ST_HOSTCHAINS) host_chains();
                if (stt == ST_CLUSTER) {  // the update left for the clustering pass; the dump of this update happens here too
                    if (dumping)
                        dump(r, dumper, runs[r].host_st.ndead, runs[r].host_st.logZ, runs[r].host_st.logZ2, runs[r].live.p,
                             runs[r].host_st.n, stream, runs[r].host_st.nlike);
                    cluster_pass();
                    PC_CUDA(cudaMemcpyAsync(&runs[r].host_st.ncl, &runs[r].st.p->ncl, sizeof(int), cudaMemcpyDeviceToHost, stream));
                    PC_CUDA(cudaStreamSynchronize(stream));
                    write_resume(false);
                }
                if (stt == ST_NEED_DEAD || stt == ST_NEED_PHANTOM || stt == ST_NEED_BOOST) { grow(r, stt); regrow = true; }
                if (stt != ST_DONE) all_done = false;
            }
            if (regrow) upload_bufs();
            if (all_done) break;
        }
        if (g_mgpu.world > 1) g_mgpu.epoch = runs[0].host_st.xepoch;
        if (L.kp.clustering && !resumed_finished) {
            // the deaths since the last pass and the final kill-off; clusters that emptied during sampling count as deleted
            HostRun& h = runs[0];
            const int n = h.host_st.n, ncl0 = (int)book.uid_of_index.size();
            h_lab.resize(n);
            h.lab.download(h_lab.data(), n, stream);
            PC_CUDA(cudaStreamSynchronize(stream));
            attribute_deaths(h.host_st.ndead);
            std::vector<int> cnt(ncl0, 0), keep;
            for (int i = 0; i < n; ++i) cnt[std::max(0, std::min(h_lab[i], ncl0 - 1))]++;
            for (int c = 0; c < ncl0; ++c) {
                if (cnt[c] > 0 || (keep.empty() && c == ncl0 - 1)) keep.push_back(book.uid_of_index[c]);
                else book.dead_order.push_back(book.uid_of_index[c]);
            }
            book.uid_of_index.swap(keep);
            publish_clusters();
        }
        for (int r = 0; r < nruns; ++r)   // nested_sampling.F90:407-409
            if (runs[r].host_st.stop_nfail && S.feedback >= 0)
                std::printf("Warning, unable to proceed after %6d: failed spawn events\n", runs[r].host_st.fail_run);
        if (g_final_live.want && nruns == 1 && runs[0].host_st.ndead >= runs[0].host_st.n) {
            // the final kill-off appended the n live points, lowest logL first: they are the last n dead records
            const size_t n = (size_t)runs[0].host_st.n, T = (size_t)L.kp.cp.T;
            g_final_live.n = (int)n;
            g_final_live.recs.resize(n * T);
            runs[0].dead.download(g_final_live.recs.data(), n * T, stream, ((size_t)runs[0].host_st.ndead - n) * T);
            PC_CUDA(cudaStreamSynchronize(stream));
            d2h += (long long)(n * T * 8);
        }
        auto tfin = now();
        if (dumping)  // the final call: every point is dead (nested_sampling.F90:392)
            for (int r = 0; r < nruns; ++r)
                dump(r, dumper, runs[r].host_st.ndead, runs[r].host_st.logZ, runs[r].host_st.logZ2, nullptr, 0, stream,
                     runs[r].host_st.nlike, true);
        if (want_files && !resumed) write_prior_info(g_files, L.kp.n0, runs[0].host_st.init_attempts);
        if (!resumed_finished) write_resume(true);
        dbg_final_ms = ms_since(tfin);
        if (std::getenv("PC_DEBUG")) {
            std::fprintf(stderr, "[pc dbg dump] copies %.3f ms, rows %.3f ms, weights %.3f ms, user dumper %.3f ms\n", g_dbg_dump[0],
                         g_dbg_dump[1], g_dbg_dump[2], g_dbg_dump[3]);
            g_dbg_dump[0] = g_dbg_dump[1] = g_dbg_dump[2] = g_dbg_dump[3] = 0;
        }
        if (std::getenv("PC_DEBUG"))
            std::fprintf(stderr, "[pc dbg run] launch %.3f ms, dumps served %llu in %.3f ms, finish(wait+download) %.3f ms, final dump %.3f ms\n",
                         dbg_launch_ms, handled, dbg_service_ms, dbg_finish_ms, dbg_final_ms);
        auto t1 = std::chrono::steady_clock::now();
        const KParams& k = L.kp;
        for (int r = 0; r < nruns; ++r) {
            const DevRun& s = runs[r].host_st;
            pc_run_info& o = out[r];
            std::memset(&o, 0, sizeof(o));
            o.status = 0;
            o.logZ = std::max(-std::numeric_limits<double>::max(), 2 * s.logZ - 0.5 * s.logZ2);  // run_time_info.f90:663-664
            o.logZerr = std::sqrt(s.logZ2 - 2 * s.logZ);
            o.logZ_raw = s.logZ; o.logZ2_raw = s.logZ2;
            o.ndead = s.ndead; o.nlike = s.nlike; o.nchains = s.nchains; o.ngenerations = s.ngen;
            o.nupdates = s.nupdates; o.nfailures = s.nfail; o.nslices = s.nslices; o.nphantoms_final = s.nphantom;
            o.ncluster_max = ncluster_max; o.ncluster_updates = ncluster_updates; o.cluster_ms = cluster_ms;
            if (std::getenv("PC_DEBUG") && ncluster_updates)
                std::fprintf(stderr, "[pc dbg cluster] %lld passes, %.3f ms in all, %.3f ms of it labelling (kNN kernel + union-find), up to %lld clusters\n",
                             ncluster_updates, cluster_ms, cluster_label_ms, ncluster_max);
            o.batch_K = k.batch_K; o.warps_per_cta = L.W; o.ctas_per_run = G; o.kernel_launches = launches;
            o.kernel_G = L.fn.G; o.kernel_DPL = L.fn.DPL; o.kernel_kind = L.fn.KIND; o.kernel_mode = k.dense;
            o.nlive_final = s.n;
            o.device_ms = device_ms;
            o.wall_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
            o.h2d_bytes = h2d; o.d2h_bytes = d2h;
            // 8T+8D per slice step, 8T per chain (seed read) + 8T (dead record), 8D^2 per generation
            {
                const int khz = sm_clock_khz();
                const long long cyc[8] = {s.cyc_wait, s.cyc_S, s.cyc_fin, s.cyc_U, s.cyc_prep, s.cyc_white, s.cyc_slice, s.cyc_total};
                for (int i = 0; i < 8; ++i) o.phase_ms[i] = khz > 0 ? (double)cyc[i] / (double)khz : 0.0;
            }
            if (std::getenv("PC_DEBUG")) {
                const int khz = std::max(1, sm_clock_khz());
                std::fprintf(stderr, "[pc dbg ms]");
                for (int i = 0; i < 24; ++i) std::fprintf(stderr, " %.3f", (double)s.dbg[i] / khz);
                std::fprintf(stderr, "\n");
            }
            o.algorithmic_bytes = s.nslices * (8LL * k.cp.T + 8LL * k.cp.D) + s.nchains * 16LL * k.cp.T + s.ngen * 8LL * k.cp.D * k.cp.D;
        }
    }
};

// ------------------------------------------------------------------------------------------
// registry: host callback pointer -> device form
// ------------------------------------------------------------------------------------------
struct LikeReg { int kind; std::vector<double> params; };
struct PriorReg { int kind; std::vector<double> params; };
static std::map<void*, LikeReg>& like_registry() { static std::map<void*, LikeReg> m; return m; }
static std::map<void*, PriorReg>& prior_registry() { static std::map<void*, PriorReg> m; return m; }

static int fail(int code, const std::string& msg) {
    std::fprintf(stderr, "\n polychord_b200: %s\n", msg.c_str());
    g_last.status = code;
    if (!g_opt.errors_return && !std::getenv("PC_ERRORS_RETURN")) std::exit(1);  // abort.F90:19-29 convention
    return code;
}

}  // namespace pc

using namespace pc;

// the facade's defaults (csrc/pc_facade.cpp; c_interface.cpp:210-213): recognised by pointer identity
void default_prior(double* cube, double* theta, int nDims);
void default_dumper(int, int, int, double*, double*, double*, double, double);

// ==========================================================================================
// C ABI
// ==========================================================================================
extern "C" {

int pc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}
const char* pc_version(void) { return "polychordlite_b200 0.1.0 (PolyChordLite 1.22.2 API)"; }

int pc_set_option(const char* name, double value) {
    std::string s(name);
    if (s == "batch_fraction") g_opt.batch_fraction = value;
    else if (s == "batch_K") g_opt.batch_K = (int)value;
    else if (s == "dense") g_opt.dense = (int)value;
    else if (s == "device") g_opt.device = (int)value;
    else if (s == "warps_per_cta") g_opt.warps_per_cta = (int)value;
    else if (s == "max_ctas") g_opt.max_ctas = (int)value;
    else if (s == "errors_return") g_opt.errors_return = (int)value;
    else if (s == "nh_global") g_opt.nh_global = (int)value;
    else if (s == "no_pairing") g_opt.no_pairing = (int)value;
    else if (s == "no_phase_d") g_opt.no_phase_d = (int)value;
    else if (s == "no_bulk") g_opt.no_bulk = (int)value;
    else if (s == "no_wave_batch") g_opt.no_wave_batch = (int)value;
    else if (s == "no_narrow") g_opt.no_narrow = (int)value;
    else if (s == "resume_text") g_opt.resume_text = (int)value;
    else if (s == "sync_dump") g_opt.sync_dump = (int)value;
    else if (s == "cap_dead0") g_opt.cap_dead0 = (long long)value;
    else if (s == "cap_ph0") g_opt.cap_ph0 = (long long)value;
    else if (s == "resume_interval") g_resume.min_interval_s = value;
    else return -1;
    return 0;
}
double pc_get_option(const char* name) {
    std::string s(name);
    if (s == "batch_fraction") return g_opt.batch_fraction;
    if (s == "batch_K") return g_opt.batch_K;
    if (s == "dense") return g_opt.dense;
    if (s == "device") return g_opt.device;
    if (s == "warps_per_cta") return g_opt.warps_per_cta;
    if (s == "max_ctas") return g_opt.max_ctas;
    if (s == "errors_return") return g_opt.errors_return;
    if (s == "nh_global") return g_opt.nh_global;
    if (s == "no_pairing") return g_opt.no_pairing;
    if (s == "no_phase_d") return g_opt.no_phase_d;
    if (s == "no_bulk") return g_opt.no_bulk;
    if (s == "no_wave_batch") return g_opt.no_wave_batch;
    if (s == "no_narrow") return g_opt.no_narrow;
    if (s == "resume_text") return g_opt.resume_text;
    if (s == "sync_dump") return g_opt.sync_dump;
    if (s == "cap_dead0") return (double)g_opt.cap_dead0;
    if (s == "cap_ph0") return (double)g_opt.cap_ph0;
    if (s == "resume_interval") return g_resume.min_interval_s;
    return NAN;
}
void pc_set_stream(void* cuda_stream) { g_stream = (cudaStream_t)cuda_stream; }
void pc_request_abort(void) { g_abort = 1; }
int pc_set_resume(const char* path, int write, int read) {
    g_resume.write = write != 0 && path;
    g_resume.read = read != 0 && path;
    g_resume.path = path ? path : "";
    if (g_resume.path.size() < 8 || g_resume.path.substr(g_resume.path.size() - 7) != ".resume") {
        g_resume.write = g_resume.read = false;
        return path ? -1 : 0;
    }
    return 0;
}
int pc_set_grades(int nGrade, const int* grade_dims, const int* grade_repeats) {
    if (nGrade < 0 || nGrade > MAX_GRADES) return -1;
    g_grade_dims.assign(grade_dims, grade_dims + nGrade);
    g_grade_reps.assign(grade_repeats, grade_repeats + nGrade);
    return 0;
}
int pc_set_nlives(const double* loglikes, const int* nlives, int m) {
    g_dyn_loglikes.clear(); g_dyn_nlives.clear();
    if (m <= 0 || !loglikes || !nlives) return 0;
    if (m > MAX_DYN) return -1;
    g_dyn_loglikes.assign(loglikes, loglikes + m);
    g_dyn_nlives.assign(nlives, nlives + m);
    return 0;
}
void pc_release_memory(void) { pool().trim(); }
int pc_auto_batch_size(int nlive, int world) {
    try {
        device_check();
        return batch_size(nlive, true, std::max(1, world));
    } catch (const std::exception&) {
        return -1;
    }
}

// ---- sharded run over the GPUs of one box ---------------------------------------------------------
int pc_mgpu_create(const pc_settings* s, int world, unsigned char* handle64) {
    try {
        device_check();
        if (world < 2 || world > MAX_RANKS) throw pc::ArgError("polychord_b200: world must be 2..8");
        if (g_mgpu.local) throw pc::ArgError("polychord_b200: pc_mgpu_create called twice (pc_mgpu_destroy first)");
        const int K = batch_size(s->nlive, true, world);
        const int T = 2 * s->nDims + s->nDerived + 2;
        g_mgpu.bytes = mgpu_block_bytes(K, T, s->nDims, world);
        g_mgpu.batch_K = K; g_mgpu.T = T; g_mgpu.D = s->nDims; g_mgpu.world = 0; g_mgpu.epoch = 0;
        PC_CUDA(cudaMalloc(&g_mgpu.local, g_mgpu.bytes));  // not from the pool: the block is exported through CUDA IPC
        PC_CUDA(cudaMemset(g_mgpu.local, 0, g_mgpu.bytes));
        PC_CUDA(cudaDeviceSynchronize());
        cudaIpcMemHandle_t h;
        PC_CUDA(cudaIpcGetMemHandle(&h, g_mgpu.local));
        static_assert(sizeof(h) == 64, "CUDA IPC handle size");
        std::memcpy(handle64, &h, 64);
        return 0;
    } catch (const std::exception& ex) {
        return fail(-7, ex.what());
    }
}
int pc_mgpu_attach(int rank, int world, const unsigned char* handles) {
    try {
        if (!g_mgpu.local) throw pc::ArgError("polychord_b200: pc_mgpu_attach before pc_mgpu_create");
        if (rank < 0 || rank >= world || world > MAX_RANKS) throw pc::ArgError("polychord_b200: bad rank/world");
        for (int q = 0; q < world; ++q) {
            if (q == rank) { g_mgpu.peer[q] = g_mgpu.local; continue; }
            cudaIpcMemHandle_t h;
            std::memcpy(&h, handles + (size_t)q * 64, 64);
            PC_CUDA(cudaIpcOpenMemHandle(&g_mgpu.peer[q], h, cudaIpcMemLazyEnablePeerAccess));
        }
        g_mgpu.rank = rank; g_mgpu.world = world;
        return 0;
    } catch (const std::exception& ex) {
        return fail(-7, ex.what());
    }
}
int pc_mgpu_destroy(void) {
    for (int q = 0; q < g_mgpu.world; ++q)
        if (q != g_mgpu.rank && g_mgpu.peer[q]) cudaIpcCloseMemHandle(g_mgpu.peer[q]);
    if (g_mgpu.local) cudaFree(g_mgpu.local);
    g_mgpu = MgpuCtx();
    return 0;
}

int pc_last_run_info(pc_run_info* out) {
    *out = g_last;
    return g_last.status;
}

int pc_register_device_likelihood(pc_loglikelihood_t fn, int kind, const double* params, int nparams) {
    if (!fn || kind < 0 || kind > PC_LIKE_CORR_GAUSSIAN) return -1;
    like_registry()[(void*)fn] = LikeReg{kind, std::vector<double>(params, params + (params ? nparams : 0))};
    return 0;
}
int pc_register_device_prior(pc_prior_t fn, int kind, const double* params, int nparams) {
    if (!fn || kind != PC_PRIOR_UNIFORM) return -1;
    prior_registry()[(void*)fn] = PriorReg{kind, std::vector<double>(params, params + (params ? nparams : 0))};
    return 0;
}
void pc_clear_registrations(void) {
    like_registry().clear();
    prior_registry().clear();
}

// ---- ready-made host callbacks (also usable by CPU-side callers; they evaluate the same formulas) ----
static std::vector<double>& reg_params(void* fn) {
    static std::vector<double> empty;
    auto it = like_registry().find(fn);
    return it == like_registry().end() ? empty : it->second.params;
}
double pc_gaussian_loglikelihood(double* theta, int nDims, double* phi, int nDerived) {
    const std::vector<double>& q = reg_params((void*)pc_gaussian_loglikelihood);
    double norm = 0, chi = 0, r2 = 0;
    for (int i = 0; i < nDims; ++i) {
        double mu = 0.5, sg = 0.1;
        if ((int)q.size() >= 2 * nDims) { mu = q[i]; sg = q[nDims + i]; } else if (q.size() == 2) { mu = q[0]; sg = q[1]; }
        norm += std::log(sg) + LOG_TWO_PI / 2.0;
        chi += (theta[i] - mu) / sg * ((theta[i] - mu) / sg);
        r2 += (theta[i] - mu) * (theta[i] - mu);
    }
    if (nDerived >= 1) phi[0] = std::sqrt(r2);
    if (nDerived >= 2) phi[1] = std::log(std::pow(phi[0], (double)nDims) * std::pow(std::sqrt(M_PI), (double)nDims) / std::tgamma(1.0 + nDims / 2.0));
    for (int i = 2; i < nDerived; ++i) phi[i] = 0.0;
    return -norm - chi / 2.0;
}
double pc_rastrigin_loglikelihood(double* theta, int nDims, double* phi, int nDerived) {
    double s = 0;
    for (int i = 0; i < nDims; ++i) s += std::log(4991.21750) + theta[i] * theta[i] - 10.0 * std::cos(2.0 * M_PI * theta[i]);
    for (int i = 0; i < nDerived; ++i) phi[i] = 0.0;
    return -s;
}
double pc_corr_gaussian_loglikelihood(double* theta, int nDims, double* phi, int nDerived) {
    const std::vector<double>& q = reg_params((void*)pc_corr_gaussian_loglikelihood);
    const int D = nDims;
    for (int i = 0; i < nDerived; ++i) phi[i] = 0.0;
    if ((int)q.size() != D + D * D + 1) return -1e30;
    double s = 0;
    for (int r = 0; r < D; ++r) {
        double y = 0;
        for (int c = 0; c < D; ++c) y += q[D + r + (size_t)c * D] * (theta[c] - q[c]);
        s += (theta[r] - q[r]) * y;
    }
    return -(D * LOG_TWO_PI + q[D + D * D]) / 2.0 - s / 2.0;
}
void pc_unit_prior(double* cube, double* theta, int nDims) {
    for (int i = 0; i < nDims; ++i) theta[i] = cube[i];
}
void pc_uniform_prior(double* cube, double* theta, int nDims) {
    auto it = prior_registry().find((void*)pc_uniform_prior);
    for (int i = 0; i < nDims; ++i) {
        double lo = 0, hi = 1;
        if (it != prior_registry().end() && (int)it->second.params.size() == 2 * nDims) { lo = it->second.params[i]; hi = it->second.params[nDims + i]; }
        theta[i] = lo + (hi - lo) * cube[i];
    }
}

static int run_common(const pc_settings* s, const ModelSpec& ms, int nruns, const int* seeds, pc_dumper_t dumper,
                      pc_run_info* out) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_abort = 0;
    auto t0 = std::chrono::steady_clock::now();
    Engine e;
    e.setup(*s, ms, nruns, seeds);
    auto t1 = std::chrono::steady_clock::now();
    e.run(dumper, out);
    if (std::getenv("PC_DEBUG"))
        std::fprintf(stderr, "[pc dbg] setup %.3f ms, run %.3f ms\n", std::chrono::duration<double, std::milli>(t1 - t0).count(),
                     std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count());
    const double wall = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    for (int r = 0; r < nruns; ++r) out[r].wall_ms = wall;  // entry to return, set-up included
    g_last = out[0];
    return 0;
}

int pc_run(const pc_settings* s, int like_kind, const double* like_params, int n_like_params, const double* prior_params,
           int n_prior_params, pc_dumper_t dumper, pc_run_info* out) {
    try {
        ModelSpec ms;
        ms.like_kind = like_kind;
        if (like_params) ms.like_params.assign(like_params, like_params + n_like_params);
        if (prior_params) ms.prior_params.assign(prior_params, prior_params + n_prior_params);
        int seed = s->seed;
        return run_common(s, ms, 1, &seed, dumper, out);
    } catch (const ArgError& ex) {
        return fail(-2, ex.what());
    } catch (const std::exception& ex) {   // these entry points are bound from C / ctypes: nothing may unwind through them
        return fail(-3, ex.what());
    } catch (...) {
        return fail(-3, "unknown exception");
    }
}

int pc_run_ensemble(const pc_settings* s, int like_kind, const double* like_params, int n_like_params,
                    const double* prior_params, int n_prior_params, int nruns, const int* seeds, pc_run_info* out) {
    try {
        ModelSpec ms;
        ms.like_kind = like_kind;
        if (like_params) ms.like_params.assign(like_params, like_params + n_like_params);
        if (prior_params) ms.prior_params.assign(prior_params, prior_params + n_prior_params);
        return run_common(s, ms, nruns, seeds, nullptr, out);
    } catch (const ArgError& ex) {
        return fail(-2, ex.what());
    } catch (const std::exception& ex) {   // these entry points are bound from C / ctypes: nothing may unwind through them
        return fail(-3, ex.what());
    } catch (...) {
        return fail(-3, "unknown exception");
    }
}

// ---- probes -------------------------------------------------------------------------------
struct ProbeCtx {
    DevModel dm;
    Layout L;
    ModelSpec ms;
};
static void probe_setup(ProbeCtx& c, const pc_settings* s, int like_kind, const double* lp, int nlp, const double* pp,
                        int npp) {
    device_check();
    c.ms.like_kind = like_kind;
    if (lp) c.ms.like_params.assign(lp, lp + nlp);
    if (pp) c.ms.prior_params.assign(pp, pp + npp);
    build_dev_model(*s, c.ms, c.dm, g_stream);
    c.L = make_layout(*s, c.ms, c.dm, std::max(1, std::min(8, g_opt.warps_per_cta)), true,
                      g_opt.dense > 0 && like_kind != PC_LIKE_CORR_GAUSSIAN && g_grade_dims.size() <= 1);
    set_smem(c.L.fn, c.L.smem);
}

int pc_slice_chains(const pc_settings* s, int like_kind, const double* like_params, int n_like_params,
                    const double* prior_params, int n_prior_params, int nchains, const double* seed_points,
                    const double* cholesky, const double* logL, const unsigned long long* uid, double* babies_out,
                    long long* nlike_out) {
    try {
        ProbeCtx c;
        probe_setup(c, s, like_kind, like_params, n_like_params, prior_params, n_prior_params);
        const KParams& k = c.L.kp;
        cudaStream_t st = g_stream;
        const int T = k.cp.T, D = k.cp.D, R = k.cp.R;
        DevArr<double> d_seed((size_t)nchains * T), d_chol((size_t)D * D), d_logL(nchains),
            d_babies((size_t)nchains * R * T), d_nh;
        DevArr<unsigned long long> d_uid(nchains);
        DevArr<long long> d_nlike(nchains);
        d_seed.upload(seed_points, (size_t)nchains * T, st);
        d_chol.upload(cholesky, (size_t)D * D, st);
        d_logL.upload(logL, nchains, st);
        d_uid.upload(uid, nchains, st);
        d_babies.zero(st);
        int W = c.L.W;
        int blocks = std::min(1024, (nchains + W - 1) / W);
        const bool dense = k.dense != 0;
        if (dense) {  // the dense chain phase: 32/G chains per warp, their slice records in global scratch
            blocks = std::min(296, (nchains + W * (32 / c.L.fn.G) - 1) / (W * (32 / c.L.fn.G)));
            d_nh.alloc((size_t)nchains * R * dense_slb(c.L.fn.G * c.L.fn.DPL));
        } else if (!k.nh_in_smem) d_nh.alloc((size_t)blocks * W * R * k.cp.LD);
        KParams kp = k;
        unsigned seed = (unsigned)s->seed;
        const double *a_seed = d_seed.p, *a_chol = d_chol.p, *a_logL = d_logL.p;
        const unsigned long long* a_uid = d_uid.p;
        double *a_babies = d_babies.p, *a_nh = d_nh.p;
        long long* a_nlike = d_nlike.p;
        void* args[] = {&kp, &nchains, &a_seed, &a_chol, &a_logL, &a_uid, &seed, &a_babies, &a_nlike, &a_nh};
        PC_CUDA(cudaLaunchKernel(dense ? c.L.fn.slice_dense : c.L.fn.slice, dim3(blocks), dim3(W * 32), args, c.L.smem, st));
        PC_CUDA(cudaGetLastError());
        d_babies.download(babies_out, (size_t)nchains * R * T, st);
        d_nlike.download(nlike_out, nchains, st);
        PC_CUDA(cudaStreamSynchronize(st));
        return 0;
    } catch (const std::exception& ex) {
        return fail(-3, ex.what());
    }
}

int pc_calculate_points(const pc_settings* s, int like_kind, const double* like_params, int n_like_params,
                        const double* prior_params, int n_prior_params, double* records, int npts) {
    try {
        ProbeCtx c;
        probe_setup(c, s, like_kind, like_params, n_like_params, prior_params, n_prior_params);
        const KParams& k = c.L.kp;
        cudaStream_t st = g_stream;
        const int T = k.cp.T;
        DevArr<double> d_rec((size_t)npts * T);
        DevArr<int> d_n(1);
        d_rec.upload(records, (size_t)npts * T, st);
        d_n.zero(st);
        int W = c.L.W;
        const int npt = 32 / c.L.fn.G;
        int blocks = std::max(1, std::min(1024, (npts + W * npt - 1) / (W * npt)));
        KParams kp = k;
        double* a_rec = d_rec.p;
        int* a_n = d_n.p;
        void* args[] = {&kp, &a_rec, &npts, &a_n};
        PC_CUDA(cudaLaunchKernel(c.L.fn.calc, dim3(blocks), dim3(W * 32), args, c.L.smem, st));
        PC_CUDA(cudaGetLastError());
        int n = 0;
        d_rec.download(records, (size_t)npts * T, st);
        d_n.download(&n, 1, st);
        PC_CUDA(cudaStreamSynchronize(st));
        return n;
    } catch (const std::exception& ex) {
        return fail(-3, ex.what());
    }
}

int pc_device_directions(int nDims, int num_repeats, unsigned seed, unsigned long long uid, double* out) {
    try {
        device_check();
        const int D = nDims, R = num_repeats, LD = D | 1;
        cudaStream_t st = g_stream;
        DevArr<double> d_nh((size_t)R * LD), d_out((size_t)R * D);
        size_t smem = chain_scratch_bytes(D, R, LD, false, LIKE_GAUSSIAN, 1);
        PC_CUDA(cudaFuncSetAttribute(pc_directions_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));
        pc_directions_kernel<<<1, 32, smem, st>>>(D, R, LD, seed, uid, d_nh.p, d_out.p);
        PC_CUDA(cudaGetLastError());
        d_out.download(out, (size_t)R * D, st);
        PC_CUDA(cudaStreamSynchronize(st));
        return 0;
    } catch (const std::exception& ex) {
        return fail(-3, ex.what());
    }
}

int pc_device_philox(const unsigned* ctr, const unsigned* key, unsigned* out4) {
    try {
        device_check();
        DevArr<unsigned> c(4), k(2), o(4);
        c.upload(ctr, 4, g_stream); k.upload(key, 2, g_stream);
        pc_philox_kernel<<<1, 1, 0, g_stream>>>(c.p, k.p, o.p);
        PC_CUDA(cudaGetLastError());
        o.download(out4, 4, g_stream);
        PC_CUDA(cudaStreamSynchronize(g_stream));
        return 0;
    } catch (const std::exception& ex) {
        return fail(-3, ex.what());
    }
}
int pc_device_uniforms(unsigned seed, unsigned tag, unsigned long long uid, unsigned a0, unsigned b, int n, double* out) {
    try {
        device_check();
        DevArr<double> o(n);
        pc_uniforms_kernel<<<(n + 255) / 256, 256, 0, g_stream>>>(seed, tag, uid, a0, b, n, o.p);
        PC_CUDA(cudaGetLastError());
        o.download(out, n, g_stream);
        PC_CUDA(cudaStreamSynchronize(g_stream));
        return 0;
    } catch (const std::exception& ex) {
        return fail(-3, ex.what());
    }
}
int pc_device_inv_normal_cdf(const double* p, int n, double* out) {
    try {
        device_check();
        DevArr<double> i(n), o(n);
        i.upload(p, n, g_stream);
        pc_inv_normal_kernel<<<(n + 255) / 256, 256, 0, g_stream>>>(i.p, n, o.p);
        PC_CUDA(cudaGetLastError());
        o.download(out, n, g_stream);
        PC_CUDA(cudaStreamSynchronize(g_stream));
        return 0;
    } catch (const std::exception& ex) {
        return fail(-3, ex.what());
    }
}
// state = {logZ, logZ2, logX, logZX, logXX}; deaths with live counts n_start, n_start-1, ...
int pc_device_evidence(double* state, const double* logLs, int count, int n_start, double* logw_out) {
    try {
        device_check();
        DevRun h;
        std::memset(&h, 0, sizeof(h));
        h.logZ = state[0]; h.logZ2 = state[1]; h.logX = state[2]; h.logZX = state[3]; h.logXX = state[4];
        DevArr<DevRun> st(1);
        DevArr<double> l(count), w(count);
        st.upload(&h, 1, g_stream);
        l.upload(logLs, count, g_stream);
        size_t smem = (size_t)(64 + count) * 8;
        PC_CUDA(cudaFuncSetAttribute(pc_evidence_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));
        pc_evidence_kernel<<<1, 256, smem, g_stream>>>(st.p, l.p, count, n_start, w.p);
        PC_CUDA(cudaGetLastError());
        st.download(&h, 1, g_stream);
        w.download(logw_out, count, g_stream);
        PC_CUDA(cudaStreamSynchronize(g_stream));
        state[0] = h.logZ; state[1] = h.logZ2; state[2] = h.logX; state[3] = h.logZX; state[4] = h.logXX;
        return 0;
    } catch (const std::exception& ex) {
        return fail(-3, ex.what());
    }
}
// measured FP64 FMA throughput in TFLOP/s (best of five launches timed with CUDA events)
double pc_measure_fp64_tflops(void) {
    try {
        device_check();
        int dev = 0, sms = 0;
        PC_CUDA(cudaGetDevice(&dev));
        PC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const int blocks = sms * 8, threads = 256, iters = 4096;
        DevArr<double> out((size_t)blocks * threads);
        cudaEvent_t e0, e1;
        PC_CUDA(cudaEventCreate(&e0));
        PC_CUDA(cudaEventCreate(&e1));
        double best = 0.0;
        for (int rep = 0; rep < 6; ++rep) {
            PC_CUDA(cudaEventRecord(e0, g_stream));
            pc_fp64_peak_kernel<<<blocks, threads, 0, g_stream>>>(out.p, iters, 0.999999, 1e-9);
            PC_CUDA(cudaEventRecord(e1, g_stream));
            PC_CUDA(cudaEventSynchronize(e1));
            float ms = 0;
            PC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            const double tf = 2.0 * 8.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
            if (rep > 0) best = std::max(best, tf);
        }
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        return best;
    } catch (const std::exception& ex) {
        fail(-3, ex.what());
        return 0.0;
    }
}

int pc_device_cholesky(const double* a, int D, double* L_out) {
    try {
        device_check();
        DevArr<double> da((size_t)D * D), dl((size_t)D * D), dl2((size_t)D * D);
        DevArr<int> fb(2);
        da.upload(a, (size_t)D * D, g_stream);
        // both factorisations of the engine: the one-warp form (per-cluster factors) and the CTA-wide form (the run
        // kernel's update); they must agree to rounding, the result returned is the run kernel's
        pc_cholesky_kernel<<<1, 32, 0, g_stream>>>(da.p, dl.p, D, fb.p, 0);
        PC_CUDA(cudaGetLastError());
        pc_cholesky_kernel<<<1, 256, 0, g_stream>>>(da.p, dl2.p, D, fb.p + 1, 1);
        PC_CUDA(cudaGetLastError());
        int f[2] = {0, 0};
        std::vector<double> Lw((size_t)D * D);
        dl.download(Lw.data(), (size_t)D * D, g_stream);
        dl2.download(L_out, (size_t)D * D, g_stream);
        fb.download(f, 2, g_stream);
        PC_CUDA(cudaStreamSynchronize(g_stream));
        if (f[0] != f[1]) throw pc::RunError("polychord_b200: the two Cholesky kernels disagree on the fallback");
        for (size_t e = 0; e < (size_t)D * D; ++e)
            if (std::fabs(Lw[e] - L_out[e]) > 1e-11 * (1.0 + std::fabs(Lw[e])))
                throw pc::RunError("polychord_b200: the two Cholesky kernels disagree");
        return f[1];
    } catch (const std::exception& ex) {
        return fail(-3, ex.what());
    }
}

// NN_clustering of m points (row-major m x D) as the engine's update runs it: device neighbour lists + host union-find
int pc_cluster_points(const double* points, int m, int D, int* labels_out) {
    try {
        device_check();
        Engine e;
        e.stream = g_stream;
        DevArr<double> d((size_t)m * D);
        d.upload(points, (size_t)m * D, g_stream);
        std::vector<int> lab;
        const int num = e.cluster_labels_of(d.p, m, D, D, lab);
        std::copy(lab.begin(), lab.end(), labels_out);
        return num;
    } catch (const std::exception& ex) {
        return fail(-3, ex.what());
    }
}

// ---- output files (host only; no device needed) -------------------------------------------------
void pc_format_e24(double v, char* out25) { format_e24(v, out25); }

// Clusters of the last run with do_clustering.  Returns their number (the clusters alive at the end of sampling first, in
// label order, then the deleted ones) and *nactive; rows (2 doubles per cluster: log<Z_p>, log<Z_p^2> -- attributed local
// evidences) and uid (the cluster's identity) when not null.
int pc_last_clusters(int* nactive, double* rows, int* uid) {
    if (nactive) *nactive = g_cluster_report.nactive;
    if (rows) std::copy(g_cluster_report.rows.begin(), g_cluster_report.rows.end(), rows);
    if (uid) std::copy(g_cluster_report.uid.begin(), g_cluster_report.uid.end(), uid);
    return (int)g_cluster_report.uid.size();
}
// ... the identity every dead point's cluster had at its death (up to cap entries; returns the number of dead points), and
// for every identity the one it was split from (-1: the initial cluster; returns the number of identities)
long long pc_last_dead_clusters(int* out, long long cap) {
    if (out) std::copy(g_cluster_report.point_uid.begin(), g_cluster_report.point_uid.begin() + std::min<long long>(cap, (long long)g_cluster_report.point_uid.size()), out);
    return (long long)g_cluster_report.point_uid.size();
}
int pc_last_cluster_tree(int* parent_out) {
    if (parent_out) std::copy(g_cluster_report.parent.begin(), g_cluster_report.parent.end(), parent_out);
    return (int)g_cluster_report.parent.size();
}

// Host-only: parse a resume file in the reference's text layout (pc_resume_text.h).  ints[8] = {nDims, nDerived, ndead,
// ncluster, ncluster_dead, live points of all clusters, phantoms of all clusters, likelihood calls}; reals[6] = {logZ,
// logZ2, log sum X_p, last update volume, lowest live logL, highest live logL}.  When out_path is given and the file
// holds one active cluster, what was read is written back in the same layout (the writer used for "resume_text").
int pc_resume_text_probe(const char* path, long long* ints, double* reals, const char* out_path) {
    try {
        pc::RefResume r;
        pc::read_reference_resume(path, r);
        long long n = 0, nph = 0, nl = 0;
        for (int v : r.nlive) n += v;
        for (int v : r.nphantom) nph += v;
        for (long long v : r.nlike) nl += v;
        const long long T = 2LL * r.nDims + r.nDerived + 2;
        double lo = INFINITY, hi = -INFINITY, mx = -INFINITY, sx = 0.0;
        for (long long i = 0; i < n; ++i) { lo = std::min(lo, r.live[(size_t)i * T + T - 1]); hi = std::max(hi, r.live[(size_t)i * T + T - 1]); }
        for (double x : r.logXp) mx = std::max(mx, x);
        for (double x : r.logXp) sx += std::exp(x - mx);
        if (ints) { ints[0] = r.nDims; ints[1] = r.nDerived; ints[2] = r.ndead; ints[3] = r.ncluster; ints[4] = r.ncluster_dead; ints[5] = n; ints[6] = nph; ints[7] = nl; }
        if (reals) { reals[0] = r.logZ; reals[1] = r.logZ2; reals[2] = mx + std::log(sx); reals[3] = r.logX_last_update; reals[4] = lo; reals[5] = hi; }
        if (out_path && *out_path) pc::write_reference_resume(out_path, r);
        return 0;
    } catch (const pc::ArgError& ex) {   // (a probe: a malformed file is reported, never fatal)
        std::fprintf(stderr, " polychord_b200: %s\n", ex.what());
        return -2;
    } catch (const std::exception& ex) {
        std::fprintf(stderr, " polychord_b200: %s\n", ex.what());
        return -3;
    }
}
static int write_files_common(const char* base_dir, const char* file_root, int flags, int nDims, int nDerived, long long ndead,
                              const double* dead_rows, const double* dead_logw, int nlive, const double* live_rows, double logZ,
                              double logZerr, long long nlike, int num_repeats, double compression_factor, unsigned seed,
                              const BoostedRows* boosted) {
    try {
        FileOpts o;
        o.enabled = true;
        o.base_dir = base_dir; o.file_root = file_root;
        o.write_stats = flags & 1; o.write_live = flags & 2; o.write_dead = flags & 4; o.write_prior = flags & 8;
        o.posteriors = flags & 16; o.equals = flags & 32;
        o.num_repeats = num_repeats; o.compression_factor = compression_factor; o.seed = seed;
        FileState st;
        return write_run_files(o, st, nDims, nDerived, ndead, dead_rows, dead_logw, nlive, live_rows, logZ, logZerr, nlike, true, boosted);
    } catch (const std::exception& ex) {
        return fail(-1, ex.what());
    }
}
int pc_write_files(const char* base_dir, const char* file_root, int flags, int nDims, int nDerived, long long ndead,
                   const double* dead_rows, const double* dead_logw, int nlive, const double* live_rows, double logZ,
                   double logZerr, long long nlike, int num_repeats, double compression_factor, unsigned seed) {
    return write_files_common(base_dir, file_root, flags, nDims, nDerived, ndead, dead_rows, dead_logw, nlive, live_rows, logZ,
                              logZerr, nlike, num_repeats, compression_factor, seed, nullptr);
}
int pc_write_files_boosted(const char* base_dir, const char* file_root, int flags, int nDims, int nDerived, long long ndead,
                           const double* dead_rows, const double* dead_logw, int nlive, const double* live_rows, double logZ,
                           double logZerr, long long nlike, int num_repeats, double compression_factor, unsigned seed,
                           long long nboosted, const double* boosted_rows, const double* boosted_logw,
                           const long long* boosted_after) {
    BoostedRows br;
    br.n = nboosted; br.rows = boosted_rows; br.logw = boosted_logw; br.after = boosted_after;
    for (long long i = 1; i < nboosted; ++i)
        if (boosted_after[i] < boosted_after[i - 1]) return fail(-1, "pc_write_files_boosted: boosted_after must ascend");
    return write_files_common(base_dir, file_root, flags, nDims, nDerived, ndead, dead_rows, dead_logw, nlive, live_rows, logZ,
                              logZerr, nlike, num_repeats, compression_factor, seed, nboosted > 0 ? &br : nullptr);
}

// cube_samples: the initial live points of the next run through polychord_c_interface (see the header)
int pc_set_initial_live(const double* cube_samples, int npoints, int nDims) {
    g_init_cubes.clear();
    g_init_n = g_init_D = 0;
    if (!cube_samples || npoints <= 0) return 0;
    if (nDims < 1) return -1;
    g_init_cubes.assign(cube_samples, cube_samples + (size_t)npoints * nDims);
    g_init_n = npoints; g_init_D = nDims;
    return 0;
}

// Host-only entry points of the maximiser (csrc/pc_maximise.cpp), see the header.
int pc_maximise(pc_loglikelihood_t loglikelihood, pc_prior_t prior, int nDims, int nDerived, double logzero,
                const double* live_records, int nlive, int posterior, double* point_out) {
    if (!loglikelihood || !prior || !live_records || !point_out || nDims < 1 || nDerived < 0) return -1;
    return do_maximisation(loglikelihood, prior, nDims, nDerived, logzero, live_records, nlive, posterior != 0, point_out) ? 0 : 1;
}
double pc_prior_log_density(pc_prior_t prior, const double* cube, int nDims) { return maximise_dXdtheta(prior, cube, nDims); }

// maximise (maximiser.F90:31-77) after a run through polychord_c_interface: likelihood and posterior maxima from the
// live points the run ended with, the posterior mean when posteriors are kept, <root>.maximum.
static void run_maximiser(pc_loglikelihood_t ll, pc_prior_t prior, const pc_settings& s, const FileOpts& fo, const pc_run_info& info,
                          int feedback) {
    const int D = s.nDims, P = s.nDerived, T = 2 * D + P + 2;
    if (!ll || !prior) throw pc::ArgError("polychord_b200: maximise needs the loglikelihood and prior callbacks");
    if (g_final_live.n < 1) throw pc::RunError("polychord_b200: maximise: the run left no live points");
    std::vector<double> mp(T, 0.0), pp(T, 0.0), mean(T, 0.0);
    if (feedback >= 1) std::printf("-------------------------------------\nMaximising Likelihood\n");
    const bool ok1 = do_maximisation(ll, prior, D, P, s.logzero, g_final_live.recs.data(), g_final_live.n, false, mp.data());
    if (feedback >= 1) std::printf("-------------------------------------\nMaximising Posterior\n");
    const bool ok2 = do_maximisation(ll, prior, D, P, s.logzero, g_final_live.recs.data(), g_final_live.n, true, pp.data());
    if ((!ok1 || !ok2) && feedback >= 0) std::printf(" Could not construct simplex\n");
    const double dx = maximise_dXdtheta(prior, pp.data(), D);
    const bool with_mean = s.posteriors && info.ndead > 0;
    if (with_mean) {  // mean (read_write.F90:912-934) over the weighted posterior the run ended with
        const DumpMirror& mr = g_mirror;
        const int npars = D + P + 2;
        double logwsum = s.logzero;
        std::vector<double> mu(D + P, 0.0);
        auto add = [&](const double* x, double lw) {
            const double m = std::max(logwsum, lw);
            logwsum = m + std::log(std::exp(logwsum - m) + std::exp(lw - m));
            const double f = std::exp(lw - logwsum);
            for (int k = 0; k < D + P; ++k) mu[k] += f * (x[k] - mu[k]);
        };
        for (long long i = 0; i < info.ndead; ++i) add(&mr.rows[(size_t)i * npars], mr.logw[(size_t)i]);
        for (size_t j = 0; j < mr.boost_logw.size(); ++j) add(&mr.boost_rows[j * npars], mr.boost_logw[j]);
        std::copy(mu.begin(), mu.end(), mean.begin() + D);
        std::vector<double> phi(std::max(P, 1), 0.0);
        mean[T - 1] = ll(mean.data() + D, D, phi.data(), P);
    }
    write_max_file(fo.base_dir + "/" + fo.file_root + ".maximum", D, P, mp.data(), pp.data(), dx, with_mean ? mean.data() : nullptr);
    std::fflush(stdout);
}

// ==========================================================================================
// Drop-in boundary
// ==========================================================================================
void polychord_c_interface(pc_loglikelihood_t loglikelihood, pc_prior_t prior, pc_dumper_t dumper, int nlive,
                           int num_repeats, int nprior, int nfail, pc_bool do_clustering, int feedback,
                           double precision_criterion, double logzero, int max_ndead, double boost_posterior,
                           pc_bool posteriors, pc_bool equals, pc_bool cluster_posteriors, pc_bool write_resume,
                           pc_bool write_paramnames, pc_bool read_resume, pc_bool write_stats, pc_bool write_live,
                           pc_bool write_dead, pc_bool write_prior, pc_bool maximise, double compression_factor,
                           pc_bool synchronous, int nDims, int nDerived, char* base_dir, char* file_root, int nGrade,
                           double* grade_frac, int* grade_dims, int n_nlives, double* loglikes, int* nlives, int seed,
                           int* comm) {
    (void)write_paramnames; (void)synchronous; (void)grade_frac; (void)comm;
    std::memset(&g_last, 0, sizeof(g_last));
    g_abort = 0;
    if (dumper == default_dumper) dumper = nullptr;   // the facade's no-op: nothing to hand the dead points to
    pc_settings s;
    std::memset(&s, 0, sizeof(s));
    s.nDims = nDims; s.nDerived = nDerived; s.nlive = nlive; s.num_repeats = num_repeats; s.nprior = nprior; s.nfail = nfail;
    s.do_clustering = do_clustering; s.feedback = feedback; s.precision_criterion = precision_criterion; s.logzero = logzero;
    s.max_ndead = max_ndead; s.boost_posterior = boost_posterior; s.posteriors = posteriors; s.equals = equals;
    s.cluster_posteriors = cluster_posteriors; s.compression_factor = compression_factor;
    if (seed < 0) {  // random_utils.F90:60-75: seed from the system clock
        seed = (int)(std::chrono::high_resolution_clock::now().time_since_epoch().count() & 0x7fffffff);
    }
    s.seed = seed;
    // fast/slow grades: repeats per grade as generate.F90:303-309 sets them -- grade_frac > 1 everywhere: the
    // fractions ARE the repeat counts; otherwise num_repeats for the slowest grade and in proportion to
    // grade_frac for the others.  The reference also scales by the measured likelihood speed of each grade
    // (time_speeds, generate.F90:330-455); this engine has no partial evaluations to time, so the speeds are equal.
    struct GradesGuard { ~GradesGuard() { g_grade_dims.clear(); g_grade_reps.clear(); } } grades_guard;
    if (nGrade > 1) {
        if (nGrade > MAX_GRADES || !grade_dims || !grade_frac) { fail(-4, "at most 8 parameter grades are supported"); return; }
        int sd = 0;
        bool counts = true;
        for (int g = 0; g < nGrade; ++g) { sd += grade_dims[g]; counts = counts && grade_frac[g] > 1.0; }
        if (sd != nDims) { fail(-4, "grade_dims must sum to nDims"); return; }
        g_grade_dims.assign(grade_dims, grade_dims + nGrade);
        g_grade_reps.assign(nGrade, 0);
        int total = 0;
        for (int g = 0; g < nGrade; ++g) {
            g_grade_reps[g] = counts ? (int)grade_frac[g]
                                     : (g == 0 ? num_repeats : (int)std::lround(grade_frac[g] / grade_frac[0] * num_repeats));
            total += g_grade_reps[g];
        }
        s.num_repeats = total;
    } else if (grade_dims && grade_dims[0] != nDims) {
        fail(-4, "grade_dims must sum to nDims");
        return;
    } else if (nGrade == 1 && grade_frac && grade_frac[0] > 1.0) {
        s.num_repeats = (int)grade_frac[0];   // generate.F90:304-309: every grade_frac > 1 -- the fractions are the repeat counts
    }
    // dynamic nlive (settings%loglikes / settings%nlives, interfaces.F90:416-422): the batched form of replace_point's
    // rule (run_time_info.f90:766-777), phase S1
    struct DynGuard { ~DynGuard() { g_dyn_loglikes.clear(); g_dyn_nlives.clear(); } } dyn_guard;
    if (n_nlives > 0) {
        if (n_nlives > MAX_DYN || !loglikes || !nlives) { fail(-4, "the nlives schedule holds at most 16 entries"); return; }
        g_dyn_loglikes.assign(loglikes, loglikes + n_nlives);
        g_dyn_nlives.assign(nlives, nlives + n_nlives);
    }
    auto li = like_registry().find((void*)loglikelihood);
    auto pi = prior_registry().find((void*)prior);
    ModelSpec ms;
    bool have_like = true, have_prior = true;
    if (li != like_registry().end()) { ms.like_kind = li->second.kind; ms.like_params = li->second.params; }
    else if (loglikelihood == pc_gaussian_loglikelihood) ms.like_kind = PC_LIKE_GAUSSIAN;
    else if (loglikelihood == pc_rastrigin_loglikelihood) ms.like_kind = PC_LIKE_RASTRIGIN;
    else have_like = false;
    if (pi != prior_registry().end()) ms.prior_params = pi->second.params;
    else if (prior == pc_unit_prior || prior == pc_uniform_prior || prior == default_prior) {}
    else have_prior = false;
    struct HostGuard {  // callbacks may throw through the engine (e.g. a Python exception): always clear
        ~HostGuard() { g_host_ll = nullptr; g_host_prior = nullptr; }
    } host_guard;
    if (!have_like || !have_prior) {
        // a callback without a device form: the generic path -- the slice state machines stay on the device, the
        // calling thread makes the prior + likelihood calls in lock step (pc_hostchain.cuh)
        if (!loglikelihood || !prior) { fail(-5, "loglikelihood and prior callbacks must not be NULL"); return; }
        ms = ModelSpec();
        ms.like_kind = PC_LIKE_HOST;
        g_host_ll = loglikelihood;
        g_host_prior = prior;
    }
    // output files (read_write.F90); the reference halts when base_dir is missing (read_write.F90:28-38)
    FileOpts fo;
    fo.write_stats = write_stats; fo.write_live = write_live; fo.write_dead = write_dead; fo.write_prior = write_prior;
    fo.posteriors = posteriors; fo.equals = equals;
    fo.enabled = write_stats || write_live || write_dead || write_prior || posteriors || equals;
    fo.base_dir = base_dir ? std::string(base_dir) : std::string(".");   // read up to the first NUL only (utils.F90:787-801)
    fo.file_root = file_root ? std::string(file_root) : std::string("test");
    fo.compression_factor = compression_factor; fo.num_repeats = s.num_repeats; fo.seed = (unsigned)seed; fo.logzero = logzero;
    if (fo.enabled) {
        FILE* probe = std::fopen((fo.base_dir + "/.pc_probe").c_str(), "w");
        if (!probe) {
            fail(-1, "base_dir '" + fo.base_dir + "' does not exist or is not writable (the reference halts here too: read_write.F90:28-38)");
            return;
        }
        std::fclose(probe);
        std::remove((fo.base_dir + "/.pc_probe").c_str());
    }
    struct ResumeGuard {
        ResumeGuard(bool w, bool r, const std::string& path) { g_resume.write = w; g_resume.read = r; g_resume.path = path; }
        ~ResumeGuard() { g_resume.write = g_resume.read = false; }
    } resume_guard(write_resume, read_resume, fo.base_dir + "/" + fo.file_root + ".resume");
    struct FilesGuard {  // exception-transparent: a throwing dumper must not leave the next run writing files
        FilesGuard(const FileOpts& o) { g_files = o; g_fstate = FileState(); }
        ~FilesGuard() { g_files.enabled = false; }
    } files_guard(fo);
    pc_run_info info;
    struct InitLiveGuard {   // cube_samples are for this run only
        InitLiveGuard(pc_loglikelihood_t l, pc_prior_t p) { g_init_ll = l; g_init_prior = p; }
        ~InitLiveGuard() { g_init_ll = nullptr; g_init_prior = nullptr; g_init_n = 0; g_init_cubes.clear(); }
    } init_live_guard(loglikelihood, prior);
    struct FinalLiveGuard {
        FinalLiveGuard(bool w) { g_final_live.want = w; g_final_live.n = 0; }
        ~FinalLiveGuard() { g_final_live.want = false; g_final_live.recs.clear(); }
    } final_live_guard(maximise);
    try {
        run_common(&s, ms, 1, &seed, dumper, &info);
        if (maximise) run_maximiser(loglikelihood, prior, s, fo, info, feedback);
    } catch (const ArgError& ex) {   // the engine's own failures only: a caller's exception passes through untouched
        fail(-2, ex.what());
        return;
    } catch (const RunError& ex) {
        fail(-3, ex.what());
        return;
    } catch (const std::bad_alloc&) {
        fail(-3, "out of host memory");
        return;
    }
    if (feedback >= 1) {
        std::printf(" log(Z) = %12.5f +/- %8.5f   ndead = %lld  nlike = %lld  [B200 engine: K=%d, %d launches, %.2f ms]\n",
                    info.logZ, info.logZerr, info.ndead, info.nlike, info.batch_K, info.kernel_launches, info.device_ms);
        std::fflush(stdout);
    }
}

// The phantoms the last run promoted to posterior samples (boost_posterior with posteriors or equals set), in the
// order they are written to the posterior files.  Returns their number; fills up to cap entries.
long long pc_last_boosted(double* rows, long long* dead_index, double* logw, long long cap, int npars) {
    const DumpMirror& mr = g_mirror;
    const long long nb = (long long)mr.boost_logw.size();
    if (nb && (long long)mr.boost_rows.size() != nb * npars) return -1;
    for (long long i = 0; i < std::min(nb, cap); ++i) {
        std::copy(mr.boost_rows.begin() + (size_t)i * npars, mr.boost_rows.begin() + (size_t)(i + 1) * npars, rows + (size_t)i * npars);
        dead_index[i] = mr.boost_dead[(size_t)i];
        logw[i] = mr.boost_logw[(size_t)i];
    }
    return nb;
}

// Host-only: the prior transform of an .ini file's parameter block (hypercube_to_physical, priors.f90:494-556) applied
// to one cube point.  Returns 0, -6 when the file cannot be read or parsed, -7 when nDims is not its parameter count.
int pc_ini_prior_transform(const char* inifile, const double* cube, double* theta, int nDims) {
    try {
        const IniConfig c = parse_ini(inifile ? std::string(inifile) : std::string());
        if ((int)c.params.size() != nDims) return -7;
        ini_prior_transform(c, cube, theta);
    } catch (const std::exception& ex) {
        std::fprintf(stderr, "pc_ini_prior_transform: %s\n", ex.what());
        return -6;
    }
    return 0;
}

static IniConfig g_ini;   // the .ini run in flight: its prior transform is reached through a plain C callback
static void ini_prior_callback(double* cube, double* theta, int nDims) {
    (void)nDims;
    ini_prior_transform(g_ini, cube, theta);
}
static void ini_noop_dumper(int, int, int, double*, double*, double*, double, double) {}

// Replaces interfaces.F90:496-519 + the generic run_polychord_ini path (:142-283): settings, parameters and priors
// come from the file (ini.f90), setup_loglikelihood() is called once, then the run goes through
// polychord_c_interface.  A likelihood pointer registered with a device form together with all-uniform priors runs
// on the GPU; anything else takes the host-callback path with the file's prior transform.
void polychord_c_interface_ini(pc_loglikelihood_t loglikelihood, void (*setup_loglikelihood)(void), char* inifile,
                               int* comm) {
    try {
        g_ini = parse_ini(inifile ? std::string(inifile) : std::string());
    } catch (const std::exception& ex) {
        fail(-6, ex.what());
        return;
    }
    if (setup_loglikelihood) setup_loglikelihood();
    const int D = (int)g_ini.params.size(), P = (int)g_ini.derived.size();
    bool all_uniform = true;
    std::vector<double> box(2 * D);
    for (int i = 0; i < D; ++i) {
        all_uniform = all_uniform && g_ini.params[i].prior_type == 1;
        box[i] = g_ini.params[i].params[0];
        box[D + i] = g_ini.params[i].params[1];
    }
    if (all_uniform) pc_register_device_prior(ini_prior_callback, PC_PRIOR_UNIFORM, box.data(), 2 * D);
    else prior_registry().erase((void*)ini_prior_callback);
    if (g_ini.write_paramnames) {  // write_paramnames_file, read_write.F90:963-1011: "<name>   <latex>", derived names starred
        FILE* f = std::fopen((g_ini.base_dir + "/" + g_ini.file_root + ".paramnames").c_str(), "w");
        if (f) {
            for (const auto& p : g_ini.params) std::fprintf(f, "%s   %s\n", p.name.c_str(), p.latex.c_str());
            for (const auto& d : g_ini.derived) std::fprintf(f, "%s*   %s\n", d.first.c_str(), d.second.c_str());
            std::fclose(f);
        }
    }
    std::vector<char> base(g_ini.base_dir.begin(), g_ini.base_dir.end()), root(g_ini.file_root.begin(), g_ini.file_root.end());
    base.push_back(0); root.push_back(0);
    polychord_c_interface(loglikelihood, ini_prior_callback, ini_noop_dumper, g_ini.nlive, g_ini.num_repeats, g_ini.nprior,
                          g_ini.nfail, g_ini.do_clustering, g_ini.feedback, g_ini.precision_criterion, g_ini.logzero,
                          g_ini.max_ndead, g_ini.boost_posterior, g_ini.posteriors, g_ini.equals, g_ini.cluster_posteriors,
                          g_ini.write_resume, false, g_ini.read_resume, g_ini.write_stats, g_ini.write_live, g_ini.write_dead,
                          g_ini.write_prior, g_ini.maximise, g_ini.compression_factor, true, D, P, base.data(), root.data(),
                          (int)g_ini.grade_dims.size(), g_ini.grade_frac.data(), g_ini.grade_dims.data(), 0, nullptr, nullptr,
                          g_ini.seed, comm);
}

}  // extern "C"
