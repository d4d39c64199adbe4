// pc_files.h -- the run's output files in the reference's on-disk formats.
//
// Replaces src/polychord/read_write.F90: write_stats_file (:809-910), write_dead_points (:679-716),
// write_phys_live_points (:621-677), write_posterior_file (:479-612, global files) and the prior_info lines of
// generate.F90:274-279.  Numbers are written in Fortran's E24.15E3 edit descriptor (utils.F90:19), which is what
// anesthetic / getdist / pypolychord.output.PolyChordOutput parse.  Host-only code: the engine hands it the same
// host arrays the dumper receives.
#pragma once
#include <chrono>
#include <string>

namespace pc {

struct FileOpts {
    bool enabled = false;  // any write_* flag set
    std::string base_dir, file_root;
    bool write_stats = false, write_live = false, write_dead = false, write_prior = false;
    bool posteriors = false, equals = false;
    double compression_factor = 0.36787944117144233;
    int num_repeats = 1;
    unsigned seed = 0;
    double logzero = -1e30;
    double min_interval_s = 0.5;  // intermediate (per-update) rewrites are rate-limited; the final write always happens
};

struct FileState {
    long long nlike_last = 0;  // likelihood calls at the previous stats write (the file reports calls since then)
    bool written = false;
    std::chrono::steady_clock::time_point last;
};

// value in Fortran E24.15E3 form, 24 characters, no terminator needed by callers (out must hold 25 bytes)
void format_e24(double v, char* out);

// boost_posterior: phantoms promoted to posterior samples; row i is written after dead row after[i] - 1
struct BoostedRows {
    long long n = 0;
    const double* rows = nullptr;    // [theta(D), phi(P), birth, logL]
    const double* logw = nullptr;    // unnormalised posterior log-weight
    const long long* after = nullptr;
};

// Clusters of a run with do_clustering: the "Local evidences" table of <root>.stats
// (read_write.F90:858-872) and, with cluster_posteriors, the files clusters/<root>_<i>.txt and
// clusters/<root>_<i>_equal_weights.txt (read_write.F90:527-607).  Clusters are listed active first, then deleted;
// frac[u] is the share log(n_i / n) of its parent's evidence identity u received when it was split off (0 for the initial
// cluster): a cluster's posterior holds its own dead points and, scaled by those shares, the ones of its ancestors (the
// reference copies the parent's posterior to every piece with the weights reduced in that proportion,
// run_time_info.f90:433-441, 497-503).
struct ClusterRows {
    int n = 0, nactive = 0;
    const double* logZp = nullptr;   // n
    const double* logZp2 = nullptr;  // n
    const int* uid = nullptr;        // n: identity of every listed cluster
    int nuid = 0;
    const int* parent = nullptr;     // nuid: identity a cluster was split from (-1: none)
    const double* frac = nullptr;    // nuid
    const int* point_uid = nullptr;  // per dead point: identity of its cluster at its death
    bool cluster_posteriors = false;
};

// Writes every requested file under base_dir.  dead_rows/live_rows: rows [theta(D), phi(P), birth, logL];
// dead_logw[i] = log-weight + logL of dead point i (unnormalised posterior log-weight).
// Returns the number of files written; throws std::runtime_error when a file cannot be opened.
int write_run_files(const FileOpts& o, FileState& st, int D, int P, long long ndead, const double* dead_rows,
                    const double* dead_logw, int nlive, const double* live_rows, double logZ, double logZerr,
                    long long nlike, bool final_call, const BoostedRows* boosted = nullptr, const ClusterRows* clusters = nullptr);

// generate.F90:274-279
void write_prior_info(const FileOpts& o, long long nprior, long long ndiscarded);

}  // namespace pc
