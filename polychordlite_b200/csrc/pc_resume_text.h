// pc_resume_text.h -- the reference's TEXT resume file.
//
// Replaces src/polychord/read_write.F90: write_resume_file (:219-288) and read_resume_file (:384-476) with their
// helpers (:130-217, :296-382): section headers "=== ... ===", integers in (nI12), reals in (nE24.15E3), a
// "-----" line in front of every cluster of a three-dimensional array.  pypolychord writes the same layout for a run
// that starts from the caller's live points (polychord.py:650-789, _make_resume_file).  Host-only code.
//
// The engine's own state is one cluster (its evidence is global); a file with several clusters is read as one: the
// volumes and cross moments are summed (X = sum X_p, <Z X> = sum <Z X_p>, <X^2> = sum <X_p X_q>), the live points and
// phantoms of all clusters are concatenated.
#pragma once
#include <string>
#include <vector>

namespace pc {

struct RefResume {
    int nDims = 0, nDerived = 0;
    long long ndead = 0;
    int ncluster = 1, ncluster_dead = 0;
    std::vector<int> grade_dims, num_repeats;
    std::vector<long long> nlike;          // per grade
    std::vector<int> nlive, nphantom;      // per cluster
    double logZ = 0, logZ2 = 0, thin_posterior = 0;
    std::vector<double> logLp, logXp, logZXp, logZp, logZp2, logZpXp, maxlogweight;   // per cluster
    std::vector<double> logXpXq;           // ncluster x ncluster, row-major as read line by line
    double logX_last_update = 0;
    std::vector<double> logZp_dead, logZp2_dead, maxlogweight_dead;
    std::vector<double> covmat, cholesky;  // ncluster x (nDims x nDims), each matrix column-major (a line is a column)
    // records [cube | theta | derived | birth contour | logL] (settings.f90:163-182), nTotal = 2 nDims + nDerived + 2 each
    std::vector<double> live;              // clusters one after the other
    std::vector<double> dead, logweights;
    std::vector<double> phantom;           // clusters one after the other
    // posterior points [logX, logL, log-weight, logZ, theta, derived] (settings.f90:189-204) of the global list
    std::vector<double> posterior_global;
    long long nposterior_global = 0;
};

// true when the file starts like the reference's text layout ("=== Number of dimensions ===")
bool is_reference_resume(const std::string& path);
// throws pc::ArgError on a malformed file
void read_reference_resume(const std::string& path, RefResume& r);
// one active cluster, no dead clusters (what this engine's state is); throws pc::RunError when the file cannot be written
void write_reference_resume(const std::string& path, const RefResume& r);

}  // namespace pc
