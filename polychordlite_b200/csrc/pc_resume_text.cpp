// pc_resume_text.cpp -- reader and writer of the reference's text resume file (see pc_resume_text.h).  Host-only.
#include "pc_resume_text.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <numeric>

#include "pc_errors.h"
#include "pc_files.h"

namespace pc {

bool is_reference_resume(const std::string& path) {
    std::ifstream f(path);
    std::string line;
    return f && std::getline(f, line) && line.rfind("===", 0) == 0;
}

namespace {

// The reads of read_write.F90:296-382: a header line is skipped, then fixed-width fields; here the fields are taken
// as blank-separated numbers (I12 and E24.15E3 always leave a blank in front of a value that fits).
struct Reader {
    std::ifstream f;
    std::string path, line;
    long long lineno = 0;
    explicit Reader(const std::string& p) : f(p), path(p) {
        if (!f) throw ArgError("polychord_b200: cannot open the resume file " + p);
    }
    [[noreturn]] void bad(const std::string& what) {
        throw ArgError("polychord_b200: resume file " + path + ", line " + std::to_string(lineno) + ": " + what);
    }
    bool next() { ++lineno; return (bool)std::getline(f, line); }
    void header() {
        if (!next() || line.rfind("===", 0) != 0) bad("a section header (=== ... ===) is expected");
    }
    void separator() {
        if (!next() || line.rfind("---", 0) != 0) bad("a cluster separator (-----) is expected");
    }
    void numbers(std::vector<double>& out, long long want) {   // want < 0: whatever the line holds (at least one)
        if (!next()) bad("the file ends inside a section");
        const char* p = line.c_str();
        long long got = 0;
        for (;;) {
            char* e = nullptr;
            const double v = std::strtod(p, &e);
            if (e == p) break;
            out.push_back(v);
            ++got;
            p = e;
        }
        while (*p == ' ' || *p == '\t' || *p == '\r') ++p;
        if (*p) bad("not a number: '" + std::string(p).substr(0, 24) + "'");
        if (want >= 0 ? got != want : got < 1) bad(std::to_string(want) + " values expected, " + std::to_string(got) + " found");
    }
    long long integer() {
        header();
        std::vector<double> v;
        numbers(v, 1);
        return (long long)std::llround(v[0]);
    }
    template <class I>
    void integers(std::vector<I>& out, long long n, bool any_count = false) {
        header();
        out.clear();
        if (n <= 0) return;
        std::vector<double> v;
        numbers(v, any_count ? -1 : n);
        for (double x : v) out.push_back((I)std::llround(x));
    }
    double real() {
        header();
        std::vector<double> v;
        numbers(v, 1);
        return v[0];
    }
    void reals(std::vector<double>& out, long long n) {
        header();
        out.clear();
        if (n > 0) numbers(out, n);
    }
    void reals2(std::vector<double>& out, long long n1, long long n2) {   // n2 lines of n1 values
        header();
        out.clear();
        for (long long i = 0; i < n2; ++i) numbers(out, n1);
    }
    template <class I>
    void reals3(std::vector<double>& out, long long n1, const std::vector<I>& counts) {   // per cluster: separator, counts[c] lines
        header();
        out.clear();
        for (size_t c = 0; c < counts.size(); ++c) {
            separator();
            for (long long i = 0; i < (long long)counts[c]; ++i) numbers(out, n1);
        }
    }
};

struct Writer {
    FILE* f;
    std::vector<char> big;
    explicit Writer(const std::string& path) : f(std::fopen(path.c_str(), "w")), big(1 << 20) {
        if (!f) throw RunError("polychord_b200: cannot write " + path);
        std::setvbuf(f, big.data(), _IOFBF, big.size());
    }
    ~Writer() { if (f) std::fclose(f); }
    void header(const char* s) { std::fprintf(f, "%s\n", s); }
    void integer(long long v, const char* s) { header(s); std::fprintf(f, "%12lld\n", v); }
    template <class I>
    void integers(const std::vector<I>& v, const char* s) {
        header(s);
        if (v.empty()) return;
        for (I x : v) std::fprintf(f, "%12lld", (long long)x);
        std::fputc('\n', f);
    }
    void row(const double* v, long long n) {
        char b[25];
        for (long long i = 0; i < n; ++i) { format_e24(v[i], b); std::fwrite(b, 1, 24, f); }
        std::fputc('\n', f);
    }
    void real(double v, const char* s) { header(s); row(&v, 1); }
    void reals(const std::vector<double>& v, const char* s) { header(s); if (!v.empty()) row(v.data(), (long long)v.size()); }
    void reals2(const double* v, long long n1, long long n2, const char* s) {
        if (s) header(s);
        for (long long i = 0; i < n2; ++i) row(v + i * n1, n1);
    }
    void separator() { std::fprintf(f, "---------------------------------------\n"); }
};

}  // namespace

void read_reference_resume(const std::string& path, RefResume& r) {
    Reader in(path);
    r = RefResume();
    r.nDims = (int)in.integer();
    r.nDerived = (int)in.integer();
    if (r.nDims < 1 || r.nDims > 4096 || r.nDerived < 0 || r.nDerived > 65536) in.bad("nDims / nDerived out of range");
    r.ndead = in.integer();
    r.ncluster = (int)in.integer();
    r.ncluster_dead = (int)in.integer();
    if (r.ndead < 0 || r.ncluster < 1 || r.ncluster > 100000 || r.ncluster_dead < 0 || r.ncluster_dead > 100000) in.bad("counts out of range");
    r.nposterior_global = in.integer();
    const long long nequals_global = in.integer();
    const int ngrades = (int)in.integer();
    if (ngrades < 1 || ngrades > 64) in.bad("number of grades out of range");
    in.integers(r.grade_dims, ngrades);
    in.integers(r.num_repeats, ngrades);
    in.integers(r.nlike, ngrades, true);   // (pypolychord writes one total whatever the number of grades, polychord.py:718)
    std::vector<long long> nposterior, nequals, imin, nposterior_dead, nequals_dead;
    in.integers(r.nlive, r.ncluster);
    in.integers(r.nphantom, r.ncluster);
    in.integers(nposterior, r.ncluster);
    in.integers(nequals, r.ncluster);
    in.integers(imin, r.ncluster);
    in.integers(nposterior_dead, r.ncluster_dead);
    in.integers(nequals_dead, r.ncluster_dead);
    for (int c = 0; c < r.ncluster; ++c)
        if (r.nlive[c] < 0 || r.nphantom[c] < 0 || nposterior[c] < 0 || nequals[c] < 0) in.bad("negative point count");
    if (r.nposterior_global < 0 || nequals_global < 0) in.bad("negative point count");
    r.logZ = in.real();
    r.logZ2 = in.real();
    r.thin_posterior = in.real();
    in.reals(r.logLp, r.ncluster);
    in.reals(r.logXp, r.ncluster);
    r.logX_last_update = in.real();
    in.reals(r.logZXp, r.ncluster);
    in.reals(r.logZp, r.ncluster);
    in.reals(r.logZp2, r.ncluster);
    in.reals(r.logZpXp, r.ncluster);
    in.reals2(r.logXpXq, r.ncluster, r.ncluster);
    in.reals(r.maxlogweight, r.ncluster);
    in.reals(r.logZp_dead, r.ncluster_dead);
    in.reals(r.logZp2_dead, r.ncluster_dead);
    in.reals(r.maxlogweight_dead, r.ncluster_dead);
    const long long D = r.nDims, nTotal = 2 * D + r.nDerived + 2, npost = 4 + D + r.nDerived, np = 2 + D + r.nDerived;
    const std::vector<long long> dd((size_t)r.ncluster, D);
    in.reals3(r.covmat, D, dd);
    in.reals3(r.cholesky, D, dd);
    in.reals3(r.live, nTotal, r.nlive);
    in.reals2(r.dead, nTotal, r.ndead);
    in.reals(r.logweights, r.ndead);
    in.reals3(r.phantom, nTotal, r.nphantom);
    std::vector<double> skip;
    in.reals3(skip, npost, nposterior);
    in.reals3(skip, npost, nposterior_dead);
    in.reals2(r.posterior_global, npost, r.nposterior_global);
    in.reals3(skip, np, nequals);
    in.reals3(skip, np, nequals_dead);
    in.reals2(skip, np, nequals_global);
}

void write_reference_resume(const std::string& path, const RefResume& r) {
    if (r.ncluster != 1 || r.ncluster_dead != 0) throw RunError("polychord_b200: the text resume writer takes one active cluster");
    Writer out(path);
    const long long D = r.nDims, nTotal = 2 * D + r.nDerived + 2, npost = 4 + D + r.nDerived;
    out.integer(r.nDims, "=== Number of dimensions ===");
    out.integer(r.nDerived, "=== Number of derived parameters ===");
    out.integer(r.ndead, "=== Number of dead points/iterations ===");
    out.integer(1, "=== Number of clusters ===");
    out.integer(0, "=== Number of dead clusters ===");
    out.integer(r.nposterior_global, "=== Number of global weighted posterior points ===");
    out.integer(0, "=== Number of global equally weighted posterior points ===");
    out.integer((long long)r.grade_dims.size(), "=== Number of grades ===");
    out.integers(r.grade_dims, "=== positions of grades ===");
    out.integers(r.num_repeats, "=== Number of repeats ===");
    out.integers(r.nlike, "=== Number of likelihood calls ===");
    out.integers(r.nlive, "=== Number of live points in each cluster ===");
    out.integers(r.nphantom, "=== Number of phantom points in each cluster ===");
    out.integers(std::vector<int>{0}, "=== Number of weighted posterior points in each cluster ===");
    out.integers(std::vector<int>{0}, "=== Number of equally weighted posterior points in each cluster ===");
    {   // position (1-based) of the lowest live point
        long long imin = 1;
        for (long long i = 1; i < r.nlive[0]; ++i)
            if (r.live[(size_t)i * nTotal + nTotal - 1] < r.live[(size_t)(imin - 1) * nTotal + nTotal - 1]) imin = i + 1;
        out.integers(std::vector<long long>{imin}, "=== Minimum loglikelihood positions ===");
    }
    out.integers(std::vector<int>{}, "=== Number of weighted posterior points in each dead cluster ===");
    out.integers(std::vector<int>{}, "=== Number of equally weighted posterior points in each dead cluster ===");
    out.real(r.logZ, "=== global evidence -- log(<Z>) ===");
    out.real(r.logZ2, "=== global evidence^2 -- log(<Z^2>) ===");
    out.real(r.thin_posterior, "=== posterior thin factor ===");
    out.reals(r.logLp, "=== local loglikelihood bounds ===");
    out.reals(r.logXp, "=== local volume -- log(<X_p>) ===");
    out.real(r.logX_last_update, "=== last update volume ===");
    out.reals(r.logZXp, "=== global evidence volume cross correlation -- log(<ZX_p>) ===");
    out.reals(r.logZp, "=== local evidence -- log(<Z_p>) ===");
    out.reals(r.logZp2, "=== local evidence^2 -- log(<Z_p^2>) ===");
    out.reals(r.logZpXp, "=== local evidence volume cross correlation -- log(<Z_pX_p>) ===");
    out.reals2(r.logXpXq.data(), 1, 1, "=== local volume cross correlation -- log(<X_pX_q>) ===");
    out.reals(r.maxlogweight, "=== maximum log weights -- log(w_p) ===");
    out.reals({}, "=== local dead evidence -- log(<Z_p>) ===");
    out.reals({}, "=== local dead evidence^2 -- log(<Z_p^2>) ===");
    out.reals({}, "=== maximum dead log weights -- log(w_p) ===");
    out.header("=== covariance matrices ===");
    out.separator();
    out.reals2(r.covmat.data(), D, D, nullptr);
    out.header("=== cholesky decompositions ===");
    out.separator();
    out.reals2(r.cholesky.data(), D, D, nullptr);
    out.header("=== live points ===");
    out.separator();
    out.reals2(r.live.data(), nTotal, r.nlive[0], nullptr);
    out.reals2(r.dead.data(), nTotal, r.ndead, "=== dead points ===");
    out.reals(r.logweights, "=== logweights of dead points ===");
    out.header("=== phantom points ===");
    out.separator();
    out.reals2(r.phantom.data(), nTotal, r.nphantom[0], nullptr);
    out.header("=== weighted posterior points ===");
    out.separator();
    out.header("=== dead weighted posterior points ===");
    out.reals2(r.posterior_global.data(), npost, r.nposterior_global, "=== global weighted posterior points ===");
    out.header("=== equally weighted posterior points ===");
    out.separator();
    out.header("=== dead equally weighted posterior points ===");
    out.header("=== global equally weighted posterior points ===");
}

}  // namespace pc
