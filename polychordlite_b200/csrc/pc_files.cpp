// pc_files.cpp -- output files in the reference's formats (see pc_files.h).  Host-only.
#include "pc_files.h"

#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <sys/stat.h>
#include <algorithm>
#include "pc_errors.h"
#include <vector>

#include "pc_device.cuh"  // the counter-based uniform stream (equally weighted posterior thinning)

namespace pc {

// Fortran E24.15E3: right-justified in 24 columns, "0.ddddddddddddddd" mantissa in [0.1, 1), exponent letter, sign
// and three exponent digits (utils.F90:19, DB_FMT).  The 15 significant digits are the ones "%.14E" produces,
// shifted by one place.
void format_e24(double v, char* out) {
    std::memset(out, ' ', 24);
    out[24] = 0;
    if (std::isnan(v)) { std::memcpy(out + 21, "NaN", 3); return; }
    if (std::isinf(v)) {
        if (v > 0) std::memcpy(out + 16, "Infinity", 8); else std::memcpy(out + 15, "-Infinity", 9);
        return;
    }
    char buf[40];
    auto res = std::to_chars(buf, buf + sizeof(buf), std::fabs(v), std::chars_format::scientific, 14);
    *res.ptr = 0;  // d.dddddddddddddde[+-]XX[X]
    const char* e = std::strchr(buf, 'e');
    int ex = std::atoi(e + 1);
    if (v != 0.0) ex += 1;
    // [0] blank, [1] sign, "0." at 2-3, the 15 digits at 4..18, 'E' at 19, exponent sign at 20, three digits at 21..23
    out[1] = std::signbit(v) ? '-' : ' ';
    out[2] = '0';
    out[3] = '.';
    out[4] = buf[0];
    std::memcpy(out + 5, buf + 2, 14);
    out[19] = 'E';
    out[20] = ex < 0 ? '-' : '+';
    const int ax = ex < 0 ? -ex : ex;
    out[21] = (char)('0' + (ax / 100) % 10);
    out[22] = (char)('0' + (ax / 10) % 10);
    out[23] = (char)('0' + ax % 10);
}

namespace {

struct Out {
    FILE* f;
    std::vector<char> big;  // stdio buffer of this file (must outlive the stream)
    explicit Out(const std::string& path) : f(std::fopen(path.c_str(), "w")), big(1 << 20) {
        if (!f) throw pc::RunError("polychord_b200: cannot open " + path + " for writing");
        std::setvbuf(f, big.data(), _IOFBF, big.size());
    }
    Out(const Out&) = delete;
    Out& operator=(const Out&) = delete;
    ~Out() { if (f) std::fclose(f); }
    void num(double v) { char b[25]; format_e24(v, b); std::fwrite(b, 1, 24, f); }
    void nl() { std::fputc('\n', f); }
    void line(const char* s) { std::fputs(s, f); std::fputc('\n', f); }
};

std::string root_of(const FileOpts& o) { return o.base_dir + "/" + o.file_root; }

double host_logaddexp(double a, double b) {
    if (a < b) std::swap(a, b);
    if (b == -std::numeric_limits<double>::infinity()) return a;
    return a + std::log1p(std::exp(b - a));
}

}  // namespace

void write_prior_info(const FileOpts& o, long long nprior, long long ndiscarded) {
    if (!o.enabled || !o.write_prior) return;
    FILE* f = std::fopen((root_of(o) + ".prior_info").c_str(), "a");
    if (!f) throw pc::RunError("polychord_b200: cannot open " + root_of(o) + ".prior_info");
    std::fprintf(f, "nprior = %12lld\n", nprior);
    std::fprintf(f, "ndiscarded = %12lld\n", ndiscarded);
    std::fclose(f);
}

// log of the sum of the local evidences (= the global log<Z>: every death belongs to exactly one cluster)
static double logZ_raw_of(const ClusterRows* c) {
    double m = -std::numeric_limits<double>::infinity(), s2 = 0.0;
    for (int i = 0; i < c->n; ++i) m = std::max(m, c->logZp[i]);
    for (int i = 0; i < c->n; ++i) s2 += std::exp(c->logZp[i] - m);
    return m + std::log(s2);
}

int write_run_files(const FileOpts& o, FileState& st, int D, int P, long long ndead, const double* dead_rows,
                    const double* dead_logw, int nlive, const double* live_rows, double logZ, double logZerr,
                    long long nlike, bool final_call, const BoostedRows* boosted, const ClusterRows* clusters) {
    if (!o.enabled) return 0;
    const auto now = std::chrono::steady_clock::now();
    if (!final_call && st.written && std::chrono::duration<double>(now - st.last).count() < o.min_interval_s) return 0;
    const int npars = D + P + 2, np = D + P;
    const std::string root = root_of(o);
    int files = 0;

    if (o.write_dead) {  // write_dead_points, read_write.F90:679-716
        {
            Out w(root + "_dead.txt");
            for (long long i = 0; i < ndead; ++i) {
                const double* r = dead_rows + (size_t)i * npars;
                w.num(r[np + 1]);
                for (int k = 0; k < np; ++k) w.num(r[k]);
                w.nl();
            }
        }
        {
            Out w(root + "_dead-birth.txt");
            for (long long i = 0; i < ndead; ++i) {
                const double* r = dead_rows + (size_t)i * npars;
                for (int k = 0; k < np; ++k) w.num(r[k]);
                w.num(r[np + 1]);
                w.num(r[np]);
                w.nl();
            }
        }
        files += 2;
    }
    if (o.write_live) {  // write_phys_live_points, read_write.F90:621-677
        Out a(root + "_phys_live.txt"), b(root + "_phys_live-birth.txt");
        for (int i = 0; i < nlive; ++i) {
            const double* r = live_rows + (size_t)i * npars;
            for (int k = 0; k < np; ++k) { a.num(r[k]); b.num(r[k]); }
            a.num(r[np + 1]); a.nl();
            b.num(r[np + 1]); b.num(r[np]); b.nl();
        }
        files += 2;
    }

    // posterior weights: log w_i + log L_i relative to the largest one (maximum weight 1.0, read_write.F90:565-566)
    if (o.write_prior) {
        // write_prior_file (read_write.F90:721-752): the points drawn from the prior, rows [1, -2 logL, theta, phi].  The
        // reference writes them once, right after GenerateLivePoints; here they are recognised by their birth contour
        // (logzero) among the dead and the live points, so the file is complete at every rewrite.
        Out w(root + "_prior.txt");
        auto rows_of = [&](const double* rows, long long n) {
            for (long long i = 0; i < n; ++i) {
                const double* r = rows + (size_t)i * npars;
                if (r[np] > o.logzero) continue;
                w.num(1.0);
                w.num(-2.0 * r[np + 1]);
                for (int k = 0; k < np; ++k) w.num(r[k]);
                w.nl();
            }
        };
        rows_of(dead_rows, ndead);
        rows_of(live_rows, nlive);
        ++files;
    }

    // The posterior samples in file order: the dead points and, after the deaths of the update that removed them, the
    // phantoms boost_posterior promoted (update_posteriors appends the stack update by update, run_time_info.f90:1036-1061)
    struct PostRef { const double* row; double lw; uint64_t uid; };
    std::vector<PostRef> post;
    if (o.posteriors || o.equals || o.write_stats) {
        const long long nb = boosted ? boosted->n : 0;
        post.reserve((size_t)(ndead + nb));
        long long j = 0;
        for (long long i = 0; i < ndead; ++i) {
            for (; j < nb && boosted->after[j] <= i; ++j)
                post.push_back({boosted->rows + (size_t)j * npars, boosted->logw[j], (1ull << 40) + (uint64_t)j});
            post.push_back({dead_rows + (size_t)i * npars, dead_logw[i], (uint64_t)i});
        }
        for (; j < nb; ++j) post.push_back({boosted->rows + (size_t)j * npars, boosted->logw[j], (1ull << 40) + (uint64_t)j});
    }
    double wmax = -std::numeric_limits<double>::infinity();
    for (const PostRef& q : post) wmax = std::max(wmax, q.lw);
    long long nposterior = 0, nequals = 0;
    for (const PostRef& q : post)
        if (std::exp(q.lw - wmax) > 0.0) ++nposterior;
    if (o.posteriors) {  // <root>.txt: weight, -2 logL, theta, phi (written to _temp, then renamed: read_write.F90:600-611)
        const std::string tmp = root + "_temp.txt";
        {
            Out w(tmp);
            for (const PostRef& q : post) {
                const double wgt = std::exp(q.lw - wmax);
                if (!(wgt > 0.0)) continue;
                const double* r = q.row;
                w.num(wgt);
                w.num(-2.0 * r[np + 1]);
                for (int k = 0; k < np; ++k) w.num(r[k]);
                w.nl();
            }
        }
        std::rename(tmp.c_str(), (root + ".txt").c_str());
        ++files;
    }
    if (o.equals || o.write_stats) {
        // equally weighted posterior: point i is kept with probability w_i / w_max (the net effect of the
        // reference's incremental thinning in update_posteriors, run_time_info.f90:955-1066); the draw is
        // addressed by the point's index, so successive rewrites agree on the points they share
        const std::string tmp = root + "_equal_weights_temp.txt";
        std::vector<const double*> keep;
        for (const PostRef& q : post) {
            const double u = uniform(o.seed, TAG_POST, q.uid, 0u, 0u);
            if (u < std::exp(q.lw - wmax)) keep.push_back(q.row);
        }
        nequals = (long long)keep.size();
        if (o.equals) {
            {
                Out w(tmp);
                for (const double* r : keep) {
                    w.num(1.0);
                    w.num(-2.0 * r[np + 1]);
                    for (int k = 0; k < np; ++k) w.num(r[k]);
                    w.nl();
                }
            }
            std::rename(tmp.c_str(), (root + "_equal_weights.txt").c_str());
            ++files;
        }
    }

    // cluster posteriors (read_write.F90:527-607): file i belongs to the cluster with the i-th largest local evidence; its
    // points are the dead points of the cluster and of its ancestors (scaled by the shares the pieces received)
    if (clusters && clusters->cluster_posteriors && clusters->n > 0 && clusters->point_uid && (o.posteriors || o.equals)) {
        const std::string cdir = o.base_dir + "/clusters";
        ::mkdir(cdir.c_str(), 0777);
        std::vector<int> ordering(clusters->n);
        for (int c = 0; c < clusters->n; ++c) ordering[c] = c;
        std::stable_sort(ordering.begin(), ordering.end(), [&](int x, int y) { return clusters->logZp[x] > clusters->logZp[y]; });
        std::vector<double> scale((size_t)clusters->nuid);
        for (int i = 0; i < clusters->n; ++i) {
            const int c = ordering[i];
            // log share of every identity's points in this cluster: 0 for itself, the product of the shares down the line
            // of descent for an ancestor, nothing for anybody else
            std::fill(scale.begin(), scale.end(), -std::numeric_limits<double>::infinity());
            double acc = 0.0;
            for (int u = clusters->uid[c]; u >= 0 && u < clusters->nuid; u = clusters->parent[u]) {
                scale[u] = acc;
                acc += clusters->frac[u];
            }
            double cmax = -std::numeric_limits<double>::infinity();
            for (long long k = 0; k < ndead; ++k) {
                const int u = clusters->point_uid[k];
                if (u >= 0 && u < clusters->nuid && std::isfinite(scale[u])) cmax = std::max(cmax, dead_logw[k] + scale[u]);
            }
            const std::string base = cdir + "/" + o.file_root + "_" + std::to_string(i + 1);
            const double rel = std::exp(clusters->logZp[c] - logZ_raw_of(clusters));   // exp(logZp - logZ): the cluster's share of the evidence
            if (o.posteriors) {
                Out w(base + ".txt");
                for (long long k = 0; k < ndead; ++k) {
                    const int u = clusters->point_uid[k];
                    if (u < 0 || u >= clusters->nuid || !std::isfinite(scale[u])) continue;
                    const double wgt = std::exp(dead_logw[k] + scale[u] - cmax) * rel;
                    if (!(wgt > 0.0)) continue;
                    const double* r = dead_rows + (size_t)k * npars;
                    w.num(wgt);
                    w.num(-2.0 * r[np + 1]);
                    for (int q = 0; q < np; ++q) w.num(r[q]);
                    w.nl();
                }
                ++files;
            }
            if (o.equals) {
                Out w(base + "_equal_weights.txt");
                for (long long k = 0; k < ndead; ++k) {
                    const int u = clusters->point_uid[k];
                    if (u < 0 || u >= clusters->nuid || !std::isfinite(scale[u])) continue;
                    const double uu = uniform(o.seed, TAG_POST, (uint64_t)k, 1u + (unsigned)c, 0u);
                    if (!(uu < std::exp(dead_logw[k] + scale[u] - cmax))) continue;
                    const double* r = dead_rows + (size_t)k * npars;
                    w.num(rel);
                    w.num(-2.0 * r[np + 1]);
                    for (int q = 0; q < np; ++q) w.num(r[q]);
                    w.nl();
                }
                ++files;
            }
        }
    }

    if (o.write_stats) {  // write_stats_file, read_write.F90:809-910 (one cluster: evidence is kept globally)
        Out w(root + ".stats");
        char a[25], b[25];
        w.line("Evidence estimates:");
        w.line("===================");
        w.line("  - The evidence Z is a log-normally distributed, with location and scale parameters mu and sigma.");
        w.line("  - We denote this as log(Z) = mu +/- sigma.");
        w.line("");
        w.line("Global evidence:");
        w.line("----------------");
        w.line("");
        format_e24(logZ, a); format_e24(logZerr, b);
        std::fprintf(w.f, "log(Z)       = %s +/- %s\n", a, b);
        w.line("");
        w.line("");
        w.line("Local evidences:");
        w.line("----------------");
        w.line("");
        int ncl_active = nlive > 0 ? 1 : 0, ncl_total = 1;
        if (clusters && clusters->n > 0) {
            // calculate_logZ_estimate (run_time_info.f90:652-678) per cluster; label "log(Z_p)" padded as read_write.F90:861-871
            ncl_active = nlive > 0 ? clusters->nactive : 0;
            ncl_total = clusters->n;
            for (int c = 0; c < clusters->n; ++c) {
                const double lz = std::max(-std::numeric_limits<double>::max(), 2 * clusters->logZp[c] - 0.5 * clusters->logZp2[c]);
                const double var = clusters->logZp2[c] - 2 * clusters->logZp[c];
                char lbl[32];
                std::snprintf(lbl, sizeof(lbl), "log(Z_%d)", c + 1);
                const int digits = (int)std::strlen(lbl) - 7;
                format_e24(lz, a); format_e24(std::sqrt(std::fabs(var)), b);
                std::fprintf(w.f, "%s%*s= %s +/- %s%s\n", lbl, std::max(0, 6 - digits), "", a, b,
                             (nlive > 0 && c < clusters->nactive) ? " (Still Active)" : "");
            }
            format_e24(logZ, a); format_e24(logZerr, b);
        } else if (nlive > 0) std::fprintf(w.f, "log(Z_1)     = %s +/- %s (Still Active)\n", a, b);
        else std::fprintf(w.f, "log(Z_1)     = %s +/- %s\n", a, b);
        w.line("");
        w.line("");
        w.line("Run-time information:");
        w.line("---------------------");
        w.line("");
        std::fprintf(w.f, " ncluster:   %8d /%8d\n", ncl_active, ncl_total);
        std::fprintf(w.f, " nposterior: %8lld\n", nposterior);
        std::fprintf(w.f, " nequals:    %8lld\n", nequals);
        std::fprintf(w.f, " ndead:      %8lld\n", ndead);
        std::fprintf(w.f, " nlive:      %8d\n", nlive);
        if (nlike < 100000000LL) std::fprintf(w.f, " nlike:      %8lld\n", nlike);
        else std::fprintf(w.f, " nlike:      ********\n");  // Fortran I8 overflow
        const double since = (double)(nlike - st.nlike_last);
        if (nlive > 0) {
            const double update_files = -(double)nlive * std::log(o.compression_factor);
            std::fprintf(w.f, " <nlike>:    %8.2f   (%8.2f per slice )\n", since / update_files,
                         since / ((double)o.num_repeats * update_files));
        } else {
            std::fprintf(w.f, " <nlike>:    %8.2f   (%8.2f per slice )\n", 0.0, 0.0);
        }
        if (o.posteriors) {  // weighted mean / variance, streaming log-space updates (read_write.F90:912-961)
            w.line("");
            w.line("");
            w.line("Dim No.       Mean        Sigma");
            std::vector<double> mu(np, 0.0), mu_old(np), logS(np, o.logzero);
            double logwsum = o.logzero;
            for (const PostRef& q : post) {
                if (!(std::exp(q.lw - wmax) > 0.0)) continue;
                const double* x = q.row;
                const double lw = q.lw;
                mu_old = mu;
                logwsum = host_logaddexp(logwsum, lw);
                const double f = std::exp(lw - logwsum);
                bool all_pos = true;
                for (int k = 0; k < np; ++k) {
                    mu[k] = mu_old[k] + f * (x[k] - mu_old[k]);
                    all_pos = all_pos && (x[k] - mu_old[k]) * (x[k] - mu[k]) > 0.0;
                }
                if (all_pos)
                    for (int k = 0; k < np; ++k) logS[k] = host_logaddexp(logS[k], lw + std::log((x[k] - mu_old[k]) * (x[k] - mu[k])));
            }
            for (int k = 0; k < np; ++k) {
                if (k == D) w.line("-------------------------------");
                format_e24(mu[k], a);
                format_e24(std::sqrt(std::exp(logS[k] - logwsum)), b);
                std::fprintf(w.f, "%3d%s +/- %s\n", k + 1, a, b);
            }
            if (np == D) w.line("-------------------------------");
        }
        ++files;
        st.nlike_last = nlike;
    }
    st.written = true;
    st.last = std::chrono::steady_clock::now();
    return files;
}

}  // namespace pc
