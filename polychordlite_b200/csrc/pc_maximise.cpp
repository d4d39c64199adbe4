// pc_maximise.cpp -- the `maximise` setting: after sampling, the likelihood and the
// posterior are maximised by Nelder-Mead started from the best nDims + 1 live points, and <root>.maximum is written.
//
// Replaces src/polychord/maximiser.F90 (maximise :31-77, do_maximisation :80-153, maximisation_func :155-170,
// dXdtheta :172-202), src/polychord/nelder_mead.f90 (nelder_mead :4-75, det :161-211) and write_max_file
// (read_write.F90:754-807).  Host-only: the likelihood and the prior are the host callbacks the caller handed to
// polychord_c_interface (the library's ready-made callbacks are host functions too), called on the calling thread.
#include "pc_maximise.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <numeric>
#include <stdexcept>
#include "pc_errors.h"

#include "pc_files.h"

namespace pc {

// det (nelder_mead.f90:161-211): Gaussian elimination without pivoting unless a zero pivot turns up; a is n x n
// column-major and is destroyed.
double maximise_det(std::vector<double>& a, int n) {
    auto A = [&](int i, int j) -> double& { return a[(size_t)i + (size_t)j * n]; };
    int sign = 1;
    for (int k = 0; k < n - 1; ++k) {
        if (A(k, k) == 0.0) {
            bool found = false;
            for (int i = k + 1; i < n; ++i)
                if (A(i, k) != 0.0) {
                    for (int j = 0; j < n; ++j) std::swap(A(i, j), A(k, j));
                    found = true;
                    sign = -sign;
                    break;
                }
            if (!found) return 0.0;
        }
        for (int j = k + 1; j < n; ++j) {
            const double m = A(j, k) / A(k, k);
            for (int i = k + 1; i < n; ++i) A(j, i) -= m * A(k, i);
        }
    }
    double d = sign;
    for (int i = 0; i < n; ++i) d *= A(i, i);
    return d;
}

// nelder_mead (nelder_mead.f90:4-75), a MAXIMISER: x is n x (n+1) column-major (one vertex per column), f the values
// at the vertices; both are updated in place.  alpha = 1, gamma = 2, rho = sigma = 0.5.  Stops when the spread of
// the values is below dl or the simplex has shrunk to dl of its first volume per dimension.  Returns the best vertex.
// The reference orders the vertices with an (unstable) quicksort; equal values are ordered by index here.
std::vector<double> nelder_mead(const std::function<double(const double*)>& func, std::vector<double>& x,
                                std::vector<double>& f, double dl, long long* nfunc, long long max_iter) {
    const int n = (int)f.size() - 1;
    std::vector<int> idx(n + 1);
    std::vector<double> xo(n), xr(n), xe(n), xc(n), m((size_t)n * n);
    auto col = [&](int v) { return &x[(size_t)v * n]; };
    // (det1/det0)**(1./n): the exponent is a single-precision quotient in the reference
    const double expo = (double)(1.0f / (float)n);
    double det0 = -1.0;
    long long calls = 0;
    for (long long iter = 0; max_iter <= 0 || iter < max_iter; ++iter) {
        std::iota(idx.begin(), idx.end(), 0);
        std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return f[a] < f[b]; });   // ascending: idx[0] worst
        const int best = idx[n], worst = idx[0];
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i) m[(size_t)i + (size_t)j * n] = col(idx[j])[i] - col(best)[i];
        const double det1 = std::fabs(maximise_det(m, n));
        if (det0 < 0.0) det0 = det1;
        if (f[best] - f[worst] < dl || std::pow(det1 / det0, expo) < dl) break;
        for (int i = 0; i < n; ++i) {   // centroid of all but the worst
            double s = 0.0;
            for (int j = 1; j <= n; ++j) s += col(idx[j])[i];
            xo[i] = s / n;
        }
        for (int i = 0; i < n; ++i) xr[i] = xo[i] + 1.0 * (xo[i] - col(worst)[i]);
        const double fr = func(xr.data()); ++calls;
        if (fr <= f[best] && f[idx[1]] < fr) {            // reflection
            f[worst] = fr;
            std::copy(xr.begin(), xr.end(), col(worst));
        } else if (fr > f[best]) {                         // expansion
            for (int i = 0; i < n; ++i) xe[i] = xo[i] + 2.0 * (xr[i] - xo[i]);
            const double fe = func(xe.data()); ++calls;
            if (fe > fr) { f[worst] = fe; std::copy(xe.begin(), xe.end(), col(worst)); }
            else { f[worst] = fr; std::copy(xr.begin(), xr.end(), col(worst)); }
        } else {                                           // contraction, else shrink towards the best vertex
            for (int i = 0; i < n; ++i) xc[i] = xo[i] + 0.5 * (col(worst)[i] - xo[i]);
            const double fc = func(xc.data()); ++calls;
            if (fc > f[worst]) {
                f[worst] = fc;
                std::copy(xc.begin(), xc.end(), col(worst));
            } else {
                for (int j = 0; j < n; ++j) {
                    double* v = col(idx[j]);
                    for (int i = 0; i < n; ++i) v[i] = col(best)[i] + 0.5 * (v[i] - col(best)[i]);
                    f[idx[j]] = func(v); ++calls;
                }
            }
        }
    }
    if (nfunc) *nfunc = calls;
    std::iota(idx.begin(), idx.end(), 0);
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return f[a] < f[b]; });
    return std::vector<double>(col(idx[n]), col(idx[n]) + n);
}

// dXdtheta (maximiser.F90:172-202): log of the prior density at a cube point, from the finite-difference Jacobian of
// the prior transform: D log(dx) - log(s det(dtheta)), stepping backwards where the cube's edge is in the way.
double maximise_dXdtheta(pc_prior_t prior, const double* cube, int D, double dx) {
    std::vector<double> c0(D), th0(D), th1(D), dth((size_t)D * D), c(cube, cube + D);
    prior(c.data(), th0.data(), D);
    int s = 1;
    for (int i = 0; i < D; ++i) {
        c0 = c;
        if (c0[i] + dx >= 1.0) { c0[i] -= dx; s = -s; } else c0[i] += dx;
        prior(c0.data(), th1.data(), D);
        for (int r = 0; r < D; ++r) dth[(size_t)r + (size_t)i * D] = th1[r] - th0[r];
    }
    return D * std::log(dx) - std::log(s * maximise_det(dth, D));
}

// calculate_point (calculate.f90:6-50) on the host: rec = [cube | theta | phi | birth | logL]
static void host_calculate_point(pc_loglikelihood_t ll, pc_prior_t prior, double* rec, int D, int P, double logzero) {
    bool in = true;
    for (int i = 0; i < D; ++i) in = in && rec[i] >= 0.0 && rec[i] <= 1.0;
    double* theta = rec + D;
    double* phi = rec + 2 * D;
    if (in) {
        prior(rec, theta, D);
        std::vector<double> dummy(1, 0.0);
        rec[2 * D + P + 1] = ll(theta, D, P > 0 ? phi : dummy.data(), P);
    } else {
        for (int i = 0; i < D + P; ++i) theta[i] = 0.0;
        rec[2 * D + P + 1] = logzero;
    }
}

// do_maximisation (maximiser.F90:80-153), one cluster: the nDims + 1 live points with the largest logL (plus the log
// prior density when the posterior is asked for) are the first simplex; the best vertex Nelder-Mead ends with is
// evaluated into a full record.
bool do_maximisation(pc_loglikelihood_t ll, pc_prior_t prior, int D, int P, double logzero, const double* live, int nlive,
                     bool posterior, double* max_point) {
    const int T = 2 * D + P + 2;
    std::fill(max_point, max_point + T, 0.0);
    if (nlive < D + 1) return false;
    std::vector<double> l(nlive);
    for (int j = 0; j < nlive; ++j) {
        l[j] = live[(size_t)j * T + T - 1];
        if (posterior) l[j] += maximise_dXdtheta(prior, live + (size_t)j * T, D);
    }
    std::vector<int> order(nlive);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return l[a] < l[b]; });
    if (!(l[order[nlive - 1]] > logzero)) return false;
    std::vector<double> simplex((size_t)D * (D + 1)), f(D + 1);
    for (int v = 0; v <= D; ++v) {
        const int src = order[nlive - 1 - D + v];
        std::copy(live + (size_t)src * T, live + (size_t)src * T + D, simplex.begin() + (size_t)v * D);
        f[v] = l[src];
    }
    std::vector<double> rec(T);
    auto func = [&](const double* x) {   // maximisation_func, maximiser.F90:155-170
        std::copy(x, x + D, rec.begin());
        host_calculate_point(ll, prior, rec.data(), D, P, logzero);
        double v = rec[T - 1];
        if (posterior && v > logzero) v += maximise_dXdtheta(prior, x, D);
        return v;
    };
    const std::vector<double> x = nelder_mead(func, simplex, f, 1e-5, nullptr, 200000);
    std::copy(x.begin(), x.end(), max_point);
    host_calculate_point(ll, prior, max_point, D, P, logzero);
    return true;
}

// write_max_file (read_write.F90:754-807)
void write_max_file(const std::string& path, int D, int P, const double* max_point, const double* max_post_point,
                    double dXdtheta, const double* mean_point) {
    FILE* f = std::fopen(path.c_str(), "w");
    if (!f) throw pc::RunError("polychord_b200: cannot write " + path);
    const int T = 2 * D + P + 2;
    char b[32];
    auto num = [&](double v) { format_e24(v, b); b[24] = 0; std::fputs(b, f); };
    auto row = [&](const double* p) { for (int k = 0; k < D + P; ++k) num(p[D + k]); std::fputc('\n', f); };
    std::fputs("Maximum LogLikelihood:\n", f); num(max_point[T - 1]); std::fputc('\n', f);
    std::fputs("Maximum Likelihood point:\n", f); row(max_point); std::fputc('\n', f);
    std::fputs("Maximum Posterior:\n", f); num(max_post_point[T - 1] + dXdtheta); std::fputc('\n', f);
    std::fputs("Maximum Likelihood at posterior:\n", f); num(max_post_point[T - 1]); std::fputc('\n', f);
    std::fputs("Maximum Posterior point:\n", f); row(max_post_point); std::fputc('\n', f);
    if (mean_point) {
        std::fputs("LogLikelihood(mean):\n", f); num(mean_point[T - 1]); std::fputc('\n', f);
        std::fputs("mean point:\n", f); row(mean_point);
    }
    std::fclose(f);
}

}  // namespace pc
