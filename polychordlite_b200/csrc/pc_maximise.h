// pc_maximise.h -- the `maximise` setting (see pc_maximise.cpp).  Host-only.
#pragma once
#include <functional>
#include <string>
#include <vector>

#include "polychord_b200.h"

namespace pc {

double maximise_det(std::vector<double>& a, int n);
std::vector<double> nelder_mead(const std::function<double(const double*)>& func, std::vector<double>& x,
                                std::vector<double>& f, double dl, long long* nfunc = nullptr, long long max_iter = 0);
double maximise_dXdtheta(pc_prior_t prior, const double* cube, int D, double dx = 1e-5);
// live: nlive records of T = 2D + P + 2 doubles [cube | theta | phi | birth | logL]; max_point: one such record out.
// Returns false when no simplex can be built (fewer than D + 1 live points, or none above logzero).
bool do_maximisation(pc_loglikelihood_t ll, pc_prior_t prior, int D, int P, double logzero, const double* live, int nlive,
                     bool posterior, double* max_point);
void write_max_file(const std::string& path, int D, int P, const double* max_point, const double* max_post_point,
                    double dXdtheta, const double* mean_point);

}  // namespace pc
