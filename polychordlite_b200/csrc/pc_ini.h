// pc_ini.h -- the .ini driver path (see pc_ini.cpp).
#pragma once
#include <string>
#include <utility>
#include <vector>

namespace pc {

struct IniParam {
    std::string name, latex;
    int speed = 1, prior_type = 1, block = 1;   // prior_type: priors.f90:5-20 numbering (1 uniform ... 6 exponential)
    std::vector<double> params;
};

struct IniConfig {
    int nlive = 0, num_repeats = 0, nprior = -1, nfail = -1, feedback = 1, max_ndead = -1, seed = -1;
    bool do_clustering = false, posteriors = false, equals = false, cluster_posteriors = false, write_resume = false,
         write_paramnames = false, read_resume = false, write_stats = true, write_live = false, write_dead = true,
         write_prior = false, maximise = false;
    double precision_criterion = 1e-3, logzero = -1e30, boost_posterior = 0.0, compression_factor = 0.36787944117144233;
    std::string base_dir = "chains", file_root = "test";
    std::vector<double> grade_frac;
    std::vector<int> grade_dims;
    std::vector<IniParam> params;
    std::vector<std::pair<std::string, std::string>> derived;   // name, latex
};

// read_params (ini.f90:44-95); throws std::invalid_argument with the reference's messages where it has them
IniConfig parse_ini(const std::string& path);
// hypercube_to_physical (priors.f90:494-556) for the separable prior types
void ini_prior_transform(const IniConfig& c, const double* cube, double* theta);

}  // namespace pc
