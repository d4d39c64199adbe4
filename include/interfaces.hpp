// interfaces.hpp -- the C++ facade of libchord.so (B200 engine).
//
// Replaces /root/reference/src/polychord/interfaces.hpp:8-92 (declarations) and c_interface.cpp:6-213
// (definitions, here polychordlite_b200/csrc/pc_facade.cpp).  Same struct layout, same overload set and same
// (mangled) symbols, so the reference's C++ callers -- src/drivers/polychord_CC.cpp:39, polychord_CC_ini.cpp:10,
// pypolychord/_pypolychord.cpp:219 -- compile unchanged against this header and link with -lchord alone
// (tests/test_facade_link.py does exactly that).  There is no MPI in this engine: a run is sharded over the GPUs of
// a box instead, so the USE_MPI overloads of the reference are not declared.
#pragma once
#include <string>
#include <vector>

// callback shapes of the C ABI (interfaces.h:2-13)
using pc_cxx_loglikelihood = double (*)(double* theta, int nDims, double* phi, int nDerived);
using pc_cxx_prior = void (*)(double* cube, double* theta, int nDims);
using pc_cxx_dumper = void (*)(int ndead, int nlive, int npars, double* live, double* dead, double* logweights,
                               double logZ, double logZerr);
using pc_cxx_setup = void (*)();

// One flat record of run settings; passed BY VALUE through run_polychord, so the member order and types are ABI
// (interfaces.hpp:8-44).  Defaults: Settings::Settings (c_interface.cpp:6-39) -- note they are the C++ defaults
// (clustering off, no files except the prior samples, maximise on), not pypolychord's.
struct Settings {
    int nDims, nDerived;               // sampled and derived parameters
    int nlive, num_repeats;            // live points; slice steps per chain (default 5 * nDims)
    int nprior, nfail;                 // prior draws for the initial live points (-1: nlive); failed spawns tolerated (-1: nlive)
    bool do_clustering;
    int feedback;
    double precision_criterion, logzero;
    int max_ndead;                     // -1: no limit
    double boost_posterior;
    bool posteriors, equals, cluster_posteriors;
    bool write_resume, write_paramnames, read_resume, write_stats, write_live, write_dead, write_prior, maximise;
    double compression_factor;
    bool synchronous;
    std::string base_dir, file_root;
    std::vector<double> grade_frac;
    std::vector<int> grade_dims;
    std::vector<double> loglikes;      // dynamic nlive: above contour loglikes[i] keep nlives[i] live points
    std::vector<int> nlives;
    int seed;                          // < 0: from the clock

    Settings(int _nDims = 0, int _nDerived = 0);
};

// every overload ends in polychord_c_interface (the first four) or polychord_c_interface_ini (the last)
void run_polychord(pc_cxx_loglikelihood loglikelihood, pc_cxx_prior prior, pc_cxx_dumper dumper, Settings settings);
void run_polychord(pc_cxx_loglikelihood loglikelihood, pc_cxx_dumper dumper, Settings settings);   // unit-cube prior
void run_polychord(pc_cxx_loglikelihood loglikelihood, pc_cxx_prior prior, Settings settings);     // no dumper
void run_polychord(pc_cxx_loglikelihood loglikelihood, Settings settings);
void run_polychord(pc_cxx_loglikelihood loglikelihood, pc_cxx_setup setup_loglikelihood, std::string inifile);

double default_loglikelihood(double* theta, int nDims, double* phi, int nDerived);   // flat: log L = 0
void default_prior(double* cube, double* theta, int nDims);                          // theta = cube
void default_dumper(int, int, int, double*, double*, double*, double, double);       // does nothing
