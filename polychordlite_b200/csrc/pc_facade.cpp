// pc_facade.cpp -- the C++ facade of libchord.so: Settings and the run_polychord overloads.
//
// Replaces /root/reference/src/polychord/c_interface.cpp:6-213 (the reference builds it into the same library:
// src/polychord/Makefile:3 globs *.cpp).  Declarations: include/interfaces.hpp.  Every overload funnels into the
// C ABI (polychord_c_interface / polychord_c_interface_ini, csrc/pc_engine.cu), exactly as the reference's do into
// the Fortran bind(c) entries, so a C++ caller of the reference library relinks against this one unchanged.
#include "../../include/interfaces.hpp"
#include "../../include/polychord_b200.h"

#include <cmath>

// Defaults of c_interface.cpp:6-39.  (They are the C++ facade's own: e.g. maximise and write_prior are ON here while
// pypolychord passes False / True explicitly; SURVEY.md appendix A.)
Settings::Settings(int _nDims, int _nDerived)
    : nDims(_nDims), nDerived(_nDerived), nlive(500), num_repeats(5 * _nDims), nprior(-1), nfail(-1),
      do_clustering(false), feedback(1), precision_criterion(1e-3), logzero(-1e30), max_ndead(-1),
      boost_posterior(0.0), posteriors(false), equals(false), cluster_posteriors(false), write_resume(false),
      write_paramnames(false), read_resume(false), write_stats(false), write_live(false), write_dead(false),
      write_prior(true), maximise(true), compression_factor(std::exp(-1.0)), synchronous(true), base_dir("chains"),
      file_root("test"), grade_frac(1, 1.0), grade_dims(1, _nDims), loglikes(), nlives(), seed(-1) {}

namespace {
// The C ABI takes mutable char*: hand it private NUL-terminated copies (c_interface.cpp:59-65).
struct CString {
    std::vector<char> buf;
    explicit CString(const std::string& s) : buf(s.begin(), s.end()) { buf.push_back('\0'); }
    char* get() { return buf.data(); }
};
}  // namespace

// c_interface.cpp:45-115.  `comm` is the MPI communicator handle of the reference ABI; this engine has no MPI
// (a run shards over the box's GPUs instead) and passes 0, as the reference does without USE_MPI (:70).
// Exceptions thrown by the callbacks unwind through this frame untouched (RAII only).
void run_polychord(pc_cxx_loglikelihood loglikelihood, pc_cxx_prior prior, pc_cxx_dumper dumper, Settings s) {
    CString base_dir(s.base_dir), file_root(s.file_root);
    int comm = 0;
    polychord_c_interface(loglikelihood, prior, dumper, s.nlive, s.num_repeats, s.nprior, s.nfail, s.do_clustering,
                          s.feedback, s.precision_criterion, s.logzero, s.max_ndead, s.boost_posterior, s.posteriors,
                          s.equals, s.cluster_posteriors, s.write_resume, s.write_paramnames, s.read_resume,
                          s.write_stats, s.write_live, s.write_dead, s.write_prior, s.maximise, s.compression_factor,
                          s.synchronous, s.nDims, s.nDerived, base_dir.get(), file_root.get(),
                          (int)s.grade_frac.size(), s.grade_frac.data(), s.grade_dims.data(), (int)s.loglikes.size(),
                          s.loglikes.data(), s.nlives.data(), s.seed, &comm);
}

// c_interface.cpp:151-165: the callbacks a caller leaves out are the defaults below
void run_polychord(pc_cxx_loglikelihood loglikelihood, pc_cxx_dumper dumper, Settings s) {
    run_polychord(loglikelihood, default_prior, dumper, s);
}
void run_polychord(pc_cxx_loglikelihood loglikelihood, pc_cxx_prior prior, Settings s) {
    run_polychord(loglikelihood, prior, default_dumper, s);
}
void run_polychord(pc_cxx_loglikelihood loglikelihood, Settings s) {
    run_polychord(loglikelihood, default_prior, default_dumper, s);
}

// c_interface.cpp:168-190: the .ini driver path
void run_polychord(pc_cxx_loglikelihood loglikelihood, pc_cxx_setup setup_loglikelihood, std::string inifile) {
    CString path(inifile);
    int comm = 0;
    polychord_c_interface_ini(loglikelihood, setup_loglikelihood, path.get(), &comm);
}

// interfaces.hpp:90 declares default_loglikelihood; the reference never defines it (a caller that used it would not
// link there).  Defined here as the flat likelihood so the declaration is usable.
double default_loglikelihood(double*, int, double* phi, int nDerived) {
    for (int i = 0; i < nDerived; ++i) phi[i] = 0.0;
    return 0.0;
}
void default_prior(double* cube, double* theta, int nDims) {   // c_interface.cpp:210-211
    for (int i = 0; i < nDims; ++i) theta[i] = cube[i];
}
void default_dumper(int, int, int, double*, double*, double*, double, double) {}   // c_interface.cpp:213
