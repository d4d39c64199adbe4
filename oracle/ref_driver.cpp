// ref_driver.cpp -- TEST INFRASTRUCTURE (like everything under oracle/).
//
// A C++ driver in the style of the reference's src/drivers/polychord_CC.cpp:39 that goes through the
// REFERENCE'S OWN C++ facade -- struct Settings and run_polychord(loglikelihood, prior, dumper, Settings),
// compiled from /root/reference/src/polychord/c_interface.cpp where it lies (see oracle/Makefile, target
// _ref/ref_driver) -- into this repository's libchord.so.  It proves the drop-in claim at the reference's
// own C++ call site: the unmodified facade marshals its Settings into polychord_c_interface(...) and the B200
// engine answers.  Prints one line: "logZ logZerr ndead nlive_last ndumps".
#include <cstdio>
#include <cstdlib>
#include "interfaces.hpp"  // the reference's header (include path set by the Makefile)

extern "C" {
double pc_gaussian_loglikelihood(double*, int, double*, int);
void pc_unit_prior(double*, double*, int);
}

static double g_logZ = 0, g_logZerr = 0;
static int g_ndead = 0, g_nlive_last = -1, g_ndumps = 0;

static void dumper(int ndead, int nlive, int npars, double* live, double* dead, double* logweights, double logZ,
                   double logZerr) {
    (void)npars; (void)live; (void)dead; (void)logweights;
    g_logZ = logZ; g_logZerr = logZerr; g_ndead = ndead; g_nlive_last = nlive; ++g_ndumps;
}

int main(int argc, char** argv) {
    const int nDims = argc > 1 ? std::atoi(argv[1]) : 20, nDerived = 2;
    Settings settings(nDims, nDerived);       // the reference's defaults (c_interface.cpp:6-39)
    settings.nlive = argc > 2 ? std::atoi(argv[2]) : 500;
    settings.num_repeats = 2 * nDims;
    settings.do_clustering = false;
    settings.precision_criterion = 1e-3;
    settings.base_dir = "chains";
    settings.file_root = "ref_driver";
    settings.write_resume = settings.read_resume = settings.write_live = settings.write_dead = false;
    settings.write_stats = settings.write_prior = settings.maximise = false;
    settings.equals = settings.posteriors = settings.cluster_posteriors = false;
    settings.feedback = 0;
    settings.seed = argc > 3 ? std::atoi(argv[3]) : 1;
    run_polychord(pc_gaussian_loglikelihood, pc_unit_prior, dumper, settings);
    std::printf("%.10f %.10f %d %d %d\n", g_logZ, g_logZerr, g_ndead, g_nlive_last, g_ndumps);
    return 0;
}
