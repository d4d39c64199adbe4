/* polychord_b200.h -- C ABI of the B200-native nested-sampling engine (libchord.so).
 *
 * Two groups of entry points:
 *
 *  (1) The drop-in boundary: the exact symbols the reference's libchord.so exports and
 *      its C++ facade / CPython shim bind (reference: src/polychord/interfaces.h:2-56,
 *      Fortran side src/polychord/interfaces.F90:285-324 and :496-497).  A caller that
 *      was linked against the reference library relinks against this one unchanged.
 *
 *  (2) Additive pc_* entry points.  The reference ABI only carries host function
 *      pointers, so the engine cannot see that a callback is "the built-in Gaussian".
 *      pc_register_device_likelihood()/pc_register_device_prior() bind a host pointer to
 *      a device-resident analytic form; polychord_c_interface() looks the pointer up by
 *      identity and runs the whole sampling loop on the GPU.  Unregistered callbacks take
 *      the lock-step host-callback path.  The remaining pc_* calls expose results
 *      (the reference entry returns nothing: interfaces.F90:129 drops output_info),
 *      engine options, and the kernel-level probes the parity tests drive.
 *
 * All pointers are plain host pointers owned by the caller unless stated otherwise; no
 * torch / CUDA types appear in any signature (streams are passed as void*).
 */
#ifndef POLYCHORD_B200_H
#define POLYCHORD_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------
 * (1) Drop-in boundary
 * ---------------------------------------------------------------------------------- */

typedef double (*pc_loglikelihood_t)(double* theta, int nDims, double* phi, int nDerived);
typedef void (*pc_prior_t)(double* cube, double* theta, int nDims);
typedef void (*pc_dumper_t)(int ndead, int nlive, int npars, double* live, double* dead, double* logweights,
                            double logZ, double logZerr);

/* Replaces the Fortran bind(c) entry src/polychord/interfaces.F90:285-436, declared for C
 * callers at src/polychord/interfaces.h:2-45 and called from src/polychord/c_interface.cpp:73-112.
 * Argument order, by-value/by-pointer convention and the `int& comm` (pointer) last argument
 * are identical.  bool arguments are C++ `bool` in the reference declaration; this header
 * spells them `unsigned char` for C callers (same size and calling convention on x86-64 SysV).
 * Fatal configuration errors follow the reference convention (abort.F90:19-29): a banner on
 * stderr and exit(1) -- unless PC_ERRORS=return is set, in which case the call returns and
 * pc_last_run_info().status carries the code. */
#ifdef __cplusplus
typedef bool pc_bool;
#else
typedef unsigned char pc_bool;
#endif

void polychord_c_interface(
    pc_loglikelihood_t loglikelihood, pc_prior_t prior, pc_dumper_t dumper,
    int nlive, int num_repeats, int nprior, int nfail, pc_bool do_clustering, int feedback,
    double precision_criterion, double logzero, int max_ndead, double boost_posterior,
    pc_bool posteriors, pc_bool equals, pc_bool cluster_posteriors, pc_bool write_resume,
    pc_bool write_paramnames, pc_bool read_resume, pc_bool write_stats, pc_bool write_live,
    pc_bool write_dead, pc_bool write_prior, pc_bool maximise, double compression_factor,
    pc_bool synchronous, int nDims, int nDerived, char* base_dir, char* file_root, int nGrade,
    double* grade_frac, int* grade_dims, int n_nlives, double* loglikes, int* nlives, int seed,
    int* comm);

/* Replaces src/polychord/interfaces.F90:496-519 (decl. interfaces.h:47-56): settings, parameters and priors are read
 * from the .ini file (src/polychord/ini.f90 format: "key = value" lines, "P : name | latex | speed | prior type |
 * block | params", "D : name | latex"), setup_loglikelihood() is called once, then the run proceeds as through
 * polychord_c_interface.  All fifteen prior families of priors.f90:5-20 are read (separable, sorted, adaptive sorted,
 * nn_adaptive_layer_gaussian; csrc/pc_ini.cpp). */
void polychord_c_interface_ini(pc_loglikelihood_t loglikelihood, void (*setup_loglikelihood)(void),
                               char* inifile, int* comm);

/* ------------------------------------------------------------------------------------
 * (2) Additive engine API
 * ---------------------------------------------------------------------------------- */

/* Analytic forms with a device implementation (likelihoods/examples/*.f90). */
enum pc_like_kind {
    PC_LIKE_GAUSSIAN = 0,      /* gaussian.f90:12-41; params: mu[D], sigma[D] (or {mu,sigma}, or none = 0.5/0.1) */
    PC_LIKE_RASTRIGIN = 1,     /* rastrigin.f90:20-35; no params */
    PC_LIKE_CORR_GAUSSIAN = 2, /* random_gaussian.f90 / utils.F90:1028-1048; params: mu[D], invcov[D*D] col-major, logdet */
    PC_LIKE_HOST = 3           /* arbitrary host callback (lock-step path) */
};
enum pc_prior_kind {
    PC_PRIOR_UNIFORM = 0, /* priors.f90:40-55 uniform_htp; params: lo[D], hi[D] (none = unit cube) */
    PC_PRIOR_HOST = 1
};

/* Bind a host callback pointer to a device-resident analytic form.  Later calls of
 * polychord_c_interface() with the same pointer run fully on the GPU.  params are copied.
 * Returns 0 on success. */
int pc_register_device_likelihood(pc_loglikelihood_t host_fn, int kind, const double* params, int nparams);
int pc_register_device_prior(pc_prior_t host_fn, int kind, const double* params, int nparams);
void pc_clear_registrations(void);

/* Ready-made host callbacks with the reference's callback signature; each is pre-registered
 * with its device form, so `polychord_c_interface(pc_gaussian_loglikelihood, pc_unit_prior, ...)`
 * is the GPU fast path.  Their parameters are set with pc_register_device_* on these pointers. */
double pc_gaussian_loglikelihood(double* theta, int nDims, double* phi, int nDerived);
double pc_rastrigin_loglikelihood(double* theta, int nDims, double* phi, int nDerived);
double pc_corr_gaussian_loglikelihood(double* theta, int nDims, double* phi, int nDerived);
void pc_unit_prior(double* cube, double* theta, int nDims);    /* theta = cube (c_interface.cpp:210) */
void pc_uniform_prior(double* cube, double* theta, int nDims); /* theta = lo + (hi-lo)*cube */

/* Engine options (name/value); unknown names return -1.
 *   "batch_fraction"  K = max(1, round(nlive*value)) lowest points die per generation; 0 (default) = automatic
 *                     (pc_auto_batch_size): about 1/2 for a run alone on the device (its wall time is the number of
 *                     generations), in whole waves of chains where that stays within [0.5, 0.6] nlive; 1/4 for the runs
 *                     of an ensemble
 *   "no_wave_batch"   1: the automatic batch of a run alone on the device is plain nlive/2
 *   "dense"           the dense chain phase (one chain per point group, csrc/pc_dense.cuh): 0 = ensembles only (default),
 *                     1 = always, -1 = never
 *   "batch_K"         absolute K (overrides batch_fraction when > 0)
 *   "device"          CUDA device ordinal
 *   "warps_per_cta"   chain warps per CTA (default: chosen from the shared-memory budget)
 *   "max_ctas"        cap on the CTAs of one run (default 0 = one warp per chain)
 *   "errors_return"   1: configuration errors return instead of exit(1)
 *   "nh_global"       1: keep the direction scratch in global memory even when it fits in shared memory
 *   "sync_dump"       1: the run kernel exits at every update, the host calls the dumper and relaunches
 *                     (also: environment PC_SYNC_DUMP).  Default 0: dumps are handed to the host through a
 *                     mapped control block while the kernel keeps sampling.  Needed under profilers that
 *                     serialise kernel launches (the host cannot acknowledge a dump from inside the launch call).
 *   "no_pairing"      1: do not use helper warps for the direction preparation (a run alone on the device
 *                     normally pairs every chain warp with a helper warp)
 *   "no_phase_d"      1: keep the order of the live points on one CTA (phase S); normally every CTA of a run
 *                     ranks its share of the live points after a regular generation (phase D)
 *   "no_bulk"         1: phase U streams the phantom records through registers; normally records of an even length
 *                     move by bulk copies (cp.async.bulk) through a shared-memory ring (same results bit for bit)
 *   "no_narrow"       1: the runs of an ensemble store and carry a phantom's theta although nothing reads it (normally
 *                     they skip those columns: less DRAM traffic, same results)
 *   "resume_text"     1: <root>.resume is written in the reference's text layout (read_write.F90:219-288) instead of the
 *                     engine's binary one.  Reading accepts both layouts whatever this option says (and the files
 *                     pypolychord writes for cube_samples, polychord.py:650-789).
 *   "resume_interval" seconds between two rewrites of the resume file at updates (default 1; 0 = every update)
 *   "cap_dead0", "cap_ph0"  initial capacity (records) of the dead / phantom pools; 0 = automatic.
 *                     The pools grow on demand either way (the kernel exits, the host reallocates, relaunches).
 */
int pc_set_option(const char* name, double value);
double pc_get_option(const char* name);
/* CUDA stream (cudaStream_t passed as void*) all engine work is enqueued on; NULL = default stream. */
void pc_set_stream(void* cuda_stream);
/* Callable from inside a host callback (loglikelihood / prior / dumper): the run in flight stops at the next round
 * of the host-callback path and polychord_c_interface reports status -3.  Language bindings whose callbacks cannot
 * unwind through C frames (ctypes, cgo, JNI) use it to turn a callback exception into a prompt return. */
void pc_request_abort(void);
/* Fast/slow parameter grades for the following pc_run()/probe calls (polychord_c_interface takes them from its own
 * grade_dims / grade_frac arguments): grade g owns grade_dims[g] consecutive dimensions and contributes
 * grade_repeats[g] slice steps per chain, drawn in the sub-space of the dimensions of grades >= g
 * (chordal_sampling.f90:94-145).  The dims must sum to nDims, the repeats to num_repeats.  nGrade = 0 clears. */
int pc_set_grades(int nGrade, const int* grade_dims, const int* grade_repeats);
/* Dynamic nlive for the following pc_run() calls (polychord_c_interface takes it from its own loglikes / nlives
 * arguments): above the contour loglikes[i] the run keeps nlives[i] live points (settings%loglikes / settings%nlives,
 * run_time_info.f90:766-777); elsewhere settings.nlive.  In the batched schedule a generation that kills K of its n live
 * points births target - (n - K) chains (none when the survivors already exceed the target, at most 2 batch_K).  At most
 * 16 entries; m = 0 clears. */
int pc_set_nlives(const double* loglikes, const int* nlives, int m);
/* Resume file for the following pc_run() calls (polychord_c_interface uses <base_dir>/<file_root>.resume with its
 * write_resume / read_resume flags; replaces read_write.F90:219-476).  `path` must end in ".resume".  write: the run's
 * state is saved between generations at the update cadence (at most once a second) and after the final kill-off;
 * read: an existing file of the same run shape continues that run -- with its own seed, so an interrupted run and
 * an uninterrupted one end bit-identical -- and a file of a different shape is a fatal error (read_write.F90:402-417).
 * The file is this engine's own binary layout, not the reference's text dump.  NULL clears. */
int pc_set_resume(const char* path, int write, int read);
/* The engine keeps the device buffers of finished runs for the next run (cudaMalloc/cudaFree cost
 * milliseconds); this returns the cached blocks to the driver. */
void pc_release_memory(void);

struct pc_settings;
/* Sharded run over the GPUs of one box, one process per GPU (replaces the reference's MPI administrator /
 * worker scheme, src/polychord/mpi_utils.F90 and nested_sampling.F90:262-303,420-500).  Every rank calls
 *   pc_mgpu_create(settings, world, handle)     allocates this rank's exchange block, returns its 64-byte CUDA IPC handle
 *   [the caller all-gathers the handles, e.g. with torch.distributed]
 *   pc_mgpu_attach(rank, world, handles)         maps every peer's block (handles: world x 64 bytes, rank order)
 * and then the same pc_run()/polychord_c_interface() call with identical settings and seed.  The ranks keep the
 * run state replicated, deal the chains of a generation k % world and exchange the new live points and the
 * covariance statistics through the mapped blocks over NVLink inside the persistent kernel.  nlike in
 * pc_run_info counts this rank's evaluations (sum over ranks for the run's total); everything else is identical
 * on every rank.  The dumper is honoured on every rank that passes one. */
int pc_mgpu_create(const struct pc_settings* s, int world, unsigned char* handle64);
int pc_mgpu_attach(int rank, int world, const unsigned char* handles);
int pc_mgpu_destroy(void);

typedef struct pc_run_info {
    int status;               /* 0 ok; <0 error code */
    double logZ, logZerr;     /* run_time_info.f90:652-678 estimate */
    double logZ_raw, logZ2_raw;
    long long ndead, nlike, nchains, ngenerations, nupdates, nfailures, nslices;
    long long nphantoms_final;
    int batch_K, warps_per_cta, ctas_per_run, kernel_launches;
    double device_ms;         /* sum of CUDA-event durations of the engine's kernels */
    double wall_ms;           /* entry to return of the call */
    long long h2d_bytes, d2h_bytes;
    long long algorithmic_bytes; /* 8T+8D per slice, 8T per chain, 8D^2 per generation */
    /* where the run kernel's time went, from SM clock counters (ms at the device's nominal SM clock):
     * [0] CTA 0 waiting for the chains, [1] phase S (select/evidence), [2] covariance+Cholesky finish,
     * [3] phase U (phantom compaction + covariance), [4..6] one representative chain warp: direction
     * preparation left on the critical path, whitening, slice steps; [7] whole kernel */
    double phase_ms[8];
    long long ncluster_max;     /* do_clustering: largest number of clusters an update found, and the number of */
    long long ncluster_updates; /* updates that ran the clustering pass */
    double cluster_ms;          /* wall time of the clustering passes (their kernels are not part of device_ms) */
    /* the run kernel that was dispatched: pc_run_kernel<G, DPL, KIND, MODE> -- G lanes per trial point, DPL dimensions
     * per lane, likelihood kind, 0 = a chain per warp / 1 = the dense chain phase (a chain per point group) */
    int kernel_G, kernel_DPL, kernel_kind, kernel_mode;
    int nlive_final;            /* live points when the sampling loop ended (dynamic nlive) */
    int pad_;
} pc_run_info;

/* Results of the most recent polychord_c_interface()/pc_run() in this process. */
int pc_last_run_info(pc_run_info* out);

typedef struct pc_settings {
    int nDims, nDerived, nlive, num_repeats, nprior, nfail;
    int do_clustering, feedback;
    double precision_criterion, logzero;
    int max_ndead;
    double boost_posterior;
    int posteriors, equals, cluster_posteriors;
    double compression_factor;
    int seed;
} pc_settings;

/* One run with explicit device forms (what polychord_c_interface dispatches to).
 * dumper may be NULL.  prior_params: lo[D],hi[D] or NULL. */
int pc_run(const pc_settings* s, int like_kind, const double* like_params, int n_like_params,
           const double* prior_params, int n_prior_params, pc_dumper_t dumper, pc_run_info* out);

/* nruns independent runs (seeds[i]) advanced concurrently by ONE persistent kernel, each run
 * owning its own group of CTAs.  out[nruns].  dead/logweight arrays are not returned. */
int pc_run_ensemble(const pc_settings* s, int like_kind, const double* like_params, int n_like_params,
                    const double* prior_params, int n_prior_params, int nruns, const int* seeds,
                    pc_run_info* out);

/* ---- kernel-level probes (parity tests call these through the same library) ---- */

/* SliceSampling for nchains chains (chordal_sampling.f90:7-92): chain c starts from
 * seed_points[c*T..] with contour logL[c], RNG stream uid[c]; cholesky is D*D column-major.
 * babies_out: nchains*R*T doubles (b0 = logL); nlike_out[nchains]. */
int pc_slice_chains(const pc_settings* s, int like_kind, const double* like_params, int n_like_params,
                    const double* prior_params, int n_prior_params, int nchains, const double* seed_points,
                    const double* cholesky, const double* logL, const unsigned long long* uid,
                    double* babies_out, long long* nlike_out);

/* calculate_point (calculate.f90:6-50) for npts records of T doubles (cube filled in). Returns nlike. */
int pc_calculate_points(const pc_settings* s, int like_kind, const double* like_params, int n_like_params,
                        const double* prior_params, int n_prior_params, double* records, int npts);

/* RNG stream probe: Philox4x32-10 block and the uniform derived from it, computed on the device. */
int pc_device_philox(const unsigned* ctr, const unsigned* key, unsigned* out4);
int pc_device_uniforms(unsigned seed, unsigned tag, unsigned long long uid, unsigned a0, unsigned b, int n,
                       double* out);
int pc_device_inv_normal_cdf(const double* p, int n, double* out);

/* Directions of one chain in use order (chordal_sampling.f90:94-145): out[i*nDims + r], i < num_repeats. */
int pc_device_directions(int nDims, int num_repeats, unsigned seed, unsigned long long uid, double* out);
/* Evidence recurrences (run_time_info.f90:211-296) for `count` deaths with live counts n_start, n_start-1, ...;
 * state = {logZ, logZ2, logX, logZX, logXX} (single cluster), updated in place; logw_out[count]. */
int pc_device_evidence(double* state, const double* logLs, int count, int n_start, double* logw_out);
/* calc_cholesky (utils.F90:621-649), column-major D x D; returns 1 when the identity fallback was taken. */
int pc_device_cholesky(const double* a, int D, double* L_out);

/* ---- output files in the reference's formats (replaces src/polychord/read_write.F90:479-716, 809-961) ----
 * polychord_c_interface() writes them itself when its write_* / posteriors / equals flags are set:
 *   <root>.stats, <root>_dead.txt, <root>_dead-birth.txt, <root>_phys_live.txt, <root>_phys_live-birth.txt,
 *   <root>.txt (weighted posterior), <root>_equal_weights.txt, <root>.prior_info, <root>_prior.txt (the prior draws)
 * with every number in Fortran's E24.15E3 edit descriptor (utils.F90:19).  The two entry points below are host-only
 * (no device needed): the formatter, and the writer driven with explicit arrays.
 * flags: 1 stats, 2 live, 4 dead, 8 prior (_prior.txt; prior_info is written by the run), 16 weighted posterior, 32 equally weighted posterior.
 * dead_rows / live_rows: rows [theta(nDims), phi(nDerived), birth contour, logL]; dead_logw[i] = log weight + logL. */
void pc_format_e24(double value, char* out25);

/* Clusters of the last run with do_clustering.  The clusters persist as in the
 * reference -- at every update each one is searched for sub-clusters and split (clustering.f90:253-324,
 * run_time_info.f90:303-505), one without live points is deleted (:507-598) -- and every death is attributed to the
 * cluster of the dying point.  pc_last_clusters returns their number: the clusters alive at the end of sampling first
 * (label order), then the deleted ones; rows = {log<Z_p>, log<Z_p^2>} per cluster, uid = its identity (0: the initial
 * cluster, a split creates new identities).  pc_last_dead_clusters: the identity every dead point's cluster had at its
 * death.  pc_last_cluster_tree: for every identity the one it was split from (-1: none). */
int pc_last_clusters(int* nactive, double* rows, int* uid);
long long pc_last_dead_clusters(int* out, long long cap);
int pc_last_cluster_tree(int* parent_out);

/* Host-only: parse a resume file in the reference's TEXT layout (src/polychord/read_write.F90:219-288 writes it,
 * :384-476 reads it; pypolychord/polychord.py:650-789 writes it for cube_samples).  ints[8] = {nDims, nDerived, ndead,
 * ncluster, ncluster_dead, live points, phantoms, likelihood calls}; reals[6] = {logZ, logZ2, log sum_p X_p, last
 * update volume, lowest and highest live logL}.  A non-empty out_path re-writes what was read in the same layout (one
 * active cluster only).  Returns 0; -2 for a malformed file, -3 for another failure (the message goes to stderr;
 * a probe never exits the process). */
int pc_resume_text_probe(const char* path, long long* ints, double* reals, const char* out_path);
int pc_write_files(const char* base_dir, const char* file_root, int flags, int nDims, int nDerived, long long ndead,
                   const double* dead_rows, const double* dead_logw, int nlive, const double* live_rows, double logZ,
                   double logZerr, long long nlike, int num_repeats, double compression_factor, unsigned seed);

/* boost_posterior (clean_phantoms, run_time_info.f90:820-877): when posteriors or equals is set and boost_posterior
 * is not 0, every phantom the update removes becomes a posterior sample with probability boost_posterior /
 * num_repeats (1 when boost_posterior < 0; generate.F90:311-316), carrying the weight of the death since the last
 * update with the smallest logL above its own.  They are written to <root>.txt and <root>_equal_weights.txt with the
 * dead points; this accessor returns those of the last run: rows[i] = [theta, phi, birth, logL] (npars = nDims +
 * nDerived + 2), dead_index[i] = the dead point whose weight sample i carries, logw[i] = that weight + logL of the
 * sample.  Returns the number of samples (fills up to cap), -1 when npars does not match the last run. */
long long pc_last_boosted(double* rows, long long* dead_index, double* logw, long long cap, int npars);

/* cube_samples (pypolychord.run, polychord.py:452, 576-579): the next run through polychord_c_interface starts from
 * these live points instead of drawing them from the prior: npoints x nDims cube coordinates, row-major, copied.  They
 * are evaluated through the run's own callbacks on the calling thread and enter with birth contour logzero.  The
 * reference passes them through a resume file (_make_resume_file, polychord.py:650-789) and lets its dynamic-nlive
 * mechanism absorb a count different from nlive; this engine wants exactly nlive of them.  One-shot; NULL or 0 clears.
 * Returns 0, -1 on bad arguments; a mismatch with the run's shape is reported by the run. */
int pc_set_initial_live(const double* cube_samples, int npoints, int nDims);

/* maximise (maximiser.F90, nelder_mead.f90): with the `maximise` flag polychord_c_interface ends by maximising the
 * likelihood and the posterior with Nelder-Mead from the best nDims + 1 live points and writes <root>.maximum
 * (write_max_file, read_write.F90:754-807).  The two entry points below are host-only (no device needed).
 * pc_maximise: live_records = nlive records of 2*nDims + nDerived + 2 doubles [cube | theta | phi | birth | logL];
 * point_out = one such record, the maximum found.  Returns 0, 1 when no simplex can be built, -1 on bad arguments.
 * pc_prior_log_density: dXdtheta (maximiser.F90:172-202), the log prior density at a cube point from the
 * finite-difference Jacobian of the prior transform. */
int pc_maximise(pc_loglikelihood_t loglikelihood, pc_prior_t prior, int nDims, int nDerived, double logzero,
                const double* live_records, int nlive, int posterior, double* point_out);
double pc_prior_log_density(pc_prior_t prior, const double* cube, int nDims);

/* Host-only (no device needed): hypercube_to_physical (priors.f90:494-556) of the parameter block of an .ini file in the
 * reference's format (ini.f90:354-458), applied to one cube point: the separable families, their sorted forms, the
 * adaptive sorted families and nn_adaptive_layer_gaussian (priors.f90:40-488).  This is the transform
 * polychord_c_interface_ini runs with.  Returns 0; -6 when the file cannot be read or parsed (message on stderr), -7
 * when nDims is not the file's parameter count. */
int pc_ini_prior_transform(const char* inifile, const double* cube, double* theta, int nDims);

/* pc_write_files with the phantoms boost_posterior promoted (see pc_last_boosted): boosted_rows like dead_rows,
 * boosted_logw[i] = log weight + logL, boosted_after[i] (ascending) = the number of dead rows that precede sample i in
 * the posterior files -- the deaths up to the update that removed it (update_posteriors, run_time_info.f90:1036-1061). */
int pc_write_files_boosted(const char* base_dir, const char* file_root, int flags, int nDims, int nDerived, long long ndead,
                           const double* dead_rows, const double* dead_logw, int nlive, const double* live_rows, double logZ,
                           double logZerr, long long nlike, int num_repeats, double compression_factor, unsigned seed,
                           long long nboosted, const double* boosted_rows, const double* boosted_logw,
                           const long long* boosted_after);

/* NN_clustering (clustering.f90:15-97) of m points (row-major m x nDims cube coordinates) exactly as the engine's
 * update runs it (device k-nearest-neighbour lists, host union-find); labels in order of first appearance.
 * Returns the number of clusters. */
int pc_cluster_points(const double* points, int m, int nDims, int* labels_out);

/* Measured FP64 fused-multiply-add throughput of the current device in TFLOP/s (a microbenchmark: eight independent
 * FMA chains per thread, best of five launches).  bench.py quotes the run's arithmetic against it. */
double pc_measure_fp64_tflops(void);

/* The batch size (deaths per generation) the engine picks by itself for a run of nlive live points alone on the device,
 * or sharded over `world` devices (about nlive/2, a whole number of waves of chains where that
 * stays below 0.6 nlive); the options batch_K / batch_fraction override it.  Needs a CUDA device (the wave is a
 * function of its SM count). */
int pc_auto_batch_size(int nlive, int world);

/* Number of CUDA devices visible; <=0 means the engine cannot run (no CPU fallback exists). */
int pc_device_count(void);
const char* pc_version(void);

#ifdef __cplusplus
}
#endif
#endif /* POLYCHORD_B200_H */
