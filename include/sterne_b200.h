/*
 * sterne_b200 -- C ABI of the B200-native astrometric log-likelihood hot path.
 *
 * Drop-in boundary for dingswin/sterne's likelihood path.  The reference has no
 * native code and no FFI of its own (pure Python); the entry points below are what
 * a ctypes binding inside the reference's `Gaussianlikelihood` (simulate.py:325-384)
 * and `positions` (model/positions.py:13-30) would bind.  Plain pointers and sizes
 * only; no torch / CUDA types in the signatures (streams travel as void*).
 *
 * All functions return 0 on success or a negative StnStatus; they never throw.
 * stn_last_error() gives a thread-local description of the last failure.
 *
 * Ownership: the caller owns every buffer it passes in.  The library owns only the
 * context (device copies of the epoch tables made by stn_create, plus staging
 * buffers used by the *_host entry points).  A context's problem is immutable after
 * creation (only the optional prior can be replaced, stn_set_prior); use a context from
 * one thread at a time.
 */
#ifndef STERNE_B200_H
#define STERNE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STN_ABI_VERSION 2

/* Parameter roots in the reference's fixed alphabetical order
 * (priors.py:270, positions.py:56).  StnSource.col[] is indexed by these. */
enum StnRoot {
    STN_DEC = 0, STN_EFAC = 1, STN_EFAD = 2, STN_INCL = 3, STN_MU_A = 4,
    STN_MU_D = 5, STN_OM_ASC = 6, STN_PX = 7, STN_RA = 8, STN_NROOTS = 9
};

/* Most output buffers one stn_loglike_scatter launch can write (this GPU + 7 peers). */
#define STN_MAX_OUTS 8

/* The reference's "parameter switched off" sentinel (positions.py:59). */
#define STN_SENTINEL (-999.0)

enum StnStatus {
    STN_OK = 0,
    STN_ERR_INVALID = -1,   /* bad argument / inconsistent problem description */
    STN_ERR_CUDA = -2,      /* a CUDA runtime call failed (message has the CUDA error) */
    STN_ERR_NOMEM = -3,
    STN_ERR_DEVICE = -4     /* no usable sm_100 device */
};

enum StnLayout {
    STN_LAYOUT_AOS = 0,     /* theta[n*ld + p] : (N, P) row-major, ld >= P  (sampler-native) */
    STN_LAYOUT_SOA = 1      /* theta[p*ld + n] : (P, N), ld >= N            (coalesced loads) */
};

enum StnMode {
    STN_MODE_FAST = 0,      /* hoisted per-(sample,source) coefficients, FMA, chi-square as a running fraction */
    STN_MODE_STRICT = 1     /* the reference's operation order, one rounding per operation
                               (positions.py:108-120,308-309; simulate.py:317-320,368-371) */
};

/* Per-epoch row of the FAST table: STN_FAST_ROW doubles, 16-byte aligned pairs.
 *  [0] dT = (epoch - refepoch) * d->yr   (positions.py:108)
 *  [1] X  [2] Y  [3] Z   Earth wrt SSB, AU, ICRS (positions.py:305)
 *  [4] observed RA (rad)   [5] observed Dec (rad)      (simulate.py:217)
 *  EFAC off: [6] 1/err_ra  [7] 1/err_dec  [8],[9] unused (0)
 *  EFAC on : [6] err_random_ra^2 [7] err_random_dec^2 [8] err_sys_ra^2 [9] err_sys_dec^2
 *                                                   (simulate.py:317)
 *  [10] b_AU*cos(theta)  [11] b_AU*sin(theta)   orbit terms, 0 when no parfile
 *                                                   (reflex_motion.py:187-200)
 */
#define STN_FAST_ROW 12

/* Per-epoch row of the STRICT table: STN_STRICT_ROW doubles.
 *  [0] dT [1] X [2] Y [3] Z [4] obs RA [5] obs Dec
 *  [6] err_ra [7] err_dec                    (VLBI 'errs'; also used when efac is sampled as -999)
 *  [8] err_random_ra [9] err_random_dec [10] err_sys_ra [11] err_sys_dec   (EFAC on)
 *  [12] b_AU [13] cos(theta) [14] sin(theta) [15] unused
 */
#define STN_STRICT_ROW 16

typedef struct StnSource {
    int64_t n_epochs;
    const double* fast_rows;     /* host, n_epochs * STN_FAST_ROW   */
    const double* strict_rows;   /* host, n_epochs * STN_STRICT_ROW */
    int32_t col[STN_NROOTS];     /* theta column of each root for this source, -1 = absent
                                    (positions.py:32-60 routing, resolved once) */
    int32_t efac_on;             /* 1: rows carry err_random/err_sys and col[STN_EFAC] >= 0 */
    int32_t binary;              /* 1: a parfile was given (dict_timing != {}), orbit terms valid */
    double cos_decj;             /* cos(DECJ of the parfile)  (reflex_motion.py:212) */
    double neg_sum_log_err;      /* -sum(log(errs)) over 2E terms, used when EFAC is off
                                    (simulate.py:371; parameter-independent) */
    double a1_lts;               /* A1 in lt-s for the a1dot constraint (kopeikin_effects.py:61) */
} StnSource;

typedef struct StnProblem {
    int32_t n_sources;
    int32_t n_params;            /* P: number of theta columns */
    const StnSource* sources;
    /* a1dot Gaussian constraints (simulate.py:373-377), already broadcast on the host:
       term k uses source a1dot_source[k], mean a1dot_mu[k], sigma a1dot_sigma[k]. */
    int32_t n_a1dot;
    const int32_t* a1dot_source;
    const double* a1dot_mu;
    const double* a1dot_sigma;
    /* sin(incl) Gaussian constraints (simulate.py:379-383): theta column, mean, sigma. */
    int32_t n_sini;
    const int32_t* sini_col;
    const double* sini_mu;
    const double* sini_sigma;
    /* Unit factors the reference takes from astropy.units at run time; 0 selects the library's
     * defaults.  A host that has astropy passes astropy's own values so that the last bit agrees:
     *   mas2rad     (u.mas).to(u.rad)        positions.py:119-120
     *   deg2rad     (u.deg).to(u.rad)        reflex_motion.py:170 (sin / cos of Om_asc in deg)
     *   masyr2rads  (u.mas/u.yr).to(u.rad/u.s)  kopeikin_effects.py:45 */
    double mas2rad;
    double deg2rad;
    double masyr2rads;
} StnProblem;

/* Independent per-parameter priors of the reference's .inits kinds (priors.py:356-366), so a
 * sampler's log-posterior is ONE launch.  Uniform(a,b); Gaussian(mu=a, sigma=b); Sine(a,b)
 * [density ~ sin x on (a,b)]; plus the limits on sin(incl) (bilby Constraint, priors.py:351-354). */
enum StnPriorKind { STN_PRIOR_UNIFORM = 1, STN_PRIOR_GAUSSIAN = 2, STN_PRIOR_SINE = 3 };

typedef struct StnPrior {
    int32_t n_params;            /* must equal the problem's P; one entry per theta column */
    const int32_t* kind;         /* StnPriorKind */
    const double* a;
    const double* b;
    int32_t n_sin_limits;        /* lo <= sin(theta[col]) <= hi */
    const int32_t* sin_col;
    const double* sin_lo;
    const double* sin_hi;
} StnPrior;

typedef struct StnCtx StnCtx;

int stn_version(void);
const char* stn_last_error(void);

/* Copies the problem to `device` (epoch tables, routing, constraints).
 * Replaces: Gaussianlikelihood.__init__ (simulate.py:326-358). */
int stn_create(const StnProblem* problem, int device, StnCtx** out_ctx);
int stn_destroy(StnCtx* ctx);

/* Batched log-likelihood, device buffers.
 * Replaces: Gaussianlikelihood.log_likelihood (simulate.py:362-384) evaluated for
 * n parameter vectors.  theta_dev: n vectors of P doubles in `layout` with leading
 * dimension ld; out_dev: n doubles.
 * plan = 0 lets the library choose the execution plan from n; otherwise
 * plan = lanes | (splits << 8): `lanes` (1,2,4,...,32) lanes of a warp share one sample,
 * `splits` (1..number of 128-epoch chunks) CTAs share one tile of samples and their partial
 * sums are added in a fixed order by a second small kernel.  The plan fixes the summation
 * order, i.e. the result bits: pass the plan chosen for the WHOLE batch (stn_auto_plan) to
 * make results independent of how a batch is sharded.  Asynchronous on `stream`
 * (a cudaStream_t, NULL = default stream); one stream at a time per context. */
int stn_loglike(StnCtx* ctx, const double* theta_dev, int64_t n, int64_t ld,
                int layout, int mode, int plan, double* out_dev, void* stream);

/* stn_loglike with the result stored to n_outs buffers at once (1 <= n_outs <= STN_MAX_OUTS):
 * outs[0] on this context's device, the others typically buffers of PEER GPUs (mapped with
 * stn_ipc_open, or device pointers of peers with peer access enabled).  This is the multi-GPU
 * gather of SURVEY 8(e) fused into the kernel: every GPU evaluates its slice of the batch and
 * its epilogue stores each log L (8 bytes per vector) over NVLink directly into the slice's place
 * in every rank's (N,) result buffer -- no collective, no second pass; the ranks only need a
 * barrier before reading. */
int stn_loglike_scatter(StnCtx* ctx, const double* theta_dev, int64_t n, int64_t ld,
                        int layout, int mode, int plan, double* const* outs, int n_outs,
                        void* stream);

/* Plain device memory (cudaMalloc, so it can be exported) and CUDA IPC plumbing for the result
 * buffers of stn_loglike_scatter when the GPUs are driven by one process each (torchrun):
 * export a 64-byte handle, ship it to the other ranks (torch.distributed), open it there. */
#define STN_IPC_HANDLE_BYTES 64
int stn_dev_alloc(int device, void** ptr, int64_t bytes);
int stn_dev_free(int device, void* ptr);
int stn_ipc_export(void* dev_ptr, void* handle64);
int stn_ipc_open(int device, const void* handle64, void** peer_ptr);
int stn_ipc_close(int device, void* peer_ptr);

/* Installs (copies) the prior used by stn_logpost.  Replaces: priors.create_priors_given_limits_dict
 * (priors.py:344-366) evaluated per walker by bilby. */
int stn_set_prior(StnCtx* ctx, const StnPrior* prior);

/* log prior + log likelihood for n vectors; -inf outside the prior's support.  Same
 * arguments as stn_loglike.  Requires stn_set_prior. */
int stn_logpost(StnCtx* ctx, const double* theta_dev, int64_t n, int64_t ld,
                int layout, int mode, int plan, double* out_dev, void* stream);

/* sum over sources and epochs of ((obs - model)/sigma)^2 for n vectors, same arguments as
 * stn_loglike.  Replaces: the chi-square of calculate_reduced_chi_square (simulate.py:290-301). */
int stn_chisq(StnCtx* ctx, const double* theta_dev, int64_t n, int64_t ld,
              int layout, int mode, int plan, double* out_dev, void* stream);

/* Batched log-likelihood, HOST buffers: H2D copy, kernel(s), D2H copy, synchronised on return.
 * Large batches are cut into chunks whose copies overlap the kernels.  Pinned
 * buffers (stn_host_alloc) are copied at full PCIe rate. */
int stn_loglike_host(StnCtx* ctx, const double* theta_host, int64_t n, int64_t ld,
                     int layout, int mode, int plan, double* out_host);

/* Model positions of one source for n parameter vectors.
 * Replaces: positions() (positions.py:13-30) -> radec_dev[n][2E] = concat(ra[E], dec[E]) rad;
 * off_mas_dev (nullable) receives the offsets in mas (positions.py:114-118). */
int stn_model(StnCtx* ctx, int source, const double* theta_dev, int64_t n, int64_t ld,
              int layout, int mode, double* radec_dev, double* off_mas_dev, void* stream);

/* The plan stn_loglike(plan = 0) would choose for a batch of n. */
int stn_auto_plan(StnCtx* ctx, int64_t n);

/* Number of kernel launches issued through this context so far. */
int64_t stn_launch_count(StnCtx* ctx);

/* Pinned host memory for the *_host entry points (portable: every device can DMA from it).
 * stn_host_register pins a buffer the caller already owns (e.g. a sampler's numpy array) in
 * place; pageable buffers are accepted too and are staged through the library's own pinned ring. */
int stn_host_alloc(void** ptr, int64_t bytes);
int stn_host_free(void* ptr);
int stn_host_register(void* ptr, int64_t bytes);
int stn_host_unregister(void* ptr);
/* Tells the host pipeline of this context how many devices copy from this host at the same time
 * (a launcher with one process per GPU knows, the library cannot): with more than one, chunks
 * grow twofold instead of fourfold so that each copy still hides under the previous kernel. */
int stn_host_hint(StnCtx* ctx, int concurrent_devices);

/* ---- several GPUs of one box driven from ONE process ---------------------------
 * The reference is one process with one sampler (simulate.py:191-192) that owns one (N, P)
 * batch; these entry points let that process use every GPU.  Each device holds the full epoch
 * tables; the batch is cut into contiguous slices (stn_multi_slice), one per device.  There is
 * no collective: every device DMAs its slice of log L straight into the caller's host array
 * (stn_multi_loglike_host), or, for device-resident slices, into one device buffer with peer
 * copies over NVLink (stn_multi_loglike_dev).  The execution plan is chosen from the WHOLE batch,
 * so a vector's log L has the same bits whatever the number of devices.  `devices` may name a
 * device more than once (two contexts on one GPU). */
typedef struct StnMulti StnMulti;
int stn_multi_create(const StnProblem* problem, const int* devices, int n_devices, StnMulti** out);
int stn_multi_destroy(StnMulti* m);
int stn_multi_size(StnMulti* m);
StnCtx* stn_multi_ctx(StnMulti* m, int i);           /* borrowed: context of the i-th device */
int64_t stn_multi_launch_count(StnMulti* m);
/* [lo, hi) of part `part` of n rows cut into n_parts contiguous slices (the first n % n_parts
 * slices hold one extra row). */
int stn_multi_slice(int64_t n, int n_parts, int part, int64_t* lo, int64_t* hi);
/* Same contract as stn_loglike_host; slices below 2^15 rows are not worth a device, so small
 * batches use fewer devices (down to one).  One host thread per device used. */
int stn_multi_loglike_host(StnMulti* m, const double* theta_host, int64_t n, int64_t ld,
                           int layout, int mode, int plan, double* out_host);
/* theta_dev[g]: counts[g] vectors resident on the g-th device (same ld / layout for all);
 * out_dev: sum(counts) doubles on device `out_device`, slices in device order.  Synchronous. */
int stn_multi_loglike_dev(StnMulti* m, const double* const* theta_dev, const int64_t* counts,
                          int64_t ld, int layout, int mode, int plan, double* out_dev,
                          int out_device);

/* One walker ensemble advanced by ALL devices of the multi-context (what the reference does in one
 * emcee process, simulate.py:191-192, for ensembles too large for one GPU).  Every device holds a full
 * copy of the walkers, x_dev[g] (W, ld >= P) and lp_dev[g] (W), identical on entry (lp = current
 * log-posteriors).  In each half-step device g proposes and evaluates the proposals of its slice of the
 * half, and its accept kernel stores the accepted walkers into EVERY device's copy through peer-mapped
 * pointers -- the exchange of the updated half over NVLink is fused into the kernel; the devices are
 * ordered by events, not by the host.  Random numbers are keyed by global walker indices and plan = 0
 * is the plan of the WHOLE half-ensemble (stn_auto_plan(W / 2)), so the chain is bit-identical to
 * stn_ensemble_run's launch engine on one device, whatever the number of devices.
 * chain_dev / lnprob_dev (nullable) live on the first device; *n_accept_host is incremented.  Needs
 * stn_multi_set_prior and peer access between all devices (STN_ERR_DEVICE otherwise).  Synchronous. */
int stn_multi_set_prior(StnMulti* m, const StnPrior* prior);
int stn_multi_ensemble_run(StnMulti* m, double* const* x_dev, double* const* lp_dev, int64_t n_walkers,
                           int64_t ld, double a, uint64_t seed, uint64_t step0, int64_t n_steps,
                           int64_t thin, double* chain_dev, double* lnprob_dev, int64_t* n_accept_host,
                           int mode, int plan);

/* FP64 pipe peak (dependent-free DFMA streams on every SM); the roofline denominator.
 * Runs `reps` launches of `iters` FMA rounds, returns the best TFLOP/s and its time. */
int stn_fp64_peak(int device, int iters, int reps, double* tflops, double* ms);

/* Affine-invariant stretch move (Goodman & Weare 2010; what emcee runs for the reference,
 * simulate.py:191-192) on device-resident walkers, so one ensemble half-step is three launches:
 * propose -> stn_logpost -> accept.  x_dev: (n_walkers, n_params) row-major with leading
 * dimension ld; the walkers of half `half` (0: rows [0, W/2), 1: rows [W/2, W)) are moved
 * using partners drawn from the other half.  Random numbers: Philox4x32-10 keyed by `seed`,
 * counter (walker, step, half) -- reproducible and independent of the launch geometry.
 * y_dev: (W/2, n_params) proposals (leading dimension n_params), z_dev: (W/2) stretch factors. */
int stn_stretch_propose(int device, const double* x_dev, int64_t n_walkers, int32_t n_params, int64_t ld,
                        int half, double a, uint64_t seed, uint64_t step, double* y_dev, double* z_dev,
                        void* stream);
/* Accepts proposal k with probability min(1, z^(P-1) exp(lpy - lp)); updates x_dev / lp_dev in
 * place and adds the number of acceptances to *n_accept_dev. */
int stn_stretch_accept(int device, double* x_dev, double* lp_dev, int64_t n_walkers, int32_t n_params,
                       int64_t ld, int half, uint64_t seed, uint64_t step, const double* y_dev,
                       const double* z_dev, const double* lpy_dev, int64_t* n_accept_dev, void* stream);

/* n_steps full ensemble steps (both halves) of the stretch move on device-resident walkers in ONE call;
 * requires stn_set_prior.  x_dev (W, ld >= P) and lp_dev (W) are updated in place (lp_dev must hold the
 * walkers' current log-posteriors, e.g. from stn_logpost); every `thin`-th state is copied to chain_dev
 * ((n_steps / thin, W, P), nullable) and lnprob_dev ((n_steps / thin, W), nullable); acceptances are added
 * to *n_accept_dev.  Random numbers and arithmetic are those of stn_stretch_propose / stn_logpost /
 * stn_stretch_accept with step counters step0, step0 + 1, ...: the chain does not depend on the engine.
 * engine 0 chooses: short series (<= 512 epochs over all sources; the reference's real-data regime,
 * simulate.py:191-192 with ~10 epochs per source) run as ONE persistent cooperative kernel with the
 * likelihood evaluated inline and a grid-wide barrier between half-steps (no launch per step at all);
 * long series use the throughput kernel, three launches per half-step issued from this call.
 * engine 1 / 2 force the persistent kernel / the launch loop.  Asynchronous on `stream`. */
int stn_ensemble_run(StnCtx* ctx, double* x_dev, double* lp_dev, int64_t n_walkers, int64_t ld, double a,
                     uint64_t seed, uint64_t step0, int64_t n_steps, int64_t thin, double* chain_dev,
                     double* lnprob_dev, int64_t* n_accept_dev, int mode, int plan, int engine, void* stream);

/* Timing helper for bench.py: runs `steps` back-to-back stn_loglike launches on an
 * internal stream bracketed by CUDA events on that stream; returns total ms. */
int stn_loglike_timed(StnCtx* ctx, const double* theta_dev, int64_t n, int64_t ld,
                      int layout, int mode, int plan, double* out_dev, int steps,
                      double* total_ms);

#ifdef __cplusplus
}
#endif
#endif /* STERNE_B200_H */
