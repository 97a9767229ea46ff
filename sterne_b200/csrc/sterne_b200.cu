// sterne_b200.cu -- B200 (sm_100a) kernels and C ABI for the astrometric log-likelihood.
//
// Path replaced (reference file:line, /root/reference/src/sterne/):
//   simulate.py:362-384   Gaussianlikelihood.log_likelihood
//   simulate.py:303-320   adjust_errs_with_efac
//   model/positions.py:13-121,266-310   positions / position / parallax offset
//   model/reflex_motion.py:201-214      (Omega_asc, incl) rotation of the orbit terms
//   model/kopeikin_effects.py:11-45     a1dot from proper motion
//
// Design (DESIGN.md has the long form):
//   * one thread owns R samples (R = 3 for EFAC-heavy problems, 2 otherwise, 1 for small batches) and
//     L lanes of a warp may share one sample (L = 1..32, epochs interleaved, fixed-order shuffle sum);
//   * per (sample, source) prologue turns the 9 routed parameters into FMA coefficients
//     (two sincos + one reciprocal); the epoch loop is then pure DFMA/DMUL/DADD -- with EFAC the
//     chi-square sum is carried as a running fraction T/Q, so there is no reciprocal in the loop
//     and Q doubles as the log-determinant;
//   * the parameter-independent epoch table streams L2 -> shared memory through the
//     bulk async-copy engine (cp.async.bulk + mbarrier, SASS: UBLKCP), double buffered,
//     and every lane reads the same row (shared-memory broadcast, LDS.128);
//   * the epilogue can store each log L into several buffers at once (peer GPUs): the multi-GPU
//     gather is part of the kernel;
//   * FP64 throughout, no tensor cores (not a contraction).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 (see build.py).

#include "sterne_b200.h"

#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <math_constants.h>
// -DSTN_DEBUG (python -m sterne_b200.build --debug): device-side asserts on every index the kernels
// form, and host-side checks that every device buffer a launch is given (the caller's and the
// library's own) lies inside ONE allocation of sufficient size (cuMemGetAddressRange).  This pool
// has no compute-sanitizer; the GPU test-suite is run once under this build instead.
#ifdef STN_DEBUG
#include <cuda.h>
#include <cassert>
#define STN_ASSERT(x) assert(x)
#else
#define STN_ASSERT(x) ((void)0)
#endif
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstdarg>
#include <cstring>
#include <cmath>
#include <new>
#include <string>
#include <thread>
#include <vector>

// ---------------------------------------------------------------------------------
// constants (astropy unit factors; see sterne_b200/constants.py for provenance)
// ---------------------------------------------------------------------------------
// defaults of the three unit factors the kernels use; StnProblem may override them (the host takes
// them from astropy.units when astropy is importable, so that the last bit matches the reference's)
#define STN_MAS2RAD   4.84813681109536e-09          /* (pi/180)/3600*0.001 */
#define STN_DEG2RAD   0.017453292519943295
#define STN_MASYR2RADS (4.84813681109536e-09 / (365.25 * 86400.0))
#define STN_LN2       0.6931471805599453

// build-time tuning knobs (defaults are the measured best; see profiles/)
// Samples per thread for large batches (R never changes a result bit).  Measured on B200 at
// N = 2^22 x 8 x 1024: without EFAC R = 2 (128 registers, 2 CTAs of 256 threads per SM) wins,
// 30.5 vs 32.4 ms; with EFAC R = 3 (204 registers, no spills), 47.0 ms.
#ifndef STN_RBIG
#define STN_RBIG 2
#endif
#ifndef STN_RBIG_EFAC
#define STN_RBIG_EFAC 3
#endif
// CTA shape of the R >= 3 kernel.  Eight warps per SM is all 204 registers allow; as FOUR CTAs of two
// warps (each with its own 2-stage table pipeline) instead of one CTA of eight, a warp waits at the
// per-chunk barrier for one other warp only and the CTAs drift apart, so one CTA's prologue (sincos,
// parameter loads) overlaps the others' epoch loops: 47.9 -> 47.0 ms (profiles/r02_kvariants.json).
#ifndef STN_THREADS_R3
#define STN_THREADS_R3 64
#endif
#ifndef STN_MINBLOCKS_R3
#define STN_MINBLOCKS_R3 4
#endif
#ifndef STN_CHUNK
#define STN_CHUNK 128
#endif
#ifndef STN_PAIR_UNROLL
#define STN_PAIR_UNROLL 1              /* unroll factor of the two-epoch loop body, EFAC off */
#endif
#ifndef STN_EFAC_UNROLL
#define STN_EFAC_UNROLL 2              /* unroll factor of the EFAC epoch loop */
#endif
// EFAC: exponent budget (bits) the running product of variances may drift between two
// renormalisations; the double range is +-1022 and T stays within 2^-100 .. 2^20 of Q.
#ifndef STN_SAMPLER_SOURCE_LANES
#define STN_SAMPLER_SOURCE_LANES 1     /* fused sampler: spread the sources of a joint fit over lanes of a warp */
#endif
#ifndef STN_SAMPLER_CLUSTER
#define STN_SAMPLER_CLUSTER 1          /* fused sampler: one thread-block cluster + hardware barrier when <= 16 CTAs */
#endif
#ifndef STN_ROW_PREFETCH
#define STN_ROW_PREFETCH 0             /* EFAC loop: fetch the next epoch's row while the current one is used */
#endif
#ifndef STN_BIGR_MIN_TENTH_WAVES
#define STN_BIGR_MIN_TENTH_WAVES 8     /* the big-R kernel is used once its CTAs fill this many tenths of a wave */
#endif
#ifndef STN_TAIL_SMALL_R
#define STN_TAIL_SMALL_R 1             /* finish the last, partly filled wave of a big-R launch with a small-tile kernel */
#endif
#ifndef STN_TAIL_R
#define STN_TAIL_R 2                   /* samples per thread of that tail kernel (64-thread CTAs) */
#endif
#ifndef STN_TAIL_MIN_PCT
#define STN_TAIL_MIN_PCT 10            /* use it when the leftover fills at least this many percent of a wave */
#endif
#ifndef STN_RENORM_BITS
#define STN_RENORM_BITS 900
#endif

// One translation unit, several files:
#include "stn_kernels_likelihood.cuh"   // device: structs, epoch loops, K1 / K1s / K2 / K4
#include "stn_kernels_sampler.cuh"      // device: Philox, stretch move, persistent ensemble kernel K5
#include "stn_context.inl"              // host: errors, StnCtx / StnMulti, plans, launches
#include "stn_host_pipeline.inl"        // host: chunked H2D / kernel / D2H pipeline
#include "stn_api_core.inl"             // extern "C": contexts, launches, priors, tooling
#include "stn_api_multi.inl"            // extern "C": several devices, one process
#include "stn_api_sampler.inl"          // extern "C": sampler on one device
