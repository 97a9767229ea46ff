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
//   * one thread owns R samples (R = 1 or 2) and L lanes of a warp may share one sample
//     (L = 1..32, epochs interleaved across the lanes, fixed-order shuffle reduction);
//   * per (sample, source) prologue turns the 9 routed parameters into FMA coefficients
//     (two sincos + one reciprocal); the epoch loop is then pure DFMA/DMUL/DADD;
//   * the parameter-independent epoch table streams L2 -> shared memory through the
//     bulk async-copy engine (cp.async.bulk + mbarrier, SASS: UBLKCP), double buffered,
//     and every lane reads the same row (shared-memory broadcast, LDS.128);
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
#ifndef STN_SAMPLER_CLUSTER
#define STN_SAMPLER_CLUSTER 1          /* fused sampler: one thread-block cluster + hardware barrier when <= 16 CTAs */
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

namespace {

constexpr int kThreads = 256;      // threads per CTA of the likelihood kernels
constexpr int kChunk = STN_CHUNK;  // epochs per shared-memory stage (128 * 96 B = 12 KiB)
constexpr int kStages = 2;
constexpr size_t kStageBytes = (size_t)kStages * kChunk * STN_FAST_ROW * sizeof(double);
#define STN_DYNAMIC_STAGE (STN_CHUNK > 240)          /* 2 x 256 x 96 B + barriers exceeds the 48 KB static limit */
constexpr size_t kStageSmem = STN_DYNAMIC_STAGE ? kStageBytes + kStages * sizeof(uint64_t) : 0;
constexpr int kPairUnroll = STN_PAIR_UNROLL;
constexpr int kEfacUnroll = STN_EFAC_UNROLL;
constexpr int kFirst = 1, kLast = 2;

struct SrcDev {
    const double* fast_rows;
    const double* strict_rows;
    int n_epochs;
    int col[STN_NROOTS];
    int efac_on;
    int binary;
    int v_shift;          // EFAC: the variance columns of fast_rows are stored times 2^v_shift
    int v_lo_bits;        // EFAC: every (scaled) variance is >= 2^-v_lo_bits
    double v_hi, v_hib;   // EFAC: every (scaled) variance is <= v_hi + g * v_hib, g = max(g_ra, g_dec)
    double chi_scale;     // EFAC: 2^v_shift, turns the scaled chi-square back
    double cos_decj;
    double inv_cos_decj;
    double neg_sum_log_err;
    double a1_lts;
};

struct ChunkDev {
    const double* rows;   // device pointer into the source's fast table
    int n_epochs;
    int source;
    int flags;            // kFirst | kLast
    int pad;
};

struct ProbDev {
    int n_sources, n_params, n_chunks, n_a1dot, n_sini;
    int n_prior, n_sinlim;            // n_prior = 0: likelihood only
    int out_chisq;                    // 1: write sum((res/sigma)^2) instead of log L (stn_chisq)
    double mas2rad, deg2rad, masyr2rads;   // unit factors (positions.py:119-120, reflex_motion.py:170, kopeikin_effects.py:45)
    int n_extra_out;                  // stn_loglike_scatter: the result also goes to these buffers (peer GPUs)
    double* extra_out[STN_MAX_OUTS - 1];
    double prior_const;               // sum of the priors' normalisation terms
    const int* prior_kind;
    const double* prior_a;
    const double* prior_b;
    const int* sinlim_col;
    const double* sinlim_lo;
    const double* sinlim_hi;
    const SrcDev* src;
    const ChunkDev* chunks;
    const int* a1_src;
    const double* a1_mu;
    const double* a1_sigma;
    const int* sini_col;
    const double* sini_mu;
    const double* sini_sigma;
};

// ---------------------------------------------------------------------------------
// small PTX wrappers: mbarrier + bulk async copy (TMA engine, 1-D form)
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// fill one stage with `n_epochs` rows (an empty chunk just completes the phase)
__device__ __forceinline__ void stage_fill(double* dst, const double* src, int n_epochs, uint64_t* bar) {
    const uint32_t bytes = (uint32_t)n_epochs * STN_FAST_ROW * (uint32_t)sizeof(double);
    STN_ASSERT(n_epochs >= 0 && n_epochs <= kChunk);
    if (bytes != 0) {
        mbar_expect_tx(bar, bytes);
        bulk_g2s(dst, src, bytes, bar);
    } else {
        mbar_arrive(bar);
    }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    const uint32_t addr = smem_u32(bar);
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}

// ---------------------------------------------------------------------------------
// parameter fetch (routing table resolved on the host: positions.py:32-60)
// ---------------------------------------------------------------------------------
// COH: plain (coherent, generic-address) loads instead of the read-only path -- for parameter vectors that
// the running kernel itself has written (the fused sampler's proposals, which may live in local memory)
template <bool COH = false>
__device__ __forceinline__ double load_param(const double* __restrict__ theta, int64_t n, int col,
                                             int64_t ld, int layout) {
    if (col < 0) return STN_SENTINEL;
    STN_ASSERT(n >= 0 && (layout == STN_LAYOUT_SOA ? n < ld : col < ld));
    const double* p = layout == STN_LAYOUT_SOA ? theta + (int64_t)col * ld + n : theta + n * ld + col;
    if (COH) return *reinterpret_cast<const volatile double*>(p);
    return __ldg(p);
}

struct Params {
    double dec, efac, efad, incl, mu_a, mu_d, om_asc, px, ra;
};

template <bool COH = false>
__device__ __forceinline__ Params load_params(const SrcDev& S, const double* __restrict__ theta,
                                              int64_t n, int64_t ld, int layout) {
    Params p;
    p.dec = load_param<COH>(theta, n, S.col[STN_DEC], ld, layout);
    p.efac = load_param<COH>(theta, n, S.col[STN_EFAC], ld, layout);
    p.efad = load_param<COH>(theta, n, S.col[STN_EFAD], ld, layout);
    p.incl = load_param<COH>(theta, n, S.col[STN_INCL], ld, layout);
    p.mu_a = load_param<COH>(theta, n, S.col[STN_MU_A], ld, layout);
    p.mu_d = load_param<COH>(theta, n, S.col[STN_MU_D], ld, layout);
    p.om_asc = load_param<COH>(theta, n, S.col[STN_OM_ASC], ld, layout);
    p.px = load_param<COH>(theta, n, S.col[STN_PX], ld, layout);
    p.ra = load_param<COH>(theta, n, S.col[STN_RA], ld, layout);
    return p;
}

// ---------------------------------------------------------------------------------
// FAST form: per-(sample, source) coefficients, mas->rad folded in.
//   offK_ra  = dT*c0 + X*c1 + Y*c2 [+ Bc*r0c + Bs*r0s]
//   offK_dec = dT*mud + X*d1 + Y*d2 + Z*d3 [+ Bc*r1c + Bs*r1s]
// which is positions.py:110-111,308-309 and reflex_motion.py:201-214 with the
// divisions and trig hoisted out of the epoch loop.
// ---------------------------------------------------------------------------------
struct Coef {
    double ra, dec;
    double c0, c1, c2;
    double mud, d1, d2, d3;
    double r0c, r0s, r1c, r1s;
    double g_ra, g_dec;   // EFAC: efac^2 and (efac*efad)^2
};

__device__ __forceinline__ Coef make_coef(const ProbDev& P, const SrcDev& S, const Params& p) {
    Coef c;
    const double K = P.mas2rad;
    double sd, cd, sa, ca;
    sincos(p.dec, &sd, &cd);
    sincos(p.ra, &sa, &ca);
    const double icd = 1.0 / cd;
    const bool px_on = (p.px != STN_SENTINEL);            // positions.py:115
    const double pxe = px_on ? p.px : 0.0;
    c.ra = p.ra;
    c.dec = p.dec;
    c.c0 = (p.mu_a * icd) * K;
    c.c1 = (pxe * sa * icd) * K;
    c.c2 = -(pxe * ca * icd) * K;
    c.mud = p.mu_d * K;
    c.d1 = (pxe * ca * sd) * K;
    c.d2 = (pxe * sa * sd) * K;
    c.d3 = -(pxe * cd) * K;
    const bool rf_on = px_on && S.binary && (p.incl != STN_SENTINEL) &&
                       (p.om_asc != STN_SENTINEL);        // positions.py:117
    if (rf_on) {
        double sO, cO;
        sincos(p.om_asc * P.deg2rad, &sO, &cO);
        const double ci = cos(p.incl);
        const double pk = pxe * K;
        c.r0c = pk * sO * S.inv_cos_decj;
        c.r0s = pk * (cO * ci) * S.inv_cos_decj;
        c.r1c = pk * cO;
        c.r1s = -(pk * (sO * ci));
    } else {
        c.r0c = c.r0s = c.r1c = c.r1s = 0.0;
    }
    if (S.efac_on) {
        // simulate.py:311-317; a sampled efac of exactly -999 switches EFAC off for that
        // vector (errs^2 is then err_random^2 + err_sys^2 up to rounding).
        const double ef = (p.efac != STN_SENTINEL) ? p.efac : 1.0;
        const double ed = (p.efad != STN_SENTINEL && p.efac != STN_SENTINEL) ? p.efad : 1.0;
        c.g_ra = ef * ef;
        const double t = ef * ed;
        c.g_dec = t * t;
    } else {
        c.g_ra = c.g_dec = 1.0;
    }
    return c;
}

template <bool BIN>
__device__ __forceinline__ void fast_offsets(const Coef& c, const double2 tx, const double2 yz,
                                             const double2 bc, double& ora, double& odec) {
    ora = tx.x * c.c0;
    ora = fma(tx.y, c.c1, ora);
    ora = fma(yz.x, c.c2, ora);
    odec = tx.x * c.mud;
    odec = fma(tx.y, c.d1, odec);
    odec = fma(yz.x, c.d2, odec);
    odec = fma(yz.y, c.d3, odec);
    if (BIN) {
        ora = fma(bc.x, c.r0c, ora);
        ora = fma(bc.y, c.r0s, ora);
        odec = fma(bc.x, c.r1c, odec);
        odec = fma(bc.y, c.r1s, odec);
    }
}

// EFAC accumulators of one (sample, source).  The chi-square sum  sum_i s_i / p_i  (one term per
// epoch, p_i = v_ra v_dec, s_i = r_ra^2 v_dec + r_dec^2 v_ra) is carried as ONE running fraction
// T / Q with  T <- T p_i + s_i Q,  Q <- Q p_i : no reciprocal in the epoch loop at all, and Q is at
// the same time the product of all variances, i.e. the log-determinant (sum ln sigma = 1/2 ln Q).
// Both are kept in range by multiplying T and Q with the same power of two every few epochs
// (renorm); that scaling is exact, so how often it happens never changes a result bit.
struct Acc {
    double chi_ra, chi_dec;   // EFAC off: sum of (res/sigma)^2 per coordinate.  EFAC on: chi_ra = T
    double prod;              // EFAC on: Q
    int esum;                 // EFAC on: exponents taken out of Q so far (unbiased)
};

// Residuals of one epoch of one sample, and for EFAC sources the two variances.
template <bool EF, bool BIN>
__device__ __forceinline__ void epoch_residuals(const double* __restrict__ rows, int e, const Coef& cf,
                                                double& res_ra, double& res_dec, double& v_ra, double& v_dec,
                                                double2& aa) {
    STN_ASSERT(e >= 0 && e < kChunk);
    const double2* row = reinterpret_cast<const double2*>(rows + e * STN_FAST_ROW);
    const double2 tx = row[0];
    const double2 yz = row[1];
    const double2 ob = row[2];
    aa = row[3];
    double2 bc = make_double2(0.0, 0.0);
    if (BIN) bc = row[5];
    double ora, odec;
    fast_offsets<BIN>(cf, tx, yz, bc, ora, odec);
    // model = ref + offset (rad); residual = obs - model   (positions.py:119-120,
    // simulate.py:368).  This order is kept: the subtraction cancels ~9 digits.
    res_ra = ob.x - (cf.ra + ora);
    res_dec = ob.y - (cf.dec + odec);
    if (EF) {
        const double2 bb = row[4];
        v_ra = fma(cf.g_ra, bb.x, aa.x);                  // simulate.py:317 (table pre-scaled by 2^k)
        v_dec = fma(cf.g_dec, bb.y, aa.y);
    }
}

// One epoch of one sample.
template <bool EF, bool BIN>
__device__ __forceinline__ void epoch_term(const double* __restrict__ rows, int e, const Coef& cf, Acc& acc) {
    double res_ra, res_dec, v_ra = 1.0, v_dec = 1.0;
    double2 aa;
    epoch_residuals<EF, BIN>(rows, e, cf, res_ra, res_dec, v_ra, v_dec, aa);
    if (!EF) {
        const double z_ra = res_ra * aa.x;
        const double z_dec = res_dec * aa.y;
        acc.chi_ra = fma(z_ra, z_ra, acc.chi_ra);
        acc.chi_dec = fma(z_dec, z_dec, acc.chi_dec);
    } else {
        const double p = v_ra * v_dec;
        double s = (res_ra * res_ra) * v_dec;
        s = fma(res_dec * res_dec, v_ra, s);
        const double sq = s * acc.prod;
        acc.chi_ra = fma(acc.chi_ra, p, sq);              // T <- T p + s Q   (one FMA on the T chain)
        acc.prod = acc.prod * p;                          // Q <- Q p
    }
}

// Take the exponent out of Q and apply the same power of two to T (exact).
__device__ __forceinline__ void renorm(Acc& acc) {
    const int hi = __double2hiint(acc.prod);
    const int ex = (hi >> 20) & 0x7ff;                    // biased exponent of Q (> 0: Q is a positive normal)
    acc.esum += ex - 1023;
    acc.prod = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(acc.prod));
    acc.chi_ra *= __hiloint2double((2046 - ex) << 20, 0); // 2^(1023 - ex)
}

// Epochs sub, sub + L, ... of one staged chunk for the R samples of this thread.  `mask` + 1 is the
// (warp-uniform) number of epochs between two renormalisations of the EFAC accumulators.
template <int R, int L, bool EF, bool BIN>
__device__ __forceinline__ void epoch_loop(const double* __restrict__ rows, int n_epochs, int sub,
                                           const Coef (&cf)[R], Acc (&acc)[R], int mask) {
    int e = sub;
    if constexpr (EF) {
        while (e < n_epochs) {
            const int span = (mask + 1) * L;
            const int stop = (n_epochs - e > span) ? e + span : n_epochs;
#pragma unroll kEfacUnroll
            for (; e < stop; e += L) {
#pragma unroll
                for (int r = 0; r < R; ++r) epoch_term<true, BIN>(rows, e, cf[r], acc[r]);
            }
#pragma unroll
            for (int r = 0; r < R; ++r) renorm(acc[r]);
        }
    } else {
#pragma unroll kPairUnroll
        for (; e + L < n_epochs; e += 2 * L) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                epoch_term<false, BIN>(rows, e, cf[r], acc[r]);
                epoch_term<false, BIN>(rows, e + L, cf[r], acc[r]);
            }
        }
        if (e < n_epochs) {
#pragma unroll
            for (int r = 0; r < R; ++r) epoch_term<false, BIN>(rows, e, cf[r], acc[r]);
        }
    }
}

// Epochs between two renormalisations for this sample: the largest power of two (minus one, as a
// mask) such that the product of that many epochs' variance pairs stays within STN_RENORM_BITS
// binary orders of magnitude.  Any variance of the source lies in [v_lo, v_hi + g v_hib].
__device__ __forceinline__ int renorm_mask(const SrcDev& S, const Coef& c) {
    const double g = fmax(c.g_ra, c.g_dec);
    const double vmax = fma(g, S.v_hib, S.v_hi);
    const int ex = ((__double2hiint(vmax) >> 20) & 0x7ff) - 1023;          // inf / nan -> 1024
    const int up = ex >= 0 ? ex + 1 : 0;
    const int bits = 2 * (up > S.v_lo_bits ? up : S.v_lo_bits);            // one epoch = two variances
    if (bits <= 0) return kChunk - 1;
    int m = STN_RENORM_BITS / bits;
    if (m < 1) m = 1;
    if (m > kChunk) m = kChunk;
    return (1 << (31 - __clz(m))) - 1;
}

template <int L>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// a1dot (kopeikin_effects.py:41-45) and sin(incl) (simulate.py:379-383) Gaussian terms
template <bool COH = false>
__device__ double constraint_terms(const ProbDev& P, const double* __restrict__ theta, int64_t n,
                                   int64_t ld, int layout) {
    double add = 0.0;
    if (P.n_a1dot > 0) {
        double s = 0.0;
        for (int k = 0; k < P.n_a1dot; ++k) {
            const SrcDev& S = P.src[P.a1_src[k]];
            const double inc = load_param<COH>(theta, n, S.col[STN_INCL], ld, layout);
            const double mu_a = load_param<COH>(theta, n, S.col[STN_MU_A], ld, layout);
            const double mu_d = load_param<COH>(theta, n, S.col[STN_MU_D], ld, layout);
            const double om = load_param<COH>(theta, n, S.col[STN_OM_ASC], ld, layout);
            const double mu = sqrt(mu_a * mu_a + mu_d * mu_d);
            const double th = atan2(mu_a, mu_d);
            double v = mu * sin(th - om * P.deg2rad) / tan(inc);
            v = (v * S.a1_lts) * P.masyr2rads;
            const double z = (v - P.a1_mu[k]) / P.a1_sigma[k];
            s += z * z;
        }
        add += -0.5 * s;
    }
    for (int k = 0; k < P.n_sini; ++k) {
        const double x = load_param<COH>(theta, n, P.sini_col[k], ld, layout);
        const double z = (P.sini_mu[k] - sin(x)) / P.sini_sigma[k];
        add += -0.5 * (z * z);
    }
    return add;
}

// log prior of one parameter vector (priors.py:356-366; -inf outside the support)
template <bool COH = false>
__device__ double prior_terms(const ProbDev& P, const double* __restrict__ theta, int64_t n, int64_t ld,
                              int layout) {
    double lp = P.prior_const;
    bool inside = true;
    for (int k = 0; k < P.n_prior; ++k) {
        const double x = load_param<COH>(theta, n, k, ld, layout);
        const double a = P.prior_a[k], b = P.prior_b[k];
        const int kind = P.prior_kind[k];
        if (kind == STN_PRIOR_GAUSSIAN) {
            const double z = (x - a) / b;
            lp += -0.5 * (z * z);
        } else {
            inside = inside && (x >= a) && (x <= b);
            if (kind == STN_PRIOR_SINE) lp += log(sin(x));
        }
    }
    for (int k = 0; k < P.n_sinlim; ++k) {
        const double sx = sin(load_param<COH>(theta, n, P.sinlim_col[k], ld, layout));
        inside = inside && (sx >= P.sinlim_lo[k]) && (sx <= P.sinlim_hi[k]);
    }
    return inside ? lp : -CUDART_INF;
}

// ---------------------------------------------------------------------------------
// Per-source building blocks shared by K1 and the fused sampler kernel: begin (parameters ->
// coefficients, empty sums), chunk (epoch loop over one chunk of the table), fold (sums -> log L).
// ---------------------------------------------------------------------------------
template <int R>
struct SrcState {
    Coef cf[R];
    Acc acc[R];
    int mask;     // EFAC: epochs between two renormalisations, minus one (warp-uniform)
    int nvar;     // EFAC: epochs this lane has folded into Q since src_begin
};

template <int R, bool COH>
__device__ __forceinline__ void src_begin(const ProbDev& P, const SrcDev& S, const double* __restrict__ theta,
                                          const int64_t (&nidx)[R], int64_t ld, int layout, SrcState<R>& st) {
    const bool ef = S.efac_on != 0;
    int m = kChunk - 1;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const Params p = load_params<COH>(S, theta, nidx[r], ld, layout);
        st.cf[r] = make_coef(P, S, p);
        st.acc[r].chi_ra = 0.0;
        st.acc[r].chi_dec = 0.0;
        st.acc[r].prod = 1.0;
        st.acc[r].esum = 0;
        if (ef) m = min(m, renorm_mask(S, st.cf[r]));
    }
    // the cadence only has to be uniform over the warp (the epoch loops stay converged); it
    // never changes a bit of the result
    st.mask = ef ? __reduce_min_sync(0xffffffffu, m) : m;
    st.nvar = 0;
}

template <int R, int L>
__device__ __forceinline__ void src_chunk(const SrcDev& S, const double* __restrict__ rows, int n_epochs, int sub,
                                          SrcState<R>& st) {
    const bool ef = S.efac_on != 0;
    const bool bin = S.binary != 0;
    if (!ef && !bin) epoch_loop<R, L, false, false>(rows, n_epochs, sub, st.cf, st.acc, st.mask);
    else if (!ef && bin) epoch_loop<R, L, false, true>(rows, n_epochs, sub, st.cf, st.acc, st.mask);
    else if (ef && !bin) epoch_loop<R, L, true, false>(rows, n_epochs, sub, st.cf, st.acc, st.mask);
    else epoch_loop<R, L, true, true>(rows, n_epochs, sub, st.cf, st.acc, st.mask);
    if (ef) st.nvar += sub < n_epochs ? (n_epochs - sub + L - 1) / L : 0;
}

// log_p += -0.5*sum((res/errs)^2); log_p += -sum(log(errs))  (simulate.py:370-371)
template <int R, int L>
__device__ __forceinline__ void src_fold(const ProbDev& P, const SrcDev& S, bool source_ends, SrcState<R>& st,
                                         double (&logl)[R]) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
        double chi, logdet;
        if (S.efac_on) {
            // chi^2 = 2^k T / Q;  -sum(log sigma) = -0.5 log(prod of variances), the table's
            // variances being stored times 2^k (two per epoch)
            renorm(st.acc[r]);                                   // Q into [1, 2): log() sees the same bits
            chi = group_sum<L>((st.acc[r].chi_ra / st.acc[r].prod) * S.chi_scale);
            const double part = log(st.acc[r].prod) + (double)(st.acc[r].esum - 2 * st.nvar * S.v_shift) * STN_LN2;
            logdet = -0.5 * group_sum<L>(part);
        } else {
            chi = group_sum<L>(st.acc[r].chi_ra + st.acc[r].chi_dec);
            // host constant, counted once: by the split that holds the source's last chunk
            logdet = source_ends ? S.neg_sum_log_err : 0.0;
        }
        if (P.out_chisq) {
            logl[r] += chi;
        } else {
            logl[r] += -0.5 * chi;
            logl[r] += logdet;
        }
    }
}

// every result is stored to `out` and to the extra (typically peer-mapped) buffers of a scatter launch
__device__ __forceinline__ void store_result(const ProbDev& P, double* __restrict__ out, int64_t n, double v) {
    out[n] = v;
    for (int k = 0; k < P.n_extra_out; ++k) P.extra_out[k][n] = v;
}

// ---------------------------------------------------------------------------------
// K1: fused FAST log-likelihood
// ---------------------------------------------------------------------------------
#ifndef STN_THREADS_R2
#define STN_THREADS_R2 256
#endif
#ifndef STN_MINBLOCKS_R2
#define STN_MINBLOCKS_R2 2
#endif
__host__ __device__ constexpr int threads_for(int R) { return R > 2 ? STN_THREADS_R3 : (R == 2 ? STN_THREADS_R2 : kThreads); }
__host__ __device__ constexpr int minblocks_for(int R) { return R > 2 ? STN_MINBLOCKS_R3 : (R == 2 ? STN_MINBLOCKS_R2 : 2); }

// TPB: threads per CTA (default: the shape chosen for R); TPB = kTailThreads is the small-CTA form of the
// R = 1 kernel that finishes the last, partly filled wave of a big launch
constexpr int kTailThreads = 64;
__host__ __device__ constexpr int minblocks_tpb(int R, int tpb) { return tpb == threads_for(R) ? minblocks_for(R) : 8; }

template <int R, int L, int TPB = threads_for(R)>
__global__ void __launch_bounds__(TPB, minblocks_tpb(R, TPB))
stn_loglike_fast_kernel(const ProbDev P, const double* __restrict__ theta, const int64_t N,
                        const int64_t ld, const int layout, double* __restrict__ out,
                        double* __restrict__ partial) {
    // two stages of the epoch table + their mbarriers: static up to 48 KB, dynamic beyond (256-epoch chunks)
#if STN_DYNAMIC_STAGE
    extern __shared__ __align__(128) unsigned char stage_smem[];
    double (*stage)[kChunk * STN_FAST_ROW] = reinterpret_cast<double (*)[kChunk * STN_FAST_ROW]>(stage_smem);
    uint64_t* full = reinterpret_cast<uint64_t*>(stage_smem + kStageBytes);
#else
    __shared__ __align__(128) double stage[kStages][kChunk * STN_FAST_ROW];
    __shared__ __align__(8) uint64_t full[kStages];
#endif

    const int tid = threadIdx.x;
    const int sub = tid % L;
    const int grp = tid / L;
    constexpr int G = TPB / L;                          // sample groups per CTA
    const int64_t tile_base = (int64_t)blockIdx.x * (G * R);

    int64_t nidx[R];
    bool valid[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int64_t n = tile_base + (int64_t)r * G + grp;
        valid[r] = n < N;
        nidx[r] = valid[r] ? n : (N - 1);
        STN_ASSERT(nidx[r] >= 0 && nidx[r] < N);
    }

    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    // gridDim.y CTAs share one tile of samples: CTA k walks chunks [c_begin, c_end) and the
    // per-split partial sums are added in split order by stn_reduce_splits_kernel.
    const int K = (int)gridDim.y;
    const int c_begin = (int)(((int64_t)P.n_chunks * blockIdx.y) / K);
    const int c_end = (int)(((int64_t)P.n_chunks * (blockIdx.y + 1)) / K);
    if (tid == 0 && c_begin < c_end) {
        const ChunkDev c0 = P.chunks[c_begin];
        stage_fill(stage[0], c0.rows, c0.n_epochs, &full[0]);
    }

    double logl[R];
    SrcState<R> st;
    st.mask = kChunk - 1;
    st.nvar = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) logl[r] = 0.0;

    for (int c = c_begin; c < c_end; ++c) {
        STN_ASSERT(c >= 0 && c < P.n_chunks);
        const ChunkDev ck = P.chunks[c];
        STN_ASSERT(ck.source >= 0 && ck.source < P.n_sources && ck.n_epochs <= kChunk);
        STN_ASSERT(ck.rows >= P.src[ck.source].fast_rows &&
                   ck.rows + (size_t)ck.n_epochs * STN_FAST_ROW <=
                       P.src[ck.source].fast_rows + (size_t)P.src[ck.source].n_epochs * STN_FAST_ROW);
        const int it = c - c_begin;                      // stage = it & 1, phase parity = (it >> 1) & 1
        if (tid == 0 && c + 1 < c_end) {
            // stage (it+1)&1 was last read in iteration it-1; the __syncthreads that ended it
            // ordered those reads before this refill.
            const ChunkDev nx = P.chunks[c + 1];
            stage_fill(stage[(it + 1) & 1], nx.rows, nx.n_epochs, &full[(it + 1) & 1]);
        }
        const SrcDev& S = P.src[ck.source];
        const bool first = (ck.flags & kFirst) || c == c_begin;   // (re)start this source's sums
        const bool last = (ck.flags & kLast) || c == c_end - 1;   // fold them into log L
        if (first) src_begin<R, false>(P, S, theta, nidx, ld, layout, st);
        mbar_wait(&full[it & 1], (uint32_t)((it >> 1) & 1));
        src_chunk<R, L>(S, stage[it & 1], ck.n_epochs, sub, st);
        if (last) src_fold<R, L>(P, S, (ck.flags & kLast) != 0, st, logl);
        __syncthreads();
    }

    if (K > 1) {
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (valid[r] && sub == 0) {
                STN_ASSERT(partial != nullptr && blockIdx.y < gridDim.y);
                partial[(int64_t)blockIdx.y * N + nidx[r]] = logl[r];
            }
        return;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (!P.out_chisq && (P.n_a1dot > 0 || P.n_sini > 0)) logl[r] += constraint_terms(P, theta, nidx[r], ld, layout);
        if (!P.out_chisq && P.n_prior > 0) {
            const double lp = prior_terms(P, theta, nidx[r], ld, layout);
            logl[r] = (lp == -CUDART_INF) ? lp : logl[r] + lp;
        }
        if (valid[r] && sub == 0) store_result(P, out, nidx[r], logl[r]);
    }
}

// second pass of a split launch: partial sums in split order, then the per-vector terms
__global__ void __launch_bounds__(kThreads)
stn_reduce_splits_kernel(const ProbDev P, const double* __restrict__ theta, const int64_t N, const int64_t ld,
                         const int layout, const int K, const double* __restrict__ partial,
                         double* __restrict__ out) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    double logl = partial[n];
    for (int k = 1; k < K; ++k) logl += partial[(int64_t)k * N + n];
    if (!P.out_chisq && (P.n_a1dot > 0 || P.n_sini > 0)) logl += constraint_terms(P, theta, n, ld, layout);
    if (!P.out_chisq && P.n_prior > 0) {
        const double lp = prior_terms(P, theta, n, ld, layout);
        logl = (lp == -CUDART_INF) ? lp : logl + lp;
    }
    store_result(P, out, n, logl);
}

// ---------------------------------------------------------------------------------
// STRICT form: the reference's operation order, one IEEE rounding per operation
// (the __d*_rn intrinsics are never contracted into FMAs).
// ---------------------------------------------------------------------------------
struct StrictTrig {
    double sd, cd, sa, ca, sO, cO, ci;
    bool px_on, rf_on;
};

__device__ __forceinline__ StrictTrig strict_trig(const ProbDev& P, const SrcDev& S, const Params& p) {
    StrictTrig t;
    t.sd = sin(p.dec);
    t.cd = cos(p.dec);
    t.sa = sin(p.ra);
    t.ca = cos(p.ra);
    t.px_on = (p.px != STN_SENTINEL);
    t.rf_on = t.px_on && S.binary && (p.incl != STN_SENTINEL) && (p.om_asc != STN_SENTINEL);
    t.sO = t.cO = t.ci = 0.0;
    if (t.rf_on) {
        const double Om = __dmul_rn(p.om_asc, P.deg2rad);
        t.sO = sin(Om);
        t.cO = cos(Om);
        t.ci = cos(p.incl);
    }
    return t;
}

// offsets in mas for one epoch (positions.py:108-118, 308-309; reflex_motion.py:201-214)
__device__ __forceinline__ void strict_offsets(const SrcDev& S, const Params& p, const StrictTrig& t,
                                               const double* __restrict__ row, double& ora,
                                               double& odec) {
    const double dT = row[0], X = row[1], Y = row[2], Z = row[3];
    ora = __ddiv_rn(__dmul_rn(dT, p.mu_a), t.cd);
    odec = __dmul_rn(dT, p.mu_d);
    if (t.px_on) {
        const double a = __dsub_rn(__dmul_rn(X, t.sa), __dmul_rn(Y, t.ca));
        const double pra = __ddiv_rn(__dmul_rn(p.px, a), t.cd);
        double b = __dadd_rn(__dmul_rn(__dmul_rn(X, t.ca), t.sd), __dmul_rn(__dmul_rn(Y, t.sa), t.sd));
        b = __dsub_rn(b, __dmul_rn(Z, t.cd));
        const double pdec = __dmul_rn(p.px, b);
        ora = __dadd_rn(ora, pra);
        odec = __dadd_rn(odec, pdec);
        if (t.rf_on) {
            const double offset = __dmul_rn(row[12], p.px);
            const double m0 = __dmul_rn(offset, row[13]);
            const double m1 = __dmul_rn(offset, row[14]);
            const double b0 = __dadd_rn(__dmul_rn(t.sO, m0), __dmul_rn(__dmul_rn(t.cO, t.ci), m1));
            const double b1 = __dadd_rn(__dmul_rn(t.cO, m0), __dmul_rn(-__dmul_rn(t.sO, t.ci), m1));
            ora = __dadd_rn(ora, __ddiv_rn(b0, S.cos_decj));
            odec = __dadd_rn(odec, b1);
        }
    }
}

__global__ void __launch_bounds__(kThreads)
stn_loglike_strict_kernel(const ProbDev P, const double* __restrict__ theta, const int64_t N,
                          const int64_t ld, const int layout, double* __restrict__ out) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    double logl = 0.0;
    for (int s = 0; s < P.n_sources; ++s) {
        const SrcDev& S = P.src[s];
        const Params p = load_params(S, theta, n, ld, layout);
        const StrictTrig t = strict_trig(P, S, p);
        const bool ef_on = S.efac_on && (p.efac != STN_SENTINEL);
        const double efad = (p.efad != STN_SENTINEL) ? p.efad : 1.0;
        double chi = 0.0, slog = 0.0;
        // RA half then Dec half, as in the concatenated [RA..., Dec...] vectors
        for (int half = 0; half < 2; ++half) {
            for (int e = 0; e < S.n_epochs; ++e) {
                const double* row = S.strict_rows + (int64_t)e * STN_STRICT_ROW;
                double ora, odec;
                strict_offsets(S, p, t, row, ora, odec);
                const double ref = half == 0 ? p.ra : p.dec;
                const double off = half == 0 ? ora : odec;
                const double model = __dadd_rn(ref, __dmul_rn(off, P.mas2rad));
                const double res = __dsub_rn(row[4 + half], model);
                double err;
                if (S.efac_on) {
                    if (ef_on) {
                        const double f = half == 0 ? __dmul_rn(p.efac, 1.0) : __dmul_rn(p.efac, efad);
                        const double q = __dmul_rn(f, row[10 + half]);
                        const double v = __dadd_rn(__dmul_rn(row[8 + half], row[8 + half]), __dmul_rn(q, q));
                        err = __dsqrt_rn(v);
                    } else {
                        // efac sampled as exactly -999: errs_new = (errs**2)**0.5 = errs   (simulate.py:318-320)
                        err = row[6 + half];
                    }
                    slog = __dadd_rn(slog, log(err));
                } else {
                    err = row[6 + half];   // sqrt(errs^2) == errs
                }
                const double z = __ddiv_rn(res, err);
                chi = __dadd_rn(chi, __dmul_rn(z, z));
            }
        }
        if (P.out_chisq) {
            logl = __dadd_rn(logl, chi);               // chi_sq += np.sum((res/errs_new)**2)  simulate.py:297
        } else {
            logl = __dadd_rn(logl, __dmul_rn(-0.5, chi));
            logl = __dadd_rn(logl, S.efac_on ? -slog : S.neg_sum_log_err);
        }
    }
    if (!P.out_chisq && (P.n_a1dot > 0 || P.n_sini > 0)) logl += constraint_terms(P, theta, n, ld, layout);
    if (!P.out_chisq && P.n_prior > 0) {
        const double lp = prior_terms(P, theta, n, ld, layout);
        logl = (lp == -CUDART_INF) ? lp : logl + lp;
    }
    store_result(P, out, n, logl);
}

// ---------------------------------------------------------------------------------
// K2: model positions of one source (positions.py:13-30), one thread per (sample, epoch)
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
stn_model_kernel(const ProbDev P, const int source, const double* __restrict__ theta, const int64_t N,
                 const int64_t ld, const int layout, const int mode, double* __restrict__ radec,
                 double* __restrict__ off_mas) {
    const SrcDev& S = P.src[source];
    const int E = S.n_epochs;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * E) return;
    const int64_t n = idx / E;
    const int e = (int)(idx - n * E);
    STN_ASSERT(n >= 0 && n < N && e >= 0 && e < E);
    const Params p = load_params(S, theta, n, ld, layout);
    double m_ra, m_dec, o_ra, o_dec;
    if (mode == STN_MODE_STRICT) {
        const StrictTrig t = strict_trig(P, S, p);
        strict_offsets(S, p, t, S.strict_rows + (int64_t)e * STN_STRICT_ROW, o_ra, o_dec);
        m_ra = __dadd_rn(p.ra, __dmul_rn(o_ra, P.mas2rad));
        m_dec = __dadd_rn(p.dec, __dmul_rn(o_dec, P.mas2rad));
    } else {
        const Coef c = make_coef(P, S, p);
        const double2* row = reinterpret_cast<const double2*>(S.fast_rows + (int64_t)e * STN_FAST_ROW);
        double ok_ra, ok_dec;
        fast_offsets<true>(c, row[0], row[1], row[5], ok_ra, ok_dec);
        m_ra = c.ra + ok_ra;
        m_dec = c.dec + ok_dec;
        o_ra = ok_ra / P.mas2rad;
        o_dec = ok_dec / P.mas2rad;
    }
    double* rd = radec + n * (2 * (int64_t)E);
    rd[e] = m_ra;
    rd[E + e] = m_dec;
    if (off_mas != nullptr) {
        double* om = off_mas + n * (2 * (int64_t)E);
        om[e] = o_ra;
        om[E + e] = o_dec;
    }
}

// ---------------------------------------------------------------------------------
// K4: FP64 pipe peak: 8 independent DFMA chains per thread, no memory traffic.  Every chain has
// its OWN multiplier and addend registers: with operands shared between the chains the stream
// stops at ~34.2 TFLOP/s on register-bank conflicts (tools/dfma_probe.cu), with private operands
// it reaches 37.1 of the nominal 37.2.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stn_dfma_peak_kernel(double* sink, int iters, double a, double b) {
    double x[8], m[8], c[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        x[k] = threadIdx.x * 1e-9 + k;
        m[k] = a + k * 1e-12;
        c[k] = b + k * 1e-13;
    }
#pragma unroll 16
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = fma(x[k], m[k], c[k]);
    }
    const double s = ((x[0] + x[1]) + (x[2] + x[3])) + ((x[4] + x[5]) + (x[6] + x[7]));
    if (s == 123.456) sink[0] = s;   // never true in practice; keeps the chains alive
}

// ---------------------------------------------------------------------------------
// Stretch-move sampler steps (sampler adaptor)
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// uniform in (0, 1) from 53 random bits
__device__ __forceinline__ double u01(uint32_t hi, uint32_t lo) {
    const uint64_t k = (((uint64_t)hi << 32) | lo) >> 11;
    return ((double)k + 0.5) * (1.0 / 9007199254740992.0);
}

__global__ void __launch_bounds__(kThreads)
stn_stretch_propose_kernel(const double* __restrict__ x, const int64_t W, const int P, const int64_t ld,
                           const int half, const double a, const uint64_t seed, const uint64_t step,
                           double* __restrict__ y, double* __restrict__ z_out) {
    const int64_t H = W / 2;
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= H) return;
    uint32_t r[4];
    philox4x32_10((uint32_t)k, (uint32_t)(k >> 32), (uint32_t)step, (uint32_t)(step >> 32) ^ (half ? 0x80000000u : 0u),
                  (uint32_t)seed, (uint32_t)(seed >> 32), r);
    const double u = u01(r[0], r[1]);
    const double t = (a - 1.0) * u + 1.0;
    const double z = t * t / a;                                       // g(z) ~ 1/sqrt(z) on [1/a, a]
    const int64_t j = (int64_t)(u01(r[2], r[3]) * (double)H);         // partner in the complementary half
    const double* xi = x + (half ? H + k : k) * ld;
    const double* xj = x + (half ? j : H + j) * ld;
    for (int p = 0; p < P; ++p) y[k * P + p] = fma(z, xi[p] - xj[p], xj[p]);
    z_out[k] = z;
}

__global__ void __launch_bounds__(kThreads)
stn_stretch_accept_kernel(double* __restrict__ x, double* __restrict__ lp, const int64_t W, const int P,
                          const int64_t ld, const int half, const uint64_t seed, const uint64_t step,
                          const double* __restrict__ y, const double* __restrict__ z, const double* __restrict__ lpy,
                          unsigned long long* __restrict__ n_accept) {
    const int64_t H = W / 2;
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool acc = false;
    if (k < H) {
        uint32_t r[4];
        philox4x32_10((uint32_t)k, (uint32_t)(k >> 32), (uint32_t)step,
                      (uint32_t)(step >> 32) ^ (half ? 0xC0000000u : 0x40000000u), (uint32_t)seed,
                      (uint32_t)(seed >> 32), r);
        const int64_t i = half ? H + k : k;
        const double lnq = (double)(P - 1) * log(z[k]) + lpy[k] - lp[i];
        acc = (lpy[k] > -CUDART_INF) && (lpy[k] == lpy[k]) && (log(u01(r[0], r[1])) < lnq);
        if (acc) {
            for (int p = 0; p < P; ++p) x[i * ld + p] = y[k * P + p];
            lp[i] = lpy[k];
        }
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, acc);
    if ((threadIdx.x & 31) == 0 && ballot) atomicAdd(n_accept, (unsigned long long)__popc(ballot));
}

// Walker-sharded forms of the two stretch kernels (several GPUs, one ensemble): this device proposes for
// the proposals [k0, k0 + cnt) of the half-step from ITS copy of the walkers, and the accept kernel
// stores every accepted walker into the copies of ALL devices (peer-mapped pointers: the exchange of
// the updated half over NVLink is the kernel's own stores).  Philox counters are global walker
// indices, so the chain does not depend on how many devices share the ensemble.
struct WalkerCopies {
    double* x[STN_MAX_OUTS];
    double* lp[STN_MAX_OUTS];
    int n;
};

__global__ void __launch_bounds__(kThreads)
stn_stretch_propose_part_kernel(const double* __restrict__ x, const int64_t W, const int P, const int64_t ld,
                                const int half, const double a, const uint64_t seed, const uint64_t step,
                                const int64_t k0, const int64_t cnt, double* __restrict__ y, double* __restrict__ z_out) {
    const int64_t H = W / 2;
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= cnt) return;
    const int64_t k = k0 + q;
    uint32_t r[4];
    philox4x32_10((uint32_t)k, (uint32_t)(k >> 32), (uint32_t)step, (uint32_t)(step >> 32) ^ (half ? 0x80000000u : 0u),
                  (uint32_t)seed, (uint32_t)(seed >> 32), r);
    const double u = u01(r[0], r[1]);
    const double t = (a - 1.0) * u + 1.0;
    const double z = t * t / a;
    const int64_t j = (int64_t)(u01(r[2], r[3]) * (double)H);
    const double* xi = x + (half ? H + k : k) * ld;
    const double* xj = x + (half ? j : H + j) * ld;
    for (int p = 0; p < P; ++p) y[q * P + p] = fma(z, xi[p] - xj[p], xj[p]);
    z_out[q] = z;
}

__global__ void __launch_bounds__(kThreads)
stn_stretch_accept_part_kernel(const WalkerCopies C, const int64_t W, const int P, const int64_t ld, const int half,
                               const uint64_t seed, const uint64_t step, const int64_t k0, const int64_t cnt,
                               const double* __restrict__ y, const double* __restrict__ z,
                               const double* __restrict__ lpy, unsigned long long* __restrict__ n_accept) {
    const int64_t H = W / 2;
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool acc = false;
    if (q < cnt) {
        const int64_t k = k0 + q;
        uint32_t r[4];
        philox4x32_10((uint32_t)k, (uint32_t)(k >> 32), (uint32_t)step,
                      (uint32_t)(step >> 32) ^ (half ? 0xC0000000u : 0x40000000u), (uint32_t)seed,
                      (uint32_t)(seed >> 32), r);
        const int64_t i = half ? H + k : k;
        const double lnq = (double)(P - 1) * log(z[q]) + lpy[q] - C.lp[0][i];
        acc = (lpy[q] > -CUDART_INF) && (lpy[q] == lpy[q]) && (log(u01(r[0], r[1])) < lnq);
        if (acc) {
            for (int c = 0; c < C.n; ++c) {
                for (int p = 0; p < P; ++p) C.x[c][i * ld + p] = y[q * P + p];
                C.lp[c][i] = lpy[q];
            }
        }
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, acc);
    if ((threadIdx.x & 31) == 0 && ballot) atomicAdd(n_accept, (unsigned long long)__popc(ballot));
}

// After a half-step, this device's slice of the updated half -- rows [row0, row0 + cnt) of x and of lp, a
// contiguous block -- goes to every other device's copy with coalesced 16-byte stores over NVLink
// (peer-mapped pointers; scattered 8-byte peer stores from the accept kernel itself were 3x slower).
__global__ void __launch_bounds__(kThreads)
stn_broadcast_rows_kernel(const WalkerCopies C, const int64_t row0, const int64_t cnt, const int64_t ld) {
    const int64_t nx = cnt * ld, stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const double* sx = C.x[0] + row0 * ld;
    const double* sl = C.lp[0] + row0;
    if ((((uintptr_t)sx) & 15) == 0 && (nx & 1) == 0) {
        const double2* s2 = reinterpret_cast<const double2*>(sx);
        for (int64_t e = t0; e < nx / 2; e += stride) {
            const double2 v = s2[e];
            for (int c = 1; c < C.n; ++c) reinterpret_cast<double2*>(C.x[c] + row0 * ld)[e] = v;
        }
    } else {
        for (int64_t e = t0; e < nx; e += stride) {
            const double v = sx[e];
            for (int c = 1; c < C.n; ++c) C.x[c][row0 * ld + e] = v;
        }
    }
    for (int64_t e = t0; e < cnt; e += stride) {
        const double v = sl[e];
        for (int c = 1; c < C.n; ++c) C.lp[c][row0 + e] = v;
    }
}

// ---------------------------------------------------------------------------------
// Fused ensemble sampler: the WHOLE stretch-move loop in one persistent cooperative kernel.
// Per half-step every thread proposes for its walkers, evaluates log prior + log L of the
// proposal inline (the same per-source building blocks as K1, rows read straight from the
// L1/L2-resident epoch tables: these are the small real-data problems, ~10 epochs per source,
// where six launches per step cost more than the arithmetic) and accepts / rejects; the two
// half-steps of a step are separated by one grid-wide barrier.  Same Philox counters and the
// same arithmetic (plan lanes = 1, splits = 1) as the three-launch path, hence the same chain.
// ---------------------------------------------------------------------------------
#define STN_SAMPLER_MAX_P 64

template <bool COH>
__device__ double logpost_one(const ProbDev& P, const double* theta_row, int64_t ld) {
    double logl[1] = {0.0};
    const int64_t nidx[1] = {0};
    SrcState<1> st;
    st.mask = kChunk - 1;
    st.nvar = 0;
    for (int c = 0; c < P.n_chunks; ++c) {
        const ChunkDev ck = P.chunks[c];
        const SrcDev& S = P.src[ck.source];
        if (ck.flags & kFirst) src_begin<1, COH>(P, S, theta_row, nidx, ld, STN_LAYOUT_AOS, st);
        src_chunk<1, 1>(S, ck.rows, ck.n_epochs, 0, st);
        if (ck.flags & kLast) src_fold<1, 1>(P, S, true, st, logl);
    }
    double v = logl[0];
    if (P.n_a1dot > 0 || P.n_sini > 0) v += constraint_terms<COH>(P, theta_row, 0, ld, STN_LAYOUT_AOS);
    if (P.n_prior > 0) {
        const double lp = prior_terms<COH>(P, theta_row, 0, ld, STN_LAYOUT_AOS);
        v = (lp == -CUDART_INF) ? lp : v + lp;
    }
    return v;
}

// Bytes of shared memory sampler_stage_problem needs (host and device agree through this one function).
__host__ __device__ inline size_t sampler_align(size_t v) { return (v + 15) & ~(size_t)15; }

__device__ __forceinline__ void block_copy(void* dst, const void* src, size_t bytes) {
    // bytes is a multiple of 4 for every array copied here
    uint32_t* d = static_cast<uint32_t*>(dst);
    const uint32_t* q = static_cast<const uint32_t*>(src);
    for (size_t i = threadIdx.x; i < bytes / 4; i += blockDim.x) d[i] = q[i];
}

// Copies everything ProbDev points to into shared memory (every CTA its own copy) and returns a ProbDev
// whose pointers refer to the copies.  Layout: SrcDev[S] | ChunkDev[C] | fast tables | constraint and prior arrays.
__device__ ProbDev sampler_stage_problem(const ProbDev& P, unsigned char* smem) {
    ProbDev Q = P;
    size_t off = 0;
    SrcDev* src = reinterpret_cast<SrcDev*>(smem + off);
    off += sampler_align(sizeof(SrcDev) * P.n_sources);
    ChunkDev* chunks = reinterpret_cast<ChunkDev*>(smem + off);
    off += sampler_align(sizeof(ChunkDev) * P.n_chunks);
    block_copy(src, P.src, sizeof(SrcDev) * P.n_sources);
    block_copy(chunks, P.chunks, sizeof(ChunkDev) * P.n_chunks);
    __syncthreads();
    // tables: one after the other in source order
    size_t table_off = off;
    for (int s = 0; s < P.n_sources; ++s) {
        const size_t bytes = sizeof(double) * STN_FAST_ROW * (size_t)src[s].n_epochs;
        block_copy(smem + table_off, P.src[s].fast_rows, bytes);
        table_off += sampler_align(bytes);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        size_t o = off;
        for (int s = 0; s < P.n_sources; ++s) {
            const double* old_rows = src[s].fast_rows;
            const double* new_rows = reinterpret_cast<const double*>(smem + o);
            for (int c = 0; c < P.n_chunks; ++c)
                if (chunks[c].source == s) chunks[c].rows = new_rows + (chunks[c].rows - old_rows);
            src[s].fast_rows = new_rows;
            o += sampler_align(sizeof(double) * STN_FAST_ROW * (size_t)src[s].n_epochs);
        }
    }
    off = table_off;
    auto take = [&](const void* g, size_t bytes) -> unsigned char* {
        unsigned char* d = smem + off;
        block_copy(d, g, bytes);
        off += sampler_align(bytes);
        return d;
    };
    Q.src = src;
    Q.chunks = chunks;
    Q.a1_src = reinterpret_cast<const int*>(take(P.a1_src, sizeof(int) * P.n_a1dot));
    Q.a1_mu = reinterpret_cast<const double*>(take(P.a1_mu, sizeof(double) * P.n_a1dot));
    Q.a1_sigma = reinterpret_cast<const double*>(take(P.a1_sigma, sizeof(double) * P.n_a1dot));
    Q.sini_col = reinterpret_cast<const int*>(take(P.sini_col, sizeof(int) * P.n_sini));
    Q.sini_mu = reinterpret_cast<const double*>(take(P.sini_mu, sizeof(double) * P.n_sini));
    Q.sini_sigma = reinterpret_cast<const double*>(take(P.sini_sigma, sizeof(double) * P.n_sini));
    Q.prior_kind = reinterpret_cast<const int*>(take(P.prior_kind, sizeof(int) * P.n_prior));
    Q.prior_a = reinterpret_cast<const double*>(take(P.prior_a, sizeof(double) * P.n_prior));
    Q.prior_b = reinterpret_cast<const double*>(take(P.prior_b, sizeof(double) * P.n_prior));
    Q.sinlim_col = reinterpret_cast<const int*>(take(P.sinlim_col, sizeof(int) * P.n_sinlim));
    Q.sinlim_lo = reinterpret_cast<const double*>(take(P.sinlim_lo, sizeof(double) * P.n_sinlim));
    Q.sinlim_hi = reinterpret_cast<const double*>(take(P.sinlim_hi, sizeof(double) * P.n_sinlim));
    __syncthreads();
    return Q;
}

constexpr int kSamplerThreads = 256;

// All CTAs of the launch meet here.  cluster_mode: the launch is ONE thread-block cluster (up to 16 CTAs
// = 4096 proposals per half-step) and this is the hardware cluster barrier (release / acquire: the walkers
// just written by other CTAs are visible afterwards); otherwise a cooperative launch and the grid barrier.
__device__ __forceinline__ void sampler_barrier(bool cluster_mode) {
    namespace cg = cooperative_groups;
    if (cluster_mode) {
        asm volatile("barrier.cluster.arrive.release;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
    } else {
        cg::this_grid().sync();
    }
}

__global__ void __launch_bounds__(kSamplerThreads)
stn_ensemble_kernel(const ProbDev P0, double* x, double* lp, const int64_t W, const int64_t ld, const double a,
                    const uint64_t seed, const uint64_t step0, const int64_t n_steps, const int64_t thin,
                    double* chain, double* lnprob, unsigned long long* n_accept, const int cluster_mode) {
    // The whole problem description -- sources, chunk list, epoch tables, constraints, prior -- is copied
    // to shared memory once: the barrier between half-steps invalidates L1 (acquire), so anything left in
    // global memory would be re-fetched from L2 (~700 cycles per dependent load) in every half-step.
    extern __shared__ __align__(16) unsigned char sampler_smem[];
    const ProbDev Q = sampler_stage_problem(P0, sampler_smem);
    const ProbDev& P = Q;
    const int np = P.n_params;
    const int64_t H = W / 2;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t Hpad = (H + 31) / 32 * 32;             // whole warps walk the loop together
    unsigned long long accepted = 0;
    double y[STN_SAMPLER_MAX_P];
    for (int64_t t = 0; t < n_steps; ++t) {
        const uint64_t step = step0 + (uint64_t)t;
        for (int half = 0; half < 2; ++half) {
            for (int64_t kk = gtid; kk < Hpad; kk += nthreads) {
                const bool valid = kk < H;
                const int64_t k = valid ? kk : H - 1;
                // this walker's row and log-posterior do not depend on the random numbers: their L2 round trip
                // (the barrier invalidated L1) overlaps the Philox rounds
                const int64_t i = half ? H + k : k;
                const double* xi = x + i * ld;
                for (int p = 0; p < np; ++p) y[p] = __ldcg(xi + p);
                const double lp_i = __ldcg(lp + i);
                uint32_t r[4];
                philox4x32_10((uint32_t)k, (uint32_t)(k >> 32), (uint32_t)step,
                              (uint32_t)(step >> 32) ^ (half ? 0x80000000u : 0u), (uint32_t)seed, (uint32_t)(seed >> 32), r);
                const double u = u01(r[0], r[1]);
                const double tt = (a - 1.0) * u + 1.0;
                const double z = tt * tt / a;
                const int64_t j = (int64_t)(u01(r[2], r[3]) * (double)H);
                const double* xj = x + (half ? j : H + j) * ld;
                uint32_t r2[4];
                philox4x32_10((uint32_t)k, (uint32_t)(k >> 32), (uint32_t)step,
                              (uint32_t)(step >> 32) ^ (half ? 0xC0000000u : 0x40000000u), (uint32_t)seed,
                              (uint32_t)(seed >> 32), r2);                        // accept draw, while the partner row arrives
                for (int p = 0; p < np; ++p) {
                    const double vj = __ldcg(xj + p);
                    y[p] = fma(z, y[p] - vj, vj);
                }
                const double lpy = logpost_one<true>(P, y, np);
                const double lnq = (double)(np - 1) * log(z) + lpy - lp_i;
                const bool acc = valid && (lpy > -CUDART_INF) && (lpy == lpy) && (log(u01(r2[0], r2[1])) < lnq);
                if (acc) {
                    for (int p = 0; p < np; ++p) x[i * ld + p] = y[p];
                    lp[i] = lpy;
                    ++accepted;
                }
            }
            sampler_barrier(cluster_mode != 0);        // the other half reads these walkers next
        }
        if (chain != nullptr && (t + 1) % thin == 0) {
            const int64_t slot = (t + 1) / thin - 1;
            for (int64_t e = gtid; e < W * np; e += nthreads) {
                const int64_t w = e / np;
                chain[slot * W * np + e] = __ldcg(x + w * ld + (e - w * np));
            }
            if (lnprob != nullptr)
                for (int64_t w = gtid; w < W; w += nthreads) lnprob[slot * W + w] = __ldcg(lp + w);
            sampler_barrier(cluster_mode != 0);        // half-step 0 of the next step overwrites rows that are being copied
        }
    }
    for (int o = 16; o > 0; o >>= 1) accepted += __shfl_xor_sync(0xffffffffu, accepted, o);
    if ((threadIdx.x & 31) == 0 && accepted) atomicAdd(n_accept, accepted);
}

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define STN_CUDA(call)                                                                        \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(STN_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                  \
    } while (0)

}  // namespace

struct StnCtx {
    int device = 0;
    int sm_count = 0;
    ProbDev prob{};
    std::vector<void*> dev_allocs;
    std::vector<int> n_epochs;          // per source
    int max_epochs = 0;
    ProbDev prior_prob{};               // copy of prob with the prior switched on (stn_set_prior)
    bool has_prior = false;
    std::vector<void*> prior_allocs;
    bool efac_heavy = false;            // at least half of the epochs belong to EFAC sources
    // zero-copy staging for small host batches (pinned, mapped: the kernel reads / writes it directly)
    double* small_in = nullptr;
    double* small_out = nullptr;
    static constexpr int kSlots = 3;    // host pipeline depth: H2D of chunk i+1, kernel of chunk i, D2H of chunk i-1
    double* scratch[1 + kSlots] = {};   // split partial sums: [0] device API, [1 + slot] host pipeline
    int64_t scratch_elems[1 + kSlots] = {};
    int64_t launches = 0;
    // *_host pipeline: one stream, one device staging pair and (for pageable caller buffers) one pinned
    // staging pair per slot
    cudaStream_t streams[kSlots] = {};
    cudaEvent_t h2d_done[kSlots] = {};
    cudaEvent_t d2h_done[kSlots] = {};
    double* stage_theta[kSlots] = {};
    double* stage_out[kSlots] = {};
    int64_t stage_rows = 0;
    double* pin_theta[kSlots] = {};
    double* pin_out[kSlots] = {};
    int64_t pin_rows = 0;
    int host_share = 1;                 // devices DMA-ing from this host at the same time (stn_host_hint)
    cudaStream_t timing_stream = nullptr;
    double* gather_buf = nullptr;       // stn_multi_loglike_dev: this device's slice before the peer copy
    int64_t gather_elems = 0;
    double* sampler_buf = nullptr;      // stn_ensemble_run (three-launch engine): proposals, stretch factors, log-posteriors
    int64_t sampler_elems = 0;
    cudaEvent_t sampler_ev[2] = {};     // stn_multi_ensemble_run: "this device finished half-step h" (ping-pong)
    unsigned long long* sampler_nacc = nullptr;
};

// Several contexts of the same problem, one per device, driven as one (stn_multi_*).
struct StnMulti {
    std::vector<StnCtx*> ctx;
    std::vector<int> peer_ok;           // [a * G + b]: kernels on ctx[a]'s device may store to ctx[b]'s device
};

namespace {

template <typename T>
int upload(StnCtx* ctx, const T* host, size_t count, const T** dev_out) {
    void* d = nullptr;
    const size_t bytes = (count ? count : 1) * sizeof(T);
    STN_CUDA(cudaMalloc(&d, bytes));
    ctx->dev_allocs.push_back(d);
    if (count) STN_CUDA(cudaMemcpy(d, host, count * sizeof(T), cudaMemcpyHostToDevice));
    *dev_out = static_cast<const T*>(d);
    return STN_OK;
}

#ifdef STN_DEBUG
// the `bytes` starting at device pointer p must lie inside one allocation known to the driver
int dbg_range(const void* p, size_t bytes, const char* what) {
    if (bytes == 0) return STN_OK;
    CUdeviceptr base = 0;
    size_t size = 0;
    const CUresult r = cuMemGetAddressRange(&base, &size, (CUdeviceptr)(uintptr_t)p);
    if (r != CUDA_SUCCESS) return fail(STN_ERR_INVALID, "STN_DEBUG: %s (%p) is not inside a CUDA allocation", what, p);
    const uintptr_t lo = (uintptr_t)p, hi = lo + bytes;
    if (lo < (uintptr_t)base || hi > (uintptr_t)base + size)
        return fail(STN_ERR_INVALID, "STN_DEBUG: %s needs %zu bytes at %p but its allocation is [%p, +%zu)", what, bytes,
                    p, (void*)(uintptr_t)base, size);
    return STN_OK;
}
size_t dbg_theta_bytes(int64_t n, int64_t ld, int layout, int P) {
    if (n == 0) return 0;
    return sizeof(double) * (size_t)(layout == STN_LAYOUT_AOS ? (n - 1) * ld + P : (int64_t)(P - 1) * ld + n);
}
#define STN_DBG_RANGE(p, bytes, what) do { const int r_ = dbg_range(p, bytes, what); if (r_) return r_; } while (0)
#else
#define STN_DBG_RANGE(p, bytes, what) ((void)0)
#endif

// ---- execution plan -------------------------------------------------------------------
// plan = lanes | (splits << 8).  lanes: lanes of a warp sharing one sample (epochs of a chunk
// interleaved over them); splits: CTAs sharing one tile of samples (the chunk list cut into
// contiguous ranges).  Both change the summation order, i.e. result bits, which is why the
// plan is explicit in the ABI; samples per thread (R) never does.
inline int plan_lanes(int plan) { return plan & 0xff; }
inline int plan_splits(int plan) { return (plan >> 8) & 0xffff; }
inline int make_plan(int lanes, int splits) { return lanes | (splits << 8); }

int big_R(const StnCtx* ctx) { return ctx->efac_heavy ? STN_RBIG_EFAC : STN_RBIG; }

// R used for a batch of n with `splits` CTAs per tile: the big R once its tiles fill the SMs
int choose_R(const StnCtx* ctx, int64_t n, int lanes, int splits) {
    if (lanes != 1) return 1;
    const int R = big_R(ctx);
    const int64_t slots = (int64_t)ctx->sm_count * minblocks_for(R);
    const int64_t tile = (int64_t)threads_for(R) * R;
    const int64_t tiles = (n + tile - 1) / tile;
    return tiles * splits >= 2 * slots ? R : 1;
}

int pick_plan(const StnCtx* ctx, int64_t n) {
    if (n <= 0) return make_plan(1, 1);
    const int T = ctx->prob.n_chunks;
    const int R = big_R(ctx);
    const int64_t slots_big = (int64_t)ctx->sm_count * minblocks_for(R);
    const int64_t tile_big = (int64_t)threads_for(R) * R;
    const int64_t tiles_big = (n + tile_big - 1) / tile_big;
    if (tiles_big >= 32 * slots_big) return make_plan(1, 1);         // >= 32 waves: the last one is < 3 %
    // choose the number of splits that wastes the least of the last wave, given the tile count
    auto eff = [&](int64_t tiles, int64_t slots, int K) {
        const int64_t ctas = tiles * K;
        const int64_t waves = (ctas + slots - 1) / slots;
        const double fill = (double)ctas / (double)(waves * slots);
        const int per = (T + K - 1) / K;                              // chunks of the longest split
        const double balance = (double)T / (double)((int64_t)per * K);
        return fill * balance;
    };
    if (tiles_big * (int64_t)T >= 2 * slots_big) {                    // the big-R kernel can fill the GPU
        int bestK = 1;
        double best = eff(tiles_big, slots_big, 1);
        for (int K = 2; K <= T && K <= 64; ++K) {
            if (tiles_big * K < 2 * slots_big && K < T) continue;
            const double e = eff(tiles_big, slots_big, K);
            if (e > best * 1.02) { best = e; bestK = K; }
            if (tiles_big * K >= 32 * slots_big) break;
        }
        if (tiles_big * bestK >= 2 * slots_big) return make_plan(1, bestK);
    }
    // small batches: one sample per thread, as many splits as there are chunks allow, then lanes
    const int64_t slots = (int64_t)ctx->sm_count * 2;
    const int64_t tiles = (n + kThreads - 1) / kThreads;
    int K = 1;
    while (K < T && K < 64 && tiles * K < 2 * slots) ++K;
    int cap = 1;                                                       // lanes never outnumber epochs/8
    while (cap < 32 && cap * 8 <= ctx->max_epochs && cap * 8 <= kChunk) cap *= 2;
    int L = 1;
    while (L < cap && ((n * L + kThreads - 1) / kThreads) * K < 2 * slots) L *= 2;
    return make_plan(L, K);
}

template <int R, int L, int TPB = threads_for(R)>
int launch_fast(StnCtx* ctx, const ProbDev& prob, const double* theta, int64_t n, int64_t ld, int layout,
                double* out, cudaStream_t st, int splits, double* partial) {
    constexpr int per_cta = (TPB / L) * R;
    const int64_t grid = (n + per_cta - 1) / per_cta;
    if (grid > 0x7fffffffLL) return fail(STN_ERR_INVALID, "batch too large for one launch");
    dim3 g((unsigned)grid, (unsigned)splits, 1);
    (void)cudaGetLastError();      // other CUDA users of this process (torch) may have left a stale error
#if STN_DYNAMIC_STAGE
    static bool attr_set = false;                    // per instantiation
    if (!attr_set) {
        STN_CUDA(cudaFuncSetAttribute(stn_loglike_fast_kernel<R, L, TPB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)kStageSmem));
        attr_set = true;
    }
#endif
    stn_loglike_fast_kernel<R, L, TPB><<<g, TPB, kStageSmem, st>>>(prob, theta, n, ld, layout, out, partial);
    STN_CUDA(cudaGetLastError());
    ctx->launches++;
    if (splits > 1) {
        const int64_t rg = (n + kThreads - 1) / kThreads;
        stn_reduce_splits_kernel<<<(unsigned)rg, kThreads, 0, st>>>(prob, theta, n, ld, layout, splits, partial, out);
        STN_CUDA(cudaGetLastError());
        ctx->launches++;
    }
    return STN_OK;
}

// scratch_slot: 0 for the device-buffer API, 1 + slot for the streams of the host pipeline
int launch_loglike(StnCtx* ctx, const double* theta, int64_t n, int64_t ld, int layout, int mode,
                   int plan, double* out, cudaStream_t st, int what = 0, int scratch_slot = 0,
                   double* const* extra_out = nullptr, int n_extra_out = 0) {
    // what: 0 = log L, 1 = log prior + log L, 2 = chi-square
    if (n == 0) return STN_OK;
    ProbDev prob = what == 1 ? ctx->prior_prob : ctx->prob;
    prob.out_chisq = what == 2 ? 1 : 0;
    prob.n_extra_out = n_extra_out;
    for (int k = 0; k < n_extra_out; ++k) prob.extra_out[k] = extra_out[k];
    if (mode == STN_MODE_STRICT) {
        STN_DBG_RANGE(theta, dbg_theta_bytes(n, ld, layout, ctx->prob.n_params), "theta");
        STN_DBG_RANGE(out, (size_t)n * sizeof(double), "out");
        const int64_t grid = (n + kThreads - 1) / kThreads;
        (void)cudaGetLastError();
        stn_loglike_strict_kernel<<<(unsigned)grid, kThreads, 0, st>>>(prob, theta, n, ld, layout, out);
        STN_CUDA(cudaGetLastError());
        ctx->launches++;
        return STN_OK;
    }
    if (plan == 0) plan = pick_plan(ctx, n);
    const int lanes = plan_lanes(plan);
    int splits = plan_splits(plan);
    if (splits == 0) splits = 1;                                   // a bare lanes value is a valid plan
    if (splits > ctx->prob.n_chunks && ctx->prob.n_chunks > 0)
        return fail(STN_ERR_INVALID, "plan asks for %d splits but the problem has %d chunks", splits,
                    ctx->prob.n_chunks);
    if (splits > 65535) return fail(STN_ERR_INVALID, "too many splits");
    double* partial = nullptr;
    if (splits > 1) {
        const int64_t need = (int64_t)splits * n;
        if (ctx->scratch_elems[scratch_slot] < need) {
            if (ctx->scratch[scratch_slot]) {
                // slot 0 may have been used from another of the caller's streams before
                if (scratch_slot == 0) STN_CUDA(cudaDeviceSynchronize());
                else STN_CUDA(cudaStreamSynchronize(st));
                STN_CUDA(cudaFree(ctx->scratch[scratch_slot]));
                ctx->scratch[scratch_slot] = nullptr;
                ctx->scratch_elems[scratch_slot] = 0;
            }
            STN_CUDA(cudaMalloc(&ctx->scratch[scratch_slot], (size_t)need * sizeof(double)));
            ctx->scratch_elems[scratch_slot] = need;
        }
        partial = ctx->scratch[scratch_slot];
    }
    STN_DBG_RANGE(theta, dbg_theta_bytes(n, ld, layout, ctx->prob.n_params), "theta");
    STN_DBG_RANGE(out, (size_t)n * sizeof(double), "out");
    if (partial) STN_DBG_RANGE(partial, (size_t)splits * n * sizeof(double), "split scratch");
    const int R = choose_R(ctx, n, lanes, splits);
    // EFAC-heavy problems use the R = STN_RBIG_EFAC kernel for large batches, others R = STN_RBIG
#define STN_GO(RR, LL) return launch_fast<RR, LL>(ctx, prob, theta, n, ld, layout, out, st, splits, partial)
    switch (lanes) {
        case 1:
            if (R > 1) {
                // Wave quantisation: every CTA of the big kernel runs for the same time, so with FEW waves (the
                // CTAs are still in lockstep) a last wave that is only partly full costs a whole wave: 4.6 waves
                // at 2^19 vectors take the time of 5.  The samples beyond the last FULL wave then go to the
                // R = 2 kernel in 64-thread CTAs: two thirds of the work per thread and many small CTAs that
                // spread evenly over the SMs (6.40 -> 6.14 ms at 2^19).  After ~10 waves the CTAs have drifted
                // apart and the partial wave costs little (23.4 ms at 2^21 without, 23.8 with): not used there.
                // R never changes a result bit, so this is invisible in the output.
                const int bigR = ctx->efac_heavy ? STN_RBIG_EFAC : STN_RBIG;
                const int64_t tile = (int64_t)threads_for(bigR) * bigR;
                const int64_t slots = (int64_t)ctx->sm_count * minblocks_for(bigR);
                const int64_t tiles = (n + tile - 1) / tile;
                const int64_t n_main = (tiles / slots) * slots * tile;
                constexpr int kTailR = STN_TAIL_R;
                const int64_t cap1 = (int64_t)ctx->sm_count * minblocks_tpb(kTailR, kTailThreads) * kTailThreads * kTailR;
                const int64_t left = tiles - (tiles / slots) * slots;                 // tiles of the last, partial wave
                const bool worth = tiles / slots <= 8 && 100 * left >= STN_TAIL_MIN_PCT * slots && 100 * left <= 85 * slots;
                if (splits == 1 && n_main > 0 && n_main < n && n - n_main <= cap1 && worth && STN_TAIL_SMALL_R) {
                    int rc = ctx->efac_heavy
                                 ? launch_fast<STN_RBIG_EFAC, 1>(ctx, prob, theta, n_main, ld, layout, out, st, 1, nullptr)
                                 : launch_fast<STN_RBIG, 1>(ctx, prob, theta, n_main, ld, layout, out, st, 1, nullptr);
                    if (rc) return rc;
                    ProbDev tail = prob;
                    for (int k = 0; k < tail.n_extra_out; ++k) tail.extra_out[k] += n_main;
                    const double* th_tail = layout == STN_LAYOUT_AOS ? theta + n_main * ld : theta + n_main;
                    return launch_fast<kTailR, 1, kTailThreads>(ctx, tail, th_tail, n - n_main, ld, layout, out + n_main, st, 1, nullptr);
                }
                if (ctx->efac_heavy) STN_GO(STN_RBIG_EFAC, 1);
                STN_GO(STN_RBIG, 1);
            }
            STN_GO(1, 1);
        case 2: STN_GO(1, 2);
        case 4: STN_GO(1, 4);
        case 8: STN_GO(1, 8);
        case 16: STN_GO(1, 16);
        case 32: STN_GO(1, 32);
        default: return fail(STN_ERR_INVALID, "plan lanes must be a power of two <= 32 (got %d)", lanes);
    }
#undef STN_GO
}

int check_call(StnCtx* ctx, const void* theta, int64_t n, int64_t ld, int layout, int mode,
               const void* out) {
    if (!ctx) return fail(STN_ERR_INVALID, "null context");
    if (n < 0) return fail(STN_ERR_INVALID, "negative batch size");
    if (n > 0 && (!theta || !out)) return fail(STN_ERR_INVALID, "null buffer");
    if (layout != STN_LAYOUT_AOS && layout != STN_LAYOUT_SOA) return fail(STN_ERR_INVALID, "bad layout %d", layout);
    if (mode != STN_MODE_FAST && mode != STN_MODE_STRICT) return fail(STN_ERR_INVALID, "bad mode %d", mode);
    if (layout == STN_LAYOUT_AOS && ld < ctx->prob.n_params)
        return fail(STN_ERR_INVALID, "AoS leading dimension %lld < P = %d", (long long)ld, ctx->prob.n_params);
    if (layout == STN_LAYOUT_SOA && ld < n)
        return fail(STN_ERR_INVALID, "SoA leading dimension %lld < n = %lld", (long long)ld, (long long)n);
    return STN_OK;
}


// ---- host-buffer pipeline ---------------------------------------------------------------
// true when the CUDA runtime knows `p` as pinned (page-locked) host memory, i.e. a DMA target
bool is_pinned(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

int host_pipe_init(StnCtx* ctx) {
    for (int i = 0; i < StnCtx::kSlots; ++i) {
        if (!ctx->streams[i]) STN_CUDA(cudaStreamCreateWithFlags(&ctx->streams[i], cudaStreamNonBlocking));
        if (!ctx->h2d_done[i]) STN_CUDA(cudaEventCreateWithFlags(&ctx->h2d_done[i], cudaEventDisableTiming));
        if (!ctx->d2h_done[i]) STN_CUDA(cudaEventCreateWithFlags(&ctx->d2h_done[i], cudaEventDisableTiming));
    }
    return STN_OK;
}

void host_pipe_drain(StnCtx* ctx) {
    for (int i = 0; i < StnCtx::kSlots; ++i)
        if (ctx->streams[i]) cudaStreamSynchronize(ctx->streams[i]);
}

// Rows of the k-th chunk: a short first chunk (its H2D copy is the only one nothing overlaps),
// growing fourfold up to 2^20 rows (every chunk boundary costs a partly filled wave of CTAs).
#ifndef STN_HOST_CHUNK_LOG2
#define STN_HOST_CHUNK_LOG2 20
#endif
constexpr int64_t kChunkRowsMax = (int64_t)1 << STN_HOST_CHUNK_LOG2;

// memcpy of a large block on a few threads: one core moves ~12 GB/s, which is below what the
// kernel consumes for short-series problems and would make pageable buffers the bottleneck
void big_memcpy(void* dst, const void* src, size_t bytes) {
    constexpr size_t kPerThread = (size_t)8 << 20;
    int nt = (int)(bytes / kPerThread);
    if (nt > 4) nt = 4;
    if (nt <= 1) {
        std::memcpy(dst, src, bytes);
        return;
    }
    const size_t part = ((bytes / nt) + 63) & ~(size_t)63;
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) {
        const size_t lo = (size_t)t * part;
        const size_t len = lo >= bytes ? 0 : (t == nt - 1 ? bytes - lo : part);
        if (len) pool.emplace_back([=]() { std::memcpy((char*)dst + lo, (const char*)src + lo, len); });
    }
    std::memcpy(dst, src, part < bytes ? part : bytes);
    for (std::thread& th : pool) th.join();
}
// Tuning overrides (environment, read once): STN_HOST_FIRST_LOG2 = log2 of the first chunk's rows,
// STN_HOST_GROWTH = factor between consecutive chunks.
struct HostSchedule {
    int64_t first = 0;
    int growth = 0;
    HostSchedule() {
        if (const char* e = std::getenv("STN_HOST_FIRST_LOG2")) first = (int64_t)1 << std::atoi(e);
        if (const char* e = std::getenv("STN_HOST_GROWTH")) growth = std::atoi(e);
    }
};
const HostSchedule& host_schedule() {
    static const HostSchedule s;
    return s;
}

int64_t first_chunk_rows(int64_t n) {
    const HostSchedule& hs = host_schedule();
    if (hs.first > 0) return hs.first < n ? hs.first : n;
    int64_t r = n / 32;
    if (r > ((int64_t)1 << 15)) r = (int64_t)1 << 15;
    if (r < 4096) r = 4096;
    return r < n ? r : n;
}

// Next chunk: fourfold growth when this device has the host's PCIe bandwidth to itself (~50 GB/s: the copy of
// a 4x larger chunk still hides under the kernel of the current one), twofold when several devices copy at
// once (stn_multi_*, or stn_host_hint from a multi-process launcher).  Chunks of a wave or more are cut at
// whole waves of the big kernel so that only the last chunk ends in a partly filled wave.
int64_t next_chunk_rows(const StnCtx* ctx, int64_t rows) {
    int g = host_schedule().growth > 1 ? host_schedule().growth : (ctx->host_share > 1 ? 2 : 4);
    int64_t r = rows >= kChunkRowsMax ? kChunkRowsMax : (rows * g < kChunkRowsMax ? rows * g : kChunkRowsMax);
    const int bigR = ctx->efac_heavy ? STN_RBIG_EFAC : STN_RBIG;
    const int64_t wave = (int64_t)ctx->sm_count * minblocks_for(bigR) * threads_for(bigR) * bigR;
    if (r >= wave) r = (r / wave) * wave;
    return r;
}

// One slice of a host batch on this context's device: H2D, kernel(s), D2H, synchronised on return.
// The caller has made ctx->device current and resolved `plan` from the WHOLE batch.
int run_host(StnCtx* ctx, const double* theta_host, int64_t n, int64_t ld, int layout, int mode, int plan,
             double* out_host) {
    const int P = ctx->prob.n_params;
    int rc = host_pipe_init(ctx);
    if (rc) return rc;
    // Small batches (a sampler evaluating one point, or a few walkers) are latency-bound: skip both
    // cudaMemcpy calls by letting the kernel read the parameters from, and write log L to, pinned
    // mapped host memory.  One launch + one stream synchronise.
    constexpr int64_t kSmallBytes = 64 * 1024, kSmallRows = 2048;
    if (layout == STN_LAYOUT_AOS && n <= kSmallRows && n * P * (int64_t)sizeof(double) <= kSmallBytes) {
        if (!ctx->small_in) STN_CUDA(cudaHostAlloc((void**)&ctx->small_in, kSmallBytes, cudaHostAllocMapped));
        if (!ctx->small_out) STN_CUDA(cudaHostAlloc((void**)&ctx->small_out, kSmallRows * sizeof(double), cudaHostAllocMapped));
        for (int64_t i = 0; i < n; ++i) std::memcpy(ctx->small_in + i * P, theta_host + i * ld, (size_t)P * sizeof(double));
        rc = launch_loglike(ctx, ctx->small_in, n, P, layout, mode, plan, ctx->small_out, ctx->streams[0], 0, 1);
        const cudaError_t e = cudaStreamSynchronize(ctx->streams[0]);
        if (rc) return rc;
        if (e != cudaSuccess) return fail(STN_ERR_CUDA, "small-batch launch failed: %s", cudaGetErrorString(e));
        std::memcpy(out_host, ctx->small_out, (size_t)n * sizeof(double));
        return STN_OK;
    }
    const bool pin_in = is_pinned(theta_host);
    const bool pin_out = is_pinned(out_host);
    const int64_t cap = n < kChunkRowsMax ? n : kChunkRowsMax;
    // device staging (and, for pageable caller buffers, the library's own pinned ring)
    if (ctx->stage_rows < cap) {
        host_pipe_drain(ctx);
        for (int i = 0; i < StnCtx::kSlots; ++i) {
            if (ctx->stage_theta[i]) cudaFree(ctx->stage_theta[i]);
            if (ctx->stage_out[i]) cudaFree(ctx->stage_out[i]);
            ctx->stage_theta[i] = ctx->stage_out[i] = nullptr;
        }
        ctx->stage_rows = 0;
        for (int i = 0; i < StnCtx::kSlots; ++i) {
            STN_CUDA(cudaMalloc(&ctx->stage_theta[i], (size_t)cap * P * sizeof(double)));
            STN_CUDA(cudaMalloc(&ctx->stage_out[i], (size_t)cap * sizeof(double)));
        }
        ctx->stage_rows = cap;
    }
    if ((!pin_in || !pin_out) && ctx->pin_rows < cap) {
        host_pipe_drain(ctx);
        for (int i = 0; i < StnCtx::kSlots; ++i) {
            if (ctx->pin_theta[i]) cudaFreeHost(ctx->pin_theta[i]);
            if (ctx->pin_out[i]) cudaFreeHost(ctx->pin_out[i]);
            ctx->pin_theta[i] = ctx->pin_out[i] = nullptr;
        }
        ctx->pin_rows = 0;
        for (int i = 0; i < StnCtx::kSlots; ++i) {
            STN_CUDA(cudaHostAlloc((void**)&ctx->pin_theta[i], (size_t)cap * P * sizeof(double), cudaHostAllocDefault));
            STN_CUDA(cudaHostAlloc((void**)&ctx->pin_out[i], (size_t)cap * sizeof(double), cudaHostAllocDefault));
        }
        ctx->pin_rows = cap;
    }

    struct Pending { int64_t lo, m; bool live; } pend[StnCtx::kSlots] = {};
    auto body = [&]() -> int {
        int64_t rows = first_chunk_rows(n);
        int k = 0;
        for (int64_t lo = 0; lo < n; ++k) {
            const int64_t m = (n - lo) < rows ? (n - lo) : rows;
            const int slot = k % StnCtx::kSlots;
            cudaStream_t st = ctx->streams[slot];
            double* dth = ctx->stage_theta[slot];
            double* dout = ctx->stage_out[slot];
            // pageable output: the chunk that used this slot last is copied out before the slot is reused
            if (pend[slot].live) {
                STN_CUDA(cudaEventSynchronize(ctx->d2h_done[slot]));
                std::memcpy(out_host + pend[slot].lo, ctx->pin_out[slot], (size_t)pend[slot].m * sizeof(double));
                pend[slot].live = false;
            }
            const double* src;
            int64_t src_ld;                 // leading dimension of the host image that is DMA'd
            if (pin_in) {
                src = layout == STN_LAYOUT_AOS ? theta_host + lo * ld : theta_host + lo;
                src_ld = ld;
            } else {
                // pageable input: compact copy into this slot's pinned buffer (the previous DMA out of it is done
                // once h2d_done has fired), then DMA from there while the next chunk is being copied
                if (k >= StnCtx::kSlots) STN_CUDA(cudaEventSynchronize(ctx->h2d_done[slot]));
                double* hp = ctx->pin_theta[slot];
                if (layout == STN_LAYOUT_AOS) {
                    if (ld == P) big_memcpy(hp, theta_host + lo * ld, (size_t)m * P * sizeof(double));
                    else for (int64_t i = 0; i < m; ++i) std::memcpy(hp + i * P, theta_host + (lo + i) * ld, (size_t)P * sizeof(double));
                    src_ld = P;
                } else {
                    for (int c = 0; c < P; ++c) big_memcpy(hp + (int64_t)c * m, theta_host + (int64_t)c * ld + lo, (size_t)m * sizeof(double));
                    src_ld = m;
                }
                src = hp;
            }
            int64_t dld;
            if (layout == STN_LAYOUT_AOS && src_ld == P) {
                STN_CUDA(cudaMemcpyAsync(dth, src, (size_t)m * P * sizeof(double), cudaMemcpyHostToDevice, st));
                dld = P;
            } else if (layout == STN_LAYOUT_AOS) {
                // padded rows: keep only the first P columns of each
                STN_CUDA(cudaMemcpy2DAsync(dth, (size_t)P * sizeof(double), src, (size_t)src_ld * sizeof(double),
                                           (size_t)P * sizeof(double), (size_t)m, cudaMemcpyHostToDevice, st));
                dld = P;
            } else {
                STN_CUDA(cudaMemcpy2DAsync(dth, (size_t)m * sizeof(double), src, (size_t)src_ld * sizeof(double),
                                           (size_t)m * sizeof(double), (size_t)P, cudaMemcpyHostToDevice, st));
                dld = m;
            }
            if (!pin_in) STN_CUDA(cudaEventRecord(ctx->h2d_done[slot], st));
            const int r = launch_loglike(ctx, dth, m, dld, layout, mode, plan, dout, st, 0, 1 + slot);
            if (r) return r;
            if (pin_out) {
                STN_CUDA(cudaMemcpyAsync(out_host + lo, dout, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, st));
            } else {
                STN_CUDA(cudaMemcpyAsync(ctx->pin_out[slot], dout, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, st));
                STN_CUDA(cudaEventRecord(ctx->d2h_done[slot], st));
                pend[slot].lo = lo; pend[slot].m = m; pend[slot].live = true;
            }
            lo += m;
            rows = next_chunk_rows(ctx, rows);
        }
        // the chunks still in flight, oldest first
        for (int j = 0; j < StnCtx::kSlots; ++j) {
            const int slot = (k + j) % StnCtx::kSlots;
            if (pend[slot].live) {
                STN_CUDA(cudaEventSynchronize(ctx->d2h_done[slot]));
                std::memcpy(out_host + pend[slot].lo, ctx->pin_out[slot], (size_t)pend[slot].m * sizeof(double));
                pend[slot].live = false;
            }
        }
        for (int i = 0; i < StnCtx::kSlots; ++i) STN_CUDA(cudaStreamSynchronize(ctx->streams[i]));
        return STN_OK;
    };
    rc = body();
    if (rc) host_pipe_drain(ctx);     // nothing of this call may still touch the caller's buffers after an error
    return rc;
}

}  // namespace

extern "C" {

int stn_version(void) { return STN_ABI_VERSION; }

const char* stn_last_error(void) { return g_err.c_str(); }

int stn_create(const StnProblem* pb, int device, StnCtx** out_ctx) {
    if (!pb || !out_ctx) return fail(STN_ERR_INVALID, "null argument");
    *out_ctx = nullptr;
    if (pb->n_sources <= 0 || !pb->sources) return fail(STN_ERR_INVALID, "problem has no sources");
    if (pb->n_params <= 0) return fail(STN_ERR_INVALID, "problem has no parameters");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(STN_ERR_DEVICE, "no CUDA device visible (this library has no CPU path)");
    if (device < 0 || device >= ndev) return fail(STN_ERR_INVALID, "device %d out of range (%d visible)", device, ndev);
    cudaDeviceProp prop;
    STN_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(STN_ERR_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                    prop.major, prop.minor);
    STN_CUDA(cudaSetDevice(device));

    StnCtx* ctx = new (std::nothrow) StnCtx();
    if (!ctx) return fail(STN_ERR_NOMEM, "out of host memory");
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;

    auto bail = [&](int code) {
        stn_destroy(ctx);
        return code;
    };

    std::vector<SrcDev> src(pb->n_sources);
    std::vector<ChunkDev> chunks;
    for (int s = 0; s < pb->n_sources; ++s) {
        const StnSource& hs = pb->sources[s];
        if (hs.n_epochs < 0 || hs.n_epochs > 0x3fffffff) return bail(fail(STN_ERR_INVALID, "source %d: bad epoch count", s));
        if (hs.n_epochs > 0 && (!hs.fast_rows || !hs.strict_rows))
            return bail(fail(STN_ERR_INVALID, "source %d: null epoch table", s));
        for (int r = 0; r < STN_NROOTS; ++r)
            if (hs.col[r] >= pb->n_params)
                return bail(fail(STN_ERR_INVALID, "source %d: column %d out of range (P = %d)", s, hs.col[r], pb->n_params));
        if (hs.efac_on && hs.col[STN_EFAC] < 0)
            return bail(fail(STN_ERR_INVALID, "source %d: efac_on without an efac column", s));
        SrcDev& d = src[s];
        std::memset(&d, 0, sizeof(d));
        int rc = upload(ctx, hs.fast_rows, (size_t)hs.n_epochs * STN_FAST_ROW, &d.fast_rows);
        if (rc) return bail(rc);
        rc = upload(ctx, hs.strict_rows, (size_t)hs.n_epochs * STN_STRICT_ROW, &d.strict_rows);
        if (rc) return bail(rc);
        d.n_epochs = (int)hs.n_epochs;
        for (int r = 0; r < STN_NROOTS; ++r) d.col[r] = hs.col[r];
        d.efac_on = hs.efac_on ? 1 : 0;
        d.binary = hs.binary ? 1 : 0;
        d.v_shift = 0;
        d.v_lo_bits = 0;
        d.v_hi = d.v_hib = 0.0;
        d.chi_scale = 1.0;
        if (d.efac_on && hs.n_epochs > 0) {
            // variance v = a + g*b with a = err_random^2, b = err_sys^2, g = efac^2 >= 0 (squared
            // radian-scale errors, ~1e-19).  The device copy of the four variance columns is stored
            // times 2^k with k chosen so that the scaled a's straddle 1: products of many variances
            // then stay in range.  The scaling is exact and undone exactly at the fold.
            double a_min = INFINITY, a_max = 0.0, b_max = 0.0;
            for (int64_t e = 0; e < hs.n_epochs; ++e) {
                const double* row = hs.fast_rows + e * STN_FAST_ROW;
                for (int k = 6; k < 8; ++k) { if (row[k] < a_min) a_min = row[k]; if (row[k] > a_max) a_max = row[k]; }
                for (int k = 8; k < 10; ++k) if (row[k] > b_max) b_max = row[k];
            }
            const bool usable = std::isfinite(a_max) && std::isfinite(b_max) && a_max > 0.0 && a_min >= 0.0;
            int k = 0;
            if (usable) {
                const double lo = a_min > 0.0 ? a_min : a_max;
                k = -(std::ilogb(lo) + std::ilogb(a_max)) / 2;
                if (k > 900) k = 900;
                if (k < -900) k = -900;
            }
            std::vector<double> scaled(hs.fast_rows, hs.fast_rows + (size_t)hs.n_epochs * STN_FAST_ROW);
            for (int64_t e = 0; e < hs.n_epochs; ++e)
                for (int c = 6; c < 10; ++c) scaled[(size_t)e * STN_FAST_ROW + c] = std::ldexp(scaled[(size_t)e * STN_FAST_ROW + c], k);
            if (cudaMemcpy(const_cast<double*>(d.fast_rows), scaled.data(), scaled.size() * sizeof(double),
                           cudaMemcpyHostToDevice) != cudaSuccess)
                return bail(fail(STN_ERR_CUDA, "source %d: upload of the scaled variance columns failed", s));
            d.v_shift = k;
            d.chi_scale = std::ldexp(1.0, k);
            d.v_hi = std::ldexp(a_max, k);
            d.v_hib = std::ldexp(b_max, k);
            const double lo_s = std::ldexp(a_min, k);
            // a == 0 somewhere (no random error): nothing bounds the variances from below, renormalise every epoch
            d.v_lo_bits = (usable && lo_s > 0.0 && std::isfinite(lo_s)) ? (lo_s >= 1.0 ? 0 : -std::ilogb(lo_s)) : 4096;
        }
        d.cos_decj = hs.binary ? hs.cos_decj : 1.0;
        d.inv_cos_decj = 1.0 / d.cos_decj;
        d.neg_sum_log_err = hs.neg_sum_log_err;
        d.a1_lts = hs.a1_lts;
        ctx->n_epochs.push_back(d.n_epochs);
        if (d.n_epochs > ctx->max_epochs) ctx->max_epochs = d.n_epochs;
        // chunk list: every source contributes at least one (possibly empty) chunk so that
        // its prologue / fold always runs
        int done = 0;
        do {
            ChunkDev c;
            c.rows = d.fast_rows + (size_t)done * STN_FAST_ROW;
            c.n_epochs = (d.n_epochs - done) < kChunk ? (d.n_epochs - done) : kChunk;
            c.source = s;
            c.flags = (done == 0 ? kFirst : 0);
            c.pad = 0;
            done += c.n_epochs;
            if (done >= d.n_epochs) c.flags |= kLast;
            chunks.push_back(c);
        } while (done < d.n_epochs);
    }
    {
        int64_t ef = 0, all = 0;
        for (const SrcDev& d : src) { all += d.n_epochs; if (d.efac_on) ef += d.n_epochs; }
        ctx->efac_heavy = all > 0 && 2 * ef >= all;
    }
    for (int k = 0; k < pb->n_a1dot; ++k)
        if (pb->a1dot_source[k] < 0 || pb->a1dot_source[k] >= pb->n_sources)
            return bail(fail(STN_ERR_INVALID, "a1dot constraint %d: bad source", k));
    for (int k = 0; k < pb->n_sini; ++k)
        if (pb->sini_col[k] < 0 || pb->sini_col[k] >= pb->n_params)
            return bail(fail(STN_ERR_INVALID, "sin(incl) constraint %d: bad column", k));

    ProbDev& P = ctx->prob;
    P.n_sources = pb->n_sources;
    P.n_params = pb->n_params;
    P.n_chunks = (int)chunks.size();
    P.n_a1dot = pb->n_a1dot;
    P.n_sini = pb->n_sini;
    P.mas2rad = pb->mas2rad > 0.0 ? pb->mas2rad : STN_MAS2RAD;
    P.deg2rad = pb->deg2rad > 0.0 ? pb->deg2rad : STN_DEG2RAD;
    P.masyr2rads = pb->masyr2rads > 0.0 ? pb->masyr2rads : STN_MASYR2RADS;
    int rc;
    if ((rc = upload(ctx, src.data(), src.size(), &P.src))) return bail(rc);
    if ((rc = upload(ctx, chunks.data(), chunks.size(), &P.chunks))) return bail(rc);
    if ((rc = upload(ctx, pb->a1dot_source, (size_t)pb->n_a1dot, &P.a1_src))) return bail(rc);
    if ((rc = upload(ctx, pb->a1dot_mu, (size_t)pb->n_a1dot, &P.a1_mu))) return bail(rc);
    if ((rc = upload(ctx, pb->a1dot_sigma, (size_t)pb->n_a1dot, &P.a1_sigma))) return bail(rc);
    if ((rc = upload(ctx, pb->sini_col, (size_t)pb->n_sini, &P.sini_col))) return bail(rc);
    if ((rc = upload(ctx, pb->sini_mu, (size_t)pb->n_sini, &P.sini_mu))) return bail(rc);
    if ((rc = upload(ctx, pb->sini_sigma, (size_t)pb->n_sini, &P.sini_sigma))) return bail(rc);
    *out_ctx = ctx;
    return STN_OK;
}

int stn_destroy(StnCtx* ctx) {
    if (!ctx) return STN_OK;
    cudaSetDevice(ctx->device);
    for (void* p : ctx->dev_allocs) cudaFree(p);
    for (void* p : ctx->prior_allocs) cudaFree(p);
    for (int i = 0; i < 1 + StnCtx::kSlots; ++i)
        if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
    if (ctx->small_in) cudaFreeHost(ctx->small_in);
    if (ctx->small_out) cudaFreeHost(ctx->small_out);
    for (int i = 0; i < StnCtx::kSlots; ++i) {
        if (ctx->stage_theta[i]) cudaFree(ctx->stage_theta[i]);
        if (ctx->stage_out[i]) cudaFree(ctx->stage_out[i]);
        if (ctx->pin_theta[i]) cudaFreeHost(ctx->pin_theta[i]);
        if (ctx->pin_out[i]) cudaFreeHost(ctx->pin_out[i]);
        if (ctx->h2d_done[i]) cudaEventDestroy(ctx->h2d_done[i]);
        if (ctx->d2h_done[i]) cudaEventDestroy(ctx->d2h_done[i]);
        if (ctx->streams[i]) cudaStreamDestroy(ctx->streams[i]);
    }
    if (ctx->gather_buf) cudaFree(ctx->gather_buf);
    if (ctx->sampler_buf) cudaFree(ctx->sampler_buf);
    if (ctx->sampler_nacc) cudaFree(ctx->sampler_nacc);
    for (int i = 0; i < 2; ++i)
        if (ctx->sampler_ev[i]) cudaEventDestroy(ctx->sampler_ev[i]);
    if (ctx->timing_stream) cudaStreamDestroy(ctx->timing_stream);
    delete ctx;
    return STN_OK;
}

int stn_auto_plan(StnCtx* ctx, int64_t n) {
    if (!ctx) return fail(STN_ERR_INVALID, "null context");
    return pick_plan(ctx, n);
}

int64_t stn_launch_count(StnCtx* ctx) { return ctx ? ctx->launches : 0; }

int stn_loglike(StnCtx* ctx, const double* theta_dev, int64_t n, int64_t ld, int layout, int mode,
                int plan, double* out_dev, void* stream) {
    int rc = check_call(ctx, theta_dev, n, ld, layout, mode, out_dev);
    if (rc) return rc;
    STN_CUDA(cudaSetDevice(ctx->device));
    return launch_loglike(ctx, theta_dev, n, ld, layout, mode, plan, out_dev, static_cast<cudaStream_t>(stream));
}

int stn_loglike_scatter(StnCtx* ctx, const double* theta_dev, int64_t n, int64_t ld, int layout, int mode,
                        int plan, double* const* outs, int n_outs, void* stream) {
    if (!outs || n_outs < 1 || n_outs > STN_MAX_OUTS) return fail(STN_ERR_INVALID, "1..%d output buffers are needed", STN_MAX_OUTS);
    for (int k = 0; k < n_outs; ++k)
        if (n > 0 && !outs[k]) return fail(STN_ERR_INVALID, "null output buffer %d", k);
    int rc = check_call(ctx, theta_dev, n, ld, layout, mode, outs[0]);
    if (rc) return rc;
    STN_CUDA(cudaSetDevice(ctx->device));
    return launch_loglike(ctx, theta_dev, n, ld, layout, mode, plan, outs[0], static_cast<cudaStream_t>(stream), 0, 0,
                          outs + 1, n_outs - 1);
}

int stn_dev_alloc(int device, void** ptr, int64_t bytes) {
    if (!ptr || bytes < 0) return fail(STN_ERR_INVALID, "bad argument");
    STN_CUDA(cudaSetDevice(device));
    STN_CUDA(cudaMalloc(ptr, (size_t)(bytes ? bytes : 1)));
    return STN_OK;
}

int stn_dev_free(int device, void* ptr) {
    if (!ptr) return STN_OK;
    STN_CUDA(cudaSetDevice(device));
    STN_CUDA(cudaFree(ptr));
    return STN_OK;
}

int stn_ipc_export(void* dev_ptr, void* handle64) {
    static_assert(sizeof(cudaIpcMemHandle_t) == STN_IPC_HANDLE_BYTES, "IPC handle size");
    if (!dev_ptr || !handle64) return fail(STN_ERR_INVALID, "null argument");
    cudaIpcMemHandle_t h;
    STN_CUDA(cudaIpcGetMemHandle(&h, dev_ptr));
    std::memcpy(handle64, &h, sizeof(h));
    return STN_OK;
}

int stn_ipc_open(int device, const void* handle64, void** peer_ptr) {
    if (!handle64 || !peer_ptr) return fail(STN_ERR_INVALID, "null argument");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, sizeof(h));
    STN_CUDA(cudaSetDevice(device));
    STN_CUDA(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return STN_OK;
}

int stn_ipc_close(int device, void* peer_ptr) {
    if (!peer_ptr) return STN_OK;
    STN_CUDA(cudaSetDevice(device));
    STN_CUDA(cudaIpcCloseMemHandle(peer_ptr));
    return STN_OK;
}

int stn_set_prior(StnCtx* ctx, const StnPrior* pr) {
    if (!ctx || !pr) return fail(STN_ERR_INVALID, "null argument");
    if (pr->n_params != ctx->prob.n_params)
        return fail(STN_ERR_INVALID, "prior has %d entries, the problem has %d parameters", pr->n_params,
                    ctx->prob.n_params);
    if (!pr->kind || !pr->a || !pr->b) return fail(STN_ERR_INVALID, "null prior arrays");
    double c = 0.0;
    for (int k = 0; k < pr->n_params; ++k) {
        const double a = pr->a[k], b = pr->b[k];
        switch (pr->kind[k]) {
            case STN_PRIOR_UNIFORM:
                if (!(b > a)) return fail(STN_ERR_INVALID, "prior %d: Uniform needs a < b", k);
                c += -std::log(b - a);
                break;
            case STN_PRIOR_GAUSSIAN:
                if (!(b > 0)) return fail(STN_ERR_INVALID, "prior %d: Gaussian needs sigma > 0", k);
                c += -std::log(b) - 0.5 * std::log(2.0 * 3.14159265358979323846);
                break;
            case STN_PRIOR_SINE:
                if (!(b > a) || a < 0.0 || b > 3.14159265358979323846)
                    return fail(STN_ERR_INVALID, "prior %d: Sine needs 0 <= a < b <= pi", k);
                c += -std::log(std::cos(a) - std::cos(b));
                break;
            default: return fail(STN_ERR_INVALID, "prior %d: unknown kind %d", k, pr->kind[k]);
        }
    }
    for (int k = 0; k < pr->n_sin_limits; ++k)
        if (pr->sin_col[k] < 0 || pr->sin_col[k] >= pr->n_params)
            return fail(STN_ERR_INVALID, "sin limit %d: bad column", k);
    STN_CUDA(cudaSetDevice(ctx->device));
    STN_CUDA(cudaDeviceSynchronize());                 // no launch may still read the old prior
    // only the previous prior's arrays are replaced; scratch / staging buffers stay as they are
    for (void* p : ctx->prior_allocs) cudaFree(p);
    ctx->prior_allocs.clear();
    ctx->has_prior = false;
    ProbDev P = ctx->prob;
    // upload() records into dev_allocs; move those entries to prior_allocs so they can be replaced
    const size_t mark = ctx->dev_allocs.size();
    int rc = STN_OK;
    if (!rc) rc = upload(ctx, pr->kind, (size_t)pr->n_params, &P.prior_kind);
    if (!rc) rc = upload(ctx, pr->a, (size_t)pr->n_params, &P.prior_a);
    if (!rc) rc = upload(ctx, pr->b, (size_t)pr->n_params, &P.prior_b);
    if (!rc) rc = upload(ctx, pr->sin_col, (size_t)pr->n_sin_limits, &P.sinlim_col);
    if (!rc) rc = upload(ctx, pr->sin_lo, (size_t)pr->n_sin_limits, &P.sinlim_lo);
    if (!rc) rc = upload(ctx, pr->sin_hi, (size_t)pr->n_sin_limits, &P.sinlim_hi);
    ctx->prior_allocs.assign(ctx->dev_allocs.begin() + mark, ctx->dev_allocs.end());
    ctx->dev_allocs.resize(mark);
    if (rc) return rc;
    P.n_prior = pr->n_params;
    P.n_sinlim = pr->n_sin_limits;
    P.prior_const = c;
    ctx->prior_prob = P;
    ctx->has_prior = true;
    return STN_OK;
}

int stn_logpost(StnCtx* ctx, const double* theta_dev, int64_t n, int64_t ld, int layout, int mode,
                int plan, double* out_dev, void* stream) {
    int rc = check_call(ctx, theta_dev, n, ld, layout, mode, out_dev);
    if (rc) return rc;
    if (!ctx->has_prior) return fail(STN_ERR_INVALID, "stn_logpost needs stn_set_prior first");
    STN_CUDA(cudaSetDevice(ctx->device));
    return launch_loglike(ctx, theta_dev, n, ld, layout, mode, plan, out_dev, static_cast<cudaStream_t>(stream),
                          1);
}

int stn_chisq(StnCtx* ctx, const double* theta_dev, int64_t n, int64_t ld, int layout, int mode,
              int plan, double* out_dev, void* stream) {
    int rc = check_call(ctx, theta_dev, n, ld, layout, mode, out_dev);
    if (rc) return rc;
    STN_CUDA(cudaSetDevice(ctx->device));
    return launch_loglike(ctx, theta_dev, n, ld, layout, mode, plan, out_dev, static_cast<cudaStream_t>(stream),
                          2);
}

int stn_loglike_timed(StnCtx* ctx, const double* theta_dev, int64_t n, int64_t ld, int layout,
                      int mode, int plan, double* out_dev, int steps, double* total_ms) {
    int rc = check_call(ctx, theta_dev, n, ld, layout, mode, out_dev);
    if (rc) return rc;
    if (steps <= 0 || !total_ms) return fail(STN_ERR_INVALID, "bad steps / null total_ms");
    STN_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->timing_stream) STN_CUDA(cudaStreamCreateWithFlags(&ctx->timing_stream, cudaStreamNonBlocking));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    float ms = 0.f;
    auto run = [&]() -> int {
        STN_CUDA(cudaEventCreate(&e0));
        STN_CUDA(cudaEventCreate(&e1));
        STN_CUDA(cudaDeviceSynchronize());
        STN_CUDA(cudaEventRecord(e0, ctx->timing_stream));
        for (int i = 0; i < steps; ++i) {
            const int r = launch_loglike(ctx, theta_dev, n, ld, layout, mode, plan, out_dev, ctx->timing_stream);
            if (r) return r;
        }
        STN_CUDA(cudaEventRecord(e1, ctx->timing_stream));
        STN_CUDA(cudaEventSynchronize(e1));
        STN_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        return STN_OK;
    };
    rc = run();
    if (rc) cudaStreamSynchronize(ctx->timing_stream);    // nothing of this call may still be running on return
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (rc) return rc;
    *total_ms = ms;
    return STN_OK;
}

int stn_loglike_host(StnCtx* ctx, const double* theta_host, int64_t n, int64_t ld, int layout,
                     int mode, int plan, double* out_host) {
    int rc = check_call(ctx, theta_host, n, ld, layout, mode, out_host);
    if (rc) return rc;
    if (n == 0) return STN_OK;
    STN_CUDA(cudaSetDevice(ctx->device));
    if (mode == STN_MODE_FAST && plan == 0) plan = pick_plan(ctx, n);    // from the whole batch
    return run_host(ctx, theta_host, n, ld, layout, mode, plan, out_host);
}

// ---------------------------------------------------------------------------------
// Several devices as one (single process, e.g. one sampler): contiguous slices of the batch, one
// per device, every device holding the full epoch tables.  No collective: for host batches each
// device DMAs its slice of the result straight into the caller's array; for device-resident
// slices the results are gathered with peer copies (NVLink) into one buffer.
// ---------------------------------------------------------------------------------
int stn_multi_create(const StnProblem* pb, const int* devices, int n_devices, StnMulti** out) {
    if (!pb || !devices || !out) return fail(STN_ERR_INVALID, "null argument");
    *out = nullptr;
    if (n_devices <= 0 || n_devices > 64) return fail(STN_ERR_INVALID, "bad device count %d", n_devices);
    StnMulti* m = new (std::nothrow) StnMulti();
    if (!m) return fail(STN_ERR_NOMEM, "out of host memory");
    for (int g = 0; g < n_devices; ++g) {
        StnCtx* c = nullptr;
        const int rc = stn_create(pb, devices[g], &c);
        if (rc) {
            stn_multi_destroy(m);
            return rc;
        }
        m->ctx.push_back(c);
    }
    // peer access for stn_multi_loglike_dev: a kernel then stores its results straight into the
    // gathering device's buffer over NVLink (an already-enabled pair is not an error)
    m->peer_ok.assign((size_t)n_devices * n_devices, 0);
    for (int a = 0; a < n_devices; ++a)
        for (int b = 0; b < n_devices; ++b) {
            if (devices[a] == devices[b]) {
                m->peer_ok[(size_t)a * n_devices + b] = 1;
                continue;
            }
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, devices[a], devices[b]) == cudaSuccess && can) {
                cudaSetDevice(devices[a]);
                const cudaError_t e = cudaDeviceEnablePeerAccess(devices[b], 0);
                if (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled) m->peer_ok[(size_t)a * n_devices + b] = 1;
            }
            (void)cudaGetLastError();
        }
    *out = m;
    return STN_OK;
}

int stn_multi_destroy(StnMulti* m) {
    if (!m) return STN_OK;
    for (StnCtx* c : m->ctx) stn_destroy(c);
    delete m;
    return STN_OK;
}

int stn_multi_size(StnMulti* m) { return m ? (int)m->ctx.size() : 0; }

StnCtx* stn_multi_ctx(StnMulti* m, int i) {
    if (!m || i < 0 || i >= (int)m->ctx.size()) return nullptr;
    return m->ctx[i];
}

int64_t stn_multi_launch_count(StnMulti* m) {
    int64_t t = 0;
    if (m) for (StnCtx* c : m->ctx) t += c->launches;
    return t;
}

int stn_multi_slice(int64_t n, int n_parts, int part, int64_t* lo, int64_t* hi) {
    if (n < 0 || n_parts <= 0 || part < 0 || part >= n_parts || !lo || !hi) return fail(STN_ERR_INVALID, "bad slice request");
    const int64_t base = n / n_parts, extra = n % n_parts;
    *lo = part * base + (part < extra ? part : extra);
    *hi = *lo + base + (part < extra ? 1 : 0);
    return STN_OK;
}

int stn_multi_loglike_host(StnMulti* m, const double* theta_host, int64_t n, int64_t ld, int layout,
                           int mode, int plan, double* out_host) {
    if (!m || m->ctx.empty()) return fail(STN_ERR_INVALID, "null multi-context");
    int rc = check_call(m->ctx[0], theta_host, n, ld, layout, mode, out_host);
    if (rc) return rc;
    if (n == 0) return STN_OK;
    // the plan fixes the summation order: chosen from the WHOLE batch so that the bits do not depend on
    // how many devices share it (R, the samples per thread, never changes a bit)
    if (mode == STN_MODE_FAST && plan == 0) plan = pick_plan(m->ctx[0], n);
    // devices used: a slice below 2^15 rows does not fill one B200
    int G = (int)m->ctx.size();
    const int64_t want = n >> 15;
    if (want < G) G = want < 1 ? 1 : (int)want;
    if (G == 1) {
        STN_CUDA(cudaSetDevice(m->ctx[0]->device));
        return run_host(m->ctx[0], theta_host, n, ld, layout, mode, plan, out_host);
    }
    std::vector<int> rcs(G, STN_OK);
    std::vector<std::string> errs(G);
    auto work = [&](int g) {
        int64_t lo = 0, hi = 0;
        stn_multi_slice(n, G, g, &lo, &hi);
        StnCtx* c = m->ctx[g];
        if (c->host_share < G) c->host_share = G;       // G devices copy from this host at once
        if (cudaSetDevice(c->device) != cudaSuccess) {
            rcs[g] = STN_ERR_CUDA;
            errs[g] = "cudaSetDevice failed";
            return;
        }
        const double* th = layout == STN_LAYOUT_AOS ? theta_host + lo * ld : theta_host + lo;
        rcs[g] = run_host(c, th, hi - lo, ld, layout, mode, plan, out_host + lo);
        if (rcs[g]) errs[g] = g_err;                    // g_err is thread-local
    };
    std::vector<std::thread> pool;
    pool.reserve(G - 1);
    for (int g = 1; g < G; ++g) pool.emplace_back(work, g);
    work(0);
    for (std::thread& t : pool) t.join();
    cudaSetDevice(m->ctx[0]->device);
    for (int g = 0; g < G; ++g)
        if (rcs[g]) return fail(rcs[g], "device slice %d: %s", g, errs[g].c_str());
    return STN_OK;
}

int stn_multi_loglike_dev(StnMulti* m, const double* const* theta_dev, const int64_t* counts, int64_t ld,
                          int layout, int mode, int plan, double* out_dev, int out_device) {
    if (!m || m->ctx.empty() || !theta_dev || !counts) return fail(STN_ERR_INVALID, "null argument");
    const int G = (int)m->ctx.size();
    int64_t n = 0;
    for (int g = 0; g < G; ++g) {
        if (counts[g] < 0) return fail(STN_ERR_INVALID, "negative slice size");
        if (counts[g] > 0 && !theta_dev[g]) return fail(STN_ERR_INVALID, "null slice");
        n += counts[g];
    }
    if (n == 0) return STN_OK;
    if (!out_dev) return fail(STN_ERR_INVALID, "null output");
    if (layout != STN_LAYOUT_AOS && layout != STN_LAYOUT_SOA) return fail(STN_ERR_INVALID, "bad layout %d", layout);
    if (mode != STN_MODE_FAST && mode != STN_MODE_STRICT) return fail(STN_ERR_INVALID, "bad mode %d", mode);
    if (mode == STN_MODE_FAST && plan == 0) plan = pick_plan(m->ctx[0], n);      // from the whole batch
    int rc = STN_OK;
    int64_t off = 0;
    for (int g = 0; g < G && !rc; ++g) {
        StnCtx* c = m->ctx[g];
        const int64_t cnt = counts[g];
        if (cnt == 0) continue;
        if ((rc = check_call(c, theta_dev[g], cnt, ld, layout, mode, out_dev))) break;
        STN_CUDA(cudaSetDevice(c->device));
        if ((rc = host_pipe_init(c))) break;
        cudaStream_t st = c->streams[0];
        bool direct = c->device == out_device;
        for (int b = 0; b < G && !direct; ++b)
            direct = m->ctx[b]->device == out_device && m->peer_ok[(size_t)g * G + b];
        if (direct) {
            // out_dev is local or peer-mapped: the kernel's own stores are the gather (8 bytes per vector over NVLink)
            rc = launch_loglike(c, theta_dev[g], cnt, ld, layout, mode, plan, out_dev + off, st, 0, 1);
        } else {
            if (c->gather_elems < cnt) {
                STN_CUDA(cudaStreamSynchronize(st));
                if (c->gather_buf) cudaFree(c->gather_buf);
                c->gather_buf = nullptr;
                c->gather_elems = 0;
                STN_CUDA(cudaMalloc(&c->gather_buf, (size_t)cnt * sizeof(double)));
                c->gather_elems = cnt;
            }
            rc = launch_loglike(c, theta_dev[g], cnt, ld, layout, mode, plan, c->gather_buf, st, 0, 1);
            if (!rc) {
                const cudaError_t e = cudaMemcpyPeerAsync(out_dev + off, out_device, c->gather_buf, c->device,
                                                          (size_t)cnt * sizeof(double), st);
                if (e != cudaSuccess) rc = fail(STN_ERR_CUDA, "peer copy %d -> %d failed: %s", c->device, out_device,
                                                cudaGetErrorString(e));
            }
        }
        off += cnt;
    }
    // synchronous on return (also after an error: nothing may still be running)
    for (int g = 0; g < G; ++g) {
        StnCtx* c = m->ctx[g];
        if (c->streams[0]) {
            cudaSetDevice(c->device);
            const cudaError_t e = cudaStreamSynchronize(c->streams[0]);
            if (e != cudaSuccess && !rc) rc = fail(STN_ERR_CUDA, "device %d: %s", c->device, cudaGetErrorString(e));
        }
    }
    cudaSetDevice(m->ctx[0]->device);
    return rc;
}

int stn_multi_set_prior(StnMulti* m, const StnPrior* prior) {
    if (!m || m->ctx.empty()) return fail(STN_ERR_INVALID, "null multi-context");
    for (StnCtx* c : m->ctx) {
        const int rc = stn_set_prior(c, prior);
        if (rc) return rc;
    }
    return STN_OK;
}

int stn_multi_ensemble_run(StnMulti* m, double* const* x_dev, double* const* lp_dev, int64_t n_walkers, int64_t ld,
                           double a, uint64_t seed, uint64_t step0, int64_t n_steps, int64_t thin, double* chain_dev,
                           double* lnprob_dev, int64_t* n_accept_host, int mode, int plan) {
    if (!m || m->ctx.empty() || !x_dev || !lp_dev) return fail(STN_ERR_INVALID, "null argument");
    const int G = (int)m->ctx.size();
    if (G > STN_MAX_OUTS) return fail(STN_ERR_INVALID, "at most %d devices can share one ensemble", STN_MAX_OUTS);
    const int P = m->ctx[0]->prob.n_params;
    if (n_walkers < 2 || (n_walkers & 1) || ld < P || !(a > 1.0) || n_steps < 0 || thin < 1)
        return fail(STN_ERR_INVALID, "stretch move needs an even number of walkers >= 2, ld >= P, a > 1, thin >= 1");
    if (mode != STN_MODE_FAST && mode != STN_MODE_STRICT) return fail(STN_ERR_INVALID, "bad mode %d", mode);
    for (int g = 0; g < G; ++g) {
        if (!m->ctx[g]->has_prior) return fail(STN_ERR_INVALID, "stn_multi_ensemble_run needs stn_multi_set_prior first");
        if (!x_dev[g] || !lp_dev[g]) return fail(STN_ERR_INVALID, "null walker copy %d", g);
        for (int b = 0; b < G; ++b)
            if (!m->peer_ok[(size_t)g * G + b])
                return fail(STN_ERR_DEVICE, "devices %d and %d have no peer access", m->ctx[g]->device, m->ctx[b]->device);
    }
    const int64_t H = n_walkers / 2;
    if (plan == 0) plan = pick_plan(m->ctx[0], H);      // from the WHOLE half-ensemble: one summation order whatever the number of devices
    // per device: proposal range of a half-step, scratch, events, acceptance counter
    std::vector<int64_t> k0(G), cnt(G);
    for (int g = 0; g < G; ++g) {
        int64_t lo, hi;
        stn_multi_slice(H, G, g, &lo, &hi);
        k0[g] = lo;
        cnt[g] = hi - lo;
        StnCtx* c = m->ctx[g];
        STN_CUDA(cudaSetDevice(c->device));
        int rc = host_pipe_init(c);
        if (rc) return rc;
        const int64_t need = cnt[g] * P + 2 * cnt[g] + 1;
        if (c->sampler_elems < need) {
            STN_CUDA(cudaDeviceSynchronize());
            if (c->sampler_buf) cudaFree(c->sampler_buf);
            c->sampler_buf = nullptr;
            c->sampler_elems = 0;
            STN_CUDA(cudaMalloc(&c->sampler_buf, (size_t)need * sizeof(double)));
            c->sampler_elems = need;
        }
        for (int i = 0; i < 2; ++i)
            if (!c->sampler_ev[i]) STN_CUDA(cudaEventCreateWithFlags(&c->sampler_ev[i], cudaEventDisableTiming));
        if (!c->sampler_nacc) STN_CUDA(cudaMalloc(&c->sampler_nacc, sizeof(unsigned long long)));
        STN_CUDA(cudaMemsetAsync(c->sampler_nacc, 0, sizeof(unsigned long long), c->streams[0]));
    }
    // One host thread per device issues that device's launches, so the per-half-step driver calls (three
    // launches, G - 1 event waits, one record) of all devices run side by side instead of G-fold in series.
    // The threads meet at a host barrier after every half-step: an event must have been RECORDED by its
    // owner's thread before another thread makes its stream wait for it.
    struct HostBarrier {
        std::atomic<int> count{0};
        std::atomic<int> gen{0};
        int n;
        explicit HostBarrier(int n_) : n(n_) {}
        void wait() {
            const int g = gen.load(std::memory_order_acquire);
            if (count.fetch_add(1, std::memory_order_acq_rel) == n - 1) {
                count.store(0, std::memory_order_relaxed);
                gen.fetch_add(1, std::memory_order_release);
            } else {
                while (gen.load(std::memory_order_acquire) == g) std::this_thread::yield();
            }
        }
    };
    HostBarrier bar(G);
    std::atomic<int> failed{0};
    std::vector<int> rcs(G, STN_OK);
    std::vector<std::string> errs(G);
    auto half_step = [&](int g, int64_t t, int half, int64_t hs) -> int {
        StnCtx* c = m->ctx[g];
        cudaStream_t st = c->streams[0];
        // every device must have finished the previous half-step: its accepted walkers are in this
        // device's copy, and it no longer reads the rows this half-step overwrites
        if (hs > 0)
            for (int b = 0; b < G; ++b)
                if (b != g) STN_CUDA(cudaStreamWaitEvent(st, m->ctx[b]->sampler_ev[(hs - 1) & 1], 0));
        if (cnt[g] > 0) {
            double* y = c->sampler_buf;
            double* z = y + cnt[g] * P;
            double* lpy = z + cnt[g];
            const unsigned grid = (unsigned)((cnt[g] + kThreads - 1) / kThreads);
            (void)cudaGetLastError();
            stn_stretch_propose_part_kernel<<<grid, kThreads, 0, st>>>(x_dev[g], n_walkers, P, ld, half, a, seed,
                                                                       step0 + (uint64_t)t, k0[g], cnt[g], y, z);
            STN_CUDA(cudaGetLastError());
            const int rc = launch_loglike(c, y, cnt[g], P, STN_LAYOUT_AOS, mode, plan, lpy, st, 1, 1);
            if (rc) return rc;
            WalkerCopies C;
            C.n = G;
            C.x[0] = x_dev[g];                   // own copy first: it also supplies the current log-posterior
            C.lp[0] = lp_dev[g];
            for (int b = 0, o = 1; b < G; ++b)
                if (b != g) {
                    C.x[o] = x_dev[b];
                    C.lp[o] = lp_dev[b];
                    ++o;
                }
            WalkerCopies own = C;
            own.n = 1;                           // accept into the local copy ...
            stn_stretch_accept_part_kernel<<<grid, kThreads, 0, st>>>(own, n_walkers, P, ld, half, seed, step0 + (uint64_t)t,
                                                                      k0[g], cnt[g], y, z, lpy, c->sampler_nacc);
            STN_CUDA(cudaGetLastError());
            c->launches += 2;
            if (G > 1) {                         // ... then the slice goes to all other copies, coalesced
                const int64_t row0 = (half ? n_walkers / 2 : 0) + k0[g];
                const int64_t work = cnt[g] * ld / 2 + 1;
                const unsigned bg = (unsigned)((work + kThreads - 1) / kThreads < 4096 ? (work + kThreads - 1) / kThreads : 4096);
                stn_broadcast_rows_kernel<<<bg, kThreads, 0, st>>>(C, row0, cnt[g], ld);
                STN_CUDA(cudaGetLastError());
                c->launches++;
            }
        }
        STN_CUDA(cudaEventRecord(c->sampler_ev[hs & 1], st));
        return STN_OK;
    };
    auto store_chain = [&](int64_t t, int64_t hs) -> int {
        // device 0's copy is complete once every device's last half-step has landed
        StnCtx* c = m->ctx[0];
        cudaStream_t st = c->streams[0];
        for (int b = 1; b < G; ++b) STN_CUDA(cudaStreamWaitEvent(st, m->ctx[b]->sampler_ev[(hs - 1) & 1], 0));
        const int64_t slot = (t + 1) / thin - 1;
        STN_CUDA(cudaMemcpy2DAsync(chain_dev + slot * n_walkers * P, (size_t)P * sizeof(double), x_dev[0],
                                   (size_t)ld * sizeof(double), (size_t)P * sizeof(double), (size_t)n_walkers,
                                   cudaMemcpyDeviceToDevice, st));
        if (lnprob_dev)
            STN_CUDA(cudaMemcpyAsync(lnprob_dev + slot * n_walkers, lp_dev[0], (size_t)n_walkers * sizeof(double),
                                     cudaMemcpyDeviceToDevice, st));
        // the next half-step of the OTHER devices overwrites rows of device 0's copy: they wait for this copy
        STN_CUDA(cudaEventRecord(c->sampler_ev[(hs - 1) & 1], st));
        return STN_OK;
    };
    auto worker = [&](int g) {
        if (cudaSetDevice(m->ctx[g]->device) != cudaSuccess) {
            rcs[g] = STN_ERR_CUDA;
            errs[g] = "cudaSetDevice failed";
            failed.store(1);
        }
        int64_t hs = 0;
        for (int64_t t = 0; t < n_steps; ++t) {
            for (int half = 0; half < 2; ++half, ++hs) {
                if (!failed.load()) {
                    const int rc = half_step(g, t, half, hs);
                    if (rc) {
                        rcs[g] = rc;
                        errs[g] = g_err;
                        failed.store(1);
                    }
                }
                bar.wait();                     // all events of this half-step are recorded
            }
            if (chain_dev && (t + 1) % thin == 0) {
                if (g == 0 && !failed.load()) {
                    const int rc = store_chain(t, hs);
                    if (rc) {
                        rcs[g] = rc;
                        errs[g] = g_err;
                        failed.store(1);
                    }
                }
                bar.wait();                     // device 0's event now covers the copy
            }
        }
    };
    {
        std::vector<std::thread> pool;
        pool.reserve(G - 1);
        for (int g = 1; g < G; ++g) pool.emplace_back(worker, g);
        worker(0);
        for (std::thread& th : pool) th.join();
    }
    int rc = STN_OK;
    for (int g = 0; g < G && !rc; ++g)
        if (rcs[g]) rc = fail(rcs[g], "device %d: %s", m->ctx[g]->device, errs[g].c_str());
    // synchronous on return (also after an error); acceptances summed over the devices
    unsigned long long total = 0;
    for (int g = 0; g < G; ++g) {
        StnCtx* c = m->ctx[g];
        cudaSetDevice(c->device);
        if (c->streams[0]) {
            const cudaError_t e = cudaStreamSynchronize(c->streams[0]);
            if (e != cudaSuccess && !rc) rc = fail(STN_ERR_CUDA, "device %d: %s", c->device, cudaGetErrorString(e));
        }
        unsigned long long v = 0;
        if (!rc && c->sampler_nacc && cudaMemcpy(&v, c->sampler_nacc, sizeof(v), cudaMemcpyDeviceToHost) == cudaSuccess) total += v;
    }
    cudaSetDevice(m->ctx[0]->device);
    if (!rc && n_accept_host) *n_accept_host += (int64_t)total;
    return rc;
}

int stn_host_hint(StnCtx* ctx, int concurrent_devices) {
    if (!ctx || concurrent_devices < 1) return fail(STN_ERR_INVALID, "bad argument");
    ctx->host_share = concurrent_devices;
    return STN_OK;
}

int stn_host_register(void* ptr, int64_t bytes) {
    if (!ptr || bytes <= 0) return fail(STN_ERR_INVALID, "bad argument");
    STN_CUDA(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable));
    return STN_OK;
}

int stn_host_unregister(void* ptr) {
    if (ptr) STN_CUDA(cudaHostUnregister(ptr));
    return STN_OK;
}

int stn_model(StnCtx* ctx, int source, const double* theta_dev, int64_t n, int64_t ld, int layout,
              int mode, double* radec_dev, double* off_mas_dev, void* stream) {
    int rc = check_call(ctx, theta_dev, n, ld, layout, mode, radec_dev);
    if (rc) return rc;
    if (source < 0 || source >= ctx->prob.n_sources) return fail(STN_ERR_INVALID, "source %d out of range", source);
    const int64_t total = n * (int64_t)ctx->n_epochs[source];
    if (total == 0) return STN_OK;
    STN_CUDA(cudaSetDevice(ctx->device));
    const int64_t grid = (total + kThreads - 1) / kThreads;
    if (grid > 0x7fffffffLL) return fail(STN_ERR_INVALID, "batch too large for one launch");
    STN_DBG_RANGE(theta_dev, dbg_theta_bytes(n, ld, layout, ctx->prob.n_params), "theta");
    STN_DBG_RANGE(radec_dev, (size_t)total * 2 * sizeof(double), "radec");
    if (off_mas_dev) STN_DBG_RANGE(off_mas_dev, (size_t)total * 2 * sizeof(double), "off_mas");
    (void)cudaGetLastError();
    stn_model_kernel<<<(unsigned)grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        ctx->prob, source, theta_dev, n, ld, layout, mode, radec_dev, off_mas_dev);
    STN_CUDA(cudaGetLastError());
    ctx->launches++;
    return STN_OK;
}

int stn_stretch_propose(int device, const double* x_dev, int64_t n_walkers, int32_t n_params, int64_t ld,
                        int half, double a, uint64_t seed, uint64_t step, double* y_dev, double* z_dev,
                        void* stream) {
    if (!x_dev || !y_dev || !z_dev) return fail(STN_ERR_INVALID, "null buffer");
    if (n_walkers < 2 || (n_walkers & 1) || n_params <= 0 || ld < n_params || !(a > 1.0) || (half != 0 && half != 1))
        return fail(STN_ERR_INVALID, "stretch move needs an even number of walkers >= 2, ld >= P and a > 1");
    STN_CUDA(cudaSetDevice(device));
    const int64_t H = n_walkers / 2;
    (void)cudaGetLastError();
    stn_stretch_propose_kernel<<<(unsigned)((H + kThreads - 1) / kThreads), kThreads, 0,
                                 static_cast<cudaStream_t>(stream)>>>(x_dev, n_walkers, n_params, ld, half, a, seed,
                                                                      step, y_dev, z_dev);
    STN_CUDA(cudaGetLastError());
    return STN_OK;
}

int stn_stretch_accept(int device, double* x_dev, double* lp_dev, int64_t n_walkers, int32_t n_params,
                       int64_t ld, int half, uint64_t seed, uint64_t step, const double* y_dev,
                       const double* z_dev, const double* lpy_dev, int64_t* n_accept_dev, void* stream) {
    if (!x_dev || !lp_dev || !y_dev || !z_dev || !lpy_dev || !n_accept_dev) return fail(STN_ERR_INVALID, "null buffer");
    if (n_walkers < 2 || (n_walkers & 1) || n_params <= 0 || ld < n_params || (half != 0 && half != 1))
        return fail(STN_ERR_INVALID, "stretch move needs an even number of walkers >= 2 and ld >= P");
    STN_CUDA(cudaSetDevice(device));
    const int64_t H = n_walkers / 2;
    (void)cudaGetLastError();
    stn_stretch_accept_kernel<<<(unsigned)((H + kThreads - 1) / kThreads), kThreads, 0,
                                static_cast<cudaStream_t>(stream)>>>(
        x_dev, lp_dev, n_walkers, n_params, ld, half, seed, step, y_dev, z_dev, lpy_dev,
        reinterpret_cast<unsigned long long*>(n_accept_dev));
    STN_CUDA(cudaGetLastError());
    return STN_OK;
}

int stn_ensemble_run(StnCtx* ctx, double* x_dev, double* lp_dev, int64_t n_walkers, int64_t ld, double a,
                     uint64_t seed, uint64_t step0, int64_t n_steps, int64_t thin, double* chain_dev,
                     double* lnprob_dev, int64_t* n_accept_dev, int mode, int plan, int engine, void* stream) {
    if (!ctx || !x_dev || !lp_dev || !n_accept_dev) return fail(STN_ERR_INVALID, "null argument");
    if (!ctx->has_prior) return fail(STN_ERR_INVALID, "stn_ensemble_run needs stn_set_prior first");
    const int P = ctx->prob.n_params;
    if (n_walkers < 2 || (n_walkers & 1) || ld < P || !(a > 1.0) || n_steps < 0 || thin < 1)
        return fail(STN_ERR_INVALID, "stretch move needs an even number of walkers >= 2, ld >= P, a > 1, thin >= 1");
    if (mode != STN_MODE_FAST && mode != STN_MODE_STRICT) return fail(STN_ERR_INVALID, "bad mode %d", mode);
    if (n_steps == 0) return STN_OK;
    STN_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t H = n_walkers / 2;
    // engine 0: choose; 1: persistent cooperative kernel; 2: three launches per half-step
    int64_t evals = 0;
    for (int e : ctx->n_epochs) evals += e;
    const bool can_fuse = mode == STN_MODE_FAST && P <= STN_SAMPLER_MAX_P && (plan == 0 || plan == make_plan(1, 1));
    if (engine == 1 && !can_fuse)
        return fail(STN_ERR_INVALID, "the persistent sampler kernel needs mode fast, plan 0 or lanes=1|splits=1, P <= %d",
                    STN_SAMPLER_MAX_P);
    bool fuse = engine == 1 || (engine == 0 && can_fuse && evals <= 512);
    int grid = 0;
    size_t smem = 0;
    if (fuse) {
        // shared-memory image of the problem (sampler_stage_problem's layout)
        const ProbDev& pp = ctx->prior_prob;
        smem = sampler_align(sizeof(SrcDev) * pp.n_sources) + sampler_align(sizeof(ChunkDev) * pp.n_chunks);
        for (int e : ctx->n_epochs) smem += sampler_align(sizeof(double) * STN_FAST_ROW * (size_t)e);
        smem += 3 * sampler_align(8 * (size_t)pp.n_a1dot) + 3 * sampler_align(8 * (size_t)pp.n_sini) +
                3 * sampler_align(8 * (size_t)pp.n_prior) + 3 * sampler_align(8 * (size_t)pp.n_sinlim);
        if (smem > 200 * 1024) {
            if (engine == 1) return fail(STN_ERR_INVALID, "the problem (%zu bytes) does not fit the persistent kernel's shared memory", smem);
            fuse = false;
        }
    }
    if (fuse) {
        int per_sm = 0;
        STN_CUDA(cudaFuncSetAttribute(stn_ensemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        STN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stn_ensemble_kernel, kSamplerThreads, smem));
        const int64_t want = (H + kSamplerThreads - 1) / kSamplerThreads;
        const int64_t cap = (int64_t)per_sm * ctx->sm_count;
        if (cap < 1) {
            if (engine == 1) return fail(STN_ERR_CUDA, "the persistent sampler kernel does not fit on this device");
            fuse = false;
        }
        grid = (int)(want < cap ? want : cap);
    }
    if (fuse) {
        ProbDev prob = ctx->prior_prob;
        prob.out_chisq = 0;
        prob.n_extra_out = 0;
        unsigned long long* nacc = reinterpret_cast<unsigned long long*>(n_accept_dev);
        (void)cudaGetLastError();
        // up to 16 CTAs: one thread-block cluster, hardware barrier between half-steps
        if (grid <= 16 && STN_SAMPLER_CLUSTER) {
            bool ok = true;
            if (grid > 8)
                ok = cudaFuncSetAttribute(stn_ensemble_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
            if (ok) {
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3((unsigned)grid);
                cfg.blockDim = dim3(kSamplerThreads);
                cfg.dynamicSmemBytes = smem;
                cfg.stream = st;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = (unsigned)grid;
                at[0].val.clusterDim.y = 1;
                at[0].val.clusterDim.z = 1;
                cfg.attrs = at;
                cfg.numAttrs = 1;
                const cudaError_t e = cudaLaunchKernelEx(&cfg, stn_ensemble_kernel, prob, x_dev, lp_dev, n_walkers, ld, a, seed,
                                                         step0, n_steps, thin, chain_dev, lnprob_dev, nacc, 1);
                if (e == cudaSuccess) {
                    ctx->launches++;
                    return STN_OK;
                }
            }
            (void)cudaGetLastError();                     // cluster launch refused: fall back to the cooperative grid
        }
        int cluster_mode = 0;
        void* args[] = {&prob, &x_dev, &lp_dev, &n_walkers, &ld, &a, &seed, &step0, &n_steps, &thin, &chain_dev,
                        &lnprob_dev, &nacc, &cluster_mode};
        STN_CUDA(cudaLaunchCooperativeKernel((const void*)stn_ensemble_kernel, dim3((unsigned)grid), dim3(kSamplerThreads), args, smem, st));
        ctx->launches++;
        return STN_OK;
    }
    // generic path (long series: the throughput kernel per half-step), launches issued from here
    // instead of from Python; proposals and their log-posteriors live in the context's scratch
    if (plan == 0) plan = pick_plan(ctx, H);
    const int64_t need = H * P + 2 * H;
    if (ctx->sampler_elems < need) {
        STN_CUDA(cudaStreamSynchronize(st));
        if (ctx->sampler_buf) cudaFree(ctx->sampler_buf);
    if (ctx->sampler_nacc) cudaFree(ctx->sampler_nacc);
    for (int i = 0; i < 2; ++i)
        if (ctx->sampler_ev[i]) cudaEventDestroy(ctx->sampler_ev[i]);
        ctx->sampler_buf = nullptr;
        ctx->sampler_elems = 0;
        STN_CUDA(cudaMalloc(&ctx->sampler_buf, (size_t)need * sizeof(double)));
        ctx->sampler_elems = need;
    }
    double* y = ctx->sampler_buf;
    double* z = y + H * P;
    double* lpy = z + H;
    for (int64_t t = 0; t < n_steps; ++t) {
        for (int half = 0; half < 2; ++half) {
            int rc = stn_stretch_propose(ctx->device, x_dev, n_walkers, P, ld, half, a, seed, step0 + (uint64_t)t, y, z, stream);
            if (rc) return rc;
            rc = launch_loglike(ctx, y, H, P, STN_LAYOUT_AOS, mode, plan, lpy, st, 1);
            if (rc) return rc;
            rc = stn_stretch_accept(ctx->device, x_dev, lp_dev, n_walkers, P, ld, half, seed, step0 + (uint64_t)t, y, z, lpy,
                                    n_accept_dev, stream);
            if (rc) return rc;
            ctx->launches += 2;
        }
        if (chain_dev && (t + 1) % thin == 0) {
            const int64_t slot = (t + 1) / thin - 1;
            STN_CUDA(cudaMemcpy2DAsync(chain_dev + slot * n_walkers * P, (size_t)P * sizeof(double), x_dev,
                                       (size_t)ld * sizeof(double), (size_t)P * sizeof(double), (size_t)n_walkers,
                                       cudaMemcpyDeviceToDevice, st));
            if (lnprob_dev)
                STN_CUDA(cudaMemcpyAsync(lnprob_dev + slot * n_walkers, lp_dev, (size_t)n_walkers * sizeof(double),
                                         cudaMemcpyDeviceToDevice, st));
        }
    }
    return STN_OK;
}

int stn_host_alloc(void** ptr, int64_t bytes) {
    if (!ptr || bytes < 0) return fail(STN_ERR_INVALID, "bad argument");
    STN_CUDA(cudaHostAlloc(ptr, (size_t)(bytes ? bytes : 1), cudaHostAllocPortable));
    return STN_OK;
}

int stn_host_free(void* ptr) {
    if (ptr) STN_CUDA(cudaFreeHost(ptr));
    return STN_OK;
}

int stn_fp64_peak(int device, int iters, int reps, double* tflops, double* ms_out) {
    if (iters <= 0 || reps <= 0 || !tflops) return fail(STN_ERR_INVALID, "bad argument");
    STN_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    STN_CUDA(cudaGetDeviceProperties(&prop, device));
    double* sink = nullptr;
    STN_CUDA(cudaMalloc(&sink, sizeof(double)));
    cudaEvent_t e0, e1;
    STN_CUDA(cudaEventCreate(&e0));
    STN_CUDA(cudaEventCreate(&e1));
    const int grid = prop.multiProcessorCount * 8;
    const int block = 256;
    double best = 1e300;
    for (int r = 0; r < reps + 1; ++r) {   // first launch is the warm-up
        STN_CUDA(cudaEventRecord(e0, 0));
        stn_dfma_peak_kernel<<<grid, block>>>(sink, iters, 0.999999, 1e-9);
        STN_CUDA(cudaGetLastError());
        STN_CUDA(cudaEventRecord(e1, 0));
        STN_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        STN_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (r > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    const double flops = 2.0 * 8.0 * (double)iters * (double)grid * (double)block;
    *tflops = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return STN_OK;
}

}  // extern "C"
