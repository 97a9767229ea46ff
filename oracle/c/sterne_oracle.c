/*
 * CPU ORACLE in plain C -- TEST / BENCH INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Restates, in the reference's operation order (one IEEE rounding per operation, no FMA
 * contraction: build with -ffp-contract=off), the batched form of
 *   simulate.py:362-384   Gaussianlikelihood.log_likelihood
 *   simulate.py:303-320   adjust_errs_with_efac
 *   model/positions.py:62-121, 305-310   position + parallax offset
 *   model/reflex_motion.py:201-214       (Om_asc, incl) rotation of precomputed orbit terms
 *   model/kopeikin_effects.py:41-45      a1dot from proper motion
 * of dingswin/sterne (paths relative to /root/reference/src/sterne/).  Parameter-independent
 * per-epoch inputs (dT, Earth vector, orbit terms) are prepared by oracle/c_oracle.py from the
 * numpy oracle's own functions.  np.sum over the concatenated [RA..., Dec...] vector is
 * reproduced with numpy's pairwise summation so that results can be compared bit for bit
 * with oracle/sterne_oracle.py where libm and numpy agree on sin/cos.
 *
 * Row layout (16 doubles per epoch):
 *  [0] dT [1] X [2] Y [3] Z [4] obs RA [5] obs Dec [6] err_ra [7] err_dec
 *  [8] err_random_ra [9] err_random_dec [10] err_sys_ra [11] err_sys_dec
 *  [12] b_AU [13] cos(theta) [14] sin(theta) [15] unused
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#define ORC_ROW 16
#define SENT (-999.0)
#define MAS2RAD 4.84813681109536e-09
#define DEG2RAD 0.017453292519943295
#define MASYR2RADS (4.84813681109536e-09 / (365.25 * 86400.0))

typedef struct {
    int64_t n_epochs;
    const double* rows;
    int32_t col[9];      /* dec efac efad incl mu_a mu_d om_asc px ra; -1 = absent (-999) */
    int32_t efac_on;     /* the efac column exists and errs_random / errs_sys are given */
    int32_t binary;      /* dict_timing != {} */
    double cos_decj;
    double a1_lts;
} OrcSource;

/* numpy's pairwise summation (numpy/core/src/umath/loops_utils.h.src, pairwise_sum_DOUBLE) */
static double pairwise_sum(const double* a, int64_t n) {
    if (n < 8) {
        double res = 0.;
        for (int64_t i = 0; i < n; i++) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8];
        int64_t i;
        for (i = 0; i < 8; i++) r[i] = a[i];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int k = 0; k < 8; k++) r[k] += a[i + k];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    } else {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        return pairwise_sum(a, n2) + pairwise_sum(a + n2, n - n2);
    }
}

static inline double par(const double* th, int32_t c) { return c < 0 ? SENT : th[c]; }

static double loglike_one(const OrcSource* src, int S, const double* th, int n_a1, const int32_t* a1_src,
                          const double* a1_mu, const double* a1_sig, int n_sini, const int32_t* sini_col,
                          const double* sini_mu, const double* sini_sig, double* z2, double* lg) {
    double log_p = 0.0;
    for (int s = 0; s < S; s++) {
        const OrcSource* q = &src[s];
        const int64_t E = q->n_epochs;
        const double dec = par(th, q->col[0]), efac = par(th, q->col[1]), efad = par(th, q->col[2]);
        const double incl = par(th, q->col[3]), mu_a = par(th, q->col[4]), mu_d = par(th, q->col[5]);
        const double om_asc = par(th, q->col[6]), px = par(th, q->col[7]), ra = par(th, q->col[8]);
        const double cd = cos(dec), sd = sin(dec), ca = cos(ra), sa = sin(ra);
        const int px_on = px != SENT;
        const int rf_on = px_on && q->binary && incl != SENT && om_asc != SENT;
        double sO = 0, cO = 0, ci = 0;
        if (rf_on) { const double Om = om_asc * DEG2RAD; sO = sin(Om); cO = cos(Om); ci = cos(incl); }
        const int ef = q->efac_on && efac != SENT;
        for (int64_t e = 0; e < E; e++) {
            const double* r = q->rows + e * ORC_ROW;
            double oRA = r[0] * mu_a / cd;                                         /* positions.py:110 */
            double oDEC = r[0] * mu_d;                                             /* :111 */
            if (px_on) {
                const double pRA = px * (r[1] * sa - r[2] * ca) / cd;              /* :308 */
                const double pDEC = px * (r[1] * ca * sd + r[2] * sa * sd - r[3] * cd);   /* :309 */
                oRA = oRA + pRA;
                oDEC = oDEC + pDEC;
                if (rf_on) {
                    const double offset = r[12] * px;                              /* reflex_motion.py:201 */
                    const double m0 = offset * r[13], m1 = offset * r[14];
                    const double b0 = sO * m0 + (cO * ci) * m1;                    /* :202-211 */
                    const double b1 = cO * m0 + (-(sO * ci)) * m1;
                    oRA = oRA + b0 / q->cos_decj;                                  /* :212 */
                    oDEC = oDEC + b1;
                }
            }
            const double m_ra = ra + oRA * MAS2RAD, m_dec = dec + oDEC * MAS2RAD;   /* positions.py:119-120 */
            const double res_ra = r[4] - m_ra, res_dec = r[5] - m_dec;              /* simulate.py:368 */
            double e_ra, e_dec;
            if (ef) {                                                              /* simulate.py:311-320 */
                const double f_dec = (efad != SENT) ? efac * (efad * 1.0) : efac * 1.0;
                const double t_ra = efac * 1.0 * r[10], t_dec = f_dec * r[11];
                e_ra = sqrt(r[8] * r[8] + t_ra * t_ra);
                e_dec = sqrt(r[9] * r[9] + t_dec * t_dec);
            } else {
                e_ra = sqrt(r[6] * r[6]);
                e_dec = sqrt(r[7] * r[7]);
            }
            const double zr = res_ra / e_ra, zd = res_dec / e_dec;
            z2[e] = zr * zr;
            z2[E + e] = zd * zd;
            lg[e] = log(e_ra);
            lg[E + e] = log(e_dec);
        }
        log_p += -0.5 * pairwise_sum(z2, 2 * E);                                   /* simulate.py:370 */
        log_p += -1 * pairwise_sum(lg, 2 * E);                                     /* :371 */
    }
    if (n_a1 > 0) {                                                                /* simulate.py:373-377 */
        double acc[64];
        for (int k = 0; k < n_a1 && k < 64; k++) {
            const OrcSource* q = &src[a1_src[k]];
            const double inc = par(th, q->col[3]), mu_a = par(th, q->col[4]), mu_d = par(th, q->col[5]);
            const double om = par(th, q->col[6]);
            const double mu = sqrt(mu_a * mu_a + mu_d * mu_d);
            const double tm = atan2(mu_a, mu_d);
            double v = mu * sin(tm - om * DEG2RAD) / tan(inc);                     /* kopeikin_effects.py:43 */
            v = v * q->a1_lts;
            v = 1e0 * (v * MASYR2RADS);
            const double z = (v - a1_mu[k]) / a1_sig[k];
            acc[k] = z * z;
        }
        log_p += -0.5 * pairwise_sum(acc, n_a1 < 64 ? n_a1 : 64);
    }
    for (int k = 0; k < n_sini; k++) {                                             /* simulate.py:379-383 */
        const double res = sini_mu[k] - sin(th[sini_col[k]]);
        const double z = res / sini_sig[k];
        log_p += -0.5 * (z * z);
    }
    return log_p;
}

/* theta: (N, P) row-major.  Returns 0, or -1 when scratch memory cannot be allocated. */
int orc_loglike_batch(const OrcSource* src, int S, const double* theta, int64_t N, int64_t P, int n_a1,
                      const int32_t* a1_src, const double* a1_mu, const double* a1_sig, int n_sini,
                      const int32_t* sini_col, const double* sini_mu, const double* sini_sig, double* out,
                      int nthreads) {
    int64_t emax = 1;
    for (int s = 0; s < S; s++) if (src[s].n_epochs > emax) emax = src[s].n_epochs;
    int failed = 0;
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
    {
        double* z2 = (double*)malloc(sizeof(double) * 2 * emax);
        double* lg = (double*)malloc(sizeof(double) * 2 * emax);
        if (!z2 || !lg) {
#pragma omp atomic write
            failed = 1;
        } else {
#pragma omp for schedule(static)
            for (int64_t n = 0; n < N; n++)
                out[n] = loglike_one(src, S, theta + n * P, n_a1, a1_src, a1_mu, a1_sig, n_sini, sini_col, sini_mu,
                                     sini_sig, z2, lg);
        }
        free(z2);
        free(lg);
    }
    return failed ? -1 : 0;
}
