/*
 * nf_oracle.c -- TEST INFRASTRUCTURE, NOT PART OF THE PRODUCT PATH.
 *
 * Plain-C float64 restatement of the reference's likelihood hot path, used only
 * by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg as the
 * checker for the CUDA kernels.  Pinned against the compiled reference itself
 * (oracle/_ref, built from /root/reference by oracle/build_ref.py) in
 * tests/test_oracle.py and against the committed fixtures in tests/golden/.
 *
 * Every function cites the reference lines it restates (paths relative to
 * /root/reference).  The reference is compiled with __APPROX = True and
 * __NEW_CONST = True (includes/model_includes.pxi:20-22).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/nf_nh3_tables.h"
#include "../include/nf_n2hp_tables.h"
#include "../include/nf_priors.h"

/* includes/model_includes.pxi:27-36 */
#define NFO_CKMS 299792.458
#define NFO_CCMS 29979245800.0
#define NFO_H 6.62607015e-27
#define NFO_KB 1.380649e-16
#define NFO_TCMB 2.72548
/* nestfit/models/ammonia.pyx:15-18 (Coudert & Roueff 2006) */
#define NFO_BROT 298192.92e6
#define NFO_CROT 186695.86e6
/* nestfit/core/core.pyx:20 */
#define NFO_FWHM 2.3548200450309493

static const double NH3_NU[NF_NH3_NTRANS] = NF_NH3_REST_FREQ_INIT;
static const double NH3_EA[NF_NH3_NTRANS] = NF_NH3_EINSTEIN_A_INIT;
static const int NH3_OFF[NF_NH3_NTRANS + 1] = NF_NH3_LINE_OFFSET_INIT;
static const int NH3_PARA[NF_NH3_NTRANS] = NF_NH3_IS_PARA_INIT;
static const double NH3_VOFF[NF_NH3_NLINES_TOTAL] = NF_NH3_LINE_VOFF_INIT;
static const double NH3_WT[NF_NH3_NLINES_TOTAL] = NF_NH3_LINE_WEIGHT_INIT;
/* nestfit/models/diazenylium.pyx:31-92 */
static const double N2HP_NU[NF_N2HP_NTRANS] = NF_N2HP_REST_FREQ_INIT;
static const int N2HP_OFF[NF_N2HP_NTRANS + 1] = NF_N2HP_LINE_OFFSET_INIT;
static const double N2HP_VOFF[NF_N2HP_NLINES_TOTAL] = NF_N2HP_LINE_VOFF_INIT;
static const double N2HP_WT[NF_N2HP_NLINES_TOTAL] = NF_N2HP_LINE_WEIGHT_INIT;

/* ------------------------------------------------------------------------ */
/* exp(-x) with the semantics of LIME's FastExp(const float)                 */
/* nestfit/core/fastexp.c:234-283, declared `double FastExp(const float)`    */
/* (core/math.pxd:17): the argument is rounded to float first; x<0 -> libm;   */
/* x==0 -> 1; x < 2^-5 -> 3rd-order nested Taylor; x >= 32 (and NaN, whose    */
/* exponent field is >= the table range) -> 0; otherwise a product of three   */
/* table entries that equals exp(-x) to 4e-16 for the float-rounded x          */
/* (SURVEY.md Appendix C).                                                    */
double nfo_fast_expn(double x)
{
    float xf = (float)x;
    double xd = (double)xf, r;
    int i;
    if (xf < 0.0f) return exp(-xd);
    if (xf == 0.0f) return 1.0;
    if (xf < 0.03125f) {
        r = 1.0;
        for (i = 3; i > 0; i--) r = 1.0 - xd * r * (1.0 / (double)i);
        return r;
    }
    if (!(xf < 32.0f)) return 0.0;
    return exp(-xd);
}

/* ------------------------------------------------------------------------ */
/* 1/(exp(x)-1) through the 1000-point linear table                          */
/* nestfit/models/hyperfine.pyx:12-45                                        */
#define T0_SIZE 1000
static double T0_X[T0_SIZE], T0_Y[T0_SIZE], T0_INV_DX, T0_XMIN, T0_XMAX;
static int t0_ready = 0;

static void t0_init(void)
{
    int i;
    double lo = NFO_H * 23.0e9 / NFO_KB, hi = NFO_H * 28.0e9 / NFO_KB, step;
    T0_XMIN = lo / 8.0;
    T0_XMAX = hi / 2.7;
    step = (T0_XMAX - T0_XMIN) / (double)(T0_SIZE - 1); /* np.linspace */
    for (i = 0; i < T0_SIZE; i++) T0_X[i] = T0_XMIN + (double)i * step;
    T0_X[T0_SIZE - 1] = T0_XMAX;
    for (i = 0; i < T0_SIZE; i++) T0_Y[i] = 1.0 / (exp(T0_X[i]) - 1.0);
    T0_INV_DX = 1.0 / (T0_X[1] - T0_X[0]);
    t0_ready = 1;
}

double nfo_iemtex_interp(double x)
{
    long i_lo;
    double slope;
    if (!t0_ready) t0_init();
    if (T0_XMIN < x && x < T0_XMAX) {
        i_lo = (long)((x - T0_XMIN) * T0_INV_DX);
        if (i_lo > T0_SIZE - 2) i_lo = T0_SIZE - 2; /* guard the one-past read */
        slope = (T0_Y[i_lo + 1] - T0_Y[i_lo]) * T0_INV_DX;
        return slope * (x - T0_X[i_lo]) + T0_Y[i_lo];
    }
    return 1.0 / expm1(x);
}

/* nestfit/models/ammonia.pyx:280-286 */
double nfo_swift_convert(double tkin)
{
    return tkin / (1.0 + (tkin / 41.18) * log(1.0 + 0.6 * exp(-15.7 / tkin)));
}

/* nestfit/models/ammonia.pyx:289-295 */
double nfo_partition_level(long j, double trot)
{
    return (double)(2 * j + 1) * nfo_fast_expn(
        NFO_H * (NFO_BROT * (double)(j * (j + 1)) + (NFO_CROT - NFO_BROT) * (double)(j * j))
        / (NFO_KB * trot));
}

/* nestfit/models/ammonia.pyx:304-315, J lists at ammonia.pyx:49-51 */
double nfo_partition_func(int para, double trot)
{
    long j;
    double q = 0.0;
    for (j = 0; j < 51; j++) {
        if (para) { if (j % 3 != 0) q += nfo_partition_level(j, trot); }
        else      { if (j % 3 == 0) q += 2.0 * nfo_partition_level(j, trot); }
    }
    return q;
}

/* Background term of one channel: nestfit/models/ammonia.pyx:273-277 */
void nfo_tbg(const double *xarr, long n, double *tbg)
{
    long i;
    for (i = 0; i < n; i++) tbg[i] = 1.0 / expm1(NFO_H * xarr[i] / NFO_KB / NFO_TCMB);
}

/* ------------------------------------------------------------------------ */
/* One hyperfine slab: nestfit/models/hyperfine.pyx:52-113.                  */
/* `tarr` is scratch [n]; `pred` is accumulated into.  counters[0] += number  */
/* of windowed Gaussian evaluations, counters[1] += radiative-transfer        */
/* channels (SURVEY.md 8d work accounting); may be NULL.                      */
static void hf_predict_lines(const double *xarr, const double *tbg, long n, double nu0,
                             const double *line_voff, const double *line_wt, long nlines,
                             double voff, double tex, double ltau_main, double sigm,
                             double *tarr, double *pred, int64_t *counters)
{
    const double nu_min = xarr[0], nu_chan = xarr[1] - xarr[0]; /* core.pyx:503,513 */
    double tau_main = pow(10.0, ltau_main);                      /* hyperfine.pyx:63 */
    long i, j, lo, hi;
    for (j = 0; j < n; j++) tarr[j] = 0.0;
    for (i = 0; i < nlines; i++) {
        double hf_freq = (1.0 - line_voff[i] / NFO_CKMS) * nu0;
        double hf_width = sigm / NFO_CKMS * hf_freq;
        double hf_offset = voff / NFO_CKMS * hf_freq;
        double hf_nucen = hf_freq - hf_offset;
        double hf_tau = tau_main * line_wt[i];
        double hf_idenom = 0.5 / (hf_width * hf_width);
        double nu_cutoff = sqrt(12.5 / hf_idenom);               /* hyperfine.pyx:82 */
        double nu_lo = hf_nucen - nu_min - nu_cutoff;
        double nu_hi = hf_nucen - nu_min + nu_cutoff;
        lo = (long)floor(nu_lo / nu_chan);
        hi = (long)floor(nu_hi / nu_chan);
        if (hi < 0 || lo > n - 1) continue;                      /* hyperfine.pyx:88 */
        if (lo < 0) lo = 0;
        if (hi > n - 1) hi = n - 1;
        for (j = lo; j < hi; j++) {                              /* upper edge excluded */
            double nu = xarr[j] - hf_nucen;
            tarr[j] += hf_tau * nfo_fast_expn(nu * nu * hf_idenom);
        }
        if (counters && hi > lo) counters[0] += hi - lo;
    }
    for (j = 0; j < n; j++) {                                    /* hyperfine.pyx:103-113 */
        double T0;
        if (tarr[j] == 0.0) continue;
        T0 = NFO_H * xarr[j] / NFO_KB;
        pred[j] += T0 * (nfo_iemtex_interp(T0 / tex) - tbg[j])
                 * (1.0 - nfo_fast_expn(tarr[j]));
        if (counters) counters[1] += 1;
    }
}

static void hf_predict(const double *xarr, const double *tbg, long n, int t,
                       double voff, double tex, double ltau_main, double sigm,
                       double *tarr, double *pred, int64_t *counters)
{
    hf_predict_lines(xarr, tbg, n, NH3_NU[t], NH3_VOFF + NH3_OFF[t], NH3_WT + NH3_OFF[t],
                     NH3_OFF[t + 1] - NH3_OFF[t], voff, tex, ltau_main, sigm, tarr, pred, counters);
}

/* nestfit/models/diazenylium.pyx:140-154: params (voff, tex, ltau, sigm), parameter-major;
 * trans_id 1..3 = J 1-0, 2-1, 3-2. */
void nfo_nnhp_predict(const double *xarr, const double *tbg, long n, int trans_id,
                      const double *params, long ncomp, double *tarr, double *pred,
                      int64_t *counters)
{
    const int t = trans_id - 1;
    long c, j;
    for (j = 0; j < n; j++) pred[j] = 0.0;
    for (c = 0; c < ncomp; c++)
        hf_predict_lines(xarr, tbg, n, N2HP_NU[t], N2HP_VOFF + N2HP_OFF[t], N2HP_WT + N2HP_OFF[t],
                         N2HP_OFF[t + 1] - N2HP_OFF[t], params[c], params[ncomp + c],
                         params[2 * ncomp + c], params[3 * ncomp + c], tarr, pred, counters);
}

/* nestfit/models/ammonia.pyx:326-361.  trans_id is 1-based ((1,1) -> 1). */
void nfo_amm_predict(const double *xarr, const double *tbg, long n, int trans_id,
                     const double *params, long ncomp, int cold, int lte,
                     double *tarr, double *pred, int64_t *counters)
{
    const int t = trans_id - 1;
    const double nu0 = NH3_NU[t];
    long c, j;
    for (j = 0; j < n; j++) pred[j] = 0.0;
    for (c = 0; c < ncomp; c++) {
        double voff = params[c], trot = params[ncomp + c], tex = params[2 * ncomp + c];
        double ntot = params[3 * ncomp + c], sigm = params[4 * ncomp + c];
        double orth = params[5 * ncomp + c];
        double zlev, qtot, frac, pop, e, expterm, fracterm, widthterm, tau_main;
        if (cold) trot = nfo_swift_convert(trot);
        if (lte) tex = trot;
        zlev = nfo_partition_level(t + 1, trot);
        qtot = nfo_partition_func(NH3_PARA[t], trot);
        frac = NH3_PARA[t] ? 1.0 - orth : orth;
        pop = pow(10.0, ntot) * frac * zlev / qtot;
        e = exp(-NFO_H * nu0 / (NFO_KB * tex));
        expterm = (1.0 - e) / (1.0 + e);
        fracterm = NFO_CCMS * NFO_CCMS * NH3_EA[t] / (8.0 * M_PI * nu0 * nu0);
        widthterm = NFO_CKMS / (sigm * nu0 * sqrt(2.0 * M_PI));
        tau_main = pop * fracterm * expterm * widthterm;
        hf_predict(xarr, tbg, n, t, voff, tex, log10(tau_main), sigm, tarr, pred, counters);
    }
}

/* nestfit/models/gaussian.pyx:17-50 (window indices are C int there). */
void nfo_gauss_predict(const double *xarr, long n, double rest_freq,
                       const double *params, long ncomp, double *pred,
                       int64_t *counters)
{
    const double nu_min = xarr[0], nu_chan = xarr[1] - xarr[0];
    long c, j;
    for (j = 0; j < n; j++) pred[j] = 0.0;
    for (c = 0; c < ncomp; c++) {
        double voff = params[c], sigm = params[ncomp + c], peak = params[2 * ncomp + c];
        double nu_width = sigm / NFO_CKMS * rest_freq;
        double nu_cen = rest_freq * (1.0 - voff / NFO_CKMS);
        double nu_denom = 0.5 / (nu_width * nu_width);
        double nu_cutoff = sqrt(12.5 / nu_denom);
        int lo = (int)floor((nu_cen - nu_min - nu_cutoff) / nu_chan);
        int hi = (int)floor((nu_cen - nu_min + nu_cutoff) / nu_chan);
        if (hi < 0 || lo > n - 1) continue;
        if (lo < 0) lo = 0;
        if (hi > n - 1) hi = (int)n - 1;
        for (j = lo; j < hi; j++) {
            double nu = xarr[j] - nu_cen;
            pred[j] += peak * nfo_fast_expn(nu * nu * nu_denom);
        }
        if (counters && hi > lo) counters[0] += hi - lo;
    }
}

/* nestfit/core/core.pyx:522-530: no Gaussian normalisation prefactor. */
double nfo_loglike(const double *data, const double *pred, long n, double noise)
{
    long j;
    double s = 0.0;
    for (j = 0; j < n; j++) { double d = data[j] - pred[j]; s += d * d; }
    return -s / (2.0 * noise * noise);
}

/* ------------------------------------------------------------------------ */
/* Batched drivers (the unit of work of AmmoniaRunner.c_loglikelihood,        */
/* nestfit/models/ammonia.pyx:423-432, minus the prior transform): vector b    */
/* is scored against pixel pix_of_vec[b].                                      */
/*   xarr  [nspec][nchan]   data [npix][nspec][nchan]   noise [npix][nspec]    */
/*   params[B][6*ncomp] (parameter-major, component-minor)                     */
/*   pred_out may be NULL, else [B][nspec][nchan]; counters NULL or int64[2].  */
int nfo_nh3_loglike_batch(long nspec, long nchan, const double *xarr,
                          const int *trans_id, const double *data,
                          const double *noise, const double *params,
                          const int *pix_of_vec, long B, long ncomp, int cold,
                          int lte, double *lnL, double *pred_out,
                          int64_t *counters)
{
    double *tbg = (double *)malloc(sizeof(double) * nspec * nchan);
    double *tarr = (double *)malloc(sizeof(double) * nchan);
    double *pred = (double *)malloc(sizeof(double) * nchan);
    long b, s;
    if (!tbg || !tarr || !pred) { free(tbg); free(tarr); free(pred); return -1; }
    for (s = 0; s < nspec; s++) nfo_tbg(xarr + s * nchan, nchan, tbg + s * nchan);
    for (b = 0; b < B; b++) {
        long p = pix_of_vec ? pix_of_vec[b] : 0;
        double acc = 0.0;
        for (s = 0; s < nspec; s++) {
            nfo_amm_predict(xarr + s * nchan, tbg + s * nchan, nchan, trans_id[s],
                            params + b * 6 * ncomp, ncomp, cold, lte, tarr, pred, counters);
            if (data)
                acc += nfo_loglike(data + (p * nspec + s) * nchan, pred, nchan,
                                   noise[p * nspec + s]);
            if (pred_out)
                memcpy(pred_out + (b * nspec + s) * nchan, pred, sizeof(double) * nchan);
        }
        if (lnL) lnL[b] = acc;
    }
    free(tbg); free(tarr); free(pred);
    return 0;
}

/* DiazenyliumRunner.c_loglikelihood (diazenylium.pyx:206-215) minus the transform:
 * params [B][4*ncomp]. */
int nfo_n2hp_loglike_batch(long nspec, long nchan, const double *xarr,
                           const int *trans_id, const double *data,
                           const double *noise, const double *params,
                           const int *pix_of_vec, long B, long ncomp,
                           double *lnL, double *pred_out, int64_t *counters)
{
    double *tbg = (double *)malloc(sizeof(double) * nspec * nchan);
    double *tarr = (double *)malloc(sizeof(double) * nchan);
    double *pred = (double *)malloc(sizeof(double) * nchan);
    long b, s;
    if (!tbg || !tarr || !pred) { free(tbg); free(tarr); free(pred); return -1; }
    for (s = 0; s < nspec; s++) nfo_tbg(xarr + s * nchan, nchan, tbg + s * nchan);
    for (b = 0; b < B; b++) {
        long p = pix_of_vec ? pix_of_vec[b] : 0;
        double acc = 0.0;
        for (s = 0; s < nspec; s++) {
            nfo_nnhp_predict(xarr + s * nchan, tbg + s * nchan, nchan, trans_id[s],
                             params + b * 4 * ncomp, ncomp, tarr, pred, counters);
            if (data)
                acc += nfo_loglike(data + (p * nspec + s) * nchan, pred, nchan,
                                   noise[p * nspec + s]);
            if (pred_out)
                memcpy(pred_out + (b * nspec + s) * nchan, pred, sizeof(double) * nchan);
        }
        if (lnL) lnL[b] = acc;
    }
    free(tbg); free(tarr); free(pred);
    return 0;
}

/* GaussianRunner.c_loglikelihood (gaussian.pyx:98-102) minus the transform. */
int nfo_gauss_loglike_batch(long nchan, const double *xarr, double rest_freq,
                            const double *data, const double *noise,
                            const double *params, const int *pix_of_vec, long B,
                            long ncomp, double *lnL, double *pred_out,
                            int64_t *counters)
{
    double *pred = (double *)malloc(sizeof(double) * nchan);
    long b;
    if (!pred) return -1;
    for (b = 0; b < B; b++) {
        long p = pix_of_vec ? pix_of_vec[b] : 0;
        nfo_gauss_predict(xarr, nchan, rest_freq, params + b * 3 * ncomp, ncomp, pred, counters);
        if (lnL && data) lnL[b] = nfo_loglike(data + p * nchan, pred, nchan, noise[p]);
        if (pred_out) memcpy(pred_out + b * nchan, pred, sizeof(double) * nchan);
    }
    free(pred);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* Prior transform on a packed plan (include/nf_priors.h).                    */

typedef struct {
    const nf_dist_desc *d;
    const double *xax, *pdf, *ppf;
    double *cdf; /* private, mutable copy: cdf_over_interval rewrites it */
} dist_view;

/* Distribution.ppf_interp, core.pyx:47-63 */
static double ppf_interp(const dist_view *v, double u)
{
    long i_lo = (long)((double)(v->d->size - 1) * u);
    double x_lo = (double)i_lo * v->d->du;
    double y_lo = v->ppf[i_lo], y_hi = v->ppf[i_lo + 1];
    double slope = (y_hi - y_lo) / v->d->du;
    return slope * (u - x_lo) + y_lo;
}

/* Distribution.cdf_interp, core.pyx:65-107 */
static double cdf_interp(const dist_view *v, double u)
{
    long size = v->d->size, i_lo = 0, i_hi = size, i = size / 2;
    double slope;
    if (u <= v->cdf[0]) u = 1e-64;
    while (i != i_lo) {
        if (u > v->cdf[i]) i_lo = i; else i_hi = i;
        i = (i_hi + i_lo) / 2;
    }
    i_lo = i < size ? i : size - 1;
    slope = (v->cdf[i_lo + 1] - v->cdf[i_lo]) / v->d->dx;
    return 1.0 / slope * (u - v->cdf[i_lo]) + v->xax[i_lo];
}

static long cast_long(double r)
{
    if (!(r == r) || r >= 9.0e18 || r <= -9.0e18) return (-9223372036854775807L - 1L);
    return (long)r;
}

/* Distribution.cdf_over_interval, core.pyx:109-161 */
static void cdf_over_interval(dist_view *v, double x_lo, double x_hi, double sfact)
{
    long size = v->d->size, i, i_lo, i_hi;
    double csum = 0.0, inv_delta_i, scale;
    if (x_lo > x_hi) { double t = x_lo; x_lo = x_hi; x_hi = t; }
    /* (long)NaN is undefined in C; the reference binary (x86-64 cvttsd2si) yields
     * LONG_MIN, made explicit here.  NaN bounds arise after a degenerate draw. */
    i_lo = cast_long((x_lo - v->d->xmin) / v->d->dx);
    if (i_lo >= size) i_lo = size - 1; else if (i_lo < 0) i_lo = 0;
    i_hi = cast_long((x_hi - v->d->xmin) / v->d->dx);
    if (i_hi == i_lo) i_hi = i_lo + 1;
    if (i_hi > size) i_hi = size; else if (i_hi < 0) i_hi = 1;
    for (i = 0; i < i_lo; i++) v->cdf[i] = 0.0;
    for (i = i_hi; i < size; i++) v->cdf[i] = 1.0;
    if (i_hi - i_lo == 1) {
        v->cdf[i_lo] = 1.0;
    } else {
        v->cdf[i_lo] = 0.0;
        inv_delta_i = 1.0 / (double)(i_hi - i_lo);
        for (i = i_lo + 1; i < i_hi; i++) {
            if (sfact == 0.0) scale = 1.0;
            else if (sfact == 1.0) scale = 1.0 - (double)(i - i_lo) * inv_delta_i;
            else if (sfact == 2.0) { scale = 1.0 - (double)(i - i_lo) * inv_delta_i; scale *= scale; }
            else scale = pow(1.0 - (double)(i - i_lo) * inv_delta_i, sfact);
            csum += 0.5 * (v->pdf[i] + v->pdf[i - 1]) * scale;
            v->cdf[i] = csum;
        }
    }
    for (i = i_lo; i < i_hi; i++) v->cdf[i] /= csum;
}

static void bind(dist_view *v, const nf_dist_desc *dd, int ix, const double *tables, double *cdf_scratch)
{
    const nf_dist_desc *d = dd + ix;
    v->d = d;
    v->xax = tables + d->offset;
    v->pdf = v->xax + d->stride;
    v->ppf = v->xax + 3 * (long)d->stride;
    v->cdf = cdf_scratch;
    memcpy(cdf_scratch, v->xax + 2 * (long)d->stride, sizeof(double) * d->stride);
}

static void prior_apply(const nf_prior_desc *pr, int k, const nf_dist_desc *dd,
                        const double *tables, double *u, long n, double **scratch)
{
    const nf_prior_desc *p = pr + k;
    dist_view a, b;
    long i, ix = (long)p->p_ix * n;
    double v;
    if (p->dist >= 0) bind(&a, dd, p->dist, tables, scratch[0]);
    if (p->dist2 >= 0) bind(&b, dd, p->dist2, tables, scratch[1]);
    switch (p->kind) {
    case NF_PRIOR_PLAIN:                                   /* core.pyx:192-197 */
        for (i = 0; i < n; i++) u[ix + i] = ppf_interp(&a, u[ix + i]);
        break;
    case NF_PRIOR_CONSTANT:                                /* core.pyx:233-238 */
        for (i = 0; i < n; i++) u[ix + i] = p->value;
        break;
    case NF_PRIOR_DUPLICATE:                               /* core.pyx:212-221 */
        for (i = 0; i < n; i++) {
            v = ppf_interp(&a, u[ix + i]);
            u[ix + i] = v;
            u[(long)p->p_ix2 * n + i] = v;
        }
        break;
    case NF_PRIOR_ORDERED: {                               /* core.pyx:242-258 */
        double umin = 0.0, uu;
        for (i = 0; i < n; i++) {
            uu = umin + (1.0 - umin) * u[ix + i];
            umin = uu;
            u[ix + i] = ppf_interp(&a, uu);
        }
        break; }
    case NF_PRIOR_SPACED:                                  /* core.pyx:280-292 */
        v = ppf_interp(&a, u[ix]);
        u[ix] = v;
        for (i = 1; i < n; i++) { v = v + ppf_interp(&b, u[ix + i]); u[ix + i] = v; }
        break;
    case NF_PRIOR_CENSEP: {                                /* core.pyx:305-318 */
        double vcen = ppf_interp(&a, u[ix]), vsep;
        if (n == 1) u[ix] = vcen;
        else if (n == 2) {
            vsep = ppf_interp(&b, u[ix + 1]);
            u[ix] = vcen - 0.5 * vsep; u[ix + 1] = vcen + 0.5 * vsep;
        }
        break; }
    case NF_PRIOR_RESOLVED_CENSEP: {                       /* core.pyx:347-366 */
        long ix_s = (long)p->p_ix2 * n;
        double vcen, vsep, min_sep;
        prior_apply(pr, p->nested, dd, tables, u, n, scratch + 2);
        vcen = ppf_interp(&a, u[ix]);
        if (n == 1) u[ix] = vcen;
        else if (n == 2) {
            vsep = ppf_interp(&b, u[ix + 1]);
            min_sep = p->value * sqrt(u[ix_s] * u[ix_s + 1]);
            if (min_sep > vsep) vsep = min_sep;
            u[ix] = vcen - 0.5 * vsep; u[ix + 1] = vcen + 0.5 * vsep;
        }
        break; }
    case NF_PRIOR_RESOLVED_PLACEMENT: {                    /* core.pyx:392-434 */
        long ix_s = (long)p->p_ix2 * n;
        double v_lo, v_hi, sep, sep_tot, f, min_seps[NF_PRIOR_MAX_COMP];
        if (n > NF_PRIOR_MAX_COMP) return;
        v_lo = a.d->xmin; v_hi = a.d->xmax;
        prior_apply(pr, p->nested, dd, tables, u, n, scratch + 2);
        if (n == 1) { u[ix] = ppf_interp(&a, u[ix]); return; }
        sep_tot = 0.0; min_seps[0] = 0.0;
        for (i = 1; i < n; i++) {
            sep = p->value * sqrt(u[ix_s + i] * u[ix_s + i - 1]);
            sep_tot += sep; min_seps[i] = sep;
        }
        if (sep_tot > v_hi - v_lo) {
            f = (v_hi - v_lo) / sep_tot; sep_tot = 0.0;
            for (i = 0; i < n; i++) { min_seps[i] *= f; sep_tot += min_seps[i]; }
        }
        v_hi -= sep_tot;
        for (i = 0; i < n; i++) {
            sep = min_seps[i];
            v_lo += sep; v_hi += sep;
            cdf_over_interval(&a, v_lo, v_hi, (double)(n - 1 - i));
            v_lo = cdf_interp(&a, u[ix + i]);
            u[ix + i] = v_lo;
        }
        break; }
    default: break;
    }
}

/* PriorTransformer.c_transform, core.pyx:459-476, applied to B vectors in
 * place (u: [B][n_model*ncomp]). */
int nfo_prior_transform(const nf_prior_desc *priors, int n_prior,
                        const nf_dist_desc *dists, int n_dist,
                        const double *tables, double *u, long B, long ndim,
                        long ncomp)
{
    int k, s, max_stride = 1;
    long b;
    double *scratch[4];
    for (k = 0; k < n_dist; k++) if (dists[k].stride > max_stride) max_stride = dists[k].stride;
    for (s = 0; s < 4; s++) {
        scratch[s] = (double *)malloc(sizeof(double) * (size_t)max_stride);
        if (!scratch[s]) return -1;
    }
    for (b = 0; b < B; b++)
        for (k = 0; k < n_prior; k++)
            if (!(priors[k].flags & NF_PRIOR_NESTED))
                prior_apply(priors, k, dists, tables, u + b * ndim, ncomp, scratch);
    for (s = 0; s < 4; s++) free(scratch[s]);
    return 0;
}
