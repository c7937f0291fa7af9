/*
 * nf_priors.h -- packed, pointer-free description of a reference
 * `PriorTransformer` (nestfit/core/core.pyx:437-483) so the unit-cube ->
 * physical-parameter map can be evaluated on the device (and by the CPU
 * oracle) without Python objects.
 *
 * A plan is
 *   - `n_dist` distributions (core.pyx:23-45 `Distribution`): for each one a
 *     `nf_dist_desc` plus NF_DIST_TABLES consecutive double tables of `stride`
 *     entries (xax, pdf, cdf, ppf, S0, S1, S2; stride = size + 1, the extra slot
 *     repeats the last value so the reference's one-past-the-end read at u == 1
 *     is defined).  S_m[i] = sum_{k<=i} k^m (pdf[k]+pdf[k-1])/2 are prefix moments
 *     of the trapezoid terms: they give the interval CDF that
 *     `cdf_over_interval` (core.pyx:109-161) rebuilds per call in closed form;
 *   - `n_prior` prior records executed in order (core.pyx:474-476).  Records
 *     flagged NF_PRIOR_NESTED are not executed at top level; they are the
 *     "sigma prior" objects the Resolved* priors call first
 *     (core.pyx:353,406).
 */
#ifndef NF_PRIORS_H
#define NF_PRIORS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    NF_PRIOR_PLAIN = 0,             /* Prior.interp                core.pyx:192-197 */
    NF_PRIOR_CONSTANT = 1,          /* ConstantPrior.interp        core.pyx:233-238 */
    NF_PRIOR_DUPLICATE = 2,         /* DuplicatePrior.interp       core.pyx:212-221 */
    NF_PRIOR_ORDERED = 3,           /* OrderedPrior.interp         core.pyx:242-258 */
    NF_PRIOR_SPACED = 4,            /* SpacedPrior.interp          core.pyx:280-292 */
    NF_PRIOR_CENSEP = 5,            /* CenSepPrior.interp          core.pyx:305-318 */
    NF_PRIOR_RESOLVED_CENSEP = 6,   /* ResolvedCenSepPrior.interp  core.pyx:347-366 */
    NF_PRIOR_RESOLVED_PLACEMENT = 7 /* ResolvedPlacementPrior      core.pyx:392-434 */
};

#define NF_PRIOR_NESTED 1u
#define NF_DIST_TABLES 7
#define NF_PRIOR_MAX_COMP 10 /* core.pyx:398-400: n > 10 is a silent no-op */

typedef struct nf_dist_desc {
    int32_t size;    /* number of samples of the distribution              */
    int32_t stride;  /* entries per table (= size + 1)                      */
    int32_t offset;  /* index of xax[0] in the tables array (doubles)       */
    int32_t pad_;
    double xmin, xmax, dx, du; /* core.pyx:28-44                            */
} nf_dist_desc;

typedef struct nf_prior_desc {
    int32_t kind;    /* NF_PRIOR_*                                          */
    uint32_t flags;  /* NF_PRIOR_NESTED                                     */
    int32_t p_ix;    /* model-parameter row written (slot p_ix*ncomp + i)   */
    int32_t p_ix2;   /* DUPLICATE: second row; RESOLVED_*: sigma row        */
    int32_t dist;    /* main distribution (centre / independent), or -1     */
    int32_t dist2;   /* dependent / separation distribution, or -1          */
    int32_t nested;  /* RESOLVED_*: index of the nested sigma prior record  */
    int32_t pad_;
    double value;    /* CONSTANT: value; RESOLVED_*: FWHM*scale             */
} nf_prior_desc;

#ifdef __cplusplus
}
#endif
#endif /* NF_PRIORS_H */
