// Unit-cube -> physical-parameter prior transform on the device (FP64).
//
// Restates PriorTransformer.c_transform and the Prior* classes of the reference
// (nestfit/core/core.pyx:47-161,192-434,459-476) on a packed plan
// (include/nf_priors.h).  One thread transforms one parameter vector in place.
//
// ResolvedPlacementPrior (core.pyx:392-434) rebuilds a 500-point CDF per
// component in the reference (`cdf_over_interval` mutates shared state).  Here
// the rebuilt CDF is never materialised: its normalisation is accumulated in a
// first sweep and the bracketing interval of `u` is found in a second sweep,
// which selects the same interval as the reference's bisection because the
// rebuilt CDF is monotone.  The degenerate one-bin interval (CDF = 0,inf,1) is
// evaluated by running the reference's bisection on the closed-form values.

#include <cstring>
#include <new>

#include "nf_internal.cuh"

namespace {

struct DistView {
    const double *xax, *pdf, *ppf, *S0, *S1, *S2;
    int size;
    double xmin, xmax, dx, du;
};

__device__ __forceinline__ DistView bind_dist(const nf_dist_desc *dd, const double *tables, int ix)
{
    DistView v;
    const nf_dist_desc d = dd[ix];
    v.xax = tables + d.offset;
    v.pdf = v.xax + d.stride;
    v.ppf = v.xax + 3 * (int64_t)d.stride;
    v.S0 = v.xax + 4 * (int64_t)d.stride;
    v.S1 = v.xax + 5 * (int64_t)d.stride;
    v.S2 = v.xax + 6 * (int64_t)d.stride;
    v.size = d.size; v.xmin = d.xmin; v.xmax = d.xmax; v.dx = d.dx; v.du = d.du;
    return v;
}

// Distribution.ppf_interp, core.pyx:47-63
__device__ double ppf_interp(const DistView &v, double u)
{
    if (!(u >= 0.0 && u <= 1.0)) return nan("");
    int i_lo = (int)((double)(v.size - 1) * u);
    double x_lo = (double)i_lo * v.du;
    double y_lo = v.ppf[i_lo], y_hi = v.ppf[i_lo + 1];
    double slope = (y_hi - y_lo) / v.du;
    return slope * (u - x_lo) + y_lo;
}

__device__ __forceinline__ double pow_int(double x, int s)
{
    double r = 1.0;
    for (int k = 0; k < s; ++k) r *= x;
    return r;
}

// cdf_over_interval(x_lo, x_hi, sfact) followed by cdf_interp(u)
// (core.pyx:109-161 and 65-107) without materialising the CDF.
__device__ double placement_draw(const DistView &v, double x_lo, double x_hi, int sfact, double u)
{
    const int size = v.size;
    if (!(u == u)) return nan("");
    if (x_lo > x_hi) { double t = x_lo; x_lo = x_hi; x_hi = t; }
    double r_lo = (x_lo - v.xmin) / v.dx, r_hi = (x_hi - v.xmin) / v.dx;
    // A NaN bound (a previous component's degenerate draw) converts to the most negative
    // integer in the reference's C cast on x86 and is then clamped; fmax(NaN, x) = x.
    r_lo = fmin(fmax(r_lo, -1.0e9), 1.0e9);
    r_hi = fmin(fmax(r_hi, -1.0e9), 1.0e9);
    int i_lo = (int)r_lo, i_hi = (int)r_hi;          // C truncation
    if (i_lo >= size) i_lo = size - 1; else if (i_lo < 0) i_lo = 0;
    if (i_hi == i_lo) i_hi = i_lo + 1;
    if (i_hi > size) i_hi = size; else if (i_hi < 0) i_hi = 1;

    if (i_hi - i_lo <= 1) {
        // rebuilt CDF is {0 below i_lo, 1/0 = inf at i_lo, 1 above}: bisection on closed form
        auto cdf = [&](int i) -> double {
            if (i < i_lo) return 0.0;
            if (i >= i_hi) return 1.0;
            return __longlong_as_double(0x7ff0000000000000LL);
        };
        if (u <= cdf(0)) u = 1e-64;
        int lo = 0, hi = size, i = size / 2;
        while (i != lo) {
            if (u > cdf(i)) lo = i; else hi = i;
            i = (hi + lo) / 2;
        }
        lo = i < size ? i : size - 1;
        const double c0 = cdf(lo), c1 = (lo + 1 < size) ? cdf(lo + 1) : 1.0;
        const double slope = (c1 - c0) / v.dx;
        return 1.0 / slope * (u - c0) + v.xax[lo];
    }

    if (sfact <= 2) {
        // closed form through prefix moments: numerator of the rebuilt CDF at index i,
        //   P(i) = sum_{k=i_lo+1}^{i} t_k ((i_hi - k) / D)^sfact
        const double D = (double)(i_hi - i_lo), ih = (double)i_hi;
        const double b0 = v.S0[i_lo], b1 = v.S1[i_lo], b2 = v.S2[i_lo];
        auto P = [&](int i) -> double {
            const double d0 = v.S0[i] - b0;
            if (sfact == 0) return d0;
            const double d1 = v.S1[i] - b1;
            if (sfact == 1) return (ih * d0 - d1) / D;
            const double d2 = v.S2[i] - b2;
            return (ih * ih * d0 - 2.0 * ih * d1 + d2) / (D * D);
        };
        const double csum = P(i_hi - 1);
        if (!(csum > 0.0)) return nan("");
        if (u <= 0.0) u = 1e-64;
        const double target = fmin(u, 1.0) * csum;   // u lies in (0, 1]
        int lo = i_lo, hi = i_hi - 1;          // P(lo) = 0 < target <= P(hi) for u <= 1
        double p_lo = 0.0, p_hi = csum;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            const double pm = P(mid);
            if (pm < target) { lo = mid; p_lo = pm; } else { hi = mid; p_hi = pm; }
        }
        const double c_lo = p_lo / csum, c_hi = p_hi / csum;
        const double slope = (c_hi - c_lo) / v.dx;
        return 1.0 / slope * (u - c_lo) + v.xax[lo];
    }
    const double inv_delta = 1.0 / (double)(i_hi - i_lo);
    // sweep 1: normalisation
    double csum = 0.0;
    for (int i = i_lo + 1; i < i_hi; ++i) {
        const double base = 1.0 - (double)(i - i_lo) * inv_delta;
        const double scale = sfact == 0 ? 1.0 : (sfact == 1 ? base : (sfact == 2 ? base * base : pow_int(base, sfact)));
        csum += 0.5 * (v.pdf[i] + v.pdf[i - 1]) * scale;
    }
    if (!(csum > 0.0)) return nan("");
    if (u <= 0.0) u = 1e-64;       // cdf[0] is 0 (or 0/csum) here, core.pyx:81-82
    // sweep 2: last index whose CDF value is < u
    double part = 0.0, prev = 0.0;
    int istar = i_hi - 1;
    double c_lo = 0.0, c_hi = 1.0;
    bool found = false;
    for (int i = i_lo + 1; i < i_hi; ++i) {
        const double base = 1.0 - (double)(i - i_lo) * inv_delta;
        const double scale = sfact == 0 ? 1.0 : (sfact == 1 ? base : (sfact == 2 ? base * base : pow_int(base, sfact)));
        prev = part;
        part += 0.5 * (v.pdf[i] + v.pdf[i - 1]) * scale;
        if (!(part / csum < u)) { istar = i - 1; c_lo = prev / csum; c_hi = part / csum; found = true; break; }
    }
    if (!found) {
        // u above the last in-interval value (== 1): bracket is [i_hi-1, i_hi]
        istar = i_hi - 1;
        c_lo = part / csum;
        c_hi = 1.0;
        if (istar + 1 >= size) { istar = size - 1; c_hi = c_lo; }
    }
    const double slope = (c_hi - c_lo) / v.dx;
    return 1.0 / slope * (u - c_lo) + v.xax[istar];
}

__device__ void prior_apply(const nf_prior_desc *pr, int k, const nf_dist_desc *dd, const double *tables,
                            double *u, int n)
{
    const nf_prior_desc p = pr[k];
    const int ix = p.p_ix * n;
    switch (p.kind) {
    case NF_PRIOR_PLAIN: {                                   // core.pyx:192-197
        const DistView a = bind_dist(dd, tables, p.dist);
        for (int i = 0; i < n; ++i) u[ix + i] = ppf_interp(a, u[ix + i]);
        break; }
    case NF_PRIOR_CONSTANT:                                  // core.pyx:233-238
        for (int i = 0; i < n; ++i) u[ix + i] = p.value;
        break;
    case NF_PRIOR_DUPLICATE: {                               // core.pyx:212-221
        const DistView a = bind_dist(dd, tables, p.dist);
        for (int i = 0; i < n; ++i) {
            const double v = ppf_interp(a, u[ix + i]);
            u[ix + i] = v;
            u[p.p_ix2 * n + i] = v;
        }
        break; }
    case NF_PRIOR_ORDERED: {                                 // core.pyx:242-258
        const DistView a = bind_dist(dd, tables, p.dist);
        double umin = 0.0;
        for (int i = 0; i < n; ++i) {
            const double uu = umin + (1.0 - umin) * u[ix + i];
            umin = uu;
            u[ix + i] = ppf_interp(a, uu);
        }
        break; }
    case NF_PRIOR_SPACED: {                                  // core.pyx:280-292
        const DistView a = bind_dist(dd, tables, p.dist);
        const DistView b = bind_dist(dd, tables, p.dist2);
        double v = ppf_interp(a, u[ix]);
        u[ix] = v;
        for (int i = 1; i < n; ++i) { v = v + ppf_interp(b, u[ix + i]); u[ix + i] = v; }
        break; }
    case NF_PRIOR_CENSEP:                                    // core.pyx:305-318
    case NF_PRIOR_RESOLVED_CENSEP: {                         // core.pyx:347-366
        const DistView a = bind_dist(dd, tables, p.dist);
        const DistView b = bind_dist(dd, tables, p.dist2);
        if (p.kind == NF_PRIOR_RESOLVED_CENSEP) prior_apply(pr, p.nested, dd, tables, u, n);
        const double vcen = ppf_interp(a, u[ix]);
        if (n == 1) u[ix] = vcen;
        else if (n == 2) {
            double vsep = ppf_interp(b, u[ix + 1]);
            if (p.kind == NF_PRIOR_RESOLVED_CENSEP) {
                const int ix_s = p.p_ix2 * n;
                const double min_sep = p.value * sqrt(u[ix_s] * u[ix_s + 1]);
                if (min_sep > vsep) vsep = min_sep;
            }
            u[ix] = vcen - 0.5 * vsep;
            u[ix + 1] = vcen + 0.5 * vsep;
        }
        break; }
    case NF_PRIOR_RESOLVED_PLACEMENT: {                      // core.pyx:392-434
        if (n > NF_PRIOR_MAX_COMP) return;
        const DistView a = bind_dist(dd, tables, p.dist);
        const int ix_s = p.p_ix2 * n;
        double v_lo = a.xmin, v_hi = a.xmax;
        prior_apply(pr, p.nested, dd, tables, u, n);
        if (n == 1) { u[ix] = ppf_interp(a, u[ix]); return; }
        double min_seps[NF_PRIOR_MAX_COMP];
        double sep_tot = 0.0;
        min_seps[0] = 0.0;
        for (int i = 1; i < n; ++i) {
            const double sep = p.value * sqrt(u[ix_s + i] * u[ix_s + i - 1]);
            sep_tot += sep;
            min_seps[i] = sep;
        }
        if (sep_tot > v_hi - v_lo) {
            const double f = (v_hi - v_lo) / sep_tot;
            sep_tot = 0.0;
            for (int i = 0; i < n; ++i) { min_seps[i] *= f; sep_tot += min_seps[i]; }
        }
        v_hi -= sep_tot;
        for (int i = 0; i < n; ++i) {
            const double sep = min_seps[i];
            v_lo += sep;
            v_hi += sep;
            v_lo = placement_draw(a, v_lo, v_hi, n - 1 - i, u[ix + i]);
            u[ix + i] = v_lo;
        }
        break; }
    default: break;
    }
}

__global__ void nf_prior_transform_kernel(const nf_prior_desc *pr, int n_prior, const nf_dist_desc *dd,
                                          const double *tables, double *u, int64_t B, int ndim, int ncomp,
                                          const int32_t *B_dev)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (B_dev) B = min(B, (int64_t)*B_dev);     // device-resident count; B only sized the grid
    if (b >= B) return;
    double *row = u + b * ndim;
    for (int k = 0; k < n_prior; ++k)
        if (!(pr[k].flags & NF_PRIOR_NESTED)) prior_apply(pr, k, dd, tables, row, ncomp);
}

}  // namespace

cudaError_t nf_launch_prior_transform(const nf_priors *pr, double *u, int64_t B, int ncomp, cudaStream_t st,
                                      const int32_t *B_dev)
{
    if (B <= 0) return cudaSuccess;
    const int threads = 128;
    const int64_t grid = (B + threads - 1) / threads;
    nf_prior_transform_kernel<<<(unsigned)grid, threads, 0, st>>>(pr->priors, pr->n_prior, pr->dists, pr->tables,
                                                                  u, B, pr->n_model * ncomp, ncomp, B_dev);
    return cudaGetLastError();
}

extern "C" {

int nf_priors_create(int device, const nf_prior_desc *priors, int n_prior, const nf_dist_desc *dists,
                     int n_dist, const double *tables, int64_t n_tables, int n_model, nf_priors **out)
{
    if (!out) return NF_EINVAL;
    *out = nullptr;
    if (!priors || n_prior < 1 || n_dist < 0 || n_model < 1 || (n_dist > 0 && (!dists || !tables))) return NF_EINVAL;
    // validate the plan before it is trusted on the device
    for (int k = 0; k < n_dist; ++k) {
        const nf_dist_desc &d = dists[k];
        if (d.size < 2 || d.stride < d.size + 1 || d.offset < 0 ||
            (int64_t)d.offset + NF_DIST_TABLES * (int64_t)d.stride > n_tables || !(d.dx > 0.0) || !(d.du > 0.0))
            return NF_EINVAL;
    }
    for (int k = 0; k < n_prior; ++k) {
        const nf_prior_desc &p = priors[k];
        if (p.kind < NF_PRIOR_PLAIN || p.kind > NF_PRIOR_RESOLVED_PLACEMENT) return NF_EINVAL;
        if (p.p_ix < 0 || p.p_ix >= n_model) return NF_EINVAL;
        const bool needs_dist = p.kind != NF_PRIOR_CONSTANT;
        if (needs_dist && (p.dist < 0 || p.dist >= n_dist)) return NF_EINVAL;
        const bool needs_dist2 = p.kind == NF_PRIOR_SPACED || p.kind == NF_PRIOR_CENSEP ||
                                 p.kind == NF_PRIOR_RESOLVED_CENSEP;
        if (needs_dist2 && (p.dist2 < 0 || p.dist2 >= n_dist)) return NF_EINVAL;
        if (p.kind == NF_PRIOR_DUPLICATE && (p.p_ix2 < 0 || p.p_ix2 >= n_model)) return NF_EINVAL;
        if (p.kind == NF_PRIOR_RESOLVED_CENSEP || p.kind == NF_PRIOR_RESOLVED_PLACEMENT) {
            if (p.p_ix2 < 0 || p.p_ix2 >= n_model) return NF_EINVAL;
            if (p.nested < 0 || p.nested >= n_prior || p.nested == k) return NF_EINVAL;
            const int nk = priors[p.nested].kind;   // nested sigma prior must be simple (no recursion)
            if (nk != NF_PRIOR_PLAIN && nk != NF_PRIOR_CONSTANT && nk != NF_PRIOR_ORDERED && nk != NF_PRIOR_DUPLICATE)
                return NF_EINVAL;
        }
    }
    int prev = -1;
    cudaGetDevice(&prev);
    if (cudaSetDevice(device) != cudaSuccess) return NF_ENODEV;
    nf_priors *pr = new (std::nothrow) nf_priors();
    if (!pr) { if (prev >= 0) cudaSetDevice(prev); return NF_ENOMEM; }
    std::memset(pr, 0, sizeof(*pr));
    pr->device = device; pr->n_prior = n_prior; pr->n_dist = n_dist; pr->n_model = n_model; pr->n_tables = n_tables;
    pr->h_priors = new (std::nothrow) nf_prior_desc[n_prior];
    if (!pr->h_priors) { delete pr; if (prev >= 0) cudaSetDevice(prev); return NF_ENOMEM; }
    std::memcpy(pr->h_priors, priors, sizeof(nf_prior_desc) * n_prior);
    cudaError_t e = cudaMalloc(&pr->priors, sizeof(nf_prior_desc) * n_prior);
    if (e == cudaSuccess) e = cudaMemcpy(pr->priors, priors, sizeof(nf_prior_desc) * n_prior, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && n_dist > 0) {
        e = cudaMalloc(&pr->dists, sizeof(nf_dist_desc) * n_dist);
        if (e == cudaSuccess) e = cudaMemcpy(pr->dists, dists, sizeof(nf_dist_desc) * n_dist, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMalloc(&pr->tables, sizeof(double) * n_tables);
        if (e == cudaSuccess) e = cudaMemcpy(pr->tables, tables, sizeof(double) * n_tables, cudaMemcpyHostToDevice);
    }
    // the uploads ran in the legacy stream; their consumers run on non-blocking streams, which do not order with it
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (prev >= 0) cudaSetDevice(prev);
    if (e != cudaSuccess) { nf_priors_free(pr); return (int)e; }
    *out = pr;
    return NF_OK;
}

int nf_priors_free(nf_priors *pr)
{
    if (!pr) return NF_OK;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(pr->device);
    if (pr->priors) cudaFree(pr->priors);
    if (pr->dists) cudaFree(pr->dists);
    if (pr->tables) cudaFree(pr->tables);
    if (prev >= 0) cudaSetDevice(prev);
    delete[] pr->h_priors;
    delete pr;
    return NF_OK;
}

int nf_prior_transform(const nf_priors *pr, double *u_dev, int64_t B, int ncomp, void *stream)
{
    if (!pr || !u_dev || B < 0 || ncomp < 1) return NF_EINVAL;
    int prev = -1;
    cudaGetDevice(&prev);
    if (cudaSetDevice(pr->device) != cudaSuccess) return NF_ENODEV;
    cudaError_t e = nf_launch_prior_transform(pr, u_dev, B, ncomp, (cudaStream_t)stream);
    if (prev >= 0) cudaSetDevice(prev);
    return (int)e;
}

int nf_prior_transform_host(const nf_priors *pr, double *u_host, int64_t B, int ncomp)
{
    if (!pr || !u_host || B < 0 || ncomp < 1) return NF_EINVAL;
    if (B == 0) return NF_OK;
    int prev = -1;
    cudaGetDevice(&prev);
    if (cudaSetDevice(pr->device) != cudaSuccess) return NF_ENODEV;
    const size_t bytes = sizeof(double) * (size_t)B * pr->n_model * ncomp;
    double *dev = nullptr;
    cudaError_t e = cudaMalloc(&dev, bytes);
    if (e == cudaSuccess) e = cudaMemcpy(dev, u_host, bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = nf_launch_prior_transform(pr, dev, B, ncomp, 0);
    if (e == cudaSuccess) e = cudaMemcpy(u_host, dev, bytes, cudaMemcpyDeviceToHost);
    if (dev) cudaFree(dev);
    if (prev >= 0) cudaSetDevice(prev);
    return (int)e;
}

}  // extern "C"
