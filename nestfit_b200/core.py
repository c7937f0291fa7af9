"""Host-side mirror of the reference's prior / spectrum / runner operator API.

Same class names, constructor arguments and error behaviour as
``nestfit.core.core`` (reference nestfit/core/core.pyx), but the objects here
only hold *descriptions* (tables, indices): all arithmetic on the hot path --
prior transform, model synthesis, chi-square -- runs in the CUDA library through
the C ABI (``include/nestfit_b200.h``).  There is no CPU fallback.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import DistDesc, PriorDesc

FWHM = 2.3548200450309493  # core.pyx:20

(KIND_PLAIN, KIND_CONSTANT, KIND_DUPLICATE, KIND_ORDERED, KIND_SPACED, KIND_CENSEP,
 KIND_RESOLVED_CENSEP, KIND_RESOLVED_PLACEMENT) = range(8)
FLAG_NESTED = 1
N_DIST_TABLES = 7   # xax, pdf, cdf, ppf, S0, S1, S2 (include/nf_priors.h)


class Distribution:
    """Numerically tabulated 1-D prior (reference core.pyx:23-45): the PDF on an
    even grid, its trapezoid CDF, and the PPF resampled on an even grid in
    probability by a cubic interpolating spline of the inverse CDF."""

    def __init__(self, xax, pdf):
        from scipy import integrate, interpolate
        xax = np.ascontiguousarray(xax, dtype=np.float64)
        pdf = np.ascontiguousarray(pdf, dtype=np.float64)
        assert xax[1] > xax[0]
        assert xax.shape == pdf.shape
        n = xax.shape[0]
        self.size = n
        self.xax, self.pdf = xax, pdf
        self.dx = float(xax[1] - xax[0])
        self.xmin, self.xmax = float(xax.min()), float(xax.max())
        cdf = integrate.cumulative_trapezoid(pdf, xax, initial=0)
        cdf /= cdf.max()
        self.cdf = cdf
        # strictly increasing copy of the CDF so it can be inverted
        strict = cdf + 1e-16 * np.arange(n)
        strict /= strict.max()
        u = np.linspace(0, 1, n)
        self.du = float(u[1] - u[0])
        self.ppf = interpolate.UnivariateSpline(strict, xax, k=3, s=0)(u)

    def tables(self):
        """xax, pdf, cdf, ppf each padded to size+1 (last value repeated), followed by the
        prefix moments S_m[i] = sum_{k<=i} k^m t_k of the trapezoid terms
        t_k = (pdf[k] + pdf[k-1]) / 2 (m = 0, 1, 2) that let the device evaluate the
        reference's interval CDF (core.pyx:109-161) in closed form."""
        def pad(a):
            return np.concatenate([a, a[-1:]])
        k = np.arange(self.size, dtype=np.float64)
        t = np.zeros(self.size)
        t[1:] = 0.5 * (self.pdf[1:] + self.pdf[:-1])
        moments = [np.cumsum(t * k**m) for m in range(3)]
        return np.concatenate([pad(self.xax), pad(self.pdf), pad(self.cdf), pad(self.ppf)] +
                              [pad(m) for m in moments])


class Prior:
    """Independent prior on model parameter row ``p_ix`` (core.pyx:169-197)."""
    kind = KIND_PLAIN

    def __init__(self, dist, p_ix):
        assert p_ix >= 0
        self.dist = dist
        self.p_ix = p_ix
        self.n_param = 1

    def _desc(self, plan, nested=False):
        return plan.add(self.kind, self.p_ix, dist=self.dist, nested_flag=nested)


class DuplicatePrior(Prior):
    """One draw written to two parameter rows (core.pyx:200-221)."""
    kind = KIND_DUPLICATE

    def __init__(self, dist, p_ix, p_ix_dup):
        assert p_ix >= 0
        assert p_ix_dup >= 0
        self.dist = dist
        self.p_ix = p_ix
        self.p_ix_dup = p_ix_dup
        self.n_param = 2

    def _desc(self, plan, nested=False):
        return plan.add(self.kind, self.p_ix, p_ix2=self.p_ix_dup, dist=self.dist, nested_flag=nested)


class ConstantPrior(Prior):
    """Parameter row fixed to a value (core.pyx:224-238)."""
    kind = KIND_CONSTANT

    def __init__(self, value, p_ix):
        self.value = float(value)
        self.p_ix = p_ix
        self.dist = None
        self.n_param = 1

    def _desc(self, plan, nested=False):
        return plan.add(self.kind, self.p_ix, value=self.value, nested_flag=nested)


class OrderedPrior(Prior):
    """Left-to-right ordered draws from one distribution (core.pyx:241-258)."""
    kind = KIND_ORDERED


class SpacedPrior(Prior):
    """First value from ``prior_indep``; running offsets from ``prior_depen``
    (core.pyx:261-292)."""
    kind = KIND_SPACED

    def __init__(self, prior_indep, prior_depen):
        self.prior_indep = prior_indep
        self.prior_depen = prior_depen
        self.p_ix = prior_indep.p_ix
        self.n_param = 1

    def _desc(self, plan, nested=False):
        return plan.add(self.kind, self.p_ix, dist=self.prior_indep.dist, dist2=self.prior_depen.dist)


class CenSepPrior(Prior):
    """Centre +/- half separation for up to two components (core.pyx:295-318)."""
    kind = KIND_CENSEP

    def __init__(self, vcen_prior, vsep_prior):
        self.vcen_prior = vcen_prior
        self.vsep_prior = vsep_prior
        self.p_ix = vcen_prior.p_ix
        self.n_param = 1

    def _desc(self, plan, nested=False):
        return plan.add(self.kind, self.p_ix, dist=self.vcen_prior.dist, dist2=self.vsep_prior.dist)


class ResolvedCenSepPrior(Prior):
    """CenSep with the separation floored at ``FWHM*scale*sqrt(s1*s2)``
    (core.pyx:321-366)."""
    kind = KIND_RESOLVED_CENSEP

    def __init__(self, vcen_prior, vsep_prior, sigm_prior, scale=1.5):
        self.vcen_prior = vcen_prior
        self.vsep_prior = vsep_prior
        self.sigm_prior = sigm_prior
        self.scale = scale
        self.sep_scale = FWHM * scale
        self.p_ix = vcen_prior.p_ix
        self.n_param = 2

    def _desc(self, plan, nested=False):
        k = self.sigm_prior._desc(plan, nested=True)
        return plan.add(self.kind, self.vcen_prior.p_ix, p_ix2=self.sigm_prior.p_ix,
                        dist=self.vcen_prior.dist, dist2=self.vsep_prior.dist, nested=k,
                        value=self.sep_scale)


class ResolvedPlacementPrior(Prior):
    """Sequential placement of up to 10 centroids separated by at least
    ``FWHM*scale*sqrt(s_i*s_{i-1})`` (core.pyx:369-434)."""
    kind = KIND_RESOLVED_PLACEMENT

    def __init__(self, vcen_prior, sigm_prior, scale=1.5):
        self.vcen_prior = vcen_prior
        self.sigm_prior = sigm_prior
        self.scale = scale
        self.sep_scale = FWHM * scale
        self.p_ix = vcen_prior.p_ix
        self.n_param = 2

    def _desc(self, plan, nested=False):
        k = self.sigm_prior._desc(plan, nested=True)
        return plan.add(self.kind, self.vcen_prior.p_ix, p_ix2=self.sigm_prior.p_ix,
                        dist=self.vcen_prior.dist, nested=k, value=self.sep_scale)


class _Plan:
    """Flattens prior objects into the packed arrays of include/nf_priors.h."""

    def __init__(self):
        self.dists = []       # Distribution objects (deduplicated by identity)
        self.records = []

    def _dist_index(self, dist):
        if dist is None:
            return -1
        for i, d in enumerate(self.dists):
            if d is dist:
                return i
        self.dists.append(dist)
        return len(self.dists) - 1

    def add(self, kind, p_ix, p_ix2=-1, dist=None, dist2=None, nested=-1, value=0.0, nested_flag=False):
        rec = PriorDesc(kind=kind, flags=FLAG_NESTED if nested_flag else 0, p_ix=int(p_ix), p_ix2=int(p_ix2),
                        dist=self._dist_index(dist), dist2=self._dist_index(dist2), nested=int(nested),
                        pad_=0, value=float(value))
        self.records.append(rec)
        return len(self.records) - 1

    def arrays(self):
        n_d = len(self.dists)
        dd = (DistDesc * max(n_d, 1))()
        chunks, off = [], 0
        for i, d in enumerate(self.dists):
            t = d.tables()
            stride = d.size + 1
            dd[i] = DistDesc(size=d.size, stride=stride, offset=off, pad_=0, xmin=d.xmin, xmax=d.xmax,
                             dx=d.dx, du=d.du)
            chunks.append(t)
            off += N_DIST_TABLES * stride
        tables = np.ascontiguousarray(np.concatenate(chunks) if chunks else np.zeros(1), dtype=np.float64)
        pp = (PriorDesc * len(self.records))(*self.records)
        return pp, len(self.records), dd, n_d, tables


class PriorTransformer:
    """Ordered list of priors mapping the unit cube to physical parameters
    (reference core.pyx:437-483).  ``transform`` runs on the device."""

    def __init__(self, priors):
        priors = np.asarray(priors, dtype=object)
        n_prior = priors.shape[0]
        assert n_prior >= 1
        self.priors = priors
        self.n_prior = n_prior
        self.n_param = int(sum(p.n_param for p in priors))
        self._handles = {}

    # -- packed description -------------------------------------------------
    def pack(self):
        plan = _Plan()
        for p in self.priors:
            p._desc(plan)
        return plan.arrays()

    def handle(self, device=0):
        """Device-resident plan (``nf_priors*``), created on first use."""
        h = self._handles.get(device)
        if h is None:
            lib = _lib.load()
            pp, n_p, dd, n_d, tables = self.pack()
            out = C.c_void_p()
            _lib.check(lib.nf_priors_create(device, C.cast(pp, C.c_void_p), n_p, C.cast(dd, C.c_void_p), n_d,
                                            _lib.ptr(tables), tables.size, self.n_param, C.byref(out)),
                       "nf_priors_create")
            h = out
            self._handles[device] = h
        return h

    def transform(self, utheta, ncomp):
        """In-place unit-cube -> physical transform of one vector (core.pyx:478-483)."""
        utheta = np.asarray(utheta)
        if self.n_param * ncomp != utheta.shape[0]:
            shape = utheta.shape[0]
            raise ValueError(f'Invalid shape for ncomp={ncomp}: {shape}')
        self.transform_batch(utheta.reshape(1, -1), ncomp)

    def transform_batch(self, u, ncomp, device=0):
        """In-place transform of ``u[B, n_param*ncomp]`` (float64, C-contiguous)."""
        if u.dtype != np.float64 or not u.flags["C_CONTIGUOUS"]:
            raise ValueError("transform_batch needs a C-contiguous float64 array")
        if u.ndim != 2 or u.shape[1] != self.n_param * ncomp:
            raise ValueError(f'Invalid shape for ncomp={ncomp}: {u.shape}')
        lib = _lib.load()
        _lib.check(lib.nf_prior_transform_host(self.handle(device), _lib.ptr(u), u.shape[0], ncomp),
                   "nf_prior_transform_host")
        return u

    def __getstate__(self):
        """Device handles are per process: a pickled copy (one worker per GPU, main.py:515-523)
        rebuilds its own plan on first use."""
        state = self.__dict__.copy()
        state['_handles'] = {}
        return state

    def __del__(self):
        try:
            lib = _lib.load()
            for h in self._handles.values():
                lib.nf_priors_free(h)
        except Exception:
            pass


class Spectrum:
    """One spectrum: ascending frequency axis [Hz], brightness temperature [K],
    scalar rms noise [K] (reference core.pyx:486-545)."""

    def __init__(self, xarr, data, noise, rest_freq=None, trans_id=None):
        xarr = np.ascontiguousarray(xarr, dtype=np.float64)
        data = np.ascontiguousarray(data, dtype=np.float64)
        assert noise > 0
        nu_chan = xarr[1] - xarr[0]
        assert nu_chan > 0
        self.xarr, self.data = xarr, data
        self.noise = float(noise)
        self.size = xarr.shape[0]
        self.rest_freq = 0 if rest_freq is None else rest_freq
        self.trans_id = -1 if trans_id is None else trans_id
        self.nu_chan = float(nu_chan)
        self.nu_min = float(xarr[0])
        self.nu_max = float(xarr[-1])
        check_uniform_axis(xarr)
        self.pred = np.zeros_like(data)
        self.prefactor = -self.size / 2 * np.log(2 * np.pi * noise**2)
        self.null_lnZ = self.loglikelihood

    @property
    def loglikelihood(self):
        """-sum (d - pred)^2 / (2 noise^2) of the *current* ``pred`` (bookkeeping
        property, core.pyx:522-530,540-542; the sampler path never uses it)."""
        dev = self.data - self.pred
        return float(-np.dot(dev, dev) / (2 * self.noise**2))

    @property
    def sum_spec(self):
        return np.nansum(self.pred)

    @property
    def max_spec(self):
        return np.nanmax(self.pred)

    def get_spec(self):
        return np.array(self.pred)


class HyperfineSpectrum(Spectrum):
    pass


def check_uniform_axis(xarr, tol=1e-3):
    """The device kernels assume x_j = x_0 + j*(x_1 - x_0) (the reference's docs
    require uniform channels, docs/limitations.rst:13).  Deviations above
    ``tol`` channels are rejected loudly."""
    n = xarr.shape[0]
    ideal = xarr[0] + np.arange(n) * (xarr[1] - xarr[0])
    dev = np.max(np.abs(xarr - ideal)) / (xarr[1] - xarr[0])
    if not dev <= tol:
        raise ValueError(f"frequency axis is not uniform (max deviation {dev:.3g} channels)")


class Runner:
    """Likelihood operator interface (reference core.pxd:63-72, core.pyx:553-561)."""
    n_model = 0
    ncomp = 0
    n_params = 0
    ndim = 0
    n_chan_tot = 0
    n_spec = 0
    null_lnZ = 0.0
    run_lnZ = np.nan

    def loglikelihood(self, utheta):
        raise NotImplementedError
