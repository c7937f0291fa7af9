"""Batched nested sampling: host wrapper of ``nf_ns_*`` (include/nestfit_b200.h)
and the drop-in ``run_multinest`` / ``Dumper`` pair of the reference
(nestfit/core/core.pyx:564-609,627-687,727-823).

The sampler itself (proposal, prior transform, fused likelihood, accept/replace,
evidence accumulation) runs on the device; this module sizes buffers, launches
the run and packages results in the reference's result contract.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import NsConfig

_VP, _I, _I64 = C.c_void_p, C.c_int, C.c_int64
_lib.register("nf_ns_create", [_VP, _VP, _I, _I, C.POINTER(NsConfig), _I64, _VP, _VP, C.POINTER(_VP)])
_lib.register("nf_ns_run", [_VP])
_lib.register("nf_ns_free", [_VP])
_lib.register("nf_ns_results", [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP])
_lib.register("nf_ns_posterior", [_VP, _I64, C.c_int32, _VP, _VP, _VP])
_lib.register("nf_ns_stats", [_VP, C.POINTER(C.c_int32), C.POINTER(_I64)])
_lib.register("nf_ns_products_rows", [_VP, _VP])
_lib.register("nf_ns_products", [_VP, _VP, _I, _VP, _VP])

from .store import MARG_COLS, MARG_QUANTILES, run_attr_columns  # noqa: E402  (names re-exported)


class NestedSamplingBatch:
    """Lock-step nested sampling of ``n_run`` (pixel, ncomp) fits on one GPU."""

    def __init__(self, block, utrans, ncomp, pix_ids=None, nlive=100, tol=1.0, efr=0.3, n_prop=32, seed=1,
                 max_iter=1_000_000, max_samples=None, cold=False, lte=False, method='auto', walks=0,
                 n_prop_max=None, target_batch=65536, keep_constant_dims=False, mmodal=False):
        """method: 'auto' = ellipsoidal rejection sampling that hands a run over to a
        constrained random walk once its acceptance stalls; 'ellipsoid' / 'rwalk' force one.
        walks: random-walk steps per new point (0 = 20 + number of dimensions the likelihood depends on).
        keep_constant_dims: keep the cube dimensions the priors overwrite (ConstantPrior rows, DuplicatePrior's
        second row) inside the bounding ellipsoid and the walk metric (the round-1 behaviour; default: they are
        drawn uniformly on their own).
        mmodal: MultiNest-style decomposition of the live set into up to 8 ellipsoids for the rejection phase (the
        reference runs MultiNest with mmodal=True, core.pyx:729); default False = one bounding ellipsoid rebuilt every
        lock-step (measured on B200: the decomposition saves 12 % of the likelihood calls of a cube fit but its
        clustering and overlap tests cost more than they save, and it raises ln Z by ~1-3 at 10-15 dimensions;
        DESIGN.md section 4.4).
        n_prop_max / target_batch: once few runs are still active each gets up to n_prop_max
        proposals per lock-step (default 16 n_prop) so that a launch keeps ~target_batch vectors."""
        lib = _lib.load()
        self.block, self.utrans, self.ncomp = block, utrans, int(ncomp)
        if pix_ids is None:
            pix_ids = np.arange(block.n_pix)
        self.pix_ids = np.ascontiguousarray(pix_ids, dtype=np.int32)
        self.n_run = int(self.pix_ids.size)
        self.nlive = np.ascontiguousarray(np.broadcast_to(np.asarray(nlive, dtype=np.int32), (self.n_run,)))
        self.ndim = block.n_model * self.ncomp
        nlive_max = int(self.nlive.max())
        if max_samples is None:
            max_samples = 64 * nlive_max
        if seed is None or seed < 0:       # reference default seed=-1 means "from the clock" (core.pyx:731)
            seed = int(np.random.SeedSequence().generate_state(1)[0])
        self.cfg = NsConfig(nlive_max=nlive_max, n_prop=int(n_prop), max_iter=int(min(max_iter, 2**31 - 1)),
                            max_samples=int(max_samples), bound_update_interval=int(walks),
                            flags={'auto': 0, 'rwalk': 1, 'ellipsoid': 2}[method] | (4 if keep_constant_dims else 0) | (0 if mmodal else 8),
                            tol=float(tol),
                            efr=float(efr), seed=int(seed),
                            n_prop_max=int(16 * n_prop if n_prop_max is None else n_prop_max),
                            target_batch=int(target_batch))
        flags = (_lib.NF_FLAG_COLD if cold else 0) | (_lib.NF_FLAG_LTE if lte else 0)
        out = C.c_void_p()
        _lib.check(lib.nf_ns_create(block.handle, utrans.handle(block.device), self.ncomp, flags,
                                    C.byref(self.cfg), self.n_run, _lib.ptr(self.pix_ids), _lib.ptr(self.nlive),
                                    C.byref(out)), "nf_ns_create")
        self.handle = out
        self._results = None

    def close(self):
        if getattr(self, "handle", None):
            _lib.load().nf_ns_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self):
        lib = _lib.load()
        _lib.check(lib.nf_ns_run(self.handle), "nf_ns_run")
        R, D = self.n_run, self.ndim
        res = dict(lnZ=np.empty(R), lnZ_err=np.empty(R), max_loglike=np.empty(R),
                   n_samples=np.empty(R, dtype=np.int32), n_iter=np.empty(R, dtype=np.int32),
                   n_evals=np.empty(R, dtype=np.int64), bestfit=np.empty((R, D)), mapfit=np.empty((R, D)))
        _lib.check(lib.nf_ns_results(self.handle, _lib.ptr(res["lnZ"]), _lib.ptr(res["lnZ_err"]),
                                     _lib.ptr(res["max_loglike"]), _lib.ptr(res["n_samples"]),
                                     _lib.ptr(res["n_iter"]), _lib.ptr(res["n_evals"]), _lib.ptr(res["bestfit"]),
                                     _lib.ptr(res["mapfit"])), "nf_ns_results")
        it, ln = C.c_int32(), C.c_int64()
        lib.nf_ns_stats(self.handle, C.byref(it), C.byref(ln))
        res["lock_iters"], res["launches"] = it.value, ln.value
        # a run that filled its share of the dead-point pool stopped before its evidence converged
        per_live = self.cfg.max_samples // self.cfg.nlive_max
        res["truncated"] = res["n_samples"] >= np.minimum(self.cfg.max_samples, per_live * self.nlive)
        self._results = res
        self._batch = None
        return res

    def products_all(self, marginals=True, posteriors=True):
        """Posterior products of every run with one device pass and one device-to-host copy (instead of one
        `posterior()` round trip per run): dict(row_offsets [n_run + 1], posteriors float32 [rows, ndim + 2] in
        the layout of the reference's `posteriors` dataset (core.pyx:680), marginals [n_run, 15, ndim])."""
        if getattr(self, "_batch", None) is not None:
            return self._batch
        self.results
        lib = _lib.load()
        off = np.empty(self.n_run + 1, dtype=np.int64)
        _lib.check(lib.nf_ns_products_rows(self.handle, _lib.ptr(off)), "nf_ns_products_rows")
        post = np.empty((int(off[-1]), self.ndim + 2), dtype=np.float32) if posteriors else None
        marg = np.empty((self.n_run, MARG_QUANTILES.size, self.ndim)) if marginals else None
        _lib.check(lib.nf_ns_products(self.handle, _lib.ptr(MARG_QUANTILES), MARG_QUANTILES.size if marginals else 0,
                                      _lib.ptr(post), _lib.ptr(marg)), "nf_ns_products")
        self._batch = dict(row_offsets=off, posteriors=post, marginals=marg)
        return self._batch

    @property
    def results(self):
        if self._results is None:
            self.run()
        return self._results

    def posterior(self, run):
        """(n_samples, ndim + 2) float64: physical parameters, lnL, posterior weight
        (MultiNest's `posterior` array convention; weights sum to 1)."""
        res = self.results
        n = int(res["n_samples"][run])
        th = np.empty((n, self.ndim), dtype=np.float32)
        lnl = np.empty(n)
        lnw = np.empty(n)
        _lib.check(_lib.load().nf_ns_posterior(self.handle, run, n, _lib.ptr(th), _lib.ptr(lnl), _lib.ptr(lnw)),
                   "nf_ns_posterior")
        keep = lnl > -np.inf          # logZero points (NaN prior draws among the first live set) carry no weight
        if not keep.all():
            th, lnl, lnw = th[keep], lnl[keep], lnw[keep]
        w = np.exp(lnl + lnw - res["lnZ"][run])
        return np.concatenate([th.astype(np.float64), lnl[:, None], w[:, None]], axis=1)

    def products(self, run, null_lnZ, n_chan_tot):
        """Attributes and datasets the reference's dumper writes for one run
        (core.pyx:645-687), as two dicts."""
        res = self.results
        return run_products(self.ncomp, self.ndim, int(self.nlive[run]), null_lnZ, n_chan_tot, self.posterior(run),
                            res["lnZ"][run], res["lnZ_err"][run], res["max_loglike"][run], res["bestfit"][run],
                            res["mapfit"][run], n_iter=res["n_iter"][run], n_evals=res["n_evals"][run],
                            truncated=res["truncated"][run])


def run_products(ncomp, ndim, nlive, null_lnZ, n_chan_tot, post, lnZ, lnZ_err, max_loglike, bestfit, mapfit,
                 n_iter=0, n_evals=0, truncated=False):
    """What `mn_dump` persists for one run (core.pyx:645-687) from the run's posterior array `post`
    [n_samples, ndim + 2] (theta, lnL, posterior weight) and its summary numbers: (attrs, datasets)."""
    res = dict(lnZ=[lnZ], lnZ_err=[lnZ_err], max_loglike=[max_loglike], n_iter=[n_iter], n_evals=[n_evals],
               truncated=[truncated])
    cols = run_attr_columns(ndim, n_chan_tot, [nlive], [null_lnZ], res, [post.shape[0]])
    attrs = {'ncomp': int(ncomp), 'n_chan_tot': int(n_chan_tot), 'n_params': int(ndim), 'marg_cols': MARG_COLS,
             'marg_quantiles': MARG_QUANTILES}
    attrs.update({k: v[0].item() for k, v in cols.items()})
    dsets = {
        'posteriors': post.astype('float32'),
        # unweighted quantiles over all rows, mirroring core.pyx:596-598
        'marginals': np.quantile(post[:, :-2], MARG_QUANTILES, axis=0),
        'marginals_weighted': weighted_quantiles(post[:, :-2], post[:, -1], MARG_QUANTILES),
        'bestfit_params': np.array(bestfit, dtype=np.float64),
        'map_params': np.array(mapfit, dtype=np.float64),
    }
    return attrs, dsets


def weighted_quantiles(x, w, q):
    """Posterior-weighted quantiles per column (an addition: the reference's
    `marginals` ignores the importance weights)."""
    out = np.empty((len(q), x.shape[1]))
    for j in range(x.shape[1]):
        o = np.argsort(x[:, j])
        cw = np.cumsum(w[o])
        cw /= cw[-1]
        out[:, j] = np.interp(q, cw, x[o, j])
    return out


class Dumper:
    """Result sink with the reference's interface (core.pyx:564-609): writes run
    attributes and datasets into an h5py-like group."""

    def __init__(self, group, no_dump=False):
        self.group = group
        self.no_dump = no_dump
        self.n_calls = 0
        self.n_samples = -1
        self.quantiles = MARG_QUANTILES
        self.marginal_cols = MARG_COLS

    def calc_marginals(self, posteriors):
        return np.quantile(posteriors[:, :-2], self.quantiles, axis=0)

    def flush(self):
        f = getattr(self.group, "file", None)
        if f is not None:
            f.flush()

    def append_attributes(self, **kwargs):
        for name, value in kwargs.items():
            self.group.attrs[name] = value

    def append_datasets(self, **kwargs):
        for name, data in kwargs.items():
            self.group.create_dataset(name, data=data)


def run_multinest(runner, dumper, IS=False, mmodal=True, ceff=False, nlive=400, tol=0.5, efr=0.3, nClsPar=None,
                  maxModes=100, updInt=10, Ztol=-1e90, root='results', seed=-1, pWrap=None, fb=False,
                  resume=False, initMPI=False, outfile=False, logZero=-1e100, maxiter=int(1e6), n_prop=32):
    """Drop-in for the reference's ``run_multinest`` (core.pyx:727-823): fits the
    runner's pixel with the batched device sampler (one run), sets
    ``runner.run_lnZ`` and writes the reference's products through ``dumper``.
    MultiNest-only switches (IS, mmodal, ceff, nClsPar, maxModes, updInt, Ztol,
    pWrap, fb, resume, initMPI, outfile, logZero, root) are accepted and ignored."""
    assert runner.ndim > 0
    assert nlive > 0
    assert tol > 0
    assert 0 < efr <= 1
    assert maxModes > 0
    assert updInt > 0
    assert Ztol is not None and np.isfinite(Ztol)
    assert logZero is not None and np.isfinite(logZero)
    assert maxiter >= 0
    if nClsPar is not None and nClsPar > runner.n_params:
        raise ValueError('Number of clustering parameters must be less than total.')
    ns = NestedSamplingBatch(runner._block, runner.utrans, runner.ncomp, pix_ids=[0], nlive=nlive, tol=tol, efr=efr,
                             n_prop=n_prop, seed=seed, max_iter=maxiter,
                             cold=getattr(runner, "cold", False), lte=getattr(runner, "lte", False))
    res = ns.run()
    runner.run_lnZ = float(res["lnZ"][0])
    if dumper is not None and not dumper.no_dump:
        attrs, dsets = ns.products(0, runner.null_lnZ, runner.n_chan_tot)
        dumper.append_attributes(**attrs)
        dumper.append_datasets(**dsets)
    dumper_calls = getattr(dumper, "n_calls", 0)
    if dumper is not None:
        dumper.n_calls = dumper_calls + 1
        dumper.n_samples = int(res["n_samples"][0])
    ns.close()
    return res
