"""The result contract pinned to the reference's own code: the compiled reference's
``run_multinest`` -> ``mn_loglikelihood`` -> ``mn_dump`` -> ``Dumper`` (core.pyx:564-687,727-823) is driven through the
``nf_oracle_ns_run`` hook of the stub ``multinest.h`` (oracle/build_ref.py) with the numpy sampler port standing where
MultiNest is, writing into a ``MemGroup``.  The same dead points then go through this package's product formatter
(``run_products``, what ``NestedSamplingBatch.products`` and -- column for column -- the device pass
``products_all`` produce) and the two trees are compared attribute by attribute, dataset by dataset.
CPU only; skipped where the compiled reference is absent."""
import ctypes as C

import numpy as np
import pytest

from oracle import ns_port
from oracle import oracle as orc
from oracle import ref as oref

pytestmark = pytest.mark.skipif(not oref.available(), reason="oracle/_ref not built")

# PYFUNCTYPE: the reference's callbacks touch Python objects, so the GIL stays held while ctypes calls them
LOGLIKE = C.PYFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_void_p)
DUMPER = C.PYFUNCTYPE(None, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.POINTER(C.c_double)),
                     C.POINTER(C.POINTER(C.c_double)), C.POINTER(C.POINTER(C.c_double)), C.POINTER(C.c_double),
                     C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p)
NS_RUN = C.CFUNCTYPE(None, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, LOGLIKE, DUMPER, C.c_void_p)


def _install_hook(core_module, captured):
    """Point the stub's `nf_oracle_ns_run` at a sampler: ns_port scored through the reference's LogLike callback,
    results handed to the reference's dumper callback in MultiNest's layout (Fortran-ordered posterior
    [n_samples, nPar + 2] = theta, lnL, posterior weight; param_constr [4 nPar] = mean, sigma, best fit, MAP)."""
    lib = C.CDLL(core_module.__file__)

    def ns_run(nlive, tol, efr, ndims, npar, seed, maxiter, loglike, dumper, context):
        nd, npar_c = C.c_int(ndims), C.c_int(npar)

        def score(U):
            out = np.empty(U.shape[0])
            for b in range(U.shape[0]):
                row = np.ascontiguousarray(U[b], dtype=np.float64).copy()
                lnl = C.c_double()
                loglike(row.ctypes.data_as(C.POINTER(C.c_double)), C.byref(nd), C.byref(npar_c), C.byref(lnl), context)
                out[b] = lnl.value if np.isfinite(row).all() else np.nan
                theta_of[U[b].tobytes()] = row          # the callback leaves the physical parameters in place
            return out

        theta_of = {}
        res = ns_port.nested_sampling(score, ndims, nlive, tol=tol, efr=efr, n_prop=16, seed=max(seed, 0),
                                      max_iter=maxiter, return_samples=True)
        keep = res['samples_lnL'] > -np.inf
        U, lnl, lnw = res['samples_u'][keep], res['samples_lnL'][keep], res['samples_lnw'][keep]
        theta = np.array([theta_of[u.tobytes()] for u in U])
        w = np.exp(lnl + lnw - res['lnZ'])
        post = np.asfortranarray(np.concatenate([theta, lnl[:, None], w[:, None]], axis=1))
        best, mapp = theta[np.argmax(lnl)], theta[np.argmax(lnl + lnw)]
        pcon = np.concatenate([np.average(theta, axis=0, weights=w), np.zeros(npar), best, mapp])
        captured.update(res=res, post=np.array(post), best=best.copy(), mapp=mapp.copy(), nlive=nlive)
        n_s, n_l, n_p = C.c_int(post.shape[0]), C.c_int(nlive), C.c_int(npar)
        pp = C.cast(post.ctypes.data, C.POINTER(C.c_double))
        pc = C.cast(pcon.ctypes.data, C.POINTER(C.c_double))
        live = np.zeros((nlive, npar + 1), order='F')
        pl = C.cast(live.ctypes.data, C.POINTER(C.c_double))
        mx, z, zi, ze = (C.c_double(v) for v in (res['max_loglike'], res['lnZ'], res['lnZ'], res['lnZ_err']))
        for _ in range(2):          # the last two calls carry the same n_samples: the second one writes (core.pyx:636-641)
            dumper(C.byref(n_s), C.byref(n_l), C.byref(n_p), C.byref(pl), C.byref(pp), C.byref(pc), C.byref(mx),
                   C.byref(z), C.byref(zi), C.byref(ze), context)

    cb = NS_RUN(ns_run)
    slot = C.c_void_p.in_dll(lib, "nf_oracle_ns_run")
    slot.value = C.cast(cb, C.c_void_p).value
    return cb, slot


@pytest.mark.parametrize("ncomp", [1, 2])
def test_reference_mn_dump_equals_run_products(ncomp):
    from nestfit_b200.sampler import run_products
    from nestfit_b200.store import MemGroup
    m = oref.load()
    rng = np.random.default_rng(5)
    xs = [orc.bench_axis(1, 380, 0.158), orc.bench_axis(2, 380, 0.158)]
    truth = np.array([-1, 1.5, 10, 15, 4, 6, 14.5, 15, .3, .6, 0, 0], dtype=float).reshape(6, 2)[:, :ncomp].ravel()
    clean = orc.nh3_batch(xs, [1, 2], truth[None], ncomp, want_pred=True)["pred"][0]
    data = clean + rng.normal(0, 0.2, clean.shape)
    specs = np.array([m.ammonia.AmmoniaSpectrum(xs[t], data[t].copy(), 0.2, trans_id=t + 1) for t in (0, 1)])
    runner = m.ammonia.AmmoniaRunner(specs, oref.make_irdc_priors(m.core), ncomp=ncomp)
    group = MemGroup(f"/pix/0/0/{ncomp}")
    dumper = m.core.Dumper(group)
    captured = {}
    cb, slot = _install_hook(m.core, captured)
    try:
        m.core.run_multinest(runner, dumper, nlive=60, seed=5, tol=1.0, efr=0.3, updInt=2000)
    finally:
        slot.value = None
    assert captured, "the hook was not called"
    res, post = captured['res'], captured['post']
    assert runner.run_lnZ == res['lnZ'] == group.attrs['global_lnZ']
    # the same dead points through this package's formatter
    attrs, dsets = run_products(ncomp, 6 * ncomp, captured['nlive'], runner.null_lnZ, runner.n_chan_tot, post,
                                res['lnZ'], res['lnZ_err'], res['max_loglike'], captured['best'], captured['mapp'])
    ref_attrs = dict(group.attrs)
    ours_only = {'n_iter', 'n_evals', 'truncated'}                 # bookkeeping extras, not in the reference
    assert set(ref_attrs) == set(attrs) - ours_only
    for k, v in ref_attrs.items():
        if isinstance(v, (list, np.ndarray)) or hasattr(v, '__len__') and not isinstance(v, str):
            assert list(np.asarray(v)) == list(np.asarray(attrs[k])), k
        else:
            assert type(attrs[k]) in (int, float) and attrs[k] == pytest.approx(v, rel=1e-14, abs=0), k
    ref_dsets = {k: np.asarray(group[k]) for k in group.keys()}
    assert set(ref_dsets) == set(dsets) - {'marginals_weighted'}    # an addition of this package
    for k, v in ref_dsets.items():
        assert dsets[k].dtype == v.dtype and dsets[k].shape == v.shape, k
        np.testing.assert_array_equal(dsets[k], v, err_msg=k)
    assert ref_dsets['posteriors'].dtype == np.float32 and ref_dsets['posteriors'].shape == (post.shape[0], 6 * ncomp + 2)
    assert ref_dsets['marginals'].shape == (15, 6 * ncomp)
