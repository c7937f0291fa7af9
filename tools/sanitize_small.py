"""Small end-to-end exercise of every kernel for compute-sanitizer."""
import sys
import numpy as np
sys.path.insert(0, '.')
import nestfit_b200 as nb
from nestfit_b200.sampler import NestedSamplingBatch
from oracle import oracle as orc
rng = np.random.default_rng(0)
ut = nb.get_irdc_priors()
xs = [orc.bench_axis(1, 200, 0.3), orc.bench_axis(2, 200, 0.3)]
for nc in (1, 2, 3, 4):
    P = ut.transform_batch(rng.uniform(size=(96, 6 * nc)), nc)
    P[~np.isfinite(P).all(axis=1)] = P[0]
    blk = nb.PixelBlock("ammonia", xs, rng.normal(0, 0.1, (5, 2, 200)).astype(np.float32), 0.1, trans_ids=[1, 2])
    pix = np.sort(rng.integers(0, 5, 96)).astype(np.int32)
    l1 = blk.loglike(P, nc, pix_of_vec=pix)
    pr = blk.predict(P[:7], nc)
    assert np.isfinite(l1).all() and np.isfinite(pr).all()
    if nc <= 2:
        ns = NestedSamplingBatch(blk, ut, nc, nlive=40, tol=1.0, n_prop=8, seed=1, max_iter=150)
        r = ns.run(); assert np.isfinite(r['lnZ']).all(); ns.posterior(0); ns.close()
    blk.close()
v = (np.arange(300) - 149.5) * 0.3
x = np.sort(orc.NU[0] * (1 - v / orc.CKMS))
Pg = np.concatenate([rng.uniform(-30, 30, (40, 5)), rng.uniform(0.3, 3, (40, 5)), rng.uniform(0.1, 4, (40, 5))], axis=1)
gb = nb.PixelBlock("gaussian", [x], rng.normal(0, 0.1, (2, 1, 300)).astype(np.float32), 0.1, rest_freq=orc.NU[0])
assert np.isfinite(gb.loglike(Pg, 5, vecs_per_pix=20)).all() and np.isfinite(gb.predict(Pg, 5)).all()
# wide band (more than one 2048-channel super-block), odd channel count, (3,3) ortho transition
xw = [orc.bench_axis(1, 2300, 0.03), orc.bench_axis(3, 2300, 0.03)]
Pw = ut.transform_batch(rng.uniform(size=(40, 12)), 2)
Pw[~np.isfinite(Pw).all(axis=1)] = Pw[0]
Pw[:, 10:] = 0.5
bw = nb.PixelBlock("ammonia", xw, rng.normal(0, 0.1, (2, 2, 2300)).astype(np.float32), 0.1, trans_ids=[1, 3])
assert np.isfinite(bw.loglike(Pw, 2, vecs_per_pix=20)).all() and np.isfinite(bw.predict(Pw[:3], 2)).all()
bw.close()
# N2H+ (45-line (3-2) transition), ragged pixel map, sampler with the tail scheduling (few runs, large K)
xn = [np.sort(orc.N2HP_NU[t - 1] * (1 - (np.arange(333) - 166.0) * 0.1 / orc.CKMS)) for t in (1, 3)]
Pn = np.concatenate([np.sort(rng.uniform(-6, 6, (50, 3)), axis=1), rng.uniform(3, 20, (50, 3)), rng.uniform(-1.5, 1.2, (50, 3)),
                     rng.uniform(0.08, 1.5, (50, 3))], axis=1)
bn = nb.PixelBlock("diazenylium", xn, rng.normal(0, 0.1, (3, 2, 333)).astype(np.float32), 0.1, trans_ids=[1, 3])
assert np.isfinite(bn.loglike(Pn, 3, pix_of_vec=rng.integers(0, 3, 50).astype(np.int32))).all()
assert np.isfinite(bn.predict(Pn[:4], 3)).all()
bn.close()
blk = nb.PixelBlock("ammonia", xs, rng.normal(0, 0.1, (2, 2, 200)).astype(np.float32), 0.1, trans_ids=[1, 2])
ns = NestedSamplingBatch(blk, ut, 2, pix_ids=np.array([0, 1, 1]), nlive=40, tol=1.0, n_prop=8, seed=2, max_iter=400,
                         method='rwalk', walks=6)
r = ns.run(); assert np.isfinite(r['lnZ']).all(); ns.close(); blk.close()
print("sanitize run ok")
# [r2] products pass (packing, column sort), the ellipsoid decomposition, and a cube fit with two host threads on two
# streams plus the background wave writer (the paths a multi-threaded caller exercises)
from nestfit_b200.synth import make_synth_stack
from nestfit_b200.models import ammonia
blk = nb.PixelBlock("ammonia", xs, rng.normal(0, 0.1, (3, 2, 200)).astype(np.float32), 0.1, trans_ids=[1, 2])
ns = NestedSamplingBatch(blk, ut, 1, nlive=[30, 40, 50], tol=1.0, n_prop=8, seed=3, max_iter=300, mmodal=True)
r = ns.run(); p = ns.products_all(); assert p['posteriors'].shape[0] == p['row_offsets'][-1]; ns.close(); blk.close()
stack = make_synth_stack((4, 4), ut, ncomp_map=np.ones((4, 4), dtype=int), n_chan=200, dv=0.3, noise=0.1, seed=2)
fitter = nb.CubeFitter(stack, ut, ammonia.AmmoniaRunner, ncomp_max=1, mn_kwargs={'nlive': 30, 'maxiter': 200}, n_prop=8,
                       n_streams=2, pixels_per_stream=4)
import tempfile
with tempfile.TemporaryDirectory() as td:
    res = fitter.fit_cube(td + '/s', nproc=1)[0]
    assert (res['nbest'] >= 0).all()
print("sanitize r2 ok")
