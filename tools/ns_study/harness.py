"""CPU experiment harness: sampler variants on synthetic NH3 pixels with the C oracle likelihood."""
import sys, math, time
sys.path.insert(0, '/root/repo')
import numpy as np
import nestfit_b200.prior_constructors as pc
from oracle import oracle as orc

def make_pixel(ncomp_true, seed, n_chan=380, dv=0.158, noise=0.1):
    rng = np.random.default_rng(seed)
    ut = pc.get_irdc_priors(); packed = ut.pack()
    xs = [orc.bench_axis(1, n_chan, dv), orc.bench_axis(2, n_chan, dv)]
    while True:
        T = orc.prior_transform(packed, rng.uniform(0.1, 0.9, size=(1, 6 * ncomp_true)), ncomp_true)
        if np.isfinite(T).all(): break
    clean = orc.nh3_batch(xs, [1, 2], T, ncomp_true, want_pred=True)["pred"][0]
    data = clean + rng.normal(0, noise, clean.shape)
    return xs, packed, data, np.full(2, noise), T[0]

def make_score(xs, packed, data, noise, ncomp):
    d3, n2 = data[None], noise[None]
    cnt = [0]
    def score(U):
        U = np.atleast_2d(U)
        cnt[0] += U.shape[0]
        th = orc.prior_transform(packed, U, ncomp)
        out = orc.nh3_batch(xs, [1, 2], np.nan_to_num(th, nan=1.0), ncomp, data=d3, noise=n2)["lnL"]
        out[~np.isfinite(th).all(axis=1)] = np.nan
        return out
    return score, cnt
