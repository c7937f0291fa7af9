"""lnL parity of the fused kernel against the C oracle under the literal north-star rule, on a large prior-drawn
sample of the bench workload (test infrastructure): |dlnL| <= 1e-3 within 1e3 of the pixel's best lnL, and the
distribution of |dlnL| / |lnL| for the poor fits beyond.  Usage: python tools/parity_strict.py [n_vectors]"""
import sys
import numpy as np
sys.path.insert(0, '.')
import nestfit_b200 as nb
from oracle import oracle as orc
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 17
rng = np.random.default_rng(99)
ut = nb.get_irdc_priors()
xs = [orc.bench_axis(1, 1000, 0.07), orc.bench_axis(2, 1000, 0.07)]
n_pix = 64
vpp = n // n_pix
for ncomp in (3, 1, 2, 4):
    T = orc.prior_transform(ut.pack(), rng.uniform(0.1, 0.9, size=(4 * n_pix, 6 * ncomp)), ncomp)
    T = T[np.isfinite(T).all(axis=1)][:n_pix]
    clean = orc.nh3_batch(xs, [1, 2], T, ncomp, want_pred=True)["pred"]
    data = (clean + rng.normal(0, 0.1, clean.shape)).astype(np.float32)
    P = orc.prior_transform(ut.pack(), rng.uniform(size=(n_pix * vpp + 4096, 6 * ncomp)), ncomp)
    P = P[np.isfinite(P).all(axis=1)][:n_pix * vpp]
    P[::vpp] = T                                     # the truth of every pixel
    for dt in (np.float32, np.float64):
        Pd = P.astype(dt)
        blk = nb.PixelBlock("ammonia", xs, data, 0.1, trans_ids=[1, 2])
        got = blk.loglike(Pd, ncomp, vecs_per_pix=vpp)
        blk.close()
        want = orc.nh3_batch(xs, [1, 2], Pd.astype(np.float64), ncomp, data=data.astype(np.float64), noise=np.full((n_pix, 2), 0.1),
                             pix_of_vec=(np.arange(Pd.shape[0]) // vpp).astype(np.int32))["lnL"]
        err = np.abs(got - want)
        near = np.repeat(want.reshape(n_pix, vpp).max(axis=1), vpp) - want <= 1e3
        rel = err[~near] / np.abs(want[~near])
        i = int(np.argmax(rel))
        print(f"ncomp {ncomp} {np.dtype(dt).name}: near {near.sum()} max |dlnL| {err[near].max():.2e}; far {rel.size}: "
              f"max rel {rel[i]:.2e} at lnL {want[~near][i]:.4g}, p99.9 {np.quantile(rel, 0.999):.2e}, median {np.median(rel):.2e}, "
              f"fraction <= 1e-6: {(rel <= 1e-6).mean():.5f}", flush=True)
