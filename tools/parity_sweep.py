"""Large randomized parity sweep of the fused NH3 kernel against the C oracle (test infrastructure):
lnL error relative to the tolerance 1e-3 + 2e-6 |lnL| and spectrum error relative to 1e-5 of the peak."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import nestfit_b200 as nb
from oracle import oracle as orc
rng = np.random.default_rng(2026)
ut = nb.get_irdc_priors()
n_pix = 32
for ncomp in (1, 2, 3, 4):
    for n_chan, dv in ((1000, 0.07), (380, 0.158)):
        xs = [orc.bench_axis(1, n_chan, dv), orc.bench_axis(2, n_chan, dv)]
        B = 16384
        P = orc.prior_transform(ut.pack(), rng.uniform(size=(B + 2000, 6 * ncomp)), ncomp)
        P = P[np.isfinite(P).all(axis=1)][:B]
        truth = P[:n_pix]
        clean = orc.nh3_batch(xs, [1, 2], truth, ncomp, want_pred=True)["pred"]
        noise = rng.uniform(0.05, 0.3, (n_pix, 2))
        data = (clean + rng.normal(size=clean.shape) * noise[:, :, None]).astype(np.float32)
        blk = nb.PixelBlock("ammonia", xs, data, noise, trans_ids=[1, 2])
        pix = rng.integers(0, n_pix, B).astype(np.int32)
        for dt in (np.float32, np.float64):
            Pd = P.astype(dt)
            want = orc.nh3_batch(xs, [1, 2], Pd.astype(np.float64), ncomp, data=data.astype(np.float64), noise=noise,
                                 pix_of_vec=pix, want_pred=True)
            got = blk.loglike(Pd, ncomp, pix_of_vec=pix)
            err = np.abs(got - want["lnL"]); lim = 1e-3 + 2e-6 * np.abs(want["lnL"])
            pr = blk.predict(Pd[:4096], ncomp)
            peak = np.maximum(np.abs(want["pred"][:4096]).max(axis=-1, keepdims=True), 1e-30)
            serr = (np.abs(pr - want["pred"][:4096]) / peak).max()
            print(f"ncomp {ncomp} n_chan {n_chan} {dt.__name__}: lnL err/tol max {np.max(err / lim):.3f} p99.9 {np.quantile(err / lim, 0.999):.3f} "
                  f"median {np.median(err / lim):.4f}; max abs err {err.max():.3g}; spectra max err/peak {serr:.2e} (tol 1e-5)", flush=True)
        blk.close()
