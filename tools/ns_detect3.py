"""Why are some true 3-component pixels selected as 2?  Compare the best likelihood the 3-component run found
with the likelihood at the truth (sampler failure) and the evidence gain with the threshold (marginal data)."""
import sys
import numpy as np
sys.path.insert(0, '.')
import nestfit_b200 as nb
from nestfit_b200.synth import make_synth_stack
from nestfit_b200.models import ammonia
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
nlive0 = int(sys.argv[2]) if len(sys.argv) > 2 else 100
snr_fact = float(sys.argv[3]) if len(sys.argv) > 3 else 5
walks = int(sys.argv[4]) if len(sys.argv) > 4 else 0
ut = nb.get_irdc_priors()
stack = make_synth_stack((n, n), ut, ncomp_map=np.full((n, n), 3), n_chan=1000, dv=0.07, noise=0.1, seed=11)
idx, T = stack.truths[3]
fitter = nb.CubeFitter(stack, ut, ammonia.AmmoniaRunner, ncomp_max=3, lnZ_thresh=11,
                       mn_kwargs={'nlive': nlive0, 'tol': 1.0, 'efr': 0.3, 'walks': walks}, nlive_snr_fact=snr_fact, n_prop=32,
                       retry_margin=(10.0 if len(sys.argv) > 5 else None))
blocks = nb.get_block_indices((n, n), 1)
res = fitter.fit_block(blocks[0], device=0)
lon, lat = blocks[0]
data, noise, valid = stack.block_arrays(lon, lat)
blk = nb.PixelBlock("ammonia", [c.xarr for c in stack.cubes], data, noise, trans_ids=[1, 2])
lnl_truth = blk.loglike(T, 3, pix_of_vec=idx.astype(np.int32))
nbest = res['nbest']
print("nbest histogram", np.bincount(nbest, minlength=4), "retried", res.get('n_retried'), "rescued", res.get('n_rescued'), "seconds", round(res['seconds'], 1))
bad = np.flatnonzero(nbest[idx] < 3)
dz = res['lnZ'][idx, 3] - res['lnZ'][idx, 2]
gap = lnl_truth - res['max_loglike'][idx, 3]
print(f"3-comp pixels selected < 3: {bad.size} of {idx.size}")
ran3 = np.isfinite(res['lnZ'][idx, 3])
print(f"  of which never tried 3 (2-comp not accepted over 1): {np.sum(~ran3[bad])}")
b = bad[ran3[bad]]
print(f"  tried 3 and rejected: {b.size}; lnZ3-lnZ2 median {np.median(dz[b]):.1f} min {dz[b].min():.1f} max {dz[b].max():.1f}")
print(f"  lnL(truth) - maxL found by the 3-comp run: median {np.median(gap[b]):.1f}, max {gap[b].max():.1f}; "
      f"fraction with gap > 10 (sampler missed the mode): {np.mean(gap[b] > 10):.2f}")
ok = np.flatnonzero((nbest[idx] == 3))
print(f"  for the pixels selected 3: gap median {np.median(gap[ok]):.1f}, max {gap[ok].max():.1f}, frac > 10: {np.mean(gap[ok] > 10):.3f}")
