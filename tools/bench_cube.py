"""Cube fit throughput (BASELINE metric M2): pixels/s with full evidence model selection."""
import argparse, json, sys, time
import numpy as np
sys.path.insert(0, '.')
import nestfit_b200 as nb
from nestfit_b200.synth import make_synth_stack
from nestfit_b200.models import ammonia

ap = argparse.ArgumentParser()
ap.add_argument('--size', type=int, default=64)
ap.add_argument('--ncomp-max', type=int, default=3)
ap.add_argument('--nprop', type=int, default=32)
ap.add_argument('--chan', type=int, default=1000)
ap.add_argument('--noise-grad', action='store_true', help='spatially varying noise 0.05..0.3 K (configs[3])')
ap.add_argument('--streams', type=int, default=1)
ap.add_argument('--pps', type=int, default=1024)
args = ap.parse_args()
ut = nb.get_irdc_priors()
n = args.size
lon, lat = np.indices((n, n))
ncomp_map = ((lon // max(1, n // 4)) + (lat // max(1, n // 4))) % 4        # spatial blocks of 0..3 components
t0 = time.perf_counter()
noise = 0.1
if args.noise_grad:      # smooth gradient across the map (NoiseMap semantics, main.py:39-65)
    noise = 0.05 + 0.25 * (lon + lat) / (2.0 * (n - 1))
stack = make_synth_stack((n, n), ut, ncomp_map=ncomp_map, n_chan=args.chan, dv=0.07, noise=noise, seed=1)
t_build = time.perf_counter() - t0
fitter = nb.CubeFitter(stack, ut, ammonia.AmmoniaRunner, ncomp_max=args.ncomp_max, lnZ_thresh=11,
                       mn_kwargs={'nlive': 100, 'tol': 1.0, 'efr': 0.3}, nlive_snr_fact=5, n_prop=args.nprop,
                       max_pixels_per_wave=16384, n_streams=args.streams, pixels_per_stream=args.pps)
idx = nb.get_block_indices((n, n), 1)[0]
res = fitter.fit_block(idx, device=0, verbose=False)
nb_map = res['nbest'].reshape(n, n)
agree = float((np.minimum(nb_map, args.ncomp_max) == np.minimum(ncomp_map, args.ncomp_max)).mean())
print(json.dumps({'size': n, 'pixels': n * n, 'seconds': res['seconds'], 'pixels_per_s': n * n / res['seconds'],
                  'n_evals': res['n_evals'], 'evals_per_s': res['n_evals'] / res['seconds'],
                  'nbest_agreement': agree, 'build_s': t_build,
                  'nbest_hist': np.bincount(nb_map.ravel() + 1, minlength=args.ncomp_max + 2).tolist()}))
