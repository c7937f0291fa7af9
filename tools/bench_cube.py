"""Cube fit throughput (BASELINE metric M2): pixels/s with full evidence model selection.
  python tools/bench_cube.py --size 64 [--ncomp-max 4 --noise-grad] [--store DIR] [--walks N] [--keep-constant-dims]"""
import argparse, json, shutil, sys, time
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
ap.add_argument('--store', default=None, help='write the store (through fit_cube) into this directory')
ap.add_argument('--no-posteriors', action='store_true')
ap.add_argument('--walks', type=int, default=0)
ap.add_argument('--method', default=None)
ap.add_argument('--wave', type=int, default=16384)
ap.add_argument('--mmodal', action='store_true', help='MultiNest-style ellipsoid decomposition')
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
mn = {'nlive': 100, 'tol': 1.0, 'efr': 0.3}
if args.walks:
    mn['walks'] = args.walks
if args.method:
    mn['method'] = args.method
if args.mmodal:
    mn['mmodal'] = True
fitter = nb.CubeFitter(stack, ut, ammonia.AmmoniaRunner, ncomp_max=args.ncomp_max, lnZ_thresh=11,
                       mn_kwargs=mn, nlive_snr_fact=5, n_prop=args.nprop, max_pixels_per_wave=args.wave,
                       n_streams=args.streams, pixels_per_stream=args.pps, store_posteriors=not args.no_posteriors)
t0 = time.perf_counter()
if args.store:
    shutil.rmtree(args.store, ignore_errors=True)
    res = fitter.fit_cube(args.store + '/cube', nproc=1)[0]
else:
    res = fitter.fit_block(nb.get_block_indices((n, n), 1)[0], device=0, verbose=False)
wall = time.perf_counter() - t0
nb_map = res['nbest'].reshape(n, n)
agree = float((np.minimum(nb_map, args.ncomp_max) == np.minimum(ncomp_map, args.ncomp_max)).mean())
out = {'size': n, 'pixels': n * n, 'seconds': wall, 'pixels_per_s': n * n / wall, 'fit_block_seconds': res['seconds'],
       'n_evals': res['n_evals'], 'evals_per_pixel': res['n_evals'] / (n * n), 'evals_per_s': res['n_evals'] / wall,
       'evals_by_ncomp': res['evals_by_ncomp'].tolist(), 'seconds_by_ncomp': res['seconds_by_ncomp'].tolist(),
       'nbest_agreement': agree, 'build_s': t_build, 'n_truncated': res['n_truncated'],
       'store_seconds': res.get('store_seconds'),
       'nbest_hist': np.bincount(nb_map.ravel() + 1, minlength=args.ncomp_max + 2).tolist()}
if args.store:
    import subprocess
    out['store_bytes'] = int(subprocess.run(['du', '-sb', args.store], capture_output=True, text=True).stdout.split()[0])
print(json.dumps(out))
