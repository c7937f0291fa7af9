"""Bias/cost of the constrained random walk against its length: many independent runs of the
same pixel (distinct RNG streams) per `walks` value; prints mean ln Z, scatter, reported error."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import nestfit_b200 as nb
from nestfit_b200.sampler import NestedSamplingBatch
from nestfit_b200.synth import make_synth_stack
ut = nb.get_irdc_priors()
nc = int(sys.argv[1]) if len(sys.argv) > 1 else 3
R = int(sys.argv[2]) if len(sys.argv) > 2 else 128
walks_list = [int(w) for w in sys.argv[3].split(',')] if len(sys.argv) > 3 else [0, 25, 16, 10]
stack = make_synth_stack((2, 2), ut, ncomp_map=np.full((2, 2), nc), n_chan=1000, dv=0.07, noise=0.1, seed=3)
lon, lat = (a.ravel() for a in np.indices((2, 2)))
data, noise, valid = stack.block_arrays(lon, lat)
blk = nb.PixelBlock("ammonia", [c.xarr for c in stack.cubes], data, noise, trans_ids=[1, 2])
pixels = [int(v) for v in sys.argv[4].split(',')] if len(sys.argv) > 4 else [0, 1]
for pix in pixels:
    for w in walks_list:
        ns = NestedSamplingBatch(blk, ut, nc, pix_ids=np.full(R, pix), nlive=300, tol=1.0, n_prop=32, seed=5, walks=w)
        t0 = time.perf_counter(); r = ns.run(); dt = time.perf_counter() - t0
        z = r['lnZ']
        print(f"pix {pix} ncomp {nc} walks {w:3d}: lnZ mean {z.mean():.3f} +- {z.std()/np.sqrt(R):.3f}  scatter {z.std():.3f}  "
              f"reported err {r['lnZ_err'].mean():.3f}  maxL {r['max_loglike'].mean():.2f}  evals/run {r['n_evals'].mean():.3g}  {dt:.1f}s", flush=True)
        ns.close()
