import sys, time
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import nestfit_b200 as nb
from nestfit_b200.sampler import NestedSamplingBatch
from nestfit_b200.synth import make_synth_stack
ut = nb.get_irdc_priors()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
nc = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mi = int(sys.argv[3]) if len(sys.argv) > 3 else 600
stack = make_synth_stack((n, n), ut, ncomp_map=np.full((n, n), nc), n_chan=1000, dv=0.07, noise=0.1, seed=1)
lon, lat = (a.ravel() for a in np.indices((n, n)))
data, noise, valid = stack.block_arrays(lon, lat)
blk = nb.PixelBlock("ammonia", [c.xarr for c in stack.cubes], data, noise, trans_ids=[1, 2])
ns = NestedSamplingBatch(blk, ut, nc, nlive=250, tol=1.0, n_prop=32, seed=1, max_iter=mi)
t0 = time.perf_counter(); r = ns.run(); dt = time.perf_counter() - t0
print(f"ncomp {nc}: runs {ns.n_run} wall {dt:.2f}s lock {r['lock_iters']} ms/lock {1e3*dt/r['lock_iters']:.3f} evals {r['n_evals'].sum():.3g} evals/s {r['n_evals'].sum()/dt:.3g}", flush=True)
