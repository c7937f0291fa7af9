"""NS driver diagnostics: efficiency, iterations, wall time for small problems."""
import sys, time
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import nestfit_b200 as nb
from nestfit_b200.sampler import NestedSamplingBatch
from test_gpu_sampler import gauss_problem, nh3_problem

def report(tag, ns):
    t0 = time.perf_counter(); r = ns.run(); dt = time.perf_counter() - t0
    eff = r['n_iter'] / np.maximum(r['n_evals'], 1)
    print(f"{tag}: runs {ns.n_run} wall {dt:.2f}s lock_iters {r['lock_iters']} n_iter {r['n_iter'].mean():.0f} "
          f"n_evals {r['n_evals'].mean():.0f} eff {eff.mean():.3f} lnZ {r['lnZ'].mean():.2f} +- {r['lnZ'].std():.2f} "
          f"(err {r['lnZ_err'].mean():.2f}) nsamp {r['n_samples'].mean():.0f}", flush=True)
    return r

blk, ut, _ = gauss_problem(nb, n_pix=1)
report("gauss rwalk K=32", NestedSamplingBatch(blk, ut, 1, pix_ids=np.zeros(24, np.int32), nlive=200, tol=0.1, n_prop=32, seed=3, method='rwalk'))
for K in (1, 32):
    report(f"gauss K={K}", NestedSamplingBatch(blk, ut, 1, pix_ids=np.zeros(24, np.int32), nlive=200, tol=0.1, n_prop=K, seed=3))
blk, ut, _ = nh3_problem(nb, 1, n_pix=64)
report("nh3 truth1 fit1", NestedSamplingBatch(blk, ut, 1, nlive=200, tol=0.5, n_prop=32, seed=1))
report("nh3 truth1 fit2", NestedSamplingBatch(blk, ut, 2, nlive=200, tol=0.5, n_prop=32, seed=1))
blk, ut, _ = nh3_problem(nb, 2, n_pix=64, noise=0.2)
report("nh3 truth2 fit1", NestedSamplingBatch(blk, ut, 1, nlive=300, tol=0.5, n_prop=32, seed=1))
report("nh3 truth2 fit2", NestedSamplingBatch(blk, ut, 2, nlive=300, tol=0.5, n_prop=32, seed=1))
report("nh3 truth2 fit2 seed2", NestedSamplingBatch(blk, ut, 2, nlive=300, tol=0.5, n_prop=32, seed=2))
report("nh3 truth2 fit3", NestedSamplingBatch(blk, ut, 3, nlive=300, tol=0.5, n_prop=32, seed=1))
