import sys, time
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import nestfit_b200 as nb
from nestfit_b200.sampler import NestedSamplingBatch
from test_gpu_sampler import nh3_problem
blk, ut, _ = nh3_problem(nb, 1, n_pix=4)
ns = NestedSamplingBatch(blk, ut, 2, nlive=200, tol=0.5, n_prop=32, seed=1, max_iter=3000)
r = ns.run(); print(r['n_iter'], r['n_evals'], r['lnZ'])
