"""Static blocks vs queue hand-out in CubeFitter.fit_cube on an unbalanced cube: the first half of
the rows is noise only (one cheap run per pixel), the second half holds three components (three expensive runs), so
one contiguous block per GPU leaves the first GPU idle most of the time.
Usage (2 GPUs): python tools/bench_handout.py [n_lon n_lat]"""
import json
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, '.')
import torch
import nestfit_b200 as nb
from nestfit_b200.models import ammonia
from nestfit_b200.synth import make_synth_stack

def main():
    n_lon = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    n_lat = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    n_gpu = min(2, torch.cuda.device_count())
    ut = nb.get_irdc_priors()
    ncomp_map = np.zeros((n_lon, n_lat), dtype=int)
    ncomp_map[n_lon // 2:] = 3
    stack = make_synth_stack((n_lon, n_lat), ut, ncomp_map=ncomp_map, n_chan=1000, dv=0.07, noise=0.1, seed=5)
    out = {"cube": [n_lon, n_lat], "gpus": n_gpu}
    for k in (1, 8):
        fitter = nb.CubeFitter(stack, ut, ammonia.AmmoniaRunner, ncomp_max=3, mn_kwargs={'nlive': 100, 'tol': 1.0},
                               store_posteriors=False)
        with tempfile.TemporaryDirectory() as tmp:
            t0 = time.perf_counter()
            fitter.fit_cube(tmp + '/cube', nproc=n_gpu, blocks_per_gpu=k)
            out[f"blocks_per_gpu={k}"] = time.perf_counter() - t0
    print(json.dumps(out))


if __name__ == '__main__':      # fit_cube spawns its workers: the entry point must be guarded
    main()
