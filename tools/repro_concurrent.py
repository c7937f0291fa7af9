"""Reproducer for concurrent block fits on one GPU: a noise-gradient cube cut into small blocks, all in flight."""
import sys, time, tempfile
import numpy as np
sys.path.insert(0, '.')
import nestfit_b200 as nb
from nestfit_b200.synth import make_synth_stack
from nestfit_b200.models import ammonia
nx, ny, nblk, nthr = (int(v) for v in sys.argv[1:5]) if len(sys.argv) > 4 else (128, 64, 8, 8)
ncomp_max = int(sys.argv[5]) if len(sys.argv) > 5 else 4
ut = nb.get_irdc_priors()
lon, lat = np.indices((nx, ny))
b = max(1, min(nx, ny) // 4)
ncomp_map = ((lon // b) + (lat // b)) % (ncomp_max + 1)
noise = 0.05 + 0.25 * (lon + lat) / float(nx + ny - 2)
stack = make_synth_stack((nx, ny), ut, ncomp_map=ncomp_map, n_chan=1000, dv=0.07, noise=noise, seed=78)
fitter = nb.CubeFitter(stack, ut, ammonia.AmmoniaRunner, ncomp_max=ncomp_max, lnZ_thresh=11,
                       mn_kwargs={'nlive': 100, 'tol': 1.0, 'efr': 0.3}, nlive_snr_fact=5, n_prop=32, store_posteriors=False)
t0 = time.perf_counter()
if len(sys.argv) > 6 and sys.argv[6] == 'nosink':
    import threading
    blocks = nb.get_block_indices((nx, ny), nblk)
    todo, res, errs, lock = list(range(nblk)), [], [], threading.Lock()

    def worker():
        while True:
            with lock:
                if not todo or errs:
                    return
                j = todo.pop(0)
            try:
                r = fitter.fit_block(blocks[j], device=0)
                r['block'] = j
                with lock:
                    res.append(r)
            except BaseException as exc:
                with lock:
                    errs.append(exc)
                return
    ths = [threading.Thread(target=worker) for _ in range(nthr)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    if errs:
        raise errs[0]
else:
    with tempfile.TemporaryDirectory() as td:
        res = fitter.fit_cube_rank(td + '/s', 0, 1, blocks_per_gpu=nblk, device=0, concurrent_blocks=nthr)
dt = time.perf_counter() - t0
print(f"ok: {nx}x{ny} in {nblk} blocks, {nthr} in flight: {nx * ny / dt:.1f} pixels/s, {dt:.1f} s, blocks {sorted(j for r in res for j in r.get('blocks', [r.get('block')]))}")
