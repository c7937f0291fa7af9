import sys; sys.path.insert(0,'/tmp/ns')
from harness import *
from slice import nested_sampling_slice
import multiprocessing as mp
def job(a):
    name, nct, nc, seed, kw = a
    xs, packed, data, noise, T = make_pixel(nct, 100+nct)
    score, cnt = make_score(xs, packed, data, noise, nc)
    active = np.ones(6*nc, bool); active[5*nc:] = False
    r = nested_sampling_slice(score, 6*nc, 300, active, tol=1.0, seed=seed, **kw)
    return name, nct, seed, r['lnZ'], r['lnZ_err'], r['n_iter'], r['n_evals'], r['max_loglike']
if __name__ == '__main__':
    cfgs = [('3c_slice3', 3, 3, dict(n_rep=3)), ('3c_slice6', 3, 3, dict(n_rep=6)), ('3c_slice12', 3, 3, dict(n_rep=12))]
    jobs = [(n, nct, nc, s, kw) for s in range(8) for n,nct,nc,kw in cfgs]
    res = {}
    with mp.Pool(8) as p:
        for r in p.imap_unordered(job, jobs):
            res.setdefault(r[0], []).append(r[3:]); print(r, flush=True)
    for n,*_ in cfgs:
        a = np.array(res[n])
        print(f"{n:10s} lnZ mean {a[:,0].mean():.3f} sd {a[:,0].std(ddof=1):.3f} (reported err {a[:,1].mean():.3f}) iters {a[:,2].mean():.0f} evals {a[:,3].mean():.0f} lmax min {a[:,4].min():.2f}")
