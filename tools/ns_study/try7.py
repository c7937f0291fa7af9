import sys; sys.path.insert(0,'/tmp/ns')
from harness import *
from multi import nested_sampling_multi
import multiprocessing as mp
def job(a):
    name, nct, nc, seed, kw = a
    xs, packed, data, noise, T = make_pixel(nct, 100+nct)
    score, cnt = make_score(xs, packed, data, noise, nc)
    active = np.ones(6*nc, bool); active[5*nc:] = False
    r = nested_sampling_multi(score, 6*nc, 300, active, tol=1.0, seed=seed, **kw)
    return name, nct, seed, r['lnZ'], r['lnZ_err'], r['n_iter'], r['n_evals'], r['max_loglike']
if __name__ == '__main__':
    cfgs = [('2c_multi_e2', 2, 2, dict(enlarge=2.0)), ('2c_multi_e4', 2, 2, dict(enlarge=4.0)), ('2c_multi_e8', 2, 2, dict(enlarge=8.0)),
            ('3c_hyb_e4', 3, 3, dict(enlarge=4.0)), ('3c_hyb_e16', 3, 3, dict(enlarge=16.0))]
    jobs = [(n, nct, nc, s, kw) for s in range(8) for n,nct,nc,kw in cfgs]
    res = {}
    with mp.Pool(8) as p:
        for r in p.imap_unordered(job, jobs):
            res.setdefault(r[0], []).append(r[3:]); print(r, flush=True)
    for n,*_ in cfgs:
        a = np.array(res[n])
        print(f"{n:10s} lnZ mean {a[:,0].mean():.3f} sd {a[:,0].std(ddof=1):.3f} (reported err {a[:,1].mean():.3f}) iters {a[:,2].mean():.0f} evals {a[:,3].mean():.0f} lmax min {a[:,4].min():.2f}")
