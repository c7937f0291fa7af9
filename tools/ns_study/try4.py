import sys; sys.path.insert(0,'/tmp/ns')
from harness import *
from multi import nested_sampling_multi
import multiprocessing as mp
def job(a):
    name, pixseed, seed, kw = a
    xs, packed, data, noise, T = make_pixel(3, pixseed)
    score, cnt = make_score(xs, packed, data, noise, 3)
    active = np.ones(18, bool); active[15:] = False
    r = nested_sampling_multi(score, 18, 300, active, tol=1.0, seed=seed, **kw)
    return name, pixseed, seed, r['lnZ'], r['lnZ_err'], r['n_iter'], r['n_evals'], r['max_loglike']
if __name__ == '__main__':
    cfgs = [('walk70', dict(eff_min=1.0, walks=70)), ('walk140', dict(eff_min=1.0, walks=140))]
    pix = 103
    jobs = [(n, pix, s, kw) for s in range(8) for n,kw in cfgs]
    res = {}
    with mp.Pool(8) as p:
        for r in p.imap_unordered(job, jobs):
            res.setdefault(r[0], []).append(r[3:]); print(r, flush=True)
    for n,_ in cfgs:
        a = np.array(res[n])
        print(f"{n:8s} lnZ mean {a[:,0].mean():.3f} sd {a[:,0].std(ddof=1):.3f} (reported err {a[:,1].mean():.3f}) iters {a[:,2].mean():.0f} evals {a[:,3].mean():.0f} lmax min {a[:,4].min():.2f}")
