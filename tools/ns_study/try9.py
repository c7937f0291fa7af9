import sys; sys.path.insert(0,'/tmp/ns')
from harness import *
from oracle import ns_port
import multiprocessing as mp
def job(a):
    name, nct, nc, seed, kw = a
    xs, packed, data, noise, T = make_pixel(nct, 100+nct)
    score, cnt = make_score(xs, packed, data, noise, nc)
    r = ns_port.nested_sampling(score, 6*nc, 300, tol=1.0, seed=seed, active=ns_port.active_dims(packed, 6, nc), **kw)
    return name, nct, seed, r['lnZ'], r['lnZ_err'], r['n_iter'], r['n_evals'], r['max_loglike']
if __name__ == '__main__':
    cfgs = [('1c_multi', 1, 1, {}), ('1c_single', 1, 1, dict(multi=False)), ('1c_on3_multi', 3, 1, {}), ('1c_on3_single', 3, 1, dict(multi=False)),
            ('2c_multi', 2, 2, {}), ('3c_multi', 3, 3, {}), ('2c_on3_multi', 3, 2, {}), ('2c_on3_single', 3, 2, dict(multi=False))]
    jobs = [(n, nct, nc, s, kw) for s in range(4) for n,nct,nc,kw in cfgs]
    res = {}
    with mp.Pool(8) as p:
        for r in p.imap_unordered(job, jobs):
            res.setdefault(r[0], []).append(r[3:])
    for n,*_ in cfgs:
        a = np.array(res[n])
        print(f"{n:14s} lnZ mean {a[:,0].mean():.3f} sd {a[:,0].std(ddof=1):.3f} (reported err {a[:,1].mean():.3f}) iters {a[:,2].mean():.0f} evals {a[:,3].mean():.0f} lmax min {a[:,4].min():.2f}")
