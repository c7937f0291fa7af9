import sys, math; sys.path.insert(0,'/tmp/ns')
import numpy as np
from multi import Ell, _logaddexp

def nested_sampling_slice(score, ndim, nlive, active, tol=1.0, n_rep=5, seed=0, w0=1.0, max_iter=10**6, upd=None, batch=16):
    """Nested sampling with slice sampling along random directions of the live set's whitened metric."""
    rng = np.random.default_rng(seed)
    act = np.flatnonzero(active); da = act.size
    U = rng.uniform(size=(nlive, ndim))
    LL = np.asarray(score(U), dtype=np.float64).copy(); LL[~(LL==LL)] = -np.inf
    st = dict(lnZ=-np.inf, H=0.0, lmax=float(LL.max()), it=0, done=False)
    n_evals = nlive
    lnshell = math.log(-math.expm1(-1.0/nlive))
    def insert(u, lc):
        im = int(np.argmin(LL)); mn = float(LL[im])
        if not lc > mn: return False
        lnw = -st['it']/nlive + lnshell; lw = mn + lnw
        new = _logaddexp(st['lnZ'], lw)
        if new > -np.inf:
            t1 = math.exp(lw-new)*mn
            t2 = math.exp(st['lnZ']-new)*(st['H']+st['lnZ']) if st['lnZ']>-np.inf else 0.0
            st['H'] = t1+t2-new
        st['lnZ'] = new; U[im]=u; LL[im]=lc; st['it']+=1; st['lmax']=max(st['lmax'],lc)
        if _logaddexp(st['lnZ'], st['lmax']-st['it']/nlive)-st['lnZ'] < tol or st['it']>=max_iter: st['done']=True
        return True
    def inside(x, lstar):
        if not ((x>0)&(x<1)).all(): return False, -np.inf
        nonlocal n_evals
        n_evals += 1
        l = float(score(x[None])[0])
        return (l > lstar), l
    while not st['done']:
        e1 = Ell(U[:,act], -np.inf, 1.0)
        lstar = float(LL.min())
        # a cohort of `batch` chains from random live points, consumed in order
        ends = []
        for c in range(batch):
            j = rng.integers(0, nlive); x = U[j].copy(); lx = LL[j]
            for rep in range(n_rep):
                z = rng.standard_normal(da); dirn = e1.L @ (z/np.linalg.norm(z))
                w = w0
                r0 = rng.uniform(); lo = -r0*w; hi = (1-r0)*w
                # stepping out
                while True:
                    xx = x.copy(); xx[act] = x[act] + lo*dirn
                    ok, _ = inside(xx, lstar)
                    if not ok: break
                    lo -= w
                while True:
                    xx = x.copy(); xx[act] = x[act] + hi*dirn
                    ok, _ = inside(xx, lstar)
                    if not ok: break
                    hi += w
                # shrink
                while True:
                    t = rng.uniform(lo, hi)
                    xx = x.copy(); xx[act] = x[act] + t*dirn
                    ok, l = inside(xx, lstar)
                    if ok:
                        x = xx; lx = l; break
                    if t < 0: lo = t
                    else: hi = t
            # dummy dims: fresh uniform
            dm = np.ones(ndim, bool); dm[act] = False
            x[dm] = rng.uniform(size=int(dm.sum()))
            ends.append((x, lx))
        for x, lx in ends:
            if st['done']: break
            insert(x, lx)
    lnw_live = -st['it']/nlive - math.log(nlive); lnZ, H = st['lnZ'], st['H']
    for l in LL:
        lw = float(l)+lnw_live; new=_logaddexp(lnZ,lw)
        if new>-np.inf:
            t1 = math.exp(lw-new)*float(l) if l>-np.inf else 0.0
            t2 = math.exp(lnZ-new)*(H+lnZ) if lnZ>-np.inf else 0.0
            H = t1+t2-new
        lnZ=new
    return dict(lnZ=lnZ, lnZ_err=math.sqrt(max(H,0)/nlive), max_loglike=st['lmax'], n_iter=st['it'], n_evals=n_evals)
