import sys, math; sys.path.insert(0,'/tmp/ns')
import numpy as np
from scipy.cluster.vq import kmeans2

def lnvol_ball(d): return 0.5*d*math.log(math.pi) - math.lgamma(0.5*d+1.0)

class Ell:
    __slots__ = ('mean','L','lnV','Linv')
    def __init__(self, pts, lnV_min, enlarge):
        d = pts.shape[1]
        self.mean = pts.mean(axis=0)
        cov = np.cov(pts, rowvar=False).reshape(d,d) + 1e-14*np.eye(d)
        L = np.linalg.cholesky(cov)
        y = np.linalg.solve(L, (pts-self.mean).T)
        f = float((y*y).sum(axis=0).max())
        lndet = float(np.log(np.diag(L)).sum())
        lnV = lnvol_ball(d) + 0.5*d*math.log(f) + lndet + math.log(enlarge)
        lnV = max(lnV, lnV_min)
        s = math.exp((lnV - lnvol_ball(d) - lndet)/d)
        self.L = s*L; self.lnV = lnV
        self.Linv = np.linalg.inv(self.L)
    def contains(self, X):
        y = (X - self.mean) @ self.Linv.T
        return (y*y).sum(axis=1) <= 1.0
    def sample(self, rng, n):
        d = self.mean.size
        z = rng.standard_normal((n,d)); r = rng.uniform(size=n)**(1.0/d)/np.sqrt((z*z).sum(axis=1))
        return self.mean + (z*r[:,None]) @ self.L.T

def split(pts, lnpointvol, enlarge, minpts, depth=0):
    n, d = pts.shape
    ell = Ell(pts, math.log(n)+lnpointvol, enlarge)
    if n < 2*minpts or depth > 6: return [ell]
    # 2-means seeded at the ends of the major axis
    w, v = np.linalg.eigh(ell.L @ ell.L.T)
    ax = v[:,-1]*math.sqrt(w[-1])
    try:
        cent, lab = kmeans2(pts, np.stack([ell.mean-0.5*ax, ell.mean+0.5*ax]), iter=10, minit='matrix')
    except Exception:
        return [ell]
    a, b = pts[lab==0], pts[lab==1]
    if a.shape[0] < minpts or b.shape[0] < minpts: return [ell]
    ea = Ell(a, math.log(a.shape[0])+lnpointvol, enlarge); eb = Ell(b, math.log(b.shape[0])+lnpointvol, enlarge)
    lnsum = np.logaddexp(ea.lnV, eb.lnV)
    if lnsum < ell.lnV + math.log(0.5) or ell.lnV > math.log(2.0)+math.log(n)+lnpointvol:
        out = split(a, lnpointvol, enlarge, minpts, depth+1) + split(b, lnpointvol, enlarge, minpts, depth+1)
        lntot = np.logaddexp.reduce([e.lnV for e in out])
        if lntot < ell.lnV: return out
    return [ell]

def _logaddexp(a,b):
    if a==-np.inf: return b
    if b==-np.inf: return a
    m=max(a,b); return m+math.log1p(math.exp(-abs(a-b)))

def nested_sampling_multi(score, ndim, nlive, active, tol=1.0, efr=0.3, n_prop=32, seed=0, upd=None, enlarge=1.2,
                          minpts=None, walks=None, max_iter=10**6, verbose=False, eff_min=None, target_acc=0.5, sc0=0.3):
    rng = np.random.default_rng(seed)
    act = np.flatnonzero(active); da = act.size
    minpts = minpts or (da+1)
    walks = walks or 20+da
    eff_min = eff_min if eff_min is not None else 1.0/(1.2*walks)
    upd = upd or max(1, int(0.1*nlive))
    U = rng.uniform(size=(nlive, ndim))
    LL = np.asarray(score(U), dtype=np.float64).copy(); LL[~(LL==LL)] = -np.inf
    st = dict(lnZ=-np.inf, H=0.0, lmax=float(LL.max()), it=0, done=False)
    n_evals = nlive
    lnshell = math.log(-math.expm1(-1.0/nlive))
    def try_insert(u, lc):
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
    ells=None; last_upd=-10**9; mode=0; ea=ep=0; sc=sc0; hist=[]; sc_hist=[]
    nell_hist=[]
    while not st['done']:
        if mode==0:
            if ells is None or st['it']-last_upd >= upd:
                lnX = -st['it']/nlive
                lnpointvol = lnX - math.log(efr) - math.log(nlive)
                if lnX - math.log(efr) > math.log(0.5):
                    ells = 'cube'
                else:
                    ells = split(U[:,act], lnpointvol, enlarge, minpts)
                    nell_hist.append(len(ells))
                last_upd = st['it']
            if isinstance(ells, str):
                cand = rng.uniform(size=(n_prop, ndim))
            else:
                lnV = np.array([e.lnV for e in ells]); p = np.exp(lnV-lnV.max()); p/=p.sum()
                ks = rng.choice(len(ells), size=n_prop, p=p)
                ca = np.empty((n_prop, da))
                for k in np.unique(ks):
                    m = ks==k; ca[m] = ells[k].sample(rng, int(m.sum()))
                q = np.zeros(n_prop)
                for e in ells: q += e.contains(ca)
                keep = (rng.uniform(size=n_prop)*q < 1.0) & ((ca>0)&(ca<1)).all(axis=1)
                ca = ca[keep]
                cand = rng.uniform(size=(ca.shape[0], ndim)); cand[:,act] = ca
            if cand.shape[0]:
                lc = score(cand); n_evals += cand.shape[0]
                for k in range(cand.shape[0]):
                    if st['done']: break
                    ea += try_insert(cand[k], float(lc[k]))
                ep += cand.shape[0]
            else: ep += 1
            if ep >= 512:
                hist.append((st['it'], ea/ep, 0 if isinstance(ells,str) else len(ells)))
                if ea < ep*eff_min:
                    mode=1; sc=sc0
                ea=ep=0
        else:
            # constrained random walk in the single-ellipsoid metric over the active dims
            e1 = Ell(U[:,act], -np.inf, 1.0)
            K = n_prop
            start = rng.integers(0, nlive, size=K)
            cu, cl = U[start].copy(), LL[start].copy()
            moved = np.zeros(K, bool); lstar=float(LL.min()); acc=0
            for _ in range(walks):
                z = rng.standard_normal((K,da)); r = rng.uniform(size=K)**(1.0/da)/np.sqrt((z*z).sum(axis=1))
                prop = cu.copy(); prop[:,act] = cu[:,act] + sc*((z*r[:,None]) @ e1.L.T)
                ok = ((prop>0)&(prop<1)).all(axis=1)
                if ok.any():
                    lp = np.full(K,-np.inf); lp[ok]=score(prop[ok]); n_evals += int(ok.sum())
                    go = ok & (lp>lstar); cu[go], cl[go] = prop[go], lp[go]; moved |= go; acc += int(go.sum())
            for k in range(K):
                if st['done']: break
                if moved[k]: try_insert(cu[k], float(cl[k]))
            facc = acc/(K*walks); sc = min(max(sc*math.exp((facc-target_acc)/(target_acc*da)),1e-5),2.0); sc_hist.append(sc)
    lnw_live = -st['it']/nlive - math.log(nlive); lnZ, H = st['lnZ'], st['H']
    for l in LL:
        lw = float(l)+lnw_live; new=_logaddexp(lnZ,lw)
        if new>-np.inf:
            t1 = math.exp(lw-new)*float(l) if l>-np.inf else 0.0
            t2 = math.exp(lnZ-new)*(H+lnZ) if lnZ>-np.inf else 0.0
            H = t1+t2-new
        lnZ=new
    return dict(lnZ=lnZ, lnZ_err=math.sqrt(max(H,0)/nlive), max_loglike=st['lmax'], n_iter=st['it'], n_evals=n_evals,
                hist=hist, mode=mode, nell=nell_hist[-5:], sc=sc_hist[::50])
