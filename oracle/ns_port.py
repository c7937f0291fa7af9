"""
TEST INFRASTRUCTURE -- not part of the product path.

CPU restatement (numpy) of the batched nested-sampling scheme of
nestfit_b200/csrc/nf_sampler.cu for ONE run, scored with the C oracle likelihood
(oracle/nf_oracle.c).  It stands where the reference has MultiNest
(nestfit/core/core.pyx:727-823 -> MultiNest `run`, cmultinest.pxd:5-33), which is an
external Fortran library absent from /root/reference: "parity unpinned" -- there is no
golden ln Z to compare with, so this port serves two purposes only:
  * an independent implementation the CUDA sampler's ln Z is compared with
    statistically (tests/test_gpu_sampler.py), and
  * the CPU baseline of the cube-fit metric ("the same batched-NS
    algorithm run on CPU with the oracle likelihood"), timed by bench.py.
Only tests/ and bench.py's cpu_baseline leg may import this module.

Scheme (same as the CUDA driver): the cube dimensions the priors overwrite (ConstantPrior rows, the second row
of a DuplicatePrior) are drawn uniformly on their own and stay out of the ellipsoid and the walk metric
(`active_dims`); live set of `nlive` unit-cube points; per step K
candidates drawn uniformly from the bounding ellipsoid of the live set (enlarged to the
larger of 1.2 x the bounding volume and X_i / efr; the unit cube while that volume is
>= 1/2) are consumed in order against the current worst live point; once the windowed
acceptance falls below 1 / (1.2 walks) the run switches to cohorts of K constrained
random walks of `walks` steps in the ellipsoid's metric (step size adapted towards an
acceptance of 1/2).  Dead points carry prior mass X_{i-1} - X_i, X_i = exp(-i / nlive);
termination ln(Z + L_max X_i) - ln Z < tol; the live points are added with X_i / nlive.
"""
import math

import numpy as np


def _logaddexp(a, b):
    if a == -np.inf:
        return b
    if b == -np.inf:
        return a
    m = max(a, b)
    return m + math.log1p(math.exp(-abs(a - b)))


def _unit_ball(rng, n, d):
    z = rng.standard_normal((n, d))
    r = rng.uniform(size=n) ** (1.0 / d) / np.sqrt((z * z).sum(axis=1))
    return z * r[:, None]


def _bound(U, it, nlive, efr):
    """mean, scaled lower Cholesky factor, use_cube flag of the live set's bounding ellipsoid."""
    d = U.shape[1]
    mean = U.mean(axis=0)
    cov = np.cov(U, rowvar=False).reshape(d, d) + 1e-12 * np.eye(d)
    try:
        L = np.linalg.cholesky(cov)
    except np.linalg.LinAlgError:
        return mean, np.eye(d), True
    y = np.linalg.solve(L, (U - mean).T)
    f = float((y * y).sum(axis=0).max())
    lndet = float(np.log(np.diag(L)).sum())
    lnVd = 0.5 * d * math.log(math.pi) - math.lgamma(0.5 * d + 1.0)
    lnV_bound = lnVd + 0.5 * d * math.log(f) + lndet + math.log(1.2) if f > 0 else -np.inf
    lnV = max(lnV_bound, -it / nlive - math.log(efr))
    scale = math.exp((lnV - lnVd - lndet) / d)
    use_cube = lnV > math.log(0.5) or not (f > 0) or not math.isfinite(scale)
    return mean, scale * L, use_cube


def active_dims(packed_priors, n_model, ncomp):
    """Indices of the cube dimensions the likelihood depends on (nf_ns_create, nf_sampler.cu): rows written
    by a ConstantPrior and the duplicate row of a DuplicatePrior are left out unless another prior writes them."""
    pp, n_p = packed_priors[0], packed_priors[1]
    dummy = np.zeros(n_model * ncomp, dtype=bool)
    for k in range(n_p):
        row = pp[k].p_ix if pp[k].kind == 1 else (pp[k].p_ix2 if pp[k].kind == 2 else -1)
        if 0 <= row < n_model:
            dummy[row * ncomp:(row + 1) * ncomp] = True
    for k in range(n_p):
        if pp[k].kind != 1:
            dummy[pp[k].p_ix * ncomp:(pp[k].p_ix + 1) * ncomp] = False
        if pp[k].kind in (6, 7) and pp[pp[k].nested].kind != 1:
            q = pp[pp[k].nested]
            dummy[q.p_ix * ncomp:(q.p_ix + 1) * ncomp] = False
    act = np.flatnonzero(~dummy)
    return act if act.size else np.arange(n_model * ncomp)


NS_MAX_ELL = 8          # ellipsoids per run (nf_sampler.cu)
NS_KMEANS_ITERS = 8


def _ellipsoid(P, lnV_min):
    """(mean, scaled lower Cholesky factor, ln V, index of the point of largest Mahalanobis radius) of the points P:
    the covariance ellipsoid through the farthest point, enlarged to 1.2 x its volume and to at least lnV_min."""
    n, d = P.shape
    mean = P.mean(axis=0)
    cov = (np.cov(P, rowvar=False).reshape(d, d) if n > 1 else np.zeros((d, d))) + 1e-12 * np.eye(d)
    L = np.linalg.cholesky(cov)
    y = np.linalg.solve(L, (P - mean).T)
    r2 = (y * y).sum(axis=0)
    far = int(np.argmax(r2))
    f = max(float(r2[far]), 1e-300)
    lndet = float(np.log(np.diag(L)).sum())
    lnVd = 0.5 * d * math.log(math.pi) - math.lgamma(0.5 * d + 1.0)
    lnV = max(lnVd + 0.5 * d * math.log(f) + lndet + math.log(1.2), lnV_min)
    return mean, math.exp((lnV - lnVd - lndet) / d) * L, lnV, far


def multi_ellipsoid_bound(U, lnX, nlive, efr):
    """MultiNest-style decomposition of the live set into at most NS_MAX_ELL ellipsoids (ns_bounds_kernel):
    breadth-first 2-means splits -- seeded at the point of largest Mahalanobis radius and the point farthest from
    it, distances scaled by the parent's per-dimension spread -- kept when the children's volumes add up to less
    than half the parent's, or when the parent is more than twice its share X n_c / (efr nlive) of the prior
    volume and the children are smaller at all.  Returns a list of (mean, L, lnV)."""
    lnpv = lnX - math.log(efr) - math.log(nlive)
    d = U.shape[1]
    minpts = d + 1
    todo, done = [np.arange(U.shape[0])], []
    while todo:
        idx = todo.pop(0)
        P = U[idx]
        mean, L, lnV, far = _ellipsoid(P, math.log(idx.size) + lnpv)
        if idx.size < 2 * minpts or len(todo) + len(done) + 2 > NS_MAX_ELL:
            done.append((mean, L, lnV))
            continue
        w = 1.0 / np.maximum(P.std(axis=0), 1e-300)
        c0 = P[far]
        c1 = P[int(np.argmax((((P - c0) * w) ** 2).sum(axis=1)))]
        for _ in range(NS_KMEANS_ITERS):
            lab = (((P - c1) * w) ** 2).sum(axis=1) < (((P - c0) * w) ** 2).sum(axis=1)
            if lab.all() or not lab.any():
                break
            c0, c1 = P[~lab].mean(axis=0), P[lab].mean(axis=0)
        na, nb = int((~lab).sum()), int(lab.sum())
        if na < minpts or nb < minpts:
            done.append((mean, L, lnV))
            continue
        va = _ellipsoid(P[~lab], math.log(na) + lnpv)[2]
        vb = _ellipsoid(P[lab], math.log(nb) + lnpv)[2]
        lnsum = np.logaddexp(va, vb)
        if lnsum < lnV + math.log(0.5) or (lnV > math.log(2.0) + math.log(idx.size) + lnpv and lnsum < lnV):
            todo += [idx[~lab], idx[lab]]
        else:
            done.append((mean, L, lnV))
    return done


def sample_union(rng, ells, n):
    """n points uniform in the union of the ellipsoids (pick one by volume, draw, keep with probability
    1 / number of ellipsoids that contain the point)."""
    lnV = np.array([e[2] for e in ells])
    p = np.exp(lnV - lnV.max())
    p /= p.sum()
    d = ells[0][0].size
    ks = rng.choice(len(ells), size=n, p=p)
    X = np.empty((n, d))
    for k in np.unique(ks):
        m = ks == k
        X[m] = ells[k][0] + _unit_ball(rng, int(m.sum()), d) @ ells[k][1].T
    q = np.zeros(n)
    for mean, L, _ in ells:
        y = np.linalg.solve(L, (X - mean).T)
        q += (y * y).sum(axis=0) <= 1.0
    return X[rng.uniform(size=n) * q < 1.0]


def nested_sampling(score, ndim, nlive, tol=1.0, efr=0.3, n_prop=32, walks=None, seed=0, max_iter=1_000_000,
                    rwalk=False, active=None, return_samples=False, multi=False, update_every=8):
    """score(U[B, ndim]) -> lnL[B] (prior transform inside; NaN = not acceptable).
    `rwalk=True` starts with the random walk (the CUDA driver's method='rwalk'); `active`: the dimensions
    inside the ellipsoid / walk metric (default all); `multi`: the MultiNest-style decomposition (the CUDA driver's
    mmodal=True) instead of one ellipsoid rebuilt every step.
    Returns dict(lnZ, lnZ_err, max_loglike, n_iter, n_evals, n_samples); with `return_samples` also the dead
    points and final live points in death order: samples_u [n, ndim], samples_lnL [n], samples_lnw [n]
    (ln of the prior-mass weight; posterior weight = exp(lnL + lnw - lnZ))."""
    rng = np.random.default_rng(seed)
    K = n_prop
    act = np.arange(ndim) if active is None else np.asarray(active)
    d = act.size
    walks = walks or 20 + d
    U = rng.uniform(size=(nlive, ndim))
    LL = np.asarray(score(U), dtype=np.float64).copy()
    LL[~(LL == LL)] = -np.inf
    st = dict(lnZ=-np.inf, H=0.0, lmax=float(LL.max()), it=0, nd=0, done=False)
    n_evals = nlive
    lnshell = math.log(-math.expm1(-1.0 / nlive))
    dead = []

    def try_insert(u, lc):
        im = int(np.argmin(LL))
        mn = float(LL[im])
        if not lc > mn:
            return False
        lnw = -st['it'] / nlive + lnshell
        lw = mn + lnw
        new = _logaddexp(st['lnZ'], lw)
        if new > -np.inf:
            t1 = math.exp(lw - new) * mn
            t2 = math.exp(st['lnZ'] - new) * (st['H'] + st['lnZ']) if st['lnZ'] > -np.inf else 0.0
            st['H'] = t1 + t2 - new
        st['lnZ'] = new
        st['nd'] += 1
        if return_samples:
            dead.append((U[im].copy(), mn, lnw))
        U[im] = u
        LL[im] = lc
        st['it'] += 1
        st['lmax'] = max(st['lmax'], lc)
        if _logaddexp(st['lnZ'], st['lmax'] - st['it'] / nlive) - st['lnZ'] < tol or st['it'] >= max_iter:
            st['done'] = True
        return True

    mode, ea, ep, sc = (1 if rwalk else 0), 0, 0, 0.3
    step, ells = 0, None
    while not st['done']:
        if mode == 0:
            lnX = -st['it'] / nlive
            use_cube = lnX - math.log(efr) > math.log(0.5)
            if multi and not use_cube and (ells is None or step % update_every == 0):
                # the decomposition is rebuilt every `update_every` lock-steps; in between the ellipsoids stay valid
                # (the constrained region only shrinks)
                ells = multi_ellipsoid_bound(U[:, act], lnX, nlive, efr)
            if not multi:
                mean, B, use_cube = _bound(U[:, act], st['it'], nlive, efr)
                ells = [(mean, B, 0.0)]
            step += 1
            if use_cube:
                cand = rng.uniform(size=(K, ndim))
            else:
                ca = sample_union(rng, ells, K)
                ca = ca[((ca > 0.0) & (ca < 1.0)).all(axis=1)]
                cand = rng.uniform(size=(ca.shape[0], ndim))
                cand[:, act] = ca
            if cand.shape[0]:
                lc = score(cand)
                n_evals += cand.shape[0]
                for k in range(cand.shape[0]):
                    if st['done']:
                        break
                    ea += try_insert(cand[k], float(lc[k]))
                ep += cand.shape[0]
            else:
                ep += 1
            if ep >= 512:
                if ea < ep / (1.2 * walks):
                    mode, sc = 1, 0.3
                ea = ep = 0
        else:
            mean, B, _ = _bound(U[:, act], st['it'], nlive, efr)
            start = rng.integers(0, nlive, size=K)
            cu, cl = U[start].copy(), LL[start].copy()
            moved = np.zeros(K, dtype=bool)
            lstar = float(LL.min())
            acc = 0
            for _ in range(walks):
                prop = rng.uniform(size=(K, ndim))
                prop[:, act] = cu[:, act] + sc * (_unit_ball(rng, K, d) @ B.T)
                ok = ((prop > 0.0) & (prop < 1.0)).all(axis=1)
                if ok.any():
                    lp = np.full(K, -np.inf)
                    lp[ok] = score(prop[ok])
                    n_evals += int(ok.sum())
                    go = ok & (lp > lstar)
                    cu[go], cl[go] = prop[go], lp[go]
                    moved |= go
                    acc += int(go.sum())
            for k in range(K):
                if st['done']:
                    break
                if moved[k]:
                    try_insert(cu[k], float(cl[k]))
            facc = acc / (K * walks)
            sc = min(max(sc * math.exp((facc - 0.5) / (0.5 * d)), 1e-5), 2.0)
    # the remaining live points
    lnw_live = -st['it'] / nlive - math.log(nlive)
    lnZ, H = st['lnZ'], st['H']
    for l in LL:
        lw = float(l) + lnw_live
        new = _logaddexp(lnZ, lw)
        if new > -np.inf:
            t1 = math.exp(lw - new) * float(l) if l > -np.inf else 0.0
            t2 = math.exp(lnZ - new) * (H + lnZ) if lnZ > -np.inf else 0.0
            H = t1 + t2 - new
        lnZ = new
    out = dict(lnZ=lnZ, lnZ_err=math.sqrt(max(H, 0.0) / nlive), max_loglike=st['lmax'], n_iter=st['it'],
               n_evals=n_evals, n_samples=st['nd'] + nlive)
    if return_samples:
        dead += [(U[p].copy(), float(LL[p]), lnw_live) for p in range(nlive)]
        out.update(samples_u=np.array([d[0] for d in dead]), samples_lnL=np.array([d[1] for d in dead]),
                   samples_lnw=np.array([d[2] for d in dead]))
    return out


def fit_pixel(xarrs, trans_ids, data, noise, packed_priors, ncomp_max=3, lnZ_thresh=11.0, nlive=100,
              nlive_snr_fact=5, tol=1.0, efr=0.3, n_prop=32, seed=0):
    """One pixel through the reference's ncomp escalation (nestfit/main.py:436-472): nlive grows with
    the peak SNR (main.py:445-447); N -> N + 1 while ln Z_N - ln Z_{N-1} >= lnZ_thresh, starting from the
    null evidence.  data [n_spec, n_chan], noise [n_spec].  Returns dict(nbest, lnZ[], n_evals)."""
    from . import oracle as orc
    data = np.asarray(data, dtype=np.float64)
    noise = np.asarray(noise, dtype=np.float64)
    d3, n2 = data[None], noise[None]
    null = float(-(data * data / (2.0 * noise[:, None] ** 2)).sum())
    nl = int(nlive + int(nlive_snr_fact * max(float((data.max(axis=1) / noise).max()), 0.0)))
    lnZ, old, nbest, n_evals = [null], null, 0, 0
    for ncomp in range(1, ncomp_max + 1):
        def score(U, ncomp=ncomp):
            th = orc.prior_transform(packed_priors, U, ncomp)
            out = orc.nh3_batch(xarrs, trans_ids, np.nan_to_num(th, nan=1.0), ncomp, data=d3, noise=n2)["lnL"]
            out[~np.isfinite(th).all(axis=1)] = np.nan
            return out
        res = nested_sampling(score, 6 * ncomp, nl, tol=tol, efr=efr, n_prop=n_prop, seed=seed + 7919 * ncomp,
                              active=active_dims(packed_priors, 6, ncomp))
        n_evals += res['n_evals']
        lnZ.append(res['lnZ'])
        if res['lnZ'] - old >= lnZ_thresh:
            old, nbest = res['lnZ'], ncomp
        else:
            break
    return dict(nbest=nbest, lnZ=lnZ, n_evals=n_evals, nlive=nl)
