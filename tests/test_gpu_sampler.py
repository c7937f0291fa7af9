"""GPU tests of the batched nested-sampling driver (nf_ns_*): evidence against
brute-force quadrature of the same likelihood, invariance of ln Z to the batching
factor K (K = 1 is plain sequential nested sampling), seed-to-seed scatter against
the reported error, posterior recovery and the reference's result contract."""
import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def flat_dist(nb, lo, hi, size=200):
    x = np.linspace(lo, hi, size)
    return nb.Distribution(x, np.ones(size) / size)


def gauss_problem(nb, seed=3, n_pix=1):
    rng = np.random.default_rng(seed)
    n_chan = 256
    v = (np.arange(n_chan) - 0.5 * (n_chan - 1)) * 0.3
    x = np.sort(orc.NU[0] * (1 - v / orc.CKMS))
    truth = np.array([1.0, 1.2, 2.0])
    clean = orc.gauss_batch(x, orc.NU[0], truth[None], 1, want_pred=True)["pred"][0]
    data = (clean[None] + rng.normal(0, 0.2, (n_pix, n_chan))).astype(np.float32)
    blk = nb.PixelBlock("gaussian", [x], data[:, None, :], 0.2, rest_freq=orc.NU[0])
    ut = nb.PriorTransformer(np.array([nb.Prior(flat_dist(nb, -10, 10), 0), nb.Prior(flat_dist(nb, 0.3, 3.0), 1),
                                       nb.Prior(flat_dist(nb, 0.0, 5.0), 2)], dtype=object))
    return blk, ut, truth


def quadrature_lnz(blk, ut, pix, center_u, half_u, n=96):
    """ln of the integral of L over the unit cube, midpoint rule on a sub-box that
    holds all of the posterior mass (checked by the caller through the edges)."""
    lo = np.clip(center_u - half_u, 0, 1)
    hi = np.clip(center_u + half_u, 0, 1)
    axes = [lo[k] + (np.arange(n) + 0.5) / n * (hi[k] - lo[k]) for k in range(3)]
    U = np.stack(np.meshgrid(*axes, indexing="ij"), axis=-1).reshape(-1, 3)
    P = ut.transform_batch(np.ascontiguousarray(U), 1)
    lnl = blk.loglike(P, 1, pix_of_vec=np.full(P.shape[0], pix, dtype=np.int32))
    m = lnl.max()
    lnz = m + np.log(np.exp(lnl - m).sum()) + np.log(np.prod(hi - lo) / U.shape[0])
    cube = lnl.reshape(n, n, n)
    edge = max(cube[0].max(), cube[-1].max(), cube[:, 0].max(), cube[:, -1].max(), cube[:, :, 0].max(),
               cube[:, :, -1].max())
    return lnz, m - edge


def test_evidence_matches_quadrature(nb):
    from nestfit_b200.sampler import NestedSamplingBatch
    blk, ut, truth = gauss_problem(nb, n_pix=4)
    ns = NestedSamplingBatch(blk, ut, 1, nlive=400, tol=0.1, efr=0.3, n_prop=32, seed=11)
    res = ns.run()
    assert np.isfinite(res["lnZ"]).all() and not res["truncated"].any()
    for p in range(4):
        post = ns.posterior(p)
        w = post[:, -1]
        assert abs(w.sum() - 1) < 1e-6
        mean = (post[:, :3] * w[:, None]).sum(0)
        std = np.sqrt((w[:, None] * (post[:, :3] - mean) ** 2).sum(0))
        # unit-cube coordinates of the posterior bulk (flat priors: linear maps)
        lo, span = np.array([-10, 0.3, 0.0]), np.array([20, 2.7, 5.0])
        cu, hu = (mean - lo) / span, 9 * std / span
        want, margin = quadrature_lnz(blk, ut, p, cu, hu)
        assert margin > 20, "quadrature box does not contain the posterior mass"
        err = max(res["lnZ_err"][p], 0.03)
        assert abs(res["lnZ"][p] - want) < 4 * err + 0.05, (res["lnZ"][p], want, err)
        assert np.all(np.abs(mean - truth) < 6 * std + 0.02)
    # reported evidence beats the null model decisively (SNR ~ 10 line)
    assert np.all(res["lnZ"] - blk.null_lnZ() > 11)


def test_batching_factor_invariance_and_seed_scatter(nb):
    from nestfit_b200.sampler import NestedSamplingBatch
    blk, ut, _ = gauss_problem(nb, n_pix=1)
    pix = np.zeros(24, dtype=np.int32)          # the same pixel fitted 24 times per configuration
    out = {}
    for K in (1, 8, 64):
        ns = NestedSamplingBatch(blk, ut, 1, pix_ids=pix, nlive=200, tol=0.1, n_prop=K, seed=100 + K)
        r = ns.run()
        out[K] = (r["lnZ"].copy(), r["lnZ_err"].copy(), r["n_evals"].mean())
        ns.close()
    means = {K: v[0].mean() for K, v in out.items()}
    for K, (lnz, err, _) in out.items():
        # scatter between independent runs is what lnZ_err claims (within a factor ~2)
        assert 0.4 * err.mean() < lnz.std(ddof=1) < 2.5 * err.mean(), (K, lnz.std(ddof=1), err.mean())
        sem = lnz.std(ddof=1) / np.sqrt(lnz.size)
        assert abs(means[K] - means[1]) < 4 * np.hypot(sem, out[1][0].std(ddof=1) / np.sqrt(24)) + 0.02
    # identical seed -> identical result (counter-based RNG, deterministic order)
    a = NestedSamplingBatch(blk, ut, 1, pix_ids=pix[:3], nlive=100, n_prop=16, seed=5).run()
    b = NestedSamplingBatch(blk, ut, 1, pix_ids=pix[:3], nlive=100, n_prop=16, seed=5).run()
    assert np.array_equal(a["lnZ"], b["lnZ"]) and np.array_equal(a["n_evals"], b["n_evals"])
    # different runs of one batch are independent streams
    assert len(np.unique(a["lnZ"])) == 3


def nh3_problem(nb, ncomp_true, n_pix, seed=21, noise=0.1):
    rng = np.random.default_rng(seed)
    ut = nb.get_irdc_priors()
    xs = [orc.bench_axis(1, 400, 0.158), orc.bench_axis(2, 400, 0.158)]
    truth = {1: np.array([0.5, 14.0, 6.0, 14.6, 0.45, 0.0]),
             2: np.array([-1.0, 1.5, 10, 15, 4, 6, 14.5, 15, 0.3, 0.6, 0, 0], dtype=float)}[ncomp_true]
    clean = orc.nh3_batch(xs, [1, 2], truth[None], ncomp_true, want_pred=True)["pred"][0]
    data = (clean[None] + rng.normal(0, noise, (n_pix,) + clean.shape)).astype(np.float32)
    blk = nb.PixelBlock("ammonia", xs, data, noise, trans_ids=[1, 2])
    return blk, ut, truth


def test_nh3_model_selection_and_recovery(nb):
    from nestfit_b200.sampler import NestedSamplingBatch
    blk, ut, truth = nh3_problem(nb, 1, n_pix=6)
    null = blk.null_lnZ()
    r1 = NestedSamplingBatch(blk, ut, 1, nlive=200, tol=0.5, n_prop=32, seed=1)
    res1 = r1.run()
    r2 = NestedSamplingBatch(blk, ut, 2, nlive=200, tol=0.5, n_prop=32, seed=2)
    res2 = r2.run()
    assert np.all(res1["lnZ"] - null > 11)               # one component is strongly detected
    assert np.all(res2["lnZ"] - res1["lnZ"] < 11)        # ... and a second one is not (main.py:464)
    for p in range(6):
        post = r1.posterior(p)
        w = post[:, -1]
        mean = (post[:, :6] * w[:, None]).sum(0)
        std = np.sqrt((w[:, None] * (post[:, :6] - mean) ** 2).sum(0))
        for k in (0, 4):                                   # velocity and line width are well constrained
            assert abs(mean[k] - truth[k]) < 5 * std[k] + 0.01
        # the best-fit vector really is the maximum-likelihood sample, evaluated by the fused kernel
        lnl = blk.loglike(res1["bestfit"][p][None], 1, pix_of_vec=np.array([p], dtype=np.int32))[0]
        assert abs(lnl - res1["max_loglike"][p]) < 0.05 + 1e-5 * abs(lnl)
    # max likelihood close to the chi-square expectation of a good fit: -N/2 +- few sqrt(N/2)
    assert np.all(np.abs(res1["max_loglike"] + 400) < 6 * np.sqrt(400) + 10)


def test_two_component_fit_prefers_two(nb):
    from nestfit_b200.sampler import NestedSamplingBatch
    blk, ut, truth = nh3_problem(nb, 2, n_pix=3, noise=0.2)
    z = {}
    for nc in (1, 2):
        res = NestedSamplingBatch(blk, ut, nc, nlive=300, tol=0.5, n_prop=32, seed=7 + nc).run()
        z[nc] = res["lnZ"]
        assert np.isfinite(res["lnZ"]).all()
    assert np.all(z[2] - z[1] > 11)                      # the second component is required


def test_run_multinest_dropin_contract(nb):
    """run_multinest(runner, dumper) + the reference's per-run products (core.pyx:645-687)."""
    from nestfit_b200.sampler import run_multinest, Dumper, MARG_COLS
    from nestfit_b200.store import MemGroup
    rng = np.random.default_rng(2)
    xs = [orc.bench_axis(1, 380, 0.158), orc.bench_axis(2, 380, 0.158)]
    truth = np.array([0.0, 12.0, 5.0, 14.4, 0.5, 0.0])
    clean = orc.nh3_batch(xs, [1, 2], truth[None], 1, want_pred=True)["pred"][0]
    data = clean + rng.normal(0, 0.2, clean.shape)
    spec_data = [[xs[0], data[0], 0.2, 1], [xs[1], data[1], 0.2, 2]]
    runner = nb.AmmoniaRunner.from_data(spec_data, nb.get_irdc_priors(), ncomp=1)
    group = MemGroup("/pix/0/0/1")
    dumper = Dumper(group)
    run_multinest(runner, dumper, nlive=100, seed=5, tol=1.0, efr=0.3, updInt=2000)
    assert np.isfinite(runner.run_lnZ) and runner.run_lnZ == group.attrs['global_lnZ']
    for key in ('ncomp', 'null_lnZ', 'n_chan_tot', 'n_samples', 'n_live', 'n_params', 'global_lnZ',
                'global_lnZ_err', 'max_loglike', 'marg_cols', 'marg_quantiles', 'BIC', 'AIC', 'AICc',
                'null_BIC', 'null_AIC', 'null_AICc'):
        assert key in group.attrs
    assert group.attrs['n_chan_tot'] == 760 and group.attrs['n_params'] == 6 and group.attrs['n_live'] == 100
    assert list(group.attrs['marg_cols']) == MARG_COLS
    n = group.attrs['n_samples']
    assert group['posteriors'].shape == (n, 8) and group['posteriors'].dtype == np.float32
    assert group['marginals'].shape == (15, 6)
    assert group['bestfit_params'].shape == (6,) and group['map_params'].shape == (6,)
    k, nn, maxL = 6.0, 760.0, group.attrs['max_loglike']
    assert group.attrs['BIC'] == pytest.approx(np.log(nn) * k - 2 * maxL)
    assert group.attrs['AICc'] == pytest.approx(2 * k - 2 * maxL + (2 * k**2 + 2 * k) / (nn - k - 1))


def test_consecutive_samplers_and_max_loglike_sanity(nb):
    """Several samplers in one process, many copies of one pixel (so each run gets the large
    per-step proposal counts of a wave's tail from the start): every run must terminate, and
    no run may report a likelihood above what the truth itself reaches plus the parameter
    count (a run whose random-walk chains were started from stale memory reported lnL = 0
    and never finished)."""
    from nestfit_b200.sampler import NestedSamplingBatch
    blk, ut, truth = nh3_problem(nb, 2, n_pix=2, noise=0.15)
    lnl_truth = blk.loglike(np.repeat(truth[None], 2, axis=0), 2, pix_of_vec=np.array([0, 1], dtype=np.int32))
    for rep in range(3):
        pix = np.full(24, rep % 2, dtype=np.int32)
        ns = NestedSamplingBatch(blk, ut, 2, pix_ids=pix, nlive=150, tol=1.0, n_prop=32, seed=40 + rep, max_iter=40000)
        res = ns.run()
        assert not res["truncated"].any()
        assert np.isfinite(res["lnZ"]).all()
        assert np.all(res["max_loglike"] < lnl_truth[rep % 2] + 12 + 20)
        assert np.all(res["max_loglike"] > lnl_truth[rep % 2] - 60)
        # best-fit vectors are finite physical parameters
        assert np.isfinite(res["bestfit"]).all()
        ns.close()


def test_lnz_agrees_with_cpu_port(nb):
    """The CUDA driver against the independent numpy port of the same scheme (oracle/ns_port.py) scored
    with the C oracle likelihood, on the same pixel: the evidences agree within their scatter.  (MultiNest
    is absent: this and the quadrature test are what pins ln Z.)"""
    from nestfit_b200.sampler import NestedSamplingBatch
    from oracle import ns_port
    blk, ut, truth = nh3_problem(nb, 1, n_pix=1, seed=33)
    xs = [orc.bench_axis(1, 400, 0.158), orc.bench_axis(2, 400, 0.158)]
    # the pixel nh3_problem uploaded (same seed, same draw order), as the FP32 values the GPU holds
    clean = orc.nh3_batch(xs, [1, 2], truth[None], 1, want_pred=True)["pred"][0]
    data = (clean[None] + np.random.default_rng(33).normal(0, 0.1, (1,) + clean.shape)).astype(np.float32)[0]
    data = data.astype(np.float64)
    lnl_gpu = blk.loglike(truth[None], 1)[0]
    lnl_cpu = orc.nh3_batch(xs, [1, 2], truth[None], 1, data=data[None], noise=np.full((1, 2), 0.1))["lnL"][0]
    assert abs(lnl_gpu - lnl_cpu) < 1e-3 + 2e-6 * abs(lnl_cpu)       # same pixel on both sides
    packed = ut.pack()

    def score(U):
        th = orc.prior_transform(packed, U, 1)
        out = orc.nh3_batch(xs, [1, 2], np.nan_to_num(th, nan=1.0), 1, data=data[None], noise=np.full((1, 2), 0.1))["lnL"]
        out[~np.isfinite(th).all(axis=1)] = np.nan
        return out
    cpu = [ns_port.nested_sampling(score, 6, 200, tol=0.5, seed=s, active=ns_port.active_dims(packed, 6, 1)) for s in range(4)]
    # eight independent GPU runs of the same pixel in one batch (distinct Philox streams per run)
    ns = NestedSamplingBatch(blk, ut, 1, pix_ids=np.zeros(8, dtype=np.int32), nlive=200, tol=0.5, n_prop=32, seed=5)
    res = ns.run()
    ns.close()
    zc, zg = np.array([c['lnZ'] for c in cpu]), res["lnZ"]
    err = np.sqrt(np.mean([c['lnZ_err'] for c in cpu]) ** 2 / 4 + np.mean(res["lnZ_err"]) ** 2 / 8)
    assert abs(zc.mean() - zg.mean()) < 4 * err + 0.1, (zc, zg, err)
    # the same maximum of the likelihood (FP32 kernel vs FP64 oracle on slightly different best samples)
    assert abs(max(c['max_loglike'] for c in cpu) - res["max_loglike"].max()) < 1.0
    # and the same cost per iteration within a factor (same proposal scheme)
    rc = np.mean([c['n_evals'] / c['n_iter'] for c in cpu])
    rg = np.mean(res["n_evals"] / res["n_iter"])
    assert 0.5 < rc / rg < 2.0, (rc, rg)


def test_products_all_matches_per_run(nb):
    """The batched products pass (device-side packing + column sort, one device-to-host copy; nf_ns_products) gives
    what the per-run path gives: the float32 `posteriors` rows in death order and numpy.quantile `marginals`
    (core.pyx:596-598,680), with runs of unequal length and unequal live sets in one pool."""
    from nestfit_b200.sampler import NestedSamplingBatch, MARG_QUANTILES
    blk, ut, _ = gauss_problem(nb, seed=8, n_pix=5)
    ns = NestedSamplingBatch(blk, ut, 1, pix_ids=[0, 1, 2, 3, 4], nlive=[40, 64, 100, 150, 64], tol=0.5, n_prop=16, seed=2)
    res = ns.run()
    assert not res["truncated"].any()
    prod = ns.products_all()
    off = prod["row_offsets"]
    assert off[0] == 0 and prod["posteriors"].shape == (off[-1], 5) and prod["posteriors"].dtype == np.float32
    for r in range(5):
        post = ns.posterior(r)
        rows = prod["posteriors"][off[r]:off[r + 1]]
        assert rows.shape[0] == post.shape[0] == res["n_samples"][r]
        np.testing.assert_array_equal(rows[:, :3], post[:, :3].astype(np.float32))
        np.testing.assert_allclose(rows[:, 3], post[:, 3], rtol=1e-6)
        np.testing.assert_allclose(rows[:, 4], post[:, 4], rtol=2e-5, atol=1e-12)
        assert abs(rows[:, 4].astype(np.float64).sum() - 1.0) < 1e-4          # posterior weights
        np.testing.assert_allclose(prod["marginals"][r], np.quantile(post[:, :3], MARG_QUANTILES, axis=0),
                                   rtol=1e-12, atol=1e-12)
    ns.close()


def test_truncated_runs_are_flagged(nb):
    """A run that fills its share of the posterior pool stops before its evidence converged and is flagged
    (CubeFitter repeats such runs with a larger share)."""
    from nestfit_b200.sampler import NestedSamplingBatch
    blk, ut, _ = gauss_problem(nb, seed=9, n_pix=2)
    ns = NestedSamplingBatch(blk, ut, 1, pix_ids=[0, 1], nlive=[50, 50], tol=0.01, n_prop=16, seed=3, max_samples=6 * 50)
    res = ns.run()
    assert res["truncated"].all()
    ns.close()
    ns = NestedSamplingBatch(blk, ut, 1, pix_ids=[0, 1], nlive=[50, 50], tol=0.01, n_prop=16, seed=3)
    full = ns.run()
    assert not full["truncated"].any() and (full["n_iter"] > res["n_iter"]).all()
    ns.close()


def test_lnz_scatter_and_bias_three_components(nb):
    """Evidence at the dimension that dominates a cube fit: three NH3 components, 18 cube dimensions of which 15 are
    active, the pixel of the sampler study (tools/ns_study/harness.py: make_pixel(3, 103), 2 x 380 channels), fitted
    32 times in one batch with the cube fitter's settings (nlive 300, tol 1, default proposal scheme).  Asserted:
    the seed-to-seed scatter of ln Z against the error the runs report, and the mean against the long-walk value of
    the study (-413.87 +- 0.14, walks of 140 steps on the CPU port, tools/ns_study/results.txt), above which the
    default scheme is known to sit.  Measured on B200: ln Z -413.02 +- 0.30 over the 32 runs,
    reported error 0.375 (ratio 0.81), bias +0.85, 6.8e5 likelihood calls per run."""
    from nestfit_b200.sampler import NestedSamplingBatch
    rng = np.random.default_rng(103)
    ut = nb.get_irdc_priors()
    packed = ut.pack()
    xs = [orc.bench_axis(1, 380, 0.158), orc.bench_axis(2, 380, 0.158)]
    while True:
        T = orc.prior_transform(packed, rng.uniform(0.1, 0.9, size=(1, 18)), 3)
        if np.isfinite(T).all():
            break
    clean = orc.nh3_batch(xs, [1, 2], T, 3, want_pred=True)["pred"][0]
    data = clean + rng.normal(0, 0.1, clean.shape)
    blk = nb.PixelBlock("ammonia", xs, data[None].astype(np.float32), 0.1, trans_ids=[1, 2])
    ns = NestedSamplingBatch(blk, ut, 3, pix_ids=np.zeros(32, dtype=np.int32), nlive=300, tol=1.0, n_prop=32, seed=7)
    res = ns.run()
    ns.close()
    blk.close()
    lnz, err = res["lnZ"], res["lnZ_err"]
    assert np.isfinite(lnz).all() and not res["truncated"].any()
    sd, ratio, bias = lnz.std(ddof=1), lnz.std(ddof=1) / err.mean(), lnz.mean() + 413.87
    print(f"3 components, 32 runs: ln Z {lnz.mean():.3f} +- {sd:.3f} (reported error {err.mean():.3f}, ratio {ratio:.2f}), "
          f"bias against the long-walk value {bias:+.2f}, evals per run {res['n_evals'].mean():.3g}")
    assert ratio < 1.3, (sd, err.mean())
    assert -0.5 < bias < 2.0, (lnz.mean(), bias)
