"""GPU parity tests: the CUDA path (through the C ABI) against the committed
golden fixtures of the compiled reference and against the C oracle on seeded
inputs.  Tolerances (BASELINE.json north_star): spectra within 1e-5 of the
spectrum's peak, lnL within 1e-3 absolute near the bulk and 2e-6 relative
elsewhere (|lnL| reaches 1e6 for poor fits where 1e-3 is below FP64-of-FP32
resolution)."""
import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu

SPEC_TOL = 1e-5
LNL_ABS, LNL_REL = 1e-3, 2e-6


def assert_spectra(got, want, tol=SPEC_TOL):
    peak = np.maximum(np.abs(want).max(axis=-1, keepdims=True), 1e-30)
    err = np.abs(got - want) / peak
    assert err.max() < tol, f"max spectrum error {err.max():.3e} of peak"


def assert_lnl(got, want):
    err = np.abs(got - want)
    lim = LNL_ABS + LNL_REL * np.abs(want)
    assert (err <= lim).all(), f"max lnL error {err.max():.3e} (worst ratio {(err / lim).max():.2f})"


def nh3_block(nb, g, ncomp):
    return nb.PixelBlock("ammonia", [g["x11"], g["x22"]], g[f"data{ncomp}"][None], 0.1, trans_ids=[1, 2])


@pytest.mark.parametrize("ncomp", [1, 2, 3, 4])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_nh3_predict_and_loglike_golden(nb, nh3_golden, ncomp, dtype):
    g = nh3_golden
    blk = nh3_block(nb, g, ncomp)
    P = g[f"params{ncomp}"].astype(dtype)
    want_pred, want_lnl = g[f"pred{ncomp}"], g[f"lnL{ncomp}"]
    if dtype == np.float32:   # identical (float32-rounded) vectors through the oracle
        o = orc.nh3_batch([g["x11"], g["x22"]], [1, 2], P.astype(np.float64), ncomp,
                          data=g[f"data{ncomp}"].astype(np.float32)[None], noise=0.1, want_pred=True)
        want_pred, want_lnl = o["pred"], o["lnL"]
    assert_spectra(blk.predict(P, ncomp), want_pred)
    assert_lnl(blk.loglike(P, ncomp, vecs_per_pix=P.shape[0]), want_lnl)
    np.testing.assert_allclose(blk.null_lnZ()[0], g[f"null_lnZ{ncomp}"], rtol=1e-6)


@pytest.mark.parametrize("name", ["hand1", "hand2", "hand3", "edge1", "narrow1", "ortho1"])
def test_nh3_hand_vectors_and_flags(nb, nh3_golden, name):
    g = nh3_golden
    p = g[f"{name}_p"]
    ncomp = p.size // 6
    blk = nb.PixelBlock("ammonia", [g["x11"], g["x22"]], np.zeros((1, 2, 1000)), 0.1, trans_ids=[1, 2])
    for cold, lte in ((0, 0), (1, 0), (0, 1)):
        got = blk.predict(p[None], ncomp, cold=bool(cold), lte=bool(lte))[0]
        assert_spectra(got, g[f"{name}_c{cold}l{lte}"])


def test_nh3_ortho_transition(nb, nh3_golden):
    g = nh3_golden
    blk = nb.PixelBlock("ammonia", [g["x33"]], np.zeros((1, 1, 1000)), 0.1, trans_ids=[3])
    assert_spectra(blk.predict(g["ortho1_p"][None], 1)[0, 0], g["ortho1_33"])


@pytest.mark.parametrize("ncomp", [1, 2, 3, 4])
def test_runner_loglikelihood_golden(nb, nh3_golden, ncomp):
    """Runner.loglikelihood: unit cube in, mutated to physical, lnL out (core.pyx:558-561)."""
    g = nh3_golden
    data = g[f"data{ncomp}"]
    spec_data = [[g["x11"], data[0], 0.1, 1], [g["x22"], data[1], 0.1, 2]]
    runner = nb.AmmoniaRunner.from_data(spec_data, nb.get_irdc_priors(), ncomp=ncomp)
    assert runner.ndim == 6 * ncomp and runner.n_chan_tot == 2000 and runner.n_spec == 2
    np.testing.assert_allclose(runner.null_lnZ, g[f"null_lnZ{ncomp}"], rtol=1e-12)
    U = g[f"run_u{ncomp}"].copy()
    got = np.array([runner.loglikelihood(row) for row in U])
    ok = np.isfinite(g[f"run_lnL{ncomp}"])
    assert_lnl(got[ok], g[f"run_lnL{ncomp}"][ok])
    np.testing.assert_allclose(U[ok], g[f"run_p{ncomp}"][ok], atol=1e-9, rtol=0)   # mutated in place
    with pytest.raises(ValueError, match="Invalid shape"):
        runner.predict(np.zeros(5))


def test_amm_predict_mutates_spectrum(nb, nh3_golden):
    g = nh3_golden
    s = nb.AmmoniaSpectrum(g["x11"], np.zeros(1000), 0.1, trans_id=1)
    nb.amm_predict(s, g["hand2_p"])
    assert_spectra(s.get_spec(), g["hand2_c0l0"][0])
    assert s.sum_spec == pytest.approx(3.687872850e+02, rel=1e-5)


def test_gauss_golden(nb, gauss_golden):
    g = gauss_golden
    blk = nb.PixelBlock("gaussian", [g["x"]], g["data"][None, None], 0.1, rest_freq=float(g["rest_freq"]))
    assert_spectra(blk.predict(g["params"], 8)[:, 0], g["pred"])
    assert_lnl(blk.loglike(g["params"], 8, vecs_per_pix=g["params"].shape[0]), g["lnL"])
    np.testing.assert_allclose(blk.null_lnZ()[0], g["null_lnZ"], rtol=1e-6)


@pytest.mark.parametrize("name,ncomps", [("irdc", (1, 2, 3, 4)), ("synth", (1, 2))])
def test_prior_transform_golden(nb, prior_golden, name, ncomps):
    ut = nb.get_irdc_priors() if name == "irdc" else nb.get_synth_priors()
    for ncomp in ncomps:
        U, want = prior_golden[f"{name}_u{ncomp}"].copy(), prior_golden[f"{name}_p{ncomp}"]
        got = ut.transform_batch(U, ncomp)
        ok = np.isfinite(want)
        assert (np.isfinite(got) == ok).all()
        np.testing.assert_allclose(got[ok], want[ok], rtol=0, atol=1e-9)


@pytest.mark.parametrize("name", ["ordered", "spaced", "censep"])
def test_prior_kinds_golden(nb, name):
    """The device prior transform for OrderedPrior, SpacedPrior and CenSepPrior (core.pyx:241-318) against
    fixtures generated from the compiled reference, ncomp 1..4 (CenSepPrior's unparametrised ncomp > 2 included)."""
    import sys
    from pathlib import Path
    import nestfit_b200.core as nbcore
    gdir = Path(__file__).resolve().parent / "golden"
    sys.path.insert(0, str(gdir))
    from prior_kind_sets import kind_prior_sets
    g = np.load(gdir / "prior_kinds_golden.npz")
    ut = kind_prior_sets(nbcore)[name]
    for ncomp in (1, 2, 3, 4):
        U, want = g[f"{name}_u{ncomp}"].copy(), g[f"{name}_p{ncomp}"]
        got = ut.transform_batch(U, ncomp)
        ok = np.isfinite(want)
        assert (np.isfinite(got) == ok).all()
        np.testing.assert_allclose(got[ok], want[ok], rtol=0, atol=1e-9)


def test_prior_transform_random_vs_oracle(nb):
    ut = nb.get_irdc_priors()
    rng = np.random.default_rng(11)
    for ncomp in (1, 2, 3, 4):
        U = rng.uniform(size=(4096, 6 * ncomp))
        want = orc.prior_transform(ut.pack(), U, ncomp)
        got = ut.transform_batch(U.copy(), ncomp)
        ok = np.isfinite(want)
        assert (np.isfinite(got) == ok).all()
        np.testing.assert_allclose(got[ok], want[ok], rtol=0, atol=1e-9)


def _random_problem(nb, rng, ncomp, n_pix, n_chan, dv):
    ut = nb.get_irdc_priors()
    xs = [orc.bench_axis(1, n_chan, dv), orc.bench_axis(2, n_chan, dv)]
    truth = orc.prior_transform(ut.pack(), rng.uniform(size=(3 * n_pix, 6 * ncomp)), ncomp)
    truth = truth[np.isfinite(truth).all(axis=1)][:n_pix]
    clean = orc.nh3_batch(xs, [1, 2], truth, ncomp, want_pred=True)["pred"]
    noise = rng.uniform(0.05, 0.3, size=(n_pix, 2))
    data = (clean + rng.normal(size=clean.shape) * noise[:, :, None]).astype(np.float32)
    return ut, xs, data, noise


@pytest.mark.parametrize("ncomp,n_chan,dv", [(1, 1000, 0.07), (2, 379, 0.158), (3, 1000, 0.07), (4, 1024, 0.05),
                                             (3, 33, 1.0), (2, 100, 0.5), (2, 65, 0.9), (3, 4500, 0.02)])
def test_nh3_random_batch_vs_oracle(nb, ncomp, n_chan, dv):
    """Ragged, unsorted pixel assignment; tiles straddling pixels; odd channel counts."""
    rng = np.random.default_rng(100 + ncomp + n_chan)
    n_pix = 7
    ut, xs, data, noise = _random_problem(nb, rng, ncomp, n_pix, n_chan, dv)
    B = 1500
    P = orc.prior_transform(ut.pack(), rng.uniform(size=(B, 6 * ncomp)), ncomp)
    P[~np.isfinite(P).all(axis=1)] = P[0]
    P[:n_pix] = P[:n_pix] * 0 + orc.prior_transform(ut.pack(), np.full((n_pix, 6 * ncomp), 0.37), ncomp)
    pix = np.sort(rng.integers(0, n_pix, size=B)).astype(np.int32)
    pix[::97] = rng.integers(0, n_pix, size=pix[::97].shape)     # break the sortedness
    blk = nb.PixelBlock("ammonia", xs, data, noise, trans_ids=[1, 2])
    want = orc.nh3_batch(xs, [1, 2], P, ncomp, data=data.astype(np.float64), noise=noise, pix_of_vec=pix,
                         want_pred=True)
    assert_lnl(blk.loglike(P, ncomp, pix_of_vec=pix), want["lnL"])
    assert_spectra(blk.predict(P[:200], ncomp), want["pred"][:200])
    # implicit contiguous layout
    vpp = 200
    want2 = orc.nh3_batch(xs, [1, 2], P[:vpp * n_pix], ncomp, data=data.astype(np.float64), noise=noise,
                          pix_of_vec=(np.arange(vpp * n_pix) // vpp).astype(np.int32))
    assert_lnl(blk.loglike(P[:vpp * n_pix], ncomp, vecs_per_pix=vpp), want2["lnL"])


def test_nh3_wide_band_ortho_mix_and_flags(nb):
    """More than 1024 channels (several dispatch super-blocks), a para + ortho transition pair with
    a free ortho fraction, one spectrum only, and the cold / lte switches through the likelihood."""
    rng = np.random.default_rng(77)
    ut = nb.get_irdc_priors()
    for trans, n_chan, dv in (((1, 3), 2500, 0.05), ((2,), 300, 0.2), ((1, 2, 3), 700, 0.1)):
        xs = [orc.bench_axis(t, n_chan, dv) for t in trans]
        for ncomp in (1, 3):
            P = orc.prior_transform(ut.pack(), rng.uniform(size=(400, 6 * ncomp)), ncomp)
            P = P[np.isfinite(P).all(axis=1)][:256]
            P[:, 5 * ncomp:] = rng.uniform(0.05, 0.9, size=(P.shape[0], ncomp))       # ortho fraction
            clean = orc.nh3_batch(xs, list(trans), P[:2], ncomp, want_pred=True)["pred"]
            data = (clean + rng.normal(0, 0.1, clean.shape)).astype(np.float32)
            pix = (np.arange(P.shape[0]) % 2).astype(np.int32)
            blk = nb.PixelBlock("ammonia", xs, data, 0.1, trans_ids=list(trans))
            for cold, lte in ((False, False), (True, False), (False, True), (True, True)):
                want = orc.nh3_batch(xs, list(trans), P, ncomp, data=data.astype(np.float64),
                                     noise=np.full((2, len(trans)), 0.1), pix_of_vec=pix, cold=cold, lte=lte,
                                     want_pred=True)
                assert_lnl(blk.loglike(P, ncomp, pix_of_vec=pix, cold=cold, lte=lte), want["lnL"])
                assert_spectra(blk.predict(P[:64], ncomp, cold=cold, lte=lte), want["pred"][:64])
            blk.close()


def test_edge_cases(nb):
    rng = np.random.default_rng(5)
    ut, xs, data, noise = _random_problem(nb, rng, 2, 3, 1000, 0.07)
    blk = nb.PixelBlock("ammonia", xs, data, noise, trans_ids=[1, 2])
    assert blk.loglike(np.zeros((0, 12)), 2, vecs_per_pix=1).shape == (0,)
    null = blk.null_lnZ()
    # lines entirely outside the band -> null-model likelihood (hyperfine.pyx:88-89)
    far = np.array([[500.0, 510.0, 10, 12, 4, 5, 14, 14, 0.3, 0.4, 0, 0]])
    np.testing.assert_allclose(blk.loglike(far, 2, pix_of_vec=np.array([1], dtype=np.int32))[0], null[1], rtol=1e-6)
    # NaN parameters do not crash; NaN voff/sigma give the null model like the reference's
    # skipped windows, NaN temperatures give NaN
    bad = far.copy(); bad[0, 0] = np.nan; bad[0, 1] = np.nan
    np.testing.assert_allclose(blk.loglike(bad, 2, vecs_per_pix=1)[0], null[0], rtol=1e-6)
    bad = np.array([[0.0, 1.0, np.nan, 12, 4, 5, 14, 14, 0.3, 0.4, 0, 0]])
    assert np.isnan(blk.loglike(bad, 2, vecs_per_pix=1)[0])
    with pytest.raises(ValueError):
        blk.loglike(np.zeros((4, 11)), 2)
    with pytest.raises(ValueError):
        blk.loglike(np.zeros((4, 12)), 2, pix_of_vec=np.array([0, 1, 2, 3], dtype=np.int32))
    from nestfit_b200 import _lib
    with pytest.raises(_lib.NfError):
        blk.loglike(np.zeros((4, 30)), 5, vecs_per_pix=4)     # ncomp > NF_MAX_NCOMP_NH3


def test_gauss_random_vs_oracle(nb):
    rng = np.random.default_rng(9)
    n_chan, ncomp, n_pix, B = 4096, 8, 3, 600
    v = (np.arange(n_chan) - 0.5 * (n_chan - 1)) * 0.05
    x = np.sort(orc.NU[0] * (1 - v / orc.CKMS))
    P = np.concatenate([np.sort(rng.uniform(-90, 90, (B, ncomp)), axis=1), rng.uniform(0.2, 3, (B, ncomp)),
                        rng.uniform(0.1, 5, (B, ncomp))], axis=1)
    clean = orc.gauss_batch(x, orc.NU[0], P[:n_pix], ncomp, want_pred=True)["pred"]
    data = (clean + rng.normal(0, 0.1, clean.shape)).astype(np.float32)
    pix = np.sort(rng.integers(0, n_pix, B)).astype(np.int32)
    blk = nb.PixelBlock("gaussian", [x], data[:, None, :], 0.1, rest_freq=orc.NU[0])
    want = orc.gauss_batch(x, orc.NU[0], P, ncomp, data=data.astype(np.float64), noise=np.full(n_pix, 0.1),
                           pix_of_vec=pix, want_pred=True)
    assert_lnl(blk.loglike(P, ncomp, pix_of_vec=pix), want["lnL"])
    assert_spectra(blk.predict(P[:100], ncomp)[:, 0], want["pred"][:100])


def test_full_size_properties(nb):
    """BASELINE config 2 size (2^20 vectors, 1024 pixels x 1024, 3 components, 2 x 1000
    channels): size-independent properties + a sampled comparison with the oracle."""
    rng = np.random.default_rng(1234)
    ncomp, n_pix, vpp = 3, 1024, 1024
    ut, xs, data, noise = _random_problem(nb, rng, ncomp, n_pix, 1000, 0.07)
    B = n_pix * vpp
    U = np.random.default_rng(4321).uniform(size=(B, 18))
    P = ut.transform_batch(U, ncomp)
    bad = ~np.isfinite(P).all(axis=1)
    P[bad] = P[np.flatnonzero(~bad)[0]]
    P32 = P.astype(np.float32)
    blk = nb.PixelBlock("ammonia", xs, data, noise, trans_ids=[1, 2])
    lnl = blk.loglike(P32, ncomp, vecs_per_pix=vpp)
    assert np.isfinite(lnl).all() and (lnl <= 0).all()
    # (a) determinism / idempotence
    assert np.array_equal(lnl, blk.loglike(P32, ncomp, vecs_per_pix=vpp))
    # (b) explicit pixel map == implicit layout; permuting vectors permutes results
    sub = rng.choice(B, size=50000, replace=False)
    pix = (sub // vpp).astype(np.int32)
    np.testing.assert_allclose(blk.loglike(P32[sub], ncomp, pix_of_vec=pix), lnl[sub], rtol=1e-12, atol=1e-9)
    # (c) sampled oracle comparison
    pick = rng.choice(B, size=1500, replace=False)
    want = orc.nh3_batch(xs, [1, 2], P32[pick].astype(np.float64), ncomp, data=data.astype(np.float64),
                         noise=noise, pix_of_vec=(pick // vpp).astype(np.int32))
    assert_lnl(lnl[pick], want["lnL"])
    # (d) never better than a perfect fit bound and never above 0; best vector per pixel beats null
    null = blk.null_lnZ()
    assert (lnl.reshape(n_pix, vpp).max(axis=1) > null - 1e-6).mean() > 0.5


def test_predict_loglike_consistency_properties(nb):
    """Properties tying the two kernels together, independent of the oracle: (a) a pixel made of the predict
    kernel's own output is fitted exactly (lnL = 0); (b) lnL equals the chi-square of data - predict summed on the
    host; (c) a component entirely outside the band changes nothing (ncomp = 3 vs ncomp = 2); (d) exchanging two
    components leaves lnL unchanged to FP32 round-off."""
    rng = np.random.default_rng(77)
    ncomp, n = 3, 2048
    ut, xs, data, noise = _random_problem(nb, rng, ncomp, 4, 1000, 0.07)
    P = ut.transform_batch(rng.uniform(size=(n, 18)), ncomp)
    P = P[np.isfinite(P).all(axis=1)]
    n = P.shape[0]
    scratch = nb.PixelBlock("ammonia", xs, data, noise, trans_ids=[1, 2])
    pred = scratch.predict(P, ncomp)                                   # [n, 2, 1000] float32
    # (b) against the first four pixels
    pix = (np.arange(n) % 4).astype(np.int32)
    lnl = scratch.loglike(P, ncomp, pix_of_vec=pix)
    d = data[pix].astype(np.float64) - pred.astype(np.float64)
    want = -(d * d / (2.0 * np.asarray(noise, dtype=np.float64)[pix][:, :, None] ** 2)).sum(axis=(1, 2))
    np.testing.assert_allclose(lnl, want, rtol=3e-6, atol=1e-3)
    # (c) third component moved out of the band == two-component model of the first two
    P3 = P.copy()
    P3[:, 2] = 900.0                                                   # voff of component 3 [km/s]
    P2 = np.ascontiguousarray(P.reshape(n, 6, 3)[:, :, :2].reshape(n, 12))
    np.testing.assert_allclose(scratch.loglike(P3, 3, pix_of_vec=pix), scratch.loglike(P2, 2, pix_of_vec=pix),
                               rtol=3e-6, atol=1e-3)
    # (d) components 1 and 2 exchanged
    Px = P.reshape(n, 6, 3)[:, :, [1, 0, 2]].reshape(n, 18).copy()
    np.testing.assert_allclose(scratch.loglike(Px, 3, pix_of_vec=pix), lnl, rtol=3e-6, atol=1e-3)
    scratch.close()
    # (a) pixels = predicted spectra, one vector per pixel
    own = nb.PixelBlock("ammonia", xs, pred, 0.1, trans_ids=[1, 2])
    z = own.loglike(P, ncomp, vecs_per_pix=1)
    assert np.all(np.abs(z) <= 1e-9), np.abs(z).max()
    own.close()


def _north_star_vectors(nb, rng, n_pix=16, per_pix=384):
    """configs[1]-shaped pixels (3 components, 2 x 1000 channels, sigma = 0.1 K) with, per pixel, vectors from
    the posterior bulk outwards: the truth, the truth perturbed on scales 1e-4 ... 0.3 of the prior width,
    and plain prior draws."""
    ncomp = 3
    ut = nb.get_irdc_priors()
    xs = [orc.bench_axis(1, 1000, 0.07), orc.bench_axis(2, 1000, 0.07)]
    U0 = rng.uniform(0.15, 0.85, size=(4 * n_pix, 6 * ncomp))
    T = orc.prior_transform(ut.pack(), U0, ncomp)
    keep = np.isfinite(T).all(axis=1)
    U0, T = U0[keep][:n_pix], T[keep][:n_pix]
    clean = orc.nh3_batch(xs, [1, 2], T, ncomp, want_pred=True)["pred"]
    data = (clean + rng.normal(0.0, 0.1, clean.shape)).astype(np.float32)
    scales = 10.0 ** rng.uniform(-4.0, -0.5, size=(n_pix, per_pix, 1))
    U = np.clip(U0[:, None, :] + scales * rng.normal(size=(n_pix, per_pix, 6 * ncomp)), 1e-6, 1 - 1e-6)
    U[:, 0] = U0
    U[:, per_pix // 2:] = rng.uniform(size=(n_pix, per_pix - per_pix // 2, 6 * ncomp))
    P = orc.prior_transform(ut.pack(), U.reshape(-1, 6 * ncomp), ncomp)
    bad = ~np.isfinite(P).all(axis=1)
    P[bad] = np.repeat(T, per_pix, axis=0)[bad]
    return xs, data, P, per_pix


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_lnl_north_star_bound(nb, dtype):
    """The literal bound of BASELINE.json's north_star on configs[1]-shaped data: |lnL_gpu - lnL_ref| <= 1e-3 for
    every vector within 1e3 of its pixel's best log-likelihood (the posterior bulk and a thousand e-folds beyond:
    everything nested sampling weights), and 1e-6 |lnL| for the poor fits outside, where 1e-3
    absolute is below the resolution the reference's own float-rounded FastExp arguments leave."""
    rng = np.random.default_rng(20261018)
    xs, data, P, per_pix = _north_star_vectors(nb, rng)
    P = P.astype(dtype)
    n_pix = data.shape[0]
    blk = nb.PixelBlock("ammonia", xs, data, 0.1, trans_ids=[1, 2])
    got = blk.loglike(P, 3, vecs_per_pix=per_pix)
    want = orc.nh3_batch(xs, [1, 2], P.astype(np.float64), 3, data=data.astype(np.float64),
                         noise=np.full((n_pix, 2), 0.1),
                         pix_of_vec=(np.arange(P.shape[0]) // per_pix).astype(np.int32))["lnL"]
    blk.close()
    err = np.abs(got - want)
    best = want.reshape(n_pix, per_pix).max(axis=1)
    near = (np.repeat(best, per_pix) - want) <= 1e3
    assert near.sum() >= 0.25 * near.size and (~near).sum() >= 0.25 * near.size      # both regimes are populated
    i_near = int(np.argmax(np.where(near, err, -1.0)))
    i_far = int(np.argmax(np.where(~near, err / np.abs(want), -1.0)))
    print(f"\nnorth-star lnL bound ({np.dtype(dtype).name}): bulk max |dlnL| {err[i_near]:.2e} at lnL {want[i_near]:.1f} "
          f"({near.sum()} vectors); outside max |dlnL|/|lnL| {err[i_far] / abs(want[i_far]):.2e} at lnL {want[i_far]:.4g}")
    assert err[near].max() <= 1e-3
    assert (err[~near] <= 1e-6 * np.abs(want[~near])).all()
