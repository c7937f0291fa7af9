"""GPU parity of the N2H+ (diazenylium) model -- the same fused hyperfine kernel with the
N2H+ line tables and four-parameter front end -- against the compiled reference's fixtures
(tests/golden/n2hp_golden.npz) and the C oracle; runner mirror of diazenylium.pyx:161-232."""
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


@pytest.mark.parametrize("trans", [1, 2, 3])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_n2hp_golden(nb, n2hp_golden, trans, dtype):
    g = n2hp_golden
    x = g[f"x{trans}"]
    for ncomp in (1, 2, 3):
        P = g[f"params{trans}_{ncomp}"].astype(dtype)
        data = g[f"data{trans}_{ncomp}"]
        blk = nb.PixelBlock("diazenylium", [x], data[None, None], 0.1, trans_ids=[trans])
        want_pred, want_lnl = g[f"pred{trans}_{ncomp}"], g[f"lnL{trans}_{ncomp}"]
        if dtype == np.float32:   # identical (float32-rounded) inputs through the oracle
            o = orc.n2hp_batch([x], [trans], P.astype(np.float64), ncomp, data=data.astype(np.float32)[None, None],
                               noise=0.1, want_pred=True)
            want_pred, want_lnl = o["pred"][:, 0], o["lnL"]
        assert_spectra(blk.predict(P, ncomp)[:, 0], want_pred)
        assert_lnl(blk.loglike(P, ncomp, vecs_per_pix=P.shape[0]), want_lnl)


def test_n2hp_two_transitions_random_vs_oracle(nb):
    """(1-0) and (3-2) scored together (45 lines: two passes of the flattened line set-up),
    four components, ragged pixel assignment."""
    rng = np.random.default_rng(77)
    n_pix, n_chan, ncomp, B = 5, 700, 4, 333
    xs = [orc.bench_axis(1, n_chan, 0.08), orc.bench_axis(1, n_chan, 0.08)]
    xs = [np.sort(orc.N2HP_NU[t - 1] * (1 - (np.arange(n_chan) - 0.5 * (n_chan - 1)) * 0.08 / orc.CKMS)) for t in (1, 3)]
    def draw(n):
        return np.concatenate([np.sort(rng.uniform(-8, 8, (n, ncomp)), axis=1), rng.uniform(3, 30, (n, ncomp)),
                               rng.uniform(-2, 1.5, (n, ncomp)), rng.uniform(0.06, 2.0, (n, ncomp))], axis=1)
    truth = draw(n_pix)
    clean = orc.n2hp_batch(xs, [1, 3], truth, ncomp, want_pred=True)["pred"]
    noise = rng.uniform(0.05, 0.3, (n_pix, 2))
    data = (clean + rng.normal(size=clean.shape) * noise[:, :, None]).astype(np.float32)
    blk = nb.PixelBlock("diazenylium", xs, data, noise, trans_ids=[1, 3])
    P = draw(B)
    pv = rng.integers(0, n_pix, B).astype(np.int32)
    want = orc.n2hp_batch(xs, [1, 3], P, ncomp, data=data.astype(np.float64), noise=noise, pix_of_vec=pv,
                          want_pred=True)
    assert_lnl(blk.loglike(P, ncomp, pix_of_vec=pv), want["lnL"])
    assert_spectra(blk.predict(P, ncomp), want["pred"])
    np.testing.assert_allclose(blk.null_lnZ(), -(data.astype(np.float64)**2 / (2 * noise[:, :, None]**2)).sum((1, 2)),
                               rtol=1e-6)


def test_diazenylium_runner_contract(nb, n2hp_golden):
    from nestfit_b200.models import diazenylium as dz
    g = n2hp_golden
    x, data = g["x1"], g["data1_2"]
    size = 200
    u = np.linspace(0, 1, size)
    flat = np.ones(size) / size
    ut = nb.PriorTransformer(np.array([
        nb.OrderedPrior(nb.Distribution(12 * u - 6, flat), 0), nb.Prior(nb.Distribution(22 * u + 3, flat), 1),
        nb.Prior(nb.Distribution(3 * u - 1.5, flat), 2), nb.Prior(nb.Distribution(1.5 * u + 0.07, flat), 3)],
        dtype=object))
    runner = dz.DiazenyliumRunner.from_data([[x, data, 0.1, 1]], ut, ncomp=2)
    assert runner.n_model == 4 and runner.ndim == 8 and runner.n_chan_tot == 800 and runner.n_spec == 1
    assert dz.NAME == 'diazenylium' and dz.IX_SIGM == 3 and dz.get_par_names(2)[:2] == ['v1', 'v2']
    rng = np.random.default_rng(2)
    U = rng.uniform(size=(6, 8))
    Uc = U.copy()
    got = np.array([runner.loglikelihood(row) for row in Uc])
    P = orc.prior_transform(ut.pack(), U, 2)
    np.testing.assert_allclose(Uc, P, atol=1e-9, rtol=0)                      # mutated in place to physical
    want = orc.n2hp_batch([x], [1], P, 2, data=data[None, None], noise=0.1)["lnL"]
    assert_lnl(got, want)
    with pytest.raises(ValueError, match="Invalid shape"):
        runner.predict(np.zeros(5))
    s = dz.DiazenyliumSpectrum(x, np.zeros(800), 0.1, trans_id=1)
    dz.nnhp_predict(s, g["params1_2"][0])
    assert_spectra(s.get_spec(), g["pred1_2"][0])


def test_n2hp_cube_fit_through_cubefitter(nb, tmp_path):
    """The N2H+ model through the whole drop-in path: CubeFitter -> batched sampler (n_model = 4) ->
    store, with the reference's escalation rule (main.py:450-469)."""
    from nestfit_b200.models import diazenylium as dz
    from nestfit_b200.main import DataCube, CubeStack
    from nestfit_b200.store import HdfStore
    rng = np.random.default_rng(9)
    n_chan = 300
    x = np.sort(orc.N2HP_NU[0] * (1 - (np.arange(n_chan) - 0.5 * (n_chan - 1)) * 0.12 / orc.CKMS))
    size = 200
    u = np.linspace(0, 1, size)
    flat = np.ones(size) / size
    ut = nb.PriorTransformer(np.array([
        nb.OrderedPrior(nb.Distribution(8 * u - 4, flat), 0), nb.Prior(nb.Distribution(17 * u + 3, flat), 1),
        nb.Prior(nb.Distribution(2.5 * u - 1.5, flat), 2), nb.Prior(nb.Distribution(1.0 * u + 0.1, flat), 3)],
        dtype=object))
    truth = np.array([0.7, 8.0, 0.3, 0.35])
    clean = orc.n2hp_batch([x], [1], truth[None], 1, want_pred=True)["pred"][0, 0]
    data = np.zeros((2, 2, n_chan))
    data[0] = clean + rng.normal(0, 0.1, (2, n_chan))        # signal in the first row, noise in the second
    data[1] = rng.normal(0, 0.1, (2, n_chan))
    stack = CubeStack([DataCube.from_arrays(data, x, 0.1, trans_id=1)])
    fitter = nb.CubeFitter(stack, ut, dz.DiazenyliumRunner, ncomp_max=2, mn_kwargs={'nlive': 100}, lnZ_thresh=11, seed=3)
    res = fitter.fit_cube(str(tmp_path / 'n2hp'), nproc=1)[0]
    nbest = res['nbest'].reshape(2, 2)
    assert np.all(nbest[0] == 1) and np.all(nbest[1] == 0)
    store = HdfStore(str(tmp_path / 'n2hp'))
    assert store.hdf.attrs['model_name'] == 'diazenylium' and store.hdf.attrs['n_params'] == 4
    run = store.hdf['/pix/0/1/1']
    assert run.attrs['n_params'] == 4 and run['marginals'].shape == (15, 4)
    bf = run['bestfit_params'][...]
    assert abs(bf[0] - 0.7) < 0.1 and abs(bf[3] - 0.35) < 0.1
    store.close()
