"""Post-processing (nestfit/main.py:664-1193): dense aggregation of a store on the host and
the two predict loops (deblend_hf_intensity, generate_predicted_profiles) batched on the GPU."""
import numpy as np
import pytest

from oracle import oracle as orc


def fake_store(tmp_path, n_lon=3, n_lat=2, ncomp_max=2, n_params=6, seed=0):
    """A store with the reference's per-run groups (docs/store_spec.rst) filled with known values."""
    from nestfit_b200.store import HdfStore
    from nestfit_b200.sampler import MARG_QUANTILES
    rng = np.random.default_rng(seed)
    store = HdfStore(str(tmp_path / 'fake'))
    hdf = store.hdf
    hdf.attrs['naxis1'], hdf.attrs['naxis2'] = n_lon, n_lat
    hdf.attrs['n_max_components'], hdf.attrs['n_params'] = ncomp_max, n_params
    hdf.attrs['lnZ_threshold'] = 11
    hdf.attrs['model_name'] = 'ammonia'
    truth = {}
    for i_lon in range(n_lon):
        for i_lat in range(n_lat):
            if (i_lon, i_lat) == (0, 0):
                continue                                    # a blanked pixel: no group at all
            nbest = (i_lon + i_lat) % (ncomp_max + 1)
            g = hdf.require_group(f'/pix/{i_lon}/{i_lat}')
            g.attrs['i_lon'], g.attrs['i_lat'], g.attrs['nbest'] = i_lon, i_lat, nbest
            for ncomp in range(1, min(nbest + 1, ncomp_max) + 1):
                sub = g.create_group(f'{ncomp}')
                lnz = -100.0 + 20 * ncomp if ncomp <= nbest else -100.0 + 20 * nbest + 1
                for k, v in dict(ncomp=ncomp, null_lnZ=-100.0, null_BIC=1.0, null_AIC=2.0, null_AICc=3.0,
                                 global_lnZ=lnz, global_lnZ_err=0.1 * ncomp, BIC=10.0 * ncomp, AIC=11.0 * ncomp,
                                 AICc=12.0 * ncomp, marg_quantiles=MARG_QUANTILES).items():
                    sub.attrs[k] = v
                nd = n_params * ncomp
                post = rng.normal(size=(50, nd + 2)).astype(np.float32)
                mp, bf = rng.normal(size=nd), rng.normal(size=nd)
                sub.create_dataset('posteriors', data=post)
                sub.create_dataset('marginals', data=np.quantile(post[:, :nd], MARG_QUANTILES, axis=0))
                sub.create_dataset('map_params', data=mp)
                sub.create_dataset('bestfit_params', data=bf)
                truth[(i_lon, i_lat, ncomp)] = (mp, bf, post)
    return store, truth


def test_aggregation_shapes_and_values(nb, tmp_path):
    from nestfit_b200 import postprocess as pp
    store, truth = fake_store(tmp_path)
    pp.aggregate_run_attributes(store)
    d = store.hdf[store.dpath]
    nbest = np.asarray(d['nbest'][...])
    assert nbest.shape == (2, 3) and nbest[0, 0] == -1 and nbest[1, 2] == 0 and nbest[1, 1] == 2
    ev = np.asarray(d['evidence'][...])
    assert ev.shape == (3, 2, 3) and np.isnan(ev[:, 0, 0]).all()
    assert ev[0, 1, 1] == -100.0 and ev[1, 1, 1] == -80.0 and ev[2, 1, 1] == -60.0
    assert np.asarray(d['evidence_err'][...])[2, 1, 1] == pytest.approx(0.2)
    pp.aggregate_run_products(store)
    mapd = np.asarray(d['nbest_MAP'][...])
    assert mapd.shape == (2, 6, 2, 3)                      # (m, p, b, l)
    mp = truth[(1, 1, 2)][0].reshape(6, 2)
    np.testing.assert_array_equal(mapd[:, :, 1, 1], mp.T)
    assert np.isnan(mapd[1, :, 0, 1]).all() and np.isfinite(mapd[0, :, 0, 1]).all()      # nbest = 1 pixel
    assert np.asarray(d['nbest_marginals'][...]).shape == (2, 6, 15, 2, 3)
    pp.aggregate_run_pdfs(store)
    pdfs = np.asarray(d['post_pdfs'][...])
    assert pdfs.shape == (2, 2, 6, 199, 2, 3)              # (r, m, p, h, b, l)
    np.testing.assert_allclose(np.nansum(pdfs[1, 0, 3, :, 1, 1]), 1.0, rtol=1e-5)
    assert np.asarray(d['pdf_bins'][...]).shape == (6, 199)
    store.close()


def test_convolve_evidence_semantics(nb, tmp_path):
    from nestfit_b200 import postprocess as pp
    # NaN-interpolating, edge-extending convolution: constants are preserved, NaNs are filled
    img = np.full((7, 9), 3.0)
    img[2, 4] = np.nan
    out = pp.convolve_nan_extend(img, pp.gaussian_kernel2d(1.0))
    np.testing.assert_allclose(out, 3.0, rtol=1e-12)
    # against a direct evaluation at one interior pixel
    rng = np.random.default_rng(1)
    img = rng.normal(size=(9, 9))
    k = pp.gaussian_kernel2d(0.7)
    h = k.shape[0] // 2
    pad = np.pad(img, h, mode='edge')
    want = sum(k[h + dy, h + dx] * pad[4 + h - dy, 5 + h - dx] for dy in range(-h, h + 1) for dx in range(-h, h + 1))
    assert pp.convolve_nan_extend(img, k)[4, 5] == pytest.approx(want, rel=1e-12)
    store, _ = fake_store(tmp_path)
    pp.aggregate_run_attributes(store)
    pp.convolve_evidence(store, 0.5)
    d = store.hdf[store.dpath]
    cn, nbest = np.asarray(d['conv_nbest'][...]), np.asarray(d['nbest'][...])
    assert cn.shape == nbest.shape and cn[0, 0] == -1 and (cn - nbest <= 1).all()
    store.close()


@pytest.mark.gpu
def test_predict_loops_match_oracle(nb, tmp_path):
    """peak/integrated intensity, deblended profiles and MAP model cubes against the oracle's
    per-vector predict on the same MAP parameters (the reference's loop, main.py:1106-1113)."""
    from nestfit_b200 import postprocess as pp
    from nestfit_b200.models import ammonia
    from nestfit_b200.main import DataCube, CubeStack
    store, truth = fake_store(tmp_path)
    rng = np.random.default_rng(5)
    ut = nb.get_irdc_priors()
    xs = [orc.bench_axis(1, nchan=400, dv=0.158), orc.bench_axis(2, nchan=400, dv=0.158)]
    # physical MAP vectors drawn from the prior
    for (i_lon, i_lat, ncomp), _ in truth.items():
        P = ut.transform_batch(rng.uniform(size=(1, 6 * ncomp)), ncomp)[0]
        g = store.hdf[f'/pix/{i_lon}/{i_lat}/{ncomp}']
        del g['map_params']
        g.create_dataset('map_params', data=P)
        truth[(i_lon, i_lat, ncomp)] = P
    pp.aggregate_run_attributes(store)
    pp.aggregate_run_products(store)
    pp.aggregate_run_pdfs(store, par_bins=np.array([np.linspace(-5, 5, 41)] * 6))
    stack = CubeStack([DataCube.from_arrays(np.zeros((3, 2, 400)), xs[t], 0.1, trans_id=t + 1) for t in range(2)])
    spec_data = [[xs[0], np.zeros(400), 0.1, 1], [xs[1], np.zeros(400), 0.1, 2]]
    runner = ammonia.AmmoniaRunner.from_data(spec_data, ut, ncomp=1)
    pp.deblend_hf_intensity(store, stack, runner)
    pp.generate_predicted_profiles(store, stack, runner)
    d = store.hdf[store.dpath]
    pk, ii = np.asarray(d['peak_intensity'][...]), np.asarray(d['integrated_intensity'][...])
    assert pk.shape == (2, 2, 2, 3) and ii.shape == pk.shape                       # (t, m, b, l)
    m11, m22 = np.asarray(d['model_spec']['trans1'][...]), np.asarray(d['model_spec']['trans2'][...])
    assert m11.shape == (2, 400, 2, 3) and m22.shape == (2, 400, 2, 3)             # (m, S, b, l)
    hf = np.asarray(d['hf_deblended'][...])
    assert hf.shape == (2, 2, 40, 2, 3)                                            # (t, m, S, b, l)
    checked = 0
    for (i_lon, i_lat, ncomp), P in truth.items():
        nbest = (i_lon + i_lat) % 3
        if ncomp != nbest:
            continue
        for i_m in range(ncomp):
            p1 = P.reshape(6, ncomp)[:, i_m]
            want = orc.nh3_batch(xs, [1, 2], p1[None], 1, want_pred=True)["pred"][0]
            for t, mc in enumerate((m11, m22)):
                peak = np.abs(want[t]).max()
                assert np.abs(mc[i_m, :, i_lat, i_lon] - want[t]).max() <= 1e-5 * max(peak, 1e-30)
                assert pk[t, i_m, i_lat, i_lon] == pytest.approx(want[t].max(), rel=1e-5, abs=1e-7)
                assert ii[t, i_m, i_lat, i_lon] == pytest.approx(want[t].sum() * stack.cubes[t].dv, rel=2e-5, abs=1e-7)
            checked += 1
    assert checked >= 4
    assert np.isnan(pk[:, :, 0, 0]).all() and np.isnan(m11[:, :, 0, 0]).all()      # blanked pixel stays NaN
    # the deblended profile integrates back to the integrated intensity
    tot = np.nansum(hf[0, 0, :, 1, 1])
    assert tot == pytest.approx(ii[0, 0, 1, 1], rel=0.05)
    store.close()
