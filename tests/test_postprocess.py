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


def test_take_by_components_and_kernels(nb):
    """Host helpers of the convolution stages (main.py:529-661); the expected numbers were produced by the
    reference's own functions (tests/golden/make_golden.py imports the reference the same way)."""
    from nestfit_b200 import postprocess as pp
    data = np.arange(24, dtype=float).reshape(2, 3, 4)                 # (m, b, l)
    comps = np.array([[-1, 0, 1, 2], [2, 1, 0, -1], [1, 1, 2, 2]])
    out = pp.take_by_components(data, comps)
    assert out.shape == (3, 4) and np.isnan(out[0, 0]) and np.isnan(out[1, 3])
    assert out[0, 1] == data[0, 0, 1] and out[0, 2] == data[0, 0, 2] and out[0, 3] == data[1, 0, 3]
    out = pp.take_by_components(data, comps, incl_zero=False)
    assert np.isnan(out[0, 1]) and np.isnan(out[1, 2]) and out[2, 3] == data[1, 2, 3]
    wide = np.arange(48, dtype=float).reshape(2, 2, 3, 4)              # (r, m, b, l), components on axis 1
    assert pp.take_by_components(wide, comps, axis=1).shape == (2, 3, 4)
    k = pp.get_indep_info_kernel(1.5, nrad=1, sigma_taper=2.0)
    np.testing.assert_allclose(k, [[0.02048616, 0.01385135, 0.02048616], [0.01385135, 1.0, 0.01385135],
                                   [0.02048616, 0.01385135, 0.02048616]], rtol=1e-6)
    assert pp.get_indep_info_kernel(0.3, nrad=0).shape == (1, 1)
    # a beam much smaller than a pixel: the neighbours carry a full unit of independent information
    np.testing.assert_allclose(pp.get_indep_info_kernel(0.05, nrad=1), np.ones((3, 3)), atol=1e-12)
    # circular aperture: exact pixel overlaps add up to the area of the disc
    mask = pp.apply_circular_mask(np.ones((9, 9)), 3.3)
    np.testing.assert_allclose(mask.sum(), np.pi * 3.3**2, rtol=1e-12)
    assert np.allclose(mask, mask.T) and np.allclose(mask, mask[::-1]) and mask[4, 4] == 1.0 and mask[0, 0] == 0.0
    # default radius = half the kernel width; corner / edge pixels against brute-force supersampling
    got = pp.apply_circular_mask(np.full((5, 5), 2.0))[0, :3]
    sub = (np.arange(400) + 0.5) / 400
    for j, g in enumerate(got):
        xx, yy = np.meshgrid(-2.5 + sub, -2.5 + j + sub, indexing='ij')
        assert abs(g - 2.0 * np.mean(xx**2 + yy**2 <= 2.5**2)) < 2e-4
    assert pp.apply_circular_mask(np.ones((3, 3)), 10.0).sum() == 9.0   # aperture larger than the kernel
    with pytest.raises(ValueError):
        pp.apply_circular_mask(np.ones((4, 5)), 1.0)


def test_pdf_convolution_and_quantiles(nb, tmp_path):
    """convolve_post_pdfs / quantize_conv_marginals / extended_masked_evidence (main.py:777-816,956-1061)."""
    from nestfit_b200 import postprocess as pp
    # fill-and-interpolate convolution: a NaN inside a constant map is interpolated, the kernel sum is kept,
    # and the zero fill outside the map shows at the edge
    img = np.full((1, 7, 9), 2.0)
    img[0, 3, 4] = np.nan
    kern = 3.0 * pp.gaussian_kernel2d(0.5)                    # 5 x 5 support
    out = pp.convolve_fill_interp(img, kern)
    np.testing.assert_allclose(out[0, 3, 4], 6.0, rtol=1e-12)
    np.testing.assert_allclose(out[0, 3, 2], 6.0, rtol=1e-12)
    assert out[0, 0, 0] < 6.0
    store, _ = fake_store(tmp_path, n_lon=5, n_lat=4, ncomp_max=2)
    pp.aggregate_run_attributes(store)
    pp.convolve_evidence(store, 0.7)
    pp.aggregate_run_products(store)
    pp.aggregate_run_pdfs(store)
    d = store.hdf[store.dpath]
    pdfs = np.asarray(d['post_pdfs'][...])
    # an identity kernel without evidence weights reproduces the PDFs (zeros floored at 1e-32)
    ident = np.zeros((3, 3)); ident[1, 1] = 1.0
    pp.convolve_post_pdfs(store, ident, evid_weight=False)
    same = np.asarray(d['conv_post_pdfs'][...])
    assert same.shape == pdfs.shape and np.array_equal(np.isnan(same), np.isnan(pdfs))
    np.testing.assert_allclose(np.nan_to_num(same), np.nan_to_num(pdfs), atol=1e-6)
    # a real kernel with evidence weights: still normalised PDFs on the fitted pixels
    del d['conv_post_pdfs']
    pp.convolve_post_pdfs(store, pp.gaussian_kernel2d(0.8) * 4.0, evid_weight=True)
    conv = np.asarray(d['conv_post_pdfs'][...])
    ok = ~np.isnan(conv[0, 0, 2, 0])
    assert ok.sum() >= 10
    np.testing.assert_allclose(np.nansum(conv[0, 0, 2], axis=0)[ok], 1.0, rtol=1e-5)
    pp.quantize_conv_marginals(store)
    margs = np.asarray(d['conv_marginals'][...])
    assert margs.shape == (2, 2, 6, 15, 4, 5)                            # (r, m, p, M, b, l)
    q = margs[0, 0, 2, :, 1, 1]
    order = np.argsort(np.asarray(d['marg_quantiles'][...]))
    assert np.all(np.diff(q[order]) >= 0) and np.isfinite(q).all()
    bins = np.asarray(d['pdf_bins'][...])[2]
    assert bins[0] <= q.min() and q.max() <= bins[-1]
    assert np.isnan(margs[1, 1, :, :, 1, 2]).all()                      # no two-component run at that pixel
    pp.extended_masked_evidence(store, 0.7, conv=True, lnz_thresh=3)
    mext = np.asarray(d['mext_evidence'][...])
    ev = np.asarray(d['conv_evidence'][...])
    assert mext.shape == (4, 5)
    assert np.isnan(mext[(ev[1] - ev[0]) > 3]).all()
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
