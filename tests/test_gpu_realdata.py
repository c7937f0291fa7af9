"""The reference's real-data fixture through the GPU path: the EVLA NH3 (1,1) + (2,2) cut-outs of
nestfit/test/data (20 x 20 pixels x 379 channels; converted once by tests/golden/make_real_cutout.py with the
minimal FITS reader tests/fits_lite.py) are fitted with `CubeFitter.fit_cube` at the reference's test noise
(NH3_RMS_K = 0.35, nestfit/test/__init__.py:12), the DataCube / CubeStack properties asserted by
nestfit/test/test_main.py:38-71 are checked, and a few pixels are compared with the CPU port of the sampler
(oracle/ns_port.py + C oracle likelihood): same number of components, ln Z within the Monte-Carlo error."""
import multiprocessing as mp
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = Path(__file__).resolve().parent / "golden"
NH3_RMS_K = 0.35
VSYS = 63.0          # km/s: the channel of the cut-out's mean-spectrum peak


def _load_stack(nb):
    z = np.load(GOLDEN / "nh3_real_cutout.npz")
    cubes = []
    for t in (1, 2):
        hdr = {k: (float(v) if v.replace('.', '', 1).replace('-', '', 1).replace('e', '', 1).replace('E', '', 1)
                   .replace('+', '', 1).isdigit() else v) for k, v in zip(z[f'hdr{t}_keys'], z[f'hdr{t}_vals'])}
        cubes.append(nb.DataCube.from_arrays(z[f'data{t}'], z[f'xarr{t}'], NH3_RMS_K, trans_id=t, header=hdr))
    return nb.CubeStack(cubes)


_JOBS = {}          # filled before the fork: the packed priors are ctypes arrays (not picklable)


def _cpu_pixel(k):
    from oracle import ns_port
    xs, data, noise, packed, seed = _JOBS[k]
    return ns_port.fit_pixel(xs, [1, 2], data, noise, packed, ncomp_max=2, lnZ_thresh=11, nlive=100, nlive_snr_fact=5,
                             tol=1.0, efr=0.3, n_prop=32, seed=seed)


def test_real_cutout_contract_and_fit(nb, tmp_path):
    from nestfit_b200.models import ammonia
    stack = _load_stack(nb)
    cube = stack.cubes[0]
    # nestfit/test/test_main.py:38-71
    assert cube.trans_id == 1 and cube.dv and cube.shape == (20, 20, 379) and cube.spatial_shape == (20, 20)
    assert cube.nchan == 379 and cube.full_header and cube.simple_header
    xarr, arr, noise, trans_id, has_nans = cube.get_spec_data(1, 1)
    assert not has_nans and xarr[1] > xarr[0] and not np.any(np.isnan(arr)) and not np.isnan(noise)
    assert stack.shape == (20, 20, 379) and stack.spatial_shape == (20, 20) and stack.full_header
    spec_data, any_nans = stack.get_spec_data(1, 1)
    assert spec_data and not any_nans and stack.get_max_snr(1, 1) > 0
    # the fit
    ut = nb.get_irdc_priors(vsys=VSYS)
    fitter = nb.CubeFitter(stack, ut, ammonia.AmmoniaRunner, ncomp_max=2, lnZ_thresh=11,
                           mn_kwargs={'nlive': 100, 'tol': 1.0, 'efr': 0.3}, nlive_snr_fact=5, seed=11)
    res = fitter.fit_cube(str(tmp_path / 'real'), nproc=1)[0]
    nbest = res['nbest'].reshape(20, 20)
    assert (nbest >= 0).all() and np.isfinite(res['lnZ'][:, 1]).all()
    assert (nbest >= 1).mean() > 0.5                    # the cut-out is centred on emission
    store = nb.HdfStore(str(tmp_path / 'real'))
    assert store.hdf.attrs['naxis1'] == 20 and len(list(store.iter_pix_groups())) == 400
    g = store.hdf['/pix/10/10/1']
    assert g['posteriors'].shape == (g.attrs['n_samples'], 8) and g['marginals'].shape == (15, 6)
    v_best = g['bestfit_params'][0]
    assert abs(v_best - VSYS) < 4.0                     # a velocity inside the prior window around the source
    store.close()
    # CPU port on three pixels: brightest, median and faintest by peak SNR
    snr = stack.block_max_snr(res['i_lon'], res['i_lat'])
    pick = [int(np.argmax(snr)), int(np.argsort(snr)[snr.size // 2]), int(np.argmin(snr))]
    data, noise, _ = stack.block_arrays(res['i_lon'][pick], res['i_lat'][pick])
    xs = [np.asarray(c.xarr) for c in stack.cubes]
    packed = ut.pack()
    for k in range(3):
        _JOBS[k] = (xs, data[k].astype(np.float64), noise[k], packed, 100 + k)
    with mp.get_context("fork").Pool(3) as pool:
        cpu = pool.map(_cpu_pixel, range(3))
    for k, ix in enumerate(pick):
        assert cpu[k]['nbest'] == res['nbest'][ix], (k, cpu[k]['lnZ'], res['lnZ'][ix])
        for n in range(1, len(cpu[k]['lnZ'])):
            err = max(3.0 * np.sqrt(2.0) * res['lnZ_err'][ix, n], 1.5)
            assert abs(cpu[k]['lnZ'][n] - res['lnZ'][ix, n]) < err, (k, n, cpu[k]['lnZ'], res['lnZ'][ix])
