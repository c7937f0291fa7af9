#!/usr/bin/env python3
"""nh3_real_cutout.npz: the reference's real-data test fixture (nestfit/test/data/ammonia_{11,22}_cutout.fits, EVLA
NH3 (1,1) and (2,2) cut-outs of 20 x 20 pixels x 380 channels) converted with tests/fits_lite.py into the array
contract of DataCube (data[lon, lat, chan] float32, ascending-or-descending Hz axis, header cards), the last --
NaN -- channel dropped like nestfit/test/__init__.py:26.  Run in the authoring container only (the GPU box has no
/root/reference):  python tests/golden/make_real_cutout.py"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
import fits_lite  # noqa: E402

SRC = Path('/root/reference/nestfit/test/data')
out = {}
for t, name in ((1, '11'), (2, '22')):
    hdr, data = fits_lite.read_primary(SRC / f'ammonia_{name}_cutout.fits')
    assert data.shape == (380, 20, 20) and hdr['BUNIT'].strip() == 'K'
    x = fits_lite.spectral_axis_hz(hdr)
    data, x = data[:-1], x[:-1]                       # last channel contains NaNs
    out[f'data{t}'] = np.ascontiguousarray(data.transpose().astype(np.float32))      # (s, b, l) -> (l, b, s)
    out[f'xarr{t}'] = x
    keys = ('NAXIS1', 'NAXIS2', 'CRPIX1', 'CRPIX2', 'CDELT1', 'CDELT2', 'CUNIT1', 'CUNIT2', 'CTYPE1', 'CTYPE2', 'CRVAL1',
            'CRVAL2', 'RADESYS', 'EQUINOX', 'RESTFRQ', 'BUNIT', 'BMAJ', 'BMIN', 'BPA')
    out[f'hdr{t}_keys'] = np.array(keys)
    out[f'hdr{t}_vals'] = np.array([str(hdr[k]) for k in keys])
np.savez_compressed(HERE / 'nh3_real_cutout.npz', **out)
print('nh3_real_cutout.npz', (HERE / 'nh3_real_cutout.npz').stat().st_size, 'bytes',
      {k: v.shape for k, v in out.items() if k.startswith('data')}, 'NaNs', int(np.isnan(out['data1']).sum()))
