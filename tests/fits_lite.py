"""Minimal FITS reader for the tests (primary HDU only): header cards and the data array.  astropy / spectral_cube,
which the reference uses for ingestion (nestfit/main.py:77-172, nestfit/test/__init__.py:11-27), are not in this image."""
import numpy as np

_DTYPES = {8: 'u1', 16: '>i2', 32: '>i4', 64: '>i8', -32: '>f4', -64: '>f8'}


def read_primary(path):
    """(header dict, data array with numpy axis order NAXISn ... NAXIS1) of the primary HDU."""
    raw = open(path, 'rb').read()
    hdr, pos = {}, 0
    while True:
        card = raw[pos:pos + 80].decode('ascii')
        pos += 80
        key = card[:8].strip()
        if key == 'END':
            break
        if card[8:10] != '= ' or key in ('COMMENT', 'HISTORY', ''):
            continue
        val = card[10:].split(' /')[0].strip() if not card[10:].lstrip().startswith("'") else card[10:]
        if val.lstrip().startswith("'"):
            val = val.lstrip()[1:].split("'")[0].rstrip()
        elif val in ('T', 'F'):
            val = val == 'T'
        else:
            val = float(val) if any(c in val for c in '.EeDd') else int(val)
        hdr[key] = val
    pos = -(-pos // 2880) * 2880
    shape = [hdr[f'NAXIS{k}'] for k in range(hdr['NAXIS'], 0, -1)]
    data = np.frombuffer(raw, dtype=_DTYPES[hdr['BITPIX']], count=int(np.prod(shape)), offset=pos).reshape(shape)
    data = data.astype(np.float64) * hdr.get('BSCALE', 1.0) + hdr.get('BZERO', 0.0)
    return hdr, data


def spectral_axis_hz(hdr):
    """Frequency axis [Hz] of a cube whose third axis is radio velocity (CTYPE3 = 'VRAD') or frequency."""
    n = hdr['NAXIS3']
    w = hdr['CRVAL3'] + (np.arange(n) + 1 - hdr['CRPIX3']) * hdr['CDELT3']
    if hdr['CTYPE3'].startswith('FREQ'):
        return w
    assert hdr['CTYPE3'].startswith('VRAD') and hdr.get('CUNIT3', 'm s-1').strip() in ('m s-1', 'm/s')
    return hdr['RESTFRQ'] * (1.0 - w / 299792458.0)
