"""Synthetic NH3 cubes built with the product's own predict kernel (the reference
builds them with pyspeckit, nestfit/synth_spectra.py:195-236, which is absent here)."""
import numpy as np

from .main import CubeStack, DataCube, NoiseMap
from .pixels import PixelBlock

CKMS = 299792.458
NU = {1: 23.6944955e9, 2: 23.722633335e9}


def velocity_axis_hz(trans_id, n_chan, dv):
    v = (np.arange(n_chan) - 0.5 * (n_chan - 1)) * dv
    return np.sort(NU[trans_id] * (1.0 - v / CKMS))


def make_synth_stack(shape, utrans, ncomp_map=None, n_chan=1000, dv=0.07, noise=0.1, seed=1234, device=0):
    """CubeStack of (1,1)+(2,2) cubes: `ncomp_map[lon, lat]` true components per pixel
    (0 = noise only), truths drawn through `utrans`; `noise` scalar or (lon, lat) map."""
    rng = np.random.default_rng(seed)
    n_lon, n_lat = shape
    n_pix = n_lon * n_lat
    if ncomp_map is None:
        ncomp_map = (np.indices(shape).sum(axis=0) // max(1, (n_lon + n_lat) // 8)) % 4
    ncomp_map = np.asarray(ncomp_map).reshape(n_pix)
    xs = [velocity_axis_hz(1, n_chan, dv), velocity_axis_hz(2, n_chan, dv)]
    clean = np.zeros((n_pix, 2, n_chan), dtype=np.float32)
    blk = PixelBlock("ammonia", xs, np.zeros((1, 2, n_chan), dtype=np.float32), 1.0, trans_ids=[1, 2], device=device)
    truths = {}
    for nc in np.unique(ncomp_map):
        if nc == 0:
            continue
        idx = np.flatnonzero(ncomp_map == nc)
        U = rng.uniform(size=(idx.size * 2 + 8, 6 * int(nc)))
        T = utrans.transform_batch(U, int(nc), device=device)
        T = T[np.isfinite(T).all(axis=1)][:idx.size]
        clean[idx] = blk.predict(T, int(nc))
        truths[int(nc)] = (idx, T)
    blk.close()
    noise_map = np.broadcast_to(np.asarray(noise, dtype=np.float64), shape)
    data = clean + (rng.normal(size=clean.shape) * noise_map.reshape(n_pix, 1, 1)).astype(np.float32)
    cubes = []
    for t in (0, 1):
        nm = NoiseMap(noise_map.T.copy()) if np.ndim(noise) else float(noise)
        cubes.append(DataCube.from_arrays(data[:, t, :].reshape(n_lon, n_lat, n_chan), xs[t], nm, trans_id=t + 1,
                                          header={'CTYPE1': 'RA---SIN', 'CTYPE2': 'DEC--SIN', 'NAXIS1': n_lon,
                                                  'NAXIS2': n_lat}))
    stack = CubeStack(cubes)
    stack.truths = truths
    stack.ncomp_map = ncomp_map.reshape(shape)
    return stack


# ---- host-side helpers of the reference's synthetic-data module (nestfit/synth_spectra.py) ----------------
# parameter sets of `get_test_spectra(kind)` (synth_spectra.py:250-267): two components, parameter-major
TEST_PARAMS = {
    0: np.array([-1.0, 1.5, 10.0, 15.0, 4.0, 6.0, 14.5, 15.0, 0.3, 0.6, 0.0, 0.0]),
    1: np.array([-1.0, 1.0, 12.0, 12.0, 6.0, 6.0, 14.5, 14.6, 0.3, 0.3, 0.0, 0.0]),
}


def test_axes(vchan=0.158):
    """Frequency axes [Hz, ascending] of the reference's test spectra: v = arange(-30, 30, vchan) km/s in the
    radio convention about the (1,1) and (2,2) rest frequencies (synth_spectra.py:246-249)."""
    v = np.arange(-30, 30, vchan)
    return [np.sort(NU[t] * (1.0 - v / CKMS)) for t in (1, 2)]


test_axes.__test__ = False      # not a pytest test


class ParamSampler:
    """Uniform draws of two-component NH3 parameter vectors: the first component at 0 km/s, the second
    `vsep` away (synth_spectra.py:165-192).  `rng`: numpy Generator (default: a fresh one)."""

    def __init__(self, vsep=(0.16, 3), trot=(3, 30), tex=(2.8, 12), ntot=(13, 16), sigm=(0.15, 2), orth=(0, 0),
                 rng=None):
        self.vsep, self.trot, self.tex, self.ntot, self.sigm, self.orth = vsep, trot, tex, ntot, sigm, orth
        self.rng = np.random.default_rng() if rng is None else rng

    def draw(self):
        voff = np.array([0.0, self.rng.uniform(*self.vsep)])
        rest = [self.rng.uniform(*r, size=2) for r in (self.trot, self.tex, self.ntot, self.sigm, self.orth)]
        return np.concatenate([voff] + rest)


def add_noise_to_cube(data, std, rng=None):
    """`data` plus independent N(0, std^2) noise (synth_spectra.py:160-162)."""
    rng = np.random.default_rng() if rng is None else rng
    return data + rng.normal(scale=std, size=data.shape) if std > 0 else data + 0.0


def make_fake_header(data, xarr, rest_freq):
    """Header cards of a synthetic (lon, lat, chan) cube with frequency axis `xarr` [Hz], as a dict
    (synth_spectra.py:12-37,149-157: the reference pins the reference pixels at the last element)."""
    return {
        'WCSAXES': 3, 'CRPIX1': data.shape[0], 'CRPIX2': data.shape[1], 'CRPIX3': xarr.shape[0],
        'CDELT1': 1e-4, 'CDELT2': 1e-4, 'CDELT3': float(xarr[1] - xarr[0]),
        'CTYPE1': 'RA---CAR', 'CTYPE2': 'DEC--CAR', 'CTYPE3': 'FREQ', 'CRVAL1': 0, 'CRVAL2': 0,
        'CRVAL3': float(rest_freq), 'CUNIT1': 'deg', 'CUNIT2': 'deg', 'CUNIT3': 'Hz', 'RESTFRQ': float(rest_freq),
        'BUNIT': 'K', 'LONPOLE': 0, 'LATPOLE': 180, 'EQUINOX': 2000.0, 'SPECSYS': 'LSRK', 'RADESYS': 'FK5',
        'SSYSOBS': 'TOPOCENT',
    }
