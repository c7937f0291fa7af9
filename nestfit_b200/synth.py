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
