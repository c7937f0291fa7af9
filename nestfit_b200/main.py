"""Cube orchestration: host mirror of the reference's ``nestfit.main`` for the
fitting path (NoiseMap, DataCube, CubeStack, CubeFitter, get_multiproc_indices;
reference nestfit/main.py:39-223,380-526,565-571).

The reference walks pixels one at a time per forked process and calls MultiNest
per (pixel, ncomp).  Here a contiguous block of pixels is uploaded to one GPU and
all its pixels advance in lock-step through the batched sampler, one wave per
number of components; the ncomp-escalation rule (main.py:450-469) selects which
pixels enter the next wave.  Multi-GPU = disjoint contiguous pixel blocks, one
process per GPU, one store chunk per process, no collective.
"""
import os
import threading
import time
from collections.abc import Iterable

import numpy as np

from .pixels import PixelBlock
from .sampler import NestedSamplingBatch
from .store import HdfStore


def nans(shape, dtype=None):
    return np.full(shape, np.nan, dtype=dtype)


class NoiseMap:
    """Per-pixel rms map; like the reference (main.py:39-65) the input is in image
    (lat, lon) order and is transposed to the cube's (lon, lat) order."""

    def __init__(self, data):
        self.data = np.asarray(data).transpose()
        self.shape = self.data.shape

    @classmethod
    def from_pbimg(cls, rms, pb_img):
        shape = pb_img.shape
        naxes = len(shape)
        if naxes == 4:
            pb_img = pb_img[0, 0]
        elif naxes == 3:
            pb_img = pb_img[0]
        elif naxes == 2:
            pass
        else:
            raise ValueError(f'Cannot parse shape : {shape}')
        img = rms / pb_img
        img[~np.isfinite(img)] = np.inf
        return cls(img)

    def get_noise(self, i_lon, i_lat):
        return self.data[i_lon, i_lat]


class NoiseMapUniform:
    def __init__(self, rms):
        self.rms = rms
        self.shape = None

    def get_noise(self, i_lon, i_lat):
        return self.rms


class DataCube:
    """Channel-contiguous cube ``data[lon, lat, chan]`` on an ascending Hz axis with
    scalar noise per pixel: the array contract of the reference's DataCube
    (main.py:77-172).  Built from arrays (`from_arrays`); FITS ingestion through
    spectral_cube stays with the reference (those packages are optional)."""

    def __init__(self, cube=None, noise_map=None, trans_id=None, *, data=None, xarr=None, header=None):
        if cube is not None:
            data, xarr, header = self._from_spectral_cube(cube)
        if isinstance(noise_map, (float, int)):
            self.noise_map = NoiseMapUniform(noise_map)
        else:
            self.noise_map = noise_map
        self.trans_id = trans_id
        self._header = dict(header or {})
        xarr = np.ascontiguousarray(xarr, dtype=np.float64)
        data = np.asarray(data)
        if xarr[1] < xarr[0]:                      # ascending frequency axis (main.py:146-149)
            xarr = xarr[::-1].copy()
            data = data[..., ::-1]
        self.data = np.ascontiguousarray(data)
        self.xarr = xarr
        self.shape = self.data.shape
        self.spatial_shape = (self.shape[0], self.shape[1])
        self.nchan = self.shape[2]
        self.dv = abs(xarr[1] - xarr[0]) / xarr.mean() * 299792.458
        if self.noise_map.shape is not None:
            assert self.spatial_shape == self.noise_map.shape

    @classmethod
    def from_arrays(cls, data, xarr, noise_map, trans_id=None, header=None):
        return cls(None, noise_map, trans_id, data=data, xarr=xarr, header=header)

    @staticmethod
    def _from_spectral_cube(cube):  # pragma: no cover - optional dependency
        from astropy import units
        if cube.unit == '':
            cube._unit = units.K
        elif cube.unit != 'K':
            cube = cube.to('K')
        if cube.spectral_axis.unit != 'Hz':
            cube = cube.with_spectral_unit('Hz')
        axis = cube.spectral_axis.value.copy()
        data = cube._data.transpose().copy()       # (s, b, l) -> (l, b, s)
        return data, axis, dict(cube.header)

    @property
    def full_header(self):
        return self._header

    @property
    def simple_header(self):
        keys = ('SIMPLE', 'BITPIX', 'NAXIS', 'NAXIS1', 'NAXIS2', 'WCSAXES', 'CRPIX1', 'CRPIX2', 'CDELT1', 'CDELT2',
                'CUNIT1', 'CUNIT2', 'CTYPE1', 'CTYPE2', 'CRVAL1', 'CRVAL2', 'RADESYS', 'EQUINOX')
        hdict = {k: self._header[k] for k in keys if k in self._header}
        hdict['NAXIS'] = 2
        hdict['WCSAXES'] = 2
        return hdict

    def get_spec_data(self, i_lon, i_lat):
        arr = self.data[i_lon, i_lat, :]
        noise = self.noise_map.get_noise(i_lon, i_lat)
        has_nans = np.isnan(arr).any() or np.isnan(noise)
        return self.xarr, arr, noise, self.trans_id, has_nans


class CubeStack:
    def __init__(self, cubes):
        assert isinstance(cubes, Iterable)
        self.cubes = list(cubes)
        self.n_cubes = len(self.cubes)

    def __iter__(self):
        for cube in self.cubes:
            yield cube

    @property
    def full_header(self):
        return self.cubes[0].full_header

    @property
    def simple_header(self):
        return self.cubes[0].simple_header

    @property
    def shape(self):
        return self.cubes[0].shape

    @property
    def spatial_shape(self):
        return self.cubes[0].spatial_shape

    def get_arrays(self, i_lon, i_lat):
        return [dcube.get_spec_data(i_lon, i_lat)[1] for dcube in self.cubes]

    def get_spec_data(self, i_lon, i_lat):
        all_spec_data = []
        any_nans = False
        for dcube in self.cubes:
            *spec_data, has_nans = dcube.get_spec_data(i_lon, i_lat)
            all_spec_data.append(spec_data)
            any_nans |= bool(has_nans)
        return all_spec_data, any_nans

    def get_max_snr(self, i_lon, i_lat):
        max_snr = 0.0
        for dcube in self.cubes:
            _, arr, noise, _, _ = dcube.get_spec_data(i_lon, i_lat)
            spec_snr = np.max(arr) / noise
            max_snr = spec_snr if spec_snr > max_snr else max_snr
        return max_snr

    # ---- batched views used by the GPU path ----------------------------------
    def block_arrays(self, lon, lat):
        """data [n, n_cubes, n_chan] float32, noise [n, n_cubes], valid [n] for pixel lists."""
        data = np.stack([c.data[lon, lat, :] for c in self.cubes], axis=1)
        noise = np.stack([np.broadcast_to(np.asarray(c.noise_map.get_noise(lon, lat), dtype=np.float64), lon.shape)
                          for c in self.cubes], axis=1)
        valid = ~(np.isnan(data).any(axis=(1, 2)) | np.isnan(noise).any(axis=1))
        return data, noise, valid


def get_multiproc_indices(shape, nproc):
    """The reference's row striping `lon_ix[i::nproc]` (main.py:565-571)."""
    lon_ix, lat_ix = np.indices(shape)
    return [(lon_ix[i::nproc, ...].flatten(), lat_ix[i::nproc, ...].flatten()) for i in range(nproc)]


def get_block_indices(shape, nblocks):
    """Contiguous pixel blocks of the (lon, lat) grid in C order, one per GPU."""
    lon_ix, lat_ix = np.indices(shape)
    lon_ix, lat_ix = lon_ix.ravel(), lat_ix.ravel()
    bounds = np.linspace(0, lon_ix.size, nblocks + 1).astype(int)
    return [(lon_ix[a:b], lat_ix[a:b]) for a, b in zip(bounds[:-1], bounds[1:])]


class CubeFitter:
    mn_default_kwargs = {
        'nlive': 100,
        'tol': 1.0,
        'efr': 0.3,
        'updInt': 2000,
    }

    def __init__(self, stack, utrans, runner_cls, runner_kwargs=None, lnZ_thresh=11, ncomp_max=2, mn_kwargs=None,
                 nlive_snr_fact=5, n_prop=32, max_pixels_per_wave=16384, seed=1234, store_posteriors=True,
                 n_streams=1, pixels_per_stream=1024, retry_margin=None, retry_chi2_sigma=6.0):
        """Same arguments as the reference (main.py:388-421) plus the batching knobs
        `n_prop` (proposals per pixel per lock-step iteration), `max_pixels_per_wave`
        (pixels in flight per device wave), `n_streams` (host threads / CUDA streams that
        escalate sub-blocks of `pixels_per_stream` pixels concurrently), `seed`, and the optional lost-mode
        guard `retry_margin` (default None = off, one run per (pixel, ncomp) like the reference; e.g. 10):
        an N-component run whose maximum likelihood falls more than this below the (N-1)-component run's
        has lost the dominant mode and is repeated once with twice the live points, as is a run (N >= 2)
        that fails the evidence threshold while its best chi-square is more than `retry_chi2_sigma` sigma
        above the channel count.  The repeats are a small extra wave per ncomp (~1-2 s of latency each)."""
        self.stack = stack
        self.utrans = utrans
        self.runner_cls = runner_cls
        self.runner_kwargs = {} if runner_kwargs is None else runner_kwargs
        self.lnZ_thresh = lnZ_thresh
        self.ncomp_max = ncomp_max
        self.mn_kwargs = self.mn_default_kwargs.copy()
        if mn_kwargs is not None:
            self.mn_kwargs.update(mn_kwargs)
        self.nlive_snr_fact = nlive_snr_fact
        self.n_prop = n_prop
        self.max_pixels_per_wave = max_pixels_per_wave
        self.n_streams = n_streams
        self.pixels_per_stream = pixels_per_stream
        self.retry_margin = retry_margin
        self.retry_chi2_sigma = retry_chi2_sigma
        self.seed = seed
        self.store_posteriors = store_posteriors
        self.stats = {}

    def _model_name(self):
        import inspect
        return inspect.getmodule(self.runner_cls).NAME

    def fit_block(self, indices, device=0, group_root=None, verbose=False):
        """Fit the pixels (all_lon, all_lat) on one GPU.  Returns a dict of per-pixel
        arrays (nbest, lnZ[ncomp_max+1], ...) and, if `group_root` (h5py-like) is
        given, writes the reference's /pix/<lon>/<lat>/<ncomp> groups into it."""
        all_lon, all_lat = (np.asarray(a) for a in indices)
        n_tot = all_lon.size
        model = self._model_name()
        n_evals = 0
        out = dict(i_lon=all_lon, i_lat=all_lat, nbest=np.full(n_tot, -1, dtype=np.int32),
                   lnZ=nans((n_tot, self.ncomp_max + 1)), lnZ_err=nans((n_tot, self.ncomp_max + 1)),
                   max_loglike=nans((n_tot, self.ncomp_max + 1)), n_samples=np.zeros((n_tot, self.ncomp_max + 1), int))
        t0 = time.perf_counter()
        for w0 in range(0, n_tot, self.max_pixels_per_wave):
            sl = slice(w0, min(n_tot, w0 + self.max_pixels_per_wave))
            lon, lat = all_lon[sl], all_lat[sl]
            data, noise, valid = self.stack.block_arrays(lon, lat)
            if not valid.any():
                continue
            vidx = np.flatnonzero(valid)
            dv = data[vidx]
            xarrs = [c.xarr for c in self.stack.cubes]
            if model in ('ammonia', 'diazenylium'):
                blk = PixelBlock(model, xarrs, dv, noise[vidx], trans_ids=[c.trans_id for c in self.stack.cubes],
                                 device=device)
            else:
                blk = PixelBlock('gaussian', xarrs, dv, noise[vidx],
                                 rest_freq=self.runner_kwargs.get('rest_freq', xarrs[0].mean()), device=device)
            null = blk.null_lnZ()
            # live points scale with the peak SNR (main.py:445-447)
            max_snr = np.max(dv.max(axis=2) / noise[vidx], axis=1)
            max_snr = np.maximum(max_snr, 0.0)
            nlive = self.mn_kwargs['nlive'] + (self.nlive_snr_fact * max_snr).astype(int)
            n_chan_tot = blk.n_spec * blk.n_chan
            old_lnZ = null.copy()
            nbest = np.zeros(vidx.size, dtype=np.int32)
            out['lnZ'][w0 + vidx, 0] = null
            evals = []
            n_retried, n_rescued = [0], [0]
            lock = threading.Lock()

            def fit_sub(active, tag):
                """ncomp escalation (main.py:450-469) of one sub-block of the wave's pixels."""
                for ncomp in range(1, self.ncomp_max + 1):
                    if active.size == 0:
                        break
                    if verbose:
                        print(f'-- wave {w0}.{tag}: N = {ncomp}: {active.size} pixels')
                    kw = {k: v for k, v in self.runner_kwargs.items() if k in ('cold', 'lte')}
                    if self.mn_kwargs.get('walks'):
                        kw['walks'] = int(self.mn_kwargs['walks'])       # random-walk steps per new point
                    ns = NestedSamplingBatch(blk, self.utrans, ncomp, pix_ids=active, nlive=nlive[active],
                                             tol=self.mn_kwargs['tol'], efr=self.mn_kwargs['efr'], n_prop=self.n_prop,
                                             seed=self.seed + 7919 * ncomp + w0 + 104729 * tag,
                                             max_iter=self.mn_kwargs.get('maxiter', 1_000_000), **kw)
                    res = ns.run()
                    evals.append(int(res['n_evals'].sum()))
                    assert np.isfinite(res['lnZ']).all()           # main.py:463
                    gi = w0 + vidx[active]
                    # Lost-mode guard (a few per cent of 3-component runs at nlive ~ 300 end in a secondary
                    # mode).  Two symptoms: (i) an N-component model can always reproduce the (N-1)-component
                    # fit, so a best likelihood clearly *below* the previous wave's is a lost mode; (ii) the
                    # run does not pass the evidence threshold although its best fit is far from the noise
                    # (chi-square many sigma above the channel count, the noise being known).  Such runs are
                    # repeated once with twice the live points and a new seed; the repeat replaces the run
                    # when its evidence is higher.
                    owner = [(ns, r) for r in range(active.size)]
                    ns_retry = None
                    if ncomp >= 2 and self.retry_margin is not None:
                        suspect = res['max_loglike'] < out['max_loglike'][gi, ncomp - 1] - self.retry_margin
                        if self.retry_chi2_sigma is not None:
                            chi2_max = n_chan_tot + self.retry_chi2_sigma * np.sqrt(2.0 * n_chan_tot)
                            suspect |= ((-2.0 * res['max_loglike'] > chi2_max) &
                                        (res['lnZ'] - old_lnZ[active] < self.lnZ_thresh))
                        lost = np.flatnonzero(suspect)
                        if lost.size:
                            ns_retry = NestedSamplingBatch(blk, self.utrans, ncomp, pix_ids=active[lost],
                                                           nlive=2 * nlive[active[lost]], tol=self.mn_kwargs['tol'],
                                                           efr=self.mn_kwargs['efr'], n_prop=self.n_prop,
                                                           seed=self.seed + 7919 * ncomp + w0 + 104729 * tag + 15485863,
                                                           max_iter=self.mn_kwargs.get('maxiter', 1_000_000), **kw)
                            res2 = ns_retry.run()
                            evals.append(int(res2['n_evals'].sum()))
                            better = np.flatnonzero(res2['lnZ'] > res['lnZ'][lost])
                            for key in ('lnZ', 'lnZ_err', 'max_loglike', 'n_samples'):
                                res[key][lost[better]] = res2[key][better]
                            for j in better:
                                owner[lost[j]] = (ns_retry, int(j))
                            n_retried[0] += int(lost.size)
                            n_rescued[0] += int(better.size)
                    out['lnZ'][gi, ncomp] = res['lnZ']
                    out['lnZ_err'][gi, ncomp] = res['lnZ_err']
                    out['max_loglike'][gi, ncomp] = res['max_loglike']
                    out['n_samples'][gi, ncomp] = res['n_samples']
                    if group_root is not None:
                        with lock:
                            for r, a in enumerate(active):
                                g = group_root.require_group(f'/pix/{lon[vidx[a]]}/{lat[vidx[a]]}')
                                sub = g.create_group(f'{ncomp}')
                                attrs, dsets = owner[r][0].products(owner[r][1], null[a], n_chan_tot)
                                for k, v in attrs.items():
                                    sub.attrs[k] = v
                                for k, v in dsets.items():
                                    if k == 'posteriors' and not self.store_posteriors:
                                        continue
                                    sub.create_dataset(k, data=v)
                    ns.close()
                    if ns_retry is not None:
                        ns_retry.close()
                    improved = res['lnZ'] - old_lnZ[active] >= self.lnZ_thresh     # main.py:464-469
                    old_lnZ[active[improved]] = res['lnZ'][improved]
                    nbest[active[improved]] = ncomp
                    active = active[improved]

            # Sub-blocks of the wave escalate independently, each on its own stream, picked from a
            # queue by `n_streams` host threads (the sampler call releases the GIL).  The workers
            # drift apart, so the thin tail of one sub-block's run overlaps the bulk of another's
            # instead of idling the GPU.
            n_sub = max(1, vidx.size // max(1, self.pixels_per_stream)) if self.n_streams > 1 else 1
            subs = np.array_split(np.arange(vidx.size), n_sub)
            if n_sub == 1:
                fit_sub(subs[0], 0)
            else:
                errors = []
                todo = list(enumerate(subs))[::-1]

                def worker():
                    while True:
                        with lock:
                            if not todo or errors:
                                return
                            tag, sub = todo.pop()
                        try:
                            fit_sub(sub, tag)
                        except BaseException as exc:   # re-raised in the caller's thread
                            errors.append(exc)

                threads = [threading.Thread(target=worker) for _ in range(min(self.n_streams, n_sub))]
                for th in threads:
                    th.start()
                for th in threads:
                    th.join()
                if errors:
                    raise errors[0]
            n_evals += sum(evals)
            out['n_retried'] = out.get('n_retried', 0) + n_retried[0]
            out['n_rescued'] = out.get('n_rescued', 0) + n_rescued[0]
            out['nbest'][w0 + vidx] = nbest
            if group_root is not None:
                for a in range(vidx.size):
                    g = group_root.require_group(f'/pix/{lon[vidx[a]]}/{lat[vidx[a]]}')
                    g.attrs['i_lon'] = int(lon[vidx[a]])
                    g.attrs['i_lat'] = int(lat[vidx[a]])
                    g.attrs['nbest'] = int(nbest[a])
            blk.close()
        out['seconds'] = time.perf_counter() - t0
        out['n_evals'] = n_evals
        return out

    def fit(self, *args):
        """Reference-compatible worker entry (main.py:423-474): ((all_lon, all_lat), chunk_path)."""
        (all_lon, all_lat), chunk_path = args
        store = getattr(self, "_store", None)
        device = getattr(self, "_device", 0)
        if store is not None:
            i = [str(p) for p in store.chunk_paths].index(str(chunk_path))
            root = store.open_chunk(i)
            res = self.fit_block((all_lon, all_lat), device=device, group_root=root)
            store.close_chunk(i, root)
            return res
        return self.fit_block((all_lon, all_lat), device=device)

    def fit_cube(self, store_name='run/test_cube', nproc=1, timeout=None, blocks_per_gpu=1, devices=None):
        """Fit every pixel and write the store.  `nproc` = number of GPUs: one process per GPU, one chunk
        file per process (main.py:476-526).  With `blocks_per_gpu` = 1 every process fits one contiguous pixel
        block; with more, the cube is cut into `nproc * blocks_per_gpu` contiguous blocks that the processes
        take from a shared queue as they finish (pixels differ in cost: nlive grows with the SNR and the
        number of model runs with the number of components).  `devices` lists the CUDA device
        of each process (default 0 .. nproc-1)."""
        n_lon = self.stack.spatial_shape[0]
        if nproc > n_lon:
            raise ValueError(f'The pixel width of the image in longitude ({n_lon}) ' +
                             f'must be greater than or equal to the number of processes ({nproc}).')
        if blocks_per_gpu < 1:
            raise ValueError(f'blocks_per_gpu must be positive: {blocks_per_gpu}')
        devices = list(range(nproc)) if devices is None else [int(d) for d in devices]
        if len(devices) != nproc:
            raise ValueError(f'devices must name one CUDA device per process: {devices}')
        store = HdfStore(store_name, nchunks=nproc)
        store.insert_header(self.stack)
        store.insert_fitter_pars(self)
        store.insert_model_metadata(self.runner_cls)
        n_blocks = store.nchunks * int(blocks_per_gpu)
        n_pix = int(np.prod(self.stack.spatial_shape))
        indices = get_block_indices(self.stack.spatial_shape, min(n_blocks, n_pix))
        self._store = store
        results = []
        if store.nchunks == 1:
            self._device = devices[0]
            lon = np.concatenate([b[0] for b in indices])
            lat = np.concatenate([b[1] for b in indices])
            results.append(self.fit((lon, lat), store.chunk_paths[0]))
        else:
            import multiprocessing as mp
            ctx = mp.get_context('spawn')
            queue = None
            if blocks_per_gpu > 1:
                queue = ctx.Queue()
                for j in range(len(indices)):
                    queue.put(j)
                for _ in range(store.nchunks):
                    queue.put(None)                 # one stop mark per process
            procs = [ctx.Process(target=_fit_worker,
                                 args=(self, i, devices[i], indices if queue is not None else indices[i],
                                       str(store.chunk_paths[i]), queue))
                     for i in range(store.nchunks)]
            for proc in procs:
                proc.start()
            failed = None
            for proc in procs:
                proc.join(timeout)
                if proc.exitcode not in (0, None) and failed is None:
                    failed = proc.exitcode
            if failed is not None:
                for proc in procs:          # do not leave the other workers running behind the error
                    if proc.is_alive():
                        proc.terminate()
                raise RuntimeError(f'GPU worker failed with exit code {failed}')
        store.link_files()
        store.close()
        self._store = None
        self.stats['results'] = results
        return results


def _fit_worker(fitter, i, device, indices, chunk_path, queue=None):
    """Process `i` of fit_cube: fits its block (or blocks taken from `queue`) on `device` into chunk `i`."""
    fitter._device = device
    fitter._store = HdfStore.__new__(HdfStore)
    # lightweight re-attachment to the already created store directory
    from pathlib import Path
    store = fitter._store
    store.store_dir = Path(chunk_path).parent
    store.nchunks = max(i + 1, len(list(store.store_dir.glob('chunk*'))) or i + 1)
    store._open = True
    store.hdf = None
    os.environ.setdefault('HDF5_USE_FILE_LOCKING', 'FALSE')
    if queue is None:
        fitter.fit(indices, store.chunk_paths[i])
        return
    root = store.open_chunk(i)
    try:
        while True:
            j = queue.get()
            if j is None:
                break
            fitter.fit_block(indices[j], device=device, group_root=root)
    finally:
        store.close_chunk(i, root)
