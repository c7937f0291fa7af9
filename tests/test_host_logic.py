"""CPU tests of the host-side logic around the hot path: store tree, cube/stack
array contract, pixel-block partitions and the world_size-2 (gloo) multi-rank path."""
import os
import socket
import subprocess
import sys
import textwrap
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def test_memgroup_tree_and_persistence(tmp_path, nb):
    from nestfit_b200.store import MemGroup
    g = MemGroup()
    p = g.require_group('/pix/3/4')
    s = p.create_group('1')
    s.attrs['ncomp'] = 1
    s.attrs['marg_quantiles'] = np.arange(3.0)
    s.create_dataset('posteriors', data=np.ones((4, 8), dtype='f4'))
    with pytest.raises(ValueError):
        p.create_group('1')            # the reference's create_group raises on re-fit (main.py:454)
    assert '/pix/3/4/1' in g and '1' in p and '2' not in p
    g.save(tmp_path / 't.npz')
    h = MemGroup.load(tmp_path / 't.npz')
    assert h['/pix/3/4/1'].attrs['ncomp'] == 1
    np.testing.assert_array_equal(h['/pix/3/4/1'].attrs['marg_quantiles'], np.arange(3.0))
    assert h['pix/3/4/1/posteriors'].dtype == np.float32 and h['pix/3/4/1/posteriors'].shape == (4, 8)


def test_hdfstore_layout_and_links(tmp_path, nb):
    from nestfit_b200.store import HdfStore
    from nestfit_b200.models import ammonia
    store = HdfStore(str(tmp_path / 'run'), nchunks=2)
    assert store.store_dir.name == 'run.store' and len(store.chunk_paths) == 2
    store.insert_model_metadata(ammonia.AmmoniaRunner)
    assert store.hdf.attrs['model_name'] == 'ammonia' and store.hdf.attrs['n_params'] == 6
    for i, lon in enumerate((0, 1)):
        root = store.open_chunk(i)
        g = root.require_group(f'/pix/{lon}/5')
        g.attrs['nbest'] = i
        g.create_group('1').attrs['global_lnZ'] = -1.0 * i
        store.close_chunk(i, root)
    store.link_files()
    assert sorted(g.attrs['nbest'] for g in store.iter_pix_groups()) == [0, 1]
    assert store.find_first_valid_group().attrs['global_lnZ'] in (0.0, -1.0)
    store.close()
    again = HdfStore(str(tmp_path / 'run'))
    assert again.nchunks == 2 and again.model is ammonia
    again.close()


def test_cube_stack_contract(nb):
    rng = np.random.default_rng(0)
    x = np.linspace(2.369e10, 2.3695e10, 50)
    cube = nb.DataCube.from_arrays(rng.normal(size=(4, 3, 50)), x[::-1].copy(), 0.35, trans_id=1)   # descending axis in
    assert cube.xarr[1] > cube.xarr[0] and cube.shape == (4, 3, 50) and cube.spatial_shape == (4, 3)
    noise = nb.NoiseMap(np.full((3, 4), 0.2))        # image (lat, lon) order -> transposed (main.py:41-44)
    cube2 = nb.DataCube.from_arrays(rng.normal(size=(4, 3, 50)), x, noise, trans_id=2)
    stack = nb.CubeStack([cube, cube2])
    spec_data, has_nans = stack.get_spec_data(1, 2)
    assert not has_nans and len(spec_data) == 2 and spec_data[1][3] == 2 and spec_data[0][2] == 0.35
    assert stack.get_max_snr(1, 2) > 0
    cube.data[0, 0, 3] = np.nan
    assert stack.get_spec_data(0, 0)[1]
    data, nz, valid = stack.block_arrays(np.array([0, 1]), np.array([0, 2]))
    assert data.shape == (2, 2, 50) and nz.shape == (2, 2) and list(valid) == [False, True]
    assert nb.NoiseMapUniform(0.35).get_noise(0, 0) == 0.35      # reference test_main.py:32-35


def test_store_header_round_trip(tmp_path, nb):
    """insert_header / read_header (main.py:329-352): the map header is the cube header cut to two axes."""
    rng = np.random.default_rng(0)
    x = np.linspace(2.369e10, 2.3695e10, 50)
    hdr = {'SIMPLE': True, 'BITPIX': -32, 'NAXIS': 3, 'NAXIS1': 4, 'NAXIS2': 3, 'NAXIS3': 50, 'CRPIX1': 2.0,
           'CDELT1': -1e-3, 'CTYPE1': 'GLON-CAR', 'CRVAL1': 30.0, 'CTYPE3': 'FREQ', 'BUNIT': 'K'}
    cube = nb.DataCube.from_arrays(rng.normal(size=(4, 3, 50)), x, 0.2, trans_id=1, header=hdr)
    store = nb.HdfStore(str(tmp_path / 'hdr'))
    store.insert_header(nb.CubeStack([cube]))
    full, simple = dict(store.read_header(full=True)), dict(store.read_header(full=False))
    assert full['NAXIS3'] == 50 and full['CTYPE3'] == 'FREQ' and full['BUNIT'] == 'K'
    assert simple['NAXIS'] == 2 and simple['WCSAXES'] == 2 and simple['CTYPE1'] == 'GLON-CAR'
    assert 'NAXIS3' not in simple and 'CTYPE3' not in simple
    assert store.hdf.attrs['naxis1'] == 4 and store.hdf.attrs['naxis2'] == 3
    store.close()


def test_partitions(nb):
    from nestfit_b200.parallel import block_bounds
    idx = nb.get_multiproc_indices((5, 3), 2)          # reference row striping (main.py:565-571)
    assert list(idx[0][0]) == [0, 0, 0, 2, 2, 2, 4, 4, 4] and list(idx[1][0]) == [1, 1, 1, 3, 3, 3]
    blocks = nb.get_block_indices((5, 3), 4)
    lon = np.concatenate([b[0] for b in blocks])
    lat = np.concatenate([b[1] for b in blocks])
    assert sorted(zip(lon, lat)) == [(i, j) for i in range(5) for j in range(3)]
    assert all(np.all(np.diff(b[0] * 3 + b[1]) == 1) for b in blocks if b[0].size > 1)   # contiguous
    for n, w in ((10, 3), (7, 8), (1024, 8)):
        bb = block_bounds(n, w)
        assert bb[0][0] == 0 and bb[-1][1] == n and all(a[1] == b[0] for a, b in zip(bb, bb[1:]))
        sizes = [b - a for a, b in bb]
        assert max(sizes) - min(sizes) <= 1


WORKER = textwrap.dedent('''
    import os, sys
    import numpy as np
    sys.path.insert(0, %r)
    from nestfit_b200.parallel import init_distributed, my_block, gather_blocks, max_over_ranks, rank_info
    rank, local_rank, world = rank_info()
    dist = init_distributed("gloo")
    n_pix = 37
    a, b = my_block(n_pix, rank, world)
    # stand-in for the per-rank fit: lnZ of pixel i is -i, nbest is i %% 3
    lnz = -np.arange(a, b, dtype=np.float64)[:, None] * np.ones((1, 3))
    nbest = (np.arange(a, b) %% 3).astype(np.int64)
    full_lnz = gather_blocks(lnz, n_pix, dist)
    full_nb = gather_blocks(nbest, n_pix, dist)
    t = max_over_ranks(1.0 + rank, dist)
    dist.barrier()
    if rank == 0:
        assert full_lnz.shape == (n_pix, 3) and np.array_equal(full_lnz[:, 0], -np.arange(n_pix))
        assert np.array_equal(full_nb, np.arange(n_pix) %% 3)
        assert t == float(world)
        print("GLOO_OK", world)
    dist.destroy_process_group()
''')


def test_two_rank_gloo_partition_and_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % str(ROOT))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert res.returncode == 0, res.stderr[-2000:]
    assert "GLOO_OK 2" in res.stdout


def test_cubefitter_pickles_for_spawned_workers(nb):
    """fit_cube(nproc > 1) hands the fitter to one spawned process per GPU (main.py:515-523): the
    fitter, its priors and the store must pickle (device handles are rebuilt per process)."""
    import pickle
    from nestfit_b200.synth import velocity_axis_hz
    from nestfit_b200.main import DataCube, CubeStack
    from nestfit_b200.models import ammonia
    ut = nb.get_irdc_priors()
    ut._handles = {0: object()}          # stands in for a live device handle
    x = [velocity_axis_hz(1, 100, 0.3), velocity_axis_hz(2, 100, 0.3)]
    stack = CubeStack([DataCube.from_arrays(np.zeros((2, 2, 100)), x[t], 0.1, trans_id=t + 1) for t in range(2)])
    fitter = nb.CubeFitter(stack, ut, ammonia.AmmoniaRunner, ncomp_max=2)
    fitter._store, fitter._device = open(__file__), 3      # an open store (unpicklable) and a device binding: left behind
    try:
        clone = pickle.loads(pickle.dumps(fitter))
        assert not hasattr(clone, '_store') and not hasattr(clone, '_device')
    finally:
        ut._handles = {}
    assert clone.utrans._handles == {} and clone.utrans.n_param == 6 and clone.ncomp_max == 2
    assert clone.stack.cubes[1].trans_id == 2


def test_fit_cube_argument_checks():
    """fit_cube validates its partitioning arguments before touching a device (main.py:476-486)."""
    import nestfit_b200 as nb
    from nestfit_b200.models import ammonia

    class _Stack:
        spatial_shape = (3, 5)
    fitter = nb.CubeFitter(_Stack(), None, ammonia.AmmoniaRunner)
    with pytest.raises(ValueError, match='number of processes'):
        fitter.fit_cube('unused', nproc=4)
    with pytest.raises(ValueError, match='blocks_per_gpu'):
        fitter.fit_cube('unused', nproc=2, blocks_per_gpu=0)
    with pytest.raises(ValueError, match='one CUDA device per process'):
        fitter.fit_cube('unused', nproc=2, devices=[0])


def test_public_names_match_reference_package():
    """Every name the reference exports at package level (nestfit/__init__.py:8-62) exists here."""
    import nestfit_b200 as nb
    names = """Distribution Prior ConstantPrior DuplicatePrior OrderedPrior SpacedPrior CenSepPrior
        ResolvedCenSepPrior ResolvedPlacementPrior PriorTransformer Spectrum Runner Dumper run_multinest
        NoiseMap NoiseMapUniform DataCube CubeStack HdfStore CubeFitter take_by_components apply_circular_mask
        get_indep_info_kernel aggregate_run_attributes convolve_evidence extended_masked_evidence
        aggregate_run_products aggregate_run_pdfs convolve_post_pdfs quantize_conv_marginals deblend_hf_intensity
        generate_predicted_profiles postprocess_run amm_predict AmmoniaSpectrum AmmoniaRunner nnhp_predict
        DiazenyliumSpectrum DiazenyliumRunner gauss_predict GaussianRunner""".split()
    missing = [n for n in names if not hasattr(nb, n)]
    assert not missing, missing


def test_synthetic_data_helpers(nb):
    """ParamSampler / add_noise_to_cube / make_fake_header / test axes (nestfit/synth_spectra.py:149-192,243-267)."""
    from nestfit_b200 import synth
    ps = synth.ParamSampler(rng=np.random.default_rng(1))
    draws = np.array([ps.draw() for _ in range(200)])
    assert draws.shape == (200, 12) and (draws[:, 0] == 0).all()
    assert (draws[:, 1] >= 0.16).all() and (draws[:, 1] <= 3).all()
    assert (draws[:, 2:4] >= 3).all() and (draws[:, 2:4] <= 30).all() and (draws[:, 10:] == 0).all()
    x11, x22 = synth.test_axes()
    assert x11.shape == (380,) and x11[1] > x11[0] and abs(x11.mean() / synth.NU[1] - 1) < 1e-6
    assert synth.TEST_PARAMS[0].shape == (12,) and synth.TEST_PARAMS[1][3] == 12.0
    cube = np.zeros((3, 4, 380))
    noisy = synth.add_noise_to_cube(cube, 0.2, rng=np.random.default_rng(2))
    assert abs(noisy.std() - 0.2) < 0.01 and np.array_equal(synth.add_noise_to_cube(cube, 0.0), cube)
    hdr = synth.make_fake_header(cube, x11, synth.NU[1])
    assert hdr['CRPIX3'] == 380 and hdr['CTYPE3'] == 'FREQ' and hdr['RESTFRQ'] == synth.NU[1] and hdr['CDELT3'] > 0
    # the header feeds DataCube.from_arrays and the store's map header
    dc = nb.DataCube.from_arrays(cube, x11, 0.2, trans_id=1, header=hdr)
    assert dc.simple_header['CTYPE1'] == 'RA---CAR' and 'CTYPE3' not in dc.simple_header


STUB_FITTER = """
import time
from nestfit_b200.main import CubeFitter


class StubStack:
    # just enough of a CubeStack for the store header and the block partition
    spatial_shape = (6, 5)
    shape = (6, 5, 8)
    simple_header = {'NAXIS': 2, 'NAXIS1': 6, 'NAXIS2': 5}
    full_header = {'NAXIS': 3, 'NAXIS1': 6, 'NAXIS2': 5, 'NAXIS3': 8}


class StubFitter(CubeFitter):
    def fit_block(self, indices, device=0, group_root=None, verbose=False):
        # stands in for the GPU fit of one block: records which worker (device) took each pixel
        lon, lat = indices
        if getattr(self, 'fail_on_device', None) == device:
            raise RuntimeError('stub failure')
        for i, j in zip(lon, lat):
            g = group_root.require_group(f'/pix/{i}/{j}')
            g.attrs['i_lon'], g.attrs['i_lat'], g.attrs['nbest'] = int(i), int(j), int(device)
        time.sleep(0.02 * len(lon))
        return {}
"""


def test_fit_cube_process_fanout_on_cpu(tmp_path, monkeypatch, nb):
    """fit_cube's process fan-out without a GPU (the block fit is stubbed): two spawned workers, static blocks
    and queue hand-out; every pixel is written exactly once, into the chunk of the worker that took its block,
    and the chunks are linked into the table."""
    import importlib
    from nestfit_b200.models import ammonia
    (tmp_path / 'nf_stubfit.py').write_text(STUB_FITTER)
    monkeypatch.syspath_prepend(str(tmp_path))          # spawned workers inherit sys.path
    stub = importlib.import_module('nf_stubfit')
    for k, name in ((1, 'static'), (4, 'queue')):
        fitter = stub.StubFitter(stub.StubStack(), None, ammonia.AmmoniaRunner, ncomp_max=1)
        fitter.fit_cube(str(tmp_path / name), nproc=2, blocks_per_gpu=k, devices=[0, 1])
        store = nb.HdfStore(str(tmp_path / name))
        assert store.nchunks == 2 and all(p.exists() for p in store.chunk_paths)
        groups = list(store.iter_pix_groups())
        assert sorted((g.attrs['i_lon'], g.attrs['i_lat']) for g in groups) == [(i, j) for i in range(6) for j in range(5)]
        assert {g.attrs['nbest'] for g in groups} == {0, 1}
        if k == 1:          # one contiguous half of the rows per worker
            assert {g.attrs['nbest'] for g in groups if g.attrs['i_lon'] < 3} == {0}
        store.close()
    # a worker that dies surfaces as an error in the caller (and the others are not left running)
    fitter = stub.StubFitter(stub.StubStack(), None, ammonia.AmmoniaRunner, ncomp_max=1)
    fitter.fail_on_device = 1
    with pytest.raises(RuntimeError, match='GPU worker failed'):
        fitter.fit_cube(str(tmp_path / 'broken'), nproc=2, blocks_per_gpu=2, devices=[0, 1])


def test_dumper_and_weighted_quantiles(nb):
    """Dumper (core.pyx:564-609): 15 unweighted quantiles of every parameter column (the last two columns of a
    posterior array are lnL and the weight), attributes and datasets into an h5py-like group."""
    from nestfit_b200.sampler import Dumper, MARG_COLS, MARG_QUANTILES, weighted_quantiles
    rng = np.random.default_rng(3)
    post = rng.normal(size=(4000, 5))
    group = nb.MemGroup()
    d = Dumper(group)
    assert not d.no_dump and len(d.marginal_cols) == 15 == len(d.quantiles) and MARG_COLS[4] == 'p50'
    m = d.calc_marginals(post)
    assert m.shape == (15, 3)
    np.testing.assert_allclose(m[0], post[:, :3].min(axis=0))
    np.testing.assert_allclose(m[8], post[:, :3].max(axis=0))
    np.testing.assert_allclose(m[4], np.median(post[:, :3], axis=0))
    np.testing.assert_allclose(m[10] - m[9], 2.0, atol=0.12)               # +-1 sigma of a unit normal
    d.append_attributes(ncomp=2, global_lnZ=-12.5)
    d.append_datasets(marginals=m, posteriors=post.astype('float32'))
    d.flush()
    assert group.attrs['ncomp'] == 2 and np.asarray(group['marginals'][...]).shape == (15, 3)
    assert np.asarray(group['posteriors'][...]).dtype == np.float32
    # equal weights: the weighted quantiles agree with the plain ones
    wq = weighted_quantiles(post[:, :3], np.full(4000, 1 / 4000), MARG_QUANTILES)
    np.testing.assert_allclose(wq[1:8], m[1:8], atol=0.03)
    # weights that keep only x > 0 move the median of the first column to the median of the positive half
    w = (post[:, 0] > 0).astype(float)
    np.testing.assert_allclose(weighted_quantiles(post[:, :1], w / w.sum(), [0.5])[0, 0],
                               np.median(post[post[:, 0] > 0, 0]), atol=0.02)


def test_fit_cube_rank_spmd_claims(tmp_path, monkeypatch, nb):
    """fit_cube_rank, the SPMD form of fit_cube for processes that already exist (one per GPU under torchrun):
    every block is claimed by exactly one rank through exclusive file creation, each rank writes its own chunk,
    rank 0 creates the store before and links the chunks after the barrier."""
    import importlib
    import threading
    from nestfit_b200.models import ammonia
    (tmp_path / 'nf_stubfit2.py').write_text(STUB_FITTER)
    monkeypatch.syspath_prepend(str(tmp_path))
    stub = importlib.import_module('nf_stubfit2')
    world = 3
    bar = threading.Barrier(world)
    out, errs = {}, []

    def rank_main(r):
        try:
            fitter = stub.StubFitter(stub.StubStack(), None, ammonia.AmmoniaRunner, ncomp_max=1)
            out[r] = fitter.fit_cube_rank(str(tmp_path / 'spmd'), r, world, blocks_per_gpu=4, device=r, barrier=bar.wait)
        except BaseException as exc:
            errs.append(exc)
            bar.abort()

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errs, errs
    blocks = sorted(j for r in out for res in out[r] for j in res['blocks'])
    assert blocks == list(range(12)) and all(len(out[r]) >= 1 for r in out)        # every block exactly once
    store = nb.HdfStore(str(tmp_path / 'spmd'))
    assert store.nchunks == world
    groups = list(store.iter_pix_groups())
    assert sorted((g.attrs['i_lon'], g.attrs['i_lat']) for g in groups) == [(i, j) for i in range(6) for j in range(5)]
    assert {g.attrs['nbest'] for g in groups} == {0, 1, 2}                         # written by all three ranks
    store.close()


def test_fit_cube_rank_failing_rank_reaches_the_barrier(tmp_path, monkeypatch, nb):
    """A rank whose block fit raises still arrives at the barrier that follows the fit (the other ranks of a torchrun
    job would otherwise hang in it) and raises afterwards; the healthy ranks return."""
    import importlib
    import threading
    from nestfit_b200.models import ammonia
    (tmp_path / 'nf_stubfit3.py').write_text(STUB_FITTER)
    monkeypatch.syspath_prepend(str(tmp_path))
    stub = importlib.import_module('nf_stubfit3')
    world = 2
    bar = threading.Barrier(world, timeout=60)
    out, errs = {}, {}

    class Failing(stub.StubFitter):
        def fit_block(self, indices, device=0, group_root=None, verbose=False):
            if device == 1:
                raise RuntimeError('device 1 is on fire')
            return super().fit_block(indices, device=device, group_root=group_root, verbose=verbose)

    def rank_main(r):
        try:
            fitter = Failing(stub.StubStack(), None, ammonia.AmmoniaRunner, ncomp_max=1)
            out[r] = fitter.fit_cube_rank(str(tmp_path / 'spmd_fail'), r, world, blocks_per_gpu=2, device=r, barrier=bar.wait)
        except BaseException as exc:
            errs[r] = exc

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
    assert not any(t.is_alive() for t in threads)
    assert 0 in out and list(errs) == [1] and 'on fire' in str(errs[1])       # no BrokenBarrierError anywhere


def test_slab_store_wave_round_trip(tmp_path, nb):
    """A finished wave (posterior pool + run offsets + attribute columns) written as a slab and read back through
    the reference's tree: /pix/<lon>/<lat>/<ncomp> groups with the dumper's attributes and datasets, lazily
    materialised; a later wave of the same (pixel, ncomp) replaces the earlier run; chunks are linked on demand."""
    from nestfit_b200.store import HdfStore, Wave, run_attr_columns, RUN_ATTR_COLUMNS, MARG_COLS
    rng = np.random.default_rng(4)
    ndim, n_run = 6, 3
    n_s = np.array([5, 9, 4])
    off = np.concatenate([[0], np.cumsum(n_s)])
    post = rng.normal(size=(off[-1], ndim + 2)).astype(np.float32)
    res = dict(lnZ=np.array([-10.0, -20.0, -30.0]), lnZ_err=np.full(3, 0.1), max_loglike=np.array([-5.0, -6.0, -7.0]),
               n_iter=np.array([100, 200, 300]), n_evals=np.array([1000, 2000, 3000]), truncated=np.zeros(3, bool))
    cols = run_attr_columns(ndim, 760, [50, 60, 70], [-100.0, -110.0, -120.0], res, n_s)
    assert tuple(cols) == RUN_ATTR_COLUMNS
    wave = Wave(1, ndim, 760, [0, 0, 2], [1, 3, 1], cols, off, post, rng.normal(size=(n_run, 15, ndim)),
                rng.normal(size=(n_run, ndim)), rng.normal(size=(n_run, ndim)))
    store = HdfStore(str(tmp_path / 'slab'), nchunks=2)
    sink = store.chunk_sink(1)
    sink.add_wave(wave)
    sink.add_pixels([0, 0, 2], [1, 3, 1], [1, 0, 1])
    # the repeat of pixel (0, 3): a second wave with one run
    cols2 = {k: v[1:2].copy() for k, v in cols.items()}
    cols2['global_lnZ'][:] = -19.0
    sink.add_wave(Wave(1, ndim, 760, [0], [3], cols2, [0, 2], post[:2], wave.marginals[1:2], wave.bestfit[1:2], wave.mapfit[1:2]))
    sink.close()
    store.link_files()
    groups = list(store.iter_pix_groups())
    assert sorted((g.attrs['i_lon'], g.attrs['i_lat'], g.attrs['nbest']) for g in groups) == [(0, 1, 1), (0, 3, 0), (2, 1, 1)]
    run = store.hdf['/pix/2/1/1']
    assert run.attrs['ncomp'] == 1 and run.attrs['n_params'] == ndim and run.attrs['n_samples'] == 4
    assert run.attrs['global_lnZ'] == -30.0 and run.attrs['n_live'] == 70 and list(run.attrs['marg_cols']) == MARG_COLS
    assert isinstance(run.attrs['n_samples'], int) and isinstance(run.attrs['BIC'], float)
    np.testing.assert_array_equal(run['posteriors'], post[off[2]:off[3]])
    np.testing.assert_array_equal(run['marginals'], wave.marginals[2])
    np.testing.assert_array_equal(run['bestfit_params'], wave.bestfit[2])
    assert store.hdf['/pix/0/3/1'].attrs['global_lnZ'] == -19.0 and store.hdf['/pix/0/3/1']['posteriors'].shape == (2, ndim + 2)
    assert store.find_first_valid_group().attrs['ncomp'] == 1
    store.close()
    again = HdfStore(str(tmp_path / 'slab'))          # linked state persists; the view is rebuilt from the slabs
    assert again.nchunks == 2 and len(list(again.iter_pix_groups())) == 3
    again.reset_pix_links()
    assert '/pix' not in again.hdf
    again.close()


def test_noise_maps_and_pbimg(nb):
    """NoiseMap keeps the image in (lon, lat) order and answers scalar and array indices; from_pbimg takes the plane of
    a 2-, 3- or 4-axis primary-beam image (main.py:39-65)."""
    img = np.arange(12.0).reshape(3, 4)                 # (lat, lon)
    nm = nb.NoiseMap(img)
    assert nm.shape == (4, 3) and nm.get_noise(2, 1) == img[1, 2]
    np.testing.assert_array_equal(nm.get_noise(np.array([0, 3]), np.array([2, 0])), [img[2, 0], img[0, 3]])
    pb = np.ones((1, 1, 3, 4)); pb[0, 0, 1, 2] = 0.5; pb[0, 0, 0, 0] = 0.0
    m = nb.NoiseMap.from_pbimg(0.1, pb)
    assert m.get_noise(2, 1) == pytest.approx(0.2) and np.isinf(m.get_noise(0, 0)) and m.get_noise(3, 2) == pytest.approx(0.1)
    assert nb.NoiseMap.from_pbimg(0.1, pb[0]).shape == (4, 3) and nb.NoiseMap.from_pbimg(0.1, pb[0, 0]).shape == (4, 3)
    with pytest.raises(ValueError, match='Cannot parse shape'):
        nb.NoiseMap.from_pbimg(0.1, np.ones(5))
    assert nb.NoiseMapUniform(0.3).shape is None


def test_convolve_nan_extend_ignores_kernel_sum(nb):
    """astropy's convolve(..., boundary='extend') normalises the kernel (normalize_kernel=True): an unnormalised
    kernel gives the same map (the conv_nbest threshold test must not scale with the kernel's sum)."""
    from nestfit_b200.postprocess import convolve_nan_extend, gaussian_kernel2d
    rng = np.random.default_rng(1)
    img = rng.normal(size=(9, 7)); img[2, 3] = np.nan
    k = gaussian_kernel2d(1.0)
    np.testing.assert_allclose(convolve_nan_extend(img, 7.5 * k), convolve_nan_extend(img, k), rtol=1e-13)
    assert np.isfinite(convolve_nan_extend(img, k)).all()                      # the NaN is interpolated over
    np.testing.assert_allclose(convolve_nan_extend(np.full((5, 5), 3.0), k), 3.0, rtol=1e-13)
