"""GPU test of the cube path: CubeFitter.fit_cube on a small synthetic cube,
ncomp escalation (main.py:450-469), NaN-pixel skipping (main.py:438-441) and the
store layout (docs/store_spec.rst)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_fit_cube_small(nb, tmp_path):
    from nestfit_b200.synth import make_synth_stack
    from nestfit_b200.store import HdfStore
    from nestfit_b200.models import ammonia
    ut = nb.get_irdc_priors()
    ncomp_map = np.zeros((6, 4), dtype=int)
    ncomp_map[2:4] = 1
    ncomp_map[4:] = 2
    stack = make_synth_stack((6, 4), ut, ncomp_map=ncomp_map, n_chan=400, dv=0.158, noise=0.1, seed=3)
    # deterministic, well separated truths so the expected nbest is unambiguous
    blk = nb.PixelBlock("ammonia", [c.xarr for c in stack.cubes], np.zeros((1, 2, 400), np.float32), 1.0,
                        trans_ids=[1, 2])
    t1 = np.array([[0.3, 14.0, 6.0, 14.6, 0.45, 0.0]])
    t2 = np.array([[-1.5, 1.5, 12, 15, 5, 6, 14.6, 14.8, 0.35, 0.5, 0, 0]], dtype=float)
    rng = np.random.default_rng(8)
    for (lo, hi), t, nc in (((2, 4), t1, 1), ((4, 6), t2, 2)):
        clean = blk.predict(t, nc)[0]
        for c in (0, 1):
            stack.cubes[c].data[lo:hi] = clean[c] + rng.normal(0, 0.1, (hi - lo, 4, 400))
    stack.cubes[0].data[0, 0, 10] = np.nan                 # a blanked pixel
    fitter = nb.CubeFitter(stack, ut, ammonia.AmmoniaRunner, ncomp_max=2, mn_kwargs={'nlive': 120}, n_prop=32,
                           lnZ_thresh=11, seed=5)
    res = fitter.fit_cube(str(tmp_path / 'cube'), nproc=1)[0]
    nbest = res['nbest'].reshape(6, 4)
    assert nbest[0, 0] == -1                                # NaN pixel skipped
    assert np.all(nbest[:2].ravel()[1:] == 0)               # noise only
    assert np.all(nbest[2:4] == 1) and np.all(nbest[4:] == 2)
    # two-component fits were only attempted where one component was significant
    assert np.isnan(res['lnZ'][:8, 2]).all() and np.isfinite(res['lnZ'][8:, 2]).all()
    store = HdfStore(str(tmp_path / 'cube'))
    assert store.hdf.attrs['n_max_components'] == 2 and store.hdf.attrs['lnZ_threshold'] == 11
    assert store.hdf.attrs['model_name'] == 'ammonia' and store.hdf.attrs['naxis1'] == 6
    groups = list(store.iter_pix_groups())
    assert len(groups) == 23
    g = store.hdf['/pix/5/3']
    assert g.attrs['nbest'] == 2 and g.attrs['i_lon'] == 5 and '1' in g and '2' in g
    run = g['2']
    assert run.attrs['ncomp'] == 2 and run.attrs['n_params'] == 12 and run['marginals'].shape == (15, 12)
    assert run['posteriors'].shape == (run.attrs['n_samples'], 14)
    assert '2' not in store.hdf['/pix/0/1']
    # fitted velocities of the 2-component pixels recover the truth ordering
    v = run['bestfit_params'][:2]
    assert abs(v[0] + 1.5) < 0.2 and abs(v[1] - 1.5) < 0.2
    store.close()


def _two_block_problem(nb):
    from nestfit_b200.synth import make_synth_stack
    from nestfit_b200.models import ammonia
    ut = nb.get_irdc_priors()
    ncomp_map = np.zeros((4, 4), dtype=int)
    ncomp_map[2:] = 1
    stack = make_synth_stack((4, 4), ut, ncomp_map=ncomp_map, n_chan=400, dv=0.158, noise=0.1, seed=4)
    blk = nb.PixelBlock("ammonia", [c.xarr for c in stack.cubes], np.zeros((1, 2, 400), np.float32), 1.0,
                        trans_ids=[1, 2])
    t1 = np.array([[0.3, 14.0, 6.0, 14.6, 0.45, 0.0]])
    clean = blk.predict(t1, 1)[0]
    blk.close()
    rng = np.random.default_rng(8)
    for c in (0, 1):
        stack.cubes[c].data[2:] = clean[c] + rng.normal(0, 0.1, (2, 4, 400))
    return nb.CubeFitter(stack, ut, ammonia.AmmoniaRunner, ncomp_max=1, mn_kwargs={'nlive': 100}, seed=6)


def _check_two_block_store(path):
    from nestfit_b200.store import HdfStore
    store = HdfStore(path)
    assert store.nchunks == 2 and all(p.exists() for p in store.chunk_paths)
    groups = list(store.iter_pix_groups())
    assert len(groups) == 16
    nbest = np.full((4, 4), -9)
    for g in groups:
        nbest[g.attrs['i_lon'], g.attrs['i_lat']] = g.attrs['nbest']
    assert np.all(nbest[:2] == 0) and np.all(nbest[2:] == 1)
    store.close()


def test_fit_cube_two_gpus(nb, tmp_path):
    """fit_cube(nproc=2): one spawned process per GPU, contiguous pixel blocks, one chunk file each,
    linked into the table (main.py:476-526, 313-322).  Skipped on single-GPU boxes."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    fitter = _two_block_problem(nb)
    fitter.fit_cube(str(tmp_path / 'cube2'), nproc=2)
    _check_two_block_store(str(tmp_path / 'cube2'))


def test_fit_cube_dynamic_blocks(nb, tmp_path):
    """fit_cube(nproc=2, blocks_per_gpu=3): six contiguous blocks handed out from a queue to two worker
    processes (both on device 0 here so that the test also runs on a single-GPU box); every pixel is fitted
    exactly once and lands in the chunk of the process that took its block."""
    fitter = _two_block_problem(nb)
    fitter.fit_cube(str(tmp_path / 'cube3'), nproc=2, blocks_per_gpu=3, devices=[0, 0])
    _check_two_block_store(str(tmp_path / 'cube3'))
    with pytest.raises(ValueError):
        fitter.fit_cube(str(tmp_path / 'cube4'), nproc=2, devices=[0])


def test_fit_cube_rank_concurrent_blocks(nb, tmp_path):
    """fit_cube_rank, the SPMD form used under torchrun, with four blocks in flight from four host threads on one
    device: blocks of different brightness (different live-set sizes, i.e. different shared-memory needs of the
    sampler kernels) run concurrently on their own streams; every pixel is fitted exactly once, into one chunk."""
    from nestfit_b200.synth import make_synth_stack
    from nestfit_b200.models import ammonia
    ut = nb.get_irdc_priors()
    ncomp_map = np.zeros((8, 4), dtype=int)
    ncomp_map[4:] = 1
    noise = 0.05 + 0.25 * np.arange(8)[:, None] / 7.0 * np.ones((8, 4))      # SNR, hence nlive, differs per block
    stack = make_synth_stack((8, 4), ut, ncomp_map=ncomp_map, n_chan=400, dv=0.158, noise=noise, seed=7)
    fitter = nb.CubeFitter(stack, ut, ammonia.AmmoniaRunner, ncomp_max=1, mn_kwargs={'nlive': 60}, seed=9)
    res = fitter.fit_cube_rank(str(tmp_path / 'spmd'), 0, 1, blocks_per_gpu=4, device=0, concurrent_blocks=4)
    assert sorted(j for r in res for j in r['blocks']) == [0, 1, 2, 3]
    store = nb.HdfStore(str(tmp_path / 'spmd'))
    groups = list(store.iter_pix_groups())
    assert len(groups) == 32 and all('1' in g for g in groups)
    nbest = np.full((8, 4), -9)
    for g in groups:
        nbest[g.attrs['i_lon'], g.attrs['i_lat']] = g.attrs['nbest']
    assert (nbest[:4] == 0).all() and (nbest[4:] == 1).mean() > 0.7
    store.close()
