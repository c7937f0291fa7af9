"""Device-pointer entry points of the C ABI (include/nestfit_b200.h) called with torch-owned device
buffers and a caller stream: every one must agree with its host-buffer twin."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _p(t):
    return C.c_void_p(t.data_ptr())


def test_device_pointer_entry_points(nb):
    import torch
    from nestfit_b200 import _lib
    lib = _lib.load()
    n = C.c_int()
    assert lib.nf_device_count(C.byref(n)) == 0 and n.value >= 1
    rng = np.random.default_rng(12)
    ut = nb.get_irdc_priors()
    dev = "cuda:0"
    stream = torch.cuda.Stream()
    n_pix, n_chan, ncomp, B = 3, 300, 2, 96
    xs = [orc.bench_axis(1, n_chan, 0.2), orc.bench_axis(2, n_chan, 0.2)]
    data = rng.normal(0, 0.1, (n_pix, 2, n_chan)).astype(np.float32)
    noise = np.full((n_pix, 2), 0.1)
    # nf_prior_transform (device) against nf_prior_transform_host
    U = rng.uniform(size=(B, 6 * ncomp))
    want_P = ut.transform_batch(U.copy(), ncomp)
    d_u = torch.from_numpy(U.copy()).to(dev)
    with torch.cuda.stream(stream):
        _lib.check(lib.nf_prior_transform(ut.handle(0), _p(d_u), B, ncomp, C.c_void_p(stream.cuda_stream)), "pt")
    stream.synchronize()
    got_P = d_u.cpu().numpy()
    ok = np.isfinite(want_P)
    assert (np.isfinite(got_P) == ok).all()
    np.testing.assert_array_equal(got_P[ok], want_P[ok])
    P = want_P.copy()
    P[~np.isfinite(P).all(axis=1)] = P[np.isfinite(P).all(axis=1)][0]
    # nf_pixels_create_from_device against nf_pixels_create
    blk = nb.PixelBlock("ammonia", xs, data, noise, trans_ids=[1, 2])
    d_data, d_noise = torch.from_numpy(data).to(dev), torch.from_numpy(noise).to(dev)
    nu_min = np.array([x[0] for x in xs]); nu_chan = np.array([x[1] - x[0] for x in xs])
    tid = np.array([1, 2], dtype=np.int32)
    h = C.c_void_p()
    _lib.check(lib.nf_pixels_create_from_device(0, _lib.NF_MODEL_NH3, n_pix, 2, n_chan, _lib.ptr(nu_min), _lib.ptr(nu_chan),
                                                _lib.ptr(tid), None, _p(d_data), _p(d_noise), C.byref(h)), "create_from_device")
    null = np.empty(n_pix)
    _lib.check(lib.nf_pixels_null_lnz(h, _lib.ptr(null)), "null")
    np.testing.assert_allclose(null, blk.null_lnZ(), rtol=1e-12)
    # nf_nh3_loglike / nf_nh3_predict (device, FP64 and FP32 parameters, explicit pixel map)
    pix = rng.integers(0, n_pix, B).astype(np.int32)
    d_pix = torch.from_numpy(pix).to(dev)
    for dt, code in ((np.float64, _lib.NF_F64), (np.float32, _lib.NF_F32)):
        Pd = np.ascontiguousarray(P.astype(dt))
        d_p = torch.from_numpy(Pd).to(dev)
        d_l = torch.empty(B, dtype=torch.float64, device=dev)
        d_pr = torch.empty((B, 2, n_chan), dtype=torch.float32, device=dev)
        with torch.cuda.stream(stream):
            _lib.check(lib.nf_nh3_loglike(h, _p(d_p), code, _p(d_pix), 0, B, ncomp, 0, _p(d_l),
                                          C.c_void_p(stream.cuda_stream)), "nh3_loglike")
            _lib.check(lib.nf_nh3_predict(h, _p(d_p), code, B, ncomp, 0, _p(d_pr), C.c_void_p(stream.cuda_stream)), "nh3_predict")
        stream.synchronize()
        np.testing.assert_array_equal(d_l.cpu().numpy(), blk.loglike(Pd, ncomp, pix_of_vec=pix))
        np.testing.assert_array_equal(d_pr.cpu().numpy(), blk.predict(Pd, ncomp))
    _lib.check(lib.nf_pixels_free(h), "free")
    # N2H+ and Gaussian device entry points
    xn = [np.sort(orc.N2HP_NU[0] * (1 - (np.arange(n_chan) - 0.5 * (n_chan - 1)) * 0.1 / orc.CKMS))]
    bn = nb.PixelBlock("diazenylium", xn, data[:, :1], 0.1, trans_ids=[1])
    Pn = np.concatenate([np.sort(rng.uniform(-5, 5, (B, 2)), axis=1), rng.uniform(3, 20, (B, 2)), rng.uniform(-1.5, 1, (B, 2)),
                         rng.uniform(0.1, 1.2, (B, 2))], axis=1)
    d_p = torch.from_numpy(Pn).to(dev); d_l = torch.empty(B, dtype=torch.float64, device=dev)
    d_pr = torch.empty((B, 1, n_chan), dtype=torch.float32, device=dev)
    _lib.check(lib.nf_n2hp_loglike(bn.handle, _p(d_p), _lib.NF_F64, _p(d_pix), 0, B, 2, _p(d_l), None), "n2hp_loglike")
    _lib.check(lib.nf_n2hp_predict(bn.handle, _p(d_p), _lib.NF_F64, B, 2, _p(d_pr), None), "n2hp_predict")
    torch.cuda.synchronize()
    np.testing.assert_array_equal(d_l.cpu().numpy(), bn.loglike(Pn, 2, pix_of_vec=pix))
    np.testing.assert_array_equal(d_pr.cpu().numpy(), bn.predict(Pn, 2))
    bg = nb.PixelBlock("gaussian", xn, data[:, :1], 0.1, rest_freq=orc.N2HP_NU[0])
    Pg = np.concatenate([rng.uniform(-10, 10, (B, 3)), rng.uniform(0.2, 2, (B, 3)), rng.uniform(0.1, 3, (B, 3))], axis=1)
    d_p = torch.from_numpy(Pg).to(dev)
    _lib.check(lib.nf_gauss_loglike(bg.handle, _p(d_p), _lib.NF_F64, _p(d_pix), 0, B, 3, _p(d_l), None), "gauss_loglike")
    _lib.check(lib.nf_gauss_predict(bg.handle, _p(d_p), _lib.NF_F64, B, 3, _p(d_pr), None), "gauss_predict")
    torch.cuda.synchronize()
    np.testing.assert_array_equal(d_l.cpu().numpy(), bg.loglike(Pg, 3, pix_of_vec=pix))
    np.testing.assert_array_equal(d_pr.cpu().numpy()[:, 0], bg.predict(Pg, 3)[:, 0])
    # error paths: wrong model for the entry point, ncomp too large
    assert lib.nf_nh3_loglike(bg.handle, _p(d_p), _lib.NF_F64, None, 1, B, 1, 0, _p(d_l), None) == -1
    assert lib.nf_n2hp_loglike(bn.handle, _p(d_p), _lib.NF_F64, None, 1, B, 5, _p(d_l), None) == -1


def test_host_call_pageable_and_pinned_buffers_agree(nb):
    """The host-buffer call takes page-locked buffers directly and bounces pageable ones (plain numpy arrays) through
    its own page-locked ring (csrc/nf_capi.cu: run_host): both give the device call's lnL bit for bit, over several
    131 072-vector chunks with a ragged last one, with the implicit and with an explicit vector -> pixel map, and
    with the three arguments pinned in any combination."""
    import torch
    from nestfit_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(5)
    ut = nb.get_irdc_priors()
    n_pix, n_chan, ncomp, vpp = 37, 128, 1, 8192
    B = n_pix * vpp - 1000                     # 2.3 chunks; the last pixel is short of vectors
    xs = [orc.bench_axis(1, n_chan, 0.4), orc.bench_axis(2, n_chan, 0.4)]
    data = rng.normal(0, 0.1, (n_pix, 2, n_chan)).astype(np.float32)
    blk = nb.PixelBlock("ammonia", xs, data, 0.1, trans_ids=[1, 2])
    P = ut.transform_batch(rng.uniform(size=(B, 6)), ncomp)
    P[~np.isfinite(P).all(axis=1)] = P[np.isfinite(P).all(axis=1)][0]
    P = np.ascontiguousarray(P.astype(np.float32))
    pix = (np.arange(B) // vpp).astype(np.int32)
    # the device call is the reference
    d_p, d_x = torch.from_numpy(P).cuda(), torch.from_numpy(pix).cuda()
    d_l = torch.empty(B, dtype=torch.float64, device="cuda")
    _lib.check(lib.nf_nh3_loglike(blk.handle, _p(d_p), _lib.NF_F32, None, vpp, B, ncomp, 0, _p(d_l), None), "device call")
    torch.cuda.synchronize()
    want = d_l.cpu().numpy()
    assert np.isfinite(want).all()

    def pinned(a):
        t = torch.from_numpy(a.copy()).pin_memory()
        assert t.is_pinned()
        return t

    tp, tx, tl = pinned(P), pinned(pix), torch.empty(B, dtype=torch.float64).pin_memory()
    for pin_p, pin_x, pin_l in ((0, 0, 0), (1, 1, 1), (1, 0, 0), (0, 1, 1), (0, 0, 1)):
        for explicit in (False, True):
            out = np.full(B, np.nan)
            tl.fill_(float("nan"))
            a_p = C.c_void_p(tp.data_ptr()) if pin_p else _lib.ptr(P)
            a_x = None if not explicit else (C.c_void_p(tx.data_ptr()) if pin_x else _lib.ptr(pix))
            a_l = C.c_void_p(tl.data_ptr()) if pin_l else _lib.ptr(out)
            _lib.check(lib.nf_nh3_loglike_host(blk.handle, a_p, _lib.NF_F32, a_x, 0 if explicit else vpp, B, ncomp, 0, a_l),
                       "host call")
            got = tl.numpy() if pin_l else out
            np.testing.assert_array_equal(got, want)
    blk.close()
