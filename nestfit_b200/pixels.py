"""`PixelBlock`: a block of pixels (spectra + noise) resident in HBM.

Thin Python owner of an opaque ``nf_pixels*`` handle (include/nestfit_b200.h).
It is the batched counterpart of the reference's per-pixel ``Spectrum`` /
``AmmoniaSpectrum`` objects (nestfit/core/core.pyx:486-520,
nestfit/models/ammonia.pyx:244-277): the cube's ``data[lon, lat, chan]`` slab
(nestfit/main.py:152) is uploaded once and every likelihood call references
pixels by index.
"""
import ctypes as C

import numpy as np

from . import _lib


class PixelBlock:
    def __init__(self, model, xarrs, data, noise, trans_ids=None, rest_freq=None, device=0):
        """
        model     : 'ammonia' | 'gaussian' | 'diazenylium'
        xarrs     : sequence of n_spec ascending uniform frequency axes [Hz], equal length
        data      : array [n_pix, n_spec, n_chan] (float32 or float64)
        noise     : array [n_pix, n_spec] rms per spectrum
        trans_ids : NH3 transition ids (1..9) / N2H+ transition ids (1..3) per spectrum
        rest_freq : Gaussian-model rest frequency [Hz]
        """
        from .core import check_uniform_axis
        lib = _lib.load()
        self.model_name = model
        self.model = {"ammonia": _lib.NF_MODEL_NH3, "gaussian": _lib.NF_MODEL_GAUSS,
                      "diazenylium": _lib.NF_MODEL_N2HP}[model]
        xarrs = [np.ascontiguousarray(x, dtype=np.float64) for x in xarrs]
        self.n_spec = len(xarrs)
        self.n_chan = int(xarrs[0].shape[0])
        for x in xarrs:
            if x.shape[0] != self.n_chan:
                raise ValueError("all spectra of a block must have the same number of channels")
            check_uniform_axis(x)
        self.xarrs = xarrs
        data = np.asarray(data)
        if data.dtype not in (np.float32, np.float64):
            data = data.astype(np.float64)
        data = np.ascontiguousarray(data)
        if data.ndim != 3 or data.shape[1] != self.n_spec or data.shape[2] != self.n_chan:
            raise ValueError(f"data must be [n_pix, {self.n_spec}, {self.n_chan}], got {data.shape}")
        self.n_pix = int(data.shape[0])
        noise = np.ascontiguousarray(np.broadcast_to(np.asarray(noise, dtype=np.float64),
                                                     (self.n_pix, self.n_spec)))
        nu_min = np.array([x[0] for x in xarrs], dtype=np.float64)
        nu_chan = np.array([x[1] - x[0] for x in xarrs], dtype=np.float64)
        tid = None
        rf = None
        if self.model in (_lib.NF_MODEL_NH3, _lib.NF_MODEL_N2HP):
            tid = np.ascontiguousarray(trans_ids, dtype=np.int32)
            if tid.shape != (self.n_spec,):
                raise ValueError("trans_ids must have one entry per spectrum")
        else:
            rf = np.ascontiguousarray(np.broadcast_to(np.asarray(rest_freq, dtype=np.float64), (self.n_spec,)))
        self.trans_ids, self.rest_freq = tid, rf
        self.device = device
        out = C.c_void_p()
        dtype = _lib.NF_F64 if data.dtype == np.float64 else _lib.NF_F32
        _lib.check(lib.nf_pixels_create(device, self.model, self.n_pix, self.n_spec, self.n_chan,
                                        _lib.ptr(nu_min), _lib.ptr(nu_chan), _lib.ptr(tid), _lib.ptr(rf),
                                        _lib.ptr(data), dtype, _lib.ptr(noise), C.byref(out)),
                   "nf_pixels_create")
        self.handle = out
        self.n_model = {_lib.NF_MODEL_NH3: 6, _lib.NF_MODEL_N2HP: 4, _lib.NF_MODEL_GAUSS: 3}[self.model]

    def close(self):
        if getattr(self, "handle", None):
            _lib.load().nf_pixels_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def null_lnZ(self):
        out = np.empty(self.n_pix, dtype=np.float64)
        _lib.check(_lib.load().nf_pixels_null_lnz(self.handle, _lib.ptr(out)), "nf_pixels_null_lnz")
        return out

    @staticmethod
    def _params(params):
        params = np.asarray(params)
        if params.dtype not in (np.float32, np.float64):
            params = params.astype(np.float64)
        params = np.ascontiguousarray(params)
        return params, (_lib.NF_F64 if params.dtype == np.float64 else _lib.NF_F32)

    def loglike(self, params, ncomp, pix_of_vec=None, vecs_per_pix=0, cold=False, lte=False, out=None):
        """Host-buffer likelihood call: params [B, n_model*ncomp] physical units -> lnL [B]."""
        lib = _lib.load()
        params, dt = self._params(params)
        if params.ndim != 2 or params.shape[1] != self.n_model * ncomp:
            raise ValueError(f'Invalid shape for ncomp={ncomp}: {params.shape}')
        B = params.shape[0]
        if out is None:
            out = np.empty(B, dtype=np.float64)
        pv = None
        if pix_of_vec is not None:
            pv = np.ascontiguousarray(pix_of_vec, dtype=np.int32)
            if pv.shape != (B,):
                raise ValueError("pix_of_vec must have one entry per vector")
            if B and (pv.min() < 0 or pv.max() >= self.n_pix):
                raise ValueError("pix_of_vec out of range")
        else:
            if vecs_per_pix < 1:
                vecs_per_pix = max(1, -(-B // self.n_pix))
            if (B + vecs_per_pix - 1) // vecs_per_pix > self.n_pix:
                raise ValueError("more vectors than pixels * vecs_per_pix")
        if self.model == _lib.NF_MODEL_NH3:
            flags = (_lib.NF_FLAG_COLD if cold else 0) | (_lib.NF_FLAG_LTE if lte else 0)
            rc = lib.nf_nh3_loglike_host(self.handle, _lib.ptr(params), dt, _lib.ptr(pv), vecs_per_pix, B,
                                         ncomp, flags, _lib.ptr(out))
        elif self.model == _lib.NF_MODEL_N2HP:
            rc = lib.nf_n2hp_loglike_host(self.handle, _lib.ptr(params), dt, _lib.ptr(pv), vecs_per_pix, B,
                                          ncomp, _lib.ptr(out))
        else:
            rc = lib.nf_gauss_loglike_host(self.handle, _lib.ptr(params), dt, _lib.ptr(pv), vecs_per_pix, B,
                                           ncomp, _lib.ptr(out))
        _lib.check(rc, "loglike")
        return out

    def predict(self, params, ncomp, cold=False, lte=False):
        """Model spectra for params [B, n_model*ncomp] -> float32 [B, n_spec, n_chan]."""
        lib = _lib.load()
        params, dt = self._params(params)
        if params.ndim != 2 or params.shape[1] != self.n_model * ncomp:
            raise ValueError(f'Invalid shape for ncomp={ncomp}: {params.shape}')
        B = params.shape[0]
        out = np.empty((B, self.n_spec, self.n_chan), dtype=np.float32)
        if self.model == _lib.NF_MODEL_NH3:
            flags = (_lib.NF_FLAG_COLD if cold else 0) | (_lib.NF_FLAG_LTE if lte else 0)
            rc = lib.nf_nh3_predict_host(self.handle, _lib.ptr(params), dt, B, ncomp, flags, _lib.ptr(out))
        elif self.model == _lib.NF_MODEL_N2HP:
            rc = lib.nf_n2hp_predict_host(self.handle, _lib.ptr(params), dt, B, ncomp, _lib.ptr(out))
        else:
            rc = lib.nf_gauss_predict_host(self.handle, _lib.ptr(params), dt, B, ncomp, _lib.ptr(out))
        _lib.check(rc, "predict")
        return out

    def last_call_stats(self):
        ms = C.c_double()
        n = C.c_int64()
        _lib.load().nf_last_call_stats(C.byref(ms), C.byref(n))
        return ms.value, n.value
