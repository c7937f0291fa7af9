"""
TEST INFRASTRUCTURE -- not part of the product path.

ctypes wrapper of the plain-C restatement (oracle/nf_oracle.c).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may import
this module; it is the checker, never the thing measured or shipped.
"""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "libnf_oracle.so"
_lib = None

CKMS = 299792.458
NU = [23.6944955e9, 23.722633335e9, 23.8701296e9, 24.1394169e9, 24.53299e9,
      25.05603e9, 25.71518e9, 26.51898e9, 27.477943e9]


def build(force=False):
    src = HERE / "nf_oracle.c"
    if force or not LIB.exists() or LIB.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["gcc", "-O2", "-fPIC", "-std=gnu11", "-fno-fast-math", "-shared", "-o", str(LIB),
                        str(src), "-lm"], check=True)
    return LIB


def load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(str(LIB))
        lib.nfo_fast_expn.restype = C.c_double
        lib.nfo_fast_expn.argtypes = [C.c_double]
        lib.nfo_iemtex_interp.restype = C.c_double
        lib.nfo_iemtex_interp.argtypes = [C.c_double]
        lib.nfo_swift_convert.restype = C.c_double
        lib.nfo_swift_convert.argtypes = [C.c_double]
        lib.nfo_partition_level.restype = C.c_double
        lib.nfo_partition_level.argtypes = [C.c_long, C.c_double]
        lib.nfo_partition_func.restype = C.c_double
        lib.nfo_partition_func.argtypes = [C.c_int, C.c_double]
        _lib = lib
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def nh3_batch(xarrs, trans_ids, params, ncomp, data=None, noise=None, pix_of_vec=None,
              cold=False, lte=False, want_pred=False, count=False):
    """xarrs [nspec][nchan]; params [B, 6*ncomp]; data [npix, nspec, nchan]; noise [npix, nspec].
    Returns dict(lnL=[B] or None, pred=[B,nspec,nchan] or None, counters=[n_gauss, n_rt])."""
    lib = load()
    x = np.ascontiguousarray(xarrs, dtype=np.float64)
    nspec, nchan = x.shape
    params = np.ascontiguousarray(params, dtype=np.float64)
    B = params.shape[0]
    assert params.shape[1] == 6 * ncomp
    tid = np.ascontiguousarray(trans_ids, dtype=np.int32)
    d = n = None
    lnL = None
    if data is not None:
        d = np.ascontiguousarray(data, dtype=np.float64)
        n = np.ascontiguousarray(np.broadcast_to(np.asarray(noise, dtype=np.float64), d.shape[:2]))
        lnL = np.empty(B, dtype=np.float64)
    pv = None if pix_of_vec is None else np.ascontiguousarray(pix_of_vec, dtype=np.int32)
    pred = np.empty((B, nspec, nchan), dtype=np.float64) if want_pred else None
    counters = np.zeros(2, dtype=np.int64) if count else None
    rc = lib.nfo_nh3_loglike_batch(C.c_long(nspec), C.c_long(nchan), _p(x), _p(tid), _p(d), _p(n), _p(params),
                                   _p(pv), C.c_long(B), C.c_long(ncomp), int(cold), int(lte), _p(lnL),
                                   _p(pred), _p(counters))
    assert rc == 0
    return dict(lnL=lnL, pred=pred, counters=counters)


N2HP_NU = [93173.7637e6, 186344.8420e6, 279511.8325e6]     # diazenylium.pyx:39-43


def n2hp_batch(xarrs, trans_ids, params, ncomp, data=None, noise=None, pix_of_vec=None, want_pred=False,
               count=False):
    """N2H+ (diazenylium) model: params [B, 4*ncomp] = (voff, tex, ltau, sigm), parameter-major."""
    lib = load()
    x = np.ascontiguousarray(xarrs, dtype=np.float64)
    nspec, nchan = x.shape
    params = np.ascontiguousarray(params, dtype=np.float64)
    B = params.shape[0]
    assert params.shape[1] == 4 * ncomp
    tid = np.ascontiguousarray(trans_ids, dtype=np.int32)
    d = n = None
    lnL = None
    if data is not None:
        d = np.ascontiguousarray(data, dtype=np.float64)
        n = np.ascontiguousarray(np.broadcast_to(np.asarray(noise, dtype=np.float64), d.shape[:2]))
        lnL = np.empty(B, dtype=np.float64)
    pv = None if pix_of_vec is None else np.ascontiguousarray(pix_of_vec, dtype=np.int32)
    pred = np.empty((B, nspec, nchan), dtype=np.float64) if want_pred else None
    counters = np.zeros(2, dtype=np.int64) if count else None
    rc = lib.nfo_n2hp_loglike_batch(C.c_long(nspec), C.c_long(nchan), _p(x), _p(tid), _p(d), _p(n), _p(params),
                                    _p(pv), C.c_long(B), C.c_long(ncomp), _p(lnL), _p(pred), _p(counters))
    assert rc == 0
    return dict(lnL=lnL, pred=pred, counters=counters)


def gauss_batch(xarr, rest_freq, params, ncomp, data=None, noise=None, pix_of_vec=None,
                want_pred=False, count=False):
    lib = load()
    x = np.ascontiguousarray(xarr, dtype=np.float64)
    nchan = x.shape[0]
    params = np.ascontiguousarray(params, dtype=np.float64)
    B = params.shape[0]
    assert params.shape[1] == 3 * ncomp
    d = n = None
    lnL = None
    if data is not None:
        d = np.ascontiguousarray(data, dtype=np.float64).reshape(-1, nchan)
        n = np.ascontiguousarray(np.broadcast_to(np.asarray(noise, dtype=np.float64), (d.shape[0],)))
        lnL = np.empty(B, dtype=np.float64)
    pv = None if pix_of_vec is None else np.ascontiguousarray(pix_of_vec, dtype=np.int32)
    pred = np.empty((B, nchan), dtype=np.float64) if want_pred else None
    counters = np.zeros(2, dtype=np.int64) if count else None
    rc = lib.nfo_gauss_loglike_batch(C.c_long(nchan), _p(x), C.c_double(rest_freq), _p(d), _p(n), _p(params),
                                     _p(pv), C.c_long(B), C.c_long(ncomp), _p(lnL), _p(pred), _p(counters))
    assert rc == 0
    return dict(lnL=lnL, pred=pred, counters=counters)


def prior_transform(packed, u, ncomp):
    """packed = PriorTransformer.pack() of nestfit_b200.core; u [B, ndim] float64 (copied)."""
    lib = load()
    pp, n_p, dd, n_d, tables = packed
    u = np.array(u, dtype=np.float64, order="C", copy=True)
    B, ndim = u.shape
    rc = lib.nfo_prior_transform(C.cast(pp, C.c_void_p), n_p, C.cast(dd, C.c_void_p), n_d, _p(tables), _p(u),
                                 C.c_long(B), C.c_long(ndim), C.c_long(ncomp))
    assert rc == 0
    return u


def bench_axis(trans_id, nchan=1000, dv=0.07):
    """Config-2 axis: v_j = (j - (nchan-1)/2) dv, x = sort(nu0 (1 - v/c))  (SURVEY.md 8d)."""
    v = (np.arange(nchan) - 0.5 * (nchan - 1)) * dv
    return np.sort(NU[trans_id - 1] * (1.0 - v / CKMS))
