"""CPU tests of the drop-in boundary: the C-ABI library builds for sm_100a,
loads, and exports every function include/*.h declares (no compute calls)."""
import ctypes

import numpy as np
import pytest


def test_library_exports_every_declared_symbol(nb):
    from nestfit_b200 import _lib
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    names = _lib.declared_symbols()
    assert "nf_nh3_loglike" in names and "nf_prior_transform" in names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"


def test_abi_version_and_error_strings(nb):
    from nestfit_b200 import _lib
    lib = _lib.load()
    assert lib.nf_abi_version() == 2
    assert lib.nf_error_string(0) == b"ok"
    assert lib.nf_error_string(-1) == b"invalid argument"


def test_struct_layouts_match_header(nb):
    from nestfit_b200 import _lib
    assert ctypes.sizeof(_lib.DistDesc) == 48
    assert ctypes.sizeof(_lib.PriorDesc) == 40


def test_prior_plan_packing(nb):
    ut = nb.get_irdc_priors()
    pp, n_p, dd, n_d, tables = ut.pack()
    assert n_p == 6 and n_d == 5          # nested sigma prior + 5 top-level records
    kinds = [pp[i].kind for i in range(n_p)]
    assert kinds == [0, 7, 0, 0, 0, 1]
    assert pp[0].flags == 1 and pp[1].nested == 0 and pp[1].p_ix == 0 and pp[1].p_ix2 == 4
    assert abs(pp[1].value - 2.3548200450309493 * 1.2) < 1e-15
    assert tables.size == 5 * 7 * 501
    assert ut.n_param == 6


def test_shape_errors_mirror_reference(nb):
    ut = nb.get_irdc_priors()
    with pytest.raises(ValueError, match="Invalid shape for ncomp=2"):
        ut.transform(np.zeros(6), 2)
    with pytest.raises(AssertionError):
        nb.Spectrum(np.arange(10.0), np.zeros(10), -1.0)
    with pytest.raises(ValueError, match="not uniform"):
        nb.Spectrum(np.array([0, 1, 2, 4.0]), np.zeros(4), 1.0)


def test_missing_library_is_an_error_not_a_fallback(nb, monkeypatch, tmp_path):
    """The product path has no CPU fallback: without the built CUDA library every entry raises."""
    from nestfit_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "libnestfit_b200.so")
    with pytest.raises(_lib.NfError, match="no CPU fallback"):
        _lib.load()
    x = np.linspace(23.69e9, 23.70e9, 64)
    with pytest.raises(_lib.NfError):
        nb.PixelBlock("ammonia", [x, x + 2.8e7], np.zeros((1, 2, 64), np.float32), 0.1, trans_ids=[1, 2])


def test_no_device_is_an_error_not_a_fallback(nb):
    """On a host without a GPU the library loads (symbols, ABI version) but creating a pixel block fails with
    the device error: nothing is computed on the CPU."""
    from nestfit_b200 import _lib
    import ctypes as C
    lib = _lib.load()
    n = C.c_int(-1)
    rc = lib.nf_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    x = np.linspace(23.69e9, 23.70e9, 64)
    with pytest.raises(_lib.NfError):
        nb.PixelBlock("ammonia", [x, x + 2.8e7], np.zeros((1, 2, 64), np.float32), 0.1, trans_ids=[1, 2])
