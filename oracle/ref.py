"""
TEST INFRASTRUCTURE -- not part of the product path.

Loader for the compiled reference (``oracle/_ref``, built by
``oracle/build_ref.py`` from /root/reference).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference arm may
import this module.

The reference calls ``scipy.integrate.cumtrapz`` (core.pyx:34) which modern
scipy renamed; the alias is installed before any ``Distribution`` is built.
"""

import sys
from pathlib import Path

import numpy as np

_REF_DIR = Path(__file__).resolve().parent / "_ref"
_mods = None

CKMS = 299792.458
NU11 = 23.6944955e9      # ammonia.pyx:69
NU22 = 23.722633335e9    # ammonia.pyx:70


def available():
    return (_REF_DIR / "nestfit" / "core").is_dir() and any(
        (_REF_DIR / "nestfit" / "core").glob("core.*.so"))


def load():
    """Return a namespace with the reference's compiled modules."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise ImportError("oracle/_ref is not built (run oracle/build_ref.py)")
    import scipy.integrate as si
    if not hasattr(si, "cumtrapz"):
        si.cumtrapz = si.cumulative_trapezoid
    if str(_REF_DIR) not in sys.path:
        sys.path.insert(0, str(_REF_DIR))
    import importlib
    core = importlib.import_module("nestfit.core.core")
    hyperfine = importlib.import_module("nestfit.models.hyperfine")
    ammonia = importlib.import_module("nestfit.models.ammonia")
    gaussian = importlib.import_module("nestfit.models.gaussian")
    diazenylium = importlib.import_module("nestfit.models.diazenylium")

    class _NS:
        pass
    ns = _NS()
    ns.core, ns.hyperfine, ns.ammonia, ns.gaussian = core, hyperfine, ammonia, gaussian
    ns.diazenylium = diazenylium
    _mods = ns
    return ns


def make_irdc_priors(core, size=500, vsys=0.0):
    """The `get_irdc_priors` prior set (prior_constructors.py:20-76) expressed
    against whichever module supplies Distribution/Prior classes (the compiled
    reference core, or nestfit_b200.core)."""
    import scipy.stats as st
    u = np.linspace(0, 1, size)
    spec = [
        ("voff", 8.00, -4.00 + vsys, (5.0, 5.0)),
        ("trot", 23.00, 7.00, (3.0, 6.7)),
        ("tex", 9.26, 2.80, (1.0, 2.5)),
        ("ntot", 4.00, 12.50, (10.0, 8.5)),
        ("sigm", 2.00, 0.067, (1.5, 5.0)),
    ]
    d = {}
    for name, a, b, (p, q) in spec:
        d[name] = core.Distribution(a * u + b, st.beta(p, q).pdf(u))
    priors = np.array([
        core.ResolvedPlacementPrior(core.Prior(d["voff"], 0),
                                    core.Prior(d["sigm"], 4), scale=1.2),
        core.Prior(d["trot"], 1),
        core.Prior(d["tex"], 2),
        core.Prior(d["ntot"], 3),
        core.ConstantPrior(0, 5),
    ])
    return core.PriorTransformer(priors)


def bench_axis(nu0, nchan=1000, dv=0.07):
    """Config-2 axis: v_j=(j-(nchan-1)/2)*dv, ascending Hz."""
    v = (np.arange(nchan) - 0.5 * (nchan - 1)) * dv
    return np.sort(nu0 * (1.0 - v / CKMS))
