"""Concrete prior sets mirroring ``nestfit.prior_constructors``
(reference nestfit/prior_constructors.py:20-141)."""
import numpy as np

from .core import (ConstantPrior, Distribution, DuplicatePrior, Prior, PriorTransformer,
                   ResolvedCenSepPrior, ResolvedPlacementPrior)


def _beta_dist(u, scale, offset, a, b):
    from scipy import stats
    return Distribution(scale * u + offset, stats.beta(a, b).pdf(u))


def get_irdc_priors(size=500, vsys=0.0):
    """IRDC prior set (prior_constructors.py:20-76): beta-shaped priors on
    voff [-4,4]+vsys, trot [7,30], tex [2.8,12.06], ntot [12.5,16.5],
    sigm [0.067,2.067]; ortho fraction fixed to 0; centroids placed by
    ResolvedPlacementPrior with scale 1.2."""
    u = np.linspace(0, 1, size)
    d_voff = _beta_dist(u, 8.00, -4.00 + vsys, 5.0, 5.0)
    d_trot = _beta_dist(u, 23.00, 7.00, 3.0, 6.7)
    d_tex = _beta_dist(u, 9.26, 2.80, 1.0, 2.5)
    d_ntot = _beta_dist(u, 4.00, 12.50, 10.0, 8.5)
    d_sigm = _beta_dist(u, 2.00, 0.067, 1.5, 5.0)
    return PriorTransformer(np.array([
        ResolvedPlacementPrior(Prior(d_voff, 0), Prior(d_sigm, 4), scale=1.2),
        Prior(d_trot, 1),
        Prior(d_tex, 2),
        Prior(d_ntot, 3),
        ConstantPrior(0, 5),
    ], dtype=object))


def get_synth_priors(size=500):
    """Synthetic-test prior set (prior_constructors.py:79-141): uniform voff,
    vsep, tkin (duplicated into tex), ntot; log-normal sigm; centre-separation
    placement with scale 1/FWHM."""
    from scipy import stats
    u = np.linspace(0, 1, size)
    flat = np.ones_like(u) / size
    d_voff = Distribution(7.800 * u - 3.90, flat.copy())
    d_vsep = Distribution(2.570 * u + 0.13, flat.copy())
    d_tkin = Distribution(17.200 * u + 7.90, flat.copy())
    d_ntot = Distribution(1.600 * u + 12.95, flat.copy())
    d_sigm = Distribution(2.025 * u + 0.075, stats.lognorm(1.0, scale=0.136).pdf(u))
    fwhm = 2 * np.sqrt(2 * np.log(2))
    return PriorTransformer(np.array([
        ResolvedCenSepPrior(Prior(d_voff, 0), Prior(d_vsep, 0), Prior(d_sigm, 4), scale=1 / fwhm),
        DuplicatePrior(d_tkin, 1, 2),
        Prior(d_ntot, 3),
        ConstantPrior(0, 5),
    ], dtype=object))
