"""Prior sets shared by the fixture generator (against the compiled reference) and the tests (against the
implementation under test).  No reference import here."""
import numpy as np


def kind_prior_sets(mod, size=500):
    """Prior sets exercising OrderedPrior, SpacedPrior and CenSepPrior (core.pyx:241-318), which neither
    get_irdc_priors nor get_synth_priors use.  `mod` is the reference's `core` module or nestfit_b200.core:
    the same definitions build the fixture and, in the tests, the implementation under test."""
    import scipy.stats as st
    u = np.linspace(0, 1, size)
    d_voff = mod.Distribution(8.0 * u - 4.0, st.beta(5.0, 5.0).pdf(u))
    d_vsep = mod.Distribution(2.57 * u + 0.13, st.beta(1.5, 3.0).pdf(u))
    d_dv = mod.Distribution(3.0 * u + 0.1, np.ones_like(u) / size)
    d_trot = mod.Distribution(23.0 * u + 7.0, st.beta(3.0, 6.7).pdf(u))
    d_tex = mod.Distribution(9.26 * u + 2.8, st.beta(1.0, 2.5).pdf(u))
    d_ntot = mod.Distribution(4.0 * u + 12.5, st.beta(10.0, 8.5).pdf(u))
    d_sigm = mod.Distribution(2.0 * u + 0.067, st.beta(1.5, 5.0).pdf(u))
    rest = [mod.Prior(d_trot, 1), mod.Prior(d_tex, 2), mod.Prior(d_ntot, 3), mod.Prior(d_sigm, 4),
            mod.ConstantPrior(0.25, 5)]
    return {
        "ordered": mod.PriorTransformer(np.array([mod.OrderedPrior(d_voff, 0)] + rest, dtype=object)),
        "spaced": mod.PriorTransformer(np.array([mod.SpacedPrior(mod.Prior(d_voff, 0), mod.Prior(d_dv, 0))] + rest,
                                                dtype=object)),
        "censep": mod.PriorTransformer(np.array([mod.CenSepPrior(mod.Prior(d_voff, 0), mod.Prior(d_vsep, 0))] + rest,
                                                dtype=object)),
    }
