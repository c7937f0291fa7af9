#!/usr/bin/env python3
"""
Generate the committed golden fixtures from the *compiled reference itself*
(oracle/_ref, built from /root/reference by oracle/build_ref.py).  Run in the
authoring container only:  python tests/golden/make_golden.py

Fixtures (float64, np.savez_compressed):
  nh3_golden.npz    parameter vectors (through the reference's own
                    get_irdc_priors-equivalent transform), reference model
                    spectra for (1,1) and (2,2) on the config-2 axes, lnL against
                    seeded synthetic data, cold / lte variants, scalar KATs
  gauss_golden.npz  8-component x 4096-channel Gaussian model spectra + lnL
  prior_golden.npz  unit-cube vectors and the reference's transforms for the
                    irdc and synth prior sets, ncomp 1..4
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import ref  # noqa: E402
sys.path.insert(0, str(Path(__file__).resolve().parent))
from prior_kind_sets import kind_prior_sets  # noqa: E402

OUT = Path(__file__).resolve().parent
m = ref.load()
amm, gau, core = m.ammonia, m.gaussian, m.core


def make_synth_priors(core, size=500):
    # prior_constructors.py:79-141 against the reference classes
    import scipy.stats as st
    u = np.linspace(0, 1, size)
    flat = np.ones_like(u) / size
    d_voff = core.Distribution(7.800 * u - 3.90, flat.copy())
    d_vsep = core.Distribution(2.570 * u + 0.13, flat.copy())
    d_tkin = core.Distribution(17.200 * u + 7.90, flat.copy())
    d_ntot = core.Distribution(1.600 * u + 12.95, flat.copy())
    d_sigm = core.Distribution(2.025 * u + 0.075, st.lognorm(1.0, scale=0.136).pdf(u))
    fwhm = 2 * np.sqrt(2 * np.log(2))
    return core.PriorTransformer(np.array([
        core.ResolvedCenSepPrior(core.Prior(d_voff, 0), core.Prior(d_vsep, 0), core.Prior(d_sigm, 4),
                                 scale=1 / fwhm),
        core.DuplicatePrior(d_tkin, 1, 2),
        core.Prior(d_ntot, 3),
        core.ConstantPrior(0, 5),
    ]))


def ref_transform(ut, U, ncomp):
    P = U.copy()
    for row in P:
        ut.transform(row, ncomp)
    return P


def main():
    rng = np.random.default_rng(20261018)
    ut = ref.make_irdc_priors(core)
    xs = [ref.bench_axis(ref.NU11), ref.bench_axis(ref.NU22)]
    out = {"x11": xs[0], "x22": xs[1]}
    # ---- NH3 spectra + lnL -------------------------------------------------
    NV = 10
    for ncomp in (1, 2, 3, 4):
        U = rng.uniform(size=(NV * 3, 6 * ncomp))
        P = ref_transform(ut, U, ncomp)
        P = P[np.isfinite(P).all(axis=1)][:NV]
        assert P.shape[0] == NV
        # synthetic data: first vector is the truth + N(0, 0.1^2)
        data = np.empty((2, 1000))
        specs = []
        for t in (0, 1):
            s0 = amm.AmmoniaSpectrum(xs[t], np.zeros(1000), 0.1, trans_id=t + 1)
            amm.amm_predict(s0, P[0].copy())
            data[t] = s0.get_spec() + rng.normal(0, 0.1, 1000)
            specs.append(amm.AmmoniaSpectrum(xs[t], data[t].copy(), 0.1, trans_id=t + 1))
        pred = np.empty((NV, 2, 1000))
        lnL = np.zeros(NV)
        for b in range(NV):
            for t in (0, 1):
                amm.amm_predict(specs[t], P[b].copy())
                pred[b, t] = specs[t].get_spec()
                lnL[b] += specs[t].loglikelihood
        out[f"params{ncomp}"] = P
        out[f"data{ncomp}"] = data
        out[f"pred{ncomp}"] = pred
        out[f"lnL{ncomp}"] = lnL
        runner = amm.AmmoniaRunner(np.array(specs), ut, ncomp=ncomp)
        out[f"null_lnZ{ncomp}"] = np.array(runner.null_lnZ)
        # Runner.loglikelihood: unit cube in (mutated to physical), lnL out (core.pyx:558-561)
        Ur = rng.uniform(size=(NV, 6 * ncomp))
        out[f"run_u{ncomp}"] = Ur.copy()
        out[f"run_lnL{ncomp}"] = np.array([runner.loglikelihood(row) for row in Ur])
        out[f"run_p{ncomp}"] = Ur
    # hand-picked vectors of SURVEY.md Appendix B (incl. optically thick) + flags
    hand = {
        "hand2": np.array([-1, 1.5, 10, 15, 4, 6, 14.5, 15, .3, .6, 0, 0], dtype=float),
        "hand3": np.array([-2, 0, 2, 10, 12, 15, 4, 5, 6, 14.5, 14.7, 15, .3, .4, .6, 0, 0, 0], dtype=float),
        "hand1": np.array([0, 25, 10, 16.4, 1.5, 0], dtype=float),
        "edge1": np.array([33.0, 12, 5, 14.8, 1.2, 0], dtype=float),      # window clipped at the band edge
        "narrow1": np.array([0.3, 9, 2.8, 13.0, 0.067, 0], dtype=float),  # Tex ~ Tcmb, narrowest line
        "ortho1": np.array([0.0, 20, 8.5, 14.5, 0.5, 0.5], dtype=float),  # ortho fraction 0.5
    }
    for name, p in hand.items():
        for cold, lte in ((0, 0), (1, 0), (0, 1)):
            pr = np.empty((2, 1000))
            for t in (0, 1):
                s0 = amm.AmmoniaSpectrum(xs[t], np.zeros(1000), 0.1, trans_id=t + 1)
                amm.amm_predict(s0, p.copy(), cold=bool(cold), lte=bool(lte))
                pr[t] = s0.get_spec()
            out[f"{name}_c{cold}l{lte}"] = pr
        out[f"{name}_p"] = p
    # (3,3) ortho transition on its own axis
    x33 = ref.bench_axis(23.8701296e9)
    s33 = amm.AmmoniaSpectrum(x33, np.zeros(1000), 0.1, trans_id=3)
    amm.amm_predict(s33, hand["ortho1"].copy())
    out["x33"] = x33
    out["ortho1_33"] = s33.get_spec()
    # scalar known-answer values (reference's own in-module tests + Appendix B)
    out["kat_partition"] = np.array([amm.partition_level(1, 10.0), amm.partition_func(True, 10.0),
                                     amm.partition_func(False, 10.0)])
    xi = np.linspace(0.05, 0.6, 23)
    out["kat_iemtex_x"] = xi
    out["kat_iemtex_y"] = np.array([m.hyperfine.iemtex_interp(v) for v in xi])
    np.savez_compressed(OUT / "nh3_golden.npz", **out)

    # ---- Gaussian model ------------------------------------------------------
    g = {}
    v = (np.arange(4096) - 2047.5) * 0.05
    xg = np.sort(ref.NU11 * (1 - v / ref.CKMS))
    g["x"] = xg
    g["rest_freq"] = np.array(ref.NU11)
    NG = 8
    Pg = np.empty((NG, 24))
    Pg[0] = np.concatenate([np.linspace(-70, 70, 8), np.linspace(.3, 2.4, 8), np.linspace(.5, 4, 8)])
    for b in range(1, NG):
        Pg[b] = np.concatenate([np.sort(rng.uniform(-90, 90, 8)), rng.uniform(0.2, 3, 8), rng.uniform(0.1, 5, 8)])
    Pg[NG - 1, 0] = -110.0   # window partly below channel 0
    s = core.Spectrum(xg, np.zeros(4096), 0.1, rest_freq=ref.NU11)
    gau.gauss_predict(s, Pg[0].copy())
    datag = s.get_spec() + rng.normal(0, 0.1, 4096)
    sd = core.Spectrum(xg, datag.copy(), 0.1, rest_freq=ref.NU11)
    predg = np.empty((NG, 4096))
    lnLg = np.empty(NG)
    for b in range(NG):
        gau.gauss_predict(sd, Pg[b].copy())
        predg[b] = sd.get_spec()
        lnLg[b] = sd.loglikelihood
    # NB: the reference's GaussianRunner cannot be constructed (gaussian.pyx:90 reads the
    # non-public cdef attribute `null_lnZ` through an untyped argument), so the null-model
    # value is taken from a fresh Spectrum whose pred is still zero (core.pyx:518-520).
    null_g = core.Spectrum(xg, datag.copy(), 0.1, rest_freq=ref.NU11).loglikelihood
    g.update(params=Pg, data=datag, pred=predg, lnL=lnLg, null_lnZ=np.array(null_g))
    np.savez_compressed(OUT / "gauss_golden.npz", **g)

    # ---- prior transforms ----------------------------------------------------
    pz = {}
    uts = {"irdc": ut, "synth": make_synth_priors(core)}
    for name, t in uts.items():
        for ncomp in (1, 2, 3, 4):
            if name == "synth" and ncomp > 2:
                continue  # CenSep priors are defined for n <= 2 (core.pyx:316-318)
            U = rng.uniform(size=(48, 6 * ncomp))
            U[0] = 0.5
            U[1] = 0.0
            pz[f"{name}_u{ncomp}"] = U
            pz[f"{name}_p{ncomp}"] = ref_transform(t, U, ncomp)
    np.savez_compressed(OUT / "prior_golden.npz", **pz)
    for f in ("nh3_golden.npz", "gauss_golden.npz", "prior_golden.npz"):
        print(f, (OUT / f).stat().st_size, "bytes")


def make_n2hp():
    """n2hp_golden.npz: N2H+ (diazenylium.pyx) spectra and lnL of the compiled reference for the
    three transitions, ncomp 1..3, incl. an optically thick and a band-edge vector."""
    dz = m.diazenylium
    rng = np.random.default_rng(20261019)
    out = {}
    nus = [93173.7637e6, 186344.8420e6, 279511.8325e6]
    nchan = 800
    for t in (1, 2, 3):
        x = ref.bench_axis(nus[t - 1], nchan=nchan, dv=0.06)
        out[f"x{t}"] = x
        for ncomp in (1, 2, 3):
            NV = 6
            P = np.concatenate([np.sort(rng.uniform(-6, 6, (NV, ncomp)), axis=1), rng.uniform(3, 25, (NV, ncomp)),
                                rng.uniform(-1.5, 1.3, (NV, ncomp)), rng.uniform(0.07, 1.6, (NV, ncomp))], axis=1)
            P[1, 2 * ncomp] = 1.8                # optically thick main component
            P[2, 0] = 21.0                       # windows clipped at the band edge
            s0 = dz.DiazenyliumSpectrum(x, np.zeros(nchan), 0.1, trans_id=t)
            dz.nnhp_predict(s0, P[0].copy())
            data = s0.get_spec() + rng.normal(0, 0.1, nchan)
            sd = dz.DiazenyliumSpectrum(x, data.copy(), 0.1, trans_id=t)
            pred = np.empty((NV, nchan))
            lnL = np.empty(NV)
            for b in range(NV):
                dz.nnhp_predict(sd, P[b].copy())
                pred[b] = sd.get_spec()
                lnL[b] = sd.loglikelihood
            out[f"params{t}_{ncomp}"] = P
            out[f"data{t}_{ncomp}"] = data
            out[f"pred{t}_{ncomp}"] = pred
            out[f"lnL{t}_{ncomp}"] = lnL
    # two transitions scored together through the reference's runner-style sum
    np.savez_compressed(OUT / "n2hp_golden.npz", **out)
    print("n2hp_golden.npz", (OUT / "n2hp_golden.npz").stat().st_size, "bytes")


def make_prior_kinds():
    """prior_kinds_golden.npz: the compiled reference's transforms through OrderedPrior, SpacedPrior and
    CenSepPrior for ncomp 1..4 (CenSepPrior leaves the unit-cube values in place for ncomp > 2, core.pyx:313-318)."""
    rng = np.random.default_rng(20261020)
    pz = {}
    for name, t in kind_prior_sets(core).items():
        for ncomp in (1, 2, 3, 4):
            U = rng.uniform(size=(64, 6 * ncomp))
            U[0] = 0.5
            U[1] = 0.0
            U[2] = 1.0 - 1e-12
            pz[f"{name}_u{ncomp}"] = U
            pz[f"{name}_p{ncomp}"] = ref_transform(t, U, ncomp)
    np.savez_compressed(OUT / "prior_kinds_golden.npz", **pz)
    print("prior_kinds_golden.npz", (OUT / "prior_kinds_golden.npz").stat().st_size, "bytes")


if __name__ == "__main__":
    if "--only-n2hp" in sys.argv:
        make_n2hp()
    elif "--only-prior-kinds" in sys.argv:
        make_prior_kinds()
    else:
        main()
        make_n2hp()
        make_prior_kinds()
