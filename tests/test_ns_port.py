"""CPU port of the batched nested-sampling scheme (oracle/ns_port.py, test infrastructure): analytic evidences.
The port is what the CUDA sampler's ln Z is compared with on the GPU (tests/test_gpu_sampler.py) and the CPU
baseline of the cube-fit metric in bench.py; MultiNest itself is absent (parity unpinned)."""
import math

import numpy as np
import pytest

from oracle import ns_port


@pytest.mark.parametrize("d,sig,rwalk", [(3, 0.05, False), (8, 0.03, False), (6, 0.04, True)])
def test_port_matches_analytic_gaussian_evidence(d, sig, rwalk):
    truth = 0.5 * d * math.log(2 * math.pi * sig * sig)      # Gaussian well inside the unit cube
    diffs = []
    for seed in range(3):
        r = ns_port.nested_sampling(lambda U: -((U - 0.5) ** 2).sum(axis=1) / (2 * sig * sig), d, 200, tol=0.5,
                                    seed=seed, rwalk=rwalk)
        assert r['n_samples'] == r['n_iter'] + 200
        assert -0.25 * d < r['max_loglike'] <= 0.0
        diffs.append((r['lnZ'] - truth) / r['lnZ_err'])
    assert np.max(np.abs(diffs)) < 4.0, diffs
    assert abs(np.mean(diffs)) < 2.5, diffs


def test_port_rejects_nan_scores():
    def score(U):
        out = -((U - 0.5) ** 2).sum(axis=1) / (2 * 0.1 ** 2)
        out[U[:, 0] < 0.2] = np.nan                           # a region the priors map to NaN
        return out
    r = ns_port.nested_sampling(score, 2, 100, tol=0.5, seed=4)
    truth = math.log(2 * math.pi * 0.01) + math.log(0.5 * (1 + math.erf(0.3 / 0.1 / math.sqrt(2))))
    assert abs(r['lnZ'] - truth) < 4 * r['lnZ_err'] + 0.05


def test_fit_pixel_escalation_small():
    """ncomp escalation of the port (main.py:450-469) on a small two-spectrum pixel: a strong single line is
    selected as one component, pure noise as none."""
    import nestfit_b200 as nb
    from oracle import oracle as orc
    ut = nb.get_irdc_priors()
    packed = ut.pack()
    xs = [orc.bench_axis(1, 160, 0.4), orc.bench_axis(2, 160, 0.4)]
    rng = np.random.default_rng(12)
    truth = np.array([[0.4, 14.0, 6.0, 14.6, 0.45, 0.0]])
    clean = orc.nh3_batch(xs, [1, 2], truth, 1, want_pred=True)["pred"][0]
    noise = np.array([0.1, 0.1])
    line = ns_port.fit_pixel(xs, [1, 2], clean + rng.normal(0, 0.1, clean.shape), noise, packed, ncomp_max=2,
                             nlive=60, nlive_snr_fact=1, seed=3)
    assert line["nbest"] == 1 and len(line["lnZ"]) == 3
    assert line["lnZ"][1] - line["lnZ"][0] > 11 and line["lnZ"][2] - line["lnZ"][1] < 11
    empty = ns_port.fit_pixel(xs, [1, 2], rng.normal(0, 0.1, clean.shape), noise, packed, ncomp_max=2,
                              nlive=60, nlive_snr_fact=1, seed=4)
    assert empty["nbest"] == 0 and len(empty["lnZ"]) == 2          # the escalation stops after the failed N = 1
