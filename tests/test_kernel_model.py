"""The FP32 record algebra of the fused likelihood kernel, modelled in numpy (tools/kernel_model.py), against the
oracle on the CPU: the symmetric-window line records, the folded weights and the Taylor / exp2 radiative-transfer
step stay well inside the 1e-5-of-peak bound, MUFU.EX2's 2^-22 error included.  (The kernel itself is compared with
the oracle on the GPU, tests/test_gpu_parity.py; this pins the design's error budget without one.)"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tools"))


def test_fp32_record_algebra_error_budget(nb):
    import kernel_model as km
    assert km.error_stats(n_vec=20, ncomp=3, n_chan=1000, dv=0.07, seed=1) < 5e-6
    # a coarse axis (windows of a few channels) and a single component
    assert km.error_stats(n_vec=12, ncomp=1, n_chan=380, dv=0.158, seed=2) < 5e-6


def test_model_windows_match_oracle_counts(nb):
    """The model uses the reference's window rule: a line entirely outside the band contributes nothing and the
    spectrum is identically zero where no window reaches."""
    import kernel_model as km
    from oracle import oracle as orc
    xs = [orc.bench_axis(1, 400, 0.158), orc.bench_axis(2, 400, 0.158)]
    far = np.array([500.0, 12.0, 5.0, 14.5, 0.4, 0.0])
    assert not km.predict(xs, [1, 2], far, 1).any()
    p = np.array([0.0, 12.0, 5.0, 14.5, 0.2, 0.0])
    got = km.predict(xs, [1, 2], p, 1)
    want = orc.nh3_batch(xs, [1, 2], p[None], 1, want_pred=True)["pred"][0]
    assert np.array_equal(got == 0, want == 0)
