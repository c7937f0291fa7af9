"""The algorithmic-work constants bench.py uses for its roofline are re-counted here with the
oracle's FP64 window rule on the benchmark's own seeded vectors."""
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tools"))


def test_bench_work_constants(nb):
    import bench
    import count_work
    g, r = count_work.nh3_counts()
    assert g == pytest.approx(bench.WORK_N_GAUSS, rel=1e-6) and r == pytest.approx(bench.WORK_N_RT, rel=1e-6)
    assert count_work.gauss_count() == pytest.approx(bench.WORK_GAUSS_MODEL, rel=2e-3)
    # SURVEY.md 8d probe values for this configuration: n_g ~ 8772, n_rt ~ 2418 (different draw of the same prior)
    assert abs(g / 8772 - 1) < 0.05 and abs(r / 2418 - 1) < 0.05
