"""Count the algorithmic work of bench.py's seeded workloads with the reference's FP64 window rule (C oracle,
test infrastructure): prints the constants WORK_N_GAUSS / WORK_N_RT / WORK_GAUSS_MODEL kept in bench.py.
Needs a GPU (the benchmark's parameter vectors come from the device prior transform)."""
import sys
import numpy as np
sys.path.insert(0, '.')
import bench
import nestfit_b200 as nb
from oracle import oracle as orc


def nh3_counts():
    xs, data, noise, P32 = bench.build_problem(nb, 0)
    pick = np.random.default_rng(7).choice(bench.B_TOTAL, size=8192, replace=False)
    cnt = orc.nh3_batch(xs, [1, 2], P32[pick].astype(np.float64), bench.NCOMP, count=True)["counters"] / 8192.0
    return float(cnt[0]), float(cnt[1])


def gauss_count():
    B, n_chan, ncomp = 1 << 20, 4096, 8
    rng = np.random.default_rng(5)
    v = (np.arange(n_chan) - 2047.5) * 0.05
    x = np.sort(bench.NU_NH3[0] * (1 - v / bench.CKMS))
    P = np.concatenate([np.sort(rng.uniform(-90, 90, (B, ncomp)), axis=1), rng.uniform(0.2, 3, (B, ncomp)),
                        rng.uniform(0.1, 5, (B, ncomp))], axis=1).astype(np.float32)
    return float(orc.gauss_batch(x, bench.NU_NH3[0], P[:2048].astype(np.float64), ncomp, count=True)["counters"][0] / 2048.0)


if __name__ == "__main__":
    g, r = nh3_counts()
    print(f"WORK_N_GAUSS = {g!r}\nWORK_N_RT = {r!r}\nWORK_GAUSS_MODEL = {gauss_count()!r}")
