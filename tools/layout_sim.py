"""Work counts of alternative main-loop layouts of the fused likelihood kernel on the bench workload (planning tool).

For a sample of the seeded config-2 vectors (bench.py: irdc priors, 3 components, 2 x 1000 channels, dv = 0.07 km/s)
the reference window [lo, hi) of every hyperfine line is computed in FP64 and the number of executed MUFU lane
slots, pair trips and (chunk | group, component) records is counted for

  v8        the shipped layout: 64-channel chunks, two lines x two channels per trip, a run of lines per
            (chunk, component) record, one-channel trips when the whole run touches only one 32-channel half
  pairhalf  the same with the half decision taken per pair instead of per run
  group     group-centric: lines of a component clustered into hyperfine groups (gap > `gap` channels), lanes over
            the group's own window in 32-channel columns from its first channel, two lines x two columns per trip
            (one-column trips for an odd last column), one record per (group, component)
  group/pair  the same with every pair walking only the columns its own two windows reach

and turned into an instruction estimate with the per-trip / per-record costs measured for v8
(profiles/r01_source_hotspots.txt).  Usage: python tools/layout_sim.py [n_vectors]
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / 'tools'))

CKMS = 299792.458


def windows(xs, trans_ids, P, ncomp, tb):
    """lo, hi [n_vec, n_spec, ncomp, max_lines] (hi <= lo: no window), lines in ascending frequency."""
    n_vec = P.shape[0]
    nl = max(tb['off'][t] - tb['off'][t - 1] for t in trans_ids)
    lo = np.zeros((n_vec, len(xs), ncomp, nl), dtype=np.int64)
    hi = np.zeros_like(lo)
    voff = P[:, 0 * ncomp:1 * ncomp]
    sigm = P[:, 4 * ncomp:5 * ncomp]
    for s, (x, t) in enumerate(zip(xs, trans_ids)):
        n = x.shape[0]
        nu_min, inv_chan = x[0], 1.0 / (x[1] - x[0])
        a, b = tb['off'][t - 1], tb['off'][t]
        f = tb['nu'][t - 1] * (1.0 - tb['voff'][a:b] / CKMS)                       # [lines]
        w = sigm[:, :, None] / CKMS * f
        nucen = f - voff[:, :, None] / CKMS * f
        rel = nucen - nu_min
        l = np.floor((rel - 5 * np.abs(w)) * inv_chan).astype(np.int64)
        h = np.floor((rel + 5 * np.abs(w)) * inv_chan).astype(np.int64)
        out = (h < 0) | (l > n - 1)
        l, h = np.clip(l, 0, None), np.clip(h, None, n - 1)
        h = np.where(out, l, h)
        lo[:, s, :, :b - a], hi[:, s, :, :b - a] = l, h
    return lo, hi


def count_v8(lo, hi, per_pair_half=False):
    """slots, full trips, half trips, records, chunks finalised -- summed over the sample."""
    tot = dict(slots=0, full=0, half=0, records=0, chunks=0, rt_slots=0)
    n_vec, n_spec, ncomp, _ = lo.shape
    for v in range(n_vec):
        for s in range(n_spec):
            touched = set()
            for c in range(ncomp):
                l, h = lo[v, s, c], hi[v, s, c]
                on = h > l
                if not on.any():
                    continue
                l, h = l[on], h[on]
                for g in range(int(l.min()) >> 6, (int(h.max()) - 1 >> 6) + 1):
                    c0, c1 = g << 6, (g << 6) + 64
                    hit = (l < c1) & (h > c0)
                    if not hit.any():
                        continue
                    idx = np.flatnonzero(hit)
                    first, end = idx[0], idx[-1] + 1              # the run of sorted lines [first, end)
                    ll, hh = l[first:end], h[first:end]
                    inA = (ll < c0 + 32) & (hh > c0)
                    inB = (ll < c1) & (hh > c0 + 32)
                    cnt = end - first
                    npairs = (cnt + 1) // 2
                    tot['records'] += 1
                    tot['rt_slots'] += 64
                    touched.add(g)
                    if per_pair_half:
                        for q in range(npairs):
                            a_ = inA[2 * q:2 * q + 2].any()
                            b_ = inB[2 * q:2 * q + 2].any()
                            if a_ and b_:
                                tot['full'] += 1; tot['slots'] += 128
                            else:
                                tot['half'] += 1; tot['slots'] += 64
                    elif inA.any() and inB.any():
                        tot['full'] += npairs; tot['slots'] += 128 * npairs
                    else:
                        tot['half'] += npairs; tot['slots'] += 64 * npairs
            tot['chunks'] += len(touched)
    return tot


def count_group(lo, hi, gap=24, per_pair_cols=False):
    tot = dict(slots=0, full=0, half=0, records=0, chunks=0, rt_slots=0)
    n_vec, n_spec, ncomp, _ = lo.shape
    for v in range(n_vec):
        for s in range(n_spec):
            for c in range(ncomp):
                l, h = lo[v, s, c], hi[v, s, c]
                on = h > l
                if not on.any():
                    continue
                l, h = l[on], h[on]
                # lines are frequency sorted: a new group starts where a window begins more than `gap` channels
                # after every earlier window ended
                start = 0
                reach = h[0]
                bounds = []
                for i in range(1, l.size):
                    if l[i] > reach + gap:
                        bounds.append((start, i)); start = i; reach = h[i]
                    else:
                        reach = max(reach, h[i])
                bounds.append((start, l.size))
                for a, b in bounds:
                    glo, ghi = int(l[a:b].min()), int(h[a:b].max())
                    ncol = -(-(ghi - glo) // 32)
                    npairs = (b - a + 1) // 2
                    tot['records'] += 1
                    tot['chunks'] += ncol                          # columns accumulated into the model spectrum
                    tot['rt_slots'] += 32 * ncol
                    if per_pair_cols:          # every pair walks only the columns its own two windows reach
                        for q in range(npairs):
                            pl = int(l[a + 2 * q:a + 2 * q + 2].min())
                            ph = int(h[a + 2 * q:a + 2 * q + 2].max())
                            nc_ = ((ph - 1 - glo) >> 5) - ((pl - glo) >> 5) + 1
                            tot['full'] += nc_ // 2; tot['slots'] += 128 * (nc_ // 2)
                            if nc_ & 1:
                                tot['half'] += 1; tot['slots'] += 64
                        continue
                    tot['full'] += npairs * (ncol // 2); tot['slots'] += 128 * npairs * (ncol // 2)
                    if ncol & 1:
                        tot['half'] += npairs; tot['slots'] += 64 * npairs
    return tot


def estimate(t, n_vec, rec_cost=34, fin_cost=14, full_cost=24, half_cost=16, fixed=676 + 624 + 207 + 330):
    """Warp instructions per eval from the counts (fixed = T + L + S + the rest of v8, per eval)."""
    return (t['full'] * full_cost + t['half'] * half_cost + t['records'] * rec_cost + t['chunks'] * fin_cost) / n_vec + fixed


def main():
    import bench
    import nestfit_b200 as nb
    from oracle import oracle as orc
    import kernel_model as km
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    ut = nb.get_irdc_priors()
    U = np.random.default_rng(4321).uniform(size=(4 * n, 18))
    P = orc.prior_transform(ut.pack(), U, 3)
    P = P[np.isfinite(P).all(axis=1)][:n]
    xs = bench.axes()
    tb = km.load_tables()
    lo, hi = windows(xs, [1, 2], P, 3, tb)
    useful = float(np.clip(hi - lo, 0, None).sum()) / P.shape[0]
    print(f"{P.shape[0]} vectors; useful Gaussian slots per eval {useful:.0f}")
    rows = [("v8", count_v8(lo, hi)), ("pairhalf", count_v8(lo, hi, per_pair_half=True))]
    rows += [(f"group gap={g}", count_group(lo, hi, gap=g)) for g in (8, 24, 48)]
    rows += [(f"group/pair {g}", count_group(lo, hi, gap=g, per_pair_cols=True)) for g in (8, 24)]
    print(f"{'layout':14s} {'slots':>8s} {'lane eff':>8s} {'full':>7s} {'half':>7s} {'records':>8s} {'fin/col':>8s} {'instr est':>10s}")
    for name, t in rows:
        m = P.shape[0]
        est = estimate(t, m) if not name.startswith('group') else estimate(t, m, fin_cost=6, fixed=624 + 207 + 330 + 250)
        print(f"{name:14s} {t['slots'] / m:8.0f} {useful / (t['slots'] / m):8.3f} {t['full'] / m:7.1f} {t['half'] / m:7.1f} "
              f"{t['records'] / m:8.1f} {t['chunks'] / m:8.1f} {est:10.0f}")


if __name__ == '__main__':
    main()
