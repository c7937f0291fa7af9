"""Times one launch of the fused NH3 likelihood kernel on the bench workload (2^20 vectors, 4096 pixels, 2 x 1000
channels) and stores the lnL it returns: `python tools/ab_kernel.py <tag> [ncomp ...]`.  Run once per kernel variant
(NF_NH3_KERNEL selects it) and compare the stored arrays with `python tools/ab_kernel.py --diff <tagA> <tagB>`.
Development tool."""
import os
import sys
import numpy as np
sys.path.insert(0, '.')


def main():
    if sys.argv[1] == '--diff':
        a, b = sys.argv[2:4]
        for nc in (1, 2, 3, 4):
            fa, fb = f'gpurun_out/ab_{a}_{nc}.npy', f'gpurun_out/ab_{b}_{nc}.npy'
            if os.path.exists(fa) and os.path.exists(fb):
                x, y = np.load(fa), np.load(fb)
                d = np.abs(x - y)
                ok = np.isfinite(x) & np.isfinite(y)
                i = int(np.argmax(np.where(ok, d, 0)))
                print(f"ncomp {nc}: max |dlnL| {d[ok].max():.3e} at lnL {x[i]:.5g}; max rel {np.max(d[ok] / np.abs(x[ok])):.3e}; "
                      f"non-finite {int((~np.isfinite(x)).sum())} / {int((~np.isfinite(y)).sum())}")
        return
    import torch
    import bench
    import nestfit_b200 as nb
    from nestfit_b200 import _lib
    tag = sys.argv[1]
    ncomps = [int(v) for v in sys.argv[2:]] or [3]
    lib = _lib.load()
    xs, data, noise, P32 = bench.build_problem(nb, 0)
    blk = nb.PixelBlock("ammonia", xs, data, noise, trans_ids=[1, 2])
    ut = nb.get_irdc_priors()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    for nc in ncomps:
        if nc == bench.NCOMP:
            P = P32
        else:
            U = np.random.default_rng(77 + nc).uniform(size=(bench.B_TOTAL, 6 * nc))
            P = ut.transform_batch(U, nc)
            bad = ~np.isfinite(P).all(axis=1)
            P[bad] = P[np.flatnonzero(~bad)[:bad.sum()]]
            P = P.astype(np.float32)
        d_params = torch.from_numpy(P).cuda()
        d_lnl = torch.empty(bench.B_TOTAL, dtype=torch.float64, device="cuda")
        ms = []
        for k in range(6):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(lib.nf_nh3_loglike(blk.handle, d_params.data_ptr(), _lib.NF_F32, None, bench.VPP, bench.B_TOTAL, nc, 0,
                                          d_lnl.data_ptr(), stream), "nf_nh3_loglike")
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        out = d_lnl.cpu().numpy()
        np.save(f'gpurun_out/ab_{tag}_{nc}.npy', out)
        best = min(ms[2:])
        print(f"[{tag}] ncomp {nc}: {best:.3f} ms per {bench.B_TOTAL} evals = {bench.B_TOTAL / best * 1e3:.4g} evals/s "
              f"(runs {', '.join(f'{m:.2f}' for m in ms)}); non-finite lnL {int((~np.isfinite(out)).sum())}", flush=True)


if __name__ == '__main__':
    main()
