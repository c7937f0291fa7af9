"""Degenerate pixel/vector layouts of the likelihood microbench:
  1024 pixels x 1024 vectors   the bench.py layout (every CTA tile shares one pixel, staged by TMA)
  1 pixel x 2^20 vectors       every tile stages the same 8 KB pixel (L2 resident)
  2^18 pixels x 1 vector       no reuse: every vector reads its own pixel rows straight from HBM
                               (2^18 and not 2^20 pixels to keep the host-side synthetic cube at 2 GB)
Prints one JSON line.  Usage: python tools/bench_layouts.py [--steps 5]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    import torch
    import bench
    import nestfit_b200 as nb
    from nestfit_b200 import _lib
    lib = _lib.load()
    dev = 0
    torch.cuda.set_device(dev)
    xs, data, noise, P32 = bench.build_problem(bench._with_device(nb, dev), 0)
    d_params = torch.from_numpy(P32).to("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
    stream = torch.cuda.current_stream().cuda_stream
    rng = np.random.default_rng(99)
    out = {}

    def timed(blk, B, vpp, pix_of_vec=None):
        d_lnl = torch.empty(B, dtype=torch.float64, device="cuda:0")
        d_pix = None if pix_of_vec is None else torch.from_numpy(pix_of_vec).to("cuda:0")

        def launch():
            _lib.check(lib.nf_nh3_loglike(blk.handle, d_params.data_ptr(), _lib.NF_F32,
                                          None if d_pix is None else d_pix.data_ptr(), vpp, B, bench.NCOMP, 0,
                                          d_lnl.data_ptr(), stream), "nf_nh3_loglike")
        for _ in range(args.warmup):
            flush.zero_()
            launch()
        torch.cuda.synchronize()
        ms = 0.0
        for _ in range(args.steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            launch()
            e1.record()
            torch.cuda.synchronize()
            ms += e0.elapsed_time(e1)
        assert torch.isfinite(d_lnl).all()
        return B * args.steps / (ms * 1e-3), ms / args.steps

    blk = nb.PixelBlock("ammonia", xs, data, noise, trans_ids=[1, 2], device=dev)
    v, ms = timed(blk, bench.B_TOTAL, bench.VPP)
    out["1024x1024"] = {"evals_per_s": v, "ms": ms}
    blk.close()

    blk = nb.PixelBlock("ammonia", xs, data[:1], noise[:1], trans_ids=[1, 2], device=dev)
    v, ms = timed(blk, bench.B_TOTAL, bench.B_TOTAL)
    out["1x2^20"] = {"evals_per_s": v, "ms": ms}
    blk.close()

    n_pix = 1 << 18
    big = np.tile(data, (n_pix // data.shape[0], 1, 1))
    big += rng.normal(0.0, 0.01, size=(n_pix, 1, 1)).astype(np.float32)      # rows differ per pixel
    import time
    t0 = time.perf_counter()
    blk = nb.PixelBlock("ammonia", xs, big, np.full((n_pix, 2), bench.NOISE), trans_ids=[1, 2], device=dev)
    torch.cuda.synchronize()
    up_s = time.perf_counter() - t0
    # the same cube already resident in HBM (configs[3] size: 2^18 pixels x 2 x 1000 ch FP32 = 2.1 GB): pack into
    # padded rows + null evidence + per-chunk sums of squares = the whole ingest of a pixel block
    import ctypes as C
    d_big = torch.from_numpy(big).to("cuda:0")
    d_noise = torch.full((n_pix, 2), bench.NOISE, dtype=torch.float64, device="cuda:0")
    nu_min = np.array([x[0] for x in xs]); nu_chan = np.array([x[1] - x[0] for x in xs])
    tid = np.array([1, 2], dtype=np.int32)
    ing = []
    for _ in range(3):
        h = C.c_void_p()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _lib.check(lib.nf_pixels_create_from_device(0, _lib.NF_MODEL_NH3, n_pix, 2, bench.N_CHAN, _lib.ptr(nu_min),
                                                    _lib.ptr(nu_chan), _lib.ptr(tid), None,
                                                    C.c_void_p(d_big.data_ptr()), C.c_void_p(d_noise.data_ptr()),
                                                    C.byref(h)), "nf_pixels_create_from_device")
        torch.cuda.synchronize()
        ing.append(time.perf_counter() - t0)
        _lib.check(lib.nf_pixels_free(h), "nf_pixels_free")
    del d_big
    # bytes moved by the three ingest kernels: pack (read n + write n_pad), null evidence (read n_pad), chunk sums (read n_pad)
    moved = big.nbytes + 3 * (big.nbytes // bench.N_CHAN) * 1024
    out["cube_ingest_2.1GB"] = {"host_upload_s": up_s, "host_upload_GBps": big.nbytes / up_s / 1e9,
                                "device_ingest_s": min(ing), "device_ingest_GBps": moved / min(ing) / 1e9,
                                "bytes_moved": int(moved)}
    v, ms = timed(blk, n_pix, 1)
    out["2^18x1"] = {"evals_per_s": v, "ms": ms, "pixel_bytes": int(big.nbytes),
                     "hbm_GBps": big.nbytes / (ms * 1e-3) / 1e9}
    # the same pixels through an explicit (shuffled) vector -> pixel map
    v, ms = timed(blk, n_pix, 1, rng.permutation(n_pix).astype(np.int32))
    out["2^18x1_shuffled_map"] = {"evals_per_s": v, "ms": ms}
    blk.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
