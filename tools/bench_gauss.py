"""Config 5 (BASELINE.json configs[4]): Gaussian model, 8 components over 4096 channels; device-resident evals/s."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import torch
import nestfit_b200 as nb
from nestfit_b200 import _lib
from oracle import oracle as orc
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
rng = np.random.default_rng(5)
n_chan, ncomp = 4096, 8
v = (np.arange(n_chan) - 2047.5) * 0.05
x = np.sort(orc.NU[0] * (1 - v / orc.CKMS))
P = np.concatenate([np.sort(rng.uniform(-90, 90, (B, ncomp)), axis=1), rng.uniform(0.2, 3, (B, ncomp)),
                    rng.uniform(0.1, 5, (B, ncomp))], axis=1).astype(np.float32)
n_pix = 256
truth = P[:n_pix].astype(np.float64)
clean = orc.gauss_batch(x, orc.NU[0], truth, ncomp, want_pred=True)["pred"]
data = (clean + rng.normal(0, 0.1, clean.shape)).astype(np.float32)
blk = nb.PixelBlock("gaussian", [x], data[:, None, :], 0.1, rest_freq=orc.NU[0])
lib = _lib.load()
d_p = torch.from_numpy(P).cuda(); d_l = torch.empty(B, dtype=torch.float64, device='cuda')
st = torch.cuda.current_stream().cuda_stream
def launch():
    _lib.check(lib.nf_gauss_loglike(blk.handle, d_p.data_ptr(), _lib.NF_F32, None, B // n_pix, B, ncomp, d_l.data_ptr(), st), "gauss")
for _ in range(3): launch()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): launch()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
cnt = orc.gauss_batch(x, orc.NU[0], P[:2048].astype(np.float64), ncomp, count=True)["counters"] / 2048.0
got = d_l.cpu().numpy()[:512]
want = orc.gauss_batch(x, orc.NU[0], P[:512].astype(np.float64), ncomp, data=data.astype(np.float64), noise=0.1,
                       pix_of_vec=(np.arange(512) // (B // n_pix)).astype(np.int32))["lnL"]
print(f"gauss 8x4096: {B / ms * 1e3:.4g} evals/s, {ms:.3f} ms per {B} evals; n_gauss/eval {cnt[0]:.0f}; "
      f"{B * cnt[0] / ms * 1e-6:.1f} G exp/s; max |dlnL| {np.abs(got - want).max():.3g} (|lnL| ~ {np.abs(want).mean():.3g})")
