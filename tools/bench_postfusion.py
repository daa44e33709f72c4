"""Times the post-fusion compose kernel at the reference's operating point (500x500 canonical face, 80x120 lip crop)
against the same ops in PyTorch eager on the same GPU (the oracle's restatement of tf_nerf.py:334-386 run on cuda)."""
import json
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import speech2lip_b200 as s2l          # noqa: E402
from oracle import s2l_oracle as O     # noqa: E402  (test infrastructure: eager-GPU comparison arm)

dev = torch.device("cuda:0")
B, h, w, lh, lw, x0, y0 = 8, 500, 500, 80, 120, 190, 300
g = torch.Generator(device="cpu").manual_seed(0)
lip = torch.rand(B, lh, lw, 3, generator=g).to(dev)
face = torch.rand(B, h, w, 3, generator=g).to(dev)
gt = torch.rand(B, h, w, 3, generator=g).to(dev)
mask = torch.zeros(B, h, w, 3, device=dev)
mask[:, y0 + 5:y0 + lh - 5, x0 + 5:x0 + lw - 5] = 1
ys, xs = torch.meshgrid(torch.linspace(-1, 1, h), torch.linspace(-1, 1, w), indexing="ij")
coord = (torch.stack([xs, ys], -1)[None].repeat(B, 1, 1, 1) * 1.02 + 0.01).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(n):
        flush.fill_(0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        tot += a.elapsed_time(b)
    return tot / n


k = lambda: s2l.post_fusion_compose(lip, face, gt, mask, coord, x0, y0, True, lw // 5, want_canonical=False)
e = lambda: O.post_fusion_compose(lip, face, gt, mask, x0, y0, coord)
fused, _ = k()
want, _ = e()
err = (fused.permute(0, 2, 3, 1) - want).abs().max().item()
tk, te = timeit(k), timeit(e)
alg_bytes = B * (h * w * (8 + 12 + 12) + h * w * 12 + h * w * 12 + lh * lw * 12)   # coord+gt+out, face, mask, lip (unique bytes)
peaks = {}
try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except OSError:
    pass
peak = peaks.get("hbm_gbs", 6650.0)
print(json.dumps({"kernel": "post_fusion_kernel", "frames": B, "ms_kernel": tk, "ms_torch_eager_gpu": te, "speedup_vs_eager": te / tk,
                  "max_abs_vs_eager": err, "algorithmic_bytes": alg_bytes, "achieved_gbs": alg_bytes / tk / 1e6,
                  "hbm_peak_gbs": peak, "frac_of_hbm_peak": alg_bytes / tk / 1e6 / peak}))
