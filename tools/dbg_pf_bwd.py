"""debug: post-fusion backward, kernel path vs PyTorch branch, loss terms separated"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import speech2lip_b200 as S
dev = torch.device("cuda:0")
cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "may_cfg.json")))
m = S.TalkingFace(device=dev, cfg=cfg, mode="eval").to(dev).eval()
for expand in (True, False):
    m.expand_lip_mask = expand
    g = torch.Generator().manual_seed(77 + int(expand))
    B, h, w, lh, lw, x0, y0 = 2, 64, 72, 12, 20, 20, 30
    lip, face, gt = torch.rand(B, lh, lw, 3, generator=g), torch.rand(B, h, w, 3, generator=g), torch.rand(B, h, w, 3, generator=g)
    mask = torch.zeros(B, h, w, 3)
    mask[:, y0 + 1:y0 + lh - 1, x0 + 1:x0 + lw - 1] = 1
    mask[:, y0 + 3, x0 + 4, 2] = 0.25
    ys, xs = torch.meshgrid(torch.linspace(-1, 1, h), torch.linspace(-1, 1, w), indexing="ij")
    coord = torch.stack([xs, ys], -1)[None].repeat(B, 1, 1, 1) * 1.05 + 0.02 * torch.randn(B, h, w, 2, generator=g)
    r1, r2 = torch.randn(B, h, w, 3, generator=g).to(dev), torch.randn(B, h, w, 3, generator=g).to(dev)
    d = lambda t: t.to(dev)
    for which in ("recon", "fused", "canon"):
        def run(force):
            lp = d(lip).requires_grad_(True)
            cd = d(coord).requires_grad_(force)
            recon, fused, canon = m.post_fusion2_onlylip(lp, d(face), d(gt), d(mask), x0, y0, cd, use_canonical_space=True)
            {"recon": (recon * r1).sum(), "fused": (fused * r1).sum(), "canon": (canon * r2).sum()}[which].backward()
            return lp.grad.clone()
        a, b = run(False), run(True)
        c = run(True)
        print(expand, which, "scale %.3e kernel-vs-torch %.3e torch-vs-torch %.3e" % (b.abs().max().item(), (a - b).abs().max().item(), (b - c).abs().max().item()))
