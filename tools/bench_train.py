"""Training-step timing of the MLP part of BASELINE.json config 5 (4 frames x 80x120 lip crop x 4 taps, forward +
backward through AudioNet -> rgb_forward -> 4-tap blend -> MSE) on one GPU:
  (a) speech2lip_b200.TalkingFace (fused fp32 forward + fused dgrad kernel + GEMM wgrads),
  (b) the same arithmetic in PyTorch eager on the same GPU (the oracle functions under autograd) — the reference's path."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import speech2lip_b200 as s2l                      # noqa: E402
from oracle import s2l_oracle as O, synth          # noqa: E402  (eager comparison arm)

dev = torch.device("cuda:0")
_te = O.time_embed
O.time_embed = lambda pos, out_dims=20, dtype=torch.float32: _te(pos.cpu(), out_dims, dtype).to(dev)   # eager arm runs on the GPU
cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "may_cfg.json")))
H, W, F = 80, 120, 4
sd_np = synth.make_state_dict(0, "kaiming")
m = s2l.TalkingFace(device=dev, cfg=cfg).to(dev).train()
m.load_state_dict({k: torch.from_numpy(v) for k, v in sd_np.items()}, strict=False)
sd = {k: torch.from_numpy(v).to(dev).requires_grad_(True) for k, v in sd_np.items()}
audio = torch.from_numpy(synth.make_audio(F, seed=3)).to(dev)
coords = O.get_coords(W, H).to(dev)
target = torch.rand(H * W, 3, device=dev)
rx, ry, eps = 0.5 / W, 0.5 / H, 0.001


def taps():
    out = []
    for vx in (-1, 1):
        for vy in (-1, 1):
            c = coords.clone()
            c[:, 0] += vx * rx + eps
            c[:, 1] += vy * ry + eps
            out.append(c.clamp_(0, 1))
    return out


TAPS = taps()


def step(fwd_audio, fwd_rgb, params):
    for p in params:
        p.grad = None
    loss = 0
    for f in range(F):
        lat = fwd_audio(audio[f:f + 1]).expand(H * W, -1)
        pred = sum(fwd_rgb(torch.cat([c, lat], -1), f) for c in TAPS) * 0.25
        loss = loss + ((pred - target) ** 2).mean()
    loss.backward()
    return loss


def timeit(fn, n=5):
    fn(); fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


mine = lambda: step(m.audio_merge_forward, lambda x, f: m.rgb_forward(x, time_pts=torch.tensor([f], device=dev)),
                    [p for n, p in m.named_parameters() if n in sd_np])
eager = lambda: step(lambda a: O.audio_merge_forward(sd, a), lambda x, f: O.rgb_forward(sd, x, torch.tensor([f])), list(sd.values()))
l1, l2 = mine().item(), eager().item()
g1 = m.pts_linears[3].weight.grad
g2 = sd["pts_linears.3.weight"].grad
rel = ((g1 - g2).abs().max() / g2.abs().max()).item()
t1, t2 = timeit(mine), timeit(eager)
pts = F * H * W * 4
print(json.dumps({"workload": "4 frames x 80x120 x 4 taps fwd+bwd (MLP part of config 5)", "point_evals": pts,
                  "ms_fused": t1, "ms_torch_eager_gpu": t2, "speedup": t2 / t1, "loss_fused": l1, "loss_eager": l2,
                  "rel_grad_err_pts3": rel, "tflops_fused_3x_fwd": pts * 1.224192e6 * 3 / (t1 * 1e-3) / 1e12}))
