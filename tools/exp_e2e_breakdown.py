"""Where the end-to-end step's extra time over the device-resident step goes (bench.py's e2e leg, headline geometry)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import speech2lip_b200 as s2l
from speech2lip_b200 import renderer as R, synth
dev = torch.device("cuda:0")
F, H, W, S = 8, 256, 256, 64
sd = {k: torch.from_numpy(v).to(dev) for k, v in synth.make_state_dict(0, "kaiming", 3, 4).items()}
w = s2l.PackedWeights(sd, 3, 4)
rend = s2l.LipRenderer(w, "fp16f8")
audio_h = torch.from_numpy(synth.make_audio(F, seed=100)).pin_memory()
index_h = torch.arange(F, dtype=torch.int64).pin_memory()
c2w_h = torch.eye(4)[:3].contiguous().pin_memory()
out_h = torch.empty(F, H, W, 3, pin_memory=True)
rgb_d = torch.empty(F, H, W, 3, device=dev)
z_d = torch.linspace(0., 1., S, device=dev)


def ev():
    return torch.cuda.Event(enable_timing=True)


def timed(fn, n=8):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    wall = 0.0
    for _ in range(n):
        e0, e1 = ev(), ev()
        t0 = time.perf_counter()
        e0.record(); fn(); e1.record(); e1.synchronize()
        wall += time.perf_counter() - t0
        tot += e0.elapsed_time(e1)
    return tot / n, wall / n * 1e3


def full():
    ad, idd, cd = audio_h.to(dev, non_blocking=True), index_h.to(dev, non_blocking=True), c2w_h.to(dev, non_blocking=True)
    ro, rd = R.get_rays(H, W, 1200.0, cd)
    rgb = rend.render_frames(ad, idd, H, W, mode="volumetric", rays_o=ro, rays_d=rd, z_vals=z_d, out=rgb_d)
    out_h.copy_(rgb, non_blocking=True)


ad, idd, cd = audio_h.to(dev), index_h.to(dev), c2w_h.to(dev)
ro, rd = R.get_rays(H, W, 1200.0, cd)


def render_only():
    rend.render_frames(ad, idd, H, W, mode="volumetric", rays_o=ro, rays_d=rd, z_vals=z_d, out=rgb_d)


def no_d2h():
    a2, i2, c2 = audio_h.to(dev, non_blocking=True), index_h.to(dev, non_blocking=True), c2w_h.to(dev, non_blocking=True)
    r2, d2 = R.get_rays(H, W, 1200.0, c2)
    rend.render_frames(a2, i2, H, W, mode="volumetric", rays_o=r2, rays_d=d2, z_vals=z_d, out=rgb_d)


for name, fn in (("render_frames only (device inputs)", render_only), ("H2D + rays + render", no_d2h), ("full e2e step", full),
                 ("D2H of the frames alone", lambda: out_h.copy_(rgb_d, non_blocking=True))):
    d, wl = timed(fn)
    print("%-40s device %.3f ms   wall %.3f ms" % (name, d, wl))
