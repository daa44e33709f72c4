"""Frames/s of the UNMODIFIED caller: the literal inference.py:144-159 call sequence (tile the audio window H*W times,
audio_merge_forward, cat with the uv grid, rgb_forward) through the TalkingFace drop-in, one frame per iteration, next to
the batched LipRenderer on the same frames.  usage: bench_dropin.py [size] [frames]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import speech2lip_b200 as s2l
from oracle import synth, s2l_oracle as O

size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n_frames = int(sys.argv[2]) if len(sys.argv) > 2 else 50
dev = torch.device("cuda:0")
cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "may_cfg.json")))
m = s2l.TalkingFace(device=dev, cfg=cfg, mode="eval").to(dev).eval()
m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_state_dict(0, "kaiming").items()}, strict=False)
H = W = size
windows = torch.from_numpy(synth.make_audio(n_frames, seed=3)).to(dev)
coords = torch.from_numpy(O.get_coords(W, H).numpy()).to(dev)


def frame(i):
    audio = windows[i:i + 1].tile(H * W, 1, 1)                                    # inference.py:144
    with torch.no_grad():
        ab = m.audio_merge_forward(audio)                                         # :150
        x = torch.cat([coords[:, None, :], ab[:, None, :]], -1)                   # :151
        out = m.rgb_forward(x.view(-1, m.audio_dims + 2), time_pts=torch.tensor([i], device=dev), rgb_pts=None)   # :158
    return out[:, :3]


for i in range(3):
    frame(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(n_frames):
    out = frame(i)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("drop-in TalkingFace loop (inference.py:144-159), %dx%d: %.1f frames/s (%.3f ms/frame)" % (H, W, n_frames / dt, dt / n_frames * 1e3))

r = m.renderer("bf16x3")
idx = torch.arange(n_frames)
ref = r.render_frames(windows, idx, H, W)
torch.cuda.synchronize()
t0 = time.perf_counter()
ref = r.render_frames(windows, idx, H, W)
torch.cuda.synchronize()
dt2 = time.perf_counter() - t0
print("batched LipRenderer (bf16x3), same frames: %.1f frames/s" % (n_frames / dt2))
last = frame(n_frames - 1).reshape(H, W, 3)
print("max-abs drop-in vs batched renderer on the last frame: %.2e" % (last - ref[-1]).abs().max().item())
