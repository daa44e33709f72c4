"""Host-to-host frames/s of a T-frame live (1 eval / pixel) sequence at 256x256: per-chunk synchronous calls
(render_frames_host) vs the pipelined render_sequence_host with fp32 and uint8-BGR outputs.  usage: bench_sequence.py [T] [precision]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import speech2lip_b200 as s2l
from speech2lip_b200 import synth

T = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
prec = sys.argv[2] if len(sys.argv) > 2 else "fp16f8"
dev = torch.device("cuda:0")
H = W = 256
sd = {k: torch.from_numpy(v).to(dev) for k, v in synth.make_state_dict(0, "kaiming").items()}
r = s2l.LipRenderer(s2l.PackedWeights(sd), prec)
audio_h = torch.from_numpy(synth.make_audio(T, seed=5)).pin_memory()
index_h = torch.arange(T).pin_memory()
out32 = torch.empty(T, H, W, 3, pin_memory=True)
out8 = torch.empty(T, H, W, 3, dtype=torch.uint8, pin_memory=True)


def naive():
    for s0 in range(0, T, 64):
        r.render_frames_host(audio_h[s0:s0 + 64], index_h[s0:s0 + 64], H, W, out_host=out32[s0:s0 + 64])


def timeit(fn):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
    return T / (time.perf_counter() - t0)


print("%d frames %dx%d live, %s" % (T, H, W, prec))
print("  per-chunk synchronous calls, fp32 frames to host : %8.0f frames/s" % timeit(naive))
print("  pipelined, fp32 frames to host                   : %8.0f frames/s" % timeit(lambda: r.render_sequence_host(audio_h, index_h, H, W, 64, "rgb32", out32)))
print("  pipelined, uint8 BGR frames to host              : %8.0f frames/s" % timeit(lambda: r.render_sequence_host(audio_h, index_h, H, W, 64, "bgr8", out8)))
dd = r.render_frames(audio_h[:64].to(dev), index_h[:64].to(dev), H, W)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(T // 64):
    r.render_frames(audio_h[:64].to(dev), index_h[:64].to(dev), H, W, out=dd)
torch.cuda.synchronize()
print("  device-resident outputs (no D2H)                 : %8.0f frames/s" % (T / (time.perf_counter() - t0)))
