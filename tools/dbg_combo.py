"""debug helper: one inference render in a given (mode, precision, H, W, frames) combination under S2L_TC_IMPL"""
import sys, torch
sys.path.insert(0, ".")
import speech2lip_b200 as s2l
from speech2lip_b200 import synth
mode, prec, H, W, F = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
dev = torch.device("cuda:0")
sd = {k: torch.from_numpy(v).to(dev) for k, v in synth.make_state_dict(0, "kaiming", 2, 3).items()}
w = s2l.PackedWeights(sd, 2, 3)
audio = torch.from_numpy(synth.make_audio(F, seed=61)).to(dev)
r = s2l.LipRenderer(w, prec)
out = r.render_frames(audio, torch.arange(F), H, W, mode=mode, eps_shift=0.001)
torch.cuda.synchronize()
ref = s2l.LipRenderer(w, "fp32").render_frames(audio, torch.arange(F), H, W, mode=mode, eps_shift=0.001)
print("ok", mode, prec, H, W, F, "max err vs fp32 %.3e" % (out - ref).abs().max().item(), flush=True)
