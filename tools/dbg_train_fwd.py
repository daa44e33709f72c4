"""debug helper: one training forward (+ optional backward) at a given size, synchronised after every stage"""
import sys, torch
sys.path.insert(0, ".")
import speech2lip_b200 as s2l
from speech2lip_b200 import synth, renderer as R
from speech2lip_b200.autograd import FusedLipRender, MLP_PARAM_NAMES
F, H, W = [int(x) for x in sys.argv[1:4]]
do_bwd = len(sys.argv) > 4
dev = torch.device("cuda:0")
sd = {k: torch.from_numpy(v).to(dev) for k, v in synth.make_state_dict(0, "kaiming", 2, 3).items()}
w = s2l.PackedWeights(sd, 2, 3)
audio = torch.from_numpy(synth.make_audio(F, seed=61)).to(dev)
idx = torch.arange(F)
latent, _ = R.audio_encode(w, audio, idx)
torch.cuda.synchronize(); print("setup ok", flush=True)
lat = latent.clone().requires_grad_(True)
params = [sd[n].clone().requires_grad_(True) for n in MLP_PARAM_NAMES]
rgb = FusedLipRender.apply(lat, idx, torch.full((F,), 0.001), H, W, w, *params)
torch.cuda.synchronize(); print("forward ok", float(rgb.abs().max()), flush=True)
if do_bwd:
    rgb.backward(torch.randn_like(rgb))
    torch.cuda.synchronize(); print("backward ok", float(params[3].grad.abs().max()), flush=True)
