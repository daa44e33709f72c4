"""Soak test of a tensor-core schedule (S2L_TC_IMPL=1|2|3 in the environment): many launches of random, ragged geometries
(different tile counts, dead tail iterations, both precisions, plain / ensemble4) compared with the exact fp32 path, plus a
long run of chip-filling launches compared with the first one bit for bit.  usage: stress_schedules.py [n_random] [n_big] [weight_seed]"""
import os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import speech2lip_b200 as s2l
from speech2lip_b200 import synth

n_random = int(sys.argv[1]) if len(sys.argv) > 1 else 200
n_big = int(sys.argv[2]) if len(sys.argv) > 2 else 100
wseed = int(sys.argv[3]) if len(sys.argv) > 3 else 0      # seed of the synthetic (kaiming) weights
dev = torch.device("cuda:0")
sd = {k: torch.from_numpy(v).to(dev) for k, v in synth.make_state_dict(wseed, "kaiming").items()}
w = s2l.PackedWeights(sd)
exact = s2l.LipRenderer(w, "fp32")
rnd = random.Random(0)
torch.manual_seed(0)
worst = {"bf16x3": 0.0, "fp16f8": 0.0}
t0 = time.time()
for it in range(n_random):
    F, H, W = rnd.randint(1, 5), rnd.randint(8, 120), rnd.randint(8, 120)
    mode = rnd.choice(["plain", "plain", "ensemble4"])
    prec = rnd.choice(["bf16x3", "fp16f8"])
    audio = torch.from_numpy(synth.make_audio(F, seed=it)).to(dev)
    idx = torch.randint(0, 5000, (F,))
    kw = dict(mode=mode, eps_shift=0.0015) if mode == "ensemble4" else {}
    got = s2l.LipRenderer(w, prec).render_frames(audio, idx, H, W, **kw)
    want = exact.render_frames(audio, idx, H, W, **kw)
    err = (got - want).abs().max().item()
    worst[prec] = max(worst[prec], err)
    assert err < (3e-4 if prec == "bf16x3" else 1e-3), (it, F, H, W, mode, prec, err)
print("weights seed %d, random geometries: %d launches ok in %.1f s, worst max-abs vs fp32 path %s" % (wseed, n_random, time.time() - t0, worst))
F, H, W = 64, 256, 256
audio = torch.from_numpy(synth.make_audio(F, seed=1)).to(dev)
idx = torch.arange(F)
r = s2l.LipRenderer(w, "fp16f8")
first = r.render_frames(audio, idx, H, W).clone()
out = torch.empty_like(first)
for it in range(n_big):
    r.render_frames(audio, idx, H, W, out=out)
    if it % 10 == 9:
        assert torch.equal(out, first), "launch %d differs" % it
torch.cuda.synchronize()
print("chip-filling launches: %d x %d frames bit-identical (schedule %d)" % (n_big, F, s2l._cabi.lib().s2l_tc_schedule(F * H * W // 128)))
