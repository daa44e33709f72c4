"""CPU emulation of the fp16f8 arithmetic (fp16 main product + e4m3/e5m2 correction products, exact accumulation) on the
reference MLP: sweeps the two power-of-two pre-scales (kScaleA, kScaleW of s2l_common.cuh) and splits the remaining error by
operand.  Result (kaiming weights, |out| <= 4.5): (8, 10) is the best pair (3.7e-4 max-abs, 9.2e-5 rms); the two e5m2 (2-bit
mantissa) operands contribute 5.7e-5 rms each, the e4m3 ones 2.6e-5 / 3.5e-5; with exact corrections 1e-6 remains.
Test/analysis tool only (imports oracle/).  usage: python tools/emulate_fp16f8.py"""
import sys, torch, numpy as np, itertools
sys.path.insert(0, '/root/repo')
from oracle import s2l_oracle as O, synth
torch.set_num_threads(16)
def e4m3(x): return x.clamp(-448, 448).to(torch.float8_e4m3fn).to(torch.float64)
def e5m2(x): return x.clamp(-57344, 57344).to(torch.float8_e5m2).to(torch.float64)
def f16(x): return x.to(torch.float16).to(torch.float64)
def lin(A, W, sA, sW):
    A32 = A.float(); W32 = W.float()
    Ah = A32.half().float(); Wh = W32.half().float()
    main = Ah.double() @ Wh.double().t()
    c1 = e4m3((A32 - Ah) * 2.0**sA) @ e5m2(Wh * 2.0**-sA).t()
    c2 = e5m2(Ah * 2.0**-sW) @ e4m3((W32 - Wh) * 2.0**sW).t()
    return (main + c1 + c2)
def forward(sd, x, t, sA, sW, exact=False):
    uv = x[:, :2]; a = x[:, 2:]
    e = O.uv_embed(uv); tt = O.time_embed(t)
    L = (lambda A, W: A.double() @ W.double().t()) if exact else (lambda A, W: lin(A, W, sA, sW))
    # per-frame terms are hoisted in fp32/fp64 in the real kernel: keep them exact
    net = L(e, sd["fc_uv.weight"]) + sd["fc_uv.bias"].double() + (a.double() @ sd["fc_audio.weight"].double().t() + sd["fc_audio.bias"].double()) + (tt.double() @ sd["fc_time.weight"].double().t() + sd["fc_time.bias"].double())
    h = net
    for i in range(8):
        h = torch.relu(L(h.float() if not exact else h, sd["pts_linears.%d.weight" % i]) + sd["pts_linears.%d.bias" % i].double())
        if i == 4:
            hs = L(e, sd["fc_uv_skip.weight"]) + sd["fc_uv_skip.bias"].double() + (a.double() @ sd["fc_audio_skip.weight"].double().t() + sd["fc_audio_skip.bias"].double()) + (tt.double() @ sd["fc_time_skip.weight"].double().t() + sd["fc_time_skip.bias"].double())
            h = torch.cat([hs, h], -1)
    return L(h.float() if not exact else h, sd["output_linear.weight"]) + sd["output_linear.bias"].double()
res = {}
for seed in (0, 1):
    sd = O.to_torch_sd(synth.make_state_dict(seed, "kaiming"))
    g = torch.Generator().manual_seed(seed)
    N = 6000
    audio = torch.from_numpy(synth.make_audio(1, seed=seed + 3))
    with torch.no_grad():
        lat = O.audio_merge_forward(sd, audio)
        x = torch.cat([torch.rand(N, 2, generator=g), lat.expand(N, -1)], -1)
        t = torch.tensor([1234 + seed])
        ref = forward(sd, x, t, 0, 0, exact=True)
        for sA, sW in itertools.product((6, 8, 10, 12), (8, 10, 12, 14)):
            out = forward(sd, x, t, sA, sW)
            res.setdefault((sA, sW), []).append((out - ref).abs().max().item())
print("absmax out", ref.abs().max().item())
for k, v in sorted(res.items(), key=lambda kv: max(kv[1])):
    print("sA=%2d sW=%2d  max-abs %s" % (k[0], k[1], ["%.2e" % e for e in v]))

# ---- which correction's quantisation dominates?  (sA, sW) = (8, 10)
def lin_var(A, W, q1, q2):
    A32 = A.float(); W32 = W.float(); Ah = A32.half().float(); Wh = W32.half().float()
    main = Ah.double() @ Wh.double().t()
    rA = (A32 - Ah); rW = (W32 - Wh)
    c1 = (e4m3(rA * 2.0**8) @ e5m2(Wh * 2.0**-8).t()) if q1 == "q" else ((e4m3(rA * 2.0**8) * 2.0**-8) @ Wh.double().t() if q1 == "qA" else (rA.double() @ (e5m2(Wh * 2.0**-8) * 2.0**8).t() if q1 == "qW" else rA.double() @ Wh.double().t()))
    c2 = (e5m2(Ah * 2.0**-10) @ e4m3(rW * 2.0**10).t()) if q2 == "q" else ((e5m2(Ah * 2.0**-10) * 2.0**10) @ rW.double().t() if q2 == "qA" else (Ah.double() @ (e4m3(rW * 2.0**10) * 2.0**-10).t() if q2 == "qW" else Ah.double() @ rW.double().t()))
    return main + c1 + c2
import functools
sd = O.to_torch_sd(synth.make_state_dict(0, "kaiming"))
g = torch.Generator().manual_seed(0)
N = 6000
audio = torch.from_numpy(synth.make_audio(1, seed=3))
with torch.no_grad():
    lat = O.audio_merge_forward(sd, audio)
    x = torch.cat([torch.rand(N, 2, generator=g), lat.expand(N, -1)], -1)
    t = torch.tensor([1234])
    ref = forward(sd, x, t, 0, 0, exact=True)
    for q1, q2 in (("q", "q"), ("x", "q"), ("q", "x"), ("x", "x"), ("qA", "x"), ("qW", "x"), ("x", "qA"), ("x", "qW")):
        lin_backup = lin
        globals()["lin"] = lambda A, W, sA, sW, q1=q1, q2=q2: lin_var(A, W, q1, q2)
        out = forward(sd, x, t, 8, 10)
        globals()["lin"] = lin_backup
        print("c1=%-2s c2=%-2s  max-abs %.2e  rms %.2e" % (q1, q2, (out - ref).abs().max().item(), (out - ref).pow(2).mean().sqrt().item()))
