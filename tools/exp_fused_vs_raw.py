"""experiment: fused-compositing launch vs the same MLP launch writing raw outputs (+ composite kernel), benched geometry"""
import ctypes as C, sys, torch
sys.path.insert(0, ".")
import speech2lip_b200 as s2l
from speech2lip_b200 import _cabi, renderer as R, synth
dev = torch.device("cuda:0")
F, H, W, S = 8, 256, 256, 64
prec = sys.argv[1] if len(sys.argv) > 1 else "fp16f8"
sd = {k: torch.from_numpy(v).to(dev) for k, v in synth.make_state_dict(0, "kaiming", 3, 4).items()}
w = s2l.PackedWeights(sd, 3, 4)
audio = torch.from_numpy(synth.make_audio(F, seed=100)).to(dev)
idx = torch.arange(F, device=dev)
ro, rd = R.get_rays(H, W, 1200.0, torch.eye(4, device=dev)[:3])
z = torch.linspace(0, 1, S, device=dev)
lib = _cabi.lib()
_, bias = R.audio_encode(w, audio, idx, want_latent=False)
g = _cabi.S2LGeom(n_frames=F, height=H, width=W, n_samples=S, pts_mode=_cabi.PTS_RAYS, uv_dims=3, out_ch=4, z_per_ray=0, rays_per_frame_shared=1)
raw = torch.empty(F * H * W * S * 4, device=dev)
rgb = torch.empty(F, H, W, 3, device=dev)
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
p = R._ptr
def run_raw():
    _cabi.check(lib.s2l_mlp_fwd(p(w.blob), C.byref(g), p(bias), None, p(ro), p(rd), p(z), p(raw), _cabi.PRECISIONS[prec], st()), "mlp")
r = s2l.LipRenderer(w, prec)
def run_fused(**kw):
    r.render_frames(audio, idx, H, W, mode="volumetric", rays_o=ro, rays_d=rd, z_vals=z, out=rgb, **kw)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / n
for name, fn in (("raw outputs (MLP only)", run_raw), ("fused compositing + re-evaluation", run_fused), ("fused, re-evaluation off", lambda: run_fused(fix_thr=-1.0)),
                 ("raw outputs (MLP only)", run_raw), ("fused compositing + re-evaluation", run_fused)):
    print("%-40s %.2f ms / 8 frames" % (name, t(fn)), flush=True)
