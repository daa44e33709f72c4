"""Debug tool (S2L_TIMELINE build only): dumps a cycle-stamped event log of CTA 0's third tile."""
import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import speech2lip_b200 as s2l
from speech2lip_b200 import _cabi
from speech2lip_b200 import synth
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
dev = torch.device("cuda:0")
sd = {k: torch.from_numpy(v).to(dev) for k, v in synth.make_state_dict(0, "kaiming").items()}
w = s2l.PackedWeights(sd)
r = s2l.LipRenderer(w, prec)
F = int(os.environ.get('S2L_TL_FRAMES', '64'))
audio = torch.from_numpy(synth.make_audio(F, seed=1)).to(dev)
idx = torch.arange(F)
lib = _cabi.lib()
lib.s2l_debug_set_timeline.argtypes = [C.c_void_p]
for _ in range(6):
    r.render_frames(audio, idx, 256, 256)
buf = torch.zeros(4 * 8192, dtype=torch.int64, device=dev)
lib.s2l_debug_set_timeline(C.c_void_p(buf.data_ptr()))
r.render_frames(audio, idx, 256, 256)
torch.cuda.synchronize()
lib.s2l_debug_set_timeline(None)
b = buf.cpu().tolist()
c0, g0, c1, g1 = b[3 * 8192: 3 * 8192 + 4]
if g1 > g0:
    print('# kernel: %d SM cycles in %.3f ms -> %.3f GHz effective SM clock' % (c1 - c0, (g1 - g0) / 1e6, (c1 - c0) / (g1 - g0)))
ev = []
for role in range(3):
    n = b[role * 8192]
    for i in range(n):
        ev.append((b[role * 8192 + 2 + 2 * i], role, b[role * 8192 + 1 + 2 * i]))
stamps = [e for e in ev if e[2] not in (2000, 3000)]
t0 = min(e[0] for e in stamps)
# log order is kept per role (codes 2000 / 3000 carry durations, not timestamps: waits of the half just logged)
for t, role, code in ev:
    if code in (2000, 3000):
        print("%8d  %s  %d  dur" % (t, ["MMA ", "EPI8", "EPI12"][role], code))
    else:
        print("%8d  %s  %d" % (t - t0, ["MMA ", "EPI8", "EPI12"][role], code))
