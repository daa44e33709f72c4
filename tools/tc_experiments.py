"""Timing of the plain 256x256 render (64 frames, 4.19 M points, one MLP launch) for debug builds of the library.
usage: tc_experiments.py [lib.so ...]   (each lib is timed in a subprocess via S2L_LIB_PATH)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] != "--child":
    for lib in sys.argv[1:]:
        env = dict(os.environ, S2L_LIB_PATH=os.path.abspath(lib))
        print("==", lib, flush=True)
        subprocess.run([sys.executable, __file__, "--child"], env=env)
    sys.exit(0)
sys.path.insert(0, ROOT)
import torch
import speech2lip_b200 as s2l
from speech2lip_b200 import synth
dev = torch.device("cuda:0")
sd = {k: torch.from_numpy(v).to(dev) for k, v in synth.make_state_dict(0, "kaiming").items()}
w = s2l.PackedWeights(sd)
F = 64
audio = torch.from_numpy(synth.make_audio(F, seed=1)).to(dev)
idx = torch.arange(F)
for prec in ("bf16x3", "fp16f8", "bf16x1"):
    r = s2l.LipRenderer(w, prec)
    for _ in range(3):
        r.render_frames(audio, idx, 256, 256)
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r.render_frames(audio, idx, 256, 256); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print("  %-7s  best %.3f ms  median %.3f ms  (%d frames, %.1f Mpts)" % (prec, ts[0], ts[2], F, F * 65536 / 1e6), flush=True)
