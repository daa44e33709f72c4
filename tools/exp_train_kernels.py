"""Per-kernel device time of one config-5 training step (4 frames x 6 renders of 80x120, one launch sequence) from the torch
profiler; run with S2L_LIB_PATH=<debug build> to compare kernel variants.   python tools/exp_train_kernels.py [H W]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import speech2lip_b200 as s2l
from speech2lip_b200 import synth
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (80, 120)
cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "may_cfg.json")))
m = s2l.TalkingFace(device=dev, cfg=cfg).to(dev).train()
m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_state_dict(0, "kaiming", 2, 3).items()}, strict=False)
F = 24
audio = torch.from_numpy(synth.make_audio(F, seed=21)).to(dev)
index = torch.arange(F)
target = torch.rand(F, H, W, 3, generator=torch.Generator().manual_seed(5)).to(dev)


def step():
    for p in m.parameters():
        p.grad = None
    ((m.render_lip_train(audio, index, H, W) - target) ** 2).mean().backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        step()
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / 5.0) for e in prof.key_averages() if e.device_time_total > 0]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print("lib:", os.environ.get("S2L_LIB_PATH", "default"), " total device time per step %.1f us" % tot)
for k, t in rows[:8]:
    print("  %8.1f us  %s" % (t, k[:90]))
