"""The UNMODIFIED caller loop (inference.py:144-159 call sequence) through the TalkingFace drop-in: frames/s, and a check that
no torch-level host synchronisation happens inside the loop body (torch.cuda.set_sync_debug_mode)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import speech2lip_b200 as s2l
from speech2lip_b200 import synth
dev = torch.device("cuda:0")
H = W = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "may_cfg.json")))
tf = s2l.TalkingFace(device=dev, cfg=cfg, mode="eval").to(dev).eval()
tf.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_state_dict(0, "kaiming", 2, 3).items()}, strict=False)
wins = torch.from_numpy(synth.make_audio(64, seed=9)).to(dev)
vv, uu = torch.meshgrid(torch.linspace(0.0, 1.0, H, device=dev), torch.linspace(0.0, 1.0, W, device=dev), indexing="ij")
coords = torch.stack([uu, vv], -1).view(-1, 2)
idx = [torch.tensor([i], device=dev) for i in range(64)]

def frame(i):
    with torch.no_grad():
        au = wins[i:i + 1].tile(H * W, 1, 1)
        ab = tf.audio_merge_forward(au)
        xx = torch.cat([coords[:, None, :], ab[:, None, :]], -1).view(-1, tf.audio_dims + 2)
        return tf.rgb_forward(xx, time_pts=idx[i], rgb_pts=None)[:, :3]
for i in range(4):
    frame(i)
torch.cuda.synchronize()
torch.cuda.set_sync_debug_mode("error")
for i in range(4):
    frame(i)                      # raises if any torch op in the body synchronises with the host
torch.cuda.set_sync_debug_mode("default")
torch.cuda.synchronize()
for prec in ("bf16x3", "fp16f8"):
    tf.dropin_precision = prec
    t0 = time.perf_counter()
    for i in range(64):
        frame(i)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("drop-in loop %dx%d %s: %.0f frames/s (%.3f ms/frame), no host sync in the body" % (H, W, prec, 64 / dt, dt / 64 * 1e3))

if "--profile" in sys.argv:
    # where the per-frame time goes: host enqueue time (loop without the final synchronise) and per-kernel device time
    tf.dropin_precision = "bf16x3"
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(64):
        frame(i)
    t_enq = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print("host enqueue %.3f ms/frame, with device drain %.3f ms/frame" % (t_enq / 64 * 1e3, t_all / 64 * 1e3))
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for i in range(16):
            frame(i)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
