#!/usr/bin/env python
"""bench.py — lip frames/s of the Speech2Lip rendering hot path on B200.

Default workload (BASELINE.json configs[1], the configuration the metric is quoted on):
    volumetric mode, 256x256 rays, 64 samples/ray, batch = 8 frames per step, 1 GPU,
    parity arithmetic fp16f8 (tcgen05: exact fp16 x fp16 main product + two fp8 correction products, fp32
    accumulate; <= 5e-4 max-abs on O(8)-magnitude outputs, tested <= 1e-3 on every parity case).  The wider-margin
    parity mode bf16x3 (3 bf16 MMAs per product) and the live 1-eval / 4-tap modes are timed in the same run and
    reported under "extras".
A "step" = one pass of the whole path (AudioNet -> per-frame biases -> ray generation -> fused MLP ->
alpha compositing) over one batch of 8 synthetic frames.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference ...                     (CPU arm: the oracle port on the host cores)

Prints ONE JSON line (rank 0).  `value` = device-resident throughput, `e2e` = the same through the public
API with host buffers (H2D of the windows/poses and D2H of the frames inside the timed region),
`roofline` = achieved algorithmic FLOP/s of the fused-MLP kernel vs the measured bf16 tensor peak,
`cpu_baseline` = the oracle timed on this box's host cores on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

FLOP_PER_POINT = {"volumetric": 1_246_208, "plain": 1_224_192, "ensemble4": 1_224_192}   # SURVEY §8(d), algorithmic
ISSUED_MAC_PER_POINT = 495_616      # folded tensor-core layer program, per MMA pass
METRIC = "lip frames/sec @256x256,64 samples/ray"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="volumetric", choices=["volumetric", "plain", "ensemble4"])
    ap.add_argument("--precision", default="fp16f8", choices=["bf16x3", "bf16x1", "fp32", "fp16f8"])
    ap.add_argument("--no-extras", action="store_true", help="skip the extra (untimed-headline) measurements")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--samples", type=int, default=64)
    ap.add_argument("--frames", type=int, default=8, help="frames per step per GPU")
    ap.add_argument("--cpu-rays", type=int, default=0, help="bound the CPU steps to this many rays of a frame (0 = whole frames)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5],
                    help="BASELINE.json configs (1-based): 1 = single 128x128 crop x 32 samples; 2 = 256x256 x 64 samples, 8 frames/step "
                         "(the headline, default); 3 = 1000-frame 256x256x64 sequence sharded over the ranks (strong scaling); "
                         "4 = 512x512 x 128 samples; 5 = 4-frame training step with the sync-window renders, bf16")
    ap.add_argument("--seq-frames", type=int, default=1000, help="config 3: frames in the sequence")
    a = ap.parse_args()
    if a.config == 1:
        a.size, a.samples, a.frames, a.mode = 128, 32, 1, "volumetric"
    elif a.config == 3:
        a.size, a.samples, a.frames, a.mode = 256, 64, 8, "volumetric"
    elif a.config == 4:
        a.size, a.samples, a.frames, a.mode = 512, 128, 1, "volumetric"
    return a


def workload_name(a):
    if a.mode == "volumetric":
        return "volumetric %dx%d rays x %d samples/ray, %d frames/step/GPU (BASELINE configs[%d])" % (a.size, a.size, a.samples, a.frames, a.config - 1)
    return "%s %dx%d, %d frames/step/GPU" % (a.mode, a.size, a.size, a.frames)


def points_per_frame(a):
    n = a.size * a.size
    return n * (a.samples if a.mode == "volumetric" else 4 if a.mode == "ensemble4" else 1)


# --------------------------------------------------------------------------------------------- reference arm
def reference_runner(a, device, n_rays=None):
    """One step of the workload on the REFERENCE'S OWN CODE (oracle/ref_runner.py: the unmodified TalkingFace / get_rays /
    density2outputs / inference-loop body imported from oracle/_ref), or — only if that copy is absent — on the oracle port.
    Returns (run() -> rgb of ONE frame or of its first n_rays rays, fraction of a frame per run, description, kind)."""
    from oracle import ref_runner as RR
    from oracle import synth
    H = W = a.size
    audio = torch.from_numpy(synth.make_audio(1, seed=1))
    if not RR.available():
        return oracle_port_runner(a, n_rays) + ("port",)
    if a.mode == "volumetric":
        m = RR.model(3, 4, device)
        c2w = torch.eye(4)[:3]
        n = H * W if n_rays is None else min(n_rays, H * W)

        def run():
            return RR.render_volumetric(m, audio, 0, H, W, a.samples, 1200.0, c2w, device, n_rays=None if n == H * W else n)
        desc = "%d of %d rays x %d samples of one frame: TalkingFace(uv_dims=3,output_ch=4).audio_merge_forward once, rgb_forward in " \
               "65536-point calls, get_rays, density2outputs" % (n, H * W, a.samples)
        return run, n / float(H * W), desc, "reference"
    m = RR.model(2, 3, device)
    n_rows = H if n_rays is None else max(1, min(H, n_rays // W))
    if a.mode == "plain":
        def run():
            return RR.render_plain(m, audio, 0, H, W, device, n_rows=None if n_rows == H else n_rows)
        desc = "%d of %d pixel rows of one frame through the loop body of inference.py:144-159 as written (window tiled per pixel)" % (n_rows, H)
        return run, n_rows / float(H), desc, "reference"
    tr = RR.trainer(m, device, n_rows, W)

    def run():
        with torch.no_grad():
            return RR.render_ensemble4(tr, audio, 0, n_rows, W, device)
    return run, n_rows / float(H), "%d of %d pixel rows of one frame through Trainer.predict_lip_image (training.py:158-251)" % (n_rows, H), "reference"


def oracle_port_runner(a, n_rays):
    """Fallback when oracle/_ref is missing: the oracle port (restated reference arithmetic, torch CPU)."""
    from oracle import s2l_oracle as O
    from oracle import synth
    H = W = a.size
    audio = torch.from_numpy(synth.make_audio(1, seed=1))
    if a.mode == "volumetric":
        sd = O.to_torch_sd(synth.make_state_dict(0, "kaiming", 3, 4))
        ro, rd = O.get_rays(H, W, 1200.0, torch.eye(4)[:3])
        n = H * W if n_rays is None else min(n_rays, H * W)
        ro, rd = ro.reshape(-1, 3)[:n], rd.reshape(-1, 3)[:n]
        z = O.z_samples(a.samples)

        def run():
            with torch.no_grad():
                lat = O.audio_merge_forward(sd, audio)
                pts = (ro[:, None, :] + rd[:, None, :] * z[None, :, None]).reshape(-1, 3)
                outs = []
                for s in range(0, pts.shape[0], 65536):
                    p = pts[s:s + 65536]
                    outs.append(O.rgb_forward(sd, torch.cat([p, lat.expand(p.shape[0], -1)], -1), torch.tensor([0]), uv_dims=3))
                raw = torch.cat(outs).reshape(n, a.samples, 4)
                return O.density2outputs(raw, z.expand(n, a.samples), rd)[0]
        return run, n / float(H * W), "oracle port, %d of %d rays x %d samples of one frame" % (n, H * W, a.samples)
    sd = O.to_torch_sd(synth.make_state_dict(0, "kaiming", 2, 3))
    n_rows = H if n_rays is None else max(1, min(H, n_rays // W))

    def run():
        with torch.no_grad():
            if a.mode == "plain":
                return O.render_plain(sd, audio, 0, n_rows, W)
            return O.render_ensemble4(sd, audio, 0, n_rows, W, 0.0)
    return run, n_rows / float(H), "oracle port, %d of %d pixel rows of one frame, mode %s" % (n_rows, H, a.mode)


def eager_gpu_reference(a, dev):
    """the reference's own modules in PyTorch eager fp32 on THIS GPU (TF32 off) — the number a user of the
    reference sees on a B200.  One whole frame per repetition, wall clock around synchronised repetitions."""
    from oracle import ref_runner as RR
    import copy
    out = {}
    if not RR.available():
        return {"reference_gpu_eager": {"error": "oracle/_ref missing", "frames_per_s": 0.0, "ms_per_step": 0.0}}
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        for mode in ("volumetric", "plain", "ensemble4"):
            b = copy.copy(a)
            b.mode = mode
            run, frac, desc, kind = reference_runner(b, dev, None)
            run()
            torch.cuda.synchronize()
            reps = 3
            t0 = time.perf_counter()
            for _ in range(reps):
                run()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / reps
            out["reference_gpu_eager_%s_%dx%d" % (mode, a.size, a.size)] = {
                "frames_per_s": frac / dt, "ms_per_step": dt * 1e3, "frames_per_step": frac, "kind": kind,
                "what": "reference modules, PyTorch eager fp32 (TF32 off) on this GPU: " + desc}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    return out


def post_fusion_extras(dev, B=64, h=500, w=500, lh=80, lw=120, x0=190, y0=300):
    """the pre-UNet part of post_fusion2_onlylip_light (tf_nerf.py:334-386) as one gather-blend kernel at the
    reference's operating point (500x500 canonical face, 80x120 lip crop), against the REFERENCE'S OWN method on the same GPU
    (UNet replaced by identity in both arms so that only the compose + warp part is timed).  HBM roofline: unique bytes / time."""
    import speech2lip_b200 as s2l
    out = {}
    try:
        g = torch.Generator(device="cpu").manual_seed(0)
        lip = torch.rand(B, lh, lw, 3, generator=g).to(dev)
        face = torch.rand(B, h, w, 3, generator=g).to(dev)
        gt = torch.rand(B, h, w, 3, generator=g).to(dev)
        mask = torch.zeros(B, h, w, 3, device=dev)
        mask[:, y0 + 5:y0 + lh - 5, x0 + 5:x0 + lw - 5] = 1
        ys, xs = torch.meshgrid(torch.linspace(-1, 1, h), torch.linspace(-1, 1, w), indexing="ij")
        coord = (torch.stack([xs, ys], -1)[None].repeat(B, 1, 1, 1) * 1.02 + 0.01).to(dev)

        def timeit(fn, n=10):
            fn(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record(); e1.synchronize()
            return e0.elapsed_time(e1) / n
        ms = timeit(lambda: s2l.post_fusion_compose(lip, face, gt, mask, coord, x0, y0, True, lw // 5, want_canonical=False))
        # unique bytes the result depends on: coord + gt + out for every pixel; face and mask only under the warped (expanded)
        # lip rectangle (tf_nerf.py:354-363) — elsewhere the output is the ground-truth pixel; the lip crop
        pad = lw // 5
        rect = (lh + 3 * pad) * (lw + 2 * pad)
        alg = B * (h * w * (8 + 12 + 12) + rect * 12 + rect * 12 + lh * lw * 12)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        peak = peaks.get("hbm_gbs", 6650.0)
        rec = {"frames_per_s": B / (ms * 1e-3), "ms_per_step": ms, "frames_per_step": B, "algorithmic_bytes": alg,
               "achieved_gbs": alg / ms / 1e6, "hbm_peak_gbs": peak, "frac_of_hbm_peak": alg / ms / 1e6 / peak}
        try:
            from oracle import ref_runner as RR
            if RR.available():
                rm = RR.model(2, 3, dev)
                rm.post_fusion_unet = torch.nn.Identity()
                with torch.no_grad():
                    ref = lambda: rm.post_fusion2_onlylip(lip, face, gt, mask, x0, y0, coord, use_canonical_space=True)
                    fused = s2l.post_fusion_compose(lip, face, gt, mask, coord, x0, y0, True, lw // 5, want_canonical=False)[0]
                    rec["max_abs_vs_reference"] = float((fused.permute(0, 2, 3, 1) - ref()[1]).abs().max().item())
                    rec["reference_eager_ms"] = timeit(ref, 5)
                rec["speedup_vs_reference_eager"] = rec["reference_eager_ms"] / ms
        except Exception as e:
            rec["reference_eager_error"] = str(e)[:200]
        out["post_fusion_%dx%d" % (h, w)] = rec
    except Exception as e:
        out["post_fusion_%dx%d" % (h, w)] = {"error": str(e)[:300], "frames_per_s": 0.0, "ms_per_step": 0.0}
    return out


def train_step_extras(dev, sizes=((80, 120), (256, 256)), frames=4, window=5, reps=3):
    """BASELINE.json configs[4]: a 4-frame training batch, each frame with its 5-frame sync-expert window
    (training.py:404-559: predict_lip_image for the frame, training.py:500-525: five more for the window), forward +
    backward + optimizer step, bf16 on tensor cores — against the reference's own modules (PyTorch eager fp32 autograd,
    TF32 off) doing the same renders through Trainer.predict_lip_image on the same GPU.  Timed: the lip-render part of the
    step (AudioNet + 4-tap MLP renders + photometric loss + backward + SGD step); the UNet / SyncNet / LPIPS tail of a
    real step is cuDNN code shared by both arms and is not included.  Wall clock, synchronised, per step."""
    import speech2lip_b200 as s2l
    from speech2lip_b200 import synth
    out = {}
    cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "may_cfg.json")))
    G = 1 + window                                    # renders per frame
    for (H, W) in sizes:
        key = "train_step_config5_%dx%d" % (H, W)
        try:
            sd = {k: torch.from_numpy(v) for k, v in synth.make_state_dict(0, "kaiming", 2, 3).items()}
            gen = torch.Generator().manual_seed(5)
            audio = torch.from_numpy(synth.make_audio(frames * G, seed=21)).to(dev)
            index = torch.arange(frames * G)
            target = torch.rand(frames * G, H, W, 3, generator=gen).to(dev)
            m = s2l.TalkingFace(device=dev, cfg=cfg).to(dev).train()
            m.load_state_dict(sd, strict=False)
            opt = torch.optim.SGD([p for n, p in m.named_parameters() if not n.startswith(("post_fusion", "canonical", "coord_"))], lr=1e-5)

            def ours_step(group):
                """group = frames per launch: 1 -> one launch per frame (its 6 renders), `frames` -> the whole batch in one launch"""
                for f0 in range(0, frames, group):
                    sl = slice(f0 * G, (f0 + group) * G)
                    opt.zero_grad(set_to_none=True)
                    rgb = m.render_lip_train(audio[sl], index[sl], H, W)
                    ((rgb - target[sl]) ** 2).mean().backward()
                    opt.step()

            def timeit(fn, n):
                fn(); torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(n):
                    fn()
                torch.cuda.synchronize()
                return (time.perf_counter() - t0) / n * 1e3
            pts = frames * G * H * W * 4
            ws_gb = pts * 9.3e3 / 1e9
            rec = {"frames": frames, "renders_per_frame": G, "point_evals_per_step": pts, "precision": "bf16 (fp32 accumulate)",
                   "ours_ms_per_frame_launches": timeit(lambda: ours_step(1), reps)}
            if ws_gb < 60:
                rec["ours_ms_one_launch"] = timeit(lambda: ours_step(frames), reps)
            best = min(v for k, v in rec.items() if k.startswith("ours_ms"))
            rec["ms_per_step"] = best
            rec["frames_per_s"] = frames / (best * 1e-3)
            rec["train_tflops_algorithmic"] = 3 * pts * FLOP_PER_POINT["ensemble4"] / (best * 1e-3) / 1e12
            # ---- the reference's own modules, eager fp32 autograd
            try:
                from oracle import ref_runner as RR
                if RR.available():
                    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
                    torch.backends.cuda.matmul.allow_tf32 = False
                    torch.backends.cudnn.allow_tf32 = False
                    rm = RR.model(2, 3, dev, mode="train")
                    tr = RR.trainer(rm, dev, H, W)
                    ropt = torch.optim.SGD([p for n, p in rm.named_parameters() if not n.startswith(("post_fusion", "canonical", "coord_"))], lr=1e-5)
                    coords = RR.ns().get_coords(W, H, dev)

                    def ref_step():
                        for f in range(frames):
                            ropt.zero_grad(set_to_none=True)
                            loss = 0
                            for j in range(G):
                                i = f * G + j
                                rgb = tr.predict_lip_image(0, coords, audio[i:i + 1], None, {"index": index[i:i + 1].to(dev)}, None, None, None)
                                loss = loss + ((rgb.view(H, W, 3) - target[i]) ** 2).mean() / G
                            loss.backward()
                            ropt.step()
                    rec["reference_eager_ms"] = timeit(ref_step, 2)
                    rec["speedup_vs_reference_eager"] = rec["reference_eager_ms"] / best
                    # the UNMODIFIED caller: the reference's own Trainer.predict_lip_image driving the drop-in, call by call
                    # (4 rgb_forward calls per render), exact fp32 kernels and the opt-in bf16 tensor-core per-call path
                    mtr = RR.trainer(m, dev, H, W)

                    def dropin_step():
                        for f in range(frames):
                            opt.zero_grad(set_to_none=True)
                            loss = 0
                            for j in range(G):
                                i = f * G + j
                                rgb = mtr.predict_lip_image(0, coords, audio[i:i + 1], None, {"index": index[i:i + 1].to(dev)}, None, None, None)
                                loss = loss + ((rgb.view(H, W, 3) - target[i]) ** 2).mean() / G
                            loss.backward()
                            opt.step()
                    for tp in ("fp32", "bf16"):
                        m.train_precision = tp
                        rec["unmodified_trainer_ms_%s" % tp] = timeit(dropin_step, 2)
                    m.train_precision = "fp32"
                    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
                    del rm, tr, ropt
            except Exception as e:
                rec["reference_eager_error"] = str(e)[:200]
            out[key] = rec
            del m, opt
            torch.cuda.empty_cache()
        except Exception as e:
            out[key] = {"error": str(e)[:300], "frames_per_s": 0.0, "ms_per_step": 0.0}
    return out


def run_reference_arm(a):
    """`--impl reference`: the reference's own CPU implementation of the path on all host cores; each step = ONE WHOLE
    frame of the workload (no extrapolation) unless --cpu-rays bounds it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    run, frac, desc, kind = reference_runner(a, torch.device("cpu"), a.cpu_rays if a.cpu_rays > 0 else None)
    for _ in range(max(1, min(a.warmup, 1))):
        run()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        run()
    dt = (time.perf_counter() - t0) / a.steps
    fps = frac / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": 1, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "mode": a.mode,
                   "reference_arm": "the reference's own modules (unmodified copy in oracle/_ref, oracle/build_ref.py) on the host cores, "
                                    "torch CPU fp32, no_grad" if kind == "reference" else "oracle port (oracle/_ref missing on this box)",
                   "frames_per_step": frac},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind, "sample": "each step = " + desc},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def stop(self, t_begin, t_end):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [l.split(", ") for (t, l) in self.samples if t_begin <= t <= t_end] or [l.split(", ") for (_, l) in self.samples[-3:]]
        sm, reasons, mx, pw = [], set(), None, []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                pw.append(float(r[2]))
                for n, v in zip(names, r[3:7]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "power_w_max": max(pw) if pw else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------- GPU arm
def run_gpu_arm(a):
    import ctypes as C
    import torch.distributed as dist
    import speech2lip_b200 as s2l
    from speech2lip_b200 import _cabi, renderer as R
    from speech2lip_b200.dist import broadcast_params
    from speech2lip_b200 import synth   # synthetic weights/inputs (generators only; nothing from oracle/ on this path)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    vol = a.mode == "volumetric"
    uvd, och = (3, 4) if vol else (2, 3)
    H = W = a.size
    F = a.frames
    # weights: rank 0 owns them, everyone else receives ONE broadcast; then zero communication
    sd = {k: torch.from_numpy(v).to(dev) for k, v in synth.make_state_dict(0 if rank == 0 else 1000 + rank, "kaiming", uvd, och).items()}
    if world > 1:
        broadcast_params(sd, src=0)
    w = s2l.PackedWeights(sd, uvd, och)
    rend = s2l.LipRenderer(w, a.precision)
    lib = _cabi.lib()

    # per-rank frames (weak scaling: F frames per GPU per step), host-resident pinned inputs for the e2e leg
    audio_h = torch.from_numpy(synth.make_audio(F, seed=100 + rank)).pin_memory()
    index_h = (torch.arange(F, dtype=torch.int64) + rank * F).pin_memory()
    c2w_h = torch.eye(4)[:3].contiguous().pin_memory()
    out_h = torch.empty(F, H, W, 3, pin_memory=True)
    audio_d, index_d, c2w_d = audio_h.to(dev), index_h.to(dev), c2w_h.to(dev)
    z_d = torch.linspace(0., 1., a.samples, device=dev) if vol else None
    rgb_d = torch.empty(F, H, W, 3, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)        # > 126 MB L2

    geom = _cabi.S2LGeom(n_frames=F, height=H, width=W, n_samples=a.samples if vol else 0,
                         pts_mode={"volumetric": _cabi.PTS_RAYS, "plain": _cabi.PTS_GRID, "ensemble4": _cabi.PTS_GRID_ENS4}[a.mode],
                         uv_dims=uvd, out_ch=och, z_per_ray=0, rays_per_frame_shared=1, pts_per_frame=0, eps_shift=0.001)
    P = points_per_frame(a)
    ro_d = torch.empty(H, W, 3, device=dev)
    rd_d = torch.empty(H, W, 3, device=dev)
    prec = _cabi.PRECISIONS[a.precision]
    st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = R._ptr
    scratch_d = torch.empty(lib.s2l_render_scratch_bytes(C.byref(geom), prec, 0), dtype=torch.uint8, device=dev)

    def step_device():
        """the whole-path C-ABI call on device-resident inputs: ray generation, then s2l_render_frames = AudioNet + per-frame
        biases -> fused MLP with the per-pixel reduction in its epilogue -> fp32 re-evaluation of the listed rays.  The
        fused-MLP launches are bracketed by CUDA events inside the library (s2l_profile_enable) for the roofline."""
        if vol:
            _cabi.check(lib.s2l_get_rays(p(c2w_d), H, W, 1200.0, p(ro_d), p(rd_d), st()), "get_rays")
        _cabi.check(lib.s2l_render_frames(p(w.blob), C.byref(geom), p(audio_d), p(index_d), p(ro_d) if vol else None,
                                          p(rd_d) if vol else None, p(z_d), p(rgb_d), None, None, p(scratch_d), prec, st()), "render")

    def step_e2e():
        """public API on HOST buffers: H2D (windows, indices, pose) -> render -> D2H (frames), all on the current stream.
        (Splitting the step into chunks whose D2H overlaps the next chunk's render — LipRenderer.render_sequence_host — was
        measured here: 49.5 ms against 49.0 ms, the per-launch fixed work of four 2-frame launches costs more than the 6 MB copy.)"""
        ad = audio_h.to(dev, non_blocking=True)
        idd = index_h.to(dev, non_blocking=True)
        if vol:
            cd = c2w_h.to(dev, non_blocking=True)
            ro, rd = R.get_rays(H, W, 1200.0, cd)
            rgb = rend.render_frames(ad, idd, H, W, mode="volumetric", rays_o=ro, rays_d=rd, z_vals=z_d, out=rgb_d)
        else:
            rgb = rend.render_frames(ad, idd, H, W, mode=a.mode, eps_shift=0.001, out=rgb_d)
        out_h.copy_(rgb, non_blocking=True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_once(fn):
        flush.fill_(1)                                        # evict L2 between timed iterations (untimed)
        e0, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e3.record()
        e3.synchronize()
        return e0.elapsed_time(e3)

    # ---- warm-up (untimed), then the timed region bracketed by barrier + synchronize
    for _ in range(max(a.warmup, 3)):
        step_device()
        step_e2e()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    t_begin = time.perf_counter()
    barrier()
    # K device-resident steps and K end-to-end steps, INTERLEAVED: under the 1 kW power cap the SM clock sags over the first
    # second of load, so two back-to-back loops would time the second leg at lower clocks than the first (that, not the
    # copies, was most of the gap between `value` and `e2e`: the copies and host work cost 0.45 ms of a 48 ms step,
    # tools/exp_e2e_breakdown.py).  Every step is bracketed by its own CUDA events; launches and the in-library kernel
    # events are collected over the device-resident steps only.
    dev_ms = e2e_ms = ker_ms = 0.0
    launches = 0
    for _ in range(a.steps):
        lib.s2l_launch_count(1)
        lib.s2l_profile_enable(1)
        dev_ms += timed_once(step_device)
        ker_ms += float(lib.s2l_profile_mlp_ms(None))
        lib.s2l_profile_enable(0)
        launches += int(lib.s2l_launch_count(0))
        e2e_ms += timed_once(step_e2e)
    barrier()
    t_end = time.perf_counter()
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None

    # ---- extras (N=1 only, after the timed headline region): other precisions of the same workload and the live modes
    extras = {}
    if world == 1 and not a.no_extras:
        def quick(fn, steps=3):
            fn(); torch.cuda.synchronize()
            tot = 0.0
            for _ in range(steps):
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); e1.synchronize()
                tot += e0.elapsed_time(e1)
            return tot / steps
        for alt in ("bf16x3", "fp16f8", "bf16x1"):
            if alt == a.precision:
                continue
            r_alt = s2l.LipRenderer(w, alt)
            if vol:
                ms = quick(lambda: r_alt.render_frames(audio_d, index_d, H, W, mode="volumetric", rays_o=ro_d, rays_d=rd_d, z_vals=z_d, out=rgb_d))
            else:
                ms = quick(lambda: r_alt.render_frames(audio_d, index_d, H, W, mode=a.mode, eps_shift=0.001, out=rgb_d))
            extras["same_workload_" + alt] = {"frames_per_s": F / (ms * 1e-3), "ms_per_step": ms,
                                             "parity_mode": alt != "bf16x1"}
        if vol:
            # the exact fp32 kernel (CUDA-core FFMA, the arithmetic every parity number is measured against) on ONE frame of
            # the same workload: frames/s and its fraction of the fp32 FFMA peak (148 SMs x 128 lanes x 2 x SM clock)
            r32 = s2l.LipRenderer(w, "fp32")
            ms = quick(lambda: r32.render_frames(audio_d[:1], index_d[:1], H, W, mode="volumetric", rays_o=ro_d, rays_d=rd_d, z_vals=z_d,
                                                 out=rgb_d[:1]), steps=2)
            tf = points_per_frame(a) * FLOP_PER_POINT["volumetric"] / (ms * 1e-3) / 1e12
            extras["same_workload_fp32_exact"] = {"frames_per_s": 1.0 / (ms * 1e-3), "ms_per_step": ms, "frames_per_step": 1,
                                                  "tflops_algorithmic": tf, "frac_of_fp32_ffma_peak_at_1965mhz": tf / (148 * 128 * 2 * 1.965e9 / 1e12)}
        if vol and a.samples % 16 == 0:
            # Early ray termination (the only results-preserving lever toward the 500 frames/s target — at 64 samples
            # per ray 500 frames/s is 2.6 PFLOP/s of algorithmic work, above the chip's measured dense bf16 peak even at ONE MMA
            # per product: 100 % of 1.66 PF = 317 frames/s).  Front-to-back chunks of 16 samples, rays with T < 1e-4 finished,
            # survivors compacted on the device.  What it buys depends on the scene's density: swept by scaling the density
            # row of output_linear (x1 = the synthetic kaiming scene, nearly transparent: nothing terminates).
            ert = {}
            for scale in (1.0, 30.0, 300.0):
                sdn = synth.make_state_dict(0, "kaiming", uvd, och)
                sdn["output_linear.weight"] = sdn["output_linear.weight"].copy()
                sdn["output_linear.bias"] = sdn["output_linear.bias"].copy()
                sdn["output_linear.weight"][3] *= scale
                sdn["output_linear.bias"][3] *= scale
                w_s = s2l.PackedWeights({k: torch.from_numpy(v).to(dev) for k, v in sdn.items()}, uvd, och)
                r_s = s2l.LipRenderer(w_s, a.precision)
                kw = dict(mode="volumetric", rays_o=ro_d, rays_d=rd_d, z_vals=z_d)
                full = torch.empty_like(rgb_d)
                ms_all = quick(lambda: r_s.render_frames(audio_d, index_d, H, W, out=full, **kw))
                ms_ert = quick(lambda: r_s.render_frames(audio_d, index_d, H, W, out=rgb_d, sample_chunks=a.samples // 16, term_thr=1e-4, **kw))
                alive = r_s.last_render_counts()["alive"].sum(1).tolist()
                ert["density_x%g" % scale] = {"all_samples_frames_per_s": F / (ms_all * 1e-3), "with_termination_frames_per_s": F / (ms_ert * 1e-3),
                                              "rays_alive_at_chunk_start": [F * H * W] + [int(x) for x in alive],
                                              "max_abs_difference": float((full - rgb_d).abs().max().item())}
            extras["early_ray_termination"] = {"frames_per_s": ert["density_x300"]["with_termination_frames_per_s"], "ms_per_step": 0.0,
                                               "term_thr": 1e-4, "chunk_samples": 16, "sweep": ert,
                                               "note": "frames_per_s = the densest scene of the sweep; all_samples is the headline mode"}
        if vol:
            sdL = {k: torch.from_numpy(v).to(dev) for k, v in synth.make_state_dict(0, "kaiming", 2, 3).items()}
            wL = s2l.PackedWeights(sdL, 2, 3)
            FL = 64
            aL = torch.from_numpy(synth.make_audio(FL, seed=7)).to(dev)
            iL = torch.arange(FL, device=dev)
            oL = torch.empty(FL, H, W, 3, device=dev)
            for mode_l in ("plain", "ensemble4"):
                for prec_l in (a.precision, "bf16x3"):
                    rl = s2l.LipRenderer(wL, prec_l)
                    ms = quick(lambda: rl.render_frames(aL, iL, H, W, mode=mode_l, eps_shift=0.001, out=oL))
                    extras["live_%s_%dx%d_%s" % (mode_l, H, W, prec_l)] = {"frames_per_s": FL / (ms * 1e-3), "ms_per_step": ms,
                                                                          "frames_per_step": FL}

        if vol:
            # live sequence HOST -> HOST: pinned windows in, uint8 BGR frames out (what inference.py:173-178 writes), D2H of chunk i
            # under the render of chunk i+1 (LipRenderer.render_sequence_host)
            Ts = 512
            ah = torch.from_numpy(synth.make_audio(Ts, seed=11)).pin_memory()
            ih = torch.arange(Ts).pin_memory()
            oh = torch.empty(Ts, H, W, 3, dtype=torch.uint8, pin_memory=True)
            rs = s2l.LipRenderer(wL, a.precision)
            rs.render_sequence_host(ah, ih, H, W, 64, "bgr8", oh)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rs.render_sequence_host(ah, ih, H, W, 64, "bgr8", oh)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            extras["live_sequence_host_to_host_bgr8_%dx%d_%s" % (H, W, a.precision)] = {
                "frames_per_s": Ts / dt, "ms_per_step": dt / (Ts / 64) * 1e3, "frames_per_step": 64,
                "d2h_bytes_per_step": 64 * H * W * 3, "what": "512 frames, pinned host windows -> uint8 BGR host frames, wall clock"}
        # the UNMODIFIED caller: inference.py:144-159 verbatim through the TalkingFace drop-in, one frame per iteration
        # (tile the window H*W times, audio_merge_forward, cat with the uv grid, rgb_forward)
        try:
            cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "may_cfg.json")))
            tf = s2l.TalkingFace(device=dev, cfg=cfg, mode="eval").to(dev).eval()
            tf.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_state_dict(0, "kaiming", 2, 3).items()}, strict=False)
            wins = torch.from_numpy(synth.make_audio(32, seed=9)).to(dev)
            vv, uu = torch.meshgrid(torch.linspace(0.0, 1.0, H, device=dev), torch.linspace(0.0, 1.0, W, device=dev), indexing="ij")
            coords = torch.stack([uu, vv], -1).view(-1, 2)                   # the caller's get_coords (rendering.py:9-28)

            def dropin_frame(i):
                with torch.no_grad():
                    au = wins[i:i + 1].tile(H * W, 1, 1)
                    ab = tf.audio_merge_forward(au)
                    xx = torch.cat([coords[:, None, :], ab[:, None, :]], -1).view(-1, tf.audio_dims + 2)
                    return tf.rgb_forward(xx, time_pts=torch.tensor([i], device=dev), rgb_pts=None)[:, :3]
            for i in range(8):
                dropin_frame(i)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(96):
                dropin_frame(i % 32)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            extras["drop_in_inference_loop_%dx%d" % (H, W)] = {"frames_per_s": 96 / dt, "ms_per_step": dt / 96 * 1e3, "frames_per_step": 1,
                                                              "precision": tf.dropin_precision,
                                                              "what": "inference.py:144-159 call sequence through TalkingFace, wall clock incl. Python"}
        except Exception as e:      # an extra must never take the headline line down
            extras["drop_in_inference_loop"] = {"error": str(e)[:200], "frames_per_s": 0.0, "ms_per_step": 0.0}
        try:
            extras.update(eager_gpu_reference(a, dev))
        except Exception as e:
            extras["reference_gpu_eager"] = {"error": str(e)[:200], "frames_per_s": 0.0, "ms_per_step": 0.0}
        extras.update(train_step_extras(dev))
        extras.update(post_fusion_extras(dev))

    t = torch.tensor([dev_ms, e2e_ms, ker_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, ker_ms = [float(x) for x in t.tolist()]
    checksum = float(rgb_d.double().sum().item())
    finite = bool(torch.isfinite(rgb_d).all().item())

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        tensor_bound = a.precision != "fp32"
        if tensor_bound:
            peak = peaks.get("bf16_tflops_sustained") or peaks.get("bf16_tflops")
            peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks else "fallback 1590 (B200_PROFILING.md)"
            peak = peak or 1590.0
        else:
            peak, peak_src = 72.0, "nominal fp32 FFMA peak 148 SM x 128 lanes x 2 x 1.9 GHz (no measured fp32 figure)"
        flop_launch = float(F) * P * FLOP_PER_POINT[a.mode]
        tc_sched = int(lib.s2l_tc_schedule((F * P + 127) // 128))      # 1 single CTAs, 2 CTA pairs, 3 multicast clusters
        ach = flop_launch / (ker_ms / a.steps * 1e-3) / 1e12
        frames_total = F * world * a.steps
        line = {
            "metric": METRIC, "value": frames_total / (dev_ms * 1e-3), "unit": "frames/s", "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": dev_ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None,
            "dtype": {"bf16x3": "bf16x3-split (fp32 accumulate)", "bf16x1": "bf16", "fp32": "f32", "fp16f8": "fp16 + 2x fp8 corrections (fp32 accumulate)"}[a.precision], "data": "synthetic",
            "config": {"workload": workload_name(a), "mode": a.mode, "precision": a.precision, "tc_schedule": tc_sched,
                       "points_per_frame": P, "weights": "synthetic kaiming-normal ('trained-like'), seed 0",
                       "l2": "flushed between timed steps (256 MiB fill, untimed); per-step CUDA events summed; device-resident and end-to-end steps interleaved",
                       "parallelism": "frames sharded across ranks, 1 NCCL weight broadcast at start, none during render"},
            "e2e": {"value": frames_total / (e2e_ms * 1e-3), "unit": "frames/s",
                    "h2d_bytes_per_step": int(audio_h.numel() * 4 + index_h.numel() * 8 + (48 if vol else 0)),
                    "d2h_bytes_per_step": int(out_h.numel() * 4), "ms_per_step": e2e_ms / a.steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor" if tensor_bound else "fp32", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                         "traffic": ncu_traffic(F, P, a.precision), "kernel": ({1: "mlp_tc_kernel", 2: "mlp_tc2_kernel (CTA pairs)", 3: "mlp_tc_kernel (2-CTA multicast weights)"}[tc_sched] if tensor_bound else "mlp_fp32_kernel"),
                         "kernel_ms_per_launch": ker_ms / a.steps, "flop_per_point_algorithmic": FLOP_PER_POINT[a.mode],
                         "mma_multiplier": {"bf16x3": 3, "fp16f8": 2}.get(a.precision, 1), "peak_source": peak_src,
                         "frac_of_burst_peak": ach / peaks["bf16_tflops"] if tensor_bound and peaks.get("bf16_tflops") else None,
                         # tensor-pipe view: MMA math actually issued (3 bf16 MMAs per product in the parity mode, after folding)
                         "issued_mma_tflops": (float(F) * P * ISSUED_MAC_PER_POINT * 2 * {"bf16x3": 3, "fp16f8": 2}.get(a.precision, 1)
                                               / (ker_ms / a.steps * 1e-3) / 1e12) if tensor_bound else None,
                         "issued_mma_frac_of_peak": (float(F) * P * ISSUED_MAC_PER_POINT * 2 * {"bf16x3": 3, "fp16f8": 2}.get(a.precision, 1)
                                                     / (ker_ms / a.steps * 1e-3) / 1e12 / peak) if tensor_bound else None},
            "clocks": clocks, "checksum": checksum, "finite": finite,
        }
        if world == 1 and not a.no_extras:
            line["extras"] = extras
        if world == 1 and not a.no_cpu_baseline:
            # the reference's own code on this box's host cores: a 4096-ray warm-up, then ONE whole frame (about 10-20 s)
            threads = os.cpu_count() or 1
            torch.set_num_threads(threads)
            warm, _, _, _ = reference_runner(a, torch.device("cpu"), 4096)
            warm()
            run, frac, desc, kind = reference_runner(a, torch.device("cpu"), a.cpu_rays if a.cpu_rays > 0 else None)
            t0 = time.perf_counter()
            run()
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": frac / dt, "unit": "frames/s", "cores": threads, "kind": kind,
                                    "sample": desc + "; one timed repetition after a 4096-ray warm-up"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def ncu_traffic(frames, pts_per_frame, precision):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the fused MLP kernel from the committed `ncu --set full`
    capture of THIS launch geometry (profiles/ncu_traffic.json: one record per captured geometry; no scaling between
    geometries — L2 residency changes with the launch size); None when that geometry was never captured."""
    try:
        import json as _j
        recs = _j.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "ncu_traffic.json")))
        for t in recs if isinstance(recs, list) else [recs]:
            if t["points_per_launch"] == frames * pts_per_frame and t.get("precision") == precision and t.get("fused_epilogue"):
                return t["dram_bytes_read"] + t["dram_bytes_write"]
    except Exception:
        pass
    return None


def run_config3(a):
    """BASELINE.json configs[2]: a 1000-frame 256x256x64 sequence sharded across the ranks in contiguous blocks
    (speech2lip_b200.dist.shard_frames), ONE NCCL broadcast of the weights, no communication during the render, an optional
    NCCL gather of the finished uint8 frames — STRONG scaling (total work fixed).  After the timed pass every rank
    re-renders probe frames owned by OTHER ranks and compares them bit for bit with the owners' per-frame checksums (all-gathered),
    and the post-broadcast weights are hashed on every rank."""
    import ctypes as C
    import torch.distributed as dist
    import speech2lip_b200 as s2l
    from speech2lip_b200 import _cabi, renderer as R, synth
    from speech2lip_b200.dist import broadcast_params, shard_frames, gather_frames
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    H = W = a.size
    T, FB = a.seq_frames, a.frames
    # every rank starts from DIFFERENT weights; only the broadcast makes them equal (a silent broadcast failure shows below)
    sd = {k: torch.from_numpy(v).to(dev) for k, v in synth.make_state_dict(0 if rank == 0 else 1000 + rank, "kaiming", 3, 4).items()}
    if world > 1:
        broadcast_params(sd, src=0)
    w = s2l.PackedWeights(sd, 3, 4)
    whash = torch.stack([v.double().sum() for v in sd.values()]).sum().reshape(1)
    rend = s2l.LipRenderer(w, a.precision)
    lo, hi = shard_frames(T, rank, world)
    audio_all = torch.from_numpy(synth.make_audio(T, seed=300)).to(dev)       # the same sequence on every rank; a rank touches its block
    ro, rd = R.get_rays(H, W, 1200.0, torch.eye(4, device=dev)[:3])
    z = torch.linspace(0., 1., a.samples, device=dev)
    out = torch.empty(max(hi - lo, 1), H, W, 3, device=dev)

    def render_block(f0, f1, dst):
        for s0 in range(f0, f1, FB):
            s1 = min(s0 + FB, f1)
            rend.render_frames(audio_all[s0:s1], torch.arange(s0, s1), H, W, mode="volumetric", rays_o=ro, rays_d=rd, z_vals=z,
                               out=dst[s0 - f0:s1 - f0])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(max(a.warmup, 3)):                         # warm-up: a few launches, not the whole sequence
        render_block(lo, min(lo + FB, hi), out)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t_begin = time.perf_counter()
    lib = _cabi.lib()
    lib.s2l_launch_count(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(a.steps):
        render_block(lo, hi, out)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = lib.s2l_launch_count(0)
    t_end = time.perf_counter()
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    # ---- optional gather of the finished frames (uint8 BGR, what the reference writes) onto every rank: one all_gather
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    u8 = R.frames_to_bgr8(out[:hi - lo])
    if world > 1:
        gather_frames(u8[:1], world)                          # warm-up: NCCL sets up the all_gather channels on first use
    barrier()
    g0.record()
    allf = gather_frames(u8, T) if world > 1 else u8
    g1.record()
    barrier()
    gather_ms = g0.elapsed_time(g1)
    # ---- cross-rank checks
    # per-frame checksum = integer sum of the fp32 bit patterns: exact and independent of the reduction order
    sums = out[:hi - lo].view(torch.int32).to(torch.int64).sum(dim=(1, 2, 3))
    per = (T + world - 1) // world
    pad = torch.zeros(per, dtype=torch.int64, device=dev)
    pad[:hi - lo] = sums
    if world > 1:
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad)
        all_sums = torch.cat(parts)[:T]
        hashes = [torch.empty_like(whash) for _ in range(world)]
        dist.all_gather(hashes, whash)
        weights_equal = all(bool(torch.equal(h, hashes[0])) for h in hashes)
    else:
        all_sums, weights_equal = sums, True
    probes = sorted({0, T // 7, T // 3, T // 2, (2 * T) // 3, T - 1})
    bad = 0
    tmp = torch.empty(1, H, W, 3, device=dev)
    for pf in probes:                                         # every rank renders every probe itself
        render_block(pf, pf + 1, tmp)
        bad += int(tmp.view(torch.int32).to(torch.int64).sum().item() != all_sums[pf].item())
        if world > 1:
            own = allf[pf].to(dev)
            bad += int(not torch.equal(R.frames_to_bgr8(tmp)[0], own))
    badt = torch.tensor([bad], device=dev)
    mx = torch.tensor([ms, gather_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(badt)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    ms, gather_ms = [float(x) for x in mx.tolist()]
    if rank == 0:
        line = {"metric": METRIC, "value": T * a.steps / (ms * 1e-3), "unit": "frames/s", "n_gpus": world, "steps": a.steps,
                "warmup": max(a.warmup, 3), "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": a.precision, "data": "synthetic",
                "config": {"workload": "%d-frame sequence, volumetric %dx%d x %d samples/ray, contiguous frame blocks per rank, %d frames per launch "
                                       "(BASELINE configs[2]); a step = the whole sequence" % (T, H, W, a.samples, FB),
                           "precision": a.precision, "frames_per_rank": per, "l2": "inputs larger than L2 (a rank's block streams through)",
                           "parallelism": "1 NCCL weight broadcast at start, none during render, 1 optional all_gather of the uint8 frames"},
                "gpu_launches": int(launches), "clocks": clocks,
                "gather_frames_ms": gather_ms, "gather_bytes": int(T * H * W * 3),
                "checks": {"weights_equal_on_all_ranks_after_broadcast": weights_equal, "probe_frames": probes,
                           "probe_mismatches_summed_over_ranks": int(badt.item()),
                           "what": "every rank re-rendered the probe frames and compared checksum (and the gathered uint8 frame) with the owner's"},
                "e2e": {"value": T * a.steps / ((ms + gather_ms * a.steps) * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": int(T * 16 * 29 * 4),
                        "d2h_bytes_per_step": 0, "what": "render + NCCL gather of the finished uint8 frames to every rank"}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_config5(a):
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    ex = train_step_extras(dev)
    r = ex.get("train_step_config5_80x120", {})
    line = {"metric": "training frames/s (4-frame batch with the 5-frame sync-window renders, lip crop 80x120, fwd+bwd+SGD)", "value": r.get("frames_per_s", 0.0),
            "unit": "frames/s", "n_gpus": 1, "steps": 3, "warmup": 1, "ms_per_step": r.get("ms_per_step", 0.0), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16 (fp32 accumulate)", "data": "synthetic",
            "config": {"workload": "BASELINE configs[4]: training step, 4 frames x (1 + 5 sync-window) renders, 4 taps", "detail": ex}}
    print(json.dumps(line))


def run_config5_dp(a, H=80, W=120, frames=4, window=5):
    """BASELINE.json configs[4] on N GPUs: data-parallel over frames as the reference trains (DistributedSampler, train.py:102;
    DDP, training.py:40) — every rank renders ITS 4 frames x (1 + 5 sync-window) renders, forward + backward on the tensor-core
    training kernels, then ONE gradient exchange of the flat hot-path bucket (speech2lip_b200.dist.GradExchange: the library's
    one-kernel all-reduce over NVLink peer memory) and the SGD step.  Weak scaling (4 frames per GPU).  Reports the step with
    the peer kernel, the same step with one NCCL all_reduce instead, both exchanges timed alone, and checks that the weights
    are bit-identical on every rank after the timed steps."""
    import torch.distributed as dist
    import speech2lip_b200 as s2l
    from speech2lip_b200 import synth
    from speech2lip_b200.dist import GradExchange, broadcast_module
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "may_cfg.json")))
    G = 1 + window
    sd = {k: torch.from_numpy(v) for k, v in synth.make_state_dict(rank, "kaiming", 2, 3).items()}   # ranks start DIFFERENT
    m = s2l.TalkingFace(device=dev, cfg=cfg).to(dev).train()
    m.load_state_dict(sd, strict=False)
    broadcast_module(m, src=0)                                            # DDP's constructor broadcast
    m.invalidate_packed()
    gen = torch.Generator().manual_seed(100 + rank)                       # every rank its own frames
    audio = torch.from_numpy(synth.make_audio(frames * G, seed=21 + rank)).to(dev)
    index = torch.arange(frames * G) + rank * frames * G
    target = torch.rand(frames * G, H, W, 3, generator=gen).to(dev)
    hot = [p for n, p in m.named_parameters() if not n.startswith(("post_fusion", "canonical", "coord_"))]
    opt = torch.optim.SGD(hot, lr=1e-5)
    exch = {"peer": GradExchange(hot, method="peer"), "nccl": GradExchange(hot, method="collective")}

    def step(method):
        opt.zero_grad(set_to_none=True)
        rgb = m.render_lip_train(audio, index, H, W)
        ((rgb - target) ** 2).mean().backward()
        exch[method].allreduce()
        opt.step()

    def timed(fn, n, warm):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize(); dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    K, Wm = max(a.steps, 1), max(a.warmup, 3)
    mon = ClockSampler(local) if rank == 0 else None
    if mon:
        mon.start()
    t_begin = time.perf_counter()
    ms_peer = timed(lambda: step("peer"), K, Wm)
    clocks = mon.stop(t_begin, time.perf_counter()) if mon else None
    ms_nccl = timed(lambda: step("nccl"), K, Wm)
    for p in hot:                                                         # the exchanges alone, on a populated bucket
        p.grad = torch.randn_like(p)
    ex_peer = timed(lambda: exch["peer"].allreduce(), 50, 5)
    ex_nccl = timed(lambda: exch["nccl"].allreduce(), 50, 5)
    # after identical averaged gradients the weights must still be bit-identical on every rank
    digest = torch.stack([p.detach().view(torch.int32).sum(dtype=torch.int64) for p in hot]).sum().reshape(1)
    alld = [torch.empty_like(digest) for _ in range(world)]
    dist.all_gather(alld, digest)
    same = all(bool(torch.equal(d, alld[0])) for d in alld)
    n_floats = exch["peer"].n
    exch["peer"].close()
    if rank == 0:
        line = {"metric": "training frames/s (4-frame batch per GPU with the 5-frame sync-window renders, lip crop 80x120, fwd+bwd+exchange+SGD)",
                "value": world * frames / (ms_peer * 1e-3), "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
                "ms_per_step": ms_peer, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16 (fp32 accumulate)", "data": "synthetic",
                "config": {"workload": "BASELINE configs[4] data-parallel: %d frames x (1 + 5 sync-window) renders x 4 taps per GPU" % frames,
                           "parallelism": "dp%d over frames; one flat gradient bucket of %d floats (%.2f MB) per step" % (world, n_floats, n_floats * 4 / 1e6)},
                "exchange": {"kernel": "s2l_allreduce_peer (one launch: signal, wait, sum %d payloads over NVLink peer loads in rank order)" % world,
                             "ms_peer_kernel": ex_peer, "ms_nccl_all_reduce": ex_nccl, "step_ms_with_peer_kernel": ms_peer,
                             "step_ms_with_nccl": ms_nccl, "bucket_bytes": n_floats * 4,
                             "weights_bit_identical_on_all_ranks_after_steps": same},
                "clocks": clocks}
        print(json.dumps(line))
    dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference_arm(a)
    elif a.config == 3:
        run_config3(a)
    elif a.config == 5:
        if int(os.environ.get("WORLD_SIZE", "1")) > 1:
            run_config5_dp(a)
        elif int(os.environ.get("RANK", "0")) == 0:
            run_config5(a)
    else:
        run_gpu_arm(a)


if __name__ == "__main__":
    main()
