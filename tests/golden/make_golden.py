"""Generates tests/golden/*.npz by running the REAL reference implementation.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

For every case the reference's own TalkingFace / Trainer.predict_lip_image /
get_coords / get_rays / density2outputs are executed on CPU (fp32) with the
synthetic weights of oracle/synth.py loaded via load_state_dict; inputs and the
reference's outputs are stored.  Weights are NOT stored — tests re-derive them
from (seed, kind) with oracle/synth.py.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import synth                      # noqa: E402
from oracle.ref_shim import load_reference    # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def ref_model(ns, sd_np, uv_dims, output_ch):
    torch.manual_seed(1234)   # only for the non-hot tensors (UNet, depth) that we do not touch
    m = ns.TalkingFace(device=torch.device("cpu"), cfg=ns.cfg, mode="eval",
                       uv_dims=uv_dims, output_ch=output_ch).eval()
    missing = m.load_state_dict({k: torch.from_numpy(v) for k, v in sd_np.items()}, strict=False)
    assert not missing.unexpected_keys, missing.unexpected_keys
    return m


def main():
    ns = load_reference()
    torch.set_num_threads(8)
    cases = {}

    for kind in ("default", "kaiming"):
        sdL = synth.make_state_dict(seed=0, kind=kind, uv_dims=2, output_ch=3)
        sdV = synth.make_state_dict(seed=0, kind=kind, uv_dims=3, output_ch=4)
        mL = ref_model(ns, sdL, 2, 3)
        mV = ref_model(ns, sdV, 3, 4)
        audio = torch.from_numpy(synth.make_audio(4, seed=1))

        with torch.no_grad():
            # ---- a1: AudioNet, both input orientations (tf_nerf.py:203-207)
            lat = mL.audio_merge_forward(audio)
            lat_t = mL.audio_merge_forward(audio.permute(0, 2, 1).contiguous())
            cases["audio_%s" % kind] = dict(audio=audio.numpy(), latent=lat.numpy(), latent_from_29x16=lat_t.numpy())

            # ---- a2/a3: embedders
            uv = torch.rand(17, 2, generator=torch.Generator().manual_seed(5))
            cases["embed"] = dict(uv=uv.numpy(), pe=mL.uv_embedder(uv).numpy(),
                                  t5=mL.time_embedder_new(torch.tensor([5])).numpy(),
                                  t6000=mL.time_embedder_new(torch.tensor([6000])).numpy(),
                                  coords_7x5=ns.get_coords(7, 5, torch.device("cpu")).numpy())

            # ---- a4 + inference.py:144-159 (plain, 1 eval / pixel), as written (audio tiled)
            for (H, W, idx) in ((24, 32, 5), (8, 8, 6000)):
                n = H * W
                a = audio[1:2].tile(n, 1, 1)
                coords = ns.get_coords(W, H, torch.device("cpu"))
                ab = mL.audio_merge_forward(a)
                x = torch.cat([coords[:, None, :], ab[:, None, :]], -1).view(-1, 66)
                out = mL.rgb_forward(x, time_pts=torch.tensor([idx]), rgb_pts=None)
                cases["plain_%s_%dx%d_i%d" % (kind, H, W, idx)] = dict(
                    audio=audio[1:2].numpy(), index=np.int64(idx), H=np.int64(H), W=np.int64(W),
                    rgb=out[:, :3].reshape(H, W, 3).numpy())

            # ---- a4 general contract: arbitrary per-row latent
            g = torch.Generator().manual_seed(9)
            xg = torch.cat([torch.rand(200, 2, generator=g), torch.randn(200, 64, generator=g) * 0.1], -1)
            cases["rowlatent_%s" % kind] = dict(
                x=xg.numpy(), index=np.int64(3),
                out=mL.rgb_forward(xg, time_pts=torch.tensor([3])).numpy())

            # ---- a5: 4-tap local ensemble through the real Trainer.predict_lip_image
            H, W, idx = 16, 24, 7
            tr = ns.Trainer(mL, None, torch.device("cpu"), "/tmp", cfg=ns.cfg, batch_rays=H * W,
                            use_audio_net=True, use_time=True, use_audio=True,
                            use_perceptual_loss=False, use_syncloss=False, multi_gpu=False)
            tr.height, tr.width = H, W            # set in train_stage1 (training.py:393-394)
            coords = ns.get_coords(W, H, torch.device("cpu"))
            for seed in (11, 12):
                torch.manual_seed(seed)
                eps_expected = (0.5 / H) * torch.rand(1) / 2.0            # training.py:198-200
                torch.manual_seed(seed)
                rgb = tr.predict_lip_image(0, coords, audio[2:3], None, {"index": torch.tensor([idx])},
                                           None, None, None)
                cases["ens4_%s_seed%d" % (kind, seed)] = dict(
                    audio=audio[2:3].numpy(), index=np.int64(idx), H=np.int64(H), W=np.int64(W),
                    eps=eps_expected.numpy(), rng_seed=np.int64(seed), rgb=rgb.reshape(H, W, 3).numpy())

            # ---- a7/a8 + Mode V (volumetric): uv_dims=3 / output_ch=4 model, get_rays, density2outputs
            for (H, W, S, focal) in ((8, 8, 16, 1200.0), (6, 10, 64, 12.0)):
                c2w = torch.eye(4)[:3, :4].clone()
                c2w[:, 3] = torch.tensor([0.05, -0.02, 0.3])
                rays_o, rays_d = ns.get_rays(H, W, focal, c2w, torch.device("cpu"))
                rays_o = rays_o.reshape(-1, 3)
                rays_d = rays_d.reshape(-1, 3)
                z = torch.linspace(0., 1., S).expand(H * W, S)
                pts = rays_o[:, None, :] + rays_d[:, None, :] * z[..., None]
                ab = mV.audio_merge_forward(audio[3:4])
                x = torch.cat([pts.reshape(-1, 3), ab.expand(H * W * S, -1)], -1)
                raw = mV.rgb_forward(x, time_pts=torch.tensor([9])).reshape(H * W, S, 4)
                if kind == "kaiming":
                    raw_used = raw
                else:
                    raw_used = raw
                rgb, weights, depth = ns.density2outputs(raw_used, z, rays_d, 0.0, torch.device("cpu"))
                cases["vol_%s_%dx%dx%d" % (kind, H, W, S)] = dict(
                    audio=audio[3:4].numpy(), index=np.int64(9), H=np.int64(H), W=np.int64(W), S=np.int64(S),
                    focal=np.float32(focal), c2w=c2w.numpy(), rays_o=rays_o.numpy(), rays_d=rays_d.numpy(),
                    raw=raw.numpy(), rgb=rgb.reshape(H, W, 3).numpy(), weights=weights.numpy(),
                    depth=depth.numpy())

    # ---- a7 alone on hand-made raw values with strong densities (exercises transmittance decay)
    g = torch.Generator().manual_seed(21)
    raw = torch.randn(33, 24, 4, generator=g) * 3.0
    z = torch.sort(torch.rand(33, 24, generator=g), -1).values
    rd = torch.randn(33, 3, generator=g)
    rgb, weights, depth = ns.density2outputs(raw, z, rd, 0.0, torch.device("cpu"))
    cases["composite_only"] = dict(raw=raw.numpy(), z=z.numpy(), rays_d=rd.numpy(), rgb=rgb.numpy(),
                                   weights=weights.numpy(), depth=depth.numpy())

    flat = {}
    for cname, d in cases.items():
        for k, v in d.items():
            flat["%s/%s" % (cname, k)] = np.asarray(v)
    path = os.path.join(OUT, "reference_golden.npz")
    np.savez_compressed(path, **flat)
    print("wrote", path, os.path.getsize(path), "bytes;", len(cases), "cases")
    for c in sorted(cases):
        print("  ", c)


def main_postfusion():
    """tests/golden/reference_golden_postfusion.npz: the reference's post_fusion2_onlylip (tf_nerf.py:287-389,
    inference branch) on random images, a random lip mask and a smooth random warp; stores the two pre-UNet outputs."""
    ns = load_reference()
    torch.manual_seed(3)
    m = ns.TalkingFace(device=torch.device("cpu"), cfg=ns.cfg, mode="eval").eval()
    g = torch.Generator().manual_seed(42)
    cases = {}
    for name, (h, w, lh, lw, x0, y0) in {"pf_a": (40, 40, 8, 12, 14, 16), "pf_b": (36, 52, 6, 10, 20, 12)}.items():
        B = 2
        lip = torch.rand(B, lh, lw, 3, generator=g)
        face = torch.rand(B, h, w, 3, generator=g)
        gt = torch.rand(B, h, w, 3, generator=g)
        face[:, :3] = 0
        mask = torch.zeros(B, h, w, 3)
        mask[:, y0 + 1:y0 + lh - 1, x0 + 1:x0 + lw - 2, :] = 1
        mask[:, y0 + 2, x0 + 3, 1] = 0
        ys, xs = torch.meshgrid(torch.linspace(-1, 1, h), torch.linspace(-1, 1, w), indexing="ij")
        coord = torch.stack([xs, ys], -1)[None].repeat(B, 1, 1, 1)
        coord = coord * 1.08 + 0.05 * torch.randn(B, h, w, 2, generator=g) * 0.3 + 0.03
        with torch.no_grad():
            recon, fused, canon = m.post_fusion2_onlylip(lip, face, gt, mask, x0, y0, coord, use_canonical_space=True)
        cases[name] = dict(lip=lip.numpy(), face=face.numpy(), gt=gt.numpy(), mask=mask.numpy(), coord=coord.numpy(),
                           x0=np.int64(x0), y0=np.int64(y0), fused=fused.numpy(), canon=canon.numpy())
    flat = {"%s/%s" % (c, k): np.asarray(v) for c, d in cases.items() for k, v in d.items()}
    np.savez_compressed(os.path.join(OUT, "reference_golden_postfusion.npz"), **flat)


def main_staging():
    """tests/golden/reference_golden_staging.npz: (a) the windowing lines of deepspeech_features.py:65-75 run verbatim
    on random logits, (b) cv2.cvtColor + cv2.imwrite(img * 255) (inference.py:173-178) through a lossless PNG."""
    import cv2
    import tempfile
    rng = np.random.Generator(np.random.PCG64(77))
    x = rng.uniform(-0.2, 1.2, size=(37, 53, 3)).astype(np.float32)
    x[0, 0] = [0.5 / 255, 1.5 / 255, 2.5 / 255]
    x[0, 1] = [254.5 / 255, 0.49999 / 255, np.nan]
    x[0, 2] = [1.0, 0.0, -0.0]
    bgr = cv2.cvtColor(x, cv2.COLOR_RGB2BGR)
    path = os.path.join(tempfile.mkdtemp(), "t.png")
    cv2.imwrite(path, bgr * 255)
    u8 = cv2.imread(path, cv2.IMREAD_COLOR)
    logits = rng.normal(size=(23, 29)).astype(np.float32)
    net_output = logits.reshape(-1, 29)
    win_size = 16
    zero_pad = np.zeros((int(win_size / 2), net_output.shape[1]))
    net_output = np.concatenate((zero_pad, net_output, zero_pad), axis=0)
    windows = []
    for window_index in range(0, net_output.shape[0] - win_size, 2):
        windows.append(net_output[window_index:window_index + win_size])
    np.savez_compressed(os.path.join(OUT, "reference_golden_staging.npz"),
                        **{"u8/rgb": x, "u8/bgr8": u8, "win/logits": logits, "win/windows": np.array(windows).astype(np.float32)})


if __name__ == "__main__":
    main()
    main_postfusion()
    main_staging()
