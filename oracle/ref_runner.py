"""TEST / BASELINE INFRASTRUCTURE ONLY — runs the reference's OWN code (imported through oracle/ref_shim.py from
/root/reference or the vendored oracle/_ref) on the workloads bench.py and the tests use.  Nothing here is arithmetic
of this repository: every function only arranges calls into the reference.

  volumetric  TalkingFace(uv_dims=3, output_ch=4) (tf_nerf.py:13-18) + get_rays (src/common.py:12-21) +
              density2outputs (rendering.py:30-62), z = linspace(near, far, S)      — SURVEY §0.2 "Mode V"
  plain       the loop body of inference.py:144-159, verbatim (audio window tiled H*W times, as written)
  ensemble4   Trainer.predict_lip_image (training.py:158-251)
"""
import torch

from . import synth
from .ref_shim import load_reference, reference_available

_ns = None


def ns():
    global _ns
    if _ns is None:
        _ns = load_reference()
    return _ns


def available():
    return reference_available()


def model(uv_dims, output_ch, device, seed=0, kind="kaiming", mode="eval"):
    """The reference's TalkingFace with this repo's deterministic synthetic hot-path weights loaded."""
    n = ns()
    torch.manual_seed(1234)          # non-hot tensors (UNet, depth) only
    m = n.TalkingFace(device=device, cfg=n.cfg, mode=mode, uv_dims=uv_dims, output_ch=output_ch).to(device)
    m = m.eval() if mode == "eval" else m.train()
    missing = m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_state_dict(seed, kind, uv_dims, output_ch).items()}, strict=False)
    assert not missing.unexpected_keys, missing.unexpected_keys
    return m


def render_volumetric(m, audio, index, H, W, S, focal, c2w, device, n_rays=None, chunk=65536, near=0.0, far=1.0):
    """One frame (or its first n_rays rays): AudioNet once, rgb_forward in `chunk`-point calls, density2outputs."""
    n = ns()
    with torch.no_grad():
        rays_o, rays_d = n.get_rays(H, W, focal, c2w.to(device), device)
        rays_o, rays_d = rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)
        if n_rays is not None:
            rays_o, rays_d = rays_o[:n_rays], rays_d[:n_rays]
        R = rays_o.shape[0]
        z = torch.linspace(near, far, S, device=device).expand(R, S)
        lat = m.audio_merge_forward(audio.to(device))
        pts = (rays_o[:, None, :] + rays_d[:, None, :] * z[..., None]).reshape(-1, 3)
        t = torch.tensor([int(index)], device=device)
        outs = []
        for s in range(0, pts.shape[0], chunk):
            p = pts[s:s + chunk]
            outs.append(m.rgb_forward(torch.cat([p, lat.expand(p.shape[0], -1)], -1), time_pts=t))
        raw = torch.cat(outs).reshape(R, S, 4)
        rgb, _, _ = n.density2outputs(raw, z, rays_d, 0.0, device)
    return rgb


def render_plain(m, audio, index, H, W, device, n_rows=None):
    """inference.py:144-159 verbatim for one frame (optionally only its first n_rows pixel rows)."""
    n = ns()
    with torch.no_grad():
        coords = n.get_coords(W, H, device)
        if n_rows is not None:
            coords = coords[:n_rows * W]
        a = audio.to(device).tile(coords.shape[0], 1, 1)
        ab = m.audio_merge_forward(a)
        x = torch.cat([coords[:, None, :], ab[:, None, :]], -1).view(-1, m.audio_dims + 2)
        return m.rgb_forward(x, time_pts=torch.tensor([int(index)], device=device), rgb_pts=None)[:, :3]


def trainer(m, device, H, W):
    n = ns()
    tr = n.Trainer(m, None, device, "/tmp", cfg=n.cfg, batch_rays=H * W, use_audio_net=True, use_time=True, use_audio=True,
                   use_perceptual_loss=False, use_syncloss=False, multi_gpu=False)
    tr.height, tr.width = H, W            # set in train_stage1 (training.py:393-394)
    return tr


def render_ensemble4(tr, audio, index, H, W, device):
    """Trainer.predict_lip_image for one whole frame (draws eps_shift from the device RNG, training.py:200)."""
    n = ns()
    coords = n.get_coords(W, H, device)
    return tr.predict_lip_image(0, coords, audio.to(device), None, {"index": torch.tensor([int(index)], device=device)}, None, None, None)
