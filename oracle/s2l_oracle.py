"""TEST INFRASTRUCTURE ONLY — CPU restatement of the Speech2Lip rendering hot path.

This is the *oracle* the CUDA path is checked against.  It is a plain,
module-free torch-CPU restatement of the reference's arithmetic (the reference
itself is 100 % PyTorch, so torch's CPU ATen ops are the faithful host port).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import it; the product package (speech2lip_b200/) never does.

Pinning: the reference ships no tests / golden vectors for this path,
so the oracle is pinned against the reference's own code *run in the build
container* — tests/golden/make_golden.py imports /root/reference (oracle/ref_shim.py),
loads the same synthetic weights, and stores the reference's outputs under
tests/golden/*.npz; tests/test_oracle_vs_golden.py then requires this file to
reproduce them (bit-exact for everything except BLAS-order effects, tolerance 2e-6).

Every function cites the reference lines it restates (paths relative to
/root/reference).  `sd` is a dict name -> tensor with the reference's state_dict
names; `dtype=torch.float64` gives a higher-precision "truth" used to rank
errors of the fp32 reference vs the CUDA kernels.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


def to_torch_sd(sd_np, dtype=torch.float32):
    return {k: torch.as_tensor(np.asarray(v)).to(dtype) for k, v in sd_np.items()}


# --------------------------------------------------------------------------- AudioNet
def audio_merge_forward(sd, audio):
    """src/face_simple/models/tf_nerf.py:197-213 with encoder_conv (:91-104) and
    encoder_fc1 (:105-109).  audio: [B,16,29] (permuted) or [B,29,16]."""
    if audio.shape[2] == 16:
        x = audio                                  # tf_nerf.py:203-204
    else:
        x = audio.permute(0, 2, 1)                 # tf_nerf.py:207
    for i in (0, 2, 4, 6):                         # 4 x Conv1d(k3,s2,p1) + LeakyReLU(0.02)
        x = F.conv1d(x, sd["encoder_conv.%d.weight" % i], sd["encoder_conv.%d.bias" % i],
                     stride=2, padding=1)
        x = F.leaky_relu(x, 0.02)
    x = x.squeeze(-1)                              # tf_nerf.py:209
    x = F.linear(x, sd["encoder_fc1.0.weight"], sd["encoder_fc1.0.bias"])
    x = F.leaky_relu(x, 0.02)
    x = F.linear(x, sd["encoder_fc1.2.weight"], sd["encoder_fc1.2.bias"])
    return x


# --------------------------------------------------------------------------- embedders
def uv_embed(x, multires=10):
    """Embedder.__call__, tf_nerf.py:404-425: [x, sin(f0 x), cos(f0 x), sin(f1 x), ...]
    with f = 2**linspace(0, multires-1, multires)."""
    freq_bands = 2.0 ** torch.linspace(0.0, multires - 1, steps=multires)
    outs = [x]
    for freq in freq_bands:
        for fn in (torch.sin, torch.cos):
            outs.append(fn(x * freq.to(x.dtype)))
    return torch.cat(outs, -1)


def time_div_term(out_dims=20):
    """PositionalEncodingTime.__init__, tf_nerf.py:431-432 (always float32)."""
    return torch.exp(torch.arange(0, out_dims, 2, dtype=torch.float) * -(math.log(10000.0) / out_dims))


def time_embed(position, out_dims=20, dtype=torch.float32):
    """PositionalEncodingTime.__call__, tf_nerf.py:434-442: uses position[0] only and
    returns ONE 1-D [out_dims] vector that broadcasts over all points."""
    div = time_div_term(out_dims)
    pe = torch.zeros(out_dims)
    p0 = position.reshape(-1)[0].float()
    pe[0::2] = torch.sin(p0 * div)
    pe[1::2] = torch.cos(p0 * div)
    return pe.to(dtype)


# --------------------------------------------------------------------------- implicit MLP
def rgb_forward(sd, uv_audio_pts, time_pts, uv_dims=2):
    """TalkingFace.rgb_forward (MLP v2), tf_nerf.py:225-285."""
    dtype = uv_audio_pts.dtype
    uv = uv_audio_pts[:, :uv_dims]
    a = uv_audio_pts[:, uv_dims:]
    e = uv_embed(uv)                                                   # :240
    t = time_embed(time_pts, dtype=dtype)                              # :247
    net = F.linear(e, sd["fc_uv.weight"], sd["fc_uv.bias"])            # :252
    net = net + F.linear(a, sd["fc_audio.weight"], sd["fc_audio.bias"])    # :254
    net = net + F.linear(t, sd["fc_time.weight"], sd["fc_time.bias"])      # :258
    h = net
    for i in range(8):                                                 # :265-281
        h = F.relu(F.linear(h, sd["pts_linears.%d.weight" % i], sd["pts_linears.%d.bias" % i]))
        if i == 4:
            hs = F.linear(e, sd["fc_uv_skip.weight"], sd["fc_uv_skip.bias"])
            hs = hs + F.linear(a, sd["fc_audio_skip.weight"], sd["fc_audio_skip.bias"])
            hs = hs + F.linear(t, sd["fc_time_skip.weight"], sd["fc_time_skip.bias"])
            h = torch.cat([hs, h], -1)
    return F.linear(h, sd["output_linear.weight"], sd["output_linear.bias"])   # :283 (raw, no sigmoid)


# --------------------------------------------------------------------------- render helpers
def get_coords(width, height, dtype=torch.float32):
    """src/face_simple/rendering.py:9-28 (no-noise branch): (u,v)=(x,y), inclusive linspace."""
    x = torch.linspace(0.0, 1.0, width)
    y = torch.linspace(0.0, 1.0, height)
    v, u = torch.meshgrid(y, x, indexing="ij")
    return torch.stack([u, v], -1).view(-1, 2).to(dtype)


def get_rays(H, W, focal, c2w):
    """src/common.py:12-21."""
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W), torch.linspace(0, H - 1, H), indexing="ij")
    i = i.t()
    j = j.t()
    dirs = torch.stack([(i - W * .5) / focal, -(j - H * .5) / focal, -torch.ones_like(i)], -1)
    rays_d = torch.sum(dirs[..., None, :] * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def density2outputs(raw, z_vals, rays_d):
    """src/face_simple/rendering.py:30-62 with raw_noise_std = 0."""
    dists = z_vals[..., 1:] - z_vals[..., :-1]
    dists = torch.cat([dists, torch.tensor([1e10], dtype=raw.dtype).expand(dists[..., :1].shape)], -1)
    dists = dists * torch.norm(rays_d[..., None, :], dim=-1)
    rgb = torch.sigmoid(raw[..., :3])
    alpha = 1. - torch.exp(-F.relu(raw[..., 3]) * dists)
    weights = alpha * torch.cumprod(
        torch.cat([torch.ones((alpha.shape[0], 1), dtype=raw.dtype), 1. - alpha + 1e-10], -1), -1)[:, :-1]
    rgb_map = torch.sum(weights[..., None] * rgb, -2)
    depth_map = torch.sum(weights * z_vals, -1)
    return rgb_map, weights, depth_map


# --------------------------------------------------------------------------- the three render modes
def render_plain(sd, audio, index, H, W, audio_once=True):
    """One frame, 1 MLP eval / pixel: inference.py:144-159 restated.
    audio: [1,16,29]; index: int.  audio_once=False reproduces the reference's tiling
    of the *input* window to H*W copies (inference.py:144) — same numbers, N x the work."""
    dtype = audio.dtype
    n = H * W
    coords = get_coords(W, H, dtype)
    if audio_once:
        lat = audio_merge_forward(sd, audio).expand(n, -1)
    else:
        lat = audio_merge_forward(sd, audio.tile(n, 1, 1))
    x = torch.cat([coords, lat], -1)
    out = rgb_forward(sd, x, torch.tensor([index]), uv_dims=2)
    return out[:, :3].reshape(H, W, 3)


def render_ensemble4(sd, audio, index, H, W, eps_shift):
    """One frame through the 4-tap local ensemble: Trainer.predict_lip_image,
    src/face_simple/training.py:158-251, with eps_shift passed explicitly (the
    reference draws it as ry*rand(1)/2, :200)."""
    dtype = audio.dtype
    n = H * W
    coords = get_coords(W, H, dtype)
    lat = audio_merge_forward(sd, audio).unsqueeze(1).tile(1, n, 1).view(-1, lat_dim(sd))   # :171
    rx = 0.5 / W
    ry = 0.5 / H
    eps = torch.as_tensor(eps_shift, dtype=dtype).reshape(1)
    preds, areas = [], []
    for vx in (-1, 1):
        for vy in (-1, 1):
            c = coords.clone()
            c[:, 0] += vx * rx + eps
            c[:, 1] += vy * ry + eps
            c.clamp_(0, 1)
            x = torch.cat([c, lat], -1)
            preds.append(rgb_forward(sd, x, torch.tensor([index]), uv_dims=2))
            area = torch.abs((c[:, 0] - coords[:, 0]) * (c[:, 1] - coords[:, 1]))
            areas.append(area + 1e-9)
    tot = torch.stack(areas).sum(dim=0)
    areas[0], areas[3] = areas[3], areas[0]        # :244
    areas[1], areas[2] = areas[2], areas[1]        # :245
    ret = 0
    for p, a in zip(preds, areas):
        ret = ret + p * (a / tot).unsqueeze(-1)
    return ret[:, :3].reshape(H, W, 3)


def lat_dim(sd):
    return sd["encoder_fc1.2.weight"].shape[0]


def z_samples(S, near=0.0, far=1.0, dtype=torch.float32):
    """Sample placement for the volumetric mode.  The reference has config keys
    (lindisp/perturb, src/face_simple/config.py:36-37) but no code; SURVEY §8(d)
    fixes z = linspace(near, far, S), no perturbation."""
    t = torch.linspace(0.0, 1.0, S)
    return (near * (1. - t) + far * t).to(dtype)


def render_volumetric(sd, audio, index, H, W, S, focal, c2w, near=0.0, far=1.0, chunk=65536,
                      return_aux=False):
    """Mode V assembled only from reference code: TalkingFace(uv_dims=3,
    output_ch=4).rgb_forward on pts = o + d*z, then density2outputs."""
    dtype = audio.dtype
    rays_o, rays_d = get_rays(H, W, focal, c2w)
    rays_o = rays_o.reshape(-1, 3).to(dtype)
    rays_d = rays_d.reshape(-1, 3).to(dtype)
    R = rays_o.shape[0]
    z = z_samples(S, near, far, dtype).expand(R, S)
    pts = rays_o[:, None, :] + rays_d[:, None, :] * z[..., None]          # [R,S,3]
    lat = audio_merge_forward(sd, audio)                                   # [1,64]
    flat = pts.reshape(-1, 3)
    outs = []
    for s in range(0, flat.shape[0], chunk):
        p = flat[s:s + chunk]
        x = torch.cat([p, lat.expand(p.shape[0], -1)], -1)
        outs.append(rgb_forward(sd, x, torch.tensor([index]), uv_dims=3))
    raw = torch.cat(outs, 0).reshape(R, S, 4)
    rgb, weights, depth = density2outputs(raw, z, rays_d)
    if return_aux:
        return rgb.reshape(H, W, 3), weights, depth, raw
    return rgb.reshape(H, W, 3)


def post_fusion_compose(rgb_lip, rgb_face_canonical, rgb_gt, mask_lip_canonical, lip_lefttop_x, lip_lefttop_y, coord,
                        data_path="dataset/may_face_crop_lip", expand_lip_mask=True):
    """Pre-UNet part of TalkingFace.post_fusion2_onlylip_light, tf_nerf.py:334-386 (inference branch: no
    black-hole augmentation).  Returns (rgb_merged_new [B,H,W,3], rgb_merged_canonical [B,h,w,3])."""
    h, w = rgb_face_canonical.shape[1:3]
    lip_h, lip_w = rgb_lip.shape[1:3]
    left = lip_lefttop_x - 1
    right = w - (left + lip_w)
    up = lip_lefttop_y - 1
    down = h - (up + lip_h)
    if any(k in data_path for k in ("macron", "obama_adnerf", "obama2_face_crop", "may")):      # :345-348
        pad = (left + 1, right - 1, up + 1, down - 1)
    else:
        pad = (left, right, up, down)
    lip_pad = F.pad(rgb_lip.permute(0, 3, 1, 2), pad=pad, mode="constant", value=0).permute(0, 2, 3, 1)
    canon = mask_lip_canonical * lip_pad + (1 - mask_lip_canonical) * rgb_face_canonical            # :352
    mask = mask_lip_canonical
    if expand_lip_mask:                                                                             # :354-363
        p = lip_w // 12 if "obama2_face_crop" in data_path else lip_w // 5
        tmp = torch.zeros_like(mask_lip_canonical)
        tmp[:, lip_lefttop_y - p:lip_lefttop_y + lip_h + 2 * p, lip_lefttop_x - p:lip_lefttop_x + lip_w + p, :] = 1
        mask = torch.ones_like(mask_lip_canonical) * tmp
    merged = F.grid_sample(canon.permute(0, 3, 1, 2), coord, align_corners=False)                   # :365
    m = F.grid_sample(mask.float().permute(0, 3, 1, 2), coord, align_corners=False)
    m[m != 0] = 1
    m = m.int()
    out = m * merged + (1 - m) * rgb_gt.permute(0, 3, 1, 2)                                          # :386
    return out.permute(0, 2, 3, 1), canon


def audio_windows(logits):
    """preprocess/deepspeech_features/deepspeech_features.py:65-75: zero-pad win_size/2 rows on both sides, then
    16-row windows with stride 2.  logits [T,29] (numpy or tensor) -> float32 tensor [ceil(T/2),16,29]."""
    x = np.asarray(logits, dtype=np.float64).reshape(-1, 29)
    win_size = 16
    zero_pad = np.zeros((int(win_size / 2), x.shape[1]))
    x = np.concatenate((zero_pad, x, zero_pad), axis=0)
    windows = [x[i:i + win_size] for i in range(0, x.shape[0] - win_size, 2)]
    return torch.from_numpy(np.array(windows).astype(np.float32).reshape(-1, 16, 29))


def frames_to_bgr8(rgb):
    """inference.py:173-178: cv2.cvtColor(img, COLOR_RGB2BGR); cv2.imwrite(path, img * 255).  imwrite converts the
    float image with Mat::convertTo(CV_8U) = saturate_cast<uchar>(cvRound(v)): round-half-even, clamp, NaN -> 0."""
    v = np.asarray(rgb, dtype=np.float32) * np.float32(255.0)
    r = np.rint(v)
    r = np.where(np.isnan(r), 0.0, r)
    return torch.from_numpy(np.clip(r, 0, 255).astype(np.uint8)[..., ::-1].copy())


def psnr(a, b, peak=1.0):
    mse = torch.mean((a.double() - b.double()) ** 2).item()
    if mse == 0:
        return float("inf")
    return 10.0 * math.log10(peak * peak / mse)
