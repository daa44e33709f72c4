"""GPU tests of the volumetric (Mode V) path at and around the headline configuration (BASELINE.json configs[1]):
the fused compositing epilogue, the fp32 re-evaluation of near-zero last-sample densities, early ray termination,
and the 64-frame soak the headline's parity claim rests on.

Why the re-evaluation exists: density2outputs (rendering.py:43-52) gives the LAST sample delta = 1e10, so
alpha_last = 1 - exp(-relu(sigma_last) * 1e10 * |d|) is a step function of sign(sigma_last): a tensor-core rounding error of
1e-4 on a sigma_last that close to zero moves the pixel by T_last * sigmoid(c) (0.2-0.4), far beyond the 1e-3 bar.  The
tensor-core kernels therefore list every ray with |sigma_last| < fix_thr (2e-3) and the exact fp32 kernel redoes that
one sample, so the sign every pixel sees is the fp32 path's.
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import s2l_oracle as O
from oracle import synth

pytestmark = pytest.mark.gpu
TOL = 1e-3                      # north_star: <= 1e-3 max-abs fp32 per pixel


@pytest.fixture(scope="module")
def S():
    import speech2lip_b200 as s2l
    assert torch.cuda.is_available() and os.path.exists(s2l.LIB_PATH)
    return s2l


def dev():
    return torch.device("cuda:0")


_cache = {}


def packed(S, kind="kaiming", seed=0):
    key = (kind, seed)
    if key not in _cache:
        sd = {k: torch.from_numpy(v).to(dev()) for k, v in synth.make_state_dict(seed, kind, 3, 4).items()}
        _cache[key] = S.PackedWeights(sd, 3, 4)
    return _cache[key]


def rays(H, W, focal):
    ro, rd = O.get_rays(H, W, focal, torch.eye(4)[:3])
    return ro.reshape(-1, 3), rd.reshape(-1, 3)


def exact_raw(S, w, audio_d, idx, ro_d, rd_d, z_d):
    """raw [F,R,S,4] from the exact fp32 kernel on explicit points (the same o + d*z roundings as the kernels' gen_point)"""
    from speech2lip_b200 import renderer as R
    _, bias = R.audio_encode(w, audio_d, idx, want_latent=False)
    pts = (ro_d[:, None, :] + rd_d[:, None, :] * z_d[None, :, None]).reshape(1, -1, 3).expand(audio_d.shape[0], -1, -1).contiguous()
    return R.mlp_points(w, bias, pts, "fp32").view(audio_d.shape[0], ro_d.shape[0], z_d.shape[0], 4)


def composite_exact(S, raw, z_d, rd_d):
    from speech2lip_b200 import renderer as R
    F, Rn, Sn, _ = raw.shape
    rgb, _, _ = R.density2outputs(raw.view(F * Rn, Sn, 4), z_d, rd_d.repeat(F, 1))
    return rgb.view(F, Rn, 3)


def oracle_rays(sdv, audio, index, ro, rd, z, sel):
    """oracle (torch CPU fp32, reference arithmetic) on the selected rays of ONE frame -> rgb [n,3], sigma_last [n],
    and the two outcomes a fp32-level perturbation of a near-zero sigma_last can produce (alpha_last = 0 / 1)."""
    n, Sn = sel.numel(), z.numel()
    pts = ro[sel][:, None, :] + rd[sel][:, None, :] * z[None, :, None]
    lat = O.audio_merge_forward(sdv, audio)
    x = torch.cat([pts.reshape(-1, 3), lat.expand(n * Sn, -1)], -1)
    raw = O.rgb_forward(sdv, x, torch.tensor([int(index)]), uv_dims=3).reshape(n, Sn, 4)
    rgb, weights, _ = O.density2outputs(raw, z.expand(n, Sn), rd[sel])
    # transmittance in front of the last sample (rendering.py:52) and the pixel without / with a fully opaque last sample
    dists = (z[1:] - z[:-1])[None, :] * rd[sel].norm(dim=-1, keepdim=True)
    alpha = 1. - torch.exp(-torch.relu(raw[:, :-1, 3]) * dists)
    T_last = torch.prod(1. - alpha + 1e-10, -1)
    c_last = torch.sigmoid(raw[:, -1, :3])
    base = rgb - weights[:, -1:] * c_last
    return rgb, raw[:, -1, 3], base, base + T_last[:, None] * c_last


# ------------------------------------------------------------------------------------------ fused epilogue
@pytest.mark.parametrize("precision", ["bf16x3", "fp16f8"])
@pytest.mark.parametrize("Sn", [4, 8, 16, 32, 64, 128])
def test_fused_compositing_equals_unfused(S, precision, Sn):
    """The reducer warp's compositing (fused epilogue, no raw tensor) against the same kernel's raw outputs composited by
    composite_kernel, re-evaluation off in both: the summation order and the reducer's hardware-approximate exp / reciprocal
    (2^-21 relative) differ (<= 1e-5).  Ragged ray counts, several
    frames, both schedules' tail handling (H*W*S not a multiple of 128)."""
    H, W, F = 9, 13, 3
    w = packed(S)
    ro, rd = rays(H, W, 30.0)
    z = O.z_samples(Sn)
    audio = torch.from_numpy(synth.make_audio(F, seed=21)).to(dev())
    idx = torch.tensor([4, 9, 2])
    r = S.LipRenderer(w, precision)
    kw = dict(mode="volumetric", rays_o=ro.to(dev()), rays_d=rd.to(dev()), z_vals=z.to(dev()), fix_thr=-1.0)
    fused = r.render_frames(audio, idx, H, W, **kw)
    unfused, weights, depth = r.render_frames(audio, idx, H, W, return_aux=True, **kw)
    e = (fused - unfused).abs().max().item()
    print("fused vs unfused S=%d %s: %.3e" % (Sn, precision, e))
    assert e < 1e-5
    # per-ray z / per-frame rays forms hit the same fused code path bit for bit
    again = r.render_frames(audio, idx, H, W, mode="volumetric", rays_o=ro.repeat(F, 1).to(dev()), rays_d=rd.repeat(F, 1).to(dev()),
                            z_vals=z.expand(F * H * W, Sn).contiguous().to(dev()), fix_thr=-1.0)
    assert torch.equal(again, fused)


@pytest.mark.parametrize("precision", ["bf16x3", "fp16f8"])
def test_fused_blend_equals_unfused(S, precision):
    """4-tap blend in the reducer warp == raw outputs + ensemble4_blend_kernel (same arithmetic, bit-exact)."""
    from speech2lip_b200 import _cabi, renderer as R
    H, W, F = 11, 19, 2
    sd = {k: torch.from_numpy(v).to(dev()) for k, v in synth.make_state_dict(0, "kaiming", 2, 3).items()}
    w = S.PackedWeights(sd, 2, 3)
    audio = torch.from_numpy(synth.make_audio(F, seed=22)).to(dev())
    idx = torch.tensor([1, 7])
    eps = torch.tensor([0.003, 0.011])
    fused = S.LipRenderer(w, precision).render_frames(audio, idx, H, W, mode="ensemble4", eps_shift=eps)
    _, bias = R.audio_encode(w, audio, idx, want_latent=False)
    eps_d = eps.to(dev())
    g = _cabi.S2LGeom(n_frames=F, height=H, width=W, n_samples=0, pts_mode=_cabi.PTS_GRID_ENS4, uv_dims=2, out_ch=3, z_per_ray=0,
                      rays_per_frame_shared=0, pts_per_frame=0, eps_shift=0.0, eps_per_frame=eps_d.data_ptr())
    raw = torch.empty(F, H * W * 4, 3, device=dev())
    rgb = torch.empty(F, H, W, 3, device=dev())
    lib = _cabi.lib()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _cabi.check(lib.s2l_mlp_fwd(R._ptr(w.blob), C.byref(g), R._ptr(bias), None, None, None, None, R._ptr(raw), _cabi.PRECISIONS[precision], st), "mlp")
    _cabi.check(lib.s2l_ensemble4_blend(R._ptr(raw), C.byref(g), R._ptr(rgb), st), "blend")
    assert torch.equal(fused, rgb)


# ------------------------------------------------------------------------------------------ last-sample re-evaluation
@pytest.mark.parametrize("precision", ["bf16x3", "fp16f8"])
@pytest.mark.parametrize("fused", [True, False])
def test_last_sample_reevaluation_removes_flips(S, precision, fused):
    """A geometry chosen so that MANY rays have a last-sample density near zero (low focal: the last samples fan out over a
    wide volume).  Without the re-evaluation some pixels flip against the exact path (reported); with it none may, in the
    fused path and in the unfused one (weights/depth requested)."""
    H, W, Sn, F = 96, 96, 32, 4
    w = packed(S)
    ro, rd = rays(H, W, 60.0)
    z = O.z_samples(Sn)
    ro_d, rd_d, z_d = ro.to(dev()), rd.to(dev()), z.to(dev())
    audio = torch.from_numpy(synth.make_audio(F, seed=31)).to(dev())
    idx = torch.arange(F) + 3
    raw32 = exact_raw(S, w, audio, idx, ro_d, rd_d, z_d)
    want = composite_exact(S, raw32, z_d, rd_d).view(F, H, W, 3)
    r = S.LipRenderer(w, precision)
    kw = dict(mode="volumetric", rays_o=ro_d, rays_d=rd_d, z_vals=z_d)

    def run(**extra):
        out = r.render_frames(audio, idx, H, W, return_aux=not fused, **kw, **extra)
        return out if fused else out[0]
    off = run(fix_thr=-1.0)
    on = run()
    n_listed = int(r.last_render_counts()["reevaluated"].sum())
    bad_off = int(((off - want).abs().amax(-1) > TOL).sum())
    bad_on = int(((on - want).abs().amax(-1) > TOL).sum())
    near = int((raw32[..., -1, 3].abs() < 2e-3).sum())
    print("%s fused=%s: rays with |sigma_last|<2e-3: %d (exact) / %d (listed); pixels > 1e-3 vs exact: %d without, %d with re-evaluation; max %.2e"
          % (precision, fused, near, n_listed, bad_off, bad_on, (on - want).abs().max().item()))
    assert bad_on == 0
    assert abs(n_listed - near) <= max(4, near // 2)
    # a wider threshold re-evaluates more rays and cannot make anything worse
    on2 = run(fix_thr=2e-2)
    assert int(r.last_render_counts()["reevaluated"].sum()) > n_listed
    assert int(((on2 - want).abs().amax(-1) > TOL).sum()) == 0


# ------------------------------------------------------------------------------------------ early ray termination
@pytest.mark.parametrize("precision", ["bf16x3", "fp16f8"])
@pytest.mark.parametrize("chunks", [2, 4, 8])
def test_sample_chunks_without_termination_equal_one_launch(S, precision, chunks):
    """C front-to-back launches with the (T, acc) carry and compacted ray lists, termination off: every ray survives every
    chunk, so the result must equal the single-launch render up to the summation order."""
    H, W, Sn, F = 17, 23, 64, 3
    w = packed(S)
    ro, rd = rays(H, W, 50.0)
    audio = torch.from_numpy(synth.make_audio(F, seed=41)).to(dev())
    idx = torch.tensor([5, 0, 8])
    r = S.LipRenderer(w, precision)
    kw = dict(mode="volumetric", rays_o=ro.to(dev()), rays_d=rd.to(dev()), z_vals=O.z_samples(Sn).to(dev()))
    one = r.render_frames(audio, idx, H, W, **kw)
    many = r.render_frames(audio, idx, H, W, sample_chunks=chunks, term_thr=0.0, **kw)
    cnt = r.last_render_counts()
    assert (cnt["alive"] == H * W).all()
    e = (one - many).abs().max().item()
    print("chunks=%d %s: %.3e" % (chunks, precision, e))
    assert e < 1e-5


@pytest.mark.parametrize("scale", [1.0, 30.0, 300.0])
def test_early_ray_termination_vs_oracle(S, scale):
    """Early ray termination on scenes of increasing density (sigma row of output_linear scaled): results stay within
    term_thr of the all-samples render and within the parity bar of the oracle; the denser the scene the fewer rays survive."""
    H, W, Sn, F = 24, 32, 64, 2
    sd_np = synth.make_state_dict(0, "kaiming", 3, 4)
    sd_np["output_linear.weight"] = sd_np["output_linear.weight"].copy()
    sd_np["output_linear.bias"] = sd_np["output_linear.bias"].copy()
    sd_np["output_linear.weight"][3] *= scale
    sd_np["output_linear.bias"][3] *= scale
    w = S.PackedWeights({k: torch.from_numpy(v).to(dev()) for k, v in sd_np.items()}, 3, 4)
    sdv = O.to_torch_sd(sd_np)
    c2w = torch.eye(4)[:3]
    ro, rd = rays(H, W, 50.0)
    audio = torch.from_numpy(synth.make_audio(F, seed=42))
    idx = torch.tensor([2, 6])
    r = S.LipRenderer(w, "bf16x3")
    kw = dict(mode="volumetric", rays_o=ro.to(dev()), rays_d=rd.to(dev()), z_vals=O.z_samples(Sn).to(dev()))
    full = r.render_frames(audio.to(dev()), idx, H, W, **kw)
    ert = r.render_frames(audio.to(dev()), idx, H, W, sample_chunks=4, term_thr=1e-4, **kw)
    alive = r.last_render_counts()["alive"]
    d = (full - ert).abs().max().item()
    want = torch.stack([O.render_volumetric(sdv, audio[i:i + 1], int(idx[i]), H, W, Sn, 50.0, c2w) for i in range(F)])
    e = (ert.cpu() - want).abs().max().item()
    print("density x%g: alive per chunk %s of %d rays, ert-vs-full %.2e, ert-vs-oracle %.2e" % (scale, alive.sum(1).tolist(), F * H * W, d, e))
    assert d < 1.5e-4          # term_thr (+ rounding)
    # sigma-scaled weights amplify the tensor-core density error by `scale`; the pixel bar vs the oracle is asserted at scale 1
    if scale == 1.0:
        assert e < TOL
    if scale >= 300.0:
        assert int(alive[-1].sum()) < F * H * W // 2


# ------------------------------------------------------------------------------------------ the soak
@pytest.mark.parametrize("precision", ["fp16f8", "bf16x3"])
def test_modev_soak_64_frames_256x256x64(S, precision):
    """64 frames of the benched geometry (256x256 rays, focal 1200, 64 samples, kaiming weights, 8 frames per launch:
    the CTA-pair schedule, fused compositing, re-evaluation on) against
      (a) the exact fp32 path over ALL rays of all frames, and
      (b) the oracle on 256 rays per frame: the 128 rays with the smallest |sigma_last| (the flip candidates) + 128 random.
    Pixels > 1e-3 are counted.  (a) must be zero.  (b) must be zero on every ray whose oracle |sigma_last| >= 1e-4; below that
    the sign of sigma_last is not determined at fp32 precision (this repo's exact kernel, the reference on cuBLAS and the
    reference on CPU differ by up to ~5e-5 there), so those rays must match ONE of the two outcomes and are reported."""
    H = W = 256
    Sn, FB, NB = 64, 8, 8
    # the driver's run uses the defaults; tools/soak_modev_seeds.sh sweeps weight seeds / kinds / audio seeds through these
    wseed, wkind = int(os.environ.get("S2L_SOAK_SEED", "0")), os.environ.get("S2L_SOAK_KIND", "kaiming")
    aseed = int(os.environ.get("S2L_SOAK_AUDIO_SEED", "51"))
    w = packed(S, wkind, wseed)
    sdv = O.to_torch_sd(synth.make_state_dict(wseed, wkind, 3, 4))
    ro, rd = rays(H, W, 1200.0)
    z = O.z_samples(Sn)
    ro_d, rd_d, z_d = ro.to(dev()), rd.to(dev()), z.to(dev())
    r = S.LipRenderer(w, precision)
    rn = S.LipRenderer(w, precision)
    audio_all = torch.from_numpy(synth.make_audio(FB * NB, seed=aseed))
    tot = dict(weights="%s/%d" % (wkind, wseed), audio_seed=aseed, frames=0, bad_exact=0, bad_exact_nofix=0, listed=0, bad_oracle=0, ill=0, ill_flipped=0, max_exact=0.0, max_oracle=0.0)
    gsel = torch.Generator().manual_seed(0)
    for b in range(NB):
        audio = audio_all[b * FB:(b + 1) * FB]
        idx = torch.arange(FB) + b * FB
        a_d = audio.to(dev())
        got = r.render_frames(a_d, idx, H, W, mode="volumetric", rays_o=ro_d, rays_d=rd_d, z_vals=z_d).view(FB, H * W, 3)
        tot["listed"] += int(r.last_render_counts()["reevaluated"].sum())
        nofix = rn.render_frames(a_d, idx, H, W, mode="volumetric", rays_o=ro_d, rays_d=rd_d, z_vals=z_d, fix_thr=-1.0).view(FB, H * W, 3)
        for f in range(FB):
            raw32 = exact_raw(S, w, a_d[f:f + 1], idx[f:f + 1], ro_d, rd_d, z_d)
            want = composite_exact(S, raw32, z_d, rd_d)[0]
            err = (got[f] - want).abs().amax(-1)
            tot["bad_exact"] += int((err > TOL).sum())
            tot["bad_exact_nofix"] += int(((nofix[f] - want).abs().amax(-1) > TOL).sum())
            tot["max_exact"] = max(tot["max_exact"], float(err.max()))
            sig_last = raw32[0, :, -1, 3].abs()
            sel = torch.cat([torch.topk(sig_last, 128, largest=False).indices.cpu(),
                             torch.randint(0, H * W, (128,), generator=gsel)])
            o_rgb, o_sig, o_a0, o_a1 = oracle_rays(sdv, audio[f:f + 1], idx[f], ro, rd, z, sel)
            g = got[f][sel.to(dev())].cpu()
            e = (g - o_rgb).abs().amax(-1)
            ill = o_sig.abs() < 1e-4
            e_alt = torch.minimum((g - o_a0).abs().amax(-1), (g - o_a1).abs().amax(-1))
            tot["bad_oracle"] += int((e[~ill] > TOL).sum()) + int((e_alt[ill] > TOL).sum())
            tot["ill"] += int(ill.sum())
            tot["ill_flipped"] += int((e[ill] > TOL).sum())
            tot["max_oracle"] = max(tot["max_oracle"], float(e[~ill].max()))
            tot["frames"] += 1
    print("MODE-V SOAK %s: %s" % (precision, tot))
    assert tot["frames"] == 64
    assert tot["bad_exact"] == 0, tot
    assert tot["bad_oracle"] == 0, tot
