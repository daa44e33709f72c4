"""CPU: the oracle (oracle/s2l_oracle.py) must reproduce the outputs of the real
reference stored in tests/golden/reference_golden.npz (the reference
has no KATs of its own, so the pin is the reference run in the build container)."""
import numpy as np
import pytest
import torch

from oracle import s2l_oracle as O
from oracle import synth

TOL = 2e-6   # fp32; differences only from BLAS blocking order ([N,..] vs broadcast shapes)


def sd(kind, uv_dims=2, och=3):
    return O.to_torch_sd(synth.make_state_dict(0, kind, uv_dims, och))


@pytest.mark.parametrize("kind", ["default", "kaiming"])
def test_audio_net(golden, kind):
    g = golden["audio_%s" % kind]
    out = O.audio_merge_forward(sd(kind), torch.from_numpy(g["audio"]))
    np.testing.assert_allclose(out.numpy(), g["latent"], atol=TOL, rtol=0)
    out2 = O.audio_merge_forward(sd(kind), torch.from_numpy(g["audio"]).permute(0, 2, 1).contiguous())
    np.testing.assert_allclose(out2.numpy(), g["latent_from_29x16"], atol=TOL, rtol=0)
    # synthetic audio generator is the one the golden was made with
    np.testing.assert_array_equal(g["audio"], synth.make_audio(4, seed=1))


def test_embedders_and_coords(golden):
    g = golden["embed"]
    np.testing.assert_array_equal(O.uv_embed(torch.from_numpy(g["uv"])).numpy(), g["pe"])
    np.testing.assert_array_equal(O.time_embed(torch.tensor([5])).numpy(), g["t5"])
    np.testing.assert_array_equal(O.time_embed(torch.tensor([6000])).numpy(), g["t6000"])
    np.testing.assert_array_equal(O.get_coords(7, 5).numpy(), g["coords_7x5"])


@pytest.mark.parametrize("kind", ["default", "kaiming"])
@pytest.mark.parametrize("shape", [(24, 32, 5), (8, 8, 6000)])
def test_plain(golden, kind, shape):
    H, W, idx = shape
    g = golden["plain_%s_%dx%d_i%d" % (kind, H, W, idx)]
    for once in (True, False):
        out = O.render_plain(sd(kind), torch.from_numpy(g["audio"]), idx, H, W, audio_once=once)
        np.testing.assert_allclose(out.numpy(), g["rgb"], atol=TOL * 10 if kind == "kaiming" else TOL, rtol=0)


@pytest.mark.parametrize("kind", ["default", "kaiming"])
def test_rowlatent(golden, kind):
    g = golden["rowlatent_%s" % kind]
    out = O.rgb_forward(sd(kind), torch.from_numpy(g["x"]), torch.tensor([int(g["index"])]))
    np.testing.assert_allclose(out.numpy(), g["out"], atol=TOL * 10 if kind == "kaiming" else TOL, rtol=0)


@pytest.mark.parametrize("kind", ["default", "kaiming"])
@pytest.mark.parametrize("seed", [11, 12])
def test_ensemble4(golden, kind, seed):
    g = golden["ens4_%s_seed%d" % (kind, seed)]
    H, W = int(g["H"]), int(g["W"])
    # RNG alignment: the reference drew eps as ry*rand(1)/2 right after manual_seed(seed)
    torch.manual_seed(seed)
    eps = (0.5 / H) * torch.rand(1) / 2.0
    np.testing.assert_array_equal(eps.numpy(), g["eps"])
    out = O.render_ensemble4(sd(kind), torch.from_numpy(g["audio"]), int(g["index"]), H, W, eps)
    np.testing.assert_allclose(out.numpy(), g["rgb"], atol=TOL * 10 if kind == "kaiming" else TOL, rtol=0)


@pytest.mark.parametrize("kind", ["default", "kaiming"])
@pytest.mark.parametrize("shape", [(8, 8, 16), (6, 10, 64)])
def test_volumetric(golden, kind, shape):
    H, W, S = shape
    g = golden["vol_%s_%dx%dx%d" % (kind, H, W, S)]
    c2w = torch.from_numpy(g["c2w"])
    ro, rd = O.get_rays(H, W, float(g["focal"]), c2w)
    np.testing.assert_array_equal(ro.reshape(-1, 3).numpy(), g["rays_o"])
    np.testing.assert_array_equal(rd.reshape(-1, 3).numpy(), g["rays_d"])
    rgb, weights, depth, raw = O.render_volumetric(sd(kind, 3, 4), torch.from_numpy(g["audio"]), int(g["index"]),
                                                   H, W, S, float(g["focal"]), c2w, return_aux=True)
    tol = TOL * 10 if kind == "kaiming" else TOL
    np.testing.assert_allclose(raw.numpy(), g["raw"], atol=tol, rtol=0)
    np.testing.assert_allclose(rgb.numpy(), g["rgb"], atol=tol, rtol=0)
    np.testing.assert_allclose(weights.numpy(), g["weights"], atol=tol, rtol=0)
    np.testing.assert_allclose(depth.numpy(), g["depth"], atol=tol, rtol=0)


def test_composite_only(golden):
    g = golden["composite_only"]
    rgb, w, d = O.density2outputs(torch.from_numpy(g["raw"]), torch.from_numpy(g["z"]), torch.from_numpy(g["rays_d"]))
    np.testing.assert_array_equal(rgb.numpy(), g["rgb"])
    np.testing.assert_array_equal(w.numpy(), g["weights"])
    np.testing.assert_array_equal(d.numpy(), g["depth"])
    # property: weights are a sub-probability along the ray
    assert (w.sum(-1) <= 1 + 1e-5).all() and (w >= 0).all()


def test_fp64_truth_close_to_fp32_reference(golden):
    """the float64 oracle is the 'truth' used to rank errors; it must sit within fp32 noise of the reference."""
    g = golden["plain_kaiming_24x32_i5"]
    sd64 = O.to_torch_sd(synth.make_state_dict(0, "kaiming"), torch.float64)
    out = O.render_plain(sd64, torch.from_numpy(g["audio"]).double(), 5, 24, 32)
    assert np.abs(out.numpy() - g["rgb"]).max() < 2e-4


@pytest.mark.parametrize("case", ["pf_a", "pf_b"])
def test_post_fusion_compose(golden, case):
    """the pre-UNet part of post_fusion2_onlylip_light (tf_nerf.py:334-386)."""
    g = golden[case]
    t = lambda k: torch.from_numpy(g[k])
    fused, canon = O.post_fusion_compose(t("lip"), t("face"), t("gt"), t("mask"), int(g["x0"]), int(g["y0"]), t("coord"))
    np.testing.assert_array_equal(canon.numpy(), g["canon"])
    np.testing.assert_allclose(fused.numpy(), g["fused"], atol=1e-7, rtol=0)
    # the drop-in module's PyTorch (CPU) branch must agree too
    import json, os
    from speech2lip_b200 import TalkingFace
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = json.load(open(os.path.join(root, "tests", "golden", "may_cfg.json")))
    m = TalkingFace(device=torch.device("cpu"), cfg=cfg, mode="eval").eval()
    with torch.no_grad():
        _, fused2, canon2 = m.post_fusion2_onlylip(t("lip"), t("face"), t("gt"), t("mask"), int(g["x0"]), int(g["y0"]), t("coord"))
    np.testing.assert_allclose(fused2.numpy(), g["fused"], atol=1e-7, rtol=0)
    np.testing.assert_array_equal(canon2.numpy(), g["canon"])


def test_staging_formats(golden):
    """the wire formats either side of the path — DeepSpeech windowing (golden = the reference's
    numpy lines run verbatim) and the uint8 BGR frame cv2.imwrite stores (golden = real cv2 round trip through PNG)."""
    w = golden["win"]
    np.testing.assert_array_equal(O.audio_windows(w["logits"]).numpy(), w["windows"])
    assert w["windows"].shape == (12, 16, 29)
    u = golden["u8"]
    np.testing.assert_array_equal(O.frames_to_bgr8(u["rgb"]).numpy(), u["bgr8"])


def test_wide_grid_case_pins_the_fused_linspace(golden):
    """get_coords at W = 256 against the REAL reference: torch.linspace's upper half is end - step*k in ONE fused
    multiply-add; the two-rounding form is 1 ulp off in ~9 % of the columns (the bug the in-kernel grid had), which this
    case would expose (the narrow golden cases cannot).  Also: the oracle reproduces the reference's wide render."""
    g = golden["grid_kaiming_40x256_i7"]
    coords = O.get_coords(256, 40).numpy()
    assert np.array_equal(coords, g["coords"])
    n = 256
    step = np.float32(1.0) / np.float32(n - 1)
    i = np.arange(n)
    k = (n - 1 - i).astype(np.float32)
    fused = np.where(i < n // 2, (step * i.astype(np.float32)).astype(np.float32), (1.0 - np.float64(step) * k).astype(np.float32))
    two = np.where(i < n // 2, fused, (np.float32(1.0) - (step * k).astype(np.float32)).astype(np.float32))
    u = g["coords"].reshape(40, 256, 2)[0, :, 0]
    assert np.array_equal(u, fused) and int((u != two).sum()) >= 10
    sd = O.to_torch_sd(synth.make_state_dict(0, "kaiming"))
    with torch.no_grad():
        rgb = O.render_plain(sd, torch.from_numpy(g["audio"]), int(g["index"]), 40, 256)
    assert np.abs(rgb.numpy() - g["rgb"]).max() < 2e-5


def test_wide_ensemble4_case_oracle(golden):
    g = golden["grid_ens4_kaiming_12x256_i9"]
    sd = O.to_torch_sd(synth.make_state_dict(0, "kaiming"))
    with torch.no_grad():
        rgb = O.render_ensemble4(sd, torch.from_numpy(g["audio"]), int(g["index"]), 12, 256, float(g["eps"][0]))
    assert np.abs(rgb.numpy() - g["rgb"]).max() < 2e-5
