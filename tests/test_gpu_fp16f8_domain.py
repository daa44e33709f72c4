"""GPU tests of the fp16f8 domain guard (include/speech2lip_b200.h, S2L_PREC_FP16F8): the arithmetic is validated for
tensor-core weights |w| < 1024 and hidden activations |a| < 4096; outside it the scaled fp8 residuals saturate silently, so
the host layer must notice (pack-time weight record, exact-path activation probe) and fall back to bf16x3 LOUDLY."""
import os
import warnings

import numpy as np
import pytest
import torch

from oracle import s2l_oracle as O
from oracle import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import speech2lip_b200 as s2l
    assert torch.cuda.is_available() and os.path.exists(s2l.LIB_PATH)
    return s2l


def dev():
    return torch.device("cuda:0")


def pack(S, sd_np, uvd=2, och=3):
    return S.PackedWeights({k: torch.from_numpy(np.ascontiguousarray(v)).to(dev()) for k, v in sd_np.items()}, uvd, och)


def test_in_domain_models_use_fp16f8_and_record_their_ranges(S):
    for kind in ("default", "kaiming", "trained"):
        w = pack(S, synth.make_state_dict(0, kind, 2, 3))
        m = w.meta()
        print("meta %s: %s" % (kind, m))
        assert 0 < m["max_abs_weight"] < 16 and m["n_saturating_weights"] == 0 and w.fp16f8_weights_ok()
        r = S.LipRenderer(w, "auto")
        with warnings.catch_warnings():
            warnings.simplefilter("error")
            assert r._resolve_precision("auto") == "fp16f8" and r._resolve_precision("fp16f8") == "fp16f8"
    wv = pack(S, synth.make_state_dict(0, "kaiming", 3, 4), 3, 4)
    mv = wv.meta()
    sig = synth.make_state_dict(0, "kaiming", 3, 4)["output_linear.weight"][3]
    assert abs(mv["density_row_norm"] - float(np.linalg.norm(sig))) < 1e-4
    assert abs(mv["auto_fix_thr"] - 2e-3 * max(1.0, float(np.linalg.norm(sig)) / 2 ** 0.5)) < 1e-7


@pytest.mark.parametrize("case", ["one huge weight", "layer x4096", "heavy tailed"])
def test_weights_outside_the_domain_fall_back_loudly(S, case):
    sd = {k: v.copy() for k, v in synth.make_state_dict(0, "kaiming", 2, 3).items()}
    if case == "one huge weight":
        sd["pts_linears.3.weight"][17, 5] = 3000.0
    elif case == "layer x4096":
        sd["pts_linears.6.weight"] *= 4096.0
        sd["pts_linears.7.weight"] /= 4096.0
    else:
        g = np.random.Generator(np.random.PCG64(5))
        t = g.standard_t(1.2, size=sd["pts_linears.2.weight"].shape).astype(np.float32)      # Cauchy-like tails
        sd["pts_linears.2.weight"] = (t * 0.05).astype(np.float32)
        sd["pts_linears.2.weight"][np.abs(sd["pts_linears.2.weight"]) > 5000] = 5000.0
        assert np.abs(sd["pts_linears.2.weight"]).max() >= 1024
    w = pack(S, sd)
    assert not w.fp16f8_weights_ok() and w.meta()["n_saturating_weights"] >= 1
    audio = torch.from_numpy(synth.make_audio(1, seed=2)).to(dev())
    r = S.LipRenderer(w, "fp16f8")
    with pytest.warns(RuntimeWarning, match="outside the validated fp16f8 domain"):
        got = r.render_frames(audio, torch.tensor([3]), 16, 24)
    want = S.LipRenderer(w, "bf16x3").render_frames(audio, torch.tensor([3]), 16, 24)
    assert torch.equal(got, want)                              # it really ran bf16x3
    exact = S.LipRenderer(w, "fp32").render_frames(audio, torch.tensor([3]), 16, 24)
    rel = ((got - exact).abs().max() / exact.abs().max().clamp_min(1e-6)).item()
    print("%s: fell back to bf16x3, relative error vs exact %.2e (meta %s)" % (case, rel, w.meta()))
    assert rel < 2e-4


def test_activation_probe_catches_growing_activations(S):
    """hidden weights x4: every weight is still far inside |w| < 1024, but the activations grow 4x per layer and leave
    |a| < 4096 after a few layers — only the exact-path probe can see that."""
    sd = {k: (v * 4.0 if k.startswith("pts_linears") and k.endswith("weight") else v).astype(np.float32)
          for k, v in synth.make_state_dict(0, "kaiming", 2, 3).items()}
    w = pack(S, sd)
    assert w.fp16f8_weights_ok()
    r = S.LipRenderer(w, "auto")
    audio = torch.from_numpy(synth.make_audio(1, seed=2)).to(dev())
    grid = torch.stack(torch.meshgrid(torch.linspace(0, 1, 24, device=dev()), torch.linspace(0, 1, 24, device=dev()), indexing="ij"), -1)
    rep = r.probe_fp16f8_domain(audio, torch.tensor([3]), grid.reshape(-1, 2))
    print("hidden weights x4: probe %s" % rep)
    assert rep["max_activation"] >= 4096 and not rep["ok"]
    with pytest.warns(RuntimeWarning, match="probed hidden activations"):
        got = r.render_frames(audio, torch.tensor([3]), 16, 24)
    assert torch.equal(got, S.LipRenderer(w, "bf16x3").render_frames(audio, torch.tensor([3]), 16, 24))
    # the in-domain model passes the same probe
    w1 = pack(S, synth.make_state_dict(0, "kaiming", 2, 3))
    r1 = S.LipRenderer(w1, "auto")
    assert r1.probe_fp16f8_domain(audio, torch.tensor([3]), grid.reshape(-1, 2))["ok"] and r1._resolve_precision("auto") == "fp16f8"


@pytest.mark.parametrize("kind", ["randn", "prob"])
def test_fp16f8_parity_on_probability_like_audio(S, kind):
    """DeepSpeech windows as softmax-like probabilities in [0,1] as well as unnormalised logits."""
    H, W = 40, 56
    sd_np = synth.make_state_dict(0, "kaiming", 2, 3)
    w = pack(S, sd_np)
    audio = torch.from_numpy(synth.make_audio(2, seed=4, kind=kind))
    want = torch.stack([O.render_plain(O.to_torch_sd(sd_np), audio[i:i + 1], 5 + i, H, W) for i in range(2)])
    got = S.LipRenderer(w, "fp16f8").render_frames(audio.to(dev()), torch.tensor([5, 6]), H, W).cpu()
    e = (got - want).abs().max().item()
    print("fp16f8 vs oracle, audio kind %s: %.2e" % (kind, e))
    assert e < 1e-3


def test_automatic_reevaluation_threshold_follows_the_density_row(S):
    """density row x30: the tensor-core density error grows 30x, and so does the automatic fix_thr — the volumetric render
    still has zero pixels beyond 1e-3 of the exact path, while the fixed 2e-3 threshold of a kaiming-scale model would not."""
    H, W, Sn, F = 64, 64, 32, 2
    sd = {k: v.copy() for k, v in synth.make_state_dict(0, "kaiming", 3, 4).items()}
    sd["output_linear.weight"][3] *= 30.0
    sd["output_linear.bias"][3] *= 30.0
    w = pack(S, sd, 3, 4)
    assert abs(w.meta()["auto_fix_thr"] / (2e-3 * 30 * np.linalg.norm(sd["output_linear.weight"][3] / 30) / 2 ** 0.5) - 1) < 1e-3
    ro, rd = O.get_rays(H, W, 60.0, torch.eye(4)[:3])
    ro, rd = ro.reshape(-1, 3).to(dev()), rd.reshape(-1, 3).to(dev())
    z = O.z_samples(Sn).to(dev())
    audio = torch.from_numpy(synth.make_audio(F, seed=6)).to(dev())
    idx = torch.tensor([2, 3])
    kw = dict(mode="volumetric", rays_o=ro, rays_d=rd, z_vals=z)
    exact = S.LipRenderer(w, "fp32").render_frames(audio, idx, H, W, **kw)
    for prec in ("bf16x3", "fp16f8"):
        r = S.LipRenderer(w, prec)
        auto = r.render_frames(audio, idx, H, W, **kw)
        n_auto = int(r.last_render_counts()["reevaluated"].sum())
        fixed = r.render_frames(audio, idx, H, W, fix_thr=2e-3, **kw)
        n_fixed = int(r.last_render_counts()["reevaluated"].sum())
        bad_auto = int(((auto - exact).abs().amax(-1) > 1e-3).sum())
        bad_fixed = int(((fixed - exact).abs().amax(-1) > 1e-3).sum())
        print("density row x30 %s: automatic threshold re-evaluates %d rays -> %d pixels > 1e-3; fixed 2e-3: %d rays -> %d pixels"
              % (prec, n_auto, bad_auto, n_fixed, bad_fixed))
        assert n_auto > n_fixed and bad_auto == 0
