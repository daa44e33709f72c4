"""CPU: the C-ABI library loads, exports every symbol include/speech2lip_b200.h declares, and rejects
bad arguments before touching the GPU.  No compute calls here."""
import ctypes as C
import json
import math
import os
import re

import pytest
import torch

from speech2lip_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "speech2lip_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(s2l_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _cabi.lib()
    names = header_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), "libs2l_b200.so does not export %s" % n
    assert sorted(_cabi.SYMBOLS) == names, "binding table and header disagree"
    assert lib.s2l_abi_version() == 2


def test_integration_md_struct_stub_matches_the_binding():
    """INTEGRATION.md shows the ctypes S2LGeom a reference maintainer would write; it must list exactly the binding's
    fields (a stub 8 bytes short once handed the kernels a garbage eps_per_frame pointer), and both must have the size
    the compiled library reports."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    stub = doc[doc.index("class S2LGeom(C.Structure)"):doc.index("assert lib.s2l_sizeof_geom()")]
    names = re.findall(r'\("([a-z0-9_]+)", C\.c_', stub)
    assert names == [n for n, _ in _cabi.S2LGeom._fields_]
    types_ = re.findall(r'\("[a-z0-9_]+", C\.(c_[a-z0-9_]+)\)', stub)
    assert [getattr(C, t) for t in types_] == [t for _, t in _cabi.S2LGeom._fields_]
    assert _cabi.lib().s2l_sizeof_geom() == C.sizeof(_cabi.S2LGeom)


def test_param_order_matches_header_enum():
    src = open(os.path.join(ROOT, "include", "speech2lip_b200.h")).read()
    assert "S2L_NUM_PARAMS" in src
    assert _cabi.NUM_PARAMS == 42 and len(_cabi.PARAM_NAMES) == 42
    assert _cabi.PARAM_NAMES[0] == "encoder_conv.0.weight" and _cabi.PARAM_NAMES[12] == "fc_uv.weight"
    assert _cabi.PARAM_NAMES[24] == "pts_linears.0.weight" and _cabi.PARAM_NAMES[40] == "output_linear.weight"


def test_blob_size_and_div_term():
    lib = _cabi.lib()
    n = lib.s2l_blob_bytes(2, 3)
    assert 4_000_000 < n < 12_000_000 and n % 1024 == 0
    buf = (C.c_float * 10)()
    lib.s2l_time_div_term(buf)
    ref = torch.exp(torch.arange(0, 20, 2, dtype=torch.float) * -(math.log(10000.0) / 20))   # tf_nerf.py:431-432
    assert torch.equal(torch.tensor(list(buf)), ref)


def test_errors_are_reported_not_crashed():
    lib = _cabi.lib()
    assert lib.s2l_pack_weights(None, None, 2, 3, None) != 0
    assert b"null" in lib.s2l_last_error()
    arr = (C.c_void_p * _cabi.NUM_PARAMS)()
    assert lib.s2l_pack_weights(arr, C.c_void_p(16), 5, 3, None) != 0
    assert b"uv_dims" in lib.s2l_last_error()
    g = _cabi.S2LGeom(n_frames=1, height=4, width=4, n_samples=0, pts_mode=_cabi.PTS_RAYS, uv_dims=2, out_ch=3)
    assert lib.s2l_render_frames(C.c_void_p(16), C.byref(g), C.c_void_p(16), None, None, None, None, C.c_void_p(16),
                                 None, None, C.c_void_p(16), 1, None) != 0
    assert b"ray mode" in lib.s2l_last_error()
    g.pts_mode = 9
    assert lib.s2l_mlp_fwd(C.c_void_p(16), C.byref(g), C.c_void_p(16), None, None, None, None, C.c_void_p(16), 0, None) != 0
    assert lib.s2l_composite_fwd(C.c_void_p(16), C.c_void_p(16), 0, C.c_void_p(16), 4, 0, 0, C.c_void_p(16), None, None, None) != 0
    with pytest.raises(RuntimeError):
        _cabi.check(3, "demo")


def test_no_fallback_for_cpu_tensors():
    from speech2lip_b200 import renderer as R
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        R._f32c(torch.zeros(3), "x")


def test_talking_face_state_dict_matches_reference_layout():
    """115 keys with the reference's names and shapes (SURVEY §8(b) checkpoint layout)."""
    from speech2lip_b200 import TalkingFace
    keys = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_keys.json")))
    cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "may_cfg.json")))
    for tag, (uvd, och) in {"live": (2, 3), "volumetric": (3, 4)}.items():
        m = TalkingFace(device=torch.device("cpu"), cfg=cfg, uv_dims=uvd, output_ch=och)
        mine = {k: list(v.shape) for k, v in m.state_dict().items()}
        assert mine == keys[tag]
    assert len(keys["live"]) == 115
    assert m.audio_dims == 64 and hasattr(m, "post_fusion_unet") and hasattr(m, "canonical_depth_head")
    bad = json.loads(json.dumps(cfg))
    bad["model"]["MLP_version"] = "v1"
    with pytest.raises(NotImplementedError):
        TalkingFace(device=torch.device("cpu"), cfg=bad)


def test_post_fusion_runs_on_cpu():
    """outside the hot path, plain PyTorch: shapes only (tf_nerf.py:287-389)."""
    from speech2lip_b200 import TalkingFace
    cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "may_cfg.json")))
    m = TalkingFace(device=torch.device("cpu"), cfg=cfg, mode="eval").eval()
    B, H, W, lh, lw = 1, 64, 64, 8, 12
    lip = torch.rand(B, lh, lw, 3)
    face = torch.rand(B, H, W, 3)
    mask = torch.zeros(B, H, W, 1)
    mask[:, 21:29, 21:33] = 1
    ys, xs = torch.meshgrid(torch.linspace(-1, 1, H), torch.linspace(-1, 1, W), indexing="ij")
    coord = torch.stack([xs, ys], -1)[None]
    with torch.no_grad():
        recon, merged, canon = m.post_fusion2_onlylip(lip, face, face.clone(), mask, 20, 20, coord)
    assert recon.shape == (B, H, W, 3) and merged.shape == (B, H, W, 3) and canon.shape == (B, H, W, 3)


def test_training_mode_audionet_matches_oracle_on_cpu():
    """in grad mode AudioNet runs as unfold+matmul on autograd (true fp32): same numbers as the oracle."""
    from oracle import s2l_oracle as O
    from oracle import synth
    from speech2lip_b200 import TalkingFace
    cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "may_cfg.json")))
    m = TalkingFace(device=torch.device("cpu"), cfg=cfg).train()
    sd = synth.make_state_dict(0, "kaiming")
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    a = torch.from_numpy(synth.make_audio(3, seed=1))
    ref = O.audio_merge_forward(O.to_torch_sd(sd), a)
    assert (m.audio_merge_forward(a) - ref).abs().max().item() < 1e-6
    assert (m.audio_merge_forward(a.permute(0, 2, 1).contiguous()) - ref).abs().max().item() < 1e-6


def test_no_library_gemm_on_the_training_path():
    """The exact per-call backward and the tensor-core training paths call the library's own GEMM kernels
    (s2l_wgrad_rows_fp32 / s2l_dx_rows_fp32 / the tcgen05 dgrad+wgrad kernels): no torch GEMM in autograd.py."""
    import re
    src = open(os.path.join(ROOT, "speech2lip_b200", "autograd.py")).read()
    code = "\n".join(l.split("#")[0] for l in src.splitlines())
    for pat in (r"\bbmm\b", r"\bmatmul\b", r" @ ", r"\baddmm\b", r"\beinsum\b", r"F\.linear", r"torch\.outer", r"\.mm\("):
        assert not re.search(pat, code), pat
    assert "s2l_wgrad_rows_fp32" in src and "s2l_dx_rows_fp32" in src


def test_synth_generators_are_shared_not_duplicated():
    """bench.py's product arm imports the synthetic generators from the package; oracle/synth.py is only a shim."""
    from oracle import synth as a
    from speech2lip_b200 import synth as b
    assert a.make_state_dict is b.make_state_dict and a.make_audio is b.make_audio
    # ... and the only place bench.py touches oracle/ is the CPU-baseline / reference-arm sampler
    import ast
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)]:
        for node in ast.walk(fn):
            if isinstance(node, ast.ImportFrom) and (node.module or "").split(".")[0] == "oracle":
                assert fn.name in ("reference_runner", "oracle_port_runner", "eager_gpu_reference", "train_step_extras", "post_fusion_extras"), "bench.py imports oracle/ in %s()" % fn.name
    for node in tree.body:
        assert not (isinstance(node, (ast.Import, ast.ImportFrom)) and "oracle" in ast.dump(node)), "module-level oracle import"


def test_bench_constants_and_traffic_helper_cpu():
    """bench.py's roofline inputs: the algorithmic FLOP counts of SURVEY 8(d), the issued-MAC count of the folded program and
    the ncu-measured DRAM traffic scaled to a launch."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("s2l_bench", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    # 2 x (fc_uv 63*256 + fc_uv_skip 63*256 + 5*256^2 + 512*256 + 2*256^2 + 256*4) for V, 42-wide PE / 3 outputs for L
    assert b.FLOP_PER_POINT["volumetric"] == 2 * (2 * 63 * 256 + 5 * 256 * 256 + 512 * 256 + 2 * 256 * 256 + 256 * 4)
    assert b.FLOP_PER_POINT["plain"] == 2 * (2 * 42 * 256 + 5 * 256 * 256 + 512 * 256 + 2 * 256 * 256 + 256 * 3)
    # folded program: G0 64x256, G1-4, G5 (64+256)x256, G6-7, G8 256x16
    assert b.ISSUED_MAC_PER_POINT == 64 * 256 + 4 * 256 * 256 + 320 * 256 + 2 * 256 * 256 + 256 * 16
    # the traffic figure is only reported for a geometry that was actually captured (no scaling between launch sizes)
    t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[0]
    assert t["frames_per_launch"] == 8 and t["points_per_launch"] == 8 * 256 * 256 * 64 and t["fused_epilogue"]
    assert b.ncu_traffic(8, 256 * 256 * 64, t["precision"]) == t["dram_bytes_read"] + t["dram_bytes_write"]
    assert b.ncu_traffic(1, 256 * 256 * 64, t["precision"]) is None and b.ncu_traffic(8, 256 * 256 * 64, "bf16x1") is None


def test_drop_in_helper_entry_points_report_errors():
    """s2l_latent_bias_fwd / s2l_rows_differ validate their arguments before touching the device."""
    lib = _cabi.lib()
    assert lib.s2l_latent_bias_fwd(None, None, 64, None, None, 1, None) == 1 and b"null" in lib.s2l_last_error()
    assert lib.s2l_latent_bias_fwd(C.c_void_p(16), C.c_void_p(16), 64, None, C.c_void_p(16), -1, None) == 2
    assert lib.s2l_latent_bias_fwd(C.c_void_p(16), C.c_void_p(16), 64, None, C.c_void_p(16), 0, None) == 0     # empty batch: no launch
    assert lib.s2l_rows_differ(None, 4, 8, 0, 8, None, None) == 1 and b"null" in lib.s2l_last_error()
    assert lib.s2l_rows_differ(C.c_void_p(16), 4, 8, 4, 8, C.c_void_p(16), None) == 2 and b"shape" in lib.s2l_last_error()
    assert lib.s2l_tc_schedule(0) in (1, 2, 3)


def test_round2_entry_points_validate_arguments_before_touching_the_device():
    """The ABI v2 additions (render options, device-gated drop-in calls, training path, AudioNet backward, blob meta) reject
    bad arguments with a message instead of launching."""
    lib = _cabi.lib()
    p16 = C.c_void_p(16)
    g = _cabi.S2LGeom(n_frames=1, height=4, width=4, n_samples=6, pts_mode=_cabi.PTS_RAYS, uv_dims=3, out_ch=4, sample_chunks=2, term_thr=1e-4)
    # 6 samples cannot be cut into chunks of 4..128 samples: refused, not silently rendered in one launch
    assert lib.s2l_render_frames(p16, C.byref(g), p16, None, p16, p16, p16, p16, None, None, p16, _cabi.PREC_FP16F8, None) == 2
    assert b"sample_chunks" in lib.s2l_last_error()
    g.sample_chunks = 0
    assert lib.s2l_render_frames(p16, C.byref(g), p16, None, p16, p16, p16, p16, None, None, p16, 9, None) == 2 and b"precision" in lib.s2l_last_error()
    # scratch: the fused volumetric render needs no raw tensor, the unfused one does
    g2 = _cabi.S2LGeom(n_frames=8, height=256, width=256, n_samples=64, pts_mode=_cabi.PTS_RAYS, uv_dims=3, out_ch=4)
    fused = lib.s2l_render_scratch_bytes(C.byref(g2), _cabi.PREC_FP16F8, 0)
    unfused = lib.s2l_render_scratch_bytes(C.byref(g2), _cabi.PREC_FP16F8, 1)
    exact = lib.s2l_render_scratch_bytes(C.byref(g2), _cabi.PREC_FP32, 0)
    assert fused < 32 << 20 and unfused > 8 * 256 * 256 * 64 * 16 and exact > 8 * 256 * 256 * 64 * 16
    assert lib.s2l_render_counts_offset(C.byref(g2)) == 8 * 4 * 256 * 4
    assert lib.s2l_rgb_forward_auto(None, None, 4, None, None, 2, 3, 1, None, None) == 1 and b"null" in lib.s2l_last_error()
    assert lib.s2l_rgb_forward_auto(p16, p16, -1, None, p16, 2, 3, 1, p16, None) == 2
    assert lib.s2l_rgb_forward_auto(p16, p16, 4, None, p16, 5, 3, 1, p16, None) == 2 and b"uv_dims" in lib.s2l_last_error()
    assert lib.s2l_audio_merge_auto(None, None, 0, 4, None, None, None) == 1
    assert lib.s2l_audio_merge_auto_scratch_bytes() >= 260 and lib.s2l_rgb_forward_auto_scratch_bytes() >= 4100
    gt = _cabi.S2LGeom(n_frames=2, height=8, width=12, pts_mode=_cabi.PTS_GRID, uv_dims=2, out_ch=3)
    assert lib.s2l_train_fwd(p16, C.byref(gt), p16, p16, p16, p16, p16, None) == 2 and b"4-tap" in lib.s2l_last_error()
    gt.pts_mode = _cabi.PTS_GRID_ENS4
    assert lib.s2l_train_fwd(p16, C.byref(gt), None, p16, p16, p16, p16, None) == 1
    rows = 2 * 3 * 128                                    # 8*12*4 = 384 points = 3 tiles per frame
    assert lib.s2l_train_workspace_bytes(C.byref(gt)) > rows * (8 * 512 * 2 + 128 + 32)
    arr = (C.c_void_p * _cabi.NUM_PARAMS)()
    assert lib.s2l_train_bwd(p16, C.byref(gt), p16, p16, p16, p16, p16, arr, p16, None) == 3 and b"gradient buffer" in lib.s2l_last_error()
    assert lib.s2l_audio_train_save_floats(3) == 3 * 640 and lib.s2l_audio_train_scratch_bytes(3) == 3 * 32800 * 4
    arr12 = (C.c_void_p * 12)()
    assert lib.s2l_audio_train_bwd(p16, p16, 0, p16, p16, arr12, p16, 2, None) == 3
    assert lib.s2l_blob_meta(None, None, None, None, None, None) == 1


def test_render_lip_train_and_dropin_refuse_cpu_tensors():
    """no CPU fallback anywhere on the product path: the training render and the drop-in's fast calls raise on CPU tensors"""
    import speech2lip_b200 as s2l
    cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "may_cfg.json")))
    m = s2l.TalkingFace(device=torch.device("cpu"), cfg=cfg)
    with pytest.raises(RuntimeError, match="CUDA"):
        m.render_lip_train(torch.zeros(1, 16, 29), torch.tensor([0]), 4, 4, torch.tensor([0.0]))
    with pytest.raises(RuntimeError, match="CUDA"):
        with torch.no_grad():
            m.rgb_forward(torch.zeros(8, 66), time_pts=torch.tensor([1]))
    with pytest.raises(ValueError):
        os.environ["S2L_DROPIN_PRECISION"] = "bf16x1"
        try:
            s2l.TalkingFace(device=torch.device("cpu"), cfg=cfg)
        finally:
            del os.environ["S2L_DROPIN_PRECISION"]


def test_documents_only_cite_profile_files_that_exist():
    """Every `profiles/<file>` (and `r2x_<file>` listed in profiles/README.md) the documents cite is committed."""
    missing = []
    for doc in ("DESIGN.md", "README.md", "INTEGRATION.md", os.path.join("profiles", "README.md")):
        text = open(os.path.join(ROOT, doc)).read()
        names = set(re.findall(r"profiles/([A-Za-z0-9_.\-]+\.(?:md|json|csv|txt))", text))
        if doc.endswith(os.path.join("profiles", "README.md")):
            names |= set(re.findall(r"`(r[0-9][a-z]_[A-Za-z0-9_.\-]+\.(?:md|json|csv|txt))`", text))
        for n in sorted(names):
            if "*" not in n and not os.path.exists(os.path.join(ROOT, "profiles", n)):
                missing.append((doc, n))
    assert not missing, missing
