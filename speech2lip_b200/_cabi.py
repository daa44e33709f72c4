"""ctypes binding of include/speech2lip_b200.h (libs2l_b200.so).

The library is built in-tree by speech2lip_b200/csrc/build.py (nvcc, sm_100a).  There is no
fallback: if the shared object is missing the import fails loudly, and every entry point raises
RuntimeError with s2l_last_error() on a non-zero status.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("S2L_LIB_PATH") or os.path.join(_HERE, "csrc", "libs2l_b200.so")   # env override: debug builds only

NUM_PARAMS = 42
PREC_FP32, PREC_BF16X3, PREC_BF16X1, PREC_FP16F8 = 0, 1, 2, 3
PTS_GRID, PTS_GRID_ENS4, PTS_RAYS, PTS_EXPLICIT = 0, 1, 2, 3
PRECISIONS = {"fp32": PREC_FP32, "bf16x3": PREC_BF16X3, "bf16x1": PREC_BF16X1, "fp16f8": PREC_FP16F8}

# reference state_dict names in S2L_P_* order (include/speech2lip_b200.h)
PARAM_NAMES = (
    ["encoder_conv.%d.%s" % (i, s) for i in (0, 2, 4, 6) for s in ("weight", "bias")]
    + ["encoder_fc1.%d.%s" % (i, s) for i in (0, 2) for s in ("weight", "bias")]
    + ["%s.%s" % (n, s) for n in ("fc_uv", "fc_uv_skip", "fc_audio", "fc_audio_skip", "fc_time", "fc_time_skip")
       for s in ("weight", "bias")]
    + ["pts_linears.%d.%s" % (i, s) for i in range(8) for s in ("weight", "bias")]
    + ["output_linear.weight", "output_linear.bias"]
)
assert len(PARAM_NAMES) == NUM_PARAMS


class S2LGeom(C.Structure):
    _fields_ = [("n_frames", C.c_int32), ("height", C.c_int32), ("width", C.c_int32), ("n_samples", C.c_int32),
                ("pts_mode", C.c_int32), ("uv_dims", C.c_int32), ("out_ch", C.c_int32), ("z_per_ray", C.c_int32),
                ("rays_per_frame_shared", C.c_int32), ("pts_per_frame", C.c_int64), ("eps_shift", C.c_float),
                ("eps_per_frame", C.c_void_p),
                # ABI v2: volumetric options (sample chunks + early ray termination, fp32 re-evaluation threshold)
                ("sample_chunks", C.c_int32), ("term_thr", C.c_float), ("fix_thr", C.c_float)]


SYMBOLS = {
    "s2l_last_error": (C.c_char_p, []),
    "s2l_abi_version": (C.c_int32, []),
    "s2l_time_div_term": (None, [C.POINTER(C.c_float)]),
    "s2l_blob_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "s2l_blob_meta": (C.c_int32, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p]),
    "s2l_pack_weights": (C.c_int32, [C.POINTER(C.c_void_p), C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "s2l_audio_encode_fwd": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "s2l_latent_bias_fwd": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "s2l_rows_differ": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "s2l_rows_differ_or": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "s2l_audio_merge_auto_scratch_bytes": (C.c_size_t, []),
    "s2l_audio_merge_auto": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "s2l_rgb_forward_auto_scratch_bytes": (C.c_size_t, []),
    "s2l_rgb_forward_auto": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                         C.c_void_p, C.c_void_p]),
    "s2l_mlp_fwd": (C.c_int32, [C.c_void_p, C.POINTER(S2LGeom), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "s2l_rgb_forward_rows": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p,
                                         C.c_int32, C.c_int32, C.c_void_p]),
    "s2l_rgb_forward_rows_train": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                               C.c_int32, C.c_int32, C.c_void_p]),
    "s2l_mlp_bwd_rows": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p]),
    "s2l_audio_train_save_floats": (C.c_size_t, [C.c_int32]),
    "s2l_audio_train_scratch_bytes": (C.c_size_t, [C.c_int32]),
    "s2l_audio_train_fwd": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "s2l_audio_train_bwd": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p,
                                        C.c_int32, C.c_void_p]),
    "s2l_train_workspace_bytes": (C.c_size_t, [C.POINTER(S2LGeom)]),
    "s2l_train_fwd": (C.c_int32, [C.c_void_p, C.POINTER(S2LGeom), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "s2l_train_bwd": (C.c_int32, [C.c_void_p, C.POINTER(S2LGeom), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p]),
    "s2l_train_rows_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "s2l_train_rows_fwd": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "s2l_train_rows_bwd": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p]),
    "s2l_peer_buffer_bytes": (C.c_size_t, [C.c_int64]),
    "s2l_peer_payload_offset": (C.c_size_t, [C.c_int64, C.c_uint32]),
    "s2l_peer_alloc": (C.c_int32, [C.c_int64, C.POINTER(C.c_void_p), C.c_void_p]),
    "s2l_peer_open": (C.c_int32, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "s2l_peer_close": (C.c_int32, [C.c_void_p]),
    "s2l_peer_free": (C.c_int32, [C.c_void_p]),
    "s2l_allreduce_peer": (C.c_int32, [C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_int64, C.c_float, C.c_uint32, C.c_void_p, C.c_void_p]),
    "s2l_wgrad_rows_scratch_bytes": (C.c_size_t, [C.c_int64, C.c_int32, C.c_int32, C.c_int32]),
    "s2l_wgrad_rows_fp32": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_void_p,
                                        C.c_void_p, C.c_void_p]),
    "s2l_dx_rows_fp32": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]),
    "s2l_embed_fwd": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "s2l_ensemble4_blend": (C.c_int32, [C.c_void_p, C.POINTER(S2LGeom), C.c_void_p, C.c_void_p]),
    "s2l_composite_fwd": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_int64, C.c_int32,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "s2l_get_rays": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "s2l_render_scratch_bytes": (C.c_size_t, [C.POINTER(S2LGeom), C.c_int32, C.c_int32]),
    "s2l_render_counts_offset": (C.c_size_t, [C.POINTER(S2LGeom)]),
    "s2l_render_frames": (C.c_int32, [C.c_void_p, C.POINTER(S2LGeom), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "s2l_post_fusion_compose": (C.c_int32, [C.c_void_p] * 5 + [C.c_int32] * 11 + [C.c_void_p] * 3),
    "s2l_audio_windows": (C.c_int32, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "s2l_post_fusion_compose_bwd": (C.c_int32, [C.c_void_p] * 4 + [C.c_int32] * 11 + [C.c_void_p, C.c_void_p]),
    "s2l_frames_to_bgr8": (C.c_int32, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "s2l_sizeof_geom": (C.c_int32, []),
    "s2l_profile_enable": (None, [C.c_int32]),
    "s2l_profile_mlp_ms": (C.c_double, [C.POINTER(C.c_int32)]),
    "s2l_launch_count": (C.c_int64, [C.c_int32]),
    "s2l_tc_schedule": (C.c_int32, [C.c_int64]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "speech2lip_b200: CUDA library %s is missing — build it with "
                "`python speech2lip_b200/csrc/build.py` (there is no CPU / PyTorch fallback)" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)          # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        if l.s2l_sizeof_geom() != C.sizeof(S2LGeom):
            raise RuntimeError("speech2lip_b200: S2LGeom is %d bytes in the binding, %d in %s (stale build?)"
                               % (C.sizeof(S2LGeom), l.s2l_sizeof_geom(), LIB_PATH))
        _lib = l
    return _lib


def check(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed (status %d): %s" % (what, rc, lib().s2l_last_error().decode()))
