"""Deterministic synthetic weights and inputs (no arithmetic of the hot path lives here: generators only).

There are no pretrained Speech2Lip checkpoints, no audio.npy and no images in the
reference tree, so every parity test, golden vector and bench
run uses weights/inputs generated here from a numpy PCG64 stream.  The generator
is independent of torch's RNG so that the fixtures under tests/golden/ can be
re-derived bit-for-bit on any box (build container or GPU box) without the
reference being present.

Parameter names / shapes follow the reference state_dict
(/root/reference/src/face_simple/models/tf_nerf.py:85-172, probe in SURVEY §8(b)).
"""
import math

import numpy as np

AUDIO_WIN = 16      # DeepSpeech window length   (tf_nerf.py:91 "n x 29 x 16")
AUDIO_FEAT = 29     # DeepSpeech logits          (tf_nerf.py:88)
LATENT = 64         # AudioNet output            (tf_nerf.py:63-64)
HIDDEN = 256        # W                          (tf_nerf.py:15)
TIME_PE = 20        # 2 * time_multires          (tf_nerf.py:77)
UV_MULTIRES = 10    # may.yaml:13 uv_embed


def pe_dims(uv_dims: int) -> int:
    # Embedder.out_dims, tf_nerf.py:399-400
    return uv_dims + 2 * UV_MULTIRES * uv_dims


def hot_path_shapes(uv_dims: int = 2, output_ch: int = 3):
    """(name -> shape) of every tensor the hot path reads, reference naming."""
    e = pe_dims(uv_dims)
    shapes = {
        "encoder_conv.0.weight": (32, AUDIO_FEAT, 3), "encoder_conv.0.bias": (32,),
        "encoder_conv.2.weight": (32, 32, 3), "encoder_conv.2.bias": (32,),
        "encoder_conv.4.weight": (64, 32, 3), "encoder_conv.4.bias": (64,),
        "encoder_conv.6.weight": (64, 64, 3), "encoder_conv.6.bias": (64,),
        "encoder_fc1.0.weight": (64, 64), "encoder_fc1.0.bias": (64,),
        "encoder_fc1.2.weight": (LATENT, 64), "encoder_fc1.2.bias": (LATENT,),
        "output_linear.weight": (output_ch, HIDDEN), "output_linear.bias": (output_ch,),
        "fc_uv.weight": (HIDDEN, e), "fc_uv.bias": (HIDDEN,),
        "fc_uv_skip.weight": (HIDDEN, e), "fc_uv_skip.bias": (HIDDEN,),
        "fc_audio.weight": (HIDDEN, LATENT), "fc_audio.bias": (HIDDEN,),
        "fc_audio_skip.weight": (HIDDEN, LATENT), "fc_audio_skip.bias": (HIDDEN,),
        "fc_time.weight": (HIDDEN, TIME_PE), "fc_time.bias": (HIDDEN,),
        "fc_time_skip.weight": (HIDDEN, TIME_PE), "fc_time_skip.bias": (HIDDEN,),
    }
    for i in range(8):
        k = 2 * HIDDEN if i == 5 else HIDDEN      # pts_linears.5 takes cat([h_skip, h]) (tf_nerf.py:170-172)
        shapes["pts_linears.%d.weight" % i] = (HIDDEN, k)
        shapes["pts_linears.%d.bias" % i] = (HIDDEN,)
    return shapes


def _fan_in(shape):
    n = 1
    for s in shape[1:]:
        n *= s
    return n


def make_state_dict(seed: int = 0, kind: str = "default", uv_dims: int = 2, output_ch: int = 3):
    """numpy float32 state dict of the hot-path tensors.

    kind="default": U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weights and biases —
        the distribution nn.Linear / nn.Conv1d draw by default.
    kind="kaiming": N(0, 2/fan_in) on every fc_*, pts_linears.*, output_linear
        weight (O(1) outputs — the harder parity case); AudioNet
        and biases stay "default".
    kind="trained": "kaiming" trunk, output layer scaled so raw RGB has sigma ~0.3.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    sd = {}
    for name, shape in hot_path_shapes(uv_dims, output_ch).items():
        fan_in = _fan_in(shape) if name.endswith("weight") else None
        if name.endswith("bias"):
            wshape = hot_path_shapes(uv_dims, output_ch)[name[:-4] + "weight"]
            bound = 1.0 / math.sqrt(_fan_in(wshape))
            v = rng.uniform(-bound, bound, size=shape)
        elif kind in ("kaiming", "trained") and not name.startswith("encoder_"):
            v = rng.normal(0.0, math.sqrt(2.0 / fan_in), size=shape)
            if kind == "trained" and name.startswith("output_linear"):
                v = v * 0.3
        else:
            bound = 1.0 / math.sqrt(fan_in)
            v = rng.uniform(-bound, bound, size=shape)
        sd[name] = np.ascontiguousarray(v, dtype=np.float32)
    return sd


def make_audio(n_frames: int, seed: int = 1, kind: str = "randn"):
    """[F,16,29] float32 DeepSpeech-like windows (unnormalised logits ~ randn;
    kind="prob" gives a softmax-like set in [0,1])."""
    rng = np.random.Generator(np.random.PCG64(seed))
    a = rng.normal(0.0, 1.0, size=(n_frames, AUDIO_WIN, AUDIO_FEAT))
    if kind == "prob":
        a = np.exp(a)
        a = a / a.sum(-1, keepdims=True)
    return np.ascontiguousarray(a, dtype=np.float32)
