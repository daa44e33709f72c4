"""Training path of TalkingFace.rgb_forward: a torch.autograd.Function whose forward is the fused fp32 kernel
(s2l_rgb_forward_rows_train, saves 10 activation tensors) and whose backward is the fused data-gradient kernel
(s2l_mlp_bwd_rows) followed by the library's own fp32 GEMM kernels for the weight gradients (dW_l = dPre_l^T h_{l-1},
s2l_wgrad_rows_fp32) and the per-row latent / coordinate gradient (s2l_dx_rows_fp32) — no torch GEMM on the path.
Replaces autograd through tf_nerf.py:225-285 (loss.backward(), training.py:559) in exact fp32 arithmetic; the tensor-core
(bf16) training paths are FusedLipRender and FusedMLPRowsTC below."""
import ctypes as C

import torch

from . import _cabi
from .renderer import _ptr, _stream

_W_ORDER = (["fc_uv", "fc_uv_skip", "fc_audio", "fc_audio_skip", "fc_time", "fc_time_skip"]
            + ["pts_linears.%d" % i for i in range(8)] + ["output_linear"])


def param_order():
    return [n + s for n in _W_ORDER for s in (".weight", ".bias")]


def _wgrad(dy, h, L, stride_dy, stride_h):
    """out[l] = dy_l^T h_l for L matrices: dy_l [N,A] at dy + l*stride_dy floats, h_l [N,B] at h + l*stride_h floats (a stride of
    0 shares the operand) -> [L,A,B].  dy / h are the FIRST matrices (contiguous [N,A] / [N,B] views).  One launch of the
    library's split-K fp32 kernel (partials summed in a fixed order: deterministic)."""
    lib = _cabi.lib()
    N, A, B = dy.shape[0], dy.shape[1], h.shape[1]
    assert dy.is_contiguous() and h.is_contiguous() and h.shape[0] == N
    out = torch.empty(L, A, B, device=dy.device)
    scratch = torch.empty(max(16, lib.s2l_wgrad_rows_scratch_bytes(N, L, A, B)), dtype=torch.uint8, device=dy.device)
    _cabi.check(lib.s2l_wgrad_rows_fp32(_ptr(dy), _ptr(h), N, L, A, B, stride_dy, stride_h, _ptr(out), _ptr(scratch), _stream()),
                "s2l_wgrad_rows_fp32")
    return out


class FusedMLPRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, time_idx, packed, div_term, *params):
        lib = _cabi.lib()
        x = x.contiguous().float()
        N = x.shape[0]
        out = torch.empty(N, packed.out_ch, device=x.device)
        acts = torch.empty(10, N, 256, device=x.device)
        has_time = time_idx is not None
        with torch.cuda.device(x.device):
            _cabi.check(lib.s2l_rgb_forward_rows_train(_ptr(packed.blob), _ptr(x), N, int(time_idx) if has_time else 0,
                                                       1 if has_time else 0, _ptr(out), _ptr(acts), packed.uv_dims,
                                                       packed.out_ch, _stream()), "s2l_rgb_forward_rows_train")
        ctx.packed = packed
        ctx.time_idx = time_idx
        ctx.div_term = div_term
        ctx.save_for_backward(x, acts, *params)
        return out

    @staticmethod
    def backward(ctx, d_out):
        with torch.cuda.device(ctx.saved_tensors[0].device):
            return FusedMLPRows._backward(ctx, d_out)

    @staticmethod
    def _backward(ctx, d_out):
        lib = _cabi.lib()
        x, acts, *params = ctx.saved_tensors
        P = dict(zip(param_order(), params))
        packed = ctx.packed
        N, D = x.shape[0], packed.uv_dims
        d_out = d_out.contiguous().float()
        dsave = torch.empty(10, N, 256, device=x.device)
        E = D + 20 * D
        pe = torch.empty(N, E, device=x.device)
        with torch.cuda.device(x.device):
            _cabi.check(lib.s2l_mlp_bwd_rows(_ptr(packed.blob), _ptr(d_out), _ptr(acts), N, _ptr(dsave), packed.out_ch,
                                             _stream()), "s2l_mlp_bwd_rows")
            _cabi.check(lib.s2l_embed_fwd(_ptr(x), N, x.shape[1], D, _ptr(pe), _stream()), "s2l_embed_fwd")
        d_net, d_skip = dsave[0], dsave[6]
        lat = x[:, D:].contiguous()
        g = {}
        mat = N * 256                                                        # floats between consecutive [N,256] buffers
        # ---- weight gradients: the library's fp32 split-K GEMM over the saved / produced [N,256] buffers
        g["output_linear.weight"] = _wgrad(d_out, acts[9], 1, 0, 0)[0]
        g["output_linear.bias"] = d_out.sum(0)
        # dsave rows: 0 d_net, 1-5 dPre of layers 0-4, 6 d_skip, 7-9 dPre of layers 5-7;  acts rows: 0-4 inputs of layers 0-4,
        # 5 h4 / 6 h_skip (the two halves of layer 5's input), 7-8 inputs of layers 6-7, 9 input of output_linear.
        # Same-shape products are batched (the step is launch-bound at the reference's 9 600-row calls).
        col = dsave.sum(1)                                                   # every bias gradient in one reduction [10,256]
        w04 = _wgrad(dsave[1], acts[0], 5, mat, mat)                         # layers 0-4
        w67 = _wgrad(dsave[8], acts[7], 2, mat, mat)                         # layers 6-7
        w5 = _wgrad(dsave[7], acts[5], 2, 0, mat)                            # layer 5 against (h4, h_skip)
        for l in range(5):
            g["pts_linears.%d.weight" % l] = w04[l]
            g["pts_linears.%d.bias" % l] = col[1 + l]
        g["pts_linears.5.weight"] = torch.cat([w5[1], w5[0]], 1)            # input order of layer 5: [h_skip | h4]
        g["pts_linears.5.bias"] = col[7]
        for i, l in enumerate((6, 7)):
            g["pts_linears.%d.weight" % l] = w67[i]
            g["pts_linears.%d.bias" % l] = col[8 + i]
        s_net, s_skip = col[0], col[6]
        wuv = _wgrad(d_net, pe, 2, 6 * mat, 0)                               # fc_uv / fc_uv_skip (d_net = dsave[0], d_skip = dsave[6])
        wau = _wgrad(d_net, lat, 2, 6 * mat, 0)                              # fc_audio / fc_audio_skip
        g["fc_uv.weight"], g["fc_uv_skip.weight"] = wuv[0], wuv[1]
        g["fc_audio.weight"], g["fc_audio_skip.weight"] = wau[0], wau[1]
        for n, s in (("fc_uv", s_net), ("fc_audio", s_net), ("fc_uv_skip", s_skip), ("fc_audio_skip", s_skip)):
            g[n + ".bias"] = s
        if ctx.time_idx is not None:
            ang = torch.tensor(float(ctx.time_idx), device=x.device) * ctx.div_term            # tf_nerf.py:439-440
            tpe = torch.stack([torch.sin(ang), torch.cos(ang)], 1).reshape(-1)
            g["fc_time.weight"] = s_net[:, None] * tpe[None, :]
            g["fc_time_skip.weight"] = s_skip[:, None] * tpe[None, :]
            g["fc_time.bias"], g["fc_time_skip.bias"] = s_net, s_skip
        else:
            for n in ("fc_time", "fc_time_skip"):
                g[n + ".weight"] = torch.zeros_like(P[n + ".weight"])
                g[n + ".bias"] = torch.zeros_like(P[n + ".bias"])
        # ---- input gradient: the latent columns and — a caller may differentiate w.r.t. the coordinates too (learnable warps,
        #      depth) — the uv columns through the positional encoding's Jacobian (1, f cos(f x), -f sin(f x)), tf_nerf.py:404-425
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            d_pe = torch.empty(N, E, device=x.device)                                         # [N, D + 20 D]
            wts = [P[n + ".weight"].detach().contiguous().float() for n in ("fc_audio", "fc_audio_skip", "fc_uv", "fc_uv_skip")]
            _cabi.check(lib.s2l_dx_rows_fp32(_ptr(d_net), _ptr(wts[0]), _ptr(d_skip), _ptr(wts[1]), N, 64,
                                             dx.data_ptr() + 4 * D, D + 64, _stream()), "s2l_dx_rows_fp32")
            _cabi.check(lib.s2l_dx_rows_fp32(_ptr(d_net), _ptr(wts[2]), _ptr(d_skip), _ptr(wts[3]), N, E, _ptr(d_pe), E, _stream()),
                        "s2l_dx_rows_fp32")
            freqs = (2.0 ** torch.arange(10, device=x.device, dtype=torch.float32))          # 2**linspace(0, 9, 10)
            ang = x[:, None, :D] * freqs[None, :, None]                                       # [N, 10, D]
            blocks = d_pe[:, D:].reshape(N, 10, 2, D)                                         # per frequency: (sin block, cos block)
            dx[:, :D] = d_pe[:, :D] + ((blocks[:, :, 0] * torch.cos(ang) - blocks[:, :, 1] * torch.sin(ang)) * freqs[None, :, None]).sum(1)
        grads = [g[n] if ctx.needs_input_grad[4 + i] else None for i, n in enumerate(param_order())]
        return (dx, None, None, None, *grads)


def rgb_forward_train(module, x, time_idx):
    """x [N, uv_dims+64] (latent columns may require grad), module = speech2lip_b200.TalkingFace."""
    sd = module._hot_params()
    params = [sd[n] for n in param_order()]
    div = module.__dict__.get("_div_term")
    if div is None or div.device != x.device:
        import math
        div = torch.exp(torch.arange(0, 20, 2, dtype=torch.float) * -(math.log(10000.0) / 20)).to(x.device)   # tf_nerf.py:431-432
        module.__dict__["_div_term"] = div
    return FusedMLPRows.apply(x, time_idx, module.packed_weights(), div, *params)


# ------------------------------------------------------------------------------------------------------------------
# Tensor-core training render: F frames x 4 taps in ONE differentiable launch sequence (include/speech2lip_b200.h,
# s2l_train_fwd / s2l_train_bwd).  Replaces autograd through Trainer.predict_lip_image (training.py:158-251) called once
# per frame — and five more times per frame for the sync-expert window (training.py:500-548).  bf16 operands, fp32 accumulate.
MLP_PARAM_NAMES = _cabi.PARAM_NAMES[12:]          # fc_uv ... output_linear, S2L_P_* order


class FusedLipRender(torch.autograd.Function):
    @staticmethod
    def forward(ctx, latent, index, eps, H, W, packed, *params):
        lib = _cabi.lib()
        if packed.uv_dims != 2 or packed.out_ch != 3:
            raise ValueError("the training render is the live 4-tap mode (uv_dims=2, output_ch=3)")
        if not latent.is_cuda:
            raise RuntimeError("speech2lip_b200: render_lip_train needs CUDA tensors (the hot path has no CPU fallback)")
        latent = latent.contiguous().float()
        F = latent.shape[0]
        dev = latent.device
        idx = torch.as_tensor(index).to(device=dev, dtype=torch.int64).reshape(-1).contiguous()
        eps = torch.as_tensor(eps, dtype=torch.float32).to(dev).reshape(-1)
        eps = (eps.expand(F) if eps.numel() == 1 else eps).contiguous()
        if idx.numel() != F or eps.numel() != F:
            raise ValueError("index / eps_shift must have one entry per frame")
        g = _cabi.S2LGeom(n_frames=F, height=int(H), width=int(W), n_samples=0, pts_mode=_cabi.PTS_GRID_ENS4, uv_dims=2, out_ch=3,
                          z_per_ray=0, rays_per_frame_shared=0, pts_per_frame=0, eps_shift=0.0, eps_per_frame=eps.data_ptr())
        ws = torch.empty(lib.s2l_train_workspace_bytes(C.byref(g)), dtype=torch.uint8, device=dev)
        rgb = torch.empty(F, H, W, 3, device=dev)
        bias = torch.empty(F, 4, 256, device=dev)
        with torch.cuda.device(dev):
            _cabi.check(lib.s2l_train_fwd(_ptr(packed.blob), C.byref(g), _ptr(latent), _ptr(idx), _ptr(rgb), _ptr(bias), _ptr(ws),
                                          _stream()), "s2l_train_fwd")
        ctx.packed, ctx.geom = packed, g
        ctx.shapes = [tuple(p.shape) for p in params]
        ctx.save_for_backward(latent, idx, eps, bias, ws)
        return rgb

    @staticmethod
    def backward(ctx, d_rgb):
        lib = _cabi.lib()
        latent, idx, eps, bias, ws = ctx.saved_tensors
        dev = latent.device
        d_rgb = d_rgb.contiguous().float()
        grads = [torch.empty(sh, device=dev) for sh in ctx.shapes]
        d_latent = torch.empty_like(latent)
        arr = (C.c_void_p * _cabi.NUM_PARAMS)(*([None] * 12 + [g.data_ptr() for g in grads]))
        with torch.cuda.device(dev):
            _cabi.check(lib.s2l_train_bwd(_ptr(ctx.packed.blob), C.byref(ctx.geom), _ptr(d_rgb), _ptr(latent), _ptr(idx), _ptr(bias),
                                          _ptr(ws), arr, _ptr(d_latent), _stream()), "s2l_train_bwd")
        return (d_latent, None, None, None, None, None, *grads)


def render_lip_train(module, audio, index, H, W, eps_shift=None):
    """Differentiable 4-tap render of F lip frames (speech2lip_b200.TalkingFace.render_lip_train).  audio [F,16,29] (or
    [F,29,16]), index [F]; eps_shift: [F] / scalar, or None to draw one per frame exactly as predict_lip_image does
    (training.py:198-200: ry * torch.rand(1, device) / 2 per call, frames in order)."""
    dev = audio.device
    F = audio.shape[0]
    if eps_shift is None:
        eps_shift = torch.cat([(0.5 / H) * torch.rand(1, device=dev) / 2.0 for _ in range(F)])
    latent = module.audio_merge_forward(audio)                      # AudioNet on autograd (67 k MAC per frame)
    sd = module._hot_params()
    params = [sd[n] for n in MLP_PARAM_NAMES]
    return FusedLipRender.apply(latent, index, eps_shift, int(H), int(W), module.packed_weights(), *params)


# ------------------------------------------------------------------------------------------------------------------
# AudioNet under autograd (tf_nerf.py:197-213): forward = the inference kernel + saved activations, backward = one CTA per
# frame + a frame reduction (s2l_audio_train_fwd / s2l_audio_train_bwd).  No library convolution / GEMM.
AUDIO_PARAM_NAMES = _cabi.PARAM_NAMES[:12]


class AudioNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, audio, packed, *params):
        lib = _cabi.lib()
        if not audio.is_cuda:
            raise RuntimeError("speech2lip_b200: AudioNet needs CUDA tensors (the hot path has no CPU fallback)")
        audio = audio.contiguous().float()
        if audio.dim() != 3 or audio.shape[1] * audio.shape[2] != 16 * 29:
            raise ValueError("audio must be [B,16,29] or [B,29,16], got %s" % (tuple(audio.shape),))
        F = audio.shape[0]
        transposed = 1 if audio.shape[2] == 16 else 0
        latent = torch.empty(F, 64, device=audio.device)
        save = torch.empty(lib.s2l_audio_train_save_floats(F), device=audio.device)
        with torch.cuda.device(audio.device):
            _cabi.check(lib.s2l_audio_train_fwd(_ptr(packed.blob), _ptr(audio), transposed, _ptr(latent), _ptr(save), F, _stream()),
                        "s2l_audio_train_fwd")
        ctx.packed, ctx.transposed = packed, transposed
        ctx.shapes = [tuple(p.shape) for p in params]
        ctx.save_for_backward(audio, save)
        return latent

    @staticmethod
    def backward(ctx, d_latent):
        lib = _cabi.lib()
        audio, save = ctx.saved_tensors
        F = audio.shape[0]
        d_latent = d_latent.contiguous().float()
        grads = [torch.empty(sh, device=audio.device) for sh in ctx.shapes]
        scratch = torch.empty(max(lib.s2l_audio_train_scratch_bytes(F), 4), dtype=torch.uint8, device=audio.device)
        arr = (C.c_void_p * 12)(*[g.data_ptr() for g in grads])
        with torch.cuda.device(audio.device):
            _cabi.check(lib.s2l_audio_train_bwd(_ptr(ctx.packed.blob), _ptr(audio), ctx.transposed, _ptr(save), _ptr(d_latent), arr,
                                                _ptr(scratch), F, _stream()), "s2l_audio_train_bwd")
        return (None, None, *grads)


def audio_merge_forward_train(module, audio):
    sd = module._hot_params()
    return AudioNetFn.apply(audio, module.packed_weights(), *[sd[n] for n in AUDIO_PARAM_NAMES])


# ------------------------------------------------------------------------------------------------------------------
# The per-call contract on the tensor-core training kernels (opt-in, bf16): one rgb_forward call whose rows share one latent
# (what the unmodified Trainer.predict_lip_image passes, training.py:216-233).  s2l_train_rows_fwd / s2l_train_rows_bwd.
class FusedMLPRowsTC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, time_idx, packed, *params):
        lib = _cabi.lib()
        x = x.contiguous().float()
        N = x.shape[0]
        dev = x.device
        tdev = None if time_idx is None else torch.as_tensor(time_idx).reshape(-1)[:1].to(device=dev, dtype=torch.int64)
        out = torch.empty(N, 3, device=dev)
        bias = torch.empty(1, 4, 256, device=dev)
        ws = torch.empty(lib.s2l_train_rows_workspace_bytes(N), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _cabi.check(lib.s2l_train_rows_fwd(_ptr(packed.blob), _ptr(x), N, _ptr(tdev), _ptr(out), _ptr(bias), _ptr(ws), _stream()),
                        "s2l_train_rows_fwd")
        ctx.packed = packed
        ctx.shapes = [tuple(p.shape) for p in params]
        ctx.has_time = tdev is not None
        ctx.save_for_backward(x, bias, ws, tdev if tdev is not None else torch.empty(0, device=dev))
        return out

    @staticmethod
    def backward(ctx, d_out):
        lib = _cabi.lib()
        x, bias, ws, tdev = ctx.saved_tensors
        dev = x.device
        d_out = d_out.contiguous().float()
        grads = [torch.empty(sh, device=dev) for sh in ctx.shapes]
        d_latent = torch.empty(64, device=dev)
        arr = (C.c_void_p * _cabi.NUM_PARAMS)(*([None] * 12 + [g.data_ptr() for g in grads]))
        with torch.cuda.device(dev):
            _cabi.check(lib.s2l_train_rows_bwd(_ptr(ctx.packed.blob), _ptr(d_out), _ptr(x), x.shape[0], _ptr(tdev if ctx.has_time else None),
                                               _ptr(bias), _ptr(ws), arr, _ptr(d_latent), _stream()), "s2l_train_rows_bwd")
        dx = None
        if ctx.needs_input_grad[0]:
            # every row carries the same latent: its gradient is the sum over the rows, returned on row 0 (whatever tiled the
            # latent — tile / expand — sums the rows' gradients anyway); no coordinate gradient on this path
            dx = torch.zeros_like(x)
            dx[0, 2:] = d_latent
        return (dx, None, None, *grads)


def rgb_forward_train_tc(module, x, time_pts):
    sd = module._hot_params()
    return FusedMLPRowsTC.apply(x, time_pts, module.packed_weights(), *[sd[n] for n in MLP_PARAM_NAMES])


# ------------------------------------------------------------------------------------------------------------------
# Post-fusion compose with a gradient to the lip crop (training.py:436-445: the lip render is pasted, warped and refined
# by the UNet; the loss gradient has to come back to the lip MLP).  Forward = the fused gather-blend kernel, backward =
# its scatter kernel.  Replaces autograd through F.pad / the mask blend / 2x F.grid_sample of tf_nerf.py:334-386.
class PostFusionCompose(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rgb_lip, face, gt, mask, coord, x0, y0, paste_shift, expand_pad):
        from .renderer import post_fusion_compose
        fused, canon = post_fusion_compose(rgb_lip, face, gt, mask, coord, x0, y0, paste_shift, expand_pad)
        m = mask.detach().contiguous().float()
        if m.shape[-1] == 1:
            m = m.expand(-1, -1, -1, 3).contiguous()
        ctx.save_for_backward(m, coord.detach().contiguous().float())
        ctx.geom = (tuple(rgb_lip.shape), tuple(face.shape), int(x0), int(y0), bool(paste_shift), int(expand_pad))
        ctx.set_materialize_grads(False)          # an output nobody differentiates arrives as None, not as a zero tensor
        return fused, canon

    @staticmethod
    def backward(ctx, d_fused, d_canon):
        lib = _cabi.lib()
        mask, coord = ctx.saved_tensors
        (B, lh, lw, _), (_, h, w, _), x0, y0, shift, pad = ctx.geom
        Hf, Wf = coord.shape[1], coord.shape[2]
        d_lip = torch.empty(B, lh, lw, 3, device=mask.device)
        df = d_fused.contiguous().float() if d_fused is not None else None
        dc = d_canon.contiguous().float() if d_canon is not None else None
        with torch.cuda.device(mask.device):
            _cabi.check(lib.s2l_post_fusion_compose_bwd(_ptr(df), _ptr(dc), _ptr(mask), _ptr(coord), B, lh, lw, h, w, Hf, Wf, x0, y0,
                                                        1 if shift else 0, pad, _ptr(d_lip), _stream()), "s2l_post_fusion_compose_bwd")
        return d_lip, None, None, None, None, None, None, None, None
