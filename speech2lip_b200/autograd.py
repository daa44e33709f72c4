"""Training path of TalkingFace.rgb_forward: a torch.autograd.Function whose forward is the fused fp32 kernel
(s2l_rgb_forward_rows_train, saves 10 activation tensors) and whose backward is the fused data-gradient kernel
(s2l_mlp_bwd_rows) followed by plain library GEMMs for the weight gradients (dW_l = dPre_l^T h_{l-1}) and
the per-row latent gradient.  Replaces autograd through tf_nerf.py:225-285 (loss.backward(), training.py:559).
SURVEY §8(f) rank 2 — first step: exact fp32 arithmetic; a tensor-core dgrad/wgrad is future work."""
import ctypes as C

import torch

from . import _cabi
from .renderer import _ptr, _stream

_W_ORDER = (["fc_uv", "fc_uv_skip", "fc_audio", "fc_audio_skip", "fc_time", "fc_time_skip"]
            + ["pts_linears.%d" % i for i in range(8)] + ["output_linear"])


def param_order():
    return [n + s for n in _W_ORDER for s in (".weight", ".bias")]


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
        lat = x[:, D:]
        g = {}
        # ---- weight gradients: plain GEMMs over the saved / produced [N,256] buffers
        g["output_linear.weight"] = d_out.t() @ acts[9]
        g["output_linear.bias"] = d_out.sum(0)
        dpre = {0: dsave[1], 1: dsave[2], 2: dsave[3], 3: dsave[4], 4: dsave[5], 5: dsave[7], 6: dsave[8], 7: dsave[9]}
        h_in = {0: acts[0], 1: acts[1], 2: acts[2], 3: acts[3], 4: acts[4], 6: acts[7], 7: acts[8]}
        for l in range(8):
            if l == 5:
                g["pts_linears.5.weight"] = torch.cat([dpre[5].t() @ acts[6], dpre[5].t() @ acts[5]], 1)   # [h_skip | h4]
            else:
                g["pts_linears.%d.weight" % l] = dpre[l].t() @ h_in[l]
            g["pts_linears.%d.bias" % l] = dpre[l].sum(0)
        s_net, s_skip = d_net.sum(0), d_skip.sum(0)
        g["fc_uv.weight"] = d_net.t() @ pe
        g["fc_uv_skip.weight"] = d_skip.t() @ pe
        g["fc_audio.weight"] = d_net.t() @ lat
        g["fc_audio_skip.weight"] = d_skip.t() @ lat
        for n, s in (("fc_uv", s_net), ("fc_audio", s_net), ("fc_uv_skip", s_skip), ("fc_audio_skip", s_skip)):
            g[n + ".bias"] = s
        if ctx.time_idx is not None:
            ang = torch.tensor(float(ctx.time_idx), device=x.device) * ctx.div_term            # tf_nerf.py:439-440
            tpe = torch.stack([torch.sin(ang), torch.cos(ang)], 1).reshape(-1)
            g["fc_time.weight"] = torch.outer(s_net, tpe)
            g["fc_time_skip.weight"] = torch.outer(s_skip, tpe)
            g["fc_time.bias"], g["fc_time_skip.bias"] = s_net, s_skip
        else:
            for n in ("fc_time", "fc_time_skip"):
                g[n + ".weight"] = torch.zeros_like(P[n + ".weight"])
                g[n + ".bias"] = torch.zeros_like(P[n + ".bias"])
        # ---- input gradient: latent columns only (coordinates are constants in the reference's training loop)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.zeros_like(x)
            dx[:, D:] = d_net @ P["fc_audio.weight"] + d_skip @ P["fc_audio_skip.weight"]
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
