"""Host side of the hot path: packed weights + the batched frame renderer.

Everything here is plumbing around the C ABI (include/speech2lip_b200.h): torch owns device memory
and streams, the CUDA library does the arithmetic.  There is no PyTorch implementation of the path in
this package — if the CUDA library is missing or the tensors are not on a CUDA device, calls raise.

Reference call sites this replaces:
  inference.py:144-159                     -> LipRenderer.render_frames(mode="plain")
  src/face_simple/training.py:158-251      -> LipRenderer.render_frames(mode="ensemble4")
  TalkingFace(uv_dims=3,output_ch=4) + src/common.py:12-21 + src/face_simple/rendering.py:30-62
                                           -> LipRenderer.render_frames(mode="volumetric")
"""
import ctypes as C

import torch

from . import _cabi
from ._cabi import S2LGeom


def _ptr(t):
    return C.c_void_p(0) if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError("speech2lip_b200: %s must be a CUDA tensor (the hot path has no CPU fallback)" % name)


def _f32c(t, name):
    _need_cuda(t, name)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class PackedWeights:
    """Kernel-layout copy of the hot-path parameters (s2l_pack_weights).  `params` maps the reference's
    state_dict names (tf_nerf.py:85-172) to CUDA fp32 tensors; call repack() after they change."""

    def __init__(self, params, uv_dims=2, out_ch=3):
        self.uv_dims = int(uv_dims)
        self.out_ch = int(out_ch)
        self.blob = None
        self.repack(params)

    def repack(self, params):
        lib = _cabi.lib()
        tensors = []
        for name in _cabi.PARAM_NAMES:
            if name not in params:
                raise KeyError("speech2lip_b200: parameter %r missing from the state dict" % name)
            t = params[name]
            t = t.detach() if isinstance(t, torch.Tensor) else torch.as_tensor(t)
            tensors.append(_f32c(t, name))
        e = self.uv_dims + 20 * self.uv_dims
        if tuple(tensors[_cabi.PARAM_NAMES.index("fc_uv.weight")].shape) != (256, e):
            raise ValueError("fc_uv.weight has shape %s, expected (256, %d) for uv_dims=%d"
                             % (tuple(params["fc_uv.weight"].shape), e, self.uv_dims))
        if tensors[_cabi.PARAM_NAMES.index("output_linear.weight")].shape[0] != self.out_ch:
            raise ValueError("output_linear.weight rows != out_ch=%d" % self.out_ch)
        dev = tensors[0].device
        if self.blob is None or self.blob.device != dev:
            self.blob = torch.empty(lib.s2l_blob_bytes(self.uv_dims, self.out_ch), dtype=torch.uint8, device=dev)
        arr = (C.c_void_p * _cabi.NUM_PARAMS)(*[t.data_ptr() for t in tensors])
        with torch.cuda.device(dev):
            _cabi.check(lib.s2l_pack_weights(arr, _ptr(self.blob), self.uv_dims, self.out_ch, _stream()),
                        "s2l_pack_weights")
        self._keepalive = tensors
        self._meta = None
        return self

    def meta(self):
        """What the pack kernels recorded (one device->host read per pack, cached): {"max_abs_weight", "n_saturating_weights"
        (tensor-core weights outside the fp16f8 domain |w| < 1024), "density_row_norm", "auto_fix_thr"}."""
        if self._meta is None:
            lib = _cabi.lib()
            mw, ns, dn, ft = C.c_float(), C.c_int32(), C.c_float(), C.c_float()
            with torch.cuda.device(self.blob.device):
                _cabi.check(lib.s2l_blob_meta(_ptr(self.blob), C.byref(mw), C.byref(ns), C.byref(dn), C.byref(ft), _stream()), "s2l_blob_meta")
            self._meta = {"max_abs_weight": mw.value, "n_saturating_weights": ns.value, "density_row_norm": dn.value,
                          "auto_fix_thr": ft.value}
        return self._meta

    def fp16f8_weights_ok(self):
        """True when every tensor-core weight lies in the validated fp16f8 domain (include/speech2lip_b200.h)."""
        m = self.meta()
        return m["n_saturating_weights"] == 0 and m["max_abs_weight"] < 1024.0

    @property
    def device(self):
        return self.blob.device


# --------------------------------------------------------------------------------------- functional API
def audio_encode(w, audio, frame_idx=None, want_latent=True, want_bias=True):
    """AudioNet (+ per-frame MLP biases).  audio [F,16,29] or [F,29,16] -> (latent [F,64], frame_bias [F,4,256])."""
    lib = _cabi.lib()
    audio = _f32c(audio, "audio")
    if audio.dim() != 3 or audio.shape[1] * audio.shape[2] != 16 * 29:
        raise ValueError("audio must be [F,16,29] or [F,29,16], got %s" % (tuple(audio.shape),))
    transposed = 1 if audio.shape[2] == 16 else 0         # tf_nerf.py:203-207
    F = audio.shape[0]
    idx = None
    if frame_idx is not None:
        idx = torch.as_tensor(frame_idx).to(device=audio.device, dtype=torch.int64).reshape(-1).contiguous()
        if idx.numel() != F:
            raise ValueError("frame_idx has %d entries for %d frames" % (idx.numel(), F))
    latent = torch.empty(F, 64, device=audio.device) if want_latent else None
    bias = torch.empty(F, 4, 256, device=audio.device) if want_bias else None
    with torch.cuda.device(audio.device):
        _cabi.check(lib.s2l_audio_encode_fwd(_ptr(w.blob), _ptr(audio), transposed, _ptr(idx), _ptr(latent), _ptr(bias),
                                             F, w.uv_dims, w.out_ch, _stream()), "s2l_audio_encode_fwd")
    return latent, bias


def rgb_forward_rows(w, x, time_idx=None):
    """General TalkingFace.rgb_forward contract (tf_nerf.py:225-285): x [N, uv_dims+64] with an arbitrary
    latent per row, time_pts -> position[0].  fp32 exact path."""
    lib = _cabi.lib()
    x = _f32c(x, "uv_audio_pts")
    if x.dim() != 2 or x.shape[1] != w.uv_dims + 64:
        raise ValueError("uv_audio_pts must be [N,%d], got %s" % (w.uv_dims + 64, tuple(x.shape)))
    out = torch.empty(x.shape[0], w.out_ch, device=x.device)
    has_time = time_idx is not None
    with torch.cuda.device(x.device):
        _cabi.check(lib.s2l_rgb_forward_rows(_ptr(w.blob), _ptr(x), x.shape[0], int(time_idx) if has_time else 0,
                                             1 if has_time else 0, _ptr(out), w.uv_dims, w.out_ch, _stream()),
                    "s2l_rgb_forward_rows")
    return out


def rows_constant(x, col0, ncols):
    """True when every row of the 2-D (or flattened-to-2-D) CUDA tensor x equals row 0 in columns [col0, col0+ncols)
    (one small kernel + one host sync)."""
    lib = _cabi.lib()
    x = _f32c(x, "x")
    x2 = x.reshape(x.shape[0], -1)
    flag = torch.empty(1, dtype=torch.int32, device=x.device)
    with torch.cuda.device(x.device):
        _cabi.check(lib.s2l_rows_differ(_ptr(x2), x2.shape[0], x2.shape[1], int(col0), int(ncols), _ptr(flag), _stream()),
                    "s2l_rows_differ")
    return int(flag.item()) == 0


def rows_differ_or(x, col0, ncols, flag):
    """Device-side, no host sync: sets the int32 device tensor `flag` to 1 when some row of x differs from row 0 in columns
    [col0, col0+ncols); never clears it (a sticky indicator checked later)."""
    lib = _cabi.lib()
    x = _f32c(x, "x")
    x2 = x.reshape(x.shape[0], -1)
    with torch.cuda.device(x.device):
        _cabi.check(lib.s2l_rows_differ_or(_ptr(x2), x2.shape[0], x2.shape[1], int(col0), int(ncols), _ptr(flag), _stream()),
                    "s2l_rows_differ_or")


def audio_merge_auto(w, audio, scratch):
    """TalkingFace.audio_merge_forward for a possibly tiled batch (inference.py:144), decided on the device: no host sync."""
    lib = _cabi.lib()
    audio = _f32c(audio, "audio")
    if audio.dim() != 3 or audio.shape[1] * audio.shape[2] != 16 * 29:
        raise ValueError("audio must be [B,16,29] or [B,29,16], got %s" % (tuple(audio.shape),))
    B = audio.shape[0]
    latent = torch.empty(B, 64, device=audio.device)
    with torch.cuda.device(audio.device):
        _cabi.check(lib.s2l_audio_merge_auto(_ptr(w.blob), _ptr(audio), 1 if audio.shape[2] == 16 else 0, B, _ptr(latent),
                                             _ptr(scratch), _stream()), "s2l_audio_merge_auto")
    return latent


def rgb_forward_auto(w, x, time_pts, precision, scratch):
    """TalkingFace.rgb_forward without a host sync: constant-latent rows -> fused tensor-core MLP in `precision`, arbitrary
    latents -> general fp32 kernel; the choice is made on the device.  time_pts: tensor / int / None (only element 0 is used,
    tf_nerf.py:439)."""
    lib = _cabi.lib()
    x = _f32c(x, "uv_audio_pts")
    out = torch.empty(x.shape[0], w.out_ch, device=x.device)
    tdev = None
    if time_pts is not None:
        tdev = torch.as_tensor(time_pts).reshape(-1)[:1].to(device=x.device, dtype=torch.int64, non_blocking=True)
    with torch.cuda.device(x.device):
        _cabi.check(lib.s2l_rgb_forward_auto(_ptr(w.blob), _ptr(x), x.shape[0], _ptr(tdev), _ptr(out), w.uv_dims, w.out_ch,
                                             _cabi.PRECISIONS[precision], _ptr(scratch), _stream()), "s2l_rgb_forward_auto")
    return out


def rgb_forward_const_latent(w, x, time_idx=None, precision="bf16x3"):
    """TalkingFace.rgb_forward for the caller pattern of inference.py:144-158 — every row of x [N, uv_dims+64] carries
    the SAME latent (checked by the caller): per-frame constants from row 0's latent, then the fused tensor-core MLP
    on the N explicit points."""
    lib = _cabi.lib()
    x = _f32c(x, "uv_audio_pts")
    N = x.shape[0]
    bias = torch.empty(1, 4, 256, device=x.device)
    idx = None if time_idx is None else torch.tensor([int(time_idx)], dtype=torch.int64, device=x.device)
    lat0 = x[0, w.uv_dims:]                                    # view: 64 contiguous floats
    with torch.cuda.device(x.device):
        _cabi.check(lib.s2l_latent_bias_fwd(_ptr(w.blob), _ptr(lat0), 64, _ptr(idx), _ptr(bias), 1, _stream()),
                    "s2l_latent_bias_fwd")
    pts = x[:, :w.uv_dims].contiguous().view(1, N, w.uv_dims)
    return mlp_points(w, bias, pts, precision)[0]


def mlp_points(w, frame_bias, pts, precision="bf16x3"):
    """rgb_forward on explicit points with per-frame-constant latent: pts [F,P,uv_dims] -> raw [F,P,out_ch]."""
    lib = _cabi.lib()
    pts = _f32c(pts, "pts")
    F, P = pts.shape[0], pts.shape[1]
    g = S2LGeom(n_frames=F, height=0, width=0, n_samples=0, pts_mode=_cabi.PTS_EXPLICIT, uv_dims=w.uv_dims,
                out_ch=w.out_ch, z_per_ray=0, rays_per_frame_shared=0, pts_per_frame=P, eps_shift=0.0)
    out = torch.empty(F, P, w.out_ch, device=pts.device)
    with torch.cuda.device(pts.device):
        _cabi.check(lib.s2l_mlp_fwd(_ptr(w.blob), C.byref(g), _ptr(frame_bias), _ptr(pts), None, None, None, _ptr(out),
                                    _cabi.PRECISIONS[precision], _stream()), "s2l_mlp_fwd")
    return out


def density2outputs(raw, z_vals, rays_d):
    """rendering.py:30-62 (raw_noise_std = 0): raw [R,S,4], z_vals [R,S] or [S], rays_d [R,3]."""
    lib = _cabi.lib()
    raw = _f32c(raw, "raw")
    z_vals = _f32c(z_vals, "z_vals")
    rays_d = _f32c(rays_d, "rays_d")
    R, S = raw.shape[0], raw.shape[1]
    if raw.shape[2] != 4:
        raise ValueError("raw must be [R,S,4]")
    rgb = torch.empty(R, 3, device=raw.device)
    weights = torch.empty(R, S, device=raw.device)
    depth = torch.empty(R, device=raw.device)
    with torch.cuda.device(raw.device):
        _cabi.check(lib.s2l_composite_fwd(_ptr(raw), _ptr(z_vals), 1 if z_vals.dim() == 2 else 0, _ptr(rays_d), R, 0, S,
                                          _ptr(rgb), _ptr(weights), _ptr(depth), _stream()), "s2l_composite_fwd")
    return rgb, weights, depth


def get_rays(H, W, focal, c2w):
    """src/common.py:12-21: c2w [3,4] (or [4,4]) CUDA tensor -> rays_o, rays_d [H,W,3]."""
    lib = _cabi.lib()
    c = _f32c(c2w, "c2w")[:3, :4].contiguous()
    ro = torch.empty(H, W, 3, device=c.device)
    rd = torch.empty(H, W, 3, device=c.device)
    with torch.cuda.device(c.device):
        _cabi.check(lib.s2l_get_rays(_ptr(c), H, W, float(focal), _ptr(ro), _ptr(rd), _stream()), "s2l_get_rays")
    return ro, rd


def post_fusion_compose(rgb_lip, face_canonical, rgb_gt, mask_lip_canonical, coord, lefttop_x, lefttop_y,
                        paste_shift=True, expand_pad=-1, want_canonical=True):
    """Pre-UNet part of post_fusion2_onlylip_light (tf_nerf.py:334-386, inference branch) as one kernel.
    Returns (fused [B,3,Hf,Wf] NCHW — the UNet input, merged_canonical [B,h,w,3] or None)."""
    lib = _cabi.lib()
    lip, face, gt = _f32c(rgb_lip, "rgb_lip"), _f32c(face_canonical, "rgb_face_canonical"), _f32c(rgb_gt, "rgb_gt")
    mask, coord = _f32c(mask_lip_canonical, "mask_lip_canonical"), _f32c(coord, "coord")
    B, lh, lw = lip.shape[0], lip.shape[1], lip.shape[2]
    h, w = face.shape[1], face.shape[2]
    Hf, Wf = coord.shape[1], coord.shape[2]
    if mask.shape[-1] == 1:
        mask = mask.expand(-1, -1, -1, 3).contiguous()
    if tuple(mask.shape) != (B, h, w, 3) or tuple(gt.shape) != (B, Hf, Wf, 3) or tuple(face.shape) != (B, h, w, 3) \
            or tuple(coord.shape) != (B, Hf, Wf, 2) or lip.shape[-1] != 3:
        raise ValueError("post_fusion_compose: inconsistent shapes lip %s face %s gt %s mask %s coord %s" % (
            tuple(lip.shape), tuple(face.shape), tuple(gt.shape), tuple(mask.shape), tuple(coord.shape)))
    fused = torch.empty(B, 3, Hf, Wf, device=lip.device)
    canon = torch.empty(B, h, w, 3, device=lip.device) if want_canonical else None
    with torch.cuda.device(lip.device):
        _cabi.check(lib.s2l_post_fusion_compose(_ptr(lip), _ptr(face), _ptr(gt), _ptr(mask), _ptr(coord), B, lh, lw, h, w, Hf, Wf,
                                                int(lefttop_x), int(lefttop_y), 1 if paste_shift else 0, int(expand_pad),
                                                _ptr(fused), _ptr(canon), _stream()), "s2l_post_fusion_compose")
    return fused, canon


def audio_windows(logits):
    """deepspeech_features.py:65-75: DeepSpeech logits [T,29] -> 16-step windows with stride 2, [ceil(T/2),16,29]."""
    lib = _cabi.lib()
    x = _f32c(logits, "logits").reshape(-1, 29)
    T = x.shape[0]
    out = torch.empty((T + 1) // 2, 16, 29, device=x.device)
    with torch.cuda.device(x.device):
        _cabi.check(lib.s2l_audio_windows(_ptr(x), T, _ptr(out), _stream()), "s2l_audio_windows")
    return out


def frames_to_bgr8(rgb, out=None):
    """inference.py:173-178 output staging: [...,3] fp32 RGB -> uint8 BGR exactly as cv2.imwrite(img * 255) stores it."""
    lib = _cabi.lib()
    x = _f32c(rgb, "rgb")
    if x.shape[-1] != 3:
        raise ValueError("rgb must have 3 channels last")
    if out is None:
        out = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    elif out.dtype != torch.uint8 or tuple(out.shape) != tuple(x.shape) or not out.is_contiguous() or out.device != x.device:
        raise ValueError("out must be a contiguous uint8 tensor of rgb's shape on rgb's device")
    with torch.cuda.device(x.device):
        _cabi.check(lib.s2l_frames_to_bgr8(_ptr(x), x.numel() // 3, _ptr(out), _stream()), "s2l_frames_to_bgr8")
    return out


class LipRenderer:
    """Batched frame renderer over one PackedWeights blob.

    render_frames(audio [F,16,29], index [F], H, W, mode=...) -> rgb [F,H,W,3]
      mode="plain"      1 MLP eval / pixel                       (inference.py:144-159)
      mode="ensemble4"  4 jittered taps / pixel, area blend      (training.py:158-251), eps_shift explicit
      mode="volumetric" S samples / ray + alpha compositing      (uv_dims=3/out_ch=4 model)
    Each frame uses its own index for the time code, i.e. the result equals calling the reference once
    per frame (the reference's time PE only ever sees position[0], tf_nerf.py:439).
    """

    def __init__(self, weights, precision="bf16x3"):
        if precision not in _cabi.PRECISIONS and precision != "auto":
            raise ValueError("precision must be one of %s or 'auto'" % sorted(_cabi.PRECISIONS))
        self.w = weights
        self.precision = precision            # "auto": fp16f8 inside its validated domain, else bf16x3
        self._scratch = None
        self._act_max = None                  # largest hidden activation seen by probe_fp16f8_domain (None: never probed)

    def _resolve_precision(self, name):
        """fp16f8 is only used inside its validated domain (weights |w| < 1024 — recorded at pack time — and, once probed,
        activations |a| < 4096): outside it the fp8 correction terms saturate silently, so the renderer switches to bf16x3
        and says so (never silently)."""
        if name not in ("fp16f8", "auto"):
            return name
        why = None
        if not self.w.fp16f8_weights_ok():
            m = self.w.meta()
            why = "max |w| = %.3g, %d tensor-core weights >= 1024" % (m["max_abs_weight"], m["n_saturating_weights"])
        elif self._act_max is not None and not self._act_max < 4096.0:
            why = "probed hidden activations reach %.3g >= 4096" % self._act_max
        if why is None:
            return "fp16f8"
        key = (id(self.w.blob), why)
        if self.__dict__.get("_warned") != key:
            self._warned = key
            import warnings
            warnings.warn("speech2lip_b200: model outside the validated fp16f8 domain (%s): using bf16x3 instead" % why, RuntimeWarning)
        return "bf16x3"

    def probe_fp16f8_domain(self, audio, index, pts):
        """Runs the EXACT fp32 path (with saved activations) on sample points pts [N, uv_dims] of the first frame and records the
        largest hidden activation; later fp16f8 / auto renders fall back to bf16x3 if it leaves the domain.  Returns
        {"max_activation", "max_abs_weight", "ok"}."""
        lib = _cabi.lib()
        audio = _f32c(audio, "audio")
        lat, _ = audio_encode(self.w, audio[:1], None, want_bias=False)
        pts = _f32c(pts, "pts").reshape(-1, self.w.uv_dims)
        x = torch.cat([pts, lat.expand(pts.shape[0], -1)], -1).contiguous()
        N = x.shape[0]
        out = torch.empty(N, self.w.out_ch, device=x.device)
        acts = torch.empty(10, N, 256, device=x.device)
        t = int(torch.as_tensor(index).reshape(-1)[0])
        with torch.cuda.device(x.device):
            _cabi.check(lib.s2l_rgb_forward_rows_train(_ptr(self.w.blob), _ptr(x), N, t, 1, _ptr(out), _ptr(acts), self.w.uv_dims,
                                                       self.w.out_ch, _stream()), "s2l_rgb_forward_rows_train")
        self._act_max = float(acts.abs().max().item())
        m = self.w.meta()
        return {"max_activation": self._act_max, "max_abs_weight": m["max_abs_weight"],
                "ok": self.w.fp16f8_weights_ok() and self._act_max < 4096.0}

    def _scratch_for(self, geom, device, prec, want_aux):
        need = _cabi.lib().s2l_render_scratch_bytes(C.byref(geom), prec, 1 if want_aux else 0)
        if self._scratch is None or self._scratch.numel() < need or self._scratch.device != device:
            self._scratch = torch.empty(need, dtype=torch.uint8, device=device)
        return self._scratch

    def render_frames(self, audio, index, H, W, mode="plain", eps_shift=0.0, n_samples=None, rays_o=None,
                      rays_d=None, z_vals=None, out=None, return_aux=False, precision=None, sample_chunks=1, term_thr=0.0,
                      fix_thr=0.0):
        lib = _cabi.lib()
        audio = _f32c(audio, "audio")
        if audio.dim() != 3 or tuple(audio.shape[1:]) != (16, 29):
            raise ValueError("audio must be [F,16,29], got %s" % (tuple(audio.shape),))
        F = audio.shape[0]
        dev = audio.device
        idx = torch.as_tensor(index).to(device=dev, dtype=torch.int64).reshape(-1).contiguous()
        if idx.numel() != F:
            raise ValueError("index has %d entries for %d frames" % (idx.numel(), F))
        prec = _cabi.PRECISIONS[self._resolve_precision(precision or self.precision)]
        eps_pf = None
        if isinstance(eps_shift, torch.Tensor) and eps_shift.numel() > 1:      # one draw per frame (sync-window render)
            eps_pf = eps_shift.to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
            if eps_pf.numel() != F:
                raise ValueError("eps_shift has %d entries for %d frames" % (eps_pf.numel(), F))
            eps_shift = 0.0
        g = S2LGeom(n_frames=F, height=H, width=W, n_samples=0, pts_mode=_cabi.PTS_GRID, uv_dims=self.w.uv_dims,
                    out_ch=self.w.out_ch, z_per_ray=0, rays_per_frame_shared=0, pts_per_frame=0, eps_shift=float(eps_shift),
                    eps_per_frame=None if eps_pf is None else eps_pf.data_ptr(), sample_chunks=int(sample_chunks),
                    term_thr=float(term_thr), fix_thr=float(fix_thr))
        weights = depth = None
        if mode == "plain":
            g.pts_mode = _cabi.PTS_GRID
        elif mode == "ensemble4":
            g.pts_mode = _cabi.PTS_GRID_ENS4
        elif mode == "volumetric":
            g.pts_mode = _cabi.PTS_RAYS
            if rays_o is None or rays_d is None or z_vals is None:
                raise ValueError("volumetric mode needs rays_o, rays_d and z_vals")
            rays_o = _f32c(rays_o, "rays_o").reshape(-1, 3)
            rays_d = _f32c(rays_d, "rays_d").reshape(-1, 3)
            z_vals = _f32c(z_vals, "z_vals")
            R = H * W
            if rays_o.shape[0] == R:
                g.rays_per_frame_shared = 1
            elif rays_o.shape[0] == R * F:
                g.rays_per_frame_shared = 0
            else:
                raise ValueError("rays must be [H*W,3] or [F*H*W,3]")
            if rays_d.shape != rays_o.shape:
                raise ValueError("rays_o / rays_d shape mismatch")
            if z_vals.dim() == 1:
                g.z_per_ray = 0
                g.n_samples = z_vals.shape[0]
            else:
                z_vals = z_vals.reshape(-1, z_vals.shape[-1])
                if z_vals.shape[0] != R * F:
                    raise ValueError("z_vals must be [S] or [F*H*W,S]")
                g.z_per_ray = 1
                g.n_samples = z_vals.shape[1]
            if n_samples is not None and n_samples != g.n_samples:
                raise ValueError("n_samples does not match z_vals")
            if return_aux:
                weights = torch.empty(F, R, g.n_samples, device=dev)
                depth = torch.empty(F, R, device=dev)
        else:
            raise ValueError("unknown mode %r" % (mode,))
        if out is None:
            out = torch.empty(F, H, W, 3, device=dev)
            if F == 0 or H * W == 0:
                return (out, weights, depth) if return_aux else out
        else:
            _need_cuda(out, "out")
            if tuple(out.shape) != (F, H, W, 3) or out.dtype != torch.float32 or not out.is_contiguous():
                raise ValueError("out must be a contiguous fp32 [F,H,W,3] tensor")
        scratch = self._scratch_for(g, dev, prec, return_aux)
        self._last_geom = g
        with torch.cuda.device(dev):
            _cabi.check(lib.s2l_render_frames(_ptr(self.w.blob), C.byref(g), _ptr(audio), _ptr(idx), _ptr(rays_o),
                                              _ptr(rays_d), _ptr(z_vals), _ptr(out), _ptr(weights), _ptr(depth),
                                              _ptr(scratch), prec, _stream()), "s2l_render_frames")
        if return_aux:
            return out, weights, depth
        return out

    def last_render_counts(self):
        """Counters the last volumetric render left in the scratch buffer (one host sync): {"alive": [C-1][F] rays of each
        frame still alive when sample chunk k = 1..C-1 started (early ray termination), "reevaluated": [F] rays whose last
        sample was re-evaluated in fp32 (near-zero density, include/speech2lip_b200.h fix_thr)}."""
        g = self.__dict__.get("_last_geom")
        if g is None or g.pts_mode != _cabi.PTS_RAYS or self._scratch is None:
            raise RuntimeError("no volumetric render has run on this renderer")
        off = _cabi.lib().s2l_render_counts_offset(C.byref(g))
        Cn, F = max(int(g.sample_chunks), 1), int(g.n_frames)
        cnt = self._scratch[off:off + (Cn + 1) * F * 4].view(torch.int32).view(Cn + 1, F).cpu()
        return {"alive": cnt[1:Cn], "reevaluated": cnt[Cn]}

    def render_sync_window(self, audio_window, index, total_frame, H, W, eps_shift):
        """The sync-expert loss window (training.py:500-525): T consecutive lip frames, each through the 4-tap
        ensemble with its own audio window, time index min(index + t, total_frame - 1) and its own eps_shift draw —
        ONE launch instead of T predict_lip_image calls.  audio_window [T,16,29], eps_shift [T] -> [T,H,W,3]."""
        T = audio_window.shape[0]
        idx = torch.clamp(torch.arange(T, device=audio_window.device) + int(index), max=int(total_frame) - 1)
        return self.render_frames(audio_window, idx, H, W, mode="ensemble4", eps_shift=torch.as_tensor(eps_shift))

    def render_sequence_host(self, audio_host, index_host, H, W, frames_per_step=64, out="bgr8", out_host=None, **kw):
        """Streams a whole T-frame sequence HOST -> HOST (the per-frame staging around the path):
        per chunk of `frames_per_step` frames  H2D (windows, indices) -> render -> [uint8 BGR staging, exactly the bytes
        cv2.imwrite(img * 255) stores, inference.py:173-178] -> D2H,  with the D2H of chunk i on a copy stream while chunk
        i+1 renders (double-buffered device outputs).  audio_host [T,16,29] / index_host [T] should be pinned;
        returns out_host [T,H,W,3] (uint8 BGR for out="bgr8", fp32 RGB for out="rgb32"), pinned if allocated here."""
        if out not in ("bgr8", "rgb32"):
            raise ValueError("out must be 'bgr8' or 'rgb32'")
        dev = self.w.device
        T = audio_host.shape[0]
        Fs = max(1, min(int(frames_per_step), max(T, 1)))
        dt = torch.uint8 if out == "bgr8" else torch.float32
        if out_host is None:
            out_host = torch.empty((T, H, W, 3), dtype=dt, pin_memory=True)
        elif tuple(out_host.shape) != (T, H, W, 3) or out_host.dtype != dt or out_host.is_cuda:
            raise ValueError("out_host must be a host %s tensor [T,H,W,3]" % dt)
        if T == 0:
            return out_host
        key = (Fs, H, W, out, dev)
        st = self.__dict__.get("_seq_state")
        if st is None or st["key"] != key:
            st = {"key": key, "copy": torch.cuda.Stream(dev),
                  "rgb": [torch.empty(Fs, H, W, 3, device=dev) for _ in range(2)],
                  "u8": [torch.empty(Fs, H, W, 3, dtype=torch.uint8, device=dev) for _ in range(2)] if out == "bgr8" else None,
                  "rendered": [torch.cuda.Event() for _ in range(2)], "copied": [torch.cuda.Event() for _ in range(2)]}
            self._seq_state = st
        compute = torch.cuda.current_stream(dev)
        for b in range(2):
            st["copied"][b].record(compute)                  # nothing in flight on the buffers yet
        for c, s0 in enumerate(range(0, T, Fs)):
            n, b = min(Fs, T - s0), c & 1
            compute.wait_event(st["copied"][b])              # the D2H that last read buffer b has finished
            a = audio_host[s0:s0 + n].to(dev, non_blocking=True)
            i = index_host[s0:s0 + n].to(dev, non_blocking=True)
            rgb = self.render_frames(a, i, H, W, out=st["rgb"][b][:n], **kw)
            src = frames_to_bgr8(rgb, out=st["u8"][b][:n]) if out == "bgr8" else rgb
            st["rendered"][b].record(compute)
            with torch.cuda.stream(st["copy"]):
                st["copy"].wait_event(st["rendered"][b])
                out_host[s0:s0 + n].copy_(src, non_blocking=True)
                st["copied"][b].record(st["copy"])
        st["copy"].synchronize()
        return out_host

    def render_frames_host(self, audio_host, index_host, H, W, out_host=None, **kw):
        """End-to-end call on HOST buffers: H2D of the audio windows / indices (pinned -> async), render,
        D2H of the frames.  This is what bench.py times as `e2e`."""
        dev = self.w.device
        a = audio_host.to(dev, non_blocking=True)
        i = index_host.to(dev, non_blocking=True)
        rgb = self.render_frames(a, i, H, W, **kw)
        if out_host is None:
            out_host = torch.empty(rgb.shape, dtype=rgb.dtype, pin_memory=True)
        out_host.copy_(rgb, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return out_host
