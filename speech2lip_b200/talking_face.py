"""Drop-in for the reference's `src.face_simple.models.tf_nerf.TalkingFace` (tf_nerf.py:12-389).

Same constructor signature, attribute names, parameter names/shapes (115 state-dict keys for may.yaml)
and method signatures, so `CheckpointIO.load` round-trips reference checkpoints and inference.py /
train.py can call it unchanged (INTEGRATION.md shows the one-line import switch).

Hot path (this repo's scope): audio_merge_forward and rgb_forward run the CUDA kernels behind the C ABI.
Outside the hot path (SURVEY §8(f) "next"): post_fusion2_onlylip and the UNet are plain PyTorch modules
kept only so that the checkpoint layout and the callers keep working.
"""
import os
import random

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import renderer as R

_HOT_PREFIXES = ("encoder_conv.", "encoder_fc1.", "fc_uv", "fc_audio", "fc_time", "pts_linears.", "output_linear.")


def _conv_bn_relu_x2(cin, cout, cmid=None):
    cmid = cmid or cout
    return nn.Sequential(nn.Conv2d(cin, cmid, 3, padding=1, bias=False), nn.BatchNorm2d(cmid), nn.ReLU(inplace=True),
                         nn.Conv2d(cmid, cout, 3, padding=1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class _Block(nn.Module):
    def __init__(self, cin, cout, cmid=None):
        super().__init__()
        self.double_conv = _conv_bn_relu_x2(cin, cout, cmid)

    def forward(self, x):
        return self.double_conv(x)


class _Down(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), _Block(cin, cout))

    def forward(self, x):
        return self.maxpool_conv(x)


class _Up(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.up = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)
        self.conv = _Block(cin, cout, cin // 2)

    def forward(self, lo, skip):
        lo = self.up(lo)
        dy, dx = skip.shape[2] - lo.shape[2], skip.shape[3] - lo.shape[3]
        lo = F.pad(lo, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])
        return self.conv(torch.cat([skip, lo], dim=1))


class _Head(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, kernel_size=1)

    def forward(self, x):
        return self.conv(x)


class PostFusionUNet(nn.Module):
    """State-dict-compatible stand-in for models/SimpleUnetLight.py:82-111 (2-down / 2-up UNet,
    64-128-128 channels).  Out of the hot path; cuDNN through PyTorch."""

    def __init__(self, cfg=None, n_channels=3, n_classes=3):
        super().__init__()
        c = 64
        self.inc = _Block(n_channels, c)
        self.down1 = _Down(c, 2 * c)
        self.down2 = _Down(2 * c, 2 * c)
        self.up1 = _Up(4 * c, c)
        self.up2 = _Up(2 * c, c)
        self.outc = _Head(c, n_classes)

    def forward(self, x, x_level1=None, x_level2=None):
        a = self.inc(x)
        b = self.down1(a)
        c = self.down2(b)
        return self.outc(self.up2(self.up1(c, b), a))


class TalkingFace(nn.Module):
    def __init__(self, device, cfg, mode='train',
                 use_viewdirs=False, coord_merge_audio=False,
                 W=256, D=8, coord_D=4, skips=[4],
                 uv_audio_dims=66, uv_dims=2, audio_dims=29, head_pose_dims=3,
                 time_multires=10,
                 output_ch=3, **args):
        super().__init__()
        m = cfg['model']
        unsupported = [k for k in ('use_head_pose', 'use_lms', 'use_text', 'use_audio_mel', 'use_attention') if m.get(k)]
        if unsupported or not m.get('audio_net', True) or not m.get('audio_not_embed', True) or not m.get('use_audio', True) \
                or m.get('MLP_version') != 'v2' or not m.get('use_time', True) or W != 256 or D != 8 or list(skips) != [4] \
                or int(m.get('uv_embed', 10)) != 10 or time_multires != 10:
            raise NotImplementedError(
                "speech2lip_b200.TalkingFace implements the may.yaml hot-path configuration (audio_net, "
                "audio_not_embed, MLP v2, use_time, W=256, D=8, skips=[4], uv_embed=10); got unsupported options %s"
                % (unsupported,))
        self.cfg = cfg
        self.device = device
        self.use_viewdirs = use_viewdirs
        self.coord_merge_audio = coord_merge_audio
        self.uv_audio_dims = uv_audio_dims
        self.use_attention = m['use_attention']
        self.use_audio_net = m['audio_net']
        self.use_uv_audio_sep = m.get('use_uv_audio_sep', True)
        self.audio_not_embed = m['audio_not_embed']
        self.skips = skips
        self.uv_dims = uv_dims
        self.audio_dims = 64                                 # tf_nerf.py:63-64 (audio_net)
        self.head_pose_dims = head_pose_dims
        self.use_audio = m['use_audio']
        self.N_sample = cfg['training']['n_sample_points']   # read, never used (tf_nerf.py:44)
        self.use_head_pose = m['use_head_pose']
        self.use_head_pose_net = m.get('use_head_pose_net', False)
        self.use_time = m['use_time']
        self.use_post_fusion = m['use_post_fusion']
        self.use_lms = m['use_lms']
        self.use_text = m['use_text']
        self.data_path = cfg['data']['path']
        self.expand_lip_mask = m['expand_lip_mask']
        self.MLP_version = m['MLP_version']
        self.output_ch = output_ch
        if self.use_post_fusion:
            self.use_light_unet = m['use_light_unet']
            self.use_resnet = m.get('use_resnet', False)
            self.post_fusion_channel = m['post_fusion_channel']
            self.post_fusion_unet = PostFusionUNet(cfg=cfg, n_channels=self.post_fusion_channel).to(device)

        lrelu = lambda: nn.LeakyReLU(0.02, True)
        self.encoder_conv = nn.Sequential(
            nn.Conv1d(29, 32, 3, stride=2, padding=1), lrelu(), nn.Conv1d(32, 32, 3, stride=2, padding=1), lrelu(),
            nn.Conv1d(32, 64, 3, stride=2, padding=1), lrelu(), nn.Conv1d(64, 64, 3, stride=2, padding=1), lrelu())
        self.encoder_fc1 = nn.Sequential(nn.Linear(64, 64), lrelu(), nn.Linear(64, self.audio_dims))
        # dead in every forward of the reference, but part of the checkpoint layout (tf_nerf.py:130-135)
        self.coord_linears = nn.ModuleList([nn.Linear(2, W)] + [nn.Linear(W, W) for _ in range(coord_D - 1)]
                                           + [nn.Linear(W, self.audio_dims)])
        e = uv_dims + 2 * 10 * uv_dims
        self.output_linear = nn.Linear(W, output_ch)
        self.fc_uv = nn.Linear(e, W)
        self.fc_uv_skip = nn.Linear(e, W)
        self.fc_audio = nn.Linear(self.audio_dims, W)
        self.fc_audio_skip = nn.Linear(self.audio_dims, W)
        self.fc_time = nn.Linear(2 * time_multires, W)
        self.fc_time_skip = nn.Linear(2 * time_multires, W)
        self.pts_linears = nn.ModuleList([nn.Linear(W, W)] + [nn.Linear(W, W) if i not in skips else nn.Linear(2 * W, W)
                                                              for i in range(D - 1)])
        if m['use_canonical_depth']:
            if 'canonical_depth_init_path' in m:
                init = torch.from_numpy(np.load(m['canonical_depth_init_path'])).float()
                depth = init.clone()
                depth[depth == 0] = depth[depth > 0].mean()
                import cv2
                mask = cv2.imread(os.path.join(cfg['data']['path'], 'canonical_head_mask.jpg')) / 255
                mask[mask > 0] = 1
                depth[torch.from_numpy(mask[:, :, 0]).int() == 0] = 0
                depth[init > 0] = init[init > 0]
            else:
                depth = torch.randn((m['canonical_depth_height'], m['canonical_depth_width']))
            self.canonical_depth_head = nn.Parameter(depth, requires_grad=True)
        self._packed = None
        self._packed_key = None
        # Drop-in acceleration of the unmodified callers (inference.py:144-158): inputs whose rows are all identical (the
        # tiled audio window, the tiled latent columns) are detected with a compare kernel and routed through the
        # per-frame / tensor-core path.  `dropin_precision` must be a parity mode ("bf16x3" or "fp16f8"); "fp32" keeps
        # the exact CUDA-core kernel; set `dropin_fast_path = False` to always take the general per-row path.
        self.dropin_fast_path = os.environ.get("S2L_DROPIN_FAST", "1") != "0"
        self.dropin_precision = os.environ.get("S2L_DROPIN_PRECISION", "bf16x3")
        if self.dropin_precision not in ("bf16x3", "fp16f8", "fp32"):
            raise ValueError("S2L_DROPIN_PRECISION must be bf16x3, fp16f8 (parity modes) or fp32, got %r" % self.dropin_precision)
        self.dropin_min_rows = 1024
        # Per-call training arithmetic (rgb_forward under autograd): "fp32" = exact kernels (default, what the reference computes),
        # "bf16" = the tensor-core training kernels for calls whose rows share one latent (S2L_TRAIN_PRECISION).  The batched
        # render_lip_train is always bf16.
        self.train_precision = os.environ.get("S2L_TRAIN_PRECISION", "fp32")
        if self.train_precision not in ("fp32", "bf16"):
            raise ValueError("S2L_TRAIN_PRECISION must be fp32 or bf16, got %r" % self.train_precision)

    # ------------------------------------------------------------------ packed weights (kernel layout)
    def _hot_params(self):
        # Parameter objects are stable (load_state_dict / optimizers update them in place), so the name -> tensor
        # map is built once; rebuilt only if a parameter object was replaced (e.g. module.to(dtype) tricks).
        cache = self.__dict__.get("_hot_cache")
        if cache is None or cache["fc_uv.weight"] is not self.fc_uv.weight or cache["pts_linears.7.bias"] is not self.pts_linears[7].bias:
            cache = {k: v for k, v in self.named_parameters() if k.startswith(_HOT_PREFIXES)}
            self.__dict__["_hot_cache"] = cache
        return cache

    def packed_weights(self):
        """PackedWeights for the current parameter values; re-packed when any hot-path tensor changed
        (in-place optimizer steps and load_state_dict bump tensor._version)."""
        hp = self._hot_params()
        # in-place updates (optimizer steps, load_state_dict, p.add_()) bump tensor._version; a replaced storage (module.to())
        # changes data_ptr.  `.data` edits bypass the version counter: call invalidate_packed() after those.
        vals = self.__dict__.get("_hot_vals")
        if vals is None or vals[0] is not hp["encoder_conv.0.weight"]:
            vals = self.__dict__["_hot_vals"] = list(hp.values())
        key = [v._version for v in vals]
        key.append(vals[0].data_ptr())
        key.append(vals[-1].data_ptr())
        if self._packed is None or key != self._packed_key:
            if self.__dict__.get("_tc_viol") is not None:
                self.check_train_rows()               # once per optimizer step: the deferred precondition check of the bf16 per-call path
            if self._packed is None:
                self._packed = R.PackedWeights(hp, self.uv_dims, self.output_ch)
            else:
                self._packed.repack(hp)
            self._packed_key = key
        return self._packed

    def invalidate_packed(self):
        """Forces the next call to re-pack the kernel-layout weights (needed only after edits through `param.data`, which do
        not bump the version counter packed_weights() watches)."""
        self._packed_key = None

    def _needs_grad(self, *tensors):
        return torch.is_grad_enabled() and (any(p.requires_grad for p in self._hot_params().values())
                                            or any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors))

    # ------------------------------------------------------------------ hot path
    def audio_merge_forward(self, audio):
        """tf_nerf.py:197-213.  audio: [B,16,29] or [B,29,16] -> [B,64]."""
        if self._needs_grad(audio):
            if audio.is_cuda and not audio.requires_grad:
                # training: the inference kernel with saved activations + a hand-written backward (one CTA per frame); the
                # latent feeds the fused MLP's autograd.Function (rows contract or render_lip_train)
                from .autograd import audio_merge_forward_train
                return audio_merge_forward_train(self, audio)
            # a caller differentiating w.r.t. the audio window itself: plain autograd (Conv1d(k3,s2,p1) as unfold + matmul)
            x = audio if audio.shape[2] == 16 else audio.permute(0, 2, 1)
            for i in (0, 2, 4, 6):
                conv = self.encoder_conv[i]
                cols = F.pad(x, (1, 1)).unfold(2, 3, 2)                                  # [B,C,T/2,3]
                cols = cols.permute(0, 2, 1, 3).reshape(x.shape[0], cols.shape[2], -1)   # [B,T/2,C*3]
                x = F.leaky_relu(cols @ conv.weight.reshape(conv.weight.shape[0], -1).t() + conv.bias, 0.02).permute(0, 2, 1)
            return self.encoder_fc1(x.squeeze(-1))
        if self.dropin_fast_path and audio.is_cuda and audio.dim() == 3 and audio.shape[0] >= self.dropin_min_rows:
            # inference.py:144 tiles ONE window H*W times: a compare kernel decides ON THE DEVICE whether row 0's latent is
            # broadcast or every row is encoded (AudioNet is bit-invariant to batch tiling) — no host synchronisation
            return R.audio_merge_auto(self.packed_weights(), audio, self._dropin_scratch(audio.device))
        latent, _ = R.audio_encode(self.packed_weights(), audio, None, want_latent=True, want_bias=False)
        return latent

    def rgb_forward(self, uv_audio_pts, time_pts=None, head_pose_pts=None, rgb_pts=None, lms_pts=None, text_pts=None):
        """tf_nerf.py:225-285.  uv_audio_pts [N, uv_dims+64]; time_pts: only element 0 is used (tf_nerf.py:439)."""
        if self._needs_grad(uv_audio_pts):
            x = uv_audio_pts
            if (self.train_precision == "bf16" and isinstance(x, torch.Tensor) and x.is_cuda and x.dim() == 2 and self.uv_dims == 2
                    and self.output_ch == 3 and x.shape[1] == 66 and x.shape[0] >= self.dropin_min_rows and self._tc_rows_ok(x)):
                # opt-in: the call's rows share one latent (training.py:216-233) -> bf16 tensor-core forward / backward
                from .autograd import rgb_forward_train_tc
                return rgb_forward_train_tc(self, x, time_pts)
            t = None if time_pts is None else int(torch.as_tensor(time_pts).reshape(-1)[0].item())
            from .autograd import rgb_forward_train      # exact fp32: fused forward (saves activations) + fused dgrad kernel
            return rgb_forward_train(self, uv_audio_pts, t)
        x = uv_audio_pts
        if isinstance(x, torch.Tensor) and x.is_cuda and x.dim() == 2 and x.shape[1] == self.uv_dims + 64:
            # No host synchronisation: the time index stays on the device and a compare kernel decides there whether the rows
            # share one latent (inference.py:150-158: hoist the audio/time terms, fused tensor-core MLP in the parity mode
            # `dropin_precision`, ~1e-4 from the exact result) or carry arbitrary latents (general exact fp32 kernel).
            # Fewer than dropin_min_rows rows, or dropin_fast_path = False, always take the exact kernel.
            fast = self.dropin_fast_path and x.shape[0] >= self.dropin_min_rows
            return R.rgb_forward_auto(self.packed_weights(), x, time_pts, self.dropin_precision if fast else "fp32",
                                      self._dropin_scratch(x.device))
        t = None if time_pts is None else int(torch.as_tensor(time_pts).reshape(-1)[0].item())
        return R.rgb_forward_rows(self.packed_weights(), uv_audio_pts, t)

    # ---- precondition of the opt-in bf16 per-call training path: every row of a call carries the same latent
    _TC_SYNC_CALLS = int(os.environ.get("S2L_TC_SYNC_CALLS", "8"))       # (a huge value = always check synchronously)

    def _tc_rows_ok(self, x):
        """The first calls are checked synchronously (a caller that does not tile the latent is routed to the exact path and
        never enters the tensor-core one).  After that the compare kernel only sets a sticky device flag — no host
        synchronisation per call, which is what made the unmodified Trainer loop host-bound — and the flag is read once per
        training step (check_train_rows, called when the packed weights are refreshed after an optimizer step)."""
        n = self.__dict__.get("_tc_calls", 0)
        self.__dict__["_tc_calls"] = n + 1
        if n < self._TC_SYNC_CALLS:
            return R.rows_constant(x, 2, 64)
        viol = self.__dict__.get("_tc_viol")
        if viol is None or viol.device != x.device:
            viol = self.__dict__["_tc_viol"] = torch.zeros(1, dtype=torch.int32, device=x.device)
        R.rows_differ_or(x, 2, 64, viol)
        return True

    def check_train_rows(self):
        """Raises if any rgb_forward call since the last check violated the bf16 per-call path's precondition."""
        viol = self.__dict__.get("_tc_viol")
        if viol is not None and int(viol.item()) != 0:
            viol.zero_()
            self.__dict__["_tc_calls"] = 0
            raise RuntimeError("speech2lip_b200: train_precision='bf16' — an rgb_forward call since the last check had rows with "
                               "DIFFERENT latents; the tensor-core per-call path needs the tiled latent of Trainer.predict_lip_image "
                               "(training.py:216-233).  The outputs and gradients of those calls are invalid: use train_precision='fp32' "
                               "for such callers.")

    def _dropin_scratch(self, device):
        sc = self.__dict__.get("_dropin_sc")
        if sc is None or sc.device != device:
            sc = torch.empty(8192, dtype=torch.uint8, device=device)
            self.__dict__["_dropin_sc"] = sc
        return sc

    def render_lip_train(self, audio, index, H, W, eps_shift=None):
        """F lip frames through the 4-tap local ensemble in ONE differentiable launch sequence on tensor cores (bf16):
        what Trainer.predict_lip_image (training.py:158-251) computes per frame, for a batch of frames / the five-frame
        sync-expert window (training.py:500-548).  Gradients reach every hot-path parameter including AudioNet.
        audio [F,16,29], index [F] (time_pts of each frame) -> rgb [F,H,W,3]."""
        from .autograd import render_lip_train
        return render_lip_train(self, audio, index, H, W, eps_shift)

    def render_sync_window_train(self, audio_window, index, total_frame, H, W, eps_shift=None):
        """The sync-expert loss window WITH gradients (training.py:500-525): T consecutive lip frames, frame t rendered with
        audio_window[t], time index min(index + t, total_frame - 1) and its own eps_shift draw (drawn in frame order from the
        device RNG exactly as T predict_lip_image calls would) — one differentiable launch sequence instead of T x 4
        rgb_forward calls.  audio_window [T,16,29] -> [T,H,W,3]."""
        T = audio_window.shape[0]
        idx = torch.clamp(torch.arange(T) + int(index), max=int(total_frame) - 1)
        return self.render_lip_train(audio_window, idx, H, W, eps_shift)

    def renderer(self, precision="bf16x3"):
        """Batched frame renderer (not expressible through the reference's per-call contract)."""
        return R.LipRenderer(self.packed_weights(), precision)

    # ------------------------------------------------------------------ outside the hot path (PyTorch)
    def post_fusion2_onlylip(self, rgb_lip_warped, rgb_face_canonical, rgb_gt, mask_lip_canonical, lip_lefttop_x,
                             lip_lefttop_y, coord, use_canonical_space=False, change_pose=-1, mask_face_canonical=None,
                             wav2lip=None, mask_head_observed=None, use_post_fusion_blackaug=False):
        """tf_nerf.py:287-389 (light-UNet branch): paste lip crop into the canonical face, warp canonical ->
        observed with `coord`, optional black-hole augmentation, UNet refine.  [B,H,W,C] layouts."""
        if not self.use_light_unet:
            return None
        shifted_ds = any(s in self.data_path for s in ('macron', 'obama_adnerf', 'obama2_face_crop', 'may'))
        aug = use_post_fusion_blackaug and random.random() > 0.5        # same RNG draw as tf_nerf.py:369
        grad_on = torch.is_grad_enabled()
        lip_grad = grad_on and rgb_lip_warped.requires_grad
        other_grad = grad_on and any(isinstance(t, torch.Tensor) and t.requires_grad
                                     for t in (rgb_face_canonical, rgb_gt, coord, mask_lip_canonical))
        if rgb_lip_warped.is_cuda and not aug and not other_grad:
            # fused gather-blend kernel; the UNet stays cuDNN.  In training the lip crop is the only
            # input with a gradient (training.py:436-445, 525-539): the kernel then runs inside an autograd.Function whose
            # backward is the scatter kernel.  A gradient w.r.t. any other input (or the black-hole augmentation) takes the
            # differentiable PyTorch branch below.
            lw_ = rgb_lip_warped.shape[2]
            pad_ = -1
            if self.expand_lip_mask:
                pad_ = lw_ // 12 if 'obama2_face_crop' in self.data_path else lw_ // 5
            if lip_grad:
                from .autograd import PostFusionCompose
                fused, canon = PostFusionCompose.apply(rgb_lip_warped, rgb_face_canonical, rgb_gt, mask_lip_canonical, coord,
                                                       int(lip_lefttop_x), int(lip_lefttop_y), shifted_ds, pad_)
            else:
                fused, canon = R.post_fusion_compose(rgb_lip_warped, rgb_face_canonical, rgb_gt, mask_lip_canonical, coord,
                                                     int(lip_lefttop_x), int(lip_lefttop_y), shifted_ds, pad_)
            recon = self.post_fusion_unet(fused)
            return recon.permute(0, 2, 3, 1), fused.permute(0, 2, 3, 1), canon
        h, w = rgb_face_canonical.shape[1:3]
        lh, lw = rgb_lip_warped.shape[1:3]
        x0, y0 = int(lip_lefttop_x), int(lip_lefttop_y)
        left, up = x0 - 1, y0 - 1
        right, down = w - (left + lw), h - (up + lh)
        pad = (left + 1, right - 1, up + 1, down - 1) if shifted_ds else (left, right, up, down)
        lip_full = F.pad(rgb_lip_warped.permute(0, 3, 1, 2), pad=pad, mode='constant', value=0).permute(0, 2, 3, 1)
        merged_canonical = mask_lip_canonical * lip_full + (1 - mask_lip_canonical) * rgb_face_canonical
        mask = mask_lip_canonical
        if self.expand_lip_mask:
            p = lw // 12 if 'obama2_face_crop' in self.data_path else lw // 5
            mask = torch.zeros_like(mask_lip_canonical)
            mask[:, y0 - p:y0 + lh + 2 * p, x0 - p:x0 + lw + p, :] = 1
        merged = F.grid_sample(merged_canonical.permute(0, 3, 1, 2), coord, align_corners=False)
        mask_obs = F.grid_sample(mask.float().permute(0, 3, 1, 2), coord, align_corners=False)
        mask_obs = (mask_obs != 0).int()
        gt = rgb_gt.permute(0, 3, 1, 2)
        if aug:
            face_obs = F.grid_sample((rgb_face_canonical > 0).float().permute(0, 3, 1, 2), coord, align_corners=False)
            face_obs = (face_obs == 1).float()

            def holes(ref):
                keep = (torch.randn(ref.shape).to(ref.device)[:, :1] >= 0.000001).float()    # CPU draw, as tf_nerf.py:309
                keep = keep * face_obs + (1 - face_obs)
                return (keep != 0).float()

            n1, n2 = holes(merged), holes(gt)
            before = merged.clone()
            merged = n1 * before + (1 - n1) * gt
            gt = n2 * gt + (1 - n2) * before
        fused_in = mask_obs * merged + (1 - mask_obs) * gt
        recon = self.post_fusion_unet(fused_in)
        return recon.permute(0, 2, 3, 1), fused_in.permute(0, 2, 3, 1), merged_canonical
