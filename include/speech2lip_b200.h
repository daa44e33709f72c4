/*
 * speech2lip_b200 — C ABI of the B200-native Speech2Lip rendering hot path.
 *
 * The reference (CVMI-Lab/Speech2Lip) is pure Python/PyTorch and has no FFI of its
 * own; its boundary for this path is the Python class
 *   src/face_simple/models/tf_nerf.py:12  class TalkingFace(nn.Module)
 * called by inference.py:144-159 and src/face_simple/training.py:158-251.
 * The entry points below are what a ctypes binding for that class binds
 * (speech2lip_b200/_cabi.py is that binding; INTEGRATION.md shows the reference-side stub).
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a *device* pointer borrowed
 *    from the caller (torch owns the memory) unless the name ends in _host.
 *  - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); all
 *    work is enqueued on it, no hidden synchronisation, no allocation.
 *  - return value: 0 on success, non-zero on error; s2l_last_error() returns a
 *    thread-local, NUL-terminated description of the last failure.
 *  - all floating-point tensors are fp32, row-major, contiguous.
 */
#ifndef SPEECH2LIP_B200_H
#define SPEECH2LIP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define S2L_ABI_VERSION 2

/* Number of parameter tensors s2l_pack_weights() reads, in this fixed order
 * (reference state_dict names, tf_nerf.py:85-172):                                  */
enum {
  S2L_P_CONV0_W = 0, S2L_P_CONV0_B,      /* encoder_conv.0  [32,29,3],[32]   tf_nerf.py:92  */
  S2L_P_CONV1_W, S2L_P_CONV1_B,          /* encoder_conv.2  [32,32,3],[32]   tf_nerf.py:95  */
  S2L_P_CONV2_W, S2L_P_CONV2_B,          /* encoder_conv.4  [64,32,3],[64]   tf_nerf.py:98  */
  S2L_P_CONV3_W, S2L_P_CONV3_B,          /* encoder_conv.6  [64,64,3],[64]   tf_nerf.py:101 */
  S2L_P_FC1_W, S2L_P_FC1_B,              /* encoder_fc1.0   [64,64],[64]     tf_nerf.py:106 */
  S2L_P_FC2_W, S2L_P_FC2_B,              /* encoder_fc1.2   [64,64],[64]     tf_nerf.py:108 */
  S2L_P_FC_UV_W, S2L_P_FC_UV_B,          /* fc_uv           [256,E],[256]    tf_nerf.py:149 */
  S2L_P_FC_UV_SKIP_W, S2L_P_FC_UV_SKIP_B,/* fc_uv_skip      [256,E],[256]    tf_nerf.py:150 */
  S2L_P_FC_AUDIO_W, S2L_P_FC_AUDIO_B,    /* fc_audio        [256,64],[256]   tf_nerf.py:152 */
  S2L_P_FC_AUDIO_SKIP_W, S2L_P_FC_AUDIO_SKIP_B, /* fc_audio_skip             tf_nerf.py:153 */
  S2L_P_FC_TIME_W, S2L_P_FC_TIME_B,      /* fc_time         [256,20],[256]   tf_nerf.py:159 */
  S2L_P_FC_TIME_SKIP_W, S2L_P_FC_TIME_SKIP_B,   /* fc_time_skip              tf_nerf.py:160 */
  S2L_P_PTS0_W, S2L_P_PTS0_B,            /* pts_linears.0..7 [256,256] ([256,512] for .5, input
                                            order [h_skip,h])                tf_nerf.py:170-172 */
  S2L_P_PTS_LAST_B = S2L_P_PTS0_W + 15,
  S2L_P_OUT_W, S2L_P_OUT_B,              /* output_linear   [out_ch,256],[out_ch] tf_nerf.py:144 */
  S2L_NUM_PARAMS
};

/* MLP arithmetic selection for s2l_mlp_fwd / s2l_render_frames. */
enum {
  S2L_PREC_FP32   = 0,  /* CUDA-core fp32 FFMA, literal (unfolded) layer order: the exact path      */
  S2L_PREC_BF16X3 = 1,  /* tcgen05 bf16 hi/lo split, 3 MMAs per product, fp32 accumulate (parity)   */
  S2L_PREC_BF16X1 = 2,  /* tcgen05 single bf16 pass (fast, NOT within the 1e-3 parity bar)          */
  S2L_PREC_FP16F8 = 3   /* tcgen05 fp16 main product + two fp8 (e4m3/e5m2) correction products, fp32 accumulate:
                           2 bf16-MMA equivalents per product, ~3e-4 max-abs on O(8) outputs (parity).
                           VALIDATED DOMAIN: every tensor-core weight |w| < 1024 and every hidden activation |a| < 4096 (beyond
                           that the scaled fp8 residuals saturate and the correction terms are silently lost; fp16 itself
                           overflows at 65504).  s2l_pack_weights records the weight side (s2l_blob_meta); the host layer
                           refuses fp16f8 outside it and falls back to S2L_PREC_BF16X3 with a logged warning
                           (speech2lip_b200.LipRenderer); activations can be probed with the exact path
                           (LipRenderer.probe_fp16f8_domain) or counted in a -DS2L_DBG_SATCOUNT build.             */
};

/* How the kernel obtains the coordinates of point-evaluation p of frame f. */
enum {
  S2L_PTS_GRID      = 0, /* (u,v) = get_coords(W,H) grid, 1 eval / pixel      rendering.py:9-28, inference.py:146 */
  S2L_PTS_GRID_ENS4 = 1, /* 4 jittered+clamped taps / pixel, tap-minor order  training.py:195-236                 */
  S2L_PTS_RAYS      = 2, /* x = o + d*z, S samples / ray, sample-minor order  (volumetric mode)      */
  S2L_PTS_EXPLICIT  = 3  /* coordinates read from `pts` [F*P, uv_dims]         rgb_forward contract, tf_nerf.py:225 */
};

typedef struct S2LGeom {
  int32_t n_frames;        /* F                                                              */
  int32_t height, width;   /* H, W (GRID*, RAYS: rays per frame = H*W)                        */
  int32_t n_samples;       /* S for RAYS, else ignored                                        */
  int32_t pts_mode;        /* S2L_PTS_*                                                       */
  int32_t uv_dims;         /* 2 (live model) or 3 (volumetric model), must match the blob     */
  int32_t out_ch;          /* 3 or 4, must match the blob                                     */
  int32_t z_per_ray;       /* RAYS: 1 -> z_vals is [F*R,S]; 0 -> z_vals is [S] shared         */
  int32_t rays_per_frame_shared; /* RAYS: 1 -> rays_o/rays_d are [R,3] shared by all frames;
                                           0 -> [F*R,3]                                       */
  int64_t pts_per_frame;   /* EXPLICIT: P (points per frame); otherwise derived               */
  float   eps_shift;       /* GRID_ENS4: the reference's eps_shift draw (training.py:200)     */
  const float* eps_per_frame; /* GRID_ENS4, optional DEVICE array [F]: one draw per frame (the sync-window
                                 render calls predict_lip_image once per frame, training.py:504-525);
                                 NULL -> eps_shift is used for every frame                             */
  /* ---- ABI v2: volumetric (RAYS) rendering options of s2l_render_frames; all-zero = v1 behaviour + the default fix_thr */
  int32_t sample_chunks;   /* 0/1: every sample of every ray in one launch.  C > 1 (C | S, S/C in {4,8,...,128}): C front-to-back
                              launches of S/C samples with EARLY RAY TERMINATION — after each chunk rays whose transmittance
                              fell below term_thr are finished (what is left would change the pixel by < term_thr) and only the
                              survivors (compacted per frame on the device) enter the next chunk                               */
  float   term_thr;        /* early-ray-termination threshold on the transmittance (e.g. 1e-4); <= 0 with C > 1: chunked, never terminates */
  float   fix_thr;         /* tensor-core precisions: rays whose LAST sample's density is within fix_thr of zero are re-evaluated
                              in exact fp32 (density2outputs gives the last sample delta = 1e10, rendering.py:44, so alpha_last is a
                              step function of sign(sigma_last) and a rounding error there flips the pixel).  0 -> automatic:
                              2e-3 * max(1, |output_linear.weight[3]|_2 / sqrt(2)) computed at pack time (2e-3 = 3x the worst
                              tensor-core density error measured on kaiming-scale weights; the error scales with the density
                              row), < 0 -> off                                                                               */
} S2LGeom;

/* Thread-local description of the last error (never NULL). */
const char* s2l_last_error(void);
int32_t     s2l_abi_version(void);

/* PositionalEncodingTime.div_term (tf_nerf.py:431-432), 10 fp32 values, host-side helper. */
void s2l_time_div_term(float* out10_host);

/* What s2l_pack_weights recorded about the packed model (synchronises `stream`; call once per pack, not per frame):
 * max |w| over the tensor-core weights incl. the folded input layers, how many of them leave the fp16f8 domain
 * (|w| >= 1024), the L2 norm of output_linear's density row and the automatic fix_thr derived from it.  Any out pointer may be NULL. */
int32_t s2l_blob_meta(const void* blob, float* max_abs_weight, int32_t* n_saturating_weights, float* density_row_norm,
                      float* auto_fix_thr, void* stream);

/* Size in bytes of the packed weight blob for a model with the given dims. */
size_t s2l_blob_bytes(int32_t uv_dims, int32_t out_ch);

/* Replaces: TalkingFace.__init__ parameter set + load_state_dict (tf_nerf.py:13-195,
 * src/checkpoints.py:97-116).  `params` is a HOST array of S2L_NUM_PARAMS DEVICE pointers
 * (fp32, PyTorch [out,in] layout).  Writes the kernel-layout blob (fp32 exact-path copies,
 * folded/split bf16 tensor-core operand images, constant tables).  Must be re-run whenever a
 * parameter changes (optimizer.step, load_state_dict). */
int32_t s2l_pack_weights(const float* const* params_host, void* blob, int32_t uv_dims, int32_t out_ch,
                         void* stream);

/* Replaces: TalkingFace.audio_merge_forward (tf_nerf.py:197-213) plus the per-frame-constant
 * terms of rgb_forward (fc_audio/fc_time and their *_skip twins, tf_nerf.py:254-258,270-276) and
 * PositionalEncodingTime (tf_nerf.py:427-442).
 *   audio      [F,16,29] (transposed=0) or [F,29,16] (transposed=1)
 *   frame_idx  [F] int64 (time_pts; NULL -> time term omitted)
 *   latent     [F,64]  out (may be NULL)
 *   frame_bias [F,4,256] out (may be NULL): rows = {bias0, bias_skip, folded bias0', folded bias5'} */
int32_t s2l_audio_encode_fwd(const void* blob, const float* audio, int32_t transposed,
                             const int64_t* frame_idx, float* latent, float* frame_bias,
                             int32_t n_frames, int32_t uv_dims, int32_t out_ch, void* stream);

/* The per-frame-constant terms of rgb_forward alone (tf_nerf.py:254-258, 270-276, 427-442), for a caller that already
 * holds AudioNet's output — the drop-in TalkingFace.rgb_forward receives it as the latent columns of uv_audio_pts
 * (inference.py:150-158):  latent [F,64] with row stride latent_stride floats -> frame_bias [F,4,256] as above. */
int32_t s2l_latent_bias_fwd(const void* blob, const float* latent, int64_t latent_stride, const int64_t* frame_idx,
                            float* frame_bias, int32_t n_frames, void* stream);

/* flag[0] (device int32) <- 1 if any of the n_rows rows of x (row stride row_stride floats) differs bitwise from row 0 in
 * columns [col0, col0+ncols), else 0.  Lets the drop-in detect inference.py's tiled inputs (inference.py:144: the same
 * audio window H*W times; :151: the same latent in every row) and run them once / through the tensor-core path. */
int32_t s2l_rows_differ_or(const float* x, int64_t n_rows, int64_t row_stride, int32_t col0, int32_t ncols, int32_t* flag,
                           void* stream);   /* as s2l_rows_differ, but *flag is only ever set to 1, never cleared (sticky across calls) */
int32_t s2l_rows_differ(const float* x, int64_t n_rows, int64_t row_stride, int32_t col0, int32_t ncols, int32_t* flag,
                        void* stream);

/* The same two caller patterns WITHOUT a host synchronisation (what the drop-in TalkingFace calls): the compare kernel
 * leaves its verdict in device memory and the kernels that follow are gated on it.
 *   s2l_audio_merge_auto : audio [B,16,29] (transposed=1: [B,29,16]) -> latent [B,64]; encodes row 0, then either
 *                          broadcasts it (all rows equal) or encodes every row.
 *   s2l_rgb_forward_auto : x [N, uv_dims+64], time_idx_dev = DEVICE int64[1] (position[0], tf_nerf.py:439; NULL: no time
 *                          term) -> out [N,out_ch].  Rows with one common latent run the fused tensor-core MLP in
 *                          `precision` (S2L_PREC_BF16X3 / FP16F8 / BF16X1) reading the coordinates in place (row stride
 *                          uv_dims+64); otherwise the general fp32 per-row kernel runs.  S2L_PREC_FP32: always the latter.
 * scratch: s2l_*_auto_scratch_bytes() bytes of device memory per in-flight call. */
size_t  s2l_audio_merge_auto_scratch_bytes(void);
int32_t s2l_audio_merge_auto(const void* blob, const float* audio, int32_t transposed, int64_t n_rows, float* latent,
                             void* scratch, void* stream);
size_t  s2l_rgb_forward_auto_scratch_bytes(void);
int32_t s2l_rgb_forward_auto(const void* blob, const float* x, int64_t n_rows, const int64_t* time_idx_dev, float* out,
                             int32_t uv_dims, int32_t out_ch, int32_t precision, void* scratch, void* stream);

/* Replaces: TalkingFace.rgb_forward (tf_nerf.py:225-285) for F frames x P points with a
 * per-frame-constant latent (frame_bias from s2l_audio_encode_fwd).  Point coordinates come from
 * geom->pts_mode.  raw_out [F*P_padless, out_ch] receives the raw linear outputs in point order. */
int32_t s2l_mlp_fwd(const void* blob, const S2LGeom* geom, const float* frame_bias,
                    const float* pts, const float* rays_o, const float* rays_d, const float* z_vals,
                    float* raw_out, int32_t precision, void* stream);

/* Replaces: TalkingFace.rgb_forward for the general contract (arbitrary latent per row):
 *   x [N, uv_dims+64], time index `time_idx` (position[0], tf_nerf.py:439) -> out [N,out_ch].
 * Always fp32 exact path. */
int32_t s2l_rgb_forward_rows(const void* blob, const float* x, int64_t n_rows, int64_t time_idx,
                             int32_t has_time, float* out, int32_t uv_dims, int32_t out_ch, void* stream);

/* Training forward of the general contract: same as s2l_rgb_forward_rows, and additionally saves the 10 activation
 * tensors the backward needs: acts [10][N][256] = {net, h0..h4, h_skip, h5, h6, h7} (tf_nerf.py:252-281). */
int32_t s2l_rgb_forward_rows_train(const void* blob, const float* x, int64_t n_rows, int64_t time_idx,
                                   int32_t has_time, float* out, float* acts, int32_t uv_dims, int32_t out_ch,
                                   void* stream);

/* Replaces: autograd's backward through rgb_forward (loss.backward(), training.py:559), data-gradient half.
 *   d_out [N,out_ch], acts from the training forward -> dsave [10][N][256] =
 *   {d net, dPre0..dPre4, d h_skip, dPre5, dPre6, dPre7}.  Weight gradients are dPre_l^T * h_{l-1} (plain GEMMs). */
int32_t s2l_mlp_bwd_rows(const void* blob, const float* d_out, const float* acts, int64_t n_rows, float* dsave,
                         int32_t out_ch, void* stream);

/* ---- Training render on tensor cores (bf16 operands, fp32 accumulate) -------------------------------------------------
 * Replaces: autograd through Trainer.predict_lip_image -> TalkingFace.rgb_forward for F frames x 4 taps
 * (training.py:158-251, loss.backward() training.py:559; the five-frame sync-expert window training.py:500-548 is the same
 * call with F = 5) as ONE differentiable launch sequence instead of 4 F rgb_forward calls:
 *   s2l_train_fwd : latent [F,64] (AudioNet's output, tf_nerf.py:197-213) + frame_idx [F] -> frame_bias [F,4,256], then the
 *                   fused bf16 MLP with the 4-tap blend in its epilogue -> rgb [F,H,W,3]; saves h0..h7 and the positional
 *                   encodings (bf16) in `workspace` for the backward.
 *   s2l_train_bwd : d_rgb [F,H,W,3] -> gradients of every MLP tensor in the reference's layouts, written (not accumulated)
 *                   to grads_host[S2L_P_FC_UV_W .. S2L_NUM_PARAMS-1] (HOST array of DEVICE pointers in S2L_P_* order, fp32;
 *                   entries of the AudioNet tensors are ignored), and d_latent [F,64] for AudioNet's own backward.
 *                   Kernels: tensor-core data-gradient chain (tcgen05, dPre kept in TMEM between layers), split-K
 *                   tensor-core weight-gradient GEMMs on MN-major operands, slab reduction + chain rule through the folded
 *                   input layers.  No library GEMM is called.
 * geom: pts_mode = S2L_PTS_GRID_ENS4, uv_dims = 2, out_ch = 3, eps_shift / eps_per_frame as for s2l_render_frames.
 * workspace must hold s2l_train_workspace_bytes(geom) bytes and stay untouched between fwd and bwd (~9.3 KB per point). */
/* AudioNet with gradients (tf_nerf.py:197-213 under autograd): the forward additionally saves its post-activation tensors
 * (s2l_audio_train_save_floats(F) floats); the backward turns d_latent [F,64] into the gradients of the 12 AudioNet tensors
 * (grads_host: HOST array of 12 DEVICE pointers, S2L_P_CONV0_W .. S2L_P_FC2_B order; written, not accumulated).  One CTA per
 * frame + a frame reduction; scratch: s2l_audio_train_scratch_bytes(F). */
size_t  s2l_audio_train_save_floats(int32_t n_frames);
size_t  s2l_audio_train_scratch_bytes(int32_t n_frames);
int32_t s2l_audio_train_fwd(const void* blob, const float* audio, int32_t transposed, float* latent, float* save, int32_t n_frames,
                            void* stream);
int32_t s2l_audio_train_bwd(const void* blob, const float* audio, int32_t transposed, const float* save, const float* d_latent,
                            float* const* grads_host, void* scratch, int32_t n_frames, void* stream);
size_t  s2l_train_workspace_bytes(const S2LGeom* geom);
int32_t s2l_train_fwd(const void* blob, const S2LGeom* geom, const float* latent, const int64_t* frame_idx, float* rgb,
                      float* frame_bias, void* workspace, void* stream);
int32_t s2l_train_bwd(const void* blob, const S2LGeom* geom, const float* d_rgb, const float* latent, const int64_t* frame_idx,
                      const float* frame_bias, void* workspace, float* const* grads_host, float* d_latent, void* stream);

/* The same kernels behind the reference's PER-CALL contract: autograd through ONE rgb_forward call (tf_nerf.py:225-285) whose
 * N rows share one latent — what Trainer.predict_lip_image passes four times per frame (training.py:216-233).  x [N,66] (the
 * coordinates are read in place, the latent from row 0), time_idx_dev = DEVICE int64[1] or NULL, out [N,3] raw outputs;
 * the backward writes the 30 MLP gradients like s2l_train_bwd and d_latent [64] (the gradient of the shared latent = the sum
 * over the rows).  bf16; the exact fp32 path of the contract is s2l_rgb_forward_rows_train / s2l_mlp_bwd_rows. */
size_t  s2l_train_rows_workspace_bytes(int64_t n_rows);
int32_t s2l_train_rows_fwd(const void* blob, const float* x, int64_t n_rows, const int64_t* time_idx_dev, float* out,
                           float* frame_bias, void* workspace, void* stream);
int32_t s2l_train_rows_bwd(const void* blob, const float* d_out, const float* x, int64_t n_rows, const int64_t* time_idx_dev,
                           const float* frame_bias, void* workspace, float* const* grads_host, float* d_latent, void* stream);

/* Data-parallel training exchange (replaces DistributedDataParallel's gradient all-reduce, training.py:40 / train.py:102;
 * SURVEY 8(e) "Training (C5)").  Each rank allocates ONE buffer (s2l_peer_alloc: flags + two payload buffers of n_floats),
 * sends its 64-byte CUDA-IPC handle to its peers (any transport — the Python host uses torch.distributed), opens theirs
 * (s2l_peer_open), and per step: writes its flat gradient bucket at s2l_peer_payload_offset(n_floats, epoch) of its own
 * buffer (stream order), then calls s2l_allreduce_peer with epoch = 1, 2, 3, ... : ONE kernel that signals, waits for the
 * peers' signals and sums all payloads over NVLink peer loads in rank order into out (x scale) — bit-identical on every
 * rank.  peer_bufs is a HOST array of `world` device pointers (this rank's own buffer at index rank).  Every rank must call
 * with the same epoch sequence; a peer that never arrives traps the kernel after a few seconds instead of hanging. */
size_t  s2l_peer_buffer_bytes(int64_t n_floats);
size_t  s2l_peer_payload_offset(int64_t n_floats, uint32_t epoch);
int32_t s2l_peer_alloc(int64_t n_floats, void** dptr, uint8_t* handle64);
int32_t s2l_peer_open(const uint8_t* handle64, void** dptr);
int32_t s2l_peer_close(void* dptr);
int32_t s2l_peer_free(void* dptr);
int32_t s2l_allreduce_peer(const void* const* peer_bufs, int32_t rank, int32_t world, int64_t n_floats, float scale, uint32_t epoch,
                           float* out, void* stream);

/* The GEMMs of the exact fp32 per-call backward (autograd through tf_nerf.py:252-283, loss.backward() training.py:559):
 *   s2l_wgrad_rows_fp32: out[l] [A,B] = dy[l]^T h[l] for n_mats matrices, dy[l] [N,A] / h[l] [N,B] row-major, matrix l at
 *                        dy + l*mat_stride_dy / h + l*mat_stride_h (a stride of 0 shares one operand between the matrices);
 *                        split over the rows, partials in scratch (s2l_wgrad_rows_scratch_bytes), summed in a fixed order.
 *   s2l_dx_rows_fp32:    out [N,b_dim] (row stride ld_out) = a1 [N,256] w1 [256,b_dim] (+ a2 w2 when a2 != NULL).  */
size_t  s2l_wgrad_rows_scratch_bytes(int64_t n_rows, int32_t n_mats, int32_t a_dim, int32_t b_dim);
int32_t s2l_wgrad_rows_fp32(const float* dy, const float* h, int64_t n_rows, int32_t n_mats, int32_t a_dim, int32_t b_dim,
                            int64_t mat_stride_dy, int64_t mat_stride_h, float* out, void* scratch, void* stream);
int32_t s2l_dx_rows_fp32(const float* a1, const float* w1, const float* a2, const float* w2, int64_t n_rows, int32_t b_dim,
                         float* out, int32_t ld_out, void* stream);

/* Replaces: Embedder.__call__ (tf_nerf.py:404-425): x rows (first uv_dims floats of each row_stride-float row) -> pe [N,E]. */
int32_t s2l_embed_fwd(const float* x, int64_t n_rows, int32_t row_stride, int32_t uv_dims, float* pe, void* stream);

/* Replaces: the 4-tap blend of Trainer.predict_lip_image (training.py:238-249).
 *   raw [F*H*W*4, out_ch] (tap-minor) -> rgb [F,H,W,3] */
int32_t s2l_ensemble4_blend(const float* raw, const S2LGeom* geom, float* rgb, void* stream);

/* Replaces: density2outputs (rendering.py:30-62, raw_noise_std=0).
 *   raw [R,S,4], z_vals ([R,S] or [S]), rays_d [R,3] -> rgb [R,3], weights [R,S] (NULL ok), depth [R] (NULL ok).
 *   rays_mod: rays_d row index = ray % rays_mod (rays shared between frames), 0 = no wrap. */
int32_t s2l_composite_fwd(const float* raw, const float* z_vals, int32_t z_per_ray, const float* rays_d,
                          int64_t n_rays, int64_t rays_mod, int32_t n_samples,
                          float* rgb, float* weights, float* depth, void* stream);

/* Replaces: get_rays (src/common.py:12-21).  c2w [3,4] row-major (device) -> rays_o, rays_d [H*W,3]. */
int32_t s2l_get_rays(const float* c2w, int32_t height, int32_t width, float focal,
                     float* rays_o, float* rays_d, void* stream);

/* Whole-path call on device buffers: AudioNet -> MLP -> per-pixel reduction for F frames.
 *   rgb [F,H,W,3] out; scratch must hold s2l_render_scratch_bytes(geom, precision, want_aux) bytes
 *   (want_aux = weights or depth requested).
 * Replaces the loop body of inference.py:144-159 (GRID), training.py:158-251 (GRID_ENS4) or the
 * assembled volumetric path (RAYS).
 * With a tensor-core precision the per-pixel reduction (4-tap blend, alpha compositing) runs inside the MLP kernel's
 * output-layer epilogue and the raw [P,out_ch] tensor never exists in memory; the unfused composition (MLP -> raw ->
 * s2l_composite_fwd / s2l_ensemble4_blend) serves S2L_PREC_FP32, weights/depth outputs and sample counts S with
 * S % 4 != 0 or 128 % S != 0. */
size_t  s2l_render_scratch_bytes(const S2LGeom* geom, int32_t precision, int32_t want_aux);
/* Byte offset inside `scratch` of the int32 counters a RAYS render leaves behind: counts[k*F + f], k = 1..C-1 = rays of
 * frame f still alive when sample chunk k started, k = C (C = max(sample_chunks,1)) = rays of frame f re-evaluated in fp32. */
size_t  s2l_render_counts_offset(const S2LGeom* geom);
int32_t s2l_render_frames(const void* blob, const S2LGeom* geom, const float* audio, const int64_t* frame_idx,
                          const float* rays_o, const float* rays_d, const float* z_vals,
                          float* rgb, float* weights, float* depth, void* scratch, int32_t precision,
                          void* stream);

/* Replaces: the pre-UNet part of TalkingFace.post_fusion2_onlylip_light (tf_nerf.py:334-386, inference branch,
 * no black-hole augmentation): paste lip crop -> blend with canonical lip mask -> grid_sample canonical->observed
 * with `coord` -> binarise warped mask -> blend with rgb_gt.  Layouts: rgb_lip [B,lh,lw,3], face_canonical [B,h,w,3],
 * rgb_gt [B,Hf,Wf,3], mask_lip_canonical [B,h,w,3], coord [B,Hf,Wf,2] in [-1,1].
 *   paste_shift = 1 for datasets whose path contains 'may'/'macron'/'obama_adnerf'/'obama2_face_crop' (tf_nerf.py:345-348)
 *   expand_pad  = lip_w/5 (or lip_w/12) when cfg expand_lip_mask, -1 to warp the lip mask itself (tf_nerf.py:354-363)
 * Outputs: fused_nchw [B,3,Hf,Wf] (the UNet input), merged_canonical [B,h,w,3] (may be NULL). */
int32_t s2l_post_fusion_compose(const float* rgb_lip, const float* face_canonical, const float* rgb_gt,
                                const float* mask_lip_canonical, const float* coord, int32_t batch, int32_t lip_h,
                                int32_t lip_w, int32_t face_h, int32_t face_w, int32_t out_h, int32_t out_w,
                                int32_t lefttop_x, int32_t lefttop_y, int32_t paste_shift, int32_t expand_pad,
                                float* fused_nchw, float* merged_canonical, void* stream);

/* Backward of s2l_post_fusion_compose w.r.t. the lip crop — the only input that carries a gradient in training
 * (autograd through tf_nerf.py:334-386 from training.py:436-445, 559).  d_fused_nchw [B,3,out_h,out_w] and/or
 * d_merged_canonical [B,face_h,face_w,3] may be NULL (no gradient through that output); d_rgb_lip [B,lip_h,lip_w,3] is
 * overwritten.  The warp term is scattered with fp32 atomics (as ATen's grid_sampler backward does). */
int32_t s2l_post_fusion_compose_bwd(const float* d_fused_nchw, const float* d_merged_canonical, const float* mask_lip_canonical,
                                    const float* coord, int32_t batch, int32_t lip_h, int32_t lip_w, int32_t face_h, int32_t face_w,
                                    int32_t out_h, int32_t out_w, int32_t lefttop_x, int32_t lefttop_y, int32_t paste_shift,
                                    int32_t expand_pad, float* d_rgb_lip, void* stream);

/* Replaces: the windowing of preprocess/deepspeech_features/deepspeech_features.py:65-75 (zero-pad 8 rows on both
 * sides, 16-row windows, stride 2): logits [T,29] -> windows [ceil(T/2),16,29]  (the hot path's input format). */
int32_t s2l_audio_windows(const float* logits, int64_t n_steps, float* windows, void* stream);

/* Replaces: the output staging of inference.py:173-178 (cv2.cvtColor RGB2BGR + cv2.imwrite(path, img * 255), i.e.
 * x*255 in fp32, round-half-even, saturate to [0,255], channel swap): rgb [N_pixels,3] fp32 -> bgr [N_pixels,3] u8.
 * Cuts the device->host bytes of a finished frame 4x. */
int32_t s2l_frames_to_bgr8(const float* rgb, int64_t n_pixels, uint8_t* bgr, void* stream);

/* sizeof(S2LGeom) as this library was compiled (bindings check their struct definition against it). */
int32_t s2l_sizeof_geom(void);

/* Measurement aid (bench.py "roofline"): while enabled, every launch of the fused tensor-core MLP kernel is bracketed by
 * CUDA events on the stream it is launched on.  s2l_profile_mlp_ms synchronises on the recorded events, returns the summed
 * kernel time in ms since the last call (and the number of launches in *n_launches, may be NULL) and clears the record. */
void    s2l_profile_enable(int32_t on);
double  s2l_profile_mlp_ms(int32_t* n_launches);

/* Number of kernels of this library launched by this thread since the last reset (bench "gpu_launches"). */
int64_t s2l_launch_count(int32_t reset);

/* Which tensor-core schedule a launch of n_tiles 128-point tiles uses: 1 independent CTAs, 2 CTA pairs (cta_group::2),
 * 3 two-CTA clusters with one multicast weight stream.  S2L_TC_IMPL=1|2|3 forces one (see s2l_mlp_tc.cu). */
int32_t s2l_tc_schedule(int64_t n_tiles);

#ifdef __cplusplus
}
#endif
#endif /* SPEECH2LIP_B200_H */
