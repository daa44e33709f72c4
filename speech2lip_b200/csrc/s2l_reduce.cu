// Per-pixel reductions after the MLP and ray generation:
//   ensemble4_blend_kernel — 4-tap area-weighted blend of Trainer.predict_lip_image (training.py:237-249)
//   composite_kernel       — density2outputs (rendering.py:30-62): one warp per ray, shuffle scan over samples
//   get_rays_kernel        — get_rays (src/common.py:12-21)
// All three are HBM-bound streaming kernels (16 B read per point); coalesced float4 loads, no staging needed.
#include "s2l_common.cuh"
#include "s2l_points.cuh"

namespace s2l {

__global__ void ensemble4_blend_kernel(const float* __restrict__ raw, PointSrc src, int n_frames, int out_ch,
                                       float* __restrict__ rgb) {
  const long long npix = (long long)src.H * src.W;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= npix * n_frames) return;
  const int f = (int)(gid / npix);
  const long long pix = gid % npix;
  // tot_area = stack(areas).sum(0); then areas 0<->3, 1<->2 are swapped (training.py:243-245)
  float wt[4];
  ens4_weights(src, f, (unsigned)pix, wt);
  const float* r = raw + (gid * 4) * out_ch;
  float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float w = wt[t];
#pragma unroll
    for (int c = 0; c < 3; ++c)
      if (c < out_ch) acc[c] = __fadd_rn(acc[c], __fmul_rn(r[t * out_ch + c], w));
  }
  for (int c = 0; c < 3; ++c) rgb[gid * 3 + c] = acc[c];
}

// One warp per ray.  weights_i = alpha_i * prod_{j<i} (1 - alpha_j + 1e-10); rgb = sum w*sigmoid(c); depth = sum w*z.
__global__ void __launch_bounds__(256) composite_kernel(const float* __restrict__ raw, const float* __restrict__ z_vals,
                                                        int z_per_ray, const float* __restrict__ rays_d,
                                                        long long n_rays, long long rays_mod, int S,
                                                        float* __restrict__ rgb, float* __restrict__ weights,
                                                        float* __restrict__ depth) {
  const int lane = threadIdx.x & 31;
  const long long ray = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (ray >= n_rays) return;
  const long long dr = rays_mod > 0 ? ray % rays_mod : ray;
  const float dx = rays_d[dr * 3 + 0], dy = rays_d[dr * 3 + 1], dz = rays_d[dr * 3 + 2];
  const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
  const float* zr = z_per_ray ? z_vals + ray * S : z_vals;
  const float4* rr = reinterpret_cast<const float4*>(raw) + ray * S;
  float carry = 1.f, ar = 0.f, ag = 0.f, ab = 0.f, ad = 0.f;
  for (int base = 0; base < S; base += 32) {
    const int i = base + lane;
    const bool valid = i < S;
    float alpha = 0.f, zi = 0.f;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) {
      v = __ldg(rr + i);
      zi = zr[i];
      const float dist = __fmul_rn((i + 1 < S) ? __fsub_rn(zr[i + 1], zi) : 1e10f, nrm);
      alpha = __fsub_rn(1.f, expf(-__fmul_rn(fmaxf(v.w, 0.f), dist)));
    }
    const float t = valid ? __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f) : 1.f;
    float incl = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl *= up;
    }
    float excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 1.f;
    const float w = alpha * (carry * excl);
    carry *= __shfl_sync(0xffffffffu, incl, 31);
    if (valid) {
      if (weights) weights[ray * S + i] = w;
      ar = fmaf(w, sigmoidf_acc(v.x), ar);
      ag = fmaf(w, sigmoidf_acc(v.y), ag);
      ab = fmaf(w, sigmoidf_acc(v.z), ab);
      ad = fmaf(w, zi, ad);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ar += __shfl_xor_sync(0xffffffffu, ar, o);
    ag += __shfl_xor_sync(0xffffffffu, ag, o);
    ab += __shfl_xor_sync(0xffffffffu, ab, o);
    ad += __shfl_xor_sync(0xffffffffu, ad, o);
  }
  if (lane == 0) {
    rgb[ray * 3 + 0] = ar;
    rgb[ray * 3 + 1] = ag;
    rgb[ray * 3 + 2] = ab;
    if (depth) depth[ray] = ad;
  }
}

__global__ void get_rays_kernel(const float* __restrict__ c2w, int H, int W, float focal, float* __restrict__ rays_o,
                                float* __restrict__ rays_d) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= H * W) return;
  const float i = (float)(gid % W), j = (float)(gid / W);
  const float d0 = __fdiv_rn(__fsub_rn(i, (float)(W * 0.5)), focal);
  const float d1 = __fdiv_rn(-__fsub_rn(j, (float)(H * 0.5)), focal);
  const float d2 = -1.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float s = __fadd_rn(__fadd_rn(__fmul_rn(d0, c2w[c * 4 + 0]), __fmul_rn(d1, c2w[c * 4 + 1])), __fmul_rn(d2, c2w[c * 4 + 2]));
    rays_d[gid * 3 + c] = s;
    rays_o[gid * 3 + c] = c2w[c * 4 + 3];
  }
}

// deepspeech_features.py:65-75: zero-pad 8 rows each side, windows of 16 rows with stride 2
__global__ void audio_windows_kernel(const float* __restrict__ logits, long long T, float* __restrict__ win) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nw = (T + 1) / 2;
  if (gid >= nw * 16 * 29) return;
  const int c = (int)(gid % 29);
  const int r = (int)((gid / 29) % 16);
  const long long w = gid / (29 * 16);
  const long long src = 2 * w + r - 8;
  win[gid] = (src >= 0 && src < T) ? logits[src * 29 + c] : 0.f;
}

// inference.py:173-178: cvtColor(RGB2BGR) then imwrite(img * 255): fp32 multiply, cvRound (half-to-even), saturate
__global__ void frames_to_bgr8_kernel(const float* __restrict__ rgb, long long n, uint8_t* __restrict__ bgr) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n) return;
  const float* p = rgb + gid * 3;
  uint8_t o[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = rintf(__fmul_rn(p[c], 255.f));
    o[2 - c] = (uint8_t)(v >= 255.f ? 255 : (v > 0.f ? (int)v : 0));      // NaN -> 0 like cv::saturate_cast
  }
  bgr[gid * 3 + 0] = o[0];
  bgr[gid * 3 + 1] = o[1];
  bgr[gid * 3 + 2] = o[2];
}

// Per-frame tile ranges of a compacted ray list: ts[f] = sum_{g<f} ceil(count[g] * Sc / T) for T = 128 (tensor-core
// kernels) and T = 64 (fp32 kernel); ts[F] = the launch's tile count, read by the persistent kernels on the device so
// that no host synchronisation sits between the launch that builds a list and the launch that consumes it.
__global__ void __launch_bounds__(256) tile_scan_kernel(const int* __restrict__ count, int F, int Sc, int* __restrict__ ts128,
                                                        int* __restrict__ ts64) {
  __shared__ int part[2][256];
  const int t = threadIdx.x;
  const int per = (F + 255) / 256;
  const int lo = min(t * per, F), hi = min(lo + per, F);
  int s128 = 0, s64 = 0;
  for (int f = lo; f < hi; ++f) {
    const long long pts = (long long)count[f] * Sc;
    s128 += (int)((pts + 127) / 128);
    s64 += (int)((pts + 63) / 64);
  }
  part[0][t] = s128;
  part[1][t] = s64;
  __syncthreads();
  if (t == 0) {
    int a128 = 0, a64 = 0;
    for (int i = 0; i < 256; ++i) {
      const int v128 = part[0][i], v64 = part[1][i];
      part[0][i] = a128;
      part[1][i] = a64;
      a128 += v128;
      a64 += v64;
    }
    ts128[F] = a128;
    ts64[F] = a64;
  }
  __syncthreads();
  s128 = part[0][t];
  s64 = part[1][t];
  for (int f = lo; f < hi; ++f) {
    ts128[f] = s128;
    ts64[f] = s64;
    const long long pts = (long long)count[f] * Sc;
    s128 += (int)((pts + 127) / 128);
    s64 += (int)((pts + 63) / 64);
  }
}

// Unfused volumetric path: list the rays whose last-sample density (raw[ray, S-1, 3]) is within thr of zero.
__global__ void flag_last_kernel(const float* __restrict__ raw, int F, int R, int S, float thr, const float* __restrict__ auto_thr,
                                 int* __restrict__ count, int* __restrict__ rays) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)F * R) return;
  if (thr < 0.f) thr = *auto_thr;
  const float sg = raw[(gid * S + (S - 1)) * 4 + 3];
  if (fabsf(sg) < thr) {
    const int f = (int)(gid / R);
    const int slot = atomicAdd(count + f, 1);
    rays[(long long)f * R + slot] = (int)(gid % R);
  }
}

int launch_tile_scan(const int* count, int F, int Sc, int* ts128, int* ts64, cudaStream_t st) {
  tile_scan_kernel<<<1, 256, 0, st>>>(count, F, Sc, ts128, ts64);
  return check_launch("tile_scan_kernel") ? 0 : 5;
}
int launch_flag_last(const float* raw, int F, int R, int S, float thr, const float* auto_thr, int* count, int* rays, cudaStream_t st) {
  const long long n = (long long)F * R;
  flag_last_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(raw, F, R, S, thr, auto_thr, count, rays);
  return check_launch("flag_last_kernel") ? 0 : 5;
}

}  // namespace s2l

using namespace s2l;

extern "C" int32_t s2l_audio_windows(const float* logits, int64_t n_steps, float* windows, void* stream) {
  if (n_steps < 0) { set_error("s2l_audio_windows: negative n_steps"); return 2; }
  if (n_steps == 0) return 0;
  if (!logits || !windows) { set_error("s2l_audio_windows: null argument"); return 1; }
  const long long n = ((long long)n_steps + 1) / 2 * 16 * 29;
  audio_windows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(logits, n_steps, windows);
  return check_launch("audio_windows_kernel") ? 0 : 5;
}

extern "C" int32_t s2l_frames_to_bgr8(const float* rgb, int64_t n_pixels, uint8_t* bgr, void* stream) {
  if (n_pixels < 0) { set_error("s2l_frames_to_bgr8: negative n_pixels"); return 2; }
  if (n_pixels == 0) return 0;
  if (!rgb || !bgr) { set_error("s2l_frames_to_bgr8: null argument"); return 1; }
  frames_to_bgr8_kernel<<<(unsigned)((n_pixels + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(rgb, n_pixels, bgr);
  return check_launch("frames_to_bgr8_kernel") ? 0 : 5;
}

extern "C" int32_t s2l_ensemble4_blend(const float* raw, const S2LGeom* g, float* rgb, void* stream) {
  if (!raw || !g || !rgb) { set_error("s2l_ensemble4_blend: null argument"); return 1; }
  PointSrc src{};
  src.mode = S2L_PTS_GRID_ENS4;
  src.H = g->height;
  src.W = g->width;
  src.uv_dims = 2;
  src.eps = g->eps_shift;
  src.eps_pf = g->eps_per_frame;
  src.P = (long long)g->height * g->width * 4;
  src.step_w = g->width > 1 ? 1.0f / (float)(g->width - 1) : 0.f;
  src.step_h = g->height > 1 ? 1.0f / (float)(g->height - 1) : 0.f;
  const long long n = (long long)g->height * g->width * g->n_frames;
  if (n == 0) return 0;
  ensemble4_blend_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      raw, src, g->n_frames, g->out_ch, rgb);
  return check_launch("ensemble4_blend_kernel") ? 0 : 5;
}

extern "C" int32_t s2l_composite_fwd(const float* raw, const float* z_vals, int32_t z_per_ray, const float* rays_d,
                                     int64_t n_rays, int64_t rays_mod, int32_t n_samples, float* rgb, float* weights,
                                     float* depth, void* stream) {
  if (!raw || !z_vals || !rays_d || !rgb) { set_error("s2l_composite_fwd: null argument"); return 1; }
  if (n_samples < 1) { set_error("s2l_composite_fwd: n_samples must be >= 1 (got %d)", n_samples); return 2; }
  if (n_rays == 0) return 0;
  const long long threads = (long long)n_rays * 32;
  composite_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      raw, z_vals, z_per_ray, rays_d, n_rays, rays_mod, n_samples, rgb, weights, depth);
  return check_launch("composite_kernel") ? 0 : 5;
}

extern "C" int32_t s2l_get_rays(const float* c2w, int32_t height, int32_t width, float focal, float* rays_o,
                                float* rays_d, void* stream) {
  if (!c2w || !rays_o || !rays_d) { set_error("s2l_get_rays: null argument"); return 1; }
  const int n = height * width;
  if (n == 0) return 0;
  get_rays_kernel<<<(n + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(c2w, height, width, focal,
                                                                                        rays_o, rays_d);
  return check_launch("get_rays_kernel") ? 0 : 5;
}
