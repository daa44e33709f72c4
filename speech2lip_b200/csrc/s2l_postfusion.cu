// Post-fusion compose + warp (SURVEY §8(f) "next" rank 1): the pre-UNet part of
// TalkingFace.post_fusion2_onlylip_light (tf_nerf.py:334-386, inference branch) as ONE gather-blend kernel:
//   paste the rendered lip crop into the canonical face (F.pad), blend with the canonical lip mask,
//   warp canonical -> observed with the per-frame `coord` grid (2x F.grid_sample, bilinear, zeros,
//   align_corners=False), binarise the warped mask, blend with the ground-truth frame.
// HBM-bound: per output pixel 8 B coord + 12 B gt read + 12 B written, plus 4-tap gathers of the
// canonical face / mask that hit L1/L2 (neighbouring pixels sample neighbouring texels); the eager
// reference materialises ~10 full-size intermediates.  One thread per observed pixel, coalesced
// float2 / scalar accesses, output written planar (NCHW) because the UNet consumes NCHW.
#include "s2l_common.cuh"

namespace s2l {

struct PfArgs {
  const float* lip;      // [B,lh,lw,3]
  const float* face;     // [B,h,w,3]   canonical face
  const float* gt;       // [B,Hf,Wf,3] observed ground truth
  const float* mask;     // [B,h,w,3]   canonical lip mask
  const float* coord;    // [B,Hf,Wf,2] sampling grid in [-1,1]
  float* fused;          // [B,3,Hf,Wf]
  int B, lh, lw, h, w, Hf, Wf;
  int px0, py0;          // canonical position of lip pixel (0,0)
  int rect;              // 1: warp the expanded rectangle mask, 0: warp the lip mask itself
  int ry0, ry1, rx0, rx1;
};

// merged_canonical[b,cy,cx,c] = m*lip_pad + (1-m)*face      (tf_nerf.py:352)
__device__ __forceinline__ float merged_canon(const PfArgs& a, int b, int cy, int cx, int c) {
  const size_t idx = (((size_t)b * a.h + cy) * a.w + cx) * 3 + c;
  const float m = a.mask[idx];
  const int ly = cy - a.py0, lx = cx - a.px0;
  float lipv = 0.f;
  if (ly >= 0 && ly < a.lh && lx >= 0 && lx < a.lw) lipv = a.lip[(((size_t)b * a.lh + ly) * a.lw + lx) * 3 + c];
  return __fadd_rn(__fmul_rn(m, lipv), __fmul_rn(__fsub_rn(1.f, m), a.face[idx]));
}

// grid = (pixel blocks, batch): no 64-bit division per thread; 32-bit offsets inside a frame (an image plane is < 2^31 bytes);
// the four taps share one base offset.  ncu on the first version: sm__throughput 73 %, DRAM 30 % — the kernel was
// instruction-bound (611 warp-instructions per 32 pixels, mostly index arithmetic), not memory-bound.  Second capture
// (profiles/r2h): 482 instructions per warp, issue slots 71 % busy, L1 41 %, DRAM 2.9 TB/s — still issue-bound, and ~90 % of
// the pixels of a 500x500 frame lie outside the warped lip mask, where the result is the ground-truth pixel.  So the warped
// mask is evaluated FIRST (analytically in the rectangle mode, from the 12 mask taps otherwise) and a pixel whose mask is
// zero in every channel returns before any face / lip / mask gather: bit-identical output (the blend is a select on
// mask != 0), ~8x fewer instructions for those pixels, and face / mask are only read around the lip.
// one observed pixel: g = its sampling coordinate, gtv = its ground-truth colour -> out[3]
template <bool RECT>
__device__ __forceinline__ void pf_pixel_value(const PfArgs& a, int b, float2 g, const float (&gtv)[3], float (&out)[3]) {
  const size_t canon = (size_t)b * a.h * a.w * 3;
  const float* __restrict__ face = a.face + canon;
  const float* __restrict__ mask = a.mask + canon;
  const float* __restrict__ lip = a.lip + (size_t)b * a.lh * a.lw * 3;
  // grid_sampler_unnormalize, align_corners=False: ((coord + 1) * size - 1) / 2
  const float ix = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(g.x, 1.f), (float)a.w), 1.f), 0.5f);
  const float iy = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(g.y, 1.f), (float)a.h), 1.f), 0.5f);
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = (int)fx, y0 = (int)fy;
  const float wx1 = __fsub_rn(ix, fx), wy1 = __fsub_rn(iy, fy);
  const float wx0 = __fsub_rn((float)(x0 + 1), ix), wy0 = __fsub_rn((float)(y0 + 1), iy);
  // accumulation order nw, ne, sw, se as in ATen's grid_sampler_2d
  const float tw[4] = {__fmul_rn(wx0, wy0), __fmul_rn(wx1, wy0), __fmul_rn(wx0, wy1), __fmul_rn(wx1, wy1)};
  const bool inx0 = (unsigned)x0 < (unsigned)a.w, inx1 = (unsigned)(x0 + 1) < (unsigned)a.w;
  const bool iny0 = (unsigned)y0 < (unsigned)a.h, iny1 = (unsigned)(y0 + 1) < (unsigned)a.h;
  const bool tin[4] = {iny0 && inx0, iny0 && inx1, iny1 && inx0, iny1 && inx1};
  const int base = (y0 * a.w + x0) * 3;
  const int toff[4] = {base, base + 3, base + a.w * 3, base + a.w * 3 + 3};
  const int lbase = ((y0 - a.py0) * a.lw + (x0 - a.px0)) * 3;
  const int loff[4] = {lbase, lbase + 3, lbase + a.lw * 3, lbase + a.lw * 3 + 3};
  const bool lx0 = (unsigned)(x0 - a.px0) < (unsigned)a.lw, lx1 = (unsigned)(x0 + 1 - a.px0) < (unsigned)a.lw;
  const bool ly0 = (unsigned)(y0 - a.py0) < (unsigned)a.lh, ly1 = (unsigned)(y0 + 1 - a.py0) < (unsigned)a.lh;
  const bool lin[4] = {ly0 && lx0, ly0 && lx1, ly1 && lx0, ly1 && lx1};
  // expanded rectangle mask (tf_nerf.py:354-363) evaluated analytically per tap
  const bool rx0 = x0 >= a.rx0 && x0 < a.rx1, rx1 = x0 + 1 >= a.rx0 && x0 + 1 < a.rx1;
  const bool ry0 = y0 >= a.ry0 && y0 < a.ry1, ry1 = y0 + 1 >= a.ry0 && y0 + 1 < a.ry1;
  const bool rin[4] = {ry0 && rx0, ry0 && rx1, ry1 && rx0, ry1 && rx1};
  // the warped mask, per channel (grid_sample of the mask: taps nw, ne, sw, se accumulated in that order)
  float mk[4][3], macc[3] = {0.f, 0.f, 0.f};
  if (RECT) {
    float m = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if (tin[t]) m = __fadd_rn(m, __fmul_rn(rin[t] ? 1.f : 0.f, tw[t]));
    macc[0] = macc[1] = macc[2] = m;
  } else {
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
      for (int c = 0; c < 3; ++c) mk[t][c] = tin[t] ? __ldg(mask + toff[t] + c) : 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (tin[t]) macc[c] = __fadd_rn(macc[c], __fmul_rn(mk[t][c], tw[t]));
  }
  // mask[mask != 0] = 1 ; out = mask*merged + (1-mask)*gt      (tf_nerf.py:367-386): outside the mask the pixel is gt
  if (macc[0] == 0.f && macc[1] == 0.f && macc[2] == 0.f) {
#pragma unroll
    for (int c = 0; c < 3; ++c) out[c] = gtv[c];
    return;
  }
  // every remaining load of the pixel is issued before the first use (coord -> address -> gather is a dependent chain)
  float fc[4][3], lp[4][3];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (RECT) mk[t][c] = tin[t] ? __ldg(mask + toff[t] + c) : 0.f;
      fc[t][c] = tin[t] ? __ldg(face + toff[t] + c) : 0.f;
      lp[t][c] = (tin[t] && lin[t]) ? __ldg(lip + loff[t] + c) : 0.f;
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      if (tin[t]) {
        // merged_canonical = m*lip_pad + (1-m)*face   (tf_nerf.py:352), then the bilinear tap
        const float mc = __fadd_rn(__fmul_rn(mk[t][c], lp[t][c]), __fmul_rn(__fsub_rn(1.f, mk[t][c]), fc[t][c]));
        acc = __fadd_rn(acc, __fmul_rn(mc, tw[t]));
      }
    }
    out[c] = (macc[c] != 0.f) ? acc : gtv[c];
  }
}

template <bool RECT>
__global__ void __launch_bounds__(256) post_fusion_kernel(PfArgs a) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  const int npix = a.Hf * a.Wf, b = blockIdx.y;
  if (pix >= npix) return;
  const float2 g = reinterpret_cast<const float2*>(a.coord)[(size_t)b * npix + pix];
  const float* __restrict__ gtp = a.gt + ((size_t)b * npix + pix) * 3;
  const float gtv[3] = {__ldg(gtp), __ldg(gtp + 1), __ldg(gtp + 2)};
  float out[3];
  pf_pixel_value<RECT>(a, b, g, gtv, out);
#pragma unroll
  for (int c = 0; c < 3; ++c) a.fused[((size_t)b * 3 + c) * npix + pix] = out[c];
}

// Four consecutive pixels per thread (plane size a multiple of 4, 16-byte aligned bases): the streamed operands move as
// 2 + 3 LDG.128 and 3 STG.128 per thread, all loads issued before the first pixel is evaluated — 80 bytes in flight per
// thread instead of 20 (the scalar kernel at 1024 threads / SM kept ~20 KB in flight per SM, half of what HBM latency needs).
template <bool RECT>
__global__ void __launch_bounds__(256) post_fusion_kernel4(PfArgs a) {
  const int npix = a.Hf * a.Wf, b = blockIdx.y;
  const int pix = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (pix >= npix) return;
  const float4* cp = reinterpret_cast<const float4*>(a.coord + ((size_t)b * npix + pix) * 2);
  const float4* gp = reinterpret_cast<const float4*>(a.gt + ((size_t)b * npix + pix) * 3);
  const float4 c0 = __ldg(cp), c1 = __ldg(cp + 1);
  const float4 g0 = __ldg(gp), g1 = __ldg(gp + 1), g2 = __ldg(gp + 2);
  const float2 cs[4] = {make_float2(c0.x, c0.y), make_float2(c0.z, c0.w), make_float2(c1.x, c1.y), make_float2(c1.z, c1.w)};
  const float gts[4][3] = {{g0.x, g0.y, g0.z}, {g0.w, g1.x, g1.y}, {g1.z, g1.w, g2.x}, {g2.y, g2.z, g2.w}};
  float o[4][3];
#pragma unroll
  for (int q = 0; q < 4; ++q) pf_pixel_value<RECT>(a, b, cs[q], gts[q], o[q]);
#pragma unroll
  for (int c = 0; c < 3; ++c)
    *reinterpret_cast<float4*>(a.fused + ((size_t)b * 3 + c) * npix + pix) = make_float4(o[0][c], o[1][c], o[2][c], o[3][c]);
}

// (A tiled variant — the block stages merged_canonical of its tile's canonical bounding box in shared memory and the pixels
// gather from there — was built and measured: 0.41 ms vs 0.31 ms for this direct kernel on 64 frames of 500x500; the bounding-box
// reduction, the fill phase and its barriers cost more than the 4x texel reuse saves while the taps already hit L1.  Removed.)

// ---- backward w.r.t. the lip crop (training: the lip image is the only input that carries a gradient, training.py:436-445):
//   fused[c]  = [macc_c != 0] * sum_t tw_t * (m_t,c * lip_t,c + (1 - m_t,c) * face_t,c)   ->  d lip_t,c += [macc_c != 0] tw_t m_t,c dF_c
//   canon[c]  = m * lip_pad + (1 - m) * face                                             ->  d lip    += m dCanon
// d_lip is first SET by the gather kernel (the canonical term or zero), then the warp term is scattered with atomics (the
// inverse of the sampling grid is not known; ATen's grid_sampler backward does the same).
__global__ void pf_bwd_init_kernel(PfArgs a, const float* __restrict__ d_canon, float* __restrict__ d_lip) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)a.B * a.lh * a.lw * 3;
  if (gid >= n) return;
  float v = 0.f;
  if (d_canon) {
    const int c = (int)(gid % 3);
    const long long p = gid / 3;
    const int lx = (int)(p % a.lw), ly = (int)((p / a.lw) % a.lh), b = (int)(p / ((long long)a.lw * a.lh));
    const size_t idx = (((size_t)b * a.h + (ly + a.py0)) * a.w + (lx + a.px0)) * 3 + c;
    v = a.mask[idx] * d_canon[idx];
  }
  d_lip[gid] = v;
}

template <bool RECT>
__global__ void __launch_bounds__(256) pf_bwd_scatter_kernel(PfArgs a, const float* __restrict__ d_fused, float* __restrict__ d_lip) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  const int npix = a.Hf * a.Wf, b = blockIdx.y;
  if (pix >= npix) return;
  const float2 g = reinterpret_cast<const float2*>(a.coord)[(size_t)b * npix + pix];
  const float* __restrict__ mask = a.mask + (size_t)b * a.h * a.w * 3;
  float* __restrict__ dl = d_lip + (size_t)b * a.lh * a.lw * 3;
  // the same tap geometry as the forward (pf_pixel_value)
  const float ix = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(g.x, 1.f), (float)a.w), 1.f), 0.5f);
  const float iy = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(g.y, 1.f), (float)a.h), 1.f), 0.5f);
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = (int)fx, y0 = (int)fy;
  const float wx1 = __fsub_rn(ix, fx), wy1 = __fsub_rn(iy, fy);
  const float wx0 = __fsub_rn((float)(x0 + 1), ix), wy0 = __fsub_rn((float)(y0 + 1), iy);
  const float tw[4] = {__fmul_rn(wx0, wy0), __fmul_rn(wx1, wy0), __fmul_rn(wx0, wy1), __fmul_rn(wx1, wy1)};
  const int tx[4] = {x0, x0 + 1, x0, x0 + 1}, ty[4] = {y0, y0, y0 + 1, y0 + 1};
  float mk[4][3], macc[3] = {0.f, 0.f, 0.f};
  bool tin[4], lin[4], any_lip = false;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    tin[t] = (unsigned)tx[t] < (unsigned)a.w && (unsigned)ty[t] < (unsigned)a.h;
    lin[t] = tin[t] && (unsigned)(tx[t] - a.px0) < (unsigned)a.lw && (unsigned)(ty[t] - a.py0) < (unsigned)a.lh;
    any_lip |= lin[t];
  }
  if (!any_lip) return;                                    // no tap of this pixel touches the lip crop
#pragma unroll
  for (int t = 0; t < 4; ++t)
#pragma unroll
    for (int c = 0; c < 3; ++c) mk[t][c] = tin[t] ? __ldg(mask + (ty[t] * a.w + tx[t]) * 3 + c) : 0.f;
  if (RECT) {
    float m = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const bool rin = tx[t] >= a.rx0 && tx[t] < a.rx1 && ty[t] >= a.ry0 && ty[t] < a.ry1;
      if (tin[t]) m = __fadd_rn(m, __fmul_rn(rin ? 1.f : 0.f, tw[t]));
    }
    macc[0] = macc[1] = macc[2] = m;
  } else {
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (tin[t]) macc[c] = __fadd_rn(macc[c], __fmul_rn(mk[t][c], tw[t]));
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    if (macc[c] == 0.f) continue;
    const float gf = d_fused[((size_t)b * 3 + c) * npix + pix];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float v = tw[t] * mk[t][c] * gf;
      if (lin[t] && v != 0.f) atomicAdd(dl + ((ty[t] - a.py0) * a.lw + (tx[t] - a.px0)) * 3 + c, v);
    }
  }
}

__global__ void merged_canonical_kernel(PfArgs a, float* __restrict__ out) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)a.B * a.h * a.w * 3;
  if (gid >= n) return;
  const int c = (int)(gid % 3);
  const long long p = gid / 3;
  const int cx = (int)(p % a.w), cy = (int)((p / a.w) % a.h), b = (int)(p / ((long long)a.w * a.h));
  out[gid] = merged_canon(a, b, cy, cx, c);
}

}  // namespace s2l

using namespace s2l;

// shared argument validation / geometry of the forward and backward entry points
static int32_t pf_setup(PfArgs& a, const char* who, int32_t batch, int32_t lip_h, int32_t lip_w, int32_t face_h, int32_t face_w, int32_t out_h,
                        int32_t out_w, int32_t lefttop_x, int32_t lefttop_y, int32_t paste_shift, int32_t expand_pad) {
  if (batch < 0 || lip_h <= 0 || lip_w <= 0 || face_h <= 0 || face_w <= 0 || out_h <= 0 || out_w <= 0) {
    set_error("%s: bad sizes", who);
    return 2;
  }
  a.B = batch; a.lh = lip_h; a.lw = lip_w; a.h = face_h; a.w = face_w; a.Hf = out_h; a.Wf = out_w;
  // F.pad(left+1, ..., up+1, ...) with left = x-1, up = y-1 for the 'may'-style datasets, else (left, up)  (tf_nerf.py:345-350)
  a.px0 = paste_shift ? lefttop_x : lefttop_x - 1;
  a.py0 = paste_shift ? lefttop_y : lefttop_y - 1;
  if (a.px0 < 0 || a.py0 < 0 || a.px0 + lip_w > face_w || a.py0 + lip_h > face_h) {
    set_error("%s: lip crop (%dx%d at %d,%d) does not fit the %dx%d canonical face", who, lip_w, lip_h, a.px0, a.py0, face_w, face_h);
    return 2;
  }
  a.rect = expand_pad >= 0;
  if (a.rect) {
    // mask[:, y-p : y+lh+2p, x-p : x+lw+p] = 1                  (tf_nerf.py:362)
    if (lefttop_y - expand_pad < 0 || lefttop_x - expand_pad < 0) { set_error("%s: expanded mask starts outside the image", who); return 2; }
    a.ry0 = lefttop_y - expand_pad; a.ry1 = min(face_h, lefttop_y + lip_h + 2 * expand_pad);
    a.rx0 = lefttop_x - expand_pad; a.rx1 = min(face_w, lefttop_x + lip_w + expand_pad);
  }
  if ((long long)face_h * face_w * 3 >= 0x7fffffffll || (long long)out_h * out_w >= 0x7fffffffll || batch > 65535) {
    set_error("%s: image planes beyond 2^31 elements / batch beyond 65535 are not supported", who);
    return 2;
  }
  return 0;
}

extern "C" int32_t s2l_post_fusion_compose(const float* rgb_lip, const float* face_canonical, const float* rgb_gt,
                                           const float* mask_lip_canonical, const float* coord, int32_t batch, int32_t lip_h,
                                           int32_t lip_w, int32_t face_h, int32_t face_w, int32_t out_h, int32_t out_w,
                                           int32_t lefttop_x, int32_t lefttop_y, int32_t paste_shift, int32_t expand_pad,
                                           float* fused_nchw, float* merged_canonical, void* stream) {
  if (!rgb_lip || !face_canonical || !rgb_gt || !mask_lip_canonical || !coord || !fused_nchw) {
    set_error("s2l_post_fusion_compose: null argument");
    return 1;
  }
  PfArgs a{};
  a.lip = rgb_lip; a.face = face_canonical; a.gt = rgb_gt; a.mask = mask_lip_canonical; a.coord = coord; a.fused = fused_nchw;
  if (int32_t rc = pf_setup(a, "s2l_post_fusion_compose", batch, lip_h, lip_w, face_h, face_w, out_h, out_w, lefttop_x, lefttop_y, paste_shift,
                            expand_pad)) return rc;
  if (batch == 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int npix = out_h * out_w;
  const bool vec4 = (npix % 4 == 0) && ((reinterpret_cast<uintptr_t>(coord) | reinterpret_cast<uintptr_t>(rgb_gt) |
                                         reinterpret_cast<uintptr_t>(fused_nchw)) & 15) == 0;
  if (vec4) {
    const dim3 grid((unsigned)((npix / 4 + 255) / 256), (unsigned)batch);
    if (a.rect) post_fusion_kernel4<true><<<grid, 256, 0, st>>>(a);
    else post_fusion_kernel4<false><<<grid, 256, 0, st>>>(a);
  } else {
    const dim3 grid((unsigned)((npix + 255) / 256), (unsigned)batch);
    if (a.rect) post_fusion_kernel<true><<<grid, 256, 0, st>>>(a);
    else post_fusion_kernel<false><<<grid, 256, 0, st>>>(a);
  }
  if (!check_launch("post_fusion_kernel")) return 5;
  if (merged_canonical) {
    const long long m = (long long)batch * face_h * face_w * 3;
    merged_canonical_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(a, merged_canonical);
    if (!check_launch("merged_canonical_kernel")) return 5;
  }
  return 0;
}

extern "C" int32_t s2l_post_fusion_compose_bwd(const float* d_fused_nchw, const float* d_merged_canonical, const float* mask_lip_canonical,
                                               const float* coord, int32_t batch, int32_t lip_h, int32_t lip_w, int32_t face_h,
                                               int32_t face_w, int32_t out_h, int32_t out_w, int32_t lefttop_x, int32_t lefttop_y,
                                               int32_t paste_shift, int32_t expand_pad, float* d_rgb_lip, void* stream) {
  if (!mask_lip_canonical || !coord || !d_rgb_lip) { set_error("s2l_post_fusion_compose_bwd: null argument"); return 1; }
  PfArgs a{};
  a.mask = mask_lip_canonical; a.coord = coord;
  if (int32_t rc = pf_setup(a, "s2l_post_fusion_compose_bwd", batch, lip_h, lip_w, face_h, face_w, out_h, out_w, lefttop_x, lefttop_y,
                            paste_shift, expand_pad)) return rc;
  if (batch == 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long n = (long long)batch * lip_h * lip_w * 3;
  pf_bwd_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a, d_merged_canonical, d_rgb_lip);
  if (!check_launch("pf_bwd_init_kernel")) return 5;
  if (d_fused_nchw) {
    const dim3 grid((unsigned)((out_h * out_w + 255) / 256), (unsigned)batch);
    if (a.rect) pf_bwd_scatter_kernel<true><<<grid, 256, 0, st>>>(a, d_fused_nchw, d_rgb_lip);
    else pf_bwd_scatter_kernel<false><<<grid, 256, 0, st>>>(a, d_fused_nchw, d_rgb_lip);
    if (!check_launch("pf_bwd_scatter_kernel")) return 5;
  }
  return 0;
}
