// Point generation shared by the fp32 and tensor-core MLP kernels: where does point-evaluation p of
// frame f sit in the MLP's input space, and its positional encoding.
#pragma once
#include "s2l_common.cuh"

namespace s2l {

struct PointSrc {
  int mode;             // S2L_PTS_*
  int H, W, S;
  int uv_dims;
  int z_per_ray;
  int rays_shared;
  float eps;            // GRID_ENS4 eps_shift
  const float* eps_pf;  // optional per-frame eps_shift [F]
  long long P;          // point evaluations per frame (RAYS: R * Sc)
  const float* pts;     // EXPLICIT: [F*P, uv_dims] (row stride pts_stride floats; 0 = uv_dims)
  int pts_stride;
  float step_w, step_h; // GRID*: 1.0f / (W - 1), 1.0f / (H - 1) (fp32 division on the host = ATen's linspace step; saves a division
                        // routine per coordinate in the kernels)
  const float* rays_o;  // RAYS
  const float* rays_d;
  const float* z;
  // RAYS, sample chunks and compacted ray lists (early ray termination, the fp32 re-evaluation of near-zero last
  // samples): a launch evaluates samples [s0, s0 + Sc) of every ray (Sc == S, s0 == 0: whole rays) or of the rays in
  // a per-frame device list; point p of frame f = sample s0 + p % Sc of ray slot p / Sc.
  int R;                     // rays per frame (H * W)
  int s0, Sc;
  const int* list_count;     // optional [F]: number of rays of frame f in the list (device)
  const int* list_rays;      // optional [F][R]: ray index within the frame
  const int* tile_start;     // with a list: [F + 1] first tile of every frame for this kernel's tile size (device)
};

// Per-pixel reduction fused into the tensor-core kernels' output-layer epilogue (reduce_tile, s2l_tc_common.cuh)
enum { EPI_RAW = 0, EPI_ENS4 = 1, EPI_COMPOSITE = 2 };
// Device-side launch gate: a kernel whose gate pointer is set does its work only when *gate == gate_value and exits at
// once otherwise.  Lets a host call enqueue BOTH implementations of a data-dependent choice (constant-latent tensor-core
// path / general per-row path of the drop-in rgb_forward) without reading the deciding flag back.
struct Gate {
  const int* flag;
  int value;
};
struct TcEpi {
  int mode;                  // EPI_*
  float* rgb;                // [F, H*W, 3]
  float4* carry;             // EPI_COMPOSITE [F*R]: (T, acc rgb) of a ray between sample chunks / for the fp32 re-evaluation
  int* next_count;           // list this launch appends to: [F] counts (zeroed by the caller) ...
  int* next_rays;            // ... and [F][R] ray indices
  float term_thr, fix_thr;
};

#ifdef __CUDACC__
// Ray (index within the frame) that point p of a RAYS launch belongs to.
__device__ __forceinline__ int ray_of_point(const PointSrc& s, int f, long long p) {
  const long long slot = p / s.Sc;
  return s.list_rays ? s.list_rays[(long long)f * s.R + slot] : (int)slot;
}

// Coordinates of point p (0 <= p < P) of frame f.  Arithmetic mirrors the reference op by op
// (separate fp32 roundings, no FMA contraction) so that inputs to the PE are bit-identical to what
// PyTorch computes on the GPU.
// One tap of the 4-tap local ensemble (training.py:195-209): taps ordered (vx,vy) = (-1,-1),(-1,1),(1,-1),(1,1);
// coord += v*r + eps (the python scalar v*r is rounded to fp32 before the add), clamp to [0,1].  (u0, v0) = the pixel centre.
__device__ __forceinline__ void ens4_tap(const PointSrc& s, float u0, float v0, float eps, int tap, float& u, float& v) {
  const float rx = (float)(0.5 / (double)s.W), ry = (float)(0.5 / (double)s.H);      // |v*r| rounded to fp32 (exact sign symmetry)
  u = fminf(fmaxf(__fadd_rn(u0, __fadd_rn((tap & 2) ? rx : -rx, eps)), 0.f), 1.f);
  v = fminf(fmaxf(__fadd_rn(v0, __fadd_rn((tap & 1) ? ry : -ry, eps)), 0.f), 1.f);
}
__device__ __forceinline__ void grid_centre(const PointSrc& s, unsigned pix, float& u0, float& v0) {
  const unsigned py = pix / (unsigned)s.W, px = pix - py * (unsigned)s.W;               // pixel indices fit 32 bits
  u0 = linspace01((int)px, s.W, s.step_w);
  v0 = linspace01((int)py, s.H, s.step_h);
}

__device__ __forceinline__ void gen_point(const PointSrc& s, int f, long long p, float x[3]) {
  x[0] = x[1] = x[2] = 0.f;
  if (s.mode == S2L_PTS_GRID) {
    // get_coords, rendering.py:18-22: coords[y*W+x] = (linspace(0,1,W)[x], linspace(0,1,H)[y])
    grid_centre(s, (unsigned)p, x[0], x[1]);
  } else if (s.mode == S2L_PTS_GRID_ENS4) {
    float u0, v0;
    grid_centre(s, (unsigned)(p >> 2), u0, v0);
    ens4_tap(s, u0, v0, s.eps_pf ? s.eps_pf[f] : s.eps, (int)(p & 3), x[0], x[1]);
  } else if (s.mode == S2L_PTS_RAYS) {
    // pts = rays_o[:,None,:] + rays_d[:,None,:] * z[...,None]   (NeRF sample placement)
    const long long ray = ray_of_point(s, f, p);
    const int smp = s.s0 + (int)(p % s.Sc);
    const long long rrow = s.rays_shared ? ray : ((long long)f * s.R + ray);
    const float zz = s.z_per_ray ? s.z[((long long)f * s.R + ray) * s.S + smp] : s.z[smp];
#pragma unroll
    for (int d = 0; d < 3; ++d)
      x[d] = __fadd_rn(s.rays_o[rrow * 3 + d], __fmul_rn(s.rays_d[rrow * 3 + d], zz));
  } else {
    const float* q = s.pts + ((long long)f * s.P + p) * (s.pts_stride ? s.pts_stride : s.uv_dims);
    for (int d = 0; d < s.uv_dims; ++d) x[d] = q[d];
  }
}

// Tile -> (frame, first point, points of that frame in this launch) for a kernel with TMX-point tiles.  Uniform launches
// have tiles_per_frame tiles in every frame; list launches read the per-frame tile ranges from tile_start (tiles visited
// by one CTA increase, so `fcur` is a running frame pointer).  Tiles >= n_tiles are dead (Pf = 0).
template <int TMX>
__device__ __forceinline__ void tile_locate(const PointSrc& s, long long tiles_per_frame, long long n_tiles, long long tile,
                                            int& fcur, int& f, long long& p0, long long& Pf) {
  if (tile >= n_tiles) { f = 0; p0 = 0; Pf = 0; return; }
  if (s.tile_start) {
    while (tile >= s.tile_start[fcur + 1]) ++fcur;
    f = fcur;
    p0 = (tile - s.tile_start[f]) * TMX;
    Pf = (long long)s.list_count[f] * s.Sc;
  } else if (n_tiles < 0x7fffffffll) {          // 32-bit division (the 64-bit one is a ~100-instruction routine)
    const unsigned t32 = (unsigned)tile, tpf = (unsigned)tiles_per_frame;
    const unsigned q = t32 / tpf;
    f = (int)q;
    p0 = (long long)(t32 - q * tpf) * TMX;
    Pf = s.P;
  } else {
    f = (int)(tile / tiles_per_frame);
    p0 = (tile % tiles_per_frame) * TMX;
    Pf = s.P;
  }
}

// The four blend weights of a pixel (training.py:237-249): area of every tap's rectangle against the pixel centre (+1e-9),
// normalised by their sum, and used SWAPPED (tap t is weighted with the area of tap 3 - t).
__device__ __forceinline__ void ens4_weights(const PointSrc& s, int f, unsigned pix, float (&w)[4]);

// 4-tap area weight of training.py:237-245 for tap `tap` of the pixel whose centre is (u0,v0) and whose
// jittered taps are c[t]; returns areas[t]+1e-9 for all taps.
__device__ __forceinline__ float ens4_area(float cu, float cv, float u0, float v0) {
  return __fadd_rn(fabsf(__fmul_rn(__fsub_rn(cu, u0), __fsub_rn(cv, v0))), 1e-9f);
}
__device__ __forceinline__ void ens4_weights(const PointSrc& s, int f, unsigned pix, float (&w)[4]) {
  float u0, v0, area[4];
  grid_centre(s, pix, u0, v0);
  const float eps = s.eps_pf ? s.eps_pf[f] : s.eps;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    float u, v;
    ens4_tap(s, u0, v0, eps, t, u, v);
    area[t] = ens4_area(u, v, u0, v0);
  }
  const float tot = __fadd_rn(__fadd_rn(__fadd_rn(area[0], area[1]), area[2]), area[3]);
#pragma unroll
  for (int t = 0; t < 4; ++t) w[t] = __fdiv_rn(area[3 - t], tot);
}
#endif

}  // namespace s2l
