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
  long long P;          // point evaluations per frame
  const float* pts;     // EXPLICIT: [F*P, uv_dims]
  const float* rays_o;  // RAYS
  const float* rays_d;
  const float* z;
};

#ifdef __CUDACC__
// Coordinates of point p (0 <= p < P) of frame f.  Arithmetic mirrors the reference op by op
// (separate fp32 roundings, no FMA contraction) so that inputs to the PE are bit-identical to what
// PyTorch computes on the GPU.
__device__ __forceinline__ void gen_point(const PointSrc& s, int f, long long p, float x[3]) {
  x[0] = x[1] = x[2] = 0.f;
  if (s.mode == S2L_PTS_GRID) {
    // get_coords, rendering.py:18-22: coords[y*W+x] = (linspace(0,1,W)[x], linspace(0,1,H)[y])
    const int px = (int)(p % s.W), py = (int)(p / s.W);
    x[0] = linspace01(px, s.W);
    x[1] = linspace01(py, s.H);
  } else if (s.mode == S2L_PTS_GRID_ENS4) {
    // training.py:195-209: taps ordered (vx,vy) = (-1,-1),(-1,1),(1,-1),(1,1);
    // coord += v*r + eps (the python scalar v*r is rounded to fp32 before the add), clamp to [0,1]
    const long long pix = p >> 2;
    const int tap = (int)(p & 3);
    const int px = (int)(pix % s.W), py = (int)(pix / s.W);
    const float vx = (tap & 2) ? 1.f : -1.f, vy = (tap & 1) ? 1.f : -1.f;
    const float rx = (float)((double)vx * (0.5 / (double)s.W));
    const float ry = (float)((double)vy * (0.5 / (double)s.H));
    const float eps = s.eps_pf ? s.eps_pf[f] : s.eps;
    const float u = __fadd_rn(linspace01(px, s.W), __fadd_rn(rx, eps));
    const float v = __fadd_rn(linspace01(py, s.H), __fadd_rn(ry, eps));
    x[0] = fminf(fmaxf(u, 0.f), 1.f);
    x[1] = fminf(fmaxf(v, 0.f), 1.f);
  } else if (s.mode == S2L_PTS_RAYS) {
    // pts = rays_o[:,None,:] + rays_d[:,None,:] * z[...,None]   (NeRF sample placement)
    const long long ray = p / s.S;
    const int smp = (int)(p % s.S);
    const long long rrow = s.rays_shared ? ray : ((long long)f * (s.P / s.S) + ray);
    const float zz = s.z_per_ray ? s.z[((long long)f * (s.P / s.S) + ray) * s.S + smp] : s.z[smp];
#pragma unroll
    for (int d = 0; d < 3; ++d)
      x[d] = __fadd_rn(s.rays_o[rrow * 3 + d], __fmul_rn(s.rays_d[rrow * 3 + d], zz));
  } else {
    const float* q = s.pts + ((long long)f * s.P + p) * s.uv_dims;
    for (int d = 0; d < s.uv_dims; ++d) x[d] = q[d];
  }
}

// 4-tap area weight of training.py:237-245 for tap `tap` of the pixel whose centre is (u0,v0) and whose
// jittered taps are c[t]; returns areas[t]+1e-9 for all taps.
__device__ __forceinline__ float ens4_area(float cu, float cv, float u0, float v0) {
  return __fadd_rn(fabsf(__fmul_rn(__fsub_rn(cu, u0), __fsub_rn(cv, v0))), 1e-9f);
}
#endif

}  // namespace s2l
