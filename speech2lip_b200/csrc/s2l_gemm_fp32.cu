// Two small fp32 CUDA-core GEMM kernels for the EXACT per-call training path (speech2lip_b200/autograd.py FusedMLPRows), so that
// no library GEMM is left on it:
//   s2l_wgrad_rows_fp32  dW[l] = dY[l]^T H[l]   ([256,N] x [N,B], reduction over the N rows: split-K slabs + ordered reduction)
//   s2l_dx_rows_fp32     dX = A1 W1 + A2 W2     ([N,256] x [256,B], the gradient w.r.t. the latent / positional-encoding columns)
// Replaces autograd's weight / input gradients of tf_nerf.py:252-283 (loss.backward(), training.py:559) in exact fp32.
// CUDA-core SGEMM tiles (up to 128x128 per CTA, 8x8 registers per thread); the tensor-core kernels (s2l_train_*.cu) are the
// throughput path, these keep the exact path self-contained.
#include "s2l_common.cuh"

namespace s2l {

constexpr int GT = 64, GK = 16;

// partial[s][l][a][b] = sum over the rows of slab s of dy[l][n][a] * h[l][n][b].
// TA x TB output tile per CTA, 256 threads as 16 x 16, each thread (TA/16) x (TB/16) outputs held as 4-wide column groups 64 apart
// (conflict-free LDS.128); 8-row stages, the next stage's global loads are issued before the current stage's FMAs (register
// prefetch) and land in the other shared-memory buffer: one __syncthreads per stage.
template <int TA, int TB>
__global__ void __launch_bounds__(256) wgrad_rows_kernel(const float* __restrict__ dy, const float* __restrict__ h, long long N, int A, int B,
                                                         long long ld_l_dy, long long ld_l_h, int slabs, float* __restrict__ partial) {
  constexpr int WK = 8;                                  // rows per stage
  constexpr int IA = TA / 64, IB = TB / 64;              // 4-wide groups per thread
  constexpr int LA = (WK * TA / 4 + 255) / 256, LB = (WK * TB / 4 + 255) / 256;   // float4 loads per thread per stage
  __shared__ __align__(16) float As[2][WK][TA], Bs[2][WK][TB];
  const int tiles_b = (B + TB - 1) / TB;
  const int a0 = (blockIdx.x / tiles_b) * TA, b0 = (blockIdx.x % tiles_b) * TB;
  const int l = blockIdx.y, s = blockIdx.z;
  const long long n0 = N * s / slabs, n1 = N * (s + 1) / slabs;
  const float* dyl = dy + (long long)l * ld_l_dy;
  const float* hl = h + (long long)l * ld_l_h;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const bool va = (A & 3) == 0, vb = (B & 3) == 0;       // rows 16-byte aligned -> vector loads
  float acc[IA * 4][IB * 4] = {};
  float4 pa[LA], pb[LB];
  auto fetch = [&](const float* src, int dim, int c_base, bool vec, long long n, int T, int i) -> float4 {
    const int e = tid + i * 256;                         // float4 index inside the stage
    const int r = e / (T / 4), c = c_base + (e % (T / 4)) * 4;
    const long long row = n + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < WK && row < n1) {
      const float* p = src + row * dim + c;
      if (vec && c + 3 < dim) v = *reinterpret_cast<const float4*>(p);
      else {
        if (c < dim) v.x = p[0];
        if (c + 1 < dim) v.y = p[1];
        if (c + 2 < dim) v.z = p[2];
        if (c + 3 < dim) v.w = p[3];
      }
    }
    return v;
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int i = 0; i < LA; ++i) { const int e = tid + i * 256; if (e < WK * TA / 4) *reinterpret_cast<float4*>(&As[buf][e / (TA / 4)][(e % (TA / 4)) * 4]) = pa[i]; }
#pragma unroll
    for (int i = 0; i < LB; ++i) { const int e = tid + i * 256; if (e < WK * TB / 4) *reinterpret_cast<float4*>(&Bs[buf][e / (TB / 4)][(e % (TB / 4)) * 4]) = pb[i]; }
  };
#pragma unroll
  for (int i = 0; i < LA; ++i) pa[i] = fetch(dyl, A, a0, va, n0, TA, i);
#pragma unroll
  for (int i = 0; i < LB; ++i) pb[i] = fetch(hl, B, b0, vb, n0, TB, i);
  stash(0);
  __syncthreads();
  int buf = 0;
  for (long long n = n0; n < n1; n += WK) {
    const bool more = n + WK < n1;
    if (more) {
#pragma unroll
      for (int i = 0; i < LA; ++i) pa[i] = fetch(dyl, A, a0, va, n + WK, TA, i);
#pragma unroll
      for (int i = 0; i < LB; ++i) pb[i] = fetch(hl, B, b0, vb, n + WK, TB, i);
    }
#pragma unroll
    for (int k = 0; k < WK; ++k) {
      float a4[IA * 4], b4[IB * 4];
#pragma unroll
      for (int i = 0; i < IA; ++i) { const float4 v = *reinterpret_cast<const float4*>(&As[buf][k][i * 64 + ty * 4]); a4[4 * i] = v.x; a4[4 * i + 1] = v.y; a4[4 * i + 2] = v.z; a4[4 * i + 3] = v.w; }
#pragma unroll
      for (int i = 0; i < IB; ++i) { const float4 v = *reinterpret_cast<const float4*>(&Bs[buf][k][i * 64 + tx * 4]); b4[4 * i] = v.x; b4[4 * i + 1] = v.y; b4[4 * i + 2] = v.z; b4[4 * i + 3] = v.w; }
#pragma unroll
      for (int i = 0; i < IA * 4; ++i)
#pragma unroll
        for (int j = 0; j < IB * 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
    }
    if (more) { stash(buf ^ 1); __syncthreads(); buf ^= 1; }
  }
  float* P = partial + ((size_t)s * gridDim.y + l) * A * B;
#pragma unroll
  for (int i = 0; i < IA * 4; ++i)
#pragma unroll
    for (int j = 0; j < IB * 4; ++j) {
      const int a = a0 + (i >> 2) * 64 + ty * 4 + (i & 3), b = b0 + (j >> 2) * 64 + tx * 4 + (j & 3);
      if (a < A && b < B) P[(size_t)a * B + b] = acc[i][j];
    }
}
__global__ void wgrad_rows_reduce_kernel(const float* __restrict__ partial, long long per_slab, int slabs, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= per_slab) return;
  float s = 0.f;
  for (int k = 0; k < slabs; ++k) s += partial[(size_t)k * per_slab + i];
  out[i] = s;
}

// out[n][c] = sum_k a1[n][k] w1[k][c] (+ a2[n][k] w2[k][c]),  k < 256, c < B; out rows have stride ld_out
__global__ void __launch_bounds__(256) dx_rows_kernel(const float* __restrict__ a1, const float* __restrict__ w1, const float* __restrict__ a2,
                                                      const float* __restrict__ w2, long long N, int B, float* __restrict__ out, int ld_out) {
  __shared__ float As[GK][GT + 4], Ws[GK][GT + 4];
  const long long n0 = (long long)blockIdx.x * GT;
  const int c0 = blockIdx.y * GT;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[4][4] = {};
  for (int pass = 0; pass < (a2 ? 2 : 1); ++pass) {
    const float* a = pass ? a2 : a1;
    const float* w = pass ? w2 : w1;
    for (int k0 = 0; k0 < 256; k0 += GK) {
      {   // A stage transposed to [k][row]: thread loads 4 consecutive k of one row
        const int r = tid >> 2, kk = (tid & 3) * 4;
        const long long row = n0 + r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < N) v = *reinterpret_cast<const float4*>(a + row * 256 + k0 + kk);
        As[kk][r] = v.x; As[kk + 1][r] = v.y; As[kk + 2][r] = v.z; As[kk + 3][r] = v.w;
        const int wr = tid >> 4, wc = (tid & 15) * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) Ws[wr][wc + j] = (c0 + wc + j < B) ? w[(size_t)(k0 + wr) * B + c0 + wc + j] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < GK; ++k) {
        const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        const float4 bv = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
        const float a4[4] = {av.x, av.y, av.z, av.w}, b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long n = n0 + ty * 4 + i;
      const int c = c0 + tx * 4 + j;
      if (n < N && c < B) out[n * ld_out + c] = acc[i][j];
    }
}

}  // namespace s2l

using namespace s2l;

// tile shape and slab count: one wave of CTAs over the 148 SMs, at least 256 rows per slab
static void wgrad_shape(long long n_rows, int n_mats, int a_dim, int b_dim, int* ta, int* tb, int* slabs) {
  *ta = a_dim > 64 ? 128 : 64;
  *tb = b_dim > 64 ? 128 : 64;
  const long long tiles = (long long)((a_dim + *ta - 1) / *ta) * ((b_dim + *tb - 1) / *tb) * n_mats;
  long long s = 148 / tiles;
  const long long cap = n_rows / 256;
  if (s > cap) s = cap;
  if (s > 64) s = 64;
  *slabs = (int)(s < 1 ? 1 : s);
}

extern "C" size_t s2l_wgrad_rows_scratch_bytes(int64_t n_rows, int32_t n_mats, int32_t a_dim, int32_t b_dim) {
  if (n_rows < 0 || n_mats < 1 || a_dim < 1 || b_dim < 1) return 0;
  int ta, tb, slabs;
  wgrad_shape(n_rows, n_mats, a_dim, b_dim, &ta, &tb, &slabs);
  return (size_t)slabs * n_mats * a_dim * b_dim * sizeof(float);
}

extern "C" int32_t s2l_wgrad_rows_fp32(const float* dy, const float* h, int64_t n_rows, int32_t n_mats, int32_t a_dim, int32_t b_dim,
                                       int64_t mat_stride_dy, int64_t mat_stride_h, float* out, void* scratch, void* stream) {
  if (!dy || !h || !out || !scratch) { set_error("s2l_wgrad_rows_fp32: null argument"); return 1; }
  if (n_rows < 0 || n_mats < 1 || a_dim < 1 || b_dim < 1) { set_error("s2l_wgrad_rows_fp32: bad sizes"); return 2; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int ta, tb, slabs;
  wgrad_shape(n_rows, n_mats, a_dim, b_dim, &ta, &tb, &slabs);
  const dim3 grid(((a_dim + ta - 1) / ta) * ((b_dim + tb - 1) / tb), n_mats, slabs);
  float* part = reinterpret_cast<float*>(scratch);
  if (ta == 128 && tb == 128) wgrad_rows_kernel<128, 128><<<grid, 256, 0, st>>>(dy, h, n_rows, a_dim, b_dim, mat_stride_dy, mat_stride_h, slabs, part);
  else if (ta == 128) wgrad_rows_kernel<128, 64><<<grid, 256, 0, st>>>(dy, h, n_rows, a_dim, b_dim, mat_stride_dy, mat_stride_h, slabs, part);
  else if (tb == 128) wgrad_rows_kernel<64, 128><<<grid, 256, 0, st>>>(dy, h, n_rows, a_dim, b_dim, mat_stride_dy, mat_stride_h, slabs, part);
  else wgrad_rows_kernel<64, 64><<<grid, 256, 0, st>>>(dy, h, n_rows, a_dim, b_dim, mat_stride_dy, mat_stride_h, slabs, part);
  if (!check_launch("wgrad_rows_kernel")) return 5;
  const long long per = (long long)n_mats * a_dim * b_dim;
  wgrad_rows_reduce_kernel<<<(unsigned)((per + 255) / 256), 256, 0, st>>>(part, per, slabs, out);
  return check_launch("wgrad_rows_reduce_kernel") ? 0 : 5;
}

extern "C" int32_t s2l_dx_rows_fp32(const float* a1, const float* w1, const float* a2, const float* w2, int64_t n_rows, int32_t b_dim,
                                    float* out, int32_t ld_out, void* stream) {
  if (!a1 || !w1 || !out || (a2 && !w2)) { set_error("s2l_dx_rows_fp32: null argument"); return 1; }
  if (n_rows < 0 || b_dim < 1 || ld_out < b_dim) { set_error("s2l_dx_rows_fp32: bad sizes"); return 2; }
  if (n_rows == 0) return 0;
  dx_rows_kernel<<<dim3((unsigned)((n_rows + GT - 1) / GT), (unsigned)((b_dim + GT - 1) / GT)), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      a1, w1, a2, w2, n_rows, b_dim, out, ld_out);
  return check_launch("dx_rows_kernel") ? 0 : 5;
}
