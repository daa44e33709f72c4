// Backward of the fused MLP, data-gradient half (first step: exact fp32 path).
//
// Given dL/d(out) [N,out_ch] and the activations saved by the training forward (s2l_rgb_forward_rows_train:
// net, h0..h4, h_skip, h5, h6, h7 — 10 x [N,256] fp32), one persistent kernel walks 64-row tiles and
// back-propagates through output_linear and pts_linears 7..0 without ever leaving shared memory:
//     dPre_l = (dPre_{l+1} * W_{l+1}) .* (h_l > 0)          (tf_nerf.py:265-283 reversed)
// with the skip split at layer 5 (d h_skip = dPre5 * W5[:, :256], d h4 = dPre5 * W5[:, 256:]).
// Every dPre_l (plus d h_skip and d net) is written once to HBM ([10][N,256]); the weight gradients are
// then plain [256,N]x[N,256] GEMMs over those buffers (library GEMMs on the host side), and the per-row
// input gradients follow from d net / d h_skip.  Same 8x8 register tiles, [K][64] smem buffers and 3-stage
// bulk-copy weight ring as the forward (s2l_fp32_core.cuh); the streamed operand is the UNtransposed
// weight ([out][in], blob section DGRAD) because dX = dY * W contracts over `out`.
#include "s2l_fp32_core.cuh"

namespace s2l {

struct BwdArgs {
  Fp32Args core;           // blob / program / n_rows (core.src.P) — only the fields gemm_seg uses
  const float* d_out;      // [N,out_ch]
  const float* acts;       // [10][N][256]
  float* dsave;            // [10][N][256]: 0 d net, 1..5 dPre0..4, 6 d h_skip, 7 dPre5, 8 dPre6, 9 dPre7
};

// Out[n][m] = mask ? acc : 0 (mask = saved activation > 0), also written row-major to global
__device__ __forceinline__ void store_grad(const float (&acc)[8][8], float* Out, const float* act_rows, float* g_rows,
                                           int rows_valid, int m0, int tn) {
  float v[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float4 a0 = make_float4(1.f, 1.f, 1.f, 1.f), a1 = a0;
    const bool ok = (m0 + i) < rows_valid;
    if (act_rows && ok) {
      const float* r = act_rows + (size_t)(m0 + i) * 256;
      a0 = *reinterpret_cast<const float4*>(r + 4 * tn);
      a1 = *reinterpret_cast<const float4*>(r + 128 + 4 * tn);
    }
    const float mk[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
    for (int jn = 0; jn < 8; ++jn) v[i][jn] = (ok && mk[jn] > 0.f) ? acc[i][jn] : 0.f;
    if (g_rows && ok) {
      float* r = g_rows + (size_t)(m0 + i) * 256;
      *reinterpret_cast<float4*>(r + 4 * tn) = make_float4(v[i][0], v[i][1], v[i][2], v[i][3]);
      *reinterpret_cast<float4*>(r + 128 + 4 * tn) = make_float4(v[i][4], v[i][5], v[i][6], v[i][7]);
    }
  }
  if (Out) {
#pragma unroll
    for (int jn = 0; jn < 8; ++jn) {
      const int n = (jn < 4) ? (4 * tn + jn) : (128 + 4 * tn + jn - 4);
      *reinterpret_cast<float4*>(Out + n * TMP + m0) = make_float4(v[0][jn], v[1][jn], v[2][jn], v[3][jn]);
      *reinterpret_cast<float4*>(Out + n * TMP + m0 + 4) = make_float4(v[4][jn], v[5][jn], v[6][jn], v[7][jn]);
    }
  }
}

__global__ void __launch_bounds__(256, 1) mlp_bwd_kernel(const __grid_constant__ BwdArgs b) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  float* X = reinterpret_cast<float*>(smem_raw);
  float* Y = X + 256 * TMP;
  float* wst = Y + 256 * TMP;
  __shared__ uint64_t full[NS];
  __shared__ float dout_s[TM * 4];
  const Fp32Args& a = b.core;
  const int tid = threadIdx.x;
  const int tn = tid & 31, m0 = (tid >> 5) * 8;
  const float* Fp = reinterpret_cast<const float*>(a.blob + a.L.off_fp32);
  const long long N = a.src.P;
  const long long n_tiles = (N + TM - 1) / TM;
  const long long my_tiles = (n_tiles > blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  Pipe ps{0, my_tiles * a.prog.n_chunks};
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0)
    for (long long c = 0; c < NS && c < ps.total; ++c) issue_chunk(a, wst, full, c);

  float acc[8][8];
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long p_base = tile * TM;
    const int rows_valid = (int)((N - p_base) < TM ? (N - p_base) : TM);
    auto act = [&](int slot) -> const float* { return b.acts + ((size_t)slot * N + p_base) * 256; };
    auto dsv = [&](int slot) -> float* { return b.dsave + ((size_t)slot * N + p_base) * 256; };
    for (int i = tid; i < TM * 4; i += 256) {
      const int m = i >> 2, c = i & 3;
      dout_s[i] = (m < rows_valid && c < a.out_ch) ? b.d_out[(p_base + m) * a.out_ch + c] : 0.f;
    }
    __syncthreads();
    // ---- d h7 = d out * W_out, masked by h7 > 0  -> dPre7 (X)
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int jn = 0; jn < 8; ++jn) {
        const int n = (jn < 4) ? (4 * tn + jn) : (128 + 4 * tn + jn - 4);
        float s = 0.f;
        for (int c = 0; c < a.out_ch; ++c) s = fmaf(dout_s[(m0 + i) * 4 + c], __ldg(Fp + F_OUT_W + c * 256 + n), s);
        acc[i][jn] = s;
      }
    store_grad(acc, X, act(9), dsv(9), rows_valid, m0, tn);
    __syncthreads();
    // ---- pts_linears.7 and .6:  dPre6 (Y) = dPre7 * W7 .* (h6 > 0);  dPre5 (X) = dPre6 * W6 .* (h5 > 0)
    zero_acc(acc);
    gemm_seg(acc, X, 16, ps, a, wst, full, m0, tn);
    store_grad(acc, Y, act(8), dsv(8), rows_valid, m0, tn);
    __syncthreads();
    zero_acc(acc);
    gemm_seg(acc, Y, 16, ps, a, wst, full, m0, tn);
    store_grad(acc, X, act(7), dsv(7), rows_valid, m0, tn);
    __syncthreads();
    // ---- skip split: d h_skip = dPre5 * W5[:, :256] (no activation on h_skip -> no mask), global only
    zero_acc(acc);
    gemm_seg(acc, X, 16, ps, a, wst, full, m0, tn);
    store_grad(acc, nullptr, nullptr, dsv(6), rows_valid, m0, tn);
    // ---- d h4 = dPre5 * W5[:, 256:], masked by h4 > 0 -> dPre4 (Y)
    zero_acc(acc);
    gemm_seg(acc, X, 16, ps, a, wst, full, m0, tn);
    store_grad(acc, Y, act(5), dsv(5), rows_valid, m0, tn);
    __syncthreads();
    // ---- pts_linears 4..1: dPre_{l-1} = dPre_l * W_l .* (h_{l-1} > 0)
    float* in = Y;
    float* out = X;
    for (int l = 4; l >= 1; --l) {
      zero_acc(acc);
      gemm_seg(acc, in, 16, ps, a, wst, full, m0, tn);
      store_grad(acc, out, act(l), dsv(l), rows_valid, m0, tn);
      __syncthreads();
      float* t = in; in = out; out = t;
    }
    // ---- d net = dPre0 * W0 (net has no activation), global only
    zero_acc(acc);
    gemm_seg(acc, in, 16, ps, a, wst, full, m0, tn);
    store_grad(acc, nullptr, nullptr, dsv(0), rows_valid, m0, tn);
    __syncthreads();
  }
}

// Embedder.__call__ (tf_nerf.py:404-425) as a standalone kernel: x rows (stride row_stride) -> pe [N, E]
__global__ void embed_kernel(const float* __restrict__ x, long long n, int row_stride, int D, float* __restrict__ pe) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int E = D + 2 * kMultires * D;
  if (gid >= n * (kMultires + 1)) return;
  const long long r = gid / (kMultires + 1);
  const int k = (int)(gid % (kMultires + 1)) - 1;
  for (int d = 0; d < D; ++d) {
    const float v = x[r * row_stride + d];
    if (k < 0) {
      pe[r * E + d] = v;
    } else {
      float sn, cs;
      sincosf(__fmul_rn(v, (float)(1 << k)), &sn, &cs);
      pe[r * E + D + (2 * k) * D + d] = sn;
      pe[r * E + D + (2 * k + 1) * D + d] = cs;
    }
  }
}

int launch_mlp_bwd_rows(const void* blob, const float* d_out, const float* acts, long long n_rows, float* dsave,
                        int out_ch, cudaStream_t st) {
  if (n_rows == 0) return 0;
  BwdArgs b{};
  b.core.blob = reinterpret_cast<const uint8_t*>(blob);
  b.core.L = blob_layout();
  b.core.src.P = n_rows;
  b.core.out_ch = out_ch;
  b.d_out = d_out;
  b.acts = acts;
  b.dsave = dsave;
  // chunk program: W7, W6, W5a, W5b, W4, W3, W2, W1, W0 (DGRAD slots 8, 7, 5, 6, 4, 3, 2, 1, 0), 16 chunks each
  const int dg = (int)(b.core.L.off_dgrad / 4);
  const int order[9] = {8, 7, 5, 6, 4, 3, 2, 1, 0};
  int n = 0;
  for (int s = 0; s < 9; ++s)
    for (int j = 0; j < 16; ++j) b.core.prog.off[n++] = dg + d_slot_off(order[s]) + j * CHUNK_FLOATS;
  b.core.prog.n_chunks = n;
  const size_t smem = sizeof(float) * (size_t)(2 * 256 * TMP + NS * CHUNK_FLOATS);
  static bool attr_set_dev[64] = {};   // cudaFuncSetAttribute is per device
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  bool& attr_set = attr_set_dev[cur_dev & 63];
  if (!attr_set) {
    if (cudaFuncSetAttribute(mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      set_error("mlp_bwd: cannot opt in to %zu B of shared memory: %s", smem, cudaGetErrorString(cudaGetLastError()));
      return 6;
    }
    attr_set = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long n_tiles = (n_rows + TM - 1) / TM;
  mlp_bwd_kernel<<<(unsigned)(n_tiles < sms ? n_tiles : sms), 256, smem, st>>>(b);
  return check_launch("mlp_bwd_kernel") ? 0 : 5;
}

int launch_embed(const float* x, long long n_rows, int row_stride, int uv_dims, float* pe, cudaStream_t st) {
  if (n_rows == 0) return 0;
  const long long n = n_rows * (kMultires + 1);
  embed_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, n_rows, row_stride, uv_dims, pe);
  return check_launch("embed_kernel") ? 0 : 5;
}

}  // namespace s2l
