// fp32 CUDA-core GEMM building blocks shared by the exact forward (s2l_mlp_fp32.cu) and the backward
// data-gradient kernel (s2l_mlp_bwd.cu): 64-point tiles, [K][TMP] fp32 activation buffers in shared
// memory, W chunks ([16][256] fp32) streamed from L2 through a 3-stage bulk-copy/mbarrier ring.
#pragma once
#include "s2l_common.cuh"
#include "s2l_points.cuh"

namespace s2l {

constexpr int TM = 64;        // points per tile
constexpr int TMP = 68;       // padded row length of the [K][TM] activation buffers (bank spread, 16B aligned)
constexpr int KC = 16;        // K rows per weight chunk
constexpr int NS = 3;         // weight ring stages
constexpr int CHUNK_FLOATS = KC * 256;
constexpr int MAX_CHUNKS = 168;

struct Fp32Program {
  int n_chunks;                 // chunks per tile
  int off[MAX_CHUNKS];          // float offset of each chunk from the blob base
};

struct Fp32Args {
  const uint8_t* blob;
  Layout L;
  PointSrc src;
  const float* frame_bias;      // [F,4,256] (non-ROWLAT)
  const float* rows;            // ROWLAT: x [N, uv_dims+64]
  long long time_idx;
  int has_time;
  float* out;                   // [F*P, out_ch]
  float* save;                  // training forward: [10][F*P][256] saved activations (net, h0..h4, h_skip, h5, h6, h7) or null
  int out_ch;
  int n_frames;
  long long tiles_per_frame;
  Fp32Program prog;
  // re-evaluation of listed rays' last samples (src is a RAYS list launch with Sc = 1, s0 = S - 1):
  const float4* fix_carry;      // [F*R] (T_last, acc rgb) left by the fused compositing epilogue -> fix_rgb [F*R,3] is rewritten
  float* fix_rgb;
  float* patch_raw;             // unfused path instead: the exact outputs overwrite raw [F*R, S, 4] at sample S - 1
  Gate gate;                    // optional device-side launch gate (s2l_points.cuh)
  const long long* time_idx_dev;   // ROWLAT: time index read on the device (overrides time_idx when set)
};

struct Pipe {
  long long c;        // chunks consumed so far by this CTA
  long long total;    // chunks this CTA will consume in total
};

__device__ __forceinline__ void issue_chunk(const Fp32Args& a, float* wst, uint64_t* full, long long c) {
  const int stage = (int)(c % NS);
  const float* src = reinterpret_cast<const float*>(a.blob) + a.prog.off[c % a.prog.n_chunks];
  mbar_arrive_expect_tx(&full[stage], CHUNK_FLOATS * 4);
  bulk_g2s(wst + stage * CHUNK_FLOATS, src, CHUNK_FLOATS * 4, &full[stage]);
}

// acc[8 m][8 n] += A[k][m] * W[k][n] over `nchunks` 16-row weight chunks; A is [K][TMP] in smem.
__device__ __forceinline__ void gemm_seg(float (&acc)[8][8], const float* Abuf, int nchunks, Pipe& ps,
                                         const Fp32Args& a, float* wst, uint64_t* full, int m0, int tn) {
  for (int j = 0; j < nchunks; ++j) {
    const int stage = (int)(ps.c % NS);
    mbar_wait(&full[stage], (uint32_t)((ps.c / NS) & 1));
    const float* Wc = wst + stage * CHUNK_FLOATS;
    const float* Ac = Abuf + (size_t)j * KC * TMP + m0;
#pragma unroll 4
    for (int kk = 0; kk < KC; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(Ac + kk * TMP);
      const float4 a1 = *reinterpret_cast<const float4*>(Ac + kk * TMP + 4);
      const float4 w0 = *reinterpret_cast<const float4*>(Wc + kk * 256 + 4 * tn);
      const float4 w1 = *reinterpret_cast<const float4*>(Wc + kk * 256 + 128 + 4 * tn);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int n = 0; n < 8; ++n) acc[i][n] = fmaf(av[i], wv[n], acc[i][n]);
    }
    __syncthreads();   // every warp is done with this stage (and, after the last chunk, with Abuf)
    if (threadIdx.x == 0 && ps.c + NS < ps.total) issue_chunk(a, wst, full, ps.c + NS);
    ps.c++;
  }
}

__device__ __forceinline__ void zero_acc(float (&acc)[8][8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int n = 0; n < 8; ++n) acc[i][n] = 0.f;
}

// Out[n][m] = act(acc + bias[n]); Out is [256][TMP] in smem.  If `grow` is non-null the same values are also
// written row-major to global memory (grow = row pointer of the tile's first point, `rows_valid` rows exist).
__device__ __forceinline__ void store_acc(const float (&acc)[8][8], float* Out, const float* bias, bool relu,
                                          int m0, int tn, float* grow = nullptr, int rows_valid = 0) {
  float v[8][8];
#pragma unroll
  for (int jn = 0; jn < 8; ++jn) {
    const int n = (jn < 4) ? (4 * tn + jn) : (128 + 4 * tn + jn - 4);
    const float b = bias[n];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[i][jn] = acc[i][jn] + b;
      if (relu) v[i][jn] = fmaxf(v[i][jn], 0.f);
    }
    *reinterpret_cast<float4*>(Out + n * TMP + m0) = make_float4(v[0][jn], v[1][jn], v[2][jn], v[3][jn]);
    *reinterpret_cast<float4*>(Out + n * TMP + m0 + 4) = make_float4(v[4][jn], v[5][jn], v[6][jn], v[7][jn]);
  }
  if (grow) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (m0 + i < rows_valid) {
        float* r = grow + (size_t)(m0 + i) * 256;
        *reinterpret_cast<float4*>(r + 4 * tn) = make_float4(v[i][0], v[i][1], v[i][2], v[i][3]);
        *reinterpret_cast<float4*>(r + 128 + 4 * tn) = make_float4(v[i][4], v[i][5], v[i][6], v[i][7]);
      }
    }
  }
}


}  // namespace s2l
