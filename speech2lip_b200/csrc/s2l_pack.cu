// s2l_pack_weights: PyTorch-layout fp32 parameters -> kernel-layout blob (see s2l_common.cuh).
// Replaces the parameter inventory of TalkingFace.__init__ (tf_nerf.py:85-172) on the device side.
#include <cmath>
#include <cstring>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include "s2l_common.cuh"

namespace s2l {

struct ParamPtrs {
  const float* p[S2L_NUM_PARAMS];
};
struct DivTerm {
  float v[16];
};

__device__ __forceinline__ const float* pts_w(const ParamPtrs& P, int i) { return P.p[S2L_P_PTS0_W + 2 * i]; }
__device__ __forceinline__ const float* pts_b(const ParamPtrs& P, int i) { return P.p[S2L_P_PTS0_W + 2 * i + 1]; }

// ---- AUDIO + CONST (except folds) + FP32 + TCBIAS: pure re-layout, one grid-stride pass per section
__global__ void pack_relayout_kernel(ParamPtrs P, uint8_t* blob, Layout L, int E, int out_ch, DivTerm div_term) {
  float* A = reinterpret_cast<float*>(blob + L.off_audio);
  float* C = reinterpret_cast<float*>(blob + L.off_const);
  float* Fp = reinterpret_cast<float*>(blob + L.off_fp32);
  float* TB = reinterpret_cast<float*>(blob + L.off_tcbias);
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nth = gridDim.x * blockDim.x;

  // AUDIO: straight copies
  const int a_off[13] = {A_CONV0_W, A_CONV0_B, A_CONV1_W, A_CONV1_B, A_CONV2_W, A_CONV2_B, A_CONV3_W,
                         A_CONV3_B, A_FC1_W,   A_FC1_B,   A_FC2_W,   A_FC2_B,   A_TOTAL};
  for (int t = 0; t < 12; ++t) {
    const float* src = P.p[S2L_P_CONV0_W + t];
    const int n = a_off[t + 1] - a_off[t];
    for (int i = tid; i < n; i += nth) A[a_off[t] + i] = src[i];
  }
  // CONST: transposed per-frame mat-vec weights
  for (int i = tid; i < 64 * 256; i += nth) {
    const int k = i / 256, n = i % 256;
    C[C_FCA_WT + i] = P.p[S2L_P_FC_AUDIO_W][n * 64 + k];
    C[C_FCAS_WT + i] = P.p[S2L_P_FC_AUDIO_SKIP_W][n * 64 + k];
  }
  for (int i = tid; i < 20 * 256; i += nth) {
    const int k = i / 256, n = i % 256;
    C[C_FCT_WT + i] = P.p[S2L_P_FC_TIME_W][n * 20 + k];
    C[C_FCTS_WT + i] = P.p[S2L_P_FC_TIME_SKIP_W][n * 20 + k];
  }
  for (int i = tid; i < 256; i += nth) {
    C[C_BIAS6 + 0 * 256 + i] = P.p[S2L_P_FC_UV_B][i];
    C[C_BIAS6 + 1 * 256 + i] = P.p[S2L_P_FC_AUDIO_B][i];
    C[C_BIAS6 + 2 * 256 + i] = P.p[S2L_P_FC_TIME_B][i];
    C[C_BIAS6 + 3 * 256 + i] = P.p[S2L_P_FC_UV_SKIP_B][i];
    C[C_BIAS6 + 4 * 256 + i] = P.p[S2L_P_FC_AUDIO_SKIP_B][i];
    C[C_BIAS6 + 5 * 256 + i] = P.p[S2L_P_FC_TIME_SKIP_B][i];
  }
  for (int i = tid; i < 16; i += nth) C[C_DIV + i] = (i < 10) ? div_term.v[i] : 0.f;
  // META: cleared here (this kernel runs first), filled by pack_tcw_kernel's atomics; the density-row norm by block 0
  {
    float* M = reinterpret_cast<float*>(blob + L.off_meta);
    for (int i = tid; i < kMetaWords; i += nth) M[i] = 0.f;
  }

  // FP32: W^T [K][256]
  for (int i = tid; i < 64 * 256; i += nth) {
    const int k = i / 256, n = i % 256;
    Fp[F_UV_WT + i] = (k < E) ? P.p[S2L_P_FC_UV_W][n * E + k] : 0.f;
    Fp[F_UVS_WT + i] = (k < E) ? P.p[S2L_P_FC_UV_SKIP_W][n * E + k] : 0.f;
  }
  for (int l = 0; l < 8; ++l) {
    const int K = (l == 5) ? 512 : 256;
    const float* w = pts_w(P, l);
    float* dst = Fp + f_pts_off(l);
    for (int i = tid; i < K * 256; i += nth) {
      const int k = i / 256, n = i % 256;
      dst[i] = w[n * K + k];
    }
    for (int i = tid; i < 256; i += nth) Fp[F_PTS_B + l * 256 + i] = pts_b(P, l)[i];
  }
  for (int i = tid; i < 4 * 256; i += nth) {
    const int n = i / 256, k = i % 256;
    Fp[F_OUT_W + i] = (n < out_ch) ? P.p[S2L_P_OUT_W][n * 256 + k] : 0.f;
  }
  for (int i = tid; i < 4; i += nth) Fp[F_OUT_B + i] = (i < out_ch) ? P.p[S2L_P_OUT_B][i] : 0.f;

  // DGRAD: untransposed weights for dX = dY * W  (slot 5 = pts_linears.5[:, :256], slot 6 = pts_linears.5[:, 256:])
  {
    float* Dg = reinterpret_cast<float*>(blob + L.off_dgrad);
    for (int slot = 0; slot < D_SLOTS; ++slot) {
      const int l = slot <= 4 ? slot : (slot <= 6 ? 5 : slot - 1);
      const int ld = (l == 5) ? 512 : 256, col0 = (slot == 6) ? 256 : 0;
      const float* w = pts_w(P, l);
      for (int i = tid; i < 65536; i += nth) Dg[d_slot_off(slot) + i] = w[(i >> 8) * ld + col0 + (i & 255)];
    }
  }

  // TCBIAS [9][256]: G0/G5 rows unused (per-frame folded bias), G8 = output bias padded
  for (int i = tid; i < kNumG * 256; i += nth) {
    const int g = i / 256, n = i % 256;
    float v = 0.f;
    if (g >= 1 && g <= 7 && g != 5) v = pts_b(P, g)[n];
    if (g == 8 && n < out_ch) v = P.p[S2L_P_OUT_B][n];
    TB[i] = v;
  }
}

// ---- folded input weights (two back-to-back Linear layers without a nonlinearity between them,
//      tf_nerf.py:252-266 and :268-281):  fold0 = W0 * Wuv,  fold5 = W5[:, :256] * Wuv_skip.  fp64 accumulate.
__global__ void pack_fold_kernel(ParamPtrs P, uint8_t* blob, Layout L, int E) {
  float* C = reinterpret_cast<float*>(blob + L.off_const);
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // 2 * 256 * 64
  if (idx >= 2 * 256 * 64) return;
  const int which = idx / (256 * 64);
  const int n = (idx / 64) % 256, e = idx % 64;
  double acc = 0.0;
  if (e < E) {
    const float* Wl = which == 0 ? pts_w(P, 0) : pts_w(P, 5);
    const int ldl = which == 0 ? 256 : 512;
    const float* Wi = which == 0 ? P.p[S2L_P_FC_UV_W] : P.p[S2L_P_FC_UV_SKIP_W];
    for (int k = 0; k < 256; ++k) acc += (double)Wl[n * ldl + k] * (double)Wi[k * E + e];
  }
  C[(which == 0 ? C_FOLD0 : C_FOLD5) + n * 64 + e] = (float)acc;
}

// ---- tensor-core operand granules: hi/lo bf16 split, K-major SW128 smem image, MMA issue order
__global__ void pack_tcw_kernel(ParamPtrs P, uint8_t* blob, Layout L, int out_ch) {
  const float* C = reinterpret_cast<const float*>(blob + L.off_const);
  uint8_t* T = blob + L.off_tcw;
  // one thread per (granule, row, k) element; granule index space = sum over layers
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  // decode: layers 0..7 have 2*nkc granules of 128x64 (half h, K-chunk kc); layer 8 has 4 granules of 16x64
  long long rem = gid;
  int g = 0, rows = kGranRows;
  for (; g < kNumG; ++g) {
    rows = (g == 8) ? kOutPad : kGranRows;
    const long long cnt = (long long)((g == 8) ? 4 : 2 * g_nkc(g)) * rows * 64;
    if (rem < cnt) break;
    rem -= cnt;
  }
  if (g >= kNumG) return;
  const int per_gran = rows * 64;
  const int gi = (int)(rem / per_gran);
  const int r = (int)(rem % per_gran);
  const int n = r / 64, k = r % 64;
  const int nkc = g_nkc(g);
  const int h = (g == 8) ? 0 : gi / nkc;
  const int kc = (g == 8) ? gi : gi % nkc;
  const int ng = h * kGranRows + n;
  float v;
  if (g == 0) {
    v = C[C_FOLD0 + ng * 64 + k];
  } else if (g == 5) {
    v = (kc == 0) ? C[C_FOLD5 + ng * 64 + k] : pts_w(P, 5)[ng * 512 + 256 + (kc - 1) * 64 + k];
  } else if (g == 8) {
    v = (n < out_ch) ? P.p[S2L_P_OUT_W][n * 256 + kc * 64 + k] : 0.f;
  } else {
    v = pts_w(P, g)[ng * 256 + kc * 64 + k];
  }
  {
    // fp16f8 domain bookkeeping (non-negative floats order like their bit patterns)
    unsigned int* M = reinterpret_cast<unsigned int*>(blob + L.off_meta);
    const float av = fabsf(v);
    const unsigned act = __activemask();
    const unsigned wmax = __reduce_max_sync(act, __float_as_uint(av));
    const unsigned nsat = __popc(__ballot_sync(act, !(av < kF8MaxWeight)));
    if ((threadIdx.x & 31) == (__ffs(act) - 1)) {      // one pair of atomics per warp
      if (wmax) atomicMax(M + 0, wmax);
      if (nsat) atomicAdd(M + 1, nsat);
    }
  }
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
  const int plane = (g == 8) ? kOutPlane : kGranPlane;
  uint8_t* base = T + g_layer_off(g) + (size_t)gi * (2 * plane);
  const int off = sw128_off(n, k);
  *reinterpret_cast<__nv_bfloat16*>(base + off) = hi;
  *reinterpret_cast<__nv_bfloat16*>(base + plane + off) = lo;
  // ---- TCW8: fp16 main operand + two scaled fp8 correction operands (see s2l_common.cuh)
  uint8_t* base8 = blob + L.off_tcw8 + g_layer_off(g) + (size_t)gi * (2 * plane);
  const __half w1 = __float2half_rn(v);
  const float w1f = __half2float(w1), w2f = v - w1f;
  *reinterpret_cast<__half*>(base8 + off) = w1;
  const int off8 = sw64_off(n, k);
  base8[plane + off8] = (uint8_t)__nv_cvt_float_to_fp8(w1f * (1.0f / (float)(1 << kScaleA)), __NV_SATFINITE, __NV_E5M2);
  base8[plane + plane / 2 + off8] = (uint8_t)__nv_cvt_float_to_fp8(w2f * (float)(1 << kScaleW), __NV_SATFINITE, __NV_E4M3);
}

// ---- TCWT: transposed bf16 granules for the backward data-gradient GEMMs (see s2l_common.cuh)
__global__ void pack_tcwt_kernel(ParamPtrs P, uint8_t* blob, Layout L, int out_ch) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // one thread per element
  const long long per_gran = 128 * 64;
  if (gid >= (long long)(2 + kTLayers * 8) * per_gran) return;
  const int gran = (int)(gid / per_gran);
  const int r = (int)(gid % per_gran);
  const int n = r / 64, k = r % 64;
  float v;
  if (gran < 2) {                       // output_linear^T: row = hidden channel, k = output channel
    const int ni = gran * 128 + n;
    v = (k < out_ch) ? P.p[S2L_P_OUT_W][k * 256 + ni] : 0.f;
  } else {
    const int i = (gran - 2) / 8, gi = (gran - 2) % 8;
    const int h = gi / 4, kc = gi % 4;
    const int l = t_layer_src(i);
    const int ni = h * 128 + n, ko = kc * 64 + k;
    v = (l == 5) ? pts_w(P, 5)[ko * 512 + 256 + ni] : pts_w(P, l)[ko * 256 + ni];
  }
  uint8_t* base = blob + L.off_tcwt + (size_t)gran * kTGran;
  *reinterpret_cast<__nv_bfloat16*>(base + sw128_off(n, k)) = __float2bfloat16_rn(v);
}

__global__ void pack_meta_kernel(ParamPtrs P, uint8_t* blob, Layout L, int out_ch) {
  float* M = reinterpret_cast<float*>(blob + L.off_meta);
  float s = 0.f;
  if (out_ch >= 4)
    for (int k = threadIdx.x; k < 256; k += 32) { const float w = P.p[S2L_P_OUT_W][3 * 256 + k]; s = fmaf(w, w, s); }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (threadIdx.x == 0) {
    const float nrm = sqrtf(s);
    M[2] = nrm;
    M[3] = 2e-3f * fmaxf(1.f, nrm / 1.41421356f);
  }
}

}  // namespace s2l

using namespace s2l;

// PositionalEncodingTime.div_term (tf_nerf.py:431-432): exp(arange(0,20,2) * -(ln(1e4)/20)) in fp32.
extern "C" void s2l_time_div_term(float* out10) {
  const float c = (float)(-(std::log(10000.0) / 20.0));
  for (int i = 0; i < 10; ++i) out10[i] = std::exp((float)(2 * i) * c);
}

extern "C" int32_t s2l_blob_meta(const void* blob, float* max_abs_weight, int32_t* n_saturating_weights, float* density_row_norm,
                                 float* auto_fix_thr, void* stream) {
  if (!blob) { set_error("s2l_blob_meta: null blob"); return 1; }
  float m[4];
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (cudaMemcpyAsync(m, reinterpret_cast<const uint8_t*>(blob) + blob_layout().off_meta, sizeof(m), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
      cudaStreamSynchronize(st) != cudaSuccess) {
    set_error("s2l_blob_meta: copy failed: %s", cudaGetErrorString(cudaGetLastError()));
    return 5;
  }
  if (max_abs_weight) *max_abs_weight = m[0];
  if (n_saturating_weights) { int32_t n; memcpy(&n, &m[1], 4); *n_saturating_weights = n; }
  if (density_row_norm) *density_row_norm = m[2];
  if (auto_fix_thr) *auto_fix_thr = m[3];
  return 0;
}

extern "C" size_t s2l_blob_bytes(int32_t uv_dims, int32_t out_ch) {
  (void)uv_dims;
  (void)out_ch;
  return blob_layout().total;
}

extern "C" int32_t s2l_pack_weights(const float* const* params_host, void* blob, int32_t uv_dims, int32_t out_ch,
                                    void* stream) {
  if (!params_host || !blob) { set_error("s2l_pack_weights: null argument"); return 1; }
  if ((uv_dims != 2 && uv_dims != 3) || out_ch < 1 || out_ch > 4) {
    set_error("s2l_pack_weights: unsupported dims uv_dims=%d out_ch=%d (need uv_dims in {2,3}, out_ch in 1..4)", uv_dims, out_ch);
    return 2;
  }
  ParamPtrs P;
  for (int i = 0; i < S2L_NUM_PARAMS; ++i) {
    if (!params_host[i]) { set_error("s2l_pack_weights: parameter %d is null", i); return 3; }
    P.p[i] = params_host[i];
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const Layout L = blob_layout();
  const int E = pe_dim(uv_dims);
  DivTerm dt;
  s2l_time_div_term(dt.v);
  for (int i = 10; i < 16; ++i) dt.v[i] = 0.f;
  uint8_t* b = reinterpret_cast<uint8_t*>(blob);
  pack_relayout_kernel<<<148, 256, 0, st>>>(P, b, L, E, out_ch, dt);
  if (!check_launch("pack_relayout_kernel")) return 5;
  pack_fold_kernel<<<(2 * 256 * 64 + 255) / 256, 256, 0, st>>>(P, b, L, E);
  if (!check_launch("pack_fold_kernel")) return 5;
  const long long n_elem = (long long)kTcwBytes / 4;   // one thread per (hi,lo) element pair
  pack_tcw_kernel<<<(unsigned)((n_elem + 255) / 256), 256, 0, st>>>(P, b, L, out_ch);
  if (!check_launch("pack_tcw_kernel")) return 5;
  const long long n_t = (long long)(2 + kTLayers * 8) * 128 * 64;
  pack_tcwt_kernel<<<(unsigned)((n_t + 255) / 256), 256, 0, st>>>(P, b, L, out_ch);
  if (!check_launch("pack_tcwt_kernel")) return 5;
  pack_meta_kernel<<<1, 32, 0, st>>>(P, b, L, out_ch);
  if (!check_launch("pack_meta_kernel")) return 5;
  return 0;
}
