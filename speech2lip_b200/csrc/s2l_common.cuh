// speech2lip_b200 — shared definitions: packed-blob layout, PTX helpers, launch bookkeeping.
//
// Data layout in HBM (one blob per model, written by s2l_pack_weights, read-only afterwards,
// ~9 MB -> L2-resident during a render):
//   AUDIO   fp32  AudioNet parameters in their PyTorch layouts (tf_nerf.py:91-109)
//   CONST   fp32  per-frame-constant mat-vec weights (fc_audio/fc_time + *_skip), bias vectors,
//                 time div_term, folded input weights fold0 = W0*Wuv, fold5 = W5a*Wuv_skip
//   FP32    fp32  exact-path weights, transposed to [K][256] so a K-chunk is one contiguous 16 KB run
//   TCBIAS  fp32  per-layer bias rows for the tensor-core path
//   TCW     bf16  tensor-core B-operand "granules": [128 rows x 64 K] tiles stored as the exact
//                 shared-memory image the tcgen05 descriptor expects (K-major, 128-byte swizzle),
//                 hi plane then lo plane (bf16 split of the fp32 weight), in MMA issue order so the
//                 producer streams them with plain 1-D bulk copies (UBLKCP), no tensor map.
#pragma once
#include <cstdint>
#include <cstddef>
#include <cuda_runtime.h>
#include "../../include/speech2lip_b200.h"

namespace s2l {

constexpr int kHidden   = 256;
constexpr int kLatent   = 64;
constexpr int kTimePE   = 20;
constexpr int kAudioWin = 16;
constexpr int kAudioFeat = 29;
constexpr int kMultires = 10;
constexpr int kEPad     = 64;      // PE width padded (42 or 63 -> 64)
constexpr int kNumG     = 9;       // tensor-core GEMM layers after folding (G0..G8)
constexpr int kOutPad   = 16;      // output layer N padded to the minimum M=128 MMA N

// ---- AUDIO section (float offsets)
constexpr int A_CONV0_W = 0;
constexpr int A_CONV0_B = A_CONV0_W + 32 * 29 * 3;
constexpr int A_CONV1_W = A_CONV0_B + 32;
constexpr int A_CONV1_B = A_CONV1_W + 32 * 32 * 3;
constexpr int A_CONV2_W = A_CONV1_B + 32;
constexpr int A_CONV2_B = A_CONV2_W + 64 * 32 * 3;
constexpr int A_CONV3_W = A_CONV2_B + 64;
constexpr int A_CONV3_B = A_CONV3_W + 64 * 64 * 3;
constexpr int A_FC1_W   = A_CONV3_B + 64;
constexpr int A_FC1_B   = A_FC1_W + 64 * 64;
constexpr int A_FC2_W   = A_FC1_B + 64;
constexpr int A_FC2_B   = A_FC2_W + 64 * 64;
constexpr int A_TOTAL   = A_FC2_B + 64;          // 32800 floats

// ---- CONST section (float offsets)
constexpr int C_FCA_WT   = 0;                            // fc_audio^T       [64][256]
constexpr int C_FCAS_WT  = C_FCA_WT + 64 * 256;          // fc_audio_skip^T  [64][256]
constexpr int C_FCT_WT   = C_FCAS_WT + 64 * 256;         // fc_time^T        [20][256]
constexpr int C_FCTS_WT  = C_FCT_WT + 20 * 256;          // fc_time_skip^T   [20][256]
constexpr int C_BIAS6    = C_FCTS_WT + 20 * 256;         // b_uv,b_a,b_t,b_uvs,b_as,b_ts [6][256]
constexpr int C_DIV      = C_BIAS6 + 6 * 256;            // div_term[10] (+6 pad)
constexpr int C_FOLD0    = C_DIV + 16;                   // (W0 * Wuv)        [256][64]  (n-major)
constexpr int C_FOLD5    = C_FOLD0 + 256 * 64;           // (W5a * Wuv_skip)  [256][64]
constexpr int C_TOTAL    = C_FOLD5 + 256 * 64;

// ---- FP32 section (float offsets); every matrix is W^T = [K][256]
constexpr int F_UV_WT    = 0;                            // fc_uv^T       [64][256] (rows >= E are zero)
constexpr int F_UVS_WT   = F_UV_WT + 64 * 256;           // fc_uv_skip^T  [64][256]
constexpr int F_PTS_WT   = F_UVS_WT + 64 * 256;          // pts_linears.i^T, i=0..7; .5 has K=512
__host__ __device__ constexpr int f_pts_off(int i) { return F_PTS_WT + (i <= 5 ? i * 65536 : (i + 1) * 65536); }
constexpr int F_PTS_B    = F_PTS_WT + 9 * 65536;         // [8][256]
constexpr int F_OUT_W    = F_PTS_B + 8 * 256;            // output_linear.weight [4][256] (row >= out_ch zero)
constexpr int F_OUT_B    = F_OUT_W + 4 * 256;            // [4]
constexpr int F_TOTAL    = F_OUT_B + 4;

// ---- TCW section (byte offsets inside the section)
constexpr int kGranRows   = 128;                         // N rows per granule (one accumulator half)
constexpr int kGranPlane  = kGranRows * 128;             // 16384 B: [128 rows][64 K] bf16, SW128 K-major
constexpr int kGranBytes  = 2 * kGranPlane;              // hi + lo
constexpr int kOutPlane   = kOutPad * 128;               // 2048 B
constexpr int kOutGranBytes = 2 * kOutPlane;
__host__ __device__ constexpr int g_nkc(int g) { return g == 0 ? 1 : (g == 5 ? 5 : 4); }
__host__ __device__ constexpr int g_layer_bytes(int g) { return g == 8 ? 4 * kOutGranBytes : 2 * g_nkc(g) * kGranBytes; }
__host__ __device__ constexpr int g_layer_off(int g) {
  int o = 0;
  for (int i = 0; i < g; ++i) o += g_layer_bytes(i);
  return o;
}
constexpr int kTcwBytes = g_layer_off(kNumG);            // 1 982 464 B per tile pass
constexpr int kGranPerTile = 2 * (1 + 4 * 4 + 5 + 2 * 4) + 4;   // 64 granules = 128 planes

// ---- DGRAD section (float offsets): untransposed [out][in=256] weights for the backward data-gradient GEMMs
//      dX = dY * W.  Slots: pts_linears 0..4, pts_linears.5[:, :256], pts_linears.5[:, 256:], pts_linears 6, 7.
constexpr int D_SLOTS = 9;
__host__ __device__ constexpr int d_slot_off(int slot) { return slot * 65536; }
constexpr int D_TOTAL = D_SLOTS * 65536;

// ---- TCW8 section: same granule order/size as TCW, for the fp16 + fp8-correction arithmetic (S2L_PREC_FP16F8):
//      plane 0 = fp16(W) [rows][64 K] SW128;  plane 1 = e5m2(fp16(W) * 2^-kScaleA) [rows][64 K] SW64  followed by
//      e4m3((W - fp16(W)) * 2^kScaleW) [rows][64 K] SW64.
constexpr int kScaleA = 8;      // activation residual (A - fp16(A)) is scaled by 2^+8 before e4m3; W1 copy by 2^-8
constexpr int kScaleW = 10;     // weight residual (W - fp16(W)) is scaled by 2^+10 before e4m3; A1 copy by 2^-10

// ---- TCWT section: bf16 B operands of the tensor-core BACKWARD data-gradient GEMMs dH_{l-1} = dPre_l * W_l
//      (s2l_train_dgrad.cu): granules [128 N-rows = input channels x 64 K = output channels] of W^T, hi plane only
//      (16 KB, same K-major SW128 image as TCW), in issue order: output_linear^T (2 granules, K = out_ch padded to 64),
//      then pts_linears 7, 6, 5[:, 256:], 4, 3, 2, 1 (8 granules each: half h, K-chunk kc).
constexpr int kTGran = kGranPlane;                        // 16 KB
constexpr int kTLayers = 7;
constexpr int kTcwtBytes = (2 + kTLayers * 8) * kTGran;   // 950 272 B
__host__ __device__ constexpr int t_layer_src(int i) { return 7 - i; }            // i-th dgrad layer reads pts_linears[7 - i]
__host__ __device__ constexpr int t_layer_off(int i) { return (2 + i * 8) * kTGran; }

// ---- META section (16 x 4 bytes, written by the pack kernels): the validated domain of the fp16f8 arithmetic and the
//      model-dependent threshold of the last-sample re-evaluation
//   [0] float  max |w| over every tensor-core weight (folded input layers included)
//   [1] int    number of tensor-core weights with |w| >= kF8MaxWeight (their e4m3 residual saturates: correction lost)
//   [2] float  L2 norm of output_linear's density row (0 when out_ch < 4)
//   [3] float  automatic fix_thr = 2e-3 * max(1, [2] / sqrt(2)): the tensor-core density error scales with that row
constexpr int kMetaWords = 16;
constexpr float kF8MaxWeight = 1024.f;       // (w - fp16 w) * 2^kScaleW must stay below e4m3's 448: half an fp16 ulp at 1024 is 0.5
constexpr float kF8MaxAct = 4096.f;          // (a - fp16 a) * 2^kScaleA below 448: half an fp16 ulp at 4096 is 2 (> 1.75)

struct Layout {
  size_t off_audio, off_const, off_fp32, off_tcbias, off_tcw, off_dgrad, off_tcw8, off_tcwt, off_meta, total;
};
__host__ __device__ inline Layout blob_layout() {
  Layout L;
  size_t o = 0;
  auto al = [](size_t x) { return (x + 1023) & ~size_t(1023); };
  L.off_audio = o;  o = al(o + sizeof(float) * A_TOTAL);
  L.off_const = o;  o = al(o + sizeof(float) * C_TOTAL);
  L.off_fp32 = o;   o = al(o + sizeof(float) * F_TOTAL);
  L.off_tcbias = o; o = al(o + sizeof(float) * kNumG * 256);
  L.off_tcw = o;    o = al(o + kTcwBytes);
  L.off_dgrad = o;  o = al(o + sizeof(float) * D_TOTAL);
  L.off_tcw8 = o;   o = al(o + kTcwBytes);
  L.off_tcwt = o;   o = al(o + kTcwtBytes);
  L.off_meta = o;   o = al(o + kMetaWords * 4);
  L.total = o;
  return L;
}

__host__ __device__ inline int pe_dim(int uv_dims) { return uv_dims + 2 * kMultires * uv_dims; }

// byte offset of element (row n, k) inside a K-major 128B-swizzled [rows][64] bf16 tile
__host__ __device__ inline int sw128_off(int n, int k) {
  return (n >> 3) * 1024 + (n & 7) * 128 + ((((k >> 3) ^ (n & 7)) & 7) << 4) + (k & 7) * 2;
}

// byte offset of element (row n, k) inside a K-major 64B-swizzled [rows][64] 8-bit tile (8-row atoms of 512 B)
__host__ __device__ inline int sw64_off(int n, int k) {
  return (n >> 3) * 512 + (n & 7) * 64 + ((((k >> 4) ^ ((n >> 1) & 3)) & 3) << 4) + (k & 15);
}

// ------------------------------------------------------------------ error / launch bookkeeping (host)
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
bool check_launch(const char* what);

// ------------------------------------------------------------------ device helpers
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}
// try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes (or ~hint ns pass)
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 1-D bulk copy global -> shared (TMA engine, SASS UBLKCP), completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// torch.linspace(0,1,n)[i] as ATen computes it on CUDA and CPU (RangeFactories: symmetric halves, fp32 step):
//   step = (end-start)/(n-1);  i < n/2 ? start + step*i : end - step*(n-1-i)
// Both back ends contract the second form into ONE fused multiply-add (nvcc's default -fmad=true on CUDA, the
// vectorised fmadd on CPU), i.e. round(1 - step*k) with a single rounding — two roundings differ by 1 ulp in ~9 % of the
// coordinates at W = 256, which the 2^9 positional-encoding frequency turns into ~1e-4 at the output.
__device__ __forceinline__ float linspace01(int i, int n, float step) {      // step = 1.0f / (float)(n - 1), IEEE fp32 division
  if (n == 1) return 0.f;
  return (i < n / 2) ? __fmul_rn(step, (float)i) : __fmaf_rn(-step, (float)(n - 1 - i), 1.0f);
}
__device__ __forceinline__ float linspace01(int i, int n) {
  return linspace01(i, n, n > 1 ? __fdiv_rn(1.0f, (float)(n - 1)) : 0.f);
}
// density2outputs pieces (rendering.py:43-58), shared by composite_kernel, the fused reducer warp and the fp32
// re-evaluation kernel so that all three round identically
__device__ __forceinline__ float sigmoidf_acc(float x) { return __fdiv_rn(1.f, 1.f + expf(-x)); }
__device__ __forceinline__ float alpha_of(float sigma, float dist) {        // 1 - exp(-relu(sigma) * dist)
  return __fsub_rn(1.f, expf(-__fmul_rn(fmaxf(sigma, 0.f), dist)));
}
__device__ __forceinline__ float ray_norm(const float* d) {
  return sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
}
#endif  // __CUDACC__

}  // namespace s2l
