// Training path on tensor cores (bf16 operands, fp32 accumulate): shared declarations of
//   s2l_mlp_tc.cu        mlp_tc_kernel<1,2,1,TRAIN>  forward, saves h0..h7 + positional encodings, fused 4-tap blend
//   s2l_train_dgrad.cu   dgrad_tc_kernel             d rgb -> dPre7..dPre0 (data gradients through the folded layer program)
//   s2l_train_wgrad.cu   wgrad_tc_kernel             split-K weight-gradient GEMMs  dW = dPre^T h  (MN-major operands)
//   s2l_train_final.cu   finalize kernels            slab reduction + the chain rule through the folded input layers
// Replaces autograd through Trainer.predict_lip_image -> TalkingFace.rgb_forward (training.py:158-251, 559; tf_nerf.py:225-285)
// for F frames x 4 taps in ONE differentiable launch sequence.
//
// Layer program and its gradients (pe = positional encoding, c_f / cs_f = the per-frame constants fc_audio(lat)+fc_time(t)+biases):
//   pre0 = (W0 Wuv) pe + W0 c_f + b0          h0 = relu(pre0)            [fold0 = W0 Wuv]
//   pre_g = W_g h_{g-1} + b_g                 h_g = relu(pre_g)          g = 1..4, 6, 7
//   pre5 = (W5a Wuvs) pe + W5a cs_f + W5b h4 + b5                        [fold5 = W5a Wuvs]
//   out = Wout h7 + bout;   rgb[pix] = sum_t w_t out[pix, t]
// dgrad:  dOut = w_t d rgb;  dPre7 = (dOut Wout) * [h7 > 0];  dPre_{g-1} = (dPre_g W_g) * [h_{g-1} > 0]  (g = 7,6,5b,4,3,2,1)
// wgrad:  dW_g = dPre_g^T h_{g-1};  M0 = dPre0^T pe;  M5 = dPre5^T pe;  dWout = dOut^T h7;  column sums of every dPre_g
//         (per frame for g = 0, 5);  the finalize step turns M0 / M5 / the per-frame sums into the gradients of
//         W0, Wuv, W5a, Wuvs, fc_audio*, fc_time*, their biases and the latent.
#pragma once
#include <cuda.h>          // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)
#include <cuda_bf16.h>
#include "s2l_common.cuh"
#include "s2l_points.cuh"

namespace s2l {

// 2-D bf16 tensor map over a row-major [rows][cols] array, box = [box_rows][box_cols]
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline bool encode_2d(CUtensorMap* tm, void* base, unsigned long long cols, unsigned long long rows, unsigned box_cols, unsigned box_rows,
                      CUtensorMapSwizzle sw) {
  static EncodeTiledFn enc = nullptr;
  if (!enc) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
      set_error("training kernels: cuTensorMapEncodeTiled is not available from this driver");
      return false;
    }
    enc = reinterpret_cast<EncodeTiledFn>(fn);
  }
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {cols * 2};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("training kernels: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return false;
  }
  return true;
}

// Buffers of one training render (device pointers, caller-owned); rows are tile-major: row = tile * 128 + r.
struct TrainBufs {
  __nv_bfloat16* h;        // [8][rows_total][256]   saved activations h0..h7 (forward)
  __nv_bfloat16* pe;       // [rows_total][64]       positional encodings (forward)
  const uint32_t* mask;    // [8][tiles][8][128]     ReLU masks of h0..h7, one word per (row, 32-column slice) (forward)
  __nv_bfloat16* dpre;     // [8][rows_total][256]   dPre0..dPre7 (dgrad)
  __nv_bfloat16* dout16;   // [rows_total][16]       dOut padded to 16 channels (dgrad prologue)
  float* partials;         // wgrad slab partials (WgPlan::total_floats)
  float* dbout_part;       // [kDboutRows][4] per-warp partial sums of dOut (output_linear bias gradient), zero-filled rows unused
  long long rows_total;
};

constexpr int kDboutRows = 148 * 4 * 2;      // one row per dOut-producer warp of every CTA (room for 296 CTAs)

// ---- wgrad work decomposition (shared by the kernel and the finalize step)
//   jobs 0..6  L: dW of pts_linears {1,2,3,4,5b,6,7}:  A = dPre_l, B = h_{l-1} (l = 5: h4), N = 256, SL slabs over all rows
//   job  7/8   M: M0 / M5:                             A = dPre0 / dPre5, B = pe, N = 64, per frame x SM slabs (column sums per frame)
//   job  9     O: dWout^T:                             A = h7, B = dOut16, N = 16, SO slabs
struct WgPlan {
  int F;                   // frames
  int chunks_per_frame;    // 64-row K chunks per frame (tiles_per_frame * 2)
  int SL, SM, SO;          // slabs per job class
  __host__ __device__ int n_items() const { return 7 * SL + 2 * F * SM + SO; }
  // float offsets of an item's partial block: L items [256][256] + [256] column sums, M items [256][64] + [256], O items [256][16]
  __host__ __device__ long long off_L(int job, int slab) const { return ((long long)job * SL + slab) * (65536 + 256); }
  __host__ __device__ long long off_M(int which, int f, int slab) const {
    return (long long)7 * SL * (65536 + 256) + (((long long)which * F + f) * SM + slab) * (16384 + 256);
  }
  __host__ __device__ long long off_O(int slab) const {
    return (long long)7 * SL * (65536 + 256) + (long long)2 * F * SM * (16384 + 256) + (long long)slab * 4096;
  }
  __host__ __device__ long long total_floats() const { return off_O(SO); }
};
inline WgPlan make_wg_plan(int F, long long tiles_per_frame, int sms) {
  WgPlan p;
  p.F = F;
  p.chunks_per_frame = (int)(tiles_per_frame * 2);
  const long long chunks = (long long)F * p.chunks_per_frame;
  // HBM-bound: an item's time ~ the bytes it loads.  A chunk (64 points) of an L job is 64 KB, of an M job 40 KB, of the O
  // job 34 KB; slabs are sized so that every item carries about total / #SMs bytes.
  const double bl = 64.0 * chunks, bm = 40.0 * chunks, bo = 34.0 * chunks;
  const double per_cta = (7 * bl + 2 * bm + bo) / (sms > 0 ? sms : 1);
  auto clampi = [](long long v, long long lo, long long hi) { return (int)(v < lo ? lo : (v > hi ? hi : v)); };
  p.SL = clampi((long long)(bl / (per_cta > 0 ? per_cta : 1) + 0.5), 1, chunks > 0 ? chunks : 1);
  p.SM = clampi((long long)(bm / (F > 0 ? F : 1) / (per_cta > 0 ? per_cta : 1) + 0.5), 1, p.chunks_per_frame > 0 ? p.chunks_per_frame : 1);
  p.SO = clampi((long long)(bo / (per_cta > 0 ? per_cta : 1) + 0.999), 1, chunks > 0 ? (chunks < 32 ? chunks : 32) : 1);
  return p;
}

// Item -> CTA assignment: items are enumerated largest first (L, then O, then M) and dealt to the persistent CTAs in
// boustrophedon order (round k forwards for even k, backwards for odd k), so a CTA that drew a large item in one round
// draws a small one in the next.  Returns the k-th item of CTA `cta`, or -1.
__host__ __device__ inline int wg_nth_item(int cta, int k, int grid, int n_items) {
  const int i = k * grid + ((k & 1) ? grid - 1 - cta : cta);
  return i < n_items ? i : -1;
}

}  // namespace s2l
