// Backward data-gradient chain of the fused MLP on tensor cores (tcgen05 + TMEM, bf16 operands, fp32 accumulate).
//
// Per 128-point tile (the forward's tiles, same persistent one-CTA-per-SM walk):
//   dOut [128,16]   = w_tap * d rgb[pixel]                       (4-tap blend backward, training.py:237-249; prologue warps)
//   dH7             = dOut * Wout                                 T8: one K = 16 MMA per accumulator half, SS form
//   dPre7           = dH7 * [h7 > 0]        -> TMEM (next A operand) + HBM (for the weight-gradient kernel)
//   dH_{l-1}        = dPre_l * W_l,  dPre_{l-1} = dH_{l-1} * [h_{l-1} > 0]      l = 7, 6, 5 (W5[:, 256:]), 4, 3, 2, 1
// i.e. autograd through tf_nerf.py:265-283 in the folded form (the chain stops at dPre0 / dPre5: the gradients of the
// folded input layers only need M = dPre^T pe and per-frame column sums, s2l_train.cuh).
//
// Same machinery as the forward (s2l_mlp_tc.cu): gradients live in TMEM only — two 256-column regions swap roles every
// layer, the epilogue converts the fp32 accumulator to the next layer's bf16 A operand IN PLACE per 64-column quarter and
// releases it to the MMA warp; W^T granules (blob section TCWT, pre-swizzled [128 N x 64 K] bf16 images) stream L2 -> SMEM
// through an 8-stage ring of 16 KB stages with one 1-D bulk copy each; the MMA issuer runs converged with a static issue
// program (every layer is 8 granules, so the ring position at each program point is a compile-time constant); the
// output layer's W^T (32 KB) stays resident in shared memory for the whole kernel.
// The ReLU masks come from the forward's saved activations: each epilogue thread prefetches the 64 B of h it needs
// before it waits for the accumulator, so the HBM latency hides behind the MMAs.
#include <type_traits>
#include "s2l_tc_common.cuh"
#include "s2l_train.cuh"

namespace s2l {

constexpr int DG_NSTG = 8;
constexpr int DG_STAGE = kTGran;                       // 16 KB
constexpr int DG_SM_DOUT = 0;                          // 2 x [128 rows][64 K] bf16 SW128 images, only K < 16 used
constexpr int DG_SM_WOUT = DG_SM_DOUT + 2 * PE_PLANE;  // output_linear^T, 2 granules, resident
constexpr int DG_SM_STG = DG_SM_WOUT + 2 * kTGran;
constexpr int DG_SM_BAR = DG_SM_STG + DG_NSTG * DG_STAGE;
// barriers: b_full[8] b_empty[8] wout_full do_full[2] do_empty[2] acc_full[2] epi_done[4] acc_t8[2]
constexpr int DG_NBAR = 2 * DG_NSTG + 1 + 2 + 2 + 2 + 4 + 2;
constexpr int DG_SM_TMEMPTR = DG_SM_BAR + DG_NBAR * 8;
constexpr int DG_SMEM_BYTES = DG_SM_TMEMPTR + 16;
static_assert(DG_SMEM_BYTES <= 232448, "shared memory budget");

struct DgArgs {
  const uint8_t* blob;
  Layout L;
  PointSrc src;               // GRID_ENS4
  int n_frames;
  long long tiles_per_frame;
  const float* d_rgb;         // [F, H*W, 3]   (4-tap render) ...
  const float* d_out_rows;    // ... or [N,3] gradient of the raw outputs (per-call rows contract, explicit points)
  TrainBufs B;
};

// fp32 accumulator slice (32 columns) + the ReLU mask word of the matching 32 saved activations -> 16 packed bf16 pairs of
// dPre = dH * [h > 0]
__device__ __forceinline__ void mask_slice(const uint32_t (&v)[32], uint32_t m, uint32_t (&o)[16]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float x0 = (m & (1u << (2 * j))) ? __uint_as_float(v[2 * j]) : 0.f;
    const float x1 = (m & (2u << (2 * j))) ? __uint_as_float(v[2 * j + 1]) : 0.f;
    o[j] = cvt_bf16x2(x0, x1);
  }
}

__global__ void __launch_bounds__(TC_THREADS, 1) dgrad_tc_kernel(const __grid_constant__ DgArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DG_SM_BAR);
  uint64_t* b_full = bars;
  uint64_t* b_empty = bars + DG_NSTG;
  uint64_t* wout_full = bars + 2 * DG_NSTG;
  uint64_t* do_full = wout_full + 1;
  uint64_t* do_empty = do_full + 2;
  uint64_t* acc_full = do_empty + 2;
  uint64_t* epi_done = acc_full + 2;
  // T8 has its own accumulator-ready barriers: it depends on nothing the epilogue produces, so the next tile's T8 (two
  // MMAs) can complete while the epilogue is still inside the previous tile's last step — on a shared barrier two
  // completions in a row would alias the 1-bit phase the epilogue is waiting for
  uint64_t* acc_t8 = epi_done + 4;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem + DG_SM_TMEMPTR);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const long long n_tiles = a.tiles_per_frame * a.n_frames;
  const uint8_t* tcwt = a.blob + a.L.off_tcwt;
  const long long RT = a.B.rows_total;

  if (tid == 0) {
    for (int s = 0; s < DG_NSTG; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    mbar_init(wout_full, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&do_full[b], 128);
      mbar_init(&do_empty[b], 1);
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_t8[b], 1);
    }
    for (int q = 0; q < 4; ++q) mbar_init(&epi_done[q], 256);
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_s, 0);

  if (warp == 0) {
    // =============================================================== weight producer
    if (elect_one()) {
      mbar_arrive_expect_tx(wout_full, 2 * kTGran);
      bulk_g2s(smem + DG_SM_WOUT, tcwt, 2 * kTGran, wout_full);
      uint32_t stage = 0, phase = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
#pragma unroll 1
        for (int gi = 0; gi < kTLayers * 8; ++gi) {
          mbar_wait_wd<true>(&b_empty[stage], phase ^ 1u, 100 + stage);
          mbar_arrive_expect_tx(&b_full[stage], kTGran);
          bulk_g2s(smem + DG_SM_STG + stage * DG_STAGE, tcwt + 2 * kTGran + (size_t)gi * kTGran, kTGran, &b_full[stage]);
          stage = (stage + 1) & (DG_NSTG - 1);
          phase ^= (stage == 0);
        }
      }
    }
  } else if (warp == 1) {
    // =============================================================== MMA issuer (converged warp, static program)
    uint32_t ph = 0, epi_par = 0;
    int rp = 0;
    long long it = 0;
    constexpr uint32_t kDescHi = 0x40004040u;   // SBO=64 | version=1 | SWIZZLE_128B
    auto mk = [](uint32_t lo) -> uint64_t { return ((uint64_t)kDescHi << 32) | lo; };
    const uint32_t stg0 = ((smem_u32(smem + DG_SM_STG) >> 4) & 0x3FFFu) | 0x10000u;
    const uint32_t wout0 = ((smem_u32(smem + DG_SM_WOUT) >> 4) & 0x3FFFu) | 0x10000u;
    uint32_t d_region = 0, a_region = 0;
    constexpr uint32_t idesc = idesc_bf16(kGranRows);
    auto granule = [&](auto stage_c, uint32_t d_addr, uint32_t a_t, uint32_t acc0, uint64_t* done0) {
      constexpr int STAGE = decltype(stage_c)::value;
      mbar_wait_trap(&b_full[STAGE], (ph >> STAGE) & 1u);
      ph ^= 1u << STAGE;
      tc_fence_after();
      const uint32_t b = stg0 + (uint32_t)STAGE * (uint32_t)(DG_STAGE >> 4);
      if (elect_one()) {
        // A chunk in TMEM (64 K = 32 columns of packed bf16 pairs, written by the epilogue as two 16-column slices at +0 / +32)
#pragma unroll
        for (int s = 0; s < 4; ++s)
          umma_ts(d_addr, a_t + (uint32_t)((s >> 1) * 32 + (s & 1) * 8), mk(b + 2 * s), idesc, s == 0 ? acc0 : 1u);
        umma_commit(&b_empty[STAGE]);
        if (done0) umma_commit(done0);
      }
      __syncwarp();
    };
    auto wait_quarter = [&](int hk) {
      mbar_wait_trap(&epi_done[hk], (epi_par >> hk) & 1u);
      epi_par ^= 1u << hk;
    };
    using std::integral_constant;
#define S2L_IC(v) integral_constant<int, (v)>{}
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int buf = (int)(it & 1);
      auto set_regions = [&]() {
        d_region = tmem_base + (rp ? 256u : 0u);
        a_region = tmem_base + (rp ? 0u : 256u);
        rp ^= 1;
      };
      // ---- T8: dH7 = dOut * Wout (K = 16), A = this tile's dOut image, B = resident output_linear^T
      set_regions();
      if (it == 0) mbar_wait_trap(wout_full, 0u);
      mbar_wait_trap(&do_full[buf], (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      {
        const uint32_t img = ((smem_u32(smem + DG_SM_DOUT + buf * PE_PLANE) >> 4) & 0x3FFFu) | 0x10000u;
        if (elect_one()) {
          umma_ss(d_region, mk(img), mk(wout0), idesc, 0u);
          umma_commit(&acc_t8[0]);
          umma_ss(d_region + 128u, mk(img), mk(wout0 + (uint32_t)(kTGran >> 4)), idesc, 0u);
          umma_commit(&acc_t8[1]);
          umma_commit(&do_empty[buf]);
        }
        __syncwarp();
      }
      // ---- 7 hidden layers, 8 granules each (ring position 0 at every layer start)
#pragma unroll 1
      for (int l = 0; l < kTLayers; ++l) {
        set_regions();
        wait_quarter(0); granule(S2L_IC(0), d_region, a_region, 0u, nullptr);
        wait_quarter(1); granule(S2L_IC(1), d_region, a_region + 64u, 1u, nullptr);
        wait_quarter(2); granule(S2L_IC(2), d_region, a_region + 128u, 1u, nullptr);
        wait_quarter(3); granule(S2L_IC(3), d_region, a_region + 192u, 1u, &acc_full[0]);
        granule(S2L_IC(4), d_region + 128u, a_region, 0u, nullptr);
        granule(S2L_IC(5), d_region + 128u, a_region + 64u, 1u, nullptr);
        granule(S2L_IC(6), d_region + 128u, a_region + 128u, 1u, nullptr);
        granule(S2L_IC(7), d_region + 128u, a_region + 192u, 1u, &acc_full[1]);
      }
    }
#undef S2L_IC
  } else if (warp >= 4 && warp < 8) {
    // =============================================================== dOut producers (one point = one tap per thread)
    const int r = tid - 128;
    long long it = 0;
    const long long npix = (long long)a.src.H * a.src.W;
    float sb0 = 0.f, sb1 = 0.f, sb2 = 0.f;            // running sums of dOut: the output_linear bias gradient
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int buf = (int)(it & 1);
      const int f = (int)(tile / a.tiles_per_frame);
      const long long p = (tile % a.tiles_per_frame) * TC_TM + r;
      float d[3] = {0.f, 0.f, 0.f};
      if (a.d_out_rows) {
        if (p < a.src.P) {
          const float* g = a.d_out_rows + ((long long)f * a.src.P + p) * 3;
          d[0] = g[0]; d[1] = g[1]; d[2] = g[2];
        }
      } else if (p < a.src.P) {
        // weight of this tap in the blend: area of the OPPOSITE tap / total area (training.py:237-249)
        const long long pix = p >> 2;
        const int tap = (int)(p & 3);
        float wt[4];
        ens4_weights(a.src, f, (unsigned)pix, wt);
        const float w = wt[tap];
        const float* g = a.d_rgb + ((long long)f * npix + pix) * 3;
        d[0] = w * g[0]; d[1] = w * g[1]; d[2] = w * g[2];
      }
      const uint4 lo = make_uint4(cvt_bf16x2(d[0], d[1]), cvt_bf16x2(d[2], 0.f), 0u, 0u);
      // the bias gradient sums the same bf16 values the weight-gradient GEMM multiplies
      sb0 += __uint_as_float(lo.x << 16); sb1 += __uint_as_float(lo.x & 0xffff0000u); sb2 += __uint_as_float(lo.y << 16);
      const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
      {
        const uint32_t w8[8] = {lo.x, lo.y, 0u, 0u, 0u, 0u, 0u, 0u};
        st_global_v8(a.B.dout16 + ((size_t)tile * TC_TM + r) * 16, w8);
      }
      mbar_wait_wd<true>(&do_empty[buf], (uint32_t)(((it >> 1) & 1) ^ 1), 500 + buf);
      uint8_t* img = smem + DG_SM_DOUT + buf * PE_PLANE + (r >> 3) * 1024 + (r & 7) * 128;
      *reinterpret_cast<uint4*>(img + ((0 ^ (r & 7)) << 4)) = lo;
      *reinterpret_cast<uint4*>(img + ((1 ^ (r & 7)) << 4)) = zero;
      fence_proxy_async();
      mbar_arrive(&do_full[buf]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sb0 += __shfl_xor_sync(0xffffffffu, sb0, o);
      sb1 += __shfl_xor_sync(0xffffffffu, sb1, o);
      sb2 += __shfl_xor_sync(0xffffffffu, sb2, o);
    }
    if (lane == 0) {
      float* dst = a.B.dbout_part + ((size_t)blockIdx.x * 4 + (warp - 4)) * 4;
      dst[0] = sb0; dst[1] = sb1; dst[2] = sb2; dst[3] = 0.f;
    }
  } else if (warp >= 8) {
    // =============================================================== epilogue: mask, convert, store
    const int quad = warp & 3, half = (warp - 8) >> 2;
    const uint32_t lane_sel = (uint32_t)(quad * 32) << 16;
    const int row = quad * 32 + lane;
    uint32_t acc_par[2] = {0, 0};
    int rp = 0;
    long long it = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const size_t grow = (size_t)tile * TC_TM + row;
#pragma unroll 1
      for (int step = 0; step < 8; ++step) {          // step s produces dPre_{7-s} from the accumulator of T8 / layer l = 8-s
        const int hl = 7 - step;
        const uint32_t d_region = tmem_base + (rp ? 256u : 0u);
        const uint32_t* mrow = a.B.mask + (((size_t)hl * n_tiles + tile) * 8) * TC_TM + row;
        __nv_bfloat16* drow = a.B.dpre + ((size_t)hl * RT + grow) * 256;
#pragma unroll 1
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t hm[2];
#pragma unroll
          for (int qq = 0; qq < 2; ++qq) hm[qq] = __ldg(mrow + (hh * 4 + qq * 2 + half) * TC_TM);
          if (step == 0) {
            mbar_wait_wd(&acc_t8[hh], (uint32_t)(it & 1), 710 + hh);
          } else {
            mbar_wait_wd(&acc_full[hh], acc_par[hh], 700 + hh);
            acc_par[hh] ^= 1;
          }
          tc_fence_after();
          const uint32_t taddr0 = d_region + lane_sel + (uint32_t)(hh * 128 + half * 32);
#pragma unroll
          for (int qq = 0; qq < 2; ++qq) {
            const int q = hh * 2 + qq;
            const uint32_t taddr = taddr0 + (uint32_t)(qq * 64);
            uint32_t v[32], o[16];
            tmem_ld32(taddr, v);
            tmem_ld_wait();
            mask_slice(v, hm[qq], o);
            if (step < 7) {                             // dPre0 feeds no further layer
              tmem_st16(taddr, o);
              tmem_st_wait();
              tc_fence_before();
              mbar_arrive(&epi_done[q]);                // released before the copy to global memory: the stores overlap the next MMAs
            }
#ifndef S2L_DBG_DG_NOSTORE     // experiment: the chain without its copies to global memory (results are garbage, timing only)
            st_rows64_paired(drow + q * 64 + half * 32, 512, o, lane);
#endif
          }
        }
        rp ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

int launch_dgrad_tc(const void* blob, const PointSrc& src, int n_frames, const float* d_rgb, const TrainBufs& B, cudaStream_t st,
                    const float* d_out_rows) {
  DgArgs a{};
  a.d_out_rows = d_out_rows;
  a.blob = reinterpret_cast<const uint8_t*>(blob);
  a.L = blob_layout();
  a.src = src;
  a.n_frames = n_frames;
  a.tiles_per_frame = (src.P + TC_TM - 1) / TC_TM;
  a.d_rgb = d_rgb;
  a.B = B;
  const long long n_tiles = a.tiles_per_frame * n_frames;
  if (n_tiles == 0) return 0;
  if (B.rows_total != n_tiles * TC_TM) { set_error("dgrad_tc: rows_total %lld != tiles * 128 = %lld", B.rows_total, n_tiles * TC_TM); return 2; }
  static bool attr_set_dev[64] = {};
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  bool& attr_set = attr_set_dev[cur_dev & 63];
  if (!attr_set) {
    if (cudaFuncSetAttribute(dgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DG_SMEM_BYTES) != cudaSuccess) {
      set_error("dgrad_tc: cannot opt in to %d B of shared memory: %s", DG_SMEM_BYTES, cudaGetErrorString(cudaGetLastError()));
      return 6;
    }
    attr_set = true;
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cur_dev);
  const unsigned grid = (unsigned)(n_tiles < sms ? n_tiles : sms);
  dgrad_tc_kernel<<<grid, TC_THREADS, DG_SMEM_BYTES, st>>>(a);
  return check_launch("dgrad_tc_kernel") ? 0 : 5;
}

}  // namespace s2l
