// Throughput path: the fused per-point MLP on 5th-gen tensor cores (tcgen05 + TMEM), sm_100a only.
//
// One persistent CTA per SM walks 128-point tiles.  Per tile it evaluates the folded layer program
//   G0  relu(fold0 * PE + bias0'[f])                 K = 64          (fc_uv folded into pts_linears.0)
//   G1-4 relu(W_g h + b_g)                            K = 256
//   G5  relu(fold5 * PE + W5b h + bias5'[f])          K = 64 + 256    (fc_uv_skip folded into pts_linears.5)
//   G6-7 relu(W_g h + b_g)                            K = 256
//   G8  W_out h + b_out                               K = 256, N = 16 (out_ch padded)
// which is algebraically TalkingFace.rgb_forward (tf_nerf.py:225-285) with the per-frame-constant
// audio/time terms hoisted into bias0'/bias5' (s2l_audio.cu).
//
// Where the data lives
//   activations : TMEM only.  The 512 columns are two 256-column regions R0/R1 that swap roles every
//                 layer: tcgen05.mma accumulates D_g (fp32, 128 lanes x 256 cols) into one region while
//                 reading A_g (bf16 hi/lo, packed 2 per column) from the other.  The epilogue converts
//                 D_g -> A_{g+1} IN PLACE, one 64-column quarter at a time (tcgen05.ld -> +bias, ReLU,
//                 bf16 hi/lo split -> tcgen05.st over the same columns), so the next layer's MMAs over
//                 K-chunk q start as soon as quarter q is converted while later quarters still compute.
//   weights     : streamed from L2 (blob TCW / TCW8 section, pre-swizzled smem images) into a 4-stage ring of
//                 32 KB granules (both planes of a [128 N-rows x 64 K] weight tile, contiguous in the blob ->
//                 ONE 1-D bulk copy, one barrier wait and one commit per 8-12 MMAs).  MMAs are M=128 x N=128
//                 (one accumulator half, 64 tensor-pipe cycles) so the epilogue of half 0 overlaps the MMAs of
//                 half 1.  The issuing warp runs converged on the uniform datapath with a STATIC program
//                 (see the MMA-issuer role below): the pipe's queue is only ~7 MMAs deep, so the issuer has to
//                 stay ahead of it.
//   PE          : computed by 4 dedicated warps one tile ahead, written as a K-major SW128 bf16 hi/lo
//                 A-operand image in shared memory (used by G0 and again by G5).
//
// Precision: NPASS = 3 issues hi*hi + lo*hi + hi*lo (bf16 split, fp32 accumulate) ~ 2^-17 relative
// per product — this is the parity mode (<= 1e-3 max-abs).  NPASS = 1 issues hi*hi only.
// NPASS = 2 ("fp16f8") is the cheaper parity mode: an exact fp16 x fp16 main product plus the two first-order
// corrections (A - fp16 A) * W and A * (W - fp16 W) evaluated on the fp8 tensor path (kind::f8f6f4, twice the
// bf16 rate) with power-of-two pre-scaling (e4m3 for the residuals, e5m2 for the full-range factors):
// 1 + 0.5 + 0.5 = 2 bf16-MMA equivalents per product instead of 3, ~2^-15 relative per product.
//
// Warp roles (512 threads): w0 weight producer (one lane), w1 MMA issuer (whole warp, one elected lane issues),
// w2 TMEM allocator, w3 idle,
// w4-7 PE producers, w8-15 epilogue (two warps per TMEM lane quadrant, 32 columns each).
#include <cuda.h>          // CUtensorMap (type only)
#include <cstdlib>
#include <type_traits>
#include "s2l_tc_common.cuh"
#include "s2l_train.cuh"

namespace s2l {

// shared-memory map of this kernel: PE images (2 x 32 KB) | 4-stage ring of 32 KB granules | biases | barriers
constexpr int T1_NSTG = 4;                          // power of two
constexpr int T1_STAGE = kGranBytes;                // 32 KB: both planes of a granule
constexpr int T1_SM_TCBIAS = SM_STG + T1_NSTG * T1_STAGE;
constexpr int T1_SM_FBIAS = T1_SM_TCBIAS + kNumG * 256 * 4;
constexpr int T1_SM_BAR = T1_SM_FBIAS + 2 * 2 * 256 * 4;
constexpr int T1_NBAR = 2 * T1_NSTG + 2 + 2 + 4 + 4 + 2 + 2 + 2;
constexpr int T1_SM_TMEMPTR = T1_SM_BAR + T1_NBAR * 8;
constexpr int T1_SM_RAW = T1_SM_TMEMPTR + 16;          // 2 x [128] float4: output tile handed to the reducer warp
constexpr int T1_SMEM_BYTES = T1_SM_RAW + 2 * TC_TM * 16;
static_assert(T1_SMEM_BYTES <= 232448, "shared memory budget");

// CL = 1: independent CTAs.  CL = 2: CTAs are launched as clusters of two that share ONE weight stream — each CTA
// fetches half of every granule and multicasts it into both CTAs' rings, halving the L2 reads / crossbar traffic per
// weight byte delivered (worth ~+3 % sustained); MMAs, TMEM and epilogues stay per CTA,
// the only coupling is the ring (a stage is refilled when BOTH CTAs released it).
// TRAIN (bf16 single pass only): additionally saves h0..h7 and the positional encodings for the backward kernels
// (s2l_train_dgrad.cu, s2l_train_wgrad.cu).
template <int NPASS, int UVD, int CL, bool TRAIN = false>
__global__ void __launch_bounds__(TC_THREADS, 1) mlp_tc_kernel(const __grid_constant__ TcArgs a, const __grid_constant__ CUtensorMap hmap) {
  // hmap (TRAIN only): [8 * rows_total][256] bf16 view of save_h, box 32 rows x 32 columns, SWIZZLE_64B — the epilogue warps
  // stage their slices in the ring stages' unused second planes (a single-pass kernel streams only the first) and the TMA
  // engine writes them out
  if (a.gate.flag && *a.gate.flag != a.gate.value) return;      // gated launch: the other implementation serves this call
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + T1_SM_BAR);
  uint64_t* b_full = bars;
  uint64_t* b_empty = bars + T1_NSTG;
  uint64_t* pe_full = bars + 2 * T1_NSTG;
  uint64_t* pe_empty = pe_full + 2;
  uint64_t* acc_full = pe_empty + 2;
  uint64_t* epi_done = acc_full + 4;
  uint64_t* raw_full = epi_done + 4;
  uint64_t* raw_empty = raw_full + 2;
  // The output layer has its own accumulator-ready barrier.  On acc_full[0] its completion would be followed by the NEXT
  // tile's G0 completion with nothing in between that depends on the epilogue; an epilogue warp whose barrier poll is
  // delayed past both (its polls queue behind global stores in the training forward) would see the 1-bit phase flip twice
  // and wait forever.
  uint64_t* acc_out = raw_empty + 2;
  float4* rawbuf = reinterpret_cast<float4*>(smem + T1_SM_RAW);
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem + T1_SM_TMEMPTR);
  float* tcbias_s = reinterpret_cast<float*>(smem + T1_SM_TCBIAS);
  float* fbias_s = reinterpret_cast<float*>(smem + T1_SM_FBIAS);   // [2 bufs][2][256]

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // provably warp-uniform: role branches and the MMA warp's loop state stay on the uniform datapath
  const long long n_tiles = launch_tiles(a);
  // every CTA runs the same number of tile iterations (a cluster shares the weight ring, so it must stay in lock step);
  // iterations past the end work on zero rows and store nothing
  const long long n_iter = (n_tiles + gridDim.x - 1) / gridDim.x;
  const long long tile_end = (long long)blockIdx.x + n_iter * gridDim.x;
  const uint32_t crank = (CL > 1) ? cluster_ctarank() : 0u;
  constexpr uint16_t kClMask = (uint16_t)((1u << CL) - 1u);
  const uint8_t* tcw = a.blob + (NPASS == 2 ? a.L.off_tcw8 : a.L.off_tcw);

#ifdef S2L_TIMELINE
  if (tid == 0 && blockIdx.x == 0 && a.dbg) {      // SM clock / wall clock pair at kernel start (effective frequency)
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    a.dbg[3 * 8192 + 0] = clock64();
    a.dbg[3 * 8192 + 1] = (long long)gt;
  }
#endif
  if (tid == 0) {
    for (int s = 0; s < T1_NSTG; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], CL);       // released by the MMA warp of every CTA sharing the stream
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&pe_full[b], 128);
      mbar_init(&pe_empty[b], 1 + 256);
      mbar_init(&raw_full[b], 128);
      mbar_init(&raw_empty[b], 1);
    }
    for (int q = 0; q < 4; ++q) {
      mbar_init(&acc_full[q], 1);
      mbar_init(&epi_done[q], 256);
    }
    mbar_init(acc_out, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {
    const float* TB = reinterpret_cast<const float*>(a.blob + a.L.off_tcbias);
    for (int i = tid; i < kNumG * 256; i += TC_THREADS) tcbias_s[i] = TB[i];
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();      // peer barriers initialised before any multicast copy / commit can land
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_s, 0);

  if (warp == 0) {
    // =============================================================== weight producer
    // One stage = one granule (both planes of a [128 N x 64 K] weight tile, contiguous in the blob) = ONE bulk copy.
    if (elect_one()) {
      uint32_t stage = 0, phase = 0;
      for (long long tile = blockIdx.x; tile < tile_end; tile += gridDim.x) {
#pragma unroll 1
        for (int g = 0; g < kNumG; ++g) {
          const int ngran = (g == 8) ? 4 : 2 * g_nkc(g);
          const uint32_t gran = (g == 8) ? kOutGranBytes : kGranBytes;
          const uint32_t bytes = (NPASS == 1) ? gran / 2 : gran;        // bf16x1 only needs the hi plane
          const uint8_t* src = tcw + g_layer_off(g);
#pragma unroll 1
          for (int gi = 0; gi < ngran; ++gi) {
#ifdef S2L_DBG_PRODSPIN
            mbar_wait_wd<false>(&b_empty[stage], phase ^ 1u, 100 + stage);
#else
            mbar_wait_wd<true>(&b_empty[stage], phase ^ 1u, 100 + stage);
#endif
#ifdef S2L_DBG_NOLOAD        // experiment: weights are never streamed (results are garbage, timing only)
            mbar_arrive(&b_full[stage]);
#else
#ifdef S2L_DBG_HALFLOAD      // experiment: only half of every granule is streamed (what a CTA pair would fetch per SM)
            mbar_arrive_expect_tx(&b_full[stage], bytes / 2);
            bulk_g2s(smem + SM_STG + stage * T1_STAGE, src + (size_t)gi * gran, bytes / 2, &b_full[stage]);
#else
            mbar_arrive_expect_tx(&b_full[stage], bytes);
            if (CL == 1) {
              bulk_g2s(smem + SM_STG + stage * T1_STAGE, src + (size_t)gi * gran, bytes, &b_full[stage]);
            } else {        // my slice of the granule, into every CTA of the cluster
              const uint32_t slice = bytes / CL;
              bulk_g2s_mc(smem + SM_STG + stage * T1_STAGE + crank * slice, src + (size_t)gi * gran + crank * slice, slice,
                          &b_full[stage], kClMask);
            }
#endif
#endif
            stage = (stage + 1) & (T1_NSTG - 1);
            phase ^= (stage == 0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================================================== MMA issuer
    // The whole warp walks the layer program in lock step, so every address / descriptor is computed on the uniform
    // datapath; one elected lane issues the tcgen05 instructions.  The program is STATIC: ring stage, operand offsets
    // and instruction descriptors of every MMA are compile-time constants relative to a few per-tile bases (the ring
    // has 4 stages and the layers use 2 / 8 / 10 / 8 / 4 granules, so the stage at each program point is
    // fixed: G0 starts at 0, G1-4 at 2, G5 at 2, G6-7 at 0, G8 at 0), which brings the issue cost per MMA well under
    // the 64 tensor-pipe cycles it covers — the issuing thread must run AHEAD of the pipe (queue depth ~7 MMAs).
    uint32_t ph = 0;                            // bit s = parity to wait for on b_full[s]
    uint32_t epi_par = 0;                       // bit hk = parity to wait for on epi_done[hk]
    int rp = 0;
    long long it = 0;
    TL_DECL;
    constexpr uint32_t kDescHi = 0x40004040u;   // SBO=64 | version=1 | SWIZZLE_128B  (upper descriptor word)
    auto mk = [](uint32_t lo) -> uint64_t { return ((uint64_t)kDescHi << 32) | lo; };
    constexpr uint32_t kDescHi64 = 0x80004020u; // SBO=32 (512 B atoms) | version=1 | SWIZZLE_64B: 8-bit operands, 64 K per row
    auto mk64 = [](uint32_t lo) -> uint64_t { return ((uint64_t)kDescHi64 << 32) | lo; };
    const uint32_t stg0 = ((smem_u32(smem + SM_STG) >> 4) & 0x3FFFu) | 0x10000u;
    uint32_t pe_hi = 0, d_region = 0, a_region = 0;
    int buf = 0;

    // One granule (ring stage STAGE): wait for its bytes, issue its 4 / 8 / 12 MMAs, release the stage.
    //   IS_PE: A operand = this tile's positional-encoding image in shared memory (SS form), else TMEM (TS form)
    //   SMALL: output layer (N = 16, 2 KB planes).  done0/done1: extra barriers committed after the last MMA.
    auto granule = [&](auto stage_c, auto is_pe_c, auto small_c, uint32_t d_addr, uint32_t a_t, uint32_t acc0,
                       uint64_t* done0, uint64_t* done1) {
      constexpr int STAGE = decltype(stage_c)::value;
      constexpr bool IS_PE = decltype(is_pe_c)::value;
      constexpr bool SMALL = decltype(small_c)::value;
      constexpr int n_mma = SMALL ? kOutPad : kGranRows;
      constexpr uint32_t idesc = (NPASS == 2) ? idesc_f16(n_mma) : idesc_bf16(n_mma);
      constexpr uint32_t idesc_rw = idesc_f8(n_mma, 0u, 1u);     // (A - fp16 A) [e4m3] x fp16(W) [e5m2]
      constexpr uint32_t idesc_wr = idesc_f8(n_mma, 1u, 0u);     // fp16(A) [e5m2] x (W - fp16 W) [e4m3]
      constexpr uint32_t plane16 = (uint32_t)((SMALL ? kOutPlane : kGranPlane) >> 4);
      TLC(3);
      mbar_wait_trap(&b_full[STAGE], (ph >> STAGE) & 1u);
      TLC(2);
      ph ^= 1u << STAGE;
      tc_fence_after();
      const uint32_t b = stg0 + (uint32_t)STAGE * (uint32_t)(T1_STAGE >> 4);      // first plane: hi / fp16
      const uint32_t b2 = b + plane16;                                            // second plane: lo / [e5m2 | e4m3]
      const uint32_t b4 = b2 + (plane16 >> 1);
      const uint32_t pe_lo = pe_hi + (PE_PLANE >> 4);
      const uint32_t pe_e5 = pe_lo, pe_e4 = pe_lo + (PE_PLANE >> 5);      // NPASS == 2: [fp16 16 KB | e5m2 8 KB | e4m3 8 KB]
      if (elect_one()) {
        if (IS_PE) {
#pragma unroll
          for (int s = 0; s < 4; ++s) {
            umma_ss(d_addr, mk(pe_hi + 2 * s), mk(b + 2 * s), idesc, s == 0 ? acc0 : 1u);
            if (NPASS == 3) umma_ss(d_addr, mk(pe_lo + 2 * s), mk(b + 2 * s), idesc, 1u);
          }
          if (NPASS == 3) {
#pragma unroll
            for (int s = 0; s < 4; ++s) umma_ss(d_addr, mk(pe_hi + 2 * s), mk(b2 + 2 * s), idesc, 1u);
          }
          if (NPASS == 2) {
#pragma unroll
            for (int t = 0; t < 2; ++t) umma8_ss(d_addr, mk64(pe_e4 + 2 * t), mk64(b2 + 2 * t), idesc_rw, 1u);
#pragma unroll
            for (int t = 0; t < 2; ++t) umma8_ss(d_addr, mk64(pe_e5 + 2 * t), mk64(b4 + 2 * t), idesc_wr, 1u);
          }
        } else {
          // A chunk layout in TMEM (64 cols): bf16: [hi K0-31 (16) | lo K0-31 (16) | hi K32-63 (16) | lo K32-63 (16)]
          //                                   fp16f8, per 32-K half: [fp16 (16 cols) | e5m2 (8) | e4m3 (8)]
#pragma unroll
          for (int s = 0; s < 4; ++s) {
            const uint32_t a_hi = a_t + (uint32_t)((s >> 1) * 32 + (s & 1) * 8);
            umma_ts(d_addr, a_hi, mk(b + 2 * s), idesc, s == 0 ? acc0 : 1u);
            if (NPASS == 3) umma_ts(d_addr, a_hi + 16u, mk(b + 2 * s), idesc, 1u);
          }
          if (NPASS == 3) {
#pragma unroll
            for (int s = 0; s < 4; ++s)
              umma_ts(d_addr, a_t + (uint32_t)((s >> 1) * 32 + (s & 1) * 8), mk(b2 + 2 * s), idesc, 1u);
          }
          if (NPASS == 2) {
            // same-format MMAs are issued back to back (operand formats live in the instruction descriptor)
#pragma unroll
            for (int t = 0; t < 2; ++t) umma8_ts(d_addr, a_t + (uint32_t)(t * 32 + 24), mk64(b2 + 2 * t), idesc_rw, 1u);
#pragma unroll
            for (int t = 0; t < 2; ++t) umma8_ts(d_addr, a_t + (uint32_t)(t * 32 + 16), mk64(b4 + 2 * t), idesc_wr, 1u);
          }
        }
        if (CL == 1) umma_commit(&b_empty[STAGE]);      // stage reusable once these MMAs retire
        else umma_commit_mc(&b_empty[STAGE], kClMask);  // ... in every CTA that multicasts into it
        if (done0) umma_commit(done0);
        if (done1) umma_commit(done1);
      }
      __syncwarp();
    };
    auto wait_quarter = [&](int hk) {       // quarter hk of the previous layer's output converted to operand form
      TLC(0);
      mbar_wait_trap(&epi_done[hk], (epi_par >> hk) & 1u);
      TLC(1);
      epi_par ^= 1u << hk;
    };
    using std::integral_constant;
#define S2L_IC(v) integral_constant<int, (v)>{}
#define S2L_BC(v) integral_constant<bool, (v)>{}
    // hidden layer (4 K-chunks per accumulator half), ring position START at entry
    auto layer_std = [&](auto start_c, int g) {
      constexpr int START = decltype(start_c)::value;
      (void)g;
      TL(0, 1000 + g * 10);
      wait_quarter(0); granule(S2L_IC((START + 0) & 3), S2L_BC(false), S2L_BC(false), d_region, a_region, 0u, nullptr, nullptr);
      wait_quarter(1); granule(S2L_IC((START + 1) & 3), S2L_BC(false), S2L_BC(false), d_region, a_region + 64u, 1u, nullptr, nullptr);
      wait_quarter(2); granule(S2L_IC((START + 2) & 3), S2L_BC(false), S2L_BC(false), d_region, a_region + 128u, 1u, nullptr, nullptr);
      wait_quarter(3); granule(S2L_IC((START + 3) & 3), S2L_BC(false), S2L_BC(false), d_region, a_region + 192u, 1u, &acc_full[0], nullptr);
      TL(0, 5000 + g * 10); TLC_FLUSH(0); TL(0, 1001 + g * 10);
      granule(S2L_IC((START + 0) & 3), S2L_BC(false), S2L_BC(false), d_region + 128u, a_region, 0u, nullptr, nullptr);
      granule(S2L_IC((START + 1) & 3), S2L_BC(false), S2L_BC(false), d_region + 128u, a_region + 64u, 1u, nullptr, nullptr);
      granule(S2L_IC((START + 2) & 3), S2L_BC(false), S2L_BC(false), d_region + 128u, a_region + 128u, 1u, nullptr, nullptr);
      granule(S2L_IC((START + 3) & 3), S2L_BC(false), S2L_BC(false), d_region + 128u, a_region + 192u, 1u, &acc_full[1], nullptr);
      TL(0, 5001 + g * 10); TLC_FLUSH(0);
    };
    for (long long tile = blockIdx.x; tile < tile_end; tile += gridDim.x, ++it) {
      buf = (int)(it & 1);
      pe_hi = ((smem_u32(smem + SM_PE + buf * PE_BUF) >> 4) & 0x3FFFu) | 0x10000u;
      auto set_regions = [&]() {
        d_region = tmem_base + (rp ? 256u : 0u);
        a_region = tmem_base + (rp ? 0u : 256u);
        rp ^= 1;
      };
      // ---- G0: PE x fold0 (stages 0, 1)
      set_regions();
      mbar_wait_trap(&pe_full[buf], (uint32_t)((it >> 1) & 1));
      TL(0, 1000);
      granule(S2L_IC(0), S2L_BC(true), S2L_BC(false), d_region, 0u, 0u, &acc_full[0], nullptr);
      TL(0, 5000); TLC_FLUSH(0); TL(0, 1001);
      granule(S2L_IC(1), S2L_BC(true), S2L_BC(false), d_region + 128u, 0u, 0u, &acc_full[1], nullptr);
      TL(0, 5001); TLC_FLUSH(0);
      // ---- G1-4 (ring position 2)
#pragma unroll 1
      for (int g = 1; g <= 4; ++g) {
        set_regions();
        layer_std(S2L_IC(2), g);
      }
      // ---- G5: PE x fold5 + W5b x h (5 granules per half; ring position 2, then 3)
      set_regions();
      TL(0, 1050);
      granule(S2L_IC(2), S2L_BC(true), S2L_BC(false), d_region, 0u, 0u, nullptr, nullptr);
      wait_quarter(0); granule(S2L_IC(3), S2L_BC(false), S2L_BC(false), d_region, a_region, 1u, nullptr, nullptr);
      wait_quarter(1); granule(S2L_IC(0), S2L_BC(false), S2L_BC(false), d_region, a_region + 64u, 1u, nullptr, nullptr);
      wait_quarter(2); granule(S2L_IC(1), S2L_BC(false), S2L_BC(false), d_region, a_region + 128u, 1u, nullptr, nullptr);
      wait_quarter(3); granule(S2L_IC(2), S2L_BC(false), S2L_BC(false), d_region, a_region + 192u, 1u, &acc_full[0], nullptr);
      TL(0, 5050); TLC_FLUSH(0); TL(0, 1051);
      granule(S2L_IC(3), S2L_BC(true), S2L_BC(false), d_region + 128u, 0u, 0u, nullptr, nullptr);
      granule(S2L_IC(0), S2L_BC(false), S2L_BC(false), d_region + 128u, a_region, 1u, nullptr, nullptr);
      granule(S2L_IC(1), S2L_BC(false), S2L_BC(false), d_region + 128u, a_region + 64u, 1u, nullptr, nullptr);
      granule(S2L_IC(2), S2L_BC(false), S2L_BC(false), d_region + 128u, a_region + 128u, 1u, nullptr, nullptr);
      granule(S2L_IC(3), S2L_BC(false), S2L_BC(false), d_region + 128u, a_region + 192u, 1u, &acc_full[1], &pe_empty[buf]);   // last reader of this tile's PE image
      TL(0, 5051); TLC_FLUSH(0);
      // ---- G6-7 (ring position 0)
#pragma unroll 1
      for (int g = 6; g <= 7; ++g) {
        set_regions();
        layer_std(S2L_IC(0), g);
      }
      // ---- G8: output layer, N = 16 (ring position 0)
      set_regions();
      TL(0, 1080);
      wait_quarter(0); granule(S2L_IC(0), S2L_BC(false), S2L_BC(true), d_region, a_region, 0u, nullptr, nullptr);
      wait_quarter(1); granule(S2L_IC(1), S2L_BC(false), S2L_BC(true), d_region, a_region + 64u, 1u, nullptr, nullptr);
      wait_quarter(2); granule(S2L_IC(2), S2L_BC(false), S2L_BC(true), d_region, a_region + 128u, 1u, nullptr, nullptr);
      wait_quarter(3); granule(S2L_IC(3), S2L_BC(false), S2L_BC(true), d_region, a_region + 192u, 1u, acc_out, nullptr);
      TL(0, 5080); TLC_FLUSH(0);
    }
#undef S2L_IC
#undef S2L_BC
  } else if (warp == 3) {
    // =============================================================== reducer (fused 4-tap blend / alpha compositing)
    if (a.epi_mode != EPI_RAW) reducer_role(a, n_tiles, tile_end, rawbuf, raw_full, raw_empty, lane);
  } else if (warp >= 4 && warp < 8) {
    // =============================================================== PE producers (one point per thread)
    const int r = tid - 128;
    long long it = 0;
    int fcur = 0;
    for (long long tile = blockIdx.x; tile < tile_end; tile += gridDim.x, ++it) {
      const int buf = (int)(it & 1);
      int f; long long p0, Pf;
      tile_locate<TC_TM>(a.src, a.tiles_per_frame, n_tiles, tile, fcur, f, p0, Pf);
      const long long p = p0 + r;
      mbar_wait_wd<true>(&pe_empty[buf], (uint32_t)(((it >> 1) & 1) ^ 1), 500 + buf);
      pe_write_row<NPASS, UVD>(a.src, f, p, p < Pf, r, smem + SM_PE + buf * PE_BUF,
#ifndef S2L_DBG_NOSAVEPE
                               (TRAIN && tile < n_tiles) ? reinterpret_cast<uint4*>(a.save_pe + ((size_t)tile * TC_TM + r) * 64) : nullptr);
#else
                               nullptr);
#endif
      {
        const float* fb = a.frame_bias + (size_t)f * 4 * 256 + 512;    // rows 2,3: folded bias0', bias5'
        float* dst = fbias_s + buf * 512;
        for (int i = r; i < 512; i += 128) dst[i] = fb[i];
      }
      fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core's async proxy
      mbar_arrive(&pe_full[buf]);
    }
  } else if (warp >= 8) {
    // =============================================================== epilogue
    const int quad = warp & 3, half = (warp - 8) >> 2;
    const uint32_t lane_sel = (uint32_t)(quad * 32) << 16;
    const int row = quad * 32 + lane;
    uint32_t acc_par[4] = {0, 0, 0, 0};
    int rp = 0;
    long long it = 0;
    int fcur = 0;
    uint32_t n_staged = 0;                                // TRAIN: slices this warp has handed to the TMA engine
    uint8_t* const stage_slot = smem + SM_STG + kGranPlane + (warp - 8) * 2048;   // + (n_staged & 3) * T1_STAGE
    TL_DECL;
    for (long long tile = blockIdx.x; tile < tile_end; tile += gridDim.x, ++it) {
      const int buf = (int)(it & 1);
      int f; long long p0, Pf;
      tile_locate<TC_TM>(a.src, a.tiles_per_frame, n_tiles, tile, fcur, f, p0, Pf);
      const long long p = p0 + row;
      mbar_wait_wd(&pe_full[buf], (uint32_t)((it >> 1) & 1), 600 + buf);   // folded per-frame biases staged
      for (int g = 0; g < 8; ++g) {
        const uint32_t d_region = tmem_base + (rp ? 256u : 0u);
        const float* bias = (g == 0) ? (fbias_s + buf * 512) : (g == 5) ? (fbias_s + buf * 512 + 256) : (tcbias_s + g * 256);
#pragma unroll 1
        for (int hh = 0; hh < 2; ++hh) {               // accumulators complete per 128-column half
          mbar_wait_wd(&acc_full[hh], acc_par[hh], 700 + hh);
          acc_par[hh] ^= 1;
          tc_fence_after();
          if (tid == 256) TL(1, 7000 + g * 10 + hh);       // epilogue observed accumulator half
          // both quarters of the half are loaded up front so the second load's latency hides behind the
          // first quarter's conversion; each quarter is released to the MMA thread as soon as it is stored
          const uint32_t taddr0 = d_region + lane_sel + (uint32_t)(hh * 128 + half * 32);
          uint32_t va[32], vb[32];
#ifdef S2L_DBG_NOEPI         // experiment: the epilogue only keeps the barrier protocol alive
          for (int qq = 0; qq < 2; ++qq) { tc_fence_before(); mbar_arrive(&epi_done[hh * 2 + qq]); }
          continue;
#endif
          tmem_ld32(taddr0, va);
          tmem_ld32(taddr0 + 64u, vb);
          tmem_ld_wait();
#pragma unroll
          for (int qq = 0; qq < 2; ++qq) {
            const int q = hh * 2 + qq;
            const uint32_t taddr = taddr0 + (uint32_t)(qq * 64);
            const float4* b4 = reinterpret_cast<const float4*>(bias + q * 64 + half * 32);
            uint32_t o[32];
            if (qq) convert_slice<NPASS>(vb, b4, o);
            else convert_slice<NPASS>(va, b4, o);
            if (NPASS != 1) tmem_st32(taddr, o);
            else tmem_st16(taddr, o);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&epi_done[q]);
            // (the quarter is released to the MMA thread BEFORE its copy goes to global memory: the stores overlap the next
            //  layer's first MMAs instead of delaying them)
#ifndef S2L_DBG_NOSAVEH
            if (TRAIN && tile < n_tiles) {     // h_g as the next layer consumes it: 32 bf16 = 64 B of this row
              uint8_t* slot = stage_slot + (n_staged & 3u) * T1_STAGE;
              if (lane == 0) bulk_wait_group_read<3>();       // the store that last read this slot (four slices ago) has drained it
              __syncwarp();
              stage_rows64_sw64(slot, lane, o);
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_2d(&hmap, slot, q * 64 + half * 32, (int)((long long)g * a.rows_total + tile * TC_TM + quad * 32));
                bulk_commit_group();
              }
              ++n_staged;
              // the ReLU mask of the same 32 values as one word: the data-gradient kernel reads 4 B instead of these 64 B
              uint32_t m = 0;
#pragma unroll
              for (int j = 0; j < 16; ++j)
                m |= ((o[j] & 0x00007FFFu) ? (1u << (2 * j)) : 0u) | ((o[j] & 0x7FFF0000u) ? (2u << (2 * j)) : 0u);
              a.save_mask[(((size_t)g * n_tiles + tile) * 8 + (q * 2 + half)) * TC_TM + row] = m;
            }
#endif
            if (tid == 256) TL(1, 9000 + g * 10 + q);      // quarter released to the MMA thread
          }
        }
        if (g == 5) mbar_arrive(&pe_empty[buf]);      // this thread no longer reads fbias_s[buf]
        rp ^= 1;
      }
      // ---- G8: raw output (no activation), tf_nerf.py:283
      {
        const uint32_t d_region = tmem_base + (rp ? 256u : 0u);
        mbar_wait_wd(acc_out, (uint32_t)(it & 1), 800);
        tc_fence_after();
        if (half == 0) {
          uint32_t v[4];
          tmem_ld4(d_region + lane_sel, v);
          tmem_ld_wait();
          const float* bo = tcbias_s + 8 * 256;
          if (a.epi_mode != EPI_RAW) {
            // hand the tile to the reducer warp (fused 4-tap blend / alpha compositing)
            mbar_wait_wd(&raw_empty[buf], (uint32_t)(((it >> 1) & 1) ^ 1), 810 + buf);
            rawbuf[buf * TC_TM + row] = make_float4(__uint_as_float(v[0]) + bo[0], __uint_as_float(v[1]) + bo[1],
                                                    __uint_as_float(v[2]) + bo[2], __uint_as_float(v[3]) + bo[3]);
            mbar_arrive(&raw_full[buf]);
          } else if (p < Pf) {
            float* o = a.out + ((long long)f * a.src.P + p) * a.out_ch;
#pragma unroll
            for (int n = 0; n < 4; ++n)
              if (n < a.out_ch) o[n] = __uint_as_float(v[n]) + bo[n];
          }
        }
        rp ^= 1;
      }
    }
    if (TRAIN && lane == 0) bulk_wait_group<0>();          // every staged slice has reached global memory before the CTA retires
  }

  tc_fence_before();
  __syncthreads();
#ifdef S2L_TIMELINE
  if (tid == 0 && blockIdx.x == 0 && a.dbg) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    a.dbg[3 * 8192 + 2] = clock64();
    a.dbg[3 * 8192 + 3] = (long long)gt;
  }
#endif
  if (CL > 1) cluster_sync_all();      // no CTA may retire while its peer can still multicast into it
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

template <int NPASS, int UVD, int CL, bool TRAIN = false>
static int launch_tc_impl(const TcArgs& a, long long n_tiles, cudaStream_t st, const CUtensorMap* hmap = nullptr) {
  static bool attr_set_dev[64] = {};   // cudaFuncSetAttribute is per device
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  bool& attr_set = attr_set_dev[cur_dev & 63];
  if (!attr_set) {
    if (cudaFuncSetAttribute(mlp_tc_kernel<NPASS, UVD, CL, TRAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, T1_SMEM_BYTES) != cudaSuccess) {
      set_error("mlp_tc: cannot opt in to %d B of shared memory: %s", T1_SMEM_BYTES, cudaGetErrorString(cudaGetLastError()));
      return 6;
    }
    attr_set = true;
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cur_dev);
  const long long want = (n_tiles + CL - 1) / CL;
  const unsigned grid = (unsigned)((want < sms / CL ? want : sms / CL) * CL);      // whole clusters, one CTA per SM
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = T1_SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute at;
  at.id = cudaLaunchAttributeClusterDimension;
  at.val.clusterDim.x = CL;
  at.val.clusterDim.y = 1;
  at.val.clusterDim.z = 1;
  cfg.attrs = &at;
  cfg.numAttrs = 1;
  static const CUtensorMap no_map = {};
  cudaLaunchKernelEx(&cfg, mlp_tc_kernel<NPASS, UVD, CL, TRAIN>, a, hmap ? *hmap : no_map);
  return check_launch(TRAIN ? "mlp_tc_kernel<train>" : "mlp_tc_kernel") ? 0 : 5;
}

template <int CL>
static int launch_tc_cl(const TcArgs& a, long long n_tiles, int npass, int uvd, cudaStream_t st) {
  if (uvd == 2)
    return npass == 3 ? launch_tc_impl<3, 2, CL>(a, n_tiles, st) : npass == 2 ? launch_tc_impl<2, 2, CL>(a, n_tiles, st) : launch_tc_impl<1, 2, CL>(a, n_tiles, st);
  return npass == 3 ? launch_tc_impl<3, 3, CL>(a, n_tiles, st) : npass == 2 ? launch_tc_impl<2, 3, CL>(a, n_tiles, st) : launch_tc_impl<1, 3, CL>(a, n_tiles, st);
}

int launch_mlp_tc2(const TcArgs& a, long long n_tiles, int npass, cudaStream_t st);   // s2l_mlp_tc2.cu (CTA pairs)
void profile_mark(cudaStream_t st, int which);                                         // s2l_capi.cu

// Which tensor-core schedule runs a launch of n_tiles 128-point tiles:
//   1 = independent CTAs (this file, CL = 1)          2 = CTA pairs, cta_group::2 MMAs sharing every B tile (s2l_mlp_tc2.cu)
//   3 = independent MMAs, 2-CTA clusters sharing one multicast weight stream (this file, CL = 2)
// S2L_TC_IMPL=1|2|3 forces one; by default launches that keep the whole chip busy for several tiles per SM use the pair
// kernel (in the sustained, power-capped regime it halves the L2 -> SMEM weight bytes per SM: +6 % SM clock, +2.5-3 %
// frames/s, round-1e measurements) and small launches the single-CTA kernel (shorter dependency chains: ~4 % faster in
// short bursts).
static int tc_forced() {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("S2L_TC_IMPL");
    forced = (e && e[0] >= '1' && e[0] <= '3') ? e[0] - '0' : 0;
  }
  return forced;
}
int tc_impl_for(long long n_tiles) {
  const int forced = tc_forced();
  if (forced) return forced;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return (n_tiles >= 8ll * sms && sms % 2 == 0) ? 2 : 1;
}
extern "C" int32_t s2l_tc_schedule(int64_t n_tiles) { return tc_impl_for(n_tiles); }

static long long* g_timeline = nullptr;
extern "C" void s2l_debug_set_timeline(long long* buf) { g_timeline = buf; }     // debug builds (tools/tc_timeline.py)

int launch_mlp_tc(const void* blob, const PointSrc& src, int n_frames, const float* frame_bias, float* out, int out_ch,
                  int npass, cudaStream_t st, const TcEpi* epi, const Gate* gate) {
  TcArgs a{};
  if (gate) a.gate = *gate;
  a.blob = reinterpret_cast<const uint8_t*>(blob);
  a.L = blob_layout();
  a.src = src;
  a.frame_bias = frame_bias;
  a.out = out;
  a.out_ch = out_ch;
  a.n_frames = n_frames;
  a.tiles_per_frame = (src.P + TC_TM - 1) / TC_TM;
  a.dbg = g_timeline;
  if (epi) {
    a.epi_mode = epi->mode;
    a.rgb = epi->rgb;
    a.carry = epi->carry;
    a.next_count = epi->next_count;
    a.next_rays = epi->next_rays;
    a.term_thr = epi->term_thr;
    a.fix_thr = epi->fix_thr;
    if (epi->mode == EPI_COMPOSITE && (src.mode != S2L_PTS_RAYS || src.Sc < 4 || (src.Sc & 3) || TC_TM % src.Sc)) {
      set_error("mlp_tc: fused compositing needs ray mode with a sample-chunk size in {4,8,...,128} (got %d)", src.Sc);
      return 2;
    }
    if (epi->mode == EPI_ENS4 && src.mode != S2L_PTS_GRID_ENS4) { set_error("mlp_tc: fused 4-tap blend needs GRID_ENS4 points"); return 2; }
  }
  if (src.list_count && (!src.tile_start || !src.list_rays || !epi || epi->mode != EPI_COMPOSITE)) {
    set_error("mlp_tc: ray lists need tile_start / list_rays and the fused compositing epilogue");
    return 2;
  }
  // list launches learn their tile count on the device (tile_start[F]); the grid is sized for the upper bound
  const long long n_tiles = a.tiles_per_frame * n_frames;
  if (n_tiles == 0) return 0;
  if (src.uv_dims != 2 && src.uv_dims != 3) { set_error("mlp_tc: unsupported uv_dims %d", src.uv_dims); return 2; }
  const int impl = tc_impl_for(n_tiles);
  struct ProfScope {       // bench.py's live kernel timing (s2l_profile_enable)
    cudaStream_t st;
    explicit ProfScope(cudaStream_t s) : st(s) { profile_mark(st, 0); }
    ~ProfScope() { profile_mark(st, 1); }
  } prof_scope(st);
  if (impl == 2) {
    const int r = launch_mlp_tc2(a, n_tiles, npass, st);
    if (r == 0 || tc_forced()) return r;
    // the automatically chosen pair schedule is unavailable here (no tensor-map encoder in the driver, a partition
    // that cannot co-schedule 2-CTA clusters, ...): same arithmetic on independent CTAs — still this library's kernel
    cudaGetLastError();
  }
  if (impl == 3) return launch_tc_cl<2>(a, n_tiles, npass, src.uv_dims, st);
  return launch_tc_cl<1>(a, n_tiles, npass, src.uv_dims, st);
}

// Training forward (s2l_train_fwd): the live 4-tap render of F frames in bf16 with the fused blend epilogue, saving the
// activations the backward needs.  Always the single-CTA schedule (training launches are a few thousand tiles).
int launch_mlp_tc_train(const void* blob, const PointSrc& src, int n_frames, const float* frame_bias, float* rgb,
                        __nv_bfloat16* save_h, __nv_bfloat16* save_pe, uint32_t* save_mask, cudaStream_t st, float* raw_out) {
  TcArgs a{};
  a.blob = reinterpret_cast<const uint8_t*>(blob);
  a.L = blob_layout();
  a.src = src;
  a.frame_bias = frame_bias;
  a.out_ch = 3;
  a.n_frames = n_frames;
  a.tiles_per_frame = (src.P + TC_TM - 1) / TC_TM;
  a.epi_mode = raw_out ? EPI_RAW : EPI_ENS4;      // raw_out: the per-call rows contract (explicit points, raw [N,3] outputs)
  a.rgb = rgb;
  a.out = raw_out;
  a.save_h = save_h;
  a.save_pe = save_pe;
  a.save_mask = save_mask;
  const long long n_tiles = a.tiles_per_frame * n_frames;
  a.rows_total = n_tiles * TC_TM;
  if (n_tiles == 0) return 0;
  if (src.uv_dims != 2 || (raw_out ? src.mode != S2L_PTS_EXPLICIT : src.mode != S2L_PTS_GRID_ENS4)) {
    set_error("mlp_tc_train: 4-tap live render (GRID_ENS4) or explicit rows, uv_dims = 2");
    return 2;
  }
  CUtensorMap hmap;
  if (!encode_2d(&hmap, save_h, 256, 8ull * (unsigned long long)a.rows_total, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B)) return 6;
  return launch_tc_impl<1, 2, 1, true>(a, n_tiles, st, &hmap);
}

}  // namespace s2l

#ifdef S2L_DBG_SATCOUNT
extern "C" unsigned long long s2l_debug_sat_count_tc(void) {
  unsigned long long v = 0, z = 0;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(&v, s2l::g_sat_count, sizeof(v));
  cudaMemcpyToSymbol(s2l::g_sat_count, &z, sizeof(z));
  return v;
}
#endif
