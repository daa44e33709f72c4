// Fused MLP on CTA PAIRS (tcgen05 cta_group::2): the layer program, TMEM residency, precision modes, static issue
// program and epilogue of s2l_mlp_tc.cu, with two CTAs of a cluster working as one 256-row tile:
//   * every tcgen05.mma is M=256 x N=128 x K=16/32 (64 tensor cycles, both SMs' tensor cores busy), issued by the
//     converged MMA warp of the pair's leader CTA; each CTA keeps the A operand / accumulators of its own 128 rows in
//     its own TMEM and supplies HALF of the B operand from its own shared memory (CTA r holds weight rows
//     [64 r, 64 r + 64) of every [128 N x 64 K] granule).  Per SM the weight bytes streamed from L2 and the
//     shared-memory operand reads are HALVED (the peer's half arrives over the SM-to-SM path); in the power-capped
//     regime the headline number is measured in, that buys +6 % SM clock / +2.5-3 % frames/s.
//   * the N=128 accumulator halves keep the epilogue / MMA overlap of the single-CTA kernel (half 0 is converted while
//     half 1 accumulates).
//   * cross-CTA flow control runs over mbarriers in the leader's shared memory (remote arrives through mapa):
//     "pair_full" (both CTAs' granule halves landed: the weight copies are TMA *tensor* copies with .cta_group::2, the
//     one bulk-copy form whose complete_tx may target the PEER CTA's mbarrier, so both producers signal the leader's
//     barrier directly — a relaying thread per CTA costs a ~1 us remote round trip per granule and starves the ring),
//     "epi_done" (both epilogues
//     converted a 64-column quarter; one arrive per warp), "pe_pair" (both positional-encoding images written);
//     tcgen05.commit multicasts accumulator-ready / stage-free / PE-free events to both CTAs.
//   * every CTA of the grid runs the same number of tile iterations (iterations past the end compute on zero rows and
//     store nothing) so the pair stays in lock step.
#include <cstdlib>
#include <type_traits>
#include <cuda.h>          // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)
#include "s2l_tc_common.cuh"

namespace s2l {

constexpr int T2_NSTG = 8;                           // power of two
constexpr int T2_STAGE = kGranBytes / 2;             // 16 KB: this CTA's 64 rows of both planes of a granule
constexpr int T2_SM_TCBIAS = SM_STG + T2_NSTG * T2_STAGE;
constexpr int T2_SM_FBIAS = T2_SM_TCBIAS + kNumG * 256 * 4;
constexpr int T2_SM_BAR = T2_SM_FBIAS + 2 * 2 * 256 * 4;
// barrier map (64-bit slots): local b_empty[8] pe_full[2] pe_empty[2] acc_full[2] (slots 0-7 unused);
// leader-side pair barriers pair_full[8] epi_done[4] pe_pair[2]
constexpr int T2_BEMPTY = 8, T2_PEFULL = 16, T2_PEEMPTY = 18, T2_ACC = 20, T2_PAIRFULL = 22, T2_EPIDONE = 30,
              T2_PEPAIR = 34, T2_RAWFULL = 36, T2_RAWEMPTY = 38, T2_ACCOUT = 40, T2_NBAR = 42;
constexpr int T2_SM_TMEMPTR = T2_SM_BAR + T2_NBAR * 8;
constexpr int T2_SM_RAW = T2_SM_TMEMPTR + 16;          // 2 x [128] float4: output tile handed to the reducer warp
constexpr int T2_SMEM_BYTES = T2_SM_RAW + 2 * TC_TM * 16;
static_assert(T2_SMEM_BYTES <= 232448, "shared memory budget");

// waits that must observe writes made by the peer CTA (acquire at cluster scope)
__device__ __forceinline__ bool mbar_try_wait_cl(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}
// bounded, printf-free (the MMA warp keeps its loop state in uniform registers: no calls inside its loop)
__device__ __forceinline__ void mbar_wait_cl(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait_cl(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_cl(bar, parity)) {
    if (clock64() - t0 > kWatchdogCycles) __trap();
  }
}
// cta_group::2 MMAs (issued by the leader CTA only) and the commit that signals both CTAs
__device__ __forceinline__ void umma2_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
               "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void umma2_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
               "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void umma2_8ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
               "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void umma2_8ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f8f6f4 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
               "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
// 2-D tensor copy global -> own shared memory whose completion bytes are credited to a barrier that may live in the
// peer CTA (shared::cluster address).  The weight section is described as [rows][256 B]; a box is 32 (tm8) or 16
// (tm4) rows, stored densely (no swizzle: the blob already holds the swizzled operand images).
__device__ __forceinline__ void tma2_load(void* dst_smem, const CUtensorMap* tm, int row, uint32_t mbar_cluster_addr) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst_smem)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(mbar_cluster_addr), "r"(0), "r"(row)
               : "memory");
}
// instruction descriptors with M = 256 (pair)
__host__ __device__ constexpr uint32_t idesc2(uint32_t a_fmt, uint32_t b_fmt, int n) {
  return (1u << 4) | (a_fmt << 7) | (b_fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}


template <int NPASS, int UVD>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1) mlp_tc2_kernel(const __grid_constant__ TcArgs a, const __grid_constant__ CUtensorMap tm8,
                                                                                            const __grid_constant__ CUtensorMap tm4) {
  if (a.gate.flag && *a.gate.flag != a.gate.value) return;      // gated launch: the other implementation serves this call
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + T2_SM_BAR);
  uint64_t* b_empty = bars + T2_BEMPTY;
  uint64_t* pe_full = bars + T2_PEFULL;
  uint64_t* pe_empty = bars + T2_PEEMPTY;
  uint64_t* acc_full = bars + T2_ACC;
  uint64_t* pair_full = bars + T2_PAIRFULL;     // used in the leader only
  uint64_t* epi_done = bars + T2_EPIDONE;       // used in the leader only
  uint64_t* pe_pair = bars + T2_PEPAIR;         // used in the leader only
  uint64_t* raw_full = bars + T2_RAWFULL;
  uint64_t* raw_empty = bars + T2_RAWEMPTY;
  uint64_t* acc_out = bars + T2_ACCOUT;         // the output layer's own accumulator-ready barrier (see s2l_mlp_tc.cu)
  float4* rawbuf = reinterpret_cast<float4*>(smem + T2_SM_RAW);
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem + T2_SM_TMEMPTR);
  float* tcbias_s = reinterpret_cast<float*>(smem + T2_SM_TCBIAS);
  float* fbias_s = reinterpret_cast<float*>(smem + T2_SM_FBIAS);   // [2 bufs][2][256]

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // provably warp-uniform
  const uint32_t rank = cluster_ctarank();
  const long long n_tiles = launch_tiles(a);
  const long long n_iter = (n_tiles + gridDim.x - 1) / gridDim.x;      // same for every CTA: pairs stay in lock step
  const long long tile_end = (long long)blockIdx.x + n_iter * gridDim.x;
  const uint8_t* tcw = a.blob + (NPASS == 2 ? a.L.off_tcw8 : a.L.off_tcw);

  if (tid == 0) {
    for (int s = 0; s < T2_NSTG; ++s) {
      mbar_init(&b_empty[s], 1);
      mbar_init(&pair_full[s], 1);        // the leader's arrive.expect_tx; both CTAs' copies complete_tx on it
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&pe_full[b], 128);
      mbar_init(&pe_empty[b], 1 + 256);
      mbar_init(&pe_pair[b], 256);
      mbar_init(&acc_full[b], 1);
      mbar_init(&raw_full[b], 128);
      mbar_init(&raw_empty[b], 1);
    }
    for (int q = 0; q < 4; ++q) mbar_init(&epi_done[q], 16);       // one arrive per epilogue warp of either CTA
    mbar_init(acc_out, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  {
    const float* TB = reinterpret_cast<const float*>(a.blob + a.L.off_tcbias);
    for (int i = tid; i < kNumG * 256; i += TC_THREADS) tcbias_s[i] = TB[i];
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // peer barriers initialised, both TMEM allocations done
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_s, 0);

  if (warp == 0) {
    // =============================================================== weight producer: this CTA's rows of every granule
    if (elect_one()) {
      uint32_t stage = 0, phase = 0;
      const uint32_t pf0 = mapa_u32(smem_u32(&pair_full[0]), 0);       // the leader's barriers (cluster addresses)
      const int sec_row = (int)(((NPASS == 2 ? a.L.off_tcw8 : a.L.off_tcw) - a.L.off_tcw) >> 8);   // tensor maps start at TCW; TCW8 lies behind it
      (void)tcw;
      for (long long tile = blockIdx.x; tile < tile_end; tile += gridDim.x) {
#pragma unroll 1
        for (int g = 0; g < kNumG; ++g) {
          const int ngran = (g == 8) ? 4 : 2 * g_nkc(g);
          const uint32_t gran = (g == 8) ? kOutGranBytes : kGranBytes;
          const int row0 = sec_row + (g_layer_off(g) >> 8);
#pragma unroll 1
          for (int gi = 0; gi < ngran; ++gi) {
            const int G = row0 + (int)((gi * gran) >> 8);                 // first 256-byte row of the granule
            uint8_t* dst = smem + SM_STG + stage * T2_STAGE;
            const uint32_t pf = pf0 + stage * 8u;
            mbar_wait_wd<false>(&b_empty[stage], phase ^ 1u, 100 + stage);     // no back-off: the refill latency of the pair's ring is on the critical path
            constexpr uint32_t kHalf = kGranPlane / 2;      // 8 KB: 64 rows of a 16-bit plane (8-row atoms of 1 KB)
#ifdef S2L_DBG_NOLOAD        // experiment: weights are never streamed (results are garbage, timing only)
            if (rank == 0) mbar_arrive(&pair_full[stage]);
            (void)dst; (void)pf; (void)G;
#else
            if (g == 8) {
              // output layer (N padded to 16): both CTAs take the same 16 rows -> the whole 4 KB granule each
              if (rank == 0) mbar_arrive_expect_tx(&pair_full[stage], 2 * kOutGranBytes);
              tma2_load(dst, &tm4, G, pf);
            } else {
              if (rank == 0) mbar_arrive_expect_tx(&pair_full[stage], (NPASS == 1) ? 2 * kHalf : 4 * kHalf);
              tma2_load(dst, &tm8, G + (int)((rank * kHalf) >> 8), pf);
              if (NPASS == 3) tma2_load(dst + kHalf, &tm8, G + (int)((kGranPlane + rank * kHalf) >> 8), pf);
              if (NPASS == 2) {   // second plane = [e5m2 8 KB | e4m3 8 KB], 64-byte rows in 8-row atoms of 512 B
                tma2_load(dst + kHalf, &tm4, G + (int)((kGranPlane + rank * (kHalf / 2)) >> 8), pf);
                tma2_load(dst + kHalf + kHalf / 2, &tm4, G + (int)((kGranPlane + kGranPlane / 2 + rank * (kHalf / 2)) >> 8), pf);
              }
            }
#endif
            stage = (stage + 1) & (T2_NSTG - 1);
            phase ^= (stage == 0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================================================== MMA issuer: leader CTA only
    // Converged warp, static program (see s2l_mlp_tc.cu): with an 8-stage ring the stage at each program point is
    // fixed: G0 at 0, G1-4 at 2, G5 at 2 (10 granules), G6-7 at 4, G8 at 4.
    if (rank == 0) {
      uint32_t ph = 0;                            // bit s = parity to wait for on pair_full[s]
      uint32_t epi_par = 0;                       // bit hk = parity to wait for on epi_done[hk]
      int rp = 0;
      long long it = 0;
      TL_DECL;
      constexpr uint32_t kDescHi = 0x40004040u;   // SBO=64 | version=1 | SWIZZLE_128B
      constexpr uint32_t kDescHi64 = 0x80004020u; // SBO=32 | version=1 | SWIZZLE_64B
      auto mk = [](uint32_t lo) -> uint64_t { return ((uint64_t)kDescHi << 32) | lo; };
      auto mk64 = [](uint32_t lo) -> uint64_t { return ((uint64_t)kDescHi64 << 32) | lo; };
      const uint32_t stg0 = ((smem_u32(smem + SM_STG) >> 4) & 0x3FFFu) | 0x10000u;
      uint32_t pe_hi = 0, d_region = 0, a_region = 0;
      int buf = 0;

      auto granule = [&](auto stage_c, auto is_pe_c, auto small_c, uint32_t d_addr, uint32_t a_t, uint32_t acc0,
                         uint64_t* done0, uint64_t* done1) {
        constexpr int STAGE = decltype(stage_c)::value;
        constexpr bool IS_PE = decltype(is_pe_c)::value;
        constexpr bool SMALL = decltype(small_c)::value;
        constexpr int n_mma = SMALL ? 2 * kOutPad : kGranRows;              // pair-wide N
        constexpr uint32_t idesc = (NPASS == 2) ? idesc2(0u, 0u, n_mma) : idesc2(1u, 1u, n_mma);   // f16 x f16 | bf16 x bf16
        constexpr uint32_t idesc_rw = idesc2(0u, 1u, n_mma);     // (A - fp16 A) [e4m3] x fp16(W) [e5m2]
        constexpr uint32_t idesc_wr = idesc2(1u, 0u, n_mma);     // fp16(A) [e5m2] x (W - fp16 W) [e4m3]
        constexpr uint32_t plane16 = (uint32_t)((SMALL ? kOutPlane : kGranPlane / 2) >> 4);   // this CTA's share of a plane
        TLC(3);
        mbar_wait_cl(&pair_full[STAGE], (ph >> STAGE) & 1u);
        TLC(2);
        ph ^= 1u << STAGE;
        tc_fence_after();
        const uint32_t b = stg0 + (uint32_t)STAGE * (uint32_t)(T2_STAGE >> 4);      // first plane: hi / fp16
        const uint32_t b2 = b + plane16;                                            // second plane: lo / [e5m2 | e4m3]
        const uint32_t b4 = b2 + (plane16 >> 1);
        const uint32_t pe_lo = pe_hi + (PE_PLANE >> 4);
        const uint32_t pe_e5 = pe_lo, pe_e4 = pe_lo + (PE_PLANE >> 5);
        if (elect_one()) {
          if (IS_PE) {
#pragma unroll
            for (int s = 0; s < 4; ++s) {
              umma2_ss(d_addr, mk(pe_hi + 2 * s), mk(b + 2 * s), idesc, s == 0 ? acc0 : 1u);
              if (NPASS == 3) umma2_ss(d_addr, mk(pe_lo + 2 * s), mk(b + 2 * s), idesc, 1u);
            }
            if (NPASS == 3) {
#pragma unroll
              for (int s = 0; s < 4; ++s) umma2_ss(d_addr, mk(pe_hi + 2 * s), mk(b2 + 2 * s), idesc, 1u);
            }
            if (NPASS == 2) {
#pragma unroll
              for (int t = 0; t < 2; ++t) umma2_8ss(d_addr, mk64(pe_e4 + 2 * t), mk64(b2 + 2 * t), idesc_rw, 1u);
#pragma unroll
              for (int t = 0; t < 2; ++t) umma2_8ss(d_addr, mk64(pe_e5 + 2 * t), mk64(b4 + 2 * t), idesc_wr, 1u);
            }
          } else {
#pragma unroll
            for (int s = 0; s < 4; ++s) {
              const uint32_t a_hi = a_t + (uint32_t)((s >> 1) * 32 + (s & 1) * 8);
              umma2_ts(d_addr, a_hi, mk(b + 2 * s), idesc, s == 0 ? acc0 : 1u);
              if (NPASS == 3) umma2_ts(d_addr, a_hi + 16u, mk(b + 2 * s), idesc, 1u);
            }
            if (NPASS == 3) {
#pragma unroll
              for (int s = 0; s < 4; ++s)
                umma2_ts(d_addr, a_t + (uint32_t)((s >> 1) * 32 + (s & 1) * 8), mk(b2 + 2 * s), idesc, 1u);
            }
            if (NPASS == 2) {
#pragma unroll
              for (int t = 0; t < 2; ++t) umma2_8ts(d_addr, a_t + (uint32_t)(t * 32 + 24), mk64(b2 + 2 * t), idesc_rw, 1u);
#pragma unroll
              for (int t = 0; t < 2; ++t) umma2_8ts(d_addr, a_t + (uint32_t)(t * 32 + 16), mk64(b4 + 2 * t), idesc_wr, 1u);
            }
          }
          umma2_commit(&b_empty[STAGE]);      // both CTAs' stage reusable once these MMAs retire
          if (done0) umma2_commit(done0);
          if (done1) umma2_commit(done1);
        }
        __syncwarp();
      };
      auto wait_quarter = [&](int hk) {       // quarter hk of the previous layer converted by BOTH epilogues
        TLC(0);
        mbar_wait_cl(&epi_done[hk], (epi_par >> hk) & 1u);
        TLC(1);
        epi_par ^= 1u << hk;
      };
      using std::integral_constant;
#define S2L_IC(v) integral_constant<int, (v)>{}
#define S2L_BC(v) integral_constant<bool, (v)>{}
      auto layer_std = [&](auto start_c, int g) {
        constexpr int START = decltype(start_c)::value;
        (void)g;
        TL(0, 1000 + g * 10);
        wait_quarter(0); granule(S2L_IC((START + 0) & 7), S2L_BC(false), S2L_BC(false), d_region, a_region, 0u, nullptr, nullptr);
        wait_quarter(1); granule(S2L_IC((START + 1) & 7), S2L_BC(false), S2L_BC(false), d_region, a_region + 64u, 1u, nullptr, nullptr);
        wait_quarter(2); granule(S2L_IC((START + 2) & 7), S2L_BC(false), S2L_BC(false), d_region, a_region + 128u, 1u, nullptr, nullptr);
        wait_quarter(3); granule(S2L_IC((START + 3) & 7), S2L_BC(false), S2L_BC(false), d_region, a_region + 192u, 1u, &acc_full[0], nullptr);
        TL(0, 5000 + g * 10); TLC_FLUSH(0); TL(0, 1001 + g * 10);
        granule(S2L_IC((START + 4) & 7), S2L_BC(false), S2L_BC(false), d_region + 128u, a_region, 0u, nullptr, nullptr);
        granule(S2L_IC((START + 5) & 7), S2L_BC(false), S2L_BC(false), d_region + 128u, a_region + 64u, 1u, nullptr, nullptr);
        granule(S2L_IC((START + 6) & 7), S2L_BC(false), S2L_BC(false), d_region + 128u, a_region + 128u, 1u, nullptr, nullptr);
        granule(S2L_IC((START + 7) & 7), S2L_BC(false), S2L_BC(false), d_region + 128u, a_region + 192u, 1u, &acc_full[1], nullptr);
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
        mbar_wait_cl(&pe_pair[buf], (uint32_t)((it >> 1) & 1));
        granule(S2L_IC(0), S2L_BC(true), S2L_BC(false), d_region, 0u, 0u, &acc_full[0], nullptr);
        granule(S2L_IC(1), S2L_BC(true), S2L_BC(false), d_region + 128u, 0u, 0u, &acc_full[1], nullptr);
        // ---- G1-4 (ring position 2)
#pragma unroll 1
        for (int g = 1; g <= 4; ++g) {
          set_regions();
          layer_std(S2L_IC(2), g);
        }
        // ---- G5: PE x fold5 + W5b x h (5 granules per half; ring positions 2..6, 7..3)
        set_regions();
        granule(S2L_IC(2), S2L_BC(true), S2L_BC(false), d_region, 0u, 0u, nullptr, nullptr);
        wait_quarter(0); granule(S2L_IC(3), S2L_BC(false), S2L_BC(false), d_region, a_region, 1u, nullptr, nullptr);
        wait_quarter(1); granule(S2L_IC(4), S2L_BC(false), S2L_BC(false), d_region, a_region + 64u, 1u, nullptr, nullptr);
        wait_quarter(2); granule(S2L_IC(5), S2L_BC(false), S2L_BC(false), d_region, a_region + 128u, 1u, nullptr, nullptr);
        wait_quarter(3); granule(S2L_IC(6), S2L_BC(false), S2L_BC(false), d_region, a_region + 192u, 1u, &acc_full[0], nullptr);
        granule(S2L_IC(7), S2L_BC(true), S2L_BC(false), d_region + 128u, 0u, 0u, nullptr, nullptr);
        granule(S2L_IC(0), S2L_BC(false), S2L_BC(false), d_region + 128u, a_region, 1u, nullptr, nullptr);
        granule(S2L_IC(1), S2L_BC(false), S2L_BC(false), d_region + 128u, a_region + 64u, 1u, nullptr, nullptr);
        granule(S2L_IC(2), S2L_BC(false), S2L_BC(false), d_region + 128u, a_region + 128u, 1u, nullptr, nullptr);
        granule(S2L_IC(3), S2L_BC(false), S2L_BC(false), d_region + 128u, a_region + 192u, 1u, &acc_full[1], &pe_empty[buf]);   // last reader of the PE images
        // ---- G6-7 (ring position 4)
#pragma unroll 1
        for (int g = 6; g <= 7; ++g) {
          set_regions();
          layer_std(S2L_IC(4), g);
        }
        // ---- G8: output layer, pair-wide N = 32 (both CTAs hold the same 16 rows; ring position 4)
        set_regions();
        wait_quarter(0); granule(S2L_IC(4), S2L_BC(false), S2L_BC(true), d_region, a_region, 0u, nullptr, nullptr);
        wait_quarter(1); granule(S2L_IC(5), S2L_BC(false), S2L_BC(true), d_region, a_region + 64u, 1u, nullptr, nullptr);
        wait_quarter(2); granule(S2L_IC(6), S2L_BC(false), S2L_BC(true), d_region, a_region + 128u, 1u, nullptr, nullptr);
        wait_quarter(3); granule(S2L_IC(7), S2L_BC(false), S2L_BC(true), d_region, a_region + 192u, 1u, acc_out, nullptr);
      }
#undef S2L_IC
#undef S2L_BC
    }
  } else if (warp == 3) {
    // =============================================================== reducer (fused 4-tap blend / alpha compositing)
    if (a.epi_mode != EPI_RAW) reducer_role(a, n_tiles, tile_end, rawbuf, raw_full, raw_empty, lane);
  } else if (warp >= 4 && warp < 8) {
    // =============================================================== PE producers (one point per thread)
    const int r = tid - 128;
    const uint32_t pp0 = mapa_u32(smem_u32(&pe_pair[0]), 0);
    long long it = 0;
    int fcur = 0;
    for (long long tile = blockIdx.x; tile < tile_end; tile += gridDim.x, ++it) {
      const int buf = (int)(it & 1);
      int f; long long p0, Pf;
      tile_locate<TC_TM>(a.src, a.tiles_per_frame, n_tiles, tile, fcur, f, p0, Pf);
      const long long p = p0 + r;
      mbar_wait_wd<true>(&pe_empty[buf], (uint32_t)(((it >> 1) & 1) ^ 1), 500 + buf);
      pe_write_row<NPASS, UVD>(a.src, f, p, p < Pf, r, smem + SM_PE + buf * PE_BUF);
      {
        const float* fb = a.frame_bias + (size_t)f * 4 * 256 + 512;    // rows 2,3: folded bias0', bias5'
        float* dst = fbias_s + buf * 512;
        for (int i = r; i < 512; i += 128) dst[i] = fb[i];
      }
      fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core's async proxy
      mbar_arrive(&pe_full[buf]);                          // local: the epilogue's per-frame biases
      mbar_arrive_remote(pp0 + (uint32_t)buf * 8u);        // leader: the pair's A operand images
    }
  } else if (warp >= 8) {
    // =============================================================== epilogue (own 128 rows)
    const int quad = warp & 3, half = (warp - 8) >> 2;
    const uint32_t lane_sel = (uint32_t)(quad * 32) << 16;
    const int row = quad * 32 + lane;
    const uint32_t ed0 = mapa_u32(smem_u32(&epi_done[0]), 0);
    uint32_t acc_par[2] = {0, 0};
    int rp = 0;
    long long it = 0;
    int fcur = 0;
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
          const uint32_t taddr0 = d_region + lane_sel + (uint32_t)(hh * 128 + half * 32);
          uint32_t va[32], vb[32];
#ifdef S2L_DBG_NOEPI         // experiment: the epilogue only keeps the barrier protocol alive
          for (int qq = 0; qq < 2; ++qq) { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive_remote(ed0 + (uint32_t)(hh * 2 + qq) * 8u); }
          (void)taddr0; (void)bias;
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
            __syncwarp();
            if (lane == 0) mbar_arrive_remote_relaxed(ed0 + (uint32_t)q * 8u);       // the leader's MMA warp counts both CTAs' warps
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
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // no CTA may retire (or free TMEM) while its peer can still signal it
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// Tensor maps over the blob's TCW..TCW8 range seen as [rows][256 B] bytes, boxes of 32 / 16 rows (8 KB / 4 KB).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static bool make_weight_maps(const TcArgs& a, CUtensorMap* tm8, CUtensorMap* tm4) {
  static EncodeTiledFn enc = nullptr;
  if (!enc) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
      set_error("mlp_tc2: cuTensorMapEncodeTiled is not available from this driver");
      return false;
    }
    enc = reinterpret_cast<EncodeTiledFn>(fn);
  }
  const size_t span = a.L.off_tcw8 + kTcwBytes - a.L.off_tcw;          // TCW, DGRAD, TCW8 (contiguous sections)
  void* base = const_cast<uint8_t*>(a.blob + a.L.off_tcw);
  const cuuint64_t dims[2] = {256, (cuuint64_t)(span >> 8)};
  const cuuint64_t strides[1] = {256};
  const cuuint32_t estr[2] = {1, 1};
  for (int i = 0; i < 2; ++i) {
    const cuuint32_t box[2] = {256, i == 0 ? 32u : 16u};
    const CUresult r = enc(i == 0 ? tm8 : tm4, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("mlp_tc2: cuTensorMapEncodeTiled failed (%d)", (int)r);
      return false;
    }
  }
  return true;
}

template <int NPASS, int UVD>
static int launch_tc2_impl(const TcArgs& a, long long n_tiles, cudaStream_t st) {
  static bool attr_set_dev[64] = {};
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  bool& attr_set = attr_set_dev[cur_dev & 63];
  if (!attr_set) {
    if (cudaFuncSetAttribute(mlp_tc2_kernel<NPASS, UVD>, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM_BYTES) != cudaSuccess) {
      set_error("mlp_tc2: cannot opt in to %d B of shared memory: %s", T2_SMEM_BYTES, cudaGetErrorString(cudaGetLastError()));
      return 6;
    }
    attr_set = true;
  }
  CUtensorMap tm8, tm4;
  if (!make_weight_maps(a, &tm8, &tm4)) return 7;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cur_dev);
  const long long want = (n_tiles + 1) / 2;
  const unsigned grid = (unsigned)((want < sms / 2 ? want : sms / 2) * 2);      // whole pairs
  mlp_tc2_kernel<NPASS, UVD><<<grid, TC_THREADS, T2_SMEM_BYTES, st>>>(a, tm8, tm4);
  return check_launch("mlp_tc2_kernel") ? 0 : 5;
}

int launch_mlp_tc2(const TcArgs& a, long long n_tiles, int npass, cudaStream_t st) {
  if (getenv("S2L_TEST_FAIL_PAIR")) {     // test hook: behave like a device / driver that cannot run the pair schedule
    set_error("mlp_tc2: pair schedule disabled by S2L_TEST_FAIL_PAIR");
    return 7;
  }
  if (a.src.uv_dims == 2)
    return npass == 3 ? launch_tc2_impl<3, 2>(a, n_tiles, st) : npass == 2 ? launch_tc2_impl<2, 2>(a, n_tiles, st) : launch_tc2_impl<1, 2>(a, n_tiles, st);
  return npass == 3 ? launch_tc2_impl<3, 3>(a, n_tiles, st) : npass == 2 ? launch_tc2_impl<2, 3>(a, n_tiles, st) : launch_tc2_impl<1, 3>(a, n_tiles, st);
}

}  // namespace s2l

#ifdef S2L_DBG_SATCOUNT
extern "C" unsigned long long s2l_debug_sat_count_tc2(void) {
  unsigned long long v = 0, z = 0;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(&v, s2l::g_sat_count, sizeof(v));
  cudaMemcpyToSymbol(s2l::g_sat_count, &z, sizeof(z));
  return v;
}
#endif
