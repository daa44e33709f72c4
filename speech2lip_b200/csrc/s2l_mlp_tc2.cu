// Fused MLP on CTA PAIRS (tcgen05 cta_group::2): same layer program, TMEM residency and precision modes as
// s2l_mlp_tc.cu, but two CTAs of a cluster work as one 256-row tile:
//   * every tcgen05.mma is M=256 x N=256 x K=16/32, issued by ONE thread of the pair's leader CTA; each CTA keeps the
//     A operand / accumulators of its own 128 rows in its own TMEM and supplies HALF of the B operand from its own
//     shared memory (CTA r holds weight rows [128 r, 128 r + 128)), so the per-SM shared-memory operand traffic and
//     the per-SM weight streaming from L2 are halved — the single-CTA kernel is bound by B-operand fetch
//     (an MMA with a fresh B tile costs ~105-120 cycles there against a 64-cycle math floor).
//   * cross-CTA flow control runs over mbarriers in the leader's shared memory (remote arrives through mapa):
//     "pair_full" (both weight planes landed — relayed by one thread per CTA), "epi_done" (both epilogues converted a
//     64-column quarter), "pe_pair" (both positional-encoding images written); tcgen05.commit multicasts
//     accumulator-ready / stage-free / PE-free events to both CTAs.
//   * every CTA of the grid runs the same number of tile iterations (iterations past the end compute on invalid rows)
//     so the pair stays in lock step.
#include "s2l_tc_common.cuh"

namespace s2l {

constexpr int T2_NSTG = 9;
// barrier map (64-bit slots from SM_BAR): local b_full[9] b_empty[9] pe_full[2] pe_empty[2] acc_full[1];
// leader-side pair barriers pair_full[9] epi_done[4] pe_pair[2]
constexpr int T2_BFULL = 0, T2_BEMPTY = 9, T2_PEFULL = 18, T2_PEEMPTY = 20, T2_ACC = 22, T2_PAIRFULL = 23, T2_EPIDONE = 32,
              T2_PEPAIR = 36, T2_NBAR = 38;
constexpr int T2_SM_TMEMPTR = SM_BAR + T2_NBAR * 8;
constexpr int T2_SMEM_BYTES = T2_SM_TMEMPTR + 16;
static_assert(T2_SMEM_BYTES <= 232448, "shared memory budget");

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
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
__device__ __forceinline__ void mbar_wait_cl(uint64_t* bar, uint32_t parity, int tag) {
  if (mbar_try_wait_cl(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_cl(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) {
      printf("s2l tc2 kernel: mbarrier wait timeout (tag %d, block %d, thread %d, parity %u)\n", tag, blockIdx.x, threadIdx.x, parity);
      __trap();
    }
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
// instruction descriptors with M = 256 (pair)
__host__ __device__ constexpr uint32_t idesc2(uint32_t a_fmt, uint32_t b_fmt, int n) {
  return (1u << 4) | (a_fmt << 7) | (b_fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

template <int NPASS, int UVD>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1) mlp_tc2_kernel(const __grid_constant__ TcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BAR);
  uint64_t* b_full = bars + T2_BFULL;
  uint64_t* b_empty = bars + T2_BEMPTY;
  uint64_t* pe_full = bars + T2_PEFULL;
  uint64_t* pe_empty = bars + T2_PEEMPTY;
  uint64_t* acc_full = bars + T2_ACC;
  uint64_t* pair_full = bars + T2_PAIRFULL;     // used in the leader only
  uint64_t* epi_done = bars + T2_EPIDONE;       // used in the leader only
  uint64_t* pe_pair = bars + T2_PEPAIR;         // used in the leader only
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem + T2_SM_TMEMPTR);
  float* tcbias_s = reinterpret_cast<float*>(smem + SM_TCBIAS);
  float* fbias_s = reinterpret_cast<float*>(smem + SM_FBIAS);   // [2 bufs][2][256]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const long long n_tiles = a.tiles_per_frame * a.n_frames;
  const long long n_iter = (n_tiles + gridDim.x - 1) / gridDim.x;      // same for every CTA: pairs stay in lock step
  const long long tile_end = (long long)blockIdx.x + n_iter * gridDim.x;
  const uint8_t* tcw = a.blob + (NPASS == 2 ? a.L.off_tcw8 : a.L.off_tcw);
  constexpr int kPl = (NPASS == 1) ? 1 : 2;                            // weight planes per granule that are streamed

  if (tid == 0) {
    for (int s = 0; s < T2_NSTG; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
      mbar_init(&pair_full[s], 2);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&pe_full[b], 128);
      mbar_init(&pe_empty[b], 1 + 256);
      mbar_init(&pe_pair[b], 256);
    }
    mbar_init(&acc_full[0], 1);
    for (int q = 0; q < 4; ++q) mbar_init(&epi_done[q], 512);
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
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == 0) {
    // =============================================================== weight producer: this CTA's half of every layer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long tile = blockIdx.x; tile < tile_end; tile += gridDim.x) {
#pragma unroll 1
        for (int g = 0; g < kNumG; ++g) {
          const int nkc = (g == 8) ? 4 : g_nkc(g);
          const uint32_t plane = (g == 8) ? kOutPlane : kGranPlane;
          // layers 0..7: granule (h = rank, kc); layer 8 (N padded to 16): both CTAs take the same 16 rows
          const uint8_t* src = tcw + g_layer_off(g) + (size_t)((g == 8) ? 0 : rank * nkc) * 2 * plane;
#pragma unroll 1
          for (int kc = 0; kc < nkc; ++kc)
            for (int pl = 0; pl < kPl; ++pl) {
              mbar_wait_wd<true>(&b_empty[stage], phase ^ 1u, 100 + stage);
              mbar_arrive_expect_tx(&b_full[stage], plane);
              bulk_g2s(smem + SM_STG + stage * kStageBytes, src + (size_t)(kc * 2 + pl) * plane, plane, &b_full[stage]);
              if (++stage == T2_NSTG) { stage = 0; phase ^= 1u; }
            }
        }
      }
    }
  } else if (warp == 3) {
    // =============================================================== relay: "my plane landed" -> leader's pair_full
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t pf0 = mapa_u32(smem_u32(&pair_full[0]), 0);
      for (long long tile = blockIdx.x; tile < tile_end; tile += gridDim.x) {
#pragma unroll 1
        for (int i = 0; i < kPl * (kGranPerTile / 2 + 2); ++i) {     // planes per tile and CTA: kPl * (30 + 4)
          mbar_wait_wd<false>(&b_full[stage], phase, 150 + stage);
          mbar_arrive_remote(pf0 + (uint32_t)stage * 8u);
          if (++stage == T2_NSTG) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================================================== MMA issuer: leader CTA only, one thread
    if (rank == 0 && elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t epi_par = 0;
      int rp = 0;
      long long it = 0;
      constexpr uint32_t kDescHi = 0x40004040u;     // SBO=64 | version=1 | SWIZZLE_128B
      constexpr uint32_t kDescHi64 = 0x80004020u;   // SBO=32 | version=1 | SWIZZLE_64B
      auto mk = [](uint32_t lo) -> uint64_t { return ((uint64_t)kDescHi << 32) | lo; };
      auto mk64 = [](uint32_t lo) -> uint64_t { return ((uint64_t)kDescHi64 << 32) | lo; };
      for (long long tile = blockIdx.x; tile < tile_end; tile += gridDim.x, ++it) {
        const int buf = (int)(it & 1);
        const uint32_t pe_hi = ((smem_u32(smem + SM_PE + buf * PE_BUF) >> 4) & 0x3FFFu) | 0x10000u;
        const uint32_t pe_lo = pe_hi + (PE_PLANE >> 4);
        const uint32_t pe_e5 = pe_lo, pe_e4 = pe_lo + (PE_PLANE >> 5);
#pragma unroll 1
        for (int g = 0; g < kNumG; ++g) {
          const uint32_t d_addr = tmem_base + (rp ? 256u : 0u);
          const uint32_t a_region = tmem_base + (rp ? 0u : 256u);
          const int n_mma = (g == 8) ? 2 * kOutPad : 256;
          const uint32_t idesc = (NPASS == 2) ? idesc2(0u, 0u, n_mma) : idesc2(1u, 1u, n_mma);   // f16 x f16 | bf16 x bf16
          const uint32_t idesc_rw = idesc2(0u, 1u, n_mma);      // e4m3 x e5m2
          const uint32_t idesc_wr = idesc2(1u, 0u, n_mma);      // e5m2 x e4m3
          const int nkc = (g == 8) ? 4 : g_nkc(g);
          const int plane8 = ((g == 8) ? kOutPlane : kGranPlane) / 2;
#pragma unroll 1
          for (int kc = 0; kc < nkc; ++kc) {
            const bool is_pe = (g == 0) || (g == 5 && kc == 0);
            const int hk = (g == 5) ? kc - 1 : kc;
            if (g == 0) mbar_wait_cl(&pe_pair[buf], (uint32_t)((it >> 1) & 1), 200 + buf);
            if (!is_pe) {
              mbar_wait_cl(&epi_done[hk], (epi_par >> hk) & 1u, 300 + hk);
              epi_par ^= 1u << hk;
            }
            const uint32_t a_t = a_region + (uint32_t)hk * 64u;
            // ---- first plane: bf16 hi / fp16
            mbar_wait_cl(&pair_full[stage], phase, 400 + stage);
            tc_fence_after();
            {
              const uint32_t b = ((smem_u32(smem + SM_STG + stage * kStageBytes) >> 4) & 0x3FFFu) | 0x10000u;
#pragma unroll
              for (int s = 0; s < 4; ++s) {
                const uint32_t acc0 = (kc == 0 && s == 0) ? 0u : 1u;
                if (is_pe) {
                  umma2_ss(d_addr, mk(pe_hi + 2 * s), mk(b + 2 * s), idesc, acc0);
                  if (NPASS == 3) umma2_ss(d_addr, mk(pe_lo + 2 * s), mk(b + 2 * s), idesc, 1u);
                } else {
                  const uint32_t a_hi = a_t + (uint32_t)((s >> 1) * 32 + (s & 1) * 8);
                  umma2_ts(d_addr, a_hi, mk(b + 2 * s), idesc, acc0);
                  if (NPASS == 3) umma2_ts(d_addr, a_hi + 16u, mk(b + 2 * s), idesc, 1u);
                }
              }
            }
            umma2_commit(&b_empty[stage]);
            if (++stage == T2_NSTG) { stage = 0; phase ^= 1u; }
            // ---- second plane: bf16 lo, or the two fp8 correction operands
            if (NPASS != 1) {
              mbar_wait_cl(&pair_full[stage], phase, 450 + stage);
              tc_fence_after();
              const uint32_t b = ((smem_u32(smem + SM_STG + stage * kStageBytes) >> 4) & 0x3FFFu) | 0x10000u;
              if (NPASS == 3) {
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                  if (is_pe) umma2_ss(d_addr, mk(pe_hi + 2 * s), mk(b + 2 * s), idesc, 1u);
                  else umma2_ts(d_addr, a_t + (uint32_t)((s >> 1) * 32 + (s & 1) * 8), mk(b + 2 * s), idesc, 1u);
                }
              } else {
                const uint32_t b4 = b + (uint32_t)(plane8 >> 4);
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                  if (is_pe) umma2_8ss(d_addr, mk64(pe_e4 + 2 * t), mk64(b + 2 * t), idesc_rw, 1u);
                  else umma2_8ts(d_addr, a_t + (uint32_t)(t * 32 + 24), mk64(b + 2 * t), idesc_rw, 1u);
                }
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                  if (is_pe) umma2_8ss(d_addr, mk64(pe_e5 + 2 * t), mk64(b4 + 2 * t), idesc_wr, 1u);
                  else umma2_8ts(d_addr, a_t + (uint32_t)(t * 32 + 16), mk64(b4 + 2 * t), idesc_wr, 1u);
                }
              }
              umma2_commit(&b_empty[stage]);
              if (++stage == T2_NSTG) { stage = 0; phase ^= 1u; }
            }
          }
          umma2_commit(&acc_full[0]);                  // whole layer accumulated (both CTAs)
          if (g == 5) umma2_commit(&pe_empty[buf]);    // last reader of this tile's PE images
          rp ^= 1;
        }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // =============================================================== PE producers (one point per thread)
    const int r = tid - 128;
    const uint32_t pp0 = mapa_u32(smem_u32(&pe_pair[0]), 0);
    long long it = 0;
    for (long long tile = blockIdx.x; tile < tile_end; tile += gridDim.x, ++it) {
      const int buf = (int)(it & 1);
      const bool live = tile < n_tiles;
      const int f = live ? (int)(tile / a.tiles_per_frame) : 0;
      const long long p = live ? (tile % a.tiles_per_frame) * TC_TM + r : a.src.P;
      mbar_wait_wd<true>(&pe_empty[buf], (uint32_t)(((it >> 1) & 1) ^ 1), 500 + buf);
      float e[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) e[i] = 0.f;
      if (p < a.src.P) {
        float x[3];
        gen_point(a.src, f, p, x);
#pragma unroll
        for (int d = 0; d < UVD; ++d) e[d] = x[d];
#pragma unroll
        for (int k = 0; k < kMultires; ++k) {
#pragma unroll
          for (int d = 0; d < UVD; ++d) {
            float sn, cs;
            sincosf(__fmul_rn(x[d], (float)(1 << k)), &sn, &cs);     // tf_nerf.py:412: p_fn(x * freq)
            e[UVD + (2 * k) * UVD + d] = sn;
            e[UVD + (2 * k + 1) * UVD + d] = cs;
          }
        }
      }
      uint8_t* hi_base = smem + SM_PE + buf * PE_BUF;
      uint8_t* lo_base = hi_base + PE_PLANE;
      const int row_off = (r >> 3) * 1024 + (r & 7) * 128;
      if (NPASS == 2) {
        constexpr float kDn = 1.0f / (float)(1 << kScaleW), kUp = (float)(1 << kScaleA);
        uint8_t* e5_base = lo_base;
        uint8_t* e4_base = lo_base + PE_PLANE / 2;
        const int row_off64 = (r >> 3) * 512 + (r & 7) * 64;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t w5[4], w4[4];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            uint32_t h[4];
            float fv[8], rs[8];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float v0 = e[16 * c + 8 * u + 2 * t], v1 = e[16 * c + 8 * u + 2 * t + 1];
              const __half2 hh = __floats2half2_rn(v0, v1);
              h[t] = *reinterpret_cast<const uint32_t*>(&hh);
              const float2 back = __half22float2(hh);
              fv[2 * t] = back.x; fv[2 * t + 1] = back.y;
              rs[2 * t] = v0 - back.x; rs[2 * t + 1] = v1 - back.y;
            }
            const int j = 2 * c + u;
            *reinterpret_cast<uint4*>(hi_base + row_off + ((j ^ (r & 7)) << 4)) = make_uint4(h[0], h[1], h[2], h[3]);
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              w5[2 * u + t] = pack_fp8x4(fv[4 * t] * kDn, fv[4 * t + 1] * kDn, fv[4 * t + 2] * kDn, fv[4 * t + 3] * kDn, __NV_E5M2);
              w4[2 * u + t] = pack_fp8x4(rs[4 * t] * kUp, rs[4 * t + 1] * kUp, rs[4 * t + 2] * kUp, rs[4 * t + 3] * kUp, __NV_E4M3);
            }
          }
          const int off64 = row_off64 + ((c ^ ((r >> 1) & 3)) << 4);
          *reinterpret_cast<uint4*>(e5_base + off64) = make_uint4(w5[0], w5[1], w5[2], w5[3]);
          *reinterpret_cast<uint4*>(e4_base + off64) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint32_t h[4], l[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float v0 = e[8 * j + 2 * t], v1 = e[8 * j + 2 * t + 1];
            const uint32_t hp = pack_bf16x2(v0, v1);
            h[t] = hp;
            l[t] = pack_bf16x2(v0 - __uint_as_float(hp << 16), v1 - __uint_as_float(hp & 0xffff0000u));
          }
          const int off = row_off + ((j ^ (r & 7)) << 4);
          *reinterpret_cast<uint4*>(hi_base + off) = make_uint4(h[0], h[1], h[2], h[3]);
          if (NPASS == 3) *reinterpret_cast<uint4*>(lo_base + off) = make_uint4(l[0], l[1], l[2], l[3]);
        }
      }
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
    uint32_t acc_par = 0;
    int rp = 0;
    long long it = 0;
    for (long long tile = blockIdx.x; tile < tile_end; tile += gridDim.x, ++it) {
      const int buf = (int)(it & 1);
      const bool live = tile < n_tiles;
      const int f = live ? (int)(tile / a.tiles_per_frame) : 0;
      const long long p = live ? (tile % a.tiles_per_frame) * TC_TM + row : a.src.P;
      mbar_wait_wd(&pe_full[buf], (uint32_t)((it >> 1) & 1), 600 + buf);
      for (int g = 0; g < 8; ++g) {
        const uint32_t d_region = tmem_base + (rp ? 256u : 0u);
        const float* bias = (g == 0) ? (fbias_s + buf * 512) : (g == 5) ? (fbias_s + buf * 512 + 256) : (tcbias_s + g * 256);
        mbar_wait_wd(&acc_full[0], acc_par, 700);
        acc_par ^= 1u;
        tc_fence_after();
#pragma unroll 1
        for (int hh = 0; hh < 2; ++hh) {
          const uint32_t taddr0 = d_region + lane_sel + (uint32_t)(hh * 128 + half * 32);
          uint32_t va[32], vb[32];
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
            mbar_arrive_remote(ed0 + (uint32_t)q * 8u);       // the leader's MMA thread counts both CTAs
          }
        }
        if (g == 5) mbar_arrive(&pe_empty[buf]);
        rp ^= 1;
      }
      // ---- G8: raw output (no activation), tf_nerf.py:283
      {
        const uint32_t d_region = tmem_base + (rp ? 256u : 0u);
        mbar_wait_wd(&acc_full[0], acc_par, 800);
        acc_par ^= 1u;
        tc_fence_after();
        if (half == 0) {
          uint32_t v[4];
          tmem_ld4(d_region + lane_sel, v);
          tmem_ld_wait();
          if (p < a.src.P) {
            float* o = a.out + ((long long)f * a.src.P + p) * a.out_ch;
            const float* bo = tcbias_s + 8 * 256;
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
  cluster_sync_all();          // no CTA leaves while its peer may still signal it or the pair's MMAs are in flight
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
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
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cur_dev);
  const long long want = (n_tiles + 1) / 2;
  const unsigned grid = (unsigned)((want < sms / 2 ? want : sms / 2) * 2);      // whole pairs
  mlp_tc2_kernel<NPASS, UVD><<<grid, TC_THREADS, T2_SMEM_BYTES, st>>>(a);
  return check_launch("mlp_tc2_kernel") ? 0 : 5;
}

int launch_mlp_tc2(const TcArgs& a, long long n_tiles, int npass, cudaStream_t st) {
  if (a.src.uv_dims == 2)
    return npass == 3 ? launch_tc2_impl<3, 2>(a, n_tiles, st) : npass == 2 ? launch_tc2_impl<2, 2>(a, n_tiles, st) : launch_tc2_impl<1, 2>(a, n_tiles, st);
  return npass == 3 ? launch_tc2_impl<3, 3>(a, n_tiles, st) : npass == 2 ? launch_tc2_impl<2, 3>(a, n_tiles, st) : launch_tc2_impl<1, 3>(a, n_tiles, st);
}

}  // namespace s2l
