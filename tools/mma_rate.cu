// Microbenchmark: sustained issue rate of tcgen05.mma for the operand forms / kinds the fused MLP kernel uses.
// One elected thread per CTA issues n_iter x 8 MMAs back to back into one accumulator, commits, waits; the
// result is SM cycles per MMA (clock64), per variant, for 1 CTA and for one CTA per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/mma_rate tools/mma_rate.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "../speech2lip_b200/csrc/s2l_tc_common.cuh"

using namespace s2l;

constexpr int SM_A = 0;                 // 32 KB: A operand images (SS forms)
constexpr int SM_B = 32768;             // 128 KB: B operand images
constexpr int SM_MISC = SM_B + 131072;  // barrier + tmem ptr
constexpr int SMEM_BYTES = SM_MISC + 64;

__device__ __forceinline__ uint64_t dsc(uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | lo; }

__global__ void __launch_bounds__(128, 1) rate_kernel(int variant, int n_iter, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + SM_MISC);
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem + SM_MISC + 16);
  const int tid = threadIdx.x, warp = tid >> 5;
  // operand bytes: 0x38 = 1.0 (e4m3) / 0.5 (e5m2 0x38) ; as fp16 pairs 0x3838 = 0.527; all finite
  for (int i = tid; i < (SM_MISC) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x38343038u + (uint32_t)(i & 3);
  if (tid == 0) {
    for (int i = 0; i < 6; ++i) mbar_init(bar + i, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  {  // finite A data in TMEM columns 256..511 (all 128 lanes)
    uint32_t v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = 0x38343038u + (uint32_t)i;
    const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;
    for (int c = 0; c < 256; c += 32) tmem_st32(tmem_base + lane_sel + 256u + (uint32_t)c, v);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 1 && elect_one()) {
    constexpr uint32_t kHi128 = 0x40004040u;   // SW128, SBO 1024 B
    constexpr uint32_t kHi64 = 0x80004020u;    // SW64, SBO 512 B
    const uint32_t a_s = ((smem_u32(smem + SM_A) >> 4) & 0x3FFFu) | 0x10000u;
    const uint32_t b_s = ((smem_u32(smem + SM_B) >> 4) & 0x3FFFu) | 0x10000u;
    const uint32_t d0 = tmem_base, a_t = tmem_base + 256u;
    const uint32_t i16_128 = idesc_f16(128), i16_256 = idesc_f16(256), ib_128 = idesc_bf16(128);
    const uint32_t rw128 = idesc_f8(128, 0u, 1u), wr128 = idesc_f8(128, 1u, 0u), ee128 = idesc_f8(128, 0u, 0u);
    const uint32_t rw256 = idesc_f8(256, 0u, 1u), wr256 = idesc_f8(256, 1u, 0u);
    const uint32_t i16_64 = idesc_f16(64), rw64 = idesc_f8(64, 0u, 1u);
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < n_iter; ++it) {
      const uint32_t bo = (uint32_t)(it & 3) * 1024u;      // walk 4 x 16 KB B planes (in 16-byte units)
      switch (variant) {
        case 0:   // bf16 TS N=128
#pragma unroll
          for (int s = 0; s < 8; ++s) umma_ts(d0, a_t + (uint32_t)(s * 8), dsc(kHi128, b_s + bo + 2 * (s & 3)), ib_128, 1u);
          break;
        case 1:   // fp16 SS N=128
#pragma unroll
          for (int s = 0; s < 8; ++s) umma_ss(d0, dsc(kHi128, a_s + 2 * (s & 3)), dsc(kHi128, b_s + bo + 2 * (s & 3)), i16_128, 1u);
          break;
        case 2:   // fp16 TS N=256
#pragma unroll
          for (int s = 0; s < 8; ++s) umma_ts(d0, a_t + (uint32_t)(s * 8), dsc(kHi128, b_s + bo + 2 * (s & 3)), i16_256, 1u);
          break;
        case 3:   // fp8 TS N=128, e4m3 x e5m2, B SW64
#pragma unroll
          for (int s = 0; s < 8; ++s) umma8_ts(d0, a_t + (uint32_t)(s * 8), dsc(kHi64, b_s + bo + 2 * (s & 1)), rw128, 1u);
          break;
        case 4:   // fp8 TS N=128, formats alternate every MMA
#pragma unroll
          for (int s = 0; s < 8; ++s) umma8_ts(d0, a_t + (uint32_t)(s * 8), dsc(kHi64, b_s + bo + 2 * (s & 1)), (s & 1) ? wr128 : rw128, 1u);
          break;
        case 5:   // fp8 SS N=128
#pragma unroll
          for (int s = 0; s < 8; ++s) umma8_ss(d0, dsc(kHi64, a_s + 2 * (s & 1)), dsc(kHi64, b_s + bo + 2 * (s & 1)), rw128, 1u);
          break;
        case 6:   // fp8 TS N=256
#pragma unroll
          for (int s = 0; s < 8; ++s) umma8_ts(d0, a_t + (uint32_t)(s * 8), dsc(kHi64, b_s + bo + 2 * (s & 1)), rw256, 1u);
          break;
        case 7:   // the kernel's K-chunk pattern: 4 fp16 + 2 fp8 (rw) + 2 fp8 (wr)
#pragma unroll
          for (int s = 0; s < 4; ++s) umma_ts(d0, a_t + (uint32_t)((s >> 1) * 32 + (s & 1) * 8), dsc(kHi128, b_s + bo + 2 * s), i16_128, 1u);
#pragma unroll
          for (int t = 0; t < 2; ++t) umma8_ts(d0, a_t + (uint32_t)(t * 32 + 24), dsc(kHi64, b_s + bo + 1024u + 2 * t), rw128, 1u);
#pragma unroll
          for (int t = 0; t < 2; ++t) umma8_ts(d0, a_t + (uint32_t)(t * 32 + 16), dsc(kHi64, b_s + bo + 1536u + 2 * t), wr128, 1u);
          break;
        case 8:   // fp8 TS N=128, B in a SW128 image (128 K per row)
#pragma unroll
          for (int s = 0; s < 8; ++s) umma8_ts(d0, a_t + (uint32_t)(s * 8), dsc(kHi128, b_s + bo + 2 * (s & 3)), rw128, 1u);
          break;
        case 9:   // fp16 TS N=128, alternating accumulator halves
#pragma unroll
          for (int s = 0; s < 8; ++s) umma_ts(d0 + (uint32_t)((s & 1) * 128), a_t + (uint32_t)(s * 8), dsc(kHi128, b_s + bo + 2 * (s & 3)), i16_128, 1u);
          break;
        case 10:  // fp8 TS N=128, e4m3 x e4m3
#pragma unroll
          for (int s = 0; s < 8; ++s) umma8_ts(d0, a_t + (uint32_t)(s * 8), dsc(kHi64, b_s + bo + 2 * (s & 1)), ee128, 1u);
          break;
        case 11:  // fp16 TS N=128 (same as 0 with fp16)
#pragma unroll
          for (int s = 0; s < 8; ++s) umma_ts(d0, a_t + (uint32_t)(s * 8), dsc(kHi128, b_s + bo + 2 * (s & 3)), i16_128, 1u);
          break;
        case 12:  // fp8 TS N=128, alternating accumulator halves
#pragma unroll
          for (int s = 0; s < 8; ++s) umma8_ts(d0 + (uint32_t)((s & 1) * 128), a_t + (uint32_t)(s * 8), dsc(kHi64, b_s + bo + 2 * (s & 1)), rw128, 1u);
          break;
        case 13:  // fp16 TS N=64
#pragma unroll
          for (int s = 0; s < 8; ++s) umma_ts(d0, a_t + (uint32_t)(s * 8), dsc(kHi128, b_s + bo + 2 * (s & 3)), i16_64, 1u);
          break;
        case 14:  // fp8 TS N=64
#pragma unroll
          for (int s = 0; s < 8; ++s) umma8_ts(d0, a_t + (uint32_t)(s * 8), dsc(kHi64, b_s + bo + 2 * (s & 1)), rw64, 1u);
          break;
        case 15:  // grouped pattern: 8 fp16 then (next iteration) 8 fp8 -> kind switch every 8 MMAs
          if (it & 1) {
#pragma unroll
            for (int s = 0; s < 8; ++s) umma8_ts(d0, a_t + (uint32_t)(s * 8), dsc(kHi64, b_s + bo + 2 * (s & 1)), (s & 4) ? wr128 : rw128, 1u);
          } else {
#pragma unroll
            for (int s = 0; s < 8; ++s) umma_ts(d0, a_t + (uint32_t)(s * 8), dsc(kHi128, b_s + bo + 2 * (s & 3)), i16_128, 1u);
          }
          break;
        case 16:  // fp8 TS N=256 accumulating over the full region, kernel-like pattern at N=256: 4 fp16 + 4 fp8
#pragma unroll
          for (int s = 0; s < 4; ++s) umma_ts(d0, a_t + (uint32_t)((s >> 1) * 32 + (s & 1) * 8), dsc(kHi128, b_s + bo + 2 * s), i16_256, 1u);
#pragma unroll
          for (int t = 0; t < 4; ++t) umma8_ts(d0, a_t + (uint32_t)((t & 1) * 32 + 24 - (t >> 1) * 8), dsc(kHi64, b_s + bo + 2048u + 2 * (t & 1)), (t >> 1) ? wr256 : rw256, 1u);
          break;
        case 17:  // 4 fp16 TS N=128 + commit
#pragma unroll
          for (int s = 0; s < 8; ++s) { umma_ts(d0, a_t + (uint32_t)(s * 8), dsc(kHi128, b_s + bo + 2 * (s & 3)), i16_128, 1u); if ((s & 3) == 3) umma_commit(bar + 1 + (s >> 2)); }
          break;
        case 18:  // 2 fp16 TS N=128 + commit
#pragma unroll
          for (int s = 0; s < 8; ++s) { umma_ts(d0, a_t + (uint32_t)(s * 8), dsc(kHi128, b_s + bo + 2 * (s & 3)), i16_128, 1u); if ((s & 1) == 1) umma_commit(bar + 1 + (s >> 1)); }
          break;
        case 19:  // 1 fp16 TS N=128 + commit
#pragma unroll
          for (int s = 0; s < 8; ++s) { umma_ts(d0, a_t + (uint32_t)(s * 8), dsc(kHi128, b_s + bo + 2 * (s & 3)), i16_128, 1u); umma_commit(bar + 1 + (s & 3)); }
          break;
        case 20:  // kernel chunk pattern with its commits: 4 fp16 + commit + 2+2 fp8 + commit
#pragma unroll
          for (int s = 0; s < 4; ++s) umma_ts(d0, a_t + (uint32_t)((s >> 1) * 32 + (s & 1) * 8), dsc(kHi128, b_s + bo + 2 * s), i16_128, 1u);
          umma_commit(bar + 1);
#pragma unroll
          for (int t = 0; t < 2; ++t) umma8_ts(d0, a_t + (uint32_t)(t * 32 + 24), dsc(kHi64, b_s + bo + 1024u + 2 * t), rw128, 1u);
#pragma unroll
          for (int t = 0; t < 2; ++t) umma8_ts(d0, a_t + (uint32_t)(t * 32 + 16), dsc(kHi64, b_s + bo + 1536u + 2 * t), wr128, 1u);
          umma_commit(bar + 2);
          break;
        case 21:  // fp16 TS N=128, first MMA of every 4 overwrites (accumulate = 0)
#pragma unroll
          for (int s = 0; s < 8; ++s) umma_ts(d0, a_t + (uint32_t)(s * 8), dsc(kHi128, b_s + bo + 2 * (s & 3)), i16_128, (s & 3) ? 1u : 0u);
          break;
        case 22:  // fp16 TS N=128 with the A operand in the same 256-column region as D (other half)
#pragma unroll
          for (int s = 0; s < 8; ++s) umma_ts(d0, d0 + 128u + (uint32_t)(s * 8), dsc(kHi128, b_s + bo + 2 * (s & 3)), i16_128, 1u);
          break;
        default: break;
      }
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// Queue-depth probe: issue n MMAs back to back (fp16 TS N=128), stamp the clock when the LAST ISSUE returns and when
// the commit lands: the issue time stays ~flat per MMA until the tensor pipe's queue is full.
__global__ void __launch_bounds__(128, 1) depth_kernel(int n, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + SM_MISC);
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem + SM_MISC + 16);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (SM_MISC) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x38343038u;
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  if (warp == 1 && elect_one()) {
    const uint32_t b_s = ((smem_u32(smem + SM_B) >> 4) & 0x3FFFu) | 0x10000u;
    const uint32_t idesc = idesc_f16(128);
    const long long t0 = clock64();
    if (n == 0) {   // best case: 12 MMAs fully unrolled with compile-time operand offsets
#pragma unroll
      for (int i = 0; i < 12; ++i) umma_ts(tmem_base, tmem_base + 256u + (uint32_t)((i & 7) * 8), dsc(0x40004040u, b_s + 2 * (i & 3) + 1024 * (i >> 2)), idesc, 1u);
    } else {
#pragma unroll 1
    for (int i = 0; i < n; ++i) umma_ts(tmem_base, tmem_base + 256u + (uint32_t)((i & 7) * 8), dsc(0x40004040u, b_s + 2 * (i & 3)), idesc, 1u);
    }
    const long long t1 = clock64();
    umma_commit(bar);
    mbar_wait(bar, 0);
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---------------------------------------------------------------- cta_group::2 (CTA pair) issue rate
__device__ __forceinline__ void umma2_ts_f16(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
               "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma2_ss_f16(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
               "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma2_ts_f8(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f8f6f4 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
               "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__host__ __device__ constexpr uint32_t idesc_pair(uint32_t a_fmt, uint32_t b_fmt, int n) {
  return (1u << 4) | (a_fmt << 7) | (b_fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) pair_kernel(int variant, int n_iter, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + SM_MISC);
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem + SM_MISC + 16);
  const int tid = threadIdx.x, warp = tid >> 5;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int i = tid; i < (SM_MISC) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x38343038u + (uint32_t)(i & 3);
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  {
    uint32_t v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = 0x38343038u + (uint32_t)i;
    const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;
    for (int c = 0; c < 256; c += 32) tmem_st32(tmem_base + lane_sel + 256u + (uint32_t)c, v);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  tc_fence_after();
  if (rank == 0 && warp == 1 && elect_one()) {
    constexpr uint32_t kHi128 = 0x40004040u, kHi64 = 0x80004020u;
    const uint32_t a_s = ((smem_u32(smem + SM_A) >> 4) & 0x3FFFu) | 0x10000u;
    const uint32_t b_s = ((smem_u32(smem + SM_B) >> 4) & 0x3FFFu) | 0x10000u;
    const uint32_t d0 = tmem_base, a_t = tmem_base + 256u;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < n_iter; ++it) {
      const uint32_t bo = (uint32_t)(it & 3) * 1024u;
      switch (variant) {
        case 0:   // fp16 TS M=256 N=128 (64 B rows per CTA)
#pragma unroll
          for (int s = 0; s < 8; ++s) umma2_ts_f16(d0, a_t + (uint32_t)(s * 8), dsc(kHi128, b_s + bo + 2 * (s & 3)), idesc_pair(0, 0, 128), 1u);
          break;
        case 1:   // fp16 TS M=256 N=256
#pragma unroll
          for (int s = 0; s < 8; ++s) umma2_ts_f16(d0, a_t + (uint32_t)(s * 8), dsc(kHi128, b_s + bo + 2 * (s & 3)), idesc_pair(0, 0, 256), 1u);
          break;
        case 2:   // fp8 TS M=256 N=128
#pragma unroll
          for (int s = 0; s < 8; ++s) umma2_ts_f8(d0, a_t + (uint32_t)(s * 8), dsc(kHi64, b_s + bo + 2 * (s & 1)), idesc_pair(0, 1, 128), 1u);
          break;
        case 3:   // fp16 SS M=256 N=128
#pragma unroll
          for (int s = 0; s < 8; ++s) umma2_ss_f16(d0, dsc(kHi128, a_s + 2 * (s & 3)), dsc(kHi128, b_s + bo + 2 * (s & 3)), idesc_pair(0, 0, 128), 1u);
          break;
        case 4:   // fp16 TS M=256 N=64
#pragma unroll
          for (int s = 0; s < 8; ++s) umma2_ts_f16(d0, a_t + (uint32_t)(s * 8), dsc(kHi128, b_s + bo + 2 * (s & 3)), idesc_pair(0, 0, 64), 1u);
          break;
        case 5:   // fp16 TS M=256 N=128 alternating accumulator halves
#pragma unroll
          for (int s = 0; s < 8; ++s) umma2_ts_f16(d0 + (uint32_t)((s & 1) * 128), a_t + (uint32_t)(s * 8), dsc(kHi128, b_s + bo + 2 * (s & 3)), idesc_pair(0, 0, 128), 1u);
          break;
        default: break;
      }
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)1) : "memory");
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    out[blockIdx.x / 2] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

int main(int argc, char** argv) {
  const int n_iter = argc > 1 ? atoi(argv[1]) : 512;
  const char* names[] = {"bf16 TS N128", "fp16 SS N128", "fp16 TS N256", "fp8 TS N128 e4m3xe5m2 SW64", "fp8 TS N128 fmt-alternating",
                         "fp8 SS N128", "fp8 TS N256", "kernel chunk pattern 4xfp16+2+2 fp8", "fp8 TS N128 B SW128", "fp16 TS N128 alt D halves",
                         "fp8 TS N128 e4m3xe4m3", "fp16 TS N128", "fp8 TS N128 alt D halves", "fp16 TS N64", "fp8 TS N64",
                         "8 fp16 / 8 fp8 alternating groups", "kernel chunk pattern at N256 (4 fp16 + 4 fp8)",
                         "4 fp16 TS + commit", "2 fp16 TS + commit", "1 fp16 TS + commit", "kernel chunk pattern incl. commits", "fp16 TS acc=0 every 4th", "fp16 TS A next to D"};
  const int nvar = 23;
  const int vfirst = argc > 2 ? atoi(argv[2]) : 0;
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  std::vector<long long> h(148);
  if (argc > 3 && argv[3][0] == 'p') {
    const char* pn[] = {"pair fp16 TS M256 N128", "pair fp16 TS M256 N256", "pair fp8 TS M256 N128", "pair fp16 SS M256 N128", "pair fp16 TS M256 N64", "pair fp16 TS N128 alt D halves"};
    cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    for (int grid : {2, 148}) {
      for (int v = 0; v < 6; ++v) {
        for (int rep = 0; rep < 2; ++rep) {
          pair_kernel<<<grid, 128, SMEM_BYTES>>>(v, n_iter, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("pair variant %d: %s\n", v, cudaGetErrorString(e)); return 1; }
        }
        cudaMemcpy(h.data(), d, (grid / 2) * sizeof(long long), cudaMemcpyDeviceToHost);
        std::sort(h.begin(), h.begin() + grid / 2);
        const double per = 1.0 / (8.0 * n_iter);
        printf("grid %3d  p%-2d %-34s cycles/MMA min %.1f med %.1f max %.1f\n", grid, v, pn[v], h[0] * per, h[grid / 4] * per, h[grid / 2 - 1] * per);
      }
    }
    return 0;
  }
  if (argc > 3) {
    cudaFuncSetAttribute(depth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    for (int n : {0, 1, 2, 3, 4, 6, 8, 10, 12, 16, 20, 24, 32, 48, 64}) {
      for (int rep = 0; rep < 2; ++rep) { depth_kernel<<<1, 128, SMEM_BYTES>>>(n, d); cudaDeviceSynchronize(); }
      cudaMemcpy(h.data(), d, 2 * sizeof(long long), cudaMemcpyDeviceToHost);
      printf("n %2d  issue-done %5lld cycles  complete %5lld cycles\n", n, h[0], h[1]);
    }
    return 0;
  }
  for (int grid : {1, 148}) {
    for (int v = vfirst; v < nvar; ++v) {
      for (int rep = 0; rep < 2; ++rep) {
        rate_kernel<<<grid, 128, SMEM_BYTES>>>(v, n_iter, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("variant %d: %s\n", v, cudaGetErrorString(e)); return 1; }
      }
      cudaMemcpy(h.data(), d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
      std::sort(h.begin(), h.begin() + grid);
      const double per = 1.0 / (8.0 * n_iter);
      printf("grid %3d  v%-2d %-44s cycles/MMA min %.1f med %.1f max %.1f\n", grid, v, names[v], h[0] * per, h[grid / 2] * per, h[grid - 1] * per);
    }
  }
  return 0;
}
