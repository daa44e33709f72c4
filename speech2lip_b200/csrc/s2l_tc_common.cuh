// Shared pieces of the tensor-core MLP kernels (s2l_mlp_tc.cu: one CTA per tile stream, cta_group::1;
// s2l_mlp_tc2.cu: CTA pairs, cta_group::2): shared-memory map, tcgen05 / TMEM wrappers, operand packing helpers.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cstdio>
#include "s2l_common.cuh"
#include "s2l_points.cuh"

namespace s2l {


constexpr int TC_TM = 128;
constexpr int TC_THREADS = 512;
constexpr int NSTG = 9;
constexpr int PE_PLANE = TC_TM * 128;              // 16 KB: [128 rows][64 K] bf16, SW128
constexpr int PE_BUF = 2 * PE_PLANE;               // hi + lo
constexpr int SM_PE = 0;                           // 2 buffers
constexpr int SM_STG = SM_PE + 2 * PE_BUF;         // 65536
constexpr int kStageBytes = kGranPlane;            // one plane per ring stage
constexpr int SM_TCBIAS = SM_STG + NSTG * kStageBytes;
constexpr int SM_FBIAS = SM_TCBIAS + kNumG * 256 * 4;
constexpr int SM_BAR = SM_FBIAS + 2 * 2 * 256 * 4;
constexpr int NBAR = 2 * NSTG + 2 + 2 + 4 + 4;
constexpr int SM_TMEMPTR = SM_BAR + NBAR * 8;
constexpr int TC_SMEM_BYTES = SM_TMEMPTR + 16;

struct TcArgs {
  const uint8_t* blob;
  Layout L;
  PointSrc src;
  const float* frame_bias;   // [F,4,256]; rows 2,3 = folded bias0', bias5'
  float* out;                // [F*P, out_ch]
  int out_ch;
  int n_frames;
  long long tiles_per_frame;
  long long* dbg;            // S2L_TIMELINE builds only: event log of CTA 0 (tools/tc_timeline.py)
  // ---- fused per-pixel reduction (the reducer warp, reduce_tile below)
  int epi_mode;              // EPI_RAW: raw outputs -> out;  EPI_ENS4: 4-tap blend -> rgb;  EPI_COMPOSITE: alpha compositing -> rgb
  float* rgb;                // [F, H*W, 3]
  float4* carry;             // EPI_COMPOSITE [F*R]: (T, acc rgb) of a ray between sample chunks / for the fp32 re-evaluation
  int* next_count;           // EPI_COMPOSITE: per-frame list this launch appends to ([F] counts, zeroed by the host side) ...
  int* next_rays;            // ... [F][R] ray indices: rays still alive after a non-final chunk, or (final chunk) rays
                             //     whose last-sample density is too close to zero to trust the tensor-core sign
  float term_thr;            // non-final chunk: a ray whose transmittance fell below term_thr is finished
  float fix_thr;             // final chunk: |sigma_last| < fix_thr -> listed for the fp32 re-evaluation (0: never)
  // ---- training forward (TRAIN instantiation, bf16): what the backward kernels need, tile-major rows (row = tile * 128 + r,
  //      rows of a tile past the frame's last point hold finite garbage that the backward multiplies by zero gradients)
  __nv_bfloat16* save_h;     // [8][rows_total][256] post-ReLU activations h0..h7 exactly as the next layer's MMA consumed them
  __nv_bfloat16* save_pe;    // [rows_total][64] positional encodings (bf16), the B operand of the fold weight gradients
  uint32_t* save_mask;       // [8][tiles][8][128] ReLU masks: bit c of word (layer, tile, slice s, row r) = h[row][32 s + c] > 0
  long long rows_total;      // n_tiles * 128
  Gate gate;                 // optional device-side launch gate (s2l_points.cuh)
};


__device__ __forceinline__ long long launch_tiles(const TcArgs& a) {
  return a.src.tile_start ? (long long)a.src.tile_start[a.n_frames] : a.tiles_per_frame * a.n_frames;
}

// Cycle-stamped event log for pipeline analysis; compiled out unless -DS2L_TIMELINE.
#ifdef S2L_TIMELINE
#define TL_DECL int tl_n = 0; long long tlc_t[3] = {0, 0, 0}, tlc_acc[2] = {0, 0}
// cheap wait accounting: TLC(i) stamps a register; TLC_FLUSH logs (epilogue-wait, weight-wait) cycles of the half as codes 2xxx/3xxx
#define TLC(i) do { if ((i) == 0) tlc_t[0] = clock64(); if ((i) == 1) tlc_acc[0] += clock64() - tlc_t[0]; if ((i) == 3) tlc_t[1] = clock64(); if ((i) == 2) tlc_acc[1] += clock64() - tlc_t[1]; } while (0)
#define TLC_FLUSH(role) do { if (a.dbg && blockIdx.x == 0 && it == 2 && tl_n < 3990 && (threadIdx.x & 31) == 0) { \
    long long* _p = a.dbg + (role) * 8192; _p[1 + 2 * tl_n] = 2000; _p[2 + 2 * tl_n] = tlc_acc[0]; ++tl_n; _p[1 + 2 * tl_n] = 3000; _p[2 + 2 * tl_n] = tlc_acc[1]; ++tl_n; _p[0] = tl_n; } \
    tlc_acc[0] = tlc_acc[1] = 0; } while (0)
#define TL(role, code) do { if (a.dbg && blockIdx.x == 0 && it == 2 && tl_n < 4000 && ((role) != 0 || (threadIdx.x & 31) == 0)) { \
    long long* _p = a.dbg + (role) * 8192; _p[1 + 2 * tl_n] = (code); _p[2 + 2 * tl_n] = clock64(); ++tl_n; _p[0] = tl_n; } } while (0)
#else
#define TL_DECL do { } while (0)
#define TLC(i) do { } while (0)
#define TLC_FLUSH(role) do { } while (0)
#define TL(role, code) do { } while (0)
#endif

// ------------------------------------------------------------------ tcgen05 wrappers
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// K-major, 128-byte-swizzled shared-memory operand descriptor (8-row atoms of 1024 B):
// start>>4 | LBO=1 (unused for swizzled K-major) | SBO=1024>>4 | version=1 (sm_100) | layout=SWIZZLE_128B
__device__ __forceinline__ uint64_t sw128_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=n
__host__ __device__ constexpr uint32_t idesc_bf16(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// kind::f16 with fp16 operands / kind::f8f6f4 instruction descriptors (D = f32, K-major, M = 128, N = n)
__host__ __device__ constexpr uint32_t idesc_f16(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__host__ __device__ constexpr uint32_t idesc_f8(int n, uint32_t a_fmt, uint32_t b_fmt) {   // 0 = e4m3, 1 = e5m2
  return (1u << 4) | (a_fmt << 7) | (b_fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma8_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
      "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
      "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// 256-bit global accesses (sm_100: LDG/STG.E.256).  The training kernels move 64-byte row segments per thread at a 512-byte
// row stride: with 128-bit accesses every warp instruction is 32 HALF-filled 32-byte sectors, with 256-bit ones 32 full sectors
// in half as many instructions.
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t* r) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]),
               "r"(r[6]), "r"(r[7])
               : "memory");
}
// Each lane holds 64 B (16 words) of ITS row; rows are row_stride_bytes apart.  A plain pair of 32-byte stores per lane makes
// every warp instruction 32 separate sectors in 32 lines (ncu: the LSU data pipe of the training kernels was 75 % busy, mostly
// these).  Lane pairs swap one sector through a shuffle so that each store instruction writes the 64 contiguous bytes of ONE
// row from two adjacent lanes: half the wavefronts for the same bytes.  All 32 lanes must be active.
__device__ __forceinline__ void st_rows64_paired(void* my_row, long long row_stride_bytes, const uint32_t* o, int lane) {
  const bool odd = lane & 1;
  uint32_t r[8], s0[8], s1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = __shfl_xor_sync(0xffffffffu, odd ? o[j] : o[8 + j], 1);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s0[j] = odd ? r[j] : o[j];          // row of the even lane: sector 0 from the even lane, sector 1 (received) from the odd one
    s1[j] = odd ? o[8 + j] : r[j];      // row of the odd lane
  }
  char* even_row = reinterpret_cast<char*>(my_row) - (odd ? row_stride_bytes : 0);
  const int sec = odd ? 32 : 0;
  st_global_v8(even_row + sec, s0);
  st_global_v8(even_row + row_stride_bytes + sec, s1);
}
// ---- asynchronous tile stores: a warp stages its [32 rows x 64 B] slice in shared memory (SWIZZLE_64B pattern: the 16-byte
// chunk index of a row is XORed with bits 1..2 of the row — conflict-free for 64-byte rows) and ONE lane hands it to the TMA
// engine.  The issuing warps never wait for the memory system: with plain st.global the epilogue warps — the critical path of
// the training kernels — stalled on store back-pressure and the copy of h / dPre did not overlap the layer chain at all
// (forward: 756 us of compute + 534 us of stores, measured with the stores compiled out).
__device__ __forceinline__ void stage_rows64_sw64(uint8_t* slot, int lane, const uint32_t* o) {
  const uint32_t base = smem_u32(slot) + (uint32_t)lane * 64u;
  const uint32_t x = ((uint32_t)lane >> 1) & 3u;
#pragma unroll
  for (uint32_t c = 0; c < 4; ++c)
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(base + ((c ^ x) << 4)), "r"(o[4 * c]), "r"(o[4 * c + 1]), "r"(o[4 * c + 2]),
                 "r"(o[4 * c + 3])
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_store_2d(const void* tmap, const void* src_smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(tmap)),
               "r"(smem_u32(src_smem)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_group_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_group() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ld_global_nc_v8(const void* p, uint32_t* r) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p));
}

// ------------------------------------------------------------------ thread-block-cluster helpers
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
// relaxed flavour for events whose payload is NOT generic memory (TMEM stores already completed by tcgen05.wait::st,
// async-proxy shared-memory writes already fenced): no cluster-scope release of the thread's earlier global stores
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// 1-D bulk copy global -> the same shared-memory offset of every CTA in ctaMask, completion bytes credited to the
// mbarrier at the same offset in each destination CTA
__device__ __forceinline__ void bulk_g2s_mc(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
               : "memory");
}
// commit whose arrival is delivered to the barrier at this offset in every CTA of ctaMask
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}

// Bounded mbarrier wait: a protocol bug must surface as a trapped kernel, never as a hung GPU box.  The bound is generous
// (~30 s of SM clocks; a launch lasts ~0.1 s): a kernel slowed 100x by compute-sanitizer, or time-sliced against another
// process, must not be mistaken for a deadlock.
constexpr long long kWatchdogCycles = 60000000000ll;
// The polling loop is a real function call, NOT inlined: these kernels are ~190 KB of straight-line code against a
// ~128 KB instruction cache (ncu: 15 % of the stall samples are instruction-fetch stalls), and an inlined copy of this loop
// with its printf at each of ~40 wait sites is pure footprint.  The fast path (barrier already complete) stays inline.
template <bool BACKOFF>
__device__ __noinline__ void mbar_wait_slow(uint64_t* bar, uint32_t parity, int tag) {
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (BACKOFF) __nanosleep(128);     // roles that run ahead (producers) must not steal issue slots while they wait
    if (clock64() - t0 > kWatchdogCycles) {
      printf("s2l tc kernel: mbarrier wait timeout (tag %d, block %d, thread %d, parity %u)\n", tag, blockIdx.x,
             threadIdx.x, parity);
      __trap();
    }
  }
}
template <bool BACKOFF = false>
__device__ __forceinline__ void mbar_wait_wd(uint64_t* bar, uint32_t parity, int tag) {
#ifdef S2L_DBG_SUSPEND       // experiment: let the hardware park the warp (up to ~20 us) instead of polling
  if (mbar_try_wait_hint(bar, parity, 20000u)) return;
#endif
  if (mbar_try_wait(bar, parity)) return;
  mbar_wait_slow<BACKOFF>(bar, parity, tag);
}

// Same bounded wait without the printf (a call inside the MMA warp's loop would force every loop-carried value
// out of the uniform registers): the trap alone still turns a protocol bug into a failed launch.
__device__ __forceinline__ void mbar_wait_trap(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > kWatchdogCycles) __trap();
  }
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);     // .x (low 16 bits) = lo, .y = hi
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);                // .x (low 16 bits) = lo
  return *reinterpret_cast<uint32_t*>(&v);
}
// four floats -> four fp8 bytes, element i in byte i
__device__ __forceinline__ uint32_t pack_fp8x4(float a, float b, float c, float d, __nv_fp8_interpretation_t kind) {
  const uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, kind);
  const uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(c, d), __NV_SATFINITE, kind);
  return lo | (hi << 16);
}

// ------------------------------------------------------------------ lean operand-split helpers (epilogue / PE)
// Blackwell-only instructions keep the per-element conversion cost low: packed fp32 math (FADD2/FMUL2), the mixed
// fp32/16-bit FMA (FHFMA: x - float(h) in ONE instruction, no unpack) and 2-wide narrowing converts (F2FP).
__device__ __forceinline__ uint32_t cvt_f16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ uint32_t cvt_bf16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
// (x0 - float(h.lo), x1 - float(h.hi)) for a packed fp16 pair
__device__ __forceinline__ float2 resid_f16x2(float x0, float x1, uint32_t h2) {
  float2 r;
  asm("{\n\t.reg .b16 lo, hi, m1;\n\tmov.b32 {lo, hi}, %4;\n\tmov.b16 m1, 0xBC00;\n\t"
      "fma.rn.f32.f16 %0, lo, m1, %2;\n\tfma.rn.f32.f16 %1, hi, m1, %3;\n\t}"
      : "=f"(r.x), "=f"(r.y) : "f"(x0), "f"(x1), "r"(h2));
  return r;
}
// same for a packed bf16 pair
__device__ __forceinline__ float2 resid_bf16x2(float x0, float x1, uint32_t h2) {
  float2 r;
  asm("{\n\t.reg .b16 lo, hi, m1;\n\tmov.b32 {lo, hi}, %4;\n\tmov.b16 m1, 0xBF80;\n\t"
      "fma.rn.f32.bf16 %0, lo, m1, %2;\n\tfma.rn.f32.bf16 %1, hi, m1, %3;\n\t}"
      : "=f"(r.x), "=f"(r.y) : "f"(x0), "f"(x1), "r"(h2));
  return r;
}
__device__ __forceinline__ uint32_t cvt_e4m3x2(float lo, float hi) {        // element 0 in the low byte
  uint16_t d;
  asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ uint32_t cvt_e5m2x2_f16x2(uint32_t h2) {
  uint16_t d;
  asm("cvt.rn.satfinite.e5m2x2.f16x2 %0, %1;" : "=h"(d) : "r"(h2));
  return d;
}
__device__ __forceinline__ uint32_t hmul2_u32(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
// fp16f8 split of four non-negative-or-any fp32 values: fp16 pairs h01/h23, e5m2(fp16(x) * 2^-kScaleW) x4, e4m3((x - fp16 x) * 2^kScaleA) x4
__device__ __forceinline__ void split4_fp16f8(float x0, float x1, float x2, float x3, uint32_t& h01, uint32_t& h23, uint32_t& w5, uint32_t& w4) {
  constexpr uint32_t kDnH2 = (uint32_t)(((15 - kScaleW) << 10) | (((15 - kScaleW) << 10) << 16));   // fp16 2^-kScaleW, twice
  constexpr float kUp = (float)(1 << kScaleA);
  h01 = cvt_f16x2(x0, x1);
  h23 = cvt_f16x2(x2, x3);
  const float2 r01 = __fmul2_rn(resid_f16x2(x0, x1, h01), make_float2(kUp, kUp));
  const float2 r23 = __fmul2_rn(resid_f16x2(x2, x3, h23), make_float2(kUp, kUp));
  w4 = cvt_e4m3x2(r01.x, r01.y) | (cvt_e4m3x2(r23.x, r23.y) << 16);
  w5 = cvt_e5m2x2_f16x2(hmul2_u32(h01, kDnH2)) | (cvt_e5m2x2_f16x2(hmul2_u32(h23, kDnH2)) << 16);
}

#ifdef S2L_DBG_SATCOUNT
// Debug builds (-DS2L_DBG_SATCOUNT, tools/check_fp16f8_domain.py): hidden activations that leave the validated fp16f8 domain
// (|a| >= 4096: the scaled e4m3 residual saturates) are counted per translation unit; s2l_debug_sat_count_* read and clear.
static __device__ unsigned long long g_sat_count = 0ull;
#endif

// One 32-column slice of an accumulator quarter -> next layer's A operand words (bias, ReLU, precision split), shared
// by the single-CTA and CTA-pair kernels so both produce bit-identical operands.
//   NPASS 3: o[0..15] bf16 hi pairs, o[16..31] bf16 lo pairs;   NPASS 1: o[0..15] bf16 pairs
//   NPASS 2: o[0..15] fp16 pairs, o[16..23] e5m2(fp16(x) * 2^-kScaleW) x4, o[24..31] e4m3((x - fp16 x) * 2^kScaleA) x4
template <int NPASS>
__device__ __forceinline__ void convert_slice(const uint32_t (&v)[32], const float4* __restrict__ b4, uint32_t (&o)[32]) {
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    const float4 bb = b4[j4];
    const float2 t01 = __fadd2_rn(make_float2(__uint_as_float(v[4 * j4 + 0]), __uint_as_float(v[4 * j4 + 1])), make_float2(bb.x, bb.y));
    const float2 t23 = __fadd2_rn(make_float2(__uint_as_float(v[4 * j4 + 2]), __uint_as_float(v[4 * j4 + 3])), make_float2(bb.z, bb.w));
    const float x0 = fmaxf(t01.x, 0.f), x1 = fmaxf(t01.y, 0.f), x2 = fmaxf(t23.x, 0.f), x3 = fmaxf(t23.y, 0.f);
    if (NPASS == 2) {
#ifdef S2L_DBG_SATCOUNT
      const int nsat = (x0 >= kF8MaxAct) + (x1 >= kF8MaxAct) + (x2 >= kF8MaxAct) + (x3 >= kF8MaxAct);
      if (nsat) atomicAdd(&g_sat_count, (unsigned long long)nsat);
#endif
      split4_fp16f8(x0, x1, x2, x3, o[2 * j4], o[2 * j4 + 1], o[16 + j4], o[24 + j4]);
    } else {
      const uint32_t h0 = cvt_bf16x2(x0, x1), h1 = cvt_bf16x2(x2, x3);
      o[2 * j4] = h0;
      o[2 * j4 + 1] = h1;
      if (NPASS == 3) {
        const float2 r01 = resid_bf16x2(x0, x1, h0), r23 = resid_bf16x2(x2, x3, h1);
        o[16 + 2 * j4] = cvt_bf16x2(r01.x, r01.y);
        o[16 + 2 * j4 + 1] = cvt_bf16x2(r23.x, r23.y);
      }
    }
  }
}

// Positional encoding of one point (Embedder.__call__, tf_nerf.py:404-425: [x, sin(2^k x), cos(2^k x)]_k, zero-padded
// to 64) written as row r of the K-major A-operand image(s) of a 128-row tile: bf16 hi (+ lo) SW128 planes, or for
// NPASS == 2 the fp16 SW128 plane + e5m2 / e4m3 SW64 planes.  `valid` = the point exists (rows past the end are zeros).
template <int NPASS, int UVD>
__device__ __forceinline__ void pe_write_row(const PointSrc& src, int f, long long p, bool valid, int r, uint8_t* hi_base,
                                             uint4* gsave = nullptr) {
  float e[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) e[i] = 0.f;
#ifdef S2L_DBG_NOPE           // experiment: what the 60 sincosf per point cost (results are garbage, timing only)
  valid = false;
#endif
  if (valid) {
    float x[3];
    gen_point(src, f, p, x);
#pragma unroll
    for (int d = 0; d < UVD; ++d) e[d] = x[d];
    // rolled over the 10 frequencies (the accurate sincosf inlines to ~50 instructions; 60 copies of it were a quarter of the
    // kernel's code): e[] then lives in local memory, which this role — one tile ahead of the MMAs — can afford
#pragma unroll 1
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
  uint8_t* lo_base = hi_base + PE_PLANE;
  const int row_off = (r >> 3) * 1024 + (r & 7) * 128;
  if (NPASS == 2) {
    // fp16 main image (SW128) + e5m2(fp16(e) * 2^-kScaleW) and e4m3((e - fp16 e) * 2^kScaleA) images (SW64)
    constexpr float kDn = 1.0f / (float)(1 << kScaleW), kUp = (float)(1 << kScaleA);
    uint8_t* e5_base = lo_base;
    uint8_t* e4_base = lo_base + PE_PLANE / 2;
    const int row_off64 = (r >> 3) * 512 + (r & 7) * 64;
#pragma unroll
    for (int c = 0; c < 4; ++c) {                 // 16 K-elements per 16-byte fp8 chunk
      uint32_t w5[4], w4[4];
#pragma unroll
      for (int u = 0; u < 2; ++u) {               // two 8-element fp16 chunks
        uint32_t h[4];
        float f[8], rs[8];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float v0 = e[16 * c + 8 * u + 2 * t], v1 = e[16 * c + 8 * u + 2 * t + 1];
          const __half2 hh = __floats2half2_rn(v0, v1);
          h[t] = *reinterpret_cast<const uint32_t*>(&hh);
          const float2 back = __half22float2(hh);
          f[2 * t] = back.x; f[2 * t + 1] = back.y;
          rs[2 * t] = v0 - back.x; rs[2 * t + 1] = v1 - back.y;
        }
        const int j = 2 * c + u;
        *reinterpret_cast<uint4*>(hi_base + row_off + ((j ^ (r & 7)) << 4)) = make_uint4(h[0], h[1], h[2], h[3]);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          w5[2 * u + t] = pack_fp8x4(f[4 * t] * kDn, f[4 * t + 1] * kDn, f[4 * t + 2] * kDn, f[4 * t + 3] * kDn, __NV_E5M2);
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
    if (gsave) gsave[j] = make_uint4(h[0], h[1], h[2], h[3]);          // training forward: the row in natural order
  }
  }
}

// ------------------------------------------------------------------ fused per-pixel reductions (reducer warp)
// The output-layer epilogue hands the tile's 128 raw outputs (float4 per point) to ONE otherwise idle warp through a
// double-buffered shared-memory tile; that warp does the per-pixel reduction off the MMA / epilogue critical path:
//   EPI_ENS4      4 taps -> 1 pixel (Trainer.predict_lip_image, training.py:237-249); lane = pixel
//   EPI_COMPOSITE density2outputs (rendering.py:43-60): lane = 4 consecutive samples of one ray, segmented shuffle scan
//                 of the transmittance over the Sc/4 lanes of a ray (Sc | 128, 4 | Sc: rays never straddle tiles).
// so the raw [P,4] tensor never exists in HBM.  With sample chunks (s0, Sc < S) the (T, acc) pair of a ray is carried
// through `carry`, rays whose transmittance fell below term_thr are finished early (early ray termination: the
// remaining samples would change the pixel by < term_thr) and the survivors are appended to the next launch's list.
// On the final chunk rays whose last-sample density is within fix_thr of zero are listed for the fp32 re-evaluation
// (rendering.py:44 gives the last sample delta = 1e10: alpha_last is a step function of sign(sigma_last), so a
// tensor-core rounding error there flips the pixel; the exact kernel redoes those few rays, s2l_mlp_fp32.cu).
// The reducer shares a warp scheduler with two epilogue warps that sit on the MMA critical path, so its per-point math is the
// hardware-approximate kind (ex2.approx / rcp: a few instructions instead of ~20 for expf / IEEE division).  Their relative
// error (2^-21) is a few 1e-7 on alpha and sigmoid — three orders of magnitude inside the 1e-3 parity bar; composite_kernel
// (the unfused path, weights / depth outputs) keeps the exact functions.
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float alpha_fast(float sigma, float dist) { return 1.f - __expf(-(fmaxf(sigma, 0.f) * dist)); }

__device__ __forceinline__ void reduce_tile(const TcArgs& a, int f, long long p0, long long Pf, const float4* rawbuf, int lane) {
  const PointSrc& s = a.src;
  const long long p = p0 + 4 * lane;
  const bool valid = p < Pf;
  if (a.epi_mode == EPI_ENS4) {
    if (!valid) return;
    const long long pix = p >> 2;
    float wt[4];
    ens4_weights(s, f, (unsigned)pix, wt);             // areas 0<->3, 1<->2 swapped (training.py:243-245)
    float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float4 v = rawbuf[4 * lane + t];
      acc[0] = __fadd_rn(acc[0], __fmul_rn(v.x, wt[t]));
      acc[1] = __fadd_rn(acc[1], __fmul_rn(v.y, wt[t]));
      acc[2] = __fadd_rn(acc[2], __fmul_rn(v.z, wt[t]));
    }
    float* o = a.rgb + ((long long)f * s.H * s.W + pix) * 3;
    o[0] = acc[0]; o[1] = acc[1]; o[2] = acc[2];
    return;
  }
  // ---- EPI_COMPOSITE  (Sc is a power of two: shifts and masks, 32-bit indices — this warp shares its issue slots with
  //      two epilogue warps, every instruction here is taken from them)
  const int L = s.Sc >> 2;                 // lanes per ray: 1, 2, 4, ..., 32
  const int sub = lane & (L - 1);
  const int lg = 31 - __clz(s.Sc);
  float Tl = 1.f, A0 = 0.f, A1 = 0.f, A2 = 0.f;
  float a_last = 0.f, T_before = 1.f, sig_last = 0.f, l0 = 0.f, l1 = 0.f, l2 = 0.f;
  float Tin = 1.f, B0 = 0.f, B1 = 0.f, B2 = 0.f;
  int ray = 0;
  bool has_last = false;
  if (valid) {
    const int p32 = (int)p;
    const int slot = p32 >> lg;
    ray = s.list_rays ? s.list_rays[(long long)f * s.R + slot] : slot;
    const long long gray = (long long)f * s.R + ray;
    const int smp = s.s0 + (p32 & (s.Sc - 1));
    const float nrm = ray_norm(s.rays_d + (s.rays_shared ? (long long)ray : gray) * 3);
    const float* zr = s.z_per_ray ? s.z + gray * s.S : s.z;
    if (s.s0 > 0) {
      const float4 c = a.carry[gray];
      Tin = c.x; B0 = c.y; B1 = c.z; B2 = c.w;
    }
    has_last = (smp + 4 == s.S);
    float zc = zr[smp];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 v = rawbuf[4 * lane + i];
      float dist;
      if (smp + i + 1 < s.S) {
        const float zn = zr[smp + i + 1];
        dist = __fmul_rn(__fsub_rn(zn, zc), nrm);
        zc = zn;
      } else {
        dist = __fmul_rn(1e10f, nrm);
      }
      const float alpha = alpha_fast(v.w, dist);
      const float c0 = sigmoid_fast(v.x), c1 = sigmoid_fast(v.y), c2 = sigmoid_fast(v.z);
      if (has_last && i == 3) {
        a_last = alpha; T_before = Tl; sig_last = v.w; l0 = c0; l1 = c1; l2 = c2;
      } else {
        const float w = alpha * Tl;
        A0 = fmaf(w, c0, A0); A1 = fmaf(w, c1, A1); A2 = fmaf(w, c2, A2);
      }
      Tl *= __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f);
    }
  }
  float incl = Tl;
  for (int o = 1; o < L; o <<= 1) {
    const float up = __shfl_up_sync(0xffffffffu, incl, o);
    if (sub >= o) incl *= up;
  }
  float excl = __shfl_up_sync(0xffffffffu, incl, 1);
  if (sub == 0) excl = 1.f;
  const float C = Tin * excl;              // transmittance in front of this lane's first sample
  float S0 = C * A0, S1 = C * A1, S2 = C * A2;
  for (int o = 1; o < L; o <<= 1) {
    S0 += __shfl_xor_sync(0xffffffffu, S0, o);
    S1 += __shfl_xor_sync(0xffffffffu, S1, o);
    S2 += __shfl_xor_sync(0xffffffffu, S2, o);
  }
  if (!valid || sub != L - 1) return;
  const long long gray = (long long)f * s.R + ray;
  float* o = a.rgb + gray * 3;
  const float acc0 = B0 + S0, acc1 = B1 + S1, acc2 = B2 + S2;
  if (has_last) {
    const float T_last = C * T_before;
    const float w = T_last * a_last;
    o[0] = fmaf(w, l0, acc0); o[1] = fmaf(w, l1, acc1); o[2] = fmaf(w, l2, acc2);
    // fix_thr < 0: the model-dependent automatic threshold the pack kernels left in the blob (s2l_common.cuh META[3])
    const float thr = a.fix_thr < 0.f ? reinterpret_cast<const float*>(a.blob + a.L.off_meta)[3] : a.fix_thr;
    if (thr > 0.f && fabsf(sig_last) < thr) {
      a.carry[gray] = make_float4(T_last, acc0, acc1, acc2);
      const int slot = atomicAdd(a.next_count + f, 1);
      a.next_rays[(long long)f * s.R + slot] = ray;
    }
  } else {
    const float T_end = Tin * incl;
    if (T_end < a.term_thr) {
      o[0] = acc0; o[1] = acc1; o[2] = acc2;             // terminated: what is left of the ray weighs < term_thr
    } else {
      a.carry[gray] = make_float4(T_end, acc0, acc1, acc2);
      const int slot = atomicAdd(a.next_count + f, 1);
      a.next_rays[(long long)f * s.R + slot] = ray;
    }
  }
}

// The reducer warp's loop, shared by both tensor-core kernels (raw_full: 128 arrivals from the output-layer epilogue
// threads, raw_empty: one arrival from the reducer).
__device__ __forceinline__ void reducer_role(const TcArgs& a, long long n_tiles, long long tile_end, const float4* rawbuf,
                                             uint64_t* raw_full, uint64_t* raw_empty, int lane) {
  long long it = 0;
  int fcur = 0;
  for (long long tile = blockIdx.x; tile < tile_end; tile += gridDim.x, ++it) {
    const int buf = (int)(it & 1);
    int f; long long p0, Pf;
    tile_locate<TC_TM>(a.src, a.tiles_per_frame, n_tiles, tile, fcur, f, p0, Pf);
    // back-off while waiting: this warp idles for a whole tile time and shares its scheduler with two epilogue warps and a
    // PE producer — a tight poll loop here takes issue slots from the critical path
    mbar_wait_wd<true>(&raw_full[buf], (uint32_t)((it >> 1) & 1), 900 + buf);
#ifndef S2L_DBG_NOREDUCE      // experiment: the hand-off protocol alone (no per-pixel work; results are garbage, timing only)
    reduce_tile(a, f, p0, Pf, rawbuf + buf * TC_TM, lane);
#endif
    __syncwarp();
    if (lane == 0) mbar_arrive(&raw_empty[buf]);
  }
}

}  // namespace s2l
