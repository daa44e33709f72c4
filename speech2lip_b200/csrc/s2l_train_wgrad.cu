// Weight-gradient GEMMs of the training path on tensor cores:  dW[m, n] = sum over points p of A[p, m] * B[p, n]
// with A, B bf16 activation / gradient matrices stored point-major ([rows][channels], what the forward and the dgrad kernel
// write).  The reduction runs over the POINTS, so both operands are "MN-major" for the tensor core (the contiguous
// dimension is M resp. N, not K): TMA loads [64 channels x 64 points] boxes with the 128-byte swizzle, which is exactly the
// canonical MN-major SW128 shared-memory layout (8-point atoms of 1 KB, 64-channel blocks 8 KB apart), and the instruction
// descriptor sets the a_major / b_major transpose bits.  No transposed copy of any activation ever exists.
//
// Split-K: the work is cut into items (job, slab of 64-point chunks), s2l_train.cuh WgPlan; a persistent CTA per SM takes
// items round-robin, accumulates a full [256 x N] fp32 tile in TMEM (two M = 128 halves x N <= 256 columns = all 512
// columns) and writes it to the item's partial block; s2l_train_final.cu reduces the slabs (deterministic, no atomics).
// Column sums of A (bias gradients; per frame for the folded layers) are taken from the staged shared-memory tiles by four
// CUDA-core warps while the MMAs run.
//
// HBM-bound by construction (an L chunk brings 64 KB for 8.4 MFLOP = 128 FLOP/B, machine balance 252): the roofline of
// this kernel is the HBM one — bytes = every dPre / h row once per job that needs it.
#include <cuda.h>
#include "s2l_tc_common.cuh"
#include "s2l_train.cuh"

namespace s2l {

constexpr int WG_NSTG = 3;
constexpr int WG_A_BYTES = 4 * 8192;                   // [64 pts][256 ch] bf16 as 4 blocks of [64 pts][64 ch]
constexpr int WG_B_BYTES = 4 * 8192;
constexpr int WG_STAGE = WG_A_BYTES + WG_B_BYTES;      // 64 KB
constexpr int WG_SM_BAR = WG_NSTG * WG_STAGE;
// barriers: full[3] empty[3] acc_done acc_free
constexpr int WG_NBAR = 2 * WG_NSTG + 2;
constexpr int WG_SM_TMEMPTR = WG_SM_BAR + WG_NBAR * 8;
constexpr int WG_SMEM_BYTES = WG_SM_TMEMPTR + 16;
constexpr int WG_THREADS = 256;
static_assert(WG_SMEM_BYTES <= 232448, "shared memory budget");

struct WgArgs {
  WgPlan plan;
  long long rows_total;
  float* partials;
};

struct WgItem {
  int kind;            // 0 L, 1 M, 2 O
  int a_layer;         // dPre index (L, M) / unused (O: A = h7)
  int b_layer;         // h index (L) / unused
  int N;
  int chunk0, chunk1;  // 64-point chunks [chunk0, chunk1) in global row units of 64
  bool colsum;
  long long out_off;   // float offset of the partial block
};

__device__ __forceinline__ WgItem decode_item(const WgPlan& pl, int i) {
  WgItem it{};
  const long long chunks = (long long)pl.F * pl.chunks_per_frame;
  if (i < 7 * pl.SL) {
    const int job = i / pl.SL, slab = i % pl.SL;
    const int l = job + 1;                       // pts_linears 1..7
    it.kind = 0;
    it.a_layer = l;
    it.b_layer = l - 1;                          // l = 5: h4 (the W5[:, 256:] half; the h_skip half is the fold, M5)
    it.N = 256;
    it.chunk0 = (int)(chunks * slab / pl.SL);
    it.chunk1 = (int)(chunks * (slab + 1) / pl.SL);
    it.colsum = (l != 5);                        // dPre5's column sums come per frame from the M5 items
    it.out_off = pl.off_L(job, slab);
    return it;
  }
  i -= 7 * pl.SL;
  if (i < pl.SO) {
    it.kind = 2;
    it.N = 16;
    it.chunk0 = (int)(chunks * i / pl.SO);
    it.chunk1 = (int)(chunks * (i + 1) / pl.SO);
    it.colsum = false;
    it.out_off = pl.off_O(i);
    return it;
  }
  i -= pl.SO;
  const int which = i / (pl.F * pl.SM);
  const int f = (i / pl.SM) % pl.F, slab = i % pl.SM;
  it.kind = 1;
  it.a_layer = which == 0 ? 0 : 5;
  it.N = 64;
  it.chunk0 = f * pl.chunks_per_frame + (int)((long long)pl.chunks_per_frame * slab / pl.SM);
  it.chunk1 = f * pl.chunks_per_frame + (int)((long long)pl.chunks_per_frame * (slab + 1) / pl.SM);
  it.colsum = true;
  it.out_off = pl.off_M(which, f, slab);
  return it;
}

// 2-D tensor copy global -> shared, completion on a CTA-local mbarrier
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst_smem)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}

// MN-major operand descriptors: LBO = distance between 64-channel blocks, SBO = distance between 8-point atoms
__device__ __forceinline__ uint64_t mn_desc_sw128(uint32_t addr) {      // blocks 8 KB apart, atoms 1 KB apart
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)(8192 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint64_t mn_desc_sw32(uint32_t addr) {       // one 16-channel block, atoms of 8 points x 32 B
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)(256 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46) | (6ull << 61);
}
// kind::f16, bf16 x bf16 -> f32, BOTH operands MN-major (transpose bits 15 / 16), M = 128
__host__ __device__ constexpr uint32_t idesc_bf16_mn(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_tc_kernel(const __grid_constant__ WgArgs a, const __grid_constant__ CUtensorMap tm_dpre,
                                                                 const __grid_constant__ CUtensorMap tm_h, const __grid_constant__ CUtensorMap tm_pe,
                                                                 const __grid_constant__ CUtensorMap tm_do) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WG_SM_BAR);
  uint64_t* full = bars;
  uint64_t* empty = bars + WG_NSTG;
  uint64_t* acc_done = bars + 2 * WG_NSTG;
  uint64_t* acc_free = acc_done + 1;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem + WG_SM_TMEMPTR);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int n_items = a.plan.n_items();
  const int rows64 = (int)(a.rows_total / 64);          // dPre / h are [8][rows_total][256]: layer l starts at chunk l * rows64

  if (tid == 0) {
    for (int s = 0; s < WG_NSTG; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1 + 4);        // the MMA commit + the four column-sum warps
    }
    mbar_init(acc_done, 1);
    mbar_init(acc_free, 128);
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == 0) {
    // =============================================================== TMA producer
    if (elect_one()) {
      uint32_t stage = 0, phase = 0;
      for (int k = 0, i; (i = wg_nth_item((int)blockIdx.x, k, (int)gridDim.x, n_items)) >= 0; ++k) {
        const WgItem it = decode_item(a.plan, i);
        const uint32_t bytes = WG_A_BYTES + (it.kind == 0 ? WG_B_BYTES : it.kind == 1 ? 8192 : 64 * 32);
        for (int c = it.chunk0; c < it.chunk1; ++c) {
          mbar_wait_wd<true>(&empty[stage], phase ^ 1u, 100 + stage);
          uint8_t* sa = smem + stage * WG_STAGE;
          uint8_t* sb = sa + WG_A_BYTES;
          mbar_arrive_expect_tx(&full[stage], bytes);
          const CUtensorMap* tma = it.kind == 2 ? &tm_h : &tm_dpre;
          const int arow = ((it.kind == 2 ? 7 : it.a_layer) * rows64 + c) * 64;
#pragma unroll
          for (int b = 0; b < 4; ++b) tma_load_2d(sa + b * 8192, tma, b * 64, arow, &full[stage]);
          if (it.kind == 0) {
            const int brow = (it.b_layer * rows64 + c) * 64;
#pragma unroll
            for (int b = 0; b < 4; ++b) tma_load_2d(sb + b * 8192, &tm_h, b * 64, brow, &full[stage]);
          } else if (it.kind == 1) {
            tma_load_2d(sb, &tm_pe, 0, c * 64, &full[stage]);
          } else {
            tma_load_2d(sb, &tm_do, 0, c * 64, &full[stage]);
          }
          stage = (stage + 1 == WG_NSTG) ? 0 : stage + 1;
          phase ^= (stage == 0);
        }
      }
    }
  } else if (warp == 1) {
    // =============================================================== MMA issuer
    uint32_t stage = 0, phase = 0, free_par = 0;
    int n_done = 0;
    for (int k = 0, i; (i = wg_nth_item((int)blockIdx.x, k, (int)gridDim.x, n_items)) >= 0; ++k, ++n_done) {
      const WgItem it = decode_item(a.plan, i);
      if (n_done > 0) {                 // the previous item's accumulators have been drained
        mbar_wait_trap(acc_free, free_par);
        free_par ^= 1u;
        tc_fence_after();
      }
      const uint32_t idesc = idesc_bf16_mn(it.N);
      for (int c = it.chunk0; c < it.chunk1; ++c) {
        mbar_wait_trap(&full[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * WG_STAGE), sb = sa + WG_A_BYTES;
        if (elect_one()) {
#pragma unroll
          for (int mh = 0; mh < 2; ++mh) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {          // K = 16 points per MMA: 2 atoms of 8 points
              const uint64_t ad = mn_desc_sw128(sa + mh * 2 * 8192 + k * 2048);
              const uint64_t bd = (it.kind == 2) ? mn_desc_sw32(sb + k * 512) : mn_desc_sw128(sb + k * 2048);
              umma_ss(tmem_base + (uint32_t)(mh * 256), ad, bd, idesc, (c > it.chunk0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit(&empty[stage]);
          if (c + 1 == it.chunk1) umma_commit(acc_done);
        }
        __syncwarp();
        stage = (stage + 1 == WG_NSTG) ? 0 : stage + 1;
        phase ^= (stage == 0);
      }
      if (it.chunk0 == it.chunk1 && elect_one()) umma_commit(acc_done);      // empty slab: nothing accumulated
      __syncwarp();
    }
  } else if (warp >= 4) {
    // =============================================================== column sums of A while the MMAs run, then the drain
    const int w4 = warp - 4;                              // TMEM lane quadrant of this warp
    const int t128 = tid - 128;                           // channels 2*t128, 2*t128+1
    uint32_t stage = 0, phase = 0, done_par = 0;
    for (int k = 0, i; (i = wg_nth_item((int)blockIdx.x, k, (int)gridDim.x, n_items)) >= 0; ++k) {
      const WgItem it = decode_item(a.plan, i);
      float s0 = 0.f, s1 = 0.f;
      for (int c = it.chunk0; c < it.chunk1; ++c) {
        mbar_wait_wd(&full[stage], phase, 300 + stage);
#ifdef S2L_DBG_WG_NOCOLSUM      // experiment: how much of a chunk's time is the column-sum pass over the staged tile
        if (false) {
#else
        if (it.colsum) {
#endif
          // A tile: block b = channel / 64, row = point: 128 B rows, 16-byte chunks XOR-swizzled with (point & 7)
          const uint8_t* blk = smem + stage * WG_STAGE + (t128 >> 5) * 8192;
          const int cb = (t128 & 31) * 4;                 // byte offset of the channel pair inside the unswizzled row
#pragma unroll 8
          for (int p = 0; p < 64; ++p) {
            const uint32_t wv = *reinterpret_cast<const uint32_t*>(blk + p * 128 + ((((cb >> 4) ^ (p & 7)) << 4) | (cb & 15)));
            s0 += __uint_as_float(wv << 16);
            s1 += __uint_as_float(wv & 0xffff0000u);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        stage = (stage + 1 == WG_NSTG) ? 0 : stage + 1;
        phase ^= (stage == 0);
      }
      // ---- drain: TMEM accumulators -> the item's partial block (row m = A channel, column n = B channel)
      mbar_wait_wd(acc_done, done_par, 400);
      done_par ^= 1u;
      tc_fence_after();
      float* out = a.partials + it.out_off;
      const bool any = it.chunk1 > it.chunk0;
#pragma unroll 1
      for (int mh = 0; mh < 2; ++mh) {
        const int m = mh * 128 + w4 * 32 + lane;
        float* orow = out + (size_t)m * it.N;
        for (int n0 = 0; n0 < it.N; n0 += 16) {
          uint32_t v[16];
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                       : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                         "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                       : "r"(tmem_base + ((uint32_t)(w4 * 32) << 16) + (uint32_t)(mh * 256 + n0))
                       : "memory");
          tmem_ld_wait();
#pragma unroll
          for (int t = 0; t < 4; ++t)
            *reinterpret_cast<float4*>(orow + n0 + 4 * t) = any ? make_float4(__uint_as_float(v[4 * t]), __uint_as_float(v[4 * t + 1]),
                                                                              __uint_as_float(v[4 * t + 2]), __uint_as_float(v[4 * t + 3]))
                                                                : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      if (it.kind != 2) {
        float* cs = out + (size_t)256 * it.N;
        cs[2 * t128] = s0;
        cs[2 * t128 + 1] = s1;
      }
      tc_fence_before();
      mbar_arrive(acc_free);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

int launch_wgrad_tc(const WgPlan& plan, const TrainBufs& B, cudaStream_t st) {
  if (plan.n_items() == 0 || B.rows_total == 0) return 0;
  WgArgs a{};
  a.plan = plan;
  a.rows_total = B.rows_total;
  a.partials = B.partials;
  CUtensorMap tm_dpre, tm_h, tm_pe, tm_do;
  const unsigned long long R8 = 8ull * (unsigned long long)B.rows_total, R = (unsigned long long)B.rows_total;
  if (!encode_2d(&tm_dpre, B.dpre, 256, R8, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B) || !encode_2d(&tm_h, B.h, 256, R8, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B) ||
      !encode_2d(&tm_pe, B.pe, 64, R, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B) || !encode_2d(&tm_do, B.dout16, 16, R, 16, 64, CU_TENSOR_MAP_SWIZZLE_32B))
    return 7;
  static bool attr_set_dev[64] = {};
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  bool& attr_set = attr_set_dev[cur_dev & 63];
  if (!attr_set) {
    if (cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_BYTES) != cudaSuccess) {
      set_error("wgrad_tc: cannot opt in to %d B of shared memory: %s", WG_SMEM_BYTES, cudaGetErrorString(cudaGetLastError()));
      return 6;
    }
    attr_set = true;
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cur_dev);
  const int items = plan.n_items();
  wgrad_tc_kernel<<<items < sms ? items : sms, WG_THREADS, WG_SMEM_BYTES, st>>>(a, tm_dpre, tm_h, tm_pe, tm_do);
  return check_launch("wgrad_tc_kernel") ? 0 : 5;
}

}  // namespace s2l
