// Data-parallel training exchange ("Training (C5)": DP over frames as in the reference — DistributedSampler,
// train.py:102, DDP training.py:40 — with the gradient all-reduce of one flat bucket after backward).
// One process per GPU.  Each rank owns ONE device buffer exported to its peers with CUDA IPC; the all-reduce is a single
// kernel per step: every rank signals "my gradients are in place" into its peers' flag slots, waits for theirs, then reads
// all `world` payloads over NVLink / NVSwitch peer loads and sums them in rank order — the same order on every rank, so all
// ranks produce bit-identical sums and their weights never drift apart.  The hot-path bucket is 2.8 MB (8.8 MB with the
// UNet): a one-shot exchange (read world x bucket per rank) is latency-bound, which is why it is one launch with in-kernel
// flags instead of a reduce-scatter + all-gather pair.  The payload is double-buffered by step parity: a rank that passed
// the barrier of step e has seen every peer's signal for e, which each peer sends only after its step e-1 kernel retired —
// so nobody can still be reading the buffer this rank fills for step e+1.
#include <cstdio>
#include <cstring>

#include "s2l_common.cuh"

namespace s2l {

constexpr int kPeerMax = 16;
constexpr size_t kPeerFlagBytes = 4096;          // flag slot r at byte 128*r (written by rank r), then the two payload buffers

struct PeerArgs {
  const float* payload[kPeerMax];
  uint32_t* flags[kPeerMax];
  float* out;
  long long n4;                                  // float4 elements
  long long n;                                   // floats
  float scale;
  uint32_t epoch;
  int rank, world;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(512) allreduce_peer_kernel(PeerArgs a) {
  const int tid = threadIdx.x;
  if (blockIdx.x == 0 && tid < a.world) {
    __threadfence_system();
    st_release_sys(a.flags[tid] + a.rank * 32, a.epoch);               // "rank a.rank has its payload of this epoch in place"
  }
  if (tid < a.world) {
    const uint32_t* mine = a.flags[a.rank] + tid * 32;
    unsigned spins = 0;
    while ((int32_t)(ld_acquire_sys(mine) - a.epoch) < 0) {
      __nanosleep(64);
      if (++spins > (1u << 26)) {                                      // ~ seconds: a peer never arrived — fail loudly, do not hang the box
        printf("s2l allreduce_peer: rank %d timed out waiting for rank %d (epoch %u)\n", a.rank, tid, a.epoch);
        __trap();
      }
    }
  }
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + tid; i < a.n4; i += stride) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < a.world; ++r) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(a.payload[r]) + i);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    s.x *= a.scale; s.y *= a.scale; s.z *= a.scale; s.w *= a.scale;
    reinterpret_cast<float4*>(a.out)[i] = s;
  }
  if (blockIdx.x == 0) {
    for (long long i = a.n4 * 4 + tid; i < a.n; i += blockDim.x) {
      float s = 0.f;
      for (int r = 0; r < a.world; ++r) s += __ldcg(a.payload[r] + i);
      a.out[i] = s * a.scale;
    }
  }
}

}  // namespace s2l

using namespace s2l;

extern "C" size_t s2l_peer_buffer_bytes(int64_t n_floats) {
  if (n_floats < 0) return 0;
  const size_t payload = (((size_t)n_floats * sizeof(float)) + 255) / 256 * 256;
  return kPeerFlagBytes + 2 * payload;
}

extern "C" int32_t s2l_peer_alloc(int64_t n_floats, void** dptr, uint8_t* handle64) {
  if (!dptr || !handle64 || n_floats <= 0) { set_error("s2l_peer_alloc: bad argument"); return 1; }
  const size_t bytes = s2l_peer_buffer_bytes(n_floats);
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) { set_error("s2l_peer_alloc: cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e)); return 3; }
  cudaMemset(p, 0, bytes);
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) { cudaFree(p); set_error("s2l_peer_alloc: cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); return 5; }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(handle64, &h, 64);
  cudaDeviceSynchronize();
  *dptr = p;
  return 0;
}

extern "C" int32_t s2l_peer_open(const uint8_t* handle64, void** dptr) {
  if (!handle64 || !dptr) { set_error("s2l_peer_open: null argument"); return 1; }
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) { set_error("s2l_peer_open: cudaIpcOpenMemHandle: %s", cudaGetErrorString(e)); return 5; }
  *dptr = p;
  return 0;
}

extern "C" int32_t s2l_peer_close(void* dptr) {
  if (!dptr) return 0;
  const cudaError_t e = cudaIpcCloseMemHandle(dptr);
  if (e != cudaSuccess) { set_error("s2l_peer_close: %s", cudaGetErrorString(e)); return 5; }
  return 0;
}

extern "C" int32_t s2l_peer_free(void* dptr) {
  if (!dptr) return 0;
  const cudaError_t e = cudaFree(dptr);
  if (e != cudaSuccess) { set_error("s2l_peer_free: %s", cudaGetErrorString(e)); return 5; }
  return 0;
}

extern "C" int32_t s2l_allreduce_peer(const void* const* peer_bufs, int32_t rank, int32_t world, int64_t n_floats, float scale,
                                      uint32_t epoch, float* out, void* stream) {
  if (!peer_bufs || !out) { set_error("s2l_allreduce_peer: null argument"); return 1; }
  if (world < 1 || world > kPeerMax || rank < 0 || rank >= world || n_floats <= 0 || epoch == 0) {
    set_error("s2l_allreduce_peer: bad rank/world/size/epoch (world <= %d, epoch >= 1)", kPeerMax);
    return 2;
  }
  if (reinterpret_cast<uintptr_t>(out) & 15) { set_error("s2l_allreduce_peer: out must be 16-byte aligned"); return 2; }
  const size_t payload = (((size_t)n_floats * sizeof(float)) + 255) / 256 * 256;
  PeerArgs a{};
  for (int r = 0; r < world; ++r) {
    if (!peer_bufs[r]) { set_error("s2l_allreduce_peer: peer buffer %d is null", r); return 1; }
    uint8_t* b = reinterpret_cast<uint8_t*>(const_cast<void*>(peer_bufs[r]));
    a.flags[r] = reinterpret_cast<uint32_t*>(b);
    a.payload[r] = reinterpret_cast<const float*>(b + kPeerFlagBytes + (epoch & 1) * payload);
  }
  a.out = out; a.n = n_floats; a.n4 = n_floats / 4; a.scale = scale; a.epoch = epoch; a.rank = rank; a.world = world;
  long long blocks = (a.n4 + 511) / 512;
  if (blocks > 148) blocks = 148;            // every CTA must be resident while it spins on the flags: one per SM
  if (blocks < 1) blocks = 1;
  allreduce_peer_kernel<<<(unsigned)blocks, 512, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  return check_launch("allreduce_peer_kernel") ? 0 : 5;
}

// byte offset of the payload a rank fills for `epoch` inside its own buffer
extern "C" size_t s2l_peer_payload_offset(int64_t n_floats, uint32_t epoch) {
  const size_t payload = (((size_t)n_floats * sizeof(float)) + 255) / 256 * 256;
  return kPeerFlagBytes + (epoch & 1) * payload;
}
