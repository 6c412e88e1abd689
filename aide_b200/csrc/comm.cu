// comm.cu -- gradient all-reduce of the data-parallel AIDE step over NVLink / NVSwitch peer memory.
//
// Reference: nn.DataParallel's gradient reduction (train_files/trainchaos_proposed_30cases1labeled.py:188-189 wraps both
// networks).  torch.distributed (NCCL) stays the plumbing -- rendezvous, broadcast of the initial weights,
// the small all_gather of the global selection -- but the one heavy collective of the step, the sum of the two flat
// gradient buffers (2 x 107 MB fp32), runs in this kernel:
//
//   * every rank maps the gradient buffers and a small flag pad of all its peers (CUDA IPC handles exchanged once);
//   * two-shot all-reduce: rank r sums slice r of the range over all ranks (peer loads, fixed rank order 0..N-1, so the
//     result is bit-identical everywhere and run to run), then copies the other ranks' reduced slices from their owners;
//   * no shared memory, 256 threads of <= 128 registers: its CTAs are co-resident with the persistent tensor-core kernels
//     (which leave 1 KB of an SM's shared memory free for exactly this reason), so the transfer rides under the backward
//     pass of the other buckets instead of taking SMs away from it -- NCCL's CTAs cannot share an SM with a 226 KB
//     conv CTA, which cost ~1 ms per step at 2 GPUs;
//   * cross-GPU barriers are per CTA (block b of every rank works on chunk b of every slice): one release-store of the
//     call's epoch into each peer's pad, one acquire-spin on the own pad.  A peer that never arrives trips a 4 s timeout
//     and traps instead of hanging the GPU.
#include "common.cuh"

#include <cstdint>
#include <cstdio>
#include <cstring>

namespace aide {
namespace {

constexpr int kMaxPeers = 8;
constexpr int kCommThreads = 256;                          // x <= 128 registers: fits beside a 192-thread conv CTA
constexpr int kCommMaxBlocks = 64;
constexpr int kPhases = 3;                                  // start, slices reduced, everybody done reading

struct P2PParams {
  float* buf[kMaxPeers];                                    // the same buffer on every rank (own entry: local pointer)
  unsigned int* pad[kMaxPeers];                             // flag pads: [channel][phase][rank][block]
  unsigned int* epoch;                                      // this rank, this channel: number of completed calls
  unsigned int* ticket;                                     // this rank, this channel: blocks finished (left at zero)
  unsigned long long lo4, n4;                               // range of the buffer in float4 units
  int rank, world, pad_off;                                 // pad_off: first flag of this channel
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// peer data: strong (L1-bypassing) loads -- the same addresses carry new values every step
__device__ __forceinline__ float4 ld_peer(const float4* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

// all ranks' block `blockIdx.x` meet: thread t < world signals rank t and waits for rank t
__device__ __forceinline__ void peer_barrier(const P2PParams& p, int phase, unsigned int e) {
  __syncthreads();
  if ((int)threadIdx.x < p.world) {
    const int t = threadIdx.x;
    __threadfence_system();
    const int nb = gridDim.x;
    st_release_sys(p.pad[t] + p.pad_off + (phase * p.world + p.rank) * nb + blockIdx.x, e);
    const unsigned int* mine = p.pad[p.rank] + p.pad_off + (phase * p.world + t) * nb + blockIdx.x;
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(mine) - e) < 0) {
      if (clock64() - t0 > 8000000000LL) {                  // ~4 s at 2 GHz: a rank died or the calls are mismatched
        printf("aide_allreduce_p2p: rank %d block %d timed out waiting for rank %d (phase %d, epoch %u)\n", p.rank,
               (int)blockIdx.x, t, phase, e);
        __trap();
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
}

// WMAX: compile-time bound of the world size, U: float4 positions per thread and iteration -- WMAX * U = 16 independent
// 128-bit peer loads in flight per thread whatever the world size (NVLink round trips are ~2 us: bandwidth = bytes in flight)
template <int WMAX, int U>
__global__ void __launch_bounds__(kCommThreads, 2) allreduce_p2p_kernel(const P2PParams p) {
  const unsigned int e = *reinterpret_cast<volatile unsigned int*>(p.epoch) + 1u;
  const int nb = gridDim.x, W = p.world;
  // slice s of the range: float4 [s * per_slice, ...); chunk b of a slice: [b * per_chunk, ...)
  const unsigned long long per_slice = (p.n4 + W - 1) / W;
  const unsigned long long per_chunk = (per_slice + nb - 1) / nb;
  auto chunk_range = [&](int s, unsigned long long& a, unsigned long long& b) {
    const unsigned long long s0 = (unsigned long long)s * per_slice;
    const unsigned long long s1 = s0 + per_slice < p.n4 ? s0 + per_slice : p.n4;
    a = s0 + (unsigned long long)blockIdx.x * per_chunk;
    b = a + per_chunk < s1 ? a + per_chunk : s1;
    if (a > b) a = b;
  };
  peer_barrier(p, 0, e);                                    // every rank's gradients of this range are final
  {
    unsigned long long a, b;
    chunk_range(p.rank, a, b);
    float4* mine = reinterpret_cast<float4*>(p.buf[p.rank]) + p.lo4;
    for (unsigned long long i = a + threadIdx.x; i < b; i += (unsigned long long)U * kCommThreads) {
      float4 v[U][WMAX];
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int r = 0; r < WMAX; ++r)
          v[u][r] = (r < W && i + u * kCommThreads < b)
                        ? ld_peer(reinterpret_cast<const float4*>(p.buf[r]) + p.lo4 + i + u * kCommThreads)
                        : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float4 s = v[u][0];                                 // rank order 0, 1, ..., W-1: the same bits on every rank
#pragma unroll
        for (int r = 1; r < WMAX; ++r)
          if (r < W) { s.x += v[u][r].x; s.y += v[u][r].y; s.z += v[u][r].z; s.w += v[u][r].w; }
        if (i + u * kCommThreads < b) mine[i + u * kCommThreads] = s;
      }
    }
  }
  peer_barrier(p, 1, e);                                    // chunk b of every slice is reduced at its owner
  {
    float4* mine = reinterpret_cast<float4*>(p.buf[p.rank]) + p.lo4;
    for (int d = 1; d < W; ++d) {
      const int s = (p.rank + d) % W;                       // start at the next rank: spread the load over the links
      unsigned long long a, b;
      chunk_range(s, a, b);
      const float4* src = reinterpret_cast<const float4*>(p.buf[s]) + p.lo4;
      for (unsigned long long i = a + threadIdx.x; i < b; i += 8 * kCommThreads) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (i + u * kCommThreads < b) v[u] = ld_peer(src + i + u * kCommThreads);
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (i + u * kCommThreads < b) mine[i + u * kCommThreads] = v[u];
      }
    }
  }
  peer_barrier(p, 2, e);                                    // nobody still reads this rank's buffer: it may be rewritten
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(p.ticket, 1u) == gridDim.x - 1u) {        // last block of this rank: the call is complete
      *p.ticket = 0u;
      *reinterpret_cast<volatile unsigned int*>(p.epoch) = e;
    }
  }
}

}  // namespace
}  // namespace aide

using namespace aide;

// ---- peer memory plumbing: plain cudaMalloc allocations exported / imported as CUDA IPC handles (64 bytes)
extern "C" int aide_comm_alloc(size_t bytes, void** ptr, unsigned char* handle /*[64]*/) {
  AIDE_REQUIRE(ptr && handle && bytes > 0, "comm_alloc: bad arguments");
  AIDE_CUDA(cudaMalloc(ptr, bytes));
  AIDE_CUDA(cudaMemset(*ptr, 0, bytes));
  cudaIpcMemHandle_t h;
  AIDE_CUDA(cudaIpcGetMemHandle(&h, *ptr));
  static_assert(sizeof(h) == 64, "CUDA IPC handle size");
  memcpy(handle, &h, 64);
  return 0;
}

extern "C" int aide_comm_open(const unsigned char* handle /*[64]*/, void** ptr) {
  AIDE_REQUIRE(handle && ptr, "comm_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  AIDE_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

extern "C" int aide_comm_close(void* ptr) {
  if (ptr) AIDE_CUDA(cudaIpcCloseMemHandle(ptr));
  return 0;
}

extern "C" int aide_comm_free(void* ptr) {
  if (ptr) AIDE_CUDA(cudaFree(ptr));
  return 0;
}

extern "C" int aide_comm_pad_words(int world, int channels) { return channels * (kPhases * world * kCommMaxBlocks + 2); }

// bufs / pads: `world` pointers each (entry `rank` = this rank's own allocation).  The range [lo, lo + count) floats must be
// 4-float aligned at both ends.  channel: calls that may be in flight at the same time need different channels; all ranks
// must issue the calls of a channel in the same order with the same (lo, count, blocks).
extern "C" int aide_allreduce_p2p(float* const* bufs, unsigned int* const* pads, int rank, int world, int channel,
                                  int channels, size_t lo, size_t count, int blocks, void* stream) {
  AIDE_REQUIRE(bufs && pads && world >= 2 && world <= kMaxPeers && rank >= 0 && rank < world, "allreduce_p2p: bad group");
  AIDE_REQUIRE(channel >= 0 && channel < channels && blocks >= 1 && blocks <= kCommMaxBlocks, "allreduce_p2p: bad channel / blocks");
  AIDE_REQUIRE(lo % 4 == 0 && count % 4 == 0, "allreduce_p2p: range must be a multiple of 4 floats");
  if (count == 0) return 0;
  P2PParams p{};
  for (int r = 0; r < world; ++r) {
    AIDE_REQUIRE(bufs[r] && pads[r], "allreduce_p2p: null peer pointer");
    p.buf[r] = bufs[r];
    p.pad[r] = pads[r];
  }
  p.rank = rank;
  p.world = world;
  p.lo4 = lo / 4;
  p.n4 = count / 4;
  const int per_channel = kPhases * world * kCommMaxBlocks + 2;
  p.pad_off = channel * per_channel;
  p.epoch = pads[rank] + p.pad_off + kPhases * world * kCommMaxBlocks;
  p.ticket = p.epoch + 1;
  if (world <= 2) allreduce_p2p_kernel<2, 8><<<blocks, kCommThreads, 0, as_stream(stream)>>>(p);
  else if (world <= 4) allreduce_p2p_kernel<4, 4><<<blocks, kCommThreads, 0, as_stream(stream)>>>(p);
  else allreduce_p2p_kernel<8, 2><<<blocks, kCommThreads, 0, as_stream(stream)>>>(p);
  AIDE_CHECK_LAUNCH();
  return 0;
}
