// Gradient all-reduce of the data-parallel training step as the library's OWN kernel over NVLink peer memory
// ("bucketed ncclAllReduce" in the plan; reference spatialModel.py:178-181 is the step it serves).
//
// The payload is the flat bf16 gradient arena of a stream (270 MB), living in a SYMMETRIC allocation: every rank holds the
// same-sized buffer and can address all of them -- through per-peer pointers (NVLink P2P) and, where the NVSwitch fabric
// offers it, through ONE multicast address whose loads are reduced and whose stores are replicated inside the switch
// (NVLS).  Two-shot, in place, one launch: rank r owns slice r of the arena;
//   * multicast form: `multimem.ld_reduce` returns the sum over all ranks of 8 bf16 values (fp32 accumulation in the
//     switch), `multimem.st` writes the bf16 result into every rank's buffer: 2 instructions per 16 bytes, 1/W of the arena
//     per rank in each direction;
//   * peer form (no multicast): W peer loads, fp32 sum in rank order, W peer stores.
// Only the slice owner computes a slice, so all replicas receive bit-identical sums.  The caller brackets the launch with
// the symmetric-memory barrier (release/acquire at system scope): before = every rank's arena is written, after = every
// slice is stored everywhere.  Few CTAs by design (the launch runs beside the other stream's persistent layer kernels on
// the SMs va_reserve_sms leaves free): NVLink latency is covered by 1024 threads x 16 bytes x unroll in flight per CTA.
#include "va_internal.h"

#include <stdio.h>

namespace va {

namespace {

constexpr int kArMaxWorld = 16;
struct ArPeers { const void* p[kArMaxWorld]; };

__device__ __forceinline__ void acc_bf16x2(float& a, float& b, uint32_t v) {
  a += __uint_as_float(v << 16);
  b += __uint_as_float(v & 0xffff0000u);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

template <int W>
__global__ void __launch_bounds__(512) allreduce_peer_kernel(ArPeers peers, int rank, long long n16) {
  constexpr int U = W <= 2 ? 8 : (W <= 4 ? 4 : 2);          // W x U 16-byte loads in flight per thread
  const long long per = (n16 + W - 1) / W;
  const long long lo = rank * per, hi = lo + per < n16 ? lo + per : n16;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < hi; i0 += stride * U) {
    uint4 v[U][W];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i < hi) {
#pragma unroll
        for (int p = 0; p < W; ++p) v[u][p] = __ldcv(reinterpret_cast<const uint4*>(peers.p[p]) + i);   // never from a stale L1 line
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i < hi) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int p = 0; p < W; ++p) {                       // fp32 sum in rank order, rounded once
          acc_bf16x2(acc[0], acc[1], v[u][p].x); acc_bf16x2(acc[2], acc[3], v[u][p].y);
          acc_bf16x2(acc[4], acc[5], v[u][p].z); acc_bf16x2(acc[6], acc[7], v[u][p].w);
        }
        const uint4 o = make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]),
                                   pack_bf16x2(acc[6], acc[7]));
#pragma unroll
        for (int p = 0; p < W; ++p) __stcg(reinterpret_cast<uint4*>(const_cast<void*>(peers.p[p])) + i, o);
      }
    }
  }
}

// kArUnroll independent 16-byte requests per thread are in flight before the first result is used: a reduction through the
// switch takes a few microseconds, and 8 CTAs x 1024 threads x 16 B x 1 was 44 GB/s (3 ms per 135 MB slice at N = 2).
constexpr int kArUnroll = 8;

__global__ void __launch_bounds__(1024) allreduce_multicast_kernel(void* mc, int world, int rank, long long n16) {
  const long long per = (n16 + world - 1) / world;
  const long long lo = rank * per, hi = lo + per < n16 ? lo + per : n16;
  uint4* base = reinterpret_cast<uint4*>(mc);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < hi; i0 += stride * kArUnroll) {
    uint4 v[kArUnroll];
#pragma unroll
    for (int u = 0; u < kArUnroll; ++u) {
      const long long i = i0 + u * stride;
      if (i < hi)
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0, %1, %2, %3}, [%4];"
                     : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w)
                     : "l"(base + i)
                     : "memory");
    }
#pragma unroll
    for (int u = 0; u < kArUnroll; ++u) {
      const long long i = i0 + u * stride;
      if (i < hi)
        asm volatile("multimem.st.relaxed.sys.global.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(base + i), "r"(v[u].x), "r"(v[u].y),
                     "r"(v[u].z), "r"(v[u].w)
                     : "memory");
    }
  }
}

thread_local char g_err_ar[200];

}  // namespace

const char* allreduce_bf16_run(const void* const* peer_ptrs, void* multicast_ptr, int world, int rank, long long n_elems, int n_ctas,
                               cudaStream_t st) {
  if (world < 1 || world > kArMaxWorld || rank < 0 || rank >= world) return "allreduce: bad world / rank";
  if (n_elems <= 0 || (n_elems & 7)) return "allreduce: the element count must be a positive multiple of 8 (16-byte blocks)";
  if (n_ctas < 1) n_ctas = 8;
  if (n_ctas > 148) n_ctas = 148;
  const long long n16 = n_elems / 8;
  count_launch();
  if (multicast_ptr != nullptr) {
    if (reinterpret_cast<uintptr_t>(multicast_ptr) & 15) return "allreduce: multicast pointer not 16-byte aligned";
    allreduce_multicast_kernel<<<n_ctas, 1024, 0, st>>>(multicast_ptr, world, rank, n16);
  } else {
    if (!peer_ptrs) return "allreduce: no peer pointers";
    ArPeers peers;
    for (int p = 0; p < kArMaxWorld; ++p) {
      peers.p[p] = p < world ? peer_ptrs[p] : nullptr;
      if (p < world && (!peers.p[p] || (reinterpret_cast<uintptr_t>(peers.p[p]) & 15))) return "allreduce: NULL or unaligned peer pointer";
    }
    switch (world) {
      case 1: allreduce_peer_kernel<1><<<n_ctas, 512, 0, st>>>(peers, rank, n16); break;
      case 2: allreduce_peer_kernel<2><<<n_ctas, 512, 0, st>>>(peers, rank, n16); break;
      case 4: allreduce_peer_kernel<4><<<n_ctas, 512, 0, st>>>(peers, rank, n16); break;
      case 8: allreduce_peer_kernel<8><<<n_ctas, 512, 0, st>>>(peers, rank, n16); break;
      default: return "allreduce (peer form): world must be 1, 2, 4 or 8";
    }
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { snprintf(g_err_ar, sizeof(g_err_ar), "allreduce kernel launch: %s", cudaGetErrorString(e)); return g_err_ar; }
  return nullptr;
}

}  // namespace va
