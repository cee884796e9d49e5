// Microbenchmark: issue rate of tcgen05.mma (SS mode, bf16, M=128) as a function of N, with 1 or 2 accumulators and
// with distinct / identical operand addresses.  One CTA per SM, no TMA, operands are whatever is in smem (zeros).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "../../video_analytics_b200/csrc/va_ptx.cuh"
using namespace va;

template <int BN, int ROWB>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(long long* out, int iters, int nacc, int spread, int kstep) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, slot, 0);
  if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(128, BN);
    const uint32_t base = smem_u32(smem);
    const uint64_t da0 = make_smem_desc<ROWB>(base);
    const uint64_t db0 = make_smem_desc<ROWB>(base + 64 * 1024);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const uint32_t which = (nacc == 2) ? (j & 1) : 0;
          // spread: each MMA reads a different 16 KB operand block (like walking pipeline stages); kstep: +32 B
          const uint32_t off = (spread ? ((j & 3) * 16384) >> 4 : 0) + (kstep ? 2 * (j & 3) : 0);
          umma_bf16(tmem + which * BN, da0 + off, db0 + off, idesc, 1u);
        }
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0, 1);
    long long t1 = clock64();
    if (threadIdx.x == 32 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int BN, int ROWB>
void run(long long* d_out, int nacc, int spread, int kstep, int grid) {
  const int iters = 200;
  cudaFuncSetAttribute(mma_rate_kernel<BN, ROWB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  mma_rate_kernel<BN, ROWB><<<grid, 128, 200 * 1024>>>(d_out, iters, nacc, spread, kstep);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
  printf("N=%3d rowB=%3d nacc=%d spread=%d kstep=%d grid=%3d : %7.1f clk/MMA  (compute floor %d)  %s\n", BN, ROWB, nacc, spread,
         kstep, grid, (double)h / (iters * 16), BN / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
}


// ---- per-stage overhead: T MMAs, then the bookkeeping a pipelined kernel does between stages -------------------
// flags: 1 = tcgen05.commit to a barrier, 2 = mbarrier.try_wait on an already-completed phase, 4 = tcgen05.fence,
//        16 = commit issued AFTER the first MMA of the next group (elect + __syncwarp always)
template <int BN, int T, int flags>
__global__ void __launch_bounds__(128, 1) stage_overhead_kernel(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar_done, bar_commit[8], bar_ready;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar_done, 1); mbar_init(&bar_ready, 1);
    for (int i = 0; i < 8; ++i) mbar_init(&bar_commit[i], 1);
    fence_mbar_init();
    mbar_arrive(&bar_ready);            // phase 0 of bar_ready is complete: try_wait(parity 0) succeeds immediately
  }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, slot, 0);
  if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(128, BN);
    const uint32_t base = smem_u32(smem);
    const uint64_t da0 = make_smem_desc<128>(base);
    const uint64_t db0 = make_smem_desc<128>(base + 64 * 1024);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (flags & 2) { while (!mbar_try_wait(&bar_ready, 0)) {} }
      if (flags & 4) tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < T; ++j) {
          const uint32_t off = (((j & 3) * 16384) >> 4) + 2 * (j & 3);
          umma_bf16(tmem, da0 + off, db0 + off, idesc, 1u);
          if ((flags & 16) && j == 0 && it > 0) umma_commit(&bar_commit[(it - 1) & 7]);
        }
        if ((flags & 1) && !(flags & 16)) umma_commit(&bar_commit[it & 7]);
      }
      __syncwarp();
    }
    if (threadIdx.x == 32) umma_commit(&bar_done);
    __syncwarp();
    mbar_wait(&bar_done, 0, 1);
    long long t1 = clock64();
    if (threadIdx.x == 32 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int BN, int T, int flags>
void run_stage(long long* d_out) {
  const int iters = 400;
  cudaFuncSetAttribute(stage_overhead_kernel<BN, T, flags>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  stage_overhead_kernel<BN, T, flags><<<148, 128, 200 * 1024>>>(d_out, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
  const double per_stage = (double)h / iters;
  const double ideal = T * (BN == 64 ? 48.0 : BN / 2.0);
  printf("stage N=%3d T=%2d flags=%2d : %7.1f clk/stage (MMA-only %6.1f, overhead %6.1f)  %s\n", BN, T, flags, per_stage, ideal,
         per_stage - ideal, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

// ---- do warps that spin on mbarrier.try_wait slow the tensor pipe?  SPIN: 0 none, 1 = 8 warps x 32 lanes poll one
// never-completing barrier, 2 = same but only lane 0 of each warp polls, 3 = poll with __nanosleep(64) back-off
template <int BN, int SPIN>
__global__ void __launch_bounds__(320, 1) spin_kernel(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar_done, bar_never;
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar_done, 1); mbar_init(&bar_never, 1); fence_mbar_init(); stop = 0; }
  if (warp == 9) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, slot, 0);
  if (warp == 9) {
    constexpr uint32_t idesc = make_idesc_bf16(128, BN);
    const uint32_t base = smem_u32(smem);
    const uint64_t da0 = make_smem_desc<128>(base);
    const uint64_t db0 = make_smem_desc<128>(base + 64 * 1024);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < 12; ++j) {
          const uint32_t off = (((j & 3) * 16384) >> 4) + 2 * (j & 3);
          umma_bf16(tmem, da0 + off, db0 + off, idesc, 1u);
        }
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&bar_done);
    __syncwarp();
    mbar_wait(&bar_done, 0, 1);
    long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    stop = 1;
  } else if (warp < 8 && SPIN > 0) {
    while (!stop) {
      if (SPIN == 1) (void)mbar_try_wait(&bar_never, 0);
      if (SPIN == 2) { if (lane == 0) (void)mbar_try_wait(&bar_never, 0); __syncwarp(); }
      if (SPIN == 3) { (void)mbar_try_wait(&bar_never, 0); __nanosleep(64); }
    }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 9) tmem_dealloc(tmem, 512);
}
template <int BN, int SPIN>
void run_spin(long long* d_out) {
  const int iters = 300;
  cudaFuncSetAttribute(spin_kernel<BN, SPIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  spin_kernel<BN, SPIN><<<148, 320, 200 * 1024>>>(d_out, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
  printf("spin N=%3d mode=%d : %7.1f clk/MMA  %s\n", BN, SPIN, (double)h / (iters * 12), e == cudaSuccess ? "" : cudaGetErrorString(e));
}

// ---- the conv kernel's exact operand addressing for an R=3 stage of conv1_2 (BN=64, CK=64): 12 MMAs reading the
// haloed 160-row A box at row offsets r*16 and k steps of 32 B, weights at r*8 KB.  MODE bits: 1 = D at TMEM column 64,
// 2 = alternate between two stage buffers, 4 = A windows do NOT overlap (r*16 KB instead of r*2 KB)
template <int MODE>
__global__ void __launch_bounds__(128, 1) pattern_kernel(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar_done;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar_done, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, slot, 0);
  if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(128, 64);
    const uint32_t base = smem_u32(smem);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t st = (MODE & 2) ? (uint32_t)(it & 1) : 0u;
      const uint64_t da0 = make_smem_desc<128>(base + 72 * 1024 + st * 20480);
      const uint64_t db0 = make_smem_desc<128>(base + (it % 3) * 24576);
      const uint32_t d = tmem + ((MODE & 1) ? 64u : 0u);
      if (elect_one()) {
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(d, da0 + ((r * ((MODE & 4) ? 16384 : 2048)) >> 4) + 2 * k, db0 + ((r * 8192) >> 4) + 2 * k, idesc, 1u);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&bar_done);
    __syncwarp();
    mbar_wait(&bar_done, 0, 1);
    long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}
template <int MODE>
void run_pattern(long long* d_out) {
  const int iters = 300;
  cudaFuncSetAttribute(pattern_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  pattern_kernel<MODE><<<148, 128, 220 * 1024>>>(d_out, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
  printf("pattern mode=%d : %7.1f clk/MMA  %s\n", MODE, (double)h / (iters * 12), e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 64);
  run_pattern<0>(d_out); run_pattern<1>(d_out); run_pattern<2>(d_out); run_pattern<3>(d_out); run_pattern<4>(d_out); run_pattern<7>(d_out);
  run_spin<64, 0>(d_out); run_spin<64, 1>(d_out); run_spin<64, 2>(d_out); run_spin<64, 3>(d_out);
  run_spin<256, 0>(d_out); run_spin<256, 1>(d_out); run_spin<128, 0>(d_out); run_spin<128, 1>(d_out);
  for (int grid : {1, 148}) {
    run<64, 128>(d_out, 1, 0, 0, grid);
    run<64, 128>(d_out, 2, 0, 0, grid);
    run<64, 128>(d_out, 1, 1, 1, grid);
    run<64, 128>(d_out, 2, 1, 1, grid);
    run<128, 128>(d_out, 1, 0, 0, grid);
    run<128, 128>(d_out, 2, 1, 1, grid);
    run<256, 128>(d_out, 1, 0, 0, grid);
    run<256, 128>(d_out, 1, 1, 1, grid);
    run<64, 64>(d_out, 1, 1, 0, grid);
    run<64, 32>(d_out, 1, 1, 0, grid);
    run<32, 128>(d_out, 1, 1, 1, grid);
    run<16, 128>(d_out, 1, 1, 1, grid);
  }
#define RUN4(F) run_stage<256, 4, F>(d_out); run_stage<64, 4, F>(d_out); run_stage<64, 12, F>(d_out); run_stage<128, 12, F>(d_out);
  RUN4(0) RUN4(1) RUN4(2) RUN4(4) RUN4(3) RUN4(7) RUN4(17) RUN4(19) RUN4(23)
  return 0;
}
