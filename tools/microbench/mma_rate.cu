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

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 64);
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
  return 0;
}
