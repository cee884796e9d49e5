// Probe: does tcgen05.mma accept a NO-SWIZZLE K-major A operand laid out as planes [k-chunk][pixel][8 bf16] with an
// arbitrary 16-byte-aligned start (pixel shift), SBO = (tile width + 2) * 16 B and LBO = plane stride?  This is the layout
// the fused gather+conv1_1 kernel writes (one haloed input patch serves all nine taps by start-address shifts).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o desc_probe desc_probe.cu && ./desc_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#include "../../video_analytics_b200/csrc/va_ptx.cuh"
using namespace va;

constexpr int PITCH = 10, ROWS = 24, NPIX = PITCH * ROWS;   // 240-pixel haloed patch
constexpr int BN = 64;

__device__ __forceinline__ uint64_t desc_noswz(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  return d;   // layout type 0 = no swizzle
}

// a: [2][NPIX][8] bf16 (planes), b: [2][BN][8] bf16; out: [128][BN] fp32
__global__ void __launch_bounds__(128, 1) probe_kernel(const __nv_bfloat16* a, const __nv_bfloat16* b, float* out, int shift,
                                                        int lbo_a, int sbo_a) {
  __shared__ __align__(1024) uint8_t sa[2 * NPIX * 16];
  __shared__ __align__(1024) uint8_t sb[2 * BN * 16];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 2 * NPIX * 8; i += 128) reinterpret_cast<__nv_bfloat16*>(sa)[i] = a[i];
  for (int i = threadIdx.x; i < 2 * BN * 8; i += 128) reinterpret_cast<__nv_bfloat16*>(sb)[i] = b[i];
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&slot, 64); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1) {
    if (elect_one()) {
      const uint64_t da = desc_noswz(smem_u32(sa) + shift * 16, lbo_a, sbo_a);
      const uint64_t db = desc_noswz(smem_u32(sb), BN * 16, 128);
      umma_bf16(tmem, da, db, make_idesc_bf16(128, BN), 0u);
      umma_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0, 1);
  tc_fence_after();
  uint32_t v0[32], v1[32];
  const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16);
  tmem_ld32(ta, v0);
  tmem_ld32(ta + 32, v1);
  tmem_ld_wait();
  const int m = threadIdx.x;
  for (int i = 0; i < 32; ++i) { out[m * BN + i] = __uint_as_float(v0[i]); out[m * BN + 32 + i] = __uint_as_float(v1[i]); }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
  std::vector<__nv_bfloat16> ha(2 * NPIX * 8), hb(2 * BN * 8);
  std::vector<float> fa(NPIX * 16), fb(BN * 16);
  srand(7);
  for (int p = 0; p < NPIX; ++p)
    for (int k = 0; k < 16; ++k) {
      const float v = bf((float)(rand() % 2001 - 1000) / 500.f);
      fa[p * 16 + k] = v;
      ha[(k / 8) * NPIX * 8 + p * 8 + (k % 8)] = __float2bfloat16(v);
    }
  for (int n = 0; n < BN; ++n)
    for (int k = 0; k < 16; ++k) {
      const float v = bf((float)(rand() % 2001 - 1000) / 500.f);
      fb[n * 16 + k] = v;
      hb[(k / 8) * BN * 8 + n * 8 + (k % 8)] = __float2bfloat16(v);
    }
  __nv_bfloat16 *da, *db;
  float* dout;
  cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dout, 128 * BN * 4);
  cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  std::vector<float> ho(128 * BN);
  int bad_total = 0;
  // (shift in pixels, sbo bytes): sbo 160 = 16x8 tile rows inside a 10-wide patch; sbo 128 = linear pixels
  const int cases[][2] = {{0, 160}, {1, 160}, {11, 160}, {22, 160}, {0, 128}, {3, 128}, {13, 192}};
  for (auto& c : cases) {
    const int shift = c[0], sbo = c[1];
    cudaMemset(dout, 0, 128 * BN * 4);
    probe_kernel<<<1, 128>>>(da, db, dout, shift, NPIX * 16, sbo);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("shift %d sbo %d: CUDA error %s\n", shift, sbo, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(ho.data(), dout, 128 * BN * 4, cudaMemcpyDeviceToHost);
    int bad = 0; double maxerr = 0;
    for (int m = 0; m < 128; ++m) {
      const int pix = (m / 8) * (sbo / 16) + (m % 8) + shift;
      for (int n = 0; n < BN; ++n) {
        double ref = 0;
        for (int k = 0; k < 16; ++k) ref += (double)fa[pix * 16 + k] * fb[n * 16 + k];
        const double err = fabs(ref - ho[m * BN + n]);
        if (err > maxerr) maxerr = err;
        if (err > 1e-3 * (1 + fabs(ref))) ++bad;
      }
    }
    printf("no-swizzle planar A: shift %2d px, SBO %3d B, LBO %d B : %s (bad %d / %d, max err %.3g)\n", shift, sbo, NPIX * 16,
           bad ? "MISMATCH" : "ok", bad, 128 * BN, maxerr);
    bad_total += bad;
  }
  printf("DESC PROBE %s\n", bad_total ? "FAILED" : "PASSED");
  return bad_total ? 2 : 0;
}
