// K5: weight-gradient GEMM of the conv / fully-connected layers on tcgen05.
//
//   dW[tap = (r,s)][co][ci] = sum_{n,h,w} dZ[n, h, w, co] * X[n, h + r - pad, w + s - pad, ci]
//
// GEMM view: M = co (128-row tiles), N = ci (BN-wide tiles), K = pixels.  UMMA operands must be K-major in shared
// memory, i.e. the PIXEL index has to be the contiguous one, so both operands are first transposed to NCHW bf16
// (nhwc_to_nchw_bf16_kernel, rows padded to a 16-byte multiple): a K block is then CKP consecutive pixels of one image
// row, fetched by ONE 4-D TMA box (w, h, channel, n) per operand.  The vertical tap offset r is the same box at row
// h + r - pad with TMA zero fill outside the image (the forward kernel's halo trick); the horizontal offset s cannot be
// a box shift (a TMA box must start 16-byte aligned; one pixel is 2 bytes), so X^T is stored ks times, copy s shifted
// by s - pad pixels, and tap (r,s) reads copy s (image index s*n + img of the same tensor map).
// Work unit = (tap, co tile, ci tile, K split); the fp32 TMEM accumulator of a unit is added to dW with
// red.global.add.f32 (split-K).  Same warp-specialised pipeline as the forward kernel (TMA producer warp, MMA warp,
// 4 epilogue warps, mbarrier full/empty rings, 2 TMEM accumulator stages).
// Reference being replaced: loss.backward() for Conv2d/Linear weights (Sheet03/spatialModel.py:180).
#include "va_internal.h"
#include "va_conv_tc.cuh"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

namespace va {

struct WgradParams {
  int n, H, wchunks;
  int ks, pad, Cout, Cin;
  int co_tiles, ci_tiles, taps;
  int total_units, rows_total, rows_per_split;
  int num_stages;
  float* dwt;                 // fp32 [taps][Cout][Cin], pre-zeroed
  int skip;                   // diagnostics: bit0 no TMA, bit1 no MMA, bit2 no TMEM load, bit3 no atomics
  int* dbg;                   // diagnostics (VA_WGRAD_DEBUG): soft watchdog records the stuck wait's tag instead of trapping
  FastDiv div_taps, div_ci, div_co, div_h;
};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

constexpr int kWgradThreads = 192;

// mbarrier wait; with a debug buffer a wait that times out records (tag, block, parity) and lets every role run to
// completion (results are garbage) so that the host can read which barrier stalled.
__device__ __forceinline__ void wg_wait(uint64_t* bar, uint32_t parity, int tag, int* dbg) {
  if (dbg == nullptr) { mbar_wait(bar, parity, tag); return; }
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (*reinterpret_cast<volatile int*>(dbg) != 0) return;
    if (clock64() - t0 > 200000000ll) {
      if (atomicCAS(dbg, 0, 1) == 0) { dbg[1] = tag; dbg[2] = (int)blockIdx.x; dbg[3] = (int)parity; dbg[4] = (int)threadIdx.x; }
      return;
    }
  }
}

template <int BN, int CKP>
__global__ void __launch_bounds__(kWgradThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const WgradParams p) {
  constexpr int ROWB = CKP * 2;
  constexpr uint32_t A_BYTES = 128 * ROWB, B_BYTES = BN * ROWB;
  constexpr uint32_t A_ALLOC = (A_BYTES + 1023) & ~1023u, B_ALLOC = (B_BYTES + 1023) & ~1023u;
  constexpr uint32_t STAGE = A_ALLOC + B_ALLOC;
  constexpr uint32_t TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.num_stages * STAGE);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tfull_bar = empty_bar + kMaxStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kProducerWarp = 4, kMmaWarp = 5;   // warps 0..3 = epilogue (TMEM lane quarter = warp)
  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < p.num_stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
    fence_mbar_init();
  }
  if (warp == kMmaWarp) { tmem_alloc(tmem_slot, TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // unit -> (split, co tile, ci tile, tap); tap fastest so that concurrently running CTAs stream the same rows
  auto decode = [&](int unit, int& tap, int& ci_t, int& co_t, int& split) {
    uint32_t q, t;
    p.div_taps.divmod((uint32_t)unit, q, t); tap = (int)t;
    p.div_ci.divmod(q, q, t); ci_t = (int)t;
    p.div_co.divmod(q, q, t); co_t = (int)t;
    split = (int)q;
  };

  if (warp == kProducerWarp) {
    uint32_t stage = 0, phase = 0;
    for (int unit = blockIdx.x; unit < p.total_units; unit += gridDim.x) {
      int tap, ci_t, co_t, split;
      decode(unit, tap, ci_t, co_t, split);
      const int s = tap / p.ks, r = tap - s * p.ks;        // tap' = s*ks + r (the forward kernel's order)
      const int row0 = split * p.rows_per_split;
      const int row1 = min(p.rows_total, row0 + p.rows_per_split);
      for (int row = row0; row < row1; ++row) {
        uint32_t img, h;
        p.div_h.divmod((uint32_t)row, img, h);
        const int hb = (int)h + r - p.pad;
        if (hb < 0 || hb >= p.H) continue;                 // the shifted row is all padding: contributes nothing
        for (int wc = 0; wc < p.wchunks; ++wc) {
          wg_wait(&empty_bar[stage], phase ^ 1, 100 + stage, p.dbg);
          if (elect_one()) {
            uint8_t* a_dst = smem + (size_t)stage * STAGE;
            if (p.skip & 1) {
              mbar_arrive(&full_bar[stage]);
            } else {
              const int z = (p.skip & 64) ? 0 : 1;
              mbar_arrive_expect_tx(&full_bar[stage], ((p.skip & 16) ? 0 : A_BYTES) + ((p.skip & 32) ? 0 : B_BYTES));
              if (!(p.skip & 16)) tma_load_4d(a_dst, &tmA, &full_bar[stage], z * wc * CKP, z * (int)h, z * co_t * 128, z * (int)img);
              if (!(p.skip & 32)) tma_load_4d(a_dst + A_ALLOC, &tmB, &full_bar[stage], z * wc * CKP, z * hb, z * ci_t * BN, z * (s * p.n + (int)img));
            }
          }
          __syncwarp();
          if (++stage == (uint32_t)p.num_stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    constexpr uint32_t idesc = make_idesc_bf16(128, BN);
    const uint32_t smem_base_u32 = smem_u32(smem);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    uint32_t stage = 0, phase = 0, as = 0, as_phase = 0;
    for (int unit = blockIdx.x; unit < p.total_units; unit += gridDim.x) {
      int tap, ci_t, co_t, split;
      decode(unit, tap, ci_t, co_t, split);
      const int s = tap / p.ks, r = tap - s * p.ks;
      const int row0 = split * p.rows_per_split;
      const int row1 = min(p.rows_total, row0 + p.rows_per_split);
      wg_wait(&tempty_bar[as], as_phase ^ 1, 200 + as, p.dbg);
      tc_fence_after();
      const uint32_t d_tmem = tmem_u + as * BN;
      uint32_t acc = 0;
      for (int row = row0; row < row1; ++row) {
        uint32_t img, h;
        p.div_h.divmod((uint32_t)row, img, h);
        const int hb = (int)h + r - p.pad;
        if (hb < 0 || hb >= p.H) continue;
        for (int wc = 0; wc < p.wchunks; ++wc) {
          wg_wait(&full_bar[stage], phase, 300 + stage, p.dbg);
          tc_fence_after();
          const uint32_t a_addr = smem_base_u32 + stage * STAGE;
          const uint64_t da0 = make_smem_desc<ROWB>(a_addr);
          const uint64_t db0 = make_smem_desc<ROWB>(a_addr + A_ALLOC);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < CKP / 16; ++k)
              if (!(p.skip & 2)) umma_bf16(d_tmem, da0 + 2 * k, db0 + 2 * k, idesc, k ? 1u : acc);
            umma_commit(&empty_bar[stage]);
          }
          __syncwarp();
          acc = 1;
          if (++stage == (uint32_t)p.num_stages) { stage = 0; phase ^= 1; }
        }
      }
      // a unit whose every row was skipped (cannot happen for H >= 2, kept for safety) would leave garbage: zero it
      if (elect_one()) {
        if (acc == 0) {
          // no MMA was issued: signal the epilogue with an "empty" accumulator flag through the barrier anyway
        }
        umma_commit(&tfull_bar[as]);
      }
      __syncwarp();
      as ^= 1;
      if (as == 0) as_phase ^= 1;
    }
  } else {
    // ===================================================== epilogue: TMEM -> fp32 atomics into dW
    const int q = warp & 3;
    const int m = q * 32 + lane;            // accumulator row == output channel inside the co tile
    uint32_t as = 0, as_phase = 0;
    for (int unit = blockIdx.x; unit < p.total_units; unit += gridDim.x) {
      int tap, ci_t, co_t, split;
      decode(unit, tap, ci_t, co_t, split);
      const int row0 = split * p.rows_per_split;
      const int row1 = min(p.rows_total, row0 + p.rows_per_split);
      wg_wait(&tfull_bar[as], as_phase, 400 + as, p.dbg);
      tc_fence_after();
      const int co = co_t * 128 + m;
      float* dst = p.dwt + ((size_t)tap * p.Cout + co) * p.Cin + ci_t * BN;
      // did this unit issue any MMA?  (same skip rule as above; H >= 2 always leaves at least one row)
      bool any = false;
      {
        const int s = tap / p.ks, r = tap - s * p.ks;
        for (int row = row0; row < row1 && !any; ++row) {
          uint32_t img, h;
          p.div_h.divmod((uint32_t)row, img, h);
          const int hb = (int)h + r - p.pad;
          any = (hb >= 0 && hb < p.H);
        }
      }
#pragma unroll 1
      for (int c16 = 0; c16 < BN / 16; ++c16) {
        uint32_t v[16];
        if (!(p.skip & 4)) {
          tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + c16 * 16, v);
          tmem_ld_wait();
        } else {
          for (int i = 0; i < 16; ++i) v[i] = 0;
        }
        if (any && co < p.Cout && !(p.skip & 8)) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int ci = ci_t * BN + c16 * 16 + i;
            if (ci < p.Cin) atomicAdd(dst + c16 * 16 + i, __uint_as_float(v[i]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      as ^= 1;
      if (as == 0) as_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, TMEM_COLS);
}

// [taps][Cout][Cin] (tap' = s*ks + r) -> OIHW fp32 [Cout][Cin][ks][ks]
__global__ void wgrad_to_oihw_kernel(const float* __restrict__ dwt, float* __restrict__ dw, int Cout, int Cin, int ks) {
  const size_t total = (size_t)Cout * Cin * ks * ks;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
    const int s = (int)(g % ks), r = (int)((g / ks) % ks);
    const int ci = (int)((g / (ks * ks)) % Cin);
    const int co = (int)(g / ((size_t)ks * ks * Cin));
    dw[g] = dwt[((size_t)(s * ks + r) * Cout + co) * Cin + ci];
  }
}

namespace {
thread_local char g_werr[512];
const char* werrf(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_werr, sizeof(g_werr), fmt, ap);
  va_end(ap);
  return g_werr;
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}
// NCHW bf16 [n][C][H][Wp] map with dims (Wp, H, C, n), box (ckp, 1, rows, 1)
const char* encode_nchw(CUtensorMap* m, const void* addr, int n, int C, int H, int Wp, int ckp, int rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return "cuTensorMapEncodeTiled not available";
  cuuint64_t gdim[4] = {(cuuint64_t)Wp, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)n};
  cuuint64_t gstr[3] = {(cuuint64_t)Wp * 2, (cuuint64_t)Wp * H * 2, (cuuint64_t)Wp * H * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)ckp, 1u, (cuuint32_t)rows, 1u};
  cuuint32_t es[4] = {1, 1, 1, 1};
  const CUtensorMapSwizzle sw = ckp == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (ckp == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(addr), gdim, gstr, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return werrf("wgrad tensor map encode failed (%d): n=%d C=%d H=%d Wp=%d ckp=%d rows=%d", (int)r, n, C, H, Wp, ckp, rows);
  return nullptr;
}

template <int BN, int CKP>
const char* launch_wgrad(const CUtensorMap& tA, const CUtensorMap& tB, WgradParams p, int grid, cudaStream_t st) {
  constexpr uint32_t A_ALLOC = (128 * CKP * 2 + 1023) & ~1023u, B_ALLOC = (BN * CKP * 2 + 1023) & ~1023u;
  int stages = (int)((227 * 1024 - 1024 - 256) / (A_ALLOC + B_ALLOC));
  if (stages > kMaxStages) stages = kMaxStages;
  p.num_stages = stages;
  const size_t smem = 1024 + (size_t)stages * (A_ALLOC + B_ALLOC) + 256;
  auto kfn = wgrad_tc_kernel<BN, CKP>;
  static size_t configured = 0;
  if (configured < smem) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return werrf("cudaFuncSetAttribute(wgrad smem=%zu): %s", smem, cudaGetErrorString(e));
    configured = smem;
  }
  count_launch();
  kfn<<<grid, kWgradThreads, smem, st>>>(tA, tB, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return werrf("wgrad_tc_kernel<%d,%d> launch: %s", BN, CKP, cudaGetErrorString(e));
  return nullptr;
}
}  // namespace

// dz_nchw: bf16 [n][Cout][H][Wp]; x_nchw: bf16 [ks][n][Cin][H][Wp] (copy s shifted by s - pad pixels); dwt_ws: fp32 [ks*ks][Cout][Cin] workspace;
// dw_oihw: fp32 [Cout][Cin][ks][ks] result.
const char* wgrad_run(const void* dz_nchw, const void* x_nchw, int n, int H, int W, int Wp, int Cout, int Cin, int ks,
                      float* dwt_ws, float* dw_oihw, cudaStream_t st) {
  if (n <= 0) return nullptr;
  const int ckp = W > 32 ? 64 : (W > 16 ? 32 : 16);
  if (Wp % 8 != 0) return "wgrad: padded row length must be a multiple of 8 elements";
  // N tile: the smallest supported width that covers Cin (16/32/64) or 128-wide tiles
  int BN = Cin <= 16 ? 16 : (Cin <= 32 ? 32 : (Cin <= 64 ? 64 : 128));
  WgradParams p;
  p.n = n; p.H = H; p.wchunks = (W + ckp - 1) / ckp;
  p.ks = ks; p.pad = (ks - 1) / 2; p.Cout = Cout; p.Cin = Cin;
  p.co_tiles = (Cout + 127) / 128; p.ci_tiles = (Cin + BN - 1) / BN; p.taps = ks * ks;
  p.rows_total = n * H;
  int sms = 148;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int base_units = p.taps * p.co_tiles * p.ci_tiles;
  int ksplit = (2 * sms + base_units - 1) / base_units;        // aim at >= 2 units per SM
  if (ksplit > p.rows_total) ksplit = p.rows_total;
  if (ksplit < 1) ksplit = 1;
  p.rows_per_split = (p.rows_total + ksplit - 1) / ksplit;
  ksplit = (p.rows_total + p.rows_per_split - 1) / p.rows_per_split;
  p.total_units = base_units * ksplit;
  p.dwt = dwt_ws;
  p.dbg = nullptr;
  p.skip = getenv("VA_WGRAD_SKIP") ? atoi(getenv("VA_WGRAD_SKIP")) : 0;
  static int* dbg_buf = nullptr;
  const bool debug = getenv("VA_WGRAD_DEBUG") != nullptr;
  if (debug) {
    if (!dbg_buf) cudaMalloc(&dbg_buf, 64);
    cudaMemsetAsync(dbg_buf, 0, 64, st);
    p.dbg = dbg_buf;
  }
  p.div_taps = FastDiv::make((uint32_t)p.taps);
  p.div_ci = FastDiv::make((uint32_t)p.ci_tiles);
  p.div_co = FastDiv::make((uint32_t)p.co_tiles);
  p.div_h = FastDiv::make((uint32_t)H);
  cudaError_t ce = cudaMemsetAsync(dwt_ws, 0, (size_t)p.taps * Cout * Cin * sizeof(float), st);
  if (ce != cudaSuccess) return werrf("wgrad memset: %s", cudaGetErrorString(ce));
  CUtensorMap tA, tB;
  if (const char* e = encode_nchw(&tA, dz_nchw, n, Cout, H, Wp, ckp, 128)) return e;
  if (const char* e = encode_nchw(&tB, x_nchw, ks * n, Cin, H, Wp, ckp, BN)) return e;
  const int grid = p.total_units < sms ? p.total_units : sms;
  const char* err = nullptr;
#define VA_W(bn, ck) if (BN == bn && ckp == ck) err = launch_wgrad<bn, ck>(tA, tB, p, grid, st); else
  VA_W(16, 64) VA_W(32, 64) VA_W(64, 64) VA_W(128, 64) VA_W(64, 32) VA_W(128, 32) VA_W(64, 16) VA_W(128, 16)
  VA_W(16, 32) VA_W(32, 32) VA_W(16, 16) VA_W(32, 16)
  err = werrf("wgrad: no kernel for BN=%d CKP=%d", BN, ckp);
#undef VA_W
  if (err) return err;
  if (debug) {
    int h[8] = {0};
    cudaStreamSynchronize(st);
    cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[va wgrad debug] BN=%d ckp=%d units=%d rows/split=%d wchunks=%d grid=%d | stalled=%d tag=%d block=%d parity=%d thread=%d\n",
            BN, ckp, p.total_units, p.rows_per_split, p.wchunks, grid, h[0], h[1], h[2], h[3], h[4]);
  }
  const size_t total = (size_t)Cout * Cin * ks * ks;
  unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, 148 * 8);
  count_launch();
  wgrad_to_oihw_kernel<<<blocks, 256, 0, st>>>(dwt_ws, dw_oihw, Cout, Cin, ks);
  ce = cudaGetLastError();
  if (ce != cudaSuccess) return werrf("wgrad_to_oihw: %s", cudaGetErrorString(ce));
  return nullptr;
}

}  // namespace va
