// K1+K2 fused: snippet preprocess gathered straight into conv1_1's shared-memory operand.
//
// Replaces, in one kernel, the reference's per-item transform pipeline (Sheet03/utils.py:137-151: RandomCrop /
// RandomHorizontalFlip / ToTensor / Normalize; per-image for the 20 flow planes, temporalModel.py:80-90) AND the first
// VGG16 layer `features[0:2]` = Conv2d(C0, 64, 3, padding=1) + ReLU (spatialModel.py:110,171; temporalModel.py:149-162).
// The separate K1 kernel wrote a channel-padded bf16 NHWC tensor (1.6 / 3.2 MB per snippet) that conv1_1 read back;
// here the u8 crop is the only HBM input of the layer.
//
// Tile = 16 x 8 output pixels of one snippet.  Four gather warps read the 18 x 10 haloed source patch of every plane
// (index-table row = image id, crop top/left, flip), normalise by table lookup (bf16(((u8/255) - mean)/std), the same
// IEEE operations as torchvision, built per CTA), and write it into a pipeline stage as PLANES of 8 channels:
//     stage[chunk j][pixel p = row*10 + col][8 x bf16]        (16-byte granules, zero outside the crop = conv padding)
// This is a NO-SWIZZLE K-major UMMA operand in which 8-pixel groups (one output row of the tile) are 160 B apart (SBO)
// and the two 8-channel halves of a K=16 MMA are LBO apart -- and because LBO is free, the second half may be ANOTHER
// TAP of the same chunk (LBO = 16 B: the pixel to the right; 160 B: the pixel below).  Filter tap (r,s) is a start
// address shift of (r*10 + s)*16 B (verified on B200 by tools/microbench/desc_probe.cu).  K is therefore packed
// densely: 3 input channels need 5 MMAs per tile (K = 9 taps x 8 channels, last slot zero-weighted) instead of 9 with
// 16-channel padding; 20 channels need 14 instead of 18.  Weights are packed to match and stay resident in smem.
//
// Warp roles (416 threads, 1 CTA/SM, persistent): warps 0-3 / 4-7 two epilogue groups (TMEM -> +bias -> ReLU -> bf16 ->
// 128B-swizzled staging -> TMA store), warps 8-11 gather, warp 12 MMA issuer + TMEM owner.
#include "va_internal.h"
#include "va_conv_tc.cuh"

#include <mutex>
#include <stdio.h>

namespace va {

namespace {

constexpr int kF1Threads = 416;
constexpr int kF1TileH = 16, kF1TileW = 8, kF1Pitch = kF1TileW + 2, kF1Rows = kF1TileH + 2;
constexpr int kF1Pix = kF1Pitch * kF1Rows;            // 180 haloed pixels
constexpr int kF1PlaneBytes = kF1Pix * 16;            // one 8-channel plane of a stage
constexpr int kF1Crop = 224;
constexpr int kF1MaxMma = 18;
constexpr int kF1Stages = 4;
constexpr int kF1StagingBytes = 128 * 128;            // 128 pixels x 64 bf16

struct Conv1FusedParams {
  const uint8_t* images;
  unsigned long long image_bytes;
  int img_w;
  const int32_t* table;          // [n][planes][4] = image id, crop top, crop left, flip
  int n_img, planes;
  FastDiv div_planes;
  int n_chunks;                  // ceil(planes * IMG_C / 8)
  int n_luts;
  float lut_mean[3], lut_std[3];
  unsigned char lut_of[32];
  const __nv_bfloat16* w_packed; // [n_mma][2][64][8]
  const float* bias;             // [64]
  int total_tiles;
  FastDiv div_w, div_h;          // by 28, 14
  int n_mma;
  unsigned long long a_delta[kF1MaxMma];   // per MMA: (first-chunk byte offset >> 4) | (LBO >> 4) << 16
};

__device__ __forceinline__ uint64_t desc_noswizzle_base(uint32_t sbo_bytes) {
  return ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);       // layout type 0, descriptor version 1
}

template <int IMG_C, int SEG, int REP, int ROUNDS>
__global__ void __launch_bounds__(kF1Threads, 1)
conv1_fused_kernel(const __grid_constant__ CUtensorMap tmO, const Conv1FusedParams p) {
  constexpr int NSEG = kF1Pitch / SEG;
  static_assert(NSEG * SEG == kF1Pitch, "SEG must divide the patch width");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem;                                               // 2 groups x 2 x 16 KB, 1024-aligned (128B swizzle)
  uint8_t* w_s = staging + 4 * kF1StagingBytes;                          // n_mma x 2 KB
  const uint32_t a_stage_bytes = (uint32_t)p.n_chunks * kF1PlaneBytes;
  uint8_t* a_ring = w_s + (size_t)p.n_mma * 2048;
  uint16_t* lut_s = reinterpret_cast<uint16_t*>(a_ring + (size_t)kF1Stages * a_stage_bytes);
  float* bias_s = reinterpret_cast<float*>(lut_s + (size_t)p.n_luts * 256 * REP);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bias_s + 64);
  uint64_t* empty_bar = full_bar + kF1Stages;
  uint64_t* tfull_bar = empty_bar + kF1Stages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int kGatherWarp0 = 8, kMmaWarp = 12;

  // ---- prologue: barriers, TMEM, resident weights, normalisation table, zeroed stages
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmO);
    for (int i = 0; i < kF1Stages; ++i) { mbar_init(&full_bar[i], 4); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
    fence_mbar_init();
  }
  if (warp == kMmaWarp) { tmem_alloc(tmem_slot, 128); tmem_relinquish(); }
  {
    const uint4* wg = reinterpret_cast<const uint4*>(p.w_packed);
    uint4* ws4 = reinterpret_cast<uint4*>(w_s);
    for (int i = threadIdx.x; i < p.n_mma * 128; i += kF1Threads) ws4[i] = __ldg(wg + i);
    // pad channels of the last chunk (and everything else) start as zero and are never written again
    uint4* a4 = reinterpret_cast<uint4*>(a_ring);
    const int n16 = (int)(kF1Stages * a_stage_bytes / 16);
    for (int i = threadIdx.x; i < n16; i += kF1Threads) a4[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = threadIdx.x; i < p.n_luts * 256; i += kF1Threads) {
      const int k = i >> 8, u = i & 255;
      // ToTensor (u8 / 255) then Normalize ((x - mean) / std) in IEEE fp32, rounded once to bf16 (utils.py:148-150)
      const float v = __fdiv_rn(__fsub_rn(__fdiv_rn((float)u, 255.0f), p.lut_mean[k]), p.lut_std[k]);
      const uint16_t b = __bfloat16_as_ushort(__float2bfloat16_rn(v));
      for (int r = 0; r < REP; ++r) lut_s[(size_t)i * REP + r] = b;
    }
    if (threadIdx.x < 64) bias_s[threadIdx.x] = __ldg(p.bias + threadIdx.x);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Every CTA walks a CONTIGUOUS range of tiles (row-major inside a snippet): consecutive tiles read neighbouring 8-pixel
  // column blocks of the same source rows and the same index-table row, so the gather's loads hit L1 (a strided
  // assignment made every tile a fresh L2 round trip for the table row AND the pixels: 1650 clk per tile measured).
  const int tiles_per = p.total_tiles / (int)gridDim.x, tiles_rem = p.total_tiles % (int)gridDim.x;
  const int tile_begin = (int)blockIdx.x * tiles_per + min((int)blockIdx.x, tiles_rem);
  const int tile_end = tile_begin + tiles_per + ((int)blockIdx.x < tiles_rem ? 1 : 0);

  if (warp >= kGatherWarp0 && warp < kGatherWarp0 + 4) {
    // ===================================================== gather + normalise -> A operand (4 warps)
    // Work item = SEG consecutive pixels of one patch row of one plane (flow: a whole 10-pixel row; RGB: 2 pixels x 3
    // channels); thread t owns items t, t+128, ... so row / plane / destination offset are per-thread constants.  The
    // source bytes of the NEXT tile are loaded into registers before the current tile is converted: the two dependent
    // L2 round trips (index-table row, then pixels) are hidden behind the table lookups and stores of the current tile
    // (first version, one tile at a time: 2.2 ms per 250 flow stacks, latency-bound).
    const int gt = threadIdx.x - kGatherWarp0 * 32;
    const int items = kF1Rows * p.planes * NSEG;
    const int4* table4 = reinterpret_cast<const int4*>(p.table);
    constexpr int NB = SEG * IMG_C;                     // source bytes per item
    constexpr uint32_t FULL = (1u << SEG) - 1u;
    int it_row[ROUNDS], it_plane[ROUNDS], it_x[ROUNDS];
    uint32_t it_off[ROUNDS];
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
      const int i = gt + 128 * r;
      uint32_t seg = 0, rp = (uint32_t)i;
      if (NSEG > 1) { rp = (uint32_t)i / NSEG; seg = (uint32_t)i - rp * NSEG; }
      uint32_t row, plane;
      p.div_planes.divmod(rp, row, plane);
      const bool valid = i < items;
      it_row[r] = valid ? (int)row : -1;
      it_plane[r] = (int)plane;
      it_x[r] = (int)seg * SEG;
      const int ch0 = (int)plane * IMG_C;
      it_off[r] = (uint32_t)(((int)row * kF1Pitch + (int)seg * SEG) * 16 + (ch0 >> 3) * kF1PlaneBytes + (ch0 & 7) * 2);
    }
    uint32_t ua[ROUNDS][NB], ub[ROUNDS][NB], ma[ROUNDS], mb[ROUNDS];

    auto load_tile = [&](int tile, uint32_t (&u)[ROUNDS][NB], uint32_t (&msk)[ROUNDS]) {
      uint32_t mt, tw, th, tn;
      p.div_w.divmod((uint32_t)tile, mt, tw);
      p.div_h.divmod(mt, tn, th);
      const int h0 = (int)th * kF1TileH, w0 = (int)tw * kF1TileW;
      // patch columns inside the crop (bit q <-> x = w0 - 1 + q): everything except the conv's zero-padding columns
      const uint32_t colmask = 0x3FFu & ~(w0 == 0 ? 1u : 0u) & ~(w0 + kF1TileW == kF1Crop ? (1u << (kF1Pitch - 1)) : 0u);
#pragma unroll
      for (int r = 0; r < ROUNDS; ++r) {
        msk[r] = 0;
        if (it_row[r] < 0) continue;
        const int y = h0 - 1 + it_row[r];
        if ((unsigned)y >= (unsigned)kF1Crop) {                            // zero-padding row: no loads, table index 0
#pragma unroll
          for (int b = 0; b < NB; ++b) u[r][b] = 0u;
          continue;
        }
        const uint32_t m = (colmask >> it_x[r]) & FULL;
        msk[r] = m;
        const int4 e = __ldg(table4 + (size_t)tn * p.planes + it_plane[r]);   // image id, crop top, crop left, flip
        const uint8_t* base = p.images + (unsigned long long)e.x * p.image_bytes +
                              ((long long)(e.y + y) * p.img_w + e.z) * IMG_C;
        const int x0 = w0 - 1 + it_x[r];
        if (m == FULL) {
          if (!e.w) {
            const uint8_t* s0 = base + x0 * IMG_C;
#pragma unroll
            for (int q = 0; q < SEG; ++q)
#pragma unroll
              for (int k = 0; k < IMG_C; ++k) u[r][q * IMG_C + k] = __ldg(s0 + q * IMG_C + k);
          } else {                                                         // hflip of the crop == reversed columns
            const uint8_t* s0 = base + (kF1Crop - 1 - x0) * IMG_C;
#pragma unroll
            for (int q = 0; q < SEG; ++q)
#pragma unroll
              for (int k = 0; k < IMG_C; ++k) u[r][q * IMG_C + k] = __ldg(s0 - q * IMG_C + k);
          }
        } else {
#pragma unroll
          for (int q = 0; q < SEG; ++q) {
            const int x = x0 + q;
            const int xs = e.w ? (kF1Crop - 1 - x) : x;
#pragma unroll
            for (int k = 0; k < IMG_C; ++k) u[r][q * IMG_C + k] = ((m >> q) & 1u) ? (uint32_t)__ldg(base + xs * IMG_C + k) : 0u;
          }
        }
      }
    };
    auto convert_tile = [&](uint8_t* a_dst, const uint32_t (&u)[ROUNDS][NB], const uint32_t (&msk)[ROUNDS]) {
#pragma unroll
      for (int r = 0; r < ROUNDS; ++r) {
        if (it_row[r] < 0) continue;
        uint8_t* dst = a_dst + it_off[r];
        const uint32_t m = msk[r];
#pragma unroll
        for (int q = 0; q < SEG; ++q) {
#pragma unroll
          for (int k = 0; k < IMG_C; ++k) {
            const int li = (REP == 32 || p.n_luts == 1) ? 0 : (int)p.lut_of[it_plane[r] * IMG_C + k];
            uint16_t v = lut_s[(size_t)(li * 256 + (int)u[r][q * IMG_C + k]) * REP + (REP == 32 ? lane : 0)];
            if (m != FULL && !((m >> q) & 1u)) v = 0;                      // outside the crop: the conv's zero padding
            *reinterpret_cast<uint16_t*>(dst + q * 16 + k * 2) = v;
          }
        }
      }
    };

    uint32_t stage = 0, phase = 0;
    int tile = tile_begin;
    if (tile < tile_end) load_tile(tile, ua, ma);
    // two tiles per iteration so that the register sets alternate without copies
    while (tile < tile_end) {
      {
        const int next = tile + 1;
        if (next < tile_end) load_tile(next, ub, mb);
        mbar_wait(&empty_bar[stage], phase ^ 1, 100 + stage);
        convert_tile(a_ring + (size_t)stage * a_stage_bytes, ua, ma);
        fence_proxy_async_smem();          // generic-proxy stores -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[stage]);
        if (++stage == kF1Stages) { stage = 0; phase ^= 1; }
        tile = next;
      }
      if (tile >= tile_end) break;
      {
        const int next = tile + 1;
        if (next < tile_end) load_tile(next, ua, ma);
        mbar_wait(&empty_bar[stage], phase ^ 1, 100 + stage);
        convert_tile(a_ring + (size_t)stage * a_stage_bytes, ub, mb);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[stage]);
        if (++stage == kF1Stages) { stage = 0; phase ^= 1; }
        tile = next;
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================================================== MMA issuer (convergent warp, one elected lane issues)
    constexpr uint32_t idesc = make_idesc_bf16(128, 64);
    const uint32_t a_ring_u32 = smem_u32(a_ring);
    const uint32_t w_u32 = smem_u32(w_s);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint64_t da_base = desc_noswizzle_base(kF1Pitch * 16);                         // 8-pixel groups 160 B apart
    const uint64_t db_base = desc_noswizzle_base(128) | ((uint64_t)(1024 >> 4) << 16);   // [k-chunk][n][8]: SBO 128, LBO 1024
    uint32_t stage = 0, phase = 0, as = 0, as_phase = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      mbar_wait(&tempty_bar[as], as_phase ^ 1, 200 + as);
      mbar_wait(&full_bar[stage], phase, 300 + stage);
      tc_fence_after();
      const uint64_t da0 = da_base + ((a_ring_u32 + stage * a_stage_bytes) >> 4);
      const uint64_t db0 = db_base + (w_u32 >> 4);
      const uint32_t d_tmem = tmem_u + as * 64;
      if (elect_one()) {
        for (int i = 0; i < p.n_mma; ++i) umma_bf16(d_tmem, da0 + p.a_delta[i], db0 + (uint64_t)(i * (2048 >> 4)), idesc, i ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        umma_commit(&tfull_bar[as]);
      }
      __syncwarp();
      if (++stage == kF1Stages) { stage = 0; phase ^= 1; }
      as ^= 1;
      if (as == 0) as_phase ^= 1;
    }
  } else if (warp < 8) {
    // ===================================================== epilogue (2 groups x 4 warps)
    const int eg = warp >> 2;
    const int q = warp & 3;
    const int m = q * 32 + lane;                  // accumulator row == pixel h_i*8 + w_i of the tile
    const int et = threadIdx.x - eg * 128;
    // TWO staging buffers per group: the layer writes 16 KB per tile, and a TMA store holds its buffer until the store
    // engine has read it (~1 us under HBM write pressure) -- with one buffer per group that turnaround, not bandwidth,
    // set the tile rate (0.98 us per tile = 2.4 TB/s of writes measured).
    uint8_t* stage_base = staging + eg * 2 * kF1StagingBytes;
    const uint32_t as = (uint32_t)eg;
    int it = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile, ++it) {
      if ((it & 1) != eg) continue;
      const uint32_t as_phase = (uint32_t)(it >> 1) & 1u;
      uint32_t mt, tw, th, tn;
      p.div_w.divmod((uint32_t)tile, mt, tw);
      p.div_h.divmod(mt, tn, th);
      mbar_wait(&tfull_bar[as], as_phase, 400 + as);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * 64;
      tmem_ld32(taddr, v0);
      tmem_ld32(taddr + 32, v1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      uint32_t pk[32];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float a = fmaxf(__uint_as_float(v0[2 * i]) + bias_s[2 * i], 0.f);
        const float b = fmaxf(__uint_as_float(v0[2 * i + 1]) + bias_s[2 * i + 1], 0.f);
        const float c = fmaxf(__uint_as_float(v1[2 * i]) + bias_s[32 + 2 * i], 0.f);
        const float d = fmaxf(__uint_as_float(v1[2 * i + 1]) + bias_s[32 + 2 * i + 1], 0.f);
        __nv_bfloat162 lo = __floats2bfloat162_rn(a, b);
        __nv_bfloat162 hi = __floats2bfloat162_rn(c, d);
        pk[i] = *reinterpret_cast<uint32_t*>(&lo);
        pk[16 + i] = *reinterpret_cast<uint32_t*>(&hi);
      }
      uint8_t* stage_out = stage_base + ((it >> 1) & 1) * kF1StagingBytes;
      if (et < 32) {                       // this buffer was last read by the group's store of two tiles ago
        if (elect_one()) tma_store_wait_read<1>();
        __syncwarp();
      }
      named_bar_sync(1 + eg, 128);
      uint8_t* rowp = stage_out + m * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint4 val = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        *reinterpret_cast<uint4*>(rowp + ((c ^ (m & 7)) << 4)) = val;     // 128B swizzle, as the TMA store expects
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + eg, 128);
      if (et < 32) {
        if (elect_one()) {
          tma_store_4d(&tmO, stage_out, 0, (int)tw * kF1TileW, (int)th * kF1TileH, (int)tn);
          tma_store_commit();
        }
        __syncwarp();
      }
    }
    if (et < 32) {
      if (elect_one()) tma_store_wait_all();
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, 128);
}

// ---- weight packing for the dense-K schedule -------------------------------------------------------------------
struct Conv1PackPlan {
  int n_mma;
  signed char tap[kF1MaxMma][2];     // r*3+s, or -1 = zero weights
  signed char chunk[kF1MaxMma][2];
};

__global__ void pack_conv1_fused_w_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int cin,
                                          const Conv1PackPlan plan) {
  // out [n_mma][2][64][8]; w OIHW [64][cin][3][3]
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= plan.n_mma * 2 * 64 * 8) return;
  const int e = idx & 7, n = (idx >> 3) & 63, kc = (idx >> 9) & 1, i = idx >> 10;
  const int tap = plan.tap[i][kc], ch = plan.chunk[i][kc] * 8 + e;
  float v = 0.f;
  if (tap >= 0 && ch < cin) v = w[((size_t)n * cin + ch) * 9 + tap];
  out[idx] = __float2bfloat16_rn(v);
}

struct Conv1Schedule {
  Conv1PackPlan plan;
  unsigned long long a_delta[kF1MaxMma];
};

Conv1Schedule make_schedule(int n_chunks) {
  Conv1Schedule s;
  s.plan.n_mma = 0;
  auto shift = [](int tap) { return (tap / 3) * kF1Pitch + (tap % 3); };
  auto add = [&](int tap0, int chunk0, bool zero0, int tap1, int chunk1) {
    const int i = s.plan.n_mma++;
    const long long off0 = (long long)chunk0 * kF1PlaneBytes + shift(tap0) * 16;
    const long long off1 = (long long)chunk1 * kF1PlaneBytes + shift(tap1) * 16;
    s.plan.tap[i][0] = (signed char)(zero0 ? -1 : tap0); s.plan.chunk[i][0] = (signed char)chunk0;
    s.plan.tap[i][1] = (signed char)tap1; s.plan.chunk[i][1] = (signed char)chunk1;
    s.a_delta[i] = (unsigned long long)(off0 >> 4) | ((unsigned long long)((off1 - off0) >> 4) << 16);
  };
  for (int tap = 0; tap < 9; ++tap)
    for (int q = 0; q + 1 < n_chunks; q += 2) add(tap, q, false, tap, q + 1);
  if (n_chunks & 1) {
    const int j = n_chunks - 1;
    for (int r = 0; r < 3; ++r) add(r * 3 + 0, j, false, r * 3 + 1, j);    // (r,0) + (r,1): second half = pixel to the right
    add(0 * 3 + 2, j, false, 1 * 3 + 2, j);                                // (0,2) + (1,2): second half = pixel below
    add(2 * 3 + 1, j, true, 2 * 3 + 2, j);                                 // zero-weighted (2,1) + (2,2)
  }
  return s;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

thread_local char g_err1[256];

template <int IMG_C, int SEG, int REP, int ROUNDS>
const char* launch_fused(const CUtensorMap& tO, const Conv1FusedParams& p, int grid, size_t smem, cudaStream_t st) {
  auto kfn = conv1_fused_kernel<IMG_C, SEG, REP, ROUNDS>;
  static size_t configured = 0;
  if (configured < smem) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { snprintf(g_err1, sizeof(g_err1), "cudaFuncSetAttribute(conv1_fused, %zu): %s", smem, cudaGetErrorString(e)); return g_err1; }
    configured = smem;
  }
  count_launch();
  kfn<<<grid, kF1Threads, smem, st>>>(tO, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { snprintf(g_err1, sizeof(g_err1), "conv1_fused_kernel launch: %s", cudaGetErrorString(e)); return g_err1; }
  return nullptr;
}

}  // namespace

int conv1_fused_packed_bytes(int cin) {
  const int n_chunks = (cin + 7) / 8;
  return make_schedule(n_chunks).plan.n_mma * 2048;
}

bool conv1_fused_supported(int planes, int img_c, int crop) {
  return crop == kF1Crop && ((planes == 1 && img_c == 3) || (img_c == 1 && planes >= 1 && planes <= 32));
}

cudaError_t launch_pack_conv1_fused_w(const float* w, void* out, int cin, cudaStream_t st) {
  const Conv1Schedule s = make_schedule((cin + 7) / 8);
  const int total = s.plan.n_mma * 2 * 64 * 8;
  count_launch();
  pack_conv1_fused_w_kernel<<<(total + 255) / 256, 256, 0, st>>>(w, static_cast<__nv_bfloat16*>(out), cin, s.plan);
  return cudaGetLastError();
}

const char* conv1_fused_run(const uint8_t* images, size_t image_bytes, int img_h, int img_w, int img_c, const int32_t* table,
                            int n, int planes, const float* mean, const float* stdv, const void* w_fused, const float* bias,
                            void* y, cudaStream_t st) {
  if (n <= 0) return nullptr;
  if (!conv1_fused_supported(planes, img_c, kF1Crop)) return "conv1_fused: unsupported plane / channel combination";
  if (img_h < kF1Crop || img_w < kF1Crop) return "conv1_fused: images smaller than the 224 crop";
  const int cin = planes * img_c;
  Conv1FusedParams p;
  p.images = images; p.image_bytes = image_bytes; p.img_w = img_w; p.table = table; p.n_img = n; p.planes = planes;
  p.div_planes = FastDiv::make((uint32_t)planes);
  p.n_chunks = (cin + 7) / 8;
  p.n_luts = 0;
  for (int i = 0; i < 3; ++i) { p.lut_mean[i] = 0.f; p.lut_std[i] = 1.f; }
  for (int i = 0; i < 32; ++i) p.lut_of[i] = 0;
  for (int c = 0; c < cin; ++c) {
    int k = 0;
    while (k < p.n_luts && !(p.lut_mean[k] == mean[c] && p.lut_std[k] == stdv[c])) ++k;
    if (k == p.n_luts) {
      if (p.n_luts == 3) return "conv1_fused: more than 3 distinct (mean, std) pairs";
      p.lut_mean[k] = mean[c]; p.lut_std[k] = stdv[c]; ++p.n_luts;
    }
    p.lut_of[c] = (unsigned char)k;
  }
  const Conv1Schedule s = make_schedule(p.n_chunks);
  p.n_mma = s.plan.n_mma;
  for (int i = 0; i < kF1MaxMma; ++i) p.a_delta[i] = i < p.n_mma ? s.a_delta[i] : 0ull;
  p.w_packed = static_cast<const __nv_bfloat16*>(w_fused);
  p.bias = bias;
  const int tiles_w = kF1Crop / kF1TileW, tiles_h = kF1Crop / kF1TileH;
  p.total_tiles = n * tiles_w * tiles_h;
  p.div_w = FastDiv::make((uint32_t)tiles_w);
  p.div_h = FastDiv::make((uint32_t)tiles_h);

  EncodeTiledFn fn = encode_fn();
  if (!fn) return "cuTensorMapEncodeTiled not available (no CUDA driver?)";
  CUtensorMap tO;
  {
    cuuint64_t gdim[4] = {64, (cuuint64_t)kF1Crop, (cuuint64_t)kF1Crop, (cuuint64_t)n};
    cuuint64_t gstr[3] = {128, (cuuint64_t)128 * kF1Crop, (cuuint64_t)128 * kF1Crop * kF1Crop};
    cuuint32_t box[4] = {64, (cuuint32_t)kF1TileW, (cuuint32_t)kF1TileH, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = fn(&tO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, y, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { snprintf(g_err1, sizeof(g_err1), "conv1_fused: cuTensorMapEncodeTiled failed (%d)", (int)r); return g_err1; }
  }
  const bool rep32 = img_c == 1 && p.n_luts == 1;
  const int rep = rep32 ? 32 : 1;
  const size_t smem = 1024 + 4 * kF1StagingBytes + (size_t)p.n_mma * 2048 + (size_t)kF1Stages * p.n_chunks * kF1PlaneBytes +
                      (size_t)p.n_luts * 256 * rep * 2 + 64 * 4 + (2 * kF1Stages + 4) * 8 + 16;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = p.total_tiles < sms ? p.total_tiles : sms;
  // ROUNDS = items per gather thread: 18 rows x planes (x 5 two-pixel segments for RGB) over 128 threads
  if (img_c == 3) return launch_fused<3, 2, 1, 1>(tO, p, grid, smem, st);
  if (planes <= 21) return rep32 ? launch_fused<1, 10, 32, 3>(tO, p, grid, smem, st) : launch_fused<1, 10, 1, 3>(tO, p, grid, smem, st);
  return rep32 ? launch_fused<1, 10, 32, 5>(tO, p, grid, smem, st) : launch_fused<1, 10, 1, 5>(tO, p, grid, smem, st);
}

}  // namespace va
