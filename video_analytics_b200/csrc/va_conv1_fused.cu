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

constexpr int kF1GatherWarps = 6;                      // 192 threads >= the 180 pixels of a haloed patch
constexpr int kF1Threads = (8 + kF1GatherWarps + 1) * 32;
constexpr int kF1TileH = 16, kF1TileW = 8, kF1Pitch = kF1TileW + 2, kF1Rows = kF1TileH + 2;
constexpr int kF1Pix = kF1Pitch * kF1Rows;            // 180 haloed pixels
constexpr int kF1PlaneBytes = kF1Pix * 32;            // one plane of a stage: [pixel][16 bf16]
constexpr int kF1Crop = 224;
constexpr int kF1MaxMma = 24;
constexpr int kF1Stages = 4;
constexpr int kF1StagingBytes = 128 * 128;            // 128 pixels x 64 bf16

struct Conv1FusedParams {
  const uint8_t* images;
  unsigned long long image_bytes;
  int img_w;
  const int32_t* table;          // [n][planes][4] = image id, crop top, crop left, flip
  int n_img, planes;
  int n_full;                    // 16-channel planes of a stage
  int has_pair;                  // + one "pair" plane holding the remaining <= 8 channels of pixel p and of pixel p+1
  int n_luts;
  float lut_mean[3], lut_std[3];
  unsigned char lut_of[32];
  const __nv_bfloat16* w_packed; // [n_mma][64][16], 32B-swizzled
  const float* bias;             // [64]
  int total_tiles;
  FastDiv div_w, div_h;          // by 28, 14
  int n_mma;
  unsigned int a_off16[kF1MaxMma];   // per MMA: byte offset of its A window inside a stage, >> 4
  long long* dbg;                    // optional [16] per-role cycle counters written by CTA 0 (diagnostics only)
};

// 32-byte swizzle (Swizzle<1,4,3>): the 16-byte half of a 32-byte row is XORed with address bit 7.  Both TMA and UMMA apply
// it to ABSOLUTE shared-memory address bits (measured: va_conv_tc.cuh HALO), so generic stores that do the same produce an
// operand the tensor core reads from ANY 32-byte-aligned start address.
__device__ __forceinline__ uint32_t swz32(uint32_t addr) { return addr ^ ((addr >> 3) & 0x10u); }

// Explicit shared-state-space loads: pointers derived from the aligned dynamic-smem base are GENERIC to the compiler, and a
// generic LD.E to shared memory (~150 clk dependent latency, measured through the role counters) made the 20 table lookups
// of a flow pixel the kernel's critical path.
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ unsigned long long lds_u64(uint32_t addr) {
  unsigned long long v;
  asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <int IMG_C, int MAXCH, int REP>
__global__ void __launch_bounds__(kF1Threads, 1)
conv1_fused_kernel(const __grid_constant__ CUtensorMap tmO, const Conv1FusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem;                                               // 2 groups x 2 x 16 KB, 1024-aligned (128B swizzle)
  uint8_t* w_s = staging + 4 * kF1StagingBytes;                          // n_mma x 2 KB
  const int n_planes = p.n_full + p.has_pair;
  const uint32_t a_stage_bytes = (uint32_t)n_planes * kF1PlaneBytes;
  uint8_t* a_ring = w_s + (size_t)p.n_mma * 2048;
  uint16_t* lut_s = reinterpret_cast<uint16_t*>(a_ring + (size_t)kF1Stages * a_stage_bytes);
  float* bias_s = reinterpret_cast<float*>(lut_s + (size_t)p.n_luts * 256 * REP);
  unsigned long long* pl_base = reinterpret_cast<unsigned long long*>(bias_s + 64);      // [32] per-plane source base of the snippet
  uint32_t* pl_flip = reinterpret_cast<uint32_t*>(pl_base + 32);                         // bit pl = flip; [1] = snippet id
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(pl_flip + 2);
  uint64_t* empty_bar = full_bar + kF1Stages;
  uint64_t* tfull_bar = empty_bar + kF1Stages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int kGatherWarp0 = 8, kMmaWarp = 8 + kF1GatherWarps;

  // ---- prologue: barriers, TMEM, resident weights, normalisation table, zeroed stages
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmO);
    for (int i = 0; i < kF1Stages; ++i) { mbar_init(&full_bar[i], kF1GatherWarps); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
    fence_mbar_init();
    pl_flip[1] = 0xFFFFFFFFu;
  }
  if (warp == kMmaWarp) { tmem_alloc(tmem_slot, 128); tmem_relinquish(); }
  {
    const uint4* wg = reinterpret_cast<const uint4*>(p.w_packed);
    uint4* ws4 = reinterpret_cast<uint4*>(w_s);
    for (int i = threadIdx.x; i < p.n_mma * 128; i += kF1Threads) ws4[i] = __ldg(wg + i);
    // channel padding, the last pair row and everything else start as zero; only real channels are ever written again
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
  // column blocks of the same source rows, so the gather's loads hit L1, and the per-plane source bases change once per snippet.
  const int tiles_per = p.total_tiles / (int)gridDim.x, tiles_rem = p.total_tiles % (int)gridDim.x;
  const int tile_begin = (int)blockIdx.x * tiles_per + min((int)blockIdx.x, tiles_rem);
  const int tile_end = tile_begin + tiles_per + ((int)blockIdx.x < tiles_rem ? 1 : 0);

  if (warp >= kGatherWarp0 && warp < kGatherWarp0 + kF1GatherWarps) {
    // ===================================================== gather + normalise -> A operand (6 warps, thread = patch pixel)
    // Thread t < 180 owns pixel (row t/10, col t%10) of the haloed patch and loops over the planes: a warp-level byte load
    // then touches ~4 cache lines (3 patch rows of one plane), not 32 as with one (plane, row) per lane.  The bytes of the
    // NEXT tile are loaded before the current tile is converted (two register sets).
    const int gt = threadIdx.x - kGatherWarp0 * 32;
    const bool has_px = gt < kF1Pix;
    const int prow = has_px ? gt / kF1Pitch : 0, pcol = has_px ? gt % kF1Pitch : 0;
    const int4* table4 = reinterpret_cast<const int4*>(p.table);
    const int nch = p.planes * IMG_C;
    uint32_t ua[MAXCH], ub[MAXCH];
    bool va_ok = false, vb_ok = false;
    const uint32_t lut_u32 = smem_u32(lut_s), plb_u32 = smem_u32(pl_base), plf_u32 = smem_u32(pl_flip);

    auto load_tile = [&](int tile, uint32_t (&u)[MAXCH], bool& ok) {
      uint32_t mt, tw, th, tn;
      p.div_w.divmod((uint32_t)tile, mt, tw);
      p.div_h.divmod(mt, tn, th);
      if (lds_u32(plf_u32 + 4) != tn) {                // first tile of a snippet (same decision in all gather threads)
        named_bar_sync(5, kF1GatherWarps * 32);        // nobody still reads the previous snippet's entries
        if (gt < p.planes) {
          const int4 e = __ldg(table4 + (size_t)tn * p.planes + gt);   // image id, crop top, crop left, flip
          pl_base[gt] = (unsigned long long)e.x * p.image_bytes + ((unsigned long long)e.y * p.img_w + e.z) * IMG_C;
          if (e.w) atomicOr(&pl_flip[0], 1u << gt); else atomicAnd(&pl_flip[0], ~(1u << gt));
        }
        if (gt == 0) pl_flip[1] = tn;
        named_bar_sync(5, kF1GatherWarps * 32);
      }
      const int y = (int)th * kF1TileH - 1 + prow, x = (int)tw * kF1TileW - 1 + pcol;
      ok = has_px && (unsigned)y < (unsigned)kF1Crop && (unsigned)x < (unsigned)kF1Crop;   // else: the conv's zero padding
      if (!ok) return;
      const uint32_t flips = lds_u32(plf_u32);
      const uint32_t off_n = (uint32_t)(y * p.img_w + x) * IMG_C, off_f = (uint32_t)(y * p.img_w + (kF1Crop - 1 - x)) * IMG_C;
#pragma unroll
      for (int pl = 0; pl < MAXCH / IMG_C; ++pl) {
        if (pl < p.planes) {
          const uint8_t* src = p.images + lds_u64(plb_u32 + 8 * pl) + (((flips >> pl) & 1u) ? off_f : off_n);   // hflip == reversed columns
#pragma unroll
          for (int k = 0; k < IMG_C; ++k) u[pl * IMG_C + k] = __ldg(src + k);
        }
      }
    };
    auto convert_tile = [&](uint32_t a_dst, const uint32_t (&u)[MAXCH], bool ok) {
      if (!has_px) return;
      uint32_t pk[MAXCH / 2];                          // bf16 pairs, channel order
#pragma unroll
      for (int c = 0; c < MAXCH; c += 2) {
        uint32_t lo = 0, hi = 0;
        if (ok && c < nch) {
          const int li = (REP == 32 || p.n_luts == 1) ? 0 : (int)p.lut_of[c];
          lo = lds_u16(lut_u32 + (uint32_t)(((li * 256 + (int)u[c]) * REP + (REP == 32 ? lane : 0)) * 2));
        }
        if (ok && c + 1 < nch) {
          const int li = (REP == 32 || p.n_luts == 1) ? 0 : (int)p.lut_of[c + 1];
          hi = lds_u16(lut_u32 + (uint32_t)(((li * 256 + (int)u[c + 1]) * REP + (REP == 32 ? lane : 0)) * 2));
        }
        pk[c / 2] = lo | (hi << 16);
      }
      const uint32_t row_addr = a_dst + (uint32_t)gt * 32u;
#pragma unroll
      for (int f = 0; f < MAXCH / 16; ++f) {           // full 16-channel planes: row p = channels 16f .. 16f+15 of pixel p
        if (f < p.n_full) {
          const uint32_t a = row_addr + (uint32_t)f * kF1PlaneBytes;
          sts128(swz32(a), make_uint4(pk[8 * f], pk[8 * f + 1], pk[8 * f + 2], pk[8 * f + 3]));
          sts128(swz32(a + 16), make_uint4(pk[8 * f + 4], pk[8 * f + 5], pk[8 * f + 6], pk[8 * f + 7]));
        }
      }
      if (p.has_pair) {                                // pair plane: row p = [remaining channels of pixel p | of pixel p+1]
        uint4 g = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int f = 0; f < MAXCH / 16; ++f)
          if (f == p.n_full) g = make_uint4(pk[8 * f], pk[8 * f + 1], pk[8 * f + 2], pk[8 * f + 3]);
        const uint32_t a = row_addr + (uint32_t)p.n_full * kF1PlaneBytes;
        sts128(swz32(a), g);
        if (gt > 0) sts128(swz32(a - 32 + 16), g);
      }
    };

    const uint32_t ring_u32 = smem_u32(a_ring);
    uint32_t stage = 0, phase = 0;
    int tile = tile_begin;
    const bool dbg = p.dbg != nullptr && blockIdx.x == 0 && gt == 0;
    long long t_load = 0, t_wait = 0, t_conv = 0, t_begin = clock64(), tq = 0;
    if (tile < tile_end) load_tile(tile, ua, va_ok);
    // two tiles per iteration so that the register sets alternate without copies
    while (tile < tile_end) {
      {
        const int next = tile + 1;
        if (dbg) tq = clock64();
        mbar_wait(&empty_bar[stage], phase ^ 1, 100 + stage);
        if (dbg) { const long long t = clock64(); t_wait += t - tq; tq = t; }
        convert_tile(ring_u32 + stage * a_stage_bytes, ua, va_ok);
        // generic-proxy stores -> visible to the tensor core's async-proxy reads.  ptxas implements this fence with a
        // MEMBAR.ALL.CTA, which also waits for every global load in flight: the next tile's loads are therefore issued
        // AFTER it (issued before, the fence exposed their whole L2 latency on every tile).
        fence_proxy_async_smem();
        if (dbg) { const long long t = clock64(); t_conv += t - tq; tq = t; }
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[stage]);
        if (++stage == kF1Stages) { stage = 0; phase ^= 1; }
        if (next < tile_end) load_tile(next, ub, vb_ok);
        if (dbg) { const long long t = clock64(); t_load += t - tq; tq = t; }
        tile = next;
      }
      if (tile >= tile_end) break;
      {
        const int next = tile + 1;
        if (dbg) tq = clock64();
        mbar_wait(&empty_bar[stage], phase ^ 1, 100 + stage);
        if (dbg) { const long long t = clock64(); t_wait += t - tq; tq = t; }
        convert_tile(ring_u32 + stage * a_stage_bytes, ub, vb_ok);
        fence_proxy_async_smem();
        if (dbg) { const long long t = clock64(); t_conv += t - tq; tq = t; }
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[stage]);
        if (++stage == kF1Stages) { stage = 0; phase ^= 1; }
        if (next < tile_end) load_tile(next, ua, va_ok);
        if (dbg) { const long long t = clock64(); t_load += t - tq; tq = t; }
        tile = next;
      }
    }
    if (dbg) { p.dbg[0] = clock64() - t_begin; p.dbg[1] = t_load; p.dbg[2] = t_wait; p.dbg[3] = t_conv; p.dbg[11] = tile_end - tile_begin; }
  } else if (warp == kMmaWarp) {
    // ===================================================== MMA issuer (convergent warp, one elected lane issues)
    constexpr uint32_t idesc = make_idesc_bf16(128, 64);
    const uint32_t a_ring_u32 = smem_u32(a_ring);
    const uint32_t w_u32 = smem_u32(w_s);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    // A: 32-byte rows (one pixel x 16 channels), 8-pixel groups (one output row of the tile) 10 pixels = 320 B apart
    const uint64_t a_sbo_fix = ((uint64_t)((kF1Pitch * 32) >> 4) << 32) - ((uint64_t)((32 * 8) >> 4) << 32);
    uint32_t stage = 0, phase = 0, as = 0, as_phase = 0;
    const bool dbg = p.dbg != nullptr && blockIdx.x == 0;
    long long t_te = 0, t_full = 0, t_begin = clock64();
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      long long tq = dbg ? clock64() : 0;
      mbar_wait(&tempty_bar[as], as_phase ^ 1, 200 + as);
      if (dbg) { const long long t = clock64(); t_te += t - tq; tq = t; }
      mbar_wait(&full_bar[stage], phase, 300 + stage);
      if (dbg) t_full += clock64() - tq;
      tc_fence_after();
      const uint64_t da0 = make_smem_desc<32>(a_ring_u32 + stage * a_stage_bytes) + a_sbo_fix;
      const uint64_t db0 = make_smem_desc<32>(w_u32);
      const uint32_t d_tmem = tmem_u + as * 64;
      if (elect_one()) {
        for (int i = 0; i < p.n_mma; ++i) umma_bf16(d_tmem, da0 + p.a_off16[i], db0 + (uint64_t)(i * (2048 >> 4)), idesc, i ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        umma_commit(&tfull_bar[as]);
      }
      __syncwarp();
      if (++stage == kF1Stages) { stage = 0; phase ^= 1; }
      as ^= 1;
      if (as == 0) as_phase ^= 1;
    }
    if (dbg && lane == 0) { p.dbg[4] = clock64() - t_begin; p.dbg[5] = t_te; p.dbg[6] = t_full; }
  } else if (warp < 8) {
    // ===================================================== epilogue (2 groups x 4 warps)
    const int eg = warp >> 2;
    const int q = warp & 3;
    const int m = q * 32 + lane;                  // accumulator row == pixel h_i*8 + w_i of the tile
    const int et = threadIdx.x - eg * 128;
    // two staging buffers per group: a TMA store holds its buffer until the store engine has read it
    uint8_t* stage_base = staging + eg * 2 * kF1StagingBytes;
    const uint32_t as = (uint32_t)eg;
    int it = 0;
    const bool dbg = p.dbg != nullptr && blockIdx.x == 0 && et == 0 && eg == 0;
    long long t_tf = 0, t_st = 0, t_begin = clock64();
    for (int tile = tile_begin; tile < tile_end; ++tile, ++it) {
      if ((it & 1) != eg) continue;
      const uint32_t as_phase = (uint32_t)(it >> 1) & 1u;
      uint32_t mt, tw, th, tn;
      p.div_w.divmod((uint32_t)tile, mt, tw);
      p.div_h.divmod(mt, tn, th);
      long long tq = dbg ? clock64() : 0;
      mbar_wait(&tfull_bar[as], as_phase, 400 + as);
      if (dbg) t_tf += clock64() - tq;
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
      tq = dbg ? clock64() : 0;
      if (et < 32) {                       // this buffer was last read by the group's store of two tiles ago
        if (elect_one()) tma_store_wait_read<1>();
        __syncwarp();
      }
      named_bar_sync(1 + eg, 128);
      if (dbg) t_st += clock64() - tq;
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
    if (dbg) { p.dbg[7] = clock64() - t_begin; p.dbg[8] = t_tf; p.dbg[9] = t_st; }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, 128);
}

// ---- weight packing for the K schedule ---------------------------------------------------------------------------
struct Conv1PackPlan {
  int n_mma;
  signed char tap[kF1MaxMma][2];     // per 8-channel half of the MMA's K = 16: tap r*3+s, or -1 = zero weights
  signed char ch0[kF1MaxMma][2];     // first input channel of that half
};

__global__ void pack_conv1_fused_w_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int cin,
                                          const Conv1PackPlan plan) {
  // out [n_mma][64 rows n][16 k] in the 32B-swizzled layout the tensor core reads; w OIHW [64][cin][3][3]
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= plan.n_mma * 64 * 16) return;
  const int k = idx & 15, n = (idx >> 4) & 63, i = idx >> 10;
  const int half = k >> 3, e = k & 7;
  const int tap = plan.tap[i][half], ch = plan.ch0[i][half] + e;
  float v = 0.f;
  if (tap >= 0 && ch < cin) v = w[((size_t)n * cin + ch) * 9 + tap];
  uint32_t off = (uint32_t)(n * 32 + half * 16);           // bytes inside the MMA's 2 KB block (2 KB-aligned in smem)
  off ^= (off >> 3) & 0x10u;
  out[(size_t)i * 1024 + off / 2 + e] = __float2bfloat16_rn(v);
}

struct Conv1Schedule {
  Conv1PackPlan plan;
  unsigned int a_off16[kF1MaxMma];
  int n_full, has_pair;
};

Conv1Schedule make_schedule(int cin) {
  Conv1Schedule s;
  s.plan.n_mma = 0;
  const int rem = cin % 16;
  s.n_full = cin / 16 + (rem > 8 ? 1 : 0);
  s.has_pair = (rem > 0 && rem <= 8) ? 1 : 0;
  auto shift = [](int r, int c) { return r * kF1Pitch + c; };
  auto add = [&](int plane, int px_shift, int tap0, int ch0, int tap1, int ch1) {
    const int i = s.plan.n_mma++;
    s.plan.tap[i][0] = (signed char)tap0; s.plan.ch0[i][0] = (signed char)ch0;
    s.plan.tap[i][1] = (signed char)tap1; s.plan.ch0[i][1] = (signed char)ch1;
    s.a_off16[i] = (unsigned int)((plane * kF1PlaneBytes + px_shift * 32) >> 4);
  };
  for (int f = 0; f < s.n_full; ++f)
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) add(f, shift(r, c), r * 3 + c, 16 * f, r * 3 + c, 16 * f + 8);
  if (s.has_pair) {
    const int c0 = 16 * s.n_full;
    for (int r = 0; r < 3; ++r) {
      add(s.n_full, shift(r, 0), r * 3 + 0, c0, r * 3 + 1, c0);      // [pixel p | pixel p+1] x taps (r,0), (r,1)
      add(s.n_full, shift(r, 2), r * 3 + 2, c0, -1, c0);             // [pixel p+2 | pixel p+3] x tap (r,2), zero weights
    }
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

template <int IMG_C, int MAXCH, int REP>
const char* launch_fused(const CUtensorMap& tO, const Conv1FusedParams& p, int grid, size_t smem, cudaStream_t st) {
  auto kfn = conv1_fused_kernel<IMG_C, MAXCH, REP>;
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

int conv1_fused_packed_bytes(int cin) { return make_schedule(cin).plan.n_mma * 2048; }

bool conv1_fused_supported(int planes, int img_c, int crop) {
  return crop == kF1Crop && ((planes == 1 && img_c == 3) || (img_c == 1 && planes >= 1 && planes <= 32));
}

cudaError_t launch_pack_conv1_fused_w(const float* w, void* out, int cin, cudaStream_t st) {
  const Conv1Schedule s = make_schedule(cin);
  const int total = s.plan.n_mma * 64 * 16;
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
  const Conv1Schedule s = make_schedule(cin);
  p.n_mma = s.plan.n_mma;
  p.n_full = s.n_full; p.has_pair = s.has_pair;
  for (int i = 0; i < kF1MaxMma; ++i) p.a_off16[i] = i < p.n_mma ? s.a_off16[i] : 0u;
  p.w_packed = static_cast<const __nv_bfloat16*>(w_fused);
  p.bias = bias;
  p.dbg = conv_get_debug_counters();
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
  const size_t smem = 1024 + 4 * kF1StagingBytes + (size_t)p.n_mma * 2048 + (size_t)kF1Stages * (p.n_full + p.has_pair) * kF1PlaneBytes +
                      (size_t)p.n_luts * 256 * rep * 2 + 64 * 4 + 32 * 8 + 8 + (2 * kF1Stages + 4) * 8 + 16;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = p.total_tiles < sms ? p.total_tiles : sms;
  // MAXCH = channel registers per gather thread (multiple of 16)
  if (img_c == 3) return launch_fused<3, 16, 1>(tO, p, grid, smem, st);
  if (planes <= 16) return rep32 ? launch_fused<1, 16, 32>(tO, p, grid, smem, st) : launch_fused<1, 16, 1>(tO, p, grid, smem, st);
  return rep32 ? launch_fused<1, 32, 32>(tO, p, grid, smem, st) : launch_fused<1, 32, 1>(tO, p, grid, smem, st);
}

}  // namespace va
