// K1+K2 fused: snippet preprocess gathered straight into conv1_1's shared-memory operand.
//
// Replaces, in one kernel, the reference's per-item transform pipeline (Sheet03/utils.py:137-151: RandomCrop /
// RandomHorizontalFlip / ToTensor / Normalize; per-image for the 20 flow planes, temporalModel.py:80-90) AND the first
// VGG16 layer `features[0:2]` = Conv2d(C0, 64, 3, padding=1) + ReLU (spatialModel.py:110,171; temporalModel.py:149-162).
// The separate K1 kernel wrote a channel-padded bf16 NHWC tensor (1.6 / 3.2 MB per snippet) that conv1_1 read back;
// here the u8 crop is the only HBM input of the layer.
//
// Work unit = a SEGMENT of 7 horizontally adjacent 16 x 8 output tiles of one snippet (a quarter of a tile row); every
// CTA walks a contiguous range of segments.  Per segment:
//   * LOADER warp: for every plane (index-table row = image id, crop top/left, flip) and each of the 18 haloed source
//     rows, ONE bulk async copy (cp.async.bulk, completion on an mbarrier) brings the 16-byte blocks covering the 58 source
//     pixels the segment touches into one of two raw strip buffers, and the byte offset of the segment's first pixel inside
//     the copied row goes into a small table.  No registers, no proxy fence in this warp: its copies stay in flight across
//     tiles (versions 1-4 loaded bytes into registers of the converting warps: ptxas implements fence.proxy.async with
//     MEMBAR.ALL.CTA, which drained those loads at every tile and exposed ~1000 clk of load latency per tile,
//     profiles/r02_conv1_fused_role_counters.log).
//   * CONVERTER warps (thread = pixel of the 18 x 10 haloed patch of a tile): byte from the strip -> normalisation table
//     (bf16(((u8/255) - mean)/std), the same IEEE operations as torchvision, built per CTA) -> operand rows
//     [pixel][16 channels] in the 32-byte-swizzle layout (the XOR is applied to absolute shared-address bits, as TMA and
//     UMMA do), zero outside the crop (= the convolution's padding).  Channel remainders <= 8 go to a "pair" plane
//     [pixel p | pixel p+1] so that two taps share one K = 16 MMA: 3 channels need 6 MMAs per tile instead of 9, 20
//     channels 15 instead of 18.  Filter tap (r,s) is a start-address shift of (r*10 + s)*32 B, 8-pixel groups are
//     320 B apart (SBO).  Weights are packed to match and stay resident in smem.
//   * MMA warp (one elected lane) and two epilogue groups (TMEM -> +bias -> ReLU -> bf16 -> 128B-swizzled staging -> TMA
//     store) as in va_conv_tc.cuh.
// Warp roles (512 threads, 1 CTA/SM, persistent): warps 0-7 epilogue, 8-13 converters, 14 loader, 15 MMA + TMEM owner.
#include "va_internal.h"
#include "va_conv_tc.cuh"

#include <mutex>
#include <stdio.h>

namespace va {

namespace {

constexpr int kF1GatherWarps = 6;                      // converter warps: 192 threads >= the 180 pixels of a haloed patch
constexpr int kF1Threads = (8 + kF1GatherWarps + 2) * 32;
constexpr int kF1SegTiles = 7;                         // tiles per segment (28 tiles per tile row = 4 segments)
constexpr int kF1SegPerSnip = 14 * 4;
constexpr int kF1SegPix = kF1SegTiles * 8 + 2;         // 58 source pixels per row of a segment
// one bulk copy per (plane, haloed row): the 16-byte blocks covering 58 pixels at any byte shift, rows bank-staggered
__host__ __device__ constexpr int f1_copy_bytes(int img_c) { return (15 + kF1SegPix * img_c + 15) / 16 * 16; }   // 80 / 192
__host__ __device__ constexpr int f1_row_pitch(int img_c) { return img_c == 1 ? 80 : 208; }
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
  int total_segments;            // n_img * 14 * 4
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

__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ int lds_s32(uint32_t addr) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void f1_bulk_copy(uint32_t smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// NCH_CT: the channel count as a compile-time constant (3 = RGB frame, 20 = the 10-pair flow stack; 0 = read it from the
// parameters).  With it the converter's per-channel slots beyond the real channels disappear instead of being issued
// predicated-off: 190 -> ~60 instructions per RGB pixel, and the tile time of a converter warp is its dependent latency.
template <int IMG_C, int MAXCH, int REP, int NCH_CT>
__global__ void __launch_bounds__(kF1Threads, 1)
conv1_fused_kernel(const __grid_constant__ CUtensorMap tmO, const Conv1FusedParams p) {
  constexpr int kPitch = f1_row_pitch(IMG_C);
  constexpr int kCopy = f1_copy_bytes(IMG_C);
  constexpr int kMaxPlanes = MAXCH / IMG_C;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem;                                               // 2 groups x 2 x 16 KB, 1024-aligned (128B swizzle)
  uint8_t* w_s = staging + 4 * kF1StagingBytes;                          // n_mma x 2 KB
  const int n_full = NCH_CT ? (NCH_CT / 16 + ((NCH_CT % 16) > 8 ? 1 : 0)) : p.n_full;
  const int has_pair = NCH_CT ? (((NCH_CT % 16) > 0 && (NCH_CT % 16) <= 8) ? 1 : 0) : p.has_pair;
  const int n_planes = n_full + has_pair;
  const uint32_t a_stage_bytes = (uint32_t)n_planes * kF1PlaneBytes;
  uint8_t* a_ring = w_s + (size_t)p.n_mma * 2048;
  uint16_t* lut_s = reinterpret_cast<uint16_t*>(a_ring + (size_t)kF1Stages * a_stage_bytes);
  float* bias_s = reinterpret_cast<float*>(lut_s + (size_t)p.n_luts * 256 * REP);
  const int units = p.planes * kF1Rows;                                  // (plane, haloed row) copies per segment
  const uint32_t strip_bytes = (uint32_t)units * kPitch;
  uint8_t* strips = reinterpret_cast<uint8_t*>(bias_s + 64);             // 2 raw strip buffers (16-byte aligned)
  int* rowoff = reinterpret_cast<int*>(strips + 2 * (size_t)strip_bytes);   // [2][units]: byte of the segment's first pixel in its row
  uint32_t* flipmask = reinterpret_cast<uint32_t*>(rowoff + 2 * units);  // [2]: bit pl = plane pl is flipped
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(flipmask + 2);
  uint64_t* empty_bar = full_bar + kF1Stages;
  uint64_t* tfull_bar = empty_bar + kF1Stages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* sfull_bar = tempty_bar + 2;                                   // strip loaded (tx bytes + one arrival per unit)
  uint64_t* sempty_bar = sfull_bar + 2;                                   // strip consumed by the 6 converter warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int kGatherWarp0 = 8, kLoaderWarp = 8 + kF1GatherWarps, kMmaWarp = kLoaderWarp + 1;

  // ---- prologue: barriers, TMEM, resident weights, normalisation table, zeroed stages
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmO);
    for (int i = 0; i < kF1Stages; ++i) { mbar_init(&full_bar[i], kF1GatherWarps); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4);
      mbar_init(&sfull_bar[i], 32 + 1); mbar_init(&sempty_bar[i], kF1GatherWarps);
    }
    fence_mbar_init();
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
  // Every CTA walks a CONTIGUOUS range of segments (row-major inside a snippet).
  const int seg_per = p.total_segments / (int)gridDim.x, seg_rem = p.total_segments % (int)gridDim.x;
  const int seg_begin = (int)blockIdx.x * seg_per + min((int)blockIdx.x, seg_rem);
  const int seg_end = seg_begin + seg_per + ((int)blockIdx.x < seg_rem ? 1 : 0);
  const int n_tiles = (seg_end - seg_begin) * kF1SegTiles;

  if (warp == kLoaderWarp) {
    // ===================================================== loader: global -> raw strips, one bulk copy per (plane, row)
    const int4* table4 = reinterpret_cast<const int4*>(p.table);
    const uint32_t strips_u32 = smem_u32(strips);
    constexpr int kChunks = kCopy / 16;                          // 5 / 12
    int segk = 0, cur_tn = -1;
    int4 ent = make_int4(0, 0, 0, 0);
    for (int seg = seg_begin; seg < seg_end; ++seg, ++segk) {
      const int buf = segk & 1;
      const uint32_t sphase = (uint32_t)(segk >> 1) & 1u;
      mbar_wait(&sempty_bar[buf], sphase ^ 1u, 500 + buf);       // the converters are done with this buffer's previous segment
      const int tn = seg / kF1SegPerSnip, rem = seg - tn * kF1SegPerSnip;
      const int th = rem >> 2, sq = rem & 3;
      const int y_first = th * kF1TileH - 1, x_first = sq * kF1SegTiles * kF1TileW - 1;
      if (tn != cur_tn) {                                          // lane pl keeps the table row of plane pl of this snippet
        cur_tn = tn;
        ent = lane < p.planes ? __ldg(table4 + (size_t)tn * p.planes + lane) : make_int4(0, 0, 0, 0);
      }
      {
        const uint32_t mask = __ballot_sync(0xffffffffu, lane < p.planes && ent.w != 0);
        if (lane == 0) flipmask[buf] = mask;
      }
      // 16-byte asynchronous copies (LDGSTS, L1 bypass), lane = one (plane, row): the row's address is computed once and
      // its 5 / 12 chunks are issued back to back.  (One cp.async.bulk per row -- 360 TMA requests per flow segment -- and,
      // after that, lane = chunk with the address arithmetic repeated per chunk both kept the loader behind the MMAs:
      // 1340 / 920 clk of strip wait per tile.)
      for (int unit0 = 0; unit0 < units; unit0 += 32) {
        const int unit = min(unit0 + lane, units - 1);
        const int pl = (unit * 3641) >> 16;                      // unit / 18 (exact below 4000)
        const int r = unit - pl * kF1Rows;
        const int y = y_first + r;
        int4 e;                                                  // image id, crop top, crop left, flip
        e.x = __shfl_sync(0xffffffffu, ent.x, pl); e.y = __shfl_sync(0xffffffffu, ent.y, pl);
        e.z = __shfl_sync(0xffffffffu, ent.z, pl); e.w = __shfl_sync(0xffffffffu, ent.w, pl);
        if (unit0 + lane < units && (unsigned)y < (unsigned)kF1Crop) {   // rows above / below the crop: zero padding
          // lowest source column of the segment: crop column x_first, or 223 - (x_first + 57) when the plane is flipped
          const int col_lo = e.z + (e.w ? (kF1Crop - 1) - (x_first + kF1SegPix - 1) : x_first);
          const long long img_lo = (long long)e.x * (long long)p.image_bytes;
          const long long off = img_lo + ((long long)(e.y + y) * p.img_w + col_lo) * IMG_C;
          long long start = off & ~15ll;
          if (start < img_lo) start = img_lo;                    // (only pixels outside the crop lie before / after the image)
          const long long img_hi = img_lo + (long long)p.image_bytes;
          if (start + kCopy > img_hi) start = img_hi - kCopy;
          // byte of segment pixel 0 inside the copied row; a flipped plane walks backwards from its last pixel
          rowoff[buf * units + unit] = (int)(off - start) + (e.w ? (kF1SegPix - 1) * IMG_C : 0);
          const uint32_t dst = strips_u32 + (uint32_t)buf * strip_bytes + (uint32_t)unit * kPitch;
          const uint8_t* src_g = p.images + start;
#pragma unroll
          for (int j = 0; j < kChunks; ++j)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * j), "l"(src_g + 16 * j) : "memory");
        }
      }
      // every lane's copies report to the strip's barrier when they land; lane 0 adds a normal (releasing) arrival after
      // the warp's row-offset / flip-mask stores
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&sfull_bar[buf])) : "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&sfull_bar[buf]);
    }
  } else if (warp >= kGatherWarp0 && warp < kGatherWarp0 + kF1GatherWarps) {
    // ===================================================== strip bytes -> normalise -> A operand (6 warps, thread = patch pixel)
    const int gt = threadIdx.x - kGatherWarp0 * 32;
    const bool has_px = gt < kF1Pix;
    const int prow = has_px ? gt / kF1Pitch : 0, pcol = has_px ? gt % kF1Pitch : 0;
    const int nch = NCH_CT ? NCH_CT : p.planes * IMG_C;
    const uint32_t lut_u32 = smem_u32(lut_s);
    const uint32_t strips_u32 = smem_u32(strips), rowoff_u32 = smem_u32(rowoff), flip_u32 = smem_u32(flipmask);
    const uint32_t ring_u32 = smem_u32(a_ring);
    uint32_t stage = 0, phase = 0;
    const bool dbg = p.dbg != nullptr && blockIdx.x == 0 && gt == 0;
    long long t_strip = 0, t_wait = 0, t_conv = 0, t_fence = 0, t_begin = clock64(), tq = 0;
    int segk = 0;
    for (int seg = seg_begin; seg < seg_end; ++seg, ++segk) {
      const int buf = segk & 1;
      const uint32_t sphase = (uint32_t)(segk >> 1) & 1u;
      const int tn = seg / kF1SegPerSnip, rem = seg - tn * kF1SegPerSnip;
      const int th = rem >> 2, sq = rem & 3;
      if (dbg) tq = clock64();
      mbar_wait(&sfull_bar[buf], sphase, 600 + buf);
      if (dbg) t_strip += clock64() - tq;
      // per-segment constants of this pixel: its source row in every plane
      const int y = th * kF1TileH - 1 + prow;
      const bool yok = has_px && (unsigned)y < (unsigned)kF1Crop;
      const uint32_t flips = lds_u32(flip_u32 + 4u * buf);
      const uint32_t all_planes = p.planes >= 32 ? 0xffffffffu : ((1u << p.planes) - 1u);
      uint32_t src[kMaxPlanes];                                  // shared address of segment pixel 0 of this row, per plane
#pragma unroll
      for (int pl = 0; pl < kMaxPlanes; ++pl) {
        src[pl] = 0;
        if (yok && pl < p.planes) {
          const int unit = pl * kF1Rows + prow;
          src[pl] = strips_u32 + (uint32_t)buf * strip_bytes + (uint32_t)unit * kPitch +
                    (uint32_t)lds_s32(rowoff_u32 + 4u * (uint32_t)(buf * units + unit));
        }
      }
      for (int t = 0; t < kF1SegTiles; ++t) {
        const int rel = t * kF1TileW + pcol;                     // pixel of the segment's 58-pixel row
        const int x = sq * kF1SegTiles * kF1TileW - 1 + rel;
        const bool ok = yok && (unsigned)x < (unsigned)kF1Crop;   // else: the conv's zero padding
        if (dbg) tq = clock64();
        mbar_wait(&empty_bar[stage], phase ^ 1, 100 + stage);
        if (dbg) { const long long tt = clock64(); t_wait += tt - tq; tq = tt; }
        if (has_px) {
          // two passes so that the loads of a pass are independent and issue back to back: source bytes, then table values.
          // ONE branch per pixel (outside the crop -> zeros; divergent only in border tiles), and warp-uniform fast paths
          // for the common flip patterns -- none / all planes, the 25 x 10 protocol -- so that the per-channel work is a byte
          // load, an index and a table load (a branch per channel around the volatile loads cost ~50 clk each).
          uint32_t pk[MAXCH / 2];                                // bf16 pairs, channel order
#pragma unroll
          for (int c = 0; c < MAXCH / 2; ++c) pk[c] = 0u;
          if (ok) {
            uint32_t u[MAXCH];
            const int relb = rel * IMG_C;
            if (flips == 0u) {
#pragma unroll
              for (int ch = 0; ch < MAXCH; ++ch)
                if (ch < nch) u[ch] = lds_u8(src[ch / IMG_C] + (uint32_t)(relb + ch % IMG_C));
            } else if (flips == all_planes) {
#pragma unroll
              for (int ch = 0; ch < MAXCH; ++ch)
                if (ch < nch) u[ch] = lds_u8(src[ch / IMG_C] + (uint32_t)(ch % IMG_C - relb));
            } else {
#pragma unroll
              for (int ch = 0; ch < MAXCH; ++ch)
                if (ch < nch) {
                  const int pl = ch / IMG_C;
                  u[ch] = lds_u8(src[pl] + (uint32_t)((((flips >> pl) & 1u) ? -relb : relb) + ch % IMG_C));
                }
            }
#pragma unroll
            for (int c = 0; c < MAXCH; c += 2) {
              uint32_t lo = 0, hi = 0;
              if (c < nch) {
                const int li = (REP == 32 || p.n_luts == 1) ? 0 : (int)p.lut_of[c];
                lo = lds_u16(lut_u32 + (uint32_t)(((li * 256 + (int)u[c]) * REP + (REP == 32 ? lane : 0)) * 2));
              }
              if (c + 1 < nch) {
                const int li = (REP == 32 || p.n_luts == 1) ? 0 : (int)p.lut_of[c + 1];
                hi = lds_u16(lut_u32 + (uint32_t)(((li * 256 + (int)u[c + 1]) * REP + (REP == 32 ? lane : 0)) * 2));
              }
              pk[c / 2] = lo | (hi << 16);
            }
          }
          const uint32_t row_addr = ring_u32 + stage * a_stage_bytes + (uint32_t)gt * 32u;
#pragma unroll
          for (int f = 0; f < MAXCH / 16; ++f) {         // full 16-channel planes: row p = channels 16f .. 16f+15 of pixel p
            if (f < n_full) {
              const uint32_t a = row_addr + (uint32_t)f * kF1PlaneBytes;
              sts128(swz32(a), make_uint4(pk[8 * f], pk[8 * f + 1], pk[8 * f + 2], pk[8 * f + 3]));
              sts128(swz32(a + 16), make_uint4(pk[8 * f + 4], pk[8 * f + 5], pk[8 * f + 6], pk[8 * f + 7]));
            }
          }
          if (has_pair) {                                // pair plane: row p = [remaining channels of pixel p | of pixel p+1]
            uint4 g = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int f = 0; f < MAXCH / 16; ++f)
              if (f == n_full) g = make_uint4(pk[8 * f], pk[8 * f + 1], pk[8 * f + 2], pk[8 * f + 3]);
            const uint32_t a = row_addr + (uint32_t)n_full * kF1PlaneBytes;
            sts128(swz32(a), g);
            if (gt > 0) sts128(swz32(a - 32 + 16), g);
          }
        }
        if (dbg) { const long long tt = clock64(); t_conv += tt - tq; tq = tt; }
        fence_proxy_async_smem();      // generic-proxy stores -> visible to the tensor core's async-proxy reads
        if (dbg) t_fence += clock64() - tq;
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[stage]);
        if (++stage == kF1Stages) { stage = 0; phase ^= 1; }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sempty_bar[buf]);      // this warp no longer reads the strip
    }
    if (dbg) { p.dbg[0] = clock64() - t_begin; p.dbg[1] = t_strip; p.dbg[2] = t_wait; p.dbg[3] = t_conv; p.dbg[12] = t_fence; p.dbg[11] = n_tiles; }
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
    for (int tile = 0; tile < n_tiles; ++tile) {
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
    const bool dbg = p.dbg != nullptr && blockIdx.x == 0 && et == 0 && eg == 0;
    long long t_tf = 0, t_st = 0, t_begin = clock64();
    for (int it = eg; it < n_tiles; it += 2) {
      const uint32_t as_phase = (uint32_t)(it >> 1) & 1u;
      const int seg = seg_begin + it / kF1SegTiles, t = it % kF1SegTiles;
      const int tn = seg / kF1SegPerSnip, rem = seg - tn * kF1SegPerSnip;
      const int th = rem >> 2, tw = (rem & 3) * kF1SegTiles + t;
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
          tma_store_4d(&tmO, stage_out, 0, tw * kF1TileW, th * kF1TileH, tn);
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

template <int IMG_C, int MAXCH, int REP, int NCH_CT>
const char* launch_fused(const CUtensorMap& tO, const Conv1FusedParams& p, int grid, size_t smem, cudaStream_t st) {
  auto kfn = conv1_fused_kernel<IMG_C, MAXCH, REP, NCH_CT>;
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
  // 1-channel stacks: two raw strip buffers of planes x 18 rows x 80 B must fit beside the operand ring (23 planes: 229.7 KB)
  return crop == kF1Crop && ((planes == 1 && img_c == 3) || (img_c == 1 && planes >= 1 && planes <= 23));
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
  p.total_segments = n * kF1SegPerSnip;
  // the loader copies 16-byte blocks clamped to the image: image bases and sizes must keep that alignment
  if ((image_bytes & 15) != 0 || (reinterpret_cast<uintptr_t>(images) & 15) != 0 || image_bytes < (size_t)f1_copy_bytes(img_c))
    return "conv1_fused: the image store must be 16-byte aligned with a 16-byte multiple per image";

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
  const size_t units = (size_t)planes * kF1Rows;
  const size_t smem = 1024 + 4 * kF1StagingBytes + (size_t)p.n_mma * 2048 + (size_t)kF1Stages * (p.n_full + p.has_pair) * kF1PlaneBytes +
                      (size_t)p.n_luts * 256 * rep * 2 + 64 * 4 + 2 * units * f1_row_pitch(img_c) + 2 * units * 4 + 8 +
                      (2 * kF1Stages + 8) * 8 + 16;
  if (smem > 232448) {
    snprintf(g_err1, sizeof(g_err1), "conv1_fused: %d planes need %zu B of shared memory (strips + operand ring), 232448 available", planes, smem);
    return g_err1;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = p.total_segments < sms ? p.total_segments : sms;
  // MAXCH = channel registers per gather thread (multiple of 16)
  if (img_c == 3) return launch_fused<3, 16, 1, 3>(tO, p, grid, smem, st);
  if (planes == 20 && rep32) return launch_fused<1, 32, 32, 20>(tO, p, grid, smem, st);       // the path's flow stack
  if (planes <= 16) return rep32 ? launch_fused<1, 16, 32, 0>(tO, p, grid, smem, st) : launch_fused<1, 16, 1, 0>(tO, p, grid, smem, st);
  return rep32 ? launch_fused<1, 32, 32, 0>(tO, p, grid, smem, st) : launch_fused<1, 32, 1, 0>(tO, p, grid, smem, st);
}

}  // namespace va
