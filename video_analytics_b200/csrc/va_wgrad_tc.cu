// K5: weight-gradient GEMM of the conv / fully-connected layers on tcgen05, straight from the NHWC tensors.
//
//   dW[tap = (r,s)][co][ci] = sum_{n,h,w} dZ[n, h, w, co] * X[n, h + r - pad, w + s - pad, ci]
//
// GEMM view: M = co (128-row tiles), N = ci (BN-wide tiles), K = pixels.  In NHWC both operands have the CHANNEL
// contiguous, i.e. they are "MN-major" UMMA operands (the K index -- the pixel -- is the strided one), which tcgen05
// takes directly: instruction-descriptor bits a_major = b_major = 1 and the MN-major canonical shared-memory layout
//     Swizzle<3,4,3> o ((8,n),(8,k)) : ((1,LBO),(8,SBO))      [units of 16 bytes, 128B swizzle]
// = rows of 64 channels (128 B), one row per pixel, 8-pixel groups SBO = 1024 B apart, 64-channel blocks LBO apart.
// That is exactly what a 4-D TMA box (64 channels, w_t, h_t, n_t) of the NHWC tensor writes, so a K block is a
// 64-pixel patch fetched by one box per 64-channel block, and filter tap (r,s) is the SAME box of X shifted by
// (s - pad, r - pad) pixels with TMA zero fill outside the image -- the forward kernel's halo trick, with no layout
// pass and no padded rows.  (A first version transposed both tensors to NCHW to get K-major operands; the pixel shift
// then fell on the contiguous dimension, where TMA cannot start a box at a 2-byte offset, and three pre-shifted copies
// of X were needed.  The transposes alone cost more than the GEMM.)
// Work unit = (tap group, ci tile, co tile, K split), tap group fastest so that the CTAs running together stream the
// same pixels.  A unit accumulates T taps at once (T accumulators of BN columns in TMEM): the dZ tile of a K block is
// loaded ONCE and multiplied with T shifted X boxes -- T = 9 for the first layer (BN = 16/32: the kernel is otherwise
// bound by re-reading dZ nine times), T = 3 (one filter column) for BN = 64/128, T = 1 for BN = 256 (TMEM is 512
// columns).  The fp32 accumulators of a unit are added to dW with red.global.add.f32 (split-K).  Warp-specialised like the
// forward kernel: TMA producer warp, MMA warp, 4 epilogue warps, mbarrier full/empty rings, 2 TMEM accumulator stages.
// Reference being replaced: loss.backward() for Conv2d/Linear weights (Sheet03/spatialModel.py:180).
#include "va_internal.h"
#include "va_conv_tc.cuh"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

namespace va {

constexpr int kWgradKP = 64;          // pixels per pipeline stage (K block)

struct WgradParams {
  int n, H, W;
  int w_t, h_t, n_t;          // pixel patch of a K block: w_t * h_t * n_t == kWgradKP
  int tiles_w, tiles_h, tiles_n, pixel_tiles;
  int ks, pad, Cout, Cin;     // Cin = real input channels (epilogue bound); the B tensor may carry padded channels
  int co_tiles, ci_tiles, taps, tap_groups;
  int total_units, ksplit;
  int num_stages;
  int row_stride;             // elements between output rows (Cin)
  int plain_store;            // one K split: every output element is written exactly once -> st.global, no memset
  int vr_rows, vr_shift;      // VR variant: rows of the (h_t+2)-row B box, rows between vertically adjacent taps (= w_t)
  float* dwt;                 // fp32 [taps][Cout][Cin], pre-zeroed
  FastDiv div_taps, div_ci, div_co, div_tw, div_th;
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

// 16-byte vector reduction: one L2 request adds four consecutive floats (sm_90+).  The split-K epilogue issues 128 x BN
// reductions per unit from 128 different rows; with scalar REDs they, not the MMAs, bounded the 256-wide layers at
// batch 64 (ncu source view: 60 % of the stall samples on the RED instructions).
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// MN-major operand descriptor: rows of ROWB bytes (one pixel each) written by TMA with the matching swizzle; 8-row
// groups are ROWB*8 apart (SBO), blocks of ROWB/2 channels are `lbo_bytes` apart (LBO).
template <int ROWB>
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  static_assert(ROWB == 32 || ROWB == 64 || ROWB == 128, "row bytes");
  constexpr uint64_t layout = ROWB == 128 ? 2ull : (ROWB == 64 ? 4ull : 6ull);
  constexpr uint64_t sbo = (ROWB * 8) >> 4;
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= sbo << 32;
  d |= 1ull << 46;
  d |= layout << 61;
  return d;
}
// kind::f16, A = B = bf16, D = fp32, both operands MN-major (bits 15, 16)
__host__ __device__ constexpr uint32_t make_idesc_bf16_mn(int M, int N) { return make_idesc_bf16(M, N) | (1u << 15) | (1u << 16); }

constexpr int kWgradThreads = 192;

// K split i covers pixel tiles [i*PT/ksplit, (i+1)*PT/ksplit): sizes differ by at most one tile
__device__ __forceinline__ int split_begin(const WgradParams& p, int split) {
  return (int)(((long long)split * p.pixel_tiles) / p.ksplit);
}

// BN = N tile (input channels), CB = channels per B block = TMA box width (64, or the padded 16/32 of the first layer),
// T = filter taps accumulated per unit (tap' = group*T + j).
// VR ("vertical reuse", T == 3, one image per K block, w_t a multiple of 8): the three taps r = 0..2 of a filter column
// read ONE B box of h_t + 2 rows at row offsets r*w_t (whole 8-row swizzle atoms) instead of three shifted boxes --
// 36 instead of 64 KB per K block at BN = 128, and the 64/128-wide layers are bound by L2 -> SM bytes.
template <int BN, int CB, int T, bool VR>
__global__ void __launch_bounds__(kWgradThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const WgradParams p) {
  static_assert(BN % CB == 0, "BN must be whole B blocks");
  constexpr int NB = BN / CB;
  constexpr int ROWB_B = CB * 2;
  constexpr uint32_t A_BLOCK = kWgradKP * 128, B_BLOCK = kWgradKP * ROWB_B;      // 8 KB, 8/4/2 KB: all multiples of 1024
  constexpr uint32_t A_BYTES = 2 * A_BLOCK, B_BYTES = NB * B_BLOCK;    // B_BYTES per tap
  static_assert(!VR || (T == 3 && CB == 64), "vertical reuse: one filter column, 64-channel blocks");
  const uint32_t b_block = VR ? (uint32_t)p.vr_rows * ROWB_B : B_BLOCK;            // bytes of one 64-channel B block
  const uint32_t STAGE = VR ? A_BYTES + NB * b_block : A_BYTES + T * B_BYTES;
  constexpr int NACC = (2 * T * BN <= 512) ? 2 : 1;                   // accumulator sets: double-buffered when TMEM allows
  constexpr uint32_t ACC_COLS = T * BN;
  constexpr uint32_t TMEM_NEED = NACC * ACC_COLS;
  constexpr uint32_t TMEM_COLS = TMEM_NEED <= 32 ? 32 : TMEM_NEED <= 64 ? 64 : TMEM_NEED <= 128 ? 128 : TMEM_NEED <= 256 ? 256 : 512;
  static_assert(TMEM_NEED <= 512, "accumulators exceed TMEM");

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

  auto decode = [&](int unit, int& tap, int& ci_t, int& co_t, int& split) {      // tap = first tap' of the unit's group
    uint32_t q, t;
    p.div_taps.divmod((uint32_t)unit, q, t); tap = (int)t * T;
    p.div_ci.divmod(q, q, t); ci_t = (int)t;
    p.div_co.divmod(q, q, t); co_t = (int)t;
    split = (int)q;
  };

  if (warp == kProducerWarp) {
    uint32_t stage = 0, phase = 0;
    for (int unit = blockIdx.x; unit < p.total_units; unit += gridDim.x) {
      int tap, ci_t, co_t, split;
      decode(unit, tap, ci_t, co_t, split);
      const int pt0 = split_begin(p, split), pt1 = split_begin(p, split + 1);
      for (int pt = pt0; pt < pt1; ++pt) {
        uint32_t q, tw, th, tn;
        p.div_tw.divmod((uint32_t)pt, q, tw);
        p.div_th.divmod(q, tn, th);
        const int w0 = (int)tw * p.w_t, h0 = (int)th * p.h_t, n0 = (int)tn * p.n_t;
        mbar_wait(&empty_bar[stage], phase ^ 1, 100 + stage);
        if (elect_one()) {
          uint8_t* a_dst = smem + (size_t)stage * STAGE;
          mbar_arrive_expect_tx(&full_bar[stage], STAGE);
          tma_load_4d(a_dst, &tmA, &full_bar[stage], co_t * 128, w0, h0, n0);
          tma_load_4d(a_dst + A_BLOCK, &tmA, &full_bar[stage], co_t * 128 + 64, w0, h0, n0);
          if (VR) {
            const int s = tap / p.ks;                                        // the group is filter column s
#pragma unroll
            for (int b = 0; b < NB; ++b)
              tma_load_4d(a_dst + A_BYTES + b * b_block, &tmB, &full_bar[stage], ci_t * BN + b * CB, w0 + s - p.pad,
                          h0 - p.pad, n0);
          } else {
#pragma unroll
            for (int j = 0; j < T; ++j) {
              const int s = (tap + j) / p.ks, r = (tap + j) - s * p.ks;      // tap' = s*ks + r (the forward kernel's order)
#pragma unroll
              for (int b = 0; b < NB; ++b)
                tma_load_4d(a_dst + A_BYTES + j * B_BYTES + b * B_BLOCK, &tmB, &full_bar[stage], ci_t * BN + b * CB,
                            w0 + s - p.pad, h0 + r - p.pad, n0);
            }
          }
        }
        __syncwarp();
        if (++stage == (uint32_t)p.num_stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == kMmaWarp) {
    // The T taps' B boxes are consecutive blocks B_BLOCK apart -- the MN-major layout's LBO -- so ONE MMA covers
    // several taps: N = NT*BN columns, tap j of the group landing in accumulator columns [j*BN, (j+1)*BN).
    constexpr int NT = (T * BN <= 256) ? T : (256 / BN);          // taps per MMA (N <= 256)
    constexpr int NG = (T + NT - 1) / NT;                          // MMAs per K step
    constexpr int NT_LAST = T - (NG - 1) * NT;
    constexpr uint32_t idesc = make_idesc_bf16_mn(128, NT * BN);
    constexpr uint32_t idesc_last = make_idesc_bf16_mn(128, NT_LAST * BN);
    const uint32_t smem_base_u32 = smem_u32(smem);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    uint32_t stage = 0, phase = 0, as = 0, as_phase = 0;
    for (int unit = blockIdx.x; unit < p.total_units; unit += gridDim.x) {
      int tap, ci_t, co_t, split;
      decode(unit, tap, ci_t, co_t, split);
      const int pt0 = split_begin(p, split), pt1 = split_begin(p, split + 1);
      mbar_wait(&tempty_bar[as], as_phase ^ 1, 200 + as);
      tc_fence_after();
      const uint32_t d_tmem = tmem_u + as * ACC_COLS;
      uint32_t acc = 0;
      for (int pt = pt0; pt < pt1; ++pt) {
        mbar_wait(&full_bar[stage], phase, 300 + stage);
        tc_fence_after();
        const uint32_t a_addr = smem_base_u32 + stage * STAGE;
        const uint64_t da0 = make_smem_desc_mn<128>(a_addr, A_BLOCK);
        if (elect_one()) {
          if (VR && NB == 1) {
            // blocks of the three taps overlap: tap r starts vr_shift rows after tap r-1 -> one N = 3*BN MMA, LBO = shift
            const uint64_t db0 = make_smem_desc_mn<ROWB_B>(a_addr + A_BYTES, (uint32_t)p.vr_shift * ROWB_B);
#pragma unroll
            for (int k = 0; k < kWgradKP / 16; ++k)
              umma_bf16(d_tmem, da0 + ((k * 16 * 128) >> 4), db0 + ((k * 16 * ROWB_B) >> 4), make_idesc_bf16_mn(128, 3 * BN),
                        k ? 1u : acc);
          } else if (VR) {
            const uint64_t db0 = make_smem_desc_mn<ROWB_B>(a_addr + A_BYTES, b_block);
            const uint32_t shift16 = ((uint32_t)p.vr_shift * ROWB_B) >> 4;
#pragma unroll
            for (int k = 0; k < kWgradKP / 16; ++k)
#pragma unroll
              for (int r = 0; r < 3; ++r)
                umma_bf16(d_tmem + r * BN, da0 + ((k * 16 * 128) >> 4), db0 + r * shift16 + ((k * 16 * ROWB_B) >> 4),
                          make_idesc_bf16_mn(128, BN), k ? 1u : acc);
          } else {
            const uint64_t db0 = make_smem_desc_mn<ROWB_B>(a_addr + A_BYTES, B_BLOCK);
#pragma unroll
            for (int k = 0; k < kWgradKP / 16; ++k)     // 16 pixels = 16 rows: 2 KB of A, 16*ROWB_B of each B block per step
#pragma unroll
              for (int gI = 0; gI < NG; ++gI)
                umma_bf16(d_tmem + gI * NT * BN, da0 + ((k * 16 * 128) >> 4),
                          db0 + ((gI * NT * B_BYTES + k * 16 * ROWB_B) >> 4), gI == NG - 1 ? idesc_last : idesc, k ? 1u : acc);
          }
          umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        acc = 1;
        if (++stage == (uint32_t)p.num_stages) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(&tfull_bar[as]);
      __syncwarp();
      if (++as == NACC) { as = 0; as_phase ^= 1; }
    }
  } else {
    // ===================================================== epilogue: TMEM -> fp32 atomics into dW
    const int q = warp & 3;
    const int m = q * 32 + lane;            // accumulator row == output channel inside the co tile
    uint32_t as = 0, as_phase = 0;
    for (int unit = blockIdx.x; unit < p.total_units; unit += gridDim.x) {
      int tap, ci_t, co_t, split;
      decode(unit, tap, ci_t, co_t, split);
      mbar_wait(&tfull_bar[as], as_phase, 400 + as);
      tc_fence_after();
      const int co = co_t * 128 + m;
#pragma unroll 1
      for (int j = 0; j < T; ++j) {
        float* dst = p.dwt + ((size_t)(tap + j) * p.Cout + co) * p.row_stride + ci_t * BN;
#pragma unroll 1
        for (int c16 = 0; c16 < BN / 16; ++c16) {
          uint32_t v[16];
          tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * ACC_COLS + j * BN + c16 * 16, v);
          tmem_ld_wait();
          if (co < p.Cout) {
            if (p.plain_store) {
              if (ci_t * BN + c16 * 16 + 16 <= p.Cin && (p.row_stride & 3) == 0) {
                float4* d4 = reinterpret_cast<float4*>(dst + c16 * 16);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  d4[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                                      __uint_as_float(v[4 * i + 3]));
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                  if (ci_t * BN + c16 * 16 + i < p.Cin) dst[c16 * 16 + i] = __uint_as_float(v[i]);
              }
            } else if (ci_t * BN + c16 * 16 + 16 <= p.Cin && (p.row_stride & 3) == 0) {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                red_add_v4(dst + c16 * 16 + 4 * i, __uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                           __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int ci = ci_t * BN + c16 * 16 + i;
                if (ci < p.Cin) atomicAdd(dst + c16 * 16 + i, __uint_as_float(v[i]));
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      if (++as == NACC) { as = 0; as_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------ CTA-pair variant
// 256-wide layers: a 2-CTA cluster computes a 256 (co) x 256 (ci) tile with tcgen05 cta_group::2.  Each CTA loads its
// own 128 output channels of dZ and only HALF of the 256 input channels of X per K block (32 instead of 48 KB -> 6
// pipeline stages, a third less L2 -> SM traffic per FLOP; the one-CTA kernel sits at ~51 % tensor-pipe activity, bound
// by operand delivery).  Same protocol as conv_tc2_kernel: both producers credit the LEADER's full barrier, the leader
// issues the MMAs, tcgen05.commit multicasts to both CTAs, both epilogues release the accumulator on the leader's
// barrier.  Work unit = (tap, ci tile, pair of co tiles, K split).
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kWgradThreads, 1)
wgrad_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const WgradParams p) {
  constexpr int BN = 256;
  constexpr uint32_t BLOCK = kWgradKP * 128;                 // one [64 px][64 ch] block: 8 KB
  constexpr uint32_t A_BYTES = 2 * BLOCK, B_HALF = 2 * BLOCK, STAGE = A_BYTES + B_HALF;   // 32 KB per CTA
  constexpr uint32_t TMEM_COLS = 512;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.num_stages * STAGE);   // used in the leader CTA
  uint64_t* empty_bar = full_bar + kMaxStages;                                               // own copy in each CTA
  uint64_t* tfull_bar = empty_bar + kMaxStages;                                              // own copy in each CTA
  uint64_t* tempty_bar = tfull_bar + 2;                                                      // used in the leader CTA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  constexpr int kProducerWarp = 4, kMmaWarp = 5;
  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < p.num_stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 8); }   // 4 warps x 2 CTAs
    fence_mbar_init();
  }
  if (warp == kMmaWarp) { tmem_alloc2(tmem_slot, TMEM_COLS); tmem_relinquish2(); }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](int unit, int& tap, int& ci_t, int& co_p, int& split) {
    uint32_t q, t;
    p.div_taps.divmod((uint32_t)unit, q, t); tap = (int)t;
    p.div_ci.divmod(q, q, t); ci_t = (int)t;
    p.div_co.divmod(q, q, t); co_p = (int)t;         // div_co divides by the number of co PAIRS here
    split = (int)q;
  };

  if (warp == kProducerWarp) {
    uint32_t stage = 0, phase = 0;
    for (int unit = cluster_id; unit < p.total_units; unit += n_clusters) {
      int tap, ci_t, co_p, split;
      decode(unit, tap, ci_t, co_p, split);
      const int s = tap / p.ks, r = tap - s * p.ks;
      const int pt0 = split_begin(p, split), pt1 = split_begin(p, split + 1);
      const int co0 = co_p * 256 + (int)rank * 128, ci0 = ci_t * BN + (int)rank * 128;
      for (int pt = pt0; pt < pt1; ++pt) {
        uint32_t q, tw, th, tn;
        p.div_tw.divmod((uint32_t)pt, q, tw);
        p.div_th.divmod(q, tn, th);
        const int w0 = (int)tw * p.w_t, h0 = (int)th * p.h_t, n0 = (int)tn * p.n_t;
        mbar_wait(&empty_bar[stage], phase ^ 1, 100 + stage);
        if (elect_one()) {
          uint8_t* a_dst = smem + (size_t)stage * STAGE;
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE);
          tma_load_4d_2sm(a_dst, &tmA, &full_bar[stage], co0, w0, h0, n0);
          tma_load_4d_2sm(a_dst + BLOCK, &tmA, &full_bar[stage], co0 + 64, w0, h0, n0);
          tma_load_4d_2sm(a_dst + A_BYTES, &tmB, &full_bar[stage], ci0, w0 + s - p.pad, h0 + r - p.pad, n0);
          tma_load_4d_2sm(a_dst + A_BYTES + BLOCK, &tmB, &full_bar[stage], ci0 + 64, w0 + s - p.pad, h0 + r - p.pad, n0);
        }
        __syncwarp();
        if (++stage == (uint32_t)p.num_stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == kMmaWarp) {
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16_mn(256, BN);
      const uint32_t smem_base_u32 = smem_u32(smem);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      uint32_t stage = 0, phase = 0, as = 0, as_phase = 0;
      for (int unit = cluster_id; unit < p.total_units; unit += n_clusters) {
        int tap, ci_t, co_p, split;
        decode(unit, tap, ci_t, co_p, split);
        const int pt0 = split_begin(p, split), pt1 = split_begin(p, split + 1);
        mbar_wait(&tempty_bar[as], as_phase ^ 1, 200 + as);
        tc_fence_after();
        const uint32_t d_tmem = tmem_u + as * BN;
        uint32_t acc = 0;
        for (int pt = pt0; pt < pt1; ++pt) {
          mbar_wait(&full_bar[stage], phase, 300 + stage);
          tc_fence_after();
          const uint32_t a_addr = smem_base_u32 + stage * STAGE;
          const uint64_t da0 = make_smem_desc_mn<128>(a_addr, BLOCK);
          const uint64_t db0 = make_smem_desc_mn<128>(a_addr + A_BYTES, BLOCK);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kWgradKP / 16; ++k)
              umma_bf16_2sm(d_tmem, da0 + ((k * 16 * 128) >> 4), db0 + ((k * 16 * 128) >> 4), idesc, k ? 1u : acc);
            umma_commit_2sm(&empty_bar[stage], 3);
            if (pt == pt1 - 1) umma_commit_2sm(&tfull_bar[as], 3);
          }
          __syncwarp();
          acc = 1;
          if (++stage == (uint32_t)p.num_stages) { stage = 0; phase ^= 1; }
        }
        as ^= 1;
        if (as == 0) as_phase ^= 1;
      }
    }
  } else {
    const int q = warp & 3;
    const int m = q * 32 + lane;
    uint32_t as = 0, as_phase = 0;
    for (int unit = cluster_id; unit < p.total_units; unit += n_clusters) {
      int tap, ci_t, co_p, split;
      decode(unit, tap, ci_t, co_p, split);
      mbar_wait(&tfull_bar[as], as_phase, 400 + as);
      tc_fence_after();
      const int co = co_p * 256 + (int)rank * 128 + m;
      float* dst = p.dwt + ((size_t)tap * p.Cout + co) * p.row_stride + ci_t * BN;
#pragma unroll 1
      for (int c16 = 0; c16 < BN / 16; ++c16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + c16 * 16, v);
        tmem_ld_wait();
        if (co < p.Cout) {
          if (ci_t * BN + c16 * 16 + 16 <= p.Cin && (p.row_stride & 3) == 0) {
            if (p.plain_store) {
              float4* d4 = reinterpret_cast<float4*>(dst + c16 * 16);
#pragma unroll
              for (int i = 0; i < 4; ++i)
                d4[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                                    __uint_as_float(v[4 * i + 3]));
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                red_add_v4(dst + c16 * 16 + 4 * i, __uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                           __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int ci = ci_t * BN + c16 * 16 + i;
              if (ci < p.Cin) {
                if (p.plain_store) dst[c16 * 16 + i] = __uint_as_float(v[i]);
                else atomicAdd(dst + c16 * 16 + i, __uint_as_float(v[i]));
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tempty_bar[as]);
      as ^= 1;
      if (as == 0) as_phase ^= 1;
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == kMmaWarp) tmem_dealloc2(tmem_base, TMEM_COLS);
}

// [taps][Cout][Cin] (tap' = s*ks + r) -> OIHW fp32 [Cout][Cin][ks][ks]; one thread per (co, ci): coalesced reads of each
// tap plane, ks*ks consecutive floats written per thread.
__global__ void wgrad_to_oihw_kernel(const float* __restrict__ dwt, float* __restrict__ dw, int Cout, int Cin, int ks) {
  const size_t plane = (size_t)Cout * Cin;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < plane; g += (size_t)gridDim.x * blockDim.x) {
    float* o = dw + g * ks * ks;
    for (int r = 0; r < ks; ++r)
      for (int s = 0; s < ks; ++s) o[r * ks + s] = __ldg(dwt + (size_t)(s * ks + r) * plane + g);
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
int sm_count_cached() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}
// NHWC bf16 [n][H][W][C] map with dims (C, W, H, n), box (cb, w_t, h_t, n_t), swizzle = cb*2 bytes
const char* encode_nhwc(CUtensorMap* m, const void* addr, int n, int H, int W, int C, int cb, int w_t, int h_t, int n_t) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return "cuTensorMapEncodeTiled not available";
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
  cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * W * 2, (cuuint64_t)C * W * H * 2};
  cuuint32_t box[4] = {(cuuint32_t)cb, (cuuint32_t)w_t, (cuuint32_t)h_t, (cuuint32_t)n_t};
  cuuint32_t es[4] = {1, 1, 1, 1};
  const CUtensorMapSwizzle sw = cb == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (cb == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(addr), gdim, gstr, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return werrf("wgrad tensor map encode failed (%d): n=%d H=%d W=%d C=%d box=(%d,%d,%d,%d)", (int)r, n, H, W, C, cb, w_t, h_t, n_t);
  return nullptr;
}

template <int BN, int CB, int T, bool VR>
const char* launch_wgrad(const CUtensorMap& tA, const CUtensorMap& tB, WgradParams p, int grid, cudaStream_t st) {
  const uint32_t STAGE = VR ? 2 * kWgradKP * 128 + (BN / CB) * p.vr_rows * CB * 2
                            : 2 * kWgradKP * 128 + T * (BN / CB) * kWgradKP * CB * 2;
  int stages = (int)((227 * 1024 - 1024 - 256) / STAGE);
  if (stages > kMaxStages) stages = kMaxStages;
  p.num_stages = stages;
  const size_t smem = 1024 + (size_t)stages * STAGE + 256;
  auto kfn = wgrad_tc_kernel<BN, CB, T, VR>;
  static size_t configured = 0;
  if (configured < smem) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return werrf("cudaFuncSetAttribute(wgrad smem=%zu): %s", smem, cudaGetErrorString(e));
    configured = smem;
  }
  count_launch();
  kfn<<<grid, kWgradThreads, smem, st>>>(tA, tB, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return werrf("wgrad_tc_kernel<%d,%d,%d,%d> launch: %s", BN, CB, T, (int)VR, cudaGetErrorString(e));
  return nullptr;
}
}  // namespace

// dz: bf16 NHWC [n][H][W][Cout]; x: bf16 NHWC [n][H][W][cin_pad] (channels >= Cin are zero); dwt_ws: fp32
// [ks*ks][Cout][Cin] workspace (unused for ks == 1: the accumulation target is dw itself); dw: fp32 [Cout][Cin][ks][ks].
const char* wgrad_run(const void* dz, const void* x, int n, int H, int W, int Cout, int Cin, int cin_pad, int ks,
                      float* dwt_ws, float* dw, cudaStream_t st) {
  if (n <= 0) return nullptr;
  if (Cout % 8 || cin_pad % 8) return "wgrad: channel counts must be multiples of 8 (16-byte NHWC rows)";
  const int CB = cin_pad >= 64 ? 64 : (cin_pad >= 32 ? 32 : 16);
  if (cin_pad < 16 || (cin_pad < 64 && cin_pad != CB)) return werrf("wgrad: unsupported padded input channel count %d", cin_pad);
  const int BN = cin_pad >= 256 ? 256 : (cin_pad >= 128 ? 128 : CB);
  WgradParams p;
  p.n = n; p.H = H; p.W = W;
  // K block = 64-pixel patch w_t x h_t x n_t; w_t and h_t divide W and H so that no patch straddles an image edge
  int w_t = 16;
  while (w_t > 1 && W % w_t) w_t >>= 1;
  int h_t = kWgradKP / w_t;
  while (h_t > 1 && H % h_t) h_t >>= 1;
  p.w_t = w_t; p.h_t = h_t; p.n_t = kWgradKP / (w_t * h_t);
  p.tiles_w = W / w_t; p.tiles_h = H / h_t; p.tiles_n = (n + p.n_t - 1) / p.n_t;
  p.pixel_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  p.ks = ks; p.pad = (ks - 1) / 2; p.Cout = Cout; p.Cin = Cin; p.row_stride = Cin;
  p.co_tiles = (Cout + 127) / 128; p.ci_tiles = (cin_pad + BN - 1) / BN; p.taps = ks * ks;
  // taps per unit: all 9 for the first layer, one filter column for the 64/128-wide N tiles, 1 when TMEM holds only one
  const int T = ks == 1 ? 1 : (BN <= 32 ? 9 : (BN <= 128 ? 3 : 1));
  p.tap_groups = p.taps / T;
  const int sms = sm_count_cached();
  // CTA pairs for the 256-wide tiles of the conv layers (the FC weight gradients are one K block per tile: no pipeline)
  const bool pair = BN == 256 && T == 1 && ks == 3 && Cout >= 256 && getenv("VA_WGRAD_NO_PAIR") == nullptr;
  const int co_units = pair ? (Cout + 255) / 256 : p.co_tiles;
  const int exec_units = pair ? sms / 2 : sms;              // clusters (pair) or CTAs that run concurrently
  const int base_units = p.tap_groups * co_units * p.ci_tiles;
  // K splits: 2..8 units per SM, picked for the fullest last wave (units are dealt round-robin to one CTA per SM, so
  // 2.07 waves cost 3); fewer splits win ties (less atomic traffic)
  int ksplit = 1;
  {
    double best = -1.0;
    const int k_lo = std::max(1, (2 * exec_units + base_units - 1) / base_units);
    const int k_hi = std::max(k_lo, (8 * exec_units) / base_units);
    for (int k = k_lo; k <= k_hi && k <= p.pixel_tiles; ++k) {
      const int units = base_units * k;
      const int waves = (units + exec_units - 1) / exec_units;
      const double eff = (double)units / ((double)waves * exec_units);
      if (eff > best + 0.01) { best = eff; ksplit = k; }
    }
    if (ksplit > p.pixel_tiles) ksplit = p.pixel_tiles;
    if (base_units >= 2 * exec_units) ksplit = 1;       // enough independent output tiles (the FC layers): no split-K
  }
  p.ksplit = ksplit;
  p.total_units = base_units * ksplit;
  float* target = ks == 1 ? dw : dwt_ws;
  p.dwt = target;
  p.div_taps = FastDiv::make((uint32_t)p.tap_groups);
  p.div_ci = FastDiv::make((uint32_t)p.ci_tiles);
  p.div_co = FastDiv::make((uint32_t)co_units);
  p.div_tw = FastDiv::make((uint32_t)p.tiles_w);
  p.div_th = FastDiv::make((uint32_t)p.tiles_h);
  p.plain_store = ksplit == 1 ? 1 : 0;
  cudaError_t ce = cudaSuccess;
  if (!p.plain_store) ce = cudaMemsetAsync(target, 0, (size_t)p.taps * Cout * Cin * sizeof(float), st);
  if (ce != cudaSuccess) return werrf("wgrad memset: %s", cudaGetErrorString(ce));
  CUtensorMap tA, tB;
  if (const char* e = encode_nhwc(&tA, dz, n, H, W, Cout, 64, p.w_t, p.h_t, p.n_t)) return e;
  // vertical reuse needs one image per K block and tap offsets of whole 8-row swizzle atoms
  const bool vr = T == 3 && CB == 64 && p.n_t == 1 && (p.w_t % 8) == 0 && getenv("VA_WGRAD_NO_VR") == nullptr;
  p.vr_rows = (p.h_t + 2) * p.w_t; p.vr_shift = p.w_t;
  if (const char* e = encode_nhwc(&tB, x, n, H, W, cin_pad, CB, p.w_t, vr ? p.h_t + 2 : p.h_t, p.n_t)) return e;
  const int grid = p.total_units < sms ? p.total_units : sms;
  const char* err = nullptr;
  if (pair) {
    constexpr uint32_t STAGE2 = 4 * kWgradKP * 128;
    int stages = (int)((227 * 1024 - 1024 - 256) / STAGE2);
    if (stages > kMaxStages) stages = kMaxStages;
    p.num_stages = stages;
    const size_t smem = 1024 + (size_t)stages * STAGE2 + 256;
    static size_t configured2 = 0;
    if (configured2 < smem) {
      cudaError_t e2 = cudaFuncSetAttribute(wgrad_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e2 != cudaSuccess) return werrf("cudaFuncSetAttribute(wgrad pair, smem=%zu): %s", smem, cudaGetErrorString(e2));
      configured2 = smem;
    }
    const int clusters = p.total_units < sms / 2 ? p.total_units : sms / 2;
    count_launch();
    wgrad_tc2_kernel<<<2 * clusters, kWgradThreads, smem, st>>>(tA, tB, p);
    cudaError_t e2 = cudaGetLastError();
    if (e2 != cudaSuccess) return werrf("wgrad_tc2_kernel launch: %s", cudaGetErrorString(e2));
  } else
#define VA_W(bn, cb, t, v) if (BN == bn && CB == cb && T == t && vr == v) err = launch_wgrad<bn, cb, t, v>(tA, tB, p, grid, st); else
  VA_W(256, 64, 1, false) VA_W(128, 64, 3, false) VA_W(64, 64, 3, false) VA_W(32, 32, 9, false) VA_W(16, 16, 9, false)
  VA_W(128, 64, 3, true) VA_W(64, 64, 3, true)
  VA_W(128, 64, 1, false) VA_W(64, 64, 1, false) VA_W(32, 32, 1, false) VA_W(16, 16, 1, false)
  err = werrf("wgrad: no kernel for BN=%d CB=%d T=%d", BN, CB, T);
#undef VA_W
  if (err) return err;
  if (ks != 1) {
    const size_t plane = (size_t)Cout * Cin;
    unsigned blocks = (unsigned)std::min<size_t>((plane + 255) / 256, 148 * 8);
    count_launch();
    wgrad_to_oihw_kernel<<<blocks, 256, 0, st>>>(dwt_ws, dw, Cout, Cin, ks);
    ce = cudaGetLastError();
    if (ce != cudaSuccess) return werrf("wgrad_to_oihw: %s", cudaGetErrorString(ce));
  }
  return nullptr;
}

}  // namespace va
