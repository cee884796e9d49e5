// K2 for the wide layers (Cout >= 256): the same implicit-GEMM convolution as va_conv_tc.cuh, on CTA PAIRS
// (tcgen05 cta_group::2, UMMA M = 256).
//
// Why: with one CTA per tile a 128x256 tile needs 16 KB (A) + 32 KB (B) per 64-channel k block, so only 4 pipeline
// stages fit in 227 KB and the MMA warp measurably waits on operands (12-15 %, profiles/r01_role_stalls_v7.log):
// 4 x 48 KB in flight cannot cover the TMA latency at the 94 B/clk/SM the tensor pipe consumes.  A CTA pair computes
// a 256-pixel x 256-channel tile; each CTA loads its own 128-pixel A patch and only HALF of the weight tile
// (128 of the 256 output channels), the tensor cores of both SMs read both halves.  Per CTA a stage is 32 KB -> 6
// stages, and the L2->SM weight traffic halves.
//
// Protocol (rank 0 = leader):
//   * TMA: both CTAs issue their loads with .cta_group::2 and the LEADER's "full" barrier (count 1: the leader's
//     producer arms it with the bytes of both CTAs).
//   * MMA: leader only, tcgen05.mma.cta_group::2, D in the TMEM of both CTAs (128 lanes x 256 columns each).
//   * tcgen05.commit ... multicast 0b11 releases the smem stage (empty barrier) and publishes the accumulator
//     (tmem-full barrier) in BOTH CTAs.
//   * epilogues of both CTAs hand the accumulator back by arriving on the LEADER's tmem-empty barrier (count 8).
// Work unit = (pair of adjacent M tiles, N tile); CTA rank r takes M tile 2*pair + r (an odd tail tile is fully
// out of bounds: TMA zero-fills its loads and clips its stores).
#pragma once
#include "va_conv_tc.cuh"

namespace va {

constexpr int kConv2Threads = 320;

inline size_t conv2_smem_bytes(uint32_t staging_bytes, int stages) {
  return 1024 + (size_t)stages * 32768 + 2 * (size_t)staging_bytes + 256;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kConv2Threads, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                const __grid_constant__ CUtensorMap tmO, const ConvKernelParams p) {
  constexpr int BN = 256, CK = 64, ROWB = 128;
  constexpr uint32_t A_BYTES = 128 * ROWB, B_HALF_BYTES = 128 * ROWB, STAGE_BYTES = A_BYTES + B_HALF_BYTES;
  constexpr uint32_t TMEM_COLS = 2 * BN;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + (size_t)p.num_stages * STAGE_BYTES;
  // (the bias slice is read straight from global/L1 in the epilogue: its 2 KB smem copy is what stands between 5 and 6
  // pipeline stages for the un-pooled layers)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + 2 * p.staging_bytes);   // used in the leader CTA
  uint64_t* empty_bar = full_bar + kMaxStages;                           // own copy in each CTA
  uint64_t* tfull_bar = empty_bar + kMaxStages;                          // own copy in each CTA
  uint64_t* tempty_bar = tfull_bar + 2;                                  // used in the leader CTA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;
  constexpr int kProducerWarp = 8, kMmaWarp = 9;

  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmO);
    for (int i = 0; i < p.num_stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 8);   // 4 epilogue warps of the owning group x 2 CTAs
    }
    fence_mbar_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc2(tmem_slot, TMEM_COLS);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync_all();          // both CTAs' barriers initialised and TMEM allocated before any cross-CTA traffic
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int groups = p.ks * p.ks;
  const int num_kb = groups * p.cin_chunks;
  const int total_units = p.total_tiles;     // host passes pairs_m * n_tiles_cout

  if (warp == kProducerWarp) {
    // ===================================================== TMA producer (both CTAs)
    uint32_t stage = 0, phase = 0;
    for (int unit = cluster_id; unit < total_units; unit += n_clusters) {
      uint32_t nt, pm, mt, tw, th, tn;
      p.div_cout.divmod((uint32_t)unit, pm, nt);
      mt = 2 * pm + rank;
      p.div_w.divmod(mt, mt, tw);
      p.div_h.divmod(mt, tn, th);
      const int n0 = tn * p.n_t, h0 = th * p.h_t, w0 = tw * p.w_t, c0 = nt * BN + (int)rank * 128;
      int s = 0, r = 0;
      for (int g = 0; g < groups; ++g) {
        const int wx = w0 + s - p.pad, hy = h0 + r - p.pad;
        for (int cc = 0; cc < p.cin_chunks; ++cc) {
          mbar_wait(&empty_bar[stage], phase ^ 1, 100 + stage);
          if (elect_one()) {
            uint8_t* a_dst = smem + (size_t)stage * STAGE_BYTES;
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
            tma_load_4d_2sm(a_dst, &tmA, &full_bar[stage], cc * CK, wx, hy, n0);
            tma_load_3d_2sm(a_dst + A_BYTES, &tmW, &full_bar[stage], cc * CK, c0, g);
          }
          __syncwarp();
          if (++stage == (uint32_t)p.num_stages) { stage = 0; phase ^= 1; }
        }
        if (++r == p.ks) { r = 0; ++s; }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================================================== MMA issuer (leader CTA only)
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(256, BN);
      const uint32_t smem_base_u32 = smem_u32(smem);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      uint32_t stage = 0, phase = 0, as = 0, as_phase = 0;
      for (int unit = cluster_id; unit < total_units; unit += n_clusters) {
        mbar_wait(&tempty_bar[as], as_phase ^ 1, 200 + as);
        tc_fence_after();
        const uint32_t d_tmem = tmem_u + as * BN;
        uint32_t acc = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase, 300 + stage);
          tc_fence_after();
          const uint32_t a_addr = smem_base_u32 + stage * STAGE_BYTES;
          const uint64_t da0 = make_smem_desc<ROWB>(a_addr);
          const uint64_t db0 = make_smem_desc<ROWB>(a_addr + A_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < CK / 16; ++k) umma_bf16_2sm(d_tmem, da0 + 2 * k, db0 + 2 * k, idesc, k ? 1u : acc);
            umma_commit_2sm(&empty_bar[stage], 3);                       // both CTAs may refill this stage
            if (kb == num_kb - 1) umma_commit_2sm(&tfull_bar[as], 3);    // accumulators complete in both CTAs
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
    // ===================================================== epilogue (2 groups x 4 warps, both CTAs)
    const int eg = warp >> 2;
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int et = threadIdx.x - eg * 128;
    uint8_t* stage_out = staging + eg * p.staging_bytes;
    const int w_i = m & (p.w_t - 1);
    const int h_i = (m >> p.log2_w_t) & (p.h_t - 1);
    const int n_i = m >> (p.log2_w_t + p.log2_h_t);
    const uint32_t as = (uint32_t)eg;
    int it = 0;
    for (int unit = cluster_id; unit < total_units; unit += n_clusters, ++it) {
      if ((it & 1) != eg) continue;
      const uint32_t as_phase = (uint32_t)(it >> 1) & 1u;
      uint32_t nt, pm, mt, tw, th, tn;
      p.div_cout.divmod((uint32_t)unit, pm, nt);
      mt = 2 * pm + rank;
      p.div_w.divmod(mt, mt, tw);
      p.div_h.divmod(mt, tn, th);
      const int n0 = tn * p.n_t, h0 = th * p.h_t, w0 = tw * p.w_t, c0 = nt * BN;

      mbar_wait(&tfull_bar[as], as_phase, 400 + as);
      tc_fence_after();

#pragma unroll 1
      for (int chunk = 0; chunk < BN / 64; ++chunk) {
        uint32_t v0[32], v1[32];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + chunk * 64;
        tmem_ld32(taddr, v0);
        tmem_ld32(taddr + 32, v1);
        tmem_ld_wait();
        if (chunk == BN / 64 - 1) {        // this CTA's half of the accumulator stage is in registers
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(&tempty_bar[as]);
        }
        const float2* bs = reinterpret_cast<const float2*>(p.bias + c0 + chunk * 64);   // warp-uniform, L1-resident
        uint32_t pk[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 b0 = __ldg(bs + i), b1 = __ldg(bs + 16 + i);
          float a = __uint_as_float(v0[2 * i]) + b0.x;
          float b = __uint_as_float(v0[2 * i + 1]) + b0.y;
          float c = __uint_as_float(v1[2 * i]) + b1.x;
          float d = __uint_as_float(v1[2 * i + 1]) + b1.y;
          if (p.relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); c = fmaxf(c, 0.f); d = fmaxf(d, 0.f); }
          __nv_bfloat162 lo = __floats2bfloat162_rn(a, b);
          __nv_bfloat162 hi = __floats2bfloat162_rn(c, d);
          pk[i] = *reinterpret_cast<uint32_t*>(&lo);
          pk[16 + i] = *reinterpret_cast<uint32_t*>(&hi);
        }
        bool writer = true;
        int row = m;
        if (p.pool) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            uint32_t u = pk[i];
            u = bf162_max(u, __shfl_xor_sync(0xffffffffu, u, 1));
            u = bf162_max(u, __shfl_xor_sync(0xffffffffu, u, p.w_t));
            pk[i] = u;
          }
          writer = ((w_i | h_i) & 1) == 0;
          row = ((n_i * (p.h_t >> 1)) + (h_i >> 1)) * (p.w_t >> 1) + (w_i >> 1);
        }
        if (et < 32) {
          if (elect_one()) tma_store_wait_read<0>();
          __syncwarp();
        }
        named_bar_sync(3 + eg, 128);
        if (writer) {
          uint8_t* rowp = stage_out + row * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            uint4 val = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
            *reinterpret_cast<uint4*>(rowp + ((c ^ (row & 7)) << 4)) = val;
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(3 + eg, 128);
        if (et < 32) {
          if (elect_one()) {
            const int sh = p.pool ? 1 : 0;
            tma_store_4d(&tmO, stage_out, c0 + chunk * 64, w0 >> sh, h0 >> sh, n0);
            tma_store_commit();
          }
          __syncwarp();
        }
      }
    }
    if (et < 32) {
      if (elect_one()) tma_store_wait_all();
      __syncwarp();
    }
  }

  tc_fence_before();
  cluster_sync_all();          // nobody leaves (or frees TMEM) while the peer may still touch this CTA's barriers / smem
  if (warp == kMmaWarp) tmem_dealloc2(tmem_base, TMEM_COLS);
}

}  // namespace va

namespace va {

// ---------------------------------------------------------------------------------------------------------------
// CTA-pair kernel for the NARROW 3x3 layers at 224^2 / 112^2 (Cout 64 / 128, Cin 64 / 128: conv1_2, conv2_1, conv2_2
// and their data gradients): HALO operand windows + RESIDENT weight halves.
//
// Why: with one CTA per tile these layers are bound by shared-memory bandwidth, not by the tensor pipe
// (profiles/r01_microbench_mma_rate.log, r02_conv1_2_variants.log): an M128 x N64 x K16 MMA reads 4 KB of A and 2 KB of B
// for 32 clk of math (48 clk at 128 B/clk), an N128 MMA 4 + 4 KB for 64 clk -- and conv2_x re-loaded its 144 / 288 KB
// of weights through TMA for EVERY tile on top.  In a CTA pair (tcgen05 cta_group::2, UMMA M = 256) each SM reads its
// own A window and only HALF of B per MMA (5 KB per 32 clk at N = 64, 6 KB per 64 clk at N = 128), and its half of the
// layer's weights (36 / 72 / 144 KB) stays resident in shared memory for the whole kernel.  Per tile the only smem
// writes left are one haloed 18 x 10-pixel box per 64-channel chunk (23 KB, serves all nine taps through start-address
// shifts, see va_conv_tc.cuh HALO) and the output staging.
//
// Protocol = conv_tc2_kernel's (leader-issued MMAs, 2-SM TMA loads crediting the leader's barriers, multicast commits).
// Tile = 16 x 8 output pixels; work unit = two adjacent tiles (CTA rank r takes tile 2*unit + r).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kConv2hThreads = 320;

inline size_t conv2h_smem_bytes(int BN, int cin_chunks, uint32_t a_box_bytes, uint32_t staging_bytes, int stages) {
  return 1024 + (size_t)cin_chunks * 9 * (BN / 2) * 128 + (size_t)stages * a_box_bytes + 2 * (size_t)staging_bytes + 256;
}

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kConv2hThreads, 1)
conv_tc2h_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                 const __grid_constant__ CUtensorMap tmO, const ConvKernelParams p) {
  static_assert(BN == 64 || BN == 128, "BN");
  constexpr int ROWB = 128;
  constexpr uint32_t HALF_TAP = (BN / 2) * ROWB;          // this CTA's rows of one tap of one channel chunk
  constexpr uint32_t TMEM_COLS = 2 * BN;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t wres_bytes = (uint32_t)p.cin_chunks * 9 * HALF_TAP;
  uint8_t* ring = smem + wres_bytes;
  uint8_t* staging = ring + (size_t)p.num_stages * p.a_box_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + 2 * p.staging_bytes);   // used in the leader CTA
  uint64_t* empty_bar = full_bar + kMaxStages;                           // own copy in each CTA
  uint64_t* tfull_bar = empty_bar + kMaxStages;                          // own copy in each CTA
  uint64_t* tempty_bar = tfull_bar + 2;                                  // used in the leader CTA
  uint64_t* wres_bar = tempty_bar + 2;                                   // used in the leader CTA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wres_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;
  constexpr int kProducerWarp = 8, kMmaWarp = 9;

  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmO);
    for (int i = 0; i < p.num_stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 8);   // 4 epilogue warps of the owning group x 2 CTAs
    }
    mbar_init(wres_bar, 1);
    fence_mbar_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc2(tmem_slot, TMEM_COLS);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_units = p.total_tiles;     // host passes ceil(tiles_m / 2) (one N tile: Cout == BN)

  if (warp == kProducerWarp) {
    // ===================================================== TMA producer (both CTAs)
    if (elect_one()) {
      // this CTA's half of every tap's weight rows, once; both CTAs' bytes are credited to the leader's barrier
      if (rank == 0) mbar_arrive_expect_tx(wres_bar, 2 * wres_bytes);
      for (int cc = 0; cc < p.cin_chunks; ++cc)
        tma_load_3d_2sm(smem + (size_t)cc * 9 * HALF_TAP, &tmW, wres_bar, cc * 64, (int)rank * (BN / 2), 0);
    }
    __syncwarp();
    uint32_t stage = 0, phase = 0;
    for (int unit = cluster_id; unit < total_units; unit += n_clusters) {
      uint32_t mt = 2 * (uint32_t)unit + rank, tw, th, tn;
      p.div_w.divmod(mt, mt, tw);
      p.div_h.divmod(mt, tn, th);
      const int h0 = th * p.h_t, w0 = tw * p.w_t;
      for (int cc = 0; cc < p.cin_chunks; ++cc) {
        mbar_wait(&empty_bar[stage], phase ^ 1, 100 + stage);
        if (elect_one()) {
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * p.a_tx_bytes);
          tma_load_4d_2sm(ring + (size_t)stage * p.a_box_bytes, &tmA, &full_bar[stage], cc * 64, w0 - 1, h0 - 1, (int)tn);
        }
        __syncwarp();
        if (++stage == (uint32_t)p.num_stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================================================== MMA issuer (leader CTA only)
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(256, BN);
      const uint32_t smem_base_u32 = smem_u32(smem);
      const uint32_t ring_u32 = smem_base_u32 + wres_bytes;
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      // tap (r,s) = window of the haloed box starting at pixel row r*pitch + s; 8-pixel groups pitch*128 B apart
      uint32_t a_win[9];
#pragma unroll
      for (int s = 0; s < 3; ++s)
#pragma unroll
        for (int r = 0; r < 3; ++r) a_win[s * 3 + r] = (uint32_t)((r * p.hb_pitch + s) * ROWB) >> 4;
      const uint64_t a_sbo_fix = ((uint64_t)(((uint32_t)p.hb_pitch * ROWB) >> 4) << 32) - ((uint64_t)((ROWB * 8) >> 4) << 32);
      mbar_wait(wres_bar, 0, 500);
      uint32_t stage = 0, phase = 0, as = 0, as_phase = 0;
      for (int unit = cluster_id; unit < total_units; unit += n_clusters) {
        mbar_wait(&tempty_bar[as], as_phase ^ 1, 200 + as);
        tc_fence_after();
        const uint32_t d_tmem = tmem_u + as * BN;
        for (int cc = 0; cc < p.cin_chunks; ++cc) {
          mbar_wait(&full_bar[stage], phase, 300 + stage);
          tc_fence_after();
          const uint64_t da0 = make_smem_desc<ROWB>(ring_u32 + stage * p.a_box_bytes) + a_sbo_fix;
          const uint64_t db0 = make_smem_desc<ROWB>(smem_base_u32 + (uint32_t)cc * 9 * HALF_TAP);
          if (elect_one()) {
#pragma unroll
            for (int t = 0; t < 9; ++t) {          // weights are packed tap' = s*3 + r, as a_win
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_2sm(d_tmem, da0 + a_win[t] + 2 * k, db0 + ((t * HALF_TAP) >> 4) + 2 * k, idesc, (cc | t | k) ? 1u : 0u);
            }
            umma_commit_2sm(&empty_bar[stage], 3);
            if (cc == p.cin_chunks - 1) umma_commit_2sm(&tfull_bar[as], 3);
          }
          __syncwarp();
          if (++stage == (uint32_t)p.num_stages) { stage = 0; phase ^= 1; }
        }
        as ^= 1;
        if (as == 0) as_phase ^= 1;
      }
    }
  } else {
    // ===================================================== epilogue (2 groups x 4 warps, both CTAs)
    const int eg = warp >> 2;
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int et = threadIdx.x - eg * 128;
    uint8_t* stage_out = staging + eg * p.staging_bytes;
    const int w_i = m & (p.w_t - 1);
    const int h_i = (m >> p.log2_w_t) & (p.h_t - 1);
    const uint32_t as = (uint32_t)eg;
    int it = 0;
    for (int unit = cluster_id; unit < total_units; unit += n_clusters, ++it) {
      if ((it & 1) != eg) continue;
      const uint32_t as_phase = (uint32_t)(it >> 1) & 1u;
      uint32_t mt = 2 * (uint32_t)unit + rank, tw, th, tn;
      p.div_w.divmod(mt, mt, tw);
      p.div_h.divmod(mt, tn, th);
      const int n0 = (int)tn, h0 = th * p.h_t, w0 = tw * p.w_t;

      mbar_wait(&tfull_bar[as], as_phase, 400 + as);
      tc_fence_after();

#pragma unroll 1
      for (int chunk = 0; chunk < BN / 64; ++chunk) {
        uint32_t v0[32], v1[32];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + chunk * 64;
        tmem_ld32(taddr, v0);
        tmem_ld32(taddr + 32, v1);
        tmem_ld_wait();
        if (chunk == BN / 64 - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(&tempty_bar[as]);
        }
        const float2* bs = reinterpret_cast<const float2*>(p.bias + chunk * 64);   // warp-uniform, L1-resident
        uint32_t pk[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 b0 = __ldg(bs + i), b1 = __ldg(bs + 16 + i);
          float a = __uint_as_float(v0[2 * i]) + b0.x;
          float b = __uint_as_float(v0[2 * i + 1]) + b0.y;
          float c = __uint_as_float(v1[2 * i]) + b1.x;
          float d = __uint_as_float(v1[2 * i + 1]) + b1.y;
          if (p.relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); c = fmaxf(c, 0.f); d = fmaxf(d, 0.f); }
          __nv_bfloat162 lo = __floats2bfloat162_rn(a, b);
          __nv_bfloat162 hi = __floats2bfloat162_rn(c, d);
          pk[i] = *reinterpret_cast<uint32_t*>(&lo);
          pk[16 + i] = *reinterpret_cast<uint32_t*>(&hi);
        }
        bool writer = true;
        int row = m;
        if (p.pool) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            uint32_t u = pk[i];
            u = bf162_max(u, __shfl_xor_sync(0xffffffffu, u, 1));
            u = bf162_max(u, __shfl_xor_sync(0xffffffffu, u, p.w_t));
            pk[i] = u;
          }
          writer = ((w_i | h_i) & 1) == 0;
          row = (h_i >> 1) * (p.w_t >> 1) + (w_i >> 1);
        }
        if (et < 32) {
          if (elect_one()) tma_store_wait_read<0>();
          __syncwarp();
        }
        named_bar_sync(3 + eg, 128);
        if (writer) {
          uint8_t* rowp = stage_out + row * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            uint4 val = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
            *reinterpret_cast<uint4*>(rowp + ((c ^ (row & 7)) << 4)) = val;
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(3 + eg, 128);
        if (et < 32) {
          if (elect_one()) {
            const int sh = p.pool ? 1 : 0;
            tma_store_4d(&tmO, stage_out, chunk * 64, w0 >> sh, h0 >> sh, n0);
            tma_store_commit();
          }
          __syncwarp();
        }
      }
    }
    if (et < 32) {
      if (elect_one()) tma_store_wait_all();
      __syncwarp();
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == kMmaWarp) tmem_dealloc2(tmem_base, TMEM_COLS);
}

}  // namespace va
