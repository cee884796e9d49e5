// K2/K3: implicit-GEMM 3x3 (and 1x1 / fully-connected) convolution on the 5th-gen tensor cores.
//
// Replaces the reference's `self.features(ip)` + classifier Linear layers, i.e. torchvision VGG16
// Conv2d(3x3,s1,p1)+bias+ReLU(+MaxPool2d(2,2)) and nn.Linear+ReLU
// (reference Sheet03/spatialModel.py:171-177, temporalModel.py:200-206; model surgery :136-152).
//
// GEMM view: D[pixel, cout] = sum_{tap, cin} X[pixel + tap, cin] * Wt[tap, cout, cin].
//   * activations: NHWC bf16, channels padded to a multiple of CK
//   * weights:     [tap' = s*ks + r][Cout][Cin_pad] bf16 (K-major "B" operand)
//   * an M tile is a 128-pixel patch n_t x h_t x w_t of the output; for tap (r,s) the A operand is the
//     same patch of the input shifted by (r-pad, s-pad), fetched by ONE 4-D tiled TMA box whose
//     out-of-bounds part (the zero halo, and partial batches) is zero-filled by the TMA unit.
//   * R==3 ("vertical reuse"): one TMA box of h_t+2 rows serves the three taps r=0..2 of a filter
//     column s; the three MMAs address it at row offsets r*w_t (whole 8-row swizzle atoms).
//   * S==3 ("whole-filter stage", first layer): a stage holds the three column-shifted boxes and all 9
//     weight taps, i.e. one pipeline stage per tile -- the K=27/180 layers are otherwise bound by the
//     per-stage latency of the single producer / MMA threads, not by bytes.
//   * WRES ("resident weights", Cout == BN and one channel chunk: conv1_1, conv1_2): all 9 taps' weights are
//     TMA-loaded into smem once per CTA instead of once per tile.  These layers are shared-memory-bandwidth
//     bound (MMA operand reads + TMA writes + epilogue staging share ~128-156 B/clk), so every byte not
//     re-written per tile is time.
//   * accumulators live in TMEM (2 stages x BN fp32 columns) so the epilogue of tile i overlaps the
//     MMAs of tile i+1; epilogue = tcgen05.ld -> +bias -> ReLU -> bf16 -> (2x2 max-pool by warp
//     shuffles) -> 128B-swizzled smem staging -> TMA store (clips partial tiles).
//
// Warp roles (320 threads, 1 CTA/SM, persistent over tiles): warps 0..3 and 4..7 = two epilogue warpgroups (TMEM
// lane quarter = warp_idx % 4), warp 8 = TMA producer, warp 9 = MMA issuer (+TMEM alloc).  The two latency-critical
// single-lane roles get the HIGHEST warp ids because the SM sub-partition scheduler prefers higher warp ids
// (B300_MICROARCH.md "hi-wid-first"): they share schedulers with busy epilogue warps and must win the issue slot.
// Group g owns
// TMEM accumulator stage g and drains the CTA's even / odd tiles, so two tile epilogues run concurrently: for
// the N=64/128 layers one 128-thread epilogue (~1.5k clk per 64 columns) was slower than the tile's MMAs.
#pragma once
#include <cuda.h>
#include "va_ptx.cuh"

namespace va {

// Division by a runtime constant as multiply-high + shift (exact for dividends < 2^31).  The tile decode runs on
// the single producer thread, where a hardware-less 32-bit divide (~40 dependent instructions) per coordinate was
// the per-tile bottleneck of the narrow layers.
struct FastDiv {
  uint32_t mul, shr, d;
  __host__ static FastDiv make(uint32_t d) {
    FastDiv f;
    f.d = d;
    if (d == 1) { f.mul = 0; f.shr = 0; return f; }
    uint32_t l = 0;
    while ((1u << l) < d) ++l;                     // ceil(log2(d))
    const uint64_t p = 31 + l;
    f.mul = (uint32_t)(((1ull << p) + d - 1) / d);
    f.shr = (uint32_t)(p - 32);
    return f;
  }
  __device__ __forceinline__ uint32_t div(uint32_t x) const { return d == 1 ? x : (__umulhi(x, mul) >> shr); }
  __device__ __forceinline__ void divmod(uint32_t x, uint32_t& q, uint32_t& r) const { q = div(x); r = x - q * d; }
};

struct ConvKernelParams {
  int n_img, H, W;                 // input (== pre-pool output) spatial size
  int h_t, w_t, n_t;               // output-pixel tile; h_t*w_t*n_t == 128, all powers of two
  int log2_w_t, log2_h_t;
  int tiles_w, tiles_h, tiles_n;   // ceil-div tile counts
  int n_tiles_cout;                // Cout / BN
  int total_tiles;
  FastDiv div_cout, div_w, div_h;  // by n_tiles_cout, tiles_w, tiles_h
  int ks, pad;                     // 3/1 or 1/0
  int cin_chunks;                  // Cin_pad / CK
  int pool, relu, out_f32;
  int split6;                      // fp32-accuracy mode: outputs written as 6 bf16 slice blocks (hi,hi,hi,mid,mid,lo)
  int Cout;
  int num_stages;
  uint32_t a_box_bytes;            // smem bytes reserved per A box (multiple of 1024); a stage holds S of them
  uint32_t a_tx_bytes;             // bytes one A TMA box delivers
  uint32_t staging_bytes;          // one output staging buffer (16 KB, or 4 KB when the 2x2 pool is fused)
  const float* bias;               // [Cout]
  float* out_f32_ptr;              // [n_img, Cout] when out_f32 (fully-connected only)
  long long* dbg;                  // optional [16] per-role cycle counters written by CTA 0 (diagnostics only)
  int hb_pitch;                    // HALO variant: pixels per row of the haloed A box (w_t + 2)
  int ksplit;                      // fully-connected split-K: number of K slices (1 = off); a tile index then decodes as
  int kb_per_split;                //   ((slice * n_tiles_cout + nt) * tiles_m + mt) and covers kb_per_split K blocks
  int tiles_m;                     // M tiles (split-K decode)
};

// tile index -> (K slice, M tile, N tile).  Without split-K the N tile is fastest (concurrent CTAs share the A patch in
// L2); with split-K (fully-connected layers, weight-bandwidth bound) the M tile is fastest, so the CTAs that run at the
// same time read the SAME weight block and HBM sees every weight byte once.
__device__ __forceinline__ void decode_tile(const ConvKernelParams& p, uint32_t tile, uint32_t& ks_i, uint32_t& mt, uint32_t& nt) {
  if (p.ksplit > 1) {
    const uint32_t rest = tile / (uint32_t)p.tiles_m;
    mt = tile - rest * (uint32_t)p.tiles_m;
    p.div_cout.divmod(rest, ks_i, nt);
  } else {
    ks_i = 0;
    p.div_cout.divmod(tile, mt, nt);
  }
}

// wait on an mbarrier, charging the stalled cycles to *acc when diagnostics are on
__device__ __forceinline__ void mbar_wait_t(uint64_t* bar, uint32_t parity, int tag, bool on, long long& acc) {
  if (!on) { mbar_wait(bar, parity, tag); return; }
  const long long t0 = clock64();
  mbar_wait(bar, parity, tag);
  acc += clock64() - t0;
}

constexpr int kConvThreads = 320;
constexpr int kMaxStages = 8;

__host__ __device__ constexpr uint32_t conv_b_stage_bytes(int BN, int CK, int R, int S) { return R * S * BN * CK * 2; }

inline size_t conv_smem_bytes(int BN, int CK, int R, int S, bool wres, int groups, uint32_t a_box_bytes,
                              uint32_t staging_bytes, int stages, int a_boxes = 0) {
  const size_t b = conv_b_stage_bytes(BN, CK, R, S);
  if (a_boxes == 0) a_boxes = S;
  return 1024 /*align slack*/ + (wres ? (size_t)groups * b : 0) + (size_t)stages * (a_boxes * a_box_bytes + (wres ? 0 : b)) +
         2 * (size_t)staging_bytes + 2 * 256 * sizeof(float) + 256;
}

__device__ __forceinline__ uint32_t bf162_max(uint32_t a, uint32_t b) {
  __nv_bfloat162 x = *reinterpret_cast<__nv_bfloat162*>(&a);
  __nv_bfloat162 y = *reinterpret_cast<__nv_bfloat162*>(&b);
  __nv_bfloat162 m = __hmax2(x, y);
  return *reinterpret_cast<uint32_t*>(&m);
}

// DRAIN (fp32-accuracy mode only): the tensor core adds into its fp32 accumulator with truncation, so a long K loop
// drifts by ~1e-5 per layer (measured: descriptors 4e-4 off after 16 layers).  With DRAIN every pipeline stage gets a
// fresh TMEM accumulator that epilogue group 0 drains and sums in registers with round-to-nearest fp32 adds.
// HALO (conv1_2: Cin = Cout = 64, resident weights): ONE TMA box of (h_t+2) x pitch pixels per tile serves all nine
// taps.  Tap (r,s) reads the box from pixel row r*pitch + s on, 8-pixel groups (one output row of the 16x8 tile) are
// pitch*128 B apart (SBO).  The start address is then no longer a multiple of the 1024-byte swizzle atom; measured on
// B200 (tools/diag_block1.py, profiles/r02_conv1_2_variants.log): the 128B swizzle of both TMA and UMMA is a function
// of the ABSOLUTE shared-memory address bits, so such a window reads exactly what TMA wrote with the descriptor's
// base-offset field left 0 (setting it to (start >> 7) & 7 gives wrong results).  36 MMAs per pipeline stage, 23 KB
// (instead of 60 KB) of smem writes per tile: 1.171 -> 0.943 ms per 250 snippets, bit-identical output.
template <int BN, int CK, int R, int S, bool WRES, bool DRAIN = false, bool HALO = false>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ CUtensorMap tmO, const ConvKernelParams p) {
  static_assert(BN == 64 || BN == 128 || BN == 256, "BN");
  static_assert(CK == 16 || CK == 32 || CK == 64, "CK");
  static_assert(R == 1 || R == 3, "R");
  static_assert(S == 1 || (S == 3 && R == 3), "S");
  static_assert(!HALO || (S == 3 && R == 3 && WRES && CK == 64), "HALO");
  constexpr int ROWB = CK * 2;
  constexpr uint32_t B_STAGE = conv_b_stage_bytes(BN, CK, R, S);
  constexpr uint32_t TMEM_COLS = 2 * BN;   // 128 / 256 / 512: powers of two >= 32

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int groups = (p.ks * p.ks) / (R * S);    // stages per channel chunk: 9 (per tap), 3 (per filter column) or 1
  const int num_kb = p.ksplit > 1 ? p.kb_per_split : groups * p.cin_chunks;      // pipeline stages consumed per tile
  const uint32_t wres_bytes = WRES ? (uint32_t)groups * B_STAGE : 0u;   // resident weights sit in front of the ring
  const uint32_t a_stage_bytes = (HALO ? 1 : S) * p.a_box_bytes;
  const uint32_t stage_bytes = a_stage_bytes + (WRES ? 0u : B_STAGE);
  uint8_t* ring = smem + wres_bytes;
  uint8_t* staging = ring + (size_t)p.num_stages * stage_bytes;
  float* bias_s = reinterpret_cast<float*>(staging + 2 * p.staging_bytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bias_s + 2 * 256);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tfull_bar = empty_bar + kMaxStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* wres_bar = tempty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wres_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

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
      mbar_init(&tempty_bar[i], 4);   // one arrive per epilogue warp
    }
    mbar_init(wres_bar, 1);
    fence_mbar_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kProducerWarp) {
    // ===================================================== TMA producer
    // The whole warp walks the loop convergently and ONE elected lane issues: inside a divergent `if (lane == 0)`
    // every uniform-register operand of UTMALDG / UTCHMMA is fetched through a per-instruction "waterfall" loop
    // (ELECT + R2UR.BROADCAST + BRA.U.ANY, ~100 clk per MMA measured), which bounded the narrow-N layers.
    const bool dbg = p.dbg != nullptr && blockIdx.x == 0;
    long long t_wait = 0, t_begin = clock64();
    uint32_t stage = 0, phase = 0;
    if (WRES) {   // the layer's whole weight set (Cout == BN, one channel chunk), once per CTA
      if (elect_one()) {
        mbar_arrive_expect_tx(wres_bar, wres_bytes);
        for (int g = 0; g < groups; ++g) tma_load_3d(smem + (size_t)g * B_STAGE, &tmW, wres_bar, 0, 0, g * (R * S));
      }
      __syncwarp();
    }
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      uint32_t nt, mt, tw, th, tn, ksl;
      decode_tile(p, (uint32_t)tile, ksl, mt, nt);
      p.div_w.divmod(mt, mt, tw);
      p.div_h.divmod(mt, tn, th);
      const int n0 = tn * p.n_t, h0 = th * p.h_t, w0 = tw * p.w_t, c0 = nt * BN;
      const int cc_begin = p.ksplit > 1 ? (int)ksl * p.kb_per_split : 0;
      const int cc_end = p.ksplit > 1 ? cc_begin + p.kb_per_split : p.cin_chunks;
      // taps are walked with counters.  R=1: g = s*ks + r; R=3: g = s; S=3: g = 0 covers the whole filter.
      int s = 0, r = 0;
      for (int g = 0; g < groups; ++g) {
        const int wx = w0 + s - p.pad, hy = h0 + r - p.pad;
        for (int cc = cc_begin; cc < cc_end; ++cc) {
          mbar_wait_t(&empty_bar[stage], phase ^ 1, 100 + stage, dbg, t_wait);
          if (elect_one()) {
            uint8_t* a_dst = ring + (size_t)stage * stage_bytes;
            mbar_arrive_expect_tx(&full_bar[stage], (HALO ? 1 : S) * p.a_tx_bytes + (WRES ? 0u : B_STAGE));
#pragma unroll
            for (int sa = 0; sa < (HALO ? 1 : S); ++sa)
              tma_load_4d(a_dst + sa * p.a_box_bytes, &tmA, &full_bar[stage], cc * CK, wx + sa, hy, n0);
            if (!WRES) tma_load_3d(a_dst + a_stage_bytes, &tmW, &full_bar[stage], cc * CK, c0, g * (R * S));
          }
          __syncwarp();
          if (++stage == (uint32_t)p.num_stages) { stage = 0; phase ^= 1; }
        }
        if (R == 1) { if (++r == p.ks) { r = 0; ++s; } } else { ++s; }
      }
    }
    if (dbg && lane == 0) { p.dbg[0] = clock64() - t_begin; p.dbg[1] = t_wait; }
  } else if (warp == kMmaWarp) {
    // ===================================================== MMA issuer (convergent warp, one elected lane issues)
    constexpr uint32_t idesc = make_idesc_bf16(128, BN);
    const uint32_t a_r_stride = (uint32_t)p.w_t * ROWB;   // bytes per input row of the A box (R==3, n_t==1)
    const uint32_t smem_base_u32 = smem_u32(smem);
    const uint32_t ring_u32 = smem_base_u32 + wres_bytes;
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);   // provably warp-uniform
    // descriptor start-address offsets (>>4) of the S*R operand windows inside a stage: launch constants, hoisted so
    // that nothing but one 64-bit add per operand separates consecutive MMAs in the issue stream
    uint64_t a_win[S * R];
#pragma unroll
    for (int sa = 0; sa < S; ++sa)
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (HALO) {
          const uint32_t prow = (uint32_t)(r * p.hb_pitch + sa);            // first pixel row of the tap's window
          a_win[sa * R + r] = (uint64_t)((prow * ROWB) >> 4);
        } else {
          a_win[sa * R + r] = (sa * p.a_box_bytes + r * a_r_stride) >> 4;
        }
      }
    const uint64_t a_sbo_fix = HALO ? ((uint64_t)(((uint32_t)p.hb_pitch * ROWB) >> 4) << 32) - ((uint64_t)((ROWB * 8) >> 4) << 32) : 0ull;
    if (WRES) mbar_wait(wres_bar, 0, 500);
    uint32_t stage = 0, phase = 0, as = 0, as_phase = 0;
    const bool dbg = p.dbg != nullptr && blockIdx.x == 0;
    long long t_full = 0, t_tempty = 0, t_begin = clock64(), n_tiles = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      ++n_tiles;
      if (!DRAIN) {
        mbar_wait_t(&tempty_bar[as], as_phase ^ 1, 200 + as, dbg, t_tempty);
        tc_fence_after();
      }
      uint32_t d_tmem = tmem_u + as * BN;
      uint32_t acc = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        if (DRAIN) {                       // a fresh accumulator per pipeline stage
          mbar_wait_t(&tempty_bar[as], as_phase ^ 1, 200 + as, dbg, t_tempty);
          d_tmem = tmem_u + as * BN;
          acc = 0;
        }
        mbar_wait_t(&full_bar[stage], phase, 300 + stage, dbg, t_full);
        tc_fence_after();
        // descriptors differ only in their 14-bit start-address field: add (byte offset >> 4) to a base
        const uint32_t a_addr = ring_u32 + stage * stage_bytes;
        const uint64_t da0 = make_smem_desc<ROWB>(a_addr) + a_sbo_fix;
        const uint64_t db0 = make_smem_desc<ROWB>(WRES ? smem_base_u32 + (uint32_t)kb * B_STAGE : a_addr + a_stage_bytes);
        if (elect_one()) {
#pragma unroll
          for (int sa = 0; sa < S; ++sa) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
              const uint64_t a_off = a_win[sa * R + r];
#pragma unroll
              for (int k = 0; k < CK / 16; ++k) {
                umma_bf16(d_tmem, da0 + a_off + 2 * k, db0 + (((sa * R + r) * (BN * ROWB)) >> 4) + 2 * k, idesc,
                          (sa | r | k) ? 1u : acc);
              }
            }
          }
          umma_commit(&empty_bar[stage]);   // smem slot reusable once these MMAs retire
          if (DRAIN || kb == num_kb - 1) umma_commit(&tfull_bar[as]);   // accumulator complete -> epilogue
        }
        __syncwarp();
        acc = 1;
        if (++stage == (uint32_t)p.num_stages) { stage = 0; phase ^= 1; }
        if (DRAIN) { as ^= 1; if (as == 0) as_phase ^= 1; }
      }
      if (!DRAIN) { as ^= 1; if (as == 0) as_phase ^= 1; }
    }
    if (dbg && lane == 0) { p.dbg[2] = clock64() - t_begin; p.dbg[3] = t_full; p.dbg[4] = t_tempty; p.dbg[11] = n_tiles; }
  } else {
    // ===================================================== epilogue (2 groups x 4 warps)
    const int eg = warp >> 2;             // epilogue group == TMEM accumulator stage it drains
    const int q = warp & 3;               // TMEM lane quarter accessible to this warp
    const int m = q * 32 + lane;          // accumulator row == pixel index inside the tile
    const int et = threadIdx.x - eg * 128;        // 0..127 within the group
    float* bias_g = bias_s + eg * 256;
    uint8_t* stage_out = staging + eg * p.staging_bytes;
    const int w_i = m & (p.w_t - 1);
    const int h_i = (m >> p.log2_w_t) & (p.h_t - 1);
    const int n_i = m >> (p.log2_w_t + p.log2_h_t);
    uint32_t as = (uint32_t)eg;
    const bool dbg = p.dbg != nullptr && blockIdx.x == 0 && et == 0;
    long long t_tfull = 0, t_stage = 0, t_begin = clock64();
    int bias_c0 = -1;
    int it = 0;
    uint32_t drain_cnt = 0;               // DRAIN: running count of drained accumulators (stage = cnt&1, phase = cnt>>1)
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      if (DRAIN ? (eg != 0) : ((it & 1) != eg)) continue;       // the other group's tile (DRAIN: group 0 does all)
      uint32_t as_phase = (uint32_t)(it >> 1) & 1u;
      uint32_t nt, mt, tw, th, tn, ksl;
      decode_tile(p, (uint32_t)tile, ksl, mt, nt);
      p.div_w.divmod(mt, mt, tw);
      p.div_h.divmod(mt, tn, th);
      const int n0 = tn * p.n_t, h0 = th * p.h_t, w0 = tw * p.w_t, c0 = nt * BN;

      if (c0 != bias_c0) {                // warp-uniform: the bias slice changes only with the Cout tile
        named_bar_sync(1 + eg, 128);      // everyone in the group is done reading the previous slice
        for (int i = et; i < BN; i += 128) bias_g[i] = __ldg(p.bias + c0 + i);
        named_bar_sync(1 + eg, 128);
        bias_c0 = c0;
      }

      if (!DRAIN) {
        mbar_wait_t(&tfull_bar[as], as_phase, 400 + as, dbg, t_tfull);
        tc_fence_after();
      }

#pragma unroll 1
      for (int chunk = 0; chunk < BN / 64; ++chunk) {
        uint32_t v0[32], v1[32];
        if (DRAIN) {
          // BN == 64: sum the per-stage accumulators with round-to-nearest fp32 adds
          float s0[32], s1[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) { s0[i] = 0.f; s1[i] = 0.f; }
          for (int g = 0; g < num_kb; ++g, ++drain_cnt) {
            as = drain_cnt & 1u;
            as_phase = (drain_cnt >> 1) & 1u;
            mbar_wait(&tfull_bar[as], as_phase, 400 + as);
            tc_fence_after();
            const uint32_t ta = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;
            tmem_ld32(ta, v0);
            tmem_ld32(ta + 32, v1);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[as]);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              s0[i] = __fadd_rn(s0[i], __uint_as_float(v0[i]));
              s1[i] = __fadd_rn(s1[i], __uint_as_float(v1[i]));
            }
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) { v0[i] = __float_as_uint(s0[i]); v1[i] = __float_as_uint(s1[i]); }
        } else {
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + chunk * 64;
          tmem_ld32(taddr, v0);
          tmem_ld32(taddr + 32, v1);
          tmem_ld_wait();
          if (chunk == BN / 64 - 1) {        // accumulator stage fully drained into registers
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[as]);
          }
        }
        const float* bs = bias_g + chunk * 64;
        if (p.out_f32) {
          // fully-connected tail (H=W=1, tile = 128 images): fp32 rows straight to global.
          const int img = n0 + m;
          if (p.ksplit > 1) {
            // split-K partial sums: raw fp32 accumulators of this K slice; bias / ReLU / rounding happen in the reduction
            if (img < p.n_img) {
              float* dst = p.out_f32_ptr + ((size_t)ksl * p.n_img + img) * p.Cout + c0 + chunk * 64;
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                *reinterpret_cast<float4*>(dst + i) = make_float4(__uint_as_float(v0[i]), __uint_as_float(v0[i + 1]),
                                                                  __uint_as_float(v0[i + 2]), __uint_as_float(v0[i + 3]));
                *reinterpret_cast<float4*>(dst + 32 + i) = make_float4(__uint_as_float(v1[i]), __uint_as_float(v1[i + 1]),
                                                                       __uint_as_float(v1[i + 2]), __uint_as_float(v1[i + 3]));
              }
            }
            continue;
          }
          if (img < p.n_img) {
            float* dst = p.out_f32_ptr + (size_t)img * p.Cout + c0 + chunk * 64;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              float4 o;
              o.x = __uint_as_float(v0[i]) + bs[i];
              o.y = __uint_as_float(v0[i + 1]) + bs[i + 1];
              o.z = __uint_as_float(v0[i + 2]) + bs[i + 2];
              o.w = __uint_as_float(v0[i + 3]) + bs[i + 3];
              if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
              *reinterpret_cast<float4*>(dst + i) = o;
            }
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              float4 o;
              o.x = __uint_as_float(v1[i]) + bs[32 + i];
              o.y = __uint_as_float(v1[i + 1]) + bs[32 + i + 1];
              o.z = __uint_as_float(v1[i + 2]) + bs[32 + i + 2];
              o.w = __uint_as_float(v1[i + 3]) + bs[32 + i + 3];
              if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
              *reinterpret_cast<float4*>(dst + 32 + i) = o;
            }
          }
          continue;
        }
        if (p.split6) {
          // fp32-accuracy ("bf16x3") mode: keep fp32 through bias/ReLU/pool, then write the value as three bf16 slices
          // hi = bf16(x), mid = bf16(x - hi), lo = bf16(x - hi - mid) into six channel blocks (hi,hi,hi,mid,mid,lo) of
          // the next layer's K dimension; its weights are packed (hi,mid,lo,hi,mid,hi), so the GEMM sums the six
          // leading cross terms with fp32 accumulation (dropped terms are O(2^-24)).
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float a = __uint_as_float(v0[i]) + bs[i], b = __uint_as_float(v1[i]) + bs[32 + i];
            if (p.relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
            if (p.pool) {
              a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, 1));
              a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, p.w_t));
              b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, 1));
              b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, p.w_t));
            }
            v0[i] = __float_as_uint(a); v1[i] = __float_as_uint(b);
          }
          const bool wr = p.pool ? (((w_i | h_i) & 1) == 0) : true;
          const int prow = p.pool ? ((n_i * (p.h_t >> 1)) + (h_i >> 1)) * (p.w_t >> 1) + (w_i >> 1) : m;
          const int sh = p.pool ? 1 : 0;
#pragma unroll 1
          for (int sl = 0; sl < 3; ++sl) {
            uint32_t ps[32];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float a = __uint_as_float(v0[2 * i]), b = __uint_as_float(v0[2 * i + 1]);
              const float c = __uint_as_float(v1[2 * i]), d = __uint_as_float(v1[2 * i + 1]);
              __nv_bfloat162 lo2 = __floats2bfloat162_rn(a, b), hi2 = __floats2bfloat162_rn(c, d);
              ps[i] = *reinterpret_cast<uint32_t*>(&lo2);
              ps[16 + i] = *reinterpret_cast<uint32_t*>(&hi2);
              // residual for the next slice (exact in fp32)
              v0[2 * i] = __float_as_uint(__fsub_rn(a, __low2float(lo2)));
              v0[2 * i + 1] = __float_as_uint(__fsub_rn(b, __high2float(lo2)));
              v1[2 * i] = __float_as_uint(__fsub_rn(c, __low2float(hi2)));
              v1[2 * i + 1] = __float_as_uint(__fsub_rn(d, __high2float(hi2)));
            }
            if (et < 32) {
              if (elect_one()) tma_store_wait_read<0>();
              __syncwarp();
            }
            named_bar_sync(3 + eg, 128);
            if (wr) {
              uint8_t* rowp = stage_out + prow * 128;
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                uint4 val = make_uint4(ps[4 * c], ps[4 * c + 1], ps[4 * c + 2], ps[4 * c + 3]);
                *reinterpret_cast<uint4*>(rowp + ((c ^ (prow & 7)) << 4)) = val;
              }
            }
            fence_proxy_async_smem();
            named_bar_sync(3 + eg, 128);
            if (et < 32) {
              if (elect_one()) {
                // slice sl goes to blocks {0,1,2} (hi), {3,4} (mid), {5} (lo)
                const int b0 = sl == 0 ? 0 : (sl == 1 ? 3 : 5), b1 = sl == 0 ? 3 : (sl == 1 ? 5 : 6);
                for (int blk = b0; blk < b1; ++blk)
                  tma_store_4d(&tmO, stage_out, blk * p.Cout + c0 + chunk * 64, w0 >> sh, h0 >> sh, n0);
                tma_store_commit();
              }
              __syncwarp();
            }
          }
          continue;
        }
        // bias + ReLU + round to bf16 (pairs of channels packed low|high)
        uint32_t pk[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float a = __uint_as_float(v0[2 * i]) + bs[2 * i];
          float b = __uint_as_float(v0[2 * i + 1]) + bs[2 * i + 1];
          float c = __uint_as_float(v1[2 * i]) + bs[32 + 2 * i];
          float d = __uint_as_float(v1[2 * i + 1]) + bs[32 + 2 * i + 1];
          if (p.relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); c = fmaxf(c, 0.f); d = fmaxf(d, 0.f); }
          __nv_bfloat162 lo = __floats2bfloat162_rn(a, b);
          __nv_bfloat162 hi = __floats2bfloat162_rn(c, d);
          pk[i] = *reinterpret_cast<uint32_t*>(&lo);
          pk[16 + i] = *reinterpret_cast<uint32_t*>(&hi);
        }
        bool writer = true;
        int row = m;
        if (p.pool) {
          // 2x2/2 max-pool: partners are the neighbouring pixel (lane^1) and the next tile row (lane^w_t).
          // max commutes with the (monotonic) bf16 rounding and ReLU, so pooling the rounded values is exact.
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
        // the group's staging buffer was last read by the TMA store of its previous chunk
        const long long ts0 = dbg ? clock64() : 0;
        if (et < 32) {                       // the group's first warp; its elected lane owns the bulk-store groups
          if (elect_one()) tma_store_wait_read<0>();
          __syncwarp();
        }
        named_bar_sync(3 + eg, 128);
        if (dbg) t_stage += clock64() - ts0;
        if (writer) {
          uint8_t* rowp = stage_out + row * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            uint4 val = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
            *reinterpret_cast<uint4*>(rowp + ((c ^ (row & 7)) << 4)) = val;   // 128B swizzle, as the TMA store expects
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
    if (dbg) { p.dbg[5 + 3 * eg] = clock64() - t_begin; p.dbg[6 + 3 * eg] = t_tfull; p.dbg[7 + 3 * eg] = t_stage; }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace va
