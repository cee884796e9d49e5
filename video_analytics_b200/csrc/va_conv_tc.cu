// Host-side planner/launcher for the tcgen05 implicit-GEMM layer kernel (va_conv_tc.cuh).
#include "va_internal.h"
#include "va_conv_tc.cuh"
#include "va_conv_tc2.cuh"

#include <mutex>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

namespace va {

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
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

thread_local char g_err[512];

const char* errf(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return g_err;
}

// bf16 tensor map of rank `rank`; dims[0] is the contiguous (channel) dimension.
const char* encode_bf16(CUtensorMap* m, const void* addr, int rank, const uint64_t* dims, const uint32_t* box,
                        int inner_bytes, CUtensorMapL2promotion promo) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return "cuTensorMapEncodeTiled not available (no CUDA driver?)";
  cuuint64_t gdim[5], gstr[5];
  cuuint32_t bx[5], es[5];
  uint64_t stride = 2;
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    stride *= dims[i];
    if (i < rank - 1) gstr[i] = stride;
    bx[i] = box[i];
    es[i] = 1;
  }
  const CUtensorMapSwizzle sw = inner_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                              : inner_bytes == 64  ? CU_TENSOR_MAP_SWIZZLE_64B
                                                   : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(addr), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return errf("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu,%llu] box [%u,%u,%u,%u] inner %dB",
                (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
                (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
                box[1], rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0, inner_bytes);
  return nullptr;
}

int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

// SM reservation (va_reserve_sms): the next `g_reserve_launches` layer-kernel launches size their persistent grids for
// `sm_total - g_reserve_sms` SMs, so that a collective launched just before (NCCL's CTAs cannot co-reside with a layer CTA
// that holds 210-227 KB of shared memory) finds free SMs instead of stalling a whole layer behind it.
int g_reserve_sms = 0;
int g_reserve_launches = 0;

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  if (g_reserve_launches > 0 && g_reserve_sms > 0) {
    int m = n - g_reserve_sms;
    m &= ~1;                                   // CTA pairs
    if (m >= 16) return m;
  }
  return n;
}

template <int BN, int CK, int R, int S, bool WRES, bool DRAIN = false, bool HALO = false>
const char* launch_variant(const CUtensorMap& tA, const CUtensorMap& tW, const CUtensorMap& tO,
                           const ConvKernelParams& p, int grid, size_t smem, cudaStream_t st) {
  auto kfn = conv_tc_kernel<BN, CK, R, S, WRES, DRAIN, HALO>;
  static size_t configured = 0;
  if (configured < smem) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return errf("cudaFuncSetAttribute(smem=%zu): %s", smem, cudaGetErrorString(e));
    configured = smem;
  }
  count_launch();
  kfn<<<grid, kConvThreads, smem, st>>>(tA, tW, tO, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return errf("conv_tc_kernel<%d,%d,%d,%d,%d> launch: %s", BN, CK, R, S, (int)WRES, cudaGetErrorString(e));
  return nullptr;
}

// Second pass of the split-K fully-connected layers: y[i][c] = act(sum_s part[s][i][c] + bias[c]), slices added in the fixed
// order s = 0..S-1 (so the result does not depend on which CTA finished first, nor on the batch size).
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float4* __restrict__ part, int S, long long n4, int cout4,
                                                            const float4* __restrict__ bias, int relu,
                                                            uint2* __restrict__ y_bf16, float4* __restrict__ y_f32) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 acc = __ldg(part + i);
    for (int s = 1; s < S; ++s) {
      const float4 v = __ldg(part + (size_t)s * n4 + i);
      acc.x = __fadd_rn(acc.x, v.x); acc.y = __fadd_rn(acc.y, v.y); acc.z = __fadd_rn(acc.z, v.z); acc.w = __fadd_rn(acc.w, v.w);
    }
    const float4 b = __ldg(bias + (int)(i % cout4));
    acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
    if (relu) { acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f); }
    if (y_f32) y_f32[i] = acc;
    if (y_bf16) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(acc.x, acc.y), hi = __floats2bfloat162_rn(acc.z, acc.w);
      y_bf16[i] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    }
  }
}

}  // namespace

static long long* g_dbg_counters = nullptr;
void conv_set_debug_counters(long long* dev_buf) { g_dbg_counters = dev_buf; }
long long* conv_get_debug_counters() { return g_dbg_counters; }

void conv_reserve_sms(int sms, int launches) {
  g_reserve_sms = sms > 0 ? sms : 0;
  g_reserve_launches = (sms > 0 && launches > 0) ? launches : 0;
}

const char* conv_layer_run(const ConvLayerDesc& d, cudaStream_t st) {
  cudaStream_t st_ = st;
  if (d.n <= 0) return nullptr;
  struct ReserveTick { ~ReserveTick() { if (g_reserve_launches > 0) --g_reserve_launches; } } reserve_tick;   // one layer = one tick
  if (d.ks != 1 && d.ks != 3) return "ks must be 1 or 3";
  const int CK = (d.cin_pad % 64 == 0) ? 64 : d.cin_pad;
  if (CK != 64 && CK != 32 && CK != 16) return errf("cin_pad %d unsupported (16, 32 or multiple of 64)", d.cin_pad);
  if (d.Cout % 64 != 0) return errf("Cout %d must be a multiple of 64", d.Cout);
  if (d.pool && ((d.H | d.W) & 1)) return "fused 2x2 pool needs even H and W";
  if (d.y_f32 && (d.H != 1 || d.W != 1 || d.pool)) return "fp32 output only for fully-connected (H=W=1) layers";
  // split-K for the fully-connected layers (weight-bandwidth bound: FC1 streams 205 MB through 16..64 CTAs otherwise).
  // The slice count depends on K alone, never on the batch: the summation order is the same for every chunk size.
  const bool splitk = d.splitk_ws != nullptr && d.H == 1 && d.W == 1 && d.ks == 1 && !d.split6 && CK == 64 &&
                      (d.cin_pad / 64) >= 32 && (d.cin_pad / 64) % kSplitK == 0;

  ConvKernelParams p;
  p.n_img = d.n; p.H = d.H; p.W = d.W;
  p.hb_pitch = 0;
  // HALO variant (one haloed A box per 16x8 tile serves all nine taps; resident weights): the 64->64 3x3 layers
  // (conv1_2 and its data gradient).  force_r = 10 requests it, any other non-zero force_r disables it.
  const bool halo_ok = d.ks == 3 && CK == 64 && d.cin_pad == 64 && d.Cout == 64 && d.H % 16 == 0 && d.W % 8 == 0 && !d.split6 &&
                       (d.force_bn == 0 || d.force_bn == 64);
  if (d.force_r == 10 && !halo_ok) return "HALO variant needs a 3x3 64->64 layer with H%16==0 and W%8==0";
  // CTA-pair HALO kernel with resident weight halves (conv_tc2h_kernel): the 3x3 layers at 224^2 / 112^2 with
  // Cout, Cin in {64, 128}.  force_r = 11 requests it, any other non-zero force_r disables it.
  const bool pairh_ok = d.ks == 3 && CK == 64 && (d.cin_pad == 64 || d.cin_pad == 128) && (d.Cout == 64 || d.Cout == 128) &&
                        d.H % 16 == 0 && d.W % 8 == 0 && !d.split6 && !d.y_f32 && (d.force_bn == 0 || d.force_bn == d.Cout);
  if (d.force_r == 11 && !pairh_ok) return "pair-HALO variant needs a 3x3 layer with Cout, Cin in {64,128}, H%16==0, W%8==0";
  if (pairh_ok && (d.force_r == 11 || d.force_r == 0)) {
    const int BNh = d.Cout, chunks = d.cin_pad / 64;
    p.h_t = 16; p.w_t = 8; p.n_t = 1; p.hb_pitch = p.w_t + 2;
    p.log2_w_t = 3; p.log2_h_t = 4;
    p.tiles_w = d.W / p.w_t; p.tiles_h = d.H / p.h_t; p.tiles_n = d.n;
    const int tiles_m_h = p.tiles_w * p.tiles_h * p.tiles_n;
    p.n_tiles_cout = 1;
    p.total_tiles = (tiles_m_h + 1) / 2;                          // work units of a CTA pair
    p.div_cout = FastDiv::make(1u);
    p.div_w = FastDiv::make((uint32_t)p.tiles_w);
    p.div_h = FastDiv::make((uint32_t)p.tiles_h);
    p.ks = 3; p.pad = 1; p.cin_chunks = chunks;
    p.pool = d.pool; p.relu = d.relu; p.out_f32 = 0; p.split6 = 0; p.Cout = d.Cout;
    p.bias = d.bias; p.out_f32_ptr = nullptr; p.dbg = nullptr;
    p.a_tx_bytes = (uint32_t)((p.h_t + 2) * p.hb_pitch * 128);
    p.a_box_bytes = (p.a_tx_bytes + 1023u) & ~1023u;
    p.staging_bytes = d.pool ? 4096u : 16384u;
    int st = (int)((227 * 1024 - conv2h_smem_bytes(BNh, chunks, p.a_box_bytes, p.staging_bytes, 0)) / p.a_box_bytes);
    if (st > kMaxStages) st = kMaxStages;
    if (st >= 2) {
      p.num_stages = st;
      const size_t smem_h = conv2h_smem_bytes(BNh, chunks, p.a_box_bytes, p.staging_bytes, st);
      CUtensorMap tA, tW, tO;
      {
        const uint64_t dims[4] = {(uint64_t)d.cin_pad, (uint64_t)d.W, (uint64_t)d.H, (uint64_t)d.n};
        const uint32_t box[4] = {64u, (uint32_t)p.hb_pitch, (uint32_t)(p.h_t + 2), 1u};
        if (const char* e = encode_bf16(&tA, d.x, 4, dims, box, 128, CU_TENSOR_MAP_L2_PROMOTION_L2_128B)) return e;
      }
      {
        const uint64_t dims[3] = {(uint64_t)d.cin_pad, (uint64_t)d.Cout, 9ull};
        const uint32_t box[3] = {64u, (uint32_t)(BNh / 2), 9u};
        if (const char* e = encode_bf16(&tW, d.w_packed, 3, dims, box, 128, CU_TENSOR_MAP_L2_PROMOTION_L2_256B)) return e;
      }
      {
        const int sh = d.pool ? 1 : 0;
        const uint64_t dims[4] = {(uint64_t)d.Cout, (uint64_t)(d.W >> sh), (uint64_t)(d.H >> sh), (uint64_t)d.n};
        const uint32_t box[4] = {64u, (uint32_t)(p.w_t >> sh), (uint32_t)(p.h_t >> sh), 1u};
        if (const char* e = encode_bf16(&tO, d.y, 4, dims, box, 128, CU_TENSOR_MAP_L2_PROMOTION_L2_128B)) return e;
      }
      const int sms_h = sm_count();
      const int clusters = p.total_tiles < sms_h / 2 ? p.total_tiles : sms_h / 2;
      static size_t configured_h[2] = {0, 0};
      const int vi = BNh == 64 ? 0 : 1;
      if (configured_h[vi] < smem_h) {
        cudaError_t e = BNh == 64 ? cudaFuncSetAttribute(conv_tc2h_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_h)
                                  : cudaFuncSetAttribute(conv_tc2h_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_h);
        if (e != cudaSuccess) return errf("cudaFuncSetAttribute(pair-halo, smem=%zu): %s", smem_h, cudaGetErrorString(e));
        configured_h[vi] = smem_h;
      }
      count_launch();
      if (BNh == 64) conv_tc2h_kernel<64><<<2 * clusters, kConv2hThreads, smem_h, st_>>>(tA, tW, tO, p);
      else conv_tc2h_kernel<128><<<2 * clusters, kConv2hThreads, smem_h, st_>>>(tA, tW, tO, p);
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) return errf("conv_tc2h_kernel<%d> launch: %s", BNh, cudaGetErrorString(e));
      return nullptr;
    }
    if (d.force_r == 11) return "pair-HALO variant: not enough shared memory for 2 stages";
  }
  const bool halo = halo_ok && (d.force_r == 10 || d.force_r == 0);
  if (halo) {
    p.h_t = 16; p.w_t = 8; p.n_t = 1;
    p.hb_pitch = p.w_t + 2;
  }
  else if (d.H == 1 && d.W == 1) { p.h_t = 1; p.w_t = 1; p.n_t = 128; }
  else if (d.W % 16 == 0 && d.H % 8 == 0) { p.h_t = 8; p.w_t = 16; p.n_t = 1; }
  else if (d.W % 8 == 0 && d.H % 8 == 0) { p.h_t = 8; p.w_t = 8; p.n_t = 2; }
  else if (d.W % 4 == 0 && d.H % 4 == 0) { p.h_t = 4; p.w_t = 4; p.n_t = 8; }
  else if (d.W % 2 == 0 && d.H % 2 == 0) { p.h_t = 2; p.w_t = 2; p.n_t = 32; }
  else return errf("unsupported spatial size %dx%d", d.H, d.W);
  if (d.pool && (p.h_t < 2 || p.w_t < 2)) return "pool needs a spatial tile";
  p.log2_w_t = ilog2(p.w_t); p.log2_h_t = ilog2(p.h_t);
  p.tiles_w = d.W / p.w_t; p.tiles_h = d.H / p.h_t; p.tiles_n = (d.n + p.n_t - 1) / p.n_t;
  const int tiles_m = p.tiles_w * p.tiles_h * p.tiles_n;

  // Kernel variant.  R=3 (one haloed A box serves a filter column) needs single-image tiles whose rows are whole
  // swizzle atoms; it pays where the N tile is narrow (Cout <= 128, i.e. the 224^2 / 112^2 layers), where the
  // layer is L2->SM bound.  S=3 (whole filter per stage) is for the first layer (Cin_pad 16/32, one channel chunk).
  const bool r3_ok = d.ks == 3 && p.n_t == 1 && p.w_t % 8 == 0;
  int R = 1, S = 1;
  if (d.split6) {
    // fp32-accuracy mode: one tap per stage, BN = 64, per-stage accumulator drain (kernel template DRAIN)
  } else if (halo) {
    R = 3;
  } else if (d.force_r == 3) {
    if (!r3_ok) return "force_r=3 needs ks=3 and a single-image tile with w_t%8==0";
    R = 3;
  } else if (d.force_r == 0 && r3_ok && d.Cout <= 128) {
    R = 3;
  }
  if (R == 3 && CK < 64 && d.force_r != 3) S = 3;
  if (d.force_r == 9 || halo) {   // whole-filter stages
    if (!r3_ok || (CK == 64 && (d.cin_pad != 64 || d.Cout != 64))) return "force_r=9 (S=3) needs an R=3-capable tile and Cin_pad 16/32, or a 64->64 layer";
    R = 3; S = 3;
  }
  const int sms = sm_count();
  int BN = d.split6 ? 64 : d.force_bn;
  if (BN == 0) {
    double best = 1e30;
    const int cands[3] = {256, 128, 64};
    for (int i = 0; i < 3; ++i) {
      const int bn = cands[i];
      if (d.Cout % bn) continue;
      if (CK != 64 && bn != 64) continue;
      if (R == 3 && bn == 256) continue;
      const long long tiles = (long long)tiles_m * (d.Cout / bn) * (splitk ? kSplitK : 1);
      const long long waves = (tiles + sms - 1) / sms;
      // relative cost per output column, measured on B200 (profiles/r01_diag_forward_call1.log): narrower N tiles
      // re-read the A patch from L2 more often and are L2->SM bandwidth bound.
      const double eff = bn == 256 ? 1.0 : (bn == 128 ? 1.5 : 2.6);
      const double cost = (double)waves * bn * eff;
      if (cost < best) { best = cost; BN = bn; }
    }
  }
  if (BN != 64 && BN != 128 && BN != 256) return errf("bad BN %d", BN);
  if (d.Cout % BN) return errf("Cout %d not divisible by BN %d", d.Cout, BN);
  if (CK != 64 && BN != 64) return "CK<64 variants are built for BN=64 only";
  if (R == 3 && BN == 256) return "R=3 not built for BN=256";

  // CTA-pair kernel (cta_group::2, UMMA M=256) for the wide 3x3 layers: 6 x 32 KB stages instead of 4 x 48 KB.
  const bool pair = BN == 256 && CK == 64 && R == 1 && S == 1 && d.ks == 3 && !d.y_f32 && d.force_r != 1 && !d.split6 &&
                    true;
  p.n_tiles_cout = d.Cout / BN;
  p.total_tiles = tiles_m * p.n_tiles_cout * (splitk ? kSplitK : 1);
  p.ksplit = splitk ? kSplitK : 1;
  p.kb_per_split = splitk ? (d.cin_pad / 64) / kSplitK : 0;
  p.tiles_m = tiles_m;
  p.div_cout = FastDiv::make((uint32_t)p.n_tiles_cout);
  p.div_w = FastDiv::make((uint32_t)p.tiles_w);
  p.div_h = FastDiv::make((uint32_t)p.tiles_h);
  p.ks = d.ks; p.pad = (d.ks - 1) / 2;
  p.cin_chunks = d.cin_pad / CK;
  p.pool = d.pool; p.relu = d.relu; p.out_f32 = (d.y_f32 || splitk) ? 1 : 0;
  p.split6 = (d.split6 && !d.y_f32) ? 1 : 0;
  p.Cout = d.Cout;
  p.bias = d.bias; p.out_f32_ptr = splitk ? d.splitk_ws : d.y_f32; p.dbg = g_dbg_counters;
  const int rowb = CK * 2;
  const int a_rows = p.n_t * (p.h_t + (R - 1)) * (halo ? p.hb_pitch : p.w_t);
  p.a_tx_bytes = (uint32_t)a_rows * rowb;
  p.a_box_bytes = (p.a_tx_bytes + 1023u) & ~1023u;
  p.staging_bytes = d.pool ? 4096u : 16384u;
  // Resident weights: the layer's whole weight set stays in smem (one Cout tile, one channel chunk, R=3 variants)
  const int groups = (d.ks * d.ks) / (R * S);
  const bool wres = !d.split6 && R == 3 && p.n_tiles_cout == 1 && p.cin_chunks == 1 && BN == 64 && true &&
                    (size_t)groups * conv_b_stage_bytes(BN, CK, R, S) <= 80 * 1024;
  if (halo && !wres) return "HALO variant needs resident weights";
  const int a_boxes = halo ? 1 : S;
  const uint32_t stage_bytes = a_boxes * p.a_box_bytes + (wres ? 0u : conv_b_stage_bytes(BN, CK, R, S));
  const size_t smem_cap = 227 * 1024;
  int stages = (int)((smem_cap - conv_smem_bytes(BN, CK, R, S, wres, groups, p.a_box_bytes, p.staging_bytes, 0, a_boxes)) / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return errf("not enough shared memory for 2 stages (stage %u B)", stage_bytes);
  p.num_stages = stages;
  const size_t smem = conv_smem_bytes(BN, CK, R, S, wres, groups, p.a_box_bytes, p.staging_bytes, stages, a_boxes);

  CUtensorMap tA, tW, tO;
  {
    const uint64_t dims[4] = {(uint64_t)d.cin_pad, (uint64_t)d.W, (uint64_t)d.H, (uint64_t)d.n};
    const uint32_t box[4] = {(uint32_t)CK, (uint32_t)(halo ? p.hb_pitch : p.w_t), (uint32_t)(p.h_t + R - 1), (uint32_t)p.n_t};
    if (const char* e = encode_bf16(&tA, d.x, 4, dims, box, rowb, CU_TENSOR_MAP_L2_PROMOTION_L2_128B)) return e;
  }
  {
    const uint64_t dims[3] = {(uint64_t)d.cin_pad, (uint64_t)d.Cout, (uint64_t)(d.ks * d.ks)};
    const uint32_t box[3] = {(uint32_t)CK, (uint32_t)BN, (uint32_t)(R * S)};
    if (const char* e = encode_bf16(&tW, d.w_packed, 3, dims, box, rowb, CU_TENSOR_MAP_L2_PROMOTION_L2_256B)) return e;
  }
  if (!d.y_f32 && !splitk) {
    const int sh = d.pool ? 1 : 0;
    const uint64_t dims[4] = {(uint64_t)d.Cout * (d.split6 ? 6 : 1), (uint64_t)(d.W >> sh), (uint64_t)(d.H >> sh), (uint64_t)d.n};
    const uint32_t box[4] = {64u, (uint32_t)(p.w_t >> sh), (uint32_t)(p.h_t >> sh), (uint32_t)p.n_t};
    if (const char* e = encode_bf16(&tO, d.y, 4, dims, box, 128, CU_TENSOR_MAP_L2_PROMOTION_L2_128B)) return e;
  } else {
    tO = tA;   // unused by the fp32 epilogue; must still be a valid descriptor
  }
  const int grid = p.total_tiles < sms ? p.total_tiles : sms;
  if (pair) {
    const int pairs_m = (tiles_m + 1) / 2;
    ConvKernelParams p2 = p;
    p2.total_tiles = pairs_m * p.n_tiles_cout;                    // work units of a CTA pair
    const size_t smem_cap2 = 227 * 1024;
    int st2 = (int)((smem_cap2 - conv2_smem_bytes(p.staging_bytes, 0)) / 32768);
    if (st2 > kMaxStages) st2 = kMaxStages;
    p2.num_stages = st2;
    const size_t smem2 = conv2_smem_bytes(p.staging_bytes, st2);
    // weights box: this CTA's half of the N tile
    const uint64_t wdims[3] = {(uint64_t)d.cin_pad, (uint64_t)d.Cout, (uint64_t)(d.ks * d.ks)};
    const uint32_t wbox[3] = {64u, 128u, 1u};
    if (const char* e = encode_bf16(&tW, d.w_packed, 3, wdims, wbox, 128, CU_TENSOR_MAP_L2_PROMOTION_L2_256B)) return e;
    static size_t configured2 = 0;
    if (configured2 < smem2) {
      cudaError_t e = cudaFuncSetAttribute(conv_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
      if (e != cudaSuccess) return errf("cudaFuncSetAttribute(pair, smem=%zu): %s", smem2, cudaGetErrorString(e));
      configured2 = smem2;
    }
    int clusters = p2.total_tiles < sms / 2 ? p2.total_tiles : sms / 2;
    count_launch();
    conv_tc2_kernel<<<2 * clusters, kConv2Threads, smem2, st>>>(tA, tW, tO, p2);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return errf("conv_tc2_kernel launch: %s", cudaGetErrorString(e));
    return nullptr;
  }

  if (d.split6) {
    if (BN == 64 && CK == 32) return launch_variant<64, 32, 1, 1, false, true>(tA, tW, tO, p, grid, smem, st);
    if (BN == 64 && CK == 64) return launch_variant<64, 64, 1, 1, false, true>(tA, tW, tO, p, grid, smem, st);
    return errf("fp32-accuracy mode: no kernel for BN=%d CK=%d", BN, CK);
  }
  if (halo) return launch_variant<64, 64, 3, 3, true, false, true>(tA, tW, tO, p, grid, smem, st);
  if (splitk) {
    const char* e = nullptr;
    if (BN == 256) e = launch_variant<256, 64, 1, 1, false>(tA, tW, tO, p, grid, smem, st);
    else if (BN == 128) e = launch_variant<128, 64, 1, 1, false>(tA, tW, tO, p, grid, smem, st);
    else e = launch_variant<64, 64, 1, 1, false>(tA, tW, tO, p, grid, smem, st);
    if (e) return e;
    const long long n4 = (long long)d.n * d.Cout / 4;
    const int blocks = (int)std::min<long long>((n4 + 255) / 256, (long long)sms * 8);
    count_launch();
    splitk_reduce_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(d.splitk_ws), kSplitK, n4, d.Cout / 4,
                                                 reinterpret_cast<const float4*>(d.bias), d.relu,
                                                 d.y_f32 ? nullptr : reinterpret_cast<uint2*>(d.y),
                                                 d.y_f32 ? reinterpret_cast<float4*>(d.y_f32) : nullptr);
    cudaError_t ce = cudaGetLastError();
    if (ce != cudaSuccess) return errf("splitk_reduce_kernel launch: %s", cudaGetErrorString(ce));
    return nullptr;
  }
#define VA_CASE(bn, ck, r, sv, wr) \
  if (BN == bn && CK == ck && R == r && S == sv && wres == wr) \
    return launch_variant<bn, ck, r, sv, wr>(tA, tW, tO, p, grid, smem, st);
  VA_CASE(64, 16, 1, 1, false) VA_CASE(64, 32, 1, 1, false) VA_CASE(64, 64, 1, 1, false) VA_CASE(128, 64, 1, 1, false)
  VA_CASE(256, 64, 1, 1, false) VA_CASE(64, 64, 3, 1, false) VA_CASE(128, 64, 3, 1, false) VA_CASE(64, 16, 3, 1, false)
  VA_CASE(64, 32, 3, 1, false) VA_CASE(64, 16, 3, 3, false) VA_CASE(64, 32, 3, 3, false)
  VA_CASE(64, 64, 3, 1, true) VA_CASE(64, 16, 3, 1, true) VA_CASE(64, 32, 3, 1, true) VA_CASE(64, 16, 3, 3, true)
  VA_CASE(64, 32, 3, 3, true) VA_CASE(64, 64, 3, 3, true)
#undef VA_CASE
  return errf("no kernel variant for BN=%d CK=%d R=%d S=%d wres=%d", BN, CK, R, S, (int)wres);
}

}  // namespace va
