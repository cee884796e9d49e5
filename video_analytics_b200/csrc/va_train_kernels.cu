// Training-step kernels that are not GEMM-shaped: max-pool forward, ReLU/pool
// backward, bias gradients, dropout, layout transposes for the weight-gradient GEMM, cross-entropy + the fp32 logit
// layer forward/backward, SGD with momentum.  Reference: Sheet03/spatialModel.py:161-181 (train(): forward in train
// mode, CrossEntropyLoss, backward, SGD(lr, momentum) step), :141-152 (Dropout after FC1..FC3).
#include "va_internal.h"

namespace va {

// ------------------------------------------------------------------------------------------------ max-pool forward
// NHWC bf16, 2x2 stride 2; 8 channels (16 B) per thread.
// codes (optional): one 4-bit code per (window, channel), 8 channels per uint32 -- the index 0..3 (scan order, first
// maximum) of the window element that receives the gradient, or 4 when the maximum is not positive (ReLU kills it).
// With the codes the backward pass needs neither a second look at the un-pooled activation nor the activation itself:
// 1/16 of its bytes are kept instead.
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y,
                                                          uint32_t* __restrict__ codes, int n, int H, int W, int C8) {
  const int Ho = H >> 1, Wo = W >> 1;
  const long long total = (long long)n * Ho * Wo * C8;
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total) return;
  const int c = (int)(g % C8);
  long long t = g / C8;
  const int wo = (int)(t % Wo); t /= Wo;
  const int ho = (int)(t % Ho);
  const int img = (int)(t / Ho);
  const uint4* base = x + (((long long)img * H + 2 * ho) * W + 2 * wo) * C8 + c;
  const uint4 a = __ldg(base), b = __ldg(base + C8), cc = __ldg(base + (long long)W * C8), d = __ldg(base + (long long)W * C8 + C8);
  uint4 o;
  auto mx = [](uint32_t p, uint32_t q) {
    __nv_bfloat162 m = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&p), *reinterpret_cast<__nv_bfloat162*>(&q));
    return *reinterpret_cast<uint32_t*>(&m);
  };
  o.x = mx(mx(a.x, b.x), mx(cc.x, d.x)); o.y = mx(mx(a.y, b.y), mx(cc.y, d.y));
  o.z = mx(mx(a.z, b.z), mx(cc.z, d.z)); o.w = mx(mx(a.w, b.w), mx(cc.w, d.w));
  y[g] = o;
  if (codes != nullptr) {
    const uint4 win[4] = {a, b, cc, d};
    uint32_t code = 0;
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) {
      float best = 0.f;
      int arg = 4;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t word = reinterpret_cast<const uint32_t*>(&win[k])[ch >> 1];
        const float v = __uint_as_float((ch & 1) ? (word & 0xFFFF0000u) : (word << 16));     // bf16 -> fp32 is a shift
        // first maximum in scan order, and only if positive: strict > against the running best (starts at 0)
        if (k == 0 ? v > 0.f : v > best) { if (v > 0.f) { best = v; arg = k; } }
      }
      code |= (uint32_t)arg << (4 * ch);
    }
    codes[g] = code;
  }
}

// Pool + ReLU backward from the codes of maxpool_fwd_kernel (+ bias gradient): reads dP and 4 bits per window-channel
// instead of the whole un-pooled activation.
__global__ void __launch_bounds__(256) pool_bwd_codes_bias_kernel(const uint4* __restrict__ dP, const uint32_t* __restrict__ codes,
                                                                  uint4* __restrict__ dZ, float* __restrict__ db, int n, int H,
                                                                  int W, int C8) {
  const int cg = threadIdx.x % C8, pl = threadIdx.x / C8, lanes = 256 / C8;
  const int Ho = H >> 1, Wo = W >> 1;
  const long long total = (long long)n * Ho * Wo;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (long long t = (long long)blockIdx.x * lanes + pl; t < total; t += (long long)gridDim.x * lanes) {
    const uint4 gp4 = __ldg(dP + t * C8 + cg);
    const uint32_t code = __ldg(codes + t * C8 + cg);
    const int wo = (int)(t % Wo);
    const long long t2 = t / Wo;
    const int ho = (int)(t2 % Ho), img = (int)(t2 / Ho);
    const long long i00 = (((long long)img * H + 2 * ho) * W + 2 * wo) * C8 + cg;
    const long long idx[4] = {i00, i00 + C8, i00 + (long long)W * C8, i00 + (long long)W * C8 + C8};
    const uint16_t* g16 = reinterpret_cast<const uint16_t*>(&gp4);
    uint32_t o[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int j = 0; j < 4; ++j) o[k][j] = 0u;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const uint32_t a = (code >> (4 * c)) & 15u;
      const uint32_t bits = (uint32_t)g16[c] << ((c & 1) * 16);
      if (a < 4) acc[c] += __uint_as_float((uint32_t)g16[c] << 16);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (a == (uint32_t)k) o[k][c >> 1] |= bits;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) dZ[idx[k]] = make_uint4(o[k][0], o[k][1], o[k][2], o[k][3]);
  }
  if (db == nullptr) return;
  __shared__ float red[256 * 8];
#pragma unroll
  for (int i = 0; i < 8; ++i) red[threadIdx.x * 8 + i] = acc[i];
  __syncthreads();
  if (pl == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float sum = 0.f;
      for (int l = 0; l < lanes; ++l) sum += red[(l * C8 + cg) * 8 + i];
      atomicAdd(db + cg * 8 + i, sum);
    }
  }
}

// ------------------------------------------------------------------------------------------------ ReLU (+pool) backward
// dZ = dY * (Y > 0) with dY = un-pooled dP: the gradient of a 2x2 window goes to its FIRST maximum in scan order
// (h-major), which is what torch's max_pool2d backward does.  One thread per window x 2 channels.
__global__ void __launch_bounds__(256) relu_pool_bwd_kernel(const __nv_bfloat162* __restrict__ dP,
                                                            const __nv_bfloat162* __restrict__ Y,
                                                            __nv_bfloat162* __restrict__ dZ, int n, int H, int W, int C2) {
  const int Ho = H >> 1, Wo = W >> 1;
  const long long total = (long long)n * Ho * Wo * C2;
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total) return;
  const int c = (int)(g % C2);
  long long t = g / C2;
  const int wo = (int)(t % Wo); t /= Wo;
  const int ho = (int)(t % Ho);
  const int img = (int)(t / Ho);
  const long long i00 = (((long long)img * H + 2 * ho) * W + 2 * wo) * C2 + c;
  const long long idx[4] = {i00, i00 + C2, i00 + (long long)W * C2, i00 + (long long)W * C2 + C2};
  const float2 gp = __bfloat1622float2(__ldg(dP + g));
  float2 v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] = __bfloat1622float2(__ldg(Y + idx[k]));
  int ax = 0, ay = 0;
#pragma unroll
  for (int k = 1; k < 4; ++k) {
    if (v[k].x > v[ax].x) ax = k;
    if (v[k].y > v[ay].y) ay = k;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float ox = (k == ax && v[k].x > 0.f) ? gp.x : 0.f;
    const float oy = (k == ay && v[k].y > 0.f) ? gp.y : 0.f;
    dZ[idx[k]] = __floats2bfloat162_rn(ox, oy);
  }
}
__global__ void __launch_bounds__(256) relu_bwd_kernel(const __nv_bfloat162* __restrict__ dY,
                                                       const __nv_bfloat162* __restrict__ Y,
                                                       __nv_bfloat162* __restrict__ dZ, long long total2) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total2) return;
  const float2 y = __bfloat1622float2(__ldg(Y + g)), d = __bfloat1622float2(__ldg(dY + g));
  dZ[g] = __floats2bfloat162_rn(y.x > 0.f ? d.x : 0.f, y.y > 0.f ? d.y : 0.f);
}

// Fused variant for the conv stack: 8 channels (16 B) per thread, a thread keeps ONE channel group and strides over
// windows / pixels, so the bias gradient db[c] = sum dZ[.., c] falls out of registers (block reduce + one atomic per
// channel per block) instead of a second pass over dZ.  Needs C % 8 == 0 and (C / 8) dividing the block size.
__device__ __forceinline__ void bf8_to_f(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
template <bool POOLED>
__global__ void __launch_bounds__(256) relu_pool_bwd_bias_kernel(const uint4* __restrict__ dP, const uint4* __restrict__ Y,
                                                                 uint4* __restrict__ dZ, float* __restrict__ db, int n, int H,
                                                                 int W, int C8) {
  const int cg = threadIdx.x % C8, pl = threadIdx.x / C8, lanes = 256 / C8;
  const int Ho = POOLED ? H >> 1 : H, Wo = POOLED ? W >> 1 : W;
  const long long total = (long long)n * Ho * Wo;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (long long t = (long long)blockIdx.x * lanes + pl; t < total; t += (long long)gridDim.x * lanes) {
    const uint4 gp4 = __ldg(dP + t * C8 + cg);
    if (POOLED) {
      const int wo = (int)(t % Wo);
      const long long t2 = t / Wo;
      const int ho = (int)(t2 % Ho), img = (int)(t2 / Ho);
      const long long i00 = (((long long)img * H + 2 * ho) * W + 2 * wo) * C8 + cg;
      const long long idx[4] = {i00, i00 + C8, i00 + (long long)W * C8, i00 + (long long)W * C8 + C8};
      float gp[8], v[4][8];
      bf8_to_f(gp4, gp);
#pragma unroll
      for (int k = 0; k < 4; ++k) bf8_to_f(__ldg(Y + idx[k]), v[k]);
      uint32_t o[4][4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) o[k][j] = 0u;
      const uint16_t* g16 = reinterpret_cast<const uint16_t*>(&gp4);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        int a = 0;
        float best = v[0][c];                 // running maximum kept in a register: no dynamically indexed array
#pragma unroll
        for (int k = 1; k < 4; ++k)
          if (v[k][c] > best) { best = v[k][c]; a = k; }
        if (best > 0.f) {
          acc[c] += gp[c];
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k == a) o[k][c >> 1] |= (uint32_t)g16[c] << ((c & 1) * 16);
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) dZ[idx[k]] = make_uint4(o[k][0], o[k][1], o[k][2], o[k][3]);
    } else {
      float y[8], g[8];
      bf8_to_f(__ldg(Y + t * C8 + cg), y);
      bf8_to_f(gp4, g);
      const uint16_t* g16 = reinterpret_cast<const uint16_t*>(&gp4);
      uint32_t o[4] = {0u, 0u, 0u, 0u};
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (y[c] > 0.f) { acc[c] += g[c]; o[c >> 1] |= (uint32_t)g16[c] << ((c & 1) * 16); }
      dZ[t * C8 + cg] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
  __shared__ float red[256 * 8];
#pragma unroll
  for (int i = 0; i < 8; ++i) red[threadIdx.x * 8 + i] = acc[i];
  __syncthreads();
  if (pl == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float sum = 0.f;
      for (int l = 0; l < lanes; ++l) sum += red[(l * C8 + cg) * 8 + i];
      atomicAdd(db + cg * 8 + i, sum);
    }
  }
}

// ------------------------------------------------------------------------------------------------ bias gradient
// db[c] = sum over rows of dZ[row][c]; block = 64 channels x 4 row lanes, grid.y slices the rows; fp32 atomics.
__global__ void __launch_bounds__(256) bias_grad_kernel(const __nv_bfloat16* __restrict__ dZ, float* __restrict__ db,
                                                        long long rows, int C) {
  const int c = blockIdx.x * 64 + (threadIdx.x & 63);
  const int lane_r = threadIdx.x >> 6;   // 0..3
  float acc = 0.f;
  if (c < C)
    for (long long r = (long long)blockIdx.y * 4 + lane_r; r < rows; r += (long long)gridDim.y * 4)
      acc += __bfloat162float(__ldg(dZ + r * C + c));
  __shared__ float red[256];
  red[threadIdx.x] = acc;
  __syncthreads();
  if (lane_r == 0 && c < C) atomicAdd(db + c, red[threadIdx.x] + red[threadIdx.x + 64] + red[threadIdx.x + 128] + red[threadIdx.x + 192]);
}

// ------------------------------------------------------------------------------------------------ dropout (fwd == bwd)
// y = x * scale where mask != 0, else 0 (nn.Dropout(p): scale = 1/(1-p)); masks are supplied by the caller so that
// the oracle can share them.
__global__ void __launch_bounds__(256) dropout_bf16_kernel(const __nv_bfloat16* __restrict__ x, const uint8_t* __restrict__ mask,
                                                           __nv_bfloat16* __restrict__ y, long long total, float scale) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total) return;
  y[g] = __float2bfloat16_rn(mask[g] ? __bfloat162float(x[g]) * scale : 0.f);
}
__global__ void __launch_bounds__(256) dropout_f32_kernel(const float* __restrict__ x, const uint8_t* __restrict__ mask,
                                                          float* __restrict__ y, long long total, float scale) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total) return;
  y[g] = mask[g] ? x[g] * scale : 0.f;
}

// ------------------------------------------------------------------------------------------------ layout transposes
// NHWC bf16 [n][HW][Cs] -> NCHW bf16 [nshift][n][C][H][Wp] (Wp >= W, zero padded) for the weight-gradient GEMM, whose K
// dimension (pixels along a row) must be contiguous.  32x32 smem tile transpose over (pixel-in-row, channel).
// Cs = channel stride of the source (>= C: the first layer's input carries zero-padded channels).
// nshift = 3 writes three copies shifted by s - 1 pixels along the row (copy s holds x[w + s - 1], zero outside the
// image): TMA cannot start a box at an address that is not 16-byte aligned, so the horizontal filter taps of the
// weight gradient cannot be reached by shifting the box by one element -- they read the pre-shifted copy instead.
__global__ void __launch_bounds__(256) nhwc_to_nchw_bf16_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                                int n, int H, int W, int Wp, int C, int Cs, int nshift) {
  __shared__ __nv_bfloat16 tile[34][33];
  const int img_h = blockIdx.z;            // img * H + h
  const int w0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  const int pad = nshift >> 1;
  for (int k = ty; k < 32 + 2 * pad; k += 8) {
    const int w = w0 - pad + k, c = c0 + tx;
    tile[k][tx] = (w >= 0 && w < W && c < C) ? x[((long long)img_h * W + w) * Cs + c] : __float2bfloat16_rn(0.f);
  }
  __syncthreads();
  const int img = img_h / H, h = img_h % H;
  for (int s = 0; s < nshift; ++s)
    for (int k = ty; k < 32; k += 8) {
      const int c = c0 + k, w = w0 + tx;
      if (c < C && w < Wp) y[((((long long)s * n + img) * C + c) * H + h) * Wp + w] = tile[tx + s][k];
    }
}

// ------------------------------------------------------------------------------------------------ casts
__global__ void __launch_bounds__(256) f32_to_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n) y[g] = __float2bfloat16_rn(x[g]);
}

// ------------------------------------------------------------------------------------------------ SGD with momentum
// torch.optim.SGD(lr, momentum, dampening=0, nesterov=False, weight_decay=0): buf = g (first step) or m*buf + g;
// p -= lr * buf   (reference spatialModel.py:116,181)
// GT = float (exact gradients) or __nv_bfloat16 (the all-reduced bf16 payload of the data-parallel step).
template <typename GT>
__device__ __forceinline__ float4 load_grad4(const GT* g, long long i4);
template <>
__device__ __forceinline__ float4 load_grad4<float>(const float* g, long long i4) { return *reinterpret_cast<const float4*>(g + i4); }
template <>
__device__ __forceinline__ float4 load_grad4<__nv_bfloat16>(const __nv_bfloat16* g, long long i4) {
  const uint2 r = *reinterpret_cast<const uint2*>(g + i4);
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&r.x), b = *reinterpret_cast<const __nv_bfloat162*>(&r.y);
  return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
}
__device__ __forceinline__ float grad1(const float* g, long long i) { return g[i]; }
__device__ __forceinline__ float grad1(const __nv_bfloat16* g, long long i) { return __bfloat162float(g[i]); }

template <typename GT>
__global__ void __launch_bounds__(256) sgd_momentum_kernel(float* __restrict__ p, const GT* __restrict__ g,
                                                           float* __restrict__ buf, long long n, float lr, float momentum,
                                                           int first_step, float grad_scale) {
  // four elements per thread (16-byte accesses: the arenas are 256-byte aligned); scalar tail
  const long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 + 3 < n) {
    const float4 gv = load_grad4<GT>(g, i4);
    float4 pv = *reinterpret_cast<float4*>(p + i4);
    float4 bv = first_step ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(buf + i4);
    const float ge[4] = {gv.x * grad_scale, gv.y * grad_scale, gv.z * grad_scale, gv.w * grad_scale};
    float be[4] = {bv.x, bv.y, bv.z, bv.w};
    float pe[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      be[k] = first_step ? ge[k] : fmaf(momentum, be[k], ge[k]);
      pe[k] = fmaf(-lr, be[k], pe[k]);
    }
    *reinterpret_cast<float4*>(buf + i4) = make_float4(be[0], be[1], be[2], be[3]);
    *reinterpret_cast<float4*>(p + i4) = make_float4(pe[0], pe[1], pe[2], pe[3]);
  } else {
    for (long long i = i4; i < n; ++i) {
      const float gi = grad1(g, i) * grad_scale;
      const float b = first_step ? gi : fmaf(momentum, buf[i], gi);
      buf[i] = b;
      p[i] = fmaf(-lr, b, p[i]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ logit layer + CE
// Forward: logits = x . W4^T + b4 (fp32), loss_i = logsumexp(logits_i) - logits_i[label_i]; backward (mean reduction):
// dlogits = (softmax - onehot) / n.  One CTA per sample.  Then dW4 = dlogits^T . x, db4 = sum dlogits,
// dx = dlogits . W4 by two small kernels.  Labels are class indices as fed by the reference (1-based ints used as-is).
__global__ void __launch_bounds__(128) ce_fwd_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w4,
                                                         const float* __restrict__ b4, const int64_t* __restrict__ labels,
                                                         int n, int D, int C, float* __restrict__ logits,
                                                         float* __restrict__ dlogits, float* __restrict__ loss_sum) {
  extern __shared__ float sm[];
  float* sx = sm;
  float* sl = sm + D;
  __shared__ float red[4];
  const int i = blockIdx.x;
  for (int d = threadIdx.x; d < D; d += blockDim.x) sx[d] = x[(size_t)i * D + d];
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = b4[c];
    const float* wr = w4 + (size_t)c * D;
    for (int d = 0; d < D; ++d) acc = fmaf(sx[d], __ldg(wr + d), acc);
    sl[c] = acc;
    if (logits) logits[(size_t)i * C + c] = acc;
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < C; c += blockDim.x) mx = fmaxf(mx, sl[c]);
  for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  float s = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) s += expf(sl[c] - mx);
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  s = red[0] + red[1] + red[2] + red[3];
  // A label outside [0, C) (torch's CrossEntropyLoss raises "Target out of bounds"; the reference's 1-based class ids
  // reach C when all 101 classes are in use) must not index shared memory: the loss becomes NaN -- loud, and it
  // poisons nothing else -- and the sample contributes no gradient.  The host layer validates labels before upload.
  const long long lab64 = labels[i];
  const bool lab_ok = lab64 >= 0 && lab64 < (long long)C;
  const int lab = lab_ok ? (int)lab64 : 0;
  const float inv_n = 1.0f / (float)n;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float pr = expf(sl[c] - mx) / s;
    dlogits[(size_t)i * C + c] = lab_ok ? (pr - (c == lab ? 1.f : 0.f)) * inv_n : 0.f;
  }
  if (threadIdx.x == 0) atomicAdd(loss_sum, lab_ok ? (logf(s) + mx - sl[lab]) * inv_n : __int_as_float(0x7fc00000));
}
// dW4[c][d] = sum_i dlogits[i][c] * x[i][d]; db4[c] = sum_i dlogits[i][c]   (grid = C, block = D threads)
__global__ void logit_wgrad_kernel(const float* __restrict__ dlogits, const float* __restrict__ x, int n, int D, int C,
                                   float* __restrict__ dw4, float* __restrict__ db4) {
  const int c = blockIdx.x;
  float bsum = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float acc = 0.f;
    for (int i = 0; i < n; ++i) acc = fmaf(dlogits[(size_t)i * C + c], x[(size_t)i * D + d], acc);
    dw4[(size_t)c * D + d] = acc;
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < n; ++i) bsum += dlogits[(size_t)i * C + c];
    db4[c] = bsum;
  }
}
// dx[i][d] = sum_c dlogits[i][c] * W4[c][d]   (grid = n, block = D threads)
__global__ void logit_dgrad_kernel(const float* __restrict__ dlogits, const float* __restrict__ w4, int D, int C,
                                   float* __restrict__ dx) {
  const int i = blockIdx.x;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float acc = 0.f;
    for (int c = 0; c < C; ++c) acc = fmaf(dlogits[(size_t)i * C + c], __ldg(w4 + (size_t)c * D + d), acc);
    dx[(size_t)i * D + d] = acc;
  }
}
// fp32 ReLU backward + cast: dz = (y > 0) ? dy : 0 -> bf16 (descriptor layer, whose forward output is fp32)
__global__ void __launch_bounds__(256) relu_bwd_f32_to_bf16_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                                   __nv_bfloat16* __restrict__ dz, long long n) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n) dz[g] = __float2bfloat16_rn(y[g] > 0.f ? dy[g] : 0.f);
}

// ------------------------------------------------------------------------------------------------ dgrad weight views
// Data gradient of a 3x3/s1/p1 convolution = the same convolution of dZ with the filter rotated by 180 degrees and its
// channel roles swapped: W'[ci][co][r][s] = W[co][ci][ks-1-r][ks-1-s]  (fp32 OIHW in, fp32 OIHW' out).
__global__ void flip_transpose_conv_w_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin, int ks) {
  const size_t total = (size_t)Cout * Cin * ks * ks;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
    const int s = (int)(g % ks), r = (int)((g / ks) % ks);
    const int co = (int)((g / (ks * ks)) % Cout);
    const int ci = (int)(g / ((size_t)ks * ks * Cout));
    out[g] = w[(((size_t)co * Cin + ci) * ks + (ks - 1 - r)) * ks + (ks - 1 - s)];
  }
}
// [out][in] fp32 -> bf16 [in][out]  (weights of the FC data-gradient GEMM dX = dY . W)
__global__ void pack_fc_w_t_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int n_out, int n_in) {
  __shared__ float tile[32][33];
  const int i0 = blockIdx.x * 32, o0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int k = ty; k < 32; k += 8) {
    const int o = o0 + k, i = i0 + tx;
    tile[k][tx] = (o < n_out && i < n_in) ? w[(size_t)o * n_in + i] : 0.f;
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    const int i = i0 + k, o = o0 + tx;
    if (i < n_in && o < n_out) out[(size_t)i * n_out + o] = __float2bfloat16_rn(tile[tx][k]);
  }
}

// ------------------------------------------------------------------------------------------------ launchers
static inline unsigned nblk(long long total) { return (unsigned)((total + 255) / 256); }

cudaError_t launch_maxpool_fwd(const void* x, void* y, void* codes, int n, int H, int W, int C, cudaStream_t st) {
  const long long total = (long long)n * (H / 2) * (W / 2) * (C / 8);
  if (total == 0) return cudaSuccess;
  count_launch();
  maxpool_fwd_kernel<<<nblk(total), 256, 0, st>>>(static_cast<const uint4*>(x), static_cast<uint4*>(y),
                                                  static_cast<uint32_t*>(codes), n, H, W, C / 8);
  return cudaGetLastError();
}
cudaError_t launch_pool_bwd_codes(const void* dP, const void* codes, void* dZ, float* db, int n, int H, int W, int C, cudaStream_t st) {
  const long long units = (long long)n * (H / 2) * (W / 2);
  if (units == 0) return cudaSuccess;
  if (C % 8 != 0 || 256 % (C / 8) != 0) return cudaErrorInvalidValue;
  if (db != nullptr) {
    cudaError_t e = cudaMemsetAsync(db, 0, (size_t)C * 4, st);
    if (e != cudaSuccess) return e;
  }
  const int C8 = C / 8, lanes = 256 / C8;
  const long long want = (units + lanes * 16 - 1) / (lanes * 16);
  const unsigned blocks = (unsigned)std::max<long long>(1, std::min<long long>(want, 148 * 8));
  count_launch();
  pool_bwd_codes_bias_kernel<<<blocks, 256, 0, st>>>(static_cast<const uint4*>(dP), static_cast<const uint32_t*>(codes),
                                                     static_cast<uint4*>(dZ), db, n, H, W, C8);
  return cudaGetLastError();
}
cudaError_t launch_relu_pool_bwd(const void* dP, const void* Y, void* dZ, float* db, int n, int H, int W, int C, int pooled,
                                 cudaStream_t st) {
  if (db != nullptr) {
    const long long rows = (long long)n * H * W;
    if (C % 8 == 0 && 256 % (C / 8) == 0 && rows > 0) {      // fused ReLU/pool backward + bias gradient
      cudaError_t e = cudaMemsetAsync(db, 0, (size_t)C * 4, st);
      if (e != cudaSuccess) return e;
      const int C8 = C / 8, lanes = 256 / C8;
      const long long units = pooled ? rows / 4 : rows;
      // >= 16 windows per thread: every block ends with C same-address atomics, which dominate on the small layers
      const long long want = (units + lanes * 16 - 1) / (lanes * 16);
      const unsigned blocks = (unsigned)std::max<long long>(1, std::min<long long>(want, 148 * 8));
      count_launch();
      if (pooled)
        relu_pool_bwd_bias_kernel<true><<<blocks, 256, 0, st>>>(static_cast<const uint4*>(dP), static_cast<const uint4*>(Y),
                                                                 static_cast<uint4*>(dZ), db, n, H, W, C8);
      else
        relu_pool_bwd_bias_kernel<false><<<blocks, 256, 0, st>>>(static_cast<const uint4*>(dP), static_cast<const uint4*>(Y),
                                                                  static_cast<uint4*>(dZ), db, n, H, W, C8);
      return cudaGetLastError();
    }
    cudaError_t e = launch_relu_pool_bwd(dP, Y, dZ, nullptr, n, H, W, C, pooled, st);
    if (e != cudaSuccess) return e;
    return launch_bias_grad(dZ, db, rows, C, st);
  }
  count_launch();
  if (pooled) {
    const long long total = (long long)n * (H / 2) * (W / 2) * (C / 2);
    if (total == 0) return cudaSuccess;
    relu_pool_bwd_kernel<<<nblk(total), 256, 0, st>>>(static_cast<const __nv_bfloat162*>(dP), static_cast<const __nv_bfloat162*>(Y),
                                                       static_cast<__nv_bfloat162*>(dZ), n, H, W, C / 2);
  } else {
    const long long total2 = (long long)n * H * W * C / 2;
    if (total2 == 0) return cudaSuccess;
    relu_bwd_kernel<<<nblk(total2), 256, 0, st>>>(static_cast<const __nv_bfloat162*>(dP), static_cast<const __nv_bfloat162*>(Y),
                                                   static_cast<__nv_bfloat162*>(dZ), total2);
  }
  return cudaGetLastError();
}
cudaError_t launch_bias_grad(const void* dZ, float* db, long long rows, int C, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(db, 0, (size_t)C * 4, st);
  if (e != cudaSuccess || rows == 0) return e;
  const unsigned gy = (unsigned)std::min<long long>((rows + 3) / 4, 148 * 8 / ((C + 63) / 64) + 1);
  count_launch();
  bias_grad_kernel<<<dim3((C + 63) / 64, gy), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(dZ), db, rows, C);
  return cudaGetLastError();
}
cudaError_t launch_dropout(const void* x, const uint8_t* mask, void* y, long long total, float scale, int is_f32, cudaStream_t st) {
  if (total == 0) return cudaSuccess;
  count_launch();
  if (is_f32) dropout_f32_kernel<<<nblk(total), 256, 0, st>>>(static_cast<const float*>(x), mask, static_cast<float*>(y), total, scale);
  else dropout_bf16_kernel<<<nblk(total), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), mask, static_cast<__nv_bfloat16*>(y), total, scale);
  return cudaGetLastError();
}
cudaError_t launch_nhwc_to_nchw_bf16(const void* x, void* y, int n, int H, int W, int Wp, int C, int Cs, int nshift,
                                     cudaStream_t st) {
  if ((long long)n * H * W * C == 0) return cudaSuccess;
  if ((long long)n * H > 65535) return cudaErrorInvalidValue;
  count_launch();
  nhwc_to_nchw_bf16_kernel<<<dim3((Wp + 31) / 32, (C + 31) / 32, n * H), 256, 0, st>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), n, H, W, Wp, C, Cs, nshift);
  return cudaGetLastError();
}
cudaError_t launch_f32_to_bf16(const float* x, void* y, long long n, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  count_launch();
  f32_to_bf16_kernel<<<nblk(n), 256, 0, st>>>(x, static_cast<__nv_bfloat16*>(y), n);
  return cudaGetLastError();
}
cudaError_t launch_sgd_momentum(float* p, const float* g, float* buf, long long n, float lr, float momentum, int first_step,
                                float grad_scale, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  count_launch();
  sgd_momentum_kernel<float><<<nblk((n + 3) / 4), 256, 0, st>>>(p, g, buf, n, lr, momentum, first_step, grad_scale);
  return cudaGetLastError();
}
cudaError_t launch_sgd_momentum_bf16g(float* p, const void* g_bf16, float* buf, long long n, float lr, float momentum,
                                      int first_step, float grad_scale, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  count_launch();
  sgd_momentum_kernel<__nv_bfloat16><<<nblk((n + 3) / 4), 256, 0, st>>>(p, static_cast<const __nv_bfloat16*>(g_bf16), buf, n, lr,
                                                                        momentum, first_step, grad_scale);
  return cudaGetLastError();
}
cudaError_t launch_ce_train(const float* x, const float* w4, const float* b4, const int64_t* labels, int n, int D, int C,
                            float* logits, float* dlogits, float* loss, float* dw4, float* db4, float* dx, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(loss, 0, 4, st);
  if (e != cudaSuccess) return e;
  count_launch();
  ce_fwd_bwd_kernel<<<n, 128, (D + C) * sizeof(float), st>>>(x, w4, b4, labels, n, D, C, logits, dlogits, loss);
  count_launch();
  logit_wgrad_kernel<<<C, 256, 0, st>>>(dlogits, x, n, D, C, dw4, db4);
  count_launch();
  logit_dgrad_kernel<<<n, 256, 0, st>>>(dlogits, w4, D, C, dx);
  return cudaGetLastError();
}
cudaError_t launch_flip_transpose_conv_w(const float* w, float* out, int Cout, int Cin, int ks, cudaStream_t st) {
  const size_t total = (size_t)Cout * Cin * ks * ks;
  unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, 148 * 8);
  count_launch();
  flip_transpose_conv_w_kernel<<<blocks, 256, 0, st>>>(w, out, Cout, Cin, ks);
  return cudaGetLastError();
}
cudaError_t launch_pack_fc_w_t(const float* w, void* out, int n_out, int n_in, cudaStream_t st) {
  count_launch();
  pack_fc_w_t_kernel<<<dim3((n_in + 31) / 32, (n_out + 31) / 32), 256, 0, st>>>(w, static_cast<__nv_bfloat16*>(out), n_out, n_in);
  return cudaGetLastError();
}
cudaError_t launch_relu_bwd_f32_to_bf16(const float* dy, const float* y, void* dz, long long n, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  count_launch();
  relu_bwd_f32_to_bf16_kernel<<<nblk(n), 256, 0, st>>>(dy, y, static_cast<__nv_bfloat16*>(dz), n);
  return cudaGetLastError();
}

}  // namespace va
