// C-ABI of libva_b200.so (see include/va_b200.h).  Owns the network handle: packed weights, activation
// workspace, and the layer schedule of the VGG16-D two-stream networks.
#include "../../include/va_b200.h"
#include "va_internal.h"

#include <atomic>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <new>

#include <vector>

namespace va {
static std::atomic<uint64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace va

// Optional live profile of the tensor-core layer kernel: event pairs bracket the conv/FC launches of every
// va_forward chunk on the launching stream; bench.py reads them after the timed region (roofline.achieved).
namespace {
struct ProfSpan { cudaEvent_t a, b; };
bool g_prof_on = false;
std::vector<ProfSpan> g_prof_spans;
uint64_t g_prof_launches = 0;
double g_prof_flops = 0.0;
}  // namespace

namespace {

thread_local char g_last_error[768] = "";

va_status fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
  return code;
}
#define VA_CUDA(expr)                                                                              \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) return fail(VA_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e));      \
  } while (0)

// Stream-ordered scratch that is released on every exit path (early error returns included).
struct AsyncScratch {
  cudaStream_t st;
  void* ptrs[4];
  int n;
  explicit AsyncScratch(cudaStream_t s) : st(s), ptrs{nullptr, nullptr, nullptr, nullptr}, n(0) {}
  cudaError_t alloc(void** out, size_t bytes) {
    cudaError_t e = cudaMallocAsync(out, bytes, st);
    if (e == cudaSuccess && n < 4) ptrs[n++] = *out;
    return e;
  }
  ~AsyncScratch() { for (int i = 0; i < n; ++i) cudaFreeAsync(ptrs[i], st); }
};

// VGG16 configuration "D" (torchvision models.vgg16; reference spatialModel.py:110): output channels per conv,
// `true` = followed by MaxPool2d(2,2).
struct ConvSpec { int cout; bool pool; };
const ConvSpec kVgg16[13] = {{64, false}, {64, true},  {128, false}, {128, true}, {256, false}, {256, false}, {256, true},
                             {512, false}, {512, false}, {512, true}, {512, false}, {512, false}, {512, true}};
constexpr int kCrop = 224;       // parameters.py:10 CROP_SIZE_TF
constexpr int kFc1In = 512 * 7 * 7;
constexpr int kFcHidden = 4096;

}  // namespace

struct va_handle {
  int stream_kind, cin, cin_pad, n_classes, desc_dim, max_batch;
  int precision;    // 0 = bf16 storage; 1 = fp32-accuracy (bf16x3 slices, six cross terms along K)
  bool loaded;
  void* wconv[13];
  void* wconv1_fused;   // conv1_1 in the dense-K layout of the fused gather kernel (precision 0 only)
  float* bconv[13];
  void* wfc[3];     // FC1, FC2, FC3 packed bf16
  float* bfc[3];
  float* w4t;       // FC4 fp32, transposed [desc_dim][n_classes]
  float* b4;
  void* act[2];     // ping-pong activation workspace
  size_t act_bytes;
  float* desc_ws;   // [max_batch][desc_dim] when the caller does not want descriptors
  float* splitk_ws; // [kSplitK][max_batch][4096] fp32 partial sums of the split-K fully-connected layers
};

extern "C" {

const char* va_last_error(void) { return g_last_error; }
int va_abi_version(void) { return 1; }
uint64_t va_launch_count(void) { return va::g_launches.load(); }

va_status va_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  VA_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  VA_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return VA_OK;
}

static va_status require_sm100() {
  int dev = 0;
  VA_CUDA(cudaGetDevice(&dev));
  int major = 0;
  VA_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) return fail(VA_ERR_UNSUPPORTED, "libva_b200 is built for sm_100a only; device has cc %d.x", major);
  // Stream-ordered scratch (cudaMallocAsync in the training primitives) must stay cached across synchronisation
  // points: the default pool releases everything on every sync, which turns each step into GBs of re-allocation.
  // Bounded: the pool may keep up to 2 GiB (the largest scratch set of a batch-256 step is ~1 GiB: FC1's bf16 repack
  // + weight-gradient planes); anything above goes back to the driver, so torch's allocator is not starved.
  static bool pool_set[64] = {false};
  if (dev < 64 && !pool_set[dev]) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t keep = 2ull << 30;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    pool_set[dev] = true;
  }
  return VA_OK;
}

va_status va_create(va_handle** out, int stream_kind, int in_channels, int n_classes, int desc_dim, int max_batch) {
  return va_create_ex(out, stream_kind, in_channels, n_classes, desc_dim, max_batch, 0);
}

va_status va_create_ex(va_handle** out, int stream_kind, int in_channels, int n_classes, int desc_dim, int max_batch,
                       int precision) {
  if (!out) return fail(VA_ERR_INVALID, "va_create: out is NULL");
  if (precision != 0 && precision != 1) return fail(VA_ERR_INVALID, "precision %d (0 = bf16, 1 = fp32x3)", precision);
  *out = nullptr;
  if (va_status s = require_sm100()) return s;
  if (in_channels < 1 || in_channels > 32) return fail(VA_ERR_INVALID, "in_channels %d not in [1,32]", in_channels);
  if (desc_dim % 64 != 0 || desc_dim <= 0) return fail(VA_ERR_INVALID, "desc_dim %d must be a multiple of 64", desc_dim);
  if (n_classes < 1 || n_classes > 1024) return fail(VA_ERR_INVALID, "n_classes %d", n_classes);
  if (max_batch < 1) return fail(VA_ERR_INVALID, "max_batch %d", max_batch);
  va_handle* h = new (std::nothrow) va_handle();
  if (!h) return fail(VA_ERR_INVALID, "out of host memory");
  memset(h, 0, sizeof(*h));
  h->stream_kind = stream_kind; h->cin = in_channels; h->cin_pad = in_channels <= 16 ? 16 : 32;
  h->precision = precision;
  const int kmul = precision ? 6 : 1;    // K multiplier of every layer but the first
  if (precision) {
    if (6 * in_channels <= 32) h->cin_pad = 32;
    else if (6 * in_channels <= 128) h->cin_pad = 128;
    else { delete h; return fail(VA_ERR_INVALID, "fp32x3 mode supports up to 21 input channels"); }
  }
  h->n_classes = n_classes; h->desc_dim = desc_dim; h->max_batch = max_batch;
  int cin = h->cin_pad;
  for (int i = 0; i < 13; ++i) {
    const size_t wbytes = (size_t)9 * kVgg16[i].cout * cin * 2;   // cin already carries the K multiplier
    if (cudaMalloc(&h->wconv[i], wbytes) != cudaSuccess || cudaMalloc(&h->bconv[i], kVgg16[i].cout * 4) != cudaSuccess) {
      va_destroy(h);
      return fail(VA_ERR_CUDA, "cudaMalloc conv weights failed");
    }
    cin = kVgg16[i].cout * kmul;
  }
  if (!precision && cudaMalloc(&h->wconv1_fused, (size_t)va::conv1_fused_packed_bytes(in_channels)) != cudaSuccess) {
    va_destroy(h);
    return fail(VA_ERR_CUDA, "cudaMalloc conv1 fused weights failed");
  }
  const int fin[3] = {kFc1In, kFcHidden, kFcHidden};
  const int fout[3] = {kFcHidden, kFcHidden, desc_dim};
  for (int i = 0; i < 3; ++i) {
    if (cudaMalloc(&h->wfc[i], (size_t)fin[i] * kmul * fout[i] * 2) != cudaSuccess ||
        cudaMalloc(&h->bfc[i], fout[i] * 4) != cudaSuccess) {
      va_destroy(h);
      return fail(VA_ERR_CUDA, "cudaMalloc fc weights failed");
    }
  }
  h->act_bytes = (size_t)max_batch * kCrop * kCrop * 64 * 2 * kmul;
  if (cudaMalloc(&h->w4t, (size_t)desc_dim * n_classes * 4) != cudaSuccess ||
      cudaMalloc(&h->b4, n_classes * 4) != cudaSuccess || cudaMalloc(&h->act[0], h->act_bytes) != cudaSuccess ||
      cudaMalloc(&h->act[1], h->act_bytes) != cudaSuccess ||
      cudaMalloc(&h->desc_ws, (size_t)max_batch * desc_dim * 4) != cudaSuccess ||
      (!precision && cudaMalloc(&h->splitk_ws, (size_t)va::kSplitK * max_batch * kFcHidden * 4) != cudaSuccess)) {
    va_destroy(h);
    return fail(VA_ERR_CUDA, "cudaMalloc workspace failed (max_batch %d needs 2 x %zu bytes)", max_batch, h->act_bytes);
  }
  *out = h;
  return VA_OK;
}

va_status va_destroy(va_handle* h) {
  if (!h) return VA_OK;
  for (int i = 0; i < 13; ++i) { cudaFree(h->wconv[i]); cudaFree(h->bconv[i]); }
  cudaFree(h->wconv1_fused);
  for (int i = 0; i < 3; ++i) { cudaFree(h->wfc[i]); cudaFree(h->bfc[i]); }
  cudaFree(h->w4t); cudaFree(h->b4); cudaFree(h->act[0]); cudaFree(h->act[1]); cudaFree(h->desc_ws); cudaFree(h->splitk_ws);
  delete h;
  return VA_OK;
}

int va_input_channels_padded(const va_handle* h) { return h ? h->cin_pad : 0; }

va_status va_load_weights(va_handle* h, const void* const* tensors, int n_tensors, va_stream_t stream) {
  if (!h || !tensors) return fail(VA_ERR_INVALID, "va_load_weights: NULL argument");
  if (n_tensors != 34) return fail(VA_ERR_INVALID, "va_load_weights: expected 34 state_dict tensors, got %d", n_tensors);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int cin = h->cin, cin_pad = h->cin_pad;
  for (int i = 0; i < 13; ++i) {
    const int cout = kVgg16[i].cout;
    if (h->precision)
      VA_CUDA(va::launch_pack_conv_w_split6(static_cast<const float*>(tensors[2 * i]), h->wconv[i], cout, cin, cin_pad, 3, st));
    else
      VA_CUDA(va::launch_pack_conv_w(static_cast<const float*>(tensors[2 * i]), h->wconv[i], cout, cin, cin_pad, 3, st));
    VA_CUDA(cudaMemcpyAsync(h->bconv[i], tensors[2 * i + 1], cout * 4, cudaMemcpyDeviceToDevice, st));
    if (i == 0 && h->wconv1_fused)
      VA_CUDA(va::launch_pack_conv1_fused_w(static_cast<const float*>(tensors[0]), h->wconv1_fused, cin, st));
    cin = cout; cin_pad = cout * (h->precision ? 6 : 1);
  }
  const int fin[3] = {kFc1In, kFcHidden, kFcHidden};
  const int fout[3] = {kFcHidden, kFcHidden, h->desc_dim};
  for (int i = 0; i < 3; ++i) {
    // FC1 consumes our NHWC flatten of the [7][7][512] feature map
    if (h->precision)
      VA_CUDA(va::launch_pack_fc_w_split6(static_cast<const float*>(tensors[26 + 2 * i]), h->wfc[i], fout[i],
                                          i == 0 ? 512 : fin[i], i == 0 ? 49 : 1, i == 0 ? 1 : 0, st));
    else
      VA_CUDA(va::launch_pack_fc_w(static_cast<const float*>(tensors[26 + 2 * i]), h->wfc[i], fout[i], fin[i],
                                   i == 0 ? 512 : 0, 49, st));
    VA_CUDA(cudaMemcpyAsync(h->bfc[i], tensors[26 + 2 * i + 1], fout[i] * 4, cudaMemcpyDeviceToDevice, st));
  }
  VA_CUDA(va::launch_transpose_f32(static_cast<const float*>(tensors[32]), h->w4t, h->n_classes, h->desc_dim, st));
  VA_CUDA(cudaMemcpyAsync(h->b4, tensors[33], h->n_classes * 4, cudaMemcpyDeviceToDevice, st));
  h->loaded = true;
  return VA_OK;
}

va_status va_preprocess(const uint8_t* images, size_t image_bytes, int img_h, int img_w, int img_c,
                        const int32_t* index_table, int n, int planes, int crop, const float* mean, const float* std,
                        int c_pad, int out_mode, void* out, va_stream_t stream) {
  if (n == 0) return VA_OK;
  if (!images || !index_table || !mean || !std || !out) return fail(VA_ERR_INVALID, "va_preprocess: NULL argument");
  if (n < 0 || planes < 1 || img_c < 1 || planes * img_c > 32) return fail(VA_ERR_INVALID, "va_preprocess: bad shape");
  if (crop > img_h || crop > img_w || crop < 1) return fail(VA_ERR_INVALID, "va_preprocess: crop %d vs image %dx%d", crop, img_h, img_w);
  if (out_mode == 0 && !(c_pad == 16 || c_pad == 32 || c_pad == 64)) return fail(VA_ERR_INVALID, "va_preprocess: c_pad %d", c_pad);
  if (va_status s = require_sm100()) return s;
  VA_CUDA(va::launch_preprocess(images, image_bytes, img_h, img_w, img_c, index_table, n, planes, crop, mean, std, c_pad,
                                out_mode, out, static_cast<cudaStream_t>(stream)));
  return VA_OK;
}

namespace {
struct StoreInput {          // front end of va_forward_store
  const uint8_t* images; size_t image_bytes; int img_h, img_w, img_c;
  const int32_t* table; int planes; const float* mean; const float* stdv;
};

// Layers of one stream over n snippets in chunks of max_batch.  `in_nhwc` (preprocessed input) or `src` (image store +
// index table: the first layer is the fused gather kernel) -- exactly one of the two.
va_status forward_impl(va_handle* h, const void* in_nhwc, const StoreInput* src, int n, float* descriptors, float* logits,
                       float* probs, int32_t* pred, cudaStream_t st) {
  const size_t in_stride = (size_t)kCrop * kCrop * h->cin_pad * 2;
  for (int off = 0; off < n; off += h->max_batch) {
    const int nb = (n - off) < h->max_batch ? (n - off) : h->max_batch;
    const void* x = in_nhwc ? static_cast<const uint8_t*>(in_nhwc) + (size_t)off * in_stride : nullptr;
    int H = kCrop, cin_pad = h->cin_pad, cur = 0;
    ProfSpan span{nullptr, nullptr};
    if (g_prof_on) {
      VA_CUDA(cudaEventCreate(&span.a));
      VA_CUDA(cudaEventCreate(&span.b));
      VA_CUDA(cudaEventRecord(span.a, st));
    }
    int cin_real = h->cin;
    for (int i = 0; i < 13; ++i) {
      const int cout = kVgg16[i].cout;
      if (i == 0 && src) {
        const int32_t* tab = src->table + (size_t)off * src->planes * 4;
        if (const char* e = va::conv1_fused_run(src->images, src->image_bytes, src->img_h, src->img_w, src->img_c, tab, nb,
                                                src->planes, src->mean, src->stdv, h->wconv1_fused, h->bconv[0], h->act[cur], st))
          return fail(VA_ERR_CUDA, "fused conv layer 0: %s", e);
      } else {
        va::ConvLayerDesc d;
        d.x = x; d.n = nb; d.H = H; d.W = H; d.cin_pad = cin_pad;
        d.w_packed = h->wconv[i]; d.bias = h->bconv[i]; d.Cout = cout; d.ks = 3;
        d.relu = 1; d.pool = kVgg16[i].pool ? 1 : 0; d.y = h->act[cur]; d.y_f32 = nullptr; d.force_bn = 0; d.force_r = 0;
        d.split6 = h->precision;
        if (const char* e = va::conv_layer_run(d, st)) return fail(VA_ERR_CUDA, "conv layer %d: %s", i, e);
      }
      x = h->act[cur]; cur ^= 1;
      if (g_prof_on) { g_prof_flops += 2.0 * nb * H * H * (double)cout * 9.0 * cin_real; ++g_prof_launches; }
      cin_pad = cout * (h->precision ? 6 : 1); cin_real = cout;
      if (kVgg16[i].pool) H >>= 1;
    }
    float* desc_out = descriptors ? descriptors + (size_t)off * h->desc_dim : h->desc_ws;
    const int fin[3] = {kFc1In, kFcHidden, kFcHidden};
    const int fout[3] = {kFcHidden, kFcHidden, h->desc_dim};
    for (int i = 0; i < 3; ++i) {
      va::ConvLayerDesc d;
      d.x = x; d.n = nb; d.H = 1; d.W = 1; d.cin_pad = fin[i] * (h->precision ? 6 : 1);
      d.split6 = h->precision;
      d.w_packed = h->wfc[i]; d.bias = h->bfc[i]; d.Cout = fout[i]; d.ks = 1;
      d.relu = 1; d.pool = 0; d.force_bn = 0; d.force_r = 0;
      d.splitk_ws = h->splitk_ws;
      d.y = (i < 2) ? h->act[cur] : nullptr;
      d.y_f32 = (i < 2) ? nullptr : desc_out;
      if (const char* e = va::conv_layer_run(d, st)) return fail(VA_ERR_CUDA, "fc layer %d: %s", i + 1, e);
      x = h->act[cur]; cur ^= 1;
      if (g_prof_on) { g_prof_flops += 2.0 * nb * (double)fin[i] * fout[i]; ++g_prof_launches; }
    }
    if (g_prof_on) {
      VA_CUDA(cudaEventRecord(span.b, st));
      g_prof_spans.push_back(span);
    }
    VA_CUDA(va::launch_head(desc_out, h->w4t, h->b4, nb, h->desc_dim, h->n_classes,
                            logits ? logits + (size_t)off * h->n_classes : nullptr,
                            probs ? probs + (size_t)off * h->n_classes : nullptr, pred ? pred + off : nullptr, st));
  }
  return VA_OK;
}
}  // namespace

va_status va_forward(va_handle* h, const void* in_nhwc, int n, float* descriptors, float* logits, float* probs,
                     int32_t* pred, va_stream_t stream) {
  if (!h || !in_nhwc) return fail(VA_ERR_INVALID, "va_forward: NULL argument");
  if (!h->loaded) return fail(VA_ERR_INVALID, "va_forward: weights not loaded");
  return forward_impl(h, in_nhwc, nullptr, n, descriptors, logits, probs, pred, static_cast<cudaStream_t>(stream));
}

va_status va_forward_store(va_handle* h, const uint8_t* images, size_t image_bytes, int img_h, int img_w, int img_c,
                           const int32_t* index_table, int n, int planes, const float* mean, const float* std,
                           float* descriptors, float* logits, float* probs, int32_t* pred, va_stream_t stream) {
  if (!h || !images || !index_table || !mean || !std) return fail(VA_ERR_INVALID, "va_forward_store: NULL argument");
  if (!h->loaded) return fail(VA_ERR_INVALID, "va_forward_store: weights not loaded");
  if (h->precision || !h->wconv1_fused)
    return fail(VA_ERR_UNSUPPORTED, "va_forward_store: precision-1 handles use va_preprocess + va_forward");
  if (planes < 1 || img_c < 1 || planes * img_c != h->cin)
    return fail(VA_ERR_INVALID, "va_forward_store: planes %d x img_c %d != in_channels %d", planes, img_c, h->cin);
  if (!va::conv1_fused_supported(planes, img_c, kCrop))
    return fail(VA_ERR_UNSUPPORTED, "va_forward_store: (planes, img_c) = (%d, %d); supported (1, 3) and (1..23, 1)", planes, img_c);
  if (img_h < kCrop || img_w < kCrop || image_bytes < (size_t)img_h * img_w * img_c)
    return fail(VA_ERR_INVALID, "va_forward_store: image %dx%dx%d (%zu bytes) cannot hold a %d crop", img_h, img_w, img_c, image_bytes, kCrop);
  const StoreInput src{images, image_bytes, img_h, img_w, img_c, index_table, planes, mean, std};
  return forward_impl(h, nullptr, &src, n, descriptors, logits, probs, pred, static_cast<cudaStream_t>(stream));
}

va_status va_conv1_fused(const uint8_t* images, size_t image_bytes, int img_h, int img_w, int img_c,
                         const int32_t* index_table, int n, int planes, const float* mean, const float* std,
                         const float* w, const float* bias, void* y, va_stream_t stream) {
  if (!images || !index_table || !mean || !std || !w || !bias || !y) return fail(VA_ERR_INVALID, "va_conv1_fused: NULL argument");
  if (n < 0 || planes < 1 || img_c < 1 || planes * img_c > 32) return fail(VA_ERR_INVALID, "va_conv1_fused: bad shape");
  if (!va::conv1_fused_supported(planes, img_c, kCrop))
    return fail(VA_ERR_UNSUPPORTED, "va_conv1_fused: (planes, img_c) = (%d, %d); supported (1, 3) and (1..23, 1)", planes, img_c);
  if (img_h < kCrop || img_w < kCrop) return fail(VA_ERR_INVALID, "va_conv1_fused: image smaller than the crop");
  if (va_status s = require_sm100()) return s;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  void* wp = nullptr;
  AsyncScratch scratch(st);
  VA_CUDA(scratch.alloc(&wp, (size_t)va::conv1_fused_packed_bytes(planes * img_c)));
  cudaError_t pe = va::launch_pack_conv1_fused_w(w, wp, planes * img_c, st);
  const char* e = pe == cudaSuccess ? va::conv1_fused_run(images, image_bytes, img_h, img_w, img_c, index_table, n, planes, mean,
                                                          std, wp, bias, y, st)
                                    : cudaGetErrorString(pe);
  if (e) return fail(VA_ERR_CUDA, "va_conv1_fused: %s", e);
  return VA_OK;
}

va_status va_conv2d_nhwc(const void* x, int n, int H, int W, int cin, int cin_pad, const float* w, const float* bias,
                         int cout, int ks, int relu, int pool, void* y, int force_bn, int force_r, va_stream_t stream) {
  if (!x || !w || !bias || !y) return fail(VA_ERR_INVALID, "va_conv2d_nhwc: NULL argument");
  if (cin > cin_pad) return fail(VA_ERR_INVALID, "va_conv2d_nhwc: cin %d > cin_pad %d", cin, cin_pad);
  if (va_status s = require_sm100()) return s;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  void* wp = nullptr;
  AsyncScratch scratch(st);
  VA_CUDA(scratch.alloc(&wp, (size_t)ks * ks * cout * cin_pad * 2));
  VA_CUDA(va::launch_pack_conv_w(w, wp, cout, cin, cin_pad, ks, st));
  va::ConvLayerDesc d;
  d.x = x; d.n = n; d.H = H; d.W = W; d.cin_pad = cin_pad; d.w_packed = wp; d.bias = bias; d.Cout = cout; d.ks = ks;
  d.relu = relu; d.pool = pool; d.y = y; d.y_f32 = nullptr; d.force_bn = force_bn; d.force_r = force_r; d.split6 = 0;
  const char* e = va::conv_layer_run(d, st);
  if (e) return fail(VA_ERR_CUDA, "va_conv2d_nhwc: %s", e);
  return VA_OK;
}

va_status va_linear(const void* x, int n, int in_features, const float* w, const float* bias, int out_features,
                    int relu, void* y_bf16, float* y_f32, int force_bn, va_stream_t stream) {
  if (!x || !w || !bias) return fail(VA_ERR_INVALID, "va_linear: NULL argument");
  if ((y_bf16 == nullptr) == (y_f32 == nullptr)) return fail(VA_ERR_INVALID, "va_linear: exactly one of y_bf16 / y_f32");
  if (in_features % 64 != 0) return fail(VA_ERR_INVALID, "va_linear: in_features %d must be a multiple of 64", in_features);
  if (va_status s = require_sm100()) return s;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  void* wp = nullptr;
  AsyncScratch scratch(st);
  VA_CUDA(scratch.alloc(&wp, (size_t)out_features * in_features * 2));
  VA_CUDA(va::launch_pack_fc_w(w, wp, out_features, in_features, 0, 0, st));
  va::ConvLayerDesc d;
  d.x = x; d.n = n; d.H = 1; d.W = 1; d.cin_pad = in_features; d.w_packed = wp; d.bias = bias; d.Cout = out_features;
  d.ks = 1; d.relu = relu; d.pool = 0; d.y = y_bf16; d.y_f32 = y_f32; d.force_bn = force_bn; d.force_r = 0; d.split6 = 0;
  void* ws = nullptr;
  if (in_features / 64 >= 32 && (in_features / 64) % va::kSplitK == 0 && force_bn == 0) {
    VA_CUDA(scratch.alloc(&ws, (size_t)va::kSplitK * n * out_features * 4));
    d.splitk_ws = static_cast<float*>(ws);
  }
  const char* e = va::conv_layer_run(d, st);
  if (e) return fail(VA_ERR_CUDA, "va_linear: %s", e);
  return VA_OK;
}

va_status va_fuse(const float* desc_s, const float* desc_t, const float* score_s, const float* score_t,
                  const int32_t* video_offsets, int V, int D, int C, int C_svm, const double* svm_w, const double* svm_b, float w_s,
                  float w_t, float* video_desc, float* video_scores, int32_t* score_pred, double* svm_scores,
                  int32_t* svm_pred, va_stream_t stream) {
  if (!video_offsets) return fail(VA_ERR_INVALID, "va_fuse: video_offsets is NULL");
  if (V < 0 || D < 1 || C < 1) return fail(VA_ERR_INVALID, "va_fuse: bad sizes V=%d D=%d C=%d", V, D, C);
  if (svm_w && (C_svm < 1 || C_svm > 4096)) return fail(VA_ERR_INVALID, "va_fuse: C_svm %d (rows of svm_w) out of range", C_svm);
  if (!svm_w) C_svm = 0;
  if ((svm_w == nullptr) != (svm_b == nullptr)) return fail(VA_ERR_INVALID, "va_fuse: svm_w and svm_b go together");
  if (svm_w && !(desc_s && desc_t)) return fail(VA_ERR_INVALID, "va_fuse: SVM scoring needs both descriptor arrays");
  if (w_s + w_t == 0.f) return fail(VA_ERR_INVALID, "va_fuse: w_s + w_t == 0");
  if (va_status s = require_sm100()) return s;
  VA_CUDA(va::launch_fuse(desc_s, desc_t, score_s, score_t, video_offsets, V, D, C, C_svm, svm_w, svm_b, w_s, w_t, video_desc,
                          video_scores, score_pred, svm_scores, svm_pred, static_cast<cudaStream_t>(stream)));
  return VA_OK;
}

va_status va_svm_decision(const double* X, int V, int F, const double* W, const double* b, int P, double* scores,
                          int32_t* pred, va_stream_t stream) {
  if (!X || !W || !b) return fail(VA_ERR_INVALID, "va_svm_decision: NULL argument");
  if (V < 0 || F < 1 || P < 1 || (size_t)(F + P) * sizeof(double) > 48 * 1024)
    return fail(VA_ERR_INVALID, "va_svm_decision: bad sizes V=%d F=%d P=%d", V, F, P);
  if (va_status s = require_sm100()) return s;
  VA_CUDA(va::launch_svm_decision(X, V, F, W, b, P, scores, pred, static_cast<cudaStream_t>(stream)));
  return VA_OK;
}

va_status va_consensus_update(float* sum, int32_t* count, const int32_t* video_ids, const float* descriptors, int B, int D,
                              va_stream_t stream) {
  if (!sum || !count || !video_ids || !descriptors) return fail(VA_ERR_INVALID, "va_consensus_update: NULL argument");
  if (B < 0 || D < 1) return fail(VA_ERR_INVALID, "va_consensus_update: bad sizes");
  if (va_status s = require_sm100()) return s;
  VA_CUDA(va::launch_consensus_update(sum, count, video_ids, descriptors, B, D, static_cast<cudaStream_t>(stream)));
  return VA_OK;
}

va_status va_pack_input_nchw(const float* x_nchw, int n, int channels, int height, int width, int c_pad, void* out_nhwc,
                             va_stream_t stream) {
  if (!x_nchw || !out_nhwc) return fail(VA_ERR_INVALID, "va_pack_input_nchw: NULL argument");
  if (channels > c_pad || !(c_pad == 16 || c_pad == 32 || c_pad == 64))
    return fail(VA_ERR_INVALID, "va_pack_input_nchw: channels %d / c_pad %d", channels, c_pad);
  if (va_status s = require_sm100()) return s;
  VA_CUDA(va::launch_nchw_to_nhwc(x_nchw, n, channels, height * width, c_pad, out_nhwc, static_cast<cudaStream_t>(stream)));
  return VA_OK;
}

va_status va_debug_conv_counters(long long* dev_counters16) {
  va::conv_set_debug_counters(dev_counters16);
  return VA_OK;
}

va_status va_profile_enable(int on) {
  g_prof_on = on != 0;
  return VA_OK;
}

va_status va_profile_read(double* tensor_ms, uint64_t* tensor_launches, double* tensor_flops) {
  double ms = 0.0;
  for (ProfSpan& sp : g_prof_spans) {
    VA_CUDA(cudaEventSynchronize(sp.b));
    float t = 0.f;
    VA_CUDA(cudaEventElapsedTime(&t, sp.a, sp.b));
    ms += t;
    cudaEventDestroy(sp.a);
    cudaEventDestroy(sp.b);
  }
  g_prof_spans.clear();
  if (tensor_ms) *tensor_ms = ms;
  if (tensor_launches) *tensor_launches = g_prof_launches;
  if (tensor_flops) *tensor_flops = g_prof_flops;
  g_prof_launches = 0;
  g_prof_flops = 0.0;
  return VA_OK;
}

va_status va_pack_input_nchw_split6(const float* x_nchw, int n, int channels, int height, int width, int k6_pad,
                                    void* out_nhwc, va_stream_t stream) {
  if (!x_nchw || !out_nhwc) return fail(VA_ERR_INVALID, "va_pack_input_nchw_split6: NULL argument");
  if (6 * channels > k6_pad || !(k6_pad == 32 || k6_pad == 128))
    return fail(VA_ERR_INVALID, "va_pack_input_nchw_split6: channels %d / k6_pad %d", channels, k6_pad);
  if (va_status s = require_sm100()) return s;
  VA_CUDA(va::launch_nchw_to_nhwc_split6(x_nchw, n, channels, height * width, k6_pad, out_nhwc,
                                         static_cast<cudaStream_t>(stream)));
  return VA_OK;
}


// ------------------------------------------------------------------------------------------------ training primitives
va_status va_maxpool2x2_nhwc(const void* x, int n, int H, int W, int C, void* y, void* codes, va_stream_t stream) {
  if (!x || !y) return fail(VA_ERR_INVALID, "va_maxpool2x2_nhwc: NULL argument");
  if ((H | W) & 1 || C % 8) return fail(VA_ERR_INVALID, "va_maxpool2x2_nhwc: H, W must be even and C a multiple of 8");
  if (va_status s = require_sm100()) return s;
  VA_CUDA(va::launch_maxpool_fwd(x, y, codes, n, H, W, C, static_cast<cudaStream_t>(stream)));
  return VA_OK;
}
va_status va_pool_bwd_codes(const void* dout, const void* codes, int n, int H, int W, int C, void* dZ, float* db,
                            va_stream_t stream) {
  if (!dout || !codes || !dZ) return fail(VA_ERR_INVALID, "va_pool_bwd_codes: NULL argument");
  if ((H | W) & 1 || C % 8 || 256 % (C / 8)) return fail(VA_ERR_INVALID, "va_pool_bwd_codes: H, W must be even and C/8 must divide 256");
  if (va_status s = require_sm100()) return s;
  VA_CUDA(va::launch_pool_bwd_codes(dout, codes, dZ, db, n, H, W, C, static_cast<cudaStream_t>(stream)));
  return VA_OK;
}
va_status va_relu_pool_bwd(const void* dout, const void* Y, int n, int H, int W, int C, int pooled, void* dZ, float* db,
                           va_stream_t stream) {
  if (!dout || !Y || !dZ) return fail(VA_ERR_INVALID, "va_relu_pool_bwd: NULL argument");
  if (C % 2 || (pooled && ((H | W) & 1))) return fail(VA_ERR_INVALID, "va_relu_pool_bwd: bad shape");
  if (va_status s = require_sm100()) return s;
  VA_CUDA(va::launch_relu_pool_bwd(dout, Y, dZ, db, n, H, W, C, pooled, static_cast<cudaStream_t>(stream)));
  return VA_OK;
}
va_status va_bias_grad(const void* dZ, long long rows, int C, float* db, va_stream_t stream) {
  if (!dZ || !db) return fail(VA_ERR_INVALID, "va_bias_grad: NULL argument");
  if (va_status s = require_sm100()) return s;
  VA_CUDA(va::launch_bias_grad(dZ, db, rows, C, static_cast<cudaStream_t>(stream)));
  return VA_OK;
}
va_status va_dropout(const void* x, const uint8_t* mask, long long total, float scale, int is_f32, void* y,
                     va_stream_t stream) {
  if (!x || !mask || !y) return fail(VA_ERR_INVALID, "va_dropout: NULL argument");
  if (va_status s = require_sm100()) return s;
  VA_CUDA(va::launch_dropout(x, mask, y, total, scale, is_f32, static_cast<cudaStream_t>(stream)));
  return VA_OK;
}
va_status va_conv2d_dgrad(const void* dZ, int n, int H, int W, int cout, const float* w, int cin, void* dX,
                          va_stream_t stream) {
  if (!dZ || !w || !dX) return fail(VA_ERR_INVALID, "va_conv2d_dgrad: NULL argument");
  if (cin % 64 || cout % 64) return fail(VA_ERR_INVALID, "va_conv2d_dgrad: channels must be multiples of 64 (got %d -> %d)", cout, cin);
  if (va_status s = require_sm100()) return s;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* wt = nullptr; void* wp = nullptr; float* zero_bias = nullptr;
  AsyncScratch scratch(st);
  VA_CUDA(scratch.alloc(reinterpret_cast<void**>(&wt), (size_t)cout * cin * 9 * 4));
  VA_CUDA(scratch.alloc(&wp, (size_t)9 * cin * cout * 2));
  VA_CUDA(scratch.alloc(reinterpret_cast<void**>(&zero_bias), (size_t)cin * 4));
  VA_CUDA(cudaMemsetAsync(zero_bias, 0, (size_t)cin * 4, st));
  VA_CUDA(va::launch_flip_transpose_conv_w(w, wt, cout, cin, 3, st));          // OIHW' [cin][cout][3][3]
  VA_CUDA(va::launch_pack_conv_w(wt, wp, cin, cout, cout, 3, st));             // the dgrad conv: cout -> cin channels
  va::ConvLayerDesc d;
  d.x = dZ; d.n = n; d.H = H; d.W = W; d.cin_pad = cout; d.w_packed = wp; d.bias = zero_bias; d.Cout = cin; d.ks = 3;
  d.relu = 0; d.pool = 0; d.y = dX; d.y_f32 = nullptr; d.force_bn = 0; d.force_r = 0; d.split6 = 0;
  const char* e = va::conv_layer_run(d, st);
  if (e) return fail(VA_ERR_CUDA, "va_conv2d_dgrad: %s", e);
  return VA_OK;
}
va_status va_linear_dgrad(const void* dY, int n, int out_features, const float* w, int in_features, void* dX,
                          va_stream_t stream) {
  if (!dY || !w || !dX) return fail(VA_ERR_INVALID, "va_linear_dgrad: NULL argument");
  if (in_features % 64 || out_features % 64) return fail(VA_ERR_INVALID, "va_linear_dgrad: features must be multiples of 64");
  if (va_status s = require_sm100()) return s;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  void* wp = nullptr; float* zero_bias = nullptr;
  AsyncScratch scratch(st);
  VA_CUDA(scratch.alloc(&wp, (size_t)in_features * out_features * 2));
  VA_CUDA(scratch.alloc(reinterpret_cast<void**>(&zero_bias), (size_t)in_features * 4));
  VA_CUDA(cudaMemsetAsync(zero_bias, 0, (size_t)in_features * 4, st));
  VA_CUDA(va::launch_pack_fc_w_t(w, wp, out_features, in_features, st));      // bf16 [in][out]
  va::ConvLayerDesc d;
  d.x = dY; d.n = n; d.H = 1; d.W = 1; d.cin_pad = out_features; d.w_packed = wp; d.bias = zero_bias; d.Cout = in_features;
  d.ks = 1; d.relu = 0; d.pool = 0; d.y = dX; d.y_f32 = nullptr; d.force_bn = 0; d.force_r = 0; d.split6 = 0;
  const char* e = va::conv_layer_run(d, st);
  if (e) return fail(VA_ERR_CUDA, "va_linear_dgrad: %s", e);
  return VA_OK;
}
va_status va_wgrad(const void* dZ, const void* X, int n, int H, int W, int cout, int cin, int cin_pad, int ks, float* dW,
                   va_stream_t stream) {
  if (!dZ || !X || !dW) return fail(VA_ERR_INVALID, "va_wgrad: NULL argument");
  if (ks != 1 && ks != 3) return fail(VA_ERR_INVALID, "va_wgrad: ks must be 1 or 3");
  if (cin_pad < cin || n <= 0) return fail(VA_ERR_INVALID, "va_wgrad: bad shape (n=%d cin=%d cin_pad=%d)", n, cin, cin_pad);
  if (va_status s = require_sm100()) return s;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* ws = nullptr;                       // [tap][cout][cin] accumulation planes of the 3x3 case
  AsyncScratch scratch(st);
  if (ks != 1) VA_CUDA(scratch.alloc(reinterpret_cast<void**>(&ws), (size_t)ks * ks * cout * cin * 4));
  const char* e = va::wgrad_run(dZ, X, n, H, W, cout, cin, cin_pad, ks, ws, dW, st);
  if (e) return fail(VA_ERR_CUDA, "va_wgrad: %s", e);
  return VA_OK;
}
va_status va_ce_train(const float* x, const float* w4, const float* b4, const int64_t* labels, int n, int D, int C,
                      float* logits, float* dlogits, float* loss, float* dw4, float* db4, float* dx, va_stream_t stream) {
  if (!x || !w4 || !b4 || !labels || !dlogits || !loss || !dw4 || !db4 || !dx) return fail(VA_ERR_INVALID, "va_ce_train: NULL argument");
  if (va_status s = require_sm100()) return s;
  VA_CUDA(va::launch_ce_train(x, w4, b4, labels, n, D, C, logits, dlogits, loss, dw4, db4, dx, static_cast<cudaStream_t>(stream)));
  return VA_OK;
}
va_status va_relu_bwd_f32_to_bf16(const float* dy, const float* y, long long n, void* dz, va_stream_t stream) {
  if (!dy || !y || !dz) return fail(VA_ERR_INVALID, "va_relu_bwd_f32_to_bf16: NULL argument");
  if (va_status s = require_sm100()) return s;
  VA_CUDA(va::launch_relu_bwd_f32_to_bf16(dy, y, dz, n, static_cast<cudaStream_t>(stream)));
  return VA_OK;
}
va_status va_sgd_momentum(float* param, const float* grad, float* momentum_buf, long long n, float lr, float momentum,
                          int first_step, float grad_scale, va_stream_t stream) {
  if (!param || !grad || !momentum_buf) return fail(VA_ERR_INVALID, "va_sgd_momentum: NULL argument");
  if (va_status s = require_sm100()) return s;
  VA_CUDA(va::launch_sgd_momentum(param, grad, momentum_buf, n, lr, momentum, first_step, grad_scale,
                                  static_cast<cudaStream_t>(stream)));
  return VA_OK;
}
va_status va_sgd_momentum_bf16g(float* param, const void* grad_bf16, float* momentum_buf, long long n, float lr, float momentum,
                                int first_step, float grad_scale, va_stream_t stream) {
  if (!param || !grad_bf16 || !momentum_buf) return fail(VA_ERR_INVALID, "va_sgd_momentum_bf16g: NULL argument");
  if (va_status s = require_sm100()) return s;
  VA_CUDA(va::launch_sgd_momentum_bf16g(param, grad_bf16, momentum_buf, n, lr, momentum, first_step, grad_scale,
                                        static_cast<cudaStream_t>(stream)));
  return VA_OK;
}
va_status va_transpose_bf16(const void* x, int n, int A, int B, void* y, va_stream_t stream) {
  if (!x || !y) return fail(VA_ERR_INVALID, "va_transpose_bf16: NULL argument");
  if (n > 65535) return fail(VA_ERR_INVALID, "va_transpose_bf16: n must be <= 65535");
  if (va_status s = require_sm100()) return s;
  VA_CUDA(va::launch_nhwc_to_nchw_bf16(x, y, n, 1, A, A, B, B, 1, static_cast<cudaStream_t>(stream)));
  return VA_OK;
}
va_status va_f32_to_bf16(const float* x, long long n, void* y, va_stream_t stream) {
  if (!x || !y) return fail(VA_ERR_INVALID, "va_f32_to_bf16: NULL argument");
  if (va_status s = require_sm100()) return s;
  VA_CUDA(va::launch_f32_to_bf16(x, y, n, static_cast<cudaStream_t>(stream)));
  return VA_OK;
}

va_status va_jpeg_decode(const uint8_t* bitstreams, const va_jpeg_image* images, int n_images, const uint16_t* qtables,
                         int n_qtables, const va_jpeg_huff* htables, int n_htables, uint8_t* out, va_stream_t stream) {
  if (n_images == 0) return VA_OK;
  if (!bitstreams || !images || !qtables || !htables || !out) return fail(VA_ERR_INVALID, "va_jpeg_decode: NULL argument");
  if (n_images < 0 || n_qtables <= 0 || n_htables <= 0) return fail(VA_ERR_INVALID, "va_jpeg_decode: bad counts");
  if (va_status s = require_sm100()) return s;
  const char* e = va::jpeg_decode_run(bitstreams, images, n_images, qtables, n_qtables, htables, n_htables, out,
                                      static_cast<cudaStream_t>(stream));
  if (e) return fail(VA_ERR_INVALID, "va_jpeg_decode: %s", e);
  return VA_OK;
}

va_status va_svm_fit(const double* X, const int32_t* class_index, int V, int F, int n_classes, double C, double bias,
                     double tol, int max_iter, double* coef, double* intercept, int32_t* epochs, double* work,
                     va_stream_t stream) {
  if (!X || !class_index || !coef || !intercept || !epochs || !work) return fail(VA_ERR_INVALID, "va_svm_fit: NULL argument");
  if (V < 1 || F < 1 || F > 1024) return fail(VA_ERR_INVALID, "va_svm_fit: bad sizes V=%d F=%d (F <= 1024)", V, F);
  if (n_classes < 2) return fail(VA_ERR_INVALID, "va_svm_fit: needs samples of at least 2 classes, got %d", n_classes);
  if (!(C > 0.0) || !(tol >= 0.0) || max_iter < 1) return fail(VA_ERR_INVALID, "va_svm_fit: bad C / tol / max_iter");
  if (va_status s = require_sm100()) return s;
  const char* e = va::svm_fit_run(X, class_index, V, F, n_classes, C, bias, tol, max_iter, coef, intercept, epochs, work,
                                  static_cast<cudaStream_t>(stream));
  if (e) return fail(VA_ERR_INVALID, "va_svm_fit: %s", e);
  return VA_OK;
}

va_status va_allreduce_bf16(const void* const* peer_ptrs, void* multicast_ptr, int world, int rank, long long n_elems, int n_ctas,
                            va_stream_t stream) {
  if (!peer_ptrs && !multicast_ptr) return fail(VA_ERR_INVALID, "va_allreduce_bf16: neither peer pointers nor a multicast pointer");
  if (va_status s = require_sm100()) return s;
  const char* e = va::allreduce_bf16_run(peer_ptrs, multicast_ptr, world, rank, n_elems, n_ctas, static_cast<cudaStream_t>(stream));
  if (e) return fail(VA_ERR_INVALID, "va_allreduce_bf16: %s", e);
  return VA_OK;
}

va_status va_reserve_sms(int sms, int launches) {
  if (sms < 0 || launches < 0) return fail(VA_ERR_INVALID, "va_reserve_sms: negative argument");
  va::conv_reserve_sms(sms, launches);
  return VA_OK;
}

va_status va_tvl1_debug_cycles(long long* dev_cycles) {
  va::tvl1_set_debug_cycles(dev_cycles);
  return VA_OK;
}

static va_tvl1_params tvl1_defaults() {
  va_tvl1_params d;
  d.tau = 0.25; d.lambda = 0.15; d.theta = 0.3; d.epsilon = 0.01; d.scale_step = 0.8; d.bound = 20.0;
  d.nscales = 5; d.warps = 5; d.iterations = 300; d.resize_w = 0; d.resize_h = 0; d.reserved = 0;
  return d;
}

size_t va_tvl1_workspace_bytes(int img_h, int img_w, const va_tvl1_params* params) {
  const va_tvl1_params d = params ? *params : tvl1_defaults();
  if (img_h < 1 || img_w < 1 || d.nscales < 1 || !(d.scale_step > 0.0 && d.scale_step < 1.0)) return 0;
  const bool rs = d.resize_w > 0 && d.resize_h > 0;
  return va::tvl1_workspace_bytes(rs ? d.resize_h : img_h, rs ? d.resize_w : img_w, d.nscales, d.scale_step);
}

va_status va_tvl1_flow(const uint8_t* images, size_t image_bytes, int img_h, int img_w, int img_c,
                       const int32_t* pair_table, int n, const va_tvl1_params* params, uint8_t* out_images,
                       size_t out_image_bytes, float* flow_f32, int32_t* iterations, void* workspace,
                       size_t workspace_bytes, va_stream_t stream) {
  if (n == 0) return VA_OK;
  if (!images || !pair_table || !out_images || !workspace) return fail(VA_ERR_INVALID, "va_tvl1_flow: NULL argument");
  const va_tvl1_params d = params ? *params : tvl1_defaults();
  if ((d.resize_w > 0) != (d.resize_h > 0) || d.resize_w < 0 || d.resize_h < 0)
    return fail(VA_ERR_INVALID, "va_tvl1_flow: resize_w / resize_h must both be set (or both 0)");
  const int fh = d.resize_h > 0 ? d.resize_h : img_h, fw = d.resize_w > 0 ? d.resize_w : img_w;     // size the flow is computed at
  if (n < 0 || fh < 16 || fw < 16 || img_h < 1 || img_w < 1 || (img_c != 1 && img_c != 3))
    return fail(VA_ERR_INVALID, "va_tvl1_flow: bad sizes n=%d %dx%dx%d -> %dx%d (flow images >= 16 pixels, 1 or 3 channels)", n, img_h,
                img_w, img_c, fh, fw);
  if (image_bytes < (size_t)img_h * img_w * img_c || out_image_bytes < (size_t)fh * fw)
    return fail(VA_ERR_INVALID, "va_tvl1_flow: image_bytes / out_image_bytes smaller than an image");
  if (!(d.tau > 0.0) || !(d.lambda > 0.0) || !(d.theta > 0.0) || !(d.epsilon >= 0.0) || !(d.bound > 0.0) ||
      !(d.scale_step > 0.0 && d.scale_step < 1.0) || d.nscales < 1 || d.nscales > 8 || d.warps < 1 || d.iterations < 1)
    return fail(VA_ERR_INVALID, "va_tvl1_flow: bad parameters");
  if (fw > 4096 || fh > 4096) return fail(VA_ERR_UNSUPPORTED, "va_tvl1_flow: %dx%d: flow images larger than 4096 pixels", fh, fw);
  if (va_status s = require_sm100()) return s;
  const char* e = va::tvl1_run(images, image_bytes, img_h, img_w, img_c, fh, fw, pair_table, n, d.tau, d.lambda, d.theta, d.nscales,
                               d.warps, d.epsilon, d.iterations, d.scale_step, d.bound, out_images, out_image_bytes, flow_f32,
                               iterations, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
  if (e) return fail(VA_ERR_INVALID, "va_tvl1_flow: %s", e);
  return VA_OK;
}

va_status va_synth_fill(uint8_t* images, size_t image_bytes, int n_images, int img_h, int img_w, int img_c, uint32_t seed,
                        uint32_t first_id, va_stream_t stream) {
  if (!images) return fail(VA_ERR_INVALID, "va_synth_fill: NULL");
  if (image_bytes < (size_t)img_h * img_w * img_c) return fail(VA_ERR_INVALID, "va_synth_fill: image_bytes too small");
  if (va_status s = require_sm100()) return s;
  VA_CUDA(va::launch_synth_fill(images, image_bytes, n_images, img_h, img_w, img_c, seed, first_id,
                                static_cast<cudaStream_t>(stream)));
  return VA_OK;
}

}  // extern "C"
