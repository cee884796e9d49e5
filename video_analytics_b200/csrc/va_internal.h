// Internal (non-ABI) declarations shared by the translation units of libva_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stddef.h>
#include <algorithm>

namespace va {

void count_launch();   // bumps the library-wide launch counter (va_launch_count)

cudaError_t launch_preprocess(const uint8_t* images, size_t image_bytes, int img_h, int img_w, int img_c,
                              const int32_t* table, int n, int planes, int crop, const float* mean,
                              const float* stdv, int c_pad, int out_mode, void* out, cudaStream_t st);
cudaError_t launch_synth_fill(uint8_t* images, size_t image_bytes, int n_images, int H, int W, int C, uint32_t seed,
                              uint32_t first_id, cudaStream_t st);
cudaError_t launch_pack_conv_w(const float* w, void* out, int Cout, int Cin, int cin_pad, int ks, cudaStream_t st);
cudaError_t launch_pack_fc_w(const float* w, void* out, int n_out, int n_in, int chan, int hw, cudaStream_t st);
cudaError_t launch_transpose_f32(const float* w, float* out, int rows, int cols, cudaStream_t st);
cudaError_t launch_head(const float* desc, const float* w4t, const float* b4, int n, int D, int C, float* logits,
                        float* probs, int32_t* pred, cudaStream_t st);
cudaError_t launch_fuse(const float* desc_s, const float* desc_t, const float* score_s, const float* score_t,
                        const int32_t* offs, int V, int D, int C, int Csvm, const double* svm_w, const double* svm_b, float w_s,
                        float w_t, float* video_desc, float* video_scores, int32_t* score_pred, double* svm_scores,
                        int32_t* svm_pred, cudaStream_t st);
cudaError_t launch_svm_decision(const double* X, int V, int F, const double* W, const double* b, int P, double* scores,
                                int32_t* pred, cudaStream_t st);

cudaError_t launch_consensus_update(float* sum, int32_t* count, const int32_t* video_ids, const float* fv, int B, int D,
                                    cudaStream_t st);

cudaError_t launch_pack_conv_w_split6(const float* w, void* out, int Cout, int Cin, int k6_pad, int ks, cudaStream_t st);
cudaError_t launch_pack_fc_w_split6(const float* w, void* out, int n_out, int chan, int hw, int permute, cudaStream_t st);
cudaError_t launch_nchw_to_nhwc_split6(const float* x, int n, int c, int hw, int k6_pad, void* out, cudaStream_t st);
cudaError_t launch_nchw_to_nhwc(const float* x, int n, int c, int hw, int c_pad, void* out, cudaStream_t st);

// ---- training-step kernels (va_train_kernels.cu, va_wgrad_tc.cu)
cudaError_t launch_maxpool_fwd(const void* x, void* y, void* codes, int n, int H, int W, int C, cudaStream_t st);
cudaError_t launch_pool_bwd_codes(const void* dP, const void* codes, void* dZ, float* db, int n, int H, int W, int C, cudaStream_t st);
cudaError_t launch_relu_pool_bwd(const void* dP, const void* Y, void* dZ, float* db, int n, int H, int W, int C, int pooled,
                                 cudaStream_t st);
cudaError_t launch_bias_grad(const void* dZ, float* db, long long rows, int C, cudaStream_t st);
cudaError_t launch_dropout(const void* x, const uint8_t* mask, void* y, long long total, float scale, int is_f32, cudaStream_t st);
cudaError_t launch_nhwc_to_nchw_bf16(const void* x, void* y, int n, int H, int W, int Wp, int C, int Cs, int nshift,
                                     cudaStream_t st);
cudaError_t launch_f32_to_bf16(const float* x, void* y, long long n, cudaStream_t st);
cudaError_t launch_sgd_momentum(float* p, const float* g, float* buf, long long n, float lr, float momentum, int first_step,
                                float grad_scale, cudaStream_t st);
cudaError_t launch_sgd_momentum_bf16g(float* p, const void* g_bf16, float* buf, long long n, float lr, float momentum,
                                      int first_step, float grad_scale, cudaStream_t st);
cudaError_t launch_ce_train(const float* x, const float* w4, const float* b4, const int64_t* labels, int n, int D, int C,
                            float* logits, float* dlogits, float* loss, float* dw4, float* db4, float* dx, cudaStream_t st);
cudaError_t launch_relu_bwd_f32_to_bf16(const float* dy, const float* y, void* dz, long long n, cudaStream_t st);
cudaError_t launch_flip_transpose_conv_w(const float* w, float* out, int Cout, int Cin, int ks, cudaStream_t st);
cudaError_t launch_pack_fc_w_t(const float* w, void* out, int n_out, int n_in, cudaStream_t st);
const char* wgrad_run(const void* dz, const void* x, int n, int H, int W, int Cout, int Cin, int cin_pad, int ks,
                      float* dwt_ws, float* dw, cudaStream_t st);

// ---- JPEG decode (va_jpeg.cu)
const char* jpeg_decode_run(const unsigned char* bitstreams, const void* images_host, int n_images,
                            const unsigned short* qtables_host, int n_q, const void* htables_host, int n_h,
                            unsigned char* out, cudaStream_t st);

// ---- linear SVM fit (va_svm_fit.cu)
const char* svm_fit_run(const double* X, const int32_t* class_index, int V, int F, int n_classes, double C, double bias,
                        double tol, int max_iter, double* coef, double* intercept, int32_t* epochs, double* work,
                        cudaStream_t st);

// ---- gradient all-reduce over NVLink peer memory / NVSwitch multicast (va_allreduce.cu)
const char* allreduce_bf16_run(const void* const* peer_ptrs, void* multicast_ptr, int world, int rank, long long n_elems, int n_ctas,
                               cudaStream_t st);

// ---- TV-L1 optical flow (va_tvl1.cu)
size_t tvl1_workspace_bytes(int h, int w, int nscales, double scale_step);
void tvl1_set_debug_cycles(long long* dev);
// src_h x src_w: the stored frames; h x w: the size the flow is computed (and written) at -- different = cv::resize first
const char* tvl1_run(const uint8_t* images, size_t image_bytes, int src_h, int src_w, int c, int h, int w, const int32_t* pairs, int n,
                     double tau, double lambda, double theta, int nscales, int warps, double epsilon, int iterations,
                     double scale_step, double bound, uint8_t* out, size_t out_bytes, float* flow, int32_t* stats,
                     void* workspace, size_t workspace_bytes, cudaStream_t st);

// ---- tensor-core conv / linear layer (va_conv_tc.cu)
struct ConvLayerDesc {
  const void* x;        // bf16 NHWC [n][H][W][cin_pad]
  int n, H, W, cin_pad;
  const void* w_packed; // bf16 [ks*ks][Cout][cin_pad]
  const float* bias;    // fp32 [Cout]
  int Cout, ks;
  int relu, pool;
  void* y;              // bf16 NHWC [n][H>>pool][W>>pool][Cout] (unless y_f32)
  float* y_f32;         // fp32 [n][Cout] (H=W=1 only)
  int force_bn, force_r;
  int split6;           // fp32-accuracy mode: y has 6*Cout channels (slice blocks), x/w_packed already carry 6x channels
  float* splitk_ws = nullptr;   // fully-connected layers: fp32 [kSplitK][n][Cout] scratch; when given (and K is long
                                // enough) the K loop is cut into kSplitK slices reduced in a fixed order by a second kernel
};
constexpr int kSplitK = 8;
// Plans (tile shape, kernel variant, tensor maps) and launches one layer.  Returns nullptr on success or a
// static/thread-local error string.
const char* conv_layer_run(const ConvLayerDesc& d, cudaStream_t st);
void conv_reserve_sms(int sms, int launches);   // the next `launches` layer launches leave `sms` SMs free (0 = off)
void conv_set_debug_counters(long long* dev_buf);
long long* conv_get_debug_counters();

// ---- fused snippet gather + conv1_1 (va_conv1_fused.cu)
bool conv1_fused_supported(int planes, int img_c, int crop);
int conv1_fused_packed_bytes(int cin);     // size of the dense-K weight pack for `cin` stacked input channels
cudaError_t launch_pack_conv1_fused_w(const float* w_oihw, void* out, int cin, cudaStream_t st);
// images/table as for launch_preprocess; y = bf16 NHWC [n][224][224][64] = ReLU(conv3x3(normalised crop) + bias)
const char* conv1_fused_run(const uint8_t* images, size_t image_bytes, int img_h, int img_w, int img_c, const int32_t* table,
                            int n, int planes, const float* mean, const float* stdv, const void* w_fused, const float* bias,
                            void* y, cudaStream_t st);   // diagnostics: per-role stall cycles of CTA 0 (nullptr = off)

}  // namespace va
