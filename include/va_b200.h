/*
 * va_b200.h -- C-ABI of libva_b200.so: the B200-native (sm_100a) two-stream action-recognition forward path.
 *
 * The reference (arindamrc/video_analytics, Sheet03) is pure Python on stock PyTorch and exposes no FFI; the
 * drop-in boundary is therefore its Python class API (see video_analytics_b200/{spatialModel,temporalModel,
 * combinedModel}.py) and THIS header is what that Python layer binds with ctypes.  Every entry point names the
 * reference lines whose work it replaces.  Conventions:
 *   - plain pointers and sizes only; every data pointer is a BORROWED DEVICE pointer unless it says "host";
 *   - every call is asynchronous on the caller's CUDA stream (va_stream_t == cudaStream_t, 0 = default stream);
 *   - return value 0 = ok, non-zero = error; va_last_error() returns the message of the calling thread's last
 *     failure.  There is no CPU fallback: without a CUDA device / the sm_100a image every call fails loudly;
 *   - one va_handle per process+GPU+stream kind; a handle is not thread-safe.
 */
#ifndef VA_B200_H_
#define VA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct va_handle va_handle;
typedef void* va_stream_t; /* cudaStream_t */
typedef int va_status;     /* 0 == VA_OK */

enum { VA_OK = 0, VA_ERR_INVALID = 1, VA_ERR_CUDA = 2, VA_ERR_UNSUPPORTED = 3 };
enum { VA_STREAM_SPATIAL = 0, VA_STREAM_TEMPORAL = 1 };

const char* va_last_error(void);
/* library/ABI version, bumped on any signature change */
int va_abi_version(void);
/* SM count of the current device (grid sizing in callers / bench) */
va_status va_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---------------------------------------------------------------------------------------------------------
 * Network handle.  Replaces SpatialNetwork.__init__ / TemporalNetwork.__init__ model construction
 * (reference Sheet03/spatialModel.py:110-113,136-152; temporalModel.py:122-126,149-181): VGG16-D features
 * (13x conv3x3+bias+ReLU, 5x maxpool2x2) + classifier 25088-4096-4096-desc_dim-n_classes.
 *   in_channels: 3 (spatial) or 2*L = 20 (temporal).  max_batch: snippets per internal chunk (workspace size).
 * --------------------------------------------------------------------------------------------------------- */
va_status va_create(va_handle** out, int stream_kind, int in_channels, int n_classes, int desc_dim, int max_batch);
/* Same with an arithmetic mode: precision 0 = bf16 storage / fp32 accumulate (the throughput path); 1 = fp32-accuracy
 * parity mode: every activation and weight is carried as three bf16 slices (hi, mid, lo) and each layer's GEMM sums
 * the six leading cross terms (K is 6x longer), reproducing an fp32 forward to ~1e-6 relative (north_star: 1e-5). */
va_status va_create_ex(va_handle** out, int stream_kind, int in_channels, int n_classes, int desc_dim, int max_batch,
                       int precision);
va_status va_destroy(va_handle* h);

/* channels the NHWC bf16 network input must be padded to for this handle (16 for 3, 32 for 20) */
int va_input_channels_padded(const va_handle* h);

/* Load the reference state_dict.  `tensors` is a HOST array of 34 DEVICE pointers to fp32 tensors in the
 * reference's state_dict order (features.{0,2,5,7,10,12,14,17,19,21,24,26,28}.{weight,bias} as OIHW / [O],
 * classifier.{0,3,6,9}.{weight,bias} as [out,in] / [out]); cf. spatialModel.py:256-261 (checkpoint "model").
 * Packs conv weights to [tap][Cout][Cin_pad] bf16, FC1..3 to bf16 (FC1 columns permuted from the reference's
 * NCHW flatten order c*49+h*7+w, spatialModel.py:172, to NHWC), FC4 kept fp32. */
va_status va_load_weights(va_handle* h, const void* const* tensors, int n_tensors, va_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * K1 fused snippet preprocess.  Replaces the per-item CPU pipeline RandomCrop(224) + RandomHorizontalFlip +
 * ToTensor + Normalize (reference utils.py:137-151) applied in SpatialDataset.__getitem__
 * (spatialModel.py:64-81) and, once per flow image, in TemporalDataset.__getitem__ (temporalModel.py:67-92),
 * including the x/y interleaved stacking of 2L flow images (temporalModel.py:80-90).
 *   images:      u8 image store; image `id` starts at images + id*image_bytes, layout [img_h][img_w][img_c]
 *   index_table: int32 [n][planes][4] = {image id, crop_i (top), crop_j (left), flip}; planes = 1 (RGB image
 *                with img_c = 3) or 2L (one 1-channel image per output channel)
 *   mean/std:    HOST float arrays [planes*img_c], applied as ((u8/255) - mean) / std in IEEE fp32
 *   out_mode 0:  bf16 NHWC [n][crop][crop][c_pad] (padding channels zero) -- the network input
 *   out_mode 1:  fp32 NCHW [n][planes*img_c][crop][crop]                  -- the reference tensor, bit-exact
 * --------------------------------------------------------------------------------------------------------- */
va_status va_preprocess(const uint8_t* images, size_t image_bytes, int img_h, int img_w, int img_c,
                        const int32_t* index_table, int n, int planes, int crop, const float* mean,
                        const float* std, int c_pad, int out_mode, void* out, va_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Network forward (eval mode).  Replaces the inlined forward of validate()
 * (spatialModel.py:212-221, temporalModel.py:241-250): features -> flatten -> classifier[:9] -> featureVectors
 * -> classifier[9] -> logits -> argmax; plus softmax class scores (notes.txt:113-116).
 *   in_nhwc: bf16 [n][224][224][c_pad] from va_preprocess.  Any output pointer may be NULL.
 *   descriptors fp32 [n][desc_dim]; logits fp32 [n][n_classes]; probs fp32 [n][n_classes]; pred int32 [n]
 *   (0-based index of the first maximum, torch semantics).
 * --------------------------------------------------------------------------------------------------------- */
va_status va_forward(va_handle* h, const void* in_nhwc, int n, float* descriptors, float* logits, float* probs,
                     int32_t* pred, va_stream_t stream);

/* Fused front end: the same forward, but fed from the image store.  The snippet transform (va_preprocess: utils.py:137-151,
 * spatialModel.py:64-81, temporalModel.py:67-92) is gathered straight into the first convolution's shared-memory operand
 * (csrc/va_conv1_fused.cu), so no preprocessed tensor exists in HBM.  Arguments as va_preprocess (crop is 224) and
 * va_forward; planes * img_c must equal the handle's in_channels; (planes, img_c) = (1, 3) or (1..23, 1); `images` must be
 * 16-byte aligned with image_bytes a multiple of 16 (the loader copies 16-byte blocks; true for 240x320x3 and 256x340).  Results are
 * those of va_preprocess + va_forward up to the summation order inside the first layer's fp32 accumulation.
 * Returns VA_ERR_UNSUPPORTED for precision-1 handles (use va_preprocess + va_forward). */
va_status va_forward_store(va_handle* h, const uint8_t* images, size_t image_bytes, int img_h, int img_w, int img_c,
                           const int32_t* index_table, int n, int planes, const float* mean, const float* std,
                           float* descriptors, float* logits, float* probs, int32_t* pred, va_stream_t stream);

/* The fused gather + first-layer kernel alone (parity tests, profiling): y bf16 NHWC [n][224][224][64] =
 * ReLU(conv3x3(normalised crop, w) + bias); w fp32 OIHW [64][planes*img_c][3][3]. */
va_status va_conv1_fused(const uint8_t* images, size_t image_bytes, int img_h, int img_w, int img_c,
                         const int32_t* index_table, int n, int planes, const float* mean, const float* std,
                         const float* w, const float* bias, void* y, va_stream_t stream);

/* Layer primitives behind va_forward, exported for per-layer parity tests and profiling.
 * conv: x bf16 NHWC [n][H][W][cin_pad]; w fp32 OIHW [cout][cin][ks][ks]; y bf16 NHWC [n][H(/2)][W(/2)][cout].
 *       ks in {1,3}, stride 1, zero pad (ks-1)/2; optional fused ReLU and 2x2/2 max-pool.
 *       force_bn in {0 (auto),64,128,256}; force_r in {0 (auto), 1 (one tap per stage), 3 (vertical tap reuse),
 *       9 (whole 3x3 filter per stage, Cin_pad 16/32 only), 10 (one haloed box per 16x8 tile, 64->64 layers)} select
 *       kernel variants.
 * linear: x bf16 [n][in]; w fp32 [out][in]; y bf16 [n][out] and/or y_f32 fp32 [n][out] (exactly one). */
va_status va_conv2d_nhwc(const void* x, int n, int H, int W, int cin, int cin_pad, const float* w,
                         const float* bias, int cout, int ks, int relu, int pool, void* y, int force_bn,
                         int force_r, va_stream_t stream);
va_status va_linear(const void* x, int n, int in_features, const float* w, const float* bias, int out_features,
                    int relu, void* y_bf16, float* y_f32, int force_bn, va_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * K4 per-video consensus + late fusion.  Replaces the AverageMeter loop (spatialModel.py:223-228,
 * utils.py:154-171: sequential fp32 sum in snippet order, then / count), combineDescriptors
 * (combinedModel.py:9-26: [spatial 256 | temporal 256]) and LinearSVC.predict (combinedModel.py:38:
 * argmax_c X.W[c] + b[c] in fp64), plus the score-averaging protocol of notes.txt:113-116,121-124,225-230.
 *   desc_s/desc_t  fp32 [N][D]; score_s/score_t fp32 [N][C] (softmax); snippets of video v are rows
 *   video_offsets[v] .. video_offsets[v+1]-1 (int32 [V+1]) of all four arrays.
 *   svm_w fp64 [C_svm][2D], svm_b fp64 [C_svm] (NULL = skip SVM scoring).  C_svm is the number of classes the SVM was
 *   FITTED on (len(np.unique(y)), as LinearSVC: 25 on mini-UCF-101) and is independent of the network's score width C (101).
 *   out: video_desc fp32 [V][2D]; video_scores fp32 [V][C] = (w_s*mean_s + w_t*mean_t)/(w_s+w_t);
 *        score_pred int32 [V]; svm_scores fp64 [V][C_svm]; svm_pred int32 [V] in [0, C_svm).  Any output may be NULL.
 * --------------------------------------------------------------------------------------------------------- */
va_status va_fuse(const float* desc_s, const float* desc_t, const float* score_s, const float* score_t,
                  const int32_t* video_offsets, int V, int D, int C, int C_svm, const double* svm_w, const double* svm_b,
                  float w_s, float w_t, float* video_desc, float* video_scores, int32_t* score_pred,
                  double* svm_scores, int32_t* svm_pred, va_stream_t stream);

/* LinearSVC.predict on fp64 rows (combinedModel.py:38 on the fp64 matrix combineDescriptors returns, :19-25):
 * scores[v][c] = X[v] . W[c] + b[c] in fp64, pred[v] = index of the first maximum.  X fp64 [V][F], W fp64 [P][F],
 * b fp64 [P]; scores fp64 [V][P] and pred int32 [V] (either may be NULL). */
va_status va_svm_decision(const double* X, int V, int F, const double* W, const double* b, int P, double* scores,
                          int32_t* pred, va_stream_t stream);

/* Running per-video descriptor sums: the device form of the reference's per-sample AverageMeter loop in
 * train()/validate() (spatialModel.py:183-188, 223-228).  For b = 0..B-1 in order: sum[video_ids[b]][:] +=
 * descriptors[b][:] (fp32, sequential -> the reference's summation order), count[video_ids[b]] += 1.
 * sum fp32 [V][D], count int32 [V], video_ids int32 [B], descriptors fp32 [B][D]. */
va_status va_consensus_update(float* sum, int32_t* count, const int32_t* video_ids, const float* descriptors, int B,
                              int D, va_stream_t stream);

/* Reference-layout network input (fp32 NCHW [n][channels][height][width], the tensor SpatialDataset/
 * TemporalDataset.__getitem__ return and train()/validate() feed as `ip`, spatialModel.py:166-171) -> bf16 NHWC
 * [n][height][width][c_pad] for va_forward. */
va_status va_pack_input_nchw(const float* x_nchw, int n, int channels, int height, int width, int c_pad,
                             void* out_nhwc, va_stream_t stream);

/* Same for a precision-1 handle: bf16 NHWC [n][height][width][k6_pad] with the six slice blocks (hi,hi,hi,mid,mid,lo)
 * of each channel; k6_pad = va_input_channels_padded(h) (32 for 3 channels, 128 for 20). */
va_status va_pack_input_nchw_split6(const float* x_nchw, int n, int channels, int height, int width, int k6_pad,
                                    void* out_nhwc, va_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Training-step primitives.  Replace loss.backward() + optimizer.step() of
 * train() (reference spatialModel.py:178-181: CrossEntropyLoss, SGD(lr, momentum=0.9)) and the train-mode forward
 * (Dropout after FC1..FC3, :141-152).  Activations / activation gradients are bf16 NHWC, parameter gradients fp32
 * in the reference's own layouts (OIHW, [out][in]) so that they line up with the fp32 master parameters.
 *   va_maxpool2x2_nhwc   : MaxPool2d(2,2) forward; with codes != NULL (uint32 [n][H/2][W/2][C/8]) it also records, in 4 bits
 *                          per window and channel, which element takes the gradient (0..3, first maximum in scan order;
 *                          4 = none: the maximum is not positive) -- 1/16 of the activation's bytes
 *   va_pool_bwd_codes    : dZ (and db) of relu+pool from those codes, without the un-pooled activation
 *   va_relu_pool_bwd     : dZ = un-pool(dout) * (Y > 0); pooled=0: dout is already the gradient of Y; with db != NULL
 *                          the bias gradient db[c] = sum dZ[.., c] is produced by the same pass
 *   va_bias_grad         : db[c] = sum_rows dZ[row][c]
 *   va_dropout           : y = x * scale where mask (u8) != 0 else 0 (forward and backward are the same map)
 *   va_conv2d_dgrad      : dX = conv3x3(dZ, rot180(W) with channel roles swapped)  -- tcgen05 layer kernel
 *   va_linear_dgrad      : dX = dY . W                                            -- tcgen05 layer kernel
 *   va_wgrad             : dW[co][ci][r][s] = sum_pixels dZ[p][co] X[p + (r,s)][ci] -- tcgen05 weight-gradient GEMM that
 *                          reads dZ [n][H][W][cout] and X [n][H][W][cin_pad] (bf16 NHWC) as MN-major operands; dW is
 *                          fp32 [cout][cin][ks][ks] with cin <= cin_pad the real input channels (channels >= cin of X
 *                          must be zero).  cout % 8 == 0; cin_pad is 16, 32 or >= 64 (% 8 == 0); ks is 3 (pad 1) or 1.
 *                          ks=1, H=W=1, n=batch gives the fully-connected dW = dY^T . X
 *   va_ce_train          : fp32 logit layer forward + mean cross-entropy + its backward (dlogits, dW4, db4, dx)
 *   va_transpose_bf16    : [n][A][B] -> [n][B][A] (NHWC <-> the reference's NCHW flatten order in front of FC1)
 *   va_relu_bwd_f32_to_bf16, va_f32_to_bf16 : glue between the fp32 descriptor layer and the bf16 stack
 *   va_sgd_momentum      : buf = g (first step) | momentum*buf + g;  p -= lr*buf   (grad_scale multiplies g first,
 *                          e.g. 1/world_size after a gradient all-reduce); va_sgd_momentum_bf16g reads the gradient as
 *                          bf16 (the compressed all-reduce payload of the data-parallel step, 270 MB instead of 541 MB)
 * --------------------------------------------------------------------------------------------------------- */
va_status va_maxpool2x2_nhwc(const void* x, int n, int H, int W, int C, void* y, void* codes, va_stream_t stream);
va_status va_pool_bwd_codes(const void* dout, const void* codes, int n, int H, int W, int C, void* dZ, float* db,
                            va_stream_t stream);
va_status va_relu_pool_bwd(const void* dout, const void* Y, int n, int H, int W, int C, int pooled, void* dZ, float* db,
                           va_stream_t stream);
va_status va_bias_grad(const void* dZ, long long rows, int C, float* db, va_stream_t stream);
va_status va_dropout(const void* x, const uint8_t* mask, long long total, float scale, int is_f32, void* y,
                     va_stream_t stream);
va_status va_conv2d_dgrad(const void* dZ, int n, int H, int W, int cout, const float* w, int cin, void* dX,
                          va_stream_t stream);
va_status va_linear_dgrad(const void* dY, int n, int out_features, const float* w, int in_features, void* dX,
                          va_stream_t stream);
va_status va_wgrad(const void* dZ, const void* X, int n, int H, int W, int cout, int cin, int cin_pad, int ks, float* dW,
                   va_stream_t stream);
va_status va_ce_train(const float* x, const float* w4, const float* b4, const int64_t* labels, int n, int D, int C,
                      float* logits, float* dlogits, float* loss, float* dw4, float* db4, float* dx, va_stream_t stream);
va_status va_relu_bwd_f32_to_bf16(const float* dy, const float* y, long long n, void* dz, va_stream_t stream);
va_status va_sgd_momentum(float* param, const float* grad, float* momentum_buf, long long n, float lr, float momentum,
                          int first_step, float grad_scale, va_stream_t stream);
va_status va_sgd_momentum_bf16g(float* param, const void* grad_bf16, float* momentum_buf, long long n, float lr, float momentum,
                                int first_step, float grad_scale, va_stream_t stream);
va_status va_transpose_bf16(const void* x, int n, int A, int B, void* y, va_stream_t stream);
va_status va_f32_to_bf16(const float* x, long long n, void* y, va_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Baseline-JPEG decode into the image store.  Replaces Image.open(...) of
 * SpatialDataset / TemporalDataset.__getitem__ (reference spatialModel.py:76-79, temporalModel.py:85-88) for the files
 * cv2.imwrite writes (utils.py:116-120): 8-bit baseline sequential Huffman JPEG, one component or YCbCr 4:2:0 / 4:4:4.
 * Pixels are bit-identical to Pillow's (libjpeg-turbo: islow IDCT, fancy h2v2 upsampling, 16-bit fixed-point YCbCr->RGB).
 * The host parses the headers (video_analytics_b200/jpeg.py) and passes, per image, where the entropy-coded segment
 * starts and which tables it uses; the whole files are in `bitstreams` (device).  Output: [H][W] (1 component) or
 * [H][W][3] RGB u8 at out + out_offset -- e.g. image id * image_bytes of a va_preprocess store.
 *   qtables: HOST uint16 [n_qtables][64] in NATURAL (row-major) order.
 *   htables: HOST derived Huffman tables (jpeg_make_d_derived_tbl): maxcode[l] = largest code of length l or -1,
 *            valoffset[l] = index of the first symbol of length l minus its code.
 * --------------------------------------------------------------------------------------------------------- */
typedef struct va_jpeg_image {
  unsigned long long scan_offset;    /* first entropy-coded byte, relative to `bitstreams` */
  unsigned long long out_offset;     /* byte offset of the decoded image in `out` */
  unsigned int scan_bytes;           /* bytes from scan_offset to the end of the file */
  unsigned int restart_interval;     /* MCUs between RSTn markers (DRI), 0 = none */
  unsigned short width, height;
  unsigned char n_comp;              /* 1 or 3 */
  unsigned char sampling;            /* 0 = one component, 1 = 4:4:4, 2 = 4:2:0 */
  unsigned char qt[3], dc[3], ac[3]; /* table indices per component */
  unsigned char pad[9];
} va_jpeg_image;                     /* 48 bytes */
typedef struct va_jpeg_huff {
  int maxcode[18];
  int valoffset[17];
  unsigned char huffval[256];
  int pad;
} va_jpeg_huff;                      /* 400 bytes */
va_status va_jpeg_decode(const uint8_t* bitstreams, const va_jpeg_image* images, int n_images, const uint16_t* qtables,
                         int n_qtables, const va_jpeg_huff* htables, int n_htables, uint8_t* out, va_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Linear SVM fit of the late-fusion step.  Replaces `svm.LinearSVC().fit(svmTrainData,
 * svmTrainLabels)` (combinedModel.py:34-35): scikit-learn's defaults = LIBLINEAR's L2-regularised L2-loss dual
 * coordinate descent, one-vs-rest, intercept as an extra constant feature `bias` (intercept_scaling, 1), C = 1,
 * tol = 1e-4 on the projected-gradient range of an epoch, max_iter = 1000 epochs.  fp64 throughout.
 *   X fp64 [V][F] row-major (F <= 1024; the path has F = 2*256); class_index int32 [V] in [0, n_classes) = index into
 *   the sorted unique labels (classes_).  n_classes == 2 fits ONE problem with class 1 positive (coef [1][F], like
 *   scikit-learn); otherwise n_classes problems.  bias <= 0: no intercept.
 *   out: coef fp64 [P][F], intercept fp64 [P], epochs int32 [P] (epochs run per problem; == max_iter means the
 *   tolerance was not reached, scikit-learn's ConvergenceWarning), P = n_classes == 2 ? 1 : n_classes.
 *   work: fp64 [(P + 1) * V + (P * V + 1) / 2] scratch (Q_ii, dual variables, active sets as int32).
 * --------------------------------------------------------------------------------------------------------- */
va_status va_svm_fit(const double* X, const int32_t* class_index, int V, int F, int n_classes, double C, double bias,
                     double tol, int max_iter, double* coef, double* intercept, int32_t* epochs, double* work,
                     va_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * TV-L1 optical-flow producer.  Produces the `flow_x_%04d` / `flow_y_%04d` images that
 * TemporalDataset.__getitem__ opens (reference temporalModel.py:76-86; parameters.py:27,38-39 -- the reference only
 * consumes them, the TSN `dense_flow` tool that wrote `..._flow_img_tvl1_gpu` is third-party): per frame pair, grey
 * conversion (cv::cvtColor BGR2GRAY), Zach/Pock/Bischof TV-L1 with OpenCV's CUDA structure and defaults, then the 8-bit
 * mapping of [-bound, bound] to [0, 255].  The arithmetic contract is oracle/tvl1.py (separately rounded fp32 operations;
 * the kernel is bit-identical to it).  A thread-block cluster (4 / 8 / 16 CTAs per pyramid level) iterates a pair on chip.
 *   images:     u8 store as for va_preprocess, image `id` at images + id*image_bytes, [img_h][img_w][img_c], img_c 1 or 3 (RGB)
 *   pair_table: DEVICE int32 [n][4] = {id of frame t-1, id of frame t, OUTPUT id of the x image, OUTPUT id of the y image};
 *               output image `k` is written as u8 [H][W] at out_images + k*out_image_bytes, (H, W) = (img_h, img_w) or the
 *               resize target of `params` (the tool's frame resize, pinned to cv2.resize in the oracle's tests)
 *   params:     HOST struct, NULL = the defaults below
 *   flow_f32:   optional fp32 [n][2][img_h][img_w] (x then y displacement); iterations: optional int32
 *               [n][nscales_used*warps] inner iterations run per (level, warp) in processing order (coarsest level first)
 *   workspace:  DEVICE scratch of va_tvl1_workspace_bytes(img_h, img_w, params) bytes
 * Pyramid levels with W <= 704 and ceil(H/16) * W <= 5504 (340x256 and 320x240 frames: all levels) are solved in shared memory;
 * larger levels run the same formulas on L2-resident fields (slower per iteration, same results); H, W <= 4096.
 * --------------------------------------------------------------------------------------------------------- */
typedef struct va_tvl1_params {
  double tau;          /* 0.25 */
  double lambda;       /* 0.15 */
  double theta;        /* 0.3  */
  double epsilon;      /* 0.01 */
  double scale_step;   /* 0.8  */
  double bound;        /* 20: dense_flow / TSN --bound */
  int nscales;         /* 5 */
  int warps;           /* 5 */
  int iterations;      /* 300 */
  int resize_w;        /* 0 = flow at the frames' own size; else cv::resize(frame, Size(resize_w, resize_h), INTER_LINEAR) first, */
  int resize_h;        /*     as dense_flow does with TSN's new_size 340 x 256; outputs are then [resize_h][resize_w]            */
  int reserved;
} va_tvl1_params;
size_t va_tvl1_workspace_bytes(int img_h, int img_w, const va_tvl1_params* params);
va_status va_tvl1_flow(const uint8_t* images, size_t image_bytes, int img_h, int img_w, int img_c,
                       const int32_t* pair_table, int n, const va_tvl1_params* params, uint8_t* out_images,
                       size_t out_image_bytes, float* flow_f32, int32_t* iterations, void* workspace,
                       size_t workspace_bytes, va_stream_t stream);

/* Gradient all-reduce (sum) of the data-parallel training step as the library's own kernel over NVLink peer memory
 * (serves loss.backward()/optimizer.step() of spatialModel.py:178-181 on several GPUs; the reference itself is single-GPU).
 * In place over a SYMMETRIC bf16 buffer of n_elems elements (multiple of 8) that every rank of `world` holds:
 *   peer_ptrs:     HOST array of `world` DEVICE pointers, entry p = rank p's buffer as addressable from this GPU (P2P);
 *   multicast_ptr: the buffer's NVSwitch multicast address, or NULL -- when given, the kernel uses in-switch reduction
 *                  (multimem.ld_reduce, fp32 accumulation) and replicated stores (multimem.st) instead of W loads / stores.
 * Two-shot: rank r reduces slice r and writes it to all ranks, so replicas receive bit-identical sums.  The CALLER brackets
 * the launch with a cross-rank barrier with system-scope release/acquire (torch symmetric memory: handle.barrier()).
 * n_ctas: CTAs of 1024 threads to use (8 by default: it runs beside the layer kernels on the SMs va_reserve_sms frees). */
va_status va_allreduce_bf16(const void* const* peer_ptrs, void* multicast_ptr, int world, int rank, long long n_elems,
                            int n_ctas, va_stream_t stream);

/* SM reservation for a collective that runs beside the layer kernels (data-parallel training: the gradient all-reduce of
 * one stream under the other stream's forward pass).  The layer kernels are persistent, one CTA per SM with 210-227 KB of
 * shared memory, so a collective's CTAs cannot co-reside with them: launched beside a full grid they hold some SMs and the
 * layer's CTAs for those SMs start only when the collective ends.  After va_reserve_sms(sms, launches) the next `launches`
 * layer launches (va_forward / va_conv2d_nhwc / va_linear / dgrad calls of THIS process) size their grids for
 * sm_count - sms SMs; (0, 0) switches it off.  Results do not depend on the grid size. */
va_status va_reserve_sms(int sms, int launches);

/* Diagnostics: when set to a device buffer of 2 * nscales * warps int64, cluster 0 of every subsequent va_tvl1_flow launch
 * writes, for its first pair and per (level, warp) in processing order, the SM cycles of the bicubic warp phase [2k] and of
 * the inner iterations [2k+1].  NULL switches it off. */
va_status va_tvl1_debug_cycles(long long* dev_cycles);

/* Synthetic image store generator (bench/test data; integer hash identical to oracle/synth.py):
 * fills n_images images of image_bytes each, image id = first_id + i. */
va_status va_synth_fill(uint8_t* images, size_t image_bytes, int n_images, int img_h, int img_w, int img_c,
                        uint32_t seed, uint32_t first_id, va_stream_t stream);

/* Live profile of the tensor-core layer kernel (the dominant kernel): while enabled, every va_forward chunk records
 * a CUDA event pair on its stream around its 16 conv/FC launches (the fp32 head launch is outside the pair).
 * va_profile_read synchronises those events and returns, then clears, the accumulated device time (ms), the number
 * of layer-kernel launches and their ALGORITHMIC FLOPs (2*MAC over real, unpadded channels). */
va_status va_profile_enable(int on);
va_status va_profile_read(double* tensor_ms, uint64_t* tensor_launches, double* tensor_flops);

/* Diagnostics: when set to a device buffer of 16 int64, CTA 0 of every subsequent layer-kernel launch writes its
 * per-role cycle counters: [0] producer total, [1] producer stalled on free slots, [2] MMA total, [3] MMA stalled on
 * operands, [4] MMA stalled on a free accumulator, [5..7]/[8..10]... epilogue group 0/1 total, stalled on the
 * accumulator, stalled on the staging buffer; [11] = tiles processed by CTA 0.  NULL switches it off. */
va_status va_debug_conv_counters(long long* dev_counters16);

/* number of kernels this library has launched since load (bench.py reports it as gpu_launches) */
uint64_t va_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* VA_B200_H_ */
