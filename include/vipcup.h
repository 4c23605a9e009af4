/*
 * vipcup.h -- C ABI of libvipcup.so: the B200 (sm_100a) implementation of the vip-cup-2022 inference hot
 * path (preprocess -> backbone forward -> head / TTA / ensemble epilogue).
 *
 * The reference (awsaf49/vip-cup-2022) is pure Python/TensorFlow and has NO FFI of its own; the seams this
 * library replaces are Python-level.  Each entry point cites the reference code whose
 * arithmetic it takes over.  INTEGRATION.md shows the ctypes binding a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success or a negative VIP_ERR_* code; vip_last_error() returns a
 *     thread-local human-readable message for the last failure on the calling thread; nothing throws
 *     across the ABI.
 *   - "device" pointers are CUDA device pointers on the current device; "host" pointers are ordinary (ideally
 *     page-locked) host memory.  The caller owns every buffer it passes in; the library keeps no state between calls
 *     (no handles: weights, workspaces and CUDA graphs live in caller-owned device memory, orchestrated from Python).
 *   - all device work is enqueued on the caller's stream (cudaStream_t passed as void*; NULL = legacy default
 *     stream); no hidden synchronisation unless the function name ends in _host / _sync.
 *   - calls are thread-compatible: the only library state is the thread-local error text / launch counter.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails with VIP_ERR_CUDA.
 */
#ifndef VIPCUP_H_
#define VIPCUP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VIP_OK 0
#define VIP_ERR_INVALID (-1)     /* bad argument (null pointer, size out of range, unknown name) */
#define VIP_ERR_CUDA (-2)        /* CUDA runtime/driver error, message carries cudaGetErrorString */
#define VIP_ERR_UNSUPPORTED (-3) /* valid request outside what the kernels were built for */
#define VIP_ERR_STATE (-4)       /* call order violated (e.g. forward before finalize) */

#define VIP_DTYPE_F32 0
#define VIP_DTYPE_BF16 1

/* per-image flag bits for vip_preprocess (dataset/augment.py:115-120,142-146) */
#define VIP_FLAG_HFLIP 1u /* tf.image.flip_left_right */
#define VIP_FLAG_VFLIP 2u /* tf.image.flip_up_down    */
#define VIP_FLAG_GRAY 4u  /* rgb_to_grayscale -> grayscale_to_rgb */

/* library / build information ------------------------------------------------------------------ */
const char* vip_version(void);
const char* vip_last_error(void);
/* number of kernels launched by this library on the calling thread since the last reset (bench.py's
 * "gpu_launches" claim is read from here). */
int64_t vip_launch_count(void);
void vip_launch_count_reset(void);

/* -------------------------------------------------------------------------------------------------
 * Preprocessing.  Replaces, per image,
 *   dataset/dataset.py:31-37   tf.cast(float32) -> tf.image.resize(bicubic) -> / 255.0
 *   dataset/augment.py:110-113 tf.image.random_jpeg_quality  (quality decided by the caller)
 *   dataset/augment.py:115-120 flip_left_right / flip_up_down
 *   dataset/augment.py:142-146 rgb_to_grayscale -> grayscale_to_rgb
 * and the "slice then resize" random-crop semantic of
 *   models/keras_cv_attention_models/imagenet/data.py:56-63.
 * Order: crop -> bicubic(Hc x Wc -> Ho x Wo) -> /255 -> [JPEG round trip at quality q] -> flips -> gray.
 *
 *   src        device u8  [N, Hs, Ws, 3] NHWC (decoded RGB)
 *   crop_yxhw  device i32 [N, 4] (y0, x0, h, w) inside the source, or NULL for the full image
 *   jpeg_q     device i32 [N], q in [1,100] enables the libjpeg 4:2:0 baseline encode->decode emulation,
 *              q < 0 skips it; NULL skips it for every image
 *   flags      device u8  [N] VIP_FLAG_* bits, or NULL
 *   dst        device f32 or bf16 [N, Ho, Wo, 3] NHWC
 * Limits: 1 <= Ws <= 1024; Ho, Wo <= 256 when any image uses JPEG emulation (plane residency in shared
 * memory), otherwise Ho, Wo <= 1024.  Results are bit-identical to oracle/preprocess.py.
 */
int vip_preprocess(const uint8_t* src, int N, int Hs, int Ws, const int32_t* crop_yxhw, const int32_t* jpeg_q,
                   const uint8_t* flags, int Ho, int Wo, void* dst, int dst_dtype, void* cuda_stream);

/* Same operation on HOST buffers: copies inputs host->device, runs vip_preprocess in chunks on internal
 * streams so that copies overlap compute, copies the result device->host and synchronises before returning.
 * This is the call that stands where dataset.build_dataset (dataset/dataset.py:64-102) hands a batch to the
 * caller.  use_jpeg: 0 = jpeg_q ignored. */
int vip_preprocess_host(const uint8_t* src, int N, int Hs, int Ws, const int32_t* crop_yxhw, const int32_t* jpeg_q,
                        const uint8_t* flags, int Ho, int Wo, void* dst, int dst_dtype);

/* Raw bf16 contraction on the tcgen05 tensor cores (the building block of every Conv2D / Dense of the backbones:
 * models/resnet_rs/resnet_rs_model.py:64-84, models/gcvit/layers/attention.py:25,33, models/gcvit/layers/feature.py:20-22):
 *   out[M,N] = act(A[M,K] x B[N,K]^T + bias[N]) * colscale[N] + residual[M,N]
 * A, B, residual: device bf16, row-major (lda / ldb / ldr elements between rows, multiples of 8); bias, colscale
 * (GCViT layer-scale gamma, models/gcvit/layers/block.py:41-56,79-80): device f32 or NULL; act: 0 none, 1 relu,
 * 2 gelu, 3 sigmoid; out: device bf16 or f32 per out_dtype (ldc multiple of 8).  K % 8 == 0, N % 8 == 0. */
int vip_gemm_bf16(const void* A, int lda, const void* B, int ldb, int M, int N, int K, const float* bias, int act,
                  const float* colscale, const void* residual, int ldr, void* out, int ldc, int out_dtype,
                  void* cuda_stream);

/* Epilogue description of vip_gemm_bf16_ex / vip_conv2d_bf16, applied per output element (m, n) in this order:
 *   v = acc
 *   v = rstd[m] * (v - mean[m] * ln_colsum[n])   if ln_stats: LayerNormalization of the A rows folded into the contraction
 *        (models/gcvit/layers/block.py:28,39 feeding attention.py:25 / feature.py:20): ln_stats[m] = (mean, rstd) of row m,
 *        written by vip_row_stats_finalize / vip_layernorm_bf16 / vip_mlp_fused_bf16; B must hold gamma-scaled weights,
 *        bias must hold beta @ W + b
 *   v += bias[n];  v = act(v);  v *= colscale[n];  v += residual[m, n];  store as bf16 or f32
 *   row_stats[m] += (sum_n (out - p), sum_n (out - p)^2, p)   if row_stats: a row statistics record (below);
 *        vip_row_stats_finalize turns the records into the ln_stats of the next contraction
 *   gap[m / gap_rows, n] += out                   if gap: GlobalAveragePooling2D partial sums (SE squeeze,
 *        models/resnet_rs/resnet_rs_model.py:149), 36.28 fixed point
 * Statistics that cross kernels are accumulated with 64-bit INTEGER atomics on fixed-point values (value * 2^28), so the
 * totals do not depend on the order in which tiles finish: outputs are bit-reproducible and independent of the batch an
 * image is in.  A row statistics record is int64[3] = { sum (v - p), sum (v - p)^2 (both fixed point), bits of the f32
 * pivot p }; p comes from row_pivot (the mean the previous LayerNorm computed for the residual row; 0 when absent), which
 * keeps the one-pass variance free of cancellation when |mean| >> sigma.
 * Two-plane residual stream (residual_lo / out_lo, bf16): the running sum x of a pre-LN transformer block
 * (models/gcvit/layers/block.py:77-81) is carried as hi + lo; v = acc + bias + hi + lo, out = bf16(v), out_lo = bf16(v - out).
 * The hi plane is the A operand of the next contraction; needs a residual, a bf16 output and a bias-only epilogue.
 * residual_lo may be NULL (zeros) with out_lo set.  A low plane is an opaque buffer of ceil(M / 32) * 32 * N bf16 elements
 * that only these epilogues read and write, stored in blocks of 32 rows x 8 columns (512 contiguous bytes, row-major inside):
 * element (m, n) lies at ((m / 32) * (N / 8) + n / 8) * 256 + (m % 32) * 8 + n % 8 -- the order in which the epilogue
 * warps (one thread per row) touch it, so that their 16-byte accesses coalesce.
 * With row_gate the order is that of an SE bottleneck tail (resnet_rs_model.py:183,278-280):
 *   v = relu((acc + bias[n]) * row_gate[m / gate_rows, n] + residual[m, n])   (needs residual, act relu, bf16 output)
 * row_stats and gap are accumulated: the caller zeroes them (vip_memset_async). */
typedef struct vip_epilogue {
  const float* bias;      /* [N] or NULL */
  int act;                /* 0 none, 1 relu, 2 gelu (Keras' erf form evaluated as a fitted tanh: |err| <= 3e-4 |x|), 3 sigmoid,
                             4 swish (x * sigmoid(x)) */
  const float* colscale;  /* [N] or NULL */
  const void* residual;   /* bf16 [M, ldr] or NULL */
  int ldr;
  void* out;              /* bf16 or f32 [M, ldc] */
  int ldc;
  int out_dtype;          /* VIP_DTYPE_* */
  const float* ln_stats;  /* [M, 2] (mean, 1 / sigma) of the A rows, or NULL */
  const float* ln_colsum; /* [N] */
  int64_t* row_stats;     /* [M, 3] row statistics records or NULL */
  int64_t* gap;           /* [ceil(M / gap_rows), N] fixed point or NULL */
  int gap_rows;
  const float* row_gate;  /* f32 [ceil(M / gate_rows), N] or NULL: squeeze-excite gate of the image a row belongs to */
  int gate_rows;
  const void* residual_lo; /* low plane of the residual (blocked layout) or NULL */
  void* out_lo;            /* low plane of the output (blocked layout) or NULL */
  const float* row_pivot;  /* f32 [M, 2] or NULL: [m][0] = pivot p of row m's statistics record -- pass the ln_stats of the
                              LayerNorm that last read the residual row (its mean is close to the new row's mean) */
} vip_epilogue_t;

/* vip_gemm_bf16 with the full epilogue. N, K, lda, ldb, ldc, ldr multiples of 8. */
int vip_gemm_bf16_ex(const void* A, int lda, const void* B, int ldb, int M, int N, int K, const vip_epilogue_t* epi,
                     void* cuda_stream);
/* Conv2D with explicit symmetric zero padding as an implicit GEMM (no im2col matrix in memory): x bf16 [N,H,W,C] NHWC,
 * w bf16 [Cout, ksize*ksize*C] (K order r,s,c = Keras kernel (kh,kw,Cin,Cout) flattened, ldw elements between rows),
 * output rows = N*Ho*Wo pixels with Ho = (H + 2 pad - ksize) / stride + 1.  C, Cout % 8 == 0.
 * models/resnet_rs/resnet_rs_model.py:64-84, model_utils.py:22-46; gcvit layers/feature.py:97-98. */
int vip_conv2d_bf16(const void* x, int N, int H, int W, int C, const void* w, int ldw, int Cout, int ksize, int stride,
                    int pad, const vip_epilogue_t* epi, void* cuda_stream);
/* vip_conv2d_bf16 on a channel slice: x points at channel c0 of a wider NHWC tensor whose pixels are ldx channels apart; the
 * C channels [c0, c0 + C) are convolved.  One call per group makes a grouped convolution (Conv2D(groups=...):
 * keras_cv_attention_models/nfnets/nfnets.py:150-153, resnest/resnest.py:36-40); the output slice is addressed through
 * epi->out / epi->ldc in the same way. */
int vip_conv2d_slice_bf16(const void* x, int N, int H, int W, int C, int ldx, const void* w, int ldw, int Cout, int ksize,
                          int stride, int pad, const vip_epilogue_t* epi, void* cuda_stream);
/* cudaMemsetAsync on the caller's stream (zeroing of row_stats / gap accumulators). */
int vip_memset_async(void* ptr, int value, size_t bytes, void* cuda_stream);

/* ---- layer kernels of the backbones; activations are device bf16 NHWC / [tokens, C], C % 8 == 0 ------------------ */
/* Conv2D with explicit zero padding as an im2col matrix [N*Ho*Wo, Kp] (K order r,s,c = Keras kernel (kh,kw,Cin,Cout)
 * flattened): models/resnet_rs/resnet_rs_model.py:64-84, model_utils.py:22-46; gcvit embedding.py:15, feature.py:98. */
int vip_im2col_bf16(const void* x, int N, int H, int W, int C, int ksize, int stride, int pad, int Ho, int Wo, void* out,
                    int Kp, void* cuda_stream);
/* AveragePooling2D(2,2,'same'), divisor = valid count: models/resnet_rs/resnet_rs_model.py:207-212. out [N,ceil(H/2),ceil(W/2),C] */
int vip_avgpool2_same_bf16(const void* x, int N, int H, int W, int C, void* out, void* cuda_stream);
/* GlobalAveragePooling2D [N,HW,C] -> [N,C] as bf16 and/or f32 (either may be NULL): resnet_rs_model.py:149,469; gcvit.py:81; feature.py:55 */
int vip_global_avgpool_bf16(const void* x, int N, int HW, int C, void* out_bf16, float* out_f32, void* cuda_stream);
/* out = act(y * gate[n,c] + shortcut); gate f32 [N,C] or NULL, shortcut bf16 or NULL, act 0 none / 1 relu:
 * SE excite + Add + ReLU, resnet_rs_model.py:183,278-280; gcvit feature.py:66,109,150 */
int vip_scale_add_act_bf16(const void* y, const float* gate, const void* shortcut, void* out, int N, int HW, int C, int act,
                           void* cuda_stream);
/* Row statistics records (int64 [M,3], see vip_epilogue_t) -> out f32 [M,2] = (mean, 1 / sqrt(var + eps)) over `cols`
 * columns: the ln_stats operand of a contraction with a folded LayerNormalization. */
int vip_row_stats_finalize(const int64_t* records, long long M, int cols, float eps, float* out, void* cuda_stream);
/* LayerNormalization(axis=-1, epsilon) over [M,C]: gcvit block.py:28,39; feature.py:100-101; gcvit.py:79.  ln_next
 * (f32 [M,2] or NULL) receives (mean, 1 / sqrt(var + next_eps)) of every (rounded) OUTPUT row: the ln_stats of a LayerNorm
 * folded into the next contraction. */
int vip_layernorm_bf16(const void* x, const float* gamma, const float* beta, void* out, float* ln_next, float next_eps, long long M,
                       int C, float eps, void* cuda_stream);
/* ZeroPadding2D(1) + DepthwiseConv2D(3,'valid',no bias) (+ GELU if gelu != 0); w f32 [3,3,C]: gcvit feature.py:92-94,132-134.
 * gap (int64 [N,C] fixed point or NULL, zeroed by the caller) accumulates the per-image channel sums of the output: the
 * SE squeeze (GlobalAveragePooling, feature.py:55) without a second pass over the map. */
int vip_dwconv3x3_bf16(const void* x, const float* w, void* out, int64_t* gap, int N, int H, int W, int C, int gelu,
                       void* cuda_stream);
/* Depthwise K x K convolution, NHWC bf16, explicit zero padding (pad_top / pad_left; the bottom / right padding follows from
 * Ho / Wo), w f32 [K, K, C], bias f32 [C] or NULL, act 0 none / 1 swish / 2 gelu (Keras' erf form evaluated as the fitted tanh of vip_epilogue_t.act) / 3 relu, gap as for
 * vip_dwconv3x3_bf16.  Built: K 3 | 5 with stride 1 | 2, K 7 with stride 1.
 * ConvNeXt: models/tfimm/architectures/convnext.py:192-198 (ZeroPadding2D(3) + DepthwiseConv2D(7) + bias);
 * EfficientNet MBConv: keras_cv_attention_models/efficientnet/efficientnet_v2.py:80-96 (BN folded by the caller). */
int vip_dwconv_bf16(const void* x, const float* w, const float* bias, void* out, int64_t* gap, int N, int H, int W, int C,
                    int ksize, int stride, int pad_top, int pad_left, int Ho, int Wo, int act, void* cuda_stream);
/* ResNeSt split attention, radix 2 (keras_cv_attention_models/resnest/resnest.py:16-24,57-62): x bf16 [N, HW, 2F] (the two
 * radix splits side by side), logits f32 [N, 2F]; a = softmax over the radix per channel;
 * out[n, p, c] = a0[n, c] x[n, p, c] + a1[n, c] x[n, p, F + c], bf16 [N, HW, F]. */
int vip_split_attention2_bf16(const void* x, const float* logits, void* out, int N, int HW, int F, void* cuda_stream);
/* ZeroPadding2D(1) + AveragePooling2D(3, strides 2): the padded zeros count (divisor 9), resnest.py:63-65. */
int vip_avgpool3s2_bf16(const void* x, void* out, int N, int H, int W, int C, void* cuda_stream);
/* out = act(x) * scale over n bf16 elements (n % 8 == 0); act 0 none, 1 relu, 4 swish: the NFNet pre-activation
 * swish(x) * beta (nfnets/nfnets.py:138). */
int vip_act_scale_bf16(const void* x, void* out, long long n, int act, float scale, void* cuda_stream);
/* Efficient Channel Attention gate (common_layers.py:335-353): gate[n, c] = out_scale * sigmoid(sum_k w[k] * mean[n, c + k - ksize/2])
 * with mean = fixed-point pooled sums (gap) * inv_hw and zeros outside [0, C). */
int vip_eca_gate_f32(const int64_t* gap, const float* w, float* gate, int N, int C, int ksize, float inv_hw, float out_scale,
                     void* cuda_stream);
/* LayerNormalization of pooled f32 vectors [M, C] -> f32 (ConvNeXt head, convnext.py:432-436). */
int vip_layernorm_f32(const float* x, const float* gamma, const float* beta, float* out, int M, int C, float eps,
                      void* cuda_stream);
/* ZeroPadding2D(1) + MaxPool2D(3,2,'valid') (padded zeros take part in the max): gcvit feature.py:139,151-152 */
int vip_maxpool3s2_bf16(const void* x, void* out, int N, int H, int W, int C, void* cuda_stream);
/* Window attention, head_dim 32, window partition/reverse folded into addressing: gcvit attention.py:52-83, window.py:3-14.
 * qkv bf16 [B*H*W, 3C] (local) or [B*H*W, 2C] = k,v (global, q_global bf16 [B, ws*ws, C] != NULL); rel_table f32
 * [heads, (2ws-1)^2] = relative_position_bias_table transposed, gathered inside the kernel with the index of
 * attention.py:39-50; out bf16 [B*H*W, C].  ws 7 and 14 are built. */
int vip_window_attention_bf16(const void* qkv, const void* q_global, const float* rel_table, void* out, int B, int H, int W,
                              int C, int ws, int heads, void* cuda_stream);
/* FitWindow padding and the crop after a level's blocks (gcvit feature.py:234-256, level.py:49,61):
 * out[n, y, x, :] = x[n, y - top, x - left, :] where that source pixel exists, zero elsewhere; x is [N,H,W,bytes_per_pixel]
 * (bytes_per_pixel % 8 == 0: bf16 activations with C % 4 == 0, or int64[3] row statistics records), out [N,Ho,Wo,...]. */
int vip_pad_crop(const void* x, int N, int H, int W, int bytes_per_pixel, void* out, int Ho, int Wo, int top, int left,
                 void* cuda_stream);
/* Classifier head on pooled f32 features [N,C]: Dense(k) (w f32 [C,k], b [k]) + softmax (sigmoid_head = 0) or sigmoid,
 * probs f32 [N,k]; when acc != NULL also acc[n] += acc_weight * P(synthetic) with P = k > 1 ? 1 - probs[n,0] : probs[n,0]
 * (TTA / fold / ensemble means of main.py:110-121,142 as one fused accumulation): resnet_rs_model.py:474-476, gcvit.py:88 */
int vip_head_f32(const float* feat, const float* w, const float* b, float* probs, double* acc, double acc_weight, int N,
                 int C, int k, int sigmoid_head, void* cuda_stream);
int vip_cast_f32_bf16(const float* x, void* out, long long n, void* cuda_stream);
/* vip_gemm_bf16_ex with one weight matrix per group of rows: rows [i * rows_per_group, (i + 1) * rows_per_group) of A are
 * contracted with B[i * N : (i + 1) * N, :] (B is [ceil(M / rows_per_group) * N, ldb]).  rows_per_group must be a multiple
 * of 128.  Used to fold a per-image input-channel scale (the SE gate of an MBConv block, feature.py:144-150) into the
 * weights of the 1x1 convolution that follows, built by vip_scale_weights_bf16. */
int vip_gemm_grouped_bf16(const void* A, int lda, const void* B, int ldb, int M, int N, int K, int rows_per_group,
                          const vip_epilogue_t* epilogue, void* cuda_stream);
/* out[g][n][k] = bf16(w[n][k] * gate[g][k]); w bf16 [N, ldw], gate fp32 [G, K], out bf16 [G * N, K] */
int vip_scale_weights_bf16(const void* w, int ldw, const float* gate, int G, int N, int K, void* out, void* cuda_stream);
/* Fused pre-LN MLP of a GCViT block (models/gcvit/layers/block.py:39-56,77-81; layers/feature.py:8-43):
 *   out[m, :] = x[m, :] + W2 gelu(W1 LayerNorm(x[m, :]) + b1) + b2     with LayerNorm folded like vip_epilogue_t.ln_stats:
 *   ln_stats f32 [M, 2] = (mean, 1 / sigma) of the rows of x, w1 bf16 [hidden, ldw1] holds gamma-scaled weights,
 *   colsum1 [hidden] their column sums, bias1 [hidden] = beta W1 + b1; w2 bf16 [C, ldw2], bias2 [C] (layer scale folded).
 *   x, out bf16 [M, C] contiguous; ln_next f32 [M, 2] (or NULL) receives (mean, 1 / sqrt(var + next_eps)) of the rows of
 *   out; x_lo / out_lo (or NULL): low planes of the two-plane residual stream (see vip_epilogue_t).
 * The hidden activations stay in TMEM / shared memory.  Built for (C, hidden) = (96, 192) and (64, 192); any other shape
 * returns VIP_ERR_UNSUPPORTED and the caller issues two vip_gemm_bf16_ex calls instead. */
int vip_mlp_fused_bf16(const void* x, const void* x_lo, long long M, int C, int hidden, const float* ln_stats, float next_eps,
                       const void* w1, int ldw1, const float* colsum1, const float* bias1, const void* w2, int ldw2,
                       const float* bias2, void* out, void* out_lo, float* ln_next, void* cuda_stream);
/* out = bf16(x * 2^-28 * scale): fixed-point pooled sums of the fused gap epilogue -> means (SE squeeze,
 * resnet_rs_model.py:149) */
int vip_scale_cast_fx_bf16(const int64_t* x, float scale, void* out, long long n, void* cuda_stream);

/* -------------------------------------------------------------------------------------------------
 * JPEG decode on the device.  Replaces tf.io.read_file -> tf.image.decode_jpeg(channels=3) of
 * dataset/dataset.py:24-28 (libjpeg-turbo defaults: JDCT_ISLOW, fancy chroma upsampling) for baseline
 * sequential Huffman files; results are bit-identical to libjpeg-turbo's (tests compare with Pillow).
 *
 * Flow: the caller reads the files, vip_jpeg_parse (HOST, no GPU) fills one descriptor per file, vip_jpeg_plan
 * assigns every image its slice of the output / coefficient buffers, the caller copies the concatenated file
 * bytes and the descriptors to the device, and vip_jpeg_decode runs two kernels: entropy decode (one warp per
 * image; lane 0 walks the Huffman stream, the warp stores each 8x8 block of coefficients with one coalesced
 * write) and dequantise -> integer IDCT -> fancy upsample -> YCbCr->RGB (8 lanes per block).
 * Files the kernels do not cover (progressive, arithmetic, 12-bit, CMYK / RGB colour spaces, sampling other than
 * 4:4:4 / 4:2:2 / 4:2:0 / grey, multi-scan) get status != 0 from the parser; the caller decodes those on the host
 * (libjpeg through Pillow, i.e. what the reference's CPU path does) into the same output slice. */
#define VIP_JPEG_OK 0
#define VIP_JPEG_NOT_JPEG 1      /* no SOI / truncated header */
#define VIP_JPEG_UNSUPPORTED 2   /* valid JPEG outside the subset above */
typedef struct vip_jpeg_desc {
  int32_t status;            /* VIP_JPEG_* (parser) */
  int32_t width, height;     /* image size; valid whenever a frame header was read, even if status != 0 */
  int32_t ncomp;             /* 1 (grey) or 3 (YCbCr) */
  int32_t hs[3], vs[3];      /* sampling factors per component */
  int32_t tq[3], td[3], ta[3]; /* quantisation / DC Huffman / AC Huffman table selectors per component */
  int32_t restart_interval;  /* MCUs between RSTn markers, 0 = none */
  int32_t scan_offset;       /* first byte of the entropy-coded segment inside the file */
  int32_t scan_bytes;        /* its length (up to the marker that ends the scan) */
  int64_t file_offset;       /* CALLER: offset of this file's first byte in the device byte buffer */
  int64_t dst_offset;        /* vip_jpeg_plan: byte offset of this image's [height, width, 3] u8 pixels in dst */
  int64_t coef_offset;       /* vip_jpeg_plan: first 8x8 block of this image in the coefficient workspace */
  uint16_t qt[4][64];        /* quantisation tables, natural (row-major) order */
  uint8_t huff_bits[4][16];  /* code-length counts: tables 0,1 = DC 0,1; 2,3 = AC 0,1 */
  uint8_t huff_vals[4][256]; /* symbols in code order */
} vip_jpeg_desc;
/* HOST: parse the headers of one JPEG file (no pixel work).  Returns VIP_OK and sets desc->status; a file that is not a
 * decodable JPEG is not an error of this call. */
int vip_jpeg_parse(const uint8_t* file, size_t len, vip_jpeg_desc* desc);
/* HOST: assign dst_offset / coef_offset of N descriptors (dense, in order; every image with a known size gets an output
 * slice, decodable ones also a coefficient slice).  Writes the total bytes of dst and the number of int16[64] blocks. */
int vip_jpeg_plan(vip_jpeg_desc* descs, int N, int64_t* dst_bytes, int64_t* coef_blocks);
/* DEVICE: decode the images with status == 0.  data = device byte buffer holding the files; descs_host / descs_dev = the
 * same N descriptors in host and device memory; coef = device workspace of coef_blocks * 64 int16; dst = device u8
 * buffer of dst_bytes (interleaved RGB, rows tightly packed); err = device int32 [N] or NULL, set to 1 for an image whose
 * entropy-coded data is corrupt (ran out of data / invalid code), 0 otherwise. */
int vip_jpeg_decode(const uint8_t* data, const vip_jpeg_desc* descs_host, const vip_jpeg_desc* descs_dev, int N,
                    int16_t* coef, uint8_t* dst, int32_t* err, void* cuda_stream);

/* Exhaustive on-device self check of the exact x/255 sequence used by the kernels against IEEE division
 * (dataset/dataset.py:37).  Writes the number of mismatching bit patterns to *mismatches. */
int vip_selftest_div255(uint64_t* mismatches, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* VIPCUP_H_ */
