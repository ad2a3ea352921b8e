/* kgdet_b200 -- C ABI of the B200-native KGDet point-set head operators.
 *
 * One shared library (kgdet_b200/_lib/libkgdet_b200.so), plain pointers and
 * sizes, no torch / ATen types.  Every entry point below replaces one pybind
 * entry point of the reference (cited as file:line under
 * mmdetection/mmdet/ops/); the Python binding that calls it is
 * kgdet_b200/ops/_capi.py (ctypes), and INTEGRATION.md shows the stub a
 * reference maintainer would add.
 *
 * Conventions
 *  - All data pointers are DEVICE pointers unless the name ends in _host.
 *  - Every call is asynchronous on `stream` (a cudaStream_t passed as void*);
 *    the reference launches on the legacy default stream
 *    (dcn/src/deform_conv_cuda_kernel.cu:264) and, for NMS, blocks on a D2H
 *    copy (nms/src/nms_kernel.cu:100) -- this library never synchronises.
 *  - No allocation inside: scratch is passed in as (workspace, workspace_bytes)
 *    and sized by the matching *_workspace_bytes query (the reference allocates
 *    `columns` / `output_buffer` / the NMS mask itself:
 *    dcn/src/deform_conv_cuda.cpp:196-214, nms/src/nms_kernel.cu:89).
 *  - Return value: 0 = ok, negative = error; kgdet_last_error() returns a
 *    thread-local message (the reference raises through AT_CHECK/AT_ERROR:
 *    dcn/src/deform_conv_cuda.cpp:61-149).
 *  - Tensors use the reference's public layouts: activations NCHW contiguous,
 *    offsets [N, dg*2*K, Ho, Wo] with (dy, dx) interleaved per tap
 *    (dcn/src/deform_conv_cuda_kernel.cu:221-224), weights [Cout, Cin/g, kh, kw].
 */
#ifndef KGDET_B200_H_
#define KGDET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KGDET_ABI_VERSION 2

#if defined(__GNUC__)
#define KGDET_API __attribute__((visibility("default")))
#else
#define KGDET_API
#endif

enum {
  KGDET_OK = 0,
  KGDET_ERR_INVALID_ARG = -1,
  KGDET_ERR_CUDA = -2,
  KGDET_ERR_WORKSPACE = -3,
  KGDET_ERR_UNSUPPORTED = -4
};

/* storage type of activations / gradients at the boundary */
enum { KGDET_F32 = 0, KGDET_BF16 = 1 };

/* arithmetic of the dense contraction
 *  FP32   : fp32 FFMA on the SIMT pipe (bit pattern of the sum order aside, the reference's SGEMM)
 *  TF32X3 : tcgen05 kind::tf32 with hi/lo split operands, 3 MMAs per k-step, and accumulator promotion (no TMEM
 *           accumulator absorbs more than 16 k-blocks; chunk results are summed in fp32 round-to-nearest):
 *           fp32-grade, measured < 1e-5 of the tensor maximum against the fp64 reference at C = 256, K = 49.
 *           Default for fp32 tensors on shapes the tensor-core path supports.
 *  BF16   : tcgen05 kind::f16 with bf16 operands, fp32 accumulation in TMEM (~1e-3)
 *  TF32   : tcgen05 kind::tf32, single pass (~1e-3) */
enum { KGDET_PREC_FP32 = 0, KGDET_PREC_TF32X3 = 1, KGDET_PREC_BF16 = 2, KGDET_PREC_TF32 = 3 };

/* output layouts of kgdet_dcn_forward_prepared */
enum { KGDET_LAYOUT_NCHW = 0, KGDET_LAYOUT_TILED = 1, KGDET_LAYOUT_TILED_SPLIT = 2 };

/* NMS comparator: the reference's two back-ends disagree at IoU == thr */
enum { KGDET_NMS_GT = 0 /* nms_kernel.cu:60 */, KGDET_NMS_GE = 1 /* nms_cpu.cpp:55 */ };

typedef struct kgdet_dcn_shape {
  int32_t N, C, H, W;        /* input  [N, C, H, W]                                   */
  int32_t Cout, kh, kw;      /* weight [Cout, C/groups, kh, kw]                       */
  int32_t stride_h, stride_w, pad_h, pad_w, dil_h, dil_w;
  int32_t groups, deformable_groups;
} kgdet_dcn_shape;

KGDET_API const char* kgdet_last_error(void);
KGDET_API int kgdet_abi_version(void);
/* number of kernels this library has launched in this process (a launch inside a CUDA-graph capture counts
 * once, at capture) -- bench.py reports it as gpu_launches */
KGDET_API uint64_t kgdet_launch_count(void);
/* 1 when the fused tcgen05 path supports (shape, precision); 0 -> the exact SIMT path runs */
KGDET_API int kgdet_dcn_fast_path_supported(const kgdet_dcn_shape* shape, int precision);

/* ---- deformable convolution --------------------------------------------------------
 * replaces deform_conv_forward_cuda            dcn/src/deform_conv_cuda.cpp:151-258
 *          modulated_deform_conv_cuda_forward  dcn/src/deform_conv_cuda.cpp:486-564
 * `mask` ([N, dg*K, Ho, Wo]) and `bias` ([Cout]) may be NULL (plain DeformConv).
 * `weight_packed` comes from kgdet_dcn_pack_weight for the same (shape, precision). */
KGDET_API size_t kgdet_dcn_packed_weight_bytes(const kgdet_dcn_shape* shape, int precision);
KGDET_API int kgdet_dcn_pack_weight(const float* weight, void* weight_packed, const kgdet_dcn_shape* shape,
                          int precision, void* stream);
KGDET_API size_t kgdet_dcn_forward_workspace_bytes(const kgdet_dcn_shape* shape, int dtype, int precision);
KGDET_API int kgdet_dcn_forward(const void* input, const float* offset, const float* mask,
                      const void* weight_packed, const float* bias, void* output,
                      const kgdet_dcn_shape* shape, int dtype, int precision, void* workspace,
                      size_t workspace_bytes, void* stream);

/* Prepared form of the forward pass (inference): the NHWC copy of an input and the sample plan of an
 * offset tensor are built once and shared by several deformable convolutions -- the six DCNs of a
 * KGDet Kp3RepBlock stage read 2 inputs and 3 offset tensors (KP3:145-163).  The epilogue can apply the
 * ReLU that follows every DCN in the head (KP3:145-150) and write straight into a channel slice
 * [out_channel_offset, out_channel_offset + Cout) of a wider NCHW tensor with out_channels_total
 * channels, which removes the torch.cat (KP3:151-153).  `prepared_input` / `plan` are opaque buffers
 * sized by the *_bytes queries; they depend on (shape, dtype, precision) only through N, C, H, W,
 * the kernel size / stride / padding / dilation and the path selected by `precision`. */
KGDET_API size_t kgdet_dcn_prepared_input_bytes(const kgdet_dcn_shape* shape, int precision);
KGDET_API int kgdet_dcn_prepare_input(const void* input, void* prepared_input, const kgdet_dcn_shape* shape,
                            int dtype, int precision, void* stream);
/* same from position-major fp32 rows [N*H*W, C] (a channels_last activation): no transpose */
KGDET_API int kgdet_dcn_prepare_input_rows(const float* rows, void* prepared_input, const kgdet_dcn_shape* shape,
                                 int precision, void* stream);
KGDET_API size_t kgdet_dcn_plan_bytes(const kgdet_dcn_shape* shape, int precision);
KGDET_API int kgdet_dcn_prepare_plan(const float* offset, const float* mask, void* plan,
                           const kgdet_dcn_shape* shape, int precision, void* stream);
/* Plan from a channel slice [channel_offset, channel_offset + 2K) of a point-set tensor
 * [N, channels_total, Ho, Wo] whose values are absolute point offsets: the head's `pts - dcn_base_offset`
 * (KP3:37-67,135-143) happens inside the plan kernel, so neither the slice nor the subtraction is a kernel.
 * gradient_mul != 0: the head's `gradient_mul * pts + (1 - gradient_mul) * pts.detach()` (KP3:135-143, the identity
 * up to fp32 rounding; the reference evaluates it in inference too) is applied first, in that operation order, with
 * the two fp32 factors given, so that the sampled locations are bit-identical to the reference's.
 * deformable_groups must be 1. */
KGDET_API int kgdet_dcn_prepare_plan_points(const float* points, int32_t channel_offset, int32_t channels_total,
                                  float gradient_mul, float one_minus_gradient_mul, void* plan,
                                  const kgdet_dcn_shape* shape, int precision, void* stream);
/* out_layout: KGDET_LAYOUT_NCHW, or (tensor-core path, dtype KGDET_BF16) KGDET_LAYOUT_TILED = position-major
 * rows [N*Ho*Wo, out_channels_total] stored as the UMMA-tiled A operand of kgdet_pointwise_conv_tiled
 * (kgdet_pointwise_tiled_bytes(M, out_channels_total, 0) bytes), or KGDET_LAYOUT_TILED_SPLIT = the same with
 * the bf16 hi parts followed by the lo parts (x - hi) for the split-precision GEMM (..._bytes(M, K, 1)). */
KGDET_API int kgdet_dcn_forward_prepared(const void* prepared_input, const void* plan, const void* weight_packed,
                               const float* bias, void* output, int32_t out_channel_offset,
                               int32_t out_channels_total, int fuse_relu, int out_layout,
                               const kgdet_dcn_shape* shape, int dtype, int precision, void* workspace,
                               size_t workspace_bytes, void* stream);
/* scratch of the prepared call (k-block split of small maps, TF32X3 accumulator promotion); 0 = none needed.
 * `workspace` may be NULL: the call then runs unsplit (TF32X3 without promotion: ~1e-4 instead of 1e-5). */
KGDET_API size_t kgdet_dcn_forward_prepared_workspace_bytes(const kgdet_dcn_shape* shape, int precision);

/* Grouped form of kgdet_dcn_forward_prepared (bf16 mode): up to KGDET_DCN_GROUP_MAX deformable convolutions with the
 * same Cout in ONE persistent launch -- the six DCNs of a Kp3RepBlock stage (KP3:145-163).  The tiles of all items
 * are handed out longest first from an atomic counter to one CTA per SM: no SM idles while another item still has
 * tiles, and launch / TMEM allocation / barrier set-up are paid once.  Results are bit-identical to the single calls.
 * `workspace`: >= 16 bytes, 16-byte aligned (the tile counter; zeroed by the call). */
#define KGDET_DCN_GROUP_MAX 6
typedef struct kgdet_dcn_group_item {
  const void* prepared_input;   /* kgdet_dcn_prepare_input(_rows) / the hi half of split planes */
  const void* plan;             /* kgdet_dcn_prepare_plan(_points) */
  const void* weight_packed;    /* kgdet_dcn_pack_weight */
  const float* bias;            /* [Cout] or NULL */
  void* output;
  int32_t out_channel_offset, out_channels_total, fuse_relu, out_layout, dtype;
  kgdet_dcn_shape shape;
} kgdet_dcn_group_item;
KGDET_API int kgdet_dcn_group_supported(const kgdet_dcn_shape* shape, int precision);
KGDET_API int kgdet_dcn_forward_prepared_group(const kgdet_dcn_group_item* items, int32_t count, int precision,
                                     void* workspace, size_t workspace_bytes, void* stream);
/* measurement hook: the next grouped call records these two CUDA events immediately around its kernel */
KGDET_API void kgdet_dcn_group_set_profile_events(void* start_event, void* stop_event);

/* Global top-k over the survivors of the batched NMS (mmdet/core/post_processing/bbox_nms_kp.py:64-70:
 * sort the concatenated per-class results by score, keep max_num).  dets [B, L, 5] (score in column 4), flags
 * [B*L] uint8 (1 = kept by kgdet_nms_batched), L = classes * candidates <= 16384.  top_s [B, k] scores in
 * descending order (ties: lower index first; -1 for empty slots), top_i [B, k] int64 index into L (0 for empty
 * slots).  One CTA per image, no host synchronisation. */
KGDET_API int kgdet_topk_flagged(const float* dets, const uint8_t* flags, int32_t B, int32_t L, int32_t k,
                       float* top_s, int64_t* top_i, void* stream);

/* ---- post-head decode around the batched NMS --------------------------
 * replaces the PyTorch glue of get_bboxes_single (KP3:843-903) and multiclass_nms_kp
 * (core/post_processing/bbox_nms_kp.py:6-75) for ONE head level, batched over images, static shapes:
 *   select   order[b, r] = position of the r-th largest max-over-classes score (topk(nms_pre), KP3:863-874;
 *            ties by ascending position; identity when n == HW).  scores: [B, C, HW] logits
 *            (apply_sigmoid = 1) or probabilities.  Ranks by counting over all positions (HW <= 16384); with
 *            kgdet_bbox_select_workspace_bytes(B, HW, n) bytes of scratch (non-zero for HW > 4096, n <= 4096),
 *            kgdet_bbox_select_ws takes the large-level path (the FPN levels of
 *            reppoints_head_kp_parallel.py:703-713; HW <= 40960): three-pass radix select of the n-th key +
 *            ranks inside the n selected -- the same order, bit for bit.
 *   decode   boxes [B, n, 4] = clamp(bbox * stride + centre) (KP3:875-886) and the dense NMS input
 *            dets [B, C, n, 5] for kgdet_nms_batched (one segment per (image, class)).  img_wh: [B, 2].
 *   finalize for top_i [B, k] (= class * n + candidate, from a top-k over the NMS-masked scores, top_s <= 0 =
 *            empty slot): out_dets [B, k, 5], out_labels [B, k] (-1 = empty), out_kpts [B, k, num_keypts * 3]
 *            = (x, y, 1) decoded from keypts [B, 2 * num_keypts, HW] (y-first pairs) only for the survivors. */
KGDET_API int kgdet_bbox_select(const float* scores, int apply_sigmoid, int32_t B, int32_t C, int32_t HW, int32_t n,
                      int32_t* order, void* stream);
KGDET_API size_t kgdet_bbox_select_workspace_bytes(int32_t B, int32_t HW, int32_t n);
KGDET_API int kgdet_bbox_select_ws(const float* scores, int apply_sigmoid, int32_t B, int32_t C, int32_t HW, int32_t n,
                         int32_t* order, void* workspace, size_t workspace_bytes, void* stream);
KGDET_API int kgdet_bbox_decode(const float* scores, int apply_sigmoid, const float* bbox, const int32_t* order,
                      const float* img_wh, float stride, int32_t map_w, int32_t B, int32_t C, int32_t HW,
                      int32_t n, float* boxes, float* dets, void* stream);
KGDET_API int kgdet_bbox_finalize(const float* boxes, const float* keypts, const int32_t* order, const int64_t* top_i,
                        const float* top_s, const float* img_wh, float stride, int32_t map_w, int32_t B,
                        int32_t HW, int32_t n, int32_t k, int32_t num_keypts, float* out_dets,
                        int64_t* out_labels, float* out_kpts, void* stream);

/* ---- plain k x k convolutions of the head towers on the tensor cores -------
 * replaces  the cuDNN calls behind ConvModule's nn.Conv2d (mmdet/models/utils/conv_module.py:96-110,156-164;
 *           KP3:292-313 towers, KP3:98-106 stage-1 3x3 convolutions): stride 1, padding (k-1)/2, no groups.
 * fp32-grade arithmetic ("bf16x3": operands split into bf16 hi + lo, three tcgen05 MMAs per k-step, fp32
 * accumulation): ~1e-5 of the tensor maximum against an fp64 convolution -- cuDNN's TF32 path is 8e-4, which the
 * deformable stages of the head amplify to 1e-1.
 * The activation is read as SPLIT PLANES: channel-blocked bf16 planes [C/64][guard | N*H*W pixels | guard][64] of
 * the hi parts -- bit-identical to kgdet_dcn_prepare_input(..., KGDET_PREC_BF16), so the hi half can be handed to
 * kgdet_dcn_forward_prepared as its prepared input -- followed by the same planes of the lo parts (x - hi).
 * Output: NHWC fp32 [N, H, W, Cout] (a channels_last tensor), optional bias and ReLU.
 * Requires C % 64 == 0, Cout in {64, 128, 192, 256}, odd k <= 7 (kgdet_conv_supported). */
KGDET_API int kgdet_conv_supported(int32_t C, int32_t Cout, int32_t ksize);
KGDET_API size_t kgdet_conv_split_planes_bytes(int32_t N, int32_t C, int32_t H, int32_t W);
KGDET_API int kgdet_conv_split_planes_from_nchw(const float* x, void* planes, int32_t N, int32_t C, int32_t H, int32_t W,
                                      void* stream);
KGDET_API int kgdet_conv_split_planes_from_rows(const float* rows_nhwc, void* planes, int32_t N, int32_t C, int32_t H,
                                      int32_t W, void* stream);
KGDET_API size_t kgdet_conv_packed_weight_bytes(int32_t Cout, int32_t Cin, int32_t ksize);
KGDET_API int kgdet_conv_pack_weight(const float* weight, void* weight_packed, int32_t Cout, int32_t Cin, int32_t ksize,
                           void* stream);
KGDET_API int kgdet_conv_forward(const void* planes, const void* weight_packed, const float* bias, float* out_nhwc,
                       int32_t N, int32_t C, int32_t H, int32_t W, int32_t Cout, int32_t ksize, int fuse_relu,
                       void* stream);
/* Two convolutions of the same geometry (different planes, weights, outputs) in ONE launch: the classification and
 * the point tower's convolution of one layer (KP3:415-420) -- a CTA runs its tile of the first, then of the second
 * into the other half of TMEM, so the first's epilogue overlaps the second's main loop. */
KGDET_API int kgdet_conv_forward_pair(const void* planes0, const void* weight_packed0, const float* bias0, float* out0,
                            const void* planes1, const void* weight_packed1, const float* bias1, float* out1,
                            int32_t N, int32_t C, int32_t H, int32_t W, int32_t Cout, int32_t ksize, int fuse_relu,
                            void* stream);
/* GroupNorm (+ ReLU) of an NHWC fp32 activation (torch.nn.GroupNorm semantics) written as split planes for the next
 * convolution / the deformable stage, and optionally (y != NULL) also as NHWC fp32. */
KGDET_API int kgdet_groupnorm_relu_nhwc_planes(const float* x, const float* gamma, const float* beta, float eps,
                                     int32_t groups, int fuse_relu, float* y, void* planes, int32_t N, int32_t H,
                                     int32_t W, int32_t C, void* stream);

/* Backward of kgdet_groupnorm_relu_nhwc for the training step of the towers (maps of at most 1600 positions):
 * dx (NHWC fp32) and per-image partials dgamma_part / dbeta_part [N, C] (their sum over N is the gradient).
 * replaces the autograd of torch.nn.GroupNorm + ReLU inside ConvModule (conv_module.py:96-110,156-164). */
KGDET_API int kgdet_groupnorm_relu_nhwc_backward(const float* x, const float* dy, const float* gamma, const float* beta,
                                       float eps, int32_t groups, int fuse_relu, float* dx, float* dgamma_part,
                                       float* dbeta_part, int32_t N, int32_t HW, int32_t C, void* stream);

/* GroupNorm (+ ReLU) of NHWC fp32 maps of ANY size (the resident kernels above need the map in shared memory:
 * <= 1600 positions): two streaming passes (per-chunk moments merged with Chan's update in chunk order, then
 * normalise).  y (NHWC fp32) and / or planes may be NULL (not both); planes_hi_only != 0 writes only the hi half
 * (= the fused deformable convolution's prepared input, kgdet_dcn_prepared_input_bytes).
 * replaces  ConvModule's norm + activation   mmdet/models/utils/conv_module.py:96-110,156-164 on the FPN levels of
 *           reppoints_head_kp_parallel.py:115-145 / reppoints_head_kp_serial.py:115-145 */
KGDET_API size_t kgdet_groupnorm_stream_workspace_bytes(int32_t N, int32_t HW, int32_t C, int32_t groups);
KGDET_API int kgdet_groupnorm_relu_nhwc_stream(const float* x, const float* gamma, const float* beta, float eps,
                                     int32_t groups, int fuse_relu, float* y, void* planes, int planes_hi_only,
                                     int32_t N, int32_t H, int32_t W, int32_t C, void* workspace,
                                     size_t workspace_bytes, void* stream);

/* ---- target assignment + the nine training losses of the KGDet head ----------
 * replaces  PointAssigner.assign    mmdet/core/bbox/assigners/point_assigner.py:23-116
 *           point_target_kp         mmdet/core/anchor/point_target_kp.py:7-169  (sampling=False, every point valid)
 *           KP3.loss / loss_single  reppoints_head_kp3rep_cas_1_assign_once.py:581-768
 * for ONE point level (point_strides=[32] of the KGDet configs) of map_h x map_w points (<= 4096), batched over
 * images, without host synchronisation.  Ground truth is padded to G boxes per image: gt_boxes [B, G, 4] fp32,
 * gt_valid [B, G] uint8, gt_labels [B, G] int64 (1-based classes), gt_keypoints [B, G, K, 3] (x, y, visibility).
 *   assign    assigned [B, P] int32 = 0 (background) or g + 1; avg_factor (device float) = sum over images of
 *             max(#positives, 1); num_visible [B, G] = visible keypoints per box.
 *   forward   losses[9] = loss_cls_1..3, loss_bbox_1..3, loss_kpt_1..3 (loss weight and 1 / avg_factor applied)
 *             from the nine head outputs outs[9] = cls_1..3 [B, NC, H, W], kpt_1..3 [B, 2K, H, W] (y-first pairs),
 *             bbox_1..3 [B, 4, H, W], all NCHW fp32, read in place.  loss_weights[9] in the same order as losses.
 *   backward  grad_outs[i] (same shapes as outs[i]; NULL = skip) = d(sum_k grad_losses[k] * losses[k]) / d outs[i]. */
KGDET_API size_t kgdet_point_assign_scratch_bytes(int32_t B, int32_t map_h, int32_t map_w);
KGDET_API int kgdet_point_assign(const float* gt_boxes, const uint8_t* gt_valid, const float* gt_keypoints, int32_t B,
                       int32_t G, int32_t num_keypoints, int32_t map_h, int32_t map_w, float stride, int32_t pos_num,
                       int32_t* assigned, float* avg_factor, float* num_visible,
                       void* scratch /* kgdet_point_assign_scratch_bytes, 16-byte aligned; initialised by the call */,
                       void* stream);
KGDET_API int kgdet_point_losses_forward(const float* const* outs, const int32_t* assigned, const float* gt_boxes,
                               const int64_t* gt_labels, const float* gt_keypoints, const float* avg_factor,
                               const float* num_visible, int32_t B, int32_t G, int32_t map_h, int32_t map_w,
                               int32_t num_classes, int32_t num_keypoints, float stride, float point_base_scale,
                               const float* loss_weights, float gamma, float alpha, float beta, float* losses,
                               void* stream);
KGDET_API int kgdet_point_losses_backward(const float* const* outs, const int32_t* assigned, const float* gt_boxes,
                                const int64_t* gt_labels, const float* gt_keypoints, const float* avg_factor,
                                const float* num_visible, const float* grad_losses, int32_t B, int32_t G, int32_t map_h,
                                int32_t map_w, int32_t num_classes, int32_t num_keypoints, float stride,
                                float point_base_scale, const float* loss_weights, float gamma, float alpha, float beta,
                                float* const* grad_outs, void* stream);

/* ---- pointwise convolutions of the Kp3RepBlock ------------------------
 * replaces  cls_out / keypts_out / reppts_out 1x1 nn.Conv2d + the cascade's residual adds
 *           (reppoints_head_kp3rep_cas_1_assign_once.py:79-96,152-171,431-432,440-441)
 * out[n, coff + c - col_begin, pos] = sum_k A[n*HW + pos, k] * W[c, k] + bias[c] (+ residual[same index])
 * A: UMMA-tiled bf16 rows [M, K] (M = N*HW) written by kgdet_dcn_forward_prepared(KGDET_LAYOUT_TILED*) or
 * kgdet_nchw_to_tiled_bf16; W: packed by kgdet_pointwise_pack_weight from fp32 [Nout, K]; bias fp32 [Nout] or
 * NULL.  The Nout columns are split into up to KGDET_POINTWISE_MAX_SEGMENTS consecutive ranges, each written
 * to its own NCHW fp32 tensor (e.g. keypoints 0..587 and point set 588..753 of one GEMM).  K % 64 == 0.
 * split = 1: split-precision operands ("bf16x3": A = hi + lo, W = hi + lo, three bf16 MMAs per k-step,
 * fp32-grade results); A and W must have been produced with the same `split`. */
#define KGDET_POINTWISE_MAX_SEGMENTS 4
typedef struct kgdet_pointwise_segment {
  float* out;              /* NCHW fp32 [N, channels_total, HW] */
  const float* residual;   /* same layout and channel range as `out`, or NULL */
  int32_t col_begin, col_end;
  int32_t channels_total, channel_offset;
} kgdet_pointwise_segment;
KGDET_API size_t kgdet_pointwise_tiled_bytes(int32_t M, int32_t K, int split);
KGDET_API size_t kgdet_pointwise_packed_weight_bytes(int32_t Nout, int32_t K, int split);
KGDET_API int kgdet_pointwise_pack_weight(const float* weight, void* packed, int32_t Nout, int32_t K, int split,
                                void* stream);
/* NCHW fp32/bf16 [N, C, S] -> UMMA-tiled bf16 rows [N*S, C] (split = 1: [hi | lo]), optional ReLU
 * (stage-1 activations after a cuDNN convolution).  C % 64 == 0. */
KGDET_API int kgdet_nchw_to_tiled_bf16(const void* src, void* dst, int32_t N, int32_t C, int32_t S, int src_dtype,
                             int fuse_relu, int split, void* stream);
/* position-major fp32 rows [M, C] (channels_last activation) -> the same tiled rows; `bias` (fp32 [C] or NULL) is
 * added before the optional ReLU (the bias of the 3x3 convolution that produced the rows, conv -> +b -> ReLU) */
KGDET_API int kgdet_rows_to_tiled_bf16(const float* rows, const float* bias, void* tiled, int64_t M, int32_t C,
                             int fuse_relu, int split, void* stream);
KGDET_API int kgdet_pointwise_conv_tiled(const void* a_tiled, const void* w_packed, const float* bias, int32_t M,
                               int32_t K, int32_t Nout, int32_t HW, int split,
                               const kgdet_pointwise_segment* segs, int32_t nseg, void* stream);

/* ---- head towers in channels_last --------------------------------------
 * GroupNorm (+ ReLU) of ConvModule (mmdet/models/utils/conv_module.py:156-164) on position-major fp32
 * [N, HW, C]: torch.nn.GroupNorm semantics (biased variance, eps inside the square root).  The towers'
 * 3x3 convolutions stay cuDNN and run in channels_last, so no NCHW<->NHWC transposes remain. */
KGDET_API int kgdet_groupnorm_relu_nhwc(const float* x, const float* gamma, const float* beta, float eps,
                              int32_t groups, int fuse_relu, float* y, int32_t N, int32_t HW, int32_t C,
                              void* stream);

/* replaces deform_conv_backward_input_cuda     dcn/src/deform_conv_cuda.cpp:260-371
 *          (+ the input/offset/mask part of modulated_deform_conv_cuda_backward :566-679)
 * grad_input / grad_offset / grad_mask are fully overwritten (no pre-zeroing needed;
 * the reference pre-zeroes in Python: dcn/deform_conv.py:73-74).  `weight` is the raw
 * fp32 [Cout, C/g, kh, kw] tensor.  grad_mask may be NULL iff mask is NULL. */
KGDET_API size_t kgdet_dcn_backward_input_workspace_bytes(const kgdet_dcn_shape* shape, int dtype,
                                                int precision);
KGDET_API int kgdet_dcn_backward_input(const void* input, const float* offset, const float* mask,
                             const float* weight, const void* grad_output, void* grad_input,
                             float* grad_offset, float* grad_mask, const kgdet_dcn_shape* shape,
                             int dtype, int precision, void* workspace, size_t workspace_bytes,
                             void* stream);

/* replaces deform_conv_backward_parameters_cuda dcn/src/deform_conv_cuda.cpp:373-484
 * grad_weight (fp32 [Cout, C/g, kh, kw]) is overwritten with scale * dL/dW
 * (the reference accumulates into a zeroed tensor with scale = 1: dcn/deform_conv.py:84-91);
 * grad_bias ([Cout], fp32) may be NULL. */
KGDET_API size_t kgdet_dcn_backward_weight_workspace_bytes(const kgdet_dcn_shape* shape, int dtype,
                                                 int precision);
KGDET_API int kgdet_dcn_backward_weight(const void* input, const float* offset, const float* mask,
                              const void* grad_output, float* grad_weight, float* grad_bias,
                              float scale, const kgdet_dcn_shape* shape, int dtype, int precision,
                              void* workspace, size_t workspace_bytes, void* stream);

/* ---- NMS ------------------------------------------------------------------------------
 * replaces nms_cuda.nms   nms/src/nms_cuda.cpp:8-12 -> nms/src/nms_kernel.cu:70-131
 *      and nms_cpu.nms    nms/src/nms_cpu.cpp:61-67
 * dets: [n, 5] fp32 (x1, y1, x2, y2, score).  keep: [n] int64, the first *num_keep
 * entries are the ORIGINAL indices of the kept boxes in ascending order
 * (nms_kernel.cu:127-130); num_keep is a device int32.  Ties between equal scores
 * are broken by ascending original index (the reference's at::sort leaves them
 * unspecified). */
KGDET_API size_t kgdet_nms_workspace_bytes(int32_t n);
KGDET_API int kgdet_nms(const float* dets, int32_t n, float iou_thr, int cmp_mode, int64_t* keep,
              int32_t* num_keep, void* workspace, size_t workspace_bytes, void* stream);

/* Batched form over `nseg` independent segments (one per (image, class)) in ONE launch --
 * replaces the per-class Python loop of multiclass_nms_kp
 * (mmdet/core/post_processing/bbox_nms_kp.py:38-52).  Segment s owns rows
 * [seg_offsets[s], seg_offsets[s+1]) of dets; keep_flags[row] = 1 if the row survives
 * NMS within its segment.  seg_offsets: device int32 [nseg+1], or NULL for `nseg` uniform
 * segments of exactly max_seg_len rows (dense mode, total == nseg * max_seg_len);
 * max_seg_len >= the longest segment (host-side bound, e.g. nms_pre).  Rows whose score
 * is not > score_thr are treated as absent (flag 0) -- the `scores > score_thr` filter of
 * bbox_nms_kp.py:39 folded into the op so that a dense, fixed-shape (CUDA-graph capturable)
 * caller needs no compaction; pass -INFINITY to keep every row. */
KGDET_API size_t kgdet_nms_batched_workspace_bytes(int32_t total, int32_t nseg, int32_t max_seg_len);
KGDET_API int kgdet_nms_batched(const float* dets, const int32_t* seg_offsets, int32_t nseg, int32_t total,
                      int32_t max_seg_len, float iou_thr, float score_thr, int cmp_mode, uint8_t* keep_flags,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ---- sigmoid focal loss -----------------------------------------------------------------
 * replaces sigmoid_focal_loss_cuda.forward / .backward
 *   sigmoid_focal_loss/src/sigmoid_focal_loss.cpp:17-39 ->
 *   sigmoid_focal_loss/src/sigmoid_focal_loss_cuda.cu:24-59, 62-98
 * logits/losses/d_losses/d_logits: [M, C] (dtype), targets: [M] int64 (0 = background,
 * t = c+1 positive for column c, t < 0 ignored). */
KGDET_API int kgdet_sigmoid_focal_loss_forward(const void* logits, const int64_t* targets, int32_t M,
                                     int32_t C, float gamma, float alpha, void* losses, int dtype,
                                     void* stream);
KGDET_API int kgdet_sigmoid_focal_loss_backward(const void* logits, const int64_t* targets,
                                      const void* d_losses, int32_t M, int32_t C, float gamma,
                                      float alpha, void* d_logits, int dtype, void* stream);
/* Fused forward + row weight + sum (the reduction FocalLoss does in Python:
 * mmdet/models/losses/focal_loss.py:28-42, losses/utils.py:41-52).
 * *loss_sum (device fp32, must be zeroed by the caller) += sum_{m,c} loss[m,c] * weight[m];
 * weight ([M] fp32) may be NULL.  Backward of that scalar:
 * d_logits[m,c] = dloss/dx * weight[m] * (*grad_scale)  (grad_scale: device fp32). */
KGDET_API int kgdet_sigmoid_focal_loss_sum_forward(const void* logits, const int64_t* targets,
                                         const float* weight, int32_t M, int32_t C, float gamma,
                                         float alpha, float* loss_sum, int dtype, void* stream);
KGDET_API int kgdet_sigmoid_focal_loss_sum_backward(const void* logits, const int64_t* targets,
                                          const float* weight, const float* grad_scale, int32_t M,
                                          int32_t C, float gamma, float alpha, void* d_logits,
                                          int dtype, void* stream);

/* ---- point set -> bbox moment transform ------------------------------------------------------
 * replaces the ~12 PyTorch kernels of points2bbox(..., 'moment')
 *   mmdet/models/anchor_heads/reppoints_head_kp3rep_cas_1_assign_once.py:373-388
 * pts: [N, 2P, S] fp32 (S = H*W for NCHW maps, S = 1 for row-major [M, 2P]); channel
 * 2i = y_i, 2i+1 = x_i when y_first, swapped otherwise.  moment_transfer: device fp32[2]
 * (width, height log-scales).  bbox: [N, 4, S] = (x1, y1, x2, y2). */
KGDET_API int kgdet_points2bbox_moment_forward(const float* pts, const float* moment_transfer, int32_t N,
                                     int32_t P, int32_t S, int y_first, float* bbox,
                                     void* stream);
/* grad_pts: [N, 2P, S] overwritten; grad_moment_transfer: device fp32[2], must be zeroed by
 * the caller, receives moment_mul * dL/dt (KP3:378-379).  */
KGDET_API int kgdet_points2bbox_moment_backward(const float* pts, const float* moment_transfer,
                                      const float* grad_bbox, int32_t N, int32_t P, int32_t S,
                                      int y_first, float moment_mul, float* grad_pts,
                                      float* grad_moment_transfer, void* stream);

/* ---- measurement hook -----------------------------------------------------------------------
 * When both events (cudaEvent_t, passed as void*) are non-NULL, the NEXT kgdet_dcn_forward call
 * records `start` immediately before and `stop` immediately after its main contraction kernel
 * (fused tcgen05 or SIMT) on the call's stream, then clears the hook.  bench.py uses it to time
 * that kernel alone inside a full step (roofline.achieved); it has no effect on results. */
KGDET_API void kgdet_dcn_set_profile_events(void* start_event, void* stop_event);

/* Development hook: the NEXT fused tensor-core forward writes clock64() stamps of its pipeline
 * (per CTA 4 * nkb + 8 int64: entry, set-up done, per-k-block "stage full" seen by the MMA
 * issuer, accumulator ready, epilogue done, per-k-block producer progress) into `device_buffer`
 * if `entries` is large enough, then clears the hook.  tools/dcn_timeline.py reads it. */
KGDET_API void kgdet_dcn_set_timeline(void* device_buffer, long long entries);

/* ---- layout helpers used by the Python mirror ---------------------------------------------- */
/* NCHW (dtype) -> NHWC fp32 or bf16 and back; S = H*W */
KGDET_API int kgdet_nchw_to_nhwc(const void* src, void* dst, int32_t N, int32_t C, int32_t S, int src_dtype,
                       int dst_dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KGDET_B200_H_ */
