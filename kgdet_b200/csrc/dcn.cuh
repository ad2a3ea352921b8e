// Internal declarations shared by the deformable-convolution translation units.
#pragma once
#include "common.cuh"

namespace kgdet {

// One record per (output position m, deformable group, tap): the four bilinear corners of
// the sampling point (deform_conv_cuda_kernel.cu:88-110) resolved ONCE per offset tensor --
// the reference recomputes them in every one of the C channel threads (:198-240).
//   pix[i]  pixel index (n*H*W + h*W + w) of corner i, or 0 when the corner is unusable
//   w[i]    bilinear weight of corner i * [corner inside the map] * [sample inside the
//           (-1,H)x(-1,W) window] * mask (modulated DCN) -- 0 for unusable corners
// corner order: (h_low,w_low) (h_low,w_high) (h_high,w_low) (h_high,w_high)
struct __align__(16) SampleRec {
  int pix[4];
  float w[4];
};
// Extra per-sample data only the backward passes need.
//   lh, lw  fractional parts;  mask  modulation value (1 for plain DCN)
//   valid   bit i = corner i usable (inside map and sample inside window)
struct __align__(16) SampleAux {
  float lh, lw, mask;
  int valid;
};

// Compact record of the tensor-core path (deformable_groups == 1): 16 bytes per (position, tap).
//   base   pixel index of the (h_low, w_low) corner, n*H*W + h_low*W + w_low; may point outside
//          the image when that corner is unusable -- corners are addressed as base + {0, 1, W, W+1}
//          and never dereferenced when their weight is 0
//   lh, lw fractional parts; the two mantissa LSBs carry the corner-validity bits of the axis:
//          bit0 = low index inside the map, bit1 = high index inside the map (costs < 4 ulp of lh/lw)
//   scale  modulation mask (1 for plain DCN) * [sample inside the (-1,H)x(-1,W) window]
struct __align__(16) SampleRec16 {
  int base;
  float lh, lw, scale;
};
// Second flavour of the same 16 bytes, used by the bf16 mode: {base, bf16x2(w0,w1), bf16x2(w2,w3), 0}
// with the four corner weights (validity, window test and mask folded in) pre-rounded to bf16.
// In both flavours `base` is always a safe address: the fused kernel loads all four corners
// unconditionally (weight 0 for unusable ones) from channel-blocked planes that each carry a zeroed guard
// band of dcn_guard_pixels() pixels on both sides.
enum { PLAN16_F32 = 0, PLAN16_BF16W = 1 };
size_t plan16_bytes(const DcnGeom& g);
// batch_stride: elements between images of `offset` (0 = contiguous [N, 2K, Ho, Wo]); points = 1: `offset`
// holds absolute point offsets and the base grid is subtracted (kgdet_dcn_prepare_plan_points)
int launch_plan16(const DcnGeom& g, const float* offset, const float* mask, SampleRec16* rec, int fmt,
                  cudaStream_t stream, long long batch_stride = 0, int points = 0, float gm = 0.f, float gm1 = 0.f);
static inline int dcn_guard_pixels(const DcnGeom& g) { return g.W + 2; }

size_t plan_rows(const DcnGeom& g);                 // M rounded up to 256
size_t plan_bytes(const DcnGeom& g);                // SampleRec array
size_t plan_aux_bytes(const DcnGeom& g);            // SampleAux array
int launch_plan(const DcnGeom& g, const float* offset, const float* mask, SampleRec* rec,
                SampleAux* aux /* may be NULL */, cudaStream_t stream, long long batch_stride = 0, int points = 0,
                float gm = 0.f, float gm1 = 0.f);

// src [B, R, Cc] -> dst [B, Cc, R] with dtype conversion (NCHW <-> NHWC)
int launch_transpose(const void* src, void* dst, int B, int R, int Cc, int src_dtype,
                     int dst_dtype, cudaStream_t stream);

// src NCHW [N][C][S] -> dst planes [C / bk][... N*S pixels ...][bk] (plane p starts plane_bytes * p after dst)
int launch_nchw_to_blocked(const void* src, void* dst, int N, int C, int S, int bk, size_t plane_bytes,
                           int src_dtype, int dst_dtype, cudaStream_t stream);

// rows [M, C] fp32 (position-major) -> the same planes (tower_nhwc.cu)
int launch_rows_to_blocked(const float* rows, void* dst, long long M, int C, int bk, size_t plane_bytes, int dst_dtype,
                           cudaStream_t stream);

// Where and how a forward kernel writes its result: channel slice [coff, coff + Cout) of an NCHW tensor
// with `ctot` channels, optional fused ReLU.
struct OutSpec {
  void* out;
  int dtype;      // KGDET_F32 / KGDET_BF16
  int coff, ctot;
  int relu;
  int nhwc;       // KGDET_LAYOUT_*: NCHW [N, ctot, Ho, Wo], or (tensor-core path only) UMMA-tiled bf16 rows
                  // [N*Ho*Wo, ctot] / split [hi | lo] -- see pointwise_umma.cu
};

// ---- exact fp32 SIMT path (any stride / dilation / groups / deformable_groups / mask) ----
size_t simt_packed_weight_bytes(const DcnGeom& g);
int simt_pack_weight(const DcnGeom& g, const float* weight, float* packed, cudaStream_t stream);
int simt_forward(const DcnGeom& g, const float* in_nhwc, const SampleRec* plan,
                 const float* packed_w, const float* bias, const OutSpec& o, cudaStream_t stream);
// weight_dgrad: [groups][K][Cout/g][C/g] fp32 (built by simt_pack_weight_dgrad)
int simt_pack_weight_dgrad(const DcnGeom& g, const float* weight, float* packed,
                           cudaStream_t stream);
int simt_backward_input(const DcnGeom& g, const float* in_nhwc, const float* go_nhwc,
                        const SampleRec* plan, const SampleAux* aux, const float* w_dgrad,
                        float* gin_nhwc /* zeroed */, float* grad_offset /* zeroed */,
                        float* grad_mask /* zeroed or NULL */, cudaStream_t stream);
int simt_backward_weight(const DcnGeom& g, const float* in_nhwc, const float* go_nhwc,
                         const SampleRec* plan, float scale, float* grad_weight,
                         float* grad_bias, cudaStream_t stream);

// ---- fused tcgen05 path (groups == deformable_groups == 1, C % 64 == 0, Cout % 16 == 0) ----
bool umma_supported(const DcnGeom& g, int precision);
size_t umma_packed_weight_bytes(const DcnGeom& g, int precision);
int umma_pack_weight(const DcnGeom& g, const float* weight, void* packed, int precision,
                     cudaStream_t stream);
// in_blocked: first pixel of plane 0 of the channel-blocked input (see dcn_api.cu), planes plane_bytes apart
// development timeline hook (kgdet_dcn_set_timeline): consumed by the next fused DCN forward or pointwise GEMM
extern thread_local long long* g_timeline;
extern thread_local long long g_timeline_entries;

int umma_forward(const DcnGeom& g, const void* in_blocked, size_t plane_bytes, const SampleRec16* plan,
                 const void* packed_w, const float* bias, const OutSpec& o, int precision, cudaStream_t stream, void* split_ws = nullptr);
int umma_splits(const DcnGeom& g, int precision);
size_t umma_split_ws_bytes(const DcnGeom& g, int precision);   // 0 when the call is not split

// ---- tensor-core backward (bf16 mode; dcn_bwd_tc.cu) ----
bool bwd_tc_supported(const DcnGeom& g, int precision);
size_t bwd_tc_input_workspace_bytes(const DcnGeom& g);
size_t bwd_tc_weight_workspace_bytes(const DcnGeom& g);
int bwd_tc_input(const DcnGeom& g, const void* input, const float* offset, const float* mask,
                 const float* weight, const void* grad_output, void* grad_input, float* grad_offset,
                 float* grad_mask, int dtype, void* ws, cudaStream_t stream);
int bwd_tc_weight(const DcnGeom& g, const void* input, const float* offset, const float* mask,
                  const void* grad_output, float* grad_weight, float* grad_bias, float scale, int dtype,
                  void* ws, cudaStream_t stream);

}  // namespace kgdet
