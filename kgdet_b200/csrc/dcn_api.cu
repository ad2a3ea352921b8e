// extern "C" entry points of the deformable convolution: validation, workspace carving,
// path selection (fused tcgen05 vs exact SIMT).  See include/kgdet_b200.h for the contract.
#include "dcn.cuh"

using namespace kgdet;

namespace {

struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* p) : base((char*)p) {}
  template <typename T> T* take(size_t bytes) {
    T* p = base ? (T*)(base + off) : nullptr;
    off += align_up(bytes, 1024);
    return p;
  }
};

thread_local cudaEvent_t g_prof_start = nullptr, g_prof_stop = nullptr;

bool valid_dtype(int d) { return d == KGDET_F32 || d == KGDET_BF16; }
bool valid_prec(int p) { return p >= KGDET_PREC_FP32 && p <= KGDET_PREC_TF32; }
bool use_umma(const DcnGeom& g, int precision) {
  return precision != KGDET_PREC_FP32 && umma_supported(g, precision);
}

// Prepared input.
//   exact SIMT path : plain NHWC copy in fp32.
//   tensor-core path: CHANNEL-BLOCKED planes  [C / BK][guard | N*H*W pixels | guard][BK channels]  in the
//                     compute type (BK = 64 bf16 / 32 fp32 channels = one 128-byte slab per pixel, the unit a
//                     k-block gathers).  The slabs one k-block touches are then contiguous 128-byte lines
//                     (NHWC with C = 256 puts them 512 bytes apart, which the L1 serves 1.4x slower:
//                     tools/micro/l1_gather_bench, 71 vs 99 B/clk/SM).  The fused kernel loads all four
//                     bilinear corners unconditionally (see dcn.cuh); each plane carries a zeroed guard band
//                     of dcn_guard_pixels() pixels on both sides for the corners that fall outside.
struct PrepIn { size_t guard_bytes, plane_bytes, in_bytes, total; int cdtype, planes, bk; };
PrepIn prep_in_layout(const DcnGeom& g, int precision) {
  PrepIn p;
  const bool fast = use_umma(g, precision);
  p.cdtype = (fast && precision == KGDET_PREC_BF16) ? KGDET_BF16 : KGDET_F32;
  const size_t esz = p.cdtype == KGDET_BF16 ? 2 : 4;
  if (fast) {
    p.bk = (int)(128 / esz);
    p.planes = g.C / p.bk;
    p.guard_bytes = (size_t)dcn_guard_pixels(g) * 128;
    p.in_bytes = (size_t)g.N * g.H * g.W * 128;
    p.plane_bytes = align_up(p.in_bytes + 2 * p.guard_bytes, 1024);
    p.total = p.plane_bytes * p.planes;
  } else {
    p.bk = g.C; p.planes = 1; p.guard_bytes = 0;
    p.in_bytes = (size_t)g.N * g.H * g.W * g.C * esz;
    p.plane_bytes = p.total = align_up(p.in_bytes, 1024);
  }
  return p;
}
size_t plan_layout_bytes(const DcnGeom& g, int precision) {
  return align_up(use_umma(g, precision) ? plan16_bytes(g) : plan_bytes(g), 1024);
}

int do_prepare_input(const DcnGeom& g, const void* input, void* prepared, int dtype, int precision,
                     cudaStream_t stream) {
  const PrepIn p = prep_in_layout(g, precision);
  char* base = (char*)prepared;
  if (!p.guard_bytes) return launch_transpose(input, base, g.N, g.C, g.H * g.W, dtype, p.cdtype, stream);
  // zero the two guard bands of every plane (2-D memsets: one row per plane)
  KG_CUDA(cudaMemset2DAsync(base, p.plane_bytes, 0, p.guard_bytes, p.planes, stream));
  KG_CUDA(cudaMemset2DAsync(base + p.guard_bytes + p.in_bytes, p.plane_bytes, 0,
                            p.plane_bytes - p.guard_bytes - p.in_bytes, p.planes, stream));
  return launch_nchw_to_blocked(input, base + p.guard_bytes, g.N, g.C, g.H * g.W, p.bk, p.plane_bytes, dtype,
                                p.cdtype, stream);
}

// same, from a position-major fp32 source (channels_last activations): no transpose, one 16-byte chunk per thread
int do_prepare_input_rows(const DcnGeom& g, const float* rows, void* prepared, int precision, cudaStream_t stream) {
  const PrepIn p = prep_in_layout(g, precision);
  char* base = (char*)prepared;
  if (!p.guard_bytes) {
    KG_CUDA(cudaMemcpyAsync(base, rows, p.in_bytes, cudaMemcpyDeviceToDevice, stream));   // SIMT path reads NHWC fp32
    return KGDET_OK;
  }
  KG_CUDA(cudaMemset2DAsync(base, p.plane_bytes, 0, p.guard_bytes, p.planes, stream));
  KG_CUDA(cudaMemset2DAsync(base + p.guard_bytes + p.in_bytes, p.plane_bytes, 0,
                            p.plane_bytes - p.guard_bytes - p.in_bytes, p.planes, stream));
  return launch_rows_to_blocked(rows, base + p.guard_bytes, (long long)g.N * g.H * g.W, g.C, p.bk, p.plane_bytes,
                                p.cdtype, stream);
}

int do_prepare_plan(const DcnGeom& g, const float* offset, const float* mask, void* plan, int precision,
                    cudaStream_t stream) {
  if (use_umma(g, precision))
    return launch_plan16(g, offset, mask, (SampleRec16*)plan,
                         precision == KGDET_PREC_BF16 ? PLAN16_BF16W : PLAN16_F32, stream);
  return launch_plan(g, offset, mask, (SampleRec*)plan, nullptr, stream);
}

int do_forward_prepared(const DcnGeom& g, const void* prepared, const void* plan, const void* weight_packed,
                        const float* bias, const OutSpec& o, int precision, cudaStream_t stream,
                        void* split_ws = nullptr) {
  const PrepIn p = prep_in_layout(g, precision);
  const char* in_nhwc = (const char*)prepared + p.guard_bytes;     // first pixel of plane 0
  cudaEvent_t ev0 = g_prof_start, ev1 = g_prof_stop;
  g_prof_start = g_prof_stop = nullptr;
  if (ev0 && ev1) KG_CUDA(cudaEventRecord(ev0, stream));
  int rc;
  if (use_umma(g, precision))
    rc = umma_forward(g, in_nhwc, p.plane_bytes, (const SampleRec16*)plan, weight_packed, bias, o, precision,
                      stream, split_ws);
  else
    rc = simt_forward(g, (const float*)in_nhwc, (const SampleRec*)plan, (const float*)weight_packed, bias, o,
                      stream);
  if (rc == KGDET_OK && ev0 && ev1) KG_CUDA(cudaEventRecord(ev1, stream));
  return rc;
}

struct BwdInWs { float *in_nhwc, *go_nhwc, *gin_nhwc, *w_dgrad; SampleRec* plan; SampleAux* aux; size_t total; };
BwdInWs carve_bwd_in(const DcnGeom& g, void* ws) {
  Carver c(ws);
  BwdInWs w;
  w.in_nhwc = c.take<float>((size_t)g.N * g.H * g.W * g.C * 4);
  w.go_nhwc = c.take<float>((size_t)g.M * g.Cout * 4);
  w.gin_nhwc = c.take<float>((size_t)g.N * g.H * g.W * g.C * 4);
  w.w_dgrad = c.take<float>(simt_packed_weight_bytes(g));
  w.plan = c.take<SampleRec>(plan_bytes(g));
  w.aux = c.take<SampleAux>(plan_aux_bytes(g));
  w.total = c.off;
  return w;
}

struct BwdWWs { float *in_nhwc, *go_nhwc; SampleRec* plan; size_t total; };
BwdWWs carve_bwd_w(const DcnGeom& g, void* ws) {
  Carver c(ws);
  BwdWWs w;
  w.in_nhwc = c.take<float>((size_t)g.N * g.H * g.W * g.C * 4);
  w.go_nhwc = c.take<float>((size_t)g.M * g.Cout * 4);
  w.plan = c.take<SampleRec>(plan_bytes(g));
  w.total = c.off;
  return w;
}

int check_ws(const char* name, void* ws, size_t have, size_t need) {
  if (!ws || have < need) {
    set_error("%s: workspace too small (%zu < %zu)", name, have, need);
    return KGDET_ERR_WORKSPACE;
  }
  if (((uintptr_t)ws & 255) != 0) {
    set_error("%s: workspace must be 256-byte aligned", name);
    return KGDET_ERR_INVALID_ARG;
  }
  return KGDET_OK;
}

}  // namespace

extern "C" int kgdet_dcn_fast_path_supported(const kgdet_dcn_shape* shape, int precision) {
  DcnGeom g;
  if (make_geom(shape, &g) != KGDET_OK) return 0;
  return use_umma(g, precision) ? 1 : 0;
}

extern "C" size_t kgdet_dcn_packed_weight_bytes(const kgdet_dcn_shape* shape, int precision) {
  DcnGeom g;
  if (make_geom(shape, &g) != KGDET_OK) return 0;
  return use_umma(g, precision) ? umma_packed_weight_bytes(g, precision) : simt_packed_weight_bytes(g);
}

extern "C" int kgdet_dcn_pack_weight(const float* weight, void* weight_packed,
                                     const kgdet_dcn_shape* shape, int precision, void* stream) {
  DcnGeom g;
  int rc = make_geom(shape, &g);
  if (rc != KGDET_OK) return rc;
  KG_CHECK_ARG(valid_prec(precision), "kgdet_dcn_pack_weight: bad precision %d", precision);
  KG_CHECK_ARG(weight && weight_packed, "kgdet_dcn_pack_weight: NULL pointer");
  if (use_umma(g, precision)) return umma_pack_weight(g, weight, weight_packed, precision, (cudaStream_t)stream);
  return simt_pack_weight(g, weight, (float*)weight_packed, (cudaStream_t)stream);
}

extern "C" size_t kgdet_dcn_forward_workspace_bytes(const kgdet_dcn_shape* shape, int, int precision) {
  DcnGeom g;
  if (make_geom(shape, &g) != KGDET_OK) return 0;
  return prep_in_layout(g, precision).total + plan_layout_bytes(g, precision) +
         (use_umma(g, precision) ? umma_split_ws_bytes(g, precision) : 0);
}

extern "C" int kgdet_dcn_forward(const void* input, const float* offset, const float* mask,
                                 const void* weight_packed, const float* bias, void* output,
                                 const kgdet_dcn_shape* shape, int dtype, int precision,
                                 void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DcnGeom g;
  int rc = make_geom(shape, &g);
  if (rc != KGDET_OK) return rc;
  KG_CHECK_ARG(valid_dtype(dtype), "kgdet_dcn_forward: bad dtype %d", dtype);
  KG_CHECK_ARG(valid_prec(precision), "kgdet_dcn_forward: bad precision %d", precision);
  KG_CHECK_ARG(input && offset && weight_packed && output, "kgdet_dcn_forward: NULL pointer");
  const size_t in_total = prep_in_layout(g, precision).total;
  const size_t plan_total = plan_layout_bytes(g, precision);
  const size_t split_total = use_umma(g, precision) ? umma_split_ws_bytes(g, precision) : 0;
  if ((rc = check_ws("kgdet_dcn_forward", workspace, workspace_bytes, in_total + plan_total + split_total)) != KGDET_OK)
    return rc;
  void* prepared = workspace;
  void* plan = (char*)workspace + in_total;
  void* split_ws = split_total ? (char*)workspace + in_total + plan_total : nullptr;
  if ((rc = do_prepare_input(g, input, prepared, dtype, precision, stream)) != KGDET_OK) return rc;
  if ((rc = do_prepare_plan(g, offset, mask, plan, precision, stream)) != KGDET_OK) return rc;
  OutSpec o{output, dtype, 0, g.Cout, 0, 0};
  return do_forward_prepared(g, prepared, plan, weight_packed, bias, o, precision, stream, split_ws);
}

extern "C" size_t kgdet_dcn_prepared_input_bytes(const kgdet_dcn_shape* shape, int precision) {
  DcnGeom g;
  if (make_geom(shape, &g) != KGDET_OK) return 0;
  return prep_in_layout(g, precision).total;
}

extern "C" int kgdet_dcn_prepare_input(const void* input, void* prepared_input, const kgdet_dcn_shape* shape,
                                       int dtype, int precision, void* stream) {
  DcnGeom g;
  int rc = make_geom(shape, &g);
  if (rc != KGDET_OK) return rc;
  KG_CHECK_ARG(valid_dtype(dtype) && valid_prec(precision), "kgdet_dcn_prepare_input: bad dtype/precision");
  KG_CHECK_ARG(input && prepared_input, "kgdet_dcn_prepare_input: NULL pointer");
  KG_CHECK_ARG(((uintptr_t)prepared_input & 255) == 0, "kgdet_dcn_prepare_input: buffer must be 256-byte aligned");
  return do_prepare_input(g, input, prepared_input, dtype, precision, (cudaStream_t)stream);
}

extern "C" int kgdet_dcn_prepare_input_rows(const float* rows, void* prepared_input, const kgdet_dcn_shape* shape,
                                            int precision, void* stream) {
  DcnGeom g;
  int rc = make_geom(shape, &g);
  if (rc != KGDET_OK) return rc;
  KG_CHECK_ARG(valid_prec(precision), "kgdet_dcn_prepare_input_rows: bad precision");
  KG_CHECK_ARG(rows && prepared_input, "kgdet_dcn_prepare_input_rows: NULL pointer");
  KG_CHECK_ARG(((uintptr_t)prepared_input & 255) == 0, "kgdet_dcn_prepare_input_rows: buffer must be 256-byte aligned");
  return do_prepare_input_rows(g, rows, prepared_input, precision, (cudaStream_t)stream);
}

extern "C" size_t kgdet_dcn_plan_bytes(const kgdet_dcn_shape* shape, int precision) {
  DcnGeom g;
  if (make_geom(shape, &g) != KGDET_OK) return 0;
  return plan_layout_bytes(g, precision);
}

extern "C" int kgdet_dcn_prepare_plan(const float* offset, const float* mask, void* plan,
                                      const kgdet_dcn_shape* shape, int precision, void* stream) {
  DcnGeom g;
  int rc = make_geom(shape, &g);
  if (rc != KGDET_OK) return rc;
  KG_CHECK_ARG(valid_prec(precision), "kgdet_dcn_prepare_plan: bad precision %d", precision);
  KG_CHECK_ARG(offset && plan, "kgdet_dcn_prepare_plan: NULL pointer");
  KG_CHECK_ARG(((uintptr_t)plan & 255) == 0, "kgdet_dcn_prepare_plan: buffer must be 256-byte aligned");
  return do_prepare_plan(g, offset, mask, plan, precision, (cudaStream_t)stream);
}

extern "C" int kgdet_dcn_prepare_plan_points(const float* points, int32_t channel_offset,
                                             int32_t channels_total, float gradient_mul, float one_minus_gradient_mul,
                                             void* plan, const kgdet_dcn_shape* shape, int precision, void* stream) {
  DcnGeom g;
  int rc = make_geom(shape, &g);
  if (rc != KGDET_OK) return rc;
  KG_CHECK_ARG(valid_prec(precision), "kgdet_dcn_prepare_plan_points: bad precision %d", precision);
  KG_CHECK_ARG(points && plan, "kgdet_dcn_prepare_plan_points: NULL pointer");
  KG_CHECK_ARG(g.dgroups == 1, "kgdet_dcn_prepare_plan_points: deformable_groups must be 1");
  KG_CHECK_ARG(channel_offset >= 0 && channel_offset + 2 * g.K <= channels_total,
               "kgdet_dcn_prepare_plan_points: channels [%d, %d) do not fit %d", channel_offset,
               channel_offset + 2 * g.K, channels_total);
  KG_CHECK_ARG(((uintptr_t)plan & 255) == 0, "kgdet_dcn_prepare_plan_points: buffer must be 256-byte aligned");
  const long long HoWo = (long long)g.Ho * g.Wo;
  const float* first = points + (long long)channel_offset * HoWo;
  const long long bstride = (long long)channels_total * HoWo;
  if (use_umma(g, precision))
    return launch_plan16(g, first, nullptr, (SampleRec16*)plan,
                         precision == KGDET_PREC_BF16 ? PLAN16_BF16W : PLAN16_F32, (cudaStream_t)stream, bstride, 1,
                         gradient_mul, one_minus_gradient_mul);
  return launch_plan(g, first, nullptr, (SampleRec*)plan, nullptr, (cudaStream_t)stream, bstride, 1, gradient_mul,
                     one_minus_gradient_mul);
}

extern "C" int kgdet_dcn_forward_prepared(const void* prepared_input, const void* plan,
                                          const void* weight_packed, const float* bias, void* output,
                                          int32_t out_channel_offset, int32_t out_channels_total,
                                          int fuse_relu, int out_layout, const kgdet_dcn_shape* shape, int dtype,
                                          int precision, void* workspace, size_t workspace_bytes, void* stream) {
  DcnGeom g;
  int rc = make_geom(shape, &g);
  if (rc != KGDET_OK) return rc;
  KG_CHECK_ARG(valid_dtype(dtype) && valid_prec(precision), "kgdet_dcn_forward_prepared: bad dtype/precision");
  KG_CHECK_ARG(prepared_input && plan && weight_packed && output, "kgdet_dcn_forward_prepared: NULL pointer");
  KG_CHECK_ARG(out_channel_offset >= 0 && out_channel_offset + g.Cout <= out_channels_total,
               "kgdet_dcn_forward_prepared: channel slice [%d, %d) does not fit %d channels",
               out_channel_offset, out_channel_offset + g.Cout, out_channels_total);
  KG_CHECK_ARG(out_layout >= KGDET_LAYOUT_NCHW && out_layout <= KGDET_LAYOUT_TILED_SPLIT,
               "kgdet_dcn_forward_prepared: bad output layout %d", out_layout);
  if (out_layout != KGDET_LAYOUT_NCHW) {
    KG_CHECK_ARG(dtype == KGDET_BF16, "kgdet_dcn_forward_prepared: the tiled layouts store bf16");
    KG_CHECK_ARG(use_umma(g, precision), "kgdet_dcn_forward_prepared: tiled output needs the tensor-core path");
    KG_CHECK_ARG(out_channel_offset % 32 == 0 && out_channels_total % 64 == 0,
                 "kgdet_dcn_forward_prepared: tiled output needs channel offset %% 32 == 0 and total %% 64 == 0");
  }
  OutSpec o{output, dtype, out_channel_offset, out_channels_total, fuse_relu ? 1 : 0,
            out_layout};
  // optional scratch for the k-block split (small maps; TF32X3 accumulator promotion): used when it is big enough
  const size_t split_total = (use_umma(g, precision) && out_layout == KGDET_LAYOUT_NCHW) ? umma_split_ws_bytes(g, precision) : 0;
  void* split_ws = nullptr;
  if (split_total && workspace) {
    if ((rc = check_ws("kgdet_dcn_forward_prepared", workspace, workspace_bytes, split_total)) != KGDET_OK) return rc;
    split_ws = workspace;
  }
  return do_forward_prepared(g, prepared_input, plan, weight_packed, bias, o, precision, (cudaStream_t)stream, split_ws);
}

extern "C" size_t kgdet_dcn_forward_prepared_workspace_bytes(const kgdet_dcn_shape* shape, int precision) {
  DcnGeom g;
  if (make_geom(shape, &g) != KGDET_OK) return 0;
  return use_umma(g, precision) ? umma_split_ws_bytes(g, precision) : 0;
}

extern "C" void kgdet_dcn_set_profile_events(void* start_event, void* stop_event) {
  g_prof_start = (cudaEvent_t)start_event;
  g_prof_stop = (cudaEvent_t)stop_event;
}

extern "C" size_t kgdet_dcn_backward_input_workspace_bytes(const kgdet_dcn_shape* shape, int, int precision) {
  DcnGeom g;
  if (make_geom(shape, &g) != KGDET_OK) return 0;
  return bwd_tc_supported(g, precision) ? bwd_tc_input_workspace_bytes(g) : carve_bwd_in(g, nullptr).total;
}

extern "C" int kgdet_dcn_backward_input(const void* input, const float* offset, const float* mask,
                                        const float* weight, const void* grad_output, void* grad_input,
                                        float* grad_offset, float* grad_mask,
                                        const kgdet_dcn_shape* shape, int dtype, int precision,
                                        void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DcnGeom g;
  int rc = make_geom(shape, &g);
  if (rc != KGDET_OK) return rc;
  KG_CHECK_ARG(valid_dtype(dtype), "kgdet_dcn_backward_input: bad dtype %d", dtype);
  KG_CHECK_ARG(valid_prec(precision), "kgdet_dcn_backward_input: bad precision %d", precision);
  KG_CHECK_ARG(input && offset && weight && grad_output && grad_input && grad_offset,
               "kgdet_dcn_backward_input: NULL pointer");
  KG_CHECK_ARG((mask == nullptr) == (grad_mask == nullptr),
               "kgdet_dcn_backward_input: mask and grad_mask must both be given or both be NULL");
  if (bwd_tc_supported(g, precision)) {       // bf16 mode: tensor-core GEMM + warp-reduction col2im
    if ((rc = check_ws("kgdet_dcn_backward_input", workspace, workspace_bytes,
                       bwd_tc_input_workspace_bytes(g))) != KGDET_OK) return rc;
    return bwd_tc_input(g, input, offset, mask, weight, grad_output, grad_input, grad_offset, grad_mask, dtype,
                        workspace, stream);
  }
  BwdInWs w = carve_bwd_in(g, workspace);
  if ((rc = check_ws("kgdet_dcn_backward_input", workspace, workspace_bytes, w.total)) != KGDET_OK) return rc;
  const int HW = g.H * g.W, HoWo = g.Ho * g.Wo;
  if ((rc = launch_transpose(input, w.in_nhwc, g.N, g.C, HW, dtype, KGDET_F32, stream)) != KGDET_OK) return rc;
  if ((rc = launch_transpose(grad_output, w.go_nhwc, g.N, g.Cout, HoWo, dtype, KGDET_F32, stream)) != KGDET_OK) return rc;
  if ((rc = launch_plan(g, offset, mask, w.plan, w.aux, stream)) != KGDET_OK) return rc;
  if ((rc = simt_pack_weight_dgrad(g, weight, w.w_dgrad, stream)) != KGDET_OK) return rc;
  KG_CUDA(cudaMemsetAsync(w.gin_nhwc, 0, (size_t)g.N * HW * g.C * 4, stream));
  KG_CUDA(cudaMemsetAsync(grad_offset, 0, (size_t)g.N * g.dgroups * 2 * g.K * HoWo * 4, stream));
  if (grad_mask) KG_CUDA(cudaMemsetAsync(grad_mask, 0, (size_t)g.N * g.dgroups * g.K * HoWo * 4, stream));
  if ((rc = simt_backward_input(g, w.in_nhwc, w.go_nhwc, w.plan, w.aux, w.w_dgrad, w.gin_nhwc, grad_offset,
                                grad_mask, stream)) != KGDET_OK) return rc;
  // NHWC fp32 -> NCHW (dtype)
  return launch_transpose(w.gin_nhwc, grad_input, g.N, HW, g.C, KGDET_F32, dtype, stream);
}

extern "C" size_t kgdet_dcn_backward_weight_workspace_bytes(const kgdet_dcn_shape* shape, int, int precision) {
  DcnGeom g;
  if (make_geom(shape, &g) != KGDET_OK) return 0;
  return bwd_tc_supported(g, precision) ? bwd_tc_weight_workspace_bytes(g) : carve_bwd_w(g, nullptr).total;
}

extern "C" int kgdet_dcn_backward_weight(const void* input, const float* offset, const float* mask,
                                         const void* grad_output, float* grad_weight, float* grad_bias,
                                         float scale, const kgdet_dcn_shape* shape, int dtype,
                                         int precision, void* workspace, size_t workspace_bytes,
                                         void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DcnGeom g;
  int rc = make_geom(shape, &g);
  if (rc != KGDET_OK) return rc;
  KG_CHECK_ARG(valid_dtype(dtype), "kgdet_dcn_backward_weight: bad dtype %d", dtype);
  KG_CHECK_ARG(valid_prec(precision), "kgdet_dcn_backward_weight: bad precision %d", precision);
  KG_CHECK_ARG(input && offset && grad_output && grad_weight, "kgdet_dcn_backward_weight: NULL pointer");
  if (bwd_tc_supported(g, precision)) {
    if ((rc = check_ws("kgdet_dcn_backward_weight", workspace, workspace_bytes,
                       bwd_tc_weight_workspace_bytes(g))) != KGDET_OK) return rc;
    return bwd_tc_weight(g, input, offset, mask, grad_output, grad_weight, grad_bias, scale, dtype, workspace,
                         stream);
  }
  BwdWWs w = carve_bwd_w(g, workspace);
  if ((rc = check_ws("kgdet_dcn_backward_weight", workspace, workspace_bytes, w.total)) != KGDET_OK) return rc;
  const int HW = g.H * g.W, HoWo = g.Ho * g.Wo;
  if ((rc = launch_transpose(input, w.in_nhwc, g.N, g.C, HW, dtype, KGDET_F32, stream)) != KGDET_OK) return rc;
  if ((rc = launch_transpose(grad_output, w.go_nhwc, g.N, g.Cout, HoWo, dtype, KGDET_F32, stream)) != KGDET_OK) return rc;
  if ((rc = launch_plan(g, offset, mask, w.plan, nullptr, stream)) != KGDET_OK) return rc;
  return simt_backward_weight(g, w.in_nhwc, w.go_nhwc, w.plan, scale, grad_weight, grad_bias, stream);
}
