// Error state, device queries and geometry validation for the C ABI.
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include "common.cuh"

namespace kgdet {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error in %s: %s", what, cudaGetErrorString(e));
  return KGDET_ERR_CUDA;
}

static std::atomic<unsigned long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
  }
  return sms;
}

// Mirrors shape_check (dcn/src/deform_conv_cuda.cpp:61-149) and the output-size rule of
// dcn/deform_conv.py:96-110.
int make_geom(const kgdet_dcn_shape* s, DcnGeom* g) {
  KG_CHECK_ARG(s != nullptr, "dcn: shape is NULL");
  KG_CHECK_ARG(s->kw > 0 && s->kh > 0, "kernel size should be greater than zero, but got kH: %d kW: %d",
               s->kh, s->kw);
  KG_CHECK_ARG(s->stride_w > 0 && s->stride_h > 0,
               "stride should be greater than zero, but got dH: %d dW: %d", s->stride_h, s->stride_w);
  KG_CHECK_ARG(s->dil_w > 0 && s->dil_h > 0,
               "dilation should be greater than 0, but got dilationH: %d dilationW: %d", s->dil_h,
               s->dil_w);
  KG_CHECK_ARG(s->pad_h >= 0 && s->pad_w >= 0, "padding must be non-negative");
  KG_CHECK_ARG(s->N >= 1 && s->C >= 1 && s->H >= 1 && s->W >= 1 && s->Cout >= 1,
               "dcn: non-positive tensor size N=%d C=%d H=%d W=%d Cout=%d", s->N, s->C, s->H, s->W,
               s->Cout);
  KG_CHECK_ARG(s->groups >= 1 && s->deformable_groups >= 1, "dcn: groups must be >= 1");
  KG_CHECK_ARG(s->C % s->groups == 0, "in_channels %d cannot be divisible by groups %d", s->C,
               s->groups);
  KG_CHECK_ARG(s->Cout % s->groups == 0, "out_channels %d cannot be divisible by groups %d", s->Cout,
               s->groups);
  KG_CHECK_ARG(s->C % s->deformable_groups == 0, "input channels must divide deformable group size");
  g->N = s->N; g->C = s->C; g->H = s->H; g->W = s->W; g->Cout = s->Cout;
  g->kh = s->kh; g->kw = s->kw; g->K = s->kh * s->kw;
  g->sh = s->stride_h; g->sw = s->stride_w; g->ph = s->pad_h; g->pw = s->pad_w;
  g->dh = s->dil_h; g->dw = s->dil_w;
  g->groups = s->groups; g->dgroups = s->deformable_groups;
  g->Ho = (s->H + 2 * s->pad_h - (s->dil_h * (s->kh - 1) + 1)) / s->stride_h + 1;
  g->Wo = (s->W + 2 * s->pad_w - (s->dil_w * (s->kw - 1) + 1)) / s->stride_w + 1;
  if (g->Ho < 1 || g->Wo < 1) {
    set_error("Given input size: (%d x %d x %d). Calculated output size: (%d x %d x %d). Output size "
              "is too small", s->C, s->H, s->W, s->Cout, g->Ho, g->Wo);
    return KGDET_ERR_INVALID_ARG;
  }
  KG_CHECK_ARG(s->H >= s->kh && s->W >= s->kw, "input image is smaller than kernel");
  long long M = (long long)s->N * g->Ho * g->Wo;
  KG_CHECK_ARG(M * g->K * g->dgroups < (1ll << 31) && (long long)s->N * s->H * s->W < (1ll << 31) &&
                   M * s->Cout < (1ll << 40) && (long long)s->N * s->H * s->W * s->C < (1ll << 40),
               "dcn: problem too large for 32-bit position indexing");
  g->M = (int)M;
  return KGDET_OK;
}

}  // namespace kgdet

extern "C" const char* kgdet_last_error(void) { return kgdet::g_err; }
extern "C" int kgdet_abi_version(void) { return KGDET_ABI_VERSION; }
extern "C" uint64_t kgdet_launch_count(void) { return kgdet::g_launches.load(std::memory_order_relaxed); }
