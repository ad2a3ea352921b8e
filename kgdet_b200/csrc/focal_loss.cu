// Sigmoid focal loss forward / backward, plus the fused "weight + sum" form.
//
// Replaces mmdet/ops/sigmoid_focal_loss/src/sigmoid_focal_loss_cuda.cu:24-59 (forward) and
// :62-98 (backward).  The arithmetic keeps the reference's float/double promotion pattern for
// scalar_t = float (its `1.` literals are doubles; expf/powf/logf are single precision) so the
// result is the same value the reference kernel produces, not merely close to it.
// The *_sum_* kernels fold in what FocalLoss does afterwards in Python
// (mmdet/models/losses/focal_loss.py:28-42, losses/utils.py:41-52): loss * weight[:, None]
// and the sum, so the [M, C] loss tensor never reaches HBM.
#include "focal.cuh"

namespace kgdet {

template <typename T> __device__ __forceinline__ float ld_f(const T* p, size_t i);
template <> __device__ __forceinline__ float ld_f<float>(const float* p, size_t i) { return p[i]; }
template <> __device__ __forceinline__ float ld_f<__nv_bfloat16>(const __nv_bfloat16* p, size_t i) {
  return __bfloat162float(p[i]);
}
template <typename T> __device__ __forceinline__ void st_f(T* p, size_t i, float v);
template <> __device__ __forceinline__ void st_f<float>(float* p, size_t i, float v) { p[i] = v; }
template <> __device__ __forceinline__ void st_f<__nv_bfloat16>(__nv_bfloat16* p, size_t i, float v) {
  p[i] = __float2bfloat16(v);
}

template <typename T>
__global__ void focal_fwd_kernel(const T* __restrict__ logits, const int64_t* __restrict__ targets,
                                 int total, int C, float gamma, float alpha, T* __restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += blockDim.x * gridDim.x) {
    int n = i / C, d = i - n * C;
    int t = (int)targets[n];                                                          // :34
    st_f(out, i, focal_fwd_value(ld_f(logits, i), t, d, gamma, alpha));
  }
}

template <typename T>
__global__ void focal_bwd_kernel(const T* __restrict__ logits, const int64_t* __restrict__ targets,
                                 const T* __restrict__ d_losses, int total, int C, float gamma,
                                 float alpha, T* __restrict__ d_logits) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += blockDim.x * gridDim.x) {
    int n = i / C, d = i - n * C;
    int t = (int)targets[n];
    float g = focal_bwd_value(ld_f(logits, i), t, d, gamma, alpha);
    st_f(d_logits, i, g * ld_f(d_losses, i));                                         // :94
  }
}

template <typename T>
__global__ void focal_sum_fwd_kernel(const T* __restrict__ logits,
                                     const int64_t* __restrict__ targets,
                                     const float* __restrict__ weight, int total, int C,
                                     float gamma, float alpha, float* __restrict__ loss_sum) {
  float acc = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += blockDim.x * gridDim.x) {
    int n = i / C, d = i - n * C;
    int t = (int)targets[n];
    float v = focal_fwd_value(ld_f(logits, i), t, d, gamma, alpha);
    acc += weight ? v * weight[n] : v;
  }
  __shared__ float part[32];
  acc = warp_sum(acc);
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) part[wid] = acc;
  __syncthreads();
  if (wid == 0) {
    float v = (lane < (blockDim.x >> 5)) ? part[lane] : 0.f;
    v = warp_sum(v);
    if (lane == 0) atomicAdd(loss_sum, v);
  }
}

template <typename T>
__global__ void focal_sum_bwd_kernel(const T* __restrict__ logits,
                                     const int64_t* __restrict__ targets,
                                     const float* __restrict__ weight,
                                     const float* __restrict__ grad_scale, int total, int C,
                                     float gamma, float alpha, T* __restrict__ d_logits) {
  const float gs = *grad_scale;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += blockDim.x * gridDim.x) {
    int n = i / C, d = i - n * C;
    int t = (int)targets[n];
    float g = focal_bwd_value(ld_f(logits, i), t, d, gamma, alpha);
    float w = weight ? weight[n] : 1.f;
    st_f(d_logits, i, g * w * gs);
  }
}

static int focal_grid(int total) {
  int blocks = ceil_div(total, 256);
  int cap = num_sms() * 8;
  return blocks < cap ? (blocks < 1 ? 1 : blocks) : cap;
}

}  // namespace kgdet

using namespace kgdet;

#define FOCAL_COMMON_CHECKS(name)                                                            \
  KG_CHECK_ARG(M >= 0 && C >= 1, name ": bad sizes M=%d C=%d", M, C);                        \
  KG_CHECK_ARG((long long)M * C < (1ll << 31), name ": M*C overflows int32");                \
  KG_CHECK_ARG(dtype == KGDET_F32 || dtype == KGDET_BF16, name ": bad dtype %d", dtype);     \
  if (M == 0) return KGDET_OK;

extern "C" int kgdet_sigmoid_focal_loss_forward(const void* logits, const int64_t* targets,
                                                int32_t M, int32_t C, float gamma, float alpha,
                                                void* losses, int dtype, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  FOCAL_COMMON_CHECKS("kgdet_sigmoid_focal_loss_forward");
  KG_CHECK_ARG(logits && targets && losses, "kgdet_sigmoid_focal_loss_forward: NULL pointer");
  int total = M * C;
  if (dtype == KGDET_F32)
    focal_fwd_kernel<float><<<focal_grid(total), 256, 0, stream>>>(
        (const float*)logits, targets, total, C, gamma, alpha, (float*)losses);
  else
    focal_fwd_kernel<__nv_bfloat16><<<focal_grid(total), 256, 0, stream>>>(
        (const __nv_bfloat16*)logits, targets, total, C, gamma, alpha, (__nv_bfloat16*)losses);
  KG_LAUNCH_CHECK("focal_fwd_kernel");
  return KGDET_OK;
}

extern "C" int kgdet_sigmoid_focal_loss_backward(const void* logits, const int64_t* targets,
                                                 const void* d_losses, int32_t M, int32_t C,
                                                 float gamma, float alpha, void* d_logits,
                                                 int dtype, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  FOCAL_COMMON_CHECKS("kgdet_sigmoid_focal_loss_backward");
  KG_CHECK_ARG(logits && targets && d_losses && d_logits,
               "kgdet_sigmoid_focal_loss_backward: NULL pointer");
  int total = M * C;
  if (dtype == KGDET_F32)
    focal_bwd_kernel<float><<<focal_grid(total), 256, 0, stream>>>(
        (const float*)logits, targets, (const float*)d_losses, total, C, gamma, alpha,
        (float*)d_logits);
  else
    focal_bwd_kernel<__nv_bfloat16><<<focal_grid(total), 256, 0, stream>>>(
        (const __nv_bfloat16*)logits, targets, (const __nv_bfloat16*)d_losses, total, C, gamma,
        alpha, (__nv_bfloat16*)d_logits);
  KG_LAUNCH_CHECK("focal_bwd_kernel");
  return KGDET_OK;
}

extern "C" int kgdet_sigmoid_focal_loss_sum_forward(const void* logits, const int64_t* targets,
                                                    const float* weight, int32_t M, int32_t C,
                                                    float gamma, float alpha, float* loss_sum,
                                                    int dtype, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  FOCAL_COMMON_CHECKS("kgdet_sigmoid_focal_loss_sum_forward");
  KG_CHECK_ARG(logits && targets && loss_sum, "kgdet_sigmoid_focal_loss_sum_forward: NULL pointer");
  int total = M * C;
  if (dtype == KGDET_F32)
    focal_sum_fwd_kernel<float><<<focal_grid(total), 256, 0, stream>>>(
        (const float*)logits, targets, weight, total, C, gamma, alpha, loss_sum);
  else
    focal_sum_fwd_kernel<__nv_bfloat16><<<focal_grid(total), 256, 0, stream>>>(
        (const __nv_bfloat16*)logits, targets, weight, total, C, gamma, alpha, loss_sum);
  KG_LAUNCH_CHECK("focal_sum_fwd_kernel");
  return KGDET_OK;
}

extern "C" int kgdet_sigmoid_focal_loss_sum_backward(const void* logits, const int64_t* targets,
                                                     const float* weight, const float* grad_scale,
                                                     int32_t M, int32_t C, float gamma,
                                                     float alpha, void* d_logits, int dtype,
                                                     void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  FOCAL_COMMON_CHECKS("kgdet_sigmoid_focal_loss_sum_backward");
  KG_CHECK_ARG(logits && targets && grad_scale && d_logits,
               "kgdet_sigmoid_focal_loss_sum_backward: NULL pointer");
  int total = M * C;
  if (dtype == KGDET_F32)
    focal_sum_bwd_kernel<float><<<focal_grid(total), 256, 0, stream>>>(
        (const float*)logits, targets, weight, grad_scale, total, C, gamma, alpha,
        (float*)d_logits);
  else
    focal_sum_bwd_kernel<__nv_bfloat16><<<focal_grid(total), 256, 0, stream>>>(
        (const __nv_bfloat16*)logits, targets, weight, grad_scale, total, C, gamma, alpha,
        (__nv_bfloat16*)d_logits);
  KG_LAUNCH_CHECK("focal_sum_bwd_kernel");
  return KGDET_OK;
}
