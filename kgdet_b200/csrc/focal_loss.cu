// Sigmoid focal loss forward / backward, plus the fused "weight + sum" form.
//
// Replaces mmdet/ops/sigmoid_focal_loss/src/sigmoid_focal_loss_cuda.cu:24-59 (forward) and
// :62-98 (backward).  Two arithmetic forms (focal.cuh): the reference's float/double promotion pattern for
// scalar_t = float mirrored operation by operation (its `1.` literals are doubles) -- the value the reference kernel
// produces, used by the training-loss kernels (point_loss.cu) and, with KGDET_FOCAL_EXACT=1, everywhere --
// and a single-precision form (<= 3e-7 from it), the default here; the elementwise fp32 entry points also process
// four elements per thread.
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
                                     float gamma, float alpha, float* __restrict__ loss_sum, int fast) {
  float acc = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += blockDim.x * gridDim.x) {
    int n = i / C, d = i - n * C;
    int t = (int)targets[n];
    const float x = ld_f(logits, i);
    float v = fast == 2 ? focal_fwd_fast<true>(x, t, d, gamma, alpha)
                        : (fast == 1 ? focal_fwd_fast<false>(x, t, d, gamma, alpha) : focal_fwd_value(x, t, d, gamma, alpha));
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
                                     float gamma, float alpha, T* __restrict__ d_logits, int fast) {
  const float gs = *grad_scale;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += blockDim.x * gridDim.x) {
    int n = i / C, d = i - n * C;
    int t = (int)targets[n];
    const float x = ld_f(logits, i);
    float g = fast == 2 ? focal_bwd_fast<true>(x, t, d, gamma, alpha)
                        : (fast == 1 ? focal_bwd_fast<false>(x, t, d, gamma, alpha) : focal_bwd_value(x, t, d, gamma, alpha));
    float w = weight ? weight[n] : 1.f;
    st_f(d_logits, i, g * w * gs);
  }
}

// fp32 tensors, four consecutive elements per thread (one 16-byte load / store; the row index is divided out once
// per four elements), single-precision arithmetic (focal.cuh): HBM-bound instead of bound by fp64 instruction issue.
// The labels of the (at most two, C >= 4) rows a quad touches are loaded TOGETHER with the quad -- a quad's label
// load used to be issued when its arithmetic started, a second dependent round trip per iteration.
template <bool BWD, bool G2>
__device__ __forceinline__ void focal_quad(const float4& x4, const float4& g4, int i, int d0, int t0, int t1, int C,
                                           float gamma, float alpha, float* __restrict__ out) {
  const float xs[4] = {x4.x, x4.y, x4.z, x4.w}, gs[4] = {g4.x, g4.y, g4.z, g4.w};
  float r[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const bool next = d0 + e >= C;                       // this element belongs to the following row
    const int d = next ? d0 + e - C : d0 + e, t = next ? t1 : t0;
    r[e] = BWD ? focal_bwd_fast<G2>(xs[e], t, d, gamma, alpha) * gs[e] : focal_fwd_fast<G2>(xs[e], t, d, gamma, alpha);
  }
  *reinterpret_cast<float4*>(out + i) = make_float4(r[0], r[1], r[2], r[3]);
}

template <bool BWD, bool G2>
__global__ void __launch_bounds__(256) focal_vec4_kernel(const float* __restrict__ logits,
                                                         const int64_t* __restrict__ targets,
                                                         const float* __restrict__ d_losses, int total, int C, int M,
                                                         float gamma, float alpha, float* __restrict__ out) {
  // FV_UNROLL quads per thread and iteration, all loads (logits, upstream gradient, row labels) issued before the
  // first use
  constexpr int FV_UNROLL = 4;
  const int quads = total >> 2;
  const int nthreads = blockDim.x * gridDim.x;
  const float4 one = make_float4(1.f, 1.f, 1.f, 1.f);
  for (int q0 = blockIdx.x * blockDim.x + threadIdx.x; q0 < quads; q0 += nthreads * FV_UNROLL) {
    float4 x4[FV_UNROLL], g4[FV_UNROLL];
    int d0[FV_UNROLL], t0[FV_UNROLL], t1[FV_UNROLL];
#pragma unroll
    for (int u = 0; u < FV_UNROLL; ++u) {
      const int q = q0 + u * nthreads;
      if (q < quads) {
        x4[u] = *reinterpret_cast<const float4*>(logits + (size_t)q * 4);
        g4[u] = BWD ? *reinterpret_cast<const float4*>(d_losses + (size_t)q * 4) : one;
        const int n = (q * 4) / C;
        d0[u] = q * 4 - n * C;
        t0[u] = (int)__ldg(targets + n);
        t1[u] = (int)__ldg(targets + (n + 1 < M ? n + 1 : n));
      }
    }
#pragma unroll
    for (int u = 0; u < FV_UNROLL; ++u) {
      const int q = q0 + u * nthreads;
      if (q < quads) focal_quad<BWD, G2>(x4[u], g4[u], q * 4, d0[u], t0[u], t1[u], C, gamma, alpha, out);
    }
  }
  // tail (total % 4 elements): one thread
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int j = quads * 4; j < total; ++j) {
      const int n = j / C, d = j - n * C;
      const int t = (int)__ldg(targets + n);
      out[j] = BWD ? focal_bwd_fast<G2>(logits[j], t, d, gamma, alpha) * d_losses[j]
                   : focal_fwd_fast<G2>(logits[j], t, d, gamma, alpha);
    }
  }
}

// KGDET_FOCAL_EXACT=1: the double-promotion mirror of the reference kernel for fp32 tensors too
static bool focal_exact() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("KGDET_FOCAL_EXACT");
    v = (e && atoi(e) != 0) ? 1 : 0;
  }
  return v == 1;
}

static int focal_grid4(int total) {
  int blocks = ceil_div(ceil_div(total, 4), 256);
  int cap = num_sms() * 8;
  return blocks < cap ? (blocks < 1 ? 1 : blocks) : cap;
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
  if (dtype == KGDET_F32 && C >= 4 && !focal_exact() && (((uintptr_t)logits | (uintptr_t)losses) & 15) == 0) {
    if (gamma == 2.0f)
      focal_vec4_kernel<false, true><<<focal_grid4(total), 256, 0, stream>>>((const float*)logits, targets, nullptr, total, C,
                                                                             M, gamma, alpha, (float*)losses);
    else
      focal_vec4_kernel<false, false><<<focal_grid4(total), 256, 0, stream>>>((const float*)logits, targets, nullptr, total, C,
                                                                              M, gamma, alpha, (float*)losses);
    KG_LAUNCH_CHECK("focal_vec4_kernel");
    return KGDET_OK;
  }
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
  if (dtype == KGDET_F32 && C >= 4 && !focal_exact() && (((uintptr_t)logits | (uintptr_t)d_losses | (uintptr_t)d_logits) & 15) == 0) {
    if (gamma == 2.0f)
      focal_vec4_kernel<true, true><<<focal_grid4(total), 256, 0, stream>>>((const float*)logits, targets,
                                                                            (const float*)d_losses, total, C, M, gamma,
                                                                            alpha, (float*)d_logits);
    else
      focal_vec4_kernel<true, false><<<focal_grid4(total), 256, 0, stream>>>((const float*)logits, targets,
                                                                             (const float*)d_losses, total, C, M, gamma,
                                                                             alpha, (float*)d_logits);
    KG_LAUNCH_CHECK("focal_vec4_kernel");
    return KGDET_OK;
  }
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
  const int fast = focal_exact() ? 0 : (gamma == 2.0f ? 2 : 1);
  if (dtype == KGDET_F32)
    focal_sum_fwd_kernel<float><<<focal_grid(total), 256, 0, stream>>>(
        (const float*)logits, targets, weight, total, C, gamma, alpha, loss_sum, fast);
  else
    focal_sum_fwd_kernel<__nv_bfloat16><<<focal_grid(total), 256, 0, stream>>>(
        (const __nv_bfloat16*)logits, targets, weight, total, C, gamma, alpha, loss_sum, fast);
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
  const int fast = focal_exact() ? 0 : (gamma == 2.0f ? 2 : 1);
  if (dtype == KGDET_F32)
    focal_sum_bwd_kernel<float><<<focal_grid(total), 256, 0, stream>>>(
        (const float*)logits, targets, weight, grad_scale, total, C, gamma, alpha,
        (float*)d_logits, fast);
  else
    focal_sum_bwd_kernel<__nv_bfloat16><<<focal_grid(total), 256, 0, stream>>>(
        (const __nv_bfloat16*)logits, targets, weight, grad_scale, total, C, gamma, alpha,
        (__nv_bfloat16*)d_logits, fast);
  KG_LAUNCH_CHECK("focal_sum_bwd_kernel");
  return KGDET_OK;
}
