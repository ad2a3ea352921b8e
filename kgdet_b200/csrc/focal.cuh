// Sigmoid focal loss element functions shared by focal_loss.cu and point_loss.cu.
// Arithmetic of mmdet/ops/sigmoid_focal_loss/src/sigmoid_focal_loss_cuda.cu:24-98 with the reference's
// float/double promotion pattern for scalar_t = float (line references in the comments below).
#pragma once
#include <cfloat>

#include "common.cuh"

namespace kgdet {

struct FocalTerms { float c1, c2, zn, zp, p; double log1m; };

__device__ __forceinline__ FocalTerms focal_terms(float x, int t, int d, float alpha) {
  FocalTerms f;
  f.c1 = (t == (d + 1)) ? 1.f : 0.f;                       // :36
  f.c2 = ((t >= 0) & (t != (d + 1))) ? 1.f : 0.f;          // :37
  f.zn = (float)(1.0 - (double)alpha);                     // :39
  f.zp = alpha;                                            // :40
  f.p = (float)(1. / (1. + (double)expf(-x)));             // :43
  double ge = (x >= 0) ? 1.0 : 0.0;
  // -x*[x>=0] - logf(1 + expf(x - 2x[x>=0]))              // :51-52
  float e = expf((float)((double)x - 2. * (double)x * ge));
  f.log1m = -1. * (double)x * ge - (double)logf((float)(1. + (double)e));
  return f;
}

__device__ __forceinline__ float focal_fwd_value(float x, int t, int d, float gamma, float alpha) {
  FocalTerms f = focal_terms(x, t, d, alpha);
  float term1 = powf((float)(1. - (double)f.p), gamma) * logf(fmaxf(f.p, FLT_MIN));   // :46
  float term2 = (float)((double)powf(f.p, gamma) * f.log1m);                          // :49-52
  float loss = 0.f;
  loss += -f.c1 * term1 * f.zp;                                                       // :55
  loss += -f.c2 * term2 * f.zn;                                                       // :56
  return loss;
}

__device__ __forceinline__ float focal_bwd_value(float x, int t, int d, float gamma, float alpha) {
  FocalTerms f = focal_terms(x, t, d, alpha);
  double p = (double)f.p;
  // (1-p)^g * (1 - p - p*g*log p)                                                    // :81-82
  float term1 = (float)((double)powf((float)(1. - p), gamma) *
                        (1. - p - (double)(f.p * gamma * logf(fmaxf(f.p, FLT_MIN)))));
  // p^g * (log(1-p)*(1-p)*g - p)                                                     // :85-90
  float term2 = (float)((double)powf(f.p, gamma) * (f.log1m * (1. - p) * (double)gamma - p));
  float g = 0.f;
  g += -f.c1 * term1 * f.zp;
  g += -f.c2 * term2 * f.zn;
  return g;
}

// ---- single-precision variants ------------------------------------------------------------------------------
// The reference's `1.` literals promote parts of the expression to double (the functions above mirror that bit for
// bit); on this GPU the double adds / multiplies / division cost 7x the rest and hold the kernel at 0.12 of the HBM
// roofline.  These evaluate the same quantities in fp32 from ONE exponential and ONE logarithm per element:
//     e = exp(-|x|), s = 1 + e, L = log(s)
//     p = 1 / s  (x >= 0)  or  e / s  (x < 0);     1 - p = e / s  or  1 / s   (no cancellation)
//     log p = min(x, 0) - L  (clamped at log(FLT_MIN), :46);      log(1 - p) = -max(x, 0) - L   (:51-52)
// and gamma == 2 -- every reference config -- as a multiplication instead of powf.  Against the reference CUDA
// kernel: <= 3e-7 of the tensor's maximum (the bar is 1e-5; tests/test_focal_moment_gpu.py); element-wise the
// difference is where the reference's own 1 - p cancels.
template <bool G2>
__device__ __forceinline__ float focal_pow(float v, float gamma) { return G2 ? v * v : powf(v, gamma); }

struct FocalFast { float p, q, logp, log1m; };

__device__ __forceinline__ FocalFast focal_fast_terms(float x) {
  // e in (0, 1], s in (1, 2]: the hardware exp2 / log2 / reciprocal approximations are at their best here (2 ulp;
  // -|x| * log2(e) adds |x| * 6e-8 of relative error to e, 1e-6 at |x| = 16 where e itself is 1e-7 of the result)
  const float e = __expf(-fabsf(x));
  const float s = 1.f + e;
  const float inv = __fdividef(1.f, s);
  const float L = __logf(s);
  const bool pos = x >= 0.f;
  FocalFast f;
  f.p = pos ? inv : e * inv;
  f.q = pos ? e * inv : inv;
  f.logp = fmaxf(fminf(x, 0.f) - L, -87.33654475f);        // logf(fmaxf(p, FLT_MIN))
  f.log1m = -fmaxf(x, 0.f) - L;
  return f;
}

template <bool G2>
__device__ __forceinline__ float focal_fwd_fast(float x, int t, int d, float gamma, float alpha) {
  if (t < 0) return 0.f;
  const FocalFast f = focal_fast_terms(x);
  return t == d + 1 ? -alpha * (focal_pow<G2>(f.q, gamma) * f.logp)                    // :46,:55
                    : -(1.f - alpha) * (focal_pow<G2>(f.p, gamma) * f.log1m);          // :49,:56
}

template <bool G2>
__device__ __forceinline__ float focal_bwd_fast(float x, int t, int d, float gamma, float alpha) {
  if (t < 0) return 0.f;
  const FocalFast f = focal_fast_terms(x);
  return t == d + 1 ? -alpha * (focal_pow<G2>(f.q, gamma) * (f.q - f.p * gamma * f.logp))               // :81-82
                    : -(1.f - alpha) * (focal_pow<G2>(f.p, gamma) * (f.log1m * f.q * gamma - f.p));     // :85-90
}

}  // namespace kgdet
