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

}  // namespace kgdet
