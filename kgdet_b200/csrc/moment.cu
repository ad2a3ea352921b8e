// Point set -> bbox "moment" transform, forward and backward, one pass each.
//
// Replaces the chain of PyTorch kernels behind points2bbox(..., 'moment')
// (mmdet/models/anchor_heads/reppoints_head_kp3rep_cas_1_assign_once.py:373-388):
// view/slice, mean x2, sub x2, std x2 (unbiased), exp, mul x2, sub/add x4, cat.
// One thread owns one (n, s) position and walks its P points twice (mean, then the
// squared deviations); for NCHW maps consecutive threads read consecutive s, so every
// load is coalesced, and the second walk hits L1/L2.
#include "common.cuh"

namespace kgdet {

__global__ void moment_fwd_kernel(const float* __restrict__ pts, const float* __restrict__ mt,
                                  int N, int P, int S, int y_first, float* __restrict__ bbox) {
  const int total = N * S;
  const float ew = expf(mt[0]), eh = expf(mt[1]);      // KP3:380-383 (mt*mul + mt*(1-mul) == mt)
  const int yo = y_first ? 0 : 1, xo = 1 - yo;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += blockDim.x * gridDim.x) {
    const int n = i / S, s = i - n * S;
    const float* base = pts + (size_t)n * 2 * P * S + s;
    float sy = 0.f, sx = 0.f;
    for (int k = 0; k < P; ++k) {
      sy += base[(size_t)(2 * k + yo) * S];
      sx += base[(size_t)(2 * k + xo) * S];
    }
    const float my = sy / (float)P, mx = sx / (float)P;  // KP3:374-375
    float vy = 0.f, vx = 0.f;
    for (int k = 0; k < P; ++k) {
      float dy = base[(size_t)(2 * k + yo) * S] - my;
      float dx = base[(size_t)(2 * k + xo) * S] - mx;
      vy += dy * dy;
      vx += dx * dx;
    }
    // torch.std: unbiased, n-1 (KP3:376-377); P == 1 gives 0/0 = NaN like torch
    const float sdy = sqrtf(vy / (float)(P - 1)), sdx = sqrtf(vx / (float)(P - 1));
    const float hw = sdx * ew, hh = sdy * eh;            // KP3:382-383
    float* o = bbox + (size_t)n * 4 * S + s;             // KP3:384-388
    o[0] = mx - hw;
    o[(size_t)S] = my - hh;
    o[(size_t)2 * S] = mx + hw;
    o[(size_t)3 * S] = my + hh;
  }
}

// Large point sets on small maps (KGDet: P = 83 on 25x42, 16 800 positions per batch of 16): one thread per
// position leaves most of the machine idle and serialises 2P dependent loads.  Here a warp covers 32
// consecutive positions (coalesced) and G warps of the CTA split the P points; every thread keeps its <= MAXK
// points in registers, so the tensor is read exactly once; the partial sums meet in shared memory and are added
// in a fixed order (deterministic).
template <int G, int MAXK>
__global__ void __launch_bounds__(32 * G) moment_fwd_split_kernel(const float* __restrict__ pts,
                                                                  const float* __restrict__ mt, int N, int P, int S,
                                                                  int y_first, float* __restrict__ bbox) {
  __shared__ float red[2][2][G][32];
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  const bool ok = i < N * S;
  const int n = ok ? i / S : 0, s = ok ? i - n * S : 0;
  const int yo = y_first ? 0 : 1, xo = 1 - yo;
  const float* base = pts + (size_t)n * 2 * P * S + s;
  float y[MAXK], x[MAXK];
#pragma unroll
  for (int j = 0; j < MAXK; ++j) {
    const int k = g + j * G;
    const bool in = ok && k < P;
    y[j] = in ? __ldg(base + (size_t)(2 * k + yo) * S) : 0.f;
    x[j] = in ? __ldg(base + (size_t)(2 * k + xo) * S) : 0.f;
  }
  float sy = 0.f, sx = 0.f;
#pragma unroll
  for (int j = 0; j < MAXK; ++j) { sy += y[j]; sx += x[j]; }
  red[0][0][g][lane] = sy;
  red[0][1][g][lane] = sx;
  __syncthreads();
  sy = 0.f; sx = 0.f;
#pragma unroll
  for (int q = 0; q < G; ++q) { sy += red[0][0][q][lane]; sx += red[0][1][q][lane]; }
  const float my = sy / (float)P, mx = sx / (float)P;          // KP3:374-375
  float vy = 0.f, vx = 0.f;
#pragma unroll
  for (int j = 0; j < MAXK; ++j) {
    if (g + j * G < P) {
      const float dy = y[j] - my, dx = x[j] - mx;
      vy += dy * dy;
      vx += dx * dx;
    }
  }
  red[1][0][g][lane] = vy;
  red[1][1][g][lane] = vx;
  __syncthreads();
  if (g == 0 && ok) {
    vy = 0.f; vx = 0.f;
#pragma unroll
    for (int q = 0; q < G; ++q) { vy += red[1][0][q][lane]; vx += red[1][1][q][lane]; }
    const float ew = expf(mt[0]), eh = expf(mt[1]);
    const float sdy = sqrtf(vy / (float)(P - 1)), sdx = sqrtf(vx / (float)(P - 1));   // unbiased, KP3:376-377
    const float hw = sdx * ew, hh = sdy * eh;
    float* o = bbox + (size_t)n * 4 * S + s;
    o[0] = mx - hw;
    o[(size_t)S] = my - hh;
    o[(size_t)2 * S] = mx + hw;
    o[(size_t)3 * S] = my + hh;
  }
}

__global__ void moment_bwd_kernel(const float* __restrict__ pts, const float* __restrict__ mt,
                                  const float* __restrict__ gbox, int N, int P, int S, int y_first,
                                  float moment_mul, float* __restrict__ gpts,
                                  float* __restrict__ gmt) {
  const int total = N * S;
  const float ew = expf(mt[0]), eh = expf(mt[1]);
  const int yo = y_first ? 0 : 1, xo = 1 - yo;
  float acc_w = 0.f, acc_h = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += blockDim.x * gridDim.x) {
    const int n = i / S, s = i - n * S;
    const float* base = pts + (size_t)n * 2 * P * S + s;
    float sy = 0.f, sx = 0.f;
    for (int k = 0; k < P; ++k) {
      sy += base[(size_t)(2 * k + yo) * S];
      sx += base[(size_t)(2 * k + xo) * S];
    }
    const float my = sy / (float)P, mx = sx / (float)P;
    float vy = 0.f, vx = 0.f;
    for (int k = 0; k < P; ++k) {
      float dy = base[(size_t)(2 * k + yo) * S] - my;
      float dx = base[(size_t)(2 * k + xo) * S] - mx;
      vy += dy * dy;
      vx += dx * dx;
    }
    const float sdy = sqrtf(vy / (float)(P - 1)), sdx = sqrtf(vx / (float)(P - 1));
    const float* g = gbox + (size_t)n * 4 * S + s;
    const float g0 = g[0], g1 = g[(size_t)S], g2 = g[(size_t)2 * S], g3 = g[(size_t)3 * S];
    const float g_mx = g0 + g2, g_my = g1 + g3;
    const float g_hw = g2 - g0, g_hh = g3 - g1;          // d/d(half width), d/d(half height)
    acc_w += g_hw * sdx * ew;                            // d/dt_w of sdx * exp(t_w)
    acc_h += g_hh * sdy * eh;
    const float g_sdx = g_hw * ew, g_sdy = g_hh * eh;
    // d std / d x_i = (x_i - mean) / ((P-1) * std); a collapsed point set (std == 0) gets a ZERO gradient, as
    // current PyTorch's std_backward does (masked_fill(result == 0, 0)); the torch 1.x the reference targeted
    // produced 0/0 = NaN there
    const float cx = sdx == 0.f ? 0.f : g_sdx / ((float)(P - 1) * sdx);
    const float cy = sdy == 0.f ? 0.f : g_sdy / ((float)(P - 1) * sdy);
    const float mxP = g_mx / (float)P, myP = g_my / (float)P;
    float* go = gpts + (size_t)n * 2 * P * S + s;
    for (int k = 0; k < P; ++k) {
      float y = base[(size_t)(2 * k + yo) * S], x = base[(size_t)(2 * k + xo) * S];
      go[(size_t)(2 * k + yo) * S] = myP + cy * (y - my);
      go[(size_t)(2 * k + xo) * S] = mxP + cx * (x - mx);
    }
  }
  // block reduce the two moment_transfer partials, one atomic pair per CTA
  __shared__ float pw[32], ph[32];
  acc_w = warp_sum(acc_w);
  acc_h = warp_sum(acc_h);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { pw[wid] = acc_w; ph[wid] = acc_h; }
  __syncthreads();
  if (wid == 0) {
    float a = (lane < (blockDim.x >> 5)) ? pw[lane] : 0.f;
    float b = (lane < (blockDim.x >> 5)) ? ph[lane] : 0.f;
    a = warp_sum(a);
    b = warp_sum(b);
    if (lane == 0 && gmt) {
      atomicAdd(&gmt[0], a * moment_mul);                // KP3:378-379: only the mul-scaled branch
      atomicAdd(&gmt[1], b * moment_mul);                // carries gradient
    }
  }
}

// Backward twin of moment_fwd_split_kernel: a warp covers 32 consecutive positions, G warps split the P points, the
// values stay in registers (one read of the point tensor), partial sums meet in shared memory in a fixed order.
// The two moment_transfer partials of a CTA are reduced over the position lanes and leave as one atomic pair.
template <int G, int MAXK>
__global__ void __launch_bounds__(32 * G) moment_bwd_split_kernel(const float* __restrict__ pts, const float* __restrict__ mt,
                                                                  const float* __restrict__ gbox, int N, int P, int S,
                                                                  int y_first, float moment_mul, float* __restrict__ gpts,
                                                                  float* __restrict__ gmt) {
  __shared__ float red[2][2][G][32];
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  const bool ok = i < N * S;
  const int n = ok ? i / S : 0, s = ok ? i - n * S : 0;
  const int yo = y_first ? 0 : 1, xo = 1 - yo;
  const float* base = pts + (size_t)n * 2 * P * S + s;
  float y[MAXK], x[MAXK];
#pragma unroll
  for (int j = 0; j < MAXK; ++j) {
    const int k = g + j * G;
    const bool in = ok && k < P;
    y[j] = in ? __ldg(base + (size_t)(2 * k + yo) * S) : 0.f;
    x[j] = in ? __ldg(base + (size_t)(2 * k + xo) * S) : 0.f;
  }
  float sy = 0.f, sx = 0.f;
#pragma unroll
  for (int j = 0; j < MAXK; ++j) { sy += y[j]; sx += x[j]; }
  red[0][0][g][lane] = sy;
  red[0][1][g][lane] = sx;
  __syncthreads();
  sy = 0.f; sx = 0.f;
#pragma unroll
  for (int q = 0; q < G; ++q) { sy += red[0][0][q][lane]; sx += red[0][1][q][lane]; }
  const float my = sy / (float)P, mx = sx / (float)P;
  float vy = 0.f, vx = 0.f;
#pragma unroll
  for (int j = 0; j < MAXK; ++j) {
    if (g + j * G < P) {
      const float dy = y[j] - my, dx = x[j] - mx;
      vy += dy * dy;
      vx += dx * dx;
    }
  }
  red[1][0][g][lane] = vy;
  red[1][1][g][lane] = vx;
  __syncthreads();
  vy = 0.f; vx = 0.f;
#pragma unroll
  for (int q = 0; q < G; ++q) { vy += red[1][0][q][lane]; vx += red[1][1][q][lane]; }
  const float ew = expf(mt[0]), eh = expf(mt[1]);
  const float sdy = sqrtf(vy / (float)(P - 1)), sdx = sqrtf(vx / (float)(P - 1));
  float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f;
  if (ok) {
    const float* gb = gbox + (size_t)n * 4 * S + s;
    g0 = gb[0]; g1 = gb[(size_t)S]; g2 = gb[(size_t)2 * S]; g3 = gb[(size_t)3 * S];
  }
  const float g_mx = g0 + g2, g_my = g1 + g3, g_hw = g2 - g0, g_hh = g3 - g1;
  const float g_sdx = g_hw * ew, g_sdy = g_hh * eh;
  // zero gradient through a collapsed point set, as torch.std's backward (see moment_bwd_kernel)
  const float cx = sdx == 0.f ? 0.f : g_sdx / ((float)(P - 1) * sdx);
  const float cy = sdy == 0.f ? 0.f : g_sdy / ((float)(P - 1) * sdy);
  const float mxP = g_mx / (float)P, myP = g_my / (float)P;
  float* go = gpts + (size_t)n * 2 * P * S + s;
#pragma unroll
  for (int j = 0; j < MAXK; ++j) {
    const int k = g + j * G;
    if (ok && k < P) {
      go[(size_t)(2 * k + yo) * S] = myP + cy * (y[j] - my);
      go[(size_t)(2 * k + xo) * S] = mxP + cx * (x[j] - mx);
    }
  }
  if (g == 0 && gmt) {
    float aw = ok ? g_hw * sdx * ew : 0.f, ah = ok ? g_hh * sdy * eh : 0.f;
    aw = warp_sum(aw);
    ah = warp_sum(ah);
    if (lane == 0) {
      atomicAdd(&gmt[0], aw * moment_mul);
      atomicAdd(&gmt[1], ah * moment_mul);
    }
  }
}

}  // namespace kgdet

using namespace kgdet;

static int moment_grid(int total) {
  int blocks = ceil_div(total, 128);
  int cap = num_sms() * 8;
  return blocks < 1 ? 1 : (blocks < cap ? blocks : cap);
}

extern "C" int kgdet_points2bbox_moment_forward(const float* pts, const float* moment_transfer,
                                                int32_t N, int32_t P, int32_t S, int y_first,
                                                float* bbox, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  KG_CHECK_ARG(N >= 0 && P >= 1 && S >= 1, "kgdet_points2bbox_moment_forward: bad sizes");
  KG_CHECK_ARG((long long)N * S * 2 * P < (1ll << 40), "kgdet_points2bbox_moment_forward: too big");
  if (N == 0) return KGDET_OK;
  KG_CHECK_ARG(pts && moment_transfer && bbox, "kgdet_points2bbox_moment_forward: NULL pointer");
  KG_CHECK_ARG((long long)N * S < (1ll << 31), "kgdet_points2bbox_moment_forward: N*S overflow");
  // split the points over the warps of a CTA when one thread per position cannot fill the machine
  constexpr int G = 8, MAXK = 16;
  if (S >= 32 && P >= 2 * G && P <= G * MAXK && (long long)N * S < 128ll * 8 * num_sms()) {
    moment_fwd_split_kernel<G, MAXK><<<ceil_div(N * S, 32), 32 * G, 0, stream>>>(pts, moment_transfer, N, P, S,
                                                                                y_first, bbox);
    KG_LAUNCH_CHECK("moment_fwd_split_kernel");
    return KGDET_OK;
  }
  moment_fwd_kernel<<<moment_grid(N * S), 128, 0, stream>>>(pts, moment_transfer, N, P, S, y_first,
                                                            bbox);
  KG_LAUNCH_CHECK("moment_fwd_kernel");
  return KGDET_OK;
}

extern "C" int kgdet_points2bbox_moment_backward(const float* pts, const float* moment_transfer,
                                                 const float* grad_bbox, int32_t N, int32_t P,
                                                 int32_t S, int y_first, float moment_mul,
                                                 float* grad_pts, float* grad_moment_transfer,
                                                 void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  KG_CHECK_ARG(N >= 0 && P >= 1 && S >= 1, "kgdet_points2bbox_moment_backward: bad sizes");
  if (N == 0) return KGDET_OK;
  KG_CHECK_ARG(pts && moment_transfer && grad_bbox && grad_pts,
               "kgdet_points2bbox_moment_backward: NULL pointer");
  KG_CHECK_ARG((long long)N * S < (1ll << 31), "kgdet_points2bbox_moment_backward: N*S overflow");
  constexpr int G = 8, MAXK = 16;
  if (S >= 32 && P >= 2 * G && P <= G * MAXK && (long long)N * S < 128ll * 8 * num_sms()) {
    moment_bwd_split_kernel<G, MAXK><<<ceil_div(N * S, 32), 32 * G, 0, stream>>>(pts, moment_transfer, grad_bbox, N, P, S,
                                                                                y_first, moment_mul, grad_pts,
                                                                                grad_moment_transfer);
    KG_LAUNCH_CHECK("moment_bwd_split_kernel");
    return KGDET_OK;
  }
  moment_bwd_kernel<<<moment_grid(N * S), 128, 0, stream>>>(pts, moment_transfer, grad_bbox, N, P,
                                                            S, y_first, moment_mul, grad_pts,
                                                            grad_moment_transfer);
  KG_LAUNCH_CHECK("moment_bwd_kernel");
  return KGDET_OK;
}
