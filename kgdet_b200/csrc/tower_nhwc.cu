// Position-major (NHWC) glue of the head towers: the towers' 3x3 convolutions
// stay cuDNN but run in channels_last, where they need no layout transposes; what this file adds is
//   * GroupNorm + ReLU fused, on NHWC fp32 (mmdet ConvModule order conv -> norm -> activation,
//     mmdet/models/utils/conv_module.py:156-164; torch.nn.GroupNorm semantics: biased variance, eps inside the
//     square root),
//   * NHWC fp32 rows -> the channel-blocked bf16/fp32 planes the fused DCN kernel gathers from,
//   * NHWC fp32 rows -> UMMA-tiled bf16 rows (optionally ReLU, optionally [hi | lo]) for the pointwise GEMM.
#include <cuda_bf16.h>

#include "dcn.cuh"

namespace kgdet {

// one CTA per (image, group); the group's HW x cpg values are staged in shared memory, so the input is read
// once: mean, then sum of squared deviations (no E[x^2] - mean^2 cancellation), then normalise + ReLU
__global__ void __launch_bounds__(256) groupnorm_relu_nhwc_kernel(const float* __restrict__ x,
                                                                  const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, float eps,
                                                                  float* __restrict__ y, int HW, int C, int cpg,
                                                                  int relu) {
  extern __shared__ float sm[];                 // [HW * cpg] + 32 reduction slots
  float* red = sm + (size_t)HW * cpg;
  const int n = blockIdx.y, g = blockIdx.x;
  const float* xg = x + (size_t)n * HW * C + (size_t)g * cpg;
  float* yg = y + (size_t)n * HW * C + (size_t)g * cpg;
  const int total = HW * cpg;
  const int vec = cpg / 4;                      // cpg % 4 == 0: float4 per (pixel, quarter-group)
  float s = 0.f;
  for (int i = threadIdx.x; i < HW * vec; i += blockDim.x) {
    const int p = i / vec, q = i - p * vec;
    const float4 v = *reinterpret_cast<const float4*>(xg + (size_t)p * C + q * 4);
    *reinterpret_cast<float4*>(sm + (size_t)p * cpg + q * 4) = v;
    s += (v.x + v.y) + (v.z + v.w);
  }
  auto block_sum = [&](float v) -> float {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    return t;
  };
  const float mean = block_sum(s) / (float)total;
  float ss = 0.f;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const float d = sm[i] - mean;
    ss = fmaf(d, d, ss);
  }
  const float var = block_sum(ss) / (float)total;
  const float rstd = rsqrtf(var + eps);
  for (int i = threadIdx.x; i < HW * vec; i += blockDim.x) {
    const int p = i / vec, q = i - p * vec;
    const float4 v = *reinterpret_cast<const float4*>(sm + (size_t)p * cpg + q * 4);
    const int c = g * cpg + q * 4;
    float4 o;
    o.x = (v.x - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
    o.y = (v.y - mean) * rstd * __ldg(gamma + c + 1) + __ldg(beta + c + 1);
    o.z = (v.z - mean) * rstd * __ldg(gamma + c + 2) + __ldg(beta + c + 2);
    o.w = (v.w - mean) * rstd * __ldg(gamma + c + 3) + __ldg(beta + c + 3);
    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    *reinterpret_cast<float4*>(yg + (size_t)p * C + q * 4) = o;
  }
}

// Wide variant: one CTA per (image, 32 consecutive channels) = 8 / VEC groups, VEC = float4 chunks per group.
// (Measured and dropped: splitting the unit over a cluster of two CTAs with a DSMEM exchange of the partial sums --
// 67 KB instead of 134 KB of shared memory, two CTAs per SM -- is slower, 13.1 vs 11.4 us per [16,256,25,42] call.)
// The one-group kernel above reads 4 * cpg bytes per pixel (32 B for the KGDet towers: 16 different 128-byte
// lines per warp request); here 8 lanes cover one pixel's 128-byte line, 1024 threads keep ~8 float4 loads each in
// flight, and the statistics of a group are reduced with xor-shuffles over the lanes that share it.
template <int VEC>
__global__ void __launch_bounds__(1024) groupnorm_relu_nhwc_wide_kernel(const float* __restrict__ x,
                                                                        const float* __restrict__ gamma,
                                                                        const float* __restrict__ beta, float eps,
                                                                        float* __restrict__ y, int HW, int C, int relu,
                                                                        unsigned char* __restrict__ hi,
                                                                        unsigned char* __restrict__ lo,
                                                                        size_t plane_bytes, int guard_bytes,
                                                                        int tail_bytes) {
  // y (NHWC fp32) and / or hi + lo ("split planes" of conv_umma.cu: channel-blocked bf16 planes of the values'
  // bf16 hi parts -- the fused DCN kernel's prepared-input layout -- and of the lo parts x - hi) may be NULL
  constexpr int GPC = 8 / VEC;                  // groups per CTA
  // guard bands of the planes (zero padding the deformable gather reads): written here by the CTAs of image 0 -- the
  // even 32-channel block of a plane clears the band in front of it, the odd one the band behind it -- instead of by
  // four memset nodes per call in front of the kernel (5 us each time on the critical path of the captured step)
  if (hi && guard_bytes > 0 && blockIdx.y == 0) {
    const int plane = blockIdx.x >> 1;
    const bool front = (blockIdx.x & 1) == 0;
    const size_t in_bytes = (size_t)gridDim.y * HW * 128;
    const size_t off = (size_t)plane * plane_bytes + (front ? 0 : (size_t)guard_bytes + in_bytes);
    const int nbytes = front ? guard_bytes : tail_bytes;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (int i = threadIdx.x * 16; i < nbytes; i += blockDim.x * 16) {
      *reinterpret_cast<uint4*>(hi - guard_bytes + off + i) = z;
      if (lo) *reinterpret_cast<uint4*>(lo - guard_bytes + off + i) = z;
    }
  }
  extern __shared__ float sm[];                 // [HW][32] values
  __shared__ float red[32][GPC];
  const int n = blockIdx.y, c0 = blockIdx.x * 32;
  const int chunk = threadIdx.x & 7, gi = chunk / VEC, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* xg = x + (size_t)n * HW * C + c0 + chunk * 4;
  float* yg = y ? y + (size_t)n * HW * C + c0 + chunk * 4 : nullptr;
  const size_t poff = (size_t)((c0 + chunk * 4) >> 6) * plane_bytes + (size_t)n * HW * 128 + (size_t)((c0 + chunk * 4) & 63) * 2;
  const int prow = threadIdx.x >> 3, pstep = blockDim.x >> 3;
  const float inv_total = 1.f / (float)(HW * VEC * 4);
  // sum over the lanes of this warp that hold the same group (other chunks of the group, other pixels), then
  // over the warps through shared memory; every thread ends up with its own group's total
  auto group_sum = [&](float v) -> float {
#pragma unroll
    for (int o = 1; o < VEC; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    __syncthreads();                            // red[] may still be read from the previous reduction
    if (lane < 8 && (lane % VEC) == 0) red[warp][gi] = v;
    __syncthreads();
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w][gi];
    return t;
  };
  float s = 0.f;
  for (int p = prow; p < HW; p += pstep) {
    const float4 v = *reinterpret_cast<const float4*>(xg + (size_t)p * C);
    *reinterpret_cast<float4*>(sm + (size_t)p * 32 + chunk * 4) = v;
    s += (v.x + v.y) + (v.z + v.w);
  }
  const float mean = group_sum(s) * inv_total;
  float ss = 0.f;
  for (int p = prow; p < HW; p += pstep) {
    const float4 v = *reinterpret_cast<const float4*>(sm + (size_t)p * 32 + chunk * 4);
    const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
    ss = fmaf(d0, d0, ss); ss = fmaf(d1, d1, ss); ss = fmaf(d2, d2, ss); ss = fmaf(d3, d3, ss);
  }
  const float rstd = rsqrtf(group_sum(ss) * inv_total + eps);
  const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + c0 + chunk * 4));
  const float4 be = __ldg(reinterpret_cast<const float4*>(beta + c0 + chunk * 4));
  for (int p = prow; p < HW; p += pstep) {
    const float4 v = *reinterpret_cast<const float4*>(sm + (size_t)p * 32 + chunk * 4);
    float4 o;
    o.x = (v.x - mean) * rstd * ga.x + be.x;
    o.y = (v.y - mean) * rstd * ga.y + be.y;
    o.z = (v.z - mean) * rstd * ga.z + be.z;
    o.w = (v.w - mean) * rstd * ga.w + be.w;
    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    if (yg) *reinterpret_cast<float4*>(yg + (size_t)p * C) = o;
    if (hi) {
      const __nv_bfloat162 h0 = __floats2bfloat162_rn(o.x, o.y), h1 = __floats2bfloat162_rn(o.z, o.w);
      const __nv_bfloat162 l0 = __floats2bfloat162_rn(o.x - __low2float(h0), o.y - __high2float(h0));
      const __nv_bfloat162 l1 = __floats2bfloat162_rn(o.z - __low2float(h1), o.w - __high2float(h1));
      uint2 hv, lv;
      hv.x = *reinterpret_cast<const uint32_t*>(&h0); hv.y = *reinterpret_cast<const uint32_t*>(&h1);
      lv.x = *reinterpret_cast<const uint32_t*>(&l0); lv.y = *reinterpret_cast<const uint32_t*>(&l1);
      *reinterpret_cast<uint2*>(hi + poff + (size_t)p * 128) = hv;
      *reinterpret_cast<uint2*>(lo + poff + (size_t)p * 128) = lv;
    }
  }
}

template <int VEC>
static int launch_groupnorm_wide(const float* x, const float* gamma, const float* beta, float eps, float* y, int N,
                                 int HW, int C, int relu, cudaStream_t stream, unsigned char* hi = nullptr,
                                 unsigned char* lo = nullptr, size_t plane_bytes = 0, int guard_bytes = 0,
                                 int tail_bytes = 0) {
  const size_t smem = (size_t)HW * 32 * sizeof(float);
  KG_CUDA(cudaFuncSetAttribute(groupnorm_relu_nhwc_wide_kernel<VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)smem));
  groupnorm_relu_nhwc_wide_kernel<VEC><<<dim3(C / 32, N), 1024, smem, stream>>>(x, gamma, beta, eps, y, HW, C, relu, hi, lo,
                                                                                plane_bytes, guard_bytes, tail_bytes);
  KG_LAUNCH_CHECK("groupnorm_relu_nhwc_wide_kernel");
  return KGDET_OK;
}

// ---- maps that do not fit shared memory (FPN levels P3 / P4 of the RepPoints-Kp heads: 16 800 / 4 200 positions) ----
// Three streaming kernels over chunks of GS_CHUNK_PX positions x all C channels (full 4 C-byte rows, coalesced):
//   stats     per (image, chunk, group): count, mean and sum of squared deviations.  Every thread accumulates
//             sum(x - K) and sum((x - K)^2) around K = its first value (no E[x^2] - mean^2 cancellation: K is a
//             sample, so (mean - K)^2 is never large against the spread); threads are merged with Chan's update;
//   finalize  per image, one warp per group: the chunk moments merged in a fixed order -> mean, 1 / std;
//   apply     normalise (+ ReLU) into NHWC fp32 and / or split planes, pure streaming.
// 3 x 4 bytes per value of HBM traffic instead of the 2 x 4 of the resident kernels above.
constexpr int GS_THREADS = 256;
constexpr int GS_SWEEPS = 16;                     // positions per thread per chunk

struct Moments { float n, mean, m2; };

__device__ __forceinline__ Moments merge_moments(Moments a, Moments b) {
  const float n = a.n + b.n;
  if (n == 0.f) return a;
  const float d = b.mean - a.mean, f = b.n / n;
  Moments r;
  r.n = n;
  r.mean = fmaf(d, f, a.mean);
  r.m2 = a.m2 + b.m2 + d * d * a.n * f;
  return r;
}

__device__ __forceinline__ Moments shfl_xor_moments(Moments m, int o) {
  Moments b;
  b.n = __shfl_xor_sync(0xffffffffu, m.n, o);
  b.mean = __shfl_xor_sync(0xffffffffu, m.mean, o);
  b.m2 = __shfl_xor_sync(0xffffffffu, m.m2, o);
  return b;
}

__global__ void __launch_bounds__(GS_THREADS) groupnorm_stream_stats_kernel(const float* __restrict__ x, int HW, int C,
                                                                             int cpg, float* __restrict__ partial) {
  extern __shared__ float red[];                  // [rows][groups][3]
  const int L = C >> 2, rows = GS_THREADS / L, chunk_px = rows * GS_SWEEPS;
  const int n = blockIdx.y, chunk = blockIdx.x, groups = C / cpg;
  const int col = threadIdx.x % L, row = threadIdx.x / L;
  const int p0 = chunk * chunk_px + row;
  const float* xg = x + (size_t)n * HW * C + col * 4;
  float K = 0.f, s = 0.f, ss = 0.f;
  int cnt = 0;
  if (p0 < HW) K = __ldg(xg + (size_t)p0 * C);
#pragma unroll 8
  for (int i = 0; i < GS_SWEEPS; ++i) {
    const int p = p0 + i * rows;
    if (p < HW) {
      const float4 v = *reinterpret_cast<const float4*>(xg + (size_t)p * C);
      const float d0 = v.x - K, d1 = v.y - K, d2 = v.z - K, d3 = v.w - K;
      s += (d0 + d1) + (d2 + d3);
      ss = fmaf(d0, d0, ss); ss = fmaf(d1, d1, ss); ss = fmaf(d2, d2, ss); ss = fmaf(d3, d3, ss);
      ++cnt;
    }
  }
  Moments m;
  m.n = 4.f * (float)cnt;
  m.mean = cnt ? K + s / m.n : 0.f;
  m.m2 = cnt ? fmaxf(ss - s * s / m.n, 0.f) : 0.f;
  // lanes of one group are adjacent (cpg / 4 of them, a power of two, inside one warp)
  for (int o = 1; o < (cpg >> 2); o <<= 1) {
    const Moments b = shfl_xor_moments(m, o);
    m = (threadIdx.x & o) ? merge_moments(b, m) : merge_moments(m, b);   // same operand order on both lanes
  }
  const int g = (col * 4) / cpg;
  if ((col * 4) % cpg == 0) {
    float* r = red + ((size_t)row * groups + g) * 3;
    r[0] = m.n; r[1] = m.mean; r[2] = m.m2;
  }
  __syncthreads();
  for (int gi = threadIdx.x; gi < groups; gi += GS_THREADS) {
    Moments a;
    a.n = red[gi * 3]; a.mean = red[gi * 3 + 1]; a.m2 = red[gi * 3 + 2];
    for (int r = 1; r < rows; ++r) {
      Moments b;
      const float* q = red + ((size_t)r * groups + gi) * 3;
      b.n = q[0]; b.mean = q[1]; b.m2 = q[2];
      a = merge_moments(a, b);
    }
    float* out = partial + (((size_t)n * gridDim.x + chunk) * groups + gi) * 3;
    out[0] = a.n; out[1] = a.mean; out[2] = a.m2;
  }
}

// one CTA per image, one warp per group: lane l merges chunks l, l + 32, ... in order, then a fixed xor tree
__global__ void __launch_bounds__(1024) groupnorm_stream_finalize_kernel(const float* __restrict__ partial, int nchunks,
                                                                         int groups, float eps, float* __restrict__ stat) {
  const int n = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int g = warp; g < groups; g += 32) {
    const float* q = partial + ((size_t)n * nchunks * groups + g) * 3;
    Moments a;
    a.n = 0.f; a.mean = 0.f; a.m2 = 0.f;
    for (int c0 = lane; c0 < nchunks; c0 += 32 * 4) {          // four loads in flight, merged in chunk order
      Moments b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = c0 + 32 * u;
        const float* r = q + (size_t)(c < nchunks ? c : 0) * groups * 3;
        b[u].n = c < nchunks ? r[0] : 0.f;                      // an empty set: merging it changes nothing
        b[u].mean = c < nchunks ? r[1] : 0.f;
        b[u].m2 = c < nchunks ? r[2] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) a = merge_moments(a, b[u]);
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const Moments b = shfl_xor_moments(a, o);
      a = (lane & o) ? merge_moments(b, a) : merge_moments(a, b);
    }
    if (lane == 0) {
      stat[((size_t)n * groups + g) * 2] = a.mean;
      stat[((size_t)n * groups + g) * 2 + 1] = rsqrtf(a.m2 / a.n + eps);
    }
  }
}

__global__ void __launch_bounds__(GS_THREADS) groupnorm_stream_apply_kernel(
    const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
    const float* __restrict__ stat, float* __restrict__ y, int HW, int C, int cpg, int relu,
    unsigned char* __restrict__ hi, unsigned char* __restrict__ lo, size_t plane_bytes) {
  const int L = C >> 2, rows = GS_THREADS / L, chunk_px = rows * GS_SWEEPS;
  const int n = blockIdx.y, chunk = blockIdx.x, groups = C / cpg;
  const int col = threadIdx.x % L, row = threadIdx.x / L, c0 = col * 4;
  const float2 st = __ldg(reinterpret_cast<const float2*>(stat) + (size_t)n * groups + c0 / cpg);
  const float mean = st.x, rstd = st.y;
  const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + c0));
  const float4 be = __ldg(reinterpret_cast<const float4*>(beta + c0));
  const float* xg = x + (size_t)n * HW * C + c0;
  float* yg = y ? y + (size_t)n * HW * C + c0 : nullptr;
  const size_t poff = (size_t)(c0 >> 6) * plane_bytes + (size_t)n * HW * 128 + (size_t)(c0 & 63) * 2;
#pragma unroll 8
  for (int i = 0; i < GS_SWEEPS; ++i) {
    const int p = chunk * chunk_px + row + i * rows;
    if (p < HW) {
      const float4 v = *reinterpret_cast<const float4*>(xg + (size_t)p * C);
      float4 o;
      o.x = (v.x - mean) * rstd * ga.x + be.x;
      o.y = (v.y - mean) * rstd * ga.y + be.y;
      o.z = (v.z - mean) * rstd * ga.z + be.z;
      o.w = (v.w - mean) * rstd * ga.w + be.w;
      if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
      if (yg) *reinterpret_cast<float4*>(yg + (size_t)p * C) = o;
      if (hi) {
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(o.x, o.y), h1 = __floats2bfloat162_rn(o.z, o.w);
        uint2 hv;
        hv.x = *reinterpret_cast<const uint32_t*>(&h0); hv.y = *reinterpret_cast<const uint32_t*>(&h1);
        *reinterpret_cast<uint2*>(hi + poff + (size_t)p * 128) = hv;
        if (lo) {
          const __nv_bfloat162 l0 = __floats2bfloat162_rn(o.x - __low2float(h0), o.y - __high2float(h0));
          const __nv_bfloat162 l1 = __floats2bfloat162_rn(o.z - __low2float(h1), o.w - __high2float(h1));
          uint2 lv;
          lv.x = *reinterpret_cast<const uint32_t*>(&l0); lv.y = *reinterpret_cast<const uint32_t*>(&l1);
          *reinterpret_cast<uint2*>(lo + poff + (size_t)p * 128) = lv;
        }
      }
    }
  }
}

static bool stream_gn_supported(int C, int groups) {
  if (C <= 0 || groups <= 0 || C % groups) return false;
  const int cpg = C / groups, L = C / 4;
  return (C % 128 == 0) && L <= GS_THREADS && GS_THREADS % L == 0 && cpg % 4 == 0 && cpg <= 128 &&
         ((cpg / 4) & (cpg / 4 - 1)) == 0 && groups <= 256;
}

static int stream_gn_chunks(int HW, int C) { return ceil_div(HW, (GS_THREADS / (C / 4)) * GS_SWEEPS); }

// Backward of GroupNorm (+ ReLU) on NHWC fp32, same CTA shape as the wide forward kernel (image x 32 channels, the
// map's x values resident in shared memory; dy is read twice, the second time from L2):
//   z = (x - mean) * rstd * gamma + beta,  y = relu(z);   dz = dy * [z > 0]
//   dgamma_c = sum dz * xhat,  dbeta_c = sum dz           (per image here: [N, C] partials, summed by the caller)
//   dx = rstd * (dz * gamma - (A + xhat * B) / m),   A = sum_group dz * gamma,  B = sum_group dz * gamma * xhat
// (torch.nn.functional.group_norm's backward, ATen/native/cuda/group_norm_kernel.cu, with the ReLU mask folded in).
template <int VEC>
__global__ void __launch_bounds__(1024) groupnorm_relu_bwd_nhwc_wide_kernel(
    const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ gamma,
    const float* __restrict__ beta, float eps, float* __restrict__ dx, float* __restrict__ dgamma_part,
    float* __restrict__ dbeta_part, int HW, int C, int relu) {
  constexpr int GPC = 8 / VEC;
  extern __shared__ float sm[];                 // [HW][32] x values
  __shared__ float red[32][GPC];
  __shared__ float cred[32][8][8];              // per warp, per chunk: 4 x (dbeta, dgamma) partials
  const int n = blockIdx.y, c0 = blockIdx.x * 32;
  const int chunk = threadIdx.x & 7, gi = chunk / VEC, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t base = (size_t)n * HW * C + c0 + chunk * 4;
  const float* xg = x + base;
  const float* dyg = dy + base;
  float* dxg = dx + base;
  const int prow = threadIdx.x >> 3, pstep = blockDim.x >> 3;
  const float inv_total = 1.f / (float)(HW * VEC * 4);
  auto group_sum = [&](float v) -> float {
#pragma unroll
    for (int o = 1; o < VEC; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    __syncthreads();
    if (lane < 8 && (lane % VEC) == 0) red[warp][gi] = v;
    __syncthreads();
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w][gi];
    return t;
  };
  float s = 0.f;
  for (int p = prow; p < HW; p += pstep) {
    const float4 v = *reinterpret_cast<const float4*>(xg + (size_t)p * C);
    *reinterpret_cast<float4*>(sm + (size_t)p * 32 + chunk * 4) = v;
    s += (v.x + v.y) + (v.z + v.w);
  }
  const float mean = group_sum(s) * inv_total;
  float ss = 0.f;
  for (int p = prow; p < HW; p += pstep) {
    const float4 v = *reinterpret_cast<const float4*>(sm + (size_t)p * 32 + chunk * 4);
    const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
    ss = fmaf(d0, d0, ss); ss = fmaf(d1, d1, ss); ss = fmaf(d2, d2, ss); ss = fmaf(d3, d3, ss);
  }
  const float rstd = rsqrtf(group_sum(ss) * inv_total + eps);
  const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + c0 + chunk * 4));
  const float4 be = __ldg(reinterpret_cast<const float4*>(beta + c0 + chunk * 4));
  const float gam[4] = {ga.x, ga.y, ga.z, ga.w}, bet[4] = {be.x, be.y, be.z, be.w};
  // pass A: per-channel sums of dz and dz * xhat
  float sb[4] = {0.f, 0.f, 0.f, 0.f}, sg[4] = {0.f, 0.f, 0.f, 0.f};
  for (int p = prow; p < HW; p += pstep) {
    const float4 v4 = *reinterpret_cast<const float4*>(sm + (size_t)p * 32 + chunk * 4);
    const float4 g4 = *reinterpret_cast<const float4*>(dyg + (size_t)p * C);
    const float v[4] = {v4.x, v4.y, v4.z, v4.w}, g[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float xh = (v[k] - mean) * rstd;
      const float dz = (!relu || xh * gam[k] + bet[k] > 0.f) ? g[k] : 0.f;
      sb[k] += dz;
      sg[k] = fmaf(dz, xh, sg[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {                 // lanes chunk, chunk + 8, + 16, + 24 hold the same channels
    sb[k] += __shfl_xor_sync(0xffffffffu, sb[k], 8);  sb[k] += __shfl_xor_sync(0xffffffffu, sb[k], 16);
    sg[k] += __shfl_xor_sync(0xffffffffu, sg[k], 8);  sg[k] += __shfl_xor_sync(0xffffffffu, sg[k], 16);
  }
  if (lane < 8) {
#pragma unroll
    for (int k = 0; k < 4; ++k) { cred[warp][chunk][k] = sb[k]; cred[warp][chunk][4 + k] = sg[k]; }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) { sb[k] = 0.f; sg[k] = 0.f; }
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
#pragma unroll
    for (int k = 0; k < 4; ++k) { sb[k] += cred[w][chunk][k]; sg[k] += cred[w][chunk][4 + k]; }
  }
  if (threadIdx.x < 8) {
    float* db = dbeta_part + (size_t)n * C + c0 + chunk * 4;
    float* dg = dgamma_part + (size_t)n * C + c0 + chunk * 4;
#pragma unroll
    for (int k = 0; k < 4; ++k) { db[k] = sb[k]; dg[k] = sg[k]; }
  }
  // group sums A, B from the channel sums (every thread holds the totals of its chunk's four channels)
  float A = 0.f, B = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) { A = fmaf(gam[k], sb[k], A); B = fmaf(gam[k], sg[k], B); }
#pragma unroll
  for (int o = 1; o < VEC; o <<= 1) {
    A += __shfl_xor_sync(0xffffffffu, A, o);
    B += __shfl_xor_sync(0xffffffffu, B, o);
  }
  A *= inv_total; B *= inv_total;
  // pass B: dx
  for (int p = prow; p < HW; p += pstep) {
    const float4 v4 = *reinterpret_cast<const float4*>(sm + (size_t)p * 32 + chunk * 4);
    const float4 g4 = *reinterpret_cast<const float4*>(dyg + (size_t)p * C);
    const float v[4] = {v4.x, v4.y, v4.z, v4.w}, g[4] = {g4.x, g4.y, g4.z, g4.w};
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float xh = (v[k] - mean) * rstd;
      const float dz = (!relu || xh * gam[k] + bet[k] > 0.f) ? g[k] : 0.f;
      o[k] = rstd * (dz * gam[k] - (A + xh * B));
    }
    *reinterpret_cast<float4*>(dxg + (size_t)p * C) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

template <int VEC>
static int launch_groupnorm_bwd_wide(const float* x, const float* dy, const float* gamma, const float* beta, float eps,
                                     float* dx, float* dgp, float* dbp, int N, int HW, int C, int relu,
                                     cudaStream_t stream) {
  const size_t smem = (size_t)HW * 32 * sizeof(float);
  KG_CUDA(cudaFuncSetAttribute(groupnorm_relu_bwd_nhwc_wide_kernel<VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)smem));
  groupnorm_relu_bwd_nhwc_wide_kernel<VEC><<<dim3(C / 32, N), 1024, smem, stream>>>(x, dy, gamma, beta, eps, dx, dgp, dbp,
                                                                                    HW, C, relu);
  KG_LAUNCH_CHECK("groupnorm_relu_bwd_nhwc_wide_kernel");
  return KGDET_OK;
}

__device__ __forceinline__ uint32_t pack2_bf16(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}

// rows [M, C] fp32 -> channel-blocked planes [C / bk][pixels][bk] in bf16 (bk = 64) or fp32 (bk = 32):
// one thread per 16-byte output chunk
template <typename Tout>
__global__ void rows_to_blocked_kernel(const float* __restrict__ rows, unsigned char* __restrict__ dst, long long M,
                                       int C, size_t plane_bytes) {
  constexpr int EPC = 16 / (int)sizeof(Tout);   // elements per chunk: 8 bf16 / 4 fp32
  const int chunks_per_row = C / EPC;
  const long long total = M * chunks_per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)blockDim.x * gridDim.x) {
    const long long m = i / chunks_per_row;
    const int ch = (int)(i - m * chunks_per_row);
    const float* src = rows + m * C + (size_t)ch * EPC;
    const int plane = ch / 8, within = ch % 8;  // 8 chunks = 128 bytes per pixel slab
    unsigned char* d = dst + (size_t)plane * plane_bytes + (size_t)m * 128 + within * 16;
    const float4 a = *reinterpret_cast<const float4*>(src);
    if constexpr (sizeof(Tout) == 2) {
      const float4 b = *reinterpret_cast<const float4*>(src + 4);
      uint4 o;
      o.x = pack2_bf16(a.x, a.y); o.y = pack2_bf16(a.z, a.w); o.z = pack2_bf16(b.x, b.y); o.w = pack2_bf16(b.z, b.w);
      *reinterpret_cast<uint4*>(d) = o;
    } else {
      *reinterpret_cast<float4*>(d) = a;
    }
  }
}

// rows [M, C] fp32 -> UMMA-tiled bf16 rows (pointwise_umma.cu), one thread per 8 channels
__global__ void rows_to_tiled_kernel(const float* __restrict__ rows, const float* __restrict__ bias,
                                     unsigned char* __restrict__ dst, long long M, int C, int relu, int split) {
  const int chunks_per_row = C / 8;
  const int a_kblocks = (split ? 2 : 1) * (C / 64);
  const long long total = M * chunks_per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)blockDim.x * gridDim.x) {
    const long long m = i / chunks_per_row;
    const int ch = (int)(i - m * chunks_per_row);
    const float* src = rows + m * C + (size_t)ch * 8;
    float v[8];
    *reinterpret_cast<float4*>(v) = *reinterpret_cast<const float4*>(src);
    *reinterpret_cast<float4*>(v + 4) = *reinterpret_cast<const float4*>(src + 4);
    if (bias) {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + ch * 8));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + ch * 8 + 4));
      v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
      v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
    }
    if (relu) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
    }
    const int r = (int)(m & 127);
    unsigned char* t = dst + (size_t)(m >> 7) * a_kblocks * (128 * 128) + (size_t)r * 128 + (((ch & 7) ^ (r & 7)) << 4);
    uint4 hi;
    hi.x = pack2_bf16(v[0], v[1]); hi.y = pack2_bf16(v[2], v[3]); hi.z = pack2_bf16(v[4], v[5]); hi.w = pack2_bf16(v[6], v[7]);
    *reinterpret_cast<uint4*>(t + (size_t)(ch >> 3) * (128 * 128)) = hi;
    if (split) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] -= __bfloat162float(__float2bfloat16(v[e]));
      uint4 lo;
      lo.x = pack2_bf16(v[0], v[1]); lo.y = pack2_bf16(v[2], v[3]); lo.z = pack2_bf16(v[4], v[5]); lo.w = pack2_bf16(v[6], v[7]);
      *reinterpret_cast<uint4*>(t + (size_t)((C / 64) + (ch >> 3)) * (128 * 128)) = lo;
    }
  }
}

static int grid_for(long long total) {
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  return (int)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
}

int launch_rows_to_blocked(const float* rows, void* dst, long long M, int C, int bk, size_t plane_bytes, int dst_dtype,
                           cudaStream_t stream) {
  KG_CHECK_ARG((dst_dtype == KGDET_BF16 && bk == 64) || (dst_dtype == KGDET_F32 && bk == 32),
               "rows_to_blocked: channel block %d does not match dtype %d", bk, dst_dtype);
  KG_CHECK_ARG(C % bk == 0, "rows_to_blocked: C %% %d != 0", bk);
  if (M <= 0) return KGDET_OK;
  if (dst_dtype == KGDET_BF16)
    rows_to_blocked_kernel<__nv_bfloat16><<<grid_for(M * (C / 8)), 256, 0, stream>>>(rows, (unsigned char*)dst, M, C, plane_bytes);
  else
    rows_to_blocked_kernel<float><<<grid_for(M * (C / 4)), 256, 0, stream>>>(rows, (unsigned char*)dst, M, C, plane_bytes);
  KG_LAUNCH_CHECK("rows_to_blocked_kernel");
  return KGDET_OK;
}

}  // namespace kgdet

using namespace kgdet;

extern "C" int kgdet_groupnorm_relu_nhwc(const float* x, const float* gamma, const float* beta, float eps,
                                         int32_t groups, int fuse_relu, float* y, int32_t N, int32_t HW, int32_t C,
                                         void* stream) {
  KG_CHECK_ARG(x && gamma && beta && y, "kgdet_groupnorm_relu_nhwc: NULL pointer");
  KG_CHECK_ARG(N >= 0 && HW > 0 && C > 0 && groups > 0 && C % groups == 0 && (C / groups) % 4 == 0,
               "kgdet_groupnorm_relu_nhwc: need C %% groups == 0 and (C / groups) %% 4 == 0");
  if (N == 0) return KGDET_OK;
  const int cpg = C / groups;
  // wide kernel: a CTA takes 32 channels (whole groups), the map must fit shared memory next to nothing else
  if (C % 32 == 0 && 32 % cpg == 0 && (size_t)HW * 128 <= 200 * 1024 && N <= 65535 &&
      (((uintptr_t)x | (uintptr_t)y | (uintptr_t)gamma | (uintptr_t)beta) & 15) == 0) {
    const int relu = fuse_relu ? 1 : 0;
    switch (cpg / 4) {
      case 1: return launch_groupnorm_wide<1>(x, gamma, beta, eps, y, N, HW, C, relu, (cudaStream_t)stream);
      case 2: return launch_groupnorm_wide<2>(x, gamma, beta, eps, y, N, HW, C, relu, (cudaStream_t)stream);
      case 4: return launch_groupnorm_wide<4>(x, gamma, beta, eps, y, N, HW, C, relu, (cudaStream_t)stream);
      case 8: return launch_groupnorm_wide<8>(x, gamma, beta, eps, y, N, HW, C, relu, (cudaStream_t)stream);
      default: break;
    }
  }
  const size_t smem = ((size_t)HW * cpg + 32) * sizeof(float);
  KG_CHECK_ARG(smem <= 200 * 1024, "kgdet_groupnorm_relu_nhwc: group of %d x %d values does not fit shared memory", HW, cpg);
  KG_CHECK_ARG(N <= 65535, "kgdet_groupnorm_relu_nhwc: batch too large");
  KG_CUDA(cudaFuncSetAttribute(groupnorm_relu_nhwc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  groupnorm_relu_nhwc_kernel<<<dim3(groups, N), 256, smem, (cudaStream_t)stream>>>(x, gamma, beta, eps, y, HW, C, cpg,
                                                                                   fuse_relu ? 1 : 0);
  KG_LAUNCH_CHECK("groupnorm_relu_nhwc_kernel");
  return KGDET_OK;
}

extern "C" int kgdet_rows_to_tiled_bf16(const float* rows, const float* bias, void* tiled, int64_t M, int32_t C,
                                        int fuse_relu, int split, void* stream) {
  KG_CHECK_ARG(rows && tiled, "kgdet_rows_to_tiled_bf16: NULL pointer");
  KG_CHECK_ARG(!bias || ((uintptr_t)bias & 15) == 0, "kgdet_rows_to_tiled_bf16: bias must be 16-byte aligned");
  KG_CHECK_ARG(M >= 0 && C >= 64 && C % 64 == 0, "kgdet_rows_to_tiled_bf16: C %% 64 == 0 required");
  if (M == 0) return KGDET_OK;
  rows_to_tiled_kernel<<<grid_for(M * (C / 8)), 256, 0, (cudaStream_t)stream>>>(rows, bias, (unsigned char*)tiled, M, C,
                                                                                fuse_relu ? 1 : 0, split ? 1 : 0);
  KG_LAUNCH_CHECK("rows_to_tiled_kernel");
  return KGDET_OK;
}

// GroupNorm (+ ReLU) of an NHWC fp32 activation written as the SPLIT PLANES the tensor-core convolution reads
// (conv_umma.cu; the hi half doubles as the fused DCN kernel's prepared input), optionally also as NHWC fp32.
// `planes` must be kgdet_conv_split_planes_bytes(N, C, H, W) bytes, 256-byte aligned; guard bands are zeroed here.
extern "C" int kgdet_groupnorm_relu_nhwc_planes(const float* x, const float* gamma, const float* beta, float eps,
                                                int32_t groups, int fuse_relu, float* y, void* planes, int32_t N, int32_t H,
                                                int32_t W, int32_t C, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  KG_CHECK_ARG(x && gamma && beta && planes, "kgdet_groupnorm_relu_nhwc_planes: NULL pointer");
  const int HW = H * W;
  KG_CHECK_ARG(N > 0 && HW > 0 && C > 0 && C % 64 == 0 && groups > 0 && C % groups == 0 && 32 % (C / groups) == 0 &&
                   (C / groups) % 4 == 0 && (size_t)HW * 128 <= 200 * 1024 && N <= 65535,
               "kgdet_groupnorm_relu_nhwc_planes: need C %% 64 == 0, 4 | C / groups | 32 and a map of at most 1600 positions");
  KG_CHECK_ARG((((uintptr_t)x | (uintptr_t)y | (uintptr_t)gamma | (uintptr_t)beta) & 15) == 0 && ((uintptr_t)planes & 255) == 0,
               "kgdet_groupnorm_relu_nhwc_planes: misaligned pointer");
  // split-plane layout of conv_umma.cu (hi planes in the DCN prepared-input layout, then the lo planes)
  const size_t guard = (size_t)(W + 2) * 128, in_bytes = (size_t)N * HW * 128;
  const size_t plane_bytes = align_up(in_bytes + 2 * guard, 1024), half = plane_bytes * (C / 64);
  unsigned char* hi = (unsigned char*)planes + guard;
  unsigned char* lo = hi + half;
  const int relu = fuse_relu ? 1 : 0;
  const int gb = (int)guard, tb = (int)(plane_bytes - guard - in_bytes);      // zeroed by the kernel (multiples of 16)
  switch ((C / groups) / 4) {
    case 1: return launch_groupnorm_wide<1>(x, gamma, beta, eps, y, N, HW, C, relu, stream, hi, lo, plane_bytes, gb, tb);
    case 2: return launch_groupnorm_wide<2>(x, gamma, beta, eps, y, N, HW, C, relu, stream, hi, lo, plane_bytes, gb, tb);
    case 4: return launch_groupnorm_wide<4>(x, gamma, beta, eps, y, N, HW, C, relu, stream, hi, lo, plane_bytes, gb, tb);
    case 8: return launch_groupnorm_wide<8>(x, gamma, beta, eps, y, N, HW, C, relu, stream, hi, lo, plane_bytes, gb, tb);
    default: break;
  }
  set_error("kgdet_groupnorm_relu_nhwc_planes: unsupported channels per group %d", C / groups);
  return KGDET_ERR_UNSUPPORTED;
}

// GroupNorm (+ ReLU) of NHWC fp32 maps of ANY size, streaming (two passes, see groupnorm_stream_*_kernel): y (NHWC
// fp32) and / or `planes` (split planes as above; with `planes_hi_only` just the hi half = the DCN prepared input,
// kgdet_dcn_prepared_input_bytes) may be NULL, not both.  `workspace`: kgdet_groupnorm_stream_workspace_bytes bytes.
// Reference: torch.nn.GroupNorm inside mmdet ConvModule (conv_module.py:96-110,156-164), as used by the towers of
// reppoints_head_kp_parallel.py:115-145.
extern "C" size_t kgdet_groupnorm_stream_workspace_bytes(int32_t N, int32_t HW, int32_t C, int32_t groups) {
  if (N <= 0 || HW <= 0 || !stream_gn_supported(C, groups)) return 0;
  return ((((size_t)N * stream_gn_chunks(HW, C) * groups * 3 + 1) & ~(size_t)1) + (size_t)N * groups * 2) * sizeof(float);
}

extern "C" int kgdet_groupnorm_relu_nhwc_stream(const float* x, const float* gamma, const float* beta, float eps,
                                                int32_t groups, int fuse_relu, float* y, void* planes, int planes_hi_only,
                                                int32_t N, int32_t H, int32_t W, int32_t C, void* workspace,
                                                size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  KG_CHECK_ARG(x && gamma && beta && (y || planes) && workspace, "kgdet_groupnorm_relu_nhwc_stream: NULL pointer");
  KG_CHECK_ARG(N > 0 && N <= 65535 && H > 0 && W > 0 && stream_gn_supported(C, groups),
               "kgdet_groupnorm_relu_nhwc_stream: need C %% 128 == 0, C <= 1024 and a power-of-two multiple of 4 channels per group");
  const int HW = H * W;
  KG_CHECK_ARG(workspace_bytes >= kgdet_groupnorm_stream_workspace_bytes(N, HW, C, groups),
               "kgdet_groupnorm_relu_nhwc_stream: workspace too small");
  KG_CHECK_ARG((((uintptr_t)x | (uintptr_t)y | (uintptr_t)gamma | (uintptr_t)beta | (uintptr_t)workspace) & 15) == 0 &&
                   ((uintptr_t)planes & 255) == 0, "kgdet_groupnorm_relu_nhwc_stream: misaligned pointer");
  unsigned char *hi = nullptr, *lo = nullptr;
  size_t plane_bytes = 0;
  if (planes) {
    const size_t guard = (size_t)(W + 2) * 128, in_bytes = (size_t)N * HW * 128;
    plane_bytes = align_up(in_bytes + 2 * guard, 1024);
    const size_t half = plane_bytes * (C / 64);
    for (int h = 0; h < (planes_hi_only ? 1 : 2); ++h) {
      unsigned char* b = (unsigned char*)planes + (size_t)h * half;
      KG_CUDA(cudaMemset2DAsync(b, plane_bytes, 0, guard, C / 64, stream));
      KG_CUDA(cudaMemset2DAsync(b + guard + in_bytes, plane_bytes, 0, plane_bytes - guard - in_bytes, C / 64, stream));
    }
    hi = (unsigned char*)planes + guard;
    lo = planes_hi_only ? nullptr : hi + half;
  }
  const int cpg = C / groups, chunks = stream_gn_chunks(HW, C), rows = GS_THREADS / (C / 4);
  const size_t smem = (size_t)rows * groups * 3 * sizeof(float);
  float* partial = (float*)workspace;
  float* stat = partial + (((size_t)N * chunks * groups * 3 + 1) & ~(size_t)1);        // 8-byte aligned pairs
  groupnorm_stream_stats_kernel<<<dim3(chunks, N), GS_THREADS, smem, stream>>>(x, HW, C, cpg, partial);
  KG_LAUNCH_CHECK("groupnorm_stream_stats_kernel");
  groupnorm_stream_finalize_kernel<<<N, 1024, 0, stream>>>(partial, chunks, groups, eps, stat);
  KG_LAUNCH_CHECK("groupnorm_stream_finalize_kernel");
  groupnorm_stream_apply_kernel<<<dim3(chunks, N), GS_THREADS, 0, stream>>>(x, gamma, beta, stat, y, HW, C, cpg,
                                                                           fuse_relu ? 1 : 0, hi, lo, plane_bytes);
  KG_LAUNCH_CHECK("groupnorm_stream_apply_kernel");
  return KGDET_OK;
}

// Backward of kgdet_groupnorm_relu_nhwc (maps of at most 1600 positions): dx (NHWC fp32) and the per-image partials
// of the affine parameters' gradients, dgamma_part / dbeta_part [N, C] (sum over N = the gradient).
// replaces torch.nn.GroupNorm's + ReLU's autograd in ConvModule (mmdet/models/utils/conv_module.py:96-110,156-164)
// for the training step of the towers (KP3:292-313).
extern "C" int kgdet_groupnorm_relu_nhwc_backward(const float* x, const float* dy, const float* gamma, const float* beta,
                                                  float eps, int32_t groups, int fuse_relu, float* dx,
                                                  float* dgamma_part, float* dbeta_part, int32_t N, int32_t HW,
                                                  int32_t C, void* stream) {
  KG_CHECK_ARG(x && dy && gamma && beta && dx && dgamma_part && dbeta_part, "kgdet_groupnorm_relu_nhwc_backward: NULL pointer");
  KG_CHECK_ARG(N > 0 && N <= 65535 && HW > 0 && C > 0 && C % 32 == 0 && groups > 0 && C % groups == 0 &&
                   32 % (C / groups) == 0 && (C / groups) % 4 == 0 && (size_t)HW * 128 <= 200 * 1024,
               "kgdet_groupnorm_relu_nhwc_backward: need C %% 32 == 0, 4 | C / groups | 32 and a map of at most 1600 positions");
  KG_CHECK_ARG((((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dx | (uintptr_t)gamma | (uintptr_t)beta) & 15) == 0,
               "kgdet_groupnorm_relu_nhwc_backward: misaligned pointer");
  const int relu = fuse_relu ? 1 : 0;
  switch ((C / groups) / 4) {
    case 1: return launch_groupnorm_bwd_wide<1>(x, dy, gamma, beta, eps, dx, dgamma_part, dbeta_part, N, HW, C, relu, (cudaStream_t)stream);
    case 2: return launch_groupnorm_bwd_wide<2>(x, dy, gamma, beta, eps, dx, dgamma_part, dbeta_part, N, HW, C, relu, (cudaStream_t)stream);
    case 4: return launch_groupnorm_bwd_wide<4>(x, dy, gamma, beta, eps, dx, dgamma_part, dbeta_part, N, HW, C, relu, (cudaStream_t)stream);
    case 8: return launch_groupnorm_bwd_wide<8>(x, dy, gamma, beta, eps, dx, dgamma_part, dbeta_part, N, HW, C, relu, (cudaStream_t)stream);
    default: break;
  }
  set_error("kgdet_groupnorm_relu_nhwc_backward: unsupported channels per group %d", C / groups);
  return KGDET_ERR_UNSUPPORTED;
}
