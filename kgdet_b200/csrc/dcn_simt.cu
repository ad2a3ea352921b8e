// Exact-fp32 deformable convolution on the SIMT pipe: the general path (any stride, dilation,
// groups, deformable_groups, mask) and the on-device cross-check for the tcgen05 path.
//
// Unlike the reference (deform_conv_cuda.cpp:220-244: im2col kernel -> `columns` in HBM ->
// cuBLAS SGEMM -> transposed copy) the column tensor never exists: every kernel gathers its
// 64x32 / 32x64 operand tile from the NHWC input straight into shared memory using the
// per-sample plan (dcn.cuh) and contracts it with FFMA.  All tiles are 64x64 per CTA,
// 4x4 per thread, 256 threads.
#include "dcn.cuh"

namespace kgdet {

static constexpr int TM = 64, TN = 64, TK = 32, NT = 256;

__device__ __forceinline__ SampleRec ld_rec(const SampleRec* p) {
  SampleRec r;
  const int4 a = __ldg(reinterpret_cast<const int4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  r.pix[0] = a.x; r.pix[1] = a.y; r.pix[2] = a.z; r.pix[3] = a.w;
  r.w[0] = b.x; r.w[1] = b.y; r.w[2] = b.z; r.w[3] = b.w;
  return r;
}

// bilinear sample of channel `c` (global channel index) with a plan record
__device__ __forceinline__ float sample(const float* __restrict__ in, const SampleRec& r, int C, int c) {
  float v = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (r.w[i] != 0.f) v = fmaf(r.w[i], __ldg(in + (size_t)r.pix[i] * C + c), v);
  return v;
}

// ---- weight repacking --------------------------------------------------------------------
// [Cout, Cg, K] -> [g][tap][c][o]   (forward:  B tile rows = (tap, c), contiguous over o)
__global__ void pack_w_fwd_kernel(const float* __restrict__ w, float* __restrict__ p, int groups,
                                  int Og, int Cg, int K) {
  const int total = groups * Og * Cg * K;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += blockDim.x * gridDim.x) {
    int o = i % Og, t = i / Og;
    int c = t % Cg; t /= Cg;
    int tap = t % K, g = t / K;
    p[i] = w[((size_t)(g * Og + o) * Cg + c) * K + tap];
  }
}
// [Cout, Cg, K] -> [g][tap][o][c]   (dgrad:    B tile rows = (tap, o), contiguous over c)
__global__ void pack_w_dgrad_kernel(const float* __restrict__ w, float* __restrict__ p, int groups,
                                    int Og, int Cg, int K) {
  const int total = groups * Og * Cg * K;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += blockDim.x * gridDim.x) {
    int c = i % Cg, t = i / Cg;
    int o = t % Og; t /= Og;
    int tap = t % K, g = t / K;
    p[i] = w[((size_t)(g * Og + o) * Cg + c) * K + tap];
  }
}

size_t simt_packed_weight_bytes(const DcnGeom& g) {
  return (size_t)g.Cout * (g.C / g.groups) * g.K * sizeof(float);
}
int simt_pack_weight(const DcnGeom& g, const float* weight, float* packed, cudaStream_t stream) {
  int total = g.Cout * (g.C / g.groups) * g.K;
  pack_w_fwd_kernel<<<ceil_div(total, 256), 256, 0, stream>>>(weight, packed, g.groups,
                                                              g.Cout / g.groups, g.C / g.groups, g.K);
  KG_LAUNCH_CHECK("pack_w_fwd_kernel");
  return KGDET_OK;
}
int simt_pack_weight_dgrad(const DcnGeom& g, const float* weight, float* packed, cudaStream_t stream) {
  int total = g.Cout * (g.C / g.groups) * g.K;
  pack_w_dgrad_kernel<<<ceil_div(total, 256), 256, 0, stream>>>(weight, packed, g.groups,
                                                                g.Cout / g.groups, g.C / g.groups, g.K);
  KG_LAUNCH_CHECK("pack_w_dgrad_kernel");
  return KGDET_OK;
}

// 4x4 register tile FMA over one TK-deep smem slab.  As: [TM][TK+1] (row, k); Bs: [TK][TN].
__device__ __forceinline__ void tile_fma_rk(const float (*As)[TK + 1], const float (*Bs)[TN], int ty,
                                            int tx, float (&acc)[4][4]) {
#pragma unroll 8
  for (int kk = 0; kk < TK; ++kk) {
    float a[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = As[ty * 4 + i][kk];
    const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
    const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
}

template <typename T> __device__ __forceinline__ void store_out(T* p, float v);
template <> __device__ __forceinline__ void store_out<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void store_out<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  *p = __float2bfloat16(v);
}

// ---- forward ---------------------------------------------------------------------------------
// out[n, o, y, x] = sum_{c, tap} W[o, c, tap] * S(n, c, tap, y, x)     (deform_conv_cuda.cpp:225-234)
template <typename Tout>
__global__ void __launch_bounds__(NT)
simt_fwd_kernel(DcnGeom g, const float* __restrict__ in, const SampleRec* __restrict__ plan,
                const float* __restrict__ wp, const float* __restrict__ bias, Tout* __restrict__ out,
                int coff, int ctot, int relu) {
  __shared__ float As[TM][TK + 1];
  __shared__ __align__(16) float Bs[TK][TN];
  const int m0 = blockIdx.x * TM, o0 = blockIdx.y * TN, grp = blockIdx.z;
  const int Cg = g.C / g.groups, Og = g.Cout / g.groups, cpdg = g.C / g.dgroups;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4] = {};
  for (int tap = 0; tap < g.K; ++tap) {
    for (int c0 = 0; c0 < Cg; c0 += TK) {
      const int c = c0 + lane;
      const int cgl = grp * Cg + c;
      const int dgi = (c < Cg) ? cgl / cpdg : 0;
#pragma unroll
      for (int rr = 0; rr < TM / 8; ++rr) {
        const int row = warp + rr * 8;
        float v = 0.f;
        if (c < Cg) {
          SampleRec r = ld_rec(plan + ((size_t)(m0 + row) * g.dgroups + dgi) * g.K + tap);
          v = sample(in, r, g.C, cgl);
        }
        As[row][lane] = v;
      }
      const float* wsrc = wp + ((size_t)(grp * g.K + tap) * Cg + c0) * Og + o0;
      for (int e = tid; e < TK * TN; e += NT) {
        int kk = e / TN, oo = e - kk * TN;
        Bs[kk][oo] = (c0 + kk < Cg && o0 + oo < Og) ? __ldg(wsrc + (size_t)kk * Og + oo) : 0.f;
      }
      __syncthreads();
      tile_fma_rk(As, Bs, ty, tx, acc);
      __syncthreads();
    }
  }
  const int HoWo = g.Ho * g.Wo;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
    const int n = m / HoWo, p = m - n * HoWo;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int o = o0 + tx * 4 + j;
      if (o >= Og) continue;
      const int og = grp * Og + o;
      float v = acc[i][j] + (bias ? bias[og] : 0.f);
      if (relu) v = fmaxf(v, 0.f);
      store_out<Tout>(out + ((size_t)n * ctot + coff + og) * HoWo + p, v);
    }
  }
}

int simt_forward(const DcnGeom& g, const float* in_nhwc, const SampleRec* plan, const float* packed_w,
                 const float* bias, const OutSpec& o, cudaStream_t stream) {
  dim3 grid(ceil_div(g.M, TM), ceil_div(g.Cout / g.groups, TN), g.groups);
  if (o.dtype == KGDET_F32)
    simt_fwd_kernel<float><<<grid, NT, 0, stream>>>(g, in_nhwc, plan, packed_w, bias, (float*)o.out, o.coff,
                                                    o.ctot, o.relu);
  else
    simt_fwd_kernel<__nv_bfloat16><<<grid, NT, 0, stream>>>(g, in_nhwc, plan, packed_w, bias,
                                                            (__nv_bfloat16*)o.out, o.coff, o.ctot, o.relu);
  KG_LAUNCH_CHECK("simt_fwd_kernel");
  return KGDET_OK;
}

// ---- backward w.r.t. input, offset (and mask) ------------------------------------------------
// One CTA = (64 positions, one tap, one group).  For each 64-channel slab it forms the column
// gradient tile  cg = gO . W  (deform_conv_cuda.cpp:330-331) in registers -> smem, then one warp
// per position row turns it into
//   grad_input  : bilinear-weighted scatter, coalesced fp32 red over channels (:318-331)
//   grad_offset : sum_c cg * dS/d(py|px), warp-shuffle reduction over channels instead of the
//                 reference's serial per-thread channel loop (:405-433)
//   grad_mask   : sum_c cg * S (:752,764)
__global__ void __launch_bounds__(NT)
simt_bwd_input_kernel(DcnGeom g, const float* __restrict__ in, const float* __restrict__ go,
                      const SampleRec* __restrict__ plan, const SampleAux* __restrict__ aux,
                      const float* __restrict__ wd, float* __restrict__ gin,
                      float* __restrict__ goff, float* __restrict__ gmask) {
  __shared__ float As[TM][TK + 1];
  __shared__ __align__(16) float Bs[TK][TN];
  __shared__ float Cs[TM][TN + 1];
  const int m0 = blockIdx.x * TM, tap = blockIdx.y, grp = blockIdx.z;
  const int Cg = g.C / g.groups, Og = g.Cout / g.groups, cpdg = g.C / g.dgroups;
  const int HoWo = g.Ho * g.Wo;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ty = tid >> 4, tx = tid & 15;
  const bool uniform_dg = (g.dgroups == 1);
  float pdy[TM / 8], pdx[TM / 8], pms[TM / 8];
#pragma unroll
  for (int rr = 0; rr < TM / 8; ++rr) pdy[rr] = pdx[rr] = pms[rr] = 0.f;

  for (int c0 = 0; c0 < Cg; c0 += TN) {
    float acc[4][4] = {};
    for (int q0 = 0; q0 < Og; q0 += TK) {
#pragma unroll
      for (int rr = 0; rr < TM / 8; ++rr) {
        const int row = warp + rr * 8, m = m0 + row;
        As[row][lane] = (m < g.M && q0 + lane < Og)
                            ? __ldg(go + (size_t)m * g.Cout + grp * Og + q0 + lane) : 0.f;
      }
      const float* wsrc = wd + ((size_t)(grp * g.K + tap) * Og + q0) * Cg + c0;
      for (int e = tid; e < TK * TN; e += NT) {
        int kk = e / TN, cc = e - kk * TN;
        Bs[kk][cc] = (q0 + kk < Og && c0 + cc < Cg) ? __ldg(wsrc + (size_t)kk * Cg + cc) : 0.f;
      }
      __syncthreads();
      tile_fma_rk(As, Bs, ty, tx, acc);
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) Cs[ty * 4 + i][tx * 4 + j] = acc[i][j];
    __syncthreads();

#pragma unroll
    for (int rr = 0; rr < TM / 8; ++rr) {
      const int row = warp + rr * 8, m = m0 + row;
      if (m >= g.M) continue;                       // warp-uniform
      const int n = m / HoWo, p = m - n * HoWo;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int cc = lane + half * 32, c = c0 + cc;
        const bool act = (c < Cg);
        const int cgl = grp * Cg + (act ? c : 0);
        const int dgi = act ? cgl / cpdg : -1;
        float ldy = 0.f, ldx = 0.f, lms = 0.f;
        if (act) {
          const size_t ridx = ((size_t)m * g.dgroups + dgi) * g.K + tap;
          const SampleRec r = ld_rec(plan + ridx);
          const float4 a4 = __ldg(reinterpret_cast<const float4*>(aux + ridx));
          const float lh = a4.x, lw = a4.y, mk = a4.z;
          const int valid = __float_as_int(a4.w);
          const float gv = Cs[row][cc];
          const float hh = 1.f - lh, hw = 1.f - lw;
          float v[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            v[i] = 0.f;
            if (valid & (1 << i)) {
              const size_t a = (size_t)r.pix[i] * g.C + cgl;
              v[i] = __ldg(in + a);
              atomicAdd(gin + a, gv * r.w[i]);      // RED.ADD.F32, lanes = consecutive channels
            }
          }
          // dS/dpy, dS/dpx  (get_coordinate_weight, deform_conv_cuda_kernel.cu:163-184)
          const float dsy = -hw * v[0] - lw * v[1] + hw * v[2] + lw * v[3];
          const float dsx = -hh * v[0] + hh * v[1] - lh * v[2] + lh * v[3];
          const float s = hh * hw * v[0] + hh * lw * v[1] + lh * hw * v[2] + lh * lw * v[3];
          ldy = gv * mk * dsy;
          ldx = gv * mk * dsx;
          lms = gv * s;
        }
        if (uniform_dg) {
          pdy[rr] += ldy; pdx[rr] += ldx; pms[rr] += lms;
        } else {
          // segmented warp reduction: lanes may belong to different deformable groups
          unsigned remaining = __ballot_sync(0xffffffffu, act);
          while (remaining) {
            const int leader = __ffs(remaining) - 1;
            const int d = __shfl_sync(0xffffffffu, dgi, leader);
            const bool mine = (dgi == d);
            const float sy = warp_sum(mine ? ldy : 0.f);
            const float sx = warp_sum(mine ? ldx : 0.f);
            const float sm = warp_sum(mine ? lms : 0.f);
            if (lane == 0) {
              atomicAdd(goff + ((size_t)(n * g.dgroups + d) * 2 * g.K + 2 * tap) * HoWo + p, sy);
              atomicAdd(goff + ((size_t)(n * g.dgroups + d) * 2 * g.K + 2 * tap + 1) * HoWo + p, sx);
              if (gmask) atomicAdd(gmask + ((size_t)(n * g.dgroups + d) * g.K + tap) * HoWo + p, sm);
            }
            remaining &= ~__ballot_sync(0xffffffffu, mine);
          }
        }
      }
    }
    __syncthreads();
  }
  if (uniform_dg) {
#pragma unroll
    for (int rr = 0; rr < TM / 8; ++rr) {
      const int m = m0 + warp + rr * 8;
      if (m >= g.M) continue;
      const float sy = warp_sum(pdy[rr]), sx = warp_sum(pdx[rr]), sm = warp_sum(pms[rr]);
      if (lane == 0) {
        const int n = m / HoWo, p = m - n * HoWo;
        // single contributor per entry when groups == 1 -> deterministic
        atomicAdd(goff + ((size_t)n * 2 * g.K + 2 * tap) * HoWo + p, sy);
        atomicAdd(goff + ((size_t)n * 2 * g.K + 2 * tap + 1) * HoWo + p, sx);
        if (gmask) atomicAdd(gmask + ((size_t)n * g.K + tap) * HoWo + p, sm);
      }
    }
  }
}

int simt_backward_input(const DcnGeom& g, const float* in_nhwc, const float* go_nhwc,
                        const SampleRec* plan, const SampleAux* aux, const float* w_dgrad,
                        float* gin_nhwc, float* grad_offset, float* grad_mask, cudaStream_t stream) {
  KG_CHECK_ARG(g.K <= 65535 && g.groups <= 65535, "dcn backward: kernel/groups too large");
  dim3 grid(ceil_div(g.M, TM), g.K, g.groups);
  simt_bwd_input_kernel<<<grid, NT, 0, stream>>>(g, in_nhwc, go_nhwc, plan, aux, w_dgrad, gin_nhwc,
                                                 grad_offset, grad_mask);
  KG_LAUNCH_CHECK("simt_bwd_input_kernel");
  return KGDET_OK;
}

// ---- backward w.r.t. weight ------------------------------------------------------------------
// grad_W[o, c, tap] = scale * sum_m gO[m, o] * S(m, c, tap)           (deform_conv_cuda.cpp:443-461)
// One CTA = (64 couts, 64 channels, one tap, one group); the reduction over all M positions
// runs inside the CTA (deterministic, no atomics); the column slab is re-gathered on the fly.
__global__ void __launch_bounds__(NT)
simt_bwd_weight_kernel(DcnGeom g, const float* __restrict__ in, const float* __restrict__ go,
                       const SampleRec* __restrict__ plan, float scale, float* __restrict__ gw) {
  __shared__ __align__(16) float As[TK][TN];   // [m][o]
  __shared__ __align__(16) float Bs[TK][TN];   // [m][c]
  const int Cg = g.C / g.groups, Og = g.Cout / g.groups, cpdg = g.C / g.dgroups;
  const int nct = ceil_div(Cg, TN);
  const int o0 = blockIdx.x * TN, ctile = blockIdx.y % nct, tap = blockIdx.y / nct, grp = blockIdx.z;
  const int c0 = ctile * TN;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ty = tid >> 4, tx = tid & 15;
  // two-level summation: `acc` absorbs 32 slabs (32 * TK positions), then is folded into `tot` -- one fp32
  // accumulator over all M positions drifts to 1.7e-5 of the tensor maximum at M = 134 400 (FPN P3, batch 8),
  // above the 1e-5 bar the reference's own SGEMM meets there (6e-6)
  float acc[4][4] = {}, tot[4][4] = {};
  int slabs = 0;
  for (int m0 = 0; m0 < g.M; m0 += TK) {
    if (++slabs == 33) {
      slabs = 1;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { tot[i][j] += acc[i][j]; acc[i][j] = 0.f; }
    }
#pragma unroll
    for (int rr = 0; rr < TK / 8; ++rr) {
      const int kk = warp + rr * 8, m = m0 + kk;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int e = lane + half * 32;
        As[kk][e] = (m < g.M && o0 + e < Og) ? __ldg(go + (size_t)m * g.Cout + grp * Og + o0 + e) : 0.f;
        float v = 0.f;
        if (m < g.M && c0 + e < Cg) {
          const int cgl = grp * Cg + c0 + e;
          SampleRec r = ld_rec(plan + ((size_t)m * g.dgroups + cgl / cpdg) * g.K + tap);
          v = sample(in, r, g.C, cgl);
        }
        Bs[kk][e] = v;
      }
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < TK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int o = o0 + ty * 4 + i;
    if (o >= Og) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + tx * 4 + j;
      if (c >= Cg) continue;
      gw[((size_t)(grp * Og + o) * Cg + c) * g.K + tap] = scale * (tot[i][j] + acc[i][j]);
    }
  }
}

// grad_bias[o] = sum_m gO[m, o]      (modulated DCN with bias: deform_conv_cuda.cpp:659-666)
__global__ void bias_grad_kernel(const float* __restrict__ go, int M, int Cout, float* __restrict__ gb) {
  __shared__ float part[8][33];
  const int o = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (o < Cout)
    for (int m = threadIdx.y; m < M; m += 8) s += go[(size_t)m * Cout + o];
  part[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && o < Cout) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += part[k][threadIdx.x];
    gb[o] = t;
  }
}

int simt_backward_weight(const DcnGeom& g, const float* in_nhwc, const float* go_nhwc,
                         const SampleRec* plan, float scale, float* grad_weight, float* grad_bias,
                         cudaStream_t stream) {
  const int Cg = g.C / g.groups, Og = g.Cout / g.groups;
  KG_CHECK_ARG((long long)ceil_div(Cg, TN) * g.K <= 65535, "dcn backward weight: grid.y overflow");
  dim3 grid(ceil_div(Og, TN), ceil_div(Cg, TN) * g.K, g.groups);
  simt_bwd_weight_kernel<<<grid, NT, 0, stream>>>(g, in_nhwc, go_nhwc, plan, scale, grad_weight);
  KG_LAUNCH_CHECK("simt_bwd_weight_kernel");
  if (grad_bias) {
    bias_grad_kernel<<<ceil_div(g.Cout, 32), dim3(32, 8), 0, stream>>>(go_nhwc, g.M, g.Cout, grad_bias);
    KG_LAUNCH_CHECK("bias_grad_kernel");
  }
  return KGDET_OK;
}

}  // namespace kgdet
