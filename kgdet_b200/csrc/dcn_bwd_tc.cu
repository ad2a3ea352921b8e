// Deformable-convolution backward on the tensor cores (bf16 mode; groups == deformable_groups == 1).
//
// Input / offset / mask gradient  (reference: deform_conv_cuda.cpp:260-371, kernels
// deform_conv_cuda_kernel.cu:278-435):
//   cg[m, (tap, c)] = sum_o gO[m, o] * W[o, c, tap]          tcgen05 GEMM (gemm_umma.cu), bf16, written in
//                                                            tap chunks small enough to stay in the 126 MB L2
//   col2im          one warp per (position, tap): lanes sweep the channels (16-byte loads of cg and of
//                   the four NHWC corner rows), grad_input by vectorised fp32 red.global.add.v4 on the
//                   NHWC gradient buffer, grad_offset / grad_mask by warp-shuffle reduction over channels
//                   -- one writer per element, no atomics, deterministic (the reference loops over the
//                   C channels serially in one thread, :405-433).
// Weight gradient  (reference: deform_conv_cuda.cpp:373-484 recomputes im2col into HBM):
//   colT[(tap, c), m] = S(m, tap, c)                         gather kernel, transposed through smem, tap chunks
//   gW^T[(tap, c), o] = sum_m colT * gO^T[o, m]              split-K tcgen05 GEMM with fp32 red epilogue
#include "dcn.cuh"

namespace kgdet {

// gather-fused weight gradient (dcn_wgrad_umma.cu)
bool wgrad_fused_supported(const DcnGeom& g);
size_t wgrad_fused_go_bytes(const DcnGeom& g);
int launch_go_to_tiled(const DcnGeom& g, const void* grad_output, void* tiled, int dtype, cudaStream_t stream);
int wgrad_fused(const DcnGeom& g, const void* in_blocked, size_t plane_bytes, const SampleRec16* plan,
                const void* go_tiled, float* gwt, cudaStream_t stream);

int umma_gemm(const __nv_bfloat16* A, long long lda, const __nv_bfloat16* B, long long ldb, void* C,
              long long ldc, int M, int N, int K, int out_dtype, int splits, float alpha, cudaStream_t stream);

// owned-slice col2im (dcn_col2im_own.cu)
bool col2im_own_supported(const DcnGeom& g);
size_t col2im_own_part_bytes(const DcnGeom& g);
size_t col2im_own_sched_bytes(const DcnGeom& g);
int col2im_own_schedule(const DcnGeom& g, const SampleRec* plan, const SampleAux* aux, unsigned* sched,
                        cudaStream_t stream);
int col2im_own(const DcnGeom& g, const __nv_bfloat16* cg, const __nv_bfloat16* in_nhwc, const SampleRec* plan,
               const SampleAux* aux, const unsigned* sched, int tap0, int ntaps, float* gin_nhwc, float* gpart,
               cudaStream_t stream);
int col2im_own_finish(const DcnGeom& g, const float* gpart, float* goff, float* gmask, cudaStream_t stream);

bool bwd_tc_supported(const DcnGeom& g, int precision) {
  return precision == KGDET_PREC_BF16 && g.groups == 1 && g.dgroups == 1 && g.C % 64 == 0 &&
         g.Cout % 64 == 0;
}

// taps per chunk so that an [M, taps*C] bf16 buffer stays around 48 MB
static int taps_per_chunk(const DcnGeom& g, size_t rows) {
  const size_t per_tap = rows * g.C * 2;
  int t = (int)((48ull << 20) / (per_tap ? per_tap : 1));
  if (t < 1) t = 1;
  if (t > g.K) t = g.K;
  return t;
}
static size_t mpad64(const DcnGeom& g) { return (size_t)ceil_div(g.M, 64) * 64; }

// (weight layout for the column-gradient GEMM, Wd[(tap*C + c), o] = W[o, c, tap]: pack_wd_tiled_kernel below)

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
  f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
  f[4] = __uint_as_float(v.z << 16); f[5] = __uint_as_float(v.z & 0xffff0000u);
  f[6] = __uint_as_float(v.w << 16); f[7] = __uint_as_float(v.w & 0xffff0000u);
}

// ---- col2im + offset / mask gradient ----------------------------------------------------------------
// cg: [M, ntaps*C] bf16 (taps tap0 .. tap0+ntaps-1).  One warp per position; the taps of the chunk are taken
// in groups of TB = 5 with everything unrolled, so a warp has five independent load -> FMA chains in flight
// (the first version took one (position, tap) pair per warp iteration and was bound by that chain's latency:
// removing its atomics changed 124 us per chunk to 101 us).  Per lane and tap: 8 channels, four dot products
// <cg, corner_i> (the three gradients are linear in them), 2 x 4 red.global.add.v4.f32 into the NHWC input
// gradient.  The 15 per-lane partials (5 taps x {dy, dx, mask}) are reduced with ONE transposing butterfly
// (16 -> 8 -> 4 -> 2 -> 1 values per lane: 16 shuffles instead of 75); lane 2j ends up owning value j and
// writes it -- one writer per element, no atomics on grad_offset / grad_mask, deterministic.
static constexpr int COL2IM_TB = 5;

__global__ void __launch_bounds__(256, 4)
col2im_tc_kernel(DcnGeom g, const __nv_bfloat16* __restrict__ cg, const __nv_bfloat16* __restrict__ in,
                 const SampleRec* __restrict__ plan, const SampleAux* __restrict__ aux, int tap0, int ntaps,
                 float* __restrict__ gin, float* __restrict__ goff, float* __restrict__ gmask) {
  constexpr int TB = COL2IM_TB;
  const int lane = threadIdx.x & 31;
  const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int HoWo = g.Ho * g.Wo;
  const int ngroups = (ntaps + TB - 1) / TB;
  const long long total = (long long)g.M * ngroups;
  for (long long wi = warp_global; wi < total; wi += nwarps) {
    const int m = (int)(wi / ngroups), tl0 = (int)(wi - (long long)m * ngroups) * TB;
    float part[16];
    float mk[TB];
#pragma unroll
    for (int i = 0; i < 16; ++i) part[i] = 0.f;
#pragma unroll
    for (int t = 0; t < TB; ++t) {
      const int tl = tl0 + t;
      mk[t] = 0.f;
      if (tl >= ntaps) continue;                                  // warp-uniform
      const size_t ridx = (size_t)m * g.K + tap0 + tl;
      const int4 pa = __ldg(reinterpret_cast<const int4*>(plan + ridx));
      const float4 pw = __ldg(reinterpret_cast<const float4*>(plan + ridx) + 1);
      const float4 a4 = __ldg(reinterpret_cast<const float4*>(aux + ridx));
      const int pix[4] = {pa.x, pa.y, pa.z, pa.w};
      const float wgt[4] = {pw.x, pw.y, pw.z, pw.w};
      const float lh = a4.x, lw = a4.y;
      mk[t] = a4.z;
      const int valid = __float_as_int(a4.w);
      if (!valid) continue;                                       // warp-uniform
      const float hh = 1.f - lh, hw = 1.f - lw;
      float d[4] = {0.f, 0.f, 0.f, 0.f};                          // <cg, corner_i> over this lane's channels
      const __nv_bfloat16* cgrow = cg + ((size_t)m * ntaps + tl) * g.C;
      for (int c0 = lane * 8; c0 < g.C; c0 += 256) {
        float gv[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(cgrow + c0)), gv);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (valid & (1 << i)) {
            float v[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(in + (size_t)pix[i] * g.C + c0)), v);
            float* dst = gin + (size_t)pix[i] * g.C + c0;
            const float wi_ = wgt[i];
            atomicAdd(reinterpret_cast<float4*>(dst),
                      make_float4(gv[0] * wi_, gv[1] * wi_, gv[2] * wi_, gv[3] * wi_));
            atomicAdd(reinterpret_cast<float4*>(dst + 4),
                      make_float4(gv[4] * wi_, gv[5] * wi_, gv[6] * wi_, gv[7] * wi_));
            float acc = d[i];
#pragma unroll
            for (int e = 0; e < 8; ++e) acc = fmaf(gv[e], v[e], acc);
            d[i] = acc;
          }
        }
      }
      // d(sample)/dy, d(sample)/dx and the sample itself, as combinations of the corner values
      // (deform_conv_cuda_kernel.cu:144-187 with the corner validity of :97-108 already in `valid`)
      part[3 * t + 0] = -hw * d[0] - lw * d[1] + hw * d[2] + lw * d[3];
      part[3 * t + 1] = -hh * d[0] + hh * d[1] - lh * d[2] + lh * d[3];
      part[3 * t + 2] = hh * hw * d[0] + hh * lw * d[1] + lh * hw * d[2] + lh * lw * d[3];
    }
    // transposing butterfly: after the level with lane bit b, a lane keeps the half of its values selected by b
#pragma unroll
    for (int lvl = 0; lvl < 4; ++lvl) {
      const int bit = 16 >> lvl, half = 8 >> lvl;                 // values kept after this level
      const bool up = (lane & bit) != 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j < half) {
          const float send = up ? part[j] : part[j + half];
          const float keep = up ? part[j + half] : part[j];
          part[j] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
        }
      }
    }
    const float tot = part[0] + __shfl_xor_sync(0xffffffffu, part[0], 1);
    const int idx = lane >> 1;                                    // value owned by this lane pair
    const int t = idx / 3, q = idx - 3 * t;
    if ((lane & 1) == 0 && idx < 3 * TB && tl0 + t < ntaps) {
      const int tap = tap0 + tl0 + t;
      const int n = m / HoWo, p = m - n * HoWo;
      float mkt = mk[0];
#pragma unroll
      for (int u = 1; u < TB; ++u) mkt = (t == u) ? mk[u] : mkt;
      if (q < 2) goff[((size_t)n * 2 * g.K + 2 * tap + q) * HoWo + p] = tot * mkt;    // assigned, like :433
      else if (gmask) gmask[((size_t)n * g.K + tap) * HoWo + p] = tot;
    }
  }
}

// ---- col2im through the bulk-copy engine (default for C = 128 / 256; KGDET_COL2IM_BULK=0 -> col2im_tc_kernel) ----
// Same mapping as col2im_tc_kernel (one warp per position and group of 5 taps, deterministic offset / mask gradient),
// but the weighted column-gradient row of a (sample, corner) is staged in shared memory (C floats) and added to the
// NHWC gradient with ONE cp.reduce.async.bulk .add.f32 of C * 4 bytes issued by lane 0, instead of C / 4
// red.global.add.v4.f32 issued by the lanes (the SM issues one red per 1.29 clk and lane: 0.86 ms per K = 49 call at
// batch 16; the bulk engine's reductions reach L2 at 47 % of the SM -> L2 write path instead).  K = 49 call,
// forward + backward: 1 722 -> 1 422 us; K = 25: 1 002 -> 850; K = 9: 555 -> 494 (profiles/r2_col2im_bulk.txt).
__device__ __forceinline__ void unpack4b(const uint2& v, float (&f)[4]) {
  f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
  f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
}
// Loads run one tap ahead of the staging (the asm statements of the staging are memory barriers for the compiler:
// without the explicit prefetch every tap's loads waited for the previous tap's bulk issue -- long_scoreboard 8.2
// stalled warps per issue, 860 -> 760 us only).  NCH = C / 128 (4 channels per lane and chunk).
template <int NCH, int BULK_SETS, bool STREAM_CG>
__global__ void __launch_bounds__(256, 2)
col2im_bulk_kernel(DcnGeom g, const __nv_bfloat16* __restrict__ cg, const __nv_bfloat16* __restrict__ in,
                   const SampleRec* __restrict__ plan, const SampleAux* __restrict__ aux, int tap0, int ntaps,
                   float* __restrict__ gin, float* __restrict__ goff, float* __restrict__ gmask) {
  constexpr int TB = COL2IM_TB;
  constexpr int C = NCH * 128;
  extern __shared__ __align__(128) float bulk_stage[];          // [8 warps][BULK_SETS][4][C]
  const int lane = threadIdx.x & 31;
  float* my = bulk_stage + (size_t)(threadIdx.x >> 5) * (BULK_SETS * 4 * C);
  const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int HoWo = g.Ho * g.Wo;
  const int ngroups = (ntaps + TB - 1) / TB;
  const long long total = (long long)g.M * ngroups;
  int set = 0;
  for (long long wi = warp_global; wi < total; wi += nwarps) {
    const int m = (int)(wi / ngroups), tl0 = (int)(wi - (long long)m * ngroups) * TB;
    float part[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) part[i] = 0.f;
    int4 pa[TB];
    float4 a4[TB];
#pragma unroll
    for (int t = 0; t < TB; ++t) {
      pa[t] = make_int4(0, 0, 0, 0);
      a4[t] = make_float4(0.f, 0.f, 0.f, 0.f);                    // valid = 0
      if (tl0 + t < ntaps) {                                      // warp-uniform
        const size_t ridx = (size_t)m * g.K + tap0 + tl0 + t;
        pa[t] = __ldg(reinterpret_cast<const int4*>(plan + ridx));
        a4[t] = __ldg(reinterpret_cast<const float4*>(aux + ridx));
      }
    }
    uint2 gq[2][NCH], vq[2][NCH][4];
    auto load_tap = [&](int t, int slot) {
      const int valid = __float_as_int(a4[t].w);
      const int pix[4] = {pa[t].x, pa[t].y, pa[t].z, pa[t].w};
      const __nv_bfloat16* cgrow = cg + ((size_t)m * ntaps + tl0 + t) * C;
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        gq[slot][ch] = valid ? (STREAM_CG ? __ldcs(reinterpret_cast<const uint2*>(cgrow + ch * 128 + lane * 4))
                                          : __ldg(reinterpret_cast<const uint2*>(cgrow + ch * 128 + lane * 4)))
                             : make_uint2(0u, 0u);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          vq[slot][ch][i] = (valid & (1 << i))
                                ? __ldg(reinterpret_cast<const uint2*>(in + (size_t)pix[i] * C + ch * 128 + lane * 4))
                                : make_uint2(0u, 0u);
      }
    };
    load_tap(0, 0);
#pragma unroll
    for (int t = 0; t < TB; ++t) {
      if (t + 1 < TB) load_tap(t + 1, (t + 1) & 1);
      const int valid = __float_as_int(a4[t].w);
      if (!valid) continue;                                       // warp-uniform
      const int pix[4] = {pa[t].x, pa[t].y, pa[t].z, pa[t].w};
      const float lh = a4[t].x, lw = a4[t].y, mkv = a4[t].z;
      const float hh = 1.f - lh, hw = 1.f - lw;
      const float wgt[4] = {hh * hw * mkv, hh * lw * mkv, lh * hw * mkv, lh * lw * mkv};   // as dcn_plan_kernel
      float d[4] = {0.f, 0.f, 0.f, 0.f};
      float* sbuf = my + (size_t)set * 4 * C;
      // the bulk reductions that last read this set have finished reading it
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(BULK_SETS - 1) : "memory");
      __syncwarp();
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        float gv[4];
        unpack4b(gq[t & 1][ch], gv);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (valid & (1 << i)) {
            float v[4];
            unpack4b(vq[t & 1][ch][i], v);
            const float wi_ = wgt[i];
            *reinterpret_cast<float4*>(sbuf + i * C + ch * 128 + lane * 4) =
                make_float4(gv[0] * wi_, gv[1] * wi_, gv[2] * wi_, gv[3] * wi_);
            d[i] = fmaf(gv[3], v[3], fmaf(gv[2], v[2], fmaf(gv[1], v[1], fmaf(gv[0], v[0], d[i]))));
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (valid & (1 << i))
            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(
                             gin + (size_t)pix[i] * C),
                         "r"(smem_u32(sbuf + i * C)), "n"(C * 4)
                         : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      set = (set + 1 == BULK_SETS) ? 0 : set + 1;
      part[3 * t + 0] = -hw * d[0] - lw * d[1] + hw * d[2] + lw * d[3];
      part[3 * t + 1] = -hh * d[0] + hh * d[1] - lh * d[2] + lh * d[3];
      part[3 * t + 2] = hh * hw * d[0] + hh * lw * d[1] + lh * hw * d[2] + lh * lw * d[3];
    }
#pragma unroll
    for (int lvl = 0; lvl < 4; ++lvl) {
      const int bit = 16 >> lvl, half = 8 >> lvl;
      const bool up = (lane & bit) != 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j < half) {
          const float send = up ? part[j] : part[j + half];
          const float keep = up ? part[j + half] : part[j];
          part[j] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
        }
      }
    }
    const float tot = part[0] + __shfl_xor_sync(0xffffffffu, part[0], 1);
    const int idx = lane >> 1;
    const int t = idx / 3, q = idx - 3 * t;
    if ((lane & 1) == 0 && idx < 3 * TB && tl0 + t < ntaps) {
      const int tap = tap0 + tl0 + t;
      const int n = m / HoWo, p = m - n * HoWo;
      float mkt = a4[0].z;
#pragma unroll
      for (int u = 1; u < TB; ++u) mkt = (t == u) ? a4[u].z : mkt;
      if (q < 2) goff[((size_t)n * 2 * g.K + 2 * tap + q) * HoWo + p] = tot * mkt;
      else if (gmask) gmask[((size_t)n * g.K + tap) * HoWo + p] = tot;
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // staging rows must outlive the copies
  __syncwarp();
}
static bool use_bulk_col2im(const DcnGeom& g) {
  if (const char* e = getenv("KGDET_COL2IM_BULK"))
    if (atoi(e) == 0) return false;
  return g.C == 128 || g.C == 256;
}

// ---- sampled columns, transposed: colT[(tl*C + c), m] for taps tap0 .. tap0+ntaps-1 ---------------
// CTA = 64 positions x one tap; loops over 64-channel blocks.  Gather is row-wise coalesced (8 lanes per
// 128-byte slab), the 64x64 tile is transposed through shared memory, rows of 64 positions (128 B) go out.
__global__ void __launch_bounds__(256)
gather_colT_kernel(DcnGeom g, const __nv_bfloat16* __restrict__ in, const SampleRec* __restrict__ plan,
                   int tap0, __nv_bfloat16* __restrict__ colT, long long mpad) {
  __shared__ __nv_bfloat16 tile[64][64 + 8];       // [channel][position]
  const int m0 = blockIdx.x * 64, tl = blockIdx.y, tap = tap0 + tl;
  const int t = threadIdx.x, chunk = t & 7, rbase = t >> 3;
  SampleRec rec[2];
#pragma unroll
  for (int ps = 0; ps < 2; ++ps) {
    const int m = m0 + rbase + ps * 32;
    const SampleRec* rp = plan + (size_t)m * g.K + tap;          // plan is padded to 128 rows
    const int4 a = __ldg(reinterpret_cast<const int4*>(rp));
    const float4 b = __ldg(reinterpret_cast<const float4*>(rp) + 1);
    rec[ps].pix[0] = a.x; rec[ps].pix[1] = a.y; rec[ps].pix[2] = a.z; rec[ps].pix[3] = a.w;
    rec[ps].w[0] = b.x; rec[ps].w[1] = b.y; rec[ps].w[2] = b.z; rec[ps].w[3] = b.w;
  }
  for (int cb = 0; cb < g.C / 64; ++cb) {
#pragma unroll
    for (int ps = 0; ps < 2; ++ps) {
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (rec[ps].w[i] != 0.f) {
          float v[8];
          unpack8(__ldg(reinterpret_cast<const uint4*>(in + (size_t)rec[ps].pix[i] * g.C + cb * 64 + chunk * 8)), v);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] = fmaf(rec[ps].w[i], v[e], acc[e]);
        }
      }
      const int pos = rbase + ps * 32;
#pragma unroll
      for (int e = 0; e < 8; ++e) tile[chunk * 8 + e][pos] = __float2bfloat16(acc[e]);
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int idx = t + r * 256;               // 512 x 16-byte pieces: 64 channels x 8 pieces
      const int c = idx >> 3, piece = idx & 7;
      const uint4 v = *reinterpret_cast<const uint4*>(&tile[c][piece * 8]);
      *reinterpret_cast<uint4*>(colT + ((size_t)tl * g.C + cb * 64 + c) * mpad + m0 + piece * 8) = v;
    }
    __syncthreads();
  }
}

// gO NCHW (fp32 / bf16) -> gOT[o, m] bf16 with m = n*HoWo + p (row stride mpad)
template <typename T>
__global__ void go_to_goT_kernel(const T* __restrict__ go, __nv_bfloat16* __restrict__ goT, int N, int Cout,
                                 int HoWo, long long mpad) {
  const long long total = (long long)N * Cout * HoWo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)blockDim.x * gridDim.x) {
    const int p = (int)(i % HoWo);
    const long long r = i / HoWo;
    const int o = (int)(r % Cout), n = (int)(r / Cout);
    float v;
    if constexpr (sizeof(T) == 4) v = go[i]; else v = __bfloat162float(go[i]);
    goT[(size_t)o * mpad + (size_t)n * HoWo + p] = __float2bfloat16(v);
  }
}

// The two weight-layout changes of the backward as 32 x 32 shared-memory transposes (one channel c per CTA row):
// element-per-thread versions read or wrote with a stride of C * Cout floats (one 4-byte access per 128-byte line)
// and together cost ~25 us per K = 49 call, every training step (weights change).
//   Wd[(tap*C + c), o] = W[o, c, tap]  (bf16, operand of the column-gradient GEMM)
//   grad_W[o, c, tap] = scale * gWT[(tap*C + c), o]
__global__ void __launch_bounds__(256) pack_wd_tiled_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wd,
                                                            int Cout, int C, int K) {
  __shared__ float tile[32][33];                       // [o][tap]
  const int c = blockIdx.y, o0 = blockIdx.x * 32, tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
  for (int t0 = 0; t0 < K; t0 += 32) {
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
      const int o = o0 + ty + k, t = t0 + tx;
      tile[ty + k][tx] = (o < Cout && t < K) ? w[((size_t)o * C + c) * K + t] : 0.f;      // contiguous over taps
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
      const int t = t0 + ty + k, o = o0 + tx;
      if (t < K && o < Cout) wd[((size_t)t * C + c) * Cout + o] = __float2bfloat16(tile[tx][ty + k]);   // contiguous over o
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) unpack_gw_tiled_kernel(const float* __restrict__ gwt, float* __restrict__ gw,
                                                              int Cout, int C, int K, float scale) {
  __shared__ float tile[32][33];                       // [tap][o]
  const int c = blockIdx.y, o0 = blockIdx.x * 32, tx = threadIdx.x, ty = threadIdx.y;
  for (int t0 = 0; t0 < K; t0 += 32) {
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
      const int t = t0 + ty + k, o = o0 + tx;
      tile[ty + k][tx] = (t < K && o < Cout) ? gwt[((size_t)t * C + c) * Cout + o] : 0.f;   // contiguous over o
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
      const int o = o0 + ty + k, t = t0 + tx;
      if (o < Cout && t < K) gw[((size_t)o * C + c) * K + t] = scale * tile[tx][ty + k];    // contiguous over taps
    }
    __syncthreads();
  }
}

static void launch_pack_wd(const float* w, __nv_bfloat16* wd, int Cout, int C, int K, cudaStream_t stream) {
  pack_wd_tiled_kernel<<<dim3((unsigned)ceil_div(Cout, 32), (unsigned)C), dim3(32, 8), 0, stream>>>(w, wd, Cout, C, K);
}
static void launch_unpack_gw(const float* gwt, float* gw, int Cout, int C, int K, float scale, cudaStream_t stream) {
  unpack_gw_tiled_kernel<<<dim3((unsigned)ceil_div(Cout, 32), (unsigned)C), dim3(32, 8), 0, stream>>>(gwt, gw, Cout, C, K,
                                                                                                   scale);
}

template <typename T>
__global__ void bias_grad_nchw_kernel(const T* __restrict__ go, int N, int Cout, int HoWo, float* __restrict__ gb) {
  const int o = blockIdx.x;
  float s = 0.f;
  for (int n = 0; n < N; ++n)
    for (int p = threadIdx.x; p < HoWo; p += blockDim.x) {
      const size_t i = ((size_t)n * Cout + o) * HoWo + p;
      if constexpr (sizeof(T) == 4) s += go[i]; else s += __bfloat162float(go[i]);
    }
  __shared__ float part[32];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) gb[o] = v;
  }
}

// ---- workspace layouts --------------------------------------------------------------------------------
struct TcBwdInWs {
  __nv_bfloat16 *in_nhwc, *go_nhwc, *wd, *cg;
  float* gin_nhwc;
  float* gpart;               // owned-slice col2im: offset / mask gradient parts per channel block
  unsigned* sched;            // owned-slice col2im: lane schedule per (image, tap, tile)
  SampleRec* plan;
  SampleAux* aux;
  size_t total;
};
static TcBwdInWs carve_tc_in(const DcnGeom& g, void* ws) {
  TcBwdInWs w;
  size_t off = 0;
  auto take = [&](size_t bytes) { void* p = ws ? (char*)ws + off : nullptr; off += align_up(bytes, 1024); return p; };
  w.in_nhwc = (__nv_bfloat16*)take((size_t)g.N * g.H * g.W * g.C * 2);
  w.go_nhwc = (__nv_bfloat16*)take((size_t)g.M * g.Cout * 2);
  w.wd = (__nv_bfloat16*)take((size_t)g.K * g.C * g.Cout * 2);
  w.cg = (__nv_bfloat16*)take((size_t)g.M * taps_per_chunk(g, g.M) * g.C * 2);
  w.gin_nhwc = (float*)take((size_t)g.N * g.H * g.W * g.C * 4);
  w.gpart = (float*)take(col2im_own_part_bytes(g));
  w.sched = (unsigned*)take(col2im_own_sched_bytes(g));
  w.plan = (SampleRec*)take(plan_bytes(g));
  w.aux = (SampleAux*)take(plan_aux_bytes(g));
  w.total = off;
  return w;
}
size_t bwd_tc_input_workspace_bytes(const DcnGeom& g) { return carve_tc_in(g, nullptr).total; }

struct TcBwdWWs {
  __nv_bfloat16 *in_nhwc, *goT, *colT;
  float* gwt;
  SampleRec* plan;
  size_t total;
};
static TcBwdWWs carve_tc_w(const DcnGeom& g, void* ws) {
  TcBwdWWs w;
  size_t off = 0;
  auto take = [&](size_t bytes) { void* p = ws ? (char*)ws + off : nullptr; off += align_up(bytes, 1024); return p; };
  const size_t mp = mpad64(g);
  w.in_nhwc = (__nv_bfloat16*)take((size_t)g.N * g.H * g.W * g.C * 2);
  w.goT = (__nv_bfloat16*)take((size_t)g.Cout * mp * 2);
  w.colT = (__nv_bfloat16*)take((size_t)taps_per_chunk(g, mp) * g.C * mp * 2);
  w.gwt = (float*)take((size_t)g.K * g.C * g.Cout * 4);
  w.plan = (SampleRec*)take(plan_bytes(g));
  w.total = off;
  return w;
}
// fused path: channel-blocked bf16 planes with guard bands (as the forward), compact plan, tiled gO, gW^T
struct TcWFusedWs {
  unsigned char* planes;
  size_t guard_bytes, in_bytes, plane_bytes;
  SampleRec16* plan;
  unsigned char* go_tiled;
  float* gwt;
  size_t total;
};
static TcWFusedWs carve_tc_w_fused(const DcnGeom& g, void* ws) {
  TcWFusedWs w;
  size_t off = 0;
  auto take = [&](size_t bytes) { void* p = ws ? (char*)ws + off : nullptr; off += align_up(bytes, 1024); return p; };
  w.guard_bytes = (size_t)dcn_guard_pixels(g) * 128;
  w.in_bytes = (size_t)g.N * g.H * g.W * 128;
  w.plane_bytes = align_up(w.in_bytes + 2 * w.guard_bytes, 1024);
  w.planes = (unsigned char*)take(w.plane_bytes * (g.C / 64));
  w.plan = (SampleRec16*)take(plan16_bytes(g));
  w.go_tiled = (unsigned char*)take(wgrad_fused_go_bytes(g));
  w.gwt = (float*)take((size_t)g.K * g.C * g.Cout * 4);
  w.total = off;
  return w;
}
static bool use_fused_wgrad(const DcnGeom& g) {
  if (const char* e = getenv("KGDET_WGRAD_FUSED")) return atoi(e) != 0 && wgrad_fused_supported(g);
  return wgrad_fused_supported(g);
}
size_t bwd_tc_weight_workspace_bytes(const DcnGeom& g) {
  const size_t a = carve_tc_w(g, nullptr).total;
  const size_t b = wgrad_fused_supported(g) ? carve_tc_w_fused(g, nullptr).total : 0;
  return a > b ? a : b;
}

static int grid_for(long long total, int threads) {
  long long b = (total + threads - 1) / threads;
  const long long cap = (long long)num_sms() * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

int bwd_tc_input(const DcnGeom& g, const void* input, const float* offset, const float* mask,
                 const float* weight, const void* grad_output, void* grad_input, float* grad_offset,
                 float* grad_mask, int dtype, void* ws, cudaStream_t stream) {
  TcBwdInWs w = carve_tc_in(g, ws);
  const int HW = g.H * g.W, HoWo = g.Ho * g.Wo;
  int rc;
  if ((rc = launch_transpose(input, w.in_nhwc, g.N, g.C, HW, dtype, KGDET_BF16, stream)) != KGDET_OK) return rc;
  if ((rc = launch_transpose(grad_output, w.go_nhwc, g.N, g.Cout, HoWo, dtype, KGDET_BF16, stream)) != KGDET_OK) return rc;
  if ((rc = launch_plan(g, offset, mask, w.plan, w.aux, stream)) != KGDET_OK) return rc;
  launch_pack_wd(weight, w.wd, g.Cout, g.C, g.K, stream);
  KG_LAUNCH_CHECK("pack_wd_tiled_kernel");
  KG_CUDA(cudaMemsetAsync(w.gin_nhwc, 0, (size_t)g.N * HW * g.C * 4, stream));
  const int tpc = taps_per_chunk(g, g.M);
  // small maps: a CTA owns its slice of the input gradient in shared memory (dcn_col2im_own.cu)
  const bool own = col2im_own_supported(g);
  if (own && (rc = col2im_own_schedule(g, w.plan, w.aux, w.sched, stream)) != KGDET_OK) return rc;
  for (int tap0 = 0; tap0 < g.K; tap0 += tpc) {
    const int nt = (g.K - tap0) < tpc ? (g.K - tap0) : tpc;
    // cg[m, (tl, c)] = go_nhwc[m, :] . Wd[(tap0+tl)*C + c, :]
    if ((rc = umma_gemm(w.go_nhwc, g.Cout, w.wd + (size_t)tap0 * g.C * g.Cout, g.Cout, w.cg, (long long)nt * g.C,
                        g.M, nt * g.C, g.Cout, KGDET_BF16, 1, 1.f, stream)) != KGDET_OK) return rc;
    if (own) {
      if ((rc = col2im_own(g, w.cg, w.in_nhwc, w.plan, w.aux, w.sched, tap0, nt, w.gin_nhwc, w.gpart, stream)) != KGDET_OK)
        return rc;
      continue;
    }
    const long long warps = (long long)g.M * ceil_div(nt, COL2IM_TB);
    if (use_bulk_col2im(g)) {
      // two staging sets per warp (three measured the same: 1 464 vs 1 479 us for the K = 49 call)
      const size_t smem = (size_t)8 * 2 * 4 * g.C * 4;
      bool stream_cg = true;
      if (const char* e = getenv("KGDET_COL2IM_STREAM_CG")) stream_cg = atoi(e) != 0;
      auto kern = g.C == 256 ? (stream_cg ? col2im_bulk_kernel<2, 2, true> : col2im_bulk_kernel<2, 2, false>)
                             : (stream_cg ? col2im_bulk_kernel<1, 2, true> : col2im_bulk_kernel<1, 2, false>);
      KG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      long long b = (warps * 32 + 255) / 256;
      const long long cap = (long long)num_sms() * 2;
      kern<<<(int)(b > cap ? cap : b), 256, smem, stream>>>(g, w.cg, w.in_nhwc, w.plan, w.aux, tap0, nt, w.gin_nhwc,
                                                           grad_offset, grad_mask);
      KG_LAUNCH_CHECK("col2im_bulk_kernel");
      continue;
    }
    col2im_tc_kernel<<<grid_for(warps * 32, 256), 256, 0, stream>>>(g, w.cg, w.in_nhwc, w.plan, w.aux, tap0, nt,
                                                                    w.gin_nhwc, grad_offset, grad_mask);
    KG_LAUNCH_CHECK("col2im_tc_kernel");
  }
  if (own && (rc = col2im_own_finish(g, w.gpart, grad_offset, grad_mask, stream)) != KGDET_OK) return rc;
  return launch_transpose(w.gin_nhwc, grad_input, g.N, HW, g.C, KGDET_F32, dtype, stream);
}

int bwd_tc_weight(const DcnGeom& g, const void* input, const float* offset, const float* mask,
                  const void* grad_output, float* grad_weight, float* grad_bias, float scale, int dtype,
                  void* ws, cudaStream_t stream) {
  const int HW = g.H * g.W, HoWo = g.Ho * g.Wo;
  int rc;
  if (use_fused_wgrad(g)) {
    TcWFusedWs f = carve_tc_w_fused(g, ws);
    const int planes = g.C / 64;
    KG_CUDA(cudaMemset2DAsync(f.planes, f.plane_bytes, 0, f.guard_bytes, planes, stream));
    KG_CUDA(cudaMemset2DAsync(f.planes + f.guard_bytes + f.in_bytes, f.plane_bytes, 0,
                              f.plane_bytes - f.guard_bytes - f.in_bytes, planes, stream));
    if ((rc = launch_nchw_to_blocked(input, f.planes + f.guard_bytes, g.N, g.C, HW, 64, f.plane_bytes, dtype,
                                     KGDET_BF16, stream)) != KGDET_OK) return rc;
    if ((rc = launch_plan16(g, offset, mask, f.plan, PLAN16_BF16W, stream)) != KGDET_OK) return rc;
    if ((rc = launch_go_to_tiled(g, grad_output, f.go_tiled, dtype, stream)) != KGDET_OK) return rc;
    KG_CUDA(cudaMemsetAsync(f.gwt, 0, (size_t)g.K * g.C * g.Cout * 4, stream));
    if ((rc = wgrad_fused(g, f.planes + f.guard_bytes, f.plane_bytes, f.plan, f.go_tiled, f.gwt, stream)) != KGDET_OK)
      return rc;
    launch_unpack_gw(f.gwt, grad_weight, g.Cout, g.C, g.K, scale, stream);
    KG_LAUNCH_CHECK("unpack_gw_tiled_kernel");
    if (grad_bias) {
      if (dtype == KGDET_F32)
        bias_grad_nchw_kernel<float><<<g.Cout, 256, 0, stream>>>((const float*)grad_output, g.N, g.Cout, HoWo, grad_bias);
      else
        bias_grad_nchw_kernel<__nv_bfloat16><<<g.Cout, 256, 0, stream>>>((const __nv_bfloat16*)grad_output, g.N, g.Cout, HoWo, grad_bias);
      KG_LAUNCH_CHECK("bias_grad_nchw_kernel");
    }
    return KGDET_OK;
  }
  TcBwdWWs w = carve_tc_w(g, ws);
  const long long mp = (long long)mpad64(g);
  if ((rc = launch_transpose(input, w.in_nhwc, g.N, g.C, HW, dtype, KGDET_BF16, stream)) != KGDET_OK) return rc;
  if ((rc = launch_plan(g, offset, mask, w.plan, nullptr, stream)) != KGDET_OK) return rc;
  KG_CUDA(cudaMemsetAsync(w.goT, 0, (size_t)g.Cout * mp * 2, stream));
  const long long gototal = (long long)g.N * g.Cout * HoWo;
  if (dtype == KGDET_F32)
    go_to_goT_kernel<float><<<grid_for(gototal, 256), 256, 0, stream>>>((const float*)grad_output, w.goT, g.N, g.Cout, HoWo, mp);
  else
    go_to_goT_kernel<__nv_bfloat16><<<grid_for(gototal, 256), 256, 0, stream>>>((const __nv_bfloat16*)grad_output, w.goT, g.N, g.Cout, HoWo, mp);
  KG_LAUNCH_CHECK("go_to_goT_kernel");
  KG_CUDA(cudaMemsetAsync(w.gwt, 0, (size_t)g.K * g.C * g.Cout * 4, stream));
  const int tpc = taps_per_chunk(g, (size_t)mp);
  for (int tap0 = 0; tap0 < g.K; tap0 += tpc) {
    const int nt = (g.K - tap0) < tpc ? (g.K - tap0) : tpc;
    gather_colT_kernel<<<dim3((unsigned)(mp / 64), nt), 256, 0, stream>>>(g, w.in_nhwc, w.plan, tap0, w.colT, mp);
    KG_LAUNCH_CHECK("gather_colT_kernel");
    // gWT[(tap0+tl)*C + c, o] += sum_m colT[(tl*C + c), m] * goT[o, m]     (split over m)
    const int mtiles = ceil_div(nt * g.C, 128) * ceil_div(g.Cout, 256);
    // one wave of CTAs: every extra split is another 128 x 256 fp32 reduction epilogue into L2
    int splits = num_sms() / mtiles;
    const int kblocks = (int)(mp / 64);
    if (splits > kblocks) splits = kblocks;
    if (splits < 1) splits = 1;
    if ((rc = umma_gemm(w.colT, mp, w.goT, mp, w.gwt + (size_t)tap0 * g.C * g.Cout, g.Cout, nt * g.C, g.Cout,
                        (int)mp, KGDET_F32, splits < 2 ? 2 : splits, 1.f, stream)) != KGDET_OK) return rc;
  }
  launch_unpack_gw(w.gwt, grad_weight, g.Cout, g.C, g.K, scale, stream);
  KG_LAUNCH_CHECK("unpack_gw_tiled_kernel");
  if (grad_bias) {
    if (dtype == KGDET_F32)
      bias_grad_nchw_kernel<float><<<g.Cout, 256, 0, stream>>>((const float*)grad_output, g.N, g.Cout, HoWo, grad_bias);
    else
      bias_grad_nchw_kernel<__nv_bfloat16><<<g.Cout, 256, 0, stream>>>((const __nv_bfloat16*)grad_output, g.N, g.Cout, HoWo, grad_bias);
    KG_LAUNCH_CHECK("bias_grad_nchw_kernel");
  }
  return KGDET_OK;
}

}  // namespace kgdet
