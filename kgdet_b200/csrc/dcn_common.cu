// Sampling plan + layout transforms shared by every deformable-convolution path.
#include "dcn.cuh"

namespace kgdet {

size_t plan_rows(const DcnGeom& g) { return (size_t)ceil_div(g.M, 256) * 256; }   // a CTA pair covers 256 rows
size_t plan_bytes(const DcnGeom& g) {
  return plan_rows(g) * g.dgroups * g.K * sizeof(SampleRec);
}
size_t plan_aux_bytes(const DcnGeom& g) {
  return plan_rows(g) * g.dgroups * g.K * sizeof(SampleAux);
}

// Sampling rule of deformable_im2col_gpu_kernel (deform_conv_cuda_kernel.cu:210-236) and
// deformable_im2col_bilinear (:83-114), evaluated once per (position, dgroup, tap).
//
// OffsetSrc: where the (dy, dx) pairs come from.  Plain case: the op's `offset` tensor [N, dg*2K, Ho, Wo].
// Points case (kgdet_dcn_prepare_plan_points): a channel slice of a wider point-set tensor whose values are
// absolute point offsets; the head turns them into DCN offsets with `pts - base`, base = the regular k x k
// grid in [-(k-1)/2, (k-1)/2] (KP3:37-67,135-143) -- done here in the same fp32 operation order.
struct OffsetSrc {
  const float* ptr;          // first channel used
  long long batch_stride;    // elements between images
  int points;                // 1: subtract the base grid
  float gm, gm1;             // points case, gm != 0: the head's gradient-mul expression first (KP3:135-143):
                             // p' = gm * p + (1 - gm) * p.detach() -- the identity up to fp32 rounding, evaluated
                             // here in the reference's operation order so that the sampled locations are bit-identical
};
__device__ __forceinline__ void load_offset(const OffsetSrc& o, const DcnGeom& g, int n, int dgi, int tap,
                                            int i, int j, int HoWo, int p, float& off_h, float& off_w) {
  const float* q = o.ptr + (long long)n * o.batch_stride + ((size_t)dgi * 2 * g.K + 2 * tap) * HoWo + p;
  off_h = q[0];                                          // :221,223
  off_w = q[HoWo];                                       // :222,224
  if (o.points) {
    if (o.gm != 0.f) {
      off_h = __fadd_rn(__fmul_rn(o.gm, off_h), __fmul_rn(o.gm1, off_h));
      off_w = __fadd_rn(__fmul_rn(o.gm, off_w), __fmul_rn(o.gm1, off_w));
    }
    off_h = off_h - (float)(i - (g.kh - 1) / 2);
    off_w = off_w - (float)(j - (g.kw - 1) / 2);
  }
}

__global__ void dcn_plan_kernel(DcnGeom g, OffsetSrc osrc,
                                const float* __restrict__ mask, SampleRec* __restrict__ rec,
                                SampleAux* __restrict__ aux, int rows_padded) {
  const int per_row = g.dgroups * g.K;
  const long long total = (long long)rows_padded * per_row;
  const int HoWo = g.Ho * g.Wo;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)blockDim.x * gridDim.x) {
    const int m = (int)(idx / per_row);
    const int rem = (int)(idx - (long long)m * per_row);
    SampleRec r;
    r.pix[0] = r.pix[1] = r.pix[2] = r.pix[3] = 0;
    r.w[0] = r.w[1] = r.w[2] = r.w[3] = 0.f;
    SampleAux a;
    a.lh = a.lw = 0.f; a.mask = 0.f; a.valid = 0;
    if (m < g.M) {
      const int dgi = rem / g.K, tap = rem - dgi * g.K;
      const int n = m / HoWo, p = m - n * HoWo;
      const int y = p / g.Wo, x = p - y * g.Wo;
      const int i = tap / g.kw, j = tap - i * g.kw;
      float off_h, off_w;
      load_offset(osrc, g, n, dgi, tap, i, j, HoWo, p, off_h, off_w);
      const float mval = mask ? mask[((size_t)(n * g.dgroups + dgi) * g.K + tap) * HoWo + p] : 1.f;
      const float h_im = (float)(y * g.sh - g.ph + i * g.dh) + off_h;   // :226
      const float w_im = (float)(x * g.sw - g.pw + j * g.dw) + off_w;   // :227
      a.mask = mval;
      if (h_im > -1.f && w_im > -1.f && h_im < (float)g.H && w_im < (float)g.W) {   // :228
        const float hf = floorf(h_im), wf = floorf(w_im);
        const int h_low = (int)hf, w_low = (int)wf;
        const int h_high = h_low + 1, w_high = w_low + 1;
        const float lh = h_im - hf, lw = w_im - wf;
        const float hh = 1.f - lh, hw = 1.f - lw;
        const bool vhl = h_low >= 0, vhh = h_high <= g.H - 1;          // :98-107
        const bool vwl = w_low >= 0, vwh = w_high <= g.W - 1;
        const int nbase = n * g.H * g.W;
        a.lh = lh; a.lw = lw;
        if (vhl && vwl) { r.pix[0] = nbase + h_low * g.W + w_low;   r.w[0] = hh * hw * mval; a.valid |= 1; }
        if (vhl && vwh) { r.pix[1] = nbase + h_low * g.W + w_high;  r.w[1] = hh * lw * mval; a.valid |= 2; }
        if (vhh && vwl) { r.pix[2] = nbase + h_high * g.W + w_low;  r.w[2] = lh * hw * mval; a.valid |= 4; }
        if (vhh && vwh) { r.pix[3] = nbase + h_high * g.W + w_high; r.w[3] = lh * lw * mval; a.valid |= 8; }
      }
    }
    rec[idx] = r;
    if (aux) aux[idx] = a;
  }
}

size_t plan16_bytes(const DcnGeom& g) { return plan_rows(g) * g.K * sizeof(SampleRec16); }

// Same sampling rule as dcn_plan_kernel, compact output for the tensor-core path (dgroups == 1).
// Threads run fastest over positions so the offset reads are coalesced (offset is [N, 2K, Ho, Wo]).
__global__ void dcn_plan16_kernel(DcnGeom g, OffsetSrc osrc,
                                  const float* __restrict__ mask, SampleRec16* __restrict__ rec,
                                  int rows_padded, int fmt) {
  const long long total = (long long)rows_padded * g.K;
  const int HoWo = g.Ho * g.Wo;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)blockDim.x * gridDim.x) {
    const int tap = (int)(idx / rows_padded);
    const int m = (int)(idx - (long long)tap * rows_padded);
    int base = 0;
    float lh = 0.f, lw = 0.f, scale = 0.f;
    unsigned vh = 0u, vw = 0u;
    if (m < g.M) {
      const int n = m / HoWo, p = m - n * HoWo;
      const int y = p / g.Wo, x = p - y * g.Wo;
      const int i = tap / g.kw, j = tap - i * g.kw;
      float off_h, off_w;
      load_offset(osrc, g, n, 0, tap, i, j, HoWo, p, off_h, off_w);
      const float mval = mask ? mask[((size_t)n * g.K + tap) * HoWo + p] : 1.f;
      const float h_im = (float)(y * g.sh - g.ph + i * g.dh) + off_h;
      const float w_im = (float)(x * g.sw - g.pw + j * g.dw) + off_w;
      base = n * g.H * g.W;                         // safe address for samples outside the window
      if (h_im > -1.f && w_im > -1.f && h_im < (float)g.H && w_im < (float)g.W) {
        const float hf = floorf(h_im), wf = floorf(w_im);
        const int h_low = (int)hf, w_low = (int)wf;
        lh = h_im - hf;
        lw = w_im - wf;
        if (h_low >= 0) vh |= 1u;
        if (h_low + 1 <= g.H - 1) vh |= 2u;
        if (w_low >= 0) vw |= 1u;
        if (w_low + 1 <= g.W - 1) vw |= 2u;
        base = n * g.H * g.W + h_low * g.W + w_low;  // within the guard band by construction
        scale = mval;
      }
    }
    SampleRec16 r;
    r.base = base;
    if (fmt == PLAN16_BF16W) {
      const float hh = 1.f - lh, hw = 1.f - lw;
      const float w0 = ((vh & 1u) && (vw & 1u)) ? hh * hw * scale : 0.f;
      const float w1 = ((vh & 1u) && (vw & 2u)) ? hh * lw * scale : 0.f;
      const float w2 = ((vh & 2u) && (vw & 1u)) ? lh * hw * scale : 0.f;
      const float w3 = ((vh & 2u) && (vw & 2u)) ? lh * lw * scale : 0.f;
      const unsigned p01 = ((unsigned)__bfloat16_as_ushort(__float2bfloat16(w1)) << 16) |
                           (unsigned)__bfloat16_as_ushort(__float2bfloat16(w0));
      const unsigned p23 = ((unsigned)__bfloat16_as_ushort(__float2bfloat16(w3)) << 16) |
                           (unsigned)__bfloat16_as_ushort(__float2bfloat16(w2));
      r.lh = __uint_as_float(p01);
      r.lw = __uint_as_float(p23);
      r.scale = 0.f;
    } else {
      r.lh = __uint_as_float((__float_as_uint(lh) & ~3u) | vh);
      r.lw = __uint_as_float((__float_as_uint(lw) & ~3u) | vw);
      r.scale = scale;
    }
    rec[idx] = r;                                    // tap-major: [K][rows_padded]
  }
}

static OffsetSrc make_offset_src(const DcnGeom& g, const float* offset, long long batch_stride, int points,
                                 float gm = 0.f, float gm1 = 0.f) {
  OffsetSrc o;
  o.gm = gm; o.gm1 = gm1;
  o.ptr = offset;
  o.batch_stride = batch_stride > 0 ? batch_stride : (long long)g.dgroups * 2 * g.K * g.Ho * g.Wo;
  o.points = points;
  return o;
}

int launch_plan16(const DcnGeom& g, const float* offset, const float* mask, SampleRec16* rec, int fmt,
                  cudaStream_t stream, long long batch_stride, int points, float gm, float gm1) {
  const int rows = (int)plan_rows(g);
  const long long total = (long long)rows * g.K;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  dcn_plan16_kernel<<<(int)blocks, 256, 0, stream>>>(g, make_offset_src(g, offset, batch_stride, points, gm, gm1), mask,
                                                     rec, rows, fmt);
  KG_LAUNCH_CHECK("dcn_plan16_kernel");
  return KGDET_OK;
}

int launch_plan(const DcnGeom& g, const float* offset, const float* mask, SampleRec* rec,
                SampleAux* aux, cudaStream_t stream, long long batch_stride, int points, float gm, float gm1) {
  const int rows = (int)plan_rows(g);
  const long long total = (long long)rows * g.dgroups * g.K;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  dcn_plan_kernel<<<(int)blocks, 256, 0, stream>>>(g, make_offset_src(g, offset, batch_stride, points, gm, gm1), mask,
                                                   rec, aux, rows);
  KG_LAUNCH_CHECK("dcn_plan_kernel");
  return KGDET_OK;
}

// ---- batched 2-D transpose with dtype conversion:  src [B, R, Cc] -> dst [B, Cc, R] ----------
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16(v); }

template <typename Tin, typename Tout>
__global__ void transpose_kernel(const Tin* __restrict__ src, Tout* __restrict__ dst, int R, int Cc) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const Tin* s = src + (size_t)b * R * Cc;
  Tout* d = dst + (size_t)b * R * Cc;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    int r = r0 + ty + k, c = c0 + tx;
    if (r < R && c < Cc) tile[ty + k][tx] = to_f<Tin>(s[(size_t)r * Cc + c]);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    int c = c0 + ty + k, r = r0 + tx;
    if (r < R && c < Cc) d[(size_t)c * R + r] = from_f<Tout>(tile[tx][ty + k]);
  }
}

int launch_transpose(const void* src, void* dst, int B, int R, int Cc, int src_dtype, int dst_dtype,
                     cudaStream_t stream) {
  if (B <= 0 || R <= 0 || Cc <= 0) return KGDET_OK;
  KG_CHECK_ARG(B <= 65535 && ceil_div(R, 32) <= 65535, "transpose: batch/rows too large");
  dim3 grid(ceil_div(Cc, 32), ceil_div(R, 32), B), block(32, 8);
  if (src_dtype == KGDET_F32 && dst_dtype == KGDET_F32)
    transpose_kernel<float, float><<<grid, block, 0, stream>>>((const float*)src, (float*)dst, R, Cc);
  else if (src_dtype == KGDET_F32 && dst_dtype == KGDET_BF16)
    transpose_kernel<float, __nv_bfloat16><<<grid, block, 0, stream>>>((const float*)src, (__nv_bfloat16*)dst, R, Cc);
  else if (src_dtype == KGDET_BF16 && dst_dtype == KGDET_F32)
    transpose_kernel<__nv_bfloat16, float><<<grid, block, 0, stream>>>((const __nv_bfloat16*)src, (float*)dst, R, Cc);
  else if (src_dtype == KGDET_BF16 && dst_dtype == KGDET_BF16)
    transpose_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, block, 0, stream>>>((const __nv_bfloat16*)src, (__nv_bfloat16*)dst, R, Cc);
  else {
    set_error("transpose: bad dtype %d -> %d", src_dtype, dst_dtype);
    return KGDET_ERR_INVALID_ARG;
  }
  KG_LAUNCH_CHECK("transpose_kernel");
  return KGDET_OK;
}

// ---- NCHW -> channel-blocked planes (input of the tensor-core path) --------------------------------
// dst element of (n, c, p): plane c / bk, pixel n*S + p, channel c % bk.  A 32-channel tile never straddles a
// plane (bk is 32 or 64).
template <typename Tin, typename Tout>
__global__ void nchw_to_blocked_kernel(const Tin* __restrict__ src, Tout* __restrict__ dst, int C, int S, int bk,
                                       size_t plane_elems) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const Tin* s = src + (size_t)n * C * S;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int c = c0 + ty + k, p = p0 + tx;
    if (c < C && p < S) tile[ty + k][tx] = to_f<Tin>(s[(size_t)c * S + p]);
  }
  __syncthreads();
  Tout* d = dst + (size_t)(c0 / bk) * plane_elems + (c0 % bk);
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int p = p0 + ty + k, c = c0 + tx;
    if (c < C && p < S) d[((size_t)n * S + p) * bk + tx] = from_f<Tout>(tile[tx][ty + k]);
  }
}

int launch_nchw_to_blocked(const void* src, void* dst, int N, int C, int S, int bk, size_t plane_bytes,
                           int src_dtype, int dst_dtype, cudaStream_t stream) {
  if (N <= 0 || C <= 0 || S <= 0) return KGDET_OK;
  KG_CHECK_ARG(N <= 65535 && ceil_div(C, 32) <= 65535, "nchw_to_blocked: batch/channels too large");
  KG_CHECK_ARG(bk % 32 == 0 && C % bk == 0, "nchw_to_blocked: bad channel block %d for %d channels", bk, C);
  dim3 grid(ceil_div(S, 32), ceil_div(C, 32), N), block(32, 8);
  if (src_dtype == KGDET_F32 && dst_dtype == KGDET_F32)
    nchw_to_blocked_kernel<float, float><<<grid, block, 0, stream>>>((const float*)src, (float*)dst, C, S, bk, plane_bytes / 4);
  else if (src_dtype == KGDET_F32 && dst_dtype == KGDET_BF16)
    nchw_to_blocked_kernel<float, __nv_bfloat16><<<grid, block, 0, stream>>>((const float*)src, (__nv_bfloat16*)dst, C, S, bk, plane_bytes / 2);
  else if (src_dtype == KGDET_BF16 && dst_dtype == KGDET_F32)
    nchw_to_blocked_kernel<__nv_bfloat16, float><<<grid, block, 0, stream>>>((const __nv_bfloat16*)src, (float*)dst, C, S, bk, plane_bytes / 4);
  else if (src_dtype == KGDET_BF16 && dst_dtype == KGDET_BF16)
    nchw_to_blocked_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, block, 0, stream>>>((const __nv_bfloat16*)src, (__nv_bfloat16*)dst, C, S, bk, plane_bytes / 2);
  else {
    set_error("nchw_to_blocked: bad dtype %d -> %d", src_dtype, dst_dtype);
    return KGDET_ERR_INVALID_ARG;
  }
  KG_LAUNCH_CHECK("nchw_to_blocked_kernel");
  return KGDET_OK;
}

}  // namespace kgdet

extern "C" int kgdet_nchw_to_nhwc(const void* src, void* dst, int32_t N, int32_t C, int32_t S,
                                  int src_dtype, int dst_dtype, void* stream) {
  KG_CHECK_ARG(N >= 0 && C >= 1 && S >= 1, "kgdet_nchw_to_nhwc: bad sizes");
  if (N == 0) return KGDET_OK;
  KG_CHECK_ARG(src && dst, "kgdet_nchw_to_nhwc: NULL pointer");
  return kgdet::launch_transpose(src, dst, N, C, S, src_dtype, dst_dtype, (cudaStream_t)stream);
}
