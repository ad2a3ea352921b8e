// Host side of the fused tensor-core deformable-convolution forward: weight packing into the layout
// tcgen05.mma reads, and parameter set-up for the kernel in dcn_umma_stream.cu.
//
// K order of the contraction is (channel block, tap, channel-in-block) so that one channel block's taps
// hit the same L1 lines back to back; the packed weights use the same order.
#include "dcn_umma.cuh"

namespace kgdet {

static int mode_of(int precision) {
  return precision == KGDET_PREC_BF16 ? MODE_BF16 : (precision == KGDET_PREC_TF32X3 ? MODE_TF32X3 : MODE_TF32);
}
static int bk_of(int precision) { return precision == KGDET_PREC_BF16 ? 64 : 32; }

bool umma_supported(const DcnGeom& g, int precision) {
  if (precision != KGDET_PREC_BF16 && precision != KGDET_PREC_TF32X3 && precision != KGDET_PREC_TF32)
    return false;
  return g.groups == 1 && g.dgroups == 1 && g.C % 64 == 0 && g.Cout % 64 == 0 && g.Cout <= 256 &&
         g.Cout >= 64;
}

// ---- weight packing ------------------------------------------------------------------------
// Packed k-block kb = cb * K + tap holds, per B tile, Cout rows of 128 bytes in the
// swizzled K-major layout: byte (o, j) -> (o/8)*1024 + (o%8)*128 + (((j*E)/16) ^ (o%8))*16 + (j*E)%16.
// TF32 modes store a hi tile (tf32-rounded) followed by a lo tile (w - hi).
size_t umma_packed_weight_bytes(const DcnGeom& g, int precision) {
  const int bk = bk_of(precision);
  const size_t nkb = (size_t)(g.C / bk) * g.K;
  const int tiles = precision == KGDET_PREC_BF16 ? 1 : 2;
  return nkb * tiles * g.Cout * 128;
}

__global__ void umma_pack_bf16_kernel(const float* __restrict__ w, unsigned char* __restrict__ p, int C,
                                      int Cout, int K) {
  // one thread per 16-byte chunk (8 channels)
  const int chunks_per_blk = Cout * 8;
  const int nkb = (C / 64) * K;
  const int total = nkb * chunks_per_blk;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += blockDim.x * gridDim.x) {
    const int kb = i / chunks_per_blk, r = i - kb * chunks_per_blk;
    const int o = r >> 3, chunk = r & 7;
    const int cb = kb / K, tap = kb - cb * K;
    __align__(16) __nv_bfloat16 v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = cb * 64 + chunk * 8 + e;
      v[e] = __float2bfloat16(w[((size_t)o * C + c) * K + tap]);
    }
    const size_t dst = (size_t)kb * Cout * 128 + (size_t)(o >> 3) * 1024 + (o & 7) * 128 +
                       ((chunk ^ (o & 7)) << 4);
    *reinterpret_cast<uint4*>(p + dst) = *reinterpret_cast<const uint4*>(v);
  }
}

// Same packing through shared memory: one CTA takes 2 output channels x one 64-channel block, reads their
// 64 * K weights each as ONE contiguous run (the layout is [Cout][C][K]) and writes, per tap, the two 128-byte rows.
// The element-per-thread kernel above reads with a stride of K floats; training repacks every step.
__global__ void __launch_bounds__(256) umma_pack_bf16_tiled_kernel(const float* __restrict__ w, unsigned char* __restrict__ p,
                                                                   int C, int Cout, int K) {
  extern __shared__ float wsm[];                       // [2][64 * K]
  const int o0 = blockIdx.x * 2, cb = blockIdx.y;
  const int run = 64 * K;
  for (int i = threadIdx.x; i < 2 * run; i += blockDim.x) {
    const int oo = i / run, j = i - oo * run;
    wsm[i] = (o0 + oo < Cout) ? w[((size_t)(o0 + oo) * C + cb * 64) * K + j] : 0.f;
  }
  __syncthreads();
  // one thread per 16-byte chunk: (tap, oo, chunk)
  for (int i = threadIdx.x; i < K * 16; i += blockDim.x) {
    const int tap = i >> 4, oo = (i >> 3) & 1, chunk = i & 7;
    const int o = o0 + oo;
    if (o >= Cout) continue;
    __align__(16) __nv_bfloat16 v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = __float2bfloat16(wsm[oo * run + (chunk * 8 + e) * K + tap]);
    const size_t dst = (size_t)(cb * K + tap) * Cout * 128 + (size_t)(o >> 3) * 1024 + (o & 7) * 128 +
                       ((chunk ^ (o & 7)) << 4);
    *reinterpret_cast<uint4*>(p + dst) = *reinterpret_cast<const uint4*>(v);
  }
}

__global__ void umma_pack_tf32_kernel(const float* __restrict__ w, unsigned char* __restrict__ p, int C,
                                      int Cout, int K) {
  // one thread per 16-byte chunk (4 channels); writes the hi and the lo tile
  const int chunks_per_blk = Cout * 8;
  const int nkb = (C / 32) * K;
  const int total = nkb * chunks_per_blk;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += blockDim.x * gridDim.x) {
    const int kb = i / chunks_per_blk, r = i - kb * chunks_per_blk;
    const int o = r >> 3, chunk = r & 7;
    const int cb = kb / K, tap = kb - cb * K;
    float hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = cb * 32 + chunk * 4 + e;
      const float x = w[((size_t)o * C + c) * K + tap];
      hi[e] = tf32_rna(x);
      lo[e] = x - hi[e];
    }
    const size_t dst = (size_t)kb * 2 * Cout * 128 + (size_t)(o >> 3) * 1024 + (o & 7) * 128 +
                       ((chunk ^ (o & 7)) << 4);
    *reinterpret_cast<float4*>(p + dst) = make_float4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<float4*>(p + dst + (size_t)Cout * 128) = make_float4(lo[0], lo[1], lo[2], lo[3]);
  }
}

int umma_pack_weight(const DcnGeom& g, const float* weight, void* packed, int precision,
                     cudaStream_t stream) {
  const int bk = bk_of(precision);
  const int total = (g.C / bk) * g.K * g.Cout * 8;
  const int blocks = ceil_div(total, 256);
  if (precision == KGDET_PREC_BF16 && (size_t)2 * 64 * g.K * sizeof(float) <= 48 * 1024)
    umma_pack_bf16_tiled_kernel<<<dim3((unsigned)ceil_div(g.Cout, 2), (unsigned)(g.C / 64)), 256,
                                  (size_t)2 * 64 * g.K * sizeof(float), stream>>>(weight, (unsigned char*)packed, g.C,
                                                                                 g.Cout, g.K);
  else if (precision == KGDET_PREC_BF16)
    umma_pack_bf16_kernel<<<blocks, 256, 0, stream>>>(weight, (unsigned char*)packed, g.C, g.Cout, g.K);
  else
    umma_pack_tf32_kernel<<<blocks, 256, 0, stream>>>(weight, (unsigned char*)packed, g.C, g.Cout, g.K);
  KG_LAUNCH_CHECK("umma_pack_kernel");
  return KGDET_OK;
}

// ---- split-K over CTAs for small maps ------------------------------------------------------------------
// The fused kernel is one CTA per 128 positions, each a serial pipeline of nkb k-blocks at ~1 155 clk: a map of
// M = 2 100 positions (KGDet training at batch 2, or FPN P6 / P7) keeps 17 of 148 SMs busy for the whole call.
// When the tiles cover less than half of the SMs the k-blocks are split over gridDim.y CTAs per tile; the partial
// accumulators go through an fp32 buffer and are combined by umma_split_reduce_kernel (deterministic order).
// TF32X3 -- accumulator promotion.  The tensor core adds every MMA into the fp32 TMEM accumulator with
// truncation, so the error of one accumulator grows linearly with the number of MMAs it absorbs (measured at
// C = 256: 4.4e-5 after 200 k-blocks, 8.2e-5 after 392, i.e. ~2e-7 per 32-channel k-block).  The fp32-grade mode
// therefore never lets one accumulator absorb more than kPromoteChunk k-blocks: the k-blocks are split into chunks
// (gridDim.y), every chunk accumulates in its own TMEM tile from zero, and the chunk results are promoted to fp32
// partial tiles that umma_split_reduce_kernel sums with round-to-nearest adds in a fixed order -- the same
// machinery that splits small maps over the SMs.  16 k-blocks per chunk keep the truncation share below 4e-6.
static constexpr int kPromoteChunk = 16;
int umma_splits(const DcnGeom& g, int precision) {
  if (!umma_supported(g, precision)) return 1;
  const int tiles = ceil_div(g.M, BM);
  const int nkb = (g.C / bk_of(precision)) * g.K;
  int s = num_sms() / tiles;
  if (s > nkb / 6) s = nkb / 6;                 // at least 6 k-blocks per CTA
  if (s > 16) s = 16;
  if (s < 4) s = 1;                             // measured: 2 splits (FPN P5 at batch 8, 66 tiles) lose to the reduction pass
  if (const char* e = getenv("KGDET_UMMA_SPLITS")) s = atoi(e);
  if (precision == KGDET_PREC_TF32X3) {
    int chunk = kPromoteChunk;
    if (const char* e = getenv("KGDET_TF32X3_CHUNK")) chunk = atoi(e) > 0 ? atoi(e) : nkb;
    const int need = ceil_div(nkb, chunk);
    if (s < need) s = need;
  }
  return s < 1 ? 1 : s;
}
size_t umma_split_ws_bytes(const DcnGeom& g, int precision) {
  const int s = umma_splits(g, precision);
  return s > 1 ? (size_t)s * ceil_div(g.M, BM) * BM * g.Cout * sizeof(float) : 0;
}

// out[n, coff + o, p] = act(bias[o] + sum_s partial[s][m][o]),  m = n * HoWo + p;  32 x 32 tiles through smem
template <typename Tout>
__global__ void umma_split_reduce_kernel(const float* __restrict__ partial, int splits, long long split_stride,
                                         const float* __restrict__ bias, Tout* __restrict__ out, int M, int Cout,
                                         int HoWo, int coff, int ctot, int relu) {
  __shared__ float tile[32][33];
  const int m0 = blockIdx.x * 32, o0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;      // 32 x 8
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int m = m0 + ty + k, o = o0 + tx;
    float v = 0.f;
    if (m < M && o < Cout) {
      const float* p = partial + (size_t)m * Cout + o;
      for (int s = 0; s < splits; ++s) v += p[(size_t)s * split_stride];
      if (bias) v += __ldg(bias + o);
      if (relu) v = fmaxf(v, 0.f);
    }
    tile[ty + k][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int o = o0 + ty + k, m = m0 + tx;
    if (m < M && o < Cout) {
      const int n = m / HoWo, pos = m - n * HoWo;
      st_out<Tout>(out + ((size_t)n * ctot + coff + o) * HoWo + pos, tile[tx][ty + k]);
    }
  }
}

thread_local long long* g_timeline = nullptr;
thread_local long long g_timeline_entries = 0;

int umma_forward(const DcnGeom& g, const void* in_blocked, size_t plane_bytes, const SampleRec16* plan,
                 const void* packed_w, const float* bias, const OutSpec& o, int precision, cudaStream_t stream,
                 void* split_ws) {
  if (!umma_supported(g, precision)) {
    set_error("dcn umma: shape/precision not supported by the tensor-core path");
    return KGDET_ERR_UNSUPPORTED;
  }
  const int mode = mode_of(precision);
  const int bk = bk_of(precision);
  UmmaParams p;
  p.in = in_blocked; p.plane_bytes = plane_bytes; p.plan = plan; p.wp = (const unsigned char*)packed_w; p.bias = bias; p.out = o.out;
  p.out_coff = o.coff; p.out_ctot = o.ctot; p.relu = o.relu; p.out_nhwc = o.nhwc;
  p.M = g.M; p.C = g.C; p.W = g.W; p.Cout = g.Cout; p.K = g.K; p.HoWo = g.Ho * g.Wo;
  p.rows_padded = (int)plan_rows(g);
  p.nkb = (g.C / bk) * g.K;
  // CTA pairs (2-SM MMA, each CTA holds half of the weight slab): implemented and parity-tested, but not
  // faster on B200 (K = 49 call: 152 us vs 144 us) because the producers, not the weight traffic, bound the
  // kernel -- opt-in with KGDET_UMMA_PAIR=1.
  bool pair = false;
  if (const char* e = getenv("KGDET_UMMA_PAIR")) pair = atoi(e) != 0;
  p.idesc = make_idesc(mode == MODE_BF16 ? 1u : 2u, pair ? 2 * BM : BM, (uint32_t)g.Cout);
  p.tmem_cols = 0;   // set by the kernel launcher
  p.timeline = nullptr;
  if (g_timeline && g_timeline_entries >= (long long)ceil_div(g.M, BM) * (12 * p.nkb + 8)) p.timeline = g_timeline;
  g_timeline = nullptr;
  p.partial = nullptr; p.kb_per_split = p.nkb; p.m_pad = ceil_div(g.M, BM) * BM;
  const int splits = (split_ws && !pair && !o.nhwc) ? umma_splits(g, precision) : 1;
  if (splits <= 1) return umma_stream_forward(g, p, mode, pair, o.dtype, stream);
  p.partial = (float*)split_ws;
  p.kb_per_split = ceil_div(p.nkb, splits);
  const int used = ceil_div(p.nkb, p.kb_per_split);             // every split owns at least one k-block
  int rc = umma_stream_forward(g, p, mode, false, o.dtype, stream, used);
  if (rc != KGDET_OK) return rc;
  const dim3 grid((unsigned)ceil_div(g.M, 32), (unsigned)ceil_div(g.Cout, 32)), block(32, 8);
  const long long stride = (long long)p.m_pad * g.Cout;
  if (o.dtype == KGDET_F32)
    umma_split_reduce_kernel<float><<<grid, block, 0, stream>>>(p.partial, used, stride, bias, (float*)o.out, g.M, g.Cout,
                                                               g.Ho * g.Wo, o.coff, o.ctot, o.relu);
  else
    umma_split_reduce_kernel<__nv_bfloat16><<<grid, block, 0, stream>>>(p.partial, used, stride, bias,
                                                                       (__nv_bfloat16*)o.out, g.M, g.Cout, g.Ho * g.Wo,
                                                                       o.coff, o.ctot, o.relu);
  KG_LAUNCH_CHECK("umma_split_reduce_kernel");
  return KGDET_OK;
}

}  // namespace kgdet

extern "C" void kgdet_dcn_set_timeline(void* device_buffer, long long entries) {
  kgdet::g_timeline = (long long*)device_buffer;
  kgdet::g_timeline_entries = entries;
}
