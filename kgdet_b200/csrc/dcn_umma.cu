// Fused deformable-convolution forward on the 5th-generation tensor cores (sm_100a).
//
//   out[m, o] = sum_{tap, c} S(m, tap, c) * W[o, c, tap]       m = (n, y, x) output position
//
// The reference materialises S as the `columns` tensor in HBM (deformable_im2col,
// deform_conv_cuda_kernel.cu:189-242: C*K x N*H*W fp32, 843 MB for one 7x7 KGDet call) and
// hands it to a cuBLAS SGEMM (deform_conv_cuda.cpp:230-233).  Here the column tile only ever
// exists in shared memory:
//
//   warps 0-7  producers : bilinear-gather a 128-position x 128-byte slab of S from the NHWC
//                          input (16-byte vector loads, 8 lanes cover one pixel's slab) using
//                          the precomputed sample plan, and store it straight into the
//                          128B-swizzled K-major layout tcgen05.mma reads;
//                          afterwards the same warps run the epilogue (TMEM -> NCHW output).
//   warp 8     loader    : one thread streams the matching pre-swizzled weight slab with
//                          cp.async.bulk (UBLKCP) -- the pack step laid it out so that a
//                          linear copy lands in UMMA layout, no tensor map needed.
//   warp 9     MMA       : one thread issues tcgen05.mma (UTCHMMA), accumulators live in TMEM;
//                          tcgen05.commit releases pipeline stages / signals the epilogue.
//
// Modes: BF16   kind::f16, bf16 operands                     (1e-3 grade)
//        TF32X3 kind::tf32, A = Ahi + Alo, B = Bhi + Blo,    (fp32 grade: the dropped term is
//               3 MMAs per k-step: Ahi.Bhi + Alo.Bhi + Ahi.Blo   Alo.Blo ~ 2^-22)
//        TF32   kind::tf32 single pass
// K order is (channel block, tap, channel-in-block) so that one channel block's taps hit the
// same L1 lines back to back; umma_pack_weight uses the same order.
#include "dcn.cuh"

namespace kgdet {

static constexpr int BM = 128;                 // positions per CTA tile (UMMA M)
static constexpr int PRODUCER_WARPS = 8;
static constexpr int UMMA_THREADS = (PRODUCER_WARPS + 2) * 32;
static constexpr int A_TILE_BYTES = BM * 128;  // 128 rows x 128 B

enum { MODE_BF16 = 0, MODE_TF32X3 = 1, MODE_TF32 = 2 };

template <int MODE> struct ModeTraits;
template <> struct ModeTraits<MODE_BF16>   { static constexpr int BK = 64, A_TILES = 1, B_TILES = 1, ELEM = 2; };
template <> struct ModeTraits<MODE_TF32X3> { static constexpr int BK = 32, A_TILES = 2, B_TILES = 2, ELEM = 4; };
template <> struct ModeTraits<MODE_TF32>   { static constexpr int BK = 32, A_TILES = 1, B_TILES = 1, ELEM = 4; };

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

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
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
  if (precision == KGDET_PREC_BF16)
    umma_pack_bf16_kernel<<<blocks, 256, 0, stream>>>(weight, (unsigned char*)packed, g.C, g.Cout, g.K);
  else
    umma_pack_tf32_kernel<<<blocks, 256, 0, stream>>>(weight, (unsigned char*)packed, g.C, g.Cout, g.K);
  KG_LAUNCH_CHECK("umma_pack_kernel");
  return KGDET_OK;
}

// ---- the fused kernel ----------------------------------------------------------------------
struct UmmaParams {
  const void* in;            // NHWC, bf16 (MODE_BF16) or fp32 (TF32 modes)
  const SampleRec* plan;     // [rows_padded][K]
  const unsigned char* wp;   // packed weights
  const float* bias;         // [Cout] or NULL
  void* out;                 // NCHW
  int M, C, Cout, K, HoWo;
  int nkb;                   // (C / BK) * K
  uint32_t idesc;
  uint32_t tmem_cols;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void fma_bf16x2(float w, uint32_t packed, float& a0, float& a1) {
  a0 = fmaf(w, __uint_as_float(packed << 16), a0);
  a1 = fmaf(w, __uint_as_float(packed & 0xffff0000u), a1);
}

template <typename T> __device__ __forceinline__ void st_out(T* p, float v);
template <> __device__ __forceinline__ void st_out<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_out<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  *p = __float2bfloat16(v);
}

template <int MODE, int NS, typename Tout>
__global__ void __launch_bounds__(UMMA_THREADS, 1) dcn_umma_fwd_kernel(const UmmaParams prm) {
  using MT = ModeTraits<MODE>;
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  // 1024-byte alignment of every tile is what the 128B swizzle pattern is anchored to
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  const int BN = prm.Cout;
  const int b_tile_bytes = BN * 128;
  const int stage_bytes = MT::A_TILES * A_TILE_BYTES + MT::B_TILES * b_tile_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)NS * stage_bytes);
  uint64_t* empty_bar = full_bar + NS;
  uint64_t* tmem_full_bar = empty_bar + NS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * BM;

  if (warp == PRODUCER_WARPS + 1) {
    if (lane == 0) {
      for (int s = 0; s < NS; ++s) {
        mbar_init(&full_bar[s], PRODUCER_WARPS + 1);
        mbar_init(&empty_bar[s], 1);
      }
      mbar_init(tmem_full_bar, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, prm.tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < PRODUCER_WARPS) {
    // ===================== producers: gather S tiles into swizzled smem =====================
    const int chunk = tid & 7;        // 16-byte chunk of the 128-byte row
    const int rbase = tid >> 3;       // 0..31
    const int C = prm.C, K = prm.K;
    for (int kb = 0; kb < prm.nkb; ++kb) {
      const int s = kb % NS, it = kb / NS;
      const int cb = kb / K, tap = kb - cb * K;
      mbar_wait(&empty_bar[s], (it & 1) ^ 1);
      unsigned char* a_tile = smem + (size_t)s * stage_bytes;
      SampleRec rec[4];
#pragma unroll
      for (int ps = 0; ps < 4; ++ps) {
        const SampleRec* rp = prm.plan + (size_t)(m0 + rbase + ps * 32) * K + tap;
        const int4 a = __ldg(reinterpret_cast<const int4*>(rp));
        const float4 b = __ldg(reinterpret_cast<const float4*>(rp) + 1);
        rec[ps].pix[0] = a.x; rec[ps].pix[1] = a.y; rec[ps].pix[2] = a.z; rec[ps].pix[3] = a.w;
        rec[ps].w[0] = b.x; rec[ps].w[1] = b.y; rec[ps].w[2] = b.z; rec[ps].w[3] = b.w;
      }
      if constexpr (MODE == MODE_BF16) {
        const __nv_bfloat16* in = reinterpret_cast<const __nv_bfloat16*>(prm.in) + cb * 64 + chunk * 8;
        uint4 v[4][4];
#pragma unroll
        for (int ps = 0; ps < 4; ++ps)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            v[ps][i] = make_uint4(0u, 0u, 0u, 0u);
            if (rec[ps].w[i] != 0.f)
              v[ps][i] = __ldg(reinterpret_cast<const uint4*>(in + (size_t)rec[ps].pix[i] * C));
          }
#pragma unroll
        for (int ps = 0; ps < 4; ++ps) {
          float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float w = rec[ps].w[i];
            fma_bf16x2(w, v[ps][i].x, acc[0], acc[1]);
            fma_bf16x2(w, v[ps][i].y, acc[2], acc[3]);
            fma_bf16x2(w, v[ps][i].z, acc[4], acc[5]);
            fma_bf16x2(w, v[ps][i].w, acc[6], acc[7]);
          }
          const int r = rbase + ps * 32;
          uint4 o;
          o.x = pack_bf16x2(acc[0], acc[1]);
          o.y = pack_bf16x2(acc[2], acc[3]);
          o.z = pack_bf16x2(acc[4], acc[5]);
          o.w = pack_bf16x2(acc[6], acc[7]);
          *reinterpret_cast<uint4*>(a_tile + r * 128 + ((chunk ^ (r & 7)) << 4)) = o;
        }
      } else {
        const float* in = reinterpret_cast<const float*>(prm.in) + cb * 32 + chunk * 4;
        float4 v[4][4];
#pragma unroll
        for (int ps = 0; ps < 4; ++ps)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            v[ps][i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (rec[ps].w[i] != 0.f)
              v[ps][i] = __ldg(reinterpret_cast<const float4*>(in + (size_t)rec[ps].pix[i] * C));
          }
#pragma unroll
        for (int ps = 0; ps < 4; ++ps) {
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float w = rec[ps].w[i];
            a0 = fmaf(w, v[ps][i].x, a0);
            a1 = fmaf(w, v[ps][i].y, a1);
            a2 = fmaf(w, v[ps][i].z, a2);
            a3 = fmaf(w, v[ps][i].w, a3);
          }
          const int r = rbase + ps * 32;
          const int off = r * 128 + ((chunk ^ (r & 7)) << 4);
          if constexpr (MODE == MODE_TF32X3) {
            const float h0 = tf32_rna(a0), h1 = tf32_rna(a1), h2 = tf32_rna(a2), h3 = tf32_rna(a3);
            *reinterpret_cast<float4*>(a_tile + off) = make_float4(h0, h1, h2, h3);
            *reinterpret_cast<float4*>(a_tile + A_TILE_BYTES + off) =
                make_float4(a0 - h0, a1 - h1, a2 - h2, a3 - h3);
          } else {
            *reinterpret_cast<float4*>(a_tile + off) = make_float4(a0, a1, a2, a3);
          }
        }
      }
      fence_proxy_async_smem();   // my generic-proxy stores -> visible to tcgen05.mma
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[s]);
    }

    // ===================== epilogue: TMEM -> registers -> NCHW global =====================
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int q = warp & 3, half = warp >> 2;      // TMEM lane quarter this warp may touch
    const int row = q * 32 + lane;
    const int m = m0 + row;
    const bool row_ok = m < prm.M;
    const int n = row_ok ? m / prm.HoWo : 0;
    const int pos = row_ok ? m - n * prm.HoWo : 0;
    Tout* obase = reinterpret_cast<Tout*>(prm.out) + (size_t)n * prm.Cout * prm.HoWo + pos;
    const int half_cols = BN >> 1;
    for (int c0 = 0; c0 < half_cols; c0 += 32) {
      const int col = half * half_cols + c0;
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)col, v);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = __uint_as_float(v[j]);
          if (prm.bias) x += __ldg(prm.bias + col + j);
          st_out<Tout>(obase + (size_t)(col + j) * prm.HoWo, x);   // lanes = consecutive positions
        }
      }
    }
  } else if (warp == PRODUCER_WARPS) {
    // ===================== weight loader (one thread, bulk async copies) =====================
    if (lane == 0) {
      const uint32_t bytes = (uint32_t)(MT::B_TILES * b_tile_bytes);
      for (int kb = 0; kb < prm.nkb; ++kb) {
        const int s = kb % NS, it = kb / NS;
        mbar_wait(&empty_bar[s], (it & 1) ^ 1);
        unsigned char* b_tile = smem + (size_t)s * stage_bytes + MT::A_TILES * A_TILE_BYTES;
        mbar_arrive_expect_tx(&full_bar[s], bytes);
        // the packed layout always carries hi+lo for tf32; single-pass TF32 copies only hi
        const size_t src_stride = (size_t)(MODE == MODE_BF16 ? 1 : 2) * b_tile_bytes;
        bulk_g2s(b_tile, prm.wp + (size_t)kb * src_stride, bytes, &full_bar[s]);
      }
    }
    __syncwarp();
  } else {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      for (int kb = 0; kb < prm.nkb; ++kb) {
        const int s = kb % NS, it = kb / NS;
        mbar_wait(&full_bar[s], it & 1);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + (size_t)s * stage_bytes);
        const uint32_t b_addr = a_addr + MT::A_TILES * A_TILE_BYTES;
        const uint64_t adesc = make_sw128_kmajor_desc(a_addr);
        const uint64_t bdesc = make_sw128_kmajor_desc(b_addr);
#pragma unroll
        for (int k = 0; k < 4; ++k) {           // 4 x 32 bytes of K per 128-byte row
          const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
          if constexpr (MODE == MODE_BF16) {
            umma_f16(tmem_base, adesc + 2 * k, bdesc + 2 * k, prm.idesc, acc);
          } else if constexpr (MODE == MODE_TF32) {
            umma_tf32(tmem_base, adesc + 2 * k, bdesc + 2 * k, prm.idesc, acc);
          } else {
            const uint64_t adesc_lo = make_sw128_kmajor_desc(a_addr + A_TILE_BYTES);
            const uint64_t bdesc_lo = make_sw128_kmajor_desc(b_addr + b_tile_bytes);
            umma_tf32(tmem_base, adesc_lo + 2 * k, bdesc + 2 * k, prm.idesc, acc);     // Alo.Bhi
            umma_tf32(tmem_base, adesc + 2 * k, bdesc_lo + 2 * k, prm.idesc, 1u);      // Ahi.Blo
            umma_tf32(tmem_base, adesc + 2 * k, bdesc + 2 * k, prm.idesc, 1u);         // Ahi.Bhi
          }
        }
        tc_commit(&empty_bar[s]);               // frees the stage when these MMAs retire
      }
      tc_commit(tmem_full_bar);                 // accumulator complete -> epilogue
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == PRODUCER_WARPS + 1) tmem_dealloc(tmem_base, prm.tmem_cols);
}

static size_t umma_smem_bytes(int mode, int ns, int Cout) {
  const int a_tiles = (mode == MODE_TF32X3) ? 2 : 1, b_tiles = a_tiles;
  const size_t stage = (size_t)a_tiles * A_TILE_BYTES + (size_t)b_tiles * Cout * 128;
  return 1024 /* alignment slack */ + ns * stage + (2 * ns + 1) * 8 + 16;
}

template <int MODE, int NS, typename Tout>
static int launch_umma(const UmmaParams& p, int grid, cudaStream_t stream) {
  const size_t smem = umma_smem_bytes(MODE, NS, p.Cout);
  KG_CUDA(cudaFuncSetAttribute(dcn_umma_fwd_kernel<MODE, NS, Tout>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dcn_umma_fwd_kernel<MODE, NS, Tout><<<grid, UMMA_THREADS, smem, stream>>>(p);
  KG_LAUNCH_CHECK("dcn_umma_fwd_kernel");
  return KGDET_OK;
}

template <int MODE, typename Tout>
static int dispatch_stages(const UmmaParams& p, int grid, int ns, cudaStream_t stream) {
  switch (ns) {
    case 2: return launch_umma<MODE, 2, Tout>(p, grid, stream);
    case 3: return launch_umma<MODE, 3, Tout>(p, grid, stream);
    case 4: return launch_umma<MODE, 4, Tout>(p, grid, stream);
    default: set_error("dcn umma: unsupported stage count %d", ns); return KGDET_ERR_INVALID_ARG;
  }
}

int umma_forward(const DcnGeom& g, const void* in_nhwc, const SampleRec* plan, const void* packed_w,
                 const float* bias, void* out_nchw, int out_dtype, int precision, cudaStream_t stream) {
  if (!umma_supported(g, precision)) {
    set_error("dcn umma: shape/precision not supported by the tensor-core path");
    return KGDET_ERR_UNSUPPORTED;
  }
  const int mode = mode_of(precision);
  const int bk = bk_of(precision);
  UmmaParams p;
  p.in = in_nhwc; p.plan = plan; p.wp = (const unsigned char*)packed_w; p.bias = bias; p.out = out_nchw;
  p.M = g.M; p.C = g.C; p.Cout = g.Cout; p.K = g.K; p.HoWo = g.Ho * g.Wo;
  p.nkb = (g.C / bk) * g.K;
  p.idesc = make_idesc(mode == MODE_BF16 ? 1u : 2u, BM, (uint32_t)g.Cout);
  p.tmem_cols = g.Cout <= 64 ? 64 : (g.Cout <= 128 ? 128 : 256);
  const int grid = ceil_div(g.M, BM);
  // pipeline depth: as deep as 227 KB allows, capped so that some L1 is left for the gather
  int ns = (mode == MODE_TF32X3) ? 2 : 3;
  if (const char* e = getenv("KGDET_UMMA_STAGES")) {
    int v = atoi(e);
    if (v >= 2 && v <= 4) ns = v;
  }
  while (ns > 2 && umma_smem_bytes(mode, ns, g.Cout) > 227 * 1024) --ns;
  if (umma_smem_bytes(mode, ns, g.Cout) > 227 * 1024) {
    set_error("dcn umma: tile does not fit shared memory");
    return KGDET_ERR_UNSUPPORTED;
  }
  const bool f32 = out_dtype == KGDET_F32;
  switch (mode) {
    case MODE_BF16:
      return f32 ? dispatch_stages<MODE_BF16, float>(p, grid, ns, stream)
                 : dispatch_stages<MODE_BF16, __nv_bfloat16>(p, grid, ns, stream);
    case MODE_TF32X3:
      return f32 ? dispatch_stages<MODE_TF32X3, float>(p, grid, ns, stream)
                 : dispatch_stages<MODE_TF32X3, __nv_bfloat16>(p, grid, ns, stream);
    default:
      return f32 ? dispatch_stages<MODE_TF32, float>(p, grid, ns, stream)
                 : dispatch_stages<MODE_TF32, __nv_bfloat16>(p, grid, ns, stream);
  }
}

}  // namespace kgdet
