// Fused deformable-convolution forward on the 5th-generation tensor cores (sm_100a).
//
//   out[m, o] = sum_{tap, c} S(m, tap, c) * W[o, c, tap]       m = (n, y, x) output position
//
// The reference materialises S as the `columns` tensor in HBM (deformable_im2col,
// deform_conv_cuda_kernel.cu:189-242: C*K x N*H*W fp32, 843 MB for one 7x7 KGDet call) and
// hands it to a cuBLAS SGEMM (deform_conv_cuda.cpp:230-233).  Here the column tile only ever
// exists in shared memory.  One CTA = 128 output positions x all Cout, 16 warps in two groups:
//
//   group g (8 warps) owns k-blocks g, g+2, g+4, ... and its own TMEM accumulator, so that one
//   group's gather loads are in flight while the other group combines/stores its tile.
//   Per k-block (64 bf16 / 32 tf32 channels of one tap):
//     all 256 threads : decode 4 compact plan records, issue 16 predicated 16-byte gathers from
//                       the NHWC input (8 lanes cover one pixel's 128-byte slab), prefetch the next
//                       records, wait for the pipeline stage, bilinear-combine and store the
//                       128-position x 128-byte A tile straight into the 128B-swizzled K-major
//                       layout tcgen05.mma reads (fence.proxy.async + mbarrier arrive);
//     group leader    : one thread streams the matching pre-swizzled weight slab with
//                       cp.async.bulk (UBLKCP; a linear copy lands in UMMA layout, no tensor map),
//                       then waits for the stage to be full and issues the tcgen05.mma's
//                       (UTCHMMA) into the group's accumulator; tcgen05.commit frees the stage.
//   Epilogue (all warps): out = acc[0] + acc[1] (tcgen05.ld), bias, NCHW store coalesced over
//   positions.
//
// Modes: BF16   kind::f16, bf16 operands                     (1e-3 grade)
//        TF32X3 kind::tf32, A = Ahi + Alo, B = Bhi + Blo,    (the dropped term Alo.Blo is ~2^-22)
//               3 MMAs per k-step: Alo.Bhi + Ahi.Blo + Ahi.Bhi
//        TF32   kind::tf32 single pass
// K order is (channel block, tap, channel-in-block) so that one channel block's taps hit the
// same L1 lines back to back; umma_pack_weight uses the same order.
#include "dcn.cuh"

namespace kgdet {

static constexpr int BM = 128;                 // positions per CTA tile (UMMA M)
static constexpr int GROUP_WARPS = 8;           // one group covers 128 rows x 8 chunks in 4 passes
static constexpr int NUM_GROUPS = 2;
static constexpr int UMMA_THREADS = NUM_GROUPS * GROUP_WARPS * 32;   // 512 -> 128 registers/thread, no spills
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
  const SampleRec16* plan;   // [rows_padded][K]
  const unsigned char* wp;   // packed weights
  const float* bias;         // [Cout] or NULL
  void* out;                 // NCHW
  int M, C, W, Cout, K, HoWo;
  int out_coff, out_ctot, relu;   // channel slice of the output tensor, fused ReLU
  int nkb;                   // (C / BK) * K
  uint32_t idesc;
  uint32_t tmem_cols;
};



template <typename T> __device__ __forceinline__ void st_out(T* p, float v);
template <> __device__ __forceinline__ void st_out<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_out<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  *p = __float2bfloat16(v);
}

__device__ __forceinline__ uint32_t bf162_bcast(float w) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %1;" : "=r"(r) : "f"(w));
  return r;
}
__device__ __forceinline__ uint32_t bf162_mul(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t bf162_fma(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void fma_bf16x2_f32(float w, uint32_t packed, float& a0, float& a1) {
  a0 = fmaf(w, __uint_as_float(packed << 16), a0);
  a1 = fmaf(w, __uint_as_float(packed & 0xffff0000u), a1);
}

// bilinear weights of the four corners from a compact plan record (see SampleRec16)
__device__ __forceinline__ void decode_rec(const float4& r, float (&w)[4]) {
  const uint32_t lhb = __float_as_uint(r.y), lwb = __float_as_uint(r.z);
  const float lh = __uint_as_float(lhb & ~3u), lw = __uint_as_float(lwb & ~3u);
  const float fh0 = (lhb & 1u) ? (1.f - lh) : 0.f, fh1 = (lhb & 2u) ? lh : 0.f;
  const float fw0 = (lwb & 1u) ? (1.f - lw) * r.w : 0.f, fw1 = (lwb & 2u) ? lw * r.w : 0.f;
  w[0] = fh0 * fw0; w[1] = fh0 * fw1; w[2] = fh1 * fw0; w[3] = fh1 * fw1;
}

// 16 warps in two groups of 8; group g owns k-blocks g, g+2, ... and TMEM accumulator g.  There are
// no dedicated control warps (a 17th warp would cut the register budget from 128 to 96 per thread):
// per k-block one warp of the group -- rotating -- additionally acts as "leader": it fetches the
// weight slab (cp.async.bulk, one k-block ahead), waits for the stage to be full and issues the MMAs.
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
  const int active_groups = prm.nkb < NUM_GROUPS ? prm.nkb : NUM_GROUPS;
  // every thread that issued MMAs makes one final tcgen05.commit on tmem_full_bar
  int num_issuers = 0;
  for (int g = 0; g < NUM_GROUPS; ++g) {
    const int n_g = (prm.nkb - g + NUM_GROUPS - 1) / NUM_GROUPS;
    num_issuers += n_g < GROUP_WARPS ? (n_g < 0 ? 0 : n_g) : GROUP_WARPS;
  }

  if (warp == 0) {
    if (lane == 0) {
      for (int s = 0; s < NS; ++s) {
        mbar_init(&full_bar[s], GROUP_WARPS + 1);   // 8 producer warps + the leader's expect_tx
        mbar_init(&empty_bar[s], 1);                // one tcgen05.commit
      }
      mbar_init(tmem_full_bar, num_issuers);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, prm.tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  {
    const int group = warp / GROUP_WARPS;
    const int wg = warp % GROUP_WARPS;
    const int t = tid - group * (GROUP_WARPS * 32);
    const int chunk = t & 7;          // 16-byte chunk of the 128-byte row
    const int rbase = t >> 3;         // 0..31; this thread owns rows rbase + 32*ps
    const int K = prm.K;
    const size_t rowb = (size_t)prm.C * MT::ELEM;         // bytes per pixel
    const size_t wrow = (size_t)prm.W * rowb;             // bytes per image row
    const unsigned char* in_base = reinterpret_cast<const unsigned char*>(prm.in) + chunk * 16;
    const uint4* plan_t = reinterpret_cast<const uint4*>(prm.plan) + (size_t)(m0 + rbase) * K;
    const uint32_t acc_tmem = tmem_base + (uint32_t)(group * BN);    // this group's accumulator
    const uint32_t b_bytes = (uint32_t)(MT::B_TILES * b_tile_bytes);
    // the packed layout always carries hi+lo for tf32; single-pass TF32 copies only hi
    const size_t b_src_stride = (size_t)(MODE == MODE_BF16 ? 1 : 2) * b_tile_bytes;
    // this thread's byte offset inside an A tile (row rbase, swizzled 16-byte chunk); rows of later
    // passes are 32 rows = 4096 bytes further and keep the same (row & 7)
    const int a_off = rbase * 128 + ((chunk ^ (rbase & 7)) << 4);
    bool issued_any = false;

    // (tap, channel block) of k-block kb = cb * K + tap, advanced incrementally (no divisions)
    int tap = group % K, cb = group / K;
    uint4 rec[4];
    if (group < prm.nkb) {
#pragma unroll
      for (int ps = 0; ps < 4; ++ps) rec[ps] = __ldg(plan_t + (size_t)ps * 32 * K + tap);
    }
    for (int kb = group; kb < prm.nkb; kb += NUM_GROUPS) {
      const int s = kb % NS, it = kb / NS;
      const bool leader = ((kb / NUM_GROUPS) % GROUP_WARPS) == wg;     // warp-uniform, rotates
      unsigned char* a_tile = smem + (size_t)s * stage_bytes;
      const unsigned char* in_cb = in_base + (size_t)cb * 128;

      // ---- issue the 16 gathers of this k-block (all four corners, unconditionally: unusable
      //      corners carry weight 0 and a guard-band-safe address) ----
      uint4 v[4][4];
      uint4 cur[4];
#pragma unroll
      for (int ps = 0; ps < 4; ++ps) {
        cur[ps] = rec[ps];
        const unsigned char* p0 = in_cb + (long long)(int)cur[ps].x * (long long)rowb;
        v[ps][0] = __ldg(reinterpret_cast<const uint4*>(p0));
        v[ps][1] = __ldg(reinterpret_cast<const uint4*>(p0 + rowb));
        v[ps][2] = __ldg(reinterpret_cast<const uint4*>(p0 + wrow));
        v[ps][3] = __ldg(reinterpret_cast<const uint4*>(p0 + wrow + rowb));
      }
      // ---- records of this group's next k-block: in flight while we combine the current one ----
      tap += NUM_GROUPS;
      while (tap >= K) { tap -= K; ++cb; }
      if (kb + NUM_GROUPS < prm.nkb) {
#pragma unroll
        for (int ps = 0; ps < 4; ++ps) rec[ps] = __ldg(plan_t + (size_t)ps * 32 * K + tap);
      }
      mbar_wait(&empty_bar[s], (it & 1) ^ 1);      // the gathers above are already in flight
      if (leader && lane == 0 && (NS < 2 * NUM_GROUPS || kb < NUM_GROUPS)) {
        // weight slab of this k-block -> smem (async).  With >= 4 stages only the group's first
        // k-block is fetched here; later slabs are prefetched one k-block ahead (below).
        mbar_arrive_expect_tx(&full_bar[s], b_bytes);
        bulk_g2s(a_tile + MT::A_TILES * A_TILE_BYTES, prm.wp + (size_t)kb * b_src_stride, b_bytes,
                 &full_bar[s]);
      }
#pragma unroll
      for (int ps = 0; ps < 4; ++ps) {
        unsigned char* dst = a_tile + a_off + ps * 4096;
        if constexpr (MODE == MODE_BF16) {
          // packed bf16 interpolation: the record carries bf16x2 (w0,w1) and (w2,w3)
          const uint32_t w0 = __byte_perm(cur[ps].y, 0u, 0x1010), w1 = __byte_perm(cur[ps].y, 0u, 0x3232);
          const uint32_t w2 = __byte_perm(cur[ps].z, 0u, 0x1010), w3 = __byte_perm(cur[ps].z, 0u, 0x3232);
          uint4 o;
          o.x = bf162_fma(w3, v[ps][3].x, bf162_fma(w2, v[ps][2].x, bf162_fma(w1, v[ps][1].x, bf162_mul(w0, v[ps][0].x))));
          o.y = bf162_fma(w3, v[ps][3].y, bf162_fma(w2, v[ps][2].y, bf162_fma(w1, v[ps][1].y, bf162_mul(w0, v[ps][0].y))));
          o.z = bf162_fma(w3, v[ps][3].z, bf162_fma(w2, v[ps][2].z, bf162_fma(w1, v[ps][1].z, bf162_mul(w0, v[ps][0].z))));
          o.w = bf162_fma(w3, v[ps][3].w, bf162_fma(w2, v[ps][2].w, bf162_fma(w1, v[ps][1].w, bf162_mul(w0, v[ps][0].w))));
          *reinterpret_cast<uint4*>(dst) = o;
        } else {
          float w[4];
          decode_rec(make_float4(0.f, __uint_as_float(cur[ps].y), __uint_as_float(cur[ps].z),
                                 __uint_as_float(cur[ps].w)), w);
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            a0 = fmaf(w[i], __uint_as_float(v[ps][i].x), a0);
            a1 = fmaf(w[i], __uint_as_float(v[ps][i].y), a1);
            a2 = fmaf(w[i], __uint_as_float(v[ps][i].z), a2);
            a3 = fmaf(w[i], __uint_as_float(v[ps][i].w), a3);
          }
          if constexpr (MODE == MODE_TF32X3) {
            const float h0 = tf32_rna(a0), h1 = tf32_rna(a1), h2 = tf32_rna(a2), h3 = tf32_rna(a3);
            *reinterpret_cast<float4*>(dst) = make_float4(h0, h1, h2, h3);
            *reinterpret_cast<float4*>(dst + A_TILE_BYTES) = make_float4(a0 - h0, a1 - h1, a2 - h2, a3 - h3);
          } else {
            *reinterpret_cast<float4*>(dst) = make_float4(a0, a1, a2, a3);
          }
        }
      }
      fence_proxy_async_smem();   // my generic-proxy stores -> visible to tcgen05.mma
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[s]);

      if (leader) {
        // ---- MMA issue for this k-block (one thread), into this group's accumulator ----
        issued_any = true;
        if (lane == 0) {
          mbar_wait(&full_bar[s], it & 1);         // A tile from 8 warps + weight bytes landed
          tc_fence_after();
          const uint32_t a_addr = smem_u32(a_tile);
          const uint32_t b_addr = a_addr + MT::A_TILES * A_TILE_BYTES;
          const uint64_t adesc = make_sw128_kmajor_desc(a_addr);
          const uint64_t bdesc = make_sw128_kmajor_desc(b_addr);
#pragma unroll
          for (int k = 0; k < 4; ++k) {            // 4 x 32 bytes of K per 128-byte row
            const uint32_t acc = (kb >= NUM_GROUPS || k > 0) ? 1u : 0u;
            if constexpr (MODE == MODE_BF16) {
              umma_f16(acc_tmem, adesc + 2 * k, bdesc + 2 * k, prm.idesc, acc);
            } else if constexpr (MODE == MODE_TF32) {
              umma_tf32(acc_tmem, adesc + 2 * k, bdesc + 2 * k, prm.idesc, acc);
            } else {
              const uint64_t adesc_lo = make_sw128_kmajor_desc(a_addr + A_TILE_BYTES);
              const uint64_t bdesc_lo = make_sw128_kmajor_desc(b_addr + b_tile_bytes);
              umma_tf32(acc_tmem, adesc_lo + 2 * k, bdesc + 2 * k, prm.idesc, acc);     // Alo.Bhi
              umma_tf32(acc_tmem, adesc + 2 * k, bdesc_lo + 2 * k, prm.idesc, 1u);      // Ahi.Blo
              umma_tf32(acc_tmem, adesc + 2 * k, bdesc + 2 * k, prm.idesc, 1u);         // Ahi.Bhi
            }
          }
          tc_commit(&empty_bar[s]);                // frees the stage when these MMAs retire
          if (NS >= 2 * NUM_GROUPS && kb + NUM_GROUPS < prm.nkb) {
            // prefetch the weight slab of this group's next k-block: its stage was freed by an MMA
            // issued two of the group's k-blocks ago, so this wait does not block in steady state
            const int kn = kb + NUM_GROUPS, sn = kn % NS, itn = kn / NS;
            mbar_wait(&empty_bar[sn], (itn & 1) ^ 1);
            mbar_arrive_expect_tx(&full_bar[sn], b_bytes);
            bulk_g2s(smem + (size_t)sn * stage_bytes + MT::A_TILES * A_TILE_BYTES,
                     prm.wp + (size_t)kn * b_src_stride, b_bytes, &full_bar[sn]);
          }
        }
        __syncwarp();
      }
    }
    if (issued_any && lane == 0) tc_commit(tmem_full_bar);   // all MMAs I issued have retired

    // ===================== epilogue: TMEM -> registers -> NCHW global =====================
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int q = warp & 3, cgrp = warp >> 2;      // TMEM lane quarter / column group of this warp
    const int row = q * 32 + lane;
    const int m = m0 + row;
    const bool row_ok = m < prm.M;
    const int n = row_ok ? m / prm.HoWo : 0;
    const int pos = row_ok ? m - n * prm.HoWo : 0;
    Tout* obase = reinterpret_cast<Tout*>(prm.out) + ((size_t)n * prm.out_ctot + prm.out_coff) * prm.HoWo + pos;
    const int cols_per_warp = (BN / 4 >= 32) ? BN / 4 : 32;
    for (int c0 = 0; c0 < cols_per_warp; c0 += 32) {
      const int col = cgrp * cols_per_warp + c0;
      if (col >= BN) break;                        // warp-uniform
      uint32_t acc0[32], acc1[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)col;
      tmem_ld32(taddr, acc0);
      if (active_groups > 1) tmem_ld32(taddr + (uint32_t)BN, acc1);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = __uint_as_float(acc0[j]);
          if (active_groups > 1) x += __uint_as_float(acc1[j]);
          if (prm.bias) x += __ldg(prm.bias + col + j);
          if (prm.relu) x = fmaxf(x, 0.f);
          st_out<Tout>(obase + (size_t)(col + j) * prm.HoWo, x);   // lanes = consecutive positions
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, prm.tmem_cols);
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
    case 4: return launch_umma<MODE, 4, Tout>(p, grid, stream);
    default: set_error("dcn umma: unsupported stage count %d", ns); return KGDET_ERR_INVALID_ARG;
  }
}

int umma_forward(const DcnGeom& g, const void* in_nhwc, const SampleRec16* plan, const void* packed_w,
                 const float* bias, const OutSpec& o, int precision, cudaStream_t stream) {
  if (!umma_supported(g, precision)) {
    set_error("dcn umma: shape/precision not supported by the tensor-core path");
    return KGDET_ERR_UNSUPPORTED;
  }
  const int mode = mode_of(precision);
  const int bk = bk_of(precision);
  UmmaParams p;
  p.in = in_nhwc; p.plan = plan; p.wp = (const unsigned char*)packed_w; p.bias = bias; p.out = o.out;
  p.out_coff = o.coff; p.out_ctot = o.ctot; p.relu = o.relu;
  p.M = g.M; p.C = g.C; p.W = g.W; p.Cout = g.Cout; p.K = g.K; p.HoWo = g.Ho * g.Wo;
  p.nkb = (g.C / bk) * g.K;
  p.idesc = make_idesc(mode == MODE_BF16 ? 1u : 2u, BM, (uint32_t)g.Cout);
  p.tmem_cols = g.Cout <= 64 ? 128 : (g.Cout <= 128 ? 256 : 512);   // two accumulators, power of two
  const int grid = ceil_div(g.M, BM);
  // pipeline depth: as deep as 227 KB allows, capped so that some L1 is left for the gather
  int ns = (mode == MODE_TF32X3) ? 2 : 4;
  if (const char* e = getenv("KGDET_UMMA_STAGES")) {
    int v = atoi(e);
    if (v == 2 || v == 4) ns = v;
  }
  // The stage count must be even: stage s then always belongs to producer group s % 2, and a warp
  // can never run a full mbarrier-parity period ahead of the MMA that frees its stage.
  while (ns > 2 && umma_smem_bytes(mode, ns, g.Cout) > 227 * 1024) ns -= 2;
  if (umma_smem_bytes(mode, ns, g.Cout) > 227 * 1024) {
    set_error("dcn umma: tile does not fit shared memory");
    return KGDET_ERR_UNSUPPORTED;
  }
  const bool f32 = o.dtype == KGDET_F32;
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
