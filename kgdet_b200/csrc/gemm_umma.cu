// Plain bf16 tensor-core GEMM used by the deformable-convolution backward pass:
//
//     C[M, N] (+)= sum_k A[m, k] * B[n, k]          A, B row-major bf16 (K contiguous), fp32 accumulate
//
// tcgen05.mma (UTCHMMA) with both operands K-major in 128B-swizzled shared memory, accumulator in TMEM.
// One CTA = 128 x BN output tile; 8 producer warps copy operand tiles global -> swizzled smem with
// 16-byte cp.async, three k-blocks in flight (rows beyond M / N are zero-filled), 1 warp issues the MMAs, the
// producers then drain TMEM.  grid.z splits the K range; split results are combined with fp32 atomics
// (red.global.add) into a zero-initialised C.  K must be a multiple of 64.
#include "dcn.cuh"

namespace kgdet {

static constexpr int G_BM = 128;
static constexpr int G_PROD_WARPS = 8;
static constexpr int G_THREADS = (G_PROD_WARPS + 1) * 32;
static constexpr int G_NS = 4;          // pipeline stages of the default instantiations

struct GemmParams {
  const __nv_bfloat16* A;
  const __nv_bfloat16* B;
  void* C;
  long long lda, ldb, ldc;
  int M, N;
  int kblocks_total;     // K / 64
  int kblocks_per_split;
  uint32_t idesc;
  uint32_t tmem_cols;
  int atomic;            // 1: red.add fp32 into C (split-K), 0: plain store
  float alpha;
};

template <typename T> __device__ __forceinline__ void st4(T* p, float a, float b, float c, float d);
template <> __device__ __forceinline__ void st4<float>(float* p, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
template <> __device__ __forceinline__ void st4<__nv_bfloat16>(__nv_bfloat16* p, float a, float b, float c, float d) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
  uint2 v;
  v.x = *reinterpret_cast<uint32_t*>(&lo);
  v.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p) = v;
}

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}

// NS pipeline stages, MINB resident CTAs per SM.  <128, bf16, 3, 2>: the column-gradient GEMM of the DCN backward has
// only K / 64 = 4 k-blocks per tile, so a tile is mostly fill + epilogue (512 clk of MMA in ~5 500); two resident
// CTAs of 128 x 128 tiles (96 KB of stages, 128 TMEM columns each) let one tile's epilogue run under the other's loads.
template <int BN, typename Tout, int NS, int MINB>
__global__ void __launch_bounds__(G_THREADS, MINB) umma_gemm_kernel(const GemmParams prm) {
  constexpr int G_NS = NS;
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  constexpr int A_BYTES = G_BM * 128, B_BYTES = BN * 128, STAGE = A_BYTES + B_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)G_NS * STAGE);
  uint64_t* empty_bar = full_bar + G_NS;
  uint64_t* tmem_full_bar = empty_bar + G_NS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * G_BM, n0 = blockIdx.y * BN;
  const int kb0 = blockIdx.z * prm.kblocks_per_split;
  int nkb = prm.kblocks_total - kb0;
  if (nkb > prm.kblocks_per_split) nkb = prm.kblocks_per_split;

  if (warp == G_PROD_WARPS) {
    if (lane == 0) {
      for (int s = 0; s < G_NS; ++s) {
        mbar_init(&full_bar[s], G_PROD_WARPS);
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

  if (warp < G_PROD_WARPS) {
    // Producers: 16-byte cp.async (LDGSTS) straight into the swizzled tiles, no register staging; DEPTH
    // k-blocks of copies stay in flight per thread (one commit group per k-block).  A thread's group for
    // k-block kb has landed after cp.async.wait_group DEPTH-1; it then makes its bytes visible to the async
    // proxy (tcgen05.mma reads) and arrives on the stage's full barrier.  The old loop held one k-block in
    // registers and paid a full L2 round trip per k-block (~2 000 clk against a 512-clk MMA).
    const int chunk = tid & 7, rbase = tid >> 3;     // 32 rows per pass
    constexpr int DEPTH = G_NS - 1;                  // < G_NS: the stage of kb + DEPTH was freed by MMA(kb + DEPTH - G_NS)
    const int off = rbase * 128 + ((chunk ^ (rbase & 7)) << 4);
    auto issue = [&](int kb) {
      const int s = kb % G_NS, it = kb / G_NS;
      mbar_wait(&empty_bar[s], (it & 1) ^ 1);
      const long long kofs = (long long)(kb0 + kb) * 64 + chunk * 8;
      const uint32_t a_dst = smem_u32(smem + (size_t)s * STAGE + off);
      const uint32_t b_dst = a_dst + A_BYTES;
#pragma unroll
      for (int p = 0; p < G_BM / 32; ++p) {
        const int r = m0 + rbase + p * 32;
        const bool ok = r < prm.M;
        const __nv_bfloat16* src = prm.A + (long long)(ok ? r : 0) * prm.lda + kofs;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(a_dst + p * 4096), "l"(src), "r"(ok ? 16 : 0) : "memory");
      }
#pragma unroll
      for (int p = 0; p < BN / 32; ++p) {
        const int r = n0 + rbase + p * 32;
        const bool ok = r < prm.N;
        const __nv_bfloat16* src = prm.B + (long long)(ok ? r : 0) * prm.ldb + kofs;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(b_dst + p * 4096), "l"(src), "r"(ok ? 16 : 0) : "memory");
      }
    };
    for (int kb = 0; kb < DEPTH - 1; ++kb) {
      if (kb < nkb) issue(kb);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int kb = 0; kb < nkb; ++kb) {
      if (kb + DEPTH - 1 < nkb) issue(kb + DEPTH - 1);
      asm volatile("cp.async.commit_group;" ::: "memory");          // (possibly empty) group of k-block kb + DEPTH - 1
      asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");   // k-block kb has landed
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[kb % G_NS]);
    }
    // ---- epilogue ----
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int q = warp & 3, half = warp >> 2;
    const int row = q * 32 + lane;
    const int m = m0 + row;
    constexpr int HALF = BN / 2;
    bool staged = false;
    if constexpr (sizeof(Tout) == 2) {
      // bf16 output, full-width tile: stage the 128 x 256 tile in shared memory (the pipeline buffers are idle
      // once the accumulator is complete) and write whole 512-byte rows, one per warp instruction.  A thread owns
      // an accumulator ROW, so direct stores put 32 different lines (16 bytes each) into every instruction; the
      // column-gradient GEMM of the DCN backward has only 4 k-blocks per tile and was bound by exactly that.
      if (!prm.atomic && n0 + BN <= prm.N && nkb > 0 && (prm.ldc & 7) == 0 &&
          (reinterpret_cast<uintptr_t>(prm.C) & 15) == 0) {
        constexpr int PITCH = BN * 2 + 16;                  // bytes; +16 keeps the 16-byte row writes conflict-free
        unsigned char* st = smem;
        for (int c0 = 0; c0 < HALF; c0 += 32) {
          const int col = half * HALF + c0;
          uint32_t acc[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)col, acc);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 v;
            v.x = pack2(prm.alpha * __uint_as_float(acc[j]), prm.alpha * __uint_as_float(acc[j + 1]));
            v.y = pack2(prm.alpha * __uint_as_float(acc[j + 2]), prm.alpha * __uint_as_float(acc[j + 3]));
            v.z = pack2(prm.alpha * __uint_as_float(acc[j + 4]), prm.alpha * __uint_as_float(acc[j + 5]));
            v.w = pack2(prm.alpha * __uint_as_float(acc[j + 6]), prm.alpha * __uint_as_float(acc[j + 7]));
            *reinterpret_cast<uint4*>(st + (size_t)row * PITCH + (col + j) * 2) = v;
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(G_PROD_WARPS * 32) : "memory");      // the eight epilogue warps only
        __nv_bfloat16* cbase = reinterpret_cast<__nv_bfloat16*>(prm.C) + n0;
        if constexpr (BN == 256) {
          for (int r = warp; r < G_BM; r += G_PROD_WARPS) {
            if (m0 + r < prm.M)
              *reinterpret_cast<uint4*>(cbase + (long long)(m0 + r) * prm.ldc + lane * 8) =
                  *reinterpret_cast<const uint4*>(st + (size_t)r * PITCH + lane * 16);
          }
        } else {                                            // 256-byte rows: two per warp instruction
          for (int r2 = warp * 2; r2 < G_BM; r2 += 2 * G_PROD_WARPS) {
            const int r = r2 + (lane >> 4), piece = lane & 15;
            if (m0 + r < prm.M)
              *reinterpret_cast<uint4*>(cbase + (long long)(m0 + r) * prm.ldc + piece * 8) =
                  *reinterpret_cast<const uint4*>(st + (size_t)r * PITCH + piece * 16);
          }
        }
        staged = true;
      }
    }
    for (int c0 = 0; c0 < HALF && !staged; c0 += 32) {
      const int col = half * HALF + c0;
      uint32_t acc[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)col, acc);
      tmem_ld_wait();
      if (m < prm.M && nkb > 0) {
        const int n = n0 + col;
        if (prm.atomic) {
          float* crow = reinterpret_cast<float*>(prm.C) + (long long)m * prm.ldc + n;
          if (n + 32 <= prm.N && ((reinterpret_cast<uintptr_t>(crow) & 15) == 0)) {
            // red.global.add.v4.f32: 8 vector reductions per thread instead of 32 scalar ones (lanes are
            // different rows, so every scalar atomic was its own L2 transaction)
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              atomicAdd(reinterpret_cast<float4*>(crow + j),
                        make_float4(prm.alpha * __uint_as_float(acc[j]), prm.alpha * __uint_as_float(acc[j + 1]),
                                    prm.alpha * __uint_as_float(acc[j + 2]), prm.alpha * __uint_as_float(acc[j + 3])));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n + j < prm.N) atomicAdd(crow + j, prm.alpha * __uint_as_float(acc[j]));
          }
        } else {
          Tout* crow = reinterpret_cast<Tout*>(prm.C) + (long long)m * prm.ldc + n;
          if (n + 32 <= prm.N && sizeof(Tout) == 2 && ((reinterpret_cast<uintptr_t>(crow) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {            // 16-byte stores: whole 32-byte sectors after two of them
              uint4 v;
              v.x = pack2(prm.alpha * __uint_as_float(acc[j]), prm.alpha * __uint_as_float(acc[j + 1]));
              v.y = pack2(prm.alpha * __uint_as_float(acc[j + 2]), prm.alpha * __uint_as_float(acc[j + 3]));
              v.z = pack2(prm.alpha * __uint_as_float(acc[j + 4]), prm.alpha * __uint_as_float(acc[j + 5]));
              v.w = pack2(prm.alpha * __uint_as_float(acc[j + 6]), prm.alpha * __uint_as_float(acc[j + 7]));
              *reinterpret_cast<uint4*>(crow + j) = v;
            }
          } else if (n + 32 <= prm.N) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              st4<Tout>(crow + j, prm.alpha * __uint_as_float(acc[j]), prm.alpha * __uint_as_float(acc[j + 1]),
                        prm.alpha * __uint_as_float(acc[j + 2]), prm.alpha * __uint_as_float(acc[j + 3]));
          } else {
            for (int j = 0; j < 32 && n + j < prm.N; ++j) {
              const float v = prm.alpha * __uint_as_float(acc[j]);
              if constexpr (sizeof(Tout) == 4) reinterpret_cast<float*>(crow)[j] = v;
              else reinterpret_cast<__nv_bfloat16*>(crow)[j] = __float2bfloat16(v);
            }
          }
        }
      }
    }
  } else {
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % G_NS, it = kb / G_NS;
        mbar_wait(&full_bar[s], it & 1);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + (size_t)s * STAGE);
        const uint64_t adesc = make_sw128_kmajor_desc(a_addr);
        const uint64_t bdesc = make_sw128_kmajor_desc(a_addr + A_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tmem_base, adesc + 2 * k, bdesc + 2 * k, prm.idesc, (kb > 0 || k > 0) ? 1u : 0u);
        tc_commit(&empty_bar[s]);
      }
      tc_commit(tmem_full_bar);
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == G_PROD_WARPS) tmem_dealloc(tmem_base, prm.tmem_cols);
}

template <int BN, typename Tout, int NS = G_NS, int MINB = 1>
static int launch_gemm(const GemmParams& p, dim3 grid, cudaStream_t stream) {
  const size_t smem = 1024 + (size_t)NS * (G_BM * 128 + BN * 128) + (2 * NS + 1) * 8 + 16;
  KG_CUDA(cudaFuncSetAttribute(umma_gemm_kernel<BN, Tout, NS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_gemm_kernel<BN, Tout, NS, MINB><<<grid, G_THREADS, smem, stream>>>(p);
  KG_LAUNCH_CHECK("umma_gemm_kernel");
  return KGDET_OK;
}

// C[M,N] = alpha * A[M,K] . B[N,K]^T.  out_dtype: KGDET_F32 / KGDET_BF16.  splits > 1 requires fp32 C that
// the caller has zeroed (results are accumulated with atomics).  K % 64 == 0, lda/ldb % 8 == 0.
int umma_gemm(const __nv_bfloat16* A, long long lda, const __nv_bfloat16* B, long long ldb, void* C,
              long long ldc, int M, int N, int K, int out_dtype, int splits, float alpha,
              cudaStream_t stream) {
  KG_CHECK_ARG(K % 64 == 0 && lda % 8 == 0 && ldb % 8 == 0, "umma_gemm: K %% 64 and ld %% 8 required");
  KG_CHECK_ARG(splits >= 1 && (splits == 1 || out_dtype == KGDET_F32), "umma_gemm: split-K needs fp32 output");
  if (M <= 0 || N <= 0 || K <= 0) return KGDET_OK;
  GemmParams p;
  p.A = A; p.B = B; p.C = C; p.lda = lda; p.ldb = ldb; p.ldc = ldc; p.M = M; p.N = N;
  p.kblocks_total = K / 64;
  p.kblocks_per_split = ceil_div(p.kblocks_total, splits);
  splits = ceil_div(p.kblocks_total, p.kblocks_per_split);
  p.atomic = splits > 1 ? 1 : 0;
  p.alpha = alpha;
  // few k-blocks, bf16 result (the DCN column gradient): 128-column tiles, two CTAs per SM
  bool narrow = out_dtype == KGDET_BF16 && splits == 1 && p.kblocks_total <= 8 && N > 128;
  if (const char* e = getenv("KGDET_GEMM_NARROW")) narrow = narrow && atoi(e) != 0;
  const int BN = (N > 128 && !narrow) ? 256 : 128;
  p.idesc = make_idesc(1u, G_BM, (uint32_t)BN);
  p.tmem_cols = BN;
  dim3 grid(ceil_div(M, G_BM), ceil_div(N, BN), splits);
  KG_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "umma_gemm: grid too large");
  if (narrow) return launch_gemm<128, __nv_bfloat16, 3, 2>(p, grid, stream);
  if (BN == 256)
    return out_dtype == KGDET_F32 ? launch_gemm<256, float>(p, grid, stream)
                                  : launch_gemm<256, __nv_bfloat16>(p, grid, stream);
  return out_dtype == KGDET_F32 ? launch_gemm<128, float>(p, grid, stream)
                                : launch_gemm<128, __nv_bfloat16>(p, grid, stream);
}

}  // namespace kgdet
