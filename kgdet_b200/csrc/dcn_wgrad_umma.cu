// Gather-fused weight gradient of the deformable convolution on the tensor cores (bf16 mode).
//
//   gW^T[(tap, c), o] = sum_m S(m, tap, c) * gO[m, o]              m = (n, y, x) output position
//
// The reference recomputes the whole im2col into HBM and calls cuBLAS (deform_conv_cuda.cpp:373-484); the
// first tensor-core version here (dcn_bwd_tc.cu) still materialised the transposed columns in L2-sized chunks
// (gather_colT_kernel 0.44 ms + GEMM 0.23 ms for the K = 49 KGDet call).  In this kernel the sampled tile only
// ever exists in shared memory, exactly as in the fused forward:
//
//   CTA            one tap x a contiguous range of 64-position blocks (the split over positions fills the SMs);
//                  D = [C channels x Cout] fp32 in TMEM, as C / 128 accumulators of 128 lanes x Cout columns.
//   warps 0..7     producers: the forward's gather (8 lanes per 128-byte pixel slab of the channel-blocked
//                  planes, all four corners unconditionally, packed HFMA2.BF16 with the plan's pre-rounded
//                  corner weights) into a [64 positions x 128 B] tile per 64-channel slab.  For THIS product the
//                  reduction runs over positions, so the same bytes are an **MN-major** A operand (M = channels
//                  contiguous, K = positions = rows): canonical SWIZZLE_128B MN-major layout, 8 K-rows x 128 B
//                  per swizzle atom, SBO = 1024 B between 8-position groups, LBO = 8 KB between 64-channel slabs.
//   warp 8         control lane: bulk-copies the gO tile of the position block (pre-tiled K-major
//                  [Cout rows x 64 positions], one contiguous cp.async.bulk, written by go_to_tiled_kernel),
//                  issues 4 K-steps x (C / 128) tcgen05.mma (M128 x N=Cout x K16), commits the stage.
//   epilogue       tcgen05.ld -> red.global.add.v4.f32 into gW^T (the position splits meet there).
#include <cuda_bf16.h>

#include "dcn_umma.cuh"

namespace kgdet {

static constexpr int WG_ROWS = 64;                 // positions per pipeline stage (= one gO k-block)
static constexpr int WG_NS = 3;
static constexpr int WG_SLAB_BYTES = WG_ROWS * 128;
static constexpr int WG_PROD_WARPS = 8;
static constexpr int WG_THREADS = (WG_PROD_WARPS + 1) * 32;

struct WgradParams {
  const unsigned char* in;     // first pixel of plane 0 of the channel-blocked bf16 planes
  size_t plane_bytes;
  const uint4* plan;           // SampleRec16 (bf16-weight flavour), tap-major [K][rows_padded]
  const unsigned char* go;     // tiled gO: [position block][Cout rows x 128 B, swizzled]
  float* gwt;                  // [K * C, Cout] fp32, zeroed by the caller
  int C, Cout, W, rows_padded;
  int nblocks;                 // ceil(M / 64)
  int blocks_per_split;
  uint32_t idesc, tmem_cols;
};

__device__ __forceinline__ uint64_t make_sw128_mnmajor_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;   // LBO: next 64-element group along M
  d |= (uint64_t)(1024 >> 4) << 32;                   // SBO: next group of 8 K rows
  d |= (uint64_t)1 << 46;                             // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ uint32_t hfma2_bf16(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
__device__ __forceinline__ uint32_t hmul2_bf16(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
// broadcast the low / high bf16 of a packed pair to both halves
__device__ __forceinline__ uint32_t bcast_lo(uint32_t v) { return __byte_perm(v, v, 0x1010); }
__device__ __forceinline__ uint32_t bcast_hi(uint32_t v) { return __byte_perm(v, v, 0x3232); }

template <int SLABS>
__global__ void __launch_bounds__(WG_THREADS, 1) dcn_wgrad_umma_kernel(const WgradParams prm) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  constexpr int A_BYTES = SLABS * WG_SLAB_BYTES;
  const int b_bytes = prm.Cout * 128;
  const int stage_bytes = A_BYTES + b_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)WG_NS * stage_bytes);
  uint64_t* empty_bar = full_bar + WG_NS;
  uint64_t* tmem_full_bar = empty_bar + WG_NS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tap = blockIdx.x;
  const int pb0 = blockIdx.y * prm.blocks_per_split;
  int nst = prm.nblocks - pb0;
  if (nst > prm.blocks_per_split) nst = prm.blocks_per_split;     // >= 1 by construction of the grid

  if (warp == WG_PROD_WARPS) {
    if (lane == 0) {
      for (int s = 0; s < WG_NS; ++s) {
        mbar_init(&full_bar[s], WG_PROD_WARPS + 1);      // producer warps + the control lane's expect_tx
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

  if (warp == WG_PROD_WARPS) {
    // ------------------------------ control lane ------------------------------
    if (lane == 0) {
      auto fetch_b = [&](int j) {
        const int s = j % WG_NS;
        unsigned char* dst = smem + (size_t)s * stage_bytes + A_BYTES;
        mbar_arrive_expect_tx(&full_bar[s], (uint32_t)b_bytes);
        bulk_g2s(dst, prm.go + (size_t)(pb0 + j) * b_bytes, (uint32_t)b_bytes, &full_bar[s]);
      };
      for (int j = 0; j < WG_NS - 1 && j < nst; ++j) fetch_b(j);
      for (int j = 0; j < nst; ++j) {
        const int s = j % WG_NS;
        mbar_wait(&full_bar[s], (uint32_t)(j / WG_NS) & 1u);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + (size_t)s * stage_bytes);
        const uint64_t bdesc = make_sw128_kmajor_desc(a_addr + A_BYTES);
#pragma unroll
        for (int kk = 0; kk < WG_ROWS / 16; ++kk) {                 // 16 positions per MMA
#pragma unroll
          for (int acc = 0; acc < SLABS / 2; ++acc) {
            const uint64_t adesc =
                make_sw128_mnmajor_desc(a_addr + acc * 2 * WG_SLAB_BYTES + kk * 2048, WG_SLAB_BYTES);
            umma_f16(tmem_base + (uint32_t)(acc * prm.Cout), adesc, bdesc + 2 * kk, prm.idesc,
                     (j > 0 || kk > 0) ? 1u : 0u);
          }
        }
        tc_commit(&empty_bar[s]);
        if (j == nst - 1) tc_commit(tmem_full_bar);
        const int jn = j + WG_NS - 1;                               // its stage was used by block j - 1
        if (jn < nst) {
          if (j >= 1) mbar_wait(&empty_bar[jn % WG_NS], (uint32_t)((j - 1) / WG_NS) & 1u);
          fetch_b(jn);
        }
      }
    }
    __syncwarp();
  } else {
    // -------------------------------- producers --------------------------------
    const int chunk = tid & 7, rbase = tid >> 3;                    // rows rbase and rbase + 32 of the stage
    const unsigned char* in_base = prm.in + chunk * 16;
    const long long wrow = (long long)prm.W * 128;
    const uint4* plan_tap = prm.plan + (size_t)tap * prm.rows_padded;
    for (int j = 0; j < nst; ++j) {
      const int s = j % WG_NS;
      const int m0 = (pb0 + j) * WG_ROWS;
      uint4 rec[2];
      rec[0] = __ldg(plan_tap + m0 + rbase);
      rec[1] = __ldg(plan_tap + m0 + rbase + 32);
      mbar_wait(&empty_bar[s], ((uint32_t)(j / WG_NS) & 1u) ^ 1u);
      unsigned char* a_tile = smem + (size_t)s * stage_bytes;
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int row = rbase + r * 32;
        const unsigned char* p0 = in_base + (long long)(int)rec[r].x * 128;
        uint4 v[SLABS][4];
#pragma unroll
        for (int sl = 0; sl < SLABS; ++sl) {
          const unsigned char* p = p0 + (size_t)sl * prm.plane_bytes;
          v[sl][0] = __ldg(reinterpret_cast<const uint4*>(p));
          v[sl][1] = __ldg(reinterpret_cast<const uint4*>(p + 128));
          v[sl][2] = __ldg(reinterpret_cast<const uint4*>(p + wrow));
          v[sl][3] = __ldg(reinterpret_cast<const uint4*>(p + wrow + 128));
        }
        const uint32_t w0 = bcast_lo(rec[r].y), w1 = bcast_hi(rec[r].y);
        const uint32_t w2 = bcast_lo(rec[r].z), w3 = bcast_hi(rec[r].z);
        const int off = row * 128 + ((chunk ^ (row & 7)) << 4);
#pragma unroll
        for (int sl = 0; sl < SLABS; ++sl) {
          uint4 o;
          o.x = hfma2_bf16(w3, v[sl][3].x, hfma2_bf16(w2, v[sl][2].x, hfma2_bf16(w1, v[sl][1].x, hmul2_bf16(w0, v[sl][0].x))));
          o.y = hfma2_bf16(w3, v[sl][3].y, hfma2_bf16(w2, v[sl][2].y, hfma2_bf16(w1, v[sl][1].y, hmul2_bf16(w0, v[sl][0].y))));
          o.z = hfma2_bf16(w3, v[sl][3].z, hfma2_bf16(w2, v[sl][2].z, hfma2_bf16(w1, v[sl][1].z, hmul2_bf16(w0, v[sl][0].z))));
          o.w = hfma2_bf16(w3, v[sl][3].w, hfma2_bf16(w2, v[sl][2].w, hfma2_bf16(w1, v[sl][1].w, hmul2_bf16(w0, v[sl][0].w))));
          *reinterpret_cast<uint4*>(a_tile + sl * WG_SLAB_BYTES + off) = o;
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[s]);
    }
    // -------------------------------- epilogue --------------------------------
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int q = warp & 3, cgrp = warp >> 2;
    const int row = q * 32 + lane;                                  // channel within the 128-channel accumulator
#pragma unroll
    for (int acc = 0; acc < SLABS / 2; ++acc) {
      float* grow = prm.gwt + ((size_t)tap * prm.C + acc * 128 + row) * prm.Cout;
      for (int col = cgrp * 32; col < prm.Cout; col += 64) {        // warp-uniform
        uint32_t a[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * prm.Cout + col), a);
        tmem_ld_wait();
#pragma unroll
        for (int jj = 0; jj < 32; jj += 4)
          atomicAdd(reinterpret_cast<float4*>(grow + col + jj),
                    make_float4(__uint_as_float(a[jj]), __uint_as_float(a[jj + 1]), __uint_as_float(a[jj + 2]),
                                __uint_as_float(a[jj + 3])));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WG_PROD_WARPS) tmem_dealloc(tmem_base, prm.tmem_cols);
}

// gO NCHW (fp32 / bf16) -> tiled K-major operand: [position block b = m / 64][row o][64 positions], every
// [Cout x 128 B] block in the 128B-swizzled layout tcgen05 reads (16-byte chunk c of row o at c ^ (o & 7)).
// Positions beyond M stay zero (the caller clears the buffer).
template <typename T>
__global__ void go_to_tiled_kernel(const T* __restrict__ go, unsigned char* __restrict__ dst, int N, int Cout,
                                   int HoWo) {
  const long long total = (long long)N * Cout * HoWo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)blockDim.x * gridDim.x) {
    const int p = (int)(i % HoWo);
    const long long r = i / HoWo;
    const int o = (int)(r % Cout), n = (int)(r / Cout);
    float v;
    if constexpr (sizeof(T) == 4) v = go[i]; else v = __bfloat162float(go[i]);
    const long long m = (long long)n * HoWo + p;
    const int k = (int)(m & 63);
    unsigned char* blk = dst + (size_t)(m >> 6) * Cout * 128;
    *reinterpret_cast<__nv_bfloat16*>(blk + (size_t)o * 128 + ((((k >> 3) ^ (o & 7))) << 4) + (k & 7) * 2) =
        __float2bfloat16(v);
  }
}

bool wgrad_fused_supported(const DcnGeom& g) {
  return g.groups == 1 && g.dgroups == 1 && (g.C == 128 || g.C == 256) && g.Cout % 16 == 0 && g.Cout >= 16 &&
         g.Cout <= 256 && (g.C / 128) * g.Cout <= 512;
}

size_t wgrad_fused_go_bytes(const DcnGeom& g) { return (size_t)ceil_div(g.M, WG_ROWS) * g.Cout * 128; }

int launch_go_to_tiled(const DcnGeom& g, const void* grad_output, void* tiled, int dtype, cudaStream_t stream) {
  KG_CUDA(cudaMemsetAsync(tiled, 0, wgrad_fused_go_bytes(g), stream));
  const long long total = (long long)g.N * g.Cout * g.Ho * g.Wo;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (dtype == KGDET_F32)
    go_to_tiled_kernel<float><<<(int)blocks, 256, 0, stream>>>((const float*)grad_output, (unsigned char*)tiled, g.N, g.Cout, g.Ho * g.Wo);
  else
    go_to_tiled_kernel<__nv_bfloat16><<<(int)blocks, 256, 0, stream>>>((const __nv_bfloat16*)grad_output, (unsigned char*)tiled, g.N, g.Cout, g.Ho * g.Wo);
  KG_LAUNCH_CHECK("go_to_tiled_kernel");
  return KGDET_OK;
}

// in_blocked: first pixel of plane 0 (after the guard band); plan: bf16-weight SampleRec16 records; gwt zeroed.
int wgrad_fused(const DcnGeom& g, const void* in_blocked, size_t plane_bytes, const SampleRec16* plan,
                const void* go_tiled, float* gwt, cudaStream_t stream) {
  WgradParams p;
  p.in = (const unsigned char*)in_blocked; p.plane_bytes = plane_bytes; p.plan = (const uint4*)plan;
  p.go = (const unsigned char*)go_tiled; p.gwt = gwt;
  p.C = g.C; p.Cout = g.Cout; p.W = g.W; p.rows_padded = (int)plan_rows(g);
  p.nblocks = ceil_div(g.M, WG_ROWS);
  int splits = num_sms() / g.K;
  if (splits < 1) splits = 1;
  if (splits > p.nblocks) splits = p.nblocks;
  p.blocks_per_split = ceil_div(p.nblocks, splits);
  splits = ceil_div(p.nblocks, p.blocks_per_split);             // every split owns at least one block
  p.idesc = make_idesc(1u, 128u, (uint32_t)g.Cout) | (1u << 15);   // A is MN-major
  const int cols = (g.C / 128) * g.Cout;
  p.tmem_cols = cols <= 32 ? 32 : (cols <= 64 ? 64 : (cols <= 128 ? 128 : (cols <= 256 ? 256 : 512)));
  const int slabs = g.C / 64;
  const size_t smem = 1024 + (size_t)WG_NS * (slabs * WG_SLAB_BYTES + g.Cout * 128) + (2 * WG_NS + 1) * 8 + 16;
  dim3 grid((unsigned)g.K, (unsigned)splits, 1);
  if (slabs == 4) {
    KG_CUDA(cudaFuncSetAttribute(dcn_wgrad_umma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dcn_wgrad_umma_kernel<4><<<grid, WG_THREADS, smem, stream>>>(p);
  } else {
    KG_CUDA(cudaFuncSetAttribute(dcn_wgrad_umma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dcn_wgrad_umma_kernel<2><<<grid, WG_THREADS, smem, stream>>>(p);
  }
  KG_LAUNCH_CHECK("dcn_wgrad_umma_kernel");
  return KGDET_OK;
}

}  // namespace kgdet
