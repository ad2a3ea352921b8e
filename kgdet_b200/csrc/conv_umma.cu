// Plain k x k (stride 1, "same" padding) convolutions of the head towers on the 5th-generation tensor cores
// (mmdet/models/utils/conv_module.py:156-164 conv -> norm -> activation,
// KP3:292-313 the 3 + 3 tower ConvModules, KP3:98-106 the two stage-1 3x3 convolutions).
//
// Why not cuDNN: with TF32 allowed its implicit-GEMM rounds both operands to 11 bits, and the two deformable
// stages that SAMPLE at the predicted points amplify that 8e-4 to 1e-1 at stage 3 of the head (measured against
// the reference golden, profiles/r2_tf32_probe.jsonl); with TF32 off it runs at 1.5 ms per convolution.  This
// kernel is fp32-grade ("bf16x3": operands split into bf16 hi + lo, three MMAs per k-step, fp32 accumulation in
// TMEM -- measured ~1e-5 of the tensor maximum) at tensor-core speed.
//
//   out[n, y, x, o] = bias[o] + sum_{i, j, c} in[n, y + i - p, x + j - p, c] * W[o, c, i, j]         (NHWC fp32 out)
//
// Operands
//   A  the activation as "split planes": channel-blocked bf16 planes [C/64][guard | N*H*W pixels | guard][64]
//      of the hi parts (bit-identical to the fused DCN kernel's prepared input, dcn_api.cu -- the last tower
//      layer's planes ARE the DCN input) followed by the same planes of the lo parts (x - hi).  A k-block = one tap
//      of one 64-channel block = a SHIFTED window of the plane: one 5-D TMA tile load {64 ch, bw, bh, 1, 1} at
//      (0, x0 + j - p, y0 + i - p, n, block) -- the tensor map zero-fills what falls outside the image, which is
//      the convolution's zero padding -- lands as the 128B-swizzled K-major A tile (UTMALDG).
//   B  weights packed per k-block as [hi tile | lo tile], Cout rows x 128 B each in the swizzled K-major layout,
//      one cp.async.bulk per tile half (UBLKCP).
// CTA pair (cta_group::2): two consecutive 128-position tiles share every weight slab -- each CTA loads the A
// tiles of its own rows and HALF of the weight rows, the even CTA issues M = 256 MMAs (UTCHMMA.2CTA) into both
// CTAs' TMEM.  L2 -> SM traffic per k-block and CTA: 32 KB of A + 32 KB of B instead of 32 + 64.
//   warp 0  producer (one lane): TMA + bulk copies, NS-stage mbarrier ring
//   warp 1  even CTA: MMA issuer; odd CTA: relays "my stage is full" to the even CTA's barrier
//   warps 2..9  epilogue: TMEM -> registers -> (+ bias, ReLU) -> NHWC fp32 rows (128 B per thread and chunk)
#include <cuda.h>
#include <cuda_bf16.h>

#include "dcn.cuh"

namespace kgdet {

static constexpr int CV_BM = 128;
static constexpr int CV_A_BYTES = CV_BM * 128;
static constexpr int CV_NS = 3;
static constexpr int CV_EPI_WARPS = 8;
static constexpr int CV_THREADS = (2 + CV_EPI_WARPS) * 32;

struct ConvParams {
  // up to two problems of the same geometry (the classification and the point tower's convolution of one layer):
  // a CTA runs its tile of problem 0, then the same tile of problem 1 into the other half of TMEM -- the epilogue
  // of the first overlaps the main loop of the second
  const unsigned char* wp[2];  // packed weights: per k-block [hi | lo], Cout rows x 128 B each
  const float* bias[2];        // [Cout] or NULL
  float* out[2];               // NHWC fp32 [N, H, W, Cout]
  int nprob;
  int N, H, W, Cout, taps, ksz, pad, ncb;     // ncb = C / 64
  int bw, bh, tiles_w, tiles_h, ntiles;       // tile = bh rows x bw columns of one image (bw * bh <= 128)
  int relu;
  uint32_t idesc, tmem_cols;
  long long* timeline;         // development hook (kgdet_dcn_set_timeline): per CTA 8 clock64() stamps, or NULL
};

__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                            int c4, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::
          "r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote_release(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CV_THREADS, 1)
conv_umma_pair_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                      const __grid_constant__ CUtensorMap map_hi1, const __grid_constant__ CUtensorMap map_lo1,
                      const ConvParams prm) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  const int b_half = (prm.Cout / 2) * 128;                 // weight rows held by THIS CTA, per hi / lo tile
  const int stage_bytes = 2 * CV_A_BYTES + 2 * b_half;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)CV_NS * stage_bytes);
  uint64_t* empty_bar = full_bar + CV_NS;
  uint64_t* tmem_full_bar = empty_bar + CV_NS;                 // [2]: one per problem
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();                 // 0 = the CTA that issues the MMAs
  // stamps: 0 entry, 1 set-up done, 2 first stage full (MMA issuer), 3 last MMA issued, 4 accumulator ready, 5 epilogue done
  long long* const tl = prm.timeline ? prm.timeline + (size_t)blockIdx.x * 8 : nullptr;
  if (tl && tid == 0) tl[0] = clock64();
  const int nkb = prm.ncb * prm.taps;
  // this CTA's tile: image n, rows [y0, y0 + bh), columns [x0, x0 + bw); a padding tile (odd tile count) repeats
  // the last real one and stores nothing
  const int tile_raw = (int)blockIdx.x;
  const bool tile_ok = tile_raw < prm.ntiles;
  const int tile = tile_ok ? tile_raw : prm.ntiles - 1;
  const int per_img = prm.tiles_w * prm.tiles_h;
  const int n = tile / per_img, tr = tile - n * per_img;
  const int y0 = (tr / prm.tiles_w) * prm.bh, x0 = (tr % prm.tiles_w) * prm.bw;

  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < CV_NS; ++s) {
        // even CTA: own producer's expect_tx arrive + the odd CTA's relay; odd CTA: own producer only
        mbar_init(&full_bar[s], rank == 0 ? 2 : 1);
        mbar_init(&empty_bar[s], 1);                       // one tcgen05.commit (multicast to both CTAs)
      }
      mbar_init(&tmem_full_bar[0], 1);
      mbar_init(&tmem_full_bar[1], 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc_pair(tmem_slot, prm.tmem_cols * prm.nprob);
  }
  tc_fence_before();
  cluster_sync_all();                                      // the peer's barriers exist before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tl && tid == 0) tl[1] = clock64();

  if (warp == 0) {
    // =========================== producer ===========================
    if (lane == 0) {
      const uint32_t a_bytes = (uint32_t)(prm.bw * prm.bh * 128);
      const uint32_t tx = 2u * a_bytes + 2u * (uint32_t)b_half;
      const size_t b_tile = (size_t)prm.Cout * 128;
      for (int t = 0; t < prm.nprob; ++t) {
        const CUtensorMap* mh = t ? &map_hi1 : &map_hi;
        const CUtensorMap* ml = t ? &map_lo1 : &map_lo;
        for (int kb = 0; kb < nkb; ++kb) {
          const int g = t * nkb + kb, s = g % CV_NS;          // the stage ring runs on across the two problems
          mbar_wait(&empty_bar[s], (((uint32_t)(g / CV_NS)) & 1u) ^ 1u);
          const int cb = kb / prm.taps, tap = kb - cb * prm.taps;
          const int i = tap / prm.ksz, j = tap - i * prm.ksz;
          unsigned char* dst = smem + (size_t)s * stage_bytes;
          mbar_arrive_expect_tx(&full_bar[s], tx);
          tma_load_5d(dst, mh, 0, x0 + j - prm.pad, y0 + i - prm.pad, n, cb, &full_bar[s]);
          tma_load_5d(dst + CV_A_BYTES, ml, 0, x0 + j - prm.pad, y0 + i - prm.pad, n, cb, &full_bar[s]);
          const unsigned char* w = prm.wp[t] + (size_t)kb * 2 * b_tile + (size_t)rank * b_half;
          bulk_g2s(dst + 2 * CV_A_BYTES, w, (uint32_t)b_half, &full_bar[s]);
          bulk_g2s(dst + 2 * CV_A_BYTES + b_half, w + b_tile, (uint32_t)b_half, &full_bar[s]);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      if (rank != 0) {
        // =========================== relay (odd CTA) ===========================
        const uint32_t remote0 = mapa_u32(smem_u32(&full_bar[0]), 0u);
        for (int g = 0; g < prm.nprob * nkb; ++g) {
          const int s = g % CV_NS;
          mbar_wait(&full_bar[s], ((uint32_t)(g / CV_NS)) & 1u);        // my A tiles and weight half have landed
          mbar_arrive_remote_release(remote0 + (uint32_t)s * 8u);
        }
      } else {
        // =========================== MMA issuer (even CTA) ===========================
        for (int t = 0; t < prm.nprob; ++t) {
          const uint32_t acc = tmem_base + (uint32_t)t * prm.tmem_cols;
          for (int kb = 0; kb < nkb; ++kb) {
            const int g = t * nkb + kb, s = g % CV_NS;
            mbar_wait_cluster(&full_bar[s], ((uint32_t)(g / CV_NS)) & 1u);
            if (tl && g == 0) tl[2] = clock64();
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem + (size_t)s * stage_bytes);
            const uint64_t a_hi = make_sw128_kmajor_desc(a_addr), a_lo = make_sw128_kmajor_desc(a_addr + CV_A_BYTES);
            const uint64_t b_hi = make_sw128_kmajor_desc(a_addr + 2 * CV_A_BYTES);
            const uint64_t b_lo = make_sw128_kmajor_desc(a_addr + 2 * CV_A_BYTES + b_half);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_f16_pair(acc, a_lo + 2 * k, b_hi + 2 * k, prm.idesc, (kb > 0 || k > 0) ? 1u : 0u);   // lo . hi
              umma_f16_pair(acc, a_hi + 2 * k, b_lo + 2 * k, prm.idesc, 1u);                            // hi . lo
              umma_f16_pair(acc, a_hi + 2 * k, b_hi + 2 * k, prm.idesc, 1u);                            // hi . hi
            }
            tc_commit_pair(&empty_bar[s], (uint16_t)3);    // frees the stage in both CTAs when these MMAs retire
          }
          tc_commit_pair(&tmem_full_bar[t], (uint16_t)3);  // this problem's accumulators of both CTAs complete
        }
        if (tl) tl[3] = clock64();
      }
    }
    __syncwarp();
  } else {
    // =========================== epilogue: TMEM -> NHWC fp32 ===========================
    const int q = warp & 3, half = (warp - 2) >> 2;        // TMEM lane quarter (hardware: warp % 4), column half
    const int row = q * 32 + lane;
    const int hh = row / prm.bw, ww = row - hh * prm.bw;
    const int y = y0 + hh, x = x0 + ww;
    const bool ok = tile_ok && row < prm.bw * prm.bh && y < prm.H && x < prm.W;
    const size_t ooff = (((size_t)n * prm.H + (ok ? y : 0)) * prm.W + (ok ? x : 0)) * prm.Cout;
    const int chalf = prm.Cout / 2;
    for (int t = 0; t < prm.nprob; ++t) {
      mbar_wait_cluster(&tmem_full_bar[t], 0);
      if (tl && tid == 64 && t == 0) tl[4] = clock64();
      tc_fence_after();
      float* orow = prm.out[t] + ooff;
      const float* bias = prm.bias[t];
      for (int c0 = 0; c0 < chalf; c0 += 32) {
        const int col = half * chalf + c0;
        uint32_t acc[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * prm.tmem_cols + col), acc);
        tmem_ld_wait();
        if (ok) {
#pragma unroll
          for (int jj = 0; jj < 32; jj += 4) {
            float4 v = make_float4(__uint_as_float(acc[jj]), __uint_as_float(acc[jj + 1]), __uint_as_float(acc[jj + 2]),
                                   __uint_as_float(acc[jj + 3]));
            if (bias) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(bias + col + jj));
              v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
            }
            if (prm.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            *reinterpret_cast<float4*>(orow + col + jj) = v;
          }
        }
      }
    }
  }
  if (tl && tid == 64) tl[5] = clock64();
  tc_fence_before();
  cluster_sync_all();            // neither CTA may free the shared TMEM allocation while the other still reads
  if (warp == 1) tmem_dealloc_pair(tmem_base, prm.tmem_cols * prm.nprob);
}

// ---- weight packing: fp32 [Cout, Cin, k, k] -> per k-block (cb * taps + tap) [hi tile | lo tile] ------------------
__global__ void conv_pack_kernel(const float* __restrict__ w, unsigned char* __restrict__ p, int Cout, int Cin, int taps) {
  const int chunks_per_blk = Cout * 8;                       // 16-byte chunks (8 channels) of one tile
  const int nkb = (Cin / 64) * taps;
  const long long total = (long long)nkb * chunks_per_blk;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)blockDim.x * gridDim.x) {
    const int kb = (int)(idx / chunks_per_blk), r = (int)(idx - (long long)kb * chunks_per_blk);
    const int o = r >> 3, chunk = r & 7;
    const int cb = kb / taps, tap = kb - cb * taps;
    __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = cb * 64 + chunk * 8 + e;
      const float x = w[((size_t)o * Cin + c) * taps + tap];
      hi[e] = __float2bfloat16(x);
      lo[e] = __float2bfloat16(x - __bfloat162float(hi[e]));
    }
    const size_t tile = (size_t)Cout * 128;
    const size_t dst = (size_t)kb * 2 * tile + (size_t)(o >> 3) * 1024 + (o & 7) * 128 + ((chunk ^ (o & 7)) << 4);
    *reinterpret_cast<uint4*>(p + dst) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(p + dst + tile) = *reinterpret_cast<const uint4*>(lo);
  }
}

// ---- split planes: fp32 activation -> bf16 hi planes (DCN prepared-input layout) + lo planes -------------------
// NCHW source: 32 x 32 transpose tiles through shared memory
__global__ void nchw_to_split_planes_kernel(const float* __restrict__ src, unsigned char* __restrict__ hi,
                                            unsigned char* __restrict__ lo, int C, int S, size_t plane_bytes) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float* s = src + (size_t)n * C * S;
  const int tx = threadIdx.x, ty = threadIdx.y;              // 32 x 8
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int c = c0 + ty + k, p = p0 + tx;
    tile[ty + k][tx] = (c < C && p < S) ? s[(size_t)c * S + p] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int p = p0 + ty + k, c = c0 + tx;
    if (c < C && p < S) {
      const float v = tile[tx][ty + k];
      const __nv_bfloat16 h = __float2bfloat16(v);
      const size_t off = (size_t)(c >> 6) * plane_bytes + ((size_t)n * S + p) * 128 + (size_t)(c & 63) * 2;
      *reinterpret_cast<__nv_bfloat16*>(hi + off) = h;
      *reinterpret_cast<__nv_bfloat16*>(lo + off) = __float2bfloat16(v - __bfloat162float(h));
    }
  }
}

// position-major (NHWC) source: one thread per 8 channels
__global__ void rows_to_split_planes_kernel(const float* __restrict__ rows, unsigned char* __restrict__ hi,
                                            unsigned char* __restrict__ lo, long long M, int C, size_t plane_bytes) {
  const int chunks = C / 8;
  const long long total = M * chunks;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)blockDim.x * gridDim.x) {
    const long long m = i / chunks;
    const int ch = (int)(i - m * chunks);
    const float* src = rows + m * C + (size_t)ch * 8;
    float v[8];
    *reinterpret_cast<float4*>(v) = *reinterpret_cast<const float4*>(src);
    *reinterpret_cast<float4*>(v + 4) = *reinterpret_cast<const float4*>(src + 4);
    uint4 h, l;
    uint32_t* hp = reinterpret_cast<uint32_t*>(&h);
    uint32_t* lp = reinterpret_cast<uint32_t*>(&l);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const __nv_bfloat162 hv = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
      const __nv_bfloat162 lv = __floats2bfloat162_rn(v[2 * e] - __low2float(hv), v[2 * e + 1] - __high2float(hv));
      hp[e] = *reinterpret_cast<const uint32_t*>(&hv);
      lp[e] = *reinterpret_cast<const uint32_t*>(&lv);
    }
    const size_t off = (size_t)(ch >> 3) * plane_bytes + (size_t)m * 128 + (size_t)(ch & 7) * 16;
    *reinterpret_cast<uint4*>(hi + off) = h;
    *reinterpret_cast<uint4*>(lo + off) = l;
  }
}

// ---- tensor maps ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

struct SplitLayout { size_t guard_bytes, in_bytes, plane_bytes, half_bytes; int planes; };
static SplitLayout split_layout(int N, int C, int H, int W) {
  SplitLayout l;
  l.planes = C / 64;
  l.guard_bytes = (size_t)(W + 2) * 128;                    // dcn_guard_pixels() slabs, as dcn_api.cu prep_in_layout
  l.in_bytes = (size_t)N * H * W * 128;
  l.plane_bytes = align_up(l.in_bytes + 2 * l.guard_bytes, 1024);
  l.half_bytes = l.plane_bytes * l.planes;
  return l;
}

// tile shape: bh rows x bw columns, bw * bh <= 128, maximising the share of useful accumulator rows
static void pick_tile(int H, int W, int* bw_out, int* bh_out) {
  double best = -1.0;
  int bbw = W < 128 ? W : 128, bbh = 1;
  for (int bw = (W < 128 ? W : 128); bw >= 8; --bw) {
    int bh = 128 / bw;
    if (bh > H) bh = H;
    if (bh > 256) bh = 256;
    const double tiles = (double)ceil_div(W, bw) * ceil_div(H, bh);
    const double eff = (double)W * H / (tiles * 128.0);
    if (eff > best + 1e-9) { best = eff; bbw = bw; bbh = bh; }
  }
  *bw_out = bbw; *bh_out = bbh;
}

}  // namespace kgdet

using namespace kgdet;

extern "C" size_t kgdet_conv_split_planes_bytes(int32_t N, int32_t C, int32_t H, int32_t W) {
  if (N <= 0 || C <= 0 || C % 64 || H <= 0 || W <= 0) return 0;
  return 2 * split_layout(N, C, H, W).half_bytes;
}

static int zero_guards(unsigned char* base, const SplitLayout& l, cudaStream_t stream) {
  for (int half = 0; half < 2; ++half) {
    unsigned char* b = base + (size_t)half * l.half_bytes;
    KG_CUDA(cudaMemset2DAsync(b, l.plane_bytes, 0, l.guard_bytes, l.planes, stream));
    KG_CUDA(cudaMemset2DAsync(b + l.guard_bytes + l.in_bytes, l.plane_bytes, 0, l.plane_bytes - l.guard_bytes - l.in_bytes,
                              l.planes, stream));
  }
  return KGDET_OK;
}

extern "C" int kgdet_conv_split_planes_from_nchw(const float* x, void* planes, int32_t N, int32_t C, int32_t H, int32_t W,
                                                 void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  KG_CHECK_ARG(x && planes, "kgdet_conv_split_planes_from_nchw: NULL pointer");
  KG_CHECK_ARG(N > 0 && C > 0 && C % 64 == 0 && H > 0 && W > 0 && N <= 65535, "kgdet_conv_split_planes_from_nchw: need C %% 64 == 0");
  KG_CHECK_ARG(((uintptr_t)planes & 255) == 0, "kgdet_conv_split_planes_from_nchw: buffer must be 256-byte aligned");
  const SplitLayout l = split_layout(N, C, H, W);
  int rc = zero_guards((unsigned char*)planes, l, stream);
  if (rc != KGDET_OK) return rc;
  unsigned char* hi = (unsigned char*)planes + l.guard_bytes;
  dim3 grid(ceil_div(H * W, 32), ceil_div(C, 32), N), block(32, 8);
  nchw_to_split_planes_kernel<<<grid, block, 0, stream>>>(x, hi, hi + l.half_bytes, C, H * W, l.plane_bytes);
  KG_LAUNCH_CHECK("nchw_to_split_planes_kernel");
  return KGDET_OK;
}

extern "C" int kgdet_conv_split_planes_from_rows(const float* rows, void* planes, int32_t N, int32_t C, int32_t H, int32_t W,
                                                 void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  KG_CHECK_ARG(rows && planes, "kgdet_conv_split_planes_from_rows: NULL pointer");
  KG_CHECK_ARG(N > 0 && C > 0 && C % 64 == 0 && H > 0 && W > 0, "kgdet_conv_split_planes_from_rows: need C %% 64 == 0");
  KG_CHECK_ARG(((uintptr_t)planes & 255) == 0 && ((uintptr_t)rows & 15) == 0,
               "kgdet_conv_split_planes_from_rows: buffers must be 256 / 16-byte aligned");
  const SplitLayout l = split_layout(N, C, H, W);
  int rc = zero_guards((unsigned char*)planes, l, stream);
  if (rc != KGDET_OK) return rc;
  unsigned char* hi = (unsigned char*)planes + l.guard_bytes;
  const long long M = (long long)N * H * W, total = M * (C / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  rows_to_split_planes_kernel<<<(int)blocks, 256, 0, stream>>>(rows, hi, hi + l.half_bytes, M, C, l.plane_bytes);
  KG_LAUNCH_CHECK("rows_to_split_planes_kernel");
  return KGDET_OK;
}

extern "C" size_t kgdet_conv_packed_weight_bytes(int32_t Cout, int32_t Cin, int32_t ksize) {
  if (Cout <= 0 || Cin <= 0 || Cin % 64 || ksize <= 0) return 0;
  return (size_t)(Cin / 64) * ksize * ksize * 2 * Cout * 128;
}

extern "C" int kgdet_conv_pack_weight(const float* weight, void* packed, int32_t Cout, int32_t Cin, int32_t ksize,
                                      void* stream) {
  KG_CHECK_ARG(weight && packed, "kgdet_conv_pack_weight: NULL pointer");
  KG_CHECK_ARG(Cout > 0 && Cout % 8 == 0 && Cin > 0 && Cin % 64 == 0 && ksize > 0 && (ksize & 1),
               "kgdet_conv_pack_weight: need Cout %% 8 == 0, Cin %% 64 == 0, odd kernel size");
  const long long total = (long long)(Cin / 64) * ksize * ksize * Cout * 8;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  conv_pack_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(weight, (unsigned char*)packed, Cout, Cin, ksize * ksize);
  KG_LAUNCH_CHECK("conv_pack_kernel");
  return KGDET_OK;
}

extern "C" int kgdet_conv_supported(int32_t C, int32_t Cout, int32_t ksize) {
  return (C > 0 && C % 64 == 0 && Cout >= 64 && Cout <= 256 && Cout % 64 == 0 && ksize >= 1 && ksize <= 7 && (ksize & 1)) ? 1 : 0;
}

static int conv_launch(const char* what, int nprob, const void* const* planes, const void* const* weight_packed,
                       const float* const* bias, float* const* out_nhwc, int32_t N, int32_t C, int32_t H, int32_t W,
                       int32_t Cout, int32_t ksize, int fuse_relu, cudaStream_t stream) {
  KG_CHECK_ARG(N > 0 && H > 0 && W > 0, "%s: bad sizes", what);
  KG_CHECK_ARG(kgdet_conv_supported(C, Cout, ksize), "%s: need C %% 64 == 0, Cout in {64, 128, 192, 256}, odd "
               "kernel size <= 7 (got C %d, Cout %d, k %d)", what, C, Cout, ksize);
  for (int t = 0; t < nprob; ++t) {
    KG_CHECK_ARG(planes[t] && weight_packed[t] && out_nhwc[t], "%s: NULL pointer", what);
    KG_CHECK_ARG(((uintptr_t)planes[t] & 255) == 0 && ((uintptr_t)out_nhwc[t] & 15) == 0 &&
                     (!bias[t] || ((uintptr_t)bias[t] & 15) == 0),
                 "%s: planes must be 256-byte aligned, out / bias 16-byte aligned", what);
  }
  EncodeTiledFn enc = encode_fn();
  if (!enc) {
    set_error("%s: cuTensorMapEncodeTiled is not available from this driver", what);
    return KGDET_ERR_UNSUPPORTED;
  }
  const SplitLayout l = split_layout(N, C, H, W);
  ConvParams p;
  p.nprob = nprob;
  for (int t = 0; t < 2; ++t) {
    const int u = t < nprob ? t : 0;
    p.wp[t] = (const unsigned char*)weight_packed[u]; p.bias[t] = bias[u]; p.out[t] = out_nhwc[u];
  }
  p.N = N; p.H = H; p.W = W; p.Cout = Cout; p.ksz = ksize; p.taps = ksize * ksize; p.pad = ksize / 2; p.ncb = C / 64;
  pick_tile(H, W, &p.bw, &p.bh);
  p.tiles_w = ceil_div(W, p.bw); p.tiles_h = ceil_div(H, p.bh);
  p.ntiles = N * p.tiles_w * p.tiles_h;
  p.relu = fuse_relu ? 1 : 0;
  p.idesc = make_idesc(1u, 2 * CV_BM, (uint32_t)Cout);
  p.tmem_cols = Cout <= 64 ? 64 : (Cout <= 128 ? 128 : 256);
  CUtensorMap maps[4];
  for (int t = 0; t < 2; ++t) {
    const int u = t < nprob ? t : 0;
    for (int half = 0; half < 2; ++half) {
      void* base = (unsigned char*)planes[u] + (size_t)half * l.half_bytes + l.guard_bytes;
      const cuuint64_t dims[5] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N, (cuuint64_t)l.planes};
      const cuuint64_t strides[4] = {128, (cuuint64_t)W * 128, (cuuint64_t)H * W * 128, (cuuint64_t)l.plane_bytes};
      const cuuint32_t box[5] = {64, (cuuint32_t)p.bw, (cuuint32_t)p.bh, 1, 1};
      const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
      const CUresult r = enc(&maps[2 * t + half], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        set_error("%s: cuTensorMapEncodeTiled failed (%d) for [%d, %d, %d, %d] box %d x %d", what, (int)r, N, C, H, W,
                  p.bw, p.bh);
        return KGDET_ERR_CUDA;
      }
    }
  }
  p.timeline = nullptr;
  if (g_timeline && g_timeline_entries >= (long long)2 * ceil_div(p.ntiles, 2) * 8) p.timeline = g_timeline;
  g_timeline = nullptr;
  const int b_half = (Cout / 2) * 128;
  const size_t smem = 1024 + (size_t)CV_NS * (2 * CV_A_BYTES + 2 * b_half) + (2 * CV_NS + 2) * 8 + 16;
  KG_CUDA(cudaFuncSetAttribute(conv_umma_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = 2 * ceil_div(p.ntiles, 2);
  conv_umma_pair_kernel<<<grid, CV_THREADS, smem, stream>>>(maps[0], maps[1], maps[2], maps[3], p);
  KG_LAUNCH_CHECK("conv_umma_pair_kernel");
  return KGDET_OK;
}

extern "C" int kgdet_conv_forward(const void* planes, const void* weight_packed, const float* bias, float* out_nhwc,
                                  int32_t N, int32_t C, int32_t H, int32_t W, int32_t Cout, int32_t ksize,
                                  int fuse_relu, void* stream_) {
  return conv_launch("kgdet_conv_forward", 1, &planes, &weight_packed, &bias, &out_nhwc, N, C, H, W, Cout, ksize, fuse_relu,
                     (cudaStream_t)stream_);
}

// Two convolutions of the same geometry in ONE launch (the classification and the point tower's convolution of a
// layer, KP3:415-420: different inputs, different weights): every CTA runs its tile of the first, then the same
// tile of the second into the other half of TMEM, so the epilogue of the first (TMEM -> 128 KB of NHWC fp32 per
// CTA) and the prologue of the second disappear under main loops.  Needs Cout <= 256 twice in TMEM: Cout <= 256.
extern "C" int kgdet_conv_forward_pair(const void* planes0, const void* weight_packed0, const float* bias0, float* out0,
                                       const void* planes1, const void* weight_packed1, const float* bias1, float* out1,
                                       int32_t N, int32_t C, int32_t H, int32_t W, int32_t Cout, int32_t ksize,
                                       int fuse_relu, void* stream_) {
  const void* planes[2] = {planes0, planes1};
  const void* weights[2] = {weight_packed0, weight_packed1};
  const float* bias[2] = {bias0, bias1};
  float* outs[2] = {out0, out1};
  return conv_launch("kgdet_conv_forward_pair", 2, planes, weights, bias, outs, N, C, H, W, Cout, ksize, fuse_relu,
                     (cudaStream_t)stream_);
}
