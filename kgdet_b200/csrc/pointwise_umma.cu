// Pointwise (1x1) convolutions of the Kp3RepBlock as one tcgen05 GEMM per branch (SURVEY.md section 8(f)
// rank 2): cls_out / keypts_out / reppts_out + the cascade's residual adds
// (reppoints_head_kp3rep_cas_1_assign_once.py:79-96,152-171,431-432,440-441).
//
//   out[n, o, pos] = sum_k A[n*HW + pos, k] * W[o, k] + bias[o] (+ residual[n, o, pos])
//
// Both operands arrive PRE-TILED in the layout tcgen05.mma reads, so the main loop is two cp.async.bulk copies
// per k-block issued by one thread (no tensor map, no register staging):
//   A  "UMMA-tiled rows": [M/128 tiles][Ka/64 k-blocks][128 rows x 128 B, 16-byte chunk c of row r at chunk
//      c ^ (r & 7)] -- written directly by the fused DCN kernel's epilogue (KGDET_LAYOUT_TILED*) or by
//      kgdet_nchw_to_tiled_bf16 (stage 1, after a cuDNN 3x3 convolution);
//   W  [Nout/256 tiles][Kw/64 k-blocks][256 rows x 128 B, same swizzle] -- packed once per weight version
//      (kgdet_pointwise_pack_weight).
// Split precision ("bf16x3", fp32-grade): A holds [hi | lo] (Ka = 2K), W holds [W_hi | W_hi | W_lo] (Kw = 3K; the
// middle copy is unused since the operands of a channel block are staged once).  A pipeline stage is one
// 64-channel block c: {A_hi(c), A_lo(c), W_hi(c), W_lo(c)} are copied ONCE (96 KB at 256 columns) and feed the
// twelve MMAs A_hi W_hi + A_lo W_hi + A_hi W_lo -- the GEMM is bound by L2->SM traffic (ncu: 6 TB/s aggregate at
// 22 % tensor-pipe activity when every product re-read both operands), so bytes per MMA is what matters.
//
// CTA = 128 rows x BN columns, BN = 256, or 64 when Nout <= 64 (cls_out has 13 columns: a 256-wide tile would
// stream 4x the weight bytes and issue 4x the MMA columns for nothing): warp 0 bulk-copy producer, warp 1 MMA
// issuer, warps 2..9 epilogue (TMEM -> registers -> bias + residual -> NCHW fp32, coalesced over positions,
// 8 residual loads in flight per thread).
#include <cuda_bf16.h>

#include "dcn.cuh"

namespace kgdet {

static constexpr int PW_BM = 128, PW_BN_MAX = 256;
static constexpr int PW_A_BYTES = PW_BM * 128;
static constexpr int PW_EPI_WARPS = 8;
// column tile: narrow outputs get a narrow tile (the packing and the GEMM derive it from Nout the same way)
__host__ __device__ constexpr int pw_bn(int nout) { return nout <= 64 ? 64 : PW_BN_MAX; }
// stages: as many as fit ~200 KB (split: 2 x 96 KB at BN = 256, 4 x 48 KB at BN = 64; plain: 4 x 48 KB / 8 x 24 KB)
static int pw_stages(int stage_bytes) {
  int ns = (200 * 1024) / stage_bytes;
  return ns > 6 ? 6 : (ns < 2 ? 2 : ns);
}
static constexpr int PW_THREADS = (2 + PW_EPI_WARPS) * 32;

struct PointwiseParams {
  const unsigned char* A;    // tiled, Ka/64 slabs per row tile
  const unsigned char* W;    // tiled, Kw/64 slabs per column tile
  const float* bias;
  int M, N, HW;
  int a_kblocks, w_kblocks;  // Ka/64, Kw/64
  int kgroups;               // K/64 channel blocks = pipeline stages' worth of work
  int tiles_m, tiles_n;      // output tiles (128 rows x bn columns); CTAs are persistent over them
  int bn, ns, stage_bytes;   // column tile, pipeline depth, bytes of one stage
  long long* timeline;       // development hook: per CTA kgroups + 8 clock64() stamps, or NULL
  int nseg;
  kgdet_pointwise_segment seg[KGDET_POINTWISE_MAX_SEGMENTS];
  uint32_t idesc;
};

template <bool SPLIT>
__global__ void __launch_bounds__(PW_THREADS, 1) pointwise_umma_kernel(const PointwiseParams prm) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  const int PW_NS = prm.ns, PW_STAGE = prm.stage_bytes, PW_BN = prm.bn, PW_B_BYTES = prm.bn * 128;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)PW_NS * PW_STAGE);
  uint64_t* empty_bar = full_bar + PW_NS;
  uint64_t* tmem_full_bar = empty_bar + PW_NS;        // [2]: accumulator a is complete
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;       // [2]: accumulator a has been drained by the epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nkb = prm.kgroups;
  // Persistent CTAs: tile t = blockIdx.x + i * gridDim.x (i = 0, 1, ...), column tile fastest so that the CTAs
  // working on one row tile at the same time share its A slabs in L2.  Two TMEM accumulators: the epilogue
  // warps drain accumulator i & 1 while the MMA warp already fills the other one for tile i + 1 -- the epilogue
  // (15-17 k clk with residuals) used to be serial with a 20 k-clk main loop.
  const int ntiles = prm.tiles_m * prm.tiles_n;
  // development stamps for the CTA's FIRST tile: [0] entry, [1] set-up done, [2 + j] issuer saw stage j full,
  // [2 + nkb] accumulator ready (warp 2), [3 + nkb] epilogue of warp 2 done
  long long* const tl = prm.timeline ? prm.timeline + (size_t)blockIdx.x * (nkb + 8) : nullptr;
  if (tl && tid == 0) tl[0] = clock64();

  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < PW_NS; ++s) {
        mbar_init(&full_bar[s], 1);      // the producer's expect_tx arrive
        mbar_init(&empty_bar[s], 1);     // one tcgen05.commit
      }
      for (int a2 = 0; a2 < 2; ++a2) {
        mbar_init(&tmem_full_bar[a2], 1);
        mbar_init(&tmem_empty_bar[a2], PW_EPI_WARPS);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, (uint32_t)(2 * PW_BN));
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tl && tid == 0) tl[1] = clock64();

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;                                         // running stage counter across tiles
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int mt = t / prm.tiles_n, nt = t - mt * prm.tiles_n;
        const unsigned char* a_tile = prm.A + (size_t)mt * prm.a_kblocks * PW_A_BYTES;
        const unsigned char* w_tile = prm.W + (size_t)nt * prm.w_kblocks * PW_B_BYTES;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % PW_NS;
          mbar_wait(&empty_bar[s], ((uint32_t)(it / PW_NS) & 1u) ^ 1u);
          unsigned char* dst = smem + (size_t)s * PW_STAGE;
          mbar_arrive_expect_tx(&full_bar[s], (uint32_t)PW_STAGE);
          // stage layout: A_hi [A_lo] W_hi [W_lo]
          bulk_g2s(dst, a_tile + (size_t)kb * PW_A_BYTES, PW_A_BYTES, &full_bar[s]);
          if constexpr (SPLIT) {
            bulk_g2s(dst + PW_A_BYTES, a_tile + (size_t)(nkb + kb) * PW_A_BYTES, PW_A_BYTES, &full_bar[s]);
            bulk_g2s(dst + 2 * PW_A_BYTES, w_tile + (size_t)kb * PW_B_BYTES, PW_B_BYTES, &full_bar[s]);
            bulk_g2s(dst + 2 * PW_A_BYTES + PW_B_BYTES, w_tile + (size_t)(2 * nkb + kb) * PW_B_BYTES, PW_B_BYTES,
                     &full_bar[s]);
          } else {
            bulk_g2s(dst + PW_A_BYTES, w_tile + (size_t)kb * PW_B_BYTES, PW_B_BYTES, &full_bar[s]);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      int it = 0, i = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++i) {
        const int a2 = i & 1;
        const uint32_t acc = tmem_base + (uint32_t)(a2 * PW_BN);
        if (i >= 2) {                                     // accumulator a2 was used by tile i - 2: wait until drained
          mbar_wait(&tmem_empty_bar[a2], (uint32_t)((i >> 1) - 1) & 1u);
          tc_fence_after();
        }
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % PW_NS;
          mbar_wait(&full_bar[s], (uint32_t)(it / PW_NS) & 1u);
          if (tl && i == 0) tl[2 + kb] = clock64();
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + (size_t)s * PW_STAGE);
          const uint64_t adesc = make_sw128_kmajor_desc(a_addr);
          if constexpr (SPLIT) {
            const uint64_t adesc_lo = make_sw128_kmajor_desc(a_addr + PW_A_BYTES);
            const uint64_t bdesc = make_sw128_kmajor_desc(a_addr + 2 * PW_A_BYTES);
            const uint64_t bdesc_lo = make_sw128_kmajor_desc(a_addr + 2 * PW_A_BYTES + PW_B_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_f16(acc, adesc_lo + 2 * k, bdesc + 2 * k, prm.idesc, (kb > 0 || k > 0) ? 1u : 0u);  // A_lo W_hi
              umma_f16(acc, adesc + 2 * k, bdesc_lo + 2 * k, prm.idesc, 1u);                          // A_hi W_lo
              umma_f16(acc, adesc + 2 * k, bdesc + 2 * k, prm.idesc, 1u);                             // A_hi W_hi
            }
          } else {
            const uint64_t bdesc = make_sw128_kmajor_desc(a_addr + PW_A_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(acc, adesc + 2 * k, bdesc + 2 * k, prm.idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          tc_commit(&empty_bar[s]);
        }
        tc_commit(&tmem_full_bar[a2]);
      }
    }
    __syncwarp();
  } else {
   int i = 0;
   for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++i) {
    const int mt = t / prm.tiles_n, nt = t - mt * prm.tiles_n;
    const int a2 = i & 1;
    const uint32_t tmem_acc = tmem_base + (uint32_t)(a2 * PW_BN);
    // ---- epilogue: TMEM -> bias + residual -> NCHW fp32 ----
    mbar_wait(&tmem_full_bar[a2], (uint32_t)(i >> 1) & 1u);
    if (tl && tid == 64 && i == 0) tl[2 + nkb] = clock64();
    tc_fence_after();
    const int q = warp & 3, half = (warp - 2) >> 2;       // TMEM lane quarter (hardware: warp % 4), column half
    const int row = q * 32 + lane;
    const int m = mt * PW_BM + row;
    const int img = m < prm.M ? m / prm.HW : 0, pos = m < prm.M ? m - img * prm.HW : 0;
    for (int c0 = 0; c0 < PW_BN / 2; c0 += 32) {                // PW_BN / 2 is 128 or 32: whole 32-column chunks
      const int col = half * (PW_BN / 2) + c0;
      if (nt * PW_BN + col >= prm.N) break;               // warp-uniform
      uint32_t acc[32];
      tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)col, acc);
      tmem_ld_wait();
      if (m < prm.M) {
        const int cbase = nt * PW_BN + col;               // first output column of this chunk (warp-uniform)
        int si0 = 0;
#pragma unroll
        for (int q2 = 0; q2 + 1 < KGDET_POINTWISE_MAX_SEGMENTS; ++q2)
          si0 += (q2 + 1 < prm.nseg && cbase >= prm.seg[q2].col_end) ? 1 : 0;
        if (cbase + 32 <= prm.N && cbase + 32 <= prm.seg[si0].col_end) {
          // Fast path (all chunks but the ones straddling a segment boundary): one segment, so a column is a
          // constant stride HW from the previous one.  The general path below spends ~30 dependent instructions
          // per element on the segment search and 64-bit index arithmetic, which made the epilogue 34-46 k
          // cycles per CTA against a 20 k-cycle main loop (tools/pointwise_timeline.py).
          const kgdet_pointwise_segment sg = prm.seg[si0];
          const size_t o0 = ((size_t)img * sg.channels_total + sg.channel_offset + (cbase - sg.col_begin)) * prm.HW + pos;
          float* __restrict__ op = sg.out + o0;
          const float* __restrict__ rp = sg.residual ? sg.residual + o0 : nullptr;
          const float* __restrict__ bp = prm.bias ? prm.bias + cbase : nullptr;
          const size_t hw = (size_t)prm.HW;
          if (rp) {
#pragma unroll
            for (int j0 = 0; j0 < 32; j0 += 16) {        // 16 unconditional residual loads in flight per thread
              float r[16];
#pragma unroll
              for (int e = 0; e < 16; ++e) r[e] = __ldg(rp + (size_t)(j0 + e) * hw);
#pragma unroll
              for (int e = 0; e < 16; ++e) acc[j0 + e] = __float_as_uint(__uint_as_float(acc[j0 + e]) + r[e]);
            }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float b = bp ? __ldg(bp + j) : 0.f;                         // warp-uniform address
            op[(size_t)j * hw] = __uint_as_float(acc[j]) + b;                 // lanes = consecutive positions
          }
        } else {
#pragma unroll
          for (int j0 = 0; j0 < 32; j0 += 8) {
            size_t o[8];
            float r[8];
            float* outp[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int c = cbase + j0 + e;
              int si = 0;
#pragma unroll
              for (int q2 = 0; q2 + 1 < KGDET_POINTWISE_MAX_SEGMENTS; ++q2)
                si += (q2 + 1 < prm.nseg && c >= prm.seg[q2].col_end) ? 1 : 0;
              const kgdet_pointwise_segment& sg = prm.seg[si];
              o[e] = ((size_t)img * sg.channels_total + sg.channel_offset + (c - sg.col_begin)) * prm.HW + pos;
              outp[e] = sg.out;
              r[e] = (c < prm.N && sg.residual) ? __ldg(sg.residual + o[e]) : 0.f;
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int c = cbase + j0 + e;
              if (c < prm.N) {
                float v = __uint_as_float(acc[j0 + e]) + r[e];
                if (prm.bias) v += __ldg(prm.bias + c);
                outp[e][o[e]] = v;
              }
            }
          }
        }
      }
    }
    // this warp's share of accumulator a2 is in registers / memory: hand the accumulator back to the MMA warp
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&tmem_empty_bar[a2]);
    if (tl && tid == 64 && i == 0) tl[3 + nkb] = clock64();
   }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)(2 * PW_BN));
}

// byte offset of element (r, k) inside a [rows x 64] bf16 slab in the swizzled K-major layout
__device__ __forceinline__ size_t tiled_off(int r, int k) {
  return (size_t)r * 128 + ((((k >> 3) ^ (r & 7)) << 4)) + (k & 7) * 2;
}

// fp32 [Nout, K] -> tiled bf16 W ([W] or, split, [W_hi | W_hi | W_lo]); rows beyond Nout are zero
__global__ void pointwise_pack_kernel(const float* __restrict__ w, unsigned char* __restrict__ p, int Nout, int K,
                                      int split) {
  const int kw = (split ? 3 : 1) * K, kblocks = kw / 64;
  const int PW_BN = pw_bn(Nout), PW_B_BYTES = PW_BN * 128;
  const int ntiles = (Nout + PW_BN - 1) / PW_BN;
  const long long total = (long long)ntiles * PW_BN * kw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)blockDim.x * gridDim.x) {
    const int o = (int)(i / kw), kk = (int)(i - (long long)o * kw);
    float v = 0.f;
    if (o < Nout) {
      const float x = w[(size_t)o * K + kk % K];
      const float hi = __bfloat162float(__float2bfloat16(x));
      v = (split && kk >= 2 * K) ? x - hi : x;
    }
    const int tile = o / PW_BN, r = o % PW_BN, kb = kk / 64;
    unsigned char* slab = p + ((size_t)tile * kblocks + kb) * PW_B_BYTES;
    *reinterpret_cast<__nv_bfloat16*>(slab + tiled_off(r, kk % 64)) = __float2bfloat16(v);
  }
}

// NCHW fp32/bf16 -> UMMA-tiled bf16 rows (optionally [hi | lo]), optional ReLU; rows beyond M stay untouched
// (they only feed accumulator rows the epilogue never stores)
template <typename Tin>
__global__ void nchw_to_tiled_kernel(const Tin* __restrict__ src, unsigned char* __restrict__ dst, int C, int S,
                                     int relu, int split) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const Tin* s = src + (size_t)n * C * S;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int c = c0 + ty + k, p = p0 + tx;
    if (c < C && p < S) {
      const float v = (float)s[(size_t)c * S + p];
      tile[ty + k][tx] = relu ? fmaxf(v, 0.f) : v;
    }
  }
  __syncthreads();
  const int a_kblocks = (split ? 2 : 1) * (C / 64);
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int p = p0 + ty + k, c = c0 + tx;
    if (c < C && p < S) {
      const int m = n * S + p;
      const float v = tile[tx][ty + k];
      const __nv_bfloat16 hi = __float2bfloat16(v);
      unsigned char* t = dst + (size_t)(m / PW_BM) * a_kblocks * PW_A_BYTES;
      *reinterpret_cast<__nv_bfloat16*>(t + (size_t)(c / 64) * PW_A_BYTES + tiled_off(m % PW_BM, c % 64)) = hi;
      if (split)
        *reinterpret_cast<__nv_bfloat16*>(t + (size_t)((C + c) / 64) * PW_A_BYTES + tiled_off(m % PW_BM, c % 64)) =
            __float2bfloat16(v - __bfloat162float(hi));
    }
  }
}

}  // namespace kgdet

using namespace kgdet;

extern "C" size_t kgdet_pointwise_tiled_bytes(int32_t M, int32_t K, int split) {
  if (M < 0 || K <= 0 || K % 64) return 0;
  return (size_t)ceil_div(M, PW_BM) * ((split ? 2 : 1) * (K / 64)) * PW_A_BYTES;
}

extern "C" size_t kgdet_pointwise_packed_weight_bytes(int32_t Nout, int32_t K, int split) {
  if (Nout <= 0 || K <= 0 || K % 64) return 0;
  return (size_t)ceil_div(Nout, pw_bn(Nout)) * ((split ? 3 : 1) * (K / 64)) * pw_bn(Nout) * 128;
}

extern "C" int kgdet_pointwise_pack_weight(const float* weight, void* packed, int32_t Nout, int32_t K, int split,
                                           void* stream) {
  KG_CHECK_ARG(weight && packed, "kgdet_pointwise_pack_weight: NULL pointer");
  KG_CHECK_ARG(Nout > 0 && K > 0 && K % 64 == 0, "kgdet_pointwise_pack_weight: Nout > 0 and K %% 64 == 0 required");
  KG_CHECK_ARG(split == 0 || split == 1, "kgdet_pointwise_pack_weight: split must be 0 or 1");
  const long long total = (long long)ceil_div(Nout, pw_bn(Nout)) * pw_bn(Nout) * (split ? 3 : 1) * K;
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)num_sms() * 16) blocks = (long long)num_sms() * 16;
  pointwise_pack_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(weight, (unsigned char*)packed, Nout, K, split);
  KG_LAUNCH_CHECK("pointwise_pack_kernel");
  return KGDET_OK;
}

extern "C" int kgdet_nchw_to_tiled_bf16(const void* src, void* dst, int32_t N, int32_t C, int32_t S, int src_dtype,
                                        int fuse_relu, int split, void* stream) {
  KG_CHECK_ARG(N >= 0 && C >= 64 && C % 64 == 0 && S >= 1, "kgdet_nchw_to_tiled_bf16: C %% 64 == 0 required");
  if (N == 0) return KGDET_OK;
  KG_CHECK_ARG(src && dst, "kgdet_nchw_to_tiled_bf16: NULL pointer");
  KG_CHECK_ARG(N <= 65535 && ceil_div(C, 32) <= 65535, "kgdet_nchw_to_tiled_bf16: batch/channels too large");
  dim3 grid(ceil_div(S, 32), ceil_div(C, 32), N), block(32, 8);
  if (src_dtype == KGDET_F32)
    nchw_to_tiled_kernel<float><<<grid, block, 0, (cudaStream_t)stream>>>((const float*)src, (unsigned char*)dst, C, S, fuse_relu, split);
  else if (src_dtype == KGDET_BF16)
    nchw_to_tiled_kernel<__nv_bfloat16><<<grid, block, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)src, (unsigned char*)dst, C, S, fuse_relu, split);
  else {
    set_error("kgdet_nchw_to_tiled_bf16: bad dtype %d", src_dtype);
    return KGDET_ERR_INVALID_ARG;
  }
  KG_LAUNCH_CHECK("nchw_to_tiled_kernel");
  return KGDET_OK;
}

extern "C" int kgdet_pointwise_conv_tiled(const void* a_tiled, const void* w_packed, const float* bias, int32_t M,
                                          int32_t K, int32_t Nout, int32_t HW, int split,
                                          const kgdet_pointwise_segment* segs, int32_t nseg, void* stream) {
  KG_CHECK_ARG(a_tiled && w_packed && segs, "kgdet_pointwise_conv_tiled: NULL pointer");
  KG_CHECK_ARG(M >= 0 && K > 0 && K % 64 == 0 && Nout > 0 && HW > 0, "kgdet_pointwise_conv_tiled: bad sizes (K %% 64 == 0)");
  KG_CHECK_ARG(split == 0 || split == 1, "kgdet_pointwise_conv_tiled: split must be 0 or 1");
  KG_CHECK_ARG(M % HW == 0, "kgdet_pointwise_conv_tiled: M (%d) must be a multiple of HW (%d)", M, HW);
  KG_CHECK_ARG(nseg >= 1 && nseg <= KGDET_POINTWISE_MAX_SEGMENTS, "kgdet_pointwise_conv_tiled: 1..%d segments",
               KGDET_POINTWISE_MAX_SEGMENTS);
  int expect = 0;
  for (int i = 0; i < nseg; ++i) {
    KG_CHECK_ARG(segs[i].out && segs[i].col_begin == expect && segs[i].col_end > segs[i].col_begin &&
                     segs[i].channel_offset >= 0 &&
                     segs[i].channel_offset + (segs[i].col_end - segs[i].col_begin) <= segs[i].channels_total,
                 "kgdet_pointwise_conv_tiled: segment %d is not contiguous / does not fit its tensor", i);
    expect = segs[i].col_end;
  }
  KG_CHECK_ARG(expect == Nout, "kgdet_pointwise_conv_tiled: segments cover %d of %d columns", expect, Nout);
  if (M == 0) return KGDET_OK;
  PointwiseParams p;
  p.A = (const unsigned char*)a_tiled; p.W = (const unsigned char*)w_packed; p.bias = bias;
  p.M = M; p.N = Nout; p.HW = HW;
  p.a_kblocks = (split ? 2 : 1) * (K / 64);
  p.w_kblocks = (split ? 3 : 1) * (K / 64);
  p.nseg = nseg;
  for (int i = 0; i < KGDET_POINTWISE_MAX_SEGMENTS; ++i) p.seg[i] = segs[i < nseg ? i : nseg - 1];
  p.kgroups = K / 64;
  p.bn = pw_bn(Nout);
  p.stage_bytes = (split ? 2 : 1) * (PW_A_BYTES + p.bn * 128);
  p.ns = pw_stages(p.stage_bytes);
  p.idesc = make_idesc(1u, PW_BM, (uint32_t)p.bn);
  p.timeline = nullptr;
  p.tiles_m = ceil_div(M, PW_BM);
  p.tiles_n = ceil_div(Nout, p.bn);
  if (g_timeline && g_timeline_entries >= (long long)p.tiles_m * p.tiles_n * (p.kgroups + 8)) p.timeline = g_timeline;
  g_timeline = nullptr;
  const size_t smem = 1024 + (size_t)p.ns * p.stage_bytes + (2 * p.ns + 4) * 8 + 16;
  auto kern = split ? pointwise_umma_kernel<true> : pointwise_umma_kernel<false>;
  KG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int ntiles = p.tiles_m * p.tiles_n;
  const int grid = ntiles < num_sms() ? ntiles : num_sms();      // one persistent CTA per SM
  kern<<<grid, PW_THREADS, smem, (cudaStream_t)stream>>>(p);
  KG_LAUNCH_CHECK("pointwise_umma_kernel");
  return KGDET_OK;
}
