// Fused deformable-convolution forward on the 5th-generation tensor cores (sm_100a).
//
//   out[m, o] = sum_{tap, c} S(m, tap, c) * W[o, c, tap]       m = (n, y, x) output position
//
// The reference materialises S as the `columns` tensor in HBM (deformable_im2col,
// deform_conv_cuda_kernel.cu:189-242: C*K x N*H*W fp32, 843 MB for one 7x7 KGDet call) and hands it to a
// cuBLAS SGEMM (deform_conv_cuda.cpp:230-233).  Here the column tile only ever exists in shared memory.
//
// One CTA = 128 output positions x all Cout, 9 warps, NS-stage mbarrier pipeline over k-blocks
// (k-block = 64 bf16 / 32 tf32 channels of one tap = one 128-byte slab per sampled pixel):
//   warps 0..7   producers.  Thread t owns 16-byte chunk (t & 7) of rows (t >> 3) + 32 i, i < 4: 8 lanes
//                cover one pixel slab, so every gather instruction reads four whole 128-byte lines of the
//                channel-blocked input (dcn_api.cu).  Per k-block a thread interpolates its four row-chunks
//                (packed HFMA2.BF16 with the plan's pre-rounded corner weights), stores them into the
//                128B-swizzled K-major A tile and immediately re-arms the same registers with the gathers of
//                the NEXT k-block (all four corners unconditionally; unusable corners carry weight 0 and a
//                guard-band-safe address), then fence.proxy.async + one mbarrier arrive per warp.
//   warp 8       control (one elected lane): streams the pre-swizzled weight slabs with cp.async.bulk
//                (UBLKCP; a linear copy lands in UMMA layout, no tensor map) NS-1 k-blocks ahead, waits for a
//                stage to be full, issues the tcgen05.mma's (UTCHMMA, M128 x N=Cout x K16/8) into the TMEM
//                accumulator and commits the stage's empty barrier.
//   Epilogue     warps 0..7: tcgen05.ld, bias, ReLU, then either an NCHW store coalesced over positions or
//                bf16 rows in the UMMA-tiled layout of the pointwise GEMM that follows (pointwise_umma.cu).
//
// Modes: BF16   kind::f16, bf16 operands                     (1e-3 grade)
//        TF32X3 kind::tf32, A = Ahi + Alo, B = Bhi + Blo,    (the dropped term Alo.Blo is ~2^-22)
//               3 MMAs per k-step: Alo.Bhi + Ahi.Blo + Ahi.Bhi
//        TF32   kind::tf32 single pass
//
// PAIR = true runs clusters of two CTAs with 2-SM MMAs (cta_group::2, M = 256): each CTA gathers its own
// 128 rows but holds only half of the weight slab, halving the weight bytes an SM pulls through L2.  The
// even CTA issues all MMAs; its full barrier collects the 16 local producer warps, the 16 remote ones
// (mbarrier.arrive on the mapa'd address), its own weight bytes and one relay arrive from the odd CTA's
// control warp once that CTA's half has landed; tcgen05.commit multicasts the empty / accumulator-ready
// arrivals to both CTAs.  Parity-tested; not the default because it is not faster (see dcn_umma.cu).
#include <cuda_bf16.h>

#include "dcn_umma.cuh"

namespace kgdet {

// RPT = rows (of the 128-row tile) per producer thread: 2 -> 16 producer warps (96 registers per thread with
// the control warp), 4 -> 8 producer warps (168 registers).  The producers are bound by instruction issue
// (ablation on B200: with gathers, record loads, fence, A-tile stores, MMAs and weight copies ALL removed the
// K = 49 call still takes 94 of its 144 us), so fewer, fatter threads -- half the per-thread loop / barrier /
// address overhead per row -- win.
template <int RPT> struct Producers {
  static constexpr int WARPS = 32 / RPT;                   // 128 rows x 8 chunks / RPT / 32 lanes
  static constexpr int THREADS = (WARPS + 1) * 32;         // + the control warp
  static constexpr int ROW_STEP = 128 / RPT;               // rows of one thread are ROW_STEP apart
};
// Warp layout.  SPREAD = false: producers are warps 0 .. WARPS-1, the control warp comes last (9 warps at RPT = 4).
// SPREAD = true (RPT = 4 only): 12 warps, the control warp is warp 0 and NO producer shares its scheduler
// (sub-partition = warp id % 4): producers are warps {1,2,3, 5,6,7, 9,10}, warps 4, 8, 11 only help with the
// epilogue.  Measured with the timeline hook (tools/dcn_timeline.py, warp_arrive_before_full): in the compact
// layout the two producer warps that share the control warp's sub-partition (0 and 4) arrive ~1 000 clk after the
// other six at every k-block and set the pace; in the spread layout the sub-partition with two producers runs
// ahead and the two with three producers set the same pace (1 110 vs 1 155 clk per k-block).  I.e. the period is a
// per-sub-partition limit of ~2.5 producer warps' worth of work -- consistent with the 32 B/clk register
// write-back port of a sub-partition (a 512-byte LDG.128 per warp occupies it for 16 clk: 20 such loads per warp
// and k-block) -- not a whole-SM L1 or tensor-pipe limit.
template <int RPT, bool SPREAD> struct Layout {
  static constexpr int THREADS = SPREAD ? 384 : Producers<RPT>::THREADS;
  static constexpr int CONTROL_WARP = SPREAD ? 0 : Producers<RPT>::WARPS;
  static constexpr int EPI_GROUPS = SPREAD ? 3 : Producers<RPT>::WARPS / 4;    // warps per TMEM lane quarter
  __device__ static __forceinline__ int producer_slot(int warp) {             // -1: not a producer
    if constexpr (!SPREAD) return warp < Producers<RPT>::WARPS ? warp : -1;
    else return ((warp & 3) != 0 && warp != 11) ? (warp >> 2) * 3 + (warp & 3) - 1 : -1;
  }
};

// bounded wait without the diagnostic printf of mbar_wait (keeps the hot loop small): a protocol bug
// still traps instead of hanging the GPU
__device__ __forceinline__ void mbar_spin(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 22)) __trap();
  }
}
__device__ __forceinline__ void mbar_spin_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (++spins > (1u << 22)) __trap();
  }
}

// shared -> global bulk store (async proxy); the caller commits the group and waits for the reads
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}

__device__ __forceinline__ __nv_bfloat162 as_bf162(uint32_t v) {
  return *reinterpret_cast<__nv_bfloat162*>(&v);
}
__device__ __forceinline__ uint32_t as_u32(__nv_bfloat162 v) { return *reinterpret_cast<uint32_t*>(&v); }

// one 16-byte chunk of one row: sum of the four corners times their weights
//   bf16: the record carries bf16x2 (w0,w1) and (w2,w3); nvcc folds the broadcasts into the operand
//         selectors of HMUL2/HFMA2.BF16_V2 (no PRMT); same rounding sequence as the two-group kernel
template <int MODE>
__device__ __forceinline__ void combine_store(const uint4 (&v)[4], uint32_t wy, uint32_t wz, uint32_t ww,
                                              unsigned char* dst) {
  if constexpr (MODE == MODE_BF16) {
    const __nv_bfloat162 w01 = as_bf162(wy), w23 = as_bf162(wz);
    const __nv_bfloat162 w0 = __low2bfloat162(w01), w1 = __high2bfloat162(w01);
    const __nv_bfloat162 w2 = __low2bfloat162(w23), w3 = __high2bfloat162(w23);
    uint4 o;
    o.x = as_u32(__hfma2(w3, as_bf162(v[3].x), __hfma2(w2, as_bf162(v[2].x), __hfma2(w1, as_bf162(v[1].x), __hmul2(w0, as_bf162(v[0].x))))));
    o.y = as_u32(__hfma2(w3, as_bf162(v[3].y), __hfma2(w2, as_bf162(v[2].y), __hfma2(w1, as_bf162(v[1].y), __hmul2(w0, as_bf162(v[0].y))))));
    o.z = as_u32(__hfma2(w3, as_bf162(v[3].z), __hfma2(w2, as_bf162(v[2].z), __hfma2(w1, as_bf162(v[1].z), __hmul2(w0, as_bf162(v[0].z))))));
    o.w = as_u32(__hfma2(w3, as_bf162(v[3].w), __hfma2(w2, as_bf162(v[2].w), __hfma2(w1, as_bf162(v[1].w), __hmul2(w0, as_bf162(v[0].w))))));
    *reinterpret_cast<uint4*>(dst) = o;
  } else {
    float w[4];
    decode_rec(make_float4(0.f, __uint_as_float(wy), __uint_as_float(wz), __uint_as_float(ww)), w);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      a0 = fmaf(w[i], __uint_as_float(v[i].x), a0);
      a1 = fmaf(w[i], __uint_as_float(v[i].y), a1);
      a2 = fmaf(w[i], __uint_as_float(v[i].z), a2);
      a3 = fmaf(w[i], __uint_as_float(v[i].w), a3);
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

// What bounds it (ablations on B200, K = 49 KGDet call, 16 warps x 2 rows, 144 us):
//   no gathers at all 143 us | no record loads 145 | no fence.proxy.async 145 | no A-tile stores 141 |
//   no weight copies 135 | no MMAs 119 | ALL of those removed 94 us.
// I.e. memory is fully hidden; two thirds of the time is the instruction / barrier skeleton of the producers
// (issue-bound: ~110 SASS instructions per thread and k-block on 4-5 warps per scheduler) and the rest is the
// tensor core's operand reads competing with the producers for the SM's shared-memory bandwidth.  Hence:
//   * 4 rows per thread (8 producer warps, 160 registers): half the per-thread loop overhead per row: 137 us;
//     8 rows per thread (4 warps): 161 us -- too few warps to cover instruction latency.
//   * built, measured and dropped: DEPTH = 2 with 16 warps at 128 registers and the control duty rotating over
//     the producer warps (17 warps cap a thread at 96 registers) 171 us; all gathers through ld.global.cg
//     166 us; 148 balanced 114-row tiles instead of 132 128-row tiles 150 us; CTA pairs 152 us.
//   * built, measured and dropped (round 1, last day): a "dual" launch that runs the 3x3 and the 5x5 point set of a
//     branch through one pipeline into two TMEM accumulators (combined plan + interleaved packed weights, the
//     producers untouched).  Bit-identical to two launches, but 119.4 us against 38.0 + 79.3 us for the pair, and
//     the extra control-lane / epilogue code made the UNCHANGED single-convolution path 7 % slower (K = 49:
//     141 -> 151 us): this hot loop is sensitive to code layout, not only to instruction count.
//   * what did help earlier: channel-blocked input planes (contiguous 128-byte slabs are served 1.4x faster by
//     the L1 than slabs 512 bytes apart, tools/micro/l1_gather_bench) and tap-major plan records.
template <int MODE, int NS, int DEPTH, typename Tout, bool PAIR, int RPT, bool SPREAD, bool TL>
__global__ void __launch_bounds__(Layout<RPT, SPREAD>::THREADS, 1) dcn_umma_stream_kernel(const UmmaParams prm) {
  using MT = ModeTraits<MODE>;
  using LY = Layout<RPT, SPREAD>;
  static_assert(!SPREAD || (RPT == 4 && !PAIR), "spread layout: 8 producer warps, single CTA");
  constexpr int PRODUCER_WARPS = Producers<RPT>::WARPS;
  constexpr int ROW_STEP = Producers<RPT>::ROW_STEP;
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  // 1024-byte alignment of every tile is what the 128B swizzle pattern is anchored to
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  const int BN = prm.Cout;
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;      // 0 = the CTA that issues the MMAs
  const int b_tile_bytes = (PAIR ? BN / 2 : BN) * 128;          // weight rows held by THIS CTA
  const int stage_bytes = MT::A_TILES * A_TILE_BYTES + MT::B_TILES * b_tile_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)NS * stage_bytes);
  uint64_t* empty_bar = full_bar + NS;
  uint64_t* tmem_full_bar = empty_bar + NS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pslot = LY::producer_slot(warp);
  const int tid = pslot >= 0 ? pslot * 32 + lane : -1;     // producer thread index (row / chunk ownership)
  const int m0 = blockIdx.x * BM;
  // k-block range of this CTA: everything, or one split of it (prm.partial != NULL)
  const int kb_lo = prm.partial ? (int)blockIdx.y * prm.kb_per_split : 0;
  const int nkb = prm.partial ? min(prm.nkb - kb_lo, prm.kb_per_split) : prm.nkb;
  constexpr int SETUP_WARP = LY::CONTROL_WARP;
  // TL = instrumented instantiation (kgdet_dcn_set_timeline); the production kernel carries none of the stamps
  long long* const tl = (TL && prm.timeline && !prm.partial) ? prm.timeline + (size_t)blockIdx.x * (12 * nkb + 8) : nullptr;
  if (tl && threadIdx.x == 0) tl[0] = clock64();

  if (warp == SETUP_WARP) {
    if (lane == 0) {
      for (int s = 0; s < NS; ++s) {
        // 16 producer warps + the control lane's expect_tx; PAIR, even CTA: + 16 remote warps + the odd
        // CTA's relay; PAIR, odd CTA: only its own weight half (expect_tx) completes on this barrier
        mbar_init(&full_bar[s], PAIR ? (cta_rank == 0 ? 2 * PRODUCER_WARPS + 2 : 1) : PRODUCER_WARPS + 1);
        mbar_init(&empty_bar[s], 1);                // one tcgen05.commit
      }
      mbar_init(tmem_full_bar, 1);
      fence_mbar_init();
    }
    __syncwarp();
    if constexpr (PAIR) tmem_alloc_pair(tmem_slot, prm.tmem_cols);
    else tmem_alloc(tmem_slot, prm.tmem_cols);
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();   // the peer's barriers must exist before any remote arrive
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tl && threadIdx.x == 0) tl[1] = clock64();
  const uint32_t full_bar0_remote = PAIR ? mapa_u32(smem_u32(&full_bar[0]), 0u) : 0u;

  // ------------------------- control duties (executed by ONE lane) -------------------------
  const uint32_t b_bytes = (uint32_t)(MT::B_TILES * b_tile_bytes);
  auto fetch_b = [&](int kq, int sq) {
    const size_t b_full_tile = (size_t)BN * 128;                          // one packed tile, all Cout rows
    // the packed layout always carries hi+lo for tf32; single-pass TF32 copies only hi
    const size_t b_src_stride = (size_t)(MODE == MODE_BF16 ? 1 : 2) * b_full_tile;
    const unsigned char* src = prm.wp + (size_t)cta_rank * b_tile_bytes + (size_t)(kb_lo + kq) * b_src_stride;
    unsigned char* dstb = smem + (size_t)sq * stage_bytes + MT::A_TILES * A_TILE_BYTES;
    mbar_arrive_expect_tx(&full_bar[sq], b_bytes);
    if (!PAIR || MT::B_TILES == 1) {
      bulk_g2s(dstb, src, b_bytes, &full_bar[sq]);
    } else {                                  // hi and lo tiles are Cout rows apart in the packed layout
      bulk_g2s(dstb, src, (uint32_t)b_tile_bytes, &full_bar[sq]);
      bulk_g2s(dstb + b_tile_bytes, src + b_full_tile, (uint32_t)b_tile_bytes, &full_bar[sq]);
    }
  };
  // k-block j: wait until its stage is full, issue its MMAs (even CTA) or relay "my weight half landed"
  // (odd CTA of a pair), then fetch the weight slab of k-block j + NS - 1 into the stage of k-block j - 1
  auto control_duty = [&](int j) {
    const int s = j % NS;
    const uint32_t ph = (uint32_t)(j / NS) & 1u;
    if (PAIR && cta_rank != 0) {
      mbar_spin(&full_bar[s], ph);
      mbar_arrive_remote(full_bar0_remote + (uint32_t)s * 8u);
    } else {
      mbar_spin(&full_bar[s], ph);     // A tile(s) from all producer warps + weight bytes (both CTAs of a pair)
      if (tl) tl[2 + j] = clock64();
      tc_fence_after();
      const uint32_t a_addr = smem_u32(smem + (size_t)s * stage_bytes);
      const uint32_t b_addr = a_addr + MT::A_TILES * A_TILE_BYTES;
      const uint64_t adesc = make_sw128_kmajor_desc(a_addr);
      const uint64_t bdesc = make_sw128_kmajor_desc(b_addr);
#pragma unroll
      for (int k = 0; k < 4; ++k) {            // 4 x 32 bytes of K per 128-byte row
        const uint32_t acc = (j > 0 || k > 0) ? 1u : 0u;
        if constexpr (MODE == MODE_BF16) {
          if constexpr (PAIR) umma_f16_pair(tmem_base, adesc + 2 * k, bdesc + 2 * k, prm.idesc, acc);
          else umma_f16(tmem_base, adesc + 2 * k, bdesc + 2 * k, prm.idesc, acc);
        } else if constexpr (MODE == MODE_TF32) {
          if constexpr (PAIR) umma_tf32_pair(tmem_base, adesc + 2 * k, bdesc + 2 * k, prm.idesc, acc);
          else umma_tf32(tmem_base, adesc + 2 * k, bdesc + 2 * k, prm.idesc, acc);
        } else {
          const uint64_t adesc_lo = make_sw128_kmajor_desc(a_addr + A_TILE_BYTES);
          const uint64_t bdesc_lo = make_sw128_kmajor_desc(b_addr + b_tile_bytes);
          if constexpr (PAIR) {
            umma_tf32_pair(tmem_base, adesc_lo + 2 * k, bdesc + 2 * k, prm.idesc, acc);   // Alo.Bhi
            umma_tf32_pair(tmem_base, adesc + 2 * k, bdesc_lo + 2 * k, prm.idesc, 1u);    // Ahi.Blo
            umma_tf32_pair(tmem_base, adesc + 2 * k, bdesc + 2 * k, prm.idesc, 1u);       // Ahi.Bhi
          } else {
            umma_tf32(tmem_base, adesc_lo + 2 * k, bdesc + 2 * k, prm.idesc, acc);     // Alo.Bhi
            umma_tf32(tmem_base, adesc + 2 * k, bdesc_lo + 2 * k, prm.idesc, 1u);      // Ahi.Blo
            umma_tf32(tmem_base, adesc + 2 * k, bdesc + 2 * k, prm.idesc, 1u);         // Ahi.Bhi
          }
        }
      }
      // frees the stage (in both CTAs of a pair) when these MMAs retire
      if constexpr (PAIR) tc_commit_pair(&empty_bar[s], (uint16_t)3);
      else tc_commit(&empty_bar[s]);
      if (j == nkb - 1) {                          // all MMAs have retired -> accumulator readable
        if constexpr (PAIR) tc_commit_pair(tmem_full_bar, (uint16_t)3);
        else tc_commit(tmem_full_bar);
      }
    }
    // weight slab of k-block j + NS - 1 goes into the stage of k-block j - 1, whose MMAs were issued one duty
    // ago: this wait ends while MMA(j) is still running, so the tensor pipe never idles
    const int kn = j + NS - 1;
    if (kn < nkb) {
      if (j >= 1) mbar_spin(&empty_bar[kn % NS], (uint32_t)((j - 1) / NS) & 1u);
      fetch_b(kn, kn % NS);
    }
  };

  if (warp == LY::CONTROL_WARP) {
    // =========================== dedicated control warp ===========================
    if (lane == 0) {
      for (int j = 0; j < NS - 1 && j < nkb; ++j) fetch_b(j, j);   // all stages start empty
      for (int kb = 0; kb < nkb; ++kb) control_duty(kb);
    }
    __syncwarp();
  } else if (pslot >= 0) {
    // ======================================= producers =======================================
    const int chunk = tid & 7;        // 16-byte chunk of the 128-byte row
    const int rbase = tid >> 3;       // this thread owns rows rbase + i * ROW_STEP, i < RPT
    const int K = prm.K;
    constexpr long long rowb = 128;                       // bytes per pixel slab (one plane = one channel block)
    const long long wrow = (long long)prm.W * rowb;       // bytes per image row
    const unsigned char* in_base = reinterpret_cast<const unsigned char*>(prm.in) + chunk * 16;
    // plan is tap-major [K][rows_padded]: the four rows a warp touches per load are 64 contiguous bytes
    const uint4* plan0 = reinterpret_cast<const uint4*>(prm.plan) + (m0 + rbase);
    const size_t tap_stride = (size_t)prm.rows_padded;
    // this thread's byte offset inside an A tile (row rbase, swizzled 16-byte chunk); its other rows are
    // ROW_STEP * 128 bytes further each and have the same (row & 7)
    const int a_off = rbase * 128 + ((chunk ^ (rbase & 7)) << 4);

    uint4 v[DEPTH][RPT][4];           // corners in flight: [k-block slot][row][corner]
    uint32_t wy[DEPTH][RPT], wz[DEPTH][RPT], ww[DEPTH][RPT];
    uint4 recn[RPT];                  // records of the next k-block to issue
    auto load_recs = [&](int tap) {
#pragma unroll
      for (int i = 0; i < RPT; ++i) recn[i] = __ldg(plan0 + (size_t)i * ROW_STEP + tap * tap_stride);
    };

    // (tap, channel block) of the k-block whose gathers are issued next, and tap of the next record fetch
    int tapI = kb_lo % K, cbI = kb_lo / K, tapR = 0;
    const unsigned char* in_plane = in_base + (size_t)cbI * prm.plane_bytes;   // plane of the k-block whose gathers are issued next
    auto issue = [&](int slot, int row, const uint4& rec) {
      const unsigned char* p0 = in_plane + (long long)(int)rec.x * rowb;
      v[slot][row][0] = __ldg(reinterpret_cast<const uint4*>(p0));
      v[slot][row][1] = __ldg(reinterpret_cast<const uint4*>(p0 + rowb));
      v[slot][row][2] = __ldg(reinterpret_cast<const uint4*>(p0 + wrow));
      v[slot][row][3] = __ldg(reinterpret_cast<const uint4*>(p0 + wrow + rowb));
      wy[slot][row] = rec.y; wz[slot][row] = rec.z;
      if constexpr (MODE != MODE_BF16) ww[slot][row] = rec.w;
    };
    auto advance =[&](int& tap, int& cb) { if (++tap == K) { tap = 0; ++cb; in_plane += prm.plane_bytes; } };

    // prologue: fill the pipeline with k-blocks 0 .. DEPTH-1, fetch the records of k-block DEPTH
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) {
      if (d < nkb) {
        load_recs(tapI);
#pragma unroll
        for (int i = 0; i < RPT; ++i) issue(d, i, recn[i]);
        advance(tapI, cbI);
      }
    }
    tapR = tapI;                      // tap of k-block DEPTH
    if (DEPTH < nkb) load_recs(tapR);
    if (++tapR == K) tapR = 0;

    auto body = [&](int kb, int slot) {
      const int s = kb % NS;
      unsigned char* a_tile = smem + (size_t)s * stage_bytes;
      mbar_spin(&empty_bar[s], ((uint32_t)(kb / NS) & 1u) ^ 1u);     // MMAs of k-block kb - NS have retired
      if (tl && tid == 0) tl[4 + 2 * nkb + kb] = clock64();           // stage acquired
      const bool more = kb + DEPTH < nkb;
#pragma unroll
      for (int row = 0; row < RPT; ++row) {
        combine_store<MODE>(v[slot][row], wy[slot][row], wz[slot][row], MODE != MODE_BF16 ? ww[slot][row] : 0u,
                            a_tile + a_off + row * (ROW_STEP * 128));
        if (more) issue(slot, row, recn[row]);                       // re-arm: k-block kb + DEPTH
      }
      if (more) advance(tapI, cbI);
      // records for the next iteration's issue.  (Fetching them a whole k-block earlier instead was measured:
      // producer thread 0's work drops 880 -> 720 clk but the k-block period does not move -- the pace is set by
      // the two producer warps that share the control warp's sub-partition, see Layout above.)
      if (kb + DEPTH + 1 < nkb) {
        load_recs(tapR);
        if (++tapR == K) tapR = 0;
      }
      if (tl && tid == 0) tl[4 + 3 * nkb + kb] = clock64();           // combine + stores + re-arm issued
      fence_proxy_async_smem();   // my generic-proxy stores -> visible to tcgen05.mma
      __syncwarp();
      if (lane == 0) {
        if (PAIR && cta_rank != 0) mbar_arrive_remote(full_bar0_remote + (uint32_t)s * 8u);
        else mbar_arrive(&full_bar[s]);
      }
      if (tl && tid == 0) tl[4 + nkb + kb] = clock64();
      if (tl && lane == 0) tl[4 + (4 + pslot) * nkb + kb] = clock64();  // every producer warp's arrival
    };

    for (int kb = 0; kb < nkb; kb += DEPTH) {
#pragma unroll
      for (int d = 0; d < DEPTH; ++d)
        if (kb + d < nkb) body(kb + d, d);
    }
  }

  if (SPREAD || pslot >= 0) {
    // ===================== epilogue: TMEM -> registers -> NCHW global =====================
    // compact layout: the producer warps; spread layout: all 12 warps (three per TMEM lane quarter)
    mbar_spin(tmem_full_bar, 0);
    if (tl && tid == 0) tl[2 + nkb] = clock64();
    tc_fence_after();
    const int q = warp & 3, cgrp = warp >> 2;      // TMEM lane quarter (hardware: warp id % 4) / column group
    const int row = q * 32 + lane;
    const int m = m0 + row;
    const bool row_ok = m < prm.M;
    const int n = row_ok ? m / prm.HoWo : 0;
    const int pos = row_ok ? m - n * prm.HoWo : 0;
    Tout* obase = reinterpret_cast<Tout*>(prm.out) + ((size_t)n * prm.out_ctot + prm.out_coff) * prm.HoWo + pos;
    // Tiled bf16 output whose channel slice starts on a 64-channel slab: the tile's slabs are staged in shared
    // memory (the pipeline stages are idle once the accumulator is complete) in exactly the global slab layout and
    // leave as 16 KB cp.async.bulk stores -- a thread owns an accumulator ROW, so direct stores put 32 different
    // lines (16 bytes each) into every store instruction.
    const bool staged = sizeof(Tout) == 2 && prm.out_nhwc && !prm.partial && (prm.out_coff & 63) == 0 && !PAIR;
    const int nslab = BN >> 6;
    // 32-column chunks of the accumulator are dealt round-robin to the PRODUCER_WARPS / 4 column groups
    // (Cout % 64 == 0, so every chunk is whole)
    for (int col = cgrp * 32; col < BN; col += 32 * LY::EPI_GROUPS) {         // warp-uniform
      uint32_t acc[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)col, acc);
      tmem_ld_wait();
      if (staged) {
        if constexpr (sizeof(Tout) == 2) {
          const bool split = prm.out_nhwc == KGDET_LAYOUT_TILED_SPLIT;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            float x[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              x[e] = __uint_as_float(acc[j + e]);
              if (prm.bias) x[e] += __ldg(prm.bias + col + j + e);
              if (prm.relu) x[e] = fmaxf(x[e], 0.f);
            }
            const int c = col + j;                                   // channel within this call's Cout
            unsigned char* dst = smem + (size_t)(c >> 6) * A_TILE_BYTES + row * 128 + ((((c & 63) >> 3) ^ (row & 7)) << 4);
            uint4 hi4;
            hi4.x = pack_bf16x2(x[0], x[1]); hi4.y = pack_bf16x2(x[2], x[3]);
            hi4.z = pack_bf16x2(x[4], x[5]); hi4.w = pack_bf16x2(x[6], x[7]);
            *reinterpret_cast<uint4*>(dst) = hi4;
            if (split) {
#pragma unroll
              for (int e = 0; e < 8; ++e) x[e] -= __bfloat162float(__float2bfloat16(x[e]));
              uint4 lo4;
              lo4.x = pack_bf16x2(x[0], x[1]); lo4.y = pack_bf16x2(x[2], x[3]);
              lo4.z = pack_bf16x2(x[4], x[5]); lo4.w = pack_bf16x2(x[6], x[7]);
              *reinterpret_cast<uint4*>(dst + (size_t)nslab * A_TILE_BYTES) = lo4;
            }
          }
        }
      } else if (prm.partial) {
        // split-K: raw accumulators, position-major (m_pad = whole tiles, so every row may be written)
        float* prow = prm.partial + ((size_t)blockIdx.y * prm.m_pad + m0 + row) * BN + col;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(prow + j) = make_float4(__uint_as_float(acc[j]), __uint_as_float(acc[j + 1]),
                                                             __uint_as_float(acc[j + 2]), __uint_as_float(acc[j + 3]));
      } else if (row_ok) {
        if (prm.out_nhwc) {
          // "UMMA-tiled rows" (pointwise_umma.cu): bf16, [M/128 tiles][k-blocks of 64 channels][128 rows x 128 B,
          // 16-byte chunk c of row r at chunk c ^ (r & 7)] -- exactly the A operand slabs of the 1x1-convolution
          // GEMM that follows, which then needs one bulk copy per k-block.  Split layout: the k-blocks of the hi
          // parts are followed by those of the lo parts (x = hi + lo).
          if constexpr (sizeof(Tout) == 2) {
            const bool split = prm.out_nhwc == KGDET_LAYOUT_TILED_SPLIT;
            const int kblocks = (split ? 2 : 1) * (prm.out_ctot >> 6);
            unsigned char* tile = reinterpret_cast<unsigned char*>(prm.out) + (size_t)(m >> 7) * kblocks * A_TILE_BYTES +
                                  (size_t)(m & 127) * 128;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              float x[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                x[e] = __uint_as_float(acc[j + e]);
                if (prm.bias) x[e] += __ldg(prm.bias + col + j + e);
                if (prm.relu) x[e] = fmaxf(x[e], 0.f);
              }
              const int k = prm.out_coff + col + j;                  // logical channel of x[0]
              const size_t off = (size_t)(k >> 6) * A_TILE_BYTES + ((((k & 63) >> 3) ^ (m & 7)) << 4);
              uint4 hi4;
              hi4.x = pack_bf16x2(x[0], x[1]); hi4.y = pack_bf16x2(x[2], x[3]);
              hi4.z = pack_bf16x2(x[4], x[5]); hi4.w = pack_bf16x2(x[6], x[7]);
              *reinterpret_cast<uint4*>(tile + off) = hi4;
              if (split) {
#pragma unroll
                for (int e = 0; e < 8; ++e) x[e] -= __bfloat162float(__float2bfloat16(x[e]));
                uint4 lo4;
                lo4.x = pack_bf16x2(x[0], x[1]); lo4.y = pack_bf16x2(x[2], x[3]);
                lo4.z = pack_bf16x2(x[4], x[5]); lo4.w = pack_bf16x2(x[6], x[7]);
                *reinterpret_cast<uint4*>(tile + off + (size_t)(prm.out_ctot >> 6) * A_TILE_BYTES) = lo4;
              }
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = __uint_as_float(acc[j]);
            if (prm.bias) x += __ldg(prm.bias + col + j);
            if (prm.relu) x = fmaxf(x, 0.f);
            st_out<Tout>(obase + (size_t)(col + j) * prm.HoWo, x);   // lanes = consecutive positions
          }
        }
      }
    }
    if (staged) {
      // every epilogue thread has written its share of the slabs; one thread ships them
      fence_proxy_async_smem();
      constexpr int EPI_THREADS = SPREAD ? LY::THREADS : PRODUCER_WARPS * 32;
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
      if (warp == 0 && lane == 0) {                 // warp 0 takes part in the epilogue in both layouts
        const bool split = prm.out_nhwc == KGDET_LAYOUT_TILED_SPLIT;
        const int kblocks = (split ? 2 : 1) * (prm.out_ctot >> 6);
        unsigned char* tile = reinterpret_cast<unsigned char*>(prm.out) + (size_t)(m0 >> 7) * kblocks * A_TILE_BYTES +
                              (size_t)(prm.out_coff >> 6) * A_TILE_BYTES;
        for (int sl = 0; sl < nslab; ++sl) {
          bulk_s2g(tile + (size_t)sl * A_TILE_BYTES, smem + (size_t)sl * A_TILE_BYTES, A_TILE_BYTES);
          if (split)
            bulk_s2g(tile + (size_t)((prm.out_ctot >> 6) + sl) * A_TILE_BYTES,
                     smem + (size_t)(nslab + sl) * A_TILE_BYTES, A_TILE_BYTES);
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory must outlive the reads
      }
    }
  }

  if (tl && tid == 0) tl[3 + nkb] = clock64();
  tc_fence_before();
  if constexpr (PAIR) {
    cluster_sync_all();          // neither CTA may free the shared allocation while the other still reads
    if (warp == SETUP_WARP) tmem_dealloc_pair(tmem_base, prm.tmem_cols);
  } else {
    __syncthreads();
    if (warp == SETUP_WARP) tmem_dealloc(tmem_base, prm.tmem_cols);
  }
}

static size_t stream_smem_bytes(int mode, int ns, int Cout, bool pair) {
  const int a_tiles = (mode == MODE_TF32X3) ? 2 : 1, b_tiles = a_tiles;
  const size_t stage = (size_t)a_tiles * A_TILE_BYTES + (size_t)b_tiles * (pair ? Cout / 2 : Cout) * 128;
  return 1024 /* alignment slack */ + ns * stage + (2 * ns + 1) * 8 + 16;   // barriers, TMEM slot, duty ticket
}

template <int MODE, int NS, int DEPTH, typename Tout, bool PAIR, int RPT, bool SPREAD>
static int launch_stream(const UmmaParams& p, int grid, cudaStream_t stream, int splits) {
  const size_t smem = stream_smem_bytes(MODE, NS, p.Cout, PAIR);
  // the instrumented twin exists for the bf16 single-CTA kernel only (tools/dcn_timeline.py)
  constexpr bool CAN_TL = MODE == MODE_BF16 && !PAIR && NS == 3;
  auto kern = dcn_umma_stream_kernel<MODE, NS, DEPTH, Tout, PAIR, RPT, SPREAD, false>;
  if constexpr (CAN_TL) {
    if (p.timeline) kern = dcn_umma_stream_kernel<MODE, NS, DEPTH, Tout, PAIR, RPT, SPREAD, true>;
  }
  KG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid, (unsigned)splits, 1);   // PAIR: even, one cluster = two consecutive 128-row tiles
  cfg.blockDim = dim3(Layout<RPT, SPREAD>::THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  KG_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  KG_LAUNCH_CHECK("dcn_umma_stream_kernel");
  return KGDET_OK;
}

template <int MODE, typename Tout, bool PAIR>
static int dispatch_stages(const UmmaParams& p, int grid, int ns, cudaStream_t stream, int splits) {
  // rows per producer thread: 4 (8 producer warps) measured best -- K = 49 call 137 us vs 145 us with 2 rows
  // (16 warps) and 161 us with 8 rows (4 warps)
  // spread warp layout: measured equal to the compact one in the bench step (701.2 vs 701.7 TFLOP/s; its 2-producer
  // sub-partition runs ahead but the two 3-producer ones then set the pace) -- opt-in with KGDET_UMMA_SPREAD=1
  bool spread = false;
  if (const char* e = getenv("KGDET_UMMA_SPREAD")) spread = !PAIR && atoi(e) != 0;
  if constexpr (!PAIR) {
    if (spread) {
      switch (ns) {
        case 2: return launch_stream<MODE, 2, 1, Tout, false, 4, true>(p, grid, stream, splits);
        case 3: return launch_stream<MODE, 3, 1, Tout, false, 4, true>(p, grid, stream, splits);
        case 4: return launch_stream<MODE, 4, 1, Tout, false, 4, true>(p, grid, stream, splits);
        default: break;
      }
    }
  }
  switch (ns) {
    case 2: return launch_stream<MODE, 2, 1, Tout, PAIR, 4, false>(p, grid, stream, splits);
    case 3: return launch_stream<MODE, 3, 1, Tout, PAIR, 4, false>(p, grid, stream, splits);
    case 4: return launch_stream<MODE, 4, 1, Tout, PAIR, 4, false>(p, grid, stream, splits);
    default: break;
  }
  set_error("dcn umma stream: unsupported stage count %d", ns);
  return KGDET_ERR_INVALID_ARG;
}

template <int MODE, typename Tout>
static int dispatch_pair(const UmmaParams& p, int grid, int ns, bool pair, cudaStream_t stream, int splits) {
  return pair ? dispatch_stages<MODE, Tout, true>(p, grid, ns, stream, splits)
              : dispatch_stages<MODE, Tout, false>(p, grid, ns, stream, splits);
}

int umma_stream_forward(const DcnGeom& g, const UmmaParams& p0, int mode, bool pair, int out_dtype,
                        cudaStream_t stream, int splits) {
  UmmaParams p = p0;
  p.tmem_cols = g.Cout <= 64 ? 64 : (g.Cout <= 128 ? 128 : 256);   // one accumulator, power of two >= 32
  const int grid = pair ? 2 * ceil_div(g.M, 2 * BM) : ceil_div(g.M, BM);
  // 3 stages measured best for every mode (more L1 for the gather than 4, enough slack for the MMA)
  int ns = (mode == MODE_TF32X3 && !pair) ? 2 : 3;
  if (const char* e = getenv("KGDET_UMMA_STAGES")) {
    const int v = atoi(e);
    if (v >= 2 && v <= 4) ns = v;
  }

  while (ns > 2 && stream_smem_bytes(mode, ns, g.Cout, pair) > 227 * 1024) --ns;
  if (stream_smem_bytes(mode, ns, g.Cout, pair) > 227 * 1024) {
    set_error("dcn umma stream: tile does not fit shared memory");
    return KGDET_ERR_UNSUPPORTED;
  }
  const bool f32 = out_dtype == KGDET_F32;
  switch (mode) {
    case MODE_BF16:
      return f32 ? dispatch_pair<MODE_BF16, float>(p, grid, ns, pair, stream, splits)
                 : dispatch_pair<MODE_BF16, __nv_bfloat16>(p, grid, ns, pair, stream, splits);
    case MODE_TF32X3:
      return f32 ? dispatch_pair<MODE_TF32X3, float>(p, grid, ns, pair, stream, splits)
                 : dispatch_pair<MODE_TF32X3, __nv_bfloat16>(p, grid, ns, pair, stream, splits);
    default:
      return f32 ? dispatch_pair<MODE_TF32, float>(p, grid, ns, pair, stream, splits)
                 : dispatch_pair<MODE_TF32, __nv_bfloat16>(p, grid, ns, pair, stream, splits);
  }
}

}  // namespace kgdet
