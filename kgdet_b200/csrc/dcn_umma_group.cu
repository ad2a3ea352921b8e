// Grouped, persistent form of the fused deformable-convolution forward (bf16 mode): ONE launch runs the tiles of
// up to KGDET_DCN_GROUP_MAX deformable convolutions -- the six DCNs of a Kp3RepBlock stage
// (reppoints_head_kp3rep_cas_1_assign_once.py:145-163: {cls, keypoint} branch x {9, 25, 49} points; they share two
// inputs and three sample plans) -- on one CTA per SM with a dynamic tile scheduler.
//
// Why: the single-problem kernel (dcn_umma_stream.cu) is one CTA per 128-position tile.  A KGDet call has 132
// tiles for 148 SMs (11 % of the machine idle for the whole call) and pays ~15 us of launch / prologue / epilogue /
// tail per call -- 39 % of the 9-point call (measured: T = 14.9 us + 0.648 us per k-block).  Here the 792 tiles of
// a stage (2 x 132 tiles of 36, 100 and 196 k-blocks) are handed out longest first from an atomic counter, every
// SM stays busy until the stage is done, and TMEM allocation / barrier set-up / launch happen once.
//
// Per tile the pipeline is the single-problem kernel's (same arithmetic, bit-identical results): 8 producer warps
// gather + interpolate the A tile of a k-block (64 channels of one tap) into the 128B-swizzled shared-memory
// stage, the control lane streams the pre-swizzled weight slab with cp.async.bulk and issues the tcgen05.mma's,
// the producer warps then run the epilogue (tcgen05.ld -> bias / ReLU -> NCHW, or bf16 [hi | lo] rows in the tiled
// layout of the pointwise GEMM, staged in shared memory and bulk-stored).  Stage / phase counters run on across
// tiles; a __syncthreads per tile publishes the next tile index.
#include <cuda_bf16.h>

#include <type_traits>

#include "dcn_umma.cuh"

namespace kgdet {

static constexpr int GP_NS = 3;                 // pipeline stages
static constexpr int GP_RPT = 4;                // rows per producer thread
static constexpr int GP_PWARPS = 8;             // producer warps
static constexpr int GP_THREADS = (GP_PWARPS + 1) * 32;
static constexpr int GP_ROW_STEP = 128 / GP_RPT;
// DEPTH = k-blocks of gathers a producer thread keeps in flight.  With one, every warp waits once per k-block for
// the slowest of its 64 cache lines (11 % of the sectors miss L1, so there is always one coming back from L2):
// `long_scoreboard` is the top stall.  Two k-blocks in flight need 64 more registers per producer thread than the
// 168 that 9 warps can have, so the DEPTH = 2 kernel is launched with 12 warps (three warpgroups) and moves the
// registers of the third one (control warp + three idle warps) to the producers with setmaxnreg: 224 / 56.
template <int DEPTH> struct GroupLayout {
  static constexpr int THREADS = DEPTH == 2 ? 384 : GP_THREADS;
  static constexpr int SYNC_THREADS = GP_THREADS;        // producers + control warp meet at the tile boundaries
};

struct GroupProblem {
  const void* in;            // channel-blocked bf16 planes (first pixel of plane 0)
  size_t plane_bytes;
  const SampleRec16* plan;   // [K][rows_padded], bf16-weight flavour
  const unsigned char* wp;   // packed weights
  const float* bias;
  void* out;
  int M, W, K, HoWo, rows_padded;
  int out_coff, out_ctot, relu, out_layout, out_dtype;
  int nkb;                   // (C / 64) * K
  int tile_begin;            // first global tile index of this problem (problems sorted by nkb, descending)
};

struct GroupParams {
  GroupProblem prob[KGDET_DCN_GROUP_MAX];
  int nprob, total_tiles;
  int Cout;
  uint32_t idesc, tmem_cols;
  int* counter;              // zeroed before the launch
  long long* timeline;       // development hook (KGDET_GROUP_TIMELINE): per CTA 64 tiles x 8 clock64() stamps, or NULL
};

__device__ __forceinline__ void gp_spin(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 22)) __trap();
  }
}
__device__ __forceinline__ __nv_bfloat162 gp_bf162(uint32_t v) { return *reinterpret_cast<__nv_bfloat162*>(&v); }
__device__ __forceinline__ uint32_t gp_u32(__nv_bfloat162 v) { return *reinterpret_cast<uint32_t*>(&v); }

// one 16-byte chunk of one row: sum of the four corners times their (pre-rounded bf16) weights; same rounding
// sequence as dcn_umma_stream.cu's combine_store<MODE_BF16>
__device__ __forceinline__ void gp_combine_store(const uint4 (&v)[4], uint32_t wy, uint32_t wz, unsigned char* dst) {
  const __nv_bfloat162 w01 = gp_bf162(wy), w23 = gp_bf162(wz);
  const __nv_bfloat162 w0 = __low2bfloat162(w01), w1 = __high2bfloat162(w01);
  const __nv_bfloat162 w2 = __low2bfloat162(w23), w3 = __high2bfloat162(w23);
  uint4 o;
  o.x = gp_u32(__hfma2(w3, gp_bf162(v[3].x), __hfma2(w2, gp_bf162(v[2].x), __hfma2(w1, gp_bf162(v[1].x), __hmul2(w0, gp_bf162(v[0].x))))));
  o.y = gp_u32(__hfma2(w3, gp_bf162(v[3].y), __hfma2(w2, gp_bf162(v[2].y), __hfma2(w1, gp_bf162(v[1].y), __hmul2(w0, gp_bf162(v[0].y))))));
  o.z = gp_u32(__hfma2(w3, gp_bf162(v[3].z), __hfma2(w2, gp_bf162(v[2].z), __hfma2(w1, gp_bf162(v[1].z), __hmul2(w0, gp_bf162(v[0].z))))));
  o.w = gp_u32(__hfma2(w3, gp_bf162(v[3].w), __hfma2(w2, gp_bf162(v[2].w), __hfma2(w1, gp_bf162(v[1].w), __hmul2(w0, gp_bf162(v[0].w))))));
  *reinterpret_cast<uint4*>(dst) = o;
}

__device__ __forceinline__ void gp_bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}

template <int DEPTH>
__global__ void __launch_bounds__(GroupLayout<DEPTH>::THREADS, 1) dcn_umma_group_kernel(const __grid_constant__ GroupParams gp) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  const int BN = gp.Cout;
  const int b_tile_bytes = BN * 128;
  const int stage_bytes = A_TILE_BYTES + b_tile_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)GP_NS * stage_bytes);
  uint64_t* empty_bar = full_bar + GP_NS;
  uint64_t* tmem_full_bar = empty_bar + GP_NS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  int* next_tile = reinterpret_cast<int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool is_control = warp == GP_PWARPS;
  const int tid = threadIdx.x;                  // producer thread index for warps 0..7

  if (is_control) {
    if (lane == 0) {
      for (int s = 0; s < GP_NS; ++s) {
        mbar_init(&full_bar[s], GP_PWARPS + 1);     // 8 producer warps + the control lane's expect_tx
        mbar_init(&empty_bar[s], 1);                // one tcgen05.commit
      }
      mbar_init(tmem_full_bar, 1);
      fence_mbar_init();
      *next_tile = atomicAdd(gp.counter, 1);
    }
    __syncwarp();
    tmem_alloc(tmem_slot, gp.tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  int tile = *next_tile;
  int it_base = 0;                              // k-blocks this CTA has pushed through the ring so far
  uint32_t tile_count = 0;
  // tile-boundary barrier of the 288 working threads (named: the idle warps of the DEPTH = 2 layout have left)
  auto sync_workers = [&]() { asm volatile("bar.sync 2, %0;" ::"n"(GroupLayout<DEPTH>::SYNC_THREADS) : "memory"); };

  // The tile loop is instantiated once per role and each instance sits in its own top-level branch, directly
  // after that role's setmaxnreg: ptxas then allocates the producer instance up to the raised register limit.
  auto tile_loop = [&](auto ctrl_tag) {
  constexpr bool CTRL = decltype(ctrl_tag)::value;
  while (tile < gp.total_tiles) {
    // ---- which problem, which rows ----
    int pi = 0;
#pragma unroll
    for (int q = 1; q < KGDET_DCN_GROUP_MAX; ++q)
      if (q < gp.nprob && tile >= gp.prob[q].tile_begin) pi = q;
    const GroupProblem& P = gp.prob[pi];
    const int m0 = (tile - P.tile_begin) * BM;
    const int nkb = P.nkb;
    long long* const tl = (gp.timeline && tile_count < 64) ? gp.timeline + ((size_t)blockIdx.x * 64 + tile_count) * 8 : nullptr;
    if (tl && threadIdx.x == 0) { tl[0] = clock64(); tl[5] = nkb; }

    if constexpr (CTRL) {
      // =========================== control lane ===========================
      if (lane == 0) {
        auto fetch_b = [&](int kq) {            // weight slab of this tile's k-block kq into its ring stage
          const int sq = (it_base + kq) % GP_NS;
          unsigned char* dstb = smem + (size_t)sq * stage_bytes + A_TILE_BYTES;
          mbar_arrive_expect_tx(&full_bar[sq], (uint32_t)b_tile_bytes);
          bulk_g2s(dstb, P.wp + (size_t)kq * b_tile_bytes, (uint32_t)b_tile_bytes, &full_bar[sq]);
        };
        for (int j = 0; j < GP_NS - 1 && j < nkb; ++j) fetch_b(j);      // the ring is drained at a tile boundary
        for (int j = 0; j < nkb; ++j) {
          const int G = it_base + j, s = G % GP_NS;
          gp_spin(&full_bar[s], (uint32_t)(G / GP_NS) & 1u);
          if (tl && j == 0) tl[1] = clock64();
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + (size_t)s * stage_bytes);
          const uint64_t adesc = make_sw128_kmajor_desc(a_addr);
          const uint64_t bdesc = make_sw128_kmajor_desc(a_addr + A_TILE_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem_base, adesc + 2 * k, bdesc + 2 * k, gp.idesc, (j > 0 || k > 0) ? 1u : 0u);
          tc_commit(&empty_bar[s]);
          if (j == nkb - 1) { tc_commit(tmem_full_bar); if (tl) tl[2] = clock64(); }
          const int kn = j + GP_NS - 1;
          if (kn < nkb) {
            // stage of k-block G - 1 (MMAs issued one duty ago); for j == 0 it was drained with the previous tile
            if (j >= 1) gp_spin(&empty_bar[(G + GP_NS - 1) % GP_NS], (uint32_t)((G - 1) / GP_NS) & 1u);
            fetch_b(kn);
          }
        }
      }
      __syncwarp();
    } else {
      // =========================== producers ===========================
      const int chunk = tid & 7, rbase = tid >> 3;
      const int K = P.K;
      constexpr long long rowb = 128;
      const long long wrow = (long long)P.W * rowb;
      const unsigned char* in_base = reinterpret_cast<const unsigned char*>(P.in) + chunk * 16;
      const uint4* plan0 = reinterpret_cast<const uint4*>(P.plan) + (m0 + rbase);
      const size_t tap_stride = (size_t)P.rows_padded;
      const int a_off = rbase * 128 + ((chunk ^ (rbase & 7)) << 4);

      uint4 v[DEPTH][GP_RPT][4];
      uint32_t wy[DEPTH][GP_RPT], wz[DEPTH][GP_RPT];
      uint4 recn[GP_RPT];
      auto load_recs = [&](int tap) {
#pragma unroll
        for (int i = 0; i < GP_RPT; ++i) recn[i] = __ldg(plan0 + (size_t)i * GP_ROW_STEP + tap * tap_stride);
      };
      int tapI = 0, tapR = 0;
      const unsigned char* in_plane = in_base;
      auto issue = [&](int slot, int row, const uint4& rec) {
        const unsigned char* p0 = in_plane + (long long)(int)rec.x * rowb;
        v[slot][row][0] = __ldg(reinterpret_cast<const uint4*>(p0));
        v[slot][row][1] = __ldg(reinterpret_cast<const uint4*>(p0 + rowb));
        v[slot][row][2] = __ldg(reinterpret_cast<const uint4*>(p0 + wrow));
        v[slot][row][3] = __ldg(reinterpret_cast<const uint4*>(p0 + wrow + rowb));
        wy[slot][row] = rec.y; wz[slot][row] = rec.z;
      };
      auto advance = [&]() { if (++tapI == K) { tapI = 0; in_plane += P.plane_bytes; } };

      // prologue: gathers of k-blocks 0 .. DEPTH-1 in flight, records of k-block DEPTH fetched
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) {
        if (d < nkb) {
          load_recs(tapI);
#pragma unroll
          for (int i = 0; i < GP_RPT; ++i) issue(d, i, recn[i]);
          advance();
        }
      }
      tapR = tapI;
      if (DEPTH < nkb) load_recs(tapR);
      if (++tapR == K) tapR = 0;

      auto body = [&](int kb, int slot) {
        const int G = it_base + kb, s = G % GP_NS;
        unsigned char* a_tile = smem + (size_t)s * stage_bytes;
        gp_spin(&empty_bar[s], ((uint32_t)(G / GP_NS) & 1u) ^ 1u);
        const bool more = kb + DEPTH < nkb;
#pragma unroll
        for (int row = 0; row < GP_RPT; ++row) {
          gp_combine_store(v[slot][row], wy[slot][row], wz[slot][row], a_tile + a_off + row * (GP_ROW_STEP * 128));
          if (more) issue(slot, row, recn[row]);                 // re-arm: k-block kb + DEPTH
        }
        if (more) advance();
        if (kb + DEPTH + 1 < nkb) {
          load_recs(tapR);
          if (++tapR == K) tapR = 0;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[s]);
      };
      for (int kb = 0; kb < nkb; kb += DEPTH) {
#pragma unroll
        for (int d = 0; d < DEPTH; ++d)
          if (kb + d < nkb) body(kb + d, d);
      }

      // =========================== epilogue ===========================
      if (tl && threadIdx.x == 0) tl[6] = clock64();          // last A tile handed over
      gp_spin(tmem_full_bar, tile_count & 1u);
      if (tl && threadIdx.x == 0) tl[3] = clock64();
      tc_fence_after();
      const int q = warp & 3, cgrp = warp >> 2;
      const int row = q * 32 + lane;
      const int m = m0 + row;
      const bool row_ok = m < P.M;
      const int n = row_ok ? m / P.HoWo : 0;
      const int pos = row_ok ? m - n * P.HoWo : 0;
      const bool tiled = P.out_layout != KGDET_LAYOUT_NCHW;
      const bool split = P.out_layout == KGDET_LAYOUT_TILED_SPLIT;
      const int nslab = BN >> 6;
      for (int col = cgrp * 32; col < BN; col += 32 * (GP_PWARPS / 4)) {
        uint32_t acc[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)col, acc);
        tmem_ld_wait();
        if (tiled) {
          // slabs staged in shared memory in the global slab layout, then bulk-stored (see dcn_umma_stream.cu)
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            float x[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              x[e] = __uint_as_float(acc[j + e]);
              if (P.bias) x[e] += __ldg(P.bias + col + j + e);
              if (P.relu) x[e] = fmaxf(x[e], 0.f);
            }
            const int c = col + j;
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
        } else if (row_ok) {
          if (P.out_dtype == KGDET_F32) {
            float* obase = reinterpret_cast<float*>(P.out) + ((size_t)n * P.out_ctot + P.out_coff) * P.HoWo + pos;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float x = __uint_as_float(acc[j]);
              if (P.bias) x += __ldg(P.bias + col + j);
              if (P.relu) x = fmaxf(x, 0.f);
              obase[(size_t)(col + j) * P.HoWo] = x;
            }
          } else {
            __nv_bfloat16* obase = reinterpret_cast<__nv_bfloat16*>(P.out) + ((size_t)n * P.out_ctot + P.out_coff) * P.HoWo + pos;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float x = __uint_as_float(acc[j]);
              if (P.bias) x += __ldg(P.bias + col + j);
              if (P.relu) x = fmaxf(x, 0.f);
              obase[(size_t)(col + j) * P.HoWo] = __float2bfloat16(x);
            }
          }
        }
      }
      if (tiled) {
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, %0;" ::"n"(GP_PWARPS * 32) : "memory");
        if (warp == 0 && lane == 0) {
          const int kblocks = (split ? 2 : 1) * (P.out_ctot >> 6);
          unsigned char* tbase = reinterpret_cast<unsigned char*>(P.out) + (size_t)(m0 >> 7) * kblocks * A_TILE_BYTES +
                                 (size_t)(P.out_coff >> 6) * A_TILE_BYTES;
          for (int sl = 0; sl < nslab; ++sl) {
            gp_bulk_s2g(tbase + (size_t)sl * A_TILE_BYTES, smem + (size_t)sl * A_TILE_BYTES, A_TILE_BYTES);
            if (split)
              gp_bulk_s2g(tbase + (size_t)((P.out_ctot >> 6) + sl) * A_TILE_BYTES,
                          smem + (size_t)(nslab + sl) * A_TILE_BYTES, A_TILE_BYTES);
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the stages are refilled right after
        }
      }
      tc_fence_before();      // the accumulator has been read: the next tile's first MMA may overwrite it
      if (tl && threadIdx.x == 0) tl[4] = clock64();
    }

    // ---- next tile: the ring is drained, every role is done with this tile ----
    if (threadIdx.x == GP_PWARPS * 32) *next_tile = atomicAdd(gp.counter, 1);
    sync_workers();
    tc_fence_after();
    tile = *next_tile;
    it_base += nkb;
    ++tile_count;
    sync_workers();           // everyone has read next_tile before the control lane overwrites it again
  }
  };

  if (!is_control && warp < GP_PWARPS) {
    if constexpr (DEPTH == 2) asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    tile_loop(std::false_type{});
  } else {
    if constexpr (DEPTH == 2) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
      if (warp > GP_PWARPS) return;             // the three idle warps of the third warpgroup
    }
    tile_loop(std::true_type{});
  }

  tc_fence_before();
  sync_workers();
  if (is_control) tmem_dealloc(tmem_base, gp.tmem_cols);
}


// ---------------------------------------------------------------------------------------------------------------
// 256-row variant: a tile is TWO 128-position halves that share every weight slab.  The loop of the kernel above is
// bound by the bytes a k-block moves through the SM's L1 / shared-memory array (64 KB of gathers, 16 KB of A stores,
// 32 KB of weight bulk-writes, 48 KB of MMA operand reads); here a weight slab is written once per TWO half-steps
// (16 KB per half-step), the pipeline shrinks to 128 KB (A ring 4 x 16 KB, B ring 2 x 32 KB -> the 132 KB carve-out
// instead of 164 KB: 96 KB of L1 for the gather instead of 64) and there are half as many tiles to fill and drain.
// Two TMEM accumulators (512 columns), one per half; the arithmetic per output element is unchanged (bit-identical).
//   step t = 2 * k-block + half:  producers gather + interpolate the A tile of (k-block, half) into the A ring;
//   the control lane issues MMA(acc[half], A[t], B[k-block]); B(k-block + 1) is fetched when B(k-block - 1) retires.
static constexpr int G2_NA = 4, G2_NB = 2;

__global__ void __launch_bounds__(GP_THREADS, 1) dcn_umma_group256_kernel(const __grid_constant__ GroupParams gp) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  const int BN = gp.Cout;
  const int b_tile_bytes = BN * 128;
  unsigned char* const a_ring = smem;
  unsigned char* const b_ring = smem + (size_t)G2_NA * A_TILE_BYTES;
  uint64_t* full_a = reinterpret_cast<uint64_t*>(b_ring + (size_t)G2_NB * b_tile_bytes);
  uint64_t* empty_a = full_a + G2_NA;
  uint64_t* full_b = empty_a + G2_NA;
  uint64_t* empty_b = full_b + G2_NB;
  uint64_t* tmem_full_bar = empty_b + G2_NB;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  int* next_tile = reinterpret_cast<int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool is_control = warp == GP_PWARPS;
  const int tid = threadIdx.x;

  if (is_control) {
    if (lane == 0) {
      for (int s = 0; s < G2_NA; ++s) { mbar_init(&full_a[s], GP_PWARPS); mbar_init(&empty_a[s], 1); }
      for (int s = 0; s < G2_NB; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
      mbar_init(tmem_full_bar, 1);
      fence_mbar_init();
      *next_tile = atomicAdd(gp.counter, 1);
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 2 * gp.tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  int tile = *next_tile;
  int step_base = 0, kb_base = 0;                 // half-steps / k-blocks this CTA has pushed through the rings
  uint32_t tile_count = 0;

  while (tile < gp.total_tiles) {
    int pi = 0;
#pragma unroll
    for (int q = 1; q < KGDET_DCN_GROUP_MAX; ++q)
      if (q < gp.nprob && tile >= gp.prob[q].tile_begin) pi = q;
    const GroupProblem& P = gp.prob[pi];
    const int m0 = (tile - P.tile_begin) * (2 * BM);
    const int nkb = P.nkb, nsteps = 2 * nkb;

    if (is_control) {
      if (lane == 0) {
        auto fetch_b = [&](int kq) {
          const int sq = (kb_base + kq) % G2_NB;
          mbar_arrive_expect_tx(&full_b[sq], (uint32_t)b_tile_bytes);
          bulk_g2s(b_ring + (size_t)sq * b_tile_bytes, P.wp + (size_t)kq * b_tile_bytes, (uint32_t)b_tile_bytes, &full_b[sq]);
        };
        fetch_b(0);                               // both rings are drained at a tile boundary
        for (int kb = 0; kb < nkb; ++kb) {
          const int KG = kb_base + kb, sb = KG % G2_NB;
          gp_spin(&full_b[sb], (uint32_t)(KG / G2_NB) & 1u);
          const uint64_t bdesc = make_sw128_kmajor_desc(smem_u32(b_ring + (size_t)sb * b_tile_bytes));
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int T = step_base + 2 * kb + h, sa = T % G2_NA;
            gp_spin(&full_a[sa], (uint32_t)(T / G2_NA) & 1u);
            tc_fence_after();
            const uint64_t adesc = make_sw128_kmajor_desc(smem_u32(a_ring + (size_t)sa * A_TILE_BYTES));
            const uint32_t acc = tmem_base + (uint32_t)(h * gp.tmem_cols);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(acc, adesc + 2 * k, bdesc + 2 * k, gp.idesc, (kb > 0 || k > 0) ? 1u : 0u);
            tc_commit(&empty_a[sa]);
            if (h == 0 && kb + 1 < nkb) {
              // weight slab of the NEXT k-block into the slot of k-block KG - 1: its MMAs were issued two half-steps
              // ago and have retired by now (this wait ends at once); the slab then has two half-steps to land
              if (kb >= 1) gp_spin(&empty_b[(KG + 1) % G2_NB], (uint32_t)((KG - 1) / G2_NB) & 1u);
              fetch_b(kb + 1);
            }
          }
          tc_commit(&empty_b[sb]);
        }
        tc_commit(tmem_full_bar);
      }
      __syncwarp();
    } else {
      // =========================== producers ===========================
      const int chunk = tid & 7, rbase = tid >> 3;
      const int K = P.K;
      constexpr long long rowb = 128;
      const long long wrow = (long long)P.W * rowb;
      const unsigned char* in_base = reinterpret_cast<const unsigned char*>(P.in) + chunk * 16;
      const uint4* plan0 = reinterpret_cast<const uint4*>(P.plan) + (m0 + rbase);
      const size_t tap_stride = (size_t)P.rows_padded;
      const int a_off = rbase * 128 + ((chunk ^ (rbase & 7)) << 4);

      uint4 v[GP_RPT][4];
      uint32_t wy[GP_RPT], wz[GP_RPT];
      uint4 recn[GP_RPT];
      auto load_recs = [&](int tap, int half) {
#pragma unroll
        for (int i = 0; i < GP_RPT; ++i) recn[i] = __ldg(plan0 + half * BM + (size_t)i * GP_ROW_STEP + tap * tap_stride);
      };
      const unsigned char* in_plane = in_base;    // plane of the step whose gathers are issued next
      auto issue = [&](int row, const uint4& rec) {
        const unsigned char* p0 = in_plane + (long long)(int)rec.x * rowb;
        v[row][0] = __ldg(reinterpret_cast<const uint4*>(p0));
        v[row][1] = __ldg(reinterpret_cast<const uint4*>(p0 + rowb));
        v[row][2] = __ldg(reinterpret_cast<const uint4*>(p0 + wrow));
        v[row][3] = __ldg(reinterpret_cast<const uint4*>(p0 + wrow + rowb));
        wy[row] = rec.y; wz[row] = rec.z;
      };
      // tapI: tap of the step whose gathers are issued next; tapR: tap of the step whose records are fetched next
      int tapI = 0, tapR = 0;
      load_recs(0, 0);
#pragma unroll
      for (int i = 0; i < GP_RPT; ++i) issue(i, recn[i]);     // step 0 = (k-block 0, half 0)
      load_recs(0, 1);                                        // records of step 1 = (k-block 0, half 1)
      // after step 1 the issue / record cursors move to k-block 1
      for (int t = 0; t < nsteps; ++t) {
        const int T = step_base + t, sa = T % G2_NA;
        unsigned char* a_tile = a_ring + (size_t)sa * A_TILE_BYTES;
        gp_spin(&empty_a[sa], ((uint32_t)(T / G2_NA) & 1u) ^ 1u);
        const bool more = t + 1 < nsteps;
        if (more && ((t + 1) & 1) == 0) {                     // next step opens a new k-block: advance tap / plane
          if (++tapI == K) { tapI = 0; in_plane += P.plane_bytes; }
        }
#pragma unroll
        for (int row = 0; row < GP_RPT; ++row) {
          gp_combine_store(v[row], wy[row], wz[row], a_tile + a_off + row * (GP_ROW_STEP * 128));
          if (more) issue(row, recn[row]);                    // re-arm: step t + 1
        }
        if (t + 2 < nsteps) {
          if (((t + 2) & 1) == 0) { if (++tapR == K) tapR = 0; }
          load_recs(tapR, (t + 2) & 1);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_a[sa]);
      }

      // =========================== epilogue: the two halves one after the other ===========================
      gp_spin(tmem_full_bar, tile_count & 1u);
      tc_fence_after();
      const int q = warp & 3, cgrp = warp >> 2;
      const int row = q * 32 + lane;
      const bool tiled = P.out_layout != KGDET_LAYOUT_NCHW;
      const bool split = P.out_layout == KGDET_LAYOUT_TILED_SPLIT;
      const int nslab = BN >> 6;
      const int m_tiles = (P.M + BM - 1) / BM;
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        const int mh = m0 + h * BM;
        if ((mh >> 7) >= m_tiles) break;                      // odd number of 128-row tiles: the last half is empty
        const int m = mh + row;
        const bool row_ok = m < P.M;
        const int n = row_ok ? m / P.HoWo : 0;
        const int pos = row_ok ? m - n * P.HoWo : 0;
        const uint32_t acc_base = tmem_base + (uint32_t)(h * gp.tmem_cols);
        for (int col = cgrp * 32; col < BN; col += 32 * (GP_PWARPS / 4)) {
          uint32_t acc[32];
          tmem_ld32(acc_base + ((uint32_t)(q * 32) << 16) + (uint32_t)col, acc);
          tmem_ld_wait();
          if (tiled) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              float x[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                x[e] = __uint_as_float(acc[j + e]);
                if (P.bias) x[e] += __ldg(P.bias + col + j + e);
                if (P.relu) x[e] = fmaxf(x[e], 0.f);
              }
              const int c = col + j;
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
          } else if (row_ok) {
            if (P.out_dtype == KGDET_F32) {
              float* obase = reinterpret_cast<float*>(P.out) + ((size_t)n * P.out_ctot + P.out_coff) * P.HoWo + pos;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                float x = __uint_as_float(acc[j]);
                if (P.bias) x += __ldg(P.bias + col + j);
                if (P.relu) x = fmaxf(x, 0.f);
                obase[(size_t)(col + j) * P.HoWo] = x;
              }
            } else {
              __nv_bfloat16* obase = reinterpret_cast<__nv_bfloat16*>(P.out) + ((size_t)n * P.out_ctot + P.out_coff) * P.HoWo + pos;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                float x = __uint_as_float(acc[j]);
                if (P.bias) x += __ldg(P.bias + col + j);
                if (P.relu) x = fmaxf(x, 0.f);
                obase[(size_t)(col + j) * P.HoWo] = __float2bfloat16(x);
              }
            }
          }
        }
        if (tiled) {
          fence_proxy_async_smem();
          asm volatile("bar.sync 1, %0;" ::"n"(GP_PWARPS * 32) : "memory");
          if (warp == 0 && lane == 0) {
            const int kblocks = (split ? 2 : 1) * (P.out_ctot >> 6);
            unsigned char* tbase = reinterpret_cast<unsigned char*>(P.out) + (size_t)(mh >> 7) * kblocks * A_TILE_BYTES +
                                   (size_t)(P.out_coff >> 6) * A_TILE_BYTES;
            for (int sl = 0; sl < nslab; ++sl) {
              gp_bulk_s2g(tbase + (size_t)sl * A_TILE_BYTES, smem + (size_t)sl * A_TILE_BYTES, A_TILE_BYTES);
              if (split)
                gp_bulk_s2g(tbase + (size_t)((P.out_ctot >> 6) + sl) * A_TILE_BYTES,
                            smem + (size_t)(nslab + sl) * A_TILE_BYTES, A_TILE_BYTES);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          }
          // the staging area is rewritten by the second half / refilled by the next tile
          asm volatile("bar.sync 1, %0;" ::"n"(GP_PWARPS * 32) : "memory");
        }
      }
      tc_fence_before();
    }

    if (threadIdx.x == GP_PWARPS * 32) *next_tile = atomicAdd(gp.counter, 1);
    __syncthreads();
    tc_fence_after();
    tile = *next_tile;
    step_base += nsteps;
    kb_base += nkb;
    ++tile_count;
    __syncthreads();
  }

  tc_fence_before();
  __syncthreads();
  if (is_control) tmem_dealloc(tmem_base, 2 * gp.tmem_cols);
}

static size_t group256_smem_bytes(int Cout) {
  return 1024 + (size_t)G2_NA * A_TILE_BYTES + (size_t)G2_NB * Cout * 128 + (2 * G2_NA + 2 * G2_NB + 1) * 8 + 32;
}

static size_t group_smem_bytes(int Cout) {
  return 1024 + (size_t)GP_NS * (A_TILE_BYTES + (size_t)Cout * 128) + (2 * GP_NS + 1) * 8 + 32;
}

int umma_group_forward(GroupParams& gp, cudaStream_t stream, bool rows256) {
  if (rows256) {
    // staging of one half's tiled split output (hi + lo slabs) must fit the two rings
    const size_t smem2 = group256_smem_bytes(gp.Cout);
    KG_CUDA(cudaFuncSetAttribute(dcn_umma_group256_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    KG_CUDA(cudaMemsetAsync(gp.counter, 0, sizeof(int), stream));
    const int grid2 = gp.total_tiles < num_sms() ? gp.total_tiles : num_sms();
    dcn_umma_group256_kernel<<<grid2, GP_THREADS, smem2, stream>>>(gp);
    KG_LAUNCH_CHECK("dcn_umma_group256_kernel");
    return KGDET_OK;
  }
  const size_t smem = group_smem_bytes(gp.Cout);
  if (smem > 227 * 1024) {
    set_error("dcn group: tile does not fit shared memory");
    return KGDET_ERR_UNSUPPORTED;
  }
  int depth = 1;
  if (const char* e = getenv("KGDET_GROUP_DEPTH")) depth = atoi(e) == 2 ? 2 : 1;
  KG_CUDA(cudaMemsetAsync(gp.counter, 0, sizeof(int), stream));
  const int grid = gp.total_tiles < num_sms() ? gp.total_tiles : num_sms();
  if (depth == 2) {
    KG_CUDA(cudaFuncSetAttribute(dcn_umma_group_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dcn_umma_group_kernel<2><<<grid, GroupLayout<2>::THREADS, smem, stream>>>(gp);
  } else {
    KG_CUDA(cudaFuncSetAttribute(dcn_umma_group_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dcn_umma_group_kernel<1><<<grid, GroupLayout<1>::THREADS, smem, stream>>>(gp);
  }
  KG_LAUNCH_CHECK("dcn_umma_group_kernel");
  return KGDET_OK;
}

}  // namespace kgdet

using namespace kgdet;

static thread_local cudaEvent_t g_group_prof_start = nullptr, g_group_prof_stop = nullptr;
extern "C" void kgdet_dcn_group_set_profile_events(void* start_event, void* stop_event) {
  g_group_prof_start = (cudaEvent_t)start_event;
  g_group_prof_stop = (cudaEvent_t)stop_event;
}

extern "C" int kgdet_dcn_group_supported(const kgdet_dcn_shape* shape, int precision) {
  DcnGeom g;
  if (make_geom(shape, &g) != KGDET_OK) return 0;
  return (precision == KGDET_PREC_BF16 && umma_supported(g, precision) && g.Cout % 64 == 0 &&
          group_smem_bytes(g.Cout) <= 227 * 1024) ? 1 : 0;
}

extern "C" int kgdet_dcn_forward_prepared_group(const kgdet_dcn_group_item* items, int32_t count, int precision,
                                                void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  KG_CHECK_ARG(items && count >= 1 && count <= KGDET_DCN_GROUP_MAX, "kgdet_dcn_forward_prepared_group: 1..%d items",
               KGDET_DCN_GROUP_MAX);
  KG_CHECK_ARG(precision == KGDET_PREC_BF16, "kgdet_dcn_forward_prepared_group: bf16 mode only");
  KG_CHECK_ARG(workspace && workspace_bytes >= 16 && ((uintptr_t)workspace & 15) == 0,
               "kgdet_dcn_forward_prepared_group: needs 16 bytes of 16-byte aligned workspace (tile counter)");
  GroupParams gp;
  gp.nprob = count;
  gp.counter = (int*)workspace;
  // 256-row tiles (two halves sharing every weight slab) when the group is big enough to keep every SM busy with
  // them and Cout = 256 (the staging of one half's [hi | lo] output needs the 128 KB of the two rings)
  bool rows256 = true;
  if (const char* e = getenv("KGDET_GROUP_ROWS")) rows256 = atoi(e) != 128;
  // longest problems first: the scheduler hands tiles out in global tile order
  int order[KGDET_DCN_GROUP_MAX];
  DcnGeom geoms[KGDET_DCN_GROUP_MAX];
  for (int i = 0; i < count; ++i) {
    int rc = make_geom(&items[i].shape, &geoms[i]);
    if (rc != KGDET_OK) return rc;
    KG_CHECK_ARG(kgdet_dcn_group_supported(&items[i].shape, precision), "kgdet_dcn_forward_prepared_group: item %d is not "
                 "supported by the fused tensor-core path", i);
    KG_CHECK_ARG(geoms[i].Cout == geoms[0].Cout, "kgdet_dcn_forward_prepared_group: all items must have the same Cout");
    const kgdet_dcn_group_item& it = items[i];
    KG_CHECK_ARG(it.prepared_input && it.plan && it.weight_packed && it.output,
                 "kgdet_dcn_forward_prepared_group: NULL pointer in item %d", i);
    KG_CHECK_ARG(it.out_layout >= KGDET_LAYOUT_NCHW && it.out_layout <= KGDET_LAYOUT_TILED_SPLIT,
                 "kgdet_dcn_forward_prepared_group: bad output layout in item %d", i);
    KG_CHECK_ARG(it.out_channel_offset >= 0 && it.out_channel_offset + geoms[i].Cout <= it.out_channels_total,
                 "kgdet_dcn_forward_prepared_group: channel slice of item %d does not fit", i);
    if (it.out_layout != KGDET_LAYOUT_NCHW)
      KG_CHECK_ARG(it.dtype == KGDET_BF16 && it.out_channel_offset % 64 == 0 && it.out_channels_total % 64 == 0,
                   "kgdet_dcn_forward_prepared_group: tiled output of item %d needs bf16 and 64-channel aligned slices", i);
    order[i] = i;
  }
  for (int a = 1; a < count; ++a)              // insertion sort by k-blocks, descending, stable
    for (int b = a; b > 0 && geoms[order[b]].K > geoms[order[b - 1]].K; --b) {
      const int t = order[b]; order[b] = order[b - 1]; order[b - 1] = t;
    }
  int tiles = 0;
  for (int k = 0; k < count; ++k) {
    const int i = order[k];
    const DcnGeom& g = geoms[i];
    const kgdet_dcn_group_item& it = items[i];
    GroupProblem& P = gp.prob[k];
    const size_t guard = (size_t)dcn_guard_pixels(g) * 128;
    const size_t in_bytes = (size_t)g.N * g.H * g.W * 128;
    P.in = (const char*)it.prepared_input + guard;
    P.plane_bytes = align_up(in_bytes + 2 * guard, 1024);
    P.plan = (const SampleRec16*)it.plan;
    P.wp = (const unsigned char*)it.weight_packed;
    P.bias = it.bias;
    P.out = it.output;
    P.M = g.M; P.W = g.W; P.K = g.K; P.HoWo = g.Ho * g.Wo; P.rows_padded = (int)plan_rows(g);
    P.out_coff = it.out_channel_offset; P.out_ctot = it.out_channels_total; P.relu = it.fuse_relu ? 1 : 0;
    P.out_layout = it.out_layout; P.out_dtype = it.dtype;
    P.nkb = (g.C / 64) * g.K;
    P.tile_begin = tiles;
    tiles += ceil_div(g.M, BM);
  }
  if (geoms[0].Cout != 256 || ceil_div(tiles, 2) < 2 * num_sms()) rows256 = false;
  if (rows256) {
    tiles = 0;
    for (int k = 0; k < count; ++k) {
      gp.prob[k].tile_begin = tiles;
      tiles += ceil_div(geoms[order[k]].M, 2 * BM);
    }
  }
  for (int k = count; k < KGDET_DCN_GROUP_MAX; ++k) gp.prob[k] = gp.prob[count - 1];
  gp.total_tiles = tiles;
  gp.Cout = geoms[0].Cout;
  gp.idesc = make_idesc(1u, BM, (uint32_t)gp.Cout);
  gp.tmem_cols = gp.Cout <= 64 ? 64 : (gp.Cout <= 128 ? 128 : 256);
  gp.timeline = nullptr;
  if (g_timeline && g_timeline_entries >= (long long)num_sms() * 64 * 8) gp.timeline = g_timeline;
  g_timeline = nullptr;
  cudaEvent_t ev0 = g_group_prof_start, ev1 = g_group_prof_stop;
  g_group_prof_start = g_group_prof_stop = nullptr;
  if (ev0 && ev1) KG_CUDA(cudaEventRecord(ev0, stream));
  int rc = umma_group_forward(gp, stream, rows256);
  if (rc == KGDET_OK && ev0 && ev1) KG_CUDA(cudaEventRecord(ev1, stream));
  return rc;
}
