// Greedy NMS on the device, bit-exact with the reference's arithmetic.
//
// Replaces mmdet/ops/nms/src/nms_kernel.cu:13-131 (bitmask kernel + HOST sweep after a
// blocking D2H copy) and mmdet/ops/nms/src/nms_cpu.cpp:4-59.  Differences by design:
//  * the greedy sweep runs on the device (no D2H of the mask, no host loop, no sync);
//  * small problems (n <= kSmallMax, every KGDet case: n <= nms_pre = 1000 per class) run
//    as ONE CTA per (image, class) segment: in-CTA bitonic sort, then 64-box blocks are
//    resolved with a 64x64 diagonal bitmask and broadcast to the remaining boxes -- the
//    n x n/64 mask never exists in memory;
//  * large problems use upper-triangular 64x64 mask tiles in the workspace plus a
//    single-CTA sweep that only reads the rows of kept boxes.
// IoU uses explicitly rounded fp32 intrinsics in the reference's operation order so the
// compiler cannot contract (Sa + Sb) - w*h into an FMA (nms_cpu.cpp:47-54).
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace kgdet {

static constexpr int kSmallMax = 4096;   // boxes per segment for the single-CTA path
static constexpr int kSmallThreads = 1024;

struct Box { float x1, y1, x2, y2; };

__device__ __forceinline__ float box_area(const Box& b) {
  // (x2 - x1 + 1) * (y2 - y1 + 1)            nms_cpu.cpp:18, nms_kernel.cu:18-19
  return __fmul_rn(__fadd_rn(__fsub_rn(b.x2, b.x1), 1.f), __fadd_rn(__fsub_rn(b.y2, b.y1), 1.f));
}

// true when box b (lower score) is suppressed by box a (higher score).
// The decision of the reference is rn(inter / union) {>, >=} thr.  IEEE division is ~10x the cost of
// the rest of the test, so pairs that are far from the threshold are decided by a multiplication with
// a 1e-6 guard band (well above the ~2e-7 worst-case relative error of the three roundings involved);
// only pairs inside the band take the exact division -- the result is bit-identical either way.
__device__ __forceinline__ bool suppresses(const Box& a, float area_a, const Box& b, float area_b,
                                           float thr, int cmp_ge) {
  float xx1 = fmaxf(a.x1, b.x1), yy1 = fmaxf(a.y1, b.y1);
  float xx2 = fminf(a.x2, b.x2), yy2 = fminf(a.y2, b.y2);
  float w = fmaxf(__fadd_rn(__fsub_rn(xx2, xx1), 1.f), 0.f);
  float h = fmaxf(__fadd_rn(__fsub_rn(yy2, yy1), 1.f), 0.f);
  float inter = __fmul_rn(w, h);
  float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
  if (uni > 0.f && thr > 0.f) {
    const float t = __fmul_rn(thr, uni);
    if (inter > __fmul_rn(t, 1.000001f)) return true;
    if (inter < __fmul_rn(t, 0.999999f)) return false;
  }
  float ovr = __fdiv_rn(inter, uni);
  return cmp_ge ? (ovr >= thr) : (ovr > thr);
}

// (score desc, index asc) ordering: "a goes before b"
__device__ __forceinline__ bool before(float sa, int ia, float sb, int ib) {
  return (sa > sb) || (sa == sb && ia < ib);
}

// ---------------------------------------------------------------------------------------
// Single-CTA NMS of one segment.  Dynamic smem layout (P = next pow2 >= n):
//   float key[P]; int idx[P]; Box box[n]; float area[n]; u64 removed[ceil(n/64)];
//   u64 diag[64]; u64 kept_word;
// Output: flags[orig_row] = 1/0.
__global__ void __launch_bounds__(kSmallThreads, 1)
nms_small_kernel(const float* __restrict__ dets, const int32_t* __restrict__ seg_offsets,
                 int single_n, float thr, int cmp_ge, float score_thr, uint8_t* __restrict__ flags,
                 int P_cap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int n_valid;
  int row0, n;
  if (seg_offsets) {
    row0 = seg_offsets[blockIdx.x];
    n = seg_offsets[blockIdx.x + 1] - row0;
  } else {                                   // uniform segments of single_n rows (dense mode)
    row0 = blockIdx.x * single_n;
    n = single_n;
  }
  if (n <= 0) return;
  if (threadIdx.x == 0) n_valid = 0;
  int P = 1;
  while (P < n) P <<= 1;
  // carve
  float* key = reinterpret_cast<float*>(smem_raw);
  int* idx = reinterpret_cast<int*>(key + P_cap);
  Box* box = reinterpret_cast<Box*>(idx + P_cap);
  float* area = reinterpret_cast<float*>(box + P_cap);
  unsigned long long* removed = reinterpret_cast<unsigned long long*>(area + P_cap);
  unsigned long long* diag = removed + (P_cap + 63) / 64;
  unsigned long long* kept_word = diag + 64;

  const int tid = threadIdx.x, nt = blockDim.x;
  const float* d = dets + (size_t)row0 * 5;
  __syncthreads();
  {
    // rows whose score does not exceed score_thr are absent (multiclass_nms_kp's `scores > score_thr`
    // filter, bbox_nms_kp.py:39, folded into the op so that the caller needs no compaction).  Present rows are
    // compacted to the front (slot order is irrelevant: the sort key (score, index) is a total order), so the
    // sort below runs on the next power of two above the PRESENT count, not above the segment length -- with
    // ~10 % of 1000 candidates above the threshold that is 128 elements and 28 passes instead of 1024 and 55.
    for (int i = tid; i < n; i += nt) {
      const float v = d[i * 5 + 4];
      if (v > score_thr) {
        const int slot = atomicAdd(&n_valid, 1);
        key[slot] = v;
        idx[slot] = i;
      } else {
        flags[row0 + i] = 0;
      }
    }
  }
  __syncthreads();
  P = 32;
  while (P < n_valid) P <<= 1;
  for (int i = n_valid + tid; i < P; i += nt) { key[i] = -INFINITY; idx[i] = 0x7fffffff; }
  __syncthreads();
  // bitonic sort, ascending in the `before` order
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < P; i += nt) {
        int l = i ^ j;
        if (l > i) {
          float si = key[i], sl = key[l];
          int ii = idx[i], il = idx[l];
          bool up = ((i & k) == 0);
          bool swap = up ? before(sl, il, si, ii) : before(si, ii, sl, il);
          if (swap) { key[i] = sl; key[l] = si; idx[i] = il; idx[l] = ii; }
        }
      }
      __syncthreads();
    }
  }
  // only the first n_valid slots hold rows (absent rows were flagged 0 above)
  n = n_valid;
  if (n == 0) return;
  // gather boxes in sorted order
  for (int i = tid; i < n; i += nt) {
    const float* s = d + (size_t)idx[i] * 5;
    Box b{s[0], s[1], s[2], s[3]};
    box[i] = b;
    area[i] = box_area(b);
  }
  const int nwords = (n + 63) / 64;
  for (int i = tid; i < nwords; i += nt) removed[i] = 0ull;
  __syncthreads();

  for (int blk = 0; blk < nwords; ++blk) {
    const int base = blk * 64;
    const int cnt = min(64, n - base);
    // 64x64 diagonal bitmask, all 1024 threads: thread t tests row t/16 against 4 columns and the 16
    // partial words of a row are merged with a shared-memory atomicOr
    if (tid < 64) diag[tid] = 0ull;
    __syncthreads();
    {
      const int r = tid >> 4, c0 = (tid & 15) * 4;
      if (r < cnt) {
        const Box a = box[base + r];
        const float aa = area[base + r];
        unsigned long long bits = 0ull;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int c = c0 + k;
          if (c > r && c < cnt && suppresses(a, aa, box[base + c], area[base + c], thr, cmp_ge))
            bits |= 1ull << c;
        }
        if (bits) atomicOr(&diag[r], bits);
      }
    }
    __syncthreads();
    // serial resolve of the block by warp 0: the 64 diagonal words live in registers (two per lane)
    // and are broadcast by shuffles that do not depend on the running `rem`, so only a short ALU
    // chain is serial
    if (tid < 32) {
      const unsigned long long d_lo = diag[tid], d_hi = diag[tid + 32];
      unsigned long long rem = removed[blk], kept = 0ull;
#pragma unroll
      for (int r = 0; r < 64; ++r) {
        const unsigned long long dr = __shfl_sync(0xffffffffu, r < 32 ? d_lo : d_hi, r & 31);
        if (r < cnt && !((rem >> r) & 1ull)) { kept |= 1ull << r; rem |= dr; }
      }
      __syncwarp();                              // every lane has read removed[blk] before lane 0 rewrites it
      if (tid == 0) { removed[blk] = rem; *kept_word = kept; }
    }
    __syncthreads();
    const unsigned long long kept = *kept_word;
    // kept boxes of this block suppress every later, still-alive box
    for (int j = base + 64 + tid; j < n; j += nt) {
      if ((removed[j >> 6] >> (j & 63)) & 1ull) continue;
      Box b = box[j];
      float ab = area[j];
      unsigned long long kk = kept;
      bool dead = false;
      while (kk && !dead) {
        int r = __ffsll((long long)kk) - 1;
        kk &= kk - 1;
        dead = suppresses(box[base + r], area[base + r], b, ab, thr, cmp_ge);
      }
      if (dead) atomicOr(&removed[j >> 6], 1ull << (j & 63));
    }
    __syncthreads();
  }
  for (int i = tid; i < n; i += nt)
    flags[row0 + idx[i]] = ((removed[i >> 6] >> (i & 63)) & 1ull) ? 0 : 1;
}

// flags[n] (0/1) -> ascending indices + count.  One CTA, chunked block scan.
__global__ void __launch_bounds__(1024, 1)
compact_flags_kernel(const uint8_t* __restrict__ flags, int n, int64_t* __restrict__ keep,
                     int32_t* __restrict__ num_keep) {
  __shared__ int warp_tot[32];
  __shared__ int carry;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    int i = base + tid;
    int f = (i < n) ? flags[i] : 0;
    unsigned bal = __ballot_sync(0xffffffffu, f);
    int pre = __popc(bal & ((1u << lane) - 1));
    if (lane == 0) warp_tot[wid] = __popc(bal);
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < wid; ++w) woff += warp_tot[w];
    int c = carry;
    if (f) keep[c + woff + pre] = i;
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int w = 0; w < 32; ++w) t += warp_tot[w];
      carry = c + t;
    }
    __syncthreads();
  }
  if (tid == 0) *num_keep = carry;
}

// ---------------------------------------------------------------------------------------
// Large-n path: CUB radix sort -> upper-triangular mask tiles -> single-CTA sweep.
__global__ void nms_prepare_sort_kernel(const float* __restrict__ dets, int n, float* keys,
                                        int* vals) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { keys[i] = dets[i * 5 + 4]; vals[i] = i; }
}

__global__ void nms_gather_sorted_kernel(const float* __restrict__ dets, const int* order, int n,
                                         Box* box, float* area) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float* s = dets + (size_t)order[i] * 5;
    Box b{s[0], s[1], s[2], s[3]};
    box[i] = b;
    area[i] = box_area(b);
  }
}

// grid (col_blocks, col_blocks), 64 threads; tiles below the diagonal exit (never read).
__global__ void nms_mask_kernel(const Box* __restrict__ box, const float* __restrict__ area, int n,
                                float thr, int cmp_ge, unsigned long long* __restrict__ mask,
                                int col_blocks) {
  const int rb = blockIdx.y, cb = blockIdx.x;
  if (cb < rb) return;
  __shared__ Box sbox[64];
  __shared__ float sarea[64];
  const int t = threadIdx.x;
  const int csize = min(64, n - cb * 64), rsize = min(64, n - rb * 64);
  if (t < csize) { sbox[t] = box[cb * 64 + t]; sarea[t] = area[cb * 64 + t]; }
  __syncthreads();
  if (t < rsize) {
    const int r = rb * 64 + t;
    Box a = box[r];
    float aa = area[r];
    unsigned long long bits = 0ull;
    int start = (rb == cb) ? t + 1 : 0;
    for (int c = start; c < csize; ++c)
      if (suppresses(a, aa, sbox[c], sarea[c], thr, cmp_ge)) bits |= 1ull << c;
    mask[(size_t)r * col_blocks + cb] = bits;
  }
}

__global__ void __launch_bounds__(1024, 1)
nms_sweep_kernel(const unsigned long long* __restrict__ mask, const int* __restrict__ order, int n,
                 int col_blocks, uint8_t* __restrict__ flags) {
  extern __shared__ unsigned long long remv[];  // col_blocks words
  __shared__ unsigned long long kept_word;
  __shared__ unsigned long long diag[64];
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < col_blocks; i += nt) remv[i] = 0ull;
  __syncthreads();
  for (int blk = 0; blk < col_blocks; ++blk) {
    const int base = blk * 64, cnt = min(64, n - base);
    if (tid < cnt) diag[tid] = mask[(size_t)(base + tid) * col_blocks + blk];
    __syncthreads();
    if (tid == 0) {
      unsigned long long rem = remv[blk], kept = 0ull;
      for (int r = 0; r < cnt; ++r)
        if (!((rem >> r) & 1ull)) {
          kept |= 1ull << r;
          rem |= diag[r];
        }
      remv[blk] = rem;
      kept_word = kept;
    }
    __syncthreads();
    const unsigned long long kept = kept_word;
    for (int w = blk + 1 + tid; w < col_blocks; w += nt) {
      unsigned long long acc = 0ull, kk = kept;
      while (kk) {
        int r = __ffsll((long long)kk) - 1;
        kk &= kk - 1;
        acc |= mask[(size_t)(base + r) * col_blocks + w];
      }
      remv[w] |= acc;
    }
    __syncthreads();
  }
  for (int i = tid; i < n; i += nt)
    flags[order[i]] = ((remv[i >> 6] >> (i & 63)) & 1ull) ? 0 : 1;
}

// Pipelined sweep.  The greedy pass is a chain over the 64-box blocks: block b can only be resolved when every kept
// box of the blocks before it has been OR-ed into word b of `remv`.  nms_sweep_kernel above runs that chain and the
// row ORs in lock step (three CTA barriers and one exposed round of global loads per block: 3.4 us per block at
// n = 65 536).  Here warp 0 runs ONLY the chain: the diagonal word and the next TWO columns of the block's 64 rows are
// requested one block ahead (they do not depend on the chain), the greedy pass takes one iteration per KEPT box in
// every lane at once (next alive box = lowest clear bit, its word broadcast by a shuffle), and the contributions of blocks b - 1 and b - 2 to word b come from those prefetched
// columns filtered by the keep words.  The other 31 warps apply every published keep word to the later words
// (words >= a + 3 of block a); each applier warp owns a contiguous range of words and reports its own progress, so
// the chain only waits for the owner of the word it is about to read, and only if that warp is more than two
// blocks behind.
static constexpr int kSweepAppliers = 31;
__global__ void __launch_bounds__(1024, 1)
nms_sweep_pipelined_kernel(const unsigned long long* __restrict__ mask, const int* __restrict__ order, int n,
                           int col_blocks, uint8_t* __restrict__ flags) {
  extern __shared__ unsigned long long sweep_smem[];
  volatile unsigned long long* remv = sweep_smem;                     // [col_blocks]
  volatile unsigned long long* kept_words = sweep_smem + col_blocks;  // [col_blocks]
  __shared__ volatile int published;
  __shared__ volatile int progress[kSweepAppliers];                   // passes completed by each applier warp
  const int tid = threadIdx.x, lane = tid & 31;
  const int wpw = (col_blocks + kSweepAppliers - 1) / kSweepAppliers; // words per applier warp
  for (int i = tid; i < col_blocks; i += blockDim.x) remv[i] = 0ull;
  if (tid == 0) published = 0;
  if (tid < kSweepAppliers) progress[tid] = 0;
  __syncthreads();
  if (tid < 32) {
    // ---- the chain (warp 0) ----
    // d: diagonal words of rows lane / lane + 32 of the block; x, y: the same rows' words of the next two columns
    unsigned long long d0, d1, x0, x1, y0, y1, nd0, nd1, nx0, nx1, ny0, ny1;
    auto load_block = [&](int blk, unsigned long long& a0, unsigned long long& a1, unsigned long long& b0,
                          unsigned long long& b1, unsigned long long& c0, unsigned long long& c1) {
      a0 = a1 = b0 = b1 = c0 = c1 = 0ull;
      if (blk >= col_blocks) return;
      const int base = blk * 64, cnt = min(64, n - base);
      const bool has1 = blk + 1 < col_blocks, has2 = blk + 2 < col_blocks;
      if (lane < cnt) {
        const unsigned long long* row = mask + (size_t)(base + lane) * col_blocks + blk;
        a0 = __ldg(row);
        if (has1) b0 = __ldg(row + 1);
        if (has2) c0 = __ldg(row + 2);
      }
      if (lane + 32 < cnt) {
        const unsigned long long* row = mask + (size_t)(base + lane + 32) * col_blocks + blk;
        a1 = __ldg(row);
        if (has1) b1 = __ldg(row + 1);
        if (has2) c1 = __ldg(row + 2);
      }
    };
    auto warp_or = [&](unsigned long long v) {
      const unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)v);
      const unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)(v >> 32));
      return ((unsigned long long)hi << 32) | lo;
    };
    load_block(0, d0, d1, x0, x1, y0, y1);
    unsigned long long carry1 = 0ull, carry2 = 0ull, carry2_next = 0ull;   // from block b - 1 / b - 2 at word b
    for (int blk = 0; blk < col_blocks; ++blk) {
      const int cnt = min(64, n - blk * 64);
      load_block(blk + 1, nd0, nd1, nx0, nx1, ny0, ny1);
      if (blk >= 3) {
        const int owner = blk / wpw;
        while (progress[owner] < blk - 2) { }                         // blocks <= blk - 3 are in word blk
      }
      unsigned long long rem = remv[blk] | carry1 | carry2, kept = 0ull;
      // greedy pass over the block, one iteration per KEPT box (every lane runs it on the same values): the next
      // alive box is the lowest clear bit of `rem` above the last kept one
      const unsigned long long valid = cnt == 64 ? ~0ull : ((1ull << cnt) - 1ull);
      unsigned long long above = ~0ull;
      while (true) {
        const unsigned long long cand = ~rem & valid & above;
        if (!cand) break;
        const int r = __ffsll((long long)cand) - 1;
        kept |= 1ull << r;
        rem |= __shfl_sync(0xffffffffu, r < 32 ? d0 : d1, r & 31);
        above = r == 63 ? 0ull : (~0ull << (r + 1));
      }
      if (lane == 0) {
        remv[blk] = rem;
        kept_words[blk] = kept;
        __threadfence_block();
        published = blk + 1;
      }
      const bool k0 = (kept >> lane) & 1ull, k1 = (kept >> (lane + 32)) & 1ull;
      carry1 = warp_or((k0 ? x0 : 0ull) | (k1 ? x1 : 0ull));          // this block at word blk + 1
      carry2 = carry2_next;                                           // block blk - 1 at word blk + 1
      carry2_next = warp_or((k0 ? y0 : 0ull) | (k1 ? y1 : 0ull));     // this block at word blk + 2
      d0 = nd0; d1 = nd1; x0 = nx0; x1 = nx1; y0 = ny0; y1 = ny1;
    }
  } else {
    // ---- the appliers (warps 1 .. 31): warp j owns words [j * wpw, (j + 1) * wpw) ----
    const int j = (tid >> 5) - 1;
    const int lo = j * wpw, hi = min(col_blocks, lo + wpw);
    for (int a = 0; a < col_blocks && a + 3 < hi; ++a) {
      if (lane == 0)                                                  // one polling lane per warp, with pauses: polls
        while (published <= a) { }                                    // share the LDS / shuffle pipe with the chain
      __syncwarp();
      __threadfence_block();
      const unsigned long long kept = kept_words[a];
      const size_t rbase = (size_t)a * 64;
      if (kept) {
        for (int w = max(lo, a + 3) + lane; w < hi; w += 32) {
          unsigned long long acc = 0ull, kk = kept;
          while (kk) {
            const int r = __ffsll((long long)kk) - 1;
            kk &= kk - 1;
            acc |= __ldg(mask + (rbase + r) * col_blocks + w);
          }
          if (acc) remv[w] |= acc;                                    // one writer per word
        }
      }
      __threadfence_block();
      __syncwarp();
      if (lane == 0) progress[j] = a + 1;
    }
    __syncwarp();
    if (lane == 0 && j < kSweepAppliers) progress[j] = 0x7fffffff;    // nothing left to apply to this warp's words
  }
  __syncthreads();
  for (int i = tid; i < n; i += blockDim.x)
    flags[order[i]] = ((remv[i >> 6] >> (i & 63)) & 1ull) ? 0 : 1;
}

static size_t small_smem_bytes(int P_cap) {
  return (size_t)P_cap * (4 + 4 + 16 + 4) + (size_t)((P_cap + 63) / 64) * 8 + 64 * 8 + 16;
}
static int pow2_cap(int n) {
  int P = 64;
  while (P < n) P <<= 1;
  return P;
}

struct LargeWs {
  float *keys_in, *keys_out, *area;
  int *vals_in, *vals_out;
  Box* box;
  unsigned long long* mask;
  uint8_t* flags;
  void* cub;
  size_t cub_bytes, total;
};
static LargeWs carve_large(void* ws, int n) {
  LargeWs w;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* p = ws ? (void*)((char*)ws + off) : nullptr;
    off += align_up(bytes, 256);
    return p;
  };
  int cb = (n + 63) / 64;
  w.keys_in = (float*)take((size_t)n * 4);
  w.keys_out = (float*)take((size_t)n * 4);
  w.vals_in = (int*)take((size_t)n * 4);
  w.vals_out = (int*)take((size_t)n * 4);
  w.box = (Box*)take((size_t)n * 16);
  w.area = (float*)take((size_t)n * 4);
  w.flags = (uint8_t*)take((size_t)n);
  w.mask = (unsigned long long*)take((size_t)n * cb * 8);
  w.cub_bytes = 0;
  cub::DeviceRadixSort::SortPairsDescending(nullptr, w.cub_bytes, (float*)nullptr, (float*)nullptr,
                                            (int*)nullptr, (int*)nullptr, n);
  w.cub = take(w.cub_bytes);
  w.total = off;
  return w;
}

}  // namespace kgdet

using namespace kgdet;

extern "C" size_t kgdet_nms_workspace_bytes(int32_t n) {
  if (n <= 0) return 256;
  if (n <= kSmallMax) return align_up((size_t)n, 256);  // flags only
  return carve_large(nullptr, n).total;
}

extern "C" int kgdet_nms(const float* dets, int32_t n, float iou_thr, int cmp_mode, int64_t* keep,
                         int32_t* num_keep, void* workspace, size_t workspace_bytes,
                         void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  KG_CHECK_ARG(n >= 0, "kgdet_nms: n must be >= 0 (got %d)", n);
  KG_CHECK_ARG(num_keep != nullptr, "kgdet_nms: num_keep is NULL");
  KG_CHECK_ARG(cmp_mode == KGDET_NMS_GT || cmp_mode == KGDET_NMS_GE, "kgdet_nms: bad cmp_mode %d",
               cmp_mode);
  if (n == 0) {  // nms_wrapper.py:39-40 / nms_cuda.cpp:10-11: empty in, empty out
    KG_CUDA(cudaMemsetAsync(num_keep, 0, sizeof(int32_t), stream));
    return KGDET_OK;
  }
  KG_CHECK_ARG(dets && keep, "kgdet_nms: NULL dets/keep");
  if (workspace_bytes < kgdet_nms_workspace_bytes(n) || !workspace) {
    set_error("kgdet_nms: workspace too small (%zu < %zu)", workspace_bytes,
              kgdet_nms_workspace_bytes(n));
    return KGDET_ERR_WORKSPACE;
  }
  if (n <= kSmallMax) {
    uint8_t* flags = (uint8_t*)workspace;
    int P_cap = pow2_cap(n);
    size_t smem = small_smem_bytes(P_cap);
    KG_CUDA(cudaFuncSetAttribute(nms_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
    nms_small_kernel<<<1, kSmallThreads, smem, stream>>>(dets, nullptr, n, iou_thr,
                                                         cmp_mode == KGDET_NMS_GE, -INFINITY, flags, P_cap);
    KG_LAUNCH_CHECK("nms_small_kernel");
    compact_flags_kernel<<<1, 1024, 0, stream>>>(flags, n, keep, num_keep);
    KG_LAUNCH_CHECK("compact_flags_kernel");
    return KGDET_OK;
  }
  LargeWs w = carve_large(workspace, n);
  const int cb = (n + 63) / 64;
  nms_prepare_sort_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(dets, n, w.keys_in, w.vals_in);
  KG_LAUNCH_CHECK("nms_prepare_sort_kernel");
  size_t cub_bytes = w.cub_bytes;
  KG_CUDA(cub::DeviceRadixSort::SortPairsDescending(w.cub, cub_bytes, w.keys_in, w.keys_out,
                                                    w.vals_in, w.vals_out, n, 0, 32, stream));
  nms_gather_sorted_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(dets, w.vals_out, n, w.box,
                                                                 w.area);
  KG_LAUNCH_CHECK("nms_gather_sorted_kernel");
  nms_mask_kernel<<<dim3(cb, cb), 64, 0, stream>>>(w.box, w.area, n, iou_thr,
                                                   cmp_mode == KGDET_NMS_GE, w.mask, cb);
  KG_LAUNCH_CHECK("nms_mask_kernel");
  bool pipelined = true;
  if (const char* e = getenv("KGDET_NMS_SWEEP_PIPELINED")) pipelined = atoi(e) != 0;
  if (pipelined && (size_t)cb * 16 <= 200 * 1024) {
    const size_t sweep_smem = (size_t)cb * 16;
    KG_CUDA(cudaFuncSetAttribute(nms_sweep_pipelined_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sweep_smem));
    nms_sweep_pipelined_kernel<<<1, 1024, sweep_smem, stream>>>(w.mask, w.vals_out, n, cb, w.flags);
    KG_LAUNCH_CHECK("nms_sweep_pipelined_kernel");
  } else {
    const size_t sweep_smem = (size_t)cb * 8;
    KG_CUDA(cudaFuncSetAttribute(nms_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sweep_smem));
    nms_sweep_kernel<<<1, 1024, sweep_smem, stream>>>(w.mask, w.vals_out, n, cb, w.flags);
    KG_LAUNCH_CHECK("nms_sweep_kernel");
  }
  compact_flags_kernel<<<1, 1024, 0, stream>>>(w.flags, n, keep, num_keep);
  KG_LAUNCH_CHECK("compact_flags_kernel");
  return KGDET_OK;
}

extern "C" size_t kgdet_nms_batched_workspace_bytes(int32_t, int32_t, int32_t) { return 256; }

extern "C" int kgdet_nms_batched(const float* dets, const int32_t* seg_offsets, int32_t nseg, int32_t total,
                                 int32_t max_seg_len, float iou_thr, float score_thr, int cmp_mode,
                                 uint8_t* keep_flags, void*, size_t, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  KG_CHECK_ARG(nseg >= 0 && total >= 0, "kgdet_nms_batched: negative sizes");
  KG_CHECK_ARG(cmp_mode == KGDET_NMS_GT || cmp_mode == KGDET_NMS_GE,
               "kgdet_nms_batched: bad cmp_mode %d", cmp_mode);
  if (nseg == 0 || total == 0) return KGDET_OK;
  KG_CHECK_ARG(dets && keep_flags, "kgdet_nms_batched: NULL pointer");
  KG_CHECK_ARG(seg_offsets || (long long)nseg * max_seg_len == total,
               "kgdet_nms_batched: dense mode needs total == nseg * max_seg_len");
  if (max_seg_len > kSmallMax) {
    set_error("kgdet_nms_batched: max_seg_len %d exceeds the single-CTA limit %d; call kgdet_nms "
              "per segment", max_seg_len, kSmallMax);
    return KGDET_ERR_UNSUPPORTED;
  }
  int P_cap = pow2_cap(max_seg_len < 1 ? 1 : max_seg_len);
  size_t smem = small_smem_bytes(P_cap);
  KG_CUDA(cudaFuncSetAttribute(nms_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)smem));
  // more segments than SMs (a KGDet batch: 16 images x 13 classes = 208): 512-thread CTAs so that two fit an SM and
  // the launch is ONE wave (1024-thread CTAs, one per SM, ran two waves for 208 segments)
  const int threads = (nseg > num_sms() && max_seg_len <= 2048) ? 512 : kSmallThreads;
  nms_small_kernel<<<nseg, threads, smem, stream>>>(dets, seg_offsets, seg_offsets ? 0 : max_seg_len,
                                                          iou_thr, cmp_mode == KGDET_NMS_GE, score_thr,
                                                          keep_flags, P_cap);
  KG_LAUNCH_CHECK("nms_small_kernel(batched)");
  return KGDET_OK;
}
