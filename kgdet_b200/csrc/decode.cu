// Post-head decode around the batched NMS: the candidate selection, box /
// keypoint decode, clamp and result gather of get_bboxes_single + multiclass_nms_kp
// (reppoints_head_kp3rep_cas_1_assign_once.py:843-903, core/post_processing/bbox_nms_kp.py:6-75) as three
// kernels with static shapes and no host synchronisation.  The reference does this with ~40 PyTorch kernels
// per image and gathers the 588-value keypoint vector of every one of the nms_pre candidates; here keypoints
// are decoded only for the max_per_img detections that survive.
//
// Arithmetic mirrors the PyTorch expressions operation by operation (separate mul / add roundings, no FMA
// contraction) so that boxes and scores are bit-identical to the reference path and NMS sees identical inputs.
#include "common.cuh"

namespace kgdet {

__device__ __forceinline__ float sigmoid_ref(float x) { return 1.f / (1.f + expf(-x)); }   // ATen sigmoid

// ---- 1. candidate selection: order[b, r] = position with the r-th largest max-over-classes score -------------
// (KP3:863-874: max_scores.topk(nms_pre); ties broken by ascending position).  Rank by counting: O(HW^2) per
// image, fine for the <= 4096 positions of a head level; order is ascending identity when HW <= nms_pre.
// Every CTA recomputes the per-position maxima of its image (ceil(HW / 256) CTAs per image): one 1024-thread CTA
// per image that computes them once was measured slower (50 vs 28 us for 16 x 1050 positions: 16 CTAs cannot
// hide the latency of the strided score reads).
// Ranking key: a TOTAL order on floats as unsigned integers (monotone bit pattern; -0 == +0; NaN above +inf,
// which is where torch.topk / torch.max put it -- KP3:866-868).  With plain `>` / `==` comparisons every
// all-NaN position (a diverged feature map) would get rank 0, ranks would collide and slots of `order` would
// stay unwritten.
__device__ __forceinline__ unsigned int rank_key(float v) {
  unsigned int b = __float_as_uint(v);
  if (v != v) return 0xFFFFFFFFu;
  if (b == 0x80000000u) b = 0u;
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__global__ void __launch_bounds__(256) bbox_select_kernel(const float* __restrict__ scores, int apply_sigmoid, int C,
                                                          int HW, int n, int* __restrict__ order) {
  extern __shared__ __align__(16) unsigned int skey[];        // [HW rounded up to 4] rank keys of the per-position maxima
  const int b = blockIdx.y;
  const float* sb = scores + (size_t)b * C * HW;
  const int HW4 = (HW + 3) & ~3;
  for (int p = threadIdx.x; p < HW4; p += blockDim.x) {
    unsigned int m = 0u;                        // below rank_key(-inf); the padding keys never outrank a real one
    if (p < HW) {
      for (int c0 = 0; c0 < C; c0 += 8) {       // eight classes in flight before the first is used
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (c0 + u < C) ? sb[(size_t)(c0 + u) * HW + p] : 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (c0 + u < C) m = max(m, rank_key(apply_sigmoid ? sigmoid_ref(v[u]) : v[u]));   // NaN propagates, like torch.max
        }
      }
    }
    skey[p] = m;
  }
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  if (n >= HW) {                                // no top-k in the reference: original order
    order[(size_t)b * n + p] = p;
    return;
  }
  const unsigned int mine = skey[p];
  int rank = 0;
  for (int j = 0; j < HW4; j += 4) {            // four keys per shared-memory load
    const uint4 o = *reinterpret_cast<const uint4*>(skey + j);
    rank += (o.x > mine || (o.x == mine && j < p)) ? 1 : 0;
    rank += (o.y > mine || (o.y == mine && j + 1 < p)) ? 1 : 0;
    rank += (o.z > mine || (o.z == mine && j + 2 < p)) ? 1 : 0;
    rank += (o.w > mine || (o.w == mine && j + 3 < p)) ? 1 : 0;
  }
  if (rank < n) order[(size_t)b * n + rank] = p;
}

// Large levels (FPN P3 / P4 of the RepPoints-Kp heads: 16 800 / 4 200 positions, nms_pre = 1000): counting ranks over
// all positions is O(HW^2).  Three kernels instead:
//   keys     the per-position maxima's rank keys, all SMs (one CTA per image cannot hide the latency of the
//            13 strided score reads per position: 150 us for 8 x 16 800 positions when it was part of `select`);
//   select   one CTA per image: the n-th largest key by a three-pass radix select (11 + 11 + 10 bits, histograms
//            in shared memory), then an ordered compaction of the n selected positions (every key above it, then
//            the first positions that equal it -- the same tie rule as above) into a list in position order;
//   rank     ranks by counting inside that list, O(n^2) spread over ceil(n / 128) CTAs per image.
// Same `order` as the kernel above, bit for bit.
constexpr int kSelThreads = 1024;
constexpr int kSelMaxHW = 40960;
constexpr int kSelMaxN = 4096;

__global__ void __launch_bounds__(256) level_keys_kernel(const float* __restrict__ scores, int apply_sigmoid, int C,
                                                         int HW, unsigned int* __restrict__ keys) {
  const int b = blockIdx.y, p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  const float* sb = scores + (size_t)b * C * HW + p;
  unsigned int m = 0u;
  for (int c0 = 0; c0 < C; c0 += 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = (c0 + u < C) ? sb[(size_t)(c0 + u) * HW] : 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (c0 + u < C) m = max(m, rank_key(apply_sigmoid ? sigmoid_ref(v[u]) : v[u]));
    }
  }
  keys[(size_t)b * HW + p] = m;
}

__global__ void __launch_bounds__(kSelThreads) bbox_select_radix_kernel(const unsigned int* __restrict__ keys, int HW,
                                                                        int n, unsigned int* __restrict__ list_key,
                                                                        int* __restrict__ list_pos) {
  extern __shared__ __align__(16) unsigned int sel_sm[];
  unsigned int* skey = sel_sm;                          // [HW]
  unsigned int* hist = skey + ((HW + 3) & ~3);          // [2048]
  __shared__ unsigned int s_prefix, s_need;
  __shared__ int warp_cnt[2][kSelThreads / 32];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n4 = (n + 3) & ~3;
  unsigned int* lkey = list_key + (size_t)b * n4;       // keys of the selected positions, ascending position
  int* lpos = list_pos + (size_t)b * n4;
  for (int p = tid; p < HW; p += kSelThreads) skey[p] = keys[(size_t)b * HW + p];
  for (int i = n + tid; i < n4; i += kSelThreads) lkey[i] = 0u;     // padding keys never outrank a real one
  unsigned int prefix = 0u, mask = 0u, need = (unsigned int)n;
#pragma unroll 1
  for (int pass = 0; pass < 3; ++pass) {
    const int shift = pass == 0 ? 21 : (pass == 1 ? 10 : 0), bits = pass == 2 ? 10 : 11, nb = 1 << bits;
    for (int i = tid; i < 2048; i += kSelThreads) hist[i] = 0u;
    __syncthreads();                                  // also: skey complete (first pass)
    for (int p = tid; p < HW; p += kSelThreads) {
      const unsigned int k = skey[p];
      if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & (nb - 1)], 1u);
    }
    __syncthreads();
    if (warp == 0) {
      // lane l owns the bins [nb - (l + 1) * per, nb - l * per): lane 0 the largest keys
      const int per = nb / 32, hi = nb - lane * per;
      unsigned int sum = 0u;
      for (int i = 1; i <= per; ++i) sum += hist[hi - i];
      unsigned int incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      unsigned int cum = incl - sum;                  // keys in the bins above mine
      if (cum < need && need <= incl) {
        for (int i = 1; i <= per; ++i) {
          const unsigned int h = hist[hi - i];
          if (cum + h >= need) {
            s_prefix = prefix | ((unsigned int)(hi - i) << shift);
            s_need = need - cum;
            break;
          }
          cum += h;
        }
      }
    }
    __syncthreads();
    prefix = s_prefix;
    need = s_need;
    mask |= (unsigned int)(nb - 1) << shift;
  }
  // prefix = the n-th largest key; `need` positions (>= 1) that equal it are taken, lowest positions first
  int base_eq = 0, base_sel = 0;
#pragma unroll 1
  for (int p0 = 0; p0 < HW; p0 += kSelThreads) {
    const int p = p0 + tid;
    const unsigned int k = p < HW ? skey[p] : 0u;
    const bool gt = p < HW && k > prefix, eq = p < HW && k == prefix;
    const unsigned int be = __ballot_sync(0xffffffffu, eq);
    if (lane == 0) warp_cnt[0][warp] = __popc(be);
    __syncthreads();
    int eq_before = base_eq + __popc(be & ((1u << lane) - 1u)), eq_total = 0;
    for (int w = 0; w < kSelThreads / 32; ++w) {
      const int c = warp_cnt[0][w];
      if (w < warp) eq_before += c;
      eq_total += c;
    }
    const bool take = gt || (eq && eq_before < (int)need);
    const unsigned int bs = __ballot_sync(0xffffffffu, take);
    if (lane == 0) warp_cnt[1][warp] = __popc(bs);
    __syncthreads();
    int idx = base_sel + __popc(bs & ((1u << lane) - 1u)), sel_total = 0;
    for (int w = 0; w < kSelThreads / 32; ++w) {
      const int c = warp_cnt[1][w];
      if (w < warp) idx += c;
      sel_total += c;
    }
    if (take && idx < n) {
      lkey[idx] = k;
      lpos[idx] = p;
    }
    base_eq += eq_total;
    base_sel += sel_total;
    __syncthreads();                                  // warp_cnt is rewritten by the next chunk
  }
}

__global__ void __launch_bounds__(128) bbox_rank_selected_kernel(const unsigned int* __restrict__ list_key,
                                                                 const int* __restrict__ list_pos, int n,
                                                                 int* __restrict__ order) {
  extern __shared__ __align__(16) unsigned int rk_sm[];       // [n4]
  const int b = blockIdx.y, n4 = (n + 3) & ~3;
  const unsigned int* lkey = list_key + (size_t)b * n4;
  for (int i = threadIdx.x; i < n4; i += blockDim.x) rk_sm[i] = lkey[i];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned int mine = rk_sm[i];
  int rank = 0;
  for (int j = 0; j < n4; j += 4) {
    const uint4 o = *reinterpret_cast<const uint4*>(rk_sm + j);
    rank += (o.x > mine || (o.x == mine && j < i)) ? 1 : 0;
    rank += (o.y > mine || (o.y == mine && j + 1 < i)) ? 1 : 0;
    rank += (o.z > mine || (o.z == mine && j + 2 < i)) ? 1 : 0;
    rank += (o.w > mine || (o.w == mine && j + 3 < i)) ? 1 : 0;
  }
  order[(size_t)b * n + rank] = list_pos[(size_t)b * n4 + i];
}

__global__ void iota_rows_kernel(int* __restrict__ order, int HW) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < HW) order[(size_t)blockIdx.y * HW + p] = p;
}

// ---- 2. decode the candidates: boxes [B, n, 4] and the dense NMS input dets [B, C, n, 5] ---------------------
__global__ void bbox_decode_kernel(const float* __restrict__ scores, int apply_sigmoid, const float* __restrict__ bbox,
                                   const int* __restrict__ order, const float* __restrict__ lim /*[B,2]: w,h*/,
                                   float stride, int Wmap, int C, int HW, int n, float* __restrict__ boxes,
                                   float* __restrict__ dets) {
  const int b = blockIdx.y;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int p = min(max(order[(size_t)b * n + r], 0), HW - 1);     // defensive: never index outside the map
  const float cx = (float)(p % Wmap) * stride, cy = (float)(p / Wmap) * stride;      // point_generator.py:14-23
  const float w = lim[b * 2], h = lim[b * 2 + 1];
  const float* bb = bbox + (size_t)b * 4 * HW + p;
  float box[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float v = __fadd_rn(__fmul_rn(bb[(size_t)i * HW], stride), (i & 1) ? cy : cx);   // KP3:875-877
    box[i] = fminf(fmaxf(v, 0.f), (i & 1) ? h : w);                                        // KP3:882-886
  }
  *reinterpret_cast<float4*>(boxes + ((size_t)b * n + r) * 4) = make_float4(box[0], box[1], box[2], box[3]);
  const float* sb = scores + (size_t)b * C * HW + p;
  for (int c = 0; c < C; ++c) {
    float s = sb[(size_t)c * HW];
    if (apply_sigmoid) s = sigmoid_ref(s);
    float* d = dets + (((size_t)b * C + c) * n + r) * 5;
    d[0] = box[0]; d[1] = box[1]; d[2] = box[2]; d[3] = box[3]; d[4] = s;
  }
}

// ---- 3. results: for the k best surviving (class, candidate) pairs of every image ---------------------------
// top_i[b, j] = class * n + candidate (bbox_nms_kp.py:64-70 order), top_s its score (<= 0: empty slot).
// out_dets [B, k, 5], out_labels [B, k] (-1 = empty), out_kpts [B, k, P*3] = (x, y, 1) decoded and clamped
// (points2kpt KP3:393-410, decode KP3:878-880,887-888, visibility KP3:856-861); one warp per detection.
__global__ void bbox_finalize_kernel(const float* __restrict__ boxes, const float* __restrict__ kp,
                                     const int* __restrict__ order, const long long* __restrict__ top_i,
                                     const float* __restrict__ top_s, const float* __restrict__ lim, float stride,
                                     int Wmap, int HW, int n, int k, int P, float* __restrict__ out_dets,
                                     long long* __restrict__ out_labels, float* __restrict__ out_kpts) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (j >= k) return;
  const long long ti = top_i[(size_t)b * k + j];
  const float s = top_s[(size_t)b * k + j];
  const bool valid = s > 0.f;
  const int r = (int)(ti % n), cls = (int)(ti / n);
  float* od = out_dets + ((size_t)b * k + j) * 5;
  float* ok = out_kpts + ((size_t)b * k + j) * P * 3;
  if (lane < 5) {
    // the reference path multiplies by the validity flag: an empty slot reads 0 (or -0 for the -1 filler score)
    const float v = lane < 4 ? boxes[((size_t)b * n + r) * 4 + lane] : s;
    od[lane] = valid ? v : __fmul_rn(v, 0.f);
  }
  if (lane == 0) out_labels[(size_t)b * k + j] = valid ? (long long)cls : -1ll;
  const int p = min(max(order[(size_t)b * n + r], 0), HW - 1);
  const float cx = (float)(p % Wmap) * stride, cy = (float)(p / Wmap) * stride;
  const float w = lim[b * 2], h = lim[b * 2 + 1];
  const float* kb = kp + (size_t)b * 2 * P * HW + p;
  for (int i = lane; i < P; i += 32) {
    const float y = kb[(size_t)(2 * i) * HW], x = kb[(size_t)(2 * i + 1) * HW];         // y-first pairs
    const float xo = fminf(fmaxf(__fadd_rn(__fmul_rn(x, stride), cx), 0.f), w);
    const float yo = fminf(fmaxf(__fadd_rn(__fmul_rn(y, stride), cy), 0.f), h);
    ok[i * 3 + 0] = valid ? xo : __fmul_rn(xo, 0.f);
    ok[i * 3 + 1] = valid ? yo : __fmul_rn(yo, 0.f);
    ok[i * 3 + 2] = valid ? 1.f : 0.f;
  }
}

// Global top-k over the survivors of the batched NMS (bbox_nms_kp.py:64-70: `scores.sort(descending)[:max_num]`
// over the concatenated per-class results): one CTA per image.  The kept (class, candidate) pairs are compacted
// into shared memory (typically a few hundred of the 13 x 1000 candidates), sorted there by (score desc,
// index asc) with a bitonic network sized to the kept count, and the first k leave.  Replaces
// torch.where + torch.topk (one 16-CTA radix-select launch of ~59 us on the critical path of the step).
// top_s of an empty slot is -1 (the PyTorch path's filler), its top_i 0.
static constexpr int kTopkCap = 16384;

__device__ __forceinline__ bool topk_before(float sa, int ia, float sb, int ib) {
  return (sa > sb) || (sa == sb && ia < ib);
}

__global__ void __launch_bounds__(1024, 1)
topk_flagged_kernel(const float* __restrict__ dets, const uint8_t* __restrict__ flags, int L, int k,
                    float* __restrict__ top_s, long long* __restrict__ top_i) {
  extern __shared__ __align__(16) unsigned char topk_smem[];
  float* key = reinterpret_cast<float*>(topk_smem);
  int* idx = reinterpret_cast<int*>(key + kTopkCap);
  __shared__ int cnt;
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  if (tid == 0) cnt = 0;
  __syncthreads();
  const uint8_t* f = flags + (size_t)b * L;
  const float* d = dets + (size_t)b * L * 5;
  for (int i = tid; i < L; i += nt) {
    if (f[i]) {
      const int slot = atomicAdd(&cnt, 1);               // order is irrelevant: the sort key is (score, index)
      key[slot] = d[(size_t)i * 5 + 4];
      idx[slot] = i;
    }
  }
  __syncthreads();
  const int m = cnt;
  int P = 32;
  while (P < m) P <<= 1;
  for (int i = m + tid; i < P; i += nt) { key[i] = -INFINITY; idx[i] = 0x7fffffff; }
  __syncthreads();
  for (int kk = 2; kk <= P; kk <<= 1) {
    for (int j = kk >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < P; i += nt) {
        const int l = i ^ j;
        if (l > i) {
          const float si = key[i], sl = key[l];
          const int ii = idx[i], il = idx[l];
          const bool up = ((i & kk) == 0);
          const bool swap = up ? topk_before(sl, il, si, ii) : topk_before(si, ii, sl, il);
          if (swap) { key[i] = sl; key[l] = si; idx[i] = il; idx[l] = ii; }
        }
      }
      __syncthreads();
    }
  }
  for (int j = tid; j < k; j += nt) {
    const bool has = j < m;
    top_s[(size_t)b * k + j] = has ? key[j] : -1.f;
    top_i[(size_t)b * k + j] = has ? (long long)idx[j] : 0ll;
  }
}

}  // namespace kgdet

using namespace kgdet;

extern "C" int kgdet_topk_flagged(const float* dets, const uint8_t* flags, int32_t B, int32_t L, int32_t k,
                                  float* top_s, int64_t* top_i, void* stream) {
  KG_CHECK_ARG(dets && flags && top_s && top_i, "kgdet_topk_flagged: NULL pointer");
  KG_CHECK_ARG(B >= 0 && L >= 1 && k >= 1 && B <= 65535, "kgdet_topk_flagged: bad sizes");
  KG_CHECK_ARG(L <= kTopkCap, "kgdet_topk_flagged: at most %d candidates per image", kTopkCap);
  if (B == 0) return KGDET_OK;
  const size_t smem = (size_t)kTopkCap * 8;
  KG_CUDA(cudaFuncSetAttribute(topk_flagged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  topk_flagged_kernel<<<B, 1024, smem, (cudaStream_t)stream>>>(dets, flags, L, k, top_s, (long long*)top_i);
  KG_LAUNCH_CHECK("topk_flagged_kernel");
  return KGDET_OK;
}

// scratch of the large-level path of kgdet_bbox_select (0: none needed)
extern "C" size_t kgdet_bbox_select_workspace_bytes(int32_t B, int32_t HW, int32_t n) {
  if (B <= 0 || HW <= 4096 || n >= HW || n > kSelMaxN || HW > kSelMaxHW) return 0;
  return ((size_t)B * HW + 2 * (size_t)B * ((n + 3) & ~3)) * 4;
}

static int bbox_select_impl(const float* scores, int apply_sigmoid, int32_t B, int32_t C, int32_t HW, int32_t n,
                            int32_t* order, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  KG_CHECK_ARG(scores && order, "kgdet_bbox_select: NULL pointer");
  KG_CHECK_ARG(B >= 0 && C >= 1 && HW >= 1 && n >= 1 && n <= HW, "kgdet_bbox_select: bad sizes");
  KG_CHECK_ARG(B <= 65535, "kgdet_bbox_select: batch too large");
  if (B == 0) return KGDET_OK;
  if (n == HW) {                                 // no top-k in the reference: original order
    iota_rows_kernel<<<dim3(ceil_div(HW, 256), B), 256, 0, stream>>>(order, HW);
    KG_LAUNCH_CHECK("iota_rows_kernel");
    return KGDET_OK;
  }
  const size_t need = kgdet_bbox_select_workspace_bytes(B, HW, n);
  if (need && workspace && workspace_bytes >= need && ((uintptr_t)workspace & 15) == 0) {
    const int n4 = (n + 3) & ~3;
    unsigned int* keys = (unsigned int*)workspace;
    unsigned int* lkey = keys + (size_t)B * HW;
    int* lpos = (int*)(lkey + (size_t)B * n4);
    level_keys_kernel<<<dim3(ceil_div(HW, 256), B), 256, 0, stream>>>(scores, apply_sigmoid, C, HW, keys);
    KG_LAUNCH_CHECK("level_keys_kernel");
    const size_t smem = ((size_t)((HW + 3) & ~3) + 2048) * 4;
    KG_CUDA(cudaFuncSetAttribute(bbox_select_radix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    bbox_select_radix_kernel<<<B, kSelThreads, smem, stream>>>(keys, HW, n, lkey, lpos);
    KG_LAUNCH_CHECK("bbox_select_radix_kernel");
    bbox_rank_selected_kernel<<<dim3(ceil_div(n, 128), B), 128, (size_t)n4 * 4, stream>>>(lkey, lpos, n, order);
    KG_LAUNCH_CHECK("bbox_rank_selected_kernel");
    return KGDET_OK;
  }
  KG_CHECK_ARG(HW <= 16384, "kgdet_bbox_select: levels of more than 16384 positions need kgdet_bbox_select_ws with "
               "kgdet_bbox_select_workspace_bytes of scratch (up to %d positions, %d candidates)", kSelMaxHW, kSelMaxN);
  const size_t smem = (size_t)(HW + 4) * sizeof(float);
  KG_CUDA(cudaFuncSetAttribute(bbox_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  bbox_select_kernel<<<dim3(ceil_div(HW, 256), B), 256, smem, stream>>>(scores, apply_sigmoid, C, HW, n, order);
  KG_LAUNCH_CHECK("bbox_select_kernel");
  return KGDET_OK;
}

extern "C" int kgdet_bbox_select(const float* scores, int apply_sigmoid, int32_t B, int32_t C, int32_t HW,
                                 int32_t n, int32_t* order, void* stream) {
  return bbox_select_impl(scores, apply_sigmoid, B, C, HW, n, order, nullptr, 0, stream);
}

extern "C" int kgdet_bbox_select_ws(const float* scores, int apply_sigmoid, int32_t B, int32_t C, int32_t HW,
                                    int32_t n, int32_t* order, void* workspace, size_t workspace_bytes, void* stream) {
  return bbox_select_impl(scores, apply_sigmoid, B, C, HW, n, order, workspace, workspace_bytes, stream);
}

extern "C" int kgdet_bbox_decode(const float* scores, int apply_sigmoid, const float* bbox, const int32_t* order,
                                 const float* img_wh, float stride, int32_t map_w, int32_t B, int32_t C, int32_t HW,
                                 int32_t n, float* boxes, float* dets, void* stream) {
  KG_CHECK_ARG(scores && bbox && order && img_wh && boxes && dets, "kgdet_bbox_decode: NULL pointer");
  KG_CHECK_ARG(B >= 0 && C >= 1 && HW >= 1 && n >= 1 && n <= HW && map_w >= 1 && B <= 65535,
               "kgdet_bbox_decode: bad sizes");
  if (B == 0) return KGDET_OK;
  bbox_decode_kernel<<<dim3(ceil_div(n, 128), B), 128, 0, (cudaStream_t)stream>>>(scores, apply_sigmoid, bbox, order,
                                                                                  img_wh, stride, map_w, C, HW, n,
                                                                                  boxes, dets);
  KG_LAUNCH_CHECK("bbox_decode_kernel");
  return KGDET_OK;
}

extern "C" int kgdet_bbox_finalize(const float* boxes, const float* keypts, const int32_t* order,
                                   const int64_t* top_i, const float* top_s, const float* img_wh, float stride,
                                   int32_t map_w, int32_t B, int32_t HW, int32_t n, int32_t k, int32_t num_keypts,
                                   float* out_dets, int64_t* out_labels, float* out_kpts, void* stream) {
  KG_CHECK_ARG(boxes && keypts && order && top_i && top_s && img_wh && out_dets && out_labels && out_kpts,
               "kgdet_bbox_finalize: NULL pointer");
  KG_CHECK_ARG(B >= 0 && HW >= 1 && n >= 1 && k >= 1 && num_keypts >= 1 && map_w >= 1 && B <= 65535,
               "kgdet_bbox_finalize: bad sizes");
  if (B == 0) return KGDET_OK;
  bbox_finalize_kernel<<<dim3(ceil_div(k, 8), B), 256, 0, (cudaStream_t)stream>>>(
      boxes, keypts, order, (const long long*)top_i, top_s, img_wh, stride, map_w, HW, n, k, num_keypts, out_dets,
      (long long*)out_labels, out_kpts);
  KG_LAUNCH_CHECK("bbox_finalize_kernel");
  return KGDET_OK;
}
