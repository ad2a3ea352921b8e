// Target assignment and the nine training losses of the KGDet head as three kernels.
//
// The reference does this on the host side of PyTorch, per image and per ground-truth box:
//   PointAssigner.assign            mmdet/core/bbox/assigners/point_assigner.py:23-116  (Python loop over the boxes,
//                                   a topk each, boolean-mask updates)
//   point_target_kp / _single       mmdet/core/anchor/point_target_kp.py:7-169          (mask indexing, nonzero())
//   KP3.loss / loss_single          reppoints_head_kp3rep_cas_1_assign_once.py:581-768  (3 focal + 6 smooth-L1 losses,
//                                   dozens of elementwise / reduction kernels, avg_factor on the host)
// ~120 PyTorch kernels forward + as many backward, each mask index a device -> host round trip.  Here:
//   point_assign_{rank,finish}_kernel   per (image, box) in parallel: the normalised centre distances of all points,
//                          the pos_num nearest by rank counting (ties: lower index); the reference's update rule "a
//                          later box takes a point only if strictly closer" (:98-99) as a 64-bit atomicMin per point;
//                          avg_factor = sum over images of max(#positives, 1) (point_target_kp.py:64) on the device.
//   point_losses_fwd_kernel  the nine sums in one pass over the nine head outputs (NCHW, read in place: no
//                          permute / reshape copies): focal terms with the reference CUDA kernel's arithmetic
//                          (focal.cuh), smooth-L1 of decoded boxes / keypoints against targets looked up through
//                          the assignment (the dense target tensors of point_target_kp are never built).
//   point_losses_bwd_kernel  the nine gradients in one pass, scaled by the upstream gradients of the nine losses.
// Single point level (the KGDet configs: point_strides=[32]); every point is valid (the map covers the padded
// image) and sampling=False (PseudoSamplerKp), as in the reference configs.
#include "focal.cuh"

namespace kgdet {

static constexpr int PA_THREADS = 256;

// assigned[b, p] = 0 (background) or g + 1; avg += max(#positives of image b, 1).
// The reference walks the boxes of an image in order (point_assigner.py:72-109) and lets a later box take a point
// only if it is strictly closer (:98-99): per point that is the minimum over the boxes whose pos_num nearest include
// it of (distance, box index) -- distances are >= 0, so their bit patterns order like the values and one 64-bit
// atomicMin on (distance bits << 32 | box) resolves it in any order.  Two kernels:
//   rank    grid (point blocks of 256, boxes, images): a CTA computes the distances of ALL points of its image to
//           its box (cheap) and ranks only its own 256 points against them (4 distances per shared-memory load) --
//           the boxes of an image run in parallel instead of one after the other (120 -> ~20 us for a batch of 2);
//   finish  grid (point blocks, images): the winners -> assigned; the last CTA of an image to finish (atomic ticket)
//           folds the image's positive count into avg_factor.
__global__ void __launch_bounds__(PA_THREADS) point_assign_rank_kernel(const float* __restrict__ gt_boxes /*[B,G,4]*/,
                                                                       const unsigned char* __restrict__ gt_valid /*[B,G]*/,
                                                                       const float* __restrict__ gt_kps /*[B,G,K,3]*/,
                                                                       int G, int K, int P, int Wmap, float stride,
                                                                       int pos_num, unsigned long long* __restrict__ best,
                                                                       float* __restrict__ nvis /*[B,G]*/) {
  extern __shared__ __align__(16) float dist[];   // [P rounded up to 4]
  const int b = blockIdx.z, g = blockIdx.y;
  const int P4 = (P + 3) & ~3;
  // visible keypoints per box (the keypoint-loss weights of a positive row are 4 / (2 * nvis), KP3:639-644)
  if (blockIdx.x == 0) {
    const float* gk = gt_kps + ((size_t)b * G + g) * K * 3;
    float c = 0.f;
    for (int j = threadIdx.x; j < K; j += blockDim.x) c += gk[(size_t)j * 3 + 2] != 0.f ? 1.f : 0.f;
    c = warp_sum(c);                              // small integers: exact in any order
    __shared__ float part[PA_THREADS / 32];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < PA_THREADS / 32; ++w) t += part[w];
      nvis[(size_t)b * G + g] = t;
    }
  }
  if (gt_valid[(size_t)b * G + g] == 0) return;              // uniform over the CTA
  const int k = pos_num < P ? pos_num : P;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;       // this thread's point
  const float* bx = gt_boxes + ((size_t)b * G + g) * 4;
  const float x1 = bx[0], y1 = bx[1], x2 = bx[2], y2 = bx[3];
  const float cx = __fmul_rn(__fadd_rn(x1, x2), 0.5f), cy = __fmul_rn(__fadd_rn(y1, y2), 0.5f);   // :59 (x / 2)
  const float w = fmaxf(__fsub_rn(x2, x1), 1e-6f), h = fmaxf(__fsub_rn(y2, y1), 1e-6f);           // :60
  for (int q = threadIdx.x; q < P4; q += blockDim.x) {
    const float px = (float)(q % Wmap) * stride, py = (float)(q / Wmap) * stride;                 // point_generator.py:14-23
    const float dx = __fdiv_rn(__fsub_rn(px, cx), w), dy = __fdiv_rn(__fsub_rn(py, cy), h);       // :84
    dist[q] = q < P ? sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy))) : INFINITY;
  }
  __syncthreads();
  if (p >= P) return;
  const float mine = dist[p];
  int rank = 0;
  for (int j = 0; j < P4; j += 4) {
    const float4 o = *reinterpret_cast<const float4*>(dist + j);
    rank += (o.x < mine || (o.x == mine && j < p)) ? 1 : 0;
    rank += (o.y < mine || (o.y == mine && j + 1 < p)) ? 1 : 0;
    rank += (o.z < mine || (o.z == mine && j + 2 < p)) ? 1 : 0;
    rank += (o.w < mine || (o.w == mine && j + 3 < p)) ? 1 : 0;
  }
  // among the k nearest of this box (:90-91); a NaN distance (degenerate box) never ranks and is never taken
  if (rank < k && mine == mine)
    atomicMin(&best[(size_t)b * P + p], ((unsigned long long)__float_as_uint(mine) << 32) | (unsigned int)g);
}

__global__ void __launch_bounds__(PA_THREADS) point_assign_finish_kernel(const unsigned long long* __restrict__ best,
                                                                         int P, int* __restrict__ assigned,
                                                                         float* __restrict__ avg,
                                                                         int* __restrict__ scratch /*[B][2], zeroed*/) {
  const int b = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  int who = 0;
  if (p < P) {
    const unsigned long long v = best[(size_t)b * P + p];
    who = v == ~0ull ? 0 : (int)(unsigned int)(v & 0xFFFFFFFFull) + 1;
    assigned[(size_t)b * P + p] = who;
  }
  int mypos = (p < P && who > 0) ? 1 : 0;
  mypos = (int)warp_sum((float)mypos);
  if ((threadIdx.x & 31) == 0 && mypos) atomicAdd(&scratch[2 * b], mypos);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int ticket = atomicAdd(&scratch[2 * b + 1], 1);
    if (ticket == (int)gridDim.x - 1) {                      // last CTA of this image
      const int npos = atomicAdd(&scratch[2 * b], 0);
      atomicAdd(avg, (float)(npos > 1 ? npos : 1));          // small integers: exact in any order
    }
  }
}

struct PointLossParams {
  const float* out[9];       // cls_1..3 [B,NC,H,W], kpt_1..3 [B,2K,H,W], bbox_1..3 [B,4,H,W]
  float* grad[9];            // same shapes (backward), any of them may be NULL
  const int* assigned;       // [B, P]
  const float* gt_boxes;     // [B, G, 4]
  const long long* gt_labels;   // [B, G]
  const float* gt_kps;       // [B, G, K, 3]
  const float* avg;          // device scalar
  const float* nvis;         // [B, G] visible keypoints per box (point_assign_rank_kernel)
  const float* grad_losses;  // [9] upstream gradients (backward)
  float* losses;             // [9] (forward): cls_1..3, bbox_1..3, kpt_1..3
  int B, G, P, Wmap, NC, K;
  float stride, nt;          // nt = point_base_scale * stride
  float w_cls[3], w_bbox[3], w_kpt[3];
  float gamma, alpha, beta;
};

__device__ __forceinline__ float smooth_l1(float diff, float beta) {        // losses/smooth_l1_loss.py:9-15
  return diff < beta ? 0.5f * diff * diff / beta : diff - 0.5f * beta;
}
__device__ __forceinline__ float smooth_l1_grad(float d, float beta) {      // d = pred - target (signed)
  const float a = fabsf(d);
  return a < beta ? d / beta : (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
}

// acc[k] += v with the accumulators staying in registers (a dynamically indexed array would live in local memory)
__device__ __forceinline__ void acc_add(float (&acc)[9], int k, float v) {
#pragma unroll
  for (int i = 0; i < 9; ++i) acc[i] += (i == k) ? v : 0.f;
}

// Unified index space: [0, nA) focal elements (stage, b, c, p), [nA, nA + nB) box rows (stage, b, p),
// [nA + nB, nA + nB + nC) keypoint elements (stage, b, ch, p); p fastest everywhere (NCHW: coalesced).
template <bool BWD>
__global__ void __launch_bounds__(256) point_losses_kernel(const PointLossParams q) {
  const long long nA = 3ll * q.B * q.NC * q.P, nB = 3ll * q.B * q.P, nC = 3ll * q.B * (2 * q.K) * q.P;
  const long long total = nA + nB + nC;
  const float inv_avg = 1.f / *q.avg;
  float acc[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) acc[i] = 0.f;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)blockDim.x * gridDim.x) {
    if (idx < nA) {
      // ---- sigmoid focal loss, label weight 1 (point_target_kp.py:144-149 with pos_weight <= 0) ----
      const int p = (int)(idx % q.P);
      long long r = idx / q.P;
      const int c = (int)(r % q.NC); r /= q.NC;
      const int b = (int)(r % q.B), st = (int)(r / q.B);
      const int a = q.assigned[(size_t)b * q.P + p];
      const int t = a > 0 ? (int)q.gt_labels[(size_t)b * q.G + a - 1] : 0;                       // :143
      const size_t off = ((size_t)b * q.NC + c) * q.P + p;
      const float x = q.out[st][off];
      if (!BWD) {
        const float v = focal_fwd_value(x, t, c, q.gamma, q.alpha);
        acc_add(acc, st, v);
      } else if (q.grad[st]) {
        q.grad[st][off] = focal_bwd_value(x, t, c, q.gamma, q.alpha) * (q.w_cls[st] * inv_avg * q.grad_losses[st]);
      }
    } else if (idx < nA + nB) {
      // ---- box loss: smooth-L1 of the decoded box / nt against the assigned ground-truth box / nt (KP3:614-636) ----
      const long long i2 = idx - nA;
      const int p = (int)(i2 % q.P);
      const long long r = i2 / q.P;
      const int b = (int)(r % q.B), st = (int)(r / q.B);
      const int a = q.assigned[(size_t)b * q.P + p];
      const float px = (float)(p % q.Wmap) * q.stride, py = (float)(p / q.Wmap) * q.stride;
      const float* o = q.out[6 + st] + (size_t)b * 4 * q.P + p;
      float* go = (BWD && q.grad[6 + st]) ? q.grad[6 + st] + (size_t)b * 4 * q.P + p : nullptr;
      if (a == 0) {
        if (go) { go[0] = 0.f; go[(size_t)q.P] = 0.f; go[(size_t)2 * q.P] = 0.f; go[(size_t)3 * q.P] = 0.f; }
        continue;
      }
      const float* gb = q.gt_boxes + ((size_t)b * q.G + a - 1) * 4;
      const float gscale = BWD ? q.w_bbox[st] * inv_avg * q.grad_losses[3 + st] * (q.stride / q.nt) : 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float pred = o[(size_t)e * q.P] * q.stride + ((e & 1) ? py : px);                 // offset_to_pts, y_first=False
        const float d = pred / q.nt - gb[e] / q.nt;
        if (!BWD) acc_add(acc, 3 + st, smooth_l1(fabsf(d), q.beta));
        else if (go) go[(size_t)e * q.P] = smooth_l1_grad(d, q.beta) * gscale;
      }
    } else {
      // ---- keypoint loss (KP3:638-665): weights = visible keypoints of the assigned box, each positive row sums to 4 ----
      const long long i3 = idx - nA - nB;
      const int p = (int)(i3 % q.P);
      long long r = i3 / q.P;
      const int ch = (int)(r % (2 * q.K)); r /= (2 * q.K);
      const int b = (int)(r % q.B), st = (int)(r / q.B);
      const int a = q.assigned[(size_t)b * q.P + p];
      const size_t off = ((size_t)b * 2 * q.K + ch) * q.P + p;
      float* go = (BWD && q.grad[3 + st]) ? q.grad[3 + st] + off : nullptr;
      if (a == 0) {
        if (go) *go = 0.f;
        continue;
      }
      const int kp = ch >> 1, is_x = ch & 1;                      // channel 2i = y_i, 2i + 1 = x_i (y-first pairs)
      const float* gk = q.gt_kps + (((size_t)b * q.G + a - 1) * q.K) * 3;
      const float vis = gk[(size_t)kp * 3 + 2];
      if (vis == 0.f) {
        if (go) *go = 0.f;
        continue;
      }
      const float wgt = 1.f / (2.f * q.nvis[(size_t)b * q.G + a - 1]) * 4.f;    // kw / kw.sum(1) * 4, kw.sum(1) = 2 * nvis
      const float px = (float)(p % q.Wmap) * q.stride, py = (float)(p / q.Wmap) * q.stride;
      const float pred = q.out[3 + st][off] * q.stride + (is_x ? px : py);
      const float tgt = gk[(size_t)kp * 3 + (is_x ? 0 : 1)];
      const float d = pred / q.nt - tgt / q.nt;
      if (!BWD) acc_add(acc, 6 + st, wgt * smooth_l1(fabsf(d), q.beta));
      else if (go) *go = smooth_l1_grad(d, q.beta) * wgt * (q.w_kpt[st] * inv_avg * q.grad_losses[6 + st] * (q.stride / q.nt));
    }
  }
  if (!BWD) {
    __shared__ float part[9][8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const float v = warp_sum(acc[i]);
      if (lane == 0) part[i][wid] = v;
    }
    __syncthreads();
    if (threadIdx.x < 9) {
      float v = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += part[threadIdx.x][w];
      const int i = threadIdx.x;
      const float lw = i < 3 ? q.w_cls[i] : (i < 6 ? q.w_bbox[i - 3] : q.w_kpt[i - 6]);
      if (v != 0.f) atomicAdd(q.losses + i, v * lw * inv_avg);
    }
  }
}

}  // namespace kgdet

using namespace kgdet;

// scratch of kgdet_point_assign: two counters per image + one 64-bit (distance, box) winner per point
extern "C" size_t kgdet_point_assign_scratch_bytes(int32_t B, int32_t map_h, int32_t map_w) {
  if (B <= 0 || map_h <= 0 || map_w <= 0) return 0;
  return align_up((size_t)B * 2 * sizeof(int), 16) + (size_t)B * map_h * map_w * sizeof(unsigned long long);
}

extern "C" int kgdet_point_assign(const float* gt_boxes, const uint8_t* gt_valid, const float* gt_keypoints, int32_t B,
                                  int32_t G, int32_t num_keypoints, int32_t map_h, int32_t map_w, float stride,
                                  int32_t pos_num, int32_t* assigned, float* avg_factor, float* num_visible,
                                  void* scratch, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  KG_CHECK_ARG(B >= 0 && G >= 0 && map_h > 0 && map_w > 0 && pos_num > 0, "kgdet_point_assign: bad sizes");
  const int P = map_h * map_w;
  KG_CHECK_ARG(P <= 4096, "kgdet_point_assign: at most 4096 points per level (got %d)", P);
  KG_CHECK_ARG(assigned && avg_factor && scratch && (G == 0 || (gt_boxes && gt_valid && gt_keypoints && num_visible)),
               "kgdet_point_assign: NULL pointer");
  KG_CHECK_ARG(((uintptr_t)scratch & 15) == 0, "kgdet_point_assign: scratch must be 16-byte aligned");
  KG_CHECK_ARG(B <= 65535 && G <= 65535, "kgdet_point_assign: batch / box count too large");
  KG_CHECK_ARG(num_keypoints >= 0, "kgdet_point_assign: bad keypoint count");
  KG_CUDA(cudaMemsetAsync(avg_factor, 0, sizeof(float), stream));
  if (B == 0) return KGDET_OK;
  const size_t counters = align_up((size_t)B * 2 * sizeof(int), 16);
  unsigned long long* best = (unsigned long long*)((unsigned char*)scratch + counters);
  KG_CUDA(cudaMemsetAsync(scratch, 0, counters, stream));
  KG_CUDA(cudaMemsetAsync(best, 0xFF, (size_t)B * P * sizeof(unsigned long long), stream));
  const int pblocks = ceil_div(P, PA_THREADS);
  if (G > 0) {
    point_assign_rank_kernel<<<dim3(pblocks, G, B), PA_THREADS, (size_t)(P + 4) * sizeof(float), stream>>>(
        gt_boxes, gt_valid, gt_keypoints, G, num_keypoints, P, map_w, stride, pos_num, best, num_visible);
    KG_LAUNCH_CHECK("point_assign_rank_kernel");
  }
  point_assign_finish_kernel<<<dim3(pblocks, B), PA_THREADS, 0, stream>>>(best, P, assigned, avg_factor, (int*)scratch);
  KG_LAUNCH_CHECK("point_assign_finish_kernel");
  return KGDET_OK;
}

static int fill_params(PointLossParams& q, const float* const* outs, const int32_t* assigned, const float* gt_boxes,
                       const int64_t* gt_labels, const float* gt_keypoints, const float* avg_factor,
                       const float* num_visible, int32_t B, int32_t G,
                       int32_t map_h, int32_t map_w, int32_t num_classes, int32_t num_keypoints, float stride,
                       float point_base_scale, const float* loss_weights, float gamma, float alpha, float beta) {
  KG_CHECK_ARG(outs && assigned && gt_boxes && gt_labels && gt_keypoints && avg_factor && num_visible && loss_weights,
               "kgdet_point_losses: NULL pointer");
  q.nvis = num_visible;
  KG_CHECK_ARG(B > 0 && G > 0 && map_h > 0 && map_w > 0 && num_classes > 0 && num_keypoints > 0,
               "kgdet_point_losses: bad sizes");
  for (int i = 0; i < 9; ++i) {
    KG_CHECK_ARG(outs[i], "kgdet_point_losses: head output %d is NULL", i);
    q.out[i] = outs[i];
    q.grad[i] = nullptr;
  }
  q.assigned = assigned; q.gt_boxes = gt_boxes; q.gt_labels = (const long long*)gt_labels; q.gt_kps = gt_keypoints;
  q.avg = avg_factor; q.grad_losses = nullptr; q.losses = nullptr;
  q.B = B; q.G = G; q.P = map_h * map_w; q.Wmap = map_w; q.NC = num_classes; q.K = num_keypoints;
  q.stride = stride; q.nt = point_base_scale * stride;
  for (int i = 0; i < 3; ++i) { q.w_cls[i] = loss_weights[i]; q.w_bbox[i] = loss_weights[3 + i]; q.w_kpt[i] = loss_weights[6 + i]; }
  q.gamma = gamma; q.alpha = alpha; q.beta = beta;
  return KGDET_OK;
}

static int loss_grid(const PointLossParams& q) {
  const long long total = 3ll * q.B * q.P * (q.NC + 1 + 2 * q.K);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  return (int)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
}

extern "C" int kgdet_point_losses_forward(const float* const* outs, const int32_t* assigned, const float* gt_boxes,
                                          const int64_t* gt_labels, const float* gt_keypoints, const float* avg_factor,
                                          const float* num_visible, int32_t B, int32_t G, int32_t map_h, int32_t map_w,
                                          int32_t num_classes,
                                          int32_t num_keypoints, float stride, float point_base_scale,
                                          const float* loss_weights, float gamma, float alpha, float beta, float* losses,
                                          void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PointLossParams q;
  int rc = fill_params(q, outs, assigned, gt_boxes, gt_labels, gt_keypoints, avg_factor, num_visible, B, G, map_h, map_w,
                       num_classes, num_keypoints, stride, point_base_scale, loss_weights, gamma, alpha, beta);
  if (rc != KGDET_OK) return rc;
  KG_CHECK_ARG(losses, "kgdet_point_losses_forward: NULL pointer");
  q.losses = losses;
  KG_CUDA(cudaMemsetAsync(losses, 0, 9 * sizeof(float), stream));
  point_losses_kernel<false><<<loss_grid(q), 256, 0, stream>>>(q);
  KG_LAUNCH_CHECK("point_losses_fwd_kernel");
  return KGDET_OK;
}

extern "C" int kgdet_point_losses_backward(const float* const* outs, const int32_t* assigned, const float* gt_boxes,
                                           const int64_t* gt_labels, const float* gt_keypoints, const float* avg_factor,
                                           const float* num_visible, const float* grad_losses, int32_t B, int32_t G, int32_t map_h, int32_t map_w,
                                           int32_t num_classes, int32_t num_keypoints, float stride,
                                           float point_base_scale, const float* loss_weights, float gamma, float alpha,
                                           float beta, float* const* grad_outs, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PointLossParams q;
  int rc = fill_params(q, outs, assigned, gt_boxes, gt_labels, gt_keypoints, avg_factor, num_visible, B, G, map_h, map_w,
                       num_classes, num_keypoints, stride, point_base_scale, loss_weights, gamma, alpha, beta);
  if (rc != KGDET_OK) return rc;
  KG_CHECK_ARG(grad_losses && grad_outs, "kgdet_point_losses_backward: NULL pointer");
  q.grad_losses = grad_losses;
  for (int i = 0; i < 9; ++i) q.grad[i] = grad_outs[i];
  point_losses_kernel<true><<<loss_grid(q), 256, 0, stream>>>(q);
  KG_LAUNCH_CHECK("point_losses_bwd_kernel");
  return KGDET_OK;
}
