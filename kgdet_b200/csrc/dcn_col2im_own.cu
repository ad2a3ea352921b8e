// col2im + offset / mask gradient of the deformable convolution with the input gradient OWNED by the CTA.
//
// Reference: deformable_col2im_gpu_kernel (deform_conv_cuda_kernel.cu:278-330: one atomicAdd per column-gradient
// element and corner) and deformable_col2im_coord_gpu_kernel (:345-435).
//
// col2im_tc_kernel (dcn_bwd_tc.cu) sends every (sample, corner, 4 channels) contribution to L2 as a
// red.global.add.v4.f32: 210 M of them per K = 49 call, and the SM issues one per 1.29 clk -- that rate IS its run time.
// Here a CTA owns (image n, 32-channel block cb): the fp32 gradient of that slice of the map ([H*W][32], 131 KB at
// 25 x 42) lives in shared memory for the whole launch and is updated with PLAIN load / add / store (an update is
// 32 bytes of shared-memory traffic at 128 B/clk instead of a 1.29-clk L2 reduction per 16 bytes); the bf16 input
// slice the offset gradient needs ([H*W][32], 66 KB) sits next to it, so the four corner gathers never leave the SM.
//   lanes  = 32 consecutive positions of one tap (their corner pixels are mostly consecutive -> conflict-free banks
//            with the chunk swizzles below)
//   warp w = channels [4w, 4w + 4) of the block -> warps never touch the same accumulator word; two lanes of a warp
//            that hit the same pixel in the same corner step are found with match.any and take turns
//   the four corners of a sample are taken one after another (neighbouring samples share pixels ACROSS corner
//   indices), with a warp barrier between them
// Which lane takes which sample, and in which turn, is decided ONCE per offset tensor by own_schedule_kernel (all
// C / 32 x 8 warps that work on a tile see the same pixels): the 64 samples of a tile are dealt into octets of lanes
// with distinct (base pixel mod 8) -- the accumulator chunk a lane touches is (w ^ (pixel & 7)), and the four corners
// are base + {0, 1, W, W + 1}, so such an octet is free of bank conflicts in all four corner steps even for random
// offsets -- and the turn of every (sample, corner) among equal pixels of its 32-lane group is stored next to it
// (match.any in the hot loop cost ~500 clk per call with 32 distinct keys: 4.1 ms per K = 49 call against 0.85).
// <column gradient, corner value> partial sums over 4 channels go through an 8 KB exchange buffer, 64 threads add the
// eight warps' parts and STORE (dy, dx, sample) of the channel block into a staging tensor [C / 32][N][3K][Ho*Wo];
// col2im_own_finish_kernel adds the C / 32 parts in fixed order: one writer per element of grad_offset / grad_mask,
// no atomics, bitwise reproducible (and no scalar reds: the SM issues one red per 1.29 clk whatever its width).
// The slice is flushed once per launch with red.global.add.v4.f32 (8 400 per CTA against 1.6 M before).
#include "dcn.cuh"

namespace kgdet {

static constexpr int OWN_TS = 64;                 // samples (positions of one tap) per tile
static constexpr int OWN_STAGE_BYTES = 2 * OWN_TS * 64 + 2 * OWN_TS * 16 + 2 * OWN_TS * 16 + 8 * OWN_TS * 16 + 3 * OWN_TS * 4;
static constexpr int OWN_SMEM_MAX = 227 * 1024;

static size_t own_smem_bytes(const DcnGeom& g) { return (size_t)g.H * g.W * 192 + OWN_STAGE_BYTES; }

bool col2im_own_geometry_ok(const DcnGeom& g) {
  return g.dgroups == 1 && g.groups == 1 && g.C % 32 == 0 && own_smem_bytes(g) <= (size_t)OWN_SMEM_MAX;
}
// Opt-in (KGDET_COL2IM_OWN=1): measured SLOWER than col2im_tc_kernel on the B200 -- K = 49 call at batch 16 with
// random offsets: 2.2 ms against 0.86 ms (profiles/r2_col2im_own_experiment.txt).  The slice update itself is cheap,
// but every (sample, corner) costs each of the 8 warps of the CTA ~100 instructions of bookkeeping (descriptor
// loads, turn logic, swizzled addresses, barriers) at 0.31 IPC with 8 warps per SM, where the red.global path pays
// one instruction per 4 channels and runs 32 warps per SM.
bool col2im_own_supported(const DcnGeom& g) {
  const char* e = getenv("KGDET_COL2IM_OWN");
  if (!e || atoi(e) == 0) return false;
  return col2im_own_geometry_ok(g);
}
// staging tensor of the offset / mask gradient parts: [C / 32][N][3K][Ho*Wo] fp32
size_t col2im_own_part_bytes(const DcnGeom& g) {
  return col2im_own_geometry_ok(g) ? (size_t)(g.C / 32) * g.N * 3 * g.K * g.Ho * g.Wo * 4 : 0;
}

__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void unpack4(const uint2& v, float (&f)[4]) {
  f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
  f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
}


// Schedule word of a lane slot: bits 0-5 sample (position p0 + sample of the tile), 6-10 / 11-15 / 16-20 / 21-25 the
// turn of corner 0..3 among the lanes of its 32-slot group that hit the same pixel in that corner step.
size_t col2im_own_sched_bytes(const DcnGeom& g) {
  return col2im_own_geometry_ok(g) ? (size_t)g.N * g.K * ceil_div(g.Ho * g.Wo, OWN_TS) * OWN_TS * 4 : 0;
}

// one warp per (image, tap, tile of 64 positions)
__global__ void __launch_bounds__(256)
own_schedule_kernel(DcnGeom g, const SampleRec* __restrict__ plan, const SampleAux* __restrict__ aux,
                    unsigned* __restrict__ sched, int permute) {
  __shared__ unsigned char src_of[8][OWN_TS];
  const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
  const int HoWo = g.Ho * g.Wo, HW = g.H * g.W;
  const int ntile = (HoWo + OWN_TS - 1) / OWN_TS;
  const long long wid = (long long)blockIdx.x * 8 + wl;
  const long long nwork = (long long)g.N * g.K * ntile;
  if (wid >= nwork) return;                                      // whole warps leave; no block barrier below
  const int tile = (int)(wid % ntile);
  const int tap = (int)((wid / ntile) % g.K), n = (int)(wid / ntile / g.K);
  const int p0 = tile * OWN_TS;
  const unsigned lt = (1u << lane) - 1u;
  // residue of the base pixel (h_low, w_low) of the two samples of this lane; 8 = no preference
  int res[2], slot[2], base[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int p = p0 + h * 32 + lane;
    res[h] = 8;
    base[h] = 0;
    if (permute && p < HoWo) {
      const size_t ridx = ((size_t)n * HoWo + p) * g.K + tap;
      const int valid = __ldg(&aux[ridx].valid) & 15;
      if (valid) {
        const int4 px = __ldg(reinterpret_cast<const int4*>(plan + ridx));
        base[h] = ((valid & 1) ? px.x : (valid & 2) ? px.y - 1 : (valid & 4) ? px.z - g.W : px.w - g.W - 1) - n * HW;
        res[h] = base[h] & 7;
      }
    }
    slot[h] = -1;
  }
  // Two samples collide in a corner step exactly when their base pixels are equal (corner i = base + const_i), i.e.
  // inside a residue class.  Order each class by base pixel and deal it alternately to the two 32-slot groups of the
  // tile (a warp takes the groups one after the other): pairs of equal bases never meet in a corner step, only
  // triples still take turns.
  int ord[2] = {0, 0};
  if (permute) {
#pragma unroll
    for (int uh = 0; uh < 2; ++uh) {
      for (int ul = 0; ul < 32; ++ul) {
        const int ub = __shfl_sync(0xffffffffu, base[uh], ul), ur = __shfl_sync(0xffffffffu, res[uh], ul);
        const int u = uh * 32 + ul;
#pragma unroll
        for (int h = 0; h < 2; ++h)
          ord[h] += (ur == res[h] && (ub < base[h] || (ub == base[h] && u < h * 32 + lane))) ? 1 : 0;
      }
    }
  }
  unsigned occ0 = 0, occ1 = 0;                                   // slots 0-31 / 32-63 taken
  if (permute) {
#pragma unroll
    for (int h = 0; h < 2; ++h)
      if (res[h] < 8 && ord[h] < 8) slot[h] = 8 * ((ord[h] & 1) * 4 + (ord[h] >> 1)) + res[h];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      occ0 |= __reduce_or_sync(0xffffffffu, (slot[h] >= 0 && slot[h] < 32) ? 1u << slot[h] : 0u);
      occ1 |= __reduce_or_sync(0xffffffffu, (slot[h] >= 32) ? 1u << (slot[h] - 32) : 0u);
    }
  }
  // the rest (overflow of a residue class, unusable samples, positions past the map) fill the free slots in order
  const unsigned o0 = __ballot_sync(0xffffffffu, slot[0] < 0), o1 = __ballot_sync(0xffffffffu, slot[1] < 0);
  const int free0 = 32 - __popc(occ0);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    if (slot[h] < 0) {
      const int k = h == 0 ? __popc(o0 & lt) : __popc(o0) + __popc(o1 & lt);
      slot[h] = k < free0 ? (int)__fns(~occ0, 0, k + 1) : 32 + (int)__fns(~occ1, 0, k - free0 + 1);
    }
    src_of[wl][slot[h]] = (unsigned char)(h * 32 + lane);
  }
  __syncwarp();
  // turns: slot order, per 32-slot group and corner
#pragma unroll
  for (int gq = 0; gq < 2; ++gq) {
    const int src = src_of[wl][gq * 32 + lane], p = p0 + src;
    unsigned word = (unsigned)src;
    int valid = 0;
    int4 px = make_int4(0, 0, 0, 0);
    if (p < HoWo) {
      const size_t ridx = ((size_t)n * HoWo + p) * g.K + tap;
      valid = __ldg(&aux[ridx].valid) & 15;
      px = __ldg(reinterpret_cast<const int4*>(plan + ridx));
    }
    const int lp[4] = {px.x, px.y, px.z, px.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const bool ok = (valid >> i) & 1;
      const unsigned mm = __match_any_sync(0xffffffffu, ok ? lp[i] : -1 - lane);
      word |= (unsigned)__popc(mm & lt) << (6 + 5 * i);
    }
    sched[(((size_t)n * g.K + tap) * ntile + tile) * OWN_TS + gq * 32 + lane] = word;
  }
}

int col2im_own_schedule(const DcnGeom& g, const SampleRec* plan, const SampleAux* aux, unsigned* sched,
                        cudaStream_t stream) {
  int permute = 1;
  if (const char* e = getenv("KGDET_COL2IM_OWN_PERMUTE")) permute = atoi(e) != 0;
  const long long nwork = (long long)g.N * g.K * ceil_div(g.Ho * g.Wo, OWN_TS);
  own_schedule_kernel<<<(unsigned)((nwork + 7) / 8), 256, 0, stream>>>(g, plan, aux, sched, permute);
  KG_LAUNCH_CHECK("own_schedule_kernel");
  return KGDET_OK;
}

// cg: [M, ntaps * C] bf16 (taps tap0 .. tap0 + ntaps - 1 of the call); grid (tap splits, C / 32, N)
__global__ void __launch_bounds__(256, 1)
col2im_own_kernel(DcnGeom g, const __nv_bfloat16* __restrict__ cg, const __nv_bfloat16* __restrict__ in,
                  const SampleRec* __restrict__ plan, const SampleAux* __restrict__ aux,
                  const unsigned* __restrict__ sched, int tap0, int ntaps, int taps_per_split,
                  float* __restrict__ gin, float* __restrict__ gpart) {
  extern __shared__ __align__(16) unsigned char own_smem[];
  const int HW = g.H * g.W, HoWo = g.Ho * g.Wo;
  float4* acc = reinterpret_cast<float4*>(own_smem);                          // [HW][8 chunks of 4 channels]
  uint2* inp = reinterpret_cast<uint2*>(own_smem + (size_t)HW * 128);         // [HW][8 slots of 4 bf16]
  uint2* cgs = reinterpret_cast<uint2*>(own_smem + (size_t)HW * 192);         // [2][TS][8 slots]
  int4* pixs = reinterpret_cast<int4*>(cgs + 2 * OWN_TS * 8);                 // [2][TS]
  float4* auxs = reinterpret_cast<float4*>(pixs + 2 * OWN_TS);                // [2][TS]
  float4* parts = auxs + 2 * OWN_TS;                                          // [8 warps][TS]
  unsigned* scheds = reinterpret_cast<unsigned*>(parts + 8 * OWN_TS);         // [3][TS]
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int n = blockIdx.z, cb = blockIdx.y;
  const int tl_begin = blockIdx.x * taps_per_split;
  const int tl_end = (tl_begin + taps_per_split < ntaps) ? tl_begin + taps_per_split : ntaps;
  if (tl_begin >= tl_end) return;
  const int nbase = n * HW;
  const int ntile = (HoWo + OWN_TS - 1) / OWN_TS;
  const int total = (tl_end - tl_begin) * ntile;

  // schedule words of tile `it` -> scheds[it % 3] (two tiles ahead of the data they steer)
  auto issue_sched = [&](int it) {
    if (it < total && tid < OWN_TS) {
      const int tl = tl_begin + it / ntile, tt = it % ntile;
      cp_async4(&scheds[(it % 3) * OWN_TS + tid],
                sched + (((size_t)n * g.K + tap0 + tl) * ntile + tt) * OWN_TS + tid);
    }
  };
  // data of tile `it`, in lane-slot order (scheds[it % 3] must have landed)
  auto issue = [&](int it, int buf) {
    const int tl = tl_begin + it / ntile, p0 = (it % ntile) * OWN_TS;
    const unsigned* sw = scheds + (it % 3) * OWN_TS;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int piece = tid + r * 256, s = piece >> 3, j = piece & 7, p = p0 + (int)(sw[s] & 63u);
      if (p < HoWo)
        cp_async8(&cgs[buf * (OWN_TS * 8) + s * 8 + (j ^ ((s >> 1) & 7))],
                  cg + (((size_t)n * HoWo + p) * ntaps + tl) * g.C + cb * 32 + j * 4);
    }
    if (tid < 2 * OWN_TS) {
      const int s = tid & (OWN_TS - 1), p = p0 + (int)(sw[s] & 63u);
      if (p < HoWo) {
        const size_t ridx = ((size_t)n * HoWo + p) * g.K + tap0 + tl;
        if (tid < OWN_TS) cp_async16(&pixs[buf * OWN_TS + s], plan + ridx);
        else cp_async16(&auxs[buf * OWN_TS + s], aux + ridx);
      }
    }
  };

  issue_sched(0);
  issue_sched(1);
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int idx = tid; idx < HW * 8; idx += 256) acc[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int idx = tid; idx < HW * 8; idx += 256) {
    const int pix = idx >> 3, j = idx & 7;
    inp[pix * 8 + (j ^ ((pix >> 1) & 7))] =
        __ldg(reinterpret_cast<const uint2*>(in + ((size_t)(nbase + pix)) * g.C + cb * 32 + j * 4));
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  issue(0, 0);
  asm volatile("cp.async.commit_group;" ::: "memory");

  for (int it = 0; it < total; ++it) {
    const int buf = it & 1;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                       // tile `it` + the schedule of tile it + 1 visible; everybody is past
                                           // the reduction of tile it - 1
    if (it + 1 < total) issue(it + 1, buf ^ 1);
    issue_sched(it + 2);
    asm volatile("cp.async.commit_group;" ::: "memory");
    const int tl = tl_begin + it / ntile, p0 = (it % ntile) * OWN_TS;
    const unsigned* sw = scheds + (it % 3) * OWN_TS;

#pragma unroll 1
    for (int sg = 0; sg < OWN_TS / 32; ++sg) {
      const int s = sg * 32 + lane;
      const unsigned word = sw[s];
      const int p = p0 + (int)(word & 63u);
      int valid = 0;
      int4 px = make_int4(0, 0, 0, 0);
      float4 ax = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p < HoWo) {
        px = pixs[buf * OWN_TS + s];
        ax = auxs[buf * OWN_TS + s];
        valid = __float_as_int(ax.w) & 15;
      }
      float d[4] = {0.f, 0.f, 0.f, 0.f};
      if (__any_sync(0xffffffffu, valid != 0)) {
        float gv[4];
        unpack4(cgs[buf * (OWN_TS * 8) + s * 8 + (w ^ ((s >> 1) & 7))], gv);
        const float lh = ax.x, lw = ax.y, mk = ax.z;
        const float hh = 1.f - lh, hw = 1.f - lw;
        const float wgt[4] = {hh * hw * mk, hh * lw * mk, lh * hw * mk, lh * lw * mk};
        const int lpix[4] = {px.x - nbase, px.y - nbase, px.z - nbase, px.w - nbase};
        int turn[4], last[4];
        uint2 iv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {                 // nothing here depends on the accumulator: off the chain below
          const bool ok = (valid >> i) & 1;
          turn[i] = ok ? (int)((word >> (6 + 5 * i)) & 31u) : 0;
          last[i] = __reduce_max_sync(0xffffffffu, turn[i]);
          const int lp = ok ? lpix[i] : 0;
          iv[i] = inp[lp * 8 + (w ^ ((lp >> 1) & 7))];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const bool ok = (valid >> i) & 1;
          const int lp = ok ? lpix[i] : 0;
          const float wi = wgt[i];
          float4* a = acc + lp * 8 + (w ^ (lp & 7));
          if (last[i] == 0) {
            if (ok) {
              float4 v = *a;
              v.x = fmaf(gv[0], wi, v.x); v.y = fmaf(gv[1], wi, v.y);
              v.z = fmaf(gv[2], wi, v.z); v.w = fmaf(gv[3], wi, v.w);
              *a = v;
            }
          } else {                                    // lanes that share a pixel take turns, lowest lane first
            for (int r = 0; r <= last[i]; ++r) {
              if (ok && turn[i] == r) {
                float4 v = *a;
                v.x = fmaf(gv[0], wi, v.x); v.y = fmaf(gv[1], wi, v.y);
                v.z = fmaf(gv[2], wi, v.z); v.w = fmaf(gv[3], wi, v.w);
                *a = v;
              }
              __syncwarp();
            }
          }
          __syncwarp();                               // corner i of every lane is in memory before corner i + 1
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if ((valid >> i) & 1) {
            float v[4];
            unpack4(iv[i], v);
            d[i] = fmaf(gv[3], v[3], fmaf(gv[2], v[2], fmaf(gv[1], v[1], gv[0] * v[0])));
          }
        }
      }
      parts[w * OWN_TS + s] = make_float4(d[0], d[1], d[2], d[3]);
    }
    __syncthreads();
    if (tid < OWN_TS) {
      const int p = p0 + (int)(sw[tid] & 63u);
      if (p < HoWo) {
        const float4 ax = auxs[buf * OWN_TS + tid];
        float dy = 0.f, dx = 0.f, sv = 0.f;
        if ((__float_as_int(ax.w) & 15) != 0) {
          float4 dsum = parts[tid];
#pragma unroll
          for (int ww = 1; ww < 8; ++ww) {
            const float4 q = parts[ww * OWN_TS + tid];
            dsum.x += q.x; dsum.y += q.y; dsum.z += q.z; dsum.w += q.w;
          }
          const float lh = ax.x, lw = ax.y, mk = ax.z, hh = 1.f - lh, hw = 1.f - lw;
          // d(sample)/dy, d(sample)/dx and the sample itself (deform_conv_cuda_kernel.cu:144-187)
          dy = (-hw * dsum.x - lw * dsum.y + hw * dsum.z + lw * dsum.w) * mk;
          dx = (-hh * dsum.x + hh * dsum.y - lh * dsum.z + lh * dsum.w) * mk;
          sv = hh * hw * dsum.x + hh * lw * dsum.y + lh * hw * dsum.z + lh * lw * dsum.w;
        }
        float* dst = gpart + (((size_t)cb * g.N + n) * 3 * g.K + 3 * (tap0 + tl)) * HoWo + p;
        dst[0] = dy;
        dst[HoWo] = dx;
        dst[2 * (size_t)HoWo] = sv;
      }
    }
  }
  __syncthreads();
  for (int idx = tid; idx < HW * 8; idx += 256) {
    const int pix = idx >> 3, j = idx & 7;
    const float4 v = acc[pix * 8 + (j ^ (pix & 7))];
    if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f)
      atomicAdd(reinterpret_cast<float4*>(gin + ((size_t)(nbase + pix)) * g.C + cb * 32 + j * 4), v);
  }
}

// grad_offset[n, 2 tap + q, p] / grad_mask[n, tap, p] = sum over the C / 32 channel blocks, in fixed order
__global__ void __launch_bounds__(256)
col2im_own_finish_kernel(int N, int K, int HoWo, int nblocks, const float* __restrict__ gpart,
                         float* __restrict__ goff, float* __restrict__ gmask) {
  const long long per_block = (long long)N * 3 * K * HoWo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per_block;
       i += (long long)blockDim.x * gridDim.x) {
    const int p = (int)(i % HoWo);
    const long long r = i / HoWo;
    const int q = (int)(r % 3);
    const long long nt = r / 3;                       // n * K + tap
    if (q == 2 && !gmask) continue;
    float s = gpart[i];
    for (int b = 1; b < nblocks; ++b) s += gpart[(long long)b * per_block + i];
    if (q < 2) {
      const long long n = nt / K, tap = nt - n * K;
      goff[((n * 2 * K) + 2 * tap + q) * HoWo + p] = s;
    } else {
      gmask[nt * HoWo + p] = s;
    }
  }
}

int col2im_own_finish(const DcnGeom& g, const float* gpart, float* goff, float* gmask, cudaStream_t stream) {
  const long long per_block = (long long)g.N * 3 * g.K * g.Ho * g.Wo;
  long long blocks = (per_block + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  col2im_own_finish_kernel<<<(unsigned)blocks, 256, 0, stream>>>(g.N, g.K, g.Ho * g.Wo, g.C / 32, gpart, goff, gmask);
  KG_LAUNCH_CHECK("col2im_own_finish_kernel");
  return KGDET_OK;
}

// One chunk of taps.  gpart: staging tensor of col2im_own_part_bytes(); col2im_own_finish after the last chunk.
int col2im_own(const DcnGeom& g, const __nv_bfloat16* cg, const __nv_bfloat16* in_nhwc, const SampleRec* plan,
               const SampleAux* aux, const unsigned* sched, int tap0, int ntaps, float* gin_nhwc, float* gpart,
               cudaStream_t stream) {
  KG_CUDA(cudaFuncSetAttribute(col2im_own_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OWN_SMEM_MAX));
  const int base = g.N * (g.C / 32);
  int splits = (num_sms() + base / 2) / base;            // fill the machine when the batch is small
  if (splits < 1) splits = 1;
  if (splits > ntaps) splits = ntaps;
  const int tps = ceil_div(ntaps, splits);
  splits = ceil_div(ntaps, tps);
  col2im_own_kernel<<<dim3((unsigned)splits, (unsigned)(g.C / 32), (unsigned)g.N), 256, own_smem_bytes(g), stream>>>(
      g, cg, in_nhwc, plan, aux, sched, tap0, ntaps, tps, gin_nhwc, gpart);
  KG_LAUNCH_CHECK("col2im_own_kernel");
  return KGDET_OK;
}

}  // namespace kgdet
