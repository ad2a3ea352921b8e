// Definitions shared by the host side (dcn_umma.cu: weight packing, parameter set-up) and the kernel
// (dcn_umma_stream.cu) of the fused tcgen05 forward.
#pragma once
#include "dcn.cuh"

namespace kgdet {

static constexpr int BM = 128;                 // positions per CTA tile (UMMA M)
static constexpr int A_TILE_BYTES = BM * 128;  // 128 rows x 128 B

enum { MODE_BF16 = 0, MODE_TF32X3 = 1, MODE_TF32 = 2 };

template <int MODE> struct ModeTraits;
template <> struct ModeTraits<MODE_BF16>   { static constexpr int BK = 64, A_TILES = 1, B_TILES = 1, ELEM = 2; };
template <> struct ModeTraits<MODE_TF32X3> { static constexpr int BK = 32, A_TILES = 2, B_TILES = 2, ELEM = 4; };
template <> struct ModeTraits<MODE_TF32>   { static constexpr int BK = 32, A_TILES = 1, B_TILES = 1, ELEM = 4; };

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// ---- the fused kernel ----------------------------------------------------------------------
struct UmmaParams {
  const void* in;            // channel-blocked planes [C/BK][pixels][BK], bf16 (MODE_BF16) or fp32 (TF32 modes)
  size_t plane_bytes;        // distance between planes
  const SampleRec16* plan;   // [K][rows_padded] (tap-major: the records of consecutive rows are contiguous)
  const unsigned char* wp;   // packed weights
  const float* bias;         // [Cout] or NULL
  void* out;                 // NCHW
  int M, C, W, Cout, K, HoWo, rows_padded;
  int out_coff, out_ctot, relu;   // channel slice of the output tensor, fused ReLU
  int out_nhwc;              // KGDET_LAYOUT_*: TILED / TILED_SPLIT = UMMA-tiled bf16 rows (pointwise_umma.cu)
  int nkb;                   // (C / BK) * K
  uint32_t idesc;
  uint32_t tmem_cols;
  // development hook (kgdet_dcn_set_timeline): per CTA 4 * nkb + 8 clock64() stamps, or NULL
  //   [0] kernel entry, [1] set-up done, [2 + j] control lane saw k-block j full, [2 + nkb] accumulator ready,
  //   [3 + nkb] epilogue done, [4 + nkb + j] producer thread 0 arrived for k-block j,
  //   [4 + 2 nkb + j] producer thread 0 acquired the stage of k-block j, [4 + 3 nkb + j] its stores are issued
  long long* timeline;
  // split-K over CTAs (small maps: fewer than half as many tiles as SMs): blockIdx.y owns k-blocks
  // [y * kb_per_split, ...), and writes its raw fp32 accumulator tile to partial[y][m][Cout]; a second
  // kernel adds the splits, bias and ReLU and writes the output layout.  NULL = no split.
  float* partial;
  int kb_per_split, m_pad;
};



template <typename T> __device__ __forceinline__ void st_out(T* p, float v);
template <> __device__ __forceinline__ void st_out<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_out<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  *p = __float2bfloat16(v);
}

template <typename T> __device__ __forceinline__ void st_out4(T* p, float a, float b, float c, float d);
template <> __device__ __forceinline__ void st_out4<float>(float* p, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
template <> __device__ __forceinline__ void st_out4<__nv_bfloat16>(__nv_bfloat16* p, float a, float b, float c, float d) {
  const __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
  uint2 v;
  v.x = *reinterpret_cast<const uint32_t*>(&lo);
  v.y = *reinterpret_cast<const uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p) = v;
}

__device__ __forceinline__ uint32_t bf162_bcast(float w) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %1;" : "=r"(r) : "f"(w));
  return r;
}
__device__ __forceinline__ uint32_t bf162_mul(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t bf162_fma(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void fma_bf16x2_f32(float w, uint32_t packed, float& a0, float& a1) {
  a0 = fmaf(w, __uint_as_float(packed << 16), a0);
  a1 = fmaf(w, __uint_as_float(packed & 0xffff0000u), a1);
}

// bilinear weights of the four corners from a compact plan record (see SampleRec16)
__device__ __forceinline__ void decode_rec(const float4& r, float (&w)[4]) {
  const uint32_t lhb = __float_as_uint(r.y), lwb = __float_as_uint(r.z);
  const float lh = __uint_as_float(lhb & ~3u), lw = __uint_as_float(lwb & ~3u);
  const float fh0 = (lhb & 1u) ? (1.f - lh) : 0.f, fh1 = (lhb & 2u) ? lh : 0.f;
  const float fw0 = (lwb & 1u) ? (1.f - lw) * r.w : 0.f, fw1 = (lwb & 2u) ? lw * r.w : 0.f;
  w[0] = fh0 * fw0; w[1] = fh0 * fw1; w[2] = fh1 * fw0; w[3] = fh1 * fw1;
}

// streaming kernel (dcn_umma_stream.cu)
int umma_stream_forward(const DcnGeom& g, const UmmaParams& p, int mode, bool pair, int out_dtype,
                        cudaStream_t stream, int splits = 1);

}  // namespace kgdet
