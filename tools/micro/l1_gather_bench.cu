// Microbenchmark: L1-resident gather bandwidth per SM for the access shapes the fused DCN producer could use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/l1_gather_bench tools/micro/l1_gather_bench.cu
// Every warp-level load instruction reads whole 128-byte "pixel slabs" (one L1 line each) at pseudo-random
// positions of a small per-CTA region (L1-resident after the first pass):
//   VEC = 16: 8 lanes per slab, 4 slabs per instruction (the v3/v4 kernels)     VEC = 8: 16 lanes per slab, 2 slabs
//   VEC = 4: 32 lanes per slab, 1 slab per instruction                          SEQ = 1: slabs of one instruction contiguous
// Output: bytes per clock per SM (clock64 around the loop, max over CTAs' elapsed).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int VEC> struct V;
template <> struct V<16> { using T = uint4; };
template <> struct V<8> { using T = uint2; };
template <> struct V<4> { using T = uint32_t; };

__device__ __forceinline__ uint4 ld_plain(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.ca.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ uint2 ld_plain(const uint2* p) {
  uint2 v;
  asm volatile("ld.global.ca.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t ld_plain(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t fold(uint4 v) { return v.x ^ v.y ^ v.z ^ v.w; }
__device__ __forceinline__ uint32_t fold(uint2 v) { return v.x ^ v.y; }
__device__ __forceinline__ uint32_t fold(uint32_t v) { return v; }

template <int VEC, int SEQ, int NC, int STRIDE>
__global__ void __launch_bounds__(512, 1) gather_kernel(const unsigned char* __restrict__ buf, int slabs_per_cta,
                                                        int iters, uint32_t* sink, long long* cycles) {
  using T = typename V<VEC>::T;
  constexpr int LPS = 128 / VEC;                 // lanes per slab
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / LPS, within = lane % LPS;
  const unsigned char* base = buf + (size_t)blockIdx.x * slabs_per_cta * 128 * STRIDE + within * VEC;
  uint32_t acc = 0;
  uint32_t state = 1234567u * (warp * 8 + sub + 1) + blockIdx.x;
  // warm the L1
  for (int i = threadIdx.x; i < slabs_per_cta * 8 * STRIDE; i += blockDim.x)
    acc ^= fold(NC ? __ldg(reinterpret_cast<const uint4*>(buf + (size_t)blockIdx.x * slabs_per_cta * 128 * STRIDE) + i)
                   : ld_plain(reinterpret_cast<const uint4*>(buf + (size_t)blockIdx.x * slabs_per_cta * 128 * STRIDE) + i));
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    T v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      state = state * 1664525u + 1013904223u;
      uint32_t slab = (state >> 8) % (uint32_t)slabs_per_cta;
      if (SEQ) slab = ((slab / (32 / LPS)) * (32 / LPS) + sub) % (uint32_t)slabs_per_cta;
      const T* p = reinterpret_cast<const T*>(base + (size_t)slab * 128 * STRIDE);
      v[u] = NC ? __ldg(p) : ld_plain(p);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) acc ^= fold(v[u]);
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
}

template <int VEC, int SEQ, int NC, int STRIDE = 1>
static void run(const char* name, const unsigned char* buf, int slabs, uint32_t* sink, long long* dcyc) {
  const int iters = 2000, grid = 148;
  gather_kernel<VEC, SEQ, NC, STRIDE><<<grid, 512>>>(buf, slabs, iters, sink, dcyc);
  gather_kernel<VEC, SEQ, NC, STRIDE><<<grid, 512>>>(buf, slabs, iters, sink, dcyc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  long long h[148];
  cudaMemcpy(h, dcyc, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  const double bytes = 512.0 * 8 * iters * VEC;
  printf("{\"shape\": \"%s\", \"slabs_per_cta\": %d, \"region_kb\": %d, \"bytes_per_clk_per_sm\": %.1f, \"cycles_per_128B\": %.2f}\n",
         name, slabs, slabs * 128 / 1024, bytes / mx, mx / (bytes / 128.0));
}

int main() {
  unsigned char* buf;
  uint32_t* sink;
  long long* dcyc;
  const size_t total = (size_t)148 * 2048 * 128 * 4;
  cudaMalloc(&buf, total);
  cudaMemset(buf, 1, total);
  cudaMalloc(&sink, 4);
  cudaMalloc(&dcyc, 148 * 8);
  for (int slabs : {256, 512, 2048}) {         // 32 KB / 64 KB (L1-resident) and 256 KB (L2-resident) per CTA
    run<16, 0, 1>("ldg128_nc_8lanes_x4slabs_random", buf, slabs, sink, dcyc);
    run<16, 0, 0>("ld128_8lanes_x4slabs_random", buf, slabs, sink, dcyc);
    run<16, 1, 1>("ldg128_nc_contiguous512B", buf, slabs, sink, dcyc);
    run<8, 0, 1>("ldg64_nc_16lanes_x2slabs_random", buf, slabs, sink, dcyc);
    run<8, 1, 1>("ldg64_nc_contiguous256B", buf, slabs, sink, dcyc);
    run<4, 0, 1>("ldg32_nc_32lanes_x1slab", buf, slabs, sink, dcyc);
    run<16, 0, 1, 4>("ldg128_nc_8lanes_x4slabs_random_stride512B", buf, slabs / 4, sink, dcyc);
    run<16, 0, 1, 2>("ldg128_nc_8lanes_x4slabs_random_stride256B", buf, slabs / 2, sink, dcyc);
  }
  return 0;
}
