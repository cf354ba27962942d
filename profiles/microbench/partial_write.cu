// Microbenchmark: do partial-sector (masked) stores cost DRAM read fills on B200?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ bool sel(uint32_t i, uint32_t thr) { return hash32(i) < thr; }

__global__ void k_read(const float4* a, const float4* b, size_t n4, float* out) {
  float s = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 x = __ldcs(a + i), y = __ldcs(b + i);
    s += x.x + x.y + x.z + x.w + y.x + y.y + y.z + y.w;
  }
  if (s == 123.456f) *out = s;
}
__global__ void k_write_full(float4* a, float4* b, size_t n4) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    a[i] = make_float4(0, 0, 0, 0);
    b[i] = make_float4(-30, -30, -30, -30);
  }
}
// masked scalar stores, nothing read
__global__ void k_write_masked(float* a, float* b, size_t n4, uint32_t thr) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const size_t e = i * 4 + q;
      if (sel((uint32_t)e, thr)) { a[e] = 0.f; b[e] = -30.f; }
    }
  }
}
// read + whole-vector write (what apply does today)
__global__ void k_rmw(float4* a, float4* b, size_t n4, uint32_t thr) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 x = __ldcs(a + i), y = __ldcs(b + i);
    const uint32_t e = (uint32_t)(i * 4);
    if (sel(e, thr)) { x.x = 0; y.x = -30; }
    if (sel(e + 1, thr)) { x.y = 0; y.y = -30; }
    if (sel(e + 2, thr)) { x.z = 0; y.z = -30; }
    if (sel(e + 3, thr)) { x.w = 0; y.w = -30; }
    a[i] = x; b[i] = y;
  }
}
// bit mask read (1 bit / element) + masked stores
__global__ void k_mask_build(uint32_t* bits, size_t n, uint32_t thr) {
  for (size_t w = blockIdx.x * (size_t)blockDim.x + threadIdx.x; w < n / 32; w += (size_t)gridDim.x * blockDim.x) {
    uint32_t m = 0;
    for (int q = 0; q < 32; ++q) m |= (sel((uint32_t)(w * 32 + q), thr) ? 1u : 0u) << q;
    bits[w] = m;
  }
}
__global__ void k_write_bits(float* a, float* b, const uint32_t* bits, size_t n) {
  // one warp handles 32*32 = 1024 consecutive elements: lane l covers element 32*j + l for j = 0..31
  const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  for (size_t g = warp; g < n / 1024; g += nwarps) {
    const uint32_t mine = bits[g * 32 + lane];
#pragma unroll 8
    for (int j = 0; j < 32; ++j) {
      const uint32_t m = __shfl_sync(0xffffffffu, mine, j);
      if ((m >> lane) & 1u) { a[g * 1024 + j * 32 + lane] = 0.f; b[g * 1024 + j * 32 + lane] = -30.f; }
    }
  }
}

int main() {
  const size_t n = size_t(1) << 28;
  float *a, *b, *out;
  uint32_t* bits;
  cudaMalloc(&a, n * 4); cudaMalloc(&b, n * 4); cudaMalloc(&out, 4); cudaMalloc(&bits, n / 8);
  cudaMemset(a, 1, n * 4); cudaMemset(b, 1, n * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = 148 * 8, th = 256;
  const double gib = double(n) * 8 / 1e9;
  for (int pi = 0; pi < 3; ++pi) {
    const double p = pi == 0 ? 0.75 : (pi == 1 ? 0.5 : 0.95);
    const uint32_t thr = (uint32_t)(p * 4294967296.0);
    k_mask_build<<<grid, th>>>(bits, n, thr);
    for (int which = 0; which < 5; ++which) {
      float best = 1e9f;
      for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        switch (which) {
          case 0: k_read<<<grid, th>>>((const float4*)a, (const float4*)b, n / 4, out); break;
          case 1: k_write_full<<<grid, th>>>((float4*)a, (float4*)b, n / 4); break;
          case 2: k_write_masked<<<grid, th>>>(a, b, n / 4, thr); break;
          case 3: k_rmw<<<grid, th>>>((float4*)a, (float4*)b, n / 4, thr); break;
          case 4: k_write_bits<<<grid, th>>>(a, b, bits, n); break;
        }
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
      }
      const char* names[] = {"read 8B/pair", "write full 8B/pair", "masked store (no read)", "read + vector write", "bitmask + masked store"};
      printf("p=%.2f %-26s %.3f ms  (%.0f GB/s if 8 B/pair)\n", p, names[which], best, gib / best * 1e3);
    }
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
