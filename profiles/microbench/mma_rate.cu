// microbenchmark: cycles per tcgen05.mma (kind::tf32, SS operands) for several tile shapes, optionally with
// concurrent shared-memory store traffic from other warps.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../bayesianneuralnetworks_b200/csrc/umma.cuh"
using namespace bnn::umma;

template <bool PAIR>
__global__ void __launch_bounds__(256, 1) rate_kernel(int n, int iters, int writers, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  for (int i = tid; i < 48 * 1024; i += 256) sts32(base + i * 4, 0u);
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) { if (PAIR) tmem_alloc_pair(&slot, 512); else tmem_alloc(&slot, 512); }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&slot);
  const uint32_t rank = PAIR ? cluster_ctarank() : 0;
  long long t0 = 0, t1 = 0;
  if (tid == 0 && rank == 0) {
    const uint32_t idesc = make_idesc_tf32(PAIR ? 256 : 128, n);
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t a = base + (it & 3) * 16384, b = base + 65536 + (it & 3) * 32768;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        if (PAIR) mma_tf32_pair(tmem, make_smem_desc(a + ks * 32), make_smem_desc(b + ks * 32), idesc, true);
        else mma_tf32(tmem, make_smem_desc(a + ks * 32), make_smem_desc(b + ks * 32), idesc, true);
      }
    }
    if (PAIR) mma_commit_pair(&bar, 3); else mma_commit(&bar);
    mbar_wait(&bar, 0);
    t1 = clock64();
    out[blockIdx.x] = (unsigned long long)(t1 - t0);
  } else if (warp >= 4 && warp < 4 + writers) {
    // concurrent store traffic into an unrelated region
    const uint32_t w = base + 196608 - 0 * 0 + 0;  // not used by the MMAs: offsets [192K, 192K+16K)
    for (int it = 0; it < iters * 2; ++it)
      sts128(base + 160 * 1024 + ((tid & 127) * 16 + (it & 7) * 2048), it, it, it, it);
    (void)w;
  }
  if (PAIR) { if (tid == 0 && rank == 1) { mbar_wait(&bar, 0); } }
  tc_fence_before_sync();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 0) { tc_fence_after_sync(); if (PAIR) tmem_dealloc_pair(tmem, 512); else tmem_dealloc(tmem, 512); }
}

int main() {
  unsigned long long* out;
  cudaMalloc(&out, 148 * 8);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(rate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(rate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 2000;
  for (int pair = 0; pair < 2; ++pair)
    for (int n : {64, 128, 256})
      for (int writers : {0, 4}) {
        cudaMemset(out, 0, 148 * 8);
        if (pair) {
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = dim3(148); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
          cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
          cfg.attrs = &at; cfg.numAttrs = 1;
          cudaLaunchKernelEx(&cfg, rate_kernel<true>, n, iters, writers, out);
        } else {
          rate_kernel<false><<<148, 256, smem>>>(n, iters, writers, out);
        }
        cudaError_t e = cudaDeviceSynchronize();
        unsigned long long h[148];
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        unsigned long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
        const double per = (double)mx / (iters * 4.0);
        const double macs = (pair ? 256.0 : 128.0) * n * 8;
        printf("pair=%d N=%3d writers=%d: %s  %.1f cycles/MMA  -> %.0f MAC/clk/SM\n", pair, n, writers, cudaGetErrorString(e), per,
               macs / per / (pair ? 2 : 1));
      }
  return 0;
}
