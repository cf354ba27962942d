// common.cuh — shared device/host helpers of libbnn_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <atomic>
#include "../../include/bnn_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libbnn_b200 is written for sm_100a only"
#endif

namespace bnn {

// ---------------------------------------------------------------------------------------------
// host side: status + thread-local error text (bnn_last_error_string)
// ---------------------------------------------------------------------------------------------
char* error_buffer();                       // defined in api.cu (thread_local, 512 bytes)
int fail(int code, const char* fmt, ...);   // formats into error_buffer(), returns code
int check_device();                         // BNN_OK iff current device is CC 10.x (cached per device)
int sm_count();                             // multiprocessor count of the current device (cached)

#define BNN_CUDA_OK(expr)                                                                   \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      return ::bnn::fail(BNN_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                         __FILE__, __LINE__);                                               \
  } while (0)

#define BNN_REQUIRE(cond, code, ...)                       \
  do {                                                     \
    if (!(cond)) return ::bnn::fail((code), __VA_ARGS__);  \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Opt-in to more than 48 KiB of dynamic shared memory.  The attribute is PER DEVICE, so the "already done" state is a
// bit per device ordinal (relaxed atomics: cudaFuncSetAttribute is idempotent, two racing threads both set it).
struct SmemOptIn {
  std::atomic<uint64_t> done{0};
};
template <typename Kernel>
inline int allow_dynamic_smem(Kernel kernel, size_t bytes, SmemOptIn* state) {
  int dev = 0;
  BNN_CUDA_OK(cudaGetDevice(&dev));
  const uint64_t bit = uint64_t(1) << (dev & 63);
  if (state->done.load(std::memory_order_relaxed) & bit) return BNN_OK;
  BNN_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
  state->done.fetch_or(bit, std::memory_order_relaxed);
  return BNN_OK;
}

// ---------------------------------------------------------------------------------------------
// single-MUFU approximations (flush-to-zero forms: no denormal range fix-ups around the MUFU op; every use below
// keeps its arguments in the normal range or is indifferent to a flushed result)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_ftz(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float lg2_ftz(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rcp_ftz(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float exp_fast(float x) { return ex2_ftz(x * 1.4426950408889634f); }
__device__ __forceinline__ float log_fast(float x) { return lg2_ftz(x) * 0.6931471805599453f; }

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11) keyed by (seed, step, tensor, sample, element/4)
// ---------------------------------------------------------------------------------------------
struct RngKey {        // resolved on the device at kernel entry (adds *step_dev when given)
  uint32_t k0, k1;     // Philox key
  uint32_t tensor_id;  // counter word 2
  uint32_t step_lo;    // counter word 3
  uint64_t elem_offset;  // added to the element index (slice of a larger tensor)
  const float* row_sign; // rank-one sign noise (Flipout): eps(n, k) = row_sign[s * rows + n] * col_sign[s * cols + k]
  const float* col_sign;
  int rows, cols;
};

__device__ __forceinline__ RngKey resolve_rng(const bnn_rng& r) {
  uint64_t step = r.step;
  if (r.step_dev != nullptr) step += *r.step_dev;
  RngKey k;
  k.k0 = static_cast<uint32_t>(r.seed);
  k.k1 = static_cast<uint32_t>(r.seed >> 32) ^ static_cast<uint32_t>(step >> 32);
  k.tensor_id = r.tensor_id;
  k.step_lo = static_cast<uint32_t>(step);
  k.elem_offset = r.elem_offset;
  k.row_sign = r.row_sign; k.col_sign = r.col_sign; k.rows = r.rows; k.cols = r.cols;
  return k;
}

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
    k0 += W0;
    k1 += W1;
  }
  return c;
}

// uniform in (0, 1]: r * 2^-32 + 2^-33 evaluated in fp32 (the product is exact, one rounding)
__device__ __forceinline__ float u01(uint32_t r) {
  return fmaf(static_cast<float>(r), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
}

// Box-Muller on a pair of words: (radius*cos, radius*sin); u1 in [2^-33, 1] is a normal number, the angle in (-pi, pi]
// is where sin/cos.approx are most accurate
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  const float u1 = u01(a), u2 = u01(b);
  float radius, s, c;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(radius) : "f"(lg2_ftz(u1) * -1.3862943611198906f));   // sqrt(-2 ln u1)
  const float angle = fmaf(u2, 6.283185307179586f, -3.141592653589793f);
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(angle));
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(angle));
  return make_float2(radius * c, radius * s);
}

// eps for elements 4*group .. 4*group+3 of (tensor, sample); group counts from the slice start and
// elem_offset must be a multiple of 4 on this path
__device__ __forceinline__ float4 eps4(const RngKey& k, uint32_t sample, uint32_t group) {
  const uint32_t g = group + static_cast<uint32_t>(k.elem_offset >> 2);
  const uint4 r = philox4x32_10(make_uint4(g, sample, k.tensor_id, k.step_lo), k.k0, k.k1);
  const float2 a = box_muller(r.x, r.y);
  const float2 b = box_muller(r.z, r.w);
  return make_float4(a.x, a.y, b.x, b.y);
}

__device__ __forceinline__ float pick4(const float4& v, int j) {
  return j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w));
}

// eps of one element (slow path: one Philox call per element)
__device__ __forceinline__ float eps1(const RngKey& k, uint32_t sample, uint64_t element) {
  const uint64_t e = element + k.elem_offset;
  const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(e >> 2), sample, k.tensor_id, k.step_lo), k.k0, k.k1);
  const float2 n = (e & 2) ? box_muller(r.z, r.w) : box_muller(r.x, r.y);
  return (e & 1) ? n.y : n.x;
}

// ---------------------------------------------------------------------------------------------
// math: torch-compatible softplus (beta 1, threshold 20) and sigma = 1e-10 + softplus(rho)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float softplus_exact(float x) {      // mirrors ATen's CUDA kernel
  return x > 20.0f ? x : log1pf(expf(x));
}
__device__ __forceinline__ float stddev_exact(float rho) { return __fadd_rn(1e-10f, softplus_exact(rho)); }

__device__ __forceinline__ float sigmoid_fast(float x) {
  // 1 / (1 + exp(-x)); saturates cleanly for |x| large
  return __fdividef(1.0f, 1.0f + __expf(-x));
}

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

// ---------------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// streaming 128-bit load that does not allocate in L1
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

// streaming 128-bit store (evict-first: the line is not needed again by this kernel)
__device__ __forceinline__ void st_stream4(float* p, const float4& v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

}  // namespace bnn
