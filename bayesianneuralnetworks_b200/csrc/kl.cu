// kl.cu — the bandwidth-bound KL sweep over the variational parameters (mu, rho):
//   bnn_kl     closed-form KL(N(mu, sigma) || N(loc, scale)) per tensor, optional gradients
//              (reference loss.py:16-38, torch/distributions/kl.py:468-471)
//   bnn_prune  log-density-at-zero key, exact top-k select, masked overwrite
//              (reference prune/prune.py:10-17, torch/distributions/normal.py:87-102)
// Many tensors are served by one launch: the host passes a table BY VALUE in the kernel
// parameters (no device allocation, no H2D copy), and a persistent grid walks fixed-size
// chunks across all tensors.  Loads are 128-bit; every chunk issues all of its loads before the
// first use.  Reductions: registers -> warp shuffle -> shared -> one slot per (tensor, block) in
// the workspace; the last block to finish adds the slots in a fixed order (deterministic).
#include "common.cuh"

namespace bnn {
namespace {

constexpr int kThreads = 256;
constexpr int kVecPerThread = 4;                               // float4 loads per thread per chunk
constexpr int kChunk = kThreads * 4 * kVecPerThread;           // 4096 elements
constexpr int kMaxTensors = 24;                                // per launch (table travels by value)

// ================================================================================== KL
struct KlDesc {
  const float* mu;
  const float* rho;
  float* gmu;
  float* grho;
  int64_t numel;
  int64_t chunk_begin;   // first global chunk id of this tensor
  float loc, inv_scale, log_scale, coeff;
  int vec;               // 16-byte aligned bases
  int pad;
};
struct KlTable {
  KlDesc t[kMaxTensors];
  int n;
  int pad;
  int64_t total_chunks;
};

struct KlTerm { float kl, gmu, grho; };

// One element, general case (any rho): precise softplus, branches allowed (rare path).
template <bool kGrad>
__device__ __forceinline__ KlTerm kl_element_slow(float mu, float rho, float loc, float inv_scale,
                                                  float log_scale, float coeff) {
  const float sp = rho > 20.0f ? rho : log1pf(__expf(rho));
  const float sigma = 1e-10f + sp;
  const float r = sigma * inv_scale;
  const float d = (mu - loc) * inv_scale;
  KlTerm out;
  out.kl = 0.5f * (fmaf(r, r, fmaf(d, d, -1.0f))) - (__logf(sigma) - log_scale);
  if (kGrad) {
    const float sig = __fdividef(1.0f, 1.0f + __expf(-rho));
    out.gmu = coeff * d * inv_scale;
    out.grho = coeff * (r * inv_scale - __fdividef(1.0f, sigma)) * sig;
  } else {
    out.gmu = 0.f;
    out.grho = 0.f;
  }
  return out;
}

// One element with e^rho <= 1/4 (rho <= -1.3863: every freshly initialised or trained posterior,
// rho ~ -2, and every pruned entry, rho = -30).  Branch-free fast-math formulation (DESIGN.md "KL
// arithmetic"): e = exp(rho) once; softplus = log1p(e) = 2 atanh(z), z = e / (2 + e), by its odd series
// (|z| <= 1/9: truncation < 1e-11 relative); log(sigma) by MUFU.LG2; sigmoid = e / (1 + e);
// 1/scale and log(scale) come from the host.
template <bool kGrad>
__device__ __forceinline__ KlTerm kl_element_fast(float mu, float rho, float loc, float inv_scale,
                                                  float log_scale, float coeff) {
  const float e = exp_fast(rho);                       // flushed to 0 below rho ~ -87: sigma = 1e-10, as in fp32 torch
  const float z = e * rcp_ftz(2.0f + e);
  const float z2 = z * z;
  float p = fmaf(z2, 0.1111111111f, 0.1428571429f);
  p = fmaf(z2, p, 0.2f);
  p = fmaf(z2, p, 0.3333333333f);
  p = fmaf(z2, p, 1.0f);
  const float sigma = fmaf(2.0f * z, p, 1e-10f);       // >= 1e-10: normal range for rcp / lg2
  const float r = sigma * inv_scale;
  const float d = (mu - loc) * inv_scale;
  KlTerm out;
  out.kl = fmaf(0.5f, fmaf(r, r, fmaf(d, d, -1.0f)), log_scale) - log_fast(sigma);
  if (kGrad) {
    const float sig = e * rcp_ftz(1.0f + e);
    out.gmu = coeff * d * inv_scale;
    out.grho = coeff * (r * inv_scale - rcp_ftz(sigma)) * sig;
  } else {
    out.gmu = 0.f;
    out.grho = 0.f;
  }
  return out;
}

constexpr float kFastRhoMax = -1.3862944f;   // ln(1/4)

// chunks are visited in increasing order by every block, so the owning tensor only moves forward
__device__ __forceinline__ int find_tensor(const int64_t* chunk_begin, int n, int64_t chunk, int t = 0) {
#pragma unroll 1
  while (t + 1 < n && chunk >= chunk_begin[t + 1]) ++t;
  return t;
}

__device__ __forceinline__ double block_sum_double(double v, double* smem8) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) smem8[threadIdx.x >> 5] = v;
  __syncthreads();
  double tot = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) tot += smem8[w];
  }
  return tot;   // valid in thread 0
}

template <bool kGrad>
__global__ void __launch_bounds__(kThreads)
kl_kernel(const __grid_constant__ KlTable tab, double* __restrict__ kl_sum, float* __restrict__ kl_total,
          int accumulate_total, const float* __restrict__ grad_scale_dev, double* __restrict__ partials,
          unsigned int* __restrict__ done_counter) {
  const bool want_sums = kl_sum != nullptr || kl_total != nullptr;
  __shared__ int64_t s_begin[kMaxTensors];
  __shared__ double s_red[kThreads / 32];
  __shared__ bool s_last;
  if (threadIdx.x < tab.n) s_begin[threadIdx.x] = tab.t[threadIdx.x].chunk_begin;
  // every (tensor, block) slot is written exactly once per launch: zero now, overwrite on flush
  if (want_sums)
    for (int t = threadIdx.x; t < tab.n; t += kThreads) partials[static_cast<int64_t>(t) * gridDim.x + blockIdx.x] = 0.0;
  __syncthreads();
  const float gscale = (kGrad && grad_scale_dev != nullptr) ? *grad_scale_dev : 1.0f;

  int cur = -1;
  float acc = 0.f;         // fp32 within a chunk run, folded into double at flush
  double acc_d = 0.0;
  for (int64_t chunk = blockIdx.x; chunk < tab.total_chunks; chunk += gridDim.x) {
    const int t = find_tensor(s_begin, tab.n, chunk, cur < 0 ? 0 : cur);
    if (t != cur) {
      if (cur >= 0 && want_sums) {
        const double tot = block_sum_double(acc_d + static_cast<double>(acc), s_red);
        if (threadIdx.x == 0) partials[static_cast<int64_t>(cur) * gridDim.x + blockIdx.x] = tot;
      }
      cur = t;
      acc = 0.f;
      acc_d = 0.0;
    }
    const KlDesc& d = tab.t[t];
    const float coeff = d.coeff * gscale;
    const int64_t base = (chunk - d.chunk_begin) * kChunk;
    if (d.vec && base + kChunk <= d.numel) {
      float4 m[kVecPerThread], r[kVecPerThread];
#pragma unroll
      for (int j = 0; j < kVecPerThread; ++j) {
        const int64_t i = base + (static_cast<int64_t>(j) * kThreads + threadIdx.x) * 4;
        m[j] = ldg_stream4(d.mu + i);
        r[j] = ldg_stream4(d.rho + i);
      }
#pragma unroll
      for (int j = 0; j < kVecPerThread; ++j) {
        KlTerm a, b, c, e;
        if (fmaxf(fmaxf(r[j].x, r[j].y), fmaxf(r[j].z, r[j].w)) <= kFastRhoMax) {   // one branch per 4 elements
          a = kl_element_fast<kGrad>(m[j].x, r[j].x, d.loc, d.inv_scale, d.log_scale, coeff);
          b = kl_element_fast<kGrad>(m[j].y, r[j].y, d.loc, d.inv_scale, d.log_scale, coeff);
          c = kl_element_fast<kGrad>(m[j].z, r[j].z, d.loc, d.inv_scale, d.log_scale, coeff);
          e = kl_element_fast<kGrad>(m[j].w, r[j].w, d.loc, d.inv_scale, d.log_scale, coeff);
        } else {
          a = kl_element_slow<kGrad>(m[j].x, r[j].x, d.loc, d.inv_scale, d.log_scale, coeff);
          b = kl_element_slow<kGrad>(m[j].y, r[j].y, d.loc, d.inv_scale, d.log_scale, coeff);
          c = kl_element_slow<kGrad>(m[j].z, r[j].z, d.loc, d.inv_scale, d.log_scale, coeff);
          e = kl_element_slow<kGrad>(m[j].w, r[j].w, d.loc, d.inv_scale, d.log_scale, coeff);
        }
        acc += (a.kl + b.kl) + (c.kl + e.kl);
        if (kGrad) {
          const int64_t i = base + (static_cast<int64_t>(j) * kThreads + threadIdx.x) * 4;
          *reinterpret_cast<float4*>(d.gmu + i) = make_float4(a.gmu, b.gmu, c.gmu, e.gmu);
          *reinterpret_cast<float4*>(d.grho + i) = make_float4(a.grho, b.grho, c.grho, e.grho);
        }
      }
    } else {   // ragged tail or unaligned tensor: scalar, bounds-checked
      for (int j = 0; j < 4 * kVecPerThread; ++j) {
        const int64_t i = base + static_cast<int64_t>(j) * kThreads + threadIdx.x;
        if (i < d.numel) {
          const KlTerm a = kl_element_slow<kGrad>(d.mu[i], d.rho[i], d.loc, d.inv_scale, d.log_scale, coeff);
          acc += a.kl;
          if (kGrad) { d.gmu[i] = a.gmu; d.grho[i] = a.grho; }
        }
      }
    }
    // keep the fp32 partial short: fold into double every chunk
    acc_d += static_cast<double>(acc);
    acc = 0.f;
  }
  if (!want_sums) return;
  if (cur >= 0) {
    const double tot = block_sum_double(acc_d, s_red);
    if (threadIdx.x == 0) partials[static_cast<int64_t>(cur) * gridDim.x + blockIdx.x] = tot;
  }
  // last block done: fixed-order sum of the per-block slots
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(done_counter, 1u);
    s_last = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double weighted = 0.0;
  for (int t = 0; t < tab.n; ++t) {
    double v = 0.0;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += kThreads)
      v += __ldcg(partials + static_cast<int64_t>(t) * gridDim.x + b);
    const double tot = block_sum_double(v, s_red);
    if (threadIdx.x == 0) {
      if (kl_sum != nullptr) kl_sum[t] = tot;
      weighted += static_cast<double>(tab.t[t].coeff) * tot;
    }
  }
  if (threadIdx.x == 0) {
    if (kl_total != nullptr)
      *kl_total = static_cast<float>((accumulate_total ? static_cast<double>(*kl_total) : 0.0) + weighted);
    *done_counter = 0u;   // ready for the next launch on this stream
  }
}

int kl_grid(bool grad, int64_t total_chunks) {
  static int occ[2] = {0, 0};
  int& o = occ[grad ? 1 : 0];
  if (o == 0) {
    int v = 0;
    cudaError_t e = grad ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kl_kernel<true>, kThreads, 0)
                         : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kl_kernel<false>, kThreads, 0);
    o = (e == cudaSuccess && v > 0) ? v : 4;
  }
  int64_t g = static_cast<int64_t>(sm_count()) * o;
  if (g > total_chunks) g = total_chunks;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

constexpr int kMaxGrid = 148 * 16;   // upper bound used for workspace sizing

}  // namespace
}  // namespace bnn

using namespace bnn;

extern "C" {

size_t bnn_kl_workspace_size(int32_t n_tensors) {
  (void)n_tensors;
  return static_cast<size_t>(kMaxTensors) * kMaxGrid * sizeof(double) + 256;
}

int bnn_kl(const bnn_kl_tensor* tensors, int32_t n_tensors, double* kl_sum, float* kl_total,
           const float* grad_scale_dev, void* workspace, size_t workspace_bytes, void* stream) {
  BNN_REQUIRE(n_tensors >= 0, BNN_ERR_BAD_ARGUMENT, "n_tensors < 0");
  if (n_tensors == 0) return BNN_OK;
  BNN_REQUIRE(tensors != nullptr, BNN_ERR_BAD_ARGUMENT, "bnn_kl: tensor table is NULL");
  BNN_REQUIRE(workspace != nullptr && workspace_bytes >= bnn_kl_workspace_size(n_tensors),
              BNN_ERR_WORKSPACE, "bnn_kl: workspace too small (%zu < %zu)", workspace_bytes,
              bnn_kl_workspace_size(n_tensors));
  BNN_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, BNN_ERR_MISALIGNED,
              "bnn_kl: workspace must be 256-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned int* counter = reinterpret_cast<unsigned int*>(workspace);
  double* partials = reinterpret_cast<double*>(static_cast<char*>(workspace) + 256);

  bool any_grad = false;
  for (int i = 0; i < n_tensors; ++i) {
    const bnn_kl_tensor& t = tensors[i];
    BNN_REQUIRE(t.numel >= 0, BNN_ERR_BAD_ARGUMENT, "bnn_kl: tensor %d has numel < 0", i);
    BNN_REQUIRE(t.numel == 0 || (t.mu && t.rho), BNN_ERR_BAD_ARGUMENT, "bnn_kl: tensor %d has NULL mu/rho", i);
    BNN_REQUIRE((t.grad_mu == nullptr) == (t.grad_rho == nullptr), BNN_ERR_BAD_ARGUMENT,
                "bnn_kl: tensor %d needs both or neither gradient pointer", i);
    BNN_REQUIRE(t.prior_scale > 0.f, BNN_ERR_BAD_ARGUMENT, "bnn_kl: tensor %d prior scale must be > 0", i);
    any_grad = any_grad || t.grad_mu != nullptr;
  }
  if (any_grad)
    for (int i = 0; i < n_tensors; ++i)
      BNN_REQUIRE(tensors[i].grad_mu != nullptr || tensors[i].numel == 0, BNN_ERR_BAD_ARGUMENT,
                  "bnn_kl: gradients requested for some tensors but not tensor %d", i);
  BNN_REQUIRE(kl_sum != nullptr || kl_total != nullptr || any_grad, BNN_ERR_BAD_ARGUMENT,
              "bnn_kl: nothing to compute");
  const bool want_sums = kl_sum != nullptr || kl_total != nullptr;
  bool total_started = false;
  int rc = check_device();          // arguments are validated first: bad calls fail the same way on any machine
  if (rc != BNN_OK) return rc;

  for (int first = 0; first < n_tensors; first += kMaxTensors) {
    const int n = (n_tensors - first < kMaxTensors) ? n_tensors - first : kMaxTensors;
    KlTable tab;
    tab.n = n;
    tab.pad = 0;
    int64_t chunks = 0;
    for (int i = 0; i < n; ++i) {
      const bnn_kl_tensor& t = tensors[first + i];
      KlDesc& d = tab.t[i];
      d.mu = t.mu; d.rho = t.rho; d.gmu = t.grad_mu; d.grho = t.grad_rho;
      d.numel = t.numel;
      d.chunk_begin = chunks;
      d.loc = t.prior_loc;
      d.inv_scale = static_cast<float>(1.0 / static_cast<double>(t.prior_scale));
      d.log_scale = static_cast<float>(log(static_cast<double>(t.prior_scale)));
      d.coeff = t.grad_coeff;
      d.vec = aligned16(t.mu) && aligned16(t.rho) && (t.grad_mu == nullptr || (aligned16(t.grad_mu) && aligned16(t.grad_rho)));
      d.pad = 0;
      chunks += (t.numel + kChunk - 1) / kChunk;
    }
    tab.total_chunks = chunks;
    double* out = kl_sum ? kl_sum + first : nullptr;
    if (chunks == 0) {
      if (out) BNN_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(double) * n, st));
      continue;
    }
    if (want_sums) BNN_CUDA_OK(cudaMemsetAsync(counter, 0, sizeof(unsigned int), st));
    const int grid = kl_grid(any_grad, chunks);
    const int acc = total_started ? 1 : 0;
    if (any_grad)
      kl_kernel<true><<<grid, kThreads, 0, st>>>(tab, out, kl_total, acc, grad_scale_dev, partials, counter);
    else
      kl_kernel<false><<<grid, kThreads, 0, st>>>(tab, out, kl_total, acc, grad_scale_dev, partials, counter);
    BNN_CUDA_OK(cudaGetLastError());
    total_started = true;
  }
  if (kl_total != nullptr && !total_started) BNN_CUDA_OK(cudaMemsetAsync(kl_total, 0, sizeof(float), st));
  return BNN_OK;
}

}  // extern "C"
