// adam.cu — bnn_adam_kl_step: the optimizer half of the ELBO tail.
//
// One pass over the variational parameters that (a) forms the closed-form KL gradient of every element
//     dKL/dmu = c (mu - loc) / scale^2,   dKL/drho = c (sigma / scale^2 - 1 / sigma) sigmoid(rho),   sigma = 1e-10 + softplus(rho)
// (the gradient of KLDivergence.forward, pytorch_bayesian/nn/loss.py:16-38, with c = 1 / (numel_t n_tensors
// n_batches)), (b) adds it to the likelihood gradient that autograd (and the gradient all-reduce) left in g_mu / g_rho, and
// (c) applies torch.optim.Adam's update (examples/MNIST/train.py:43,65: Adam over all parameters, no weight decay, no
// amsgrad) to mu, rho and their moment buffers.  The separate path reads and writes the gradients three more times (KL
// gradient kernel, autograd accumulation, optimizer).  Bandwidth bound: 8 loads + 6 stores of 4 bytes per (mu, rho) pair.
#include "common.cuh"

namespace bnn {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxTensors = 16;       // per launch: the table travels by value in the kernel parameters

struct AdamDesc {
  float* mu; float* rho;
  const float* g_mu; const float* g_rho;
  float* m_mu; float* v_mu; float* m_rho; float* v_rho;
  int64_t numel;
  int64_t block_begin;       // first global block-chunk of this tensor
  float loc, inv_scale2, coeff;
  int vec;                   // every pointer 16-byte aligned
};
struct AdamTable {
  AdamDesc t[kMaxTensors];
  int n;
  int pad;
  int64_t total_chunks;
  float lr, beta1, beta2, eps;
  const float* step_dev;     // device scalar: number of this step (1, 2, ...), as torch's capturable Adam keeps it
  float step_host;           // used when step_dev is NULL
};

constexpr int kPerThread = 4;                               // elements per thread per chunk (one float4 per array)
constexpr int kChunk = kThreads * kPerThread;               // 1024 elements

struct Hyper { float lr_over_bc1, inv_sqrt_bc2, beta1, beta2, eps; };

__device__ __forceinline__ void adam_element(float& p, float g, float& m, float& v, const Hyper& h) {
  m = fmaf(g - m, 1.0f - h.beta1, m);                      // exp_avg.lerp_(grad, 1 - beta1)
  v = fmaf(h.beta2, v, (1.0f - h.beta2) * g * g);          // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  const float denom = fmaf(sqrtf(v), h.inv_sqrt_bc2, h.eps);
  p = fmaf(-h.lr_over_bc1, m / denom, p);
}

__device__ __forceinline__ void elbo_adam_pair(float& mu, float& rho, float g_mu, float g_rho, float& m_mu, float& v_mu,
                                               float& m_rho, float& v_rho, float loc, float inv_scale2, float coeff,
                                               const Hyper& h) {
  if (coeff != 0.0f) {
    const float sigma = stddev_exact(rho);
    const float sig = 1.0f / (1.0f + expf(-rho));
    g_mu = fmaf(coeff * (mu - loc), inv_scale2, g_mu);
    g_rho = fmaf(coeff * (sigma * inv_scale2 - 1.0f / sigma), sig, g_rho);
  }
  adam_element(mu, g_mu, m_mu, v_mu, h);
  adam_element(rho, g_rho, m_rho, v_rho, h);
}

__global__ void __launch_bounds__(kThreads) adam_kl_kernel(const __grid_constant__ AdamTable tab) {
  __shared__ int64_t s_begin[kMaxTensors];
  if (threadIdx.x < tab.n) s_begin[threadIdx.x] = tab.t[threadIdx.x].block_begin;
  __syncthreads();
  const double step = tab.step_dev != nullptr ? static_cast<double>(*tab.step_dev) : static_cast<double>(tab.step_host);
  Hyper h;
  h.beta1 = tab.beta1; h.beta2 = tab.beta2; h.eps = tab.eps;
  h.lr_over_bc1 = static_cast<float>(static_cast<double>(tab.lr) / (1.0 - pow(static_cast<double>(tab.beta1), step)));
  h.inv_sqrt_bc2 = static_cast<float>(1.0 / sqrt(1.0 - pow(static_cast<double>(tab.beta2), step)));
  int t = 0;
  for (int64_t chunk = blockIdx.x; chunk < tab.total_chunks; chunk += gridDim.x) {
    while (t + 1 < tab.n && chunk >= s_begin[t + 1]) ++t;
    const AdamDesc& d = tab.t[t];
    const int64_t i0 = (chunk - d.block_begin) * kChunk + threadIdx.x * kPerThread;
    if (i0 >= d.numel) continue;
    if (d.vec && i0 + kPerThread <= d.numel) {
      float4 mu = *reinterpret_cast<const float4*>(d.mu + i0), rho = *reinterpret_cast<const float4*>(d.rho + i0);
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 gm = d.g_mu ? ldg_stream4(d.g_mu + i0) : z, gr = d.g_rho ? ldg_stream4(d.g_rho + i0) : z;
      float4 mm = *reinterpret_cast<const float4*>(d.m_mu + i0), vm = *reinterpret_cast<const float4*>(d.v_mu + i0);
      float4 mr = *reinterpret_cast<const float4*>(d.m_rho + i0), vr = *reinterpret_cast<const float4*>(d.v_rho + i0);
      elbo_adam_pair(mu.x, rho.x, gm.x, gr.x, mm.x, vm.x, mr.x, vr.x, d.loc, d.inv_scale2, d.coeff, h);
      elbo_adam_pair(mu.y, rho.y, gm.y, gr.y, mm.y, vm.y, mr.y, vr.y, d.loc, d.inv_scale2, d.coeff, h);
      elbo_adam_pair(mu.z, rho.z, gm.z, gr.z, mm.z, vm.z, mr.z, vr.z, d.loc, d.inv_scale2, d.coeff, h);
      elbo_adam_pair(mu.w, rho.w, gm.w, gr.w, mm.w, vm.w, mr.w, vr.w, d.loc, d.inv_scale2, d.coeff, h);
      *reinterpret_cast<float4*>(d.mu + i0) = mu; *reinterpret_cast<float4*>(d.rho + i0) = rho;
      *reinterpret_cast<float4*>(d.m_mu + i0) = mm; *reinterpret_cast<float4*>(d.v_mu + i0) = vm;
      *reinterpret_cast<float4*>(d.m_rho + i0) = mr; *reinterpret_cast<float4*>(d.v_rho + i0) = vr;
    } else {
      for (int e = 0; e < kPerThread; ++e) {
        const int64_t i = i0 + e;
        if (i >= d.numel) break;
        float mu = d.mu[i], rho = d.rho[i], mm = d.m_mu[i], vm = d.v_mu[i], mr = d.m_rho[i], vr = d.v_rho[i];
        elbo_adam_pair(mu, rho, d.g_mu ? d.g_mu[i] : 0.f, d.g_rho ? d.g_rho[i] : 0.f, mm, vm, mr, vr, d.loc, d.inv_scale2,
                       d.coeff, h);
        d.mu[i] = mu; d.rho[i] = rho; d.m_mu[i] = mm; d.v_mu[i] = vm; d.m_rho[i] = mr; d.v_rho[i] = vr;
      }
    }
  }
}

}  // namespace
}  // namespace bnn

using namespace bnn;

extern "C" int bnn_adam_kl_step(const bnn_adam_tensor* tensors, int32_t n_tensors, float lr, float beta1, float beta2,
                                float eps, const float* step_dev, int64_t step_host, void* stream) {
  BNN_REQUIRE(n_tensors >= 0, BNN_ERR_BAD_ARGUMENT, "bnn_adam_kl_step: n_tensors < 0");
  if (n_tensors == 0) return BNN_OK;
  BNN_REQUIRE(tensors != nullptr, BNN_ERR_BAD_ARGUMENT, "bnn_adam_kl_step: tensor table is NULL");
  BNN_REQUIRE(lr >= 0.f && beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f, BNN_ERR_BAD_ARGUMENT,
              "bnn_adam_kl_step: need lr >= 0, 0 <= beta < 1, eps >= 0");
  BNN_REQUIRE(step_dev != nullptr || step_host >= 1, BNN_ERR_BAD_ARGUMENT,
              "bnn_adam_kl_step: the step number starts at 1 (pass step_dev or step_host >= 1)");
  for (int i = 0; i < n_tensors; ++i) {
    const bnn_adam_tensor& t = tensors[i];
    BNN_REQUIRE(t.numel >= 0, BNN_ERR_BAD_ARGUMENT, "bnn_adam_kl_step: tensor %d has numel < 0", i);
    BNN_REQUIRE(t.numel == 0 || (t.mu && t.rho && t.m_mu && t.v_mu && t.m_rho && t.v_rho), BNN_ERR_BAD_ARGUMENT,
                "bnn_adam_kl_step: tensor %d has a NULL parameter or moment pointer", i);
    BNN_REQUIRE(t.kl_coeff == 0.f || t.prior_scale > 0.f, BNN_ERR_BAD_ARGUMENT,
                "bnn_adam_kl_step: tensor %d needs prior scale > 0", i);
  }
  int rc = check_device();
  if (rc != BNN_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int max_grid = sm_count() * 8;
  for (int first = 0; first < n_tensors; first += kMaxTensors) {
    const int n = n_tensors - first < kMaxTensors ? n_tensors - first : kMaxTensors;
    AdamTable tab;
    tab.n = 0; tab.pad = 0;
    tab.lr = lr; tab.beta1 = beta1; tab.beta2 = beta2; tab.eps = eps;
    tab.step_dev = step_dev; tab.step_host = static_cast<float>(step_host);
    int64_t chunks = 0;
    for (int i = 0; i < n; ++i) {
      const bnn_adam_tensor& t = tensors[first + i];
      if (t.numel == 0) continue;
      AdamDesc& d = tab.t[tab.n++];
      d.mu = t.mu; d.rho = t.rho; d.g_mu = t.g_mu; d.g_rho = t.g_rho;
      d.m_mu = t.m_mu; d.v_mu = t.v_mu; d.m_rho = t.m_rho; d.v_rho = t.v_rho;
      d.numel = t.numel; d.block_begin = chunks;
      d.loc = t.prior_loc; d.coeff = t.kl_coeff;
      d.inv_scale2 = t.kl_coeff != 0.f ? 1.0f / (t.prior_scale * t.prior_scale) : 0.f;
      d.vec = aligned16(t.mu) && aligned16(t.rho) && aligned16(t.m_mu) && aligned16(t.v_mu) && aligned16(t.m_rho) &&
              aligned16(t.v_rho) && (t.g_mu == nullptr || aligned16(t.g_mu)) && (t.g_rho == nullptr || aligned16(t.g_rho));
      chunks += (t.numel + kChunk - 1) / kChunk;
    }
    if (tab.n == 0) continue;
    tab.total_chunks = chunks;
    const int grid = static_cast<int>(chunks < max_grid ? chunks : max_grid);
    adam_kl_kernel<<<grid, kThreads, 0, st>>>(tab);
    BNN_CUDA_OK(cudaGetLastError());
  }
  return BNN_OK;
}
