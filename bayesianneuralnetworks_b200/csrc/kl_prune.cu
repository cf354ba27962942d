// kl_prune.cu — the bandwidth-bound sweeps over the variational parameters (mu, rho):
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

// One element. Fast-math formulation (see DESIGN.md "KL arithmetic"): e = exp(rho) once;
// softplus = log1p(e) through the atanh series for e <= 1/4, log(1+e) otherwise; log(sigma) by
// MUFU.LG2; 1/scale and log(scale) come from the host.
template <bool kGrad>
__device__ __forceinline__ KlTerm kl_element(float mu, float rho, float loc, float inv_scale,
                                             float log_scale, float coeff) {
  float sp, e;
  if (rho > 20.0f) {
    sp = rho;
    e = 0.f;   // unused; sigmoid -> 1 below
  } else {
    e = __expf(rho);
    if (e <= 0.25f) {
      const float z = __fdividef(e, 2.0f + e);
      const float z2 = z * z;
      // log1p(e) = 2*atanh(z) = 2z(1 + z^2/3 + z^4/5 + z^6/7 + z^8/9)
      float p = fmaf(z2, 0.1111111111f, 0.1428571429f);
      p = fmaf(z2, p, 0.2f);
      p = fmaf(z2, p, 0.3333333333f);
      p = fmaf(z2, p, 1.0f);
      sp = 2.0f * z * p;
    } else {
      sp = __logf(1.0f + e);
    }
  }
  const float sigma = 1e-10f + sp;
  const float r = sigma * inv_scale;
  const float d = (mu - loc) * inv_scale;
  KlTerm out;
  out.kl = 0.5f * (fmaf(r, r, fmaf(d, d, -1.0f))) - (__logf(sigma) - log_scale);
  if (kGrad) {
    const float sig = rho > 20.0f ? 1.0f : __fdividef(e, 1.0f + e);
    out.gmu = coeff * d * inv_scale;
    out.grho = coeff * (r * inv_scale - __fdividef(1.0f, sigma)) * sig;
  } else {
    out.gmu = 0.f;
    out.grho = 0.f;
  }
  return out;
}

__device__ __forceinline__ int find_tensor(const int64_t* chunk_begin, int n, int64_t chunk) {
  int t = 0;
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
    const int t = find_tensor(s_begin, tab.n, chunk);
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
        const KlTerm a = kl_element<kGrad>(m[j].x, r[j].x, d.loc, d.inv_scale, d.log_scale, coeff);
        const KlTerm b = kl_element<kGrad>(m[j].y, r[j].y, d.loc, d.inv_scale, d.log_scale, coeff);
        const KlTerm c = kl_element<kGrad>(m[j].z, r[j].z, d.loc, d.inv_scale, d.log_scale, coeff);
        const KlTerm e = kl_element<kGrad>(m[j].w, r[j].w, d.loc, d.inv_scale, d.log_scale, coeff);
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
          const KlTerm a = kl_element<kGrad>(d.mu[i], d.rho[i], d.loc, d.inv_scale, d.log_scale, coeff);
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

// ================================================================================== prune
// ordered key: unsigned order == float order (larger float -> larger uint)
__device__ __forceinline__ uint32_t order_key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unorder_key(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// key = Normal(mu, sigma).log_prob(0) with torch's op order and NO fused multiply-add:
//   var = sigma*sigma; t = (0 - mu); t = t*t; t = -t; t = t / (2*var); t = t - log(sigma);
//   t = t - float(log(sqrt(2*pi)))
__device__ __forceinline__ float prune_key(float mu, float rho) {
  const float sigma = stddev_exact(rho);
  const float var = __fmul_rn(sigma, sigma);
  const float nmu = __fsub_rn(0.0f, mu);
  const float sq = __fmul_rn(nmu, nmu);
  const float q = __fdiv_rn(-sq, __fmul_rn(2.0f, var));
  const float a = __fsub_rn(q, logf(sigma));
  return __fsub_rn(a, 0.9189385332046727f);
}

struct PruneState {          // one per tensor, in the workspace
  uint32_t prefix;           // bits of the k-th largest ordered key decided so far
  uint32_t need_ranks;       // ties at the threshold must be ranked by index
  int64_t k_rem;             // how many still to take inside the current prefix class
  int64_t eq_total;          // elements equal to the final threshold
};

struct PruneDesc {
  float* mu;
  float* rho;
  uint8_t* mask;
  float* keys_out;
  uint32_t* keys;            // workspace: ordered keys
  uint32_t* hist;            // workspace: 2048 bins
  int64_t* chunk_cnt;        // workspace: per-chunk count of keys equal to the threshold (then offsets)
  PruneState* state;
  int64_t numel;
  int64_t k;
  int64_t chunk_begin;
  int64_t n_chunks;
};
struct PruneTable {
  PruneDesc t[kMaxTensors];
  int n;
  int pad;
  int64_t total_chunks;
};

constexpr int kBins = 2048;

__device__ __forceinline__ void flush_hist(uint32_t* s_hist, uint32_t* g_hist) {
  __syncthreads();
  for (int b = threadIdx.x; b < kBins; b += kThreads) {
    const uint32_t c = s_hist[b];
    if (c) atomicAdd(g_hist + b, c);
    s_hist[b] = 0;
  }
  __syncthreads();
}

// pass 0: keys from (mu, rho) -> workspace, histogram of the top 11 bits
// pass 1/2: histogram of the next digit among keys matching the prefix decided so far
template <int kPass>
__global__ void __launch_bounds__(kThreads)
prune_hist_kernel(const __grid_constant__ PruneTable tab) {
  __shared__ uint32_t s_hist[kBins];
  __shared__ int64_t s_begin[kMaxTensors];
  if (threadIdx.x < tab.n) s_begin[threadIdx.x] = tab.t[threadIdx.x].chunk_begin;
  for (int b = threadIdx.x; b < kBins; b += kThreads) s_hist[b] = 0;
  __syncthreads();
  constexpr int shift = kPass == 0 ? 21 : (kPass == 1 ? 10 : 0);
  constexpr uint32_t digit_mask = kPass == 2 ? 0x3ffu : 0x7ffu;
  constexpr uint32_t prefix_mask = kPass == 0 ? 0u : (kPass == 1 ? 0xffe00000u : 0xfffffc00u);
  int cur = -1;
  uint32_t prefix = 0;
  bool active = true;
  for (int64_t chunk = blockIdx.x; chunk < tab.total_chunks; chunk += gridDim.x) {
    const int t = find_tensor(s_begin, tab.n, chunk);
    if (t != cur) {
      if (cur >= 0 && active) flush_hist(s_hist, tab.t[cur].hist);
      cur = t;
      const PruneDesc& d0 = tab.t[t];
      active = d0.k > 0 && d0.k < d0.numel;      // k == 0 / k == numel need no selection
      if (kPass > 0) prefix = d0.state->prefix;
    }
    const PruneDesc& d = tab.t[t];
    if (kPass > 0 && !active) continue;
    const int64_t base = (chunk - d.chunk_begin) * kChunk;
#pragma unroll 4
    for (int j = 0; j < 4 * kVecPerThread; ++j) {
      const int64_t i = base + static_cast<int64_t>(j) * kThreads + threadIdx.x;
      if (i < d.numel) {
        uint32_t ok;
        if (kPass == 0) {
          const float key = prune_key(d.mu[i], d.rho[i]);
          ok = order_key(key);
          d.keys[i] = ok;
          if (d.keys_out != nullptr) d.keys_out[i] = key;
        } else {
          ok = d.keys[i];
        }
        if (active && ((ok ^ prefix) & prefix_mask) == 0u)
          atomicAdd(&s_hist[(ok >> shift) & digit_mask], 1u);
      }
    }
  }
  if (cur >= 0 && active) flush_hist(s_hist, tab.t[cur].hist);
}

// one block per tensor: walk the histogram from the top bin down until k_rem is covered
template <int kPass>
__global__ void __launch_bounds__(kThreads) prune_select_kernel(const __grid_constant__ PruneTable tab) {
  const PruneDesc& d = tab.t[blockIdx.x];
  __shared__ uint32_t s_hist[kBins];
  constexpr int shift = kPass == 0 ? 21 : (kPass == 1 ? 10 : 0);
  constexpr int bins = kPass == 2 ? 1024 : 2048;
  for (int b = threadIdx.x; b < kBins; b += kThreads) {
    s_hist[b] = d.hist[b];
    d.hist[b] = 0;                       // clean for the next pass / next call
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  PruneState st = *d.state;
  if (kPass == 0) {
    st.prefix = 0;
    st.need_ranks = 0;
    st.k_rem = d.k;
    st.eq_total = 0;
  }
  if (d.k > 0 && d.k < d.numel) {
    int64_t above = 0;
    int b = bins - 1;
    for (; b > 0; --b) {
      if (above + s_hist[b] >= st.k_rem) break;
      above += s_hist[b];
    }
    st.prefix |= static_cast<uint32_t>(b) << shift;
    st.k_rem -= above;
    if (kPass == 2) {
      st.eq_total = s_hist[b];
      st.need_ranks = (st.k_rem < st.eq_total) ? 1u : 0u;
    }
  }
  *d.state = st;
}

// per-chunk count of keys equal to the threshold (only for tensors with ties at the boundary)
__global__ void __launch_bounds__(kThreads) prune_count_eq_kernel(const __grid_constant__ PruneTable tab) {
  __shared__ int64_t s_begin[kMaxTensors];
  __shared__ int s_cnt[kThreads / 32];
  if (threadIdx.x < tab.n) s_begin[threadIdx.x] = tab.t[threadIdx.x].chunk_begin;
  __syncthreads();
  for (int64_t chunk = blockIdx.x; chunk < tab.total_chunks; chunk += gridDim.x) {
    const int t = find_tensor(s_begin, tab.n, chunk);
    const PruneDesc& d = tab.t[t];
    const PruneState st = *d.state;
    if (!st.need_ranks) continue;
    const int64_t local = chunk - d.chunk_begin;
    const int64_t base = local * kChunk;
    int c = 0;
    for (int j = 0; j < 4 * kVecPerThread; ++j) {
      const int64_t i = base + static_cast<int64_t>(j) * kThreads + threadIdx.x;
      if (i < d.numel && d.keys[i] == st.prefix) ++c;
    }
    c = warp_sum(c);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < kThreads / 32; ++w) tot += s_cnt[w];
      d.chunk_cnt[local] = tot;
    }
  }
}

// exclusive scan of the per-chunk counts, one block per tensor (serial over 256-wide strips)
__global__ void __launch_bounds__(kThreads) prune_scan_kernel(const __grid_constant__ PruneTable tab) {
  const PruneDesc& d = tab.t[blockIdx.x];
  if (!d.state->need_ranks) return;
  __shared__ int64_t s_val[kThreads];
  __shared__ int64_t s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < d.n_chunks; base += kThreads) {
    const int64_t i = base + threadIdx.x;
    const int64_t v = i < d.n_chunks ? d.chunk_cnt[i] : 0;
    s_val[threadIdx.x] = v;
    __syncthreads();
    // Hillis-Steele inclusive scan over 256 entries
    for (int off = 1; off < kThreads; off <<= 1) {
      const int64_t add = threadIdx.x >= off ? s_val[threadIdx.x - off] : 0;
      __syncthreads();
      s_val[threadIdx.x] += add;
      __syncthreads();
    }
    const int64_t incl = s_val[threadIdx.x];
    const int64_t carry = s_carry;
    if (i < d.n_chunks) d.chunk_cnt[i] = carry + incl - v;   // exclusive offset
    __syncthreads();
    if (threadIdx.x == kThreads - 1) s_carry = carry + incl;
    __syncthreads();
  }
}

// apply: key > T pruned; key == T pruned while its index rank among equals is < k_rem
__global__ void __launch_bounds__(kThreads) prune_apply_kernel(const __grid_constant__ PruneTable tab) {
  __shared__ int64_t s_begin[kMaxTensors];
  __shared__ int s_warp[kThreads / 32];
  if (threadIdx.x < tab.n) s_begin[threadIdx.x] = tab.t[threadIdx.x].chunk_begin;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t chunk = blockIdx.x; chunk < tab.total_chunks; chunk += gridDim.x) {
    const int t = find_tensor(s_begin, tab.n, chunk);
    const PruneDesc& d = tab.t[t];
    const PruneState st = *d.state;
    const bool none = d.k <= 0, all = d.k >= d.numel;
    const int64_t local = chunk - d.chunk_begin;
    const int64_t base = local * kChunk;
    int64_t rank_base = st.need_ranks ? d.chunk_cnt[local] : 0;
    for (int j = 0; j < 4 * kVecPerThread; ++j) {      // index order: j major, then thread
      const int64_t i = base + static_cast<int64_t>(j) * kThreads + threadIdx.x;
      const bool in = i < d.numel;
      const uint32_t ok = in ? d.keys[i] : 0u;
      bool take = in && (all || (!none && ok > st.prefix));
      const bool eq = in && !all && !none && ok == st.prefix;
      if (st.need_ranks) {     // uniform per tensor
        const unsigned int bal = __ballot_sync(0xffffffffu, eq);
        const int before = __popc(bal & ((1u << lane) - 1u));
        __syncthreads();
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        int wbase = 0, tot = 0;
        for (int w = 0; w < kThreads / 32; ++w) {
          const int c = s_warp[w];
          if (w < warp) wbase += c;
          tot += c;
        }
        if (eq && rank_base + wbase + before < st.k_rem) take = true;
        rank_base += tot;
      } else if (eq) {
        take = true;           // every key equal to the threshold is inside the top k
      }
      if (in) {
        if (take) { d.mu[i] = 0.0f; d.rho[i] = -30.0f; }
        if (d.mask != nullptr) d.mask[i] = take ? 1 : 0;
      }
    }
  }
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

size_t prune_ws_one(int64_t numel) {
  const int64_t chunks = (numel + kChunk - 1) / kChunk;
  return align_up(static_cast<size_t>(numel) * 4, 256) + align_up(kBins * 4, 256) +
         align_up(static_cast<size_t>(chunks) * 8, 256) + align_up(sizeof(PruneState), 256);
}

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
  int rc = check_device();
  if (rc != BNN_OK) return rc;
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

size_t bnn_prune_workspace_size(const bnn_prune_tensor* tensors, int32_t n_tensors) {
  size_t total = 256;
  if (tensors == nullptr) return total;
  for (int i = 0; i < n_tensors; ++i) total += prune_ws_one(tensors[i].numel > 0 ? tensors[i].numel : 0);
  return total;
}

int bnn_prune(const bnn_prune_tensor* tensors, int32_t n_tensors, void* workspace,
              size_t workspace_bytes, void* stream) {
  BNN_REQUIRE(n_tensors >= 0, BNN_ERR_BAD_ARGUMENT, "n_tensors < 0");
  if (n_tensors == 0) return BNN_OK;
  BNN_REQUIRE(tensors != nullptr, BNN_ERR_BAD_ARGUMENT, "bnn_prune: tensor table is NULL");
  BNN_REQUIRE(workspace != nullptr && workspace_bytes >= bnn_prune_workspace_size(tensors, n_tensors),
              BNN_ERR_WORKSPACE, "bnn_prune: workspace too small (%zu < %zu)", workspace_bytes,
              bnn_prune_workspace_size(tensors, n_tensors));
  BNN_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, BNN_ERR_MISALIGNED,
              "bnn_prune: workspace must be 256-byte aligned");
  for (int i = 0; i < n_tensors; ++i) {
    BNN_REQUIRE(tensors[i].numel >= 0 && tensors[i].k >= 0 && tensors[i].k <= tensors[i].numel,
                BNN_ERR_BAD_ARGUMENT, "bnn_prune: tensor %d needs 0 <= k <= numel", i);
    BNN_REQUIRE(tensors[i].numel == 0 || (tensors[i].mu && tensors[i].rho), BNN_ERR_BAD_ARGUMENT,
                "bnn_prune: tensor %d has NULL mu/rho", i);
  }
  int rc = check_device();
  if (rc != BNN_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* ws = static_cast<char*>(workspace) + 256;
  const int max_grid = sm_count() * 8;

  for (int first = 0; first < n_tensors; first += kMaxTensors) {
    const int n = (n_tensors - first < kMaxTensors) ? n_tensors - first : kMaxTensors;
    PruneTable tab;
    tab.n = 0;
    tab.pad = 0;
    int64_t chunks = 0;
    char* hist_first = nullptr;
    for (int i = 0; i < n; ++i) {
      const bnn_prune_tensor& t = tensors[first + i];
      if (t.numel == 0) continue;
      PruneDesc& d = tab.t[tab.n++];
      const int64_t nch = (t.numel + kChunk - 1) / kChunk;
      d.mu = t.mu; d.rho = t.rho; d.mask = t.mask_out; d.keys_out = t.keys_out;
      d.numel = t.numel; d.k = t.k; d.chunk_begin = chunks; d.n_chunks = nch;
      d.keys = reinterpret_cast<uint32_t*>(ws); ws += align_up(static_cast<size_t>(t.numel) * 4, 256);
      d.hist = reinterpret_cast<uint32_t*>(ws);
      if (!hist_first) hist_first = ws;
      ws += align_up(kBins * 4, 256);
      d.chunk_cnt = reinterpret_cast<int64_t*>(ws); ws += align_up(static_cast<size_t>(nch) * 8, 256);
      d.state = reinterpret_cast<PruneState*>(ws); ws += align_up(sizeof(PruneState), 256);
      BNN_CUDA_OK(cudaMemsetAsync(d.hist, 0, kBins * 4, st));
      chunks += nch;
    }
    if (tab.n == 0) continue;
    tab.total_chunks = chunks;
    const int grid = static_cast<int>(chunks < max_grid ? chunks : max_grid);
    prune_hist_kernel<0><<<grid, kThreads, 0, st>>>(tab);
    prune_select_kernel<0><<<tab.n, kThreads, 0, st>>>(tab);
    prune_hist_kernel<1><<<grid, kThreads, 0, st>>>(tab);
    prune_select_kernel<1><<<tab.n, kThreads, 0, st>>>(tab);
    prune_hist_kernel<2><<<grid, kThreads, 0, st>>>(tab);
    prune_select_kernel<2><<<tab.n, kThreads, 0, st>>>(tab);
    prune_count_eq_kernel<<<grid, kThreads, 0, st>>>(tab);
    prune_scan_kernel<<<tab.n, kThreads, 0, st>>>(tab);
    prune_apply_kernel<<<grid, kThreads, 0, st>>>(tab);
    BNN_CUDA_OK(cudaGetLastError());
  }
  return BNN_OK;
}

}  // extern "C"
