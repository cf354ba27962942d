// ce.cu — bnn_mc_cross_entropy_fwd / _bwd: the likelihood half of the ELBO tail.
//
// The reference's loop body (examples/MNIST/train.py:59-61) is
//     loss = torch.stack([criterion(pred, y) for pred in preds]).mean()        criterion = CrossEntropyLoss()
// over the S Monte-Carlo predictions.  With the S predictions stored as the row blocks of ONE [S*B, C] matrix this is
// the mean over all R = S*B rows of  logsumexp(x[r, :]) - x[r, y[r % B]]  (torch.nn.functional.cross_entropy, reduction
// 'mean', ignore_index semantics: ignored rows contribute nothing and do not count).  torch evaluates it as log_softmax
// (read x, write R*C), nll_loss (read), and the same again backwards: 6 passes over R*C floats and 5 launches; here the
// forward reads x once (online max / sum, one warp per row) and the backward reads x once and writes dx once.
// Bandwidth bound at C4 (32768 x 4096 rows: 0.5 + 1.1 GB), launch bound at C2 (2048 x 10).
#include "common.cuh"

namespace bnn {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kGridPerSm = 8;

struct CeWorkspace {          // zero when handed to the first call; every call leaves `ticket` at zero again
  unsigned int ticket;
  unsigned int pad[3];
  double partial[1];          // [2 * grid]: loss sums, then valid-row counts
};

__device__ __forceinline__ void online_add(float& m, float& s, float v) {
  if (v > m) {
    s = s * exp_fast(m - v) + 1.0f;          // m = -inf on the first element: s = 0 * 0 + 1
    m = v;
  } else {
    s += exp_fast(v - m);
  }
}

// (max, sum of exp(x - max)) of one row, one warp per row
template <bool kVec>
__device__ __forceinline__ void row_max_sum(const float* __restrict__ row, int C, int lane, float& m_out, float& s_out) {
  float m = -INFINITY, s = 0.f;
  if (kVec) {
    for (int j = lane * 4; j < C; j += 128) {
      const float4 v = ldg_stream4(row + j);
      const float mx = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
      if (mx > m) { s *= exp_fast(m - mx); m = mx; }          // exp(-inf) = 0 on the first vector
      s += exp_fast(v.x - m) + exp_fast(v.y - m) + exp_fast(v.z - m) + exp_fast(v.w - m);
    }
  } else {
    for (int j = lane; j < C; j += 32) online_add(m, s, __ldg(row + j));
  }
  // combine the 32 lanes (lanes without elements hold m = -inf, s = 0)
  float m_all = m;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m_all = fmaxf(m_all, __shfl_xor_sync(0xffffffffu, m_all, o));
  float t = (m == -INFINITY) ? 0.f : s * exp_fast(m - m_all);
  t = warp_sum(t);
  m_out = m_all;
  s_out = t;
}

template <bool kVec>
__global__ void __launch_bounds__(kThreads)
ce_fwd_kernel(const float* __restrict__ x, int64_t ldx, const int64_t* __restrict__ target, int64_t rows, int64_t labels,
              int C, int64_t ignore_index, float* __restrict__ lse_out, float* __restrict__ loss_out,
              float* __restrict__ count_out, CeWorkspace* ws) {
  __shared__ double s_sum[kWarps], s_cnt[kWarps];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double sum = 0.0, cnt = 0.0;                       // lane 0 of each warp accumulates its rows
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * kWarps + warp; r < rows; r += static_cast<int64_t>(gridDim.x) * kWarps) {
    const float* row = x + r * ldx;
    float m, s;
    row_max_sum<kVec>(row, C, lane, m, s);
    if (lane == 0) {
      const float lse = m + logf(s);
      lse_out[r] = lse;
      const int64_t y = target[r % labels];
      if (y != ignore_index) {
        const bool ok = y >= 0 && y < C;            // torch asserts on the device; here the loss turns NaN
        sum += ok ? static_cast<double>(lse - __ldg(row + y)) : static_cast<double>(NAN);
        cnt += 1.0;
      }
    }
  }
  if (lane == 0) { s_sum[warp] = sum; s_cnt[warp] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, c = 0.0;
    for (int w = 0; w < kWarps; ++w) { a += s_sum[w]; c += s_cnt[w]; }        // fixed order
    ws->partial[blockIdx.x] = a;
    ws->partial[gridDim.x + blockIdx.x] = c;
    __threadfence();
    s_last = atomicAdd(&ws->ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last) {                                       // the whole last block folds the per-block partials, fixed order
    __threadfence();
    double a = 0.0, c = 0.0;
    const volatile double* p = ws->partial;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += kThreads) { a += p[b]; c += p[gridDim.x + b]; }
    a = warp_sum(a);
    c = warp_sum(c);
    __syncthreads();                                  // s_sum / s_cnt were read by thread 0 above
    if (lane == 0) { s_sum[warp] = a; s_cnt[warp] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
      a = 0.0; c = 0.0;
      for (int w = 0; w < kWarps; ++w) { a += s_sum[w]; c += s_cnt[w]; }
      *loss_out = static_cast<float>(a / c);           // 0 / 0 = NaN when every row is ignored, as torch returns
      *count_out = static_cast<float>(c);
      ws->ticket = 0u;                                  // ready for the next call on this stream
    }
  }
}

template <bool kVec>
__global__ void __launch_bounds__(kThreads)
ce_bwd_kernel(const float* __restrict__ x, int64_t ldx, const int64_t* __restrict__ target, int64_t rows, int64_t labels,
              int C, int64_t ignore_index, const float* __restrict__ lse, const float* __restrict__ count,
              const float* __restrict__ grad_loss, float* __restrict__ dx, int64_t lddx) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float scale = __ldg(grad_loss) / __ldg(count);
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * kWarps + warp; r < rows; r += static_cast<int64_t>(gridDim.x) * kWarps) {
    const float* row = x + r * ldx;
    float* out = dx + r * lddx;
    const int64_t y64 = target[r % labels];
    const float g = y64 == ignore_index ? 0.f : scale;
    const int y = (y64 >= 0 && y64 < C) ? static_cast<int>(y64) : -1;
    const float l = __ldg(lse + r);
    if (kVec) {
      for (int j = lane * 4; j < C; j += 128) {
        const float4 v = ldg_stream4(row + j);
        float4 d;
        d.x = g * (exp_fast(v.x - l) - (j == y ? 1.f : 0.f));
        d.y = g * (exp_fast(v.y - l) - (j + 1 == y ? 1.f : 0.f));
        d.z = g * (exp_fast(v.z - l) - (j + 2 == y ? 1.f : 0.f));
        d.w = g * (exp_fast(v.w - l) - (j + 3 == y ? 1.f : 0.f));
        *reinterpret_cast<float4*>(out + j) = d;
      }
    } else {
      for (int j = lane; j < C; j += 32) out[j] = g * (exp_fast(__ldg(row + j) - l) - (j == y ? 1.f : 0.f));
    }
  }
}

int ce_grid(int64_t rows) {
  const int64_t want = (rows + kWarps - 1) / kWarps;
  const int64_t cap = static_cast<int64_t>(sm_count()) * kGridPerSm;
  return static_cast<int>(want < cap ? (want < 1 ? 1 : want) : cap);
}

int check_ce(const float* x, int64_t ldx, const int64_t* target, int64_t rows, int64_t labels, int32_t classes) {
  BNN_REQUIRE(rows >= 1 && labels >= 1 && classes >= 1, BNN_ERR_BAD_ARGUMENT, "bnn_mc_cross_entropy: empty problem");
  BNN_REQUIRE(rows % labels == 0, BNN_ERR_BAD_ARGUMENT,
              "bnn_mc_cross_entropy: %lld rows are not a whole number of sample blocks of %lld labels",
              static_cast<long long>(rows), static_cast<long long>(labels));
  BNN_REQUIRE(x && target, BNN_ERR_BAD_ARGUMENT, "bnn_mc_cross_entropy: NULL pointer");
  BNN_REQUIRE(ldx >= classes, BNN_ERR_BAD_ARGUMENT, "bnn_mc_cross_entropy: row pitch < classes");
  return BNN_OK;
}

}  // namespace
}  // namespace bnn

using namespace bnn;

extern "C" {

size_t bnn_mc_cross_entropy_workspace_size(void) {
  const int sms = sm_count();
  return sizeof(CeWorkspace) + 2 * sizeof(double) * static_cast<size_t>(sms > 0 ? sms : 148) * kGridPerSm;
}

int bnn_mc_cross_entropy_fwd(const float* x, int64_t ldx, const int64_t* target, int64_t rows, int64_t labels,
                             int32_t classes, int64_t ignore_index, float* lse, float* loss, float* count,
                             void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_ce(x, ldx, target, rows, labels, classes);
  if (rc != BNN_OK) return rc;
  BNN_REQUIRE(lse && loss && count, BNN_ERR_BAD_ARGUMENT, "bnn_mc_cross_entropy_fwd: NULL output");
  rc = check_device();
  if (rc != BNN_OK) return rc;
  BNN_REQUIRE(workspace != nullptr && workspace_bytes >= bnn_mc_cross_entropy_workspace_size(), BNN_ERR_WORKSPACE,
              "bnn_mc_cross_entropy_fwd: workspace too small (need %zu bytes)", bnn_mc_cross_entropy_workspace_size());
  BNN_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15u) == 0, BNN_ERR_MISALIGNED,
              "bnn_mc_cross_entropy_fwd: workspace must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = ce_grid(rows);
  CeWorkspace* ws = static_cast<CeWorkspace*>(workspace);
  if (classes % 4 == 0 && ldx % 4 == 0 && aligned16(x))
    ce_fwd_kernel<true><<<grid, kThreads, 0, st>>>(x, ldx, target, rows, labels, classes, ignore_index, lse, loss, count, ws);
  else
    ce_fwd_kernel<false><<<grid, kThreads, 0, st>>>(x, ldx, target, rows, labels, classes, ignore_index, lse, loss, count, ws);
  BNN_CUDA_OK(cudaGetLastError());
  return BNN_OK;
}

int bnn_mc_cross_entropy_bwd(const float* x, int64_t ldx, const int64_t* target, int64_t rows, int64_t labels,
                             int32_t classes, int64_t ignore_index, const float* lse, const float* count,
                             const float* grad_loss, float* dx, int64_t lddx, void* stream) {
  int rc = check_ce(x, ldx, target, rows, labels, classes);
  if (rc != BNN_OK) return rc;
  BNN_REQUIRE(lse && count && grad_loss && dx, BNN_ERR_BAD_ARGUMENT, "bnn_mc_cross_entropy_bwd: NULL pointer");
  BNN_REQUIRE(lddx >= classes, BNN_ERR_BAD_ARGUMENT, "bnn_mc_cross_entropy_bwd: row pitch < classes");
  rc = check_device();
  if (rc != BNN_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = ce_grid(rows);
  if (classes % 4 == 0 && ldx % 4 == 0 && lddx % 4 == 0 && aligned16(x) && aligned16(dx))
    ce_bwd_kernel<true><<<grid, kThreads, 0, st>>>(x, ldx, target, rows, labels, classes, ignore_index, lse, count, grad_loss, dx, lddx);
  else
    ce_bwd_kernel<false><<<grid, kThreads, 0, st>>>(x, ldx, target, rows, labels, classes, ignore_index, lse, count, grad_loss, dx, lddx);
  BNN_CUDA_OK(cudaGetLastError());
  return BNN_OK;
}

}  // extern "C"
