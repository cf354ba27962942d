// prune.cu — bnn_prune: exact top-k of the log-density-at-zero key + masked overwrite
// (reference prune/prune.py:10-17; key arithmetic of torch/distributions/normal.py:87-102).
//
// key_i = log N(0; mu_i, sigma_i), the k largest get mu <- 0, rho <- -30; ties at the k-th key go to the
// lowest element index.  Two implementations behind one entry point, chosen per tensor ON THE DEVICE:
//
//  * sampled path (default; ~2 reads of (mu, rho), no key workspace traffic):
//      1. sample      exact keys of <= 32768 strided elements -> two order statistics (lo, hi) that bracket the
//                     k-th key with ~6 sigma of the sampling distribution
//      2. partition   one sweep: cheap fast-math key with a rigorous error margin classifies each element as
//                     above / below the bracket; only elements inside (or within the margin of) the bracket get
//                     the exact key; they are compacted as (key, index) candidates, the rest is counted
//      3. resolve     exact radix select among the candidates (a few % of the tensor): threshold T, number of
//                     ties to take, and the index bound for them
//      4. apply       second sweep: fast key against T (exact key only within the margin), vector stores
//    Any surprise — bracket missed, candidate buffer overflow, >= 2^32 elements — flags the tensor for
//  * the general path (exact 3-pass radix select over a stored key workspace + per-chunk tie ranking),
//    which also serves BNN_PRUNE_GENERAL and keys_out requests.  Its kernels return at once for tensors
//    that the sampled path has finished.
// All tensors of a call share the launches (table of <= 24 descriptors by value in the kernel parameters).
#include "common.cuh"

namespace bnn {
namespace {

constexpr int kThreads = 256;
constexpr int kVecPerThread = 4;
constexpr int kChunk = kThreads * 4 * kVecPerThread;           // 4096 elements
constexpr int kMaxTensors = 24;
constexpr int kBins = 2048;
constexpr int kSample = 32768;                                 // sampled keys per tensor (128 KiB of smem)
constexpr int kSmallTensor = 65536;                            // at or below: every element is a candidate
constexpr int kResolveThreads = 1024;

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

// Fast-math key and a bound on |fast - exact|.  sigma carries <= ~3e-6 relative error (ex2.approx with the
// argument scaling, approximate division, series), so q = mu^2 / (2 sigma^2) is within ~7e-6 relative and
// log(sigma) within ~4e-6 absolute; the margin 2e-5 * |q| + 2e-5 covers both with room to spare.
struct FastKey { float key, margin; };
__device__ __forceinline__ FastKey prune_key_fast(float mu, float rho) {
  float sp;
  if (rho <= -1.3862944f) {
    const float e = exp_fast(rho);
    const float z = e * rcp_ftz(2.0f + e);
    const float z2 = z * z;
    float p = fmaf(z2, 0.1111111111f, 0.1428571429f);
    p = fmaf(z2, p, 0.2f);
    p = fmaf(z2, p, 0.3333333333f);
    p = fmaf(z2, p, 1.0f);
    sp = 2.0f * z * p;
  } else {
    sp = softplus_exact(rho);
  }
  const float sigma = 1e-10f + sp;
  const float t = mu * rcp_ftz(sigma);
  const float q = 0.5f * t * t;
  FastKey k;
  k.key = -q - log_fast(sigma) - 0.9189385332046727f;
  k.margin = fmaf(2e-5f, q, 2e-5f);
  return k;
}

// per-tensor state in the workspace
struct PruneState {
  // general path
  uint32_t prefix;           // bits of the k-th largest ordered key decided so far
  uint32_t need_ranks;       // ties at the threshold must be ranked by index
  int64_t k_rem;             // how many still to take inside the current prefix class
  int64_t eq_total;          // elements equal to the final threshold
  // sampled path
  uint32_t lo, hi;           // candidate bracket (ordered keys, inclusive)
  uint32_t n_cand;           // candidates appended (may exceed the capacity: overflow)
  uint32_t general;          // 1: this tensor goes through the general path
  unsigned long long count_gt;   // elements with key > hi
  uint32_t T;                // exact threshold (ordered key of the k-th largest)
  uint32_t idx_bound;        // among key == T take the elements with index <= idx_bound
  uint32_t take_all_eq;      // every key == T is taken
  uint32_t pad;
};

struct PruneDesc {
  float* mu;
  float* rho;
  uint8_t* mask;
  float* keys_out;
  uint32_t* keys;            // workspace: ordered keys (general path) / (key, index) candidates (sampled path)
  uint32_t* hist;            // workspace: 2048 bins
  int64_t* chunk_cnt;        // workspace: per-chunk count of keys equal to the threshold (then offsets)
  PruneState* state;
  int64_t numel;
  int64_t k;
  int64_t chunk_begin;
  int64_t n_chunks;
  uint32_t cand_cap;         // capacity of the candidate buffer in (key, index) pairs
  uint32_t force_general;
  int vec;                   // mu / rho (and mask) aligned for 128-bit access
  int pad;
};
struct PruneTable {
  PruneDesc t[kMaxTensors];
  int n;
  int pad;
  int64_t total_chunks;
  uint32_t* any_general;     // workspace header: != 0 once some tensor needs the general path
};

__device__ __forceinline__ int find_tensor(const int64_t* chunk_begin, int n, int64_t chunk, int t = 0) {
#pragma unroll 1
  while (t + 1 < n && chunk >= chunk_begin[t + 1]) ++t;
  return t;
}

// ================================================================================== sampled path
// Locates the histogram bin holding the element of 0-based rank `rem` when the bins are walked downwards
// (descending = true: from bin 2047) or upwards, and the number of elements in the bins walked before it.
// Executed by warp 0 of the block (all 32 lanes); hist has kBins entries (unused high bins are zero).
// Replaces a 2048-step serial walk (each step a dependent shared load) by two 32-lane levels.
__device__ __forceinline__ void warp_find_bin(const uint32_t* hist, bool descending, uint64_t rem, uint32_t* bin_out,
                                              uint64_t* before_out) {
  const int lane = threadIdx.x & 31;
  // level 1: 32 groups of 64 bins; group g of the walk order maps to bins [64 * pos, 64 * pos + 64)
  const int pos = descending ? 31 - lane : lane;          // lane = position in walk order
  uint64_t gs = 0;
#pragma unroll 8
  for (int b = 0; b < 64; ++b) gs += hist[pos * 64 + b];
  uint64_t incl = gs;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint64_t v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  const unsigned int hit = __ballot_sync(0xffffffffu, incl > rem);
  const int g = hit ? __ffs(hit) - 1 : 31;                // first group (walk order) whose cumulative count exceeds rem
  const uint64_t before_g = __shfl_sync(0xffffffffu, incl - gs, g);
  const int gpos = descending ? 31 - g : g;
  // level 2: the 64 bins of that group, two per lane, again in walk order
  const int b0 = descending ? gpos * 64 + 63 - 2 * lane : gpos * 64 + 2 * lane;       // first of my two bins in walk order
  const int b1 = descending ? b0 - 1 : b0 + 1;
  const uint64_t c0 = hist[b0], c1 = hist[b1];
  uint64_t incl2 = c0 + c1;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint64_t v = __shfl_up_sync(0xffffffffu, incl2, o);
    if (lane >= o) incl2 += v;
  }
  const uint64_t rem2 = rem - before_g;
  const unsigned int hit2 = __ballot_sync(0xffffffffu, incl2 > rem2);
  const int l2 = hit2 ? __ffs(hit2) - 1 : 31;
  const uint64_t excl2 = __shfl_sync(0xffffffffu, incl2 - (c0 + c1), l2);
  const uint64_t c0s = __shfl_sync(0xffffffffu, c0, l2);
  const int b0s = __shfl_sync(0xffffffffu, b0, l2), b1s = __shfl_sync(0xffffffffu, b1, l2);
  const bool first = !hit2 ? false : (excl2 + c0s > rem2);
  *bin_out = static_cast<uint32_t>(first ? b0s : b1s);
  *before_out = before_g + excl2 + (first ? 0 : c0s);
}

// descending rank r (0 = largest) -> value, by <= 3 histogram passes over `n` keys in shared memory.  Digits are
// taken relative to the minimum key [mn, mx] so that a narrow key range still spreads over the 2048 bins
// (trained posteriors put every key into two or three top-bit bins: same-address shared atomics).
__device__ uint32_t smem_select_desc(const uint32_t* keys, int n, int rank, uint32_t mn, uint32_t mx, uint32_t* hist,
                                     uint32_t* bcast) {
  const uint32_t range = mx - mn;
  const int bits = range == 0u ? 0 : 32 - __clz(range);
  const int s0 = max(bits - 11, 0), s1 = max(bits - 22, 0);
  uint32_t prefix = 0, prefix_mask = 0;
  int rem = rank;
  for (int pass = 0; pass < 3; ++pass) {
    const int shift = pass == 0 ? s0 : (pass == 1 ? s1 : 0);
    const int top = pass == 0 ? bits : (pass == 1 ? s0 : s1);
    if (top == shift) continue;
    const uint32_t digit_mask = (1u << (top - shift)) - 1u;
    for (int b = threadIdx.x; b < kBins; b += blockDim.x) hist[b] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const uint32_t k = keys[i] - mn;
      if (((k ^ prefix) & prefix_mask) == 0u) atomicAdd(&hist[(k >> shift) & digit_mask], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      uint32_t bin;
      uint64_t before;
      warp_find_bin(hist, true, static_cast<uint64_t>(rem), &bin, &before);
      if (threadIdx.x == 0) { bcast[0] = bin; bcast[1] = static_cast<uint32_t>(before); }
    }
    __syncthreads();
    prefix |= bcast[0] << shift;
    prefix_mask |= digit_mask << shift;
    rem -= static_cast<int>(bcast[1]);
    __syncthreads();
  }
  return prefix + mn;
}

// 1. one block per tensor: bracket (lo, hi) of the k-th largest key from a strided sample
__global__ void __launch_bounds__(kResolveThreads) prune_sample_kernel(const __grid_constant__ PruneTable tab) {
  extern __shared__ uint32_t s_keys[];          // kSample keys
  __shared__ uint32_t s_hist[kBins];
  __shared__ uint32_t s_bcast[2];
  const PruneDesc& d = tab.t[blockIdx.x];
  PruneState st;
  st.prefix = 0; st.need_ranks = 0; st.k_rem = d.k; st.eq_total = 0;
  st.lo = 0u; st.hi = 0xffffffffu; st.n_cand = 0; st.count_gt = 0ull;
  st.T = 0; st.idx_bound = 0xffffffffu; st.take_all_eq = 1; st.pad = 0;
  const bool trivial = d.k <= 0 || d.k >= d.numel;
  st.general = (d.force_general || d.numel >= (int64_t(1) << 32)) ? 1u : 0u;
  if (!trivial && !st.general && d.numel > kSmallTensor) {
    const int m = kSample;
    for (int j0 = 0; j0 < m; j0 += 8 * kResolveThreads) {          // 8 independent strided loads in flight
      float mv[8], rv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = j0 + u * kResolveThreads + threadIdx.x;
        const int64_t i = static_cast<int64_t>((static_cast<uint64_t>(j) * static_cast<uint64_t>(d.numel)) / m);   // numel < 2^32
        mv[u] = __ldg(d.mu + i);
        rv[u] = __ldg(d.rho + i);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) s_keys[j0 + u * kResolveThreads + threadIdx.x] = order_key(prune_key(mv[u], rv[u]));
    }
    __syncthreads();
    const double p = static_cast<double>(d.k) / static_cast<double>(d.numel);
    const int r = static_cast<int>(p * m);                              // descending rank of the k-th key
    const int margin = static_cast<int>(6.0 * sqrt(m * p * (1.0 - p))) + 8;
    const int r_hi = r - margin, r_lo = r + margin;
    // block min / max of the sampled keys
    uint32_t mn = 0xffffffffu, mx = 0u;
    for (int j = threadIdx.x; j < m; j += blockDim.x) { mn = min(mn, s_keys[j]); mx = max(mx, s_keys[j]); }
    for (int o = 16; o > 0; o >>= 1) {
      mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) { s_hist[threadIdx.x >> 5] = mn; s_hist[64 + (threadIdx.x >> 5)] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t a = 0xffffffffu, b = 0u;
      for (int w = 0; w < kResolveThreads / 32; ++w) { a = min(a, s_hist[w]); b = max(b, s_hist[64 + w]); }
      s_bcast[0] = a; s_bcast[1] = b;
    }
    __syncthreads();
    mn = s_bcast[0]; mx = s_bcast[1];
    __syncthreads();
    if (r_hi >= 0) st.hi = smem_select_desc(s_keys, m, r_hi, mn, mx, s_hist, s_bcast);
    if (r_lo < m) st.lo = smem_select_desc(s_keys, m, r_lo, mn, mx, s_hist, s_bcast);
  }
  if (threadIdx.x == 0) {
    *d.state = st;
    if (st.general) atomicOr(tab.any_general, 1u);
  }
}

// Fast classification of one element against the bracket: 2 = above hi, 0 = below lo, 1 = needs the exact key.
// With Q = (mu / sigma)^2 = 2q and a = -log(sigma):  key = -Q/2 + a - c.  The fast value is within
// margin = 2e-5 * q + 2e-5 of the exact key (see prune_key_fast), so
//   key - margin > hi  <=>  fma(Q, -(1 + 2e-5)/2, a) > hi + c + 2e-5 =: hi_t
//   key + margin < lo  <=>  fma(Q, -(1 - 2e-5)/2, a) < lo + c - 2e-5 =: lo_t
// Branch-free for rho <= ln(1/4) (series softplus); 4 MUFU + ~16 FP32 operations per element.
__device__ __forceinline__ int classify_fast_small_rho(float mu, float rho, float lo_t, float hi_t) {
  const float e = exp_fast(rho);                       // flushed to 0 below rho ~ -87: sigma = 1e-10, as in fp32 torch
  const float z = e * rcp_ftz(2.0f + e);
  const float z2 = z * z;
  float p = fmaf(z2, 0.1111111111f, 0.1428571429f);
  p = fmaf(z2, p, 0.2f);
  p = fmaf(z2, p, 0.3333333333f);
  p = fmaf(z2, p, 1.0f);
  const float sigma = fmaf(2.0f * z, p, 1e-10f);       // >= 1e-10: normal range for rcp / lg2
  const float t = mu * rcp_ftz(sigma);
  const float Q = t * t;
  const float a = lg2_ftz(sigma) * -0.6931471805599453f;
  const bool above = fmaf(Q, -0.50001f, a) > hi_t;
  const bool below = fmaf(Q, -0.49999f, a) < lo_t;
  return above ? 2 : (below ? 0 : 1);
}
__device__ __forceinline__ int classify_fast(float mu, float rho, float lo_t, float hi_t) {
  if (rho <= -1.3862944f) return classify_fast_small_rho(mu, rho, lo_t, hi_t);
  const FastKey f = prune_key_fast(mu, rho);
  // same thresholds, general form: key - margin > hi  <=>  key + c - margin + 2e-5 > hi_t
  if (f.key + 0.9189385332046727f - f.margin + 2e-5f > hi_t) return 2;
  if (f.key + 0.9189385332046727f + f.margin - 2e-5f < lo_t) return 0;
  return 1;
}

// 2. sweep: count the elements above the bracket, compact the candidates
__global__ void __launch_bounds__(kThreads) prune_partition_kernel(const __grid_constant__ PruneTable tab) {
  __shared__ int64_t s_begin[kMaxTensors];
  __shared__ unsigned int s_cnt[kThreads / 32];
  __shared__ uint16_t s_work[kChunk];          // element offsets (within the chunk) that need the exact key
  __shared__ uint32_t s_cand[2 * kChunk];      // (key, index) candidates of this chunk
  __shared__ unsigned int s_nwork, s_n;
  __shared__ uint32_t s_base;
  if (threadIdx.x < tab.n) s_begin[threadIdx.x] = tab.t[threadIdx.x].chunk_begin;
  if (threadIdx.x == 0) { s_n = 0; s_nwork = 0; }
  __syncthreads();
  int cur = -1;
  bool active = false;
  uint32_t lo = 0, hi = 0;
  float lo_f = 0.f, hi_f = 0.f;
  int hshift = 0;
  unsigned int gt = 0;
  auto flush = [&]() {
    // block-wide (uniform) — adds this block's count of "above" elements of tensor `cur`
    unsigned int v = warp_sum(gt);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long tot = 0;
      for (int w = 0; w < kThreads / 32; ++w) tot += s_cnt[w];
      if (tot) atomicAdd(&tab.t[cur].state->count_gt, tot);
    }
    gt = 0;
  };
  for (int64_t chunk = blockIdx.x; chunk < tab.total_chunks; chunk += gridDim.x) {
    const int t = find_tensor(s_begin, tab.n, chunk, cur < 0 ? 0 : cur);
    if (t != cur) {
      if (cur >= 0 && active) flush();
      cur = t;
      const PruneDesc& d0 = tab.t[t];
      const PruneState st = *d0.state;
      active = !st.general && d0.k > 0 && d0.k < d0.numel;
      lo = st.lo; hi = st.hi;
      // thresholds of classify_fast (open brackets: +-inf) and the digit shift of the candidates' first-level histogram
      lo_f = lo == 0u ? -INFINITY : unorder_key(lo) + 0.9189385332046727f - 2e-5f;
      hi_f = hi == 0xffffffffu ? INFINITY : unorder_key(hi) + 0.9189385332046727f + 2e-5f;
      const uint32_t range = hi - lo;
      const int bits = range == 0u ? 0 : 32 - __clz(range);
      hshift = bits > 11 ? bits - 11 : 0;
    }
    if (!active) continue;
    const PruneDesc& d = tab.t[t];
    const int64_t base = (chunk - d.chunk_begin) * kChunk;
    // phase 1: fast keys; decided elements are counted, the undecided ones are remembered in a per-thread
    // bit mask and listed with ONE warp scan + ONE shared atomic per warp
    uint32_t undecided = 0;          // bit e: this thread's e-th element of the chunk needs the exact key
    const bool vec_chunk = d.vec && base + kChunk <= d.numel;
    if (vec_chunk) {
      float4 m[kVecPerThread], r[kVecPerThread];
#pragma unroll
      for (int j = 0; j < kVecPerThread; ++j) {
        const int64_t i = base + (static_cast<int64_t>(j) * kThreads + threadIdx.x) * 4;
        m[j] = ldg_stream4(d.mu + i);
        r[j] = ldg_stream4(d.rho + i);
      }
#pragma unroll
      for (int j = 0; j < kVecPerThread; ++j) {
        const float mm[4] = {m[j].x, m[j].y, m[j].z, m[j].w};
        const float rr[4] = {r[j].x, r[j].y, r[j].z, r[j].w};
        if (fmaxf(fmaxf(rr[0], rr[1]), fmaxf(rr[2], rr[3])) <= -1.3862944f) {      // one branch per 4 elements
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int c = classify_fast_small_rho(mm[q], rr[q], lo_f, hi_f);
            gt += (c == 2);
            undecided |= (c == 1 ? 1u : 0u) << (j * 4 + q);
          }
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int c = classify_fast(mm[q], rr[q], lo_f, hi_f);
            gt += (c == 2);
            undecided |= (c == 1 ? 1u : 0u) << (j * 4 + q);
          }
        }
      }
    } else {
#pragma unroll 4
      for (int j = 0; j < 4 * kVecPerThread; ++j) {
        const uint32_t off = static_cast<uint32_t>(j * kThreads + threadIdx.x);
        int c = 0;
        if (base + off < d.numel) c = classify_fast(d.mu[base + off], d.rho[base + off], lo_f, hi_f);
        gt += (c == 2);
        undecided |= (c == 1 ? 1u : 0u) << j;
      }
    }
    {
      const int lane = threadIdx.x & 31;
      const unsigned int mine = __popc(undecided);
      unsigned int incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      const unsigned int total = __shfl_sync(0xffffffffu, incl, 31);
      unsigned int wbase = 0;
      if (total != 0u) {           // warp-uniform
        if (lane == 31) wbase = atomicAdd(&s_nwork, total);
        wbase = __shfl_sync(0xffffffffu, wbase, 31);
        unsigned int slot = wbase + incl - mine;
        while (undecided != 0u) {
          const int e = __ffs(undecided) - 1;
          undecided &= undecided - 1u;
          const uint32_t off = vec_chunk ? static_cast<uint32_t>(((e >> 2) * kThreads + threadIdx.x) * 4 + (e & 3))
                                         : static_cast<uint32_t>(e * kThreads + threadIdx.x);
          s_work[slot++] = static_cast<uint16_t>(off);
        }
      }
    }
    __syncthreads();
    // phase 2: exact keys of the listed elements, evaluated densely
    const unsigned int n_work = s_nwork;
    for (unsigned int w0 = 0; w0 < n_work; w0 += kThreads) {       // uniform trip count
      const unsigned int w = w0 + threadIdx.x;
      int c = 0;
      uint32_t ok = 0, idx = 0;
      if (w < n_work) {
        const int64_t i = base + s_work[w];
        idx = static_cast<uint32_t>(i);
        ok = order_key(prune_key(d.mu[i], d.rho[i]));
        c = ok > hi ? 2 : (ok >= lo ? 1 : 0);
        if (c == 1) {      // first-level histogram of the candidates (digits relative to lo), read by the resolve kernel
          const uint32_t bin = (ok - lo) >> hshift;
          atomicAdd(d.hist + (bin < static_cast<uint32_t>(kBins) ? bin : static_cast<uint32_t>(kBins - 1)), 1u);
        }
      }
      gt += (c == 2);
      const unsigned int bal = __ballot_sync(0xffffffffu, c == 1);
      if (bal != 0u) {
        const int lane = threadIdx.x & 31;
        const int leader = __ffs(bal) - 1;
        unsigned int sb = 0;
        if (lane == leader) sb = atomicAdd(&s_n, static_cast<unsigned int>(__popc(bal)));
        sb = __shfl_sync(0xffffffffu, sb, leader);
        if (c == 1) {
          const unsigned int slot = sb + __popc(bal & ((1u << lane) - 1u));
          s_cand[2 * slot] = ok;
          s_cand[2 * slot + 1] = idx;
        }
      }
    }
    __syncthreads();
    // phase 3: one global reservation per block and chunk, coalesced copy-out
    const unsigned int n_here = s_n;
    if (n_here != 0u) {           // uniform
      if (threadIdx.x == 0) s_base = atomicAdd(&d.state->n_cand, n_here);
      __syncthreads();
      const uint32_t gbase = s_base;
      for (unsigned int c = threadIdx.x; c < n_here; c += kThreads) {
        const uint64_t slot = static_cast<uint64_t>(gbase) + c;
        if (slot < d.cand_cap)
          *reinterpret_cast<uint2*>(d.keys + 2 * slot) = make_uint2(s_cand[2 * c], s_cand[2 * c + 1]);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) { s_n = 0; s_nwork = 0; }
    __syncthreads();
  }
  if (cur >= 0 && active) flush();
}

// rank-th (0-based) value of the `n` candidates' field (`field` 0 = key descending, 1 = index ascending), restricted
// to entries with key == only_key when field == 1.  Digits are taken relative to the minimum so that a narrow
// bracket still spreads over the 2048 bins.  Returns the value; *count_before = entries strictly before it in the
// order, *count_equal = entries equal to it.
__device__ uint32_t cand_select(const uint32_t* cand, uint32_t n, int field, uint32_t only_key, uint64_t rank,
                                uint32_t* hist, uint32_t* bcast, uint64_t* count_before, uint32_t* count_equal) {
  // min / max of the field
  uint32_t mn = 0xffffffffu, mx = 0u;
  for (uint32_t i0 = 0; i0 < n; i0 += 8u * blockDim.x) {
    uint2 e[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const uint32_t i = i0 + u * blockDim.x + threadIdx.x;
      e[u] = i < n ? reinterpret_cast<const uint2*>(cand)[i] : make_uint2(0u, 0u);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const uint32_t i = i0 + u * blockDim.x + threadIdx.x;
      if (i >= n || (field == 1 && e[u].x != only_key)) continue;
      const uint32_t v = field == 0 ? e[u].x : e[u].y;
      mn = min(mn, v);
      mx = max(mx, v);
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { hist[threadIdx.x >> 5] = mn; hist[64 + (threadIdx.x >> 5)] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t a = 0xffffffffu, b = 0u;
    for (unsigned int w = 0; w < blockDim.x / 32; ++w) { a = min(a, hist[w]); b = max(b, hist[64 + w]); }
    bcast[0] = a; bcast[1] = b;
  }
  __syncthreads();
  mn = bcast[0]; mx = bcast[1];
  __syncthreads();
  const uint32_t range = mx - mn;
  const int bits = range == 0u ? 0 : 32 - __clz(range);
  uint32_t prefix = 0, prefix_mask = 0;      // on v - mn
  uint64_t rem = rank, before = 0;
  uint32_t equal = 0;
  const int s0 = max(bits - 11, 0), s1 = max(bits - 22, 0);
  for (int pass = 0; pass < 3; ++pass) {
    // pass p decides the bits [shift, top) of v - mn; at most 11 bits wide
    const int shift = pass == 0 ? s0 : (pass == 1 ? s1 : 0);
    const int top = pass == 0 ? bits : (pass == 1 ? s0 : s1);
    if (top == shift) continue;                                   // nothing left to decide in this pass
    const uint32_t digit_mask = (1u << (top - shift)) - 1u;
    for (int b = threadIdx.x; b < kBins; b += blockDim.x) hist[b] = 0;
    __syncthreads();
    for (uint32_t i0 = 0; i0 < n; i0 += 8u * blockDim.x) {
      uint2 e[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const uint32_t i = i0 + u * blockDim.x + threadIdx.x;
        e[u] = i < n ? reinterpret_cast<const uint2*>(cand)[i] : make_uint2(0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const uint32_t i = i0 + u * blockDim.x + threadIdx.x;
        if (i >= n || (field == 1 && e[u].x != only_key)) continue;
        const uint32_t v = (field == 0 ? e[u].x : e[u].y) - mn;
        if (((v ^ prefix) & prefix_mask) == 0u) atomicAdd(&hist[(v >> shift) & digit_mask], 1u);
      }
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      uint32_t bin;
      uint64_t acc;
      warp_find_bin(hist, field == 0, rem, &bin, &acc);
      if (threadIdx.x == 0) {
        bcast[0] = bin;
        bcast[1] = static_cast<uint32_t>(acc);
        bcast[2] = static_cast<uint32_t>(acc >> 32);
        bcast[3] = hist[bin];
      }
    }
    __syncthreads();
    prefix |= bcast[0] << shift;
    prefix_mask |= digit_mask << shift;
    const uint64_t acc = static_cast<uint64_t>(bcast[1]) | (static_cast<uint64_t>(bcast[2]) << 32);
    rem -= acc;
    before += acc;
    equal = bcast[3];
    __syncthreads();
  }
  if (bits == 0) {           // every entry has the same value
    uint32_t c = 0;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
      if (field == 0 || cand[2 * static_cast<size_t>(i)] == only_key) ++c;
    c = warp_sum(c);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) hist[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t tot = 0;
      for (unsigned int w = 0; w < blockDim.x / 32; ++w) tot += hist[w];
      bcast[0] = tot;
    }
    __syncthreads();
    equal = bcast[0];
    before = 0;
    __syncthreads();
  }
  *count_before = before;
  *count_equal = equal;
  return prefix + mn;
}

// 3. one block per tensor: exact threshold among the candidates.  The partition sweep left a 2048-bin histogram of the
// candidates; the bin holding the wanted rank is found from it, its members (a few hundred) are filtered into shared
// memory in ONE pass over the candidate list, and the exact select runs there.  A bin with more members than the
// shared list holds (massive ties) takes the multi-pass select over the whole list instead.
constexpr uint32_t kResolveList = 4096;      // (key, index) pairs in shared memory

__global__ void __launch_bounds__(kResolveThreads) prune_resolve_kernel(const __grid_constant__ PruneTable tab) {
  __shared__ uint32_t s_hist[kBins];
  __shared__ uint32_t s_bcast[4];
  __shared__ uint32_t s_list[2 * kResolveList];
  __shared__ unsigned int s_count;
  const PruneDesc& d = tab.t[blockIdx.x];
  PruneState st = *d.state;
  if (st.general || d.k <= 0 || d.k >= d.numel) return;
  const uint64_t k = static_cast<uint64_t>(d.k);
  const bool ok = st.n_cand <= d.cand_cap && st.count_gt < k && k <= st.count_gt + st.n_cand;
  for (int b = threadIdx.x; b < kBins; b += blockDim.x) {
    s_hist[b] = d.hist[b];
    d.hist[b] = 0;                                   // clean for the general path / the next call
  }
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();
  if (ok) {
    const uint64_t rank = k - st.count_gt - 1;                  // 0-based, descending, among the candidates
    const uint32_t range = st.hi - st.lo;
    const int bits = range == 0u ? 0 : 32 - __clz(range);
    const int hshift = bits > 11 ? bits - 11 : 0;
    if (threadIdx.x < 32) {
      uint32_t bin;
      uint64_t before;
      warp_find_bin(s_hist, true, rank, &bin, &before);
      if (threadIdx.x == 0) { s_bcast[0] = bin; s_bcast[1] = static_cast<uint32_t>(before); s_bcast[2] = s_hist[bin]; }
    }
    __syncthreads();
    const uint32_t bin = s_bcast[0], before_bins = s_bcast[1], in_bin = s_bcast[2];
    __syncthreads();
    uint64_t above = 0;
    uint32_t eq = 0, T = 0;
    const bool small = in_bin <= kResolveList;
    if (small) {
      // one pass: members of the bin -> shared list
      for (uint32_t i0 = 0; i0 < st.n_cand; i0 += 8u * blockDim.x) {
        uint2 e[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint32_t i = i0 + u * blockDim.x + threadIdx.x;
          e[u] = i < st.n_cand ? reinterpret_cast<const uint2*>(d.keys)[i] : make_uint2(0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint32_t i = i0 + u * blockDim.x + threadIdx.x;
          if (i >= st.n_cand) continue;
          uint32_t b = (e[u].x - st.lo) >> hshift;
          b = b < static_cast<uint32_t>(kBins) ? b : static_cast<uint32_t>(kBins - 1);
          if (b == bin) {
            const unsigned int slot = atomicAdd(&s_count, 1u);
            if (slot < kResolveList) { s_list[2 * slot] = e[u].x; s_list[2 * slot + 1] = e[u].y; }
          }
        }
      }
      __syncthreads();
      const uint32_t n_list = s_count < kResolveList ? s_count : kResolveList;
      T = cand_select(s_list, n_list, 0, 0u, rank - before_bins, s_hist, s_bcast, &above, &eq);
      above += before_bins;
      const uint64_t take_eq = rank - above + 1;
      st.T = T;
      st.take_all_eq = take_eq >= eq ? 1u : 0u;
      st.idx_bound = 0xffffffffu;
      if (!st.take_all_eq) {
        uint64_t b2 = 0;
        uint32_t e2 = 0;
        st.idx_bound = cand_select(s_list, n_list, 1, T, take_eq - 1, s_hist, s_bcast, &b2, &e2);
      }
    } else {
      T = cand_select(d.keys, st.n_cand, 0, 0u, rank, s_hist, s_bcast, &above, &eq);
      const uint64_t take_eq = rank - above + 1;                // how many of the key == T entries are taken
      st.T = T;
      st.take_all_eq = take_eq >= eq ? 1u : 0u;
      st.idx_bound = 0xffffffffu;
      if (!st.take_all_eq) {
        uint64_t b2 = 0;
        uint32_t e2 = 0;
        st.idx_bound = cand_select(d.keys, st.n_cand, 1, T, take_eq - 1, s_hist, s_bcast, &b2, &e2);
      }
    }
  } else {
    st.general = 1u;
  }
  if (threadIdx.x == 0) {
    *d.state = st;
    if (st.general) atomicOr(tab.any_general, 1u);
  }
}

// 4. second sweep: apply the mask
__global__ void __launch_bounds__(kThreads) prune_apply_sampled_kernel(const __grid_constant__ PruneTable tab) {
  __shared__ int64_t s_begin[kMaxTensors];
  if (threadIdx.x < tab.n) s_begin[threadIdx.x] = tab.t[threadIdx.x].chunk_begin;
  __syncthreads();
  int cur = -1;
  int mode = 0;            // 0 skip tensor, 1 none, 2 all, 3 select
  uint32_t T = 0, idx_bound = 0, all_eq = 0;
  float T_f = 0.f;
  for (int64_t chunk = blockIdx.x; chunk < tab.total_chunks; chunk += gridDim.x) {
    const int t = find_tensor(s_begin, tab.n, chunk, cur < 0 ? 0 : cur);
    if (t != cur) {
      cur = t;
      const PruneDesc& d0 = tab.t[t];
      const PruneState st = *d0.state;
      mode = st.general ? 0 : (d0.k <= 0 ? 1 : (d0.k >= d0.numel ? 2 : 3));
      T = st.T; idx_bound = st.idx_bound; all_eq = st.take_all_eq;
      T_f = unorder_key(T);
    }
    if (mode == 0) continue;
    const PruneDesc& d = tab.t[t];
    const int64_t base = (chunk - d.chunk_begin) * kChunk;
    auto decide = [&](float mu, float rho, uint32_t idx) -> bool {
      if (mode != 3) return mode == 2;
      const FastKey f = prune_key_fast(mu, rho);
      if (f.key - f.margin > T_f) return true;
      if (f.key + f.margin < T_f) return false;
      const uint32_t ok = order_key(prune_key(mu, rho));
      return ok > T || (ok == T && (all_eq || idx <= idx_bound));
    };
    if (d.vec && base + kChunk <= d.numel) {
      float4 m[kVecPerThread], r[kVecPerThread];
      if (mode == 3) {
#pragma unroll
        for (int j = 0; j < kVecPerThread; ++j) {
          const int64_t i = base + (static_cast<int64_t>(j) * kThreads + threadIdx.x) * 4;
          m[j] = ldg_stream4(d.mu + i);
          r[j] = ldg_stream4(d.rho + i);
        }
      }
#pragma unroll
      for (int j = 0; j < kVecPerThread; ++j) {
        const int64_t i = base + (static_cast<int64_t>(j) * kThreads + threadIdx.x) * 4;
        const uint32_t i32 = static_cast<uint32_t>(i);
        bool tk[4];
        if (mode == 3) {
          tk[0] = decide(m[j].x, r[j].x, i32);
          tk[1] = decide(m[j].y, r[j].y, i32 + 1);
          tk[2] = decide(m[j].z, r[j].z, i32 + 2);
          tk[3] = decide(m[j].w, r[j].w, i32 + 3);
        } else {
          tk[0] = tk[1] = tk[2] = tk[3] = (mode == 2);
        }
        const bool any = tk[0] || tk[1] || tk[2] || tk[3];
        const bool all = tk[0] && tk[1] && tk[2] && tk[3];
        if (all) {
          *reinterpret_cast<float4*>(d.mu + i) = make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(d.rho + i) = make_float4(-30.f, -30.f, -30.f, -30.f);
        } else if (any) {      // mode 3 only: the old values are in registers, store whole vectors
          *reinterpret_cast<float4*>(d.mu + i) =
              make_float4(tk[0] ? 0.f : m[j].x, tk[1] ? 0.f : m[j].y, tk[2] ? 0.f : m[j].z, tk[3] ? 0.f : m[j].w);
          *reinterpret_cast<float4*>(d.rho + i) = make_float4(tk[0] ? -30.f : r[j].x, tk[1] ? -30.f : r[j].y,
                                                              tk[2] ? -30.f : r[j].z, tk[3] ? -30.f : r[j].w);
        }
        if (d.mask != nullptr)
          *reinterpret_cast<uchar4*>(d.mask + i) = make_uchar4(tk[0], tk[1], tk[2], tk[3]);
      }
    } else {
      for (int j = 0; j < 4 * kVecPerThread; ++j) {
        const int64_t i = base + static_cast<int64_t>(j) * kThreads + threadIdx.x;
        if (i < d.numel) {
          const bool take = mode == 3 ? decide(d.mu[i], d.rho[i], static_cast<uint32_t>(i)) : (mode == 2);
          if (take) { d.mu[i] = 0.0f; d.rho[i] = -30.0f; }
          if (d.mask != nullptr) d.mask[i] = take ? 1 : 0;
        }
      }
    }
  }
}

// ================================================================================== general path
__device__ __forceinline__ void flush_hist(uint32_t* s_hist, uint32_t* g_hist) {
  __syncthreads();
  for (int b = threadIdx.x; b < kBins; b += kThreads) {
    const uint32_t c = s_hist[b];
    if (c) atomicAdd(g_hist + b, c);
    s_hist[b] = 0;
  }
  __syncthreads();
}

#define BNN_RETURN_UNLESS_GENERAL(tab)                                   \
  if (*reinterpret_cast<volatile uint32_t*>((tab).any_general) == 0u) return

// pass 0: keys from (mu, rho) -> workspace, histogram of the top 11 bits
// pass 1/2: histogram of the next digit among keys matching the prefix decided so far
template <int kPass>
__global__ void __launch_bounds__(kThreads) prune_hist_kernel(const __grid_constant__ PruneTable tab) {
  BNN_RETURN_UNLESS_GENERAL(tab);
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
  bool general = false, active = false;
  for (int64_t chunk = blockIdx.x; chunk < tab.total_chunks; chunk += gridDim.x) {
    const int t = find_tensor(s_begin, tab.n, chunk, cur < 0 ? 0 : cur);
    if (t != cur) {
      if (cur >= 0 && active) flush_hist(s_hist, tab.t[cur].hist);
      cur = t;
      const PruneDesc& d0 = tab.t[t];
      general = d0.state->general != 0u;
      active = general && d0.k > 0 && d0.k < d0.numel;      // k == 0 / k == numel need no selection
      if (kPass > 0) prefix = d0.state->prefix;
    }
    const PruneDesc& d = tab.t[t];
    if (!general || (kPass > 0 && !active)) continue;
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
  BNN_RETURN_UNLESS_GENERAL(tab);
  const PruneDesc& d = tab.t[blockIdx.x];
  if (d.state->general == 0u) return;
  __shared__ uint32_t s_hist[kBins];
  constexpr int shift = kPass == 0 ? 21 : (kPass == 1 ? 10 : 0);
  for (int b = threadIdx.x; b < kBins; b += kThreads) {
    s_hist[b] = d.hist[b];
    d.hist[b] = 0;                       // clean for the next pass / next call
  }
  __syncthreads();
  if (threadIdx.x >= 32) return;
  PruneState st = *d.state;
  if (kPass == 0) {
    st.prefix = 0;
    st.need_ranks = 0;
    st.k_rem = d.k;
    st.eq_total = 0;
  }
  if (d.k > 0 && d.k < d.numel) {
    uint32_t b;
    uint64_t above;
    warp_find_bin(s_hist, true, static_cast<uint64_t>(st.k_rem - 1), &b, &above);     // rank k_rem - 1, 0-based
    st.prefix |= b << shift;
    st.k_rem -= static_cast<int64_t>(above);
    if (kPass == 2) {
      st.eq_total = s_hist[b];
      st.need_ranks = (st.k_rem < st.eq_total) ? 1u : 0u;
    }
  }
  if (threadIdx.x == 0) *d.state = st;
}

// per-chunk count of keys equal to the threshold (only for tensors with ties at the boundary)
__global__ void __launch_bounds__(kThreads) prune_count_eq_kernel(const __grid_constant__ PruneTable tab) {
  BNN_RETURN_UNLESS_GENERAL(tab);
  __shared__ int64_t s_begin[kMaxTensors];
  __shared__ int s_cnt[kThreads / 32];
  if (threadIdx.x < tab.n) s_begin[threadIdx.x] = tab.t[threadIdx.x].chunk_begin;
  __syncthreads();
  int cur = -1;
  for (int64_t chunk = blockIdx.x; chunk < tab.total_chunks; chunk += gridDim.x) {
    const int t = find_tensor(s_begin, tab.n, chunk, cur < 0 ? 0 : cur);
    cur = t;
    const PruneDesc& d = tab.t[t];
    const PruneState st = *d.state;
    if (!st.general || !st.need_ranks) continue;
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
  BNN_RETURN_UNLESS_GENERAL(tab);
  const PruneDesc& d = tab.t[blockIdx.x];
  if (!d.state->general || !d.state->need_ranks) return;
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
  BNN_RETURN_UNLESS_GENERAL(tab);
  __shared__ int64_t s_begin[kMaxTensors];
  __shared__ int s_warp[kThreads / 32];
  if (threadIdx.x < tab.n) s_begin[threadIdx.x] = tab.t[threadIdx.x].chunk_begin;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int cur = -1;
  for (int64_t chunk = blockIdx.x; chunk < tab.total_chunks; chunk += gridDim.x) {
    const int t = find_tensor(s_begin, tab.n, chunk, cur < 0 ? 0 : cur);
    cur = t;
    const PruneDesc& d = tab.t[t];
    const PruneState st = *d.state;
    if (!st.general) continue;
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

size_t keys_bytes(int64_t numel) {       // ordered keys (general) or (key, index) candidates (sampled)
  const int64_t small = numel < kSmallTensor ? numel : kSmallTensor;
  const int64_t a = numel * 4, b = small * 8;
  return static_cast<size_t>(a > b ? a : b);
}

size_t prune_ws_one(int64_t numel) {
  const int64_t chunks = (numel + kChunk - 1) / kChunk;
  return align_up(keys_bytes(numel), 256) + align_up(kBins * 4, 256) +
         align_up(static_cast<size_t>(chunks) * 8, 256) + align_up(sizeof(PruneState), 256);
}

}  // namespace
}  // namespace bnn

using namespace bnn;

extern "C" {

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
  uint32_t* any_general = reinterpret_cast<uint32_t*>(workspace);
  char* ws = static_cast<char*>(workspace) + 256;
  const int max_grid = sm_count() * 8;
  static bool smem_set = false;
  const size_t sample_smem = static_cast<size_t>(kSample) * sizeof(uint32_t);
  if (!smem_set) {
    BNN_CUDA_OK(cudaFuncSetAttribute(prune_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(sample_smem)));
    smem_set = true;
  }

  for (int first = 0; first < n_tensors; first += kMaxTensors) {
    const int n = (n_tensors - first < kMaxTensors) ? n_tensors - first : kMaxTensors;
    PruneTable tab;
    tab.n = 0;
    tab.pad = 0;
    tab.any_general = any_general;
    int64_t chunks = 0;
    for (int i = 0; i < n; ++i) {
      const bnn_prune_tensor& t = tensors[first + i];
      if (t.numel == 0) continue;
      PruneDesc& d = tab.t[tab.n++];
      const int64_t nch = (t.numel + kChunk - 1) / kChunk;
      d.mu = t.mu; d.rho = t.rho; d.mask = t.mask_out; d.keys_out = t.keys_out;
      d.numel = t.numel; d.k = t.k; d.chunk_begin = chunks; d.n_chunks = nch;
      d.keys = reinterpret_cast<uint32_t*>(ws);
      d.cand_cap = static_cast<uint32_t>(keys_bytes(t.numel) / 8);
      ws += align_up(keys_bytes(t.numel), 256);
      d.hist = reinterpret_cast<uint32_t*>(ws); ws += align_up(kBins * 4, 256);
      d.chunk_cnt = reinterpret_cast<int64_t*>(ws); ws += align_up(static_cast<size_t>(nch) * 8, 256);
      d.state = reinterpret_cast<PruneState*>(ws); ws += align_up(sizeof(PruneState), 256);
      d.force_general = ((t.flags & BNN_PRUNE_GENERAL) != 0u || t.keys_out != nullptr) ? 1u : 0u;
      d.vec = aligned16(t.mu) && aligned16(t.rho) && (t.mask_out == nullptr || (reinterpret_cast<uintptr_t>(t.mask_out) & 3u) == 0);
      d.pad = 0;
      BNN_CUDA_OK(cudaMemsetAsync(d.hist, 0, kBins * 4, st));
      chunks += nch;
    }
    if (tab.n == 0) continue;
    tab.total_chunks = chunks;
    BNN_CUDA_OK(cudaMemsetAsync(any_general, 0, sizeof(uint32_t), st));
    const int grid = static_cast<int>(chunks < max_grid ? chunks : max_grid);
    // sampled path
    prune_sample_kernel<<<tab.n, kResolveThreads, sample_smem, st>>>(tab);
    prune_partition_kernel<<<grid, kThreads, 0, st>>>(tab);
    prune_resolve_kernel<<<tab.n, kResolveThreads, 0, st>>>(tab);
    prune_apply_sampled_kernel<<<grid, kThreads, 0, st>>>(tab);
    // general path (kernels return immediately unless a tensor asked for it)
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
