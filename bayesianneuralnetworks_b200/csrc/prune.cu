// prune.cu — bnn_prune: exact top-k of the log-density-at-zero key + masked overwrite
// (reference prune/prune.py:10-17; key arithmetic of torch/distributions/normal.py:87-102).
//
// key_i = log N(0; mu_i, sigma_i), the k largest get mu <- 0, rho <- -30; ties at the k-th key go to the
// lowest element index.  Two implementations behind one entry point, chosen per tensor ON THE DEVICE:
//
//  * sampled path (default; two reads of (mu, rho), one write, no key workspace traffic).  Every element gets a cheap
//    CERTIFIED INTERVAL for its key (fast math + a rigorous error margin), expressed in grid coordinates y: a 2048-bin
//    grid laid over a bracket of the k-th key.
//      1. sample    fast keys of 32768 strided elements (gathered by a machine-wide kernel) -> two order statistics
//                   that bracket the k-th key with ~6 sigma of the sampling distribution -> the grid
//      2. bin       sweep 1 (read only): elements whose interval lies above / below the grid are counted, the others
//                   (~3 %) enter two histograms: bin of the interval's lower end, bin of its upper end
//      3. bracket   from the two histograms: bins j_lo <= j_hi that PROVABLY enclose the k-th key, the exact number of
//                   elements above them (all pruned) and of elements that overlap them (~1000 per tensor: deferred)
//      4. apply     sweep 2: same interval arithmetic; above -> pruned (vector stores), below -> kept, overlapping ->
//                   (index, mu, rho) appended to a short list
//      5. finish    exact keys (torch op order) of the deferred elements, exact select of the missing k - |above|
//                   among them (ties -> lowest index), scattered writes
//    Nothing is modified before step 3 has proven the bracket and sized the list; any surprise — bracket missed,
//    massive ties, >= 2^32 elements — flags the tensor for
//  * the general path (exact 3-pass radix select over a stored key workspace + per-chunk tie ranking),
//    which also serves BNN_PRUNE_GENERAL and keys_out requests.  Its kernels return at once for tensors
//    that the sampled path has finished.
// All tensors of a call share the launches (table of <= 24 descriptors by value in the kernel parameters).
//
//  * bnn_prune_into — the same selection OUT OF PLACE in ONE sweep (8 B read + 8 B written per pair instead of two reads
//    and a write).  In place nothing may be modified before the bracket is proven, which costs the read-only sweep 1;
//    with a separate output the sweep can commit at once to what lies outside the sample's grid — above: pruned, below:
//    copied — copies the ~3 % inside the grid unchanged while histogramming and listing them, and the bracket step then
//    runs on the histograms as before: a resolve kernel walks the listed elements (certainly above the bracket ->
//    scattered prune in the output, overlapping it -> the short exact list), finish as before.  If the sample's grid turns
//    out wrong (or anything else surprises) the INPUT is intact: the tensor is copied and goes through the general path
//    on the output.  The Python layer swaps the parameters' storage for the output (prune/prune.py).
#include <cooperative_groups.h>

#include <mutex>
#include <vector>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace bnn {
namespace {

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

constexpr int kThreads = 256;
constexpr int kVecPerThread = 4;
constexpr int kChunk = kThreads * 4 * kVecPerThread;           // 4096 elements
constexpr int kMaxTensors = 24;
constexpr int kBins = 2048;
constexpr int kHistStride = kBins + 64;                        // a histogram slot: 2049 used entries, 256-byte multiple
constexpr int kSample = 32768;                                 // sampled keys per tensor: single elements at a fixed stride
constexpr int kSampleBig = 262144;                             // tensors of >= 4 * kSampleBig elements: 2048 runs of 128
constexpr int kSampleRun = 128;                                //   consecutive elements (whole sectors; the bracket is
                                                               //   2.8x narrower: fewer elements to defer / list)
__host__ __device__ inline int sample_size(int64_t numel) { return numel >= 4 * static_cast<int64_t>(kSampleBig) ? kSampleBig : kSample; }
constexpr int kSmallTensor = 65536;                            // at or below: every element is deferred to the exact select
constexpr int kResolveThreads = 1024;
constexpr int kFinishThreads = 512;                             // finish kernel: leaves room on an SM that is busy sweeping
constexpr int kSweepChunks = 8;                                // chunks per block of the out-of-place sweep
constexpr float kMidGrid = 1024.0f;                            // bnn_prune_into: provisional decision inside the grid
constexpr int kUnit = 512;                                     // elements one warp handles per pass (8 units per chunk)

// ordered key: unsigned order == float order (larger float -> larger uint)
__device__ __forceinline__ uint32_t order_key(float f) {
  const uint32_t u = __float_as_uint(f);
  if ((u & 0x7fffffffu) > 0x7f800000u) return 0xffffffffu;      // NaN of either sign ranks first, as in torch.topk
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

// per-tensor state in the workspace
struct PruneState {
  // general path
  uint32_t prefix;           // bits of the k-th largest ordered key decided so far
  uint32_t need_ranks;       // ties at the threshold must be ranked by index
  int64_t k_rem;             // how many still to take inside the current prefix class
  int64_t eq_total;          // elements equal to the final threshold
  // sampled path
  float scale, offs_minus, offs_plus;    // grid: y = fma(key2 bound, scale, offs), bins [0, 2048)
  uint32_t general;          // 1: this tensor goes through the general path
  unsigned long long count_above;   // sweep 1: elements certified above the grid
  float a_thr, b_thr;        // sweep 2: y_minus >= a_thr -> pruned; y_plus < b_thr -> kept; else deferred
  uint32_t defer_all;        // small tensor: every element is deferred (exact select over the whole tensor)
  uint32_t n_take;           // how many of the deferred elements are pruned
  uint32_t expect_deferred;  // size of the deferred list, known from the histograms before sweep 2
  uint32_t n_deferred;       // entries appended by sweep 2 (bnn_prune_into: in-grid elements listed by the single sweep)
  uint32_t n_deferred2;      // bnn_prune_into: entries the resolve kernel passed on to the exact select
};

struct PruneDesc {
  float* mu;
  float* rho;
  uint8_t* mask;
  float* keys_out;
  uint32_t* keys;            // workspace: ordered keys (general path) / on the sampled path first the kSample sampled keys,
                             // later the deferred [index | mu | rho] x defer_cap, then (key, index) pairs x defer_cap
  uint32_t* hist;            // workspace: kHistStride entries (lower interval ends: 2049 used / general path digits: 2048)
  uint32_t* hist_plus;       // workspace: kHistStride entries (upper interval ends: 2049 used)
  int64_t* chunk_cnt;        // workspace: per-chunk count of keys equal to the threshold (then offsets)
  PruneState* state;
  int64_t numel;
  int64_t k;
  int64_t chunk_begin;
  int64_t n_chunks;
  uint32_t defer_cap;        // capacity of the deferred list in entries
  uint32_t force_general;
  int vec;                   // mu / rho (and mask) aligned for 128-bit access
  int pad;
  float* mu_w;               // where results are written: mu / rho themselves (in place) or the output tensors
  float* rho_w;
  uint32_t* list2;           // bnn_prune_into: the exact-select list [index | mu | rho] x cap2, then (key, index) pairs x cap2
  uint32_t cap2;
  uint32_t pad2;
  double* kl_out;            // bnn_prune_into, optional: the sweep also accumulates the tensor's KL element sum here
  float kl_loc, kl_inv_scale, kl_log_scale;
  uint32_t pad3;
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

// The same search over a histogram that is spread over the CTAs of a cluster (the sample kernel): every CTA holds a
// partial histogram `hist` (kBins) and its 32 group sums `grp` (64 bins each); warp 0 of ONE CTA walks them through
// distributed shared memory — 8 remote loads per lane for the group level, 16 for the two bins of the group level —
// instead of pulling all kBins x kCtas counters.  Bins are walked downwards (descending rank).
template <int kCtas>
__device__ __forceinline__ void cluster_find_bin(cg::cluster_group& cluster, uint32_t* hist, uint32_t* grp, uint64_t rem,
                                                 uint32_t* bin_out, uint64_t* before_out) {
  const int lane = threadIdx.x & 31;
  const int pos = 31 - lane;
  uint64_t gs = 0;
#pragma unroll
  for (int q = 0; q < kCtas; ++q) gs += cluster.map_shared_rank(grp, q)[pos];
  uint64_t incl = gs;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint64_t v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  const unsigned int hit = __ballot_sync(0xffffffffu, incl > rem);
  const int g = hit ? __ffs(hit) - 1 : 31;
  const uint64_t before_g = __shfl_sync(0xffffffffu, incl - gs, g);
  const int gpos = 31 - g;
  const int b0 = gpos * 64 + 63 - 2 * lane, b1 = b0 - 1;
  uint64_t c0 = 0, c1 = 0;
#pragma unroll
  for (int q = 0; q < kCtas; ++q) {
    const uint32_t* h = cluster.map_shared_rank(hist, q);
    c0 += h[b0];
    c1 += h[b1];
  }
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

// key2 = (key + log sqrt(2 pi)) * log2(e): the domain of the certified intervals and of the grid
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kDelta2 = 4e-5f * 1.4426950408889634f;      // absolute part of the key margin, in key2 units

// Certified interval of one element's key in grid coordinates.  With Q = (mu / sigma)^2 = 2q and L = log2(sigma):
// key2 = -Q * log2e / 2 - L.  The fast value is within 2e-5 * q + 4e-5 (key units) of the exact fp32 key that torch's
// op order produces: sigma carries <= ~1.5e-6 relative error (ex2.approx and its argument rounding, the degree-4
// minimax polynomial of log1p(e)/e on [0, 1/4] — 2.8e-7 in fp32 Horner form —, lg2.approx for e > 1/4), rcp.approx 1 ulp, lg2.approx
// <= 2^-22 relative (8e-6 at sigma = 1e-10); the exact key itself rounds within a few ulp.  So
//   y_minus = ((-Q * 0.50001 * log2e - L) - delta - g_lo) * scale  <=  y(exact key)  <=
//   y_plus  = ((-Q * 0.49999 * log2e - L) + delta - g_lo) * scale
// Both sweeps call THIS function: only explicitly rounded operations and MUFU instructions, so the two sweeps compute
// bit-identical intervals (the histograms of sweep 1 predict sweep 2 exactly).  Every step is monotone, NaN stays NaN
// (such elements are always deferred).  3 MUFU + 15 FP32 operations per element when every rho <= ln(1/4).
struct Grid { float sA, sB, nscale, offs_minus, offs_plus; };
__device__ __forceinline__ Grid make_grid(const PruneState* st) {
  Grid g;
  const float scale = st->scale;
  g.sA = __fmul_rn(-0.50001f * kLog2e, scale);
  g.sB = __fmul_rn(-0.49999f * kLog2e, scale);
  g.nscale = -scale;
  g.offs_minus = st->offs_minus;
  g.offs_plus = st->offs_plus;
  return g;
}
template <bool kAnyRho>
__device__ __forceinline__ void key_interval(float mu, float rho, const Grid& g, float& y_minus, float& y_plus,
                                             float* sigma_out = nullptr, float* log2_sigma_out = nullptr) {
  const float e = ex2_ftz(__fmul_rn(rho, kLog2e));       // flushed to 0 below rho ~ -87: sigma = 1e-10, as in fp32 torch
  float p = __fmaf_rn(e, 0.1237151250243187f, -0.23501642048358917f);
  p = __fmaf_rn(e, p, 0.3320590555667877f);
  p = __fmaf_rn(e, p, -0.4999610483646393f);
  p = __fmaf_rn(e, p, 0.9999998211860657f);
  float sigma = __fmaf_rn(e, p, 1e-10f);                 // >= 1e-10: normal range for rcp / lg2
  if (kAnyRho) {
    // e > 1/4: ln2 * lg2(1 + e); rho > 15: softplus = rho + exp(-rho) to 1e-13 (torch returns rho itself above 20)
    const float big = __fmaf_rn(lg2_ftz(__fadd_rn(1.0f, e)), 0.6931471805599453f, 1e-10f);
    const float huge = __fadd_rn(rho, ex2_ftz(__fmul_rn(rho, -kLog2e)));
    sigma = rho > 15.0f ? huge : (e > 0.25f ? big : sigma);
  }
  const float t = __fmul_rn(mu, rcp_ftz(sigma));
  const float Q = __fmul_rn(t, t);
  const float L = lg2_ftz(sigma);
  y_minus = __fmaf_rn(Q, g.sA, __fmaf_rn(L, g.nscale, g.offs_minus));
  y_plus = __fmaf_rn(Q, g.sB, __fmaf_rn(L, g.nscale, g.offs_plus));
  if (sigma_out != nullptr) { *sigma_out = sigma; *log2_sigma_out = L; }      // by-products for the fused KL sum
}

// KL(N(mu, sigma) || N(loc, scale)) of one element from the interval arithmetic's by-products (reference loss.py:16-38,
// torch/distributions/kl.py:468-471): 0.5 ((sigma / scale)^2 + ((mu - loc) / scale)^2 - 1) - ln(sigma / scale).  sigma
// carries <= 1.5e-6 relative error here, lg2.approx <= 2^-22: the sum agrees with bnn_kl to a few 1e-6 relative.
struct KlPrior { float loc, inv_scale, log_scale; };
__device__ __forceinline__ float kl_from_sigma(float mu, float sigma, float log2_sigma, const KlPrior& pr) {
  const float r = sigma * pr.inv_scale;
  const float dd = (mu - pr.loc) * pr.inv_scale;
  return fmaf(0.5f, fmaf(r, r, fmaf(dd, dd, -1.0f)), pr.log_scale) - log2_sigma * 0.6931471805599453f;
}
// The same sum over a run of elements with the constants factored out — three instructions per element in the sweep:
//   sum KL = 0.5 / scale^2 * sum(sigma^2 + (mu - loc)^2) - ln 2 * sum(log2 sigma) + n (ln scale - 0.5)
struct KlRun {
  float sq = 0.f, lg = 0.f, n = 0.f;
  __device__ __forceinline__ void add(float mu, float sigma, float log2_sigma, const KlPrior& pr) {
    const float dd = mu - pr.loc;
    sq = fmaf(sigma, sigma, fmaf(dd, dd, sq));
    lg += log2_sigma;
    n += 1.0f;
  }
  __device__ __forceinline__ float total(const KlPrior& pr) const {
    return fmaf(0.5f * pr.inv_scale * pr.inv_scale, sq, fmaf(-0.6931471805599453f, lg, n * (pr.log_scale - 0.5f)));
  }
};

// The sample only PROPOSES the grid (sweep 1 and the bracket step prove or reject it), so its keys are the fast ones.
__device__ __forceinline__ float key2_fast(float mu, float rho) {
  const Grid g = {-0.5f * kLog2e, -0.5f * kLog2e, -1.0f, 0.0f, 0.0f};
  float lo, hi;
  key_interval<true>(mu, rho, g, lo, hi);
  return lo;
}

// 1. the sample -> the grid.  One CLUSTER of kSampleCtas CTAs per tensor: every thread gathers its share of the strided
// sample (kSample single elements, or kSampleBig elements in 2048 runs of 128 for large tensors), turns it into ordered
// fast keys and KEEPS THEM IN REGISTERS for the three passes (min / max, an 11-bit histogram of the key range, an 11-bit
// sub-histogram of the two bins holding the bracketing order statistics).  Each pass bins into the CTA's own shared
// memory; CTA 0 sums the eight histograms through distributed shared memory, locates the bins and broadcasts them.  The
// earlier form — a machine-wide gather kernel that wrote the keys to the workspace plus ONE block per tensor that re-read
// them three times — cost 19 + 154 us per group of 16 tensors; this one is bounded by the gather latency.
static_assert((kSample & (kSample - 1)) == 0 && (kSampleBig & (kSampleBig - 1)) == 0, "sample sizes must be powers of two");
constexpr int kSampleCtas = 8;
constexpr int kSamplePerThread = kSampleBig / (kSampleCtas * kResolveThreads);      // 32 keys per thread (4 for kSample)
static_assert(kSamplePerThread * kSampleCtas * kResolveThreads == kSampleBig && kSample % (kSampleCtas * kResolveThreads) == 0,
              "the sample must split evenly over the cluster");

__global__ void __cluster_dims__(kSampleCtas, 1, 1) __launch_bounds__(kResolveThreads)
prune_sample_kernel(const __grid_constant__ PruneTable tab) {
  __shared__ uint32_t s_hist[kBins], s_hist2[kBins];
  __shared__ uint32_t s_grp[32], s_grp2[32];      // sums over groups of 64 bins of s_hist / s_hist2
  __shared__ uint32_t s_red[2 * (kResolveThreads / 32)];
  __shared__ uint32_t s_mm[2];          // this CTA's min / max (read by the whole cluster)
  __shared__ uint32_t s_sel[6];         // CTA 0: bins / counts of the two order statistics (read by the whole cluster)
  __shared__ uint32_t s_loc[6];
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned int rank = cluster.block_rank();
  const PruneDesc& d = tab.t[blockIdx.x / kSampleCtas];
  PruneState st;
  st.prefix = 0; st.need_ranks = 0; st.k_rem = d.k; st.eq_total = 0;
  st.scale = 0.f; st.offs_minus = 0.f; st.offs_plus = 0.f;
  st.count_above = 0ull; st.a_thr = INFINITY; st.b_thr = -INFINITY;
  st.defer_all = 0; st.n_take = 0; st.expect_deferred = 0; st.n_deferred = 0; st.n_deferred2 = 0;
  const bool trivial = d.k <= 0 || d.k >= d.numel;
  st.general = (d.force_general || d.numel >= (int64_t(1) << 32)) ? 1u : 0u;
  if (!trivial && !st.general && d.numel <= kSmallTensor) {
    st.defer_all = 1;
    st.n_take = static_cast<uint32_t>(d.k);
    st.expect_deferred = static_cast<uint32_t>(d.numel);
  }
  if (rank == 0 && threadIdx.x == 0 && d.kl_out != nullptr) *d.kl_out = 0.0;      // the sweep accumulates into it
  if (trivial || st.general || d.numel <= kSmallTensor) {      // uniform over the cluster: nobody reaches a cluster.sync
    if (rank == 0 && threadIdx.x == 0) {
      *d.state = st;
      if (st.general) atomicOr(tab.any_general, 1u);
    }
    return;
  }
  const int m = sample_size(d.numel);
  const int per = m / (kSampleCtas * kResolveThreads);        // 4 or 32
  const int cta_base = static_cast<int>(rank) * (m / kSampleCtas);
  uint32_t keys[kSamplePerThread];
#pragma unroll
  for (int u0 = 0; u0 < kSamplePerThread; u0 += 8) {
    float mv[8], rv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      mv[u] = 0.f; rv[u] = 0.f;
      if (u0 + u < per) {
        const uint64_t j = static_cast<uint64_t>(cta_base + (u0 + u) * kResolveThreads + static_cast<int>(threadIdx.x));
        int64_t i;
        if (m == kSample) {
          i = static_cast<int64_t>((j * static_cast<uint64_t>(d.numel)) / kSample);      // a shift; numel < 2^32
        } else {                                             // run j / 128 starts at run * numel / 2048 (4-element aligned)
          const uint64_t run = j / kSampleRun, within = j % kSampleRun;
          i = static_cast<int64_t>(((run * static_cast<uint64_t>(d.numel)) / (kSampleBig / kSampleRun)) & ~uint64_t(3)) +
              static_cast<int64_t>(within);
        }
        mv[u] = __ldg(d.mu + i);
        rv[u] = __ldg(d.rho + i);
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) keys[u0 + u] = order_key(key2_fast(mv[u], rv[u]));
  }
  // min / max over the cluster
  uint32_t mn = 0xffffffffu, mx = 0u;
#pragma unroll
  for (int u = 0; u < kSamplePerThread; ++u)
    if (u < per) { mn = min(mn, keys[u]); mx = max(mx, keys[u]); }
  for (int o = 16; o > 0; o >>= 1) {
    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) { s_red[threadIdx.x >> 5] = mn; s_red[kResolveThreads / 32 + (threadIdx.x >> 5)] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t a = 0xffffffffu, b = 0u;
    for (int w = 0; w < kResolveThreads / 32; ++w) { a = min(a, s_red[w]); b = max(b, s_red[kResolveThreads / 32 + w]); }
    s_mm[0] = a; s_mm[1] = b;
  }
  cluster.sync();
  if (threadIdx.x == 0) {
    uint32_t a = 0xffffffffu, b = 0u;
    for (int r = 0; r < kSampleCtas; ++r) {
      const uint32_t* remote = cluster.map_shared_rank(s_mm, r);
      a = min(a, remote[0]); b = max(b, remote[1]);
    }
    s_loc[0] = a; s_loc[1] = b;
  }
  __syncthreads();
  mn = s_loc[0]; mx = s_loc[1];
  const double p = static_cast<double>(d.k) / static_cast<double>(d.numel);
  const int r = static_cast<int>(p * m);                              // descending rank of the k-th key
  const int margin = static_cast<int>(6.0 * sqrt(m * p * (1.0 - p))) + 8;
  const int r_hi = r - margin, r_lo = r + margin;
  // both order statistics in two histogram passes (11 + 11 bits of the key range; below that the bracket ends are
  // sub-bin edges, which only widens the bracket by < 2^-22 of the range)
  const uint32_t range_u = mx - mn;
  const int bits = range_u == 0u ? 0 : 32 - __clz(range_u);
  const int s0 = max(bits - 11, 0), s1 = max(bits - 22, 0);
  for (int b = threadIdx.x; b < kBins; b += blockDim.x) { s_hist[b] = 0; s_hist2[b] = 0; }
  __syncthreads();
#pragma unroll
  for (int u = 0; u < kSamplePerThread; ++u)
    if (u < per) atomicAdd(&s_hist[(keys[u] - mn) >> s0], 1u);
  __syncthreads();
  {
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;      // 32 warps: warp w sums the 64 bins of group w
    const uint32_t v = warp_sum(s_hist[64 * w + l] + s_hist[64 * w + 32 + l]);
    if (l == 0) s_grp[w] = v;
  }
  cluster.sync();
  if (rank == 0 && threadIdx.x < 32) {
    uint32_t bin_a = 0, bin_b = 0;
    uint64_t before_a = 0, before_b = 0;
    if (r_hi >= 0) cluster_find_bin<kSampleCtas>(cluster, s_hist, s_grp, static_cast<uint64_t>(r_hi), &bin_a, &before_a);
    if (r_lo < m) cluster_find_bin<kSampleCtas>(cluster, s_hist, s_grp, static_cast<uint64_t>(r_lo), &bin_b, &before_b);
    if (threadIdx.x == 0) {
      s_sel[0] = bin_a; s_sel[1] = static_cast<uint32_t>(before_a);
      s_sel[2] = bin_b; s_sel[3] = static_cast<uint32_t>(before_b);
    }
  }
  cluster.sync();              // CTA 0 has read every histogram and published the bins
  if (threadIdx.x < 4) s_loc[2 + threadIdx.x] = cluster.map_shared_rank(s_sel, 0)[threadIdx.x];
  __syncthreads();
  const uint32_t bin_a = s_loc[2], before_a = s_loc[3], bin_b = s_loc[4], before_b = s_loc[5];
  uint32_t sub_a = 0, sub_b = 0;
  if (s0 > 0) {                // uniform over the cluster (a function of the common min / max)
    for (int b = threadIdx.x; b < kBins; b += blockDim.x) { s_hist[b] = 0; s_hist2[b] = 0; }
    __syncthreads();
    const uint32_t sub_mask = (1u << (s0 - s1)) - 1u;
#pragma unroll
    for (int u = 0; u < kSamplePerThread; ++u) {
      if (u < per) {
        const uint32_t v = keys[u] - mn, dgt = v >> s0, sub = (v >> s1) & sub_mask;
        if (dgt == bin_a) atomicAdd(&s_hist[sub], 1u);
        if (dgt == bin_b) atomicAdd(&s_hist2[sub], 1u);
      }
    }
    __syncthreads();
    {
      const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
      const uint32_t va = warp_sum(s_hist[64 * w + l] + s_hist[64 * w + 32 + l]);
      const uint32_t vb = warp_sum(s_hist2[64 * w + l] + s_hist2[64 * w + 32 + l]);
      if (l == 0) { s_grp[w] = va; s_grp2[w] = vb; }
    }
    cluster.sync();
    if (rank == 0 && threadIdx.x < 32) {
      uint32_t a = 0, b = 0;
      uint64_t before = 0;
      if (r_hi >= 0) cluster_find_bin<kSampleCtas>(cluster, s_hist, s_grp, static_cast<uint64_t>(r_hi) - before_a, &a, &before);
      if (r_lo < m) cluster_find_bin<kSampleCtas>(cluster, s_hist2, s_grp2, static_cast<uint64_t>(r_lo) - before_b, &b, &before);
      sub_a = a; sub_b = b;          // used by thread 0 below
    }
  }
  if (rank == 0 && threadIdx.x == 0) {
    // bracket ends; a rank outside the sample leaves that side at the sample extreme (the thin tail beyond it simply
    // lands in the outermost bin or outside the grid — both are handled exactly by the bracket step)
    uint32_t hi = mx, lo = mn;
    if (r_hi >= 0) {
      const uint64_t top = static_cast<uint64_t>(mn) + ((static_cast<uint64_t>(bin_a) << s0) | (static_cast<uint64_t>(sub_a) << s1)) +
                           ((uint64_t(1) << s1) - 1);
      hi = top < mx ? static_cast<uint32_t>(top) : mx;
    }
    if (r_lo < m) lo = mn + ((bin_b << s0) | (sub_b << s1));
    const float g_lo = unorder_key(lo), g_hi = unorder_key(hi);      // the sampled keys are key2 values already
    const float range = fmaxf(g_hi - g_lo, 2e-4f + 1e-5f * fabsf(g_lo));      // never degenerate
    if (g_lo == g_lo && range < INFINITY) {                                   // no NaN / inf among the bracket keys
      st.scale = static_cast<float>(kBins) / range;
      st.offs_minus = -(g_lo + kDelta2) * st.scale;
      st.offs_plus = -(g_lo - kDelta2) * st.scale;
    } else {
      st.general = 1u;
    }
    *d.state = st;
    if (st.general) atomicOr(tab.any_general, 1u);
  }
  cluster.sync();              // no CTA exits while CTA 0 may still read its shared memory
}

// self test: the interval on the identity grid (scale 1, origin 0), i.e. in key2 units
template <bool kAnyRho>
__global__ void prune_interval_selftest_kernel(const float* mu, const float* rho, int64_t n, float* lo_out, float* hi_out) {
  const Grid g = {-0.50001f * kLog2e, -0.49999f * kLog2e, -1.0f, -kDelta2, kDelta2};
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float a, b;
    key_interval<kAnyRho>(mu[i], rho[i], g, a, b);
    lo_out[i] = a;
    hi_out[i] = b;
  }
}

// 2. sweep 1 (read only): count the elements certified above the grid, histogram the interval ends of the others.
// Warps are independent (a warp owns one 512-element unit per chunk): no block-level synchronisation in the loop.
__global__ void __launch_bounds__(kThreads) prune_bin_kernel(const __grid_constant__ PruneTable tab) {
  __shared__ int64_t s_begin[kMaxTensors];
  __shared__ float2 s_queue[(4 * kVecPerThread + 1) * kThreads];      // per lane: 16 slots + 1 for the trailing store
  if (threadIdx.x < tab.n) s_begin[threadIdx.x] = tab.t[threadIdx.x].chunk_begin;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int cur = -1;
  bool active = false;
  Grid g = {0.f, 0.f, 0.f, 0.f, 0.f};
  unsigned int above = 0;
  auto flush = [&]() {      // warp-wide: adds this warp's count of "above" elements of tensor `cur`
    const unsigned int v = warp_sum(above);
    if (lane == 0 && v != 0u) atomicAdd(&tab.t[cur].state->count_above, static_cast<unsigned long long>(v));
    above = 0;
  };
  for (int64_t chunk = blockIdx.x; chunk < tab.total_chunks; chunk += gridDim.x) {
    const int t = find_tensor(s_begin, tab.n, chunk, cur < 0 ? 0 : cur);
    if (t != cur) {
      if (cur >= 0 && active) flush();
      cur = t;
      const PruneDesc& d0 = tab.t[t];
      const PruneState* st = d0.state;
      active = !st->general && !st->defer_all && d0.k > 0 && d0.k < d0.numel;
      g = make_grid(st);
    }
    if (!active) continue;
    const PruneDesc& d = tab.t[t];
    const int64_t ubase = (chunk - d.chunk_begin) * kChunk + warp * kUnit;
    if (ubase >= d.numel) continue;        // warp-uniform
    // ~3 % of the elements overlap the grid, but in almost every warp-wide slot SOME lane does: a branch per element
    // would be entered all the time.  Instead every element stores its interval UNCONDITIONALLY into the lane's own
    // shared-memory queue and only the queue position advances when it overlaps (a following store overwrites a
    // skipped one); the histogram updates run afterwards over the few queued intervals.
    float2* const q_lane = s_queue + threadIdx.x;          // slot s of this lane: q_lane[s * kThreads]
    const uint32_t q_base = static_cast<uint32_t>(__cvta_generic_to_shared(q_lane));
    uint32_t q_addr = q_base;                              // shared-memory address of the lane's next slot
    // five instructions per element (two compares, the store, two predicated adds) — written as PTX because the
    // compiler turns the conditional increments into add + select pairs:
    //   below2048 = !(y_minus >= 2048) [or NaN];  overlap = below2048 && !(y_plus < 0) [or NaN]
    auto account = [&](float y_minus, float y_plus) {
      asm volatile(
          "{\n\t.reg .pred p1, p2;\n\t"
          "st.shared.v2.f32 [%0], {%2, %3};\n\t"
          "setp.ltu.f32 p1, %2, 2048.0;\n\t"
          "setp.geu.and.f32 p2, %3, 0.0, p1;\n\t"
          "@!p1 add.u32 %1, %1, 1;\n\t"
          "@p2 add.u32 %0, %0, %4;\n\t}"
          : "+r"(q_addr), "+r"(above)
          : "f"(y_minus), "f"(y_plus), "n"(kThreads * 8)
          : "memory");
    };
    if (d.vec && ubase + kUnit <= d.numel) {
      float4 m[kVecPerThread], r[kVecPerThread];
#pragma unroll
      for (int j = 0; j < kVecPerThread; ++j) {
        const int64_t i = ubase + (j * 32 + lane) * 4;
        m[j] = ldg_stream4(d.mu + i);
        r[j] = ldg_stream4(d.rho + i);
      }
#pragma unroll
      for (int j = 0; j < kVecPerThread; ++j) {
        const float mm[4] = {m[j].x, m[j].y, m[j].z, m[j].w};
        const float rr[4] = {r[j].x, r[j].y, r[j].z, r[j].w};
        float ym[4], yp[4];
        if (fmaxf(fmaxf(rr[0], rr[1]), fmaxf(rr[2], rr[3])) <= -1.3862944f) {      // one branch per 4 elements
#pragma unroll
          for (int q = 0; q < 4; ++q) key_interval<false>(mm[q], rr[q], g, ym[q], yp[q]);
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) key_interval<true>(mm[q], rr[q], g, ym[q], yp[q]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) account(ym[q], yp[q]);
      }
    } else {
#pragma unroll 1
      for (int j = 0; j < 4 * kVecPerThread; ++j) {
        const int64_t i = ubase + j * 32 + lane;
        float ym = -INFINITY, yp = -INFINITY;          // past the end: below the grid, no effect
        if (i < d.numel) key_interval<true>(d.mu[i], d.rho[i], g, ym, yp);
        account(ym, yp);
      }
    }
    const int pos = static_cast<int>((q_addr - q_base) / (kThreads * 8));
    for (int s = 0; s < pos; ++s) {                    // rarely more than two rounds
      // lower end: index 0 = below the grid (or NaN), index b + 1 = bin b; upper end: index b = bin b, index 2048 =
      // above the grid (or NaN).  No clamping INTO the grid: an interval that sticks out may hold a key outside it.
      const float2 y = q_lane[s * kThreads];
      const int im = !(y.x >= 0.0f) ? 0 : __float2int_rd(y.x) + 1;                            // y.x < 2048 in the queue
      const int ip = !(y.y < static_cast<float>(kBins)) ? kBins : __float2int_rd(y.y);        // y.y >= 0 in the queue
      atomicAdd(d.hist + im, 1u);
      atomicAdd(d.hist_plus + ip, 1u);
    }
  }
  if (cur >= 0 && active) flush();
}

// 3. one block per tensor: the bins that provably enclose the k-th key.  Bins are numbered -1 (below the grid),
// 0 .. 2047, 2048 (above the grid); every element has a lower-end bin bm <= bin(exact key) <= bp, its upper-end bin
// (elements counted as "above" in sweep 1: bm = bp = 2048; the uncounted rest below the grid: bm = bp = -1; NaN
// intervals: bm = -1, bp = 2048).  With need = k - |above|:
//   j_hi = smallest j with |{binned: bp > j}| < need: fewer than k elements can have a key in a bin above j_hi, so the
//          k-th key is in a bin <= j_hi (j_hi = 2048 when even the grid's top cannot be excluded);
//   j_lo = largest j with |{binned: bm >= j}| >= need: at least k elements have a key in a bin >= j_lo, so the k-th key
//          is in a bin >= j_lo (j_lo = -1 when the grid's bottom cannot be excluded).
// Elements with bm > j_hi are in the top k, elements with bp < j_lo are not; the others (counted exactly from the
// histograms) are resolved by exact keys in the finish kernel.
__global__ void __launch_bounds__(kThreads) prune_bracket_kernel(const __grid_constant__ PruneTable tab) {
  __shared__ uint32_t s_minus[kHistStride], s_plus[kHistStride];     // index conventions of the bin kernel
  __shared__ uint32_t s_out[2];
  __shared__ unsigned long long s_cnt[4];
  const PruneDesc& d = tab.t[blockIdx.x];
  for (int b = threadIdx.x; b < kHistStride; b += blockDim.x) {
    s_minus[b] = d.hist[b];
    s_plus[b] = d.hist_plus[b];
    d.hist[b] = 0;                                   // clean for the general path / the next call
    d.hist_plus[b] = 0;
  }
  if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0ull;
  __syncthreads();
  PruneState st = *d.state;
  if (st.general || st.defer_all || d.k <= 0 || d.k >= d.numel) return;
  // totals: all binned elements, and those whose lower end is inside the grid (bm >= 0)
  unsigned long long part = 0, part_in = 0;
  for (int b = threadIdx.x; b <= kBins; b += blockDim.x) {
    part += s_plus[b];
    if (b >= 1) part_in += s_minus[b];
  }
  part = warp_sum(part);
  part_in = warp_sum(part_in);
  if ((threadIdx.x & 31) == 0) {
    if (part) atomicAdd(&s_cnt[0], part);
    if (part_in) atomicAdd(&s_cnt[1], part_in);
  }
  __syncthreads();
  const uint64_t k = static_cast<uint64_t>(d.k), n_binned = s_cnt[0], n_lower_in = s_cnt[1];
  const uint64_t n_above = st.count_above, n_below = static_cast<uint64_t>(d.numel) - n_above - n_binned;
  bool ok = n_above < k && k <= n_above + n_binned;
  if (ok) {
    const uint64_t need = k - n_above;
    const uint64_t top_out = s_plus[kBins];                        // upper end above the grid
    if (threadIdx.x < 32) {
      uint32_t j_hi = kBins, j_lo = 0;
      uint64_t before;
      if (top_out < need) warp_find_bin(s_plus, true, need - 1 - top_out, &j_hi, &before);       // bins 0 .. 2047
      if (n_lower_in >= need) warp_find_bin(s_minus + 1, true, need - 1, &j_lo, &before);        // bins 0 .. 2047
      if (threadIdx.x == 0) { s_out[0] = j_hi; s_out[1] = j_lo; }
    }
    __syncthreads();
    const int j_hi = static_cast<int>(s_out[0]);                          // 2048: the grid's top is not excluded
    const int j_lo = n_lower_in >= need ? static_cast<int>(s_out[1]) : -1;
    // |binned: bm > j_hi| (lower-end index >= j_hi + 2) and |binned: bp >= j_lo|
    unsigned long long a = 0, c = 0;
    for (int b = threadIdx.x; b <= kBins; b += blockDim.x) {
      if (b >= j_hi + 2) a += s_minus[b];
      if (b >= j_lo) c += s_plus[b];
    }
    a = warp_sum(a);
    c = warp_sum(c);
    if ((threadIdx.x & 31) == 0) {
      if (a) atomicAdd(&s_cnt[2], a);
      if (c) atomicAdd(&s_cnt[3], c);
    }
    __syncthreads();
    const bool top_open = j_hi >= kBins, bottom_open = j_lo < 0;
    const uint64_t certain = top_open ? 0 : n_above + s_cnt[2];            // pruned by sweep 2 without an exact key
    const uint64_t deferred = s_cnt[3] - s_cnt[2] + (top_open ? n_above : 0) + (bottom_open ? n_below : 0);
    ok = j_lo <= j_hi && certain < k && deferred <= (d.list2 != nullptr ? d.cap2 : d.defer_cap) && k - certain <= deferred;
    // bnn_prune_into has already committed to everything outside the grid: the bracket must lie inside it, and the list
    // of in-grid elements must be complete
    if (d.list2 != nullptr && (top_open || bottom_open || st.n_deferred > d.defer_cap)) ok = false;
    if (ok) {
      st.a_thr = top_open ? INFINITY : static_cast<float>(j_hi + 1);
      st.b_thr = bottom_open ? -INFINITY : static_cast<float>(j_lo);
      st.n_take = static_cast<uint32_t>(k - certain);
      st.expect_deferred = static_cast<uint32_t>(deferred);
    }
  }
  if (!ok) st.general = 1u;
  if (threadIdx.x == 0) {
    *d.state = st;
    if (st.general) atomicOr(tab.any_general, 1u);
  }
}

// 4. sweep 2: prune what is certified above the bracket, defer what overlaps it
__global__ void __launch_bounds__(kThreads, 4) prune_apply_sampled_kernel(const __grid_constant__ PruneTable tab) {
  __shared__ int64_t s_begin[kMaxTensors];
  if (threadIdx.x < tab.n) s_begin[threadIdx.x] = tab.t[threadIdx.x].chunk_begin;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int cur = -1;
  int mode = 0;            // 0 skip tensor, 1 none, 2 all, 3 select
  Grid g = {0.f, 0.f, 0.f, 0.f, 0.f};
  float a_thr = 0.f, b_thr = 0.f;
  for (int64_t chunk = blockIdx.x; chunk < tab.total_chunks; chunk += gridDim.x) {
    const int t = find_tensor(s_begin, tab.n, chunk, cur < 0 ? 0 : cur);
    if (t != cur) {
      cur = t;
      const PruneDesc& d0 = tab.t[t];
      const PruneState* st = d0.state;
      mode = st->general ? 0 : (d0.k <= 0 ? 1 : (d0.k >= d0.numel ? 2 : 3));
      g = make_grid(st);
      a_thr = st->a_thr; b_thr = st->b_thr;          // defer_all: +inf / -inf, every element is deferred
    }
    if (mode == 0) continue;
    const PruneDesc& d = tab.t[t];
    const int64_t ubase = (chunk - d.chunk_begin) * kChunk + warp * kUnit;
    if (ubase >= d.numel) continue;        // warp-uniform
    // deferred elements are rare: one warp-aggregated reservation per occurrence
    auto defer = [&](bool mine, int64_t i, float mu, float rho) {
      const unsigned int bal = __ballot_sync(0xffffffffu, mine);
      if (bal == 0u) return;
      const int leader = __ffs(bal) - 1;
      unsigned int base = 0;
      if (lane == leader) base = atomicAdd(&d.state->n_deferred, static_cast<unsigned int>(__popc(bal)));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (mine) {
        const unsigned int slot = base + __popc(bal & ((1u << lane) - 1u));
        if (slot < d.defer_cap) {
          d.keys[slot] = static_cast<uint32_t>(i);            // numel < 2^32 on this path
          d.keys[d.defer_cap + slot] = __float_as_uint(mu);
          d.keys[2u * d.defer_cap + slot] = __float_as_uint(rho);
        }
      }
    };
    if (d.vec && ubase + kUnit <= d.numel) {
      if (mode != 3) {
#pragma unroll
        for (int j = 0; j < kVecPerThread; ++j) {
          const int64_t i = ubase + (j * 32 + lane) * 4;
          if (mode == 2) {
            *reinterpret_cast<float4*>(d.mu + i) = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(d.rho + i) = make_float4(-30.f, -30.f, -30.f, -30.f);
          }
          if (d.mask != nullptr) {
            const unsigned char v = mode == 2 ? 1 : 0;
            *reinterpret_cast<uchar4*>(d.mask + i) = make_uchar4(v, v, v, v);
          }
        }
        continue;
      }
      float4 m[kVecPerThread], r[kVecPerThread];
#pragma unroll
      for (int j = 0; j < kVecPerThread; ++j) {
        const int64_t i = ubase + (j * 32 + lane) * 4;
        m[j] = ldg_stream4(d.mu + i);
        r[j] = ldg_stream4(d.rho + i);
      }
#pragma unroll
      for (int j = 0; j < kVecPerThread; ++j) {
        const int64_t i = ubase + (j * 32 + lane) * 4;
        const float mm[4] = {m[j].x, m[j].y, m[j].z, m[j].w};
        const float rr[4] = {r[j].x, r[j].y, r[j].z, r[j].w};
        float ym[4], yp[4];
        if (fmaxf(fmaxf(rr[0], rr[1]), fmaxf(rr[2], rr[3])) <= -1.3862944f) {      // one branch per 4 elements
#pragma unroll
          for (int q = 0; q < 4; ++q) key_interval<false>(mm[q], rr[q], g, ym[q], yp[q]);
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) key_interval<true>(mm[q], rr[q], g, ym[q], yp[q]);
        }
        bool tk[4], df[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          tk[q] = ym[q] >= a_thr;
          df[q] = !tk[q] && !(yp[q] < b_thr);
        }
        if (__any_sync(0xffffffffu, df[0] || df[1] || df[2] || df[3])) {
#pragma unroll
          for (int q = 0; q < 4; ++q) defer(df[q], i + q, mm[q], rr[q]);
        }
        if (tk[0] || tk[1] || tk[2] || tk[3]) {      // the old values are in registers: whole-vector stores
          *reinterpret_cast<float4*>(d.mu + i) =
              make_float4(tk[0] ? 0.f : mm[0], tk[1] ? 0.f : mm[1], tk[2] ? 0.f : mm[2], tk[3] ? 0.f : mm[3]);
          *reinterpret_cast<float4*>(d.rho + i) = make_float4(tk[0] ? -30.f : rr[0], tk[1] ? -30.f : rr[1],
                                                              tk[2] ? -30.f : rr[2], tk[3] ? -30.f : rr[3]);
        }
        if (d.mask != nullptr)
          *reinterpret_cast<uchar4*>(d.mask + i) = make_uchar4(tk[0], tk[1], tk[2], tk[3]);
      }
    } else {
#pragma unroll 1
      for (int j = 0; j < 4 * kVecPerThread; ++j) {       // warp-uniform trip count (defer() votes)
        const int64_t i = ubase + j * 32 + lane;
        const bool in = i < d.numel;
        bool take = mode == 2, df = false;
        float mu = 0.f, rho = 0.f;
        if (in && mode == 3) {
          mu = d.mu[i]; rho = d.rho[i];
          float ym, yp;
          key_interval<true>(mu, rho, g, ym, yp);
          take = ym >= a_thr;
          df = !take && !(yp < b_thr);
        }
        defer(df, i, mu, rho);
        if (in) {
          if (take) { d.mu[i] = 0.0f; d.rho[i] = -30.0f; }
          if (d.mask != nullptr) d.mask[i] = take ? 1 : 0;
        }
      }
    }
  }
}

// ---- bnn_prune_into: the single out-of-place sweep.  Per element: interval in grid coordinates; above the grid ->
// (0, -30) in the output and counted; inside the grid -> copied, both histograms updated, (index, mu, rho) listed; below ->
// copied.  Same warp-private queue as the bin kernel for the histogram updates; the list append is one reservation per
// warp and 512-element unit (a warp-wide prefix sum of the lanes' counts), not one per element.
template <bool kKl>
__global__ void __launch_bounds__(kThreads) prune_sweep_into_kernel(const __grid_constant__ PruneTable tab) {
  __shared__ int64_t s_begin[kMaxTensors];
  __shared__ float2 s_queue[(4 * kVecPerThread + 1) * kThreads];
  // per (chunk of this block, warp): elements certified above the grid and, kKl, the KL partial sum — combined over the
  // warps at the end of the block: one atomic per (block, chunk) instead of one per (warp, chunk)
  __shared__ unsigned int s_above[kSweepChunks * (kThreads / 32)];
  __shared__ double s_kl[kSweepChunks * (kThreads / 32)];
  if (threadIdx.x < tab.n) s_begin[threadIdx.x] = tab.t[threadIdx.x].chunk_begin;
  if (threadIdx.x < kSweepChunks * (kThreads / 32)) { s_above[threadIdx.x] = 0u; s_kl[threadIdx.x] = 0.0; }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int cur = -1;
  int mode = 0;            // 0 general (left to the fallback), 1 copy, 2 all pruned, 3 select, 4 defer all (small tensor)
  Grid g = {0.f, 0.f, 0.f, 0.f, 0.f};
  unsigned int above = 0;
  double kl_acc = 0.0;     // kKl: this thread's share of the current chunk's KL element sum
  bool want_kl = false;
  KlPrior prior = {0.f, 0.f, 0.f};
  const Grid g_unit = {-0.5f * kLog2e, -0.5f * kLog2e, -1.0f, 0.0f, 0.0f};      // any grid: only sigma / log2 sigma are used
  // chunk it of block b = b + it * gridDim.x: at any moment the resident blocks work on one compact window of addresses
  // (walking consecutive chunks per block instead — each block its own 128 KiB stream — measured 8 % slower)
  for (int it = 0; it < kSweepChunks; ++it) {
    const int64_t chunk = blockIdx.x + static_cast<int64_t>(it) * gridDim.x;
    if (chunk >= tab.total_chunks) break;
    if (it > 0) {          // publish the previous chunk's counts
      const unsigned int va = warp_sum(above);
      if (lane == 0) s_above[(it - 1) * (kThreads / 32) + warp] = va;
      if (kKl) {
        const double vk = warp_sum(kl_acc);
        if (lane == 0) s_kl[(it - 1) * (kThreads / 32) + warp] = vk;
      }
      above = 0;
      kl_acc = 0.0;
    }
    const int t = find_tensor(s_begin, tab.n, chunk, cur < 0 ? 0 : cur);
    if (t != cur) {
      cur = t;
      const PruneDesc& d0 = tab.t[t];
      const PruneState* st = d0.state;
      mode = st->general ? 0 : (d0.k <= 0 ? 1 : (d0.k >= d0.numel ? 2 : (st->defer_all ? 4 : 3)));
      g = make_grid(st);
      want_kl = kKl && d0.kl_out != nullptr;
      prior.loc = d0.kl_loc; prior.inv_scale = d0.kl_inv_scale; prior.log_scale = d0.kl_log_scale;
    }
    if (mode == 0 && !want_kl) continue;
    const PruneDesc& d = tab.t[t];
    const int64_t ubase = (chunk - d.chunk_begin) * kChunk + warp * kUnit;
    if (ubase >= d.numel) continue;        // warp-uniform
    float2* const q_lane = s_queue + threadIdx.x;          // slot s of this lane: q_lane[s * kThreads]
    int qpos = 0;
    uint32_t listed = 0;                   // bit e: element ordinal e of this lane lies inside the grid
    const bool full = d.vec && ubase + kUnit <= d.numel;
    if (mode == 0) {                       // left to the fallback, which copies and selects on the output; KL only
      float part = 0.f;
      for (int j = 0; j < 4 * kVecPerThread; ++j) {
        const int64_t i = ubase + j * 32 + lane;
        if (i >= d.numel) continue;
        float a, b, sg, lg;
        const float mu = d.mu[i];
        key_interval<true>(mu, d.rho[i], g_unit, a, b, &sg, &lg);
        part += kl_from_sigma(mu, sg, lg, prior);
      }
      kl_acc += static_cast<double>(part);
      continue;
    }
    if (full && mode == 2 && !want_kl) {   // k == numel: write only
#pragma unroll
      for (int j = 0; j < kVecPerThread; ++j) {
        const int64_t i = ubase + (j * 32 + lane) * 4;
        st_stream4(d.mu_w + i, make_float4(0.f, 0.f, 0.f, 0.f));
        st_stream4(d.rho_w + i, make_float4(-30.f, -30.f, -30.f, -30.f));
        if (d.mask != nullptr) *reinterpret_cast<uchar4*>(d.mask + i) = make_uchar4(1, 1, 1, 1);
      }
      continue;
    }
    KlRun kl_run;
    if (full) {
      float4 m[kVecPerThread], r[kVecPerThread];
#pragma unroll
      for (int j = 0; j < kVecPerThread; ++j) {
        const int64_t i = ubase + (j * 32 + lane) * 4;
        m[j] = ldg_stream4(d.mu + i);
        r[j] = ldg_stream4(d.rho + i);
      }
#pragma unroll
      for (int j = 0; j < kVecPerThread; ++j) {
        const int64_t i = ubase + (j * 32 + lane) * 4;
        const float mm[4] = {m[j].x, m[j].y, m[j].z, m[j].w};
        const float rr[4] = {r[j].x, r[j].y, r[j].z, r[j].w};
        bool tk[4] = {mode == 2, mode == 2, mode == 2, mode == 2};
        float sg[4], lg[4];
        if (mode == 3) {
          float ym[4], yp[4];
          if (fmaxf(fmaxf(rr[0], rr[1]), fmaxf(rr[2], rr[3])) <= -1.3862944f) {
#pragma unroll
            for (int q = 0; q < 4; ++q) key_interval<false>(mm[q], rr[q], g, ym[q], yp[q], &sg[q], &lg[q]);
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) key_interval<true>(mm[q], rr[q], g, ym[q], yp[q], &sg[q], &lg[q]);
          }
          if (want_kl) {
#pragma unroll
            for (int q = 0; q < 4; ++q) kl_run.add(mm[q], sg[q], lg[q], prior);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            // above = certified above the grid (NaN never is); in-grid = neither above nor certified below (NaN is).
            // What is written for an in-grid element is the PROVISIONAL decision "upper half of the grid": the resolve
            // and finish kernels correct it where the proven bracket says otherwise (see kMidGrid)
            const bool is_above = ym[q] >= static_cast<float>(kBins);
            tk[q] = ym[q] >= kMidGrid;
            const bool in_grid = !is_above && !(yp[q] < 0.0f);
            q_lane[qpos * kThreads] = make_float2(mm[q], rr[q]);
            if (in_grid) { ++qpos; listed |= 1u << (j * 4 + q); }
            above += is_above ? 1u : 0u;
          }
        } else {
          if (mode == 4) {
#pragma unroll
            for (int q = 0; q < 4; ++q) q_lane[qpos++ * kThreads] = make_float2(mm[q], rr[q]);
            listed |= 0xfu << (j * 4);
          }
          if (want_kl) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float a, b;
              key_interval<true>(mm[q], rr[q], g_unit, a, b, &sg[q], &lg[q]);
              kl_run.add(mm[q], sg[q], lg[q], prior);
            }
          }
        }
        st_stream4(d.mu_w + i, make_float4(tk[0] ? 0.f : mm[0], tk[1] ? 0.f : mm[1], tk[2] ? 0.f : mm[2], tk[3] ? 0.f : mm[3]));
        st_stream4(d.rho_w + i, make_float4(tk[0] ? -30.f : rr[0], tk[1] ? -30.f : rr[1],
                                            tk[2] ? -30.f : rr[2], tk[3] ? -30.f : rr[3]));
        if (d.mask != nullptr) *reinterpret_cast<uchar4*>(d.mask + i) = make_uchar4(tk[0], tk[1], tk[2], tk[3]);
      }
    } else {
#pragma unroll 1
      for (int j = 0; j < 4 * kVecPerThread; ++j) {
        const int64_t i = ubase + j * 32 + lane;
        if (i >= d.numel) continue;
        const float mu = d.mu[i], rho = d.rho[i];
        bool take = mode == 2;
        if (mode == 3) {
          float ym, yp;
          float sg, lg;
          key_interval<true>(mu, rho, g, ym, yp, &sg, &lg);
          if (want_kl) kl_run.add(mu, sg, lg, prior);
          const bool is_above = ym >= static_cast<float>(kBins);
          take = ym >= kMidGrid;
          const bool in_grid = !is_above && !(yp < 0.0f);
          q_lane[qpos * kThreads] = make_float2(mu, rho);
          if (in_grid) { ++qpos; listed |= 1u << j; }
          above += is_above ? 1u : 0u;
        } else {
          if (mode == 4) {
            q_lane[qpos++ * kThreads] = make_float2(mu, rho);
            listed |= 1u << j;
          }
          if (want_kl) {
            float a, b, sg, lg;
            key_interval<true>(mu, rho, g_unit, a, b, &sg, &lg);
            kl_run.add(mu, sg, lg, prior);
          }
        }
        d.mu_w[i] = take ? 0.0f : mu;
        d.rho_w[i] = take ? -30.0f : rho;
        if (d.mask != nullptr) d.mask[i] = take ? 1 : 0;
      }
    }
    if (kKl && want_kl) kl_acc += static_cast<double>(kl_run.total(prior));
    if (mode != 3 && mode != 4) continue;
    // histogram updates of the queued in-grid elements (bin kernel conventions); the queue holds their (mu, rho) — the
    // interval is recomputed (same arithmetic, ~1 % of the elements) so that the list below needs no second global read
    if (mode == 3) {
      for (int s = 0; s < qpos; ++s) {
        const float2 e = q_lane[s * kThreads];
        float ym, yp;
        key_interval<true>(e.x, e.y, g, ym, yp);
        const int im = !(ym >= 0.0f) ? 0 : __float2int_rd(ym) + 1;
        const int ip = !(yp < static_cast<float>(kBins)) ? kBins : __float2int_rd(yp);
        atomicAdd(d.hist + im, 1u);
        atomicAdd(d.hist_plus + ip, 1u);
      }
    }
    // list append: one reservation per warp
    const unsigned int mine = __popc(listed);
    unsigned int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    const unsigned int total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0u) continue;           // warp-uniform
    unsigned int base = 0;
    if (lane == 0) base = atomicAdd(&d.state->n_deferred, total);
    base = __shfl_sync(0xffffffffu, base, 0);
    unsigned int slot = base + incl - mine;
    int s = 0;                           // queue entries are in element order, like the bits of `listed`
    while (listed != 0u) {
      const int e = __ffs(listed) - 1;
      listed &= listed - 1u;
      const int64_t i = full ? ubase + ((e >> 2) * 32 + lane) * 4 + (e & 3) : ubase + e * 32 + lane;
      const float2 v = q_lane[s * kThreads];
      ++s;
      if (slot < d.defer_cap)            // one 16-byte record per listed element: (index, mu, rho, -); numel < 2^32 here
        reinterpret_cast<uint4*>(d.keys)[slot] =
            make_uint4(static_cast<uint32_t>(i), __float_as_uint(v.x), __float_as_uint(v.y), 0u);
      ++slot;
    }
  }
  {                        // the last chunk's counts, then one atomic per (block, chunk)
    int last = 0;
    while (last + 1 < kSweepChunks && blockIdx.x + static_cast<int64_t>(last + 1) * gridDim.x < tab.total_chunks) ++last;
    if (blockIdx.x < tab.total_chunks) {
      const unsigned int va = warp_sum(above);
      if (lane == 0) s_above[last * (kThreads / 32) + warp] = va;
      if (kKl) {
        const double vk = warp_sum(kl_acc);
        if (lane == 0) s_kl[last * (kThreads / 32) + warp] = vk;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < kSweepChunks) {
    const int64_t chunk = blockIdx.x + static_cast<int64_t>(threadIdx.x) * gridDim.x;
    if (chunk < tab.total_chunks) {
      const PruneDesc& d = tab.t[find_tensor(s_begin, tab.n, chunk, 0)];
      unsigned int va = 0;
      double vk = 0.0;
      for (int w = 0; w < kThreads / 32; ++w) {
        va += s_above[threadIdx.x * (kThreads / 32) + w];
        vk += s_kl[threadIdx.x * (kThreads / 32) + w];
      }
      if (va != 0u) atomicAdd(&d.state->count_above, static_cast<unsigned long long>(va));
      if (kKl && d.kl_out != nullptr) atomicAdd(d.kl_out, vk);
    }
  }
}

// bnn_prune_into: after the bracket step, walk the listed in-grid elements of every tensor: certainly above the bracket ->
// pruned in the output (scattered stores), overlapping it -> the exact-select list, below -> nothing (already copied)
__global__ void __launch_bounds__(kThreads) prune_resolve_kernel(const __grid_constant__ PruneTable tab) {
  const PruneDesc& d = tab.t[blockIdx.y];
  const PruneState* stp = d.state;
  if (stp->general || d.k <= 0 || d.k >= d.numel) return;
  const uint32_t n = stp->n_deferred;
  if (n > d.defer_cap) return;             // overflow: the bracket kernel has flagged the tensor (see there)
  const Grid g = make_grid(stp);
  const float a_thr = stp->a_thr, b_thr = stp->b_thr;
  const int lane = threadIdx.x & 31;
  for (uint32_t i0 = blockIdx.x * kThreads; i0 < n; i0 += gridDim.x * kThreads) {
    const uint32_t i = i0 + threadIdx.x;
    bool take = false, defer = false, provisional = false;
    uint32_t idx = 0;
    float mu = 0.f, rho = 0.f;
    if (i < n) {
      const uint4 rec = reinterpret_cast<const uint4*>(d.keys)[i];
      idx = rec.x;
      mu = __uint_as_float(rec.y);
      rho = __uint_as_float(rec.z);
      if (stp->defer_all) {
        defer = true;
      } else {
        float ym, yp;
        key_interval<true>(mu, rho, g, ym, yp);        // the same arithmetic as the sweep: identical intervals
        take = ym >= a_thr;
        defer = !take && !(yp < b_thr);
        provisional = ym >= kMidGrid;                  // what the sweep wrote
      }
    }
    if (take && !provisional) {
      d.mu_w[idx] = 0.0f;
      d.rho_w[idx] = -30.0f;
      if (d.mask != nullptr) d.mask[idx] = 1;
    } else if (i < n && !take && !defer && provisional) {
      d.mu_w[idx] = mu;
      d.rho_w[idx] = rho;
      if (d.mask != nullptr) d.mask[idx] = 0;
    }
    const unsigned int bal = __ballot_sync(0xffffffffu, defer);
    if (bal != 0u) {
      const int leader = __ffs(bal) - 1;
      unsigned int base = 0;
      if (lane == leader) base = atomicAdd(&d.state->n_deferred2, static_cast<unsigned int>(__popc(bal)));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (defer) {
        const unsigned int slot = base + __popc(bal & ((1u << lane) - 1u));
        if (slot < d.cap2) {
          d.list2[slot] = idx;
          d.list2[d.cap2 + slot] = __float_as_uint(mu);
          d.list2[2u * static_cast<size_t>(d.cap2) + slot] = __float_as_uint(rho);
        }
      }
    }
  }
}

// bnn_prune_into, fallback: tensors flagged for the general path are first copied to the output unchanged
__global__ void __launch_bounds__(kThreads) prune_copy_general_kernel(const __grid_constant__ PruneTable tab) {
  if (*reinterpret_cast<volatile uint32_t*>(tab.any_general) == 0u) return;
  __shared__ int64_t s_begin[kMaxTensors];
  if (threadIdx.x < tab.n) s_begin[threadIdx.x] = tab.t[threadIdx.x].chunk_begin;
  __syncthreads();
  int cur = -1;
  for (int64_t chunk = blockIdx.x; chunk < tab.total_chunks; chunk += gridDim.x) {
    cur = find_tensor(s_begin, tab.n, chunk, cur < 0 ? 0 : cur);
    const PruneDesc& d = tab.t[cur];
    if (d.state->general == 0u) continue;
    const int64_t base = (chunk - d.chunk_begin) * kChunk;
    for (int j = 0; j < 4 * kVecPerThread; ++j) {
      const int64_t i = base + static_cast<int64_t>(j) * kThreads + threadIdx.x;
      if (i < d.numel) { d.mu_w[i] = d.mu[i]; d.rho_w[i] = d.rho[i]; }
    }
  }
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

// 5. one block per tensor: exact keys of the deferred elements, exact select of the n_take largest (ties -> lowest
// index), scattered writes.  The usual list (~1000 entries) is handled in shared memory; longer ones (ties, small
// tensors that defer everything) go through (key, index) pairs in the workspace.
constexpr uint32_t kResolveList = 4096;      // (key, index) pairs in shared memory

__global__ void __launch_bounds__(kFinishThreads) prune_finish_kernel(const __grid_constant__ PruneTable tab) {
  __shared__ uint32_t s_hist[kBins];
  __shared__ uint32_t s_bcast[4];
  __shared__ __align__(8) uint32_t s_list[2 * kResolveList];
  const PruneDesc& d = tab.t[blockIdx.x];
  const PruneState st = *d.state;
  if (st.general || d.k <= 0 || d.k >= d.numel) return;
  const bool two_level = d.list2 != nullptr;                 // bnn_prune_into: the list the resolve kernel wrote
  const uint32_t cap = two_level ? d.cap2 : d.defer_cap;
  const uint32_t listed = two_level ? st.n_deferred2 : st.n_deferred;
  uint32_t n = listed < cap ? listed : cap;                   // == expect_deferred by construction
  const uint32_t* base = two_level ? d.list2 : d.keys;
  const uint32_t* e_idx = base;
  const uint32_t* e_mu = base + cap;
  const uint32_t* e_rho = base + 2u * static_cast<size_t>(cap);
  uint32_t* pairs = n <= kResolveList ? s_list : const_cast<uint32_t*>(base) + 3u * static_cast<size_t>(cap);
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    pairs[2 * static_cast<size_t>(i)] = order_key(prune_key(__uint_as_float(e_mu[i]), __uint_as_float(e_rho[i])));
    pairs[2 * static_cast<size_t>(i) + 1] = e_idx[i];
  }
  __syncthreads();
  uint32_t take = st.n_take < n ? st.n_take : n;
  if (take == 0u) {
    if (two_level)             // the sweep's provisional decision may have pruned some of them
      for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t idx = e_idx[i];
        d.mu_w[idx] = __uint_as_float(e_mu[i]);
        d.rho_w[idx] = __uint_as_float(e_rho[i]);
        if (d.mask != nullptr) d.mask[idx] = 0;
      }
    return;
  }
  uint64_t above = 0;
  uint32_t eq = 0;
  const uint32_t T = cand_select(pairs, n, 0, 0u, static_cast<uint64_t>(take) - 1, s_hist, s_bcast, &above, &eq);
  const uint64_t take_eq = take - above;                    // how many of the key == T entries are taken
  const bool all_eq = take_eq >= eq;
  uint32_t idx_bound = 0xffffffffu;
  if (!all_eq) {
    uint64_t b2 = 0;
    uint32_t e2 = 0;
    idx_bound = cand_select(pairs, n, 1, T, take_eq - 1, s_hist, s_bcast, &b2, &e2);
  }
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    const uint32_t key = pairs[2 * static_cast<size_t>(i)], idx = pairs[2 * static_cast<size_t>(i) + 1];
    if (key > T || (key == T && (all_eq || idx <= idx_bound))) {
      d.mu_w[idx] = 0.0f;
      d.rho_w[idx] = -30.0f;
      if (d.mask != nullptr) d.mask[idx] = 1;
    } else if (two_level) {    // bnn_prune_into: undo a provisional decision of the sweep
      d.mu_w[idx] = __uint_as_float(e_mu[i]);
      d.rho_w[idx] = __uint_as_float(e_rho[i]);
      if (d.mask != nullptr) d.mask[idx] = 0;
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

// Workspace: [group flags][state + two histograms per tensor] — cleared by ONE memset — then per tensor the keys
// region and the chunk counts.  The keys region holds the ordered keys of the general path (4 B/element) or, on the
// sampled path, the deferred list ([index | mu | rho] x defer_cap) followed by (key, index) pairs x defer_cap.
constexpr size_t kStateBytes = 256;                            // PruneState, padded
constexpr size_t kSmallBytes = kStateBytes + 2 * kHistStride * 4;    // state + histograms of one tensor
static_assert(sizeof(PruneState) <= kStateBytes, "PruneState outgrew its slot");

uint32_t defer_cap_for(int64_t numel) {
  if (numel <= kSmallTensor) return static_cast<uint32_t>(numel);
  int64_t cap = numel / 16;                                    // far above the ~3 % that overlap the first bracket
  if (cap < 4096) cap = 4096;
  if (cap > (int64_t(1) << 28)) cap = int64_t(1) << 28;
  return static_cast<uint32_t>(cap);
}
uint32_t cap2_for(int64_t numel) {            // bnn_prune_into: capacity of the exact-select list
  const uint32_t cap = defer_cap_for(numel);
  return cap < 262144u ? cap : 262144u;
}
size_t keys_bytes(int64_t numel) {
  const size_t general = static_cast<size_t>(numel) * 4, sampled = static_cast<size_t>(defer_cap_for(numel)) * 24;
  return align_up(general > sampled ? general : sampled, 256);
}
size_t header_bytes(int n_tensors) {       // one "some tensor needs the general path" flag per group of kMaxTensors
  const int groups = (n_tensors + kMaxTensors - 1) / kMaxTensors;
  return align_up(static_cast<size_t>(groups > 0 ? groups : 1) * 4, 256);
}
size_t prune_ws_one(int64_t numel) {
  const int64_t chunks = (numel + kChunk - 1) / kChunk;
  return keys_bytes(numel) + align_up(static_cast<size_t>(chunks) * 8, 256);
}

}  // namespace
}  // namespace bnn

using namespace bnn;

extern "C" {

size_t bnn_prune_workspace_size(const bnn_prune_tensor* tensors, int32_t n_tensors) {
  if (tensors == nullptr || n_tensors <= 0) return 256;
  size_t total = header_bytes(n_tensors) + static_cast<size_t>(n_tensors) * kSmallBytes;
  for (int i = 0; i < n_tensors; ++i) total += prune_ws_one(tensors[i].numel > 0 ? tensors[i].numel : 0);
  return total;
}

int bnn_selftest_prune_interval(const float* mu, const float* rho, int64_t numel, float* lo_out, float* hi_out,
                                int32_t variant, void* stream) {
  BNN_REQUIRE(numel >= 0 && (variant == 0 || variant == 1), BNN_ERR_BAD_ARGUMENT,
              "bnn_selftest_prune_interval: numel >= 0 and variant in {0, 1}");
  if (numel == 0) return BNN_OK;
  BNN_REQUIRE(mu && rho && lo_out && hi_out, BNN_ERR_BAD_ARGUMENT, "bnn_selftest_prune_interval: NULL pointer");
  int rc = check_device();
  if (rc != BNN_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = static_cast<int>((numel + 255) / 256 < sm_count() * 8 ? (numel + 255) / 256 : sm_count() * 8);
  if (variant == 0) prune_interval_selftest_kernel<false><<<grid, 256, 0, st>>>(mu, rho, numel, lo_out, hi_out);
  else prune_interval_selftest_kernel<true><<<grid, 256, 0, st>>>(mu, rho, numel, lo_out, hi_out);
  BNN_CUDA_OK(cudaGetLastError());
  return BNN_OK;
}

namespace {
struct PruneIo {                       // one tensor of either entry point
  float* mu; float* rho; float* mu_out; float* rho_out; uint8_t* mask; float* keys_out;
  int64_t numel, k;
  uint32_t flags;
  double* kl_out; float prior_loc, prior_scale;      // bnn_prune_into only
};

size_t into_extra_bytes(int64_t numel) {       // the exact-select list of bnn_prune_into: 20 bytes per entry
  return align_up(static_cast<size_t>(cap2_for(numel)) * 20, 256);
}

// A second stream per device for bnn_prune_into calls that span several groups of kMaxTensors tensors: the short
// single-block steps of group g (bracket, resolve, finish, idle fallback launches) run there while the main stream
// already sweeps group g + 1.  Created lazily; if anything fails the call runs on the caller's stream alone.
struct SideLane {
  cudaStream_t stream = nullptr;
  cudaStream_t fallback = nullptr;      // the (normally idle) general-path launches of a group, off the critical path
  std::vector<cudaEvent_t> events;
  bool failed = false;
};
std::mutex g_side_mutex;
SideLane g_side[64];

SideLane* side_lane(int n_events) {      // call with g_side_mutex held
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  SideLane& lane = g_side[dev];
  if (lane.failed) return nullptr;
  if (lane.stream == nullptr) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);            // hi = numerically lowest = highest priority
    if (cudaStreamCreateWithPriority(&lane.stream, cudaStreamNonBlocking, hi) != cudaSuccess) {
      cudaGetLastError();
      lane.stream = nullptr;
      lane.failed = true;
      return nullptr;
    }
    if (cudaStreamCreateWithPriority(&lane.fallback, cudaStreamNonBlocking, hi) != cudaSuccess) {
      cudaGetLastError();
      lane.failed = true;
      return nullptr;
    }
  }
  while (static_cast<int>(lane.events.size()) < n_events) {
    cudaEvent_t e;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      lane.failed = true;
      return nullptr;
    }
    lane.events.push_back(e);
  }
  return &lane;
}

void launch_general_path(PruneTable tab, bool into, int grid, cudaStream_t st) {
  // general path (kernels return immediately unless a tensor asked for it; when no tensor is known to need it they
  // are launched with one block per SM — all of them loop over the chunks — so that the idle launches stay cheap)
  bool forced = false;
  for (int i = 0; i < tab.n; ++i) forced = forced || tab.t[i].force_general != 0u || tab.t[i].numel >= (int64_t(1) << 32);
  const int ggrid = forced ? grid : (grid < sm_count() ? grid : sm_count());
  if (into) {
    // the input is intact: flagged tensors are copied, then selected in place on the OUTPUT
    prune_copy_general_kernel<<<ggrid, kThreads, 0, st>>>(tab);
    for (int i = 0; i < tab.n; ++i) { tab.t[i].mu = tab.t[i].mu_w; tab.t[i].rho = tab.t[i].rho_w; }
  }
  prune_hist_kernel<0><<<ggrid, kThreads, 0, st>>>(tab);
  prune_select_kernel<0><<<tab.n, kThreads, 0, st>>>(tab);
  prune_hist_kernel<1><<<ggrid, kThreads, 0, st>>>(tab);
  prune_select_kernel<1><<<tab.n, kThreads, 0, st>>>(tab);
  prune_hist_kernel<2><<<ggrid, kThreads, 0, st>>>(tab);
  prune_select_kernel<2><<<tab.n, kThreads, 0, st>>>(tab);
  prune_count_eq_kernel<<<ggrid, kThreads, 0, st>>>(tab);
  prune_scan_kernel<<<tab.n, kThreads, 0, st>>>(tab);
  prune_apply_kernel<<<ggrid, kThreads, 0, st>>>(tab);
}

int prune_run(const PruneIo* io, int32_t n_tensors, bool into, void* workspace, cudaStream_t st) {
  uint32_t* group_flags = reinterpret_cast<uint32_t*>(workspace);
  char* small = static_cast<char*>(workspace) + header_bytes(n_tensors);
  char* ws = small + static_cast<size_t>(n_tensors) * kSmallBytes;
  BNN_CUDA_OK(cudaMemsetAsync(workspace, 0, static_cast<size_t>(ws - static_cast<char*>(workspace)), st));
  const int max_grid = sm_count() * 8;

  std::vector<PruneTable> tabs;
  // groups of at most kMaxTensors tensors (one descriptor table each), balanced: 64 tensors -> 22 + 21 + 21
  // (cutting ONE table of large tensors in two, to overlap the first half's short steps with the second half's sweep, was
  // tried: the extra launches cost the host as much as the overlap saves on a 0.5 ms call)
  const int n_tables = (n_tensors + kMaxTensors - 1) / kMaxTensors;
  const int per_table = (n_tensors + n_tables - 1) / n_tables;
  for (int first = 0, gi = 0; first < n_tensors; first += per_table, ++gi) {
    const int n = (n_tensors - first < per_table) ? n_tensors - first : per_table;
    PruneTable tab;
    tab.n = 0;
    tab.pad = 0;
    tab.any_general = group_flags + gi;
    int64_t chunks = 0;
    for (int i = 0; i < n; ++i) {
      const PruneIo& t = io[first + i];
      if (t.numel == 0) continue;
      PruneDesc& d = tab.t[tab.n++];
      const int64_t nch = (t.numel + kChunk - 1) / kChunk;
      char* mine = small + static_cast<size_t>(first + i) * kSmallBytes;
      d.mu = t.mu; d.rho = t.rho; d.mask = t.mask; d.keys_out = t.keys_out;
      d.mu_w = into ? t.mu_out : t.mu; d.rho_w = into ? t.rho_out : t.rho;
      d.numel = t.numel; d.k = t.k; d.chunk_begin = chunks; d.n_chunks = nch;
      d.state = reinterpret_cast<PruneState*>(mine);
      d.hist = reinterpret_cast<uint32_t*>(mine + kStateBytes);
      d.hist_plus = d.hist + kHistStride;
      d.keys = reinterpret_cast<uint32_t*>(ws); ws += keys_bytes(t.numel);
      d.defer_cap = defer_cap_for(t.numel);
      d.chunk_cnt = reinterpret_cast<int64_t*>(ws); ws += align_up(static_cast<size_t>(nch) * 8, 256);
      d.list2 = nullptr; d.cap2 = 0; d.pad2 = 0;
      if (into) { d.list2 = reinterpret_cast<uint32_t*>(ws); d.cap2 = cap2_for(t.numel); ws += into_extra_bytes(t.numel); }
      d.kl_out = into ? t.kl_out : nullptr;
      d.kl_loc = t.prior_loc;
      d.kl_inv_scale = d.kl_out != nullptr ? 1.0f / t.prior_scale : 0.f;
      d.kl_log_scale = d.kl_out != nullptr ? logf(t.prior_scale) : 0.f;
      d.pad3 = 0;
      d.force_general = ((t.flags & BNN_PRUNE_GENERAL) != 0u || t.keys_out != nullptr) ? 1u : 0u;
      d.vec = aligned16(t.mu) && aligned16(t.rho) && aligned16(d.mu_w) && aligned16(d.rho_w) &&
              (t.mask == nullptr || (reinterpret_cast<uintptr_t>(t.mask) & 3u) == 0);
      d.pad = 0;
      chunks += nch;
    }
    if (tab.n == 0) continue;
    tab.total_chunks = chunks;
    tabs.push_back(tab);
  }
  const int n_groups = static_cast<int>(tabs.size());
  if (n_groups == 0) return BNN_OK;
  auto grid_of = [&](const PruneTable& tab) { return static_cast<int>(tab.total_chunks < max_grid ? tab.total_chunks : max_grid); };
  // short-lived sweep blocks (kSweepChunks chunks each) instead of a persistent grid: SM slots free up all the time, so
  // the side lane's small kernels of the previous group get onto the machine while this group is being swept
  auto sweep_grid_of = [&](const PruneTable& tab) {
    const int64_t g = (tab.total_chunks + kSweepChunks - 1) / kSweepChunks;
    return static_cast<int>(g < 1 ? 1 : (g > (int64_t(1) << 30) ? (int64_t(1) << 30) : g));
  };

  if (!into) {
    for (int g = 0; g < n_groups; ++g) {
      const PruneTable& tab = tabs[g];
      const int grid = grid_of(tab);
      prune_sample_kernel<<<tab.n * kSampleCtas, kResolveThreads, 0, st>>>(tab);
      prune_bin_kernel<<<grid, kThreads, 0, st>>>(tab);
      prune_bracket_kernel<<<tab.n, kThreads, 0, st>>>(tab);
      prune_apply_sampled_kernel<<<grid, kThreads, 0, st>>>(tab);
      prune_finish_kernel<<<tab.n, kFinishThreads, 0, st>>>(tab);
      launch_general_path(tab, false, grid, st);
    }
    BNN_CUDA_OK(cudaGetLastError());
    return BNN_OK;
  }

  // bnn_prune_into over several groups: the main stream runs sample(0), sweep(0), sweep(1), ...; the side lane runs the
  // samples of the later groups (hidden behind sweep(0)) and then, per group, bracket / resolve / finish / fallback as
  // soon as that group's sweep is done — while the main stream sweeps the next group.
  std::unique_lock<std::mutex> lock(g_side_mutex, std::defer_lock);
  SideLane* lane = nullptr;
  if (n_groups > 1) {
    lock.lock();
    lane = side_lane(4 * n_groups + 2);
    if (lane == nullptr) lock.unlock();
  }
  cudaEvent_t* ev_sweep = lane ? lane->events.data() : nullptr;
  cudaEvent_t* ev_sample = lane ? lane->events.data() + n_groups : nullptr;
  prune_sample_kernel<<<tabs[0].n * kSampleCtas, kResolveThreads, 0, st>>>(tabs[0]);
  if (lane != nullptr) {
    cudaEvent_t fork = lane->events[2 * n_groups];
    BNN_CUDA_OK(cudaEventRecord(fork, st));
    BNN_CUDA_OK(cudaStreamWaitEvent(lane->stream, fork, 0));
  }
  for (int g = 1; g < n_groups; ++g) {
    prune_sample_kernel<<<tabs[g].n * kSampleCtas, kResolveThreads, 0, lane ? lane->stream : st>>>(tabs[g]);
    if (lane != nullptr) BNN_CUDA_OK(cudaEventRecord(ev_sample[g], lane->stream));
  }
  for (int g = 0; g < n_groups; ++g) {
    const PruneTable& tab = tabs[g];
    if (lane != nullptr && g > 0) BNN_CUDA_OK(cudaStreamWaitEvent(st, ev_sample[g], 0));
    bool any_kl = false;
    for (int i = 0; i < tab.n; ++i) any_kl = any_kl || tab.t[i].kl_out != nullptr;
    if (any_kl) prune_sweep_into_kernel<true><<<sweep_grid_of(tab), kThreads, 0, st>>>(tab);
    else prune_sweep_into_kernel<false><<<sweep_grid_of(tab), kThreads, 0, st>>>(tab);
    cudaStream_t post = st;
    if (lane != nullptr) {
      BNN_CUDA_OK(cudaEventRecord(ev_sweep[g], st));
      BNN_CUDA_OK(cudaStreamWaitEvent(lane->stream, ev_sweep[g], 0));
      post = lane->stream;
    }
    prune_bracket_kernel<<<tab.n, kThreads, 0, post>>>(tab);
    cudaStream_t gen = post;
    if (lane != nullptr) {     // the fallback launches only need the bracket step's verdict: third stream, beside resolve / finish
      cudaEvent_t bracketed = lane->events[2 * n_groups + 2 + 2 * g];
      BNN_CUDA_OK(cudaEventRecord(bracketed, post));
      BNN_CUDA_OK(cudaStreamWaitEvent(lane->fallback, bracketed, 0));
      gen = lane->fallback;
    }
    prune_resolve_kernel<<<dim3(64, tab.n), kThreads, 0, post>>>(tab);
    prune_finish_kernel<<<tab.n, kFinishThreads, 0, post>>>(tab);
    launch_general_path(tab, true, grid_of(tab), gen);
    if (lane != nullptr) {     // ... and rejoin the side lane (one join with the caller's stream at the end)
      cudaEvent_t done = lane->events[2 * n_groups + 3 + 2 * g];
      BNN_CUDA_OK(cudaEventRecord(done, gen));
      BNN_CUDA_OK(cudaStreamWaitEvent(post, done, 0));
    }
  }
  if (lane != nullptr) {
    BNN_CUDA_OK(cudaEventRecord(lane->events[2 * n_groups + 1], lane->stream));
    BNN_CUDA_OK(cudaStreamWaitEvent(st, lane->events[2 * n_groups + 1], 0));
  }
  BNN_CUDA_OK(cudaGetLastError());
  return BNN_OK;
}
}  // namespace

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
  std::vector<PruneIo> io(static_cast<size_t>(n_tensors));
  for (int i = 0; i < n_tensors; ++i) {
    BNN_REQUIRE(tensors[i].numel >= 0 && tensors[i].k >= 0 && tensors[i].k <= tensors[i].numel,
                BNN_ERR_BAD_ARGUMENT, "bnn_prune: tensor %d needs 0 <= k <= numel", i);
    BNN_REQUIRE(tensors[i].numel == 0 || (tensors[i].mu && tensors[i].rho), BNN_ERR_BAD_ARGUMENT,
                "bnn_prune: tensor %d has NULL mu/rho", i);
    io[i] = PruneIo{tensors[i].mu, tensors[i].rho, nullptr, nullptr, tensors[i].mask_out, tensors[i].keys_out,
                    tensors[i].numel, tensors[i].k, tensors[i].flags, nullptr, 0.f, 1.f};
  }
  int rc = check_device();
  if (rc != BNN_OK) return rc;
  return prune_run(io.data(), n_tensors, false, workspace, static_cast<cudaStream_t>(stream));
}

size_t bnn_prune_into_workspace_size(const bnn_prune_into_tensor* tensors, int32_t n_tensors) {
  if (tensors == nullptr || n_tensors <= 0) return 256;
  size_t total = header_bytes(n_tensors) + static_cast<size_t>(n_tensors) * kSmallBytes;
  for (int i = 0; i < n_tensors; ++i) {
    const int64_t numel = tensors[i].numel > 0 ? tensors[i].numel : 0;
    total += prune_ws_one(numel) + into_extra_bytes(numel);
  }
  return total;
}

int bnn_prune_into(const bnn_prune_into_tensor* tensors, int32_t n_tensors, void* workspace, size_t workspace_bytes,
                   void* stream) {
  BNN_REQUIRE(n_tensors >= 0, BNN_ERR_BAD_ARGUMENT, "n_tensors < 0");
  if (n_tensors == 0) return BNN_OK;
  BNN_REQUIRE(tensors != nullptr, BNN_ERR_BAD_ARGUMENT, "bnn_prune_into: tensor table is NULL");
  BNN_REQUIRE(workspace != nullptr && workspace_bytes >= bnn_prune_into_workspace_size(tensors, n_tensors),
              BNN_ERR_WORKSPACE, "bnn_prune_into: workspace too small (%zu < %zu)", workspace_bytes,
              bnn_prune_into_workspace_size(tensors, n_tensors));
  BNN_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, BNN_ERR_MISALIGNED,
              "bnn_prune_into: workspace must be 256-byte aligned");
  std::vector<PruneIo> io(static_cast<size_t>(n_tensors));
  for (int i = 0; i < n_tensors; ++i) {
    const bnn_prune_into_tensor& t = tensors[i];
    BNN_REQUIRE(t.numel >= 0 && t.k >= 0 && t.k <= t.numel, BNN_ERR_BAD_ARGUMENT,
                "bnn_prune_into: tensor %d needs 0 <= k <= numel", i);
    BNN_REQUIRE(t.numel == 0 || (t.mu && t.rho && t.mu_out && t.rho_out), BNN_ERR_BAD_ARGUMENT,
                "bnn_prune_into: tensor %d has a NULL pointer", i);
    BNN_REQUIRE(t.numel == 0 || (t.mu != t.mu_out && t.rho != t.rho_out), BNN_ERR_BAD_ARGUMENT,
                "bnn_prune_into: tensor %d: outputs must not alias the inputs (use bnn_prune in place)", i);
    BNN_REQUIRE(t.kl_sum_out == nullptr || (t.prior_scale > 0.f && (reinterpret_cast<uintptr_t>(t.kl_sum_out) & 7u) == 0),
                BNN_ERR_BAD_ARGUMENT, "bnn_prune_into: tensor %d: kl_sum_out needs prior_scale > 0 and 8-byte alignment", i);
    io[i] = PruneIo{const_cast<float*>(t.mu), const_cast<float*>(t.rho), t.mu_out, t.rho_out, t.mask_out, nullptr,
                    t.numel, t.k, t.flags, t.kl_sum_out, t.prior_loc, t.prior_scale};
  }
  int rc = check_device();
  if (rc != BNN_OK) return rc;
  return prune_run(io.data(), n_tensors, true, workspace, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
