// sampled_gemm_tma.cu — TMA-fed TF32 variants of the sample-and-contract kernels for aligned row-major operands (the
// large-shape path: BASELINE config C4 and the im2col matrices of the conv layers).
//
//   forward / data gradient   Activations (or dY) arrive by TMA (cp.async.bulk.tensor.3d, 128B swizzle, TF32 conversion in
//     contract_tma_kernel     the copy engine): the MB tiles of a k-block land in one of two slot groups on one
//     contract_pair_kernel    transaction barrier.  Sixteen generator warps in four groups do nothing but produce
//                             W_s = mu + sigma * eps_s (Philox) — group j owns weight slot j and the k-blocks
//                             it % 4 == j, so four tiles are in flight.  One thread issues tcgen05.mma and ONE
//                             tcgen05.commit per k-block ("k-block consumed": frees the weight slot and, two k-blocks
//                             later, the activation group).  Four warps drain TMEM.  The data gradient stores the
//                             generated tile in the MN-major canonical layout, i.e. in W's natural [n][k] orientation.
//                             The pair kernel runs the same protocol over a cluster of two CTAs (cta_group::2): each
//                             generates half of the weight tile, the leader issues M = 256 MMAs and multicast commits.
//   weight gradient           BOTH operands (dY^T and A^T) are MN-major tiles loaded by TMA — no thread touches
//     wgrad_tma_kernel        operand data; twelve epilogue warps regenerate eps per work unit and keep the running
//                             sums of G_s and G_s o eps_s in TMEM.
// Eligibility (else the caller uses sampled_gemm.cu): TF32 precision, row-major operands (view P == 1 for everything
// that is READ), 16-byte aligned bases and leading dimensions.
// Profiling builds (-DBNN_PROFILE_WAITS) add wait-cycle counters and elimination switches (BNN_EXP_FLAGS).
#include <cuda.h>

#include <cstdlib>
#include <mutex>

#include <algorithm>
#include <array>
#include <vector>

#include "contract.cuh"

namespace bnn {
namespace contract {
#ifdef BNN_PROFILE_WAITS
// profiling builds: cycles spent waiting, summed over CTAs: [0] kernel, [1] MMA on full_w, [2] MMA on full_a,
// [3] generator warp 0 on empty_w, [4] TMA thread on the consumed barrier, [5] CTAs
__device__ unsigned long long g_wait_cycles[8];
// CTA timeline of the CTA-pair kernels, cycles since the CTA's first instruction summed over CTAs: [0] prologue done,
// [1] generator warp 0 done, [2] accumulator complete (epilogue warp 0), [3] epilogue done, [4] CTA end, [5] CTAs;
// [6] / [7] globaltimer (ns) of the first CTA start / last CTA end
__device__ unsigned long long g_stage_cycles[8];
#define BNN_STAGE(i, t0) atomicAdd(&g_stage_cycles[i], (unsigned long long)(clock64() - (t0)))
#define BNN_T0() const long long _t0 = clock64()
#define BNN_ACC(var) var += clock64() - _t0
#else
#define BNN_T0()
#define BNN_ACC(var)
#endif
namespace {

constexpr int kASlots = 8;                   // activation tile slots (16 KiB each): two k-block groups of up to 4 tiles
constexpr int kWSlots = 4;                   // generated weight tile ring (the k-block bookkeeping assumes 4)
// warp roles of the forward / data-gradient kernel: 16 weight generators (the Philox chains are latency bound:
// four warps per scheduler hide them), MMA issuer, four epilogue warps (TMEM lane quarter = warp % 4), TMA issuer
constexpr int kGenWarps = 16;
constexpr int kGenGroups = 4;                       // generator groups: group j produces the k-blocks it % 4 == j into
constexpr int kGroupWarps = kGenWarps / kGenGroups; // weight slot j, so four tiles are in flight at different phases
constexpr int kGroupThreads = kGroupWarps * 32;     // (one tile's latency — loads, Philox chain — is hidden by the others)
constexpr int kMmaWarpT = kGenWarps;
constexpr int kEpiWarp0T = kGenWarps + 1;
constexpr int kTmaWarp = kGenWarps + 5;
constexpr int kThreadsTma = (kGenWarps + 6) * 32;
constexpr uint32_t kMnLbo = 4096, kMnSbo = 512;    // MN-major 128 x 32 tile: 4 groups of 32 MN, 8 atoms of 4 K-rows
constexpr uint32_t kMnKStep = 2 * kMnSbo;            // one UMMA K-step (8 TF32) = two 4-row atoms

// ---------------------------------------------------------------------------------------------- tensor maps
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn encode_fn() {
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(p);
  });
  return fn;
}

// fp32 matrix [samples][rows][cols] (row stride ld, sample stride ss floats) -> 3-d map with a box of
// 32 columns x box_rows rows, 128B swizzle, TF32 conversion, zero fill outside [cols) x [rows) x [samples)
int make_map(CUtensorMap* map, const float* base, int64_t cols, int64_t rows, int64_t samples, int64_t ld, int64_t ss,
             int box_rows, bool mn_major = false) {
  EncodeFn fn = encode_fn();
  if (fn == nullptr) return kNotEligible;
  const cuuint64_t dims[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows),
                              static_cast<cuuint64_t>(samples)};
  const cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * 4u,
                                 static_cast<cuuint64_t>(samples > 1 ? ss : ld * rows) * 4u};
  const cuuint32_t box[3] = {32u, static_cast<cuuint32_t>(box_rows), 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return kNotEligible;
  return BNN_OK;
}

// NHWC fp32 tensor [n_imgs][H][W][C] seen through the filter's bounding box (cuTensorMapEncodeIm2col): a load delivers
// `box_pixels` base pixels (traversed w, h, n with the given stride, starting at the lower corner) x 32 channels, TF32
// conversion, zero fill outside the image.  lower / upper corner: first base pixel and (last base pixel + 1 - extent).
using EncodeIm2colFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);
EncodeIm2colFn encode_im2col_fn() {
  static EncodeIm2colFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeIm2colFn>(p);
  });
  return fn;
}

int make_im2col_map(CUtensorMap* map, const float* base, int64_t C, int64_t W, int64_t H, int64_t n_imgs, int lower_w,
                    int lower_h, int upper_w, int upper_h, int stride_w, int stride_h, int box_pixels, bool mn_major) {
  EncodeIm2colFn fn = encode_im2col_fn();
  if (fn == nullptr) return kNotEligible;
  const cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                              static_cast<cuuint64_t>(n_imgs)};
  const cuuint64_t strides[3] = {static_cast<cuuint64_t>(C) * 4u, static_cast<cuuint64_t>(W * C) * 4u,
                                 static_cast<cuuint64_t>(H * W * C) * 4u};
  const int lower[2] = {lower_w, lower_h}, upper[2] = {upper_w, upper_h};
  const cuuint32_t estr[4] = {1u, static_cast<cuuint32_t>(stride_w), static_cast<cuuint32_t>(stride_h), 1u};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 4, const_cast<float*>(base), dims, strides, lower, upper,
                        32u, static_cast<cuuint32_t>(box_pixels), estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return kNotEligible;
  // Known driver issue (CUDA <= 13.1, the workaround CUTLASS applies in cute/atom/copy_traits_sm90_im2col.hpp): for tensors
  // smaller than 128 KiB the encoder sets bit 21 of the descriptor's second word, which makes the copy engine misread the
  // extent; clear it.
  int driver = 0;
  if (cudaDriverGetVersion(&driver) == cudaSuccess && driver <= 13010 && n_imgs * H * W * C * 4 < 131072)
    reinterpret_cast<uint64_t*>(map)[1] &= ~(1ull << 21);
  return BNN_OK;
}

bool tma_ok(const void* base, int64_t ld, int64_t ss) {
  return aligned16(base) && ld % 4 == 0 && ss % 4 == 0 && ld > 0;
}

// ---------------------------------------------------------------------------------------------- forward / dgrad
// Implicit-GEMM addressing of the TMA-loaded operand (conv layers): the operand is never an im2col MATRIX, it is the NHWC
// tensor itself seen through an im2col tensor map; row m of the GEMM is the pixel (img, p, q) of a P = rows x OW grid
// (output pixels for forward / weight gradient, input pixels for the data gradient), k-block rb is (filter tap, 32
// channels).  on == 0: plain row-major matrix.
struct ConvCoords {
  int on;
  int P, OW;                // pixel grid of the GEMM rows
  int imgs;                 // images per sample: n coordinate = sample * imgs + img
  int sh, sw;               // traversal stride (the conv stride; 1 for the data gradient)
  int lh, lw;               // lower corner: base pixel of grid pixel 0
  int dh, dw;               // dilation
  int KW, taps;             // filter width, KH * KW
  int cblocks;              // 32-channel k-blocks per tap (channels of the LOADED tensor / 32)
  int w_rows;               // data gradient: rows of the weight matrix (Cout); its row pitch is taps * K floats
  int chans;                // channels of the loaded tensor (weight gradient: groups beyond K load channel `chans` = zeros)
};
__device__ __forceinline__ void conv_pixel(const ConvCoords& c, int m, int smp, int* w, int* h, int* n) {
  const int img = m / c.P, rem = m - img * c.P;
  const int ph = rem / c.OW, q = rem - ph * c.OW;
  *w = q * c.sw + c.lw;
  *h = ph * c.sh + c.lh;
  *n = smp * c.imgs + img;
}

// Heterogeneous tile list of the CTA-pair kernel (narrow layers whose uniform 1024-row tiles need between one and two
// waves): a sample's rows are cut into `a` tiles of 8 row blocks followed by `b` tiles of 6; samples [0, s1) use (a1, b1),
// the others (a2, b2).  The grid lists every 8-block tile first, then the 6-block ones, so that a pair slot runs one of
// each instead of two full ones (C3 conv forward: 74 + 72 tiles on 74 slots: 1.75 wave-equivalents instead of 2).
struct TilePlan {
  int on;
  int n_a;                  // number of 8-block tiles of the launch
  int s1, a1, b1, a2, b2;
};

struct TmaContractParams {
  TilePlan plan;
  CUtensorMap map_l;        // fwd: activations [S or 1][M][K]; dgrad: dY [S][M][N]; conv: the NHWC tensor (im2col map)
  ConvCoords conv;
  int64_t w_numel;          // elements of the weight tensor (injected eps: stride between samples)
  View out;                 // fwd: y view; dgrad: da as a row-major view
  int64_t out_sample_stride;
  const float* mu_w;
  const float* sigma_w;
  const float* eps_w;
  const float* mu_b;
  const float* sigma_b;
  const float* eps_b;
  int M, N, K, S;
  uint32_t sample_begin;
  bnn_rng rng_w, rng_b;
  int shared_l;             // all samples read sample 0 of the L operand (shared activations)
  int sum_samples;          // dgrad with shared activations: one output, summed over the samples
  int z_per;                // sum_samples: samples per CTA (grid.z = sample groups); more than one group -> the partial
  int atomic_out;           //   sums are ADDED to the (zeroed) output
  int vec_out;
  int exp_flags;            // profiling builds: 1 = skip weight generation, 2 = skip TMA loads, 4 = skip MMA issue
  // balanced schedule (contract_pair_sk_kernel): a line of sk_total items cut into sk_slots equal ranges, one per
  // resident CTA pair.  sk_msplit: items = 256-row units of the (sample, column tile) columns, sk_mtiles units each;
  // else items = k-block iterations of the 1024-row tiles (sk_its each; sk_mtiles tiles per column)
  int sk_slots;             // > 0: pair slots (clusters) of the launch
  int sk_msplit;
  int sk_mtiles, sk_gx;     // column = z * sk_gx + y  (y: 128-column tile, z: sample)
  int sk_its;               // k-block iterations per tile (reduction blocks x samples summed into the tile)
  int sk_total;
};

struct TmaPipe {
  uint64_t* full_a;      // [2]        TMA transaction barriers: the MB activation tiles of one k-block group
  uint64_t* full_w;      // [kWSlots]  one arrival per warp of the generator group that owns the slot
  uint64_t* empty_w;     // [kWSlots]  "k-block consumed": one tcgen05.commit per k-block
  uint64_t* accum_full;  // accumulators complete
  uint32_t* tmem_slot;
  float* aux;            // 128 floats: the sampled bias row of this CTA's columns
  uint32_t ring_a, ring_w;
};

__device__ __forceinline__ TmaPipe carve_tma(uint8_t* smem_raw) {
  TmaPipe p;
  p.full_a = reinterpret_cast<uint64_t*>(smem_raw);
  p.full_w = p.full_a + 2;
  p.empty_w = p.full_w + kWSlots;
  p.accum_full = p.empty_w + kWSlots;
  p.tmem_slot = reinterpret_cast<uint32_t*>(p.accum_full + 1);
  p.aux = reinterpret_cast<float*>(smem_raw + 512);
  const uint32_t base = smem_u32(smem_raw) + kSmemAux;
  p.ring_a = (base + 1023u) & ~1023u;
  p.ring_w = p.ring_a + kASlots * kTileBytes;
  return p;
}

// W_s = mu + sigma * eps_s for a (kRows x 32) piece of the weight matrix, written as an operand tile by ONE generator
// group (kGroupThreads threads, four float4 per thread and trip: all loads first, then the Philox chains).
//   kMnMajor = false: K-major tile, tile rows = n (kRows of them), columns = k          (forward)
//   kMnMajor = true : MN-major tile, K-rows = n (32), MN = k (kRows of them)            (data gradient)
// Element (n, k) of the weight matrix lives at n * ldw + woff + k (ldw = K, woff = 0 for a plain [N][K] matrix; the
// conv data gradient walks the (o, kh, kw, c) tensor with n = o, ldw = taps * C, woff = tap * C, k = c).
template <int kRows, bool kMnMajor, int kSigns>
__device__ __forceinline__ void gen_w_tile(uint32_t tile, const float* __restrict__ mu, const float* __restrict__ sigma,
                                           const EpsSrc& eps, int n0, int N, int k0, int K, int tid, int64_t ldw,
                                           int64_t woff) {
  constexpr int kItems = kRows * 8;                       // float4 items
  constexpr int kTrips = kItems / (4 * kGroupThreads);
  static_assert(kItems % (4 * kGroupThreads) == 0, "tile does not divide over the generator group");
#pragma unroll 1
  for (int trip = 0; trip < kTrips; ++trip) {
    float4 m[4], sg[4];
    int64_t idx[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int item = (trip * 4 + u) * kGroupThreads + tid;
      int n, k;
      if (!kMnMajor) { n = n0 + (item >> 3); k = k0 + ((item & 7) << 2); }
      else { n = n0 + item / (kRows / 4); k = k0 + ((item % (kRows / 4)) << 2); }
      idx[u] = -1;
      if (n < N && k < K) {
        idx[u] = static_cast<int64_t>(n) * ldw + woff + k;
        m[u] = __ldg(reinterpret_cast<const float4*>(mu + idx[u]));
        sg[u] = __ldg(reinterpret_cast<const float4*>(sigma + idx[u]));
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int item = (trip * 4 + u) * kGroupThreads + tid;
      float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx[u] >= 0) {
        const float4 e = eps_vec4<kSigns>(eps, idx[u]);
        w.x = fmaf(sg[u].x, e.x, m[u].x);
        w.y = fmaf(sg[u].y, e.y, m[u].y);
        w.z = fmaf(sg[u].z, e.z, m[u].z);
        w.w = fmaf(sg[u].w, e.w, m[u].w);
      }
      uint32_t off;
      if (!kMnMajor) off = tile_offset(item >> 3, item & 7);
      else off = tile_offset_mn((item % (kRows / 4)) << 2, item / (kRows / 4), kMnLbo, kMnSbo);
      sts128(tile + off, to_tf32(w.x), to_tf32(w.y), to_tf32(w.z), to_tf32(w.w));
    }
  }
}

// ---------------------------------------------------------------------------------------------- staged epilogue
// A TMEM lane is an output row, so an epilogue thread that stores its own 16-column chunk writes 16 bytes of 32
// different rows per warp instruction: 32 memory wavefronts per 512 bytes (measured: 25 k cycles to drain the 512 x 128
// accumulator of one CTA, a quarter of a C3 conv tile's time).  Each epilogue warp therefore passes its 32 rows through
// a private shared-memory buffer, 64 columns at a time, and writes whole row segments:
//   row-major output (view P == 1): staging rows of 64 + 4 floats (float4 writes of a quarter warp hit 32 different
//     banks); two rows per store instruction, 256 contiguous bytes each;
//   NCHW output (P = OH*OW, P % 4 == 0): staging [column][32 rows + 4]; four rows of a lane's group are four consecutive
//     pixels of one image, i.e. one aligned float4 of out[img][n][p..p+3]; a store instruction covers 4 columns x 8 row
//     groups: for P = 16 two runs of 256 contiguous bytes (was: two runs of 64 bytes per scalar store).
// kStageWarpBytes per warp; the uniform-grid kernels alias the operand ring (dead once the accumulator is complete), the
// balanced schedule keeps its own region because its producers run ahead into the next segment.
constexpr int kStageWarpBytes = 64 * 36 * 4;     // 9216: [64 columns][32 + 4] floats; the row-major form needs 32 x 68 x 4 = 8704
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ bool stage_eligible(const View& out, bool vec_out, int cols_here) {
  if ((cols_here & 3) != 0 || (reinterpret_cast<uintptr_t>(out.base) & 15) != 0 || (out.bs & 3) != 0) return false;
  return out.P == 1 ? vec_out : (out.P & 3) == 0;
}
// Halves [h_begin, h_end) (64 columns each) of one 128-row block, drained by one warp (TMEM lanes [32 quad, +32), rows
// m0 .. m0 + 31 of the output, columns col0 .. col0 + cols_here): two tcgen05.ld per round trip, then (+ bias) ->
// staging -> coalesced stores / adds.
template <bool kBias>
__device__ __forceinline__ void drain_block_staged(uint32_t taddr, uint32_t stage, const float* aux, const View& out,
                                                   int m0, int M, int col0, int cols_here, int lane, bool add,
                                                   int h_begin = 0, int h_end = 2) {
  const bool nchw = out.P != 1;
  // per-lane destination of the write-out phase
  float* dst;
  bool ok;
  if (!nchw) {
    dst = out.base + static_cast<int64_t>(m0 + (lane >> 4)) * out.bs + col0 + ((lane & 15) << 2);
    ok = true;
  } else {
    const int m = m0 + ((lane & 7) << 2);                  // first of the lane's four rows (one image: P % 4 == 0)
    const int b = m / out.P, px = m - b * out.P;
    dst = out.base + static_cast<int64_t>(b) * out.bs + static_cast<int64_t>(col0 + (lane >> 3)) * out.P + px;
    ok = m < M;
  }
  for (int h = h_begin; h < h_end && h * 64 < cols_here; ++h) {
#pragma unroll 1
    for (int pr = 0; pr < 2 && (h * 4 + pr * 2) * 16 < cols_here; ++pr) {      // two 16-column chunks per TMEM round trip
      uint32_t r[2][16];
      tmem_ld16_nowait(taddr + (h * 4 + pr * 2) * 16, r[0]);
      if ((h * 4 + pr * 2 + 1) * 16 < cols_here) tmem_ld16_nowait(taddr + (h * 4 + pr * 2 + 1) * 16, r[1]);
      tmem_ld_wait();
#pragma unroll
      for (int q2 = 0; q2 < 2; ++q2) {
        const int cc = pr * 2 + q2;
        if ((h * 4 + cc) * 16 >= cols_here) break;
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[q2][j]);
        if (kBias) {
          const float4* a4 = reinterpret_cast<const float4*>(aux + (h * 4 + cc) * 16);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 q = a4[j];
            v[4 * j] += q.x; v[4 * j + 1] += q.y; v[4 * j + 2] += q.z; v[4 * j + 3] += q.w;
          }
        }
        if (!nchw) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            sts128(stage + lane * 272 + cc * 64 + j * 16, __float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
                   __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) sts32(stage + ((cc * 16 + j) * 36 + lane) * 4, __float_as_uint(v[j]));
        }
      }
    }
    __syncwarp();
    const int cols_half = cols_here - h * 64 < 64 ? cols_here - h * 64 : 64;
    if (!nchw) {
      const int c4 = lane & 15;
#pragma unroll 8
      for (int i = 0; i < 16; ++i) {
        const int row = 2 * i + (lane >> 4);
        if (c4 * 4 < cols_half && m0 + row < M) {
          const float4 q = lds128(stage + row * 272 + c4 * 16);
          float* d = dst + static_cast<int64_t>(2 * i) * out.bs + h * 64;
          if (add) red_add4(d, q.x, q.y, q.z, q.w);
          else *reinterpret_cast<float4*>(d) = q;
        }
      }
    } else {
      const int rg = lane & 7, ns = lane >> 3;
#pragma unroll 8
      for (int i = 0; i < 16; ++i) {
        const int n = i * 4 + ns;
        if (n < cols_half && ok) {
          const float4 q = lds128(stage + (n * 36 + rg * 4) * 4);
          float* d = dst + static_cast<int64_t>(h * 64 + i * 4) * out.P;
          if (add) red_add4(d, q.x, q.y, q.z, q.w);
          else *reinterpret_cast<float4*>(d) = q;
        }
      }
    }
    __syncwarp();
  }
}
// The accumulator of one CTA (mb_count row blocks of 128 rows x cols_here columns) drained by kDrainPerQuad warps per
// TMEM lane quarter — the epilogue warp of the quarter (idx 0) and generator warps with the same warp % 4, idle once
// their loop has ended: work items = (row block, 64-column half), item i goes to warp i % kDrainPerQuad.
constexpr int kDrainPerQuad = 4;
constexpr int kDrainThreads = 4 * kDrainPerQuad * 32;
constexpr int kDrainBarrier = 2;
template <bool kBias>
__device__ __forceinline__ void drain_tile_shared(uint32_t tmem, uint32_t stage_base, const float* aux, const View& out,
                                                  int row0, int M, int mb_count, int col0, int cols_here, int quad,
                                                  int idx, int lane, bool add) {
  const uint32_t stage = stage_base + static_cast<uint32_t>(idx * 4 + quad) * kStageWarpBytes;
  const int halves = (cols_here + 63) / 64;
  for (int item = idx; item < mb_count * halves; item += kDrainPerQuad) {
    const int mb = item / halves, h = item - mb * halves;
    if (row0 + mb * 128 >= M) break;
    drain_block_staged<kBias>(tmem + (static_cast<uint32_t>(quad * 32) << 16) + mb * 128, stage, aux, out,
                              row0 + mb * 128 + quad * 32, M, col0, cols_here, lane, add, h, h + 1);
  }
}

template <int MB, bool kDgrad, int kMode>      // kMode: 0 plain, 1 conv (im2col maps), 2 rank-one sign noise (Flipout)
__global__ void __launch_bounds__(kThreadsTma, 1) contract_tma_kernel(const __grid_constant__ TmaContractParams p) {
  constexpr bool kConv = kMode == 1;
  constexpr int kSigns = kMode == 2 ? 1 : 0;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  long long w_kernel = 0, w_mma_w = 0, w_mma_a = 0, w_gen = 0, w_tma = 0;
  (void)w_kernel; (void)w_mma_w; (void)w_mma_a; (void)w_gen; (void)w_tma;
#ifdef BNN_PROFILE_WAITS
  const long long t_kernel0 = clock64();
#endif
  constexpr uint32_t kTmemCols = tmem_cols_pow2(MB * 128);
  const TmaPipe pipe = carve_tma(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    // full_a[g]: the MB activation tiles of k-block group g (two groups of MB slots); full_w[j]: generated weight
    // tile j (one arrival per generator warp); empty_w[j] = "k-block consumed": ONE tcgen05.commit per k-block frees
    // weight slot j for the generators and, two k-blocks later, activation group j & 1 for the TMA thread
    for (int i = 0; i < 2; ++i) mbar_init(pipe.full_a + i, 1);
    for (int i = 0; i < kWSlots; ++i) { mbar_init(pipe.full_w + i, kGroupWarps); mbar_init(pipe.empty_w + i, 1); }
    mbar_init(pipe.accum_full, 1);
    fence_mbar_init();
  }
  if (warp == kTmaWarp && lane == 0) tma_prefetch_desc(&p.map_l);
  if (warp == kMmaWarpT) tmem_alloc(pipe.tmem_slot, kTmemCols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(pipe.tmem_slot);

  const int col0 = blockIdx.x * 128;                   // output columns (n for fwd, k for dgrad)
  const int row0 = blockIdx.y * (MB * 128);            // output rows (m)
  const int n_cols = kDgrad ? p.K : p.N;
  const int n_red = kDgrad ? p.N : p.K;
  const int s_begin = p.sum_samples ? static_cast<int>(blockIdx.z) * p.z_per : blockIdx.z;
  const int s_end = p.sum_samples ? (s_begin + p.z_per < p.S ? s_begin + p.z_per : p.S) : blockIdx.z + 1;
  const int red_blocks = (n_red + kBK - 1) / kBK;
  int mb_used = (p.M - row0 + 127) / 128;
  if (mb_used > MB) mb_used = MB;

  // the accumulator is drained by the epilogue warps AND, once their loop has ended, three generator warps per TMEM lane
  // quarter (staged, coalesced write-out through the operand rings, which are dead by then); uniform per CTA
  View out_tile = p.out;
  out_tile.base += (p.sum_samples ? 0 : static_cast<int64_t>(blockIdx.z) * p.out_sample_stride);
  const int cols_tile = n_cols - col0 < 128 ? n_cols - col0 : 128;
  const bool staged = stage_eligible(out_tile, p.vec_out != 0, cols_tile) && (out_tile.P == 1 || (p.M & 3) == 0);

  if (warp < kGenWarps) {
    // ------------------------------------------------------------------ weight generators (four groups, slot = group)
    const int group = warp / kGroupWarps, tid = threadIdx.x - group * kGroupThreads;
    const RngKey key = resolve_rng(p.rng_w);
    const int total = red_blocks * (s_end - s_begin);
    for (int it = group; it < total; it += kGenGroups) {
      const int s = s_begin + it / red_blocks, rb = it - (it / red_blocks) * red_blocks;
      EpsSrc eps;
      eps.inj = p.eps_w ? p.eps_w + static_cast<int64_t>(s) * p.w_numel : nullptr;
      eps.key = key;
      eps.sample = p.sample_begin + s;
      { BNN_T0(); mbar_wait(pipe.empty_w + group, ((it >> 2) & 1) ^ 1); BNN_ACC(w_gen); }
      const uint32_t tile = pipe.ring_w + group * kTileBytes;
#ifdef BNN_PROFILE_WAITS
      if (p.exp_flags & 1) { /* skip */ } else
#endif
      if (!kDgrad)
        gen_w_tile<128, false, kSigns>(tile, p.mu_w, p.sigma_w, eps, col0, p.N, rb * kBK, p.K, tid, p.K, 0);
      else if (!kConv)
        gen_w_tile<128, true, kSigns>(tile, p.mu_w, p.sigma_w, eps, rb * kBK, p.N, col0, p.K, tid, p.K, 0);
      else {          // conv data gradient: k-block = (flipped tap, 32 output channels o0 ..)
        const int tapf = rb / p.conv.cblocks, o0 = (rb - tapf * p.conv.cblocks) * kBK;
        gen_w_tile<128, true, kSigns>(tile, p.mu_w, p.sigma_w, eps, o0, p.conv.w_rows, col0, p.K, tid,
                              static_cast<int64_t>(p.conv.taps) * p.K, static_cast<int64_t>(p.conv.taps - 1 - tapf) * p.K);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(pipe.full_w + group);
    }
    if (staged && (warp >> 2) < kDrainPerQuad - 1) {     // join the drain (TMEM lane quarter = warp % 4)
      if (!kDgrad) named_bar_sync(kDrainBarrier, kDrainThreads);          // the bias row is in shared memory
      mbar_wait(pipe.accum_full, 0);
      tc_fence_after_sync();
      drain_tile_shared<!kDgrad>(tmem, pipe.ring_a, pipe.aux, out_tile, row0, p.M, mb_used, col0, cols_tile, warp & 3,
                                 1 + (warp >> 2), lane, kDgrad && p.atomic_out);
    }
  } else if (warp == kTmaWarp) {
    // ------------------------------------------------------------------ TMA issuer
    if (lane == 0) {
      int it = 0;
      for (int s = s_begin; s < s_end; ++s) {
        const int smp = p.shared_l ? 0 : s;
        int cw[MB], ch[MB], cn[MB];                // conv: first pixel of each row block (fixed over the k-blocks)
        if (kConv) {
#pragma unroll
          for (int mb = 0; mb < MB; ++mb) conv_pixel(p.conv, row0 + mb * 128, smp, &cw[mb], &ch[mb], &cn[mb]);
        }
        int tap = 0, cb = 0;                       // conv: k-block rb = (tap, 32-channel block cb), walked incrementally
        for (int rb = 0; rb < red_blocks; ++rb, ++it) {
          const int g = it & 1;
          if (it >= 2) { BNN_T0(); mbar_wait(pipe.empty_w + ((it - 2) & 3), ((it - 2) >> 2) & 1); BNN_ACC(w_tma); }
#ifdef BNN_PROFILE_WAITS
          if (p.exp_flags & 2) { mbar_arrive(pipe.full_a + g); continue; }
#endif
          mbar_arrive_expect_tx(pipe.full_a + g, mb_used * kTileBytes);
          if (!kConv) {
            for (int mb = 0; mb < mb_used; ++mb)
              tma_load_3d(pipe.ring_a + (g * MB + mb) * kTileBytes, &p.map_l, rb * kBK, row0 + mb * 128, smp, pipe.full_a + g);
          } else {
            const int kh = tap / p.conv.KW, kw = tap - kh * p.conv.KW;
#pragma unroll
            for (int mb = 0; mb < MB; ++mb)
              if (mb < mb_used)
                tma_load_im2col_4d(pipe.ring_a + (g * MB + mb) * kTileBytes, &p.map_l, cb * kBK, cw[mb], ch[mb], cn[mb],
                                   static_cast<uint16_t>(kw * p.conv.dw), static_cast<uint16_t>(kh * p.conv.dh), pipe.full_a + g);
            if (++cb == p.conv.cblocks) { cb = 0; ++tap; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == kMmaWarpT) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32(128, mma_n(n_cols - col0), false, kDgrad);
      const uint64_t desc_a0 = make_smem_desc(pipe.ring_a);
      const uint64_t desc_w0 = kDgrad ? make_smem_desc_mn(pipe.ring_w, kMnLbo, kMnSbo) : make_smem_desc(pipe.ring_w);
      int it = 0;
      for (int s = s_begin; s < s_end; ++s) {
        for (int rb = 0; rb < red_blocks; ++rb, ++it) {
          const int wslot = it & 3, g = it & 1;
          { BNN_T0(); mbar_wait(pipe.full_w + wslot, (it >> 2) & 1); BNN_ACC(w_mma_w); }
          { BNN_T0(); mbar_wait(pipe.full_a + g, (it >> 1) & 1); BNN_ACC(w_mma_a); }
          tc_fence_after_sync();
          const uint64_t da0 = desc_advance(desc_a0, static_cast<uint32_t>(g * MB) * (kTileBytes >> 4));
          const uint64_t db0 = desc_advance(desc_w0, static_cast<uint32_t>(wslot) * (kTileBytes >> 4));
#ifdef BNN_PROFILE_WAITS
          if (!(p.exp_flags & 4))
#endif
          for (int mb = 0; mb < mb_used; ++mb) {
#pragma unroll
            for (int ks = 0; ks < kBK / 8; ++ks)
              mma_tf32(tmem + mb * 128, desc_advance(da0, mb * (kTileBytes >> 4) + ks * 2),
                       desc_advance(db0, ks * ((kDgrad ? kMnKStep : 32u) >> 4)), idesc, it > 0 || ks > 0);
          }
          mma_commit(pipe.empty_w + wslot);           // the one commit of this k-block
        }
      }
      mma_commit(pipe.accum_full);
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue
    const int et = threadIdx.x - kEpiWarp0T * 32;   // 0..127
    const int quad = warp & 3;
    if (!kDgrad) {
      float b = 0.f;
      const int n = col0 + et;
      if (p.mu_b != nullptr && n < p.N) {
        const int s = blockIdx.z;
        const float e = p.eps_b ? __ldg(p.eps_b + static_cast<int64_t>(s) * p.N + n)
                                : eps1(resolve_rng(p.rng_b), p.sample_begin + s, static_cast<uint64_t>(n));
        b = fmaf(__ldg(p.sigma_b + n), e, __ldg(p.mu_b + n));
      }
      pipe.aux[et] = b;
      named_bar_sync(kEpiBarrier, kEpiThreads);
      if (staged) named_bar_sync(kDrainBarrier, kDrainThreads);
    }
    mbar_wait(pipe.accum_full, 0);
    tc_fence_after_sync();
    if (staged) {
      drain_tile_shared<!kDgrad>(tmem, pipe.ring_a, pipe.aux, out_tile, row0, p.M, mb_used, col0, cols_tile, quad, 0, lane,
                                 kDgrad && p.atomic_out);
    } else {
      const View& out = out_tile;
      const int cols_here = cols_tile;
      for (int mb = 0; mb < mb_used; ++mb) {
        const int m = row0 + mb * 128 + quad * 32 + lane;
        for (int c = 0; c * 16 < cols_here; ++c) {
          float v[16];
          tmem_ld16(tmem + (static_cast<uint32_t>(quad * 32) << 16) + mb * 128 + c * 16, v);
          if (!kDgrad) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += pipe.aux[c * 16 + j];
          }
          if (m < p.M) {
            if (kDgrad && p.atomic_out) add_chunk(out, m, col0 + c * 16, n_cols, v, p.vec_out != 0);
            else store_chunk(out, m, col0 + c * 16, n_cols, v, p.vec_out != 0);
          }
        }
      }
    }
  }
#ifdef BNN_PROFILE_WAITS
  {
    const int w_ = threadIdx.x >> 5, l_ = threadIdx.x & 31;
    if (l_ == 0) {
      if (w_ == 0) { atomicAdd(&g_wait_cycles[3], (unsigned long long)w_gen); atomicAdd(&g_wait_cycles[0], (unsigned long long)(clock64() - t_kernel0)); atomicAdd(&g_wait_cycles[5], 1ull); }
      if (w_mma_w | w_mma_a) { atomicAdd(&g_wait_cycles[1], (unsigned long long)w_mma_w); atomicAdd(&g_wait_cycles[2], (unsigned long long)w_mma_a); }
      if (w_tma) atomicAdd(&g_wait_cycles[4], (unsigned long long)w_tma);
    }
  }
#endif
  tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarpT) {
    tc_fence_after_sync();
    tmem_dealloc(tmem, kTmemCols);
  }
}

constexpr size_t kContractSmem = kSmemAux + 1024 + static_cast<size_t>(kASlots + kWSlots) * kTileBytes;

// The conv variants (im2col tensor maps, tap-major weight rows) are separate instantiations: folded into one kernel, the
// conv state kept live across the generator loop pushed every input-gradient variant over its register budget (116-152
// bytes of spill stores in the hot loop: C4 input gradient 2.33 -> 2.86 ms).
template <int MB, bool kDgrad, int kMode>
int launch_tma_contract_v(const TmaContractParams& p, dim3 grid, cudaStream_t st) {
  static SmemOptIn opt_in;
  const int rc = allow_dynamic_smem(contract_tma_kernel<MB, kDgrad, kMode>, kContractSmem, &opt_in);
  if (rc != BNN_OK) return rc;
  contract_tma_kernel<MB, kDgrad, kMode><<<grid, kThreadsTma, kContractSmem, st>>>(p);
  BNN_CUDA_OK(cudaGetLastError());
  return BNN_OK;
}
template <int MB, bool kDgrad>
int launch_tma_contract(const TmaContractParams& p, dim3 grid, cudaStream_t st) {
  if (p.conv.on) return launch_tma_contract_v<MB, kDgrad, 1>(p, grid, st);
  if (p.rng_w.row_sign != nullptr) return launch_tma_contract_v<MB, kDgrad, 2>(p, grid, st);
  return launch_tma_contract_v<MB, kDgrad, 0>(p, grid, st);
}

// ---------------------------------------------------------------------------------------------- CTA-pair forward / dgrad
// Same contraction, two SMs per tile (cta_group::2): the pair computes 2 x MB x 128 output rows against ONE 128-column
// weight tile of which each CTA generates only its half (64 n-rows forward / 64 k-columns dgrad) — half the Philox work
// per SM and per MMA cycle.  Each CTA loads its own activation tiles by TMA (transaction bytes land on the LEADER's
// barrier), the generator warps of the peer arrive on the leader's barrier through the cluster address space, the
// leader's single MMA thread issues tcgen05.mma.cta_group::2 (M = 256) and frees the slots of BOTH CTAs with
// multicast commits.  Epilogue: every CTA drains its own TMEM.
constexpr int kPairWSlots = 4;               // 8 KiB half-tiles: slot = k-block % kPairWSlots (a power of two, a multiple of
constexpr int kPairWShift = 2;               // the four generator groups).  Eight slots were measured: the generators wait
                                             // less (19 k instead of 26 k cycles per C3 tile), the MMA thread's 8.8 k cycles of
                                             // waiting for weights do not move — they are the cold start of a tile (first
                                             // mu / sigma loads, instruction cache), not a lack of slack in the ring
constexpr int kSkWSlots = 4;                 // balanced schedule: four slots (its epilogue staging needs the space)
constexpr int kHalfTileBytes = kTileBytes / 2;
constexpr size_t kPairSmem = kSmemAux + 1024 + static_cast<size_t>(kASlots) * kTileBytes + kPairWSlots * kHalfTileBytes;

struct PairPipe {
  uint64_t* full_a;      // [2]  leader: 1 arrival (expect_tx) + the bytes of both CTAs' tiles of one k-block group
  uint64_t* full_w;      // [kPairWSlots]  leader: one arrival per warp of the owning generator group of both CTAs
  uint64_t* empty_w;     // [kPairWSlots]  both CTAs: "k-block consumed", one multicast tcgen05.commit per k-block
  uint64_t* accum_full;  // both CTAs
  uint32_t* tmem_slot;
  float* aux;
  uint32_t ring_a, ring_w;
};

__device__ __forceinline__ PairPipe carve_pair(uint8_t* smem_raw) {
  PairPipe p;
  p.full_a = reinterpret_cast<uint64_t*>(smem_raw);
  p.full_w = p.full_a + 2;
  p.empty_w = p.full_w + kPairWSlots;
  p.accum_full = p.empty_w + kPairWSlots;
  p.tmem_slot = reinterpret_cast<uint32_t*>(p.accum_full + 1);
  p.aux = reinterpret_cast<float*>(smem_raw + 512);
  const uint32_t base = smem_u32(smem_raw) + kSmemAux;
  p.ring_a = (base + 1023u) & ~1023u;
  p.ring_w = p.ring_a + kASlots * kTileBytes;
  return p;
}

template <int MB, bool kDgrad, int kMode>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreadsTma, 1)
contract_pair_kernel(const __grid_constant__ TmaContractParams p) {
  constexpr bool kConv = kMode == 1;
  constexpr int kSigns = kMode == 2 ? 1 : 0;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  long long w_kernel = 0, w_mma_w = 0, w_mma_a = 0, w_gen = 0, w_tma = 0;
  (void)w_kernel; (void)w_mma_w; (void)w_mma_a; (void)w_gen; (void)w_tma;
#ifdef BNN_PROFILE_WAITS
  const long long t_kernel0 = clock64();
  if (threadIdx.x == 0) {
    unsigned long long g;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
    atomicMin(&g_stage_cycles[6], g);
  }
#endif
  constexpr uint32_t kTmemCols = tmem_cols_pow2(MB * 128);
  const PairPipe pipe = carve_pair(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();          // 0 = leader
  if (threadIdx.x == 0) {
    // leader: full_a[g] = activation tiles of k-block group g of BOTH CTAs (1 arrival + bytes), full_w[j] = both halves of
    // weight tile j (one arrival per generator warp of both CTAs); both CTAs: empty_w[j] = "k-block consumed", ONE
    // multicast tcgen05.commit per k-block (frees weight slot j, and activation group j & 1 two k-blocks later)
    for (int i = 0; i < 2; ++i) mbar_init(pipe.full_a + i, 1);
    for (int i = 0; i < kPairWSlots; ++i) { mbar_init(pipe.full_w + i, 2 * kGroupWarps); mbar_init(pipe.empty_w + i, 1); }
    mbar_init(pipe.accum_full, 1);
    fence_mbar_init();
  }
  if (warp == kTmaWarp && lane == 0) tma_prefetch_desc(&p.map_l);
  if (warp == kMmaWarpT) tmem_alloc_pair(pipe.tmem_slot, kTmemCols);
  tc_fence_before_sync();
  cluster_sync_all();                                // barriers of both CTAs initialised before any remote arrive
  tc_fence_after_sync();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(pipe.tmem_slot);
#ifdef BNN_PROFILE_WAITS
  if (threadIdx.x == 0) { BNN_STAGE(0, t_kernel0); atomicAdd(&g_stage_cycles[5], 1ull); }
#endif

  const int col0 = blockIdx.y * 128;                   // output columns of the PAIR (n for fwd, k for dgrad)
  // the pair's tile: sample, first row of the leader, row blocks per CTA (uniform grid: MB; tile plan: 4 or 3)
  int tile_s = blockIdx.z, lead_row0 = (blockIdx.x & ~1u) * (MB * 128), mb_cap = MB;
  if (MB == 4 && p.plan.on) {
    const TilePlan& pl = p.plan;
    const int t = blockIdx.x >> 1;
    if (t < pl.n_a) {
      const int first = pl.s1 * pl.a1;
      int j;
      if (t < first) { tile_s = t / pl.a1; j = t - tile_s * pl.a1; }
      else { const int u = t - first; tile_s = pl.s1 + u / pl.a2; j = u - (u / pl.a2) * pl.a2; }
      lead_row0 = j * (8 * 128);
    } else {
      const int v = t - pl.n_a, first = pl.s1 * pl.b1;
      int j, a;
      if (v < first) { tile_s = v / pl.b1; j = v - tile_s * pl.b1; a = pl.a1; }
      else { const int u = v - first; tile_s = pl.s1 + u / pl.b2; j = u - (u / pl.b2) * pl.b2; a = pl.a2; }
      lead_row0 = (a * 8 + j * 6) * 128;
      mb_cap = 3;
    }
  }
  const int row0 = lead_row0 + static_cast<int>(blockIdx.x & 1u) * (mb_cap * 128);      // output rows of THIS CTA
  const int n_cols = kDgrad ? p.K : p.N;
  const int n_red = kDgrad ? p.N : p.K;
  const int s_begin = p.sum_samples ? static_cast<int>(blockIdx.z) * p.z_per : tile_s;
  const int s_end = p.sum_samples ? (s_begin + p.z_per < p.S ? s_begin + p.z_per : p.S) : tile_s + 1;
  const int red_blocks = (n_red + kBK - 1) / kBK;
  // both CTAs walk the same number of M-blocks (the leader holds the lower rows, so its count is the larger one);
  // tiles beyond M are zero-filled by TMA and never stored
  int mb_pair = (p.M - lead_row0 + 127) / 128;
  if (mb_pair > mb_cap) mb_pair = mb_cap;

  // the accumulator is drained by the epilogue warps AND, once their loop has ended, three generator warps per TMEM lane
  // quarter (staged, coalesced write-out through the operand rings, which are dead by then); uniform per CTA
  View out_tile = p.out;
  out_tile.base += (p.sum_samples ? 0 : static_cast<int64_t>(tile_s) * p.out_sample_stride);
  const int cols_tile = n_cols - col0 < 128 ? n_cols - col0 : 128;
  const bool staged = stage_eligible(out_tile, p.vec_out != 0, cols_tile) && (out_tile.P == 1 || (p.M & 3) == 0);

  if (warp < kGenWarps) {
    // ------------------------------------------------------------------ weight generators (half tile per CTA, four groups)
    const int group = warp / kGroupWarps, tid = threadIdx.x - group * kGroupThreads;
    const RngKey key = resolve_rng(p.rng_w);
    const int half0 = col0 + static_cast<int>(rank) * 64;
    const int total = red_blocks * (s_end - s_begin);
    const uint32_t lead_full_w0 = mapa_u32(smem_u32(pipe.full_w), 0);
    for (int it = group; it < total; it += kGenGroups) {
      const int wslot = it & (kPairWSlots - 1);
      const int s = s_begin + it / red_blocks, rb = it - (it / red_blocks) * red_blocks;
      EpsSrc eps;
      eps.inj = p.eps_w ? p.eps_w + static_cast<int64_t>(s) * p.w_numel : nullptr;
      eps.key = key;
      eps.sample = p.sample_begin + s;
      { BNN_T0(); mbar_wait(pipe.empty_w + wslot, ((it >> kPairWShift) & 1) ^ 1); BNN_ACC(w_gen); }
      const uint32_t tile = pipe.ring_w + wslot * kHalfTileBytes;
#ifdef BNN_PROFILE_WAITS
      if (p.exp_flags & 2) eps.inj = p.sigma_w;   // elimination run: eps read from memory (any valid array): loads and stores, no Philox
      if (p.exp_flags & 1) { /* elimination run: barrier traffic only, no generation */ } else
#endif
      if (!kDgrad)
        gen_w_tile<64, false, kSigns>(tile, p.mu_w, p.sigma_w, eps, half0, p.N, rb * kBK, p.K, tid, p.K, 0);
      else if (!kConv)
        gen_w_tile<64, true, kSigns>(tile, p.mu_w, p.sigma_w, eps, rb * kBK, p.N, half0, p.K, tid, p.K, 0);
      else {
        const int tapf = rb / p.conv.cblocks, o0 = (rb - tapf * p.conv.cblocks) * kBK;
        gen_w_tile<64, true, kSigns>(tile, p.mu_w, p.sigma_w, eps, o0, p.conv.w_rows, half0, p.K, tid,
                             static_cast<int64_t>(p.conv.taps) * p.K, static_cast<int64_t>(p.conv.taps - 1 - tapf) * p.K);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(lead_full_w0 + wslot * 8);
    }
    if (staged && (warp >> 2) < kDrainPerQuad - 1) {     // join the drain (TMEM lane quarter = warp % 4)
      if (!kDgrad) named_bar_sync(kDrainBarrier, kDrainThreads);          // the bias row is in shared memory
      mbar_wait(pipe.accum_full, 0);
      tc_fence_after_sync();
      drain_tile_shared<!kDgrad>(tmem, pipe.ring_a, pipe.aux, out_tile, row0, p.M, mb_pair, col0, cols_tile, warp & 3,
                                 1 + (warp >> 2), lane, kDgrad && p.atomic_out);
    }
  } else if (warp == kTmaWarp) {
    // ------------------------------------------------------------------ TMA issuer (own rows; bytes counted by the leader)
    if (lane == 0) {
      int it = 0;
      const uint32_t lead_full_a = mapa_u32(smem_u32(pipe.full_a), 0);
      for (int s = s_begin; s < s_end; ++s) {
        const int smp = p.shared_l ? 0 : s;
        int cw[MB], ch[MB], cn[MB];
        if (kConv) {
#pragma unroll
          for (int mb = 0; mb < MB; ++mb) conv_pixel(p.conv, row0 + mb * 128, smp, &cw[mb], &ch[mb], &cn[mb]);
        }
        int tap = 0, cb = 0;
        for (int rb = 0; rb < red_blocks; ++rb, ++it) {
          const int g = it & 1;
          if (it >= 2) { BNN_T0(); mbar_wait(pipe.empty_w + ((it - 2) & (kPairWSlots - 1)), ((it - 2) >> kPairWShift) & 1); BNN_ACC(w_tma); }
          if (rank == 0) mbar_arrive_expect_tx(pipe.full_a + g, 2 * mb_pair * kTileBytes);
          if (!kConv) {
            for (int mb = 0; mb < mb_pair; ++mb)
              tma_load_3d_pair(pipe.ring_a + (g * MB + mb) * kTileBytes, &p.map_l, rb * kBK, row0 + mb * 128, smp,
                               lead_full_a + g * 8);
          } else {
            const int kh = tap / p.conv.KW, kw = tap - kh * p.conv.KW;
#pragma unroll
            for (int mb = 0; mb < MB; ++mb)
              if (mb < mb_pair)
                tma_load_im2col_4d_pair(pipe.ring_a + (g * MB + mb) * kTileBytes, &p.map_l, cb * kBK, cw[mb], ch[mb], cn[mb],
                                        static_cast<uint16_t>(kw * p.conv.dw), static_cast<uint16_t>(kh * p.conv.dh),
                                        lead_full_a + g * 8);
            if (++cb == p.conv.cblocks) { cb = 0; ++tap; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == kMmaWarpT) {
    // ------------------------------------------------------------------ MMA issuer (leader only)
    if (rank == 0 && lane == 0) {
      // N = 128 always: each CTA contributes its 64-row half of the weight tile (zero rows beyond the matrix edge)
      const uint32_t idesc = make_idesc_tf32(256, 128, false, kDgrad);
      const uint64_t desc_a0 = make_smem_desc(pipe.ring_a);
      const uint64_t desc_w0 = kDgrad ? make_smem_desc_mn(pipe.ring_w, kMnLbo, kMnSbo) : make_smem_desc(pipe.ring_w);
      int it = 0;
      for (int s = s_begin; s < s_end; ++s) {
        for (int rb = 0; rb < red_blocks; ++rb, ++it) {
          const int wslot = it & (kPairWSlots - 1), g = it & 1;
          { BNN_T0(); mbar_wait_cluster(pipe.full_w + wslot, (it >> kPairWShift) & 1); BNN_ACC(w_mma_w); }
#ifdef BNN_PROFILE_WAITS
          if (it == 0) atomicAdd(&g_wait_cycles[6], (unsigned long long)(clock64() - t_kernel0));    // first weight tile ready
          if (it == 4) atomicAdd(&g_wait_cycles[7], (unsigned long long)w_mma_w);                    // weight waits of k-blocks 0..4
#endif
          { BNN_T0(); mbar_wait_cluster(pipe.full_a + g, (it >> 1) & 1); BNN_ACC(w_mma_a); }
          tc_fence_after_sync();
          const uint64_t da0 = desc_advance(desc_a0, static_cast<uint32_t>(g * MB) * (kTileBytes >> 4));
          const uint64_t db0 = desc_advance(desc_w0, static_cast<uint32_t>(wslot) * (kHalfTileBytes >> 4));
          for (int mb = 0; mb < mb_pair; ++mb) {
#pragma unroll
            for (int ks = 0; ks < kBK / 8; ++ks)
              mma_tf32_pair(tmem + mb * 128, desc_advance(da0, mb * (kTileBytes >> 4) + ks * 2),
                            desc_advance(db0, ks * ((kDgrad ? kMnKStep : 32u) >> 4)), idesc, it > 0 || ks > 0);
          }
          mma_commit_pair(pipe.empty_w + wslot, 3);   // the one (multicast) commit of this k-block
        }
      }
      mma_commit_pair(pipe.accum_full, 3);
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (own TMEM, own rows)
    const int et = threadIdx.x - kEpiWarp0T * 32;   // 0..127
    const int quad = warp & 3;
    if (!kDgrad) {
      float b = 0.f;
      const int n = col0 + et;
      if (p.mu_b != nullptr && n < p.N) {
        const int s = tile_s;
        const float e = p.eps_b ? __ldg(p.eps_b + static_cast<int64_t>(s) * p.N + n)
                                : eps1(resolve_rng(p.rng_b), p.sample_begin + s, static_cast<uint64_t>(n));
        b = fmaf(__ldg(p.sigma_b + n), e, __ldg(p.mu_b + n));
      }
      pipe.aux[et] = b;
      named_bar_sync(kEpiBarrier, kEpiThreads);
      if (staged) named_bar_sync(kDrainBarrier, kDrainThreads);
    }
    mbar_wait(pipe.accum_full, 0);
    tc_fence_after_sync();
#ifdef BNN_PROFILE_WAITS
    if (threadIdx.x == kEpiWarp0T * 32) BNN_STAGE(2, t_kernel0);
#endif
    if (staged) {
      drain_tile_shared<!kDgrad>(tmem, pipe.ring_a, pipe.aux, out_tile, row0, p.M, mb_pair, col0, cols_tile, quad, 0, lane,
                                 kDgrad && p.atomic_out);
    } else {
      const View& out = out_tile;
      const int cols_here = cols_tile;
      for (int mb = 0; mb < mb_pair; ++mb) {
        const int m = row0 + mb * 128 + quad * 32 + lane;
        if (row0 + mb * 128 >= p.M) break;               // uniform per CTA
        for (int c = 0; c * 16 < cols_here; ++c) {
          float v[16];
          tmem_ld16(tmem + (static_cast<uint32_t>(quad * 32) << 16) + mb * 128 + c * 16, v);
          if (!kDgrad) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += pipe.aux[c * 16 + j];
          }
          if (m < p.M) {
            if (kDgrad && p.atomic_out) add_chunk(out, m, col0 + c * 16, n_cols, v, p.vec_out != 0);
            else store_chunk(out, m, col0 + c * 16, n_cols, v, p.vec_out != 0);
          }
        }
      }
    }
  }
#ifdef BNN_PROFILE_WAITS
  if (threadIdx.x == kEpiWarp0T * 32) BNN_STAGE(3, t_kernel0);
  if (threadIdx.x == 0) BNN_STAGE(1, t_kernel0);
  {
    const int w_ = threadIdx.x >> 5, l_ = threadIdx.x & 31;
    if (l_ == 0) {
      if (w_ == 0) { atomicAdd(&g_wait_cycles[3], (unsigned long long)w_gen); atomicAdd(&g_wait_cycles[0], (unsigned long long)(clock64() - t_kernel0)); atomicAdd(&g_wait_cycles[5], 1ull); }
      if (w_mma_w | w_mma_a) { atomicAdd(&g_wait_cycles[1], (unsigned long long)w_mma_w); atomicAdd(&g_wait_cycles[2], (unsigned long long)w_mma_a); }
      if (w_tma) atomicAdd(&g_wait_cycles[4], (unsigned long long)w_tma);
    }
  }
#endif
  tc_fence_before_sync();
  cluster_sync_all();                                // nobody leaves while the partner may still signal or read
  if (warp == kMmaWarpT) {
    tc_fence_after_sync();
    tmem_dealloc_pair(tmem, kTmemCols);
  }
#ifdef BNN_PROFILE_WAITS
  if (threadIdx.x == 0) {
    BNN_STAGE(4, t_kernel0);
    unsigned long long g;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
    atomicMax(&g_stage_cycles[7], g);
  }
#endif
}

template <int MB, bool kDgrad, int kMode>
int launch_pair_contract_v(const TmaContractParams& p, dim3 grid, cudaStream_t st) {
  static SmemOptIn opt_in;
  const int rc = allow_dynamic_smem(contract_pair_kernel<MB, kDgrad, kMode>, kPairSmem, &opt_in);
  if (rc != BNN_OK) return rc;
  contract_pair_kernel<MB, kDgrad, kMode><<<grid, kThreadsTma, kPairSmem, st>>>(p);
  BNN_CUDA_OK(cudaGetLastError());
  return BNN_OK;
}
template <int MB, bool kDgrad>
int launch_pair_contract(const TmaContractParams& p, dim3 grid, cudaStream_t st) {
  if (p.conv.on) return launch_pair_contract_v<MB, kDgrad, 1>(p, grid, st);
  if (p.rng_w.row_sign != nullptr) return launch_pair_contract_v<MB, kDgrad, 2>(p, grid, st);
  return launch_pair_contract_v<MB, kDgrad, 0>(p, grid, st);
}

// ---------------------------------------------------------------------------------------------- balanced schedule
// contract_pair_sk_kernel: the CTA-pair kernel as a PERSISTENT grid of `sk_slots` clusters (one per resident SM pair).
// The uniform grid runs whole 1024-row tiles, so a layer with 128 tiles on 74 pair slots (the example conv layers)
// pays two full waves for 1.73 waves of work, and the shared-input data gradient (64 sample-group tiles) leaves 10 of
// 74 slots idle.  Here all work forms one line of items that is cut into sk_slots equal ranges; a slot walks its
// range segment by segment, keeping barriers, TMEM and the producer pipelines alive across segments — generators and
// the TMA thread run ahead into the next segment while the epilogue warps drain the accumulator of the previous one
// (the MMA thread waits on `tmem_empty`).  Two ways to cut, neither needs scratch memory or a hand-over between slots:
//   sk_msplit == 1 (forward, per-sample data gradient): items are UNITS of 256 rows (one 128-row block per CTA) of the
//     (sample, column tile) columns of the output; a segment is a tile of 1..4 consecutive units over the whole
//     reduction, stored like a uniform tile (C3 conv forward: 512 units on 74 slots = 7 or 6 units per slot, run as
//     4 + 3 / 3 + 3 instead of two waves of 4);
//   sk_msplit == 0 (summed data gradient: one output for all samples, zeroed first): items are the k-block
//     iterations (reduction blocks x samples) of the 1024-row tiles; every segment ADDS its partial sums
//     (red.global.add), as the uniform grid does when it splits the samples over grid.z.
struct SkPipe {
  uint64_t* full_a;      // [2]  as PairPipe
  uint64_t* full_w;      // [4]
  uint64_t* empty_w;     // [4]
  uint64_t* accum_full;  // both CTAs: one phase per segment
  uint64_t* tmem_empty;  // leader: one arrival per epilogue warp of both CTAs, one phase per segment
  uint32_t* tmem_slot;
  float* aux;            // [2][128]: sampled bias rows, double buffered by segment parity
  uint32_t ring_a, ring_w;
  uint32_t stage;        // four staging buffers of the epilogue warps (the rings stay live across segments)
};
constexpr size_t kPairSkSmem = kSmemAux + 1024 + static_cast<size_t>(kASlots) * kTileBytes + kSkWSlots * kHalfTileBytes +
                               4 * static_cast<size_t>(kStageWarpBytes);
__device__ __forceinline__ SkPipe carve_sk(uint8_t* smem_raw) {
  SkPipe p;
  p.full_a = reinterpret_cast<uint64_t*>(smem_raw);
  p.full_w = p.full_a + 2;
  p.empty_w = p.full_w + kSkWSlots;
  p.accum_full = p.empty_w + kSkWSlots;
  p.tmem_empty = p.accum_full + 1;
  p.tmem_slot = reinterpret_cast<uint32_t*>(p.tmem_empty + 1);
  p.aux = reinterpret_cast<float*>(smem_raw + 512);
  const uint32_t base = smem_u32(smem_raw) + kSmemAux;
  p.ring_a = (base + 1023u) & ~1023u;
  p.ring_w = p.ring_a + kASlots * kTileBytes;
  p.stage = p.ring_w + kSkWSlots * kHalfTileBytes;
  return p;
}
struct SkSeg {
  int lead_row0;         // first output row of the leader CTA
  int mb_cap;            // row blocks per CTA of this tile (1..4); the peer's rows start mb_cap * 128 further
  int y, z;              // column tile, sample (0 for a summed tile)
  int i0, n;             // first k-block iteration inside the tile, iterations of this segment
};
// units of a tile that starts a run of `run` units: never a lone unit after a full tile (5 -> 3 + 2, 6 -> 3 + 3, 7 -> 4 + 3)
__host__ __device__ __forceinline__ int sk_tile_units(int run) {
  return (run >= 8 || run <= 4) ? (run < 4 ? run : 4) : (run + 1) / 2;
}
// next segment of the range [pos, end); advances pos
__host__ __device__ __forceinline__ bool sk_next(const TmaContractParams& p, int& pos, int end, SkSeg* s) {
  if (pos >= end) return false;
  if (p.sk_msplit) {
    const int col = pos / p.sk_mtiles, u0 = pos - col * p.sk_mtiles;        // sk_mtiles: units per column
    int run = p.sk_mtiles - u0;
    if (run > end - pos) run = end - pos;
    const int u = sk_tile_units(run);
    s->lead_row0 = u0 * 256;
    s->mb_cap = u;
    s->z = col / p.sk_gx;
    s->y = col - s->z * p.sk_gx;
    s->i0 = 0;
    s->n = p.sk_its;
    pos += u;
    return true;
  }
  const int t = pos / p.sk_its;
  s->i0 = pos - t * p.sk_its;
  const int left = end - pos, room = p.sk_its - s->i0;
  s->n = left < room ? left : room;
  const int r = t / p.sk_mtiles;                                            // sk_mtiles: 1024-row tiles per column
  s->lead_row0 = (t - r * p.sk_mtiles) * 1024;
  s->mb_cap = 4;
  s->z = r / p.sk_gx;
  s->y = r - s->z * p.sk_gx;
  pos += s->n;
  return true;
}

template <bool kDgrad, int kMode>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreadsTma, 1)
contract_pair_sk_kernel(const __grid_constant__ TmaContractParams p) {
  constexpr bool kConv = kMode == 1;
  constexpr int kSigns = kMode == 2 ? 1 : 0;
  constexpr int MB = 4;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  long long w_mma_w = 0, w_mma_a = 0, w_mma_t = 0, w_gen = 0, w_tma = 0, w_epi = 0;
  (void)w_mma_w; (void)w_mma_a; (void)w_mma_t; (void)w_gen; (void)w_tma; (void)w_epi;
#ifdef BNN_PROFILE_WAITS
  const long long t_kernel0 = clock64();
  if (threadIdx.x == 0) {
    unsigned long long g;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
    atomicMin(&g_stage_cycles[6], g);
  }
#endif
  const SkPipe pipe = carve_sk(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();          // 0 = leader
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) mbar_init(pipe.full_a + i, 1);
    for (int i = 0; i < 4; ++i) { mbar_init(pipe.full_w + i, 2 * kGroupWarps); mbar_init(pipe.empty_w + i, 1); }
    mbar_init(pipe.accum_full, 1);
    mbar_init(pipe.tmem_empty, 2 * (kEpiThreads / 32));
    fence_mbar_init();
  }
  if (warp == kTmaWarp && lane == 0) tma_prefetch_desc(&p.map_l);
  if (warp == kMmaWarpT) tmem_alloc_pair(pipe.tmem_slot, 512);
  tc_fence_before_sync();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(pipe.tmem_slot);
#ifdef BNN_PROFILE_WAITS
  if (threadIdx.x == 0) { BNN_STAGE(0, t_kernel0); atomicAdd(&g_stage_cycles[5], 1ull); }
#endif

  const int slot = static_cast<int>(blockIdx.x >> 1);
  const int range_a = static_cast<int>(static_cast<long long>(p.sk_total) * slot / p.sk_slots);
  const int range_b = static_cast<int>(static_cast<long long>(p.sk_total) * (slot + 1) / p.sk_slots);
  const int n_cols = kDgrad ? p.K : p.N;
  const int n_red = kDgrad ? p.N : p.K;
  const int red_blocks = (n_red + kBK - 1) / kBK;

  if (warp < kGenWarps) {
    // ------------------------------------------------------------------ weight generators (half tile per CTA, four groups)
    const int group = warp / kGroupWarps, tid = threadIdx.x - group * kGroupThreads;
    const RngKey key = resolve_rng(p.rng_w);
    const uint32_t lead_full_w = mapa_u32(smem_u32(pipe.full_w + group), 0);
    const uint32_t tile = pipe.ring_w + group * kHalfTileBytes;
    int pos = range_a, gi0 = 0;
    SkSeg sg;
    while (sk_next(p, pos, range_b, &sg)) {
      const int half0 = sg.y * 128 + static_cast<int>(rank) * 64;
      for (int j = (group - gi0) & 3; j < sg.n; j += kGenGroups) {       // iterations gi = gi0 + j with gi % 4 == group
        const int gi = gi0 + j, it = sg.i0 + j;
        const int ds = it / red_blocks, rb = it - ds * red_blocks;
        EpsSrc eps;
        eps.inj = p.eps_w ? p.eps_w + static_cast<int64_t>(sg.z + ds) * p.w_numel : nullptr;
        eps.key = key;
        eps.sample = p.sample_begin + sg.z + ds;
        { BNN_T0(); mbar_wait(pipe.empty_w + group, ((gi >> 2) & 1) ^ 1); BNN_ACC(w_gen); }
        if (!kDgrad)
          gen_w_tile<64, false, kSigns>(tile, p.mu_w, p.sigma_w, eps, half0, p.N, rb * kBK, p.K, tid, p.K, 0);
        else if (!kConv)
          gen_w_tile<64, true, kSigns>(tile, p.mu_w, p.sigma_w, eps, rb * kBK, p.N, half0, p.K, tid, p.K, 0);
        else {
          const int tapf = rb / p.conv.cblocks, o0 = (rb - tapf * p.conv.cblocks) * kBK;
          gen_w_tile<64, true, kSigns>(tile, p.mu_w, p.sigma_w, eps, o0, p.conv.w_rows, half0, p.K, tid,
                               static_cast<int64_t>(p.conv.taps) * p.K, static_cast<int64_t>(p.conv.taps - 1 - tapf) * p.K);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(lead_full_w);
      }
      gi0 += sg.n;
    }
  } else if (warp == kTmaWarp) {
    // ------------------------------------------------------------------ TMA issuer (own rows; bytes counted by the leader)
    if (lane == 0) {
      const uint32_t lead_full_a = mapa_u32(smem_u32(pipe.full_a), 0);
      int pos = range_a, gi = 0;
      SkSeg sg;
      while (sk_next(p, pos, range_b, &sg)) {
        const int row0 = sg.lead_row0 + static_cast<int>(rank) * (sg.mb_cap * 128);
        int mb_pair = (p.M - sg.lead_row0 + 127) / 128;
        if (mb_pair > sg.mb_cap) mb_pair = sg.mb_cap;
        int ds = sg.i0 / red_blocks, rb = sg.i0 - ds * red_blocks;
        int smp = p.shared_l ? 0 : sg.z + ds;
        int cw[MB], ch[MB], cn[MB];
        int tap = 0, cb = 0;
        if (kConv) {
#pragma unroll
          for (int mb = 0; mb < MB; ++mb) conv_pixel(p.conv, row0 + mb * 128, smp, &cw[mb], &ch[mb], &cn[mb]);
          tap = rb / p.conv.cblocks;
          cb = rb - tap * p.conv.cblocks;
        }
        for (int j = 0; j < sg.n; ++j, ++gi) {
          const int g = gi & 1;
          if (gi >= 2) { BNN_T0(); mbar_wait(pipe.empty_w + ((gi - 2) & 3), ((gi - 2) >> 2) & 1); BNN_ACC(w_tma); }
          if (rank == 0) mbar_arrive_expect_tx(pipe.full_a + g, 2 * mb_pair * kTileBytes);
          if (!kConv) {
            for (int mb = 0; mb < mb_pair; ++mb)
              tma_load_3d_pair(pipe.ring_a + (g * MB + mb) * kTileBytes, &p.map_l, rb * kBK, row0 + mb * 128, smp,
                               lead_full_a + g * 8);
          } else {
            const int kh = tap / p.conv.KW, kw = tap - kh * p.conv.KW;
#pragma unroll
            for (int mb = 0; mb < MB; ++mb)
              if (mb < mb_pair)
                tma_load_im2col_4d_pair(pipe.ring_a + (g * MB + mb) * kTileBytes, &p.map_l, cb * kBK, cw[mb], ch[mb], cn[mb],
                                        static_cast<uint16_t>(kw * p.conv.dw), static_cast<uint16_t>(kh * p.conv.dh),
                                        lead_full_a + g * 8);
            if (++cb == p.conv.cblocks) { cb = 0; ++tap; }
          }
          if (++rb == red_blocks) {                  // next sample of a summed tile
            rb = 0; ++ds; tap = 0; cb = 0;
            if (!p.shared_l) {
              smp = sg.z + ds;
              if (kConv) {
#pragma unroll
                for (int mb = 0; mb < MB; ++mb) cn[mb] += p.conv.imgs;
              }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == kMmaWarpT) {
    // ------------------------------------------------------------------ MMA issuer (leader only)
    if (rank == 0 && lane == 0) {
      const uint32_t idesc = make_idesc_tf32(256, 128, false, kDgrad);
      const uint64_t desc_a0 = make_smem_desc(pipe.ring_a);
      const uint64_t desc_w0 = kDgrad ? make_smem_desc_mn(pipe.ring_w, kMnLbo, kMnSbo) : make_smem_desc(pipe.ring_w);
      int pos = range_a, gi = 0, seg = 0;
      SkSeg sg;
      while (sk_next(p, pos, range_b, &sg)) {
        int mb_pair = (p.M - sg.lead_row0 + 127) / 128;
        if (mb_pair > sg.mb_cap) mb_pair = sg.mb_cap;
        if (seg > 0) {                                 // the epilogue warps of both CTAs have drained the last segment
          { BNN_T0(); mbar_wait_cluster(pipe.tmem_empty, (seg - 1) & 1); BNN_ACC(w_mma_t); }
          tc_fence_after_sync();
        }
        for (int j = 0; j < sg.n; ++j, ++gi) {
          const int wslot = gi & 3, g = gi & 1;
          { BNN_T0(); mbar_wait_cluster(pipe.full_w + wslot, (gi >> 2) & 1); BNN_ACC(w_mma_w); }
          { BNN_T0(); mbar_wait_cluster(pipe.full_a + g, (gi >> 1) & 1); BNN_ACC(w_mma_a); }
          tc_fence_after_sync();
          const uint64_t da0 = desc_advance(desc_a0, static_cast<uint32_t>(g * MB) * (kTileBytes >> 4));
          const uint64_t db0 = desc_advance(desc_w0, static_cast<uint32_t>(wslot) * (kHalfTileBytes >> 4));
          for (int mb = 0; mb < mb_pair; ++mb) {
#pragma unroll
            for (int ks = 0; ks < kBK / 8; ++ks)
              mma_tf32_pair(tmem + mb * 128, desc_advance(da0, mb * (kTileBytes >> 4) + ks * 2),
                            desc_advance(db0, ks * ((kDgrad ? kMnKStep : 32u) >> 4)), idesc, j > 0 || ks > 0);
          }
          mma_commit_pair(pipe.empty_w + wslot, 3);
        }
        mma_commit_pair(pipe.accum_full, 3);
        ++seg;
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (own TMEM, own rows)
    const int et = threadIdx.x - kEpiWarp0T * 32;   // 0..127
    const int quad = warp & 3;
    const uint32_t lead_tmem_empty = mapa_u32(smem_u32(pipe.tmem_empty), 0);
    int pos = range_a, seg = 0;
    SkSeg sg;
    while (sk_next(p, pos, range_b, &sg)) {
      const int col0 = sg.y * 128;
      const int row0 = sg.lead_row0 + static_cast<int>(rank) * (sg.mb_cap * 128);
      int mb_pair = (p.M - sg.lead_row0 + 127) / 128;
      if (mb_pair > sg.mb_cap) mb_pair = sg.mb_cap;
      const bool add = p.atomic_out != 0;              // summed tiles: every segment adds into the zeroed output
      float* aux = pipe.aux + (seg & 1) * 128;
      if (!kDgrad) {
        float b = 0.f;
        const int n = col0 + et;
        if (p.mu_b != nullptr && n < p.N) {
          const float e = p.eps_b ? __ldg(p.eps_b + static_cast<int64_t>(sg.z) * p.N + n)
                                  : eps1(resolve_rng(p.rng_b), p.sample_begin + sg.z, static_cast<uint64_t>(n));
          b = fmaf(__ldg(p.sigma_b + n), e, __ldg(p.mu_b + n));
        }
        aux[et] = b;
        named_bar_sync(kEpiBarrier, kEpiThreads);
      }
      mbar_wait(pipe.accum_full, seg & 1);
      tc_fence_after_sync();
#ifdef BNN_PROFILE_WAITS
      const long long t_epi0 = clock64();
#endif
      View out = p.out;
      out.base += (p.sum_samples ? 0 : static_cast<int64_t>(sg.z) * p.out_sample_stride);
      const int cols_here = n_cols - col0 < 128 ? n_cols - col0 : 128;
      const bool staged = stage_eligible(out, p.vec_out != 0, cols_here) && (out.P == 1 || (p.M & 3) == 0);
      for (int mb = 0; mb < mb_pair; ++mb) {
        const int m = row0 + mb * 128 + quad * 32 + lane;
        if (row0 + mb * 128 >= p.M) break;               // uniform per CTA
        if (staged) {
          drain_block_staged<!kDgrad>(tmem + (static_cast<uint32_t>(quad * 32) << 16) + mb * 128,
                                      pipe.stage + quad * kStageWarpBytes, aux, out, row0 + mb * 128 + quad * 32,
                                      p.M, col0, cols_here, lane, add);
          continue;
        }
        for (int c = 0; c * 16 < cols_here; ++c) {
          float v[16];
          tmem_ld16(tmem + (static_cast<uint32_t>(quad * 32) << 16) + mb * 128 + c * 16, v);
          if (!kDgrad) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += aux[c * 16 + j];
          }
          if (m < p.M) {
            if (add) add_chunk(out, m, col0 + c * 16, n_cols, v, p.vec_out != 0);
            else store_chunk(out, m, col0 + c * 16, n_cols, v, p.vec_out != 0);
          }
        }
      }
      tc_fence_before_sync();                            // all tcgen05.ld of this segment have completed (wait::ld)
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(lead_tmem_empty);
#ifdef BNN_PROFILE_WAITS
      w_epi += clock64() - t_epi0;
#endif
      ++seg;
    }
  }
#ifdef BNN_PROFILE_WAITS
  if (threadIdx.x == kEpiWarp0T * 32) BNN_STAGE(3, t_kernel0);
  if (threadIdx.x == 0) BNN_STAGE(1, t_kernel0);
  if (lane == 0) {      // [0] kernel [1] MMA: weights [2] MMA: activations [3] generator warp 0 [4] TMA [5] CTAs [6] MMA: TMEM drain [7] epilogue busy
    if (warp == 0) { atomicAdd(&g_wait_cycles[3], (unsigned long long)w_gen); atomicAdd(&g_wait_cycles[0], (unsigned long long)(clock64() - t_kernel0)); atomicAdd(&g_wait_cycles[5], 1ull); }
    if (w_mma_w | w_mma_a | w_mma_t) { atomicAdd(&g_wait_cycles[1], (unsigned long long)w_mma_w); atomicAdd(&g_wait_cycles[2], (unsigned long long)w_mma_a); atomicAdd(&g_wait_cycles[6], (unsigned long long)w_mma_t); }
    if (w_tma) atomicAdd(&g_wait_cycles[4], (unsigned long long)w_tma);
    if (warp == kEpiWarp0T) atomicAdd(&g_wait_cycles[7], (unsigned long long)w_epi);
  }
#endif
  tc_fence_before_sync();
  cluster_sync_all();                                // nobody leaves while the partner may still signal or read
  if (warp == kMmaWarpT) {
    tc_fence_after_sync();
    tmem_dealloc_pair(tmem, 512);
  }
#ifdef BNN_PROFILE_WAITS
  if (threadIdx.x == 0) {
    BNN_STAGE(4, t_kernel0);
    unsigned long long g;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
    atomicMax(&g_stage_cycles[7], g);
  }
#endif
}

template <bool kDgrad, int kMode>
int launch_pair_sk_v(const TmaContractParams& p, cudaStream_t st) {
  static SmemOptIn opt_in;
  const int rc = allow_dynamic_smem(contract_pair_sk_kernel<kDgrad, kMode>, kPairSkSmem, &opt_in);
  if (rc != BNN_OK) return rc;
  contract_pair_sk_kernel<kDgrad, kMode><<<dim3(2 * p.sk_slots, 1, 1), kThreadsTma, kPairSkSmem, st>>>(p);
  BNN_CUDA_OK(cudaGetLastError());
  return BNN_OK;
}
template <bool kDgrad>
int launch_pair_sk(const TmaContractParams& p, cudaStream_t st) {
  if (p.conv.on) return launch_pair_sk_v<kDgrad, 1>(p, st);
  if (p.rng_w.row_sign != nullptr) return launch_pair_sk_v<kDgrad, 2>(p, st);
  return launch_pair_sk_v<kDgrad, 0>(p, st);
}

bool pair_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("BNN_DISABLE_CTA_PAIRS");
    v = (e != nullptr && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

// Test aid: force one kernel variant (0 = pair, 1 / 2 / 4 = rows blocks per CTA, -1 = cost model) so that small test
// shapes reach every variant.  Set through bnn_debug_force_contract_variant (tests) — the environment variable
// BNN_CONTRACT_VARIANT = pair | balanced | mb4 | mb2 | mb1 only provides the initial value, read once.  8 = the CTA-pair
// kernel's balanced schedule (contract_pair_sk_kernel) whenever the pair kernel is eligible.
std::atomic<int> g_forced_variant{-2};
int forced_variant() {
  int v = g_forced_variant.load(std::memory_order_relaxed);
  if (v != -2) return v;
  v = -1;
  const char* e = getenv("BNN_CONTRACT_VARIANT");
  if (e != nullptr && e[0] == 'p') v = 0;
  else if (e != nullptr && e[0] == 'b') v = 8;                  // "balanced"
  else if (e != nullptr && e[0] == 'm' && e[1] == 'b') v = e[2] == '4' ? 4 : (e[2] == '2' ? 2 : (e[2] == '1' ? 1 : -1));
  g_forced_variant.store(v, std::memory_order_relaxed);
  return v;
}

// Tile plan of the CTA-pair kernel (TilePlan): when the uniform grid needs a fraction of a wave more than a whole
// number (1 < tiles / slots < 2 is the case that matters: the example conv layers), look for a cut of every sample into
// 8- and 6-row-block tiles — at most two kinds of samples — whose list schedule (all 8-block tiles first) finishes
// earlier.  Cost unit: one row block of MMAs per k-block; the fixed cost per tile is ignored, so a plan must win by 8 %.
bool tile_plan_enabled() {                                     // BNN_TILE_PLAN=0 (read once) keeps the uniform grid: A/B runs
  static const bool on = [] { const char* e = getenv("BNN_TILE_PLAN"); return !(e != nullptr && e[0] == '0'); }();
  return on;
}
bool solve_pair_tiles(TilePlan* plan, int m_blocks, int S, int slots);
bool plan_pair_tiles(TilePlan* plan, int m_blocks, int S, int gx, int gz, bool sum_samples, int slots) {
  plan->on = 0;
  if (!tile_plan_enabled() || sum_samples || gx != 1 || gz != S || S < 1 || slots < 1) return false;
  // the search costs milliseconds: one result per (rows, samples, machine) for the life of the process
  static std::mutex mu;
  static std::vector<std::pair<std::array<int, 3>, TilePlan>> cache;
  const std::array<int, 3> key = {m_blocks, S, slots};
  std::lock_guard<std::mutex> lock(mu);
  for (const auto& e : cache)
    if (e.first == key) { *plan = e.second; return plan->on != 0; }
  TilePlan found{};
  solve_pair_tiles(&found, m_blocks, S, slots);
  cache.emplace_back(key, found);
  *plan = found;
  return plan->on != 0;
}
bool solve_pair_tiles(TilePlan* plan, int m_blocks, int S, int slots) {
  plan->on = 0;
  const int units = (m_blocks + 1) / 2;                        // row-block pairs per sample (one row block per CTA)
  const int uniform_tiles = S * ((units + 3) / 4);
  if (uniform_tiles <= slots || uniform_tiles > 4 * slots) return false;
  auto makespan = [&](int n_a, int n_b) {                      // greedy list schedule on `slots` machines
    std::vector<int> load(static_cast<size_t>(slots), 0);
    auto place = [&](int cost) { auto it = std::min_element(load.begin(), load.end()); *it += cost; };
    for (int i = 0; i < n_a; ++i) place(4);
    for (int i = 0; i < n_b; ++i) place(3);
    return *std::max_element(load.begin(), load.end());
  };
  const int base = makespan(uniform_tiles, 0);
  struct Cut { int a, b; };
  std::vector<Cut> cuts;                                       // 4 a + 3 b covers the units with at most 2 to spare
  for (int a = 0; a * 4 <= units + 3; ++a)
    for (int b = 0; b <= (units + 2) / 3 + 1; ++b) {
      const int cover = 4 * a + 3 * b;
      if (cover >= units && cover <= units + 2 && (a > 0 || b > 0)) cuts.push_back({a, b});
    }
  int best = base;
  TilePlan found = *plan;
  for (const Cut& c1 : cuts)
    for (const Cut& c2 : cuts)
      for (int s1 = 0; s1 <= S; ++s1) {
        if ((s1 == 0 && (&c1 != &cuts[0])) || (s1 == S && (&c2 != &cuts[0]))) continue;      // unused class: one representative
        const int n_a = s1 * c1.a + (S - s1) * c2.a, n_b = s1 * c1.b + (S - s1) * c2.b;
        const int ms = makespan(n_a, n_b);
        if (ms < best) { best = ms; found = TilePlan{1, n_a, s1, c1.a, c1.b, c2.a, c2.b}; }
      }
  if (best * 100 > base * 92) return false;
  *plan = found;
  return true;
}

// ---- balanced schedule, host side
std::atomic<int> g_sk_slot_cap{0};              // test aid: fewer slots than the machine has (0 = no cap)
std::atomic<int> g_sk_launches{0};              // launches that took the balanced schedule (tests, bench)
std::atomic<int> g_sk_on{-1};                   // -1: not decided yet (off unless BNN_BALANCED=1 is in the environment)
bool sk_enabled() {                             // opt-in (bnn_contract_set_balanced / BNN_BALANCED=1): measured slower on B200, DESIGN §4
  int v = g_sk_on.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("BNN_BALANCED");
    v = (e != nullptr && e[0] == '1') ? 1 : 0;
    g_sk_on.store(v, std::memory_order_relaxed);
  }
  return v == 1;
}
// clusters of contract_pair_sk_kernel that are resident at the same time on the current device (cached per device);
// the schedule does not depend on co-residency (slots never wait for each other), it only sizes the grid
int sk_max_slots() {
  static std::mutex mu;
  static int cached[64];
  static bool known[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  std::lock_guard<std::mutex> lock(mu);
  if (!known[dev]) {
    known[dev] = true;
    cached[dev] = 0;
    static SmemOptIn opt_in;
    if (allow_dynamic_smem(contract_pair_sk_kernel<false, 0>, kPairSkSmem, &opt_in) == BNN_OK) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(static_cast<unsigned>(sm_count() & ~1), 1, 1);
      cfg.blockDim = dim3(kThreadsTma, 1, 1);
      cfg.dynamicSmemBytes = kPairSkSmem;
      cudaLaunchAttribute attr{};
      attr.id = cudaLaunchAttributeClusterDimension;
      attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
      cfg.attrs = &attr;
      cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, contract_pair_sk_kernel<false, 0>, &cfg) == cudaSuccess) cached[dev] = n;
      else (void)cudaGetLastError();
    }
  }
  return cached[dev];
}

// k-block cost of a tile of u units (one row block per CTA each), in cycles (CTA timeline of a profiling build at the C3
// conv shape: the leader issues a 256 x 128 x 8 MMA per ~117 cycles, a CTA generates its half tile in ~1200)
inline double sk_kblock_cycles(int u) {
  const double mma = 117.0 * 4 * u + 130.0;
  return mma > 1250.0 ? mma : 1250.0;
}

template <bool kDgrad>
bool plan_sk(TmaContractParams& p, int gx, int m_blocks, int uniform_pairs, bool forced) {
  p.sk_slots = 0;
  if (!forced && !sk_enabled()) return false;
  if (p.sum_samples && p.out.P != 1) return false;
  const int slots = sk_max_slots();
  if (slots < 2) return false;
  const int red_blocks = ((kDgrad ? p.N : p.K) + kBK - 1) / kBK;
  const int cap = g_sk_slot_cap.load(std::memory_order_relaxed);
  const double fixed = 6000.0, epilogue = 3500.0;         // launch + prologue + pipeline fill; per row block of a tile
  const double waves = static_cast<double>((uniform_pairs + slots - 1) / slots);
  if (p.sum_samples) {
    // items = k-block iterations of the 1024-row tiles over ALL samples; every segment adds
    const int m_tiles = (m_blocks + 7) / 8;
    const int64_t n_tiles = static_cast<int64_t>(gx) * m_tiles;
    const int64_t its = static_cast<int64_t>(red_blocks) * p.S;
    const int64_t total = n_tiles * its;
    if (total <= 0 || total > (int64_t(1) << 30)) return false;
    int64_t q = slots;
    if (q > total / 8) q = total / 8;           // at least eight k-blocks per slot
    if (cap > 0 && q > cap) q = cap;
    if (q < 1) return false;
    if (!forced) {
      const double per_slot = static_cast<double>((total + q - 1) / q);
      const double segs = per_slot / static_cast<double>(its) + 1.0;
      const double sk_cost = fixed + per_slot * sk_kblock_cycles(4) + segs * 4 * epilogue * 1.4;
      const double uniform_cost = waves * (fixed + static_cast<double>(red_blocks) * p.z_per * sk_kblock_cycles(4) + 4 * epilogue * 1.4);
      if (sk_cost > 0.95 * uniform_cost) return false;
    }
    p.sk_msplit = 0;
    p.sk_slots = static_cast<int>(q);
    p.sk_mtiles = m_tiles;
    p.sk_gx = gx;
    p.sk_its = static_cast<int>(its);
    p.sk_total = static_cast<int>(total);
  } else {
    // items = 256-row units of the (sample, column tile) columns; a segment = a tile of 1..4 units, whole reduction
    const int m_units = (m_blocks + 1) / 2;
    const int64_t total = static_cast<int64_t>(gx) * p.S * m_units;
    if (total <= 0 || total > (int64_t(1) << 30)) return false;
    int64_t q = slots;
    if (q > total) q = total;
    if (cap > 0 && q > cap) q = cap;
    if (q < 1) return false;
    if (!forced) {
      // the busiest slot: ceil(total / q) units cut by sk_tile_units (ranges that straddle a column cost one more cut)
      int left = static_cast<int>((total + q - 1) / q);
      double sk_cost = fixed;
      while (left > 0) {
        const int u = sk_tile_units(left < m_units ? left : m_units);
        sk_cost += red_blocks * sk_kblock_cycles(u) + u * epilogue;
        left -= u;
      }
      const int u_uniform = m_units < 4 ? m_units : 4;
      const double uniform_cost = waves * (fixed + red_blocks * sk_kblock_cycles(u_uniform) + u_uniform * epilogue);
      if (sk_cost > 0.95 * uniform_cost) return false;
    }
    p.sk_msplit = 1;
    p.sk_slots = static_cast<int>(q);
    p.sk_mtiles = m_units;
    p.sk_gx = gx;
    p.sk_its = red_blocks;
    p.sk_total = static_cast<int>(total);
  }
  p.plan.on = 0;
  return true;
}

template <bool kDgrad>
int dispatch_tma_contract(TmaContractParams& p, int n_cols, cudaStream_t st) {
  const int m_blocks = (p.M + 127) / 128;
  const int gx = (n_cols + 127) / 128;
  const int sms = sm_count();
  const int n_red = kDgrad ? p.N : p.K;
  // Shared activations (one output summed over the samples): a CTA walks `z_per` samples; when the output has too few
  // tiles to fill the machine the samples are split over grid.z and the partial sums are added to the zeroed output.
  int gz = p.S;
  if (p.sum_samples) {
    const int64_t tiles = static_cast<int64_t>(gx) * ((m_blocks + 3) / 4);
    int groups = static_cast<int>((sms + tiles - 1) / tiles);
    if (groups > p.S) groups = p.S;
    if (groups < 1 || p.out.P != 1) groups = 1;
    p.z_per = (p.S + groups - 1) / groups;
    gz = (p.S + p.z_per - 1) / p.z_per;
    p.atomic_out = gz > 1 ? 1 : 0;
    if (p.atomic_out)
      BNN_CUDA_OK(cudaMemset2DAsync(p.out.base, static_cast<size_t>(p.out.bs) * 4, 0, static_cast<size_t>(n_cols) * 4,
                                    static_cast<size_t>(p.M), st));
  }
  // Rows per CTA by a cost model of the measured kernels, in cycles per 32-wide k-block: a CTA generates
  // its weight tile at ~1.86 weights/clk whatever the number of 128-row blocks that reuse it (128 x 32 tile: ~2200
  // cycles; half of it per CTA of a pair), its MMAs + operand delivery cost ~425 cycles per block (~468 per block pair
  // with cta_group::2), so a k-block takes the larger of the two; a launch takes waves x (fixed cost + k-blocks x that).
  // Small problems thereby keep few rows per CTA (every SM gets a short k-loop) and large ones share each generated
  // tile between 1024 rows — without insisting on a CTA for every SM: 144 busy SMs with 4x the reuse beat 148.
  const double iters = static_cast<double>((n_red + kBK - 1) / kBK) * (p.sum_samples ? p.z_per : 1);
  const double w_rows = n_cols < 128 ? n_cols : 128;
  auto cost_single = [&](int mb) {
    const int64_t ctas = static_cast<int64_t>(gx) * ((m_blocks + mb - 1) / mb) * gz;
    const double waves = static_cast<double>((ctas + sms - 1) / sms);
    const double gen = w_rows * kBK / 1.86, mma = 425.0 * (mb < m_blocks ? mb : m_blocks);
    return waves * (4000.0 + 600.0 * mb + iters * ((gen > mma ? gen : mma) + 100.0));
  };
  auto cost_pair = [&]() {
    const int64_t pairs = static_cast<int64_t>(gx) * ((m_blocks + 7) / 8) * gz;
    const int slots = sms / 2;
    const double waves = static_cast<double>((pairs + slots - 1) / slots);
    const double gen = (w_rows > 64 ? 64 : w_rows) * kBK / 1.86, mma = 468.0 * 4;
    return waves * (5000.0 + 600.0 * 4 + iters * ((gen > mma ? gen : mma) + 100.0));
  };
  int best = 1;
  double best_cost = cost_single(1);
  for (int mb = 2; mb <= 4; mb *= 2) {
    if (m_blocks < mb) break;
    const double c = cost_single(mb);
    if (c < best_cost) { best_cost = c; best = mb; }
  }
  if (m_blocks > 4 && pair_enabled() && cost_pair() < 0.9 * best_cost) best = 0;       // a clear win only: the model is coarse
  const int forced = forced_variant();
  const bool force_sk = forced == 8 && m_blocks > 4;
  if (force_sk) best = 0;
  else if ((forced > 0 && forced != 8) || (forced == 0 && m_blocks > 4)) best = forced;
  if (best == 0) {                                                // two CTAs (one cluster) per 1024 rows
    const int pairs = (m_blocks + 7) / 8;
    if ((forced < 0 || force_sk) && plan_sk<kDgrad>(p, gx, m_blocks, pairs * gx * gz, force_sk)) {
      if (p.sum_samples && !p.atomic_out) {                       // every segment adds: the output starts from zero
        p.atomic_out = 1;
        BNN_CUDA_OK(cudaMemset2DAsync(p.out.base, static_cast<size_t>(p.out.bs) * 4, 0, static_cast<size_t>(n_cols) * 4,
                                      static_cast<size_t>(p.M), st));
      }
      g_sk_launches.fetch_add(1, std::memory_order_relaxed);
      return launch_pair_sk<kDgrad>(p, st);
    }
    if (plan_pair_tiles(&p.plan, m_blocks, p.S, gx, gz, p.sum_samples != 0, sms / 2))
      return launch_pair_contract<4, kDgrad>(p, dim3(2 * (p.plan.n_a + p.plan.s1 * p.plan.b1 + (p.S - p.plan.s1) * p.plan.b2), 1, 1), st);
    return launch_pair_contract<4, kDgrad>(p, dim3(2 * pairs, gx, gz), st);
  }
  if (best == 4) return launch_tma_contract<4, kDgrad>(p, dim3(gx, (m_blocks + 3) / 4, gz), st);
  if (best == 2) return launch_tma_contract<2, kDgrad>(p, dim3(gx, (m_blocks + 1) / 2, gz), st);
  return launch_tma_contract<1, kDgrad>(p, dim3(gx, m_blocks, gz), st);
}

// ---------------------------------------------------------------------------------------------- weight gradient
// CTA (k-tile, n-tile, sample group).  Stage = dY^T tile (MN = n) + A^T tile (MN = k), each 4 TMA boxes of
// 32 MN x 32 K-rows (m).  TMEM columns: [0,128) G0 | [128,256) sum G | [256,384) sum G o eps | [384,512) G1.
constexpr int kWgStages = 3;                 // stages of 64 contraction rows: 8 MMAs per barrier wait / tcgen05.commit
constexpr int kWgRows = 64;                  // contraction rows (m) per stage
constexpr int kWgOperandBytes = 4 * kWgRows * kRowBytes;      // 4 MN groups x 64 K-rows x 128 B = 32 KiB per operand
constexpr uint32_t kWgLbo = kWgRows * kRowBytes;              // next 32 MN elements: 8 KiB further
constexpr int kWgEpiWarps = 12;              // three warps per TMEM lane quarter share the 8 column chunks of a tile:
constexpr int kWgThreads = (2 + kWgEpiWarps) * 32;   // regenerating eps for 128 x 128 per sample is the long pole
                                             // warp 0 TMA, warp 1 MMA (+TMEM owner), warps 2.. epilogue
constexpr size_t kWgradSmem = kSmemAux + 1024 + static_cast<size_t>(kWgStages) * 2 * kWgOperandBytes;

struct TmaWgradParams {
  CUtensorMap map_dy;       // dY [S][M][N], box 32 x 32
  CUtensorMap map_a;        // A  [S or 1][M][K], box 32 x 32; conv: the NHWC input (im2col map, 32 channels x 64 pixels)
  ConvCoords conv;
  const float* rho_w;
  const float* eps_w;
  float* dmu_w;
  float* drho_w;
  int M, N, K, S;
  uint32_t sample_begin;
  bnn_rng rng_w;
  int shared_a;
  int n_chunks;             // the M reduction of every sample is cut into n_chunks pieces of chunk_blocks k-blocks:
  int chunk_blocks;         // a work unit = (sample, chunk); (sum_chunks G) o eps = sum_chunks (G o eps), so units are independent
};

// kPair: the same kernel over a cluster of two CTAs (launched with cluster dimension (2,1,1)): the pair owns two adjacent
// n-tiles and ONE k-tile, tcgen05.mma.cta_group::2 computes 256 (n) x 128 (k) per instruction, and each CTA loads its own
// dY^T tile but only HALF of the shared A^T tile (48 instead of 64 KiB per stage and CTA: the kernel is paced by operand
// delivery).  TMA bytes and the epilogue's buffer releases land on the leader's barriers; commits are multicast.
template <bool kPair, bool kSignNoise>
__global__ void __launch_bounds__(kWgThreads, 1) wgrad_tma_kernel(const __grid_constant__ TmaWgradParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  constexpr uint32_t kTmemCols = 512;
  constexpr uint32_t kColG0 = 0, kColMu = 128, kColRho = 256, kColG1 = 384;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* empty = full + kWgStages;
  uint64_t* accum_full = empty + kWgStages;
  uint64_t* accum_empty = accum_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_empty + 2);
  const uint32_t ring = (smem_u32(smem_raw) + kSmemAux + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kWgStages; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(accum_full + i, 1); mbar_init(accum_empty + i, (kPair ? 2 : 1) * kWgEpiWarps); }
    fence_mbar_init();
    tma_prefetch_desc(&p.map_dy);
    tma_prefetch_desc(&p.map_a);
  }
  if (warp == 1) { if (kPair) tmem_alloc_pair(tmem_slot, kTmemCols); else tmem_alloc(tmem_slot, kTmemCols); }
  tc_fence_before_sync();
  if (kPair) cluster_sync_all(); else __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;      // 0 = leader

  const int n0 = blockIdx.x * 128;                          // this CTA's rows of dmu / drho (pairs are adjacent in x)
  const int k0 = blockIdx.y * 128;
  const int groups = gridDim.z;
  const int units = p.S * p.n_chunks;
  const int per = (units + groups - 1) / groups;
  const int s_begin = blockIdx.z * per;                          // work units [s_begin, s_end)
  const int s_end = s_begin + per < units ? s_begin + per : units;
  const int m_total = (p.M + kWgRows - 1) / kWgRows;        // stages of 64 rows
  auto unit_blocks = [&](int u, int* mb0) {                      // k-block range of unit u
    const int c = u % p.n_chunks;
    *mb0 = c * p.chunk_blocks;
    const int left = m_total - *mb0;
    return left < p.chunk_blocks ? left : p.chunk_blocks;
  };

  if (s_begin < s_end) {
    if (warp == 0) {
      if (lane == 0) {
        int it = 0;
        // conv: the 32-column group kc .. kc + 31 of the im2col matrix = (tap, channels c0 ..) of the NHWC input — fixed per
        // CTA; groups beyond K read channel `chans` (outside the tensor: zeros, the transaction still completes)
        int g_c0[4] = {0, 0, 0, 0};
        uint16_t g_ow[4] = {0, 0, 0, 0}, g_oh[4] = {0, 0, 0, 0};
        if (p.conv.on) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int kc = k0 + g * 32;
            if (kc >= p.K) { g_c0[g] = p.conv.chans; continue; }
            const int tap = kc / p.conv.chans;
            g_c0[g] = kc - tap * p.conv.chans;
            const int kh = tap / p.conv.KW, kw = tap - kh * p.conv.KW;
            g_ow[g] = static_cast<uint16_t>(kw * p.conv.dw);
            g_oh[g] = static_cast<uint16_t>(kh * p.conv.dh);
          }
        }
        for (int u = s_begin; u < s_end; ++u) {
          const int s = u / p.n_chunks;
          const int sa = p.shared_a ? 0 : s;
          int mb0;
          const int m_blocks = unit_blocks(u, &mb0);
          for (int mb = mb0; mb < mb0 + m_blocks; ++mb, ++it) {
            const int stage = it % kWgStages;
            mbar_wait(empty + stage, ((it / kWgStages) & 1) ^ 1);
            const uint32_t base = ring + stage * 2 * kWgOperandBytes;
            int cw = 0, ch = 0, cn = 0;
            if (p.conv.on) conv_pixel(p.conv, mb * kWgRows, sa, &cw, &ch, &cn);
            if (!kPair) {
              mbar_arrive_expect_tx(full + stage, 2 * kWgOperandBytes);
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                tma_load_3d(base + g * kWgLbo, &p.map_dy, n0 + g * 32, mb * kWgRows, s, full + stage);
                if (!p.conv.on) {
                  tma_load_3d(base + kWgOperandBytes + g * kWgLbo, &p.map_a, k0 + g * 32, mb * kWgRows, sa, full + stage);
                } else {
                  tma_load_im2col_4d(base + kWgOperandBytes + g * kWgLbo, &p.map_a, g_c0[g], cw, ch, cn, g_ow[g], g_oh[g],
                                     full + stage);
                }
              }
            } else {
              const uint32_t lead_full = mapa_u32(smem_u32(full + stage), 0);
              if (rank == 0) mbar_arrive_expect_tx(full + stage, 2 * (kWgOperandBytes + kWgOperandBytes / 2));
#pragma unroll
              for (int g = 0; g < 4; ++g)
                tma_load_3d_pair(base + g * kWgLbo, &p.map_dy, n0 + g * 32, mb * kWgRows, s, lead_full);
#pragma unroll
              for (int g = 0; g < 2; ++g) {     // this CTA's half (64 columns) of the shared A^T tile
                const int kc = k0 + static_cast<int>(rank) * 64 + g * 32;
                if (!p.conv.on) {
                  tma_load_3d_pair(base + kWgOperandBytes + g * kWgLbo, &p.map_a, kc, mb * kWgRows, sa, lead_full);
                } else {
                  const int gi = static_cast<int>(rank) * 2 + g;
                  tma_load_im2col_4d_pair(base + kWgOperandBytes + g * kWgLbo, &p.map_a, g_c0[gi], cw, ch, cn, g_ow[gi],
                                          g_oh[gi], lead_full);
                }
              }
            }
          }
        }
      }
      __syncwarp();
    } else if (warp == 1) {
      if (lane == 0 && rank == 0) {
        const uint32_t idesc = kPair ? make_idesc_tf32(256, 128, true, true) : make_idesc_tf32(128, mma_n(p.K - k0), true, true);
        const uint64_t desc_dy0 = make_smem_desc_mn(ring, kWgLbo, kMnSbo);
        const uint64_t desc_a0 = make_smem_desc_mn(ring + kWgOperandBytes, kWgLbo, kMnSbo);
        int it = 0;
        for (int u = s_begin, i = 0; u < s_end; ++u, ++i) {
          const int buf = i & 1;
          if (kPair) mbar_wait_cluster(accum_empty + buf, ((i >> 1) & 1) ^ 1); else mbar_wait(accum_empty + buf, ((i >> 1) & 1) ^ 1);
          tc_fence_after_sync();
          const uint32_t d = tmem + (buf ? kColG1 : kColG0);
          int mb0;
          const int m_blocks = unit_blocks(u, &mb0);
          for (int mb = 0; mb < m_blocks; ++mb, ++it) {
            const int stage = it % kWgStages;
            if (kPair) mbar_wait_cluster(full + stage, (it / kWgStages) & 1); else mbar_wait(full + stage, (it / kWgStages) & 1);
            tc_fence_after_sync();
            const uint64_t da0 = desc_advance(desc_dy0, static_cast<uint32_t>(stage) * (2 * kWgOperandBytes >> 4));
            const uint64_t db0 = desc_advance(desc_a0, static_cast<uint32_t>(stage) * (2 * kWgOperandBytes >> 4));
#pragma unroll
            for (int ks = 0; ks < kWgRows / 8; ++ks) {
              if (kPair)
                mma_tf32_pair(d, desc_advance(da0, ks * (kMnKStep >> 4)), desc_advance(db0, ks * (kMnKStep >> 4)), idesc,
                              mb > 0 || ks > 0);
              else
                mma_tf32(d, desc_advance(da0, ks * (kMnKStep >> 4)), desc_advance(db0, ks * (kMnKStep >> 4)), idesc,
                         mb > 0 || ks > 0);
            }
            if (kPair) mma_commit_pair(empty + stage, 3); else mma_commit(empty + stage);
          }
          if (kPair) mma_commit_pair(accum_full + buf, 3); else mma_commit(accum_full + buf);
        }
      }
      __syncwarp();
    } else {
      const int quad = warp & 3;                    // TMEM lane quarter this warp may access
      const int sub = (warp - 2) >> 2;              // which third of the column chunks it handles
      const int n = n0 + quad * 32 + lane;
      const uint32_t lane_addr = tmem + (static_cast<uint32_t>(quad * 32) << 16);
      const RngKey key = resolve_rng(p.rng_w);
      const int cols_here = p.K - k0 < 128 ? p.K - k0 : 128;
      for (int u = s_begin, i = 0; u < s_end; ++u, ++i) {
        const int s = u / p.n_chunks;
        const int buf = i & 1;
        const bool first = (u == s_begin), last = (u + 1 == s_end);
        EpsSrc eps;
        eps.inj = p.eps_w ? p.eps_w + static_cast<int64_t>(s) * p.N * p.K : nullptr;
        eps.key = key;
        eps.sample = p.sample_begin + s;
        mbar_wait(accum_full + buf, (i >> 1) & 1);
        tc_fence_after_sync();
        for (int c = sub; c * 16 < cols_here; c += kWgEpiWarps / 4) {
          float g[16], dm[16], dr[16];
          tmem_ld16(lane_addr + (buf ? kColG1 : kColG0) + c * 16, g);
          if (!first) {
            tmem_ld16(lane_addr + kColMu + c * 16, dm);
            tmem_ld16(lane_addr + kColRho + c * 16, dr);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) dm[j] = dr[j] = 0.f;
          }
          const int k = k0 + c * 16;
          if (n < p.N) {
            const int64_t row = static_cast<int64_t>(n) * p.K;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (k + q * 4 < p.K) {
                const float4 e = eps_vec4<kSignNoise ? 1 : 0>(eps, row + k + q * 4);
                dm[q * 4 + 0] += g[q * 4 + 0]; dr[q * 4 + 0] = fmaf(g[q * 4 + 0], e.x, dr[q * 4 + 0]);
                dm[q * 4 + 1] += g[q * 4 + 1]; dr[q * 4 + 1] = fmaf(g[q * 4 + 1], e.y, dr[q * 4 + 1]);
                dm[q * 4 + 2] += g[q * 4 + 2]; dr[q * 4 + 2] = fmaf(g[q * 4 + 2], e.z, dr[q * 4 + 2]);
                dm[q * 4 + 3] += g[q * 4 + 3]; dr[q * 4 + 3] = fmaf(g[q * 4 + 3], e.w, dr[q * 4 + 3]);
              }
            }
          }
          if (!last) {
            tmem_st16(lane_addr + kColMu + c * 16, dm);
            tmem_st16(lane_addr + kColRho + c * 16, dr);
          } else if (n < p.N) {
            const int64_t row = static_cast<int64_t>(n) * p.K;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (k + q * 4 < p.K) {
                const int64_t i0 = row + k + q * 4;
                const float4 r = __ldg(reinterpret_cast<const float4*>(p.rho_w + i0));
                red_add4(p.dmu_w + i0, dm[q * 4], dm[q * 4 + 1], dm[q * 4 + 2], dm[q * 4 + 3]);
                red_add4(p.drho_w + i0, dr[q * 4] * sigmoid_fast(r.x), dr[q * 4 + 1] * sigmoid_fast(r.y),
                         dr[q * 4 + 2] * sigmoid_fast(r.z), dr[q * 4 + 3] * sigmoid_fast(r.w));
              }
            }
          }
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) {
          if (kPair) mbar_arrive_cluster(mapa_u32(smem_u32(accum_empty + buf), 0)); else mbar_arrive(accum_empty + buf);
        }
      }
    }
  }
  tc_fence_before_sync();
  if (kPair) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    if (kPair) tmem_dealloc_pair(tmem, kTmemCols); else tmem_dealloc(tmem, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------- self test
// MN-major operand tiles: D(128 x 64) = A^T-tile(MN-major, 128 x 32) * B^T-tile(MN-major, 64 x 32)^T against a
// serial fp32 loop over the TF32-rounded values.
__device__ __forceinline__ float st_a(int r, int c) {
  return __uint_as_float(to_tf32(0.01f * static_cast<float>((r * 7 + c * 13) % 97) - 0.4f));
}
__device__ __forceinline__ float st_b(int r, int c) {
  return __uint_as_float(to_tf32(0.02f * static_cast<float>((r * 5 + c * 3) % 89) - 0.7f));
}
__global__ void __launch_bounds__(128, 1) selftest_mn_kernel(float* max_err, int variant) {
  __shared__ __align__(1024) uint8_t tiles[2 * kTileBytes];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ float red[4];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t ta = smem_u32(tiles), tb = ta + kTileBytes;
  const bool a_mn = variant & 1, b_mn = variant & 2;
  const uint32_t lbo = kMnLbo, sbo = kMnSbo;
  for (int i = tid; i < 128 * 32; i += 128) {
    const int mn = i & 127, k = i >> 7;
    const uint32_t va = __float_as_uint(st_a(mn, k)), vb = __float_as_uint(mn < 64 ? st_b(mn, k) : 0.f);
    if (a_mn) sts32(ta + tile_offset_mn(mn, k, lbo, sbo), va);
    else sts32(ta + tile_offset(mn, k >> 2) + ((k & 3) << 2), va);
    if (b_mn) sts32(tb + tile_offset_mn(mn, k, lbo, sbo), vb);
    else sts32(tb + tile_offset(mn, k >> 2) + ((k & 3) << 2), vb);
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&slot, 64);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&slot);
  if (tid == 0) {
    const uint32_t idesc = make_idesc_tf32(128, 64, a_mn, b_mn);
    for (int ks = 0; ks < 4; ++ks) {
      const uint64_t da = a_mn ? make_smem_desc_mn(ta + ks * kMnKStep, lbo, sbo) : make_smem_desc(ta + ks * 32);
      const uint64_t db = b_mn ? make_smem_desc_mn(tb + ks * kMnKStep, lbo, sbo) : make_smem_desc(tb + ks * 32);
      mma_tf32(tmem, da, db, idesc, ks > 0);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after_sync();
  float err = (ta & 1023u) ? 1e30f : 0.f;
  for (int c = 0; c < 4; ++c) {
    float v[16];
    tmem_ld16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c * 16, v);
    for (int j = 0; j < 16; ++j) {
      float ref = 0.f;
      for (int k = 0; k < 32; ++k) ref = fmaf(st_a(tid, k), st_b(c * 16 + j, k), ref);
      err = fmaxf(err, fabsf(ref - v[j]));
    }
  }
  for (int o = 16; o > 0; o >>= 1) err = fmaxf(err, __shfl_xor_sync(0xffffffffu, err, o));
  if (lane == 0) red[warp] = err;
  tc_fence_before_sync();
  __syncthreads();
  if (tid == 0) *max_err = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc(tmem, 64);
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------- host entry points
int tma_fwd(const float* a, int64_t lda, int64_t a_sample_stride, const float* mu_w, const float* sigma_w,
            const float* mu_b, const float* sigma_b, const float* eps_w, const float* eps_b, bnn_view y,
            int64_t y_sample_stride, int M, int N, int K, int S, uint32_t sample_begin, const bnn_rng* rng_w,
            const bnn_rng* rng_b, cudaStream_t st) {
  const bool shared = a_sample_stride == 0;
  if (!tma_ok(a, lda, a_sample_stride) || K % 4 != 0 || rng_w->elem_offset % 4 != 0 || !aligned16(mu_w) ||
      !aligned16(sigma_w) || (eps_w != nullptr && !aligned16(eps_w)))
    return kNotEligible;
  TmaContractParams p{};
  int rc = make_map(&p.map_l, a, K, M, shared ? 1 : S, lda, a_sample_stride, 128);
  if (rc != BNN_OK) return rc;
  p.out.base = y.base; p.out.bs = y.batch_stride; p.out.P = y.P; p.out_sample_stride = y_sample_stride;
  p.mu_w = mu_w; p.sigma_w = sigma_w; p.eps_w = eps_w; p.mu_b = mu_b; p.sigma_b = sigma_b; p.eps_b = eps_b;
  p.M = M; p.N = N; p.K = K; p.S = S; p.sample_begin = sample_begin;
  p.w_numel = static_cast<int64_t>(N) * K;
  p.rng_w = *rng_w; p.rng_b = rng_b ? *rng_b : *rng_w;
  p.shared_l = shared ? 1 : 0;
  p.sum_samples = 0;
#ifdef BNN_PROFILE_WAITS
  { const char* e = getenv("BNN_EXP_FLAGS"); p.exp_flags = e ? atoi(e) : 0; }
#endif
  p.vec_out = (y.P == 1) && (y.batch_stride % 4 == 0) && (y_sample_stride % 4 == 0) && aligned16(y.base);
  return dispatch_tma_contract<false>(p, N, st);
}

int tma_dgrad(bnn_view dy, int64_t dy_sample_stride, const float* mu_w, const float* sigma_w, const float* eps_w,
              float* da, int64_t lda, int64_t a_sample_stride, int M, int N, int K, int S, uint32_t sample_begin,
              const bnn_rng* rng_w, cudaStream_t st) {
  if (dy.P != 1 || !tma_ok(dy.base, dy.batch_stride, dy_sample_stride) || K % 4 != 0 || rng_w->elem_offset % 4 != 0 ||
      !aligned16(mu_w) || !aligned16(sigma_w) || (eps_w != nullptr && !aligned16(eps_w)))
    return kNotEligible;
  TmaContractParams p{};
  int rc = make_map(&p.map_l, dy.base, N, M, S, dy.batch_stride, dy_sample_stride, 128);
  if (rc != BNN_OK) return rc;
  p.out.base = da; p.out.bs = lda; p.out.P = 1; p.out_sample_stride = a_sample_stride;
  p.mu_w = mu_w; p.sigma_w = sigma_w; p.eps_w = eps_w;
  p.M = M; p.N = N; p.K = K; p.S = S; p.sample_begin = sample_begin;
  p.w_numel = static_cast<int64_t>(N) * K;
  p.rng_w = *rng_w; p.rng_b = *rng_w;
  p.shared_l = 0;
  p.sum_samples = (a_sample_stride == 0) ? 1 : 0;
  p.vec_out = (lda % 4 == 0) && (a_sample_stride % 4 == 0) && aligned16(da);
  return dispatch_tma_contract<true>(p, K, st);
}

// Work-unit plan + launch shared by the matrix and the implicit-GEMM (conv) weight gradient
int wgrad_plan_and_launch(TmaWgradParams& p, int M, int N, int K, int S, cudaStream_t st) {
  int rc = BNN_OK;
  const int tiles = ((N + 127) / 128) * ((K + 127) / 128);
  const int m_total = (M + kWgRows - 1) / kWgRows;         // 64-row stages
  // Work units = (sample, M-chunk); a CTA = (tile, group) walks `per` consecutive units, the epilogue of one unit
  // (TMEM drain + eps regeneration for 128 x 128 weights, ~6 stage times) overlapping the MMAs of the next.  The plan
  // minimises  waves x (per x (stages per unit + 0.5) + 6)  over the number of M-chunks and units per CTA: small layers
  // get ONE wave of CTAs with long units instead of several waves of short ones (every extra unit costs an epilogue),
  // and no CTA is launched without work.  Units keep >= 4 stages (256 rows).
  const int sms = sm_count();
  const int max_chunks = (m_total + 3) / 4 < 64 ? (m_total + 3) / 4 : 64;
  double best_cost = 0.0;
  int best_cb = m_total, best_n = 1, best_per = S;
  bool have = false;
  for (int n = 1; n <= (max_chunks < 1 ? 1 : max_chunks); ++n) {
    const int cb = (m_total + n - 1) / n, n_eff = (m_total + cb - 1) / cb;
    if (n_eff != n) continue;
    const int64_t units = static_cast<int64_t>(S) * n;
    for (int c = 0; c < 32; ++c) {
      const int64_t per = c < 16 ? c + 1 : (units + (c - 16)) / (c - 15);        // 1..16, then units / (1..16) rounded up
      if (per < 1 || per > units) continue;
      const int64_t groups = (units + per - 1) / per;
      if (groups > 65535) continue;
      const int64_t ctas = static_cast<int64_t>(tiles) * groups;
      const double waves = static_cast<double>((ctas + sms - 1) / sms);
      const double cost = waves * (static_cast<double>(per) * (cb + 0.5) + 6.0);
      if (!have || cost < best_cost * 0.999 || (cost <= best_cost * 1.001 && ctas < static_cast<int64_t>(tiles) * ((static_cast<int64_t>(S) * best_n + best_per - 1) / best_per))) {
        have = true; best_cost = cost; best_cb = cb; best_n = n; best_per = static_cast<int>(per);
      }
    }
  }
  p.chunk_blocks = best_cb;
  p.n_chunks = best_n;
  const int groups = static_cast<int>((static_cast<int64_t>(S) * best_n + best_per - 1) / best_per);
  static SmemOptIn opt_in[4];
  const bool signs = p.rng_w.row_sign != nullptr;
  rc = allow_dynamic_smem(wgrad_tma_kernel<false, false>, kWgradSmem, &opt_in[0]);
  if (rc == BNN_OK) rc = allow_dynamic_smem(wgrad_tma_kernel<true, false>, kWgradSmem, &opt_in[1]);
  if (rc == BNN_OK) rc = allow_dynamic_smem(wgrad_tma_kernel<false, true>, kWgradSmem, &opt_in[2]);
  if (rc == BNN_OK) rc = allow_dynamic_smem(wgrad_tma_kernel<true, true>, kWgradSmem, &opt_in[3]);
  if (rc != BNN_OK) return rc;
  const int n_tiles = (N + 127) / 128, k_tiles = (K + 127) / 128;
  if (n_tiles >= 2 && pair_enabled() && static_cast<int64_t>(tiles) * groups >= 2 * sm_count()) {
    // CTA pairs over adjacent n-tiles (large problems: the kernel is paced by operand delivery, a pair halves the A^T traffic)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((n_tiles + 1) / 2 * 2, k_tiles, groups);
    cfg.blockDim = dim3(kWgThreads);
    cfg.dynamicSmemBytes = kWgradSmem;
    cfg.stream = st;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    if (signs) BNN_CUDA_OK(cudaLaunchKernelEx(&cfg, wgrad_tma_kernel<true, true>, p));
    else BNN_CUDA_OK(cudaLaunchKernelEx(&cfg, wgrad_tma_kernel<true, false>, p));
  } else {
    if (signs) wgrad_tma_kernel<false, true><<<dim3(n_tiles, k_tiles, groups), kWgThreads, kWgradSmem, st>>>(p);
    else wgrad_tma_kernel<false, false><<<dim3(n_tiles, k_tiles, groups), kWgThreads, kWgradSmem, st>>>(p);
  }
  BNN_CUDA_OK(cudaGetLastError());
  return BNN_OK;
}


int tma_wgrad(bnn_view dy, int64_t dy_sample_stride, const float* a, int64_t lda, int64_t a_sample_stride,
              const float* rho_w, const float* eps_w, float* dmu_w, float* drho_w, int M, int N, int K, int S,
              uint32_t sample_begin, const bnn_rng* rng_w, cudaStream_t st) {
  const bool shared = a_sample_stride == 0;
  if (dy.P != 1 || !tma_ok(dy.base, dy.batch_stride, dy_sample_stride) || !tma_ok(a, lda, a_sample_stride) ||
      K % 4 != 0 || rng_w->elem_offset % 4 != 0 || !aligned16(rho_w) || !aligned16(dmu_w) || !aligned16(drho_w) ||
      (eps_w != nullptr && !aligned16(eps_w)))
    return kNotEligible;
  TmaWgradParams p{};
  int rc = make_map(&p.map_dy, dy.base, N, M, S, dy.batch_stride, dy_sample_stride, kWgRows, true);
  if (rc != BNN_OK) return rc;
  rc = make_map(&p.map_a, a, K, M, shared ? 1 : S, lda, a_sample_stride, kWgRows, true);
  if (rc != BNN_OK) return rc;
  p.rho_w = rho_w; p.eps_w = eps_w; p.dmu_w = dmu_w; p.drho_w = drho_w;
  p.M = M; p.N = N; p.K = K; p.S = S; p.sample_begin = sample_begin;
  p.rng_w = *rng_w;
  p.shared_a = shared ? 1 : 0;
  return wgrad_plan_and_launch(p, M, N, K, S, st);
}

// ---------------------------------------------------------------------------------------------- implicit-GEMM convolution
// The same kernels with the activation operand addressed through an im2col tensor map over the NHWC tensor: no im2col
// matrix exists.  Weights are [Cout][KH][KW][C] (k = (kh, kw, c), c fastest), so a 32-wide k-block is one filter tap and
// 32 consecutive channels — C % 32 == 0 is the eligibility condition.
namespace {
ConvCoords conv_coords_fwd(const bnn_conv2d_nhwc& g) {
  ConvCoords c{};
  c.on = 1; c.P = g.OH * g.OW; c.OW = g.OW; c.imgs = g.B;
  c.sh = g.sh; c.sw = g.sw; c.lh = -g.ph; c.lw = -g.pw; c.dh = g.dh; c.dw = g.dw;
  c.KW = g.KW; c.taps = g.KH * g.KW; c.cblocks = g.C / kBK; c.w_rows = g.Cout; c.chans = g.C;
  return c;
}
bool conv_ok(const bnn_conv2d_nhwc& g) {
  const int lo_h = -g.ph, lo_w = -g.pw, up_h = g.ph - (g.KH - 1) * g.dh, up_w = g.pw - (g.KW - 1) * g.dw;
  return g.C % kBK == 0 && lo_h >= -128 && lo_w >= -128 && up_h >= -128 && up_w >= -128 && up_h <= 127 && up_w <= 127 &&
         (g.KH - 1) * g.dh < 256 && (g.KW - 1) * g.dw < 256 && g.sh >= 1 && g.sw >= 1 && g.sh <= 8 && g.sw <= 8;
}
}  // namespace

int tma_conv_fwd(const float* x, int64_t x_sample_stride, const float* mu_w, const float* sigma_w, const float* mu_b,
                 const float* sigma_b, const float* eps_w, const float* eps_b, bnn_view y, int64_t y_sample_stride,
                 const bnn_conv2d_nhwc* g, int S, uint32_t sample_begin, const bnn_rng* rng_w, const bnn_rng* rng_b,
                 cudaStream_t st) {
  const bool shared = x_sample_stride == 0;
  if (!conv_ok(*g) || !aligned16(x) || rng_w->elem_offset % 4 != 0 || !aligned16(mu_w) || !aligned16(sigma_w) ||
      (eps_w != nullptr && !aligned16(eps_w)))
    return kNotEligible;
  TmaContractParams p{};
  int rc = make_im2col_map(&p.map_l, x, g->C, g->W, g->H, static_cast<int64_t>(shared ? 1 : S) * g->B, -g->pw, -g->ph,
                           g->pw - (g->KW - 1) * g->dw, g->ph - (g->KH - 1) * g->dh, g->sw, g->sh, 128, false);
  if (rc != BNN_OK) return rc;
  p.conv = conv_coords_fwd(*g);
  p.out.base = y.base; p.out.bs = y.batch_stride; p.out.P = y.P; p.out_sample_stride = y_sample_stride;
  p.mu_w = mu_w; p.sigma_w = sigma_w; p.eps_w = eps_w; p.mu_b = mu_b; p.sigma_b = sigma_b; p.eps_b = eps_b;
  p.M = g->B * g->OH * g->OW; p.N = g->Cout; p.K = g->KH * g->KW * g->C; p.S = S; p.sample_begin = sample_begin;
  p.w_numel = static_cast<int64_t>(p.N) * p.K;
  p.rng_w = *rng_w; p.rng_b = rng_b ? *rng_b : *rng_w;
  p.shared_l = shared ? 1 : 0;
  p.sum_samples = 0;
#ifdef BNN_PROFILE_WAITS
  { const char* e = getenv("BNN_EXP_FLAGS"); p.exp_flags = e ? atoi(e) : 0; }
#endif
  p.vec_out = (y.P == 1) && (y.batch_stride % 4 == 0) && (y_sample_stride % 4 == 0) && aligned16(y.base);
  return dispatch_tma_contract<false>(p, p.N, st);
}

// dX[n][h][w][c] = sum_{kh, kw, o} dY[n][h + ph - kh dh][w + pw - kw dw][o] W[o][kh][kw][c]   (stride 1): the transposed-
// filter form — an im2col load of dY with lower corner p - (K - 1) d, taps walked in flipped order — so the input gradient
// is a gather like the forward pass: no column matrix, no scatter.
int tma_conv_dgrad(const float* dy, const float* mu_w, const float* sigma_w, const float* eps_w, float* dx,
                   int64_t x_sample_stride, const bnn_conv2d_nhwc* g, int S, uint32_t sample_begin, const bnn_rng* rng_w,
                   cudaStream_t st) {
  if (g->sh != 1 || g->sw != 1 || g->Cout % kBK != 0 || g->C % 4 != 0 || !aligned16(dy) || !aligned16(dx) ||
      rng_w->elem_offset % 4 != 0 || !aligned16(mu_w) || !aligned16(sigma_w) || (eps_w != nullptr && !aligned16(eps_w)))
    return kNotEligible;
  const int lo_w = g->pw - (g->KW - 1) * g->dw, lo_h = g->ph - (g->KH - 1) * g->dh;
  const int up_w = lo_w + g->W - g->OW, up_h = lo_h + g->H - g->OH;
  if (lo_w < -128 || lo_h < -128 || lo_w > 127 || lo_h > 127 || up_w < -128 || up_h < -128 || up_w > 127 || up_h > 127 ||
      (g->KH - 1) * g->dh >= 256 || (g->KW - 1) * g->dw >= 256)
    return kNotEligible;
  TmaContractParams p{};
  int rc = make_im2col_map(&p.map_l, dy, g->Cout, g->OW, g->OH, static_cast<int64_t>(S) * g->B, lo_w, lo_h, up_w, up_h, 1,
                           1, 128, false);
  if (rc != BNN_OK) return rc;
  ConvCoords c{};
  c.on = 1; c.P = g->H * g->W; c.OW = g->W; c.imgs = g->B; c.sh = 1; c.sw = 1; c.lh = lo_h; c.lw = lo_w;
  c.dh = g->dh; c.dw = g->dw; c.KW = g->KW; c.taps = g->KH * g->KW; c.cblocks = g->Cout / kBK; c.w_rows = g->Cout;
  c.chans = g->Cout;
  p.conv = c;
  p.out.base = dx; p.out.bs = g->C; p.out.P = 1; p.out_sample_stride = x_sample_stride;
  p.mu_w = mu_w; p.sigma_w = sigma_w; p.eps_w = eps_w;
  p.M = g->B * g->H * g->W; p.N = c.taps * g->Cout; p.K = g->C; p.S = S; p.sample_begin = sample_begin;
  p.w_numel = static_cast<int64_t>(g->Cout) * c.taps * g->C;
  p.rng_w = *rng_w; p.rng_b = *rng_w;
  p.shared_l = 0;
  p.sum_samples = (x_sample_stride == 0) ? 1 : 0;
  p.vec_out = (x_sample_stride % 4 == 0);
  return dispatch_tma_contract<true>(p, p.K, st);
}

int tma_conv_wgrad(const float* dy, const float* x, int64_t x_sample_stride, const float* rho_w, const float* eps_w,
                   float* dmu_w, float* drho_w, const bnn_conv2d_nhwc* g, int S, uint32_t sample_begin,
                   const bnn_rng* rng_w, cudaStream_t st) {
  const bool shared = x_sample_stride == 0;
  const int M = g->B * g->OH * g->OW, N = g->Cout, K = g->KH * g->KW * g->C;
  if (!conv_ok(*g) || N % 4 != 0 || !aligned16(dy) || !aligned16(x) || rng_w->elem_offset % 4 != 0 || !aligned16(rho_w) ||
      !aligned16(dmu_w) || !aligned16(drho_w) || (eps_w != nullptr && !aligned16(eps_w)))
    return kNotEligible;
  TmaWgradParams p{};
  int rc = make_map(&p.map_dy, dy, N, M, S, N, static_cast<int64_t>(M) * N, kWgRows, true);
  if (rc != BNN_OK) return rc;
  rc = make_im2col_map(&p.map_a, x, g->C, g->W, g->H, static_cast<int64_t>(shared ? 1 : S) * g->B, -g->pw, -g->ph,
                       g->pw - (g->KW - 1) * g->dw, g->ph - (g->KH - 1) * g->dh, g->sw, g->sh, kWgRows, true);
  if (rc != BNN_OK) return rc;
  p.conv = conv_coords_fwd(*g);
  p.rho_w = rho_w; p.eps_w = eps_w; p.dmu_w = dmu_w; p.drho_w = drho_w;
  p.M = M; p.N = N; p.K = K; p.S = S; p.sample_begin = sample_begin;
  p.rng_w = *rng_w;
  p.shared_a = shared ? 1 : 0;
  return wgrad_plan_and_launch(p, M, N, K, S, st);
}

int tma_wait_counters(unsigned long long* out8, int reset) {
#ifdef BNN_PROFILE_WAITS
  BNN_CUDA_OK(cudaMemcpyFromSymbol(out8, g_wait_cycles, 8 * sizeof(unsigned long long)));
  if (reset) {
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    BNN_CUDA_OK(cudaMemcpyToSymbol(g_wait_cycles, z, sizeof(z)));
  }
  return BNN_OK;
#else
  (void)out8; (void)reset;
  return kNotEligible;
#endif
}

int tma_stage_counters(unsigned long long* out8, int reset) {
#ifdef BNN_PROFILE_WAITS
  BNN_CUDA_OK(cudaMemcpyFromSymbol(out8, g_stage_cycles, 8 * sizeof(unsigned long long)));
  if (reset) {
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, ~0ull, 0};
    BNN_CUDA_OK(cudaMemcpyToSymbol(g_stage_cycles, z, sizeof(z)));
  }
  return BNN_OK;
#else
  (void)out8; (void)reset;
  return kNotEligible;
#endif
}

// host only: the segments contract_pair_sk_kernel's slot `slot` of `slots` would walk (the kernel's own sk_next), six ints
// each: {first row of the leader, row blocks per CTA, column tile, sample, first k-block iteration, iterations}
int tma_balanced_plan(int m_blocks, int S, int gx, int red_blocks, int sum_samples, int slots, int slot, int32_t* out,
                      int max_segments) {
  TmaContractParams p{};
  if (sum_samples) {
    p.sk_msplit = 0; p.sk_mtiles = (m_blocks + 7) / 8; p.sk_its = red_blocks * S;
    p.sk_total = gx * p.sk_mtiles * p.sk_its;
  } else {
    p.sk_msplit = 1; p.sk_mtiles = (m_blocks + 1) / 2; p.sk_its = red_blocks;
    p.sk_total = gx * S * p.sk_mtiles;
  }
  p.sk_gx = gx; p.sk_slots = slots;
  int pos = static_cast<int>(static_cast<long long>(p.sk_total) * slot / slots);
  const int end = static_cast<int>(static_cast<long long>(p.sk_total) * (slot + 1) / slots);
  SkSeg sg;
  int n = 0;
  while (sk_next(p, pos, end, &sg)) {
    if (n < max_segments) {
      int32_t* o = out + 6 * n;
      o[0] = sg.lead_row0; o[1] = sg.mb_cap; o[2] = sg.y; o[3] = sg.z; o[4] = sg.i0; o[5] = sg.n;
    }
    ++n;
  }
  return n;
}

int tma_pair_tile_plan(int m_blocks, int S, int slots, int32_t* out7) {      // host only: the plan the launcher would use
  TilePlan plan{};
  solve_pair_tiles(&plan, m_blocks, S, slots);
  out7[0] = plan.on; out7[1] = plan.n_a; out7[2] = plan.s1; out7[3] = plan.a1; out7[4] = plan.b1; out7[5] = plan.a2; out7[6] = plan.b2;
  return BNN_OK;
}

int tma_force_variant(int variant) {
  const bool ok = variant == -1 || variant == 0 || variant == 1 || variant == 2 || variant == 4 || variant == 8;
  g_forced_variant.store(ok ? variant : -1, std::memory_order_relaxed);
  return BNN_OK;
}

int contract_set_balanced(int on) {
  g_sk_on.store(on ? 1 : 0, std::memory_order_relaxed);
  return BNN_OK;
}
int contract_balanced_state(int slot_cap, int* launches_out, int* slots_out) {
  if (slot_cap >= 0) g_sk_slot_cap.store(slot_cap, std::memory_order_relaxed);
  if (launches_out) *launches_out = g_sk_launches.load(std::memory_order_relaxed);
  if (slots_out) *slots_out = sk_max_slots();
  return BNN_OK;
}

int tma_selftest(float* max_err_dev, cudaStream_t st) {
  selftest_mn_kernel<<<1, 128, 0, st>>>(max_err_dev, 3);
  BNN_CUDA_OK(cudaGetLastError());
  return BNN_OK;
}

}  // namespace contract
}  // namespace bnn
