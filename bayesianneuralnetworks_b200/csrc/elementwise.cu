// elementwise.cu — bandwidth-bound helpers of the variational path:
//   bnn_stddev       sigma = 1e-10 + softplus(rho)                (reference core.py:25-27)
//   bnn_materialize  W_s = mu + sigma * eps(s)                    (reference core.py:44-45)
//   bnn_bias_grad    reparameterised bias gradient                (autograd of dense.py:46-60)
//   bnn_im2col / bnn_col2im  conv2d lowering for the sampled GEMM (conv.py:112-119): staged through shared memory per
//                    (image, channel slice) with table-driven indexing; generic gather kernels as the fallback
#include "common.cuh"

namespace bnn {
namespace {

constexpr int kThreads = 256;

inline int grid_for(int64_t work_items, int per_block, int max_waves = 16) {
  int64_t blocks = (work_items + per_block - 1) / per_block;
  int64_t cap = static_cast<int64_t>(sm_count()) * max_waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

// ------------------------------------------------------------------ stddev
__global__ void __launch_bounds__(kThreads) stddev_kernel(const float* __restrict__ rho,
                                                          float* __restrict__ sigma, int64_t n,
                                                          bool vec) {
  const int64_t tid = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  if (vec) {
    const int64_t n4 = n >> 2;
    for (int64_t i = tid; i < n4; i += stride) {
      const float4 r = ldg_stream4(rho + 4 * i);
      float4 s;
      s.x = stddev_exact(r.x); s.y = stddev_exact(r.y); s.z = stddev_exact(r.z); s.w = stddev_exact(r.w);
      *reinterpret_cast<float4*>(sigma + 4 * i) = s;
    }
    for (int64_t i = (n4 << 2) + tid; i < n; i += stride) sigma[i] = stddev_exact(rho[i]);
  } else {
    for (int64_t i = tid; i < n; i += stride) sigma[i] = stddev_exact(rho[i]);
  }
}

// ------------------------------------------------------------------ materialize
// one thread per (sample, Philox group of 4 consecutive elements)
__global__ void __launch_bounds__(kThreads)
materialize_kernel(const float* __restrict__ mu, const float* __restrict__ sigma,
                   const float* __restrict__ eps_in, float* __restrict__ out,
                   float* __restrict__ eps_out, int64_t numel, int S, uint32_t sample_begin,
                   bnn_rng rng, bool vec) {
  const RngKey key = resolve_rng(rng);
  const int64_t groups = (numel + 3) >> 2;
  const int64_t total = groups * S;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < total; t += stride) {
    const int s = static_cast<int>(t / groups);
    const int64_t g = t - static_cast<int64_t>(s) * groups;
    const int64_t e0 = g << 2;
    const int64_t so = static_cast<int64_t>(s) * numel;
    float4 e;
    if (eps_in != nullptr) {
      e.x = eps_in[so + e0];
      e.y = (e0 + 1 < numel) ? eps_in[so + e0 + 1] : 0.f;
      e.z = (e0 + 2 < numel) ? eps_in[so + e0 + 2] : 0.f;
      e.w = (e0 + 3 < numel) ? eps_in[so + e0 + 3] : 0.f;
    } else {
      e = eps4(key, sample_begin + static_cast<uint32_t>(s), static_cast<uint32_t>(g));
    }
    if (vec) {   // numel % 4 == 0 and all bases 16-byte aligned
      const float4 m = *reinterpret_cast<const float4*>(mu + e0);
      const float4 sd = *reinterpret_cast<const float4*>(sigma + e0);
      float4 w;
      w.x = fmaf(sd.x, e.x, m.x); w.y = fmaf(sd.y, e.y, m.y);
      w.z = fmaf(sd.z, e.z, m.z); w.w = fmaf(sd.w, e.w, m.w);
      *reinterpret_cast<float4*>(out + so + e0) = w;
      if (eps_out != nullptr) *reinterpret_cast<float4*>(eps_out + so + e0) = e;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (e0 + j < numel) {
          const float ej = pick4(e, j);
          out[so + e0 + j] = fmaf(sigma[e0 + j], ej, mu[e0 + j]);
          if (eps_out != nullptr) eps_out[so + e0 + j] = ej;
        }
      }
    }
  }
}

// ------------------------------------------------------------------ bias gradient
// grid (ceil(N/32), S). Row-major dY (P == 1): lanes along n, warps stride over m.
__global__ void __launch_bounds__(kThreads)
bias_grad_rowmajor_kernel(const float* __restrict__ dy, int64_t ld, int64_t sample_stride,
                          const float* __restrict__ rho_b, const float* __restrict__ eps_b,
                          float* __restrict__ dmu_b, float* __restrict__ drho_b, int M, int N,
                          uint32_t sample_begin, bnn_rng rng) {
  __shared__ float part[8][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + lane;
  const int s = blockIdx.y;
  const float* base = dy + static_cast<int64_t>(s) * sample_stride;
  // blockIdx.z splits the rows: the chain c * eps * sigmoid(rho) is linear in c, so partial sums add up
  const int rows_per = (M + gridDim.z - 1) / gridDim.z;
  const int m_begin = blockIdx.z * rows_per;
  const int m_end = m_begin + rows_per < M ? m_begin + rows_per : M;
  float acc = 0.f;
  if (n < N) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;      // four independent loads in flight per thread
    int m = m_begin + warp;
    for (; m + 24 < m_end; m += 32) {
      a0 += base[static_cast<int64_t>(m) * ld + n];
      a1 += base[static_cast<int64_t>(m + 8) * ld + n];
      a2 += base[static_cast<int64_t>(m + 16) * ld + n];
      a3 += base[static_cast<int64_t>(m + 24) * ld + n];
    }
    for (; m < m_end; m += 8) a0 += base[static_cast<int64_t>(m) * ld + n];
    acc = (a0 + a1) + (a2 + a3);
  }
  part[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && n < N) {
    float c = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) c += part[w][lane];
    const RngKey key = resolve_rng(rng);
    const float e = eps_b != nullptr ? eps_b[static_cast<int64_t>(s) * N + n]
                                     : eps1(key, sample_begin + s, static_cast<uint64_t>(n));
    atomicAdd(dmu_b + n, c);
    atomicAdd(drho_b + n, c * e * sigmoid_fast(rho_b[n]));
  }
}

// NCHW dY (P > 1): one block per (n, s); threads stride over m = (b, p), contiguous in p.
__global__ void __launch_bounds__(kThreads)
bias_grad_nchw_kernel(const float* __restrict__ dy, int64_t batch_stride, int P,
                      int64_t sample_stride, const float* __restrict__ rho_b,
                      const float* __restrict__ eps_b, float* __restrict__ dmu_b,
                      float* __restrict__ drho_b, int M, int N, uint32_t sample_begin, bnn_rng rng) {
  __shared__ float part[8];
  const int n = blockIdx.x, s = blockIdx.y;
  const float* base = dy + static_cast<int64_t>(s) * sample_stride + static_cast<int64_t>(n) * P;
  float acc = 0.f;
  for (int m = threadIdx.x; m < M; m += kThreads) {
    const int b = m / P, p = m - b * P;
    acc += base[static_cast<int64_t>(b) * batch_stride + p];
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float c = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) c += part[w];
    const RngKey key = resolve_rng(rng);
    const float e = eps_b != nullptr ? eps_b[static_cast<int64_t>(s) * N + n]
                                     : eps1(key, sample_begin + s, static_cast<uint64_t>(n));
    atomicAdd(dmu_b + n, c);
    atomicAdd(drho_b + n, c * e * sigmoid_fast(rho_b[n]));
  }
}

// ------------------------------------------------------------------ im2col / col2im
__global__ void __launch_bounds__(kThreads)
im2col_kernel(const float* __restrict__ x, float* __restrict__ col, bnn_conv2d_geom g) {
  const int Kg = g.Cg * g.KH * g.KW;
  const int64_t total = static_cast<int64_t>(g.B) * g.OH * g.OW * Kg;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < total; t += stride) {
    const int k = static_cast<int>(t % Kg);
    const int64_t m = t / Kg;
    const int kw = k % g.KW, kh = (k / g.KW) % g.KH, c = k / (g.KW * g.KH);
    const int ow = static_cast<int>(m % g.OW), oh = static_cast<int>((m / g.OW) % g.OH);
    const int b = static_cast<int>(m / (static_cast<int64_t>(g.OW) * g.OH));
    const int ih = oh * g.sh - g.ph + kh * g.dh, iw = ow * g.sw - g.pw + kw * g.dw;
    float v = 0.f;
    if (ih >= 0 && ih < g.H && iw >= 0 && iw < g.W)
      v = x[((static_cast<int64_t>(b) * g.C + g.c0 + c) * g.H + ih) * g.W + iw];
    col[t] = v;
  }
}

__global__ void __launch_bounds__(kThreads)
col2im_kernel(const float* __restrict__ dcol, float* __restrict__ dx, bnn_conv2d_geom g, int accumulate) {
  const int Kg = g.Cg * g.KH * g.KW;
  const int64_t total = static_cast<int64_t>(g.B) * g.Cg * g.H * g.W;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < total; t += stride) {
    const int w = static_cast<int>(t % g.W), h = static_cast<int>((t / g.W) % g.H);
    const int c = static_cast<int>((t / (static_cast<int64_t>(g.W) * g.H)) % g.Cg);
    const int b = static_cast<int>(t / (static_cast<int64_t>(g.W) * g.H * g.Cg));
    float acc = 0.f;
    for (int kh = 0; kh < g.KH; ++kh) {
      const int hn = h + g.ph - kh * g.dh;
      if (hn < 0 || hn % g.sh != 0) continue;
      const int oh = hn / g.sh;
      if (oh >= g.OH) continue;
      for (int kw = 0; kw < g.KW; ++kw) {
        const int wn = w + g.pw - kw * g.dw;
        if (wn < 0 || wn % g.sw != 0) continue;
        const int ow = wn / g.sw;
        if (ow >= g.OW) continue;
        const int64_t m = (static_cast<int64_t>(b) * g.OH + oh) * g.OW + ow;
        acc += dcol[m * Kg + (c * g.KH + kh) * g.KW + kw];
      }
    }
    const int64_t o = ((static_cast<int64_t>(b) * g.C + g.c0 + c) * g.H + h) * g.W + w;
    dx[o] = accumulate ? dx[o] + acc : acc;
  }
}

// Staged variants (the shapes of the examples: small feature maps, many channels).  One block per (image, channel
// slice): the block's input — im2col: the slice's cs x H x W activations; col2im: the slice's columns of the image's
// OH*OW matrix rows — is read once, coalesced, into shared memory; every output element is then produced from shared
// memory with 32-bit index arithmetic and written coalesced (128-bit stores for the matrix rows).  The generic kernels
// above (64-bit div/mod per element, gathers from global memory) remain the fallback for images that do not fit.
constexpr int kStageThreads = 512;
constexpr size_t kStageSmemTarget = 48 * 1024;      // preferred slice size: several blocks per SM
constexpr size_t kStageSmemHard = 160 * 1024;

// Channels per slice `cs` and the number of slices for B images of Cg channels, `bytes_per_channel` of shared memory each.
inline bool plan_slices(int B, int Cg, size_t bytes_per_channel, int* cs_out, int* slices_out) {
  if (bytes_per_channel == 0 || bytes_per_channel > kStageSmemHard) return false;
  int64_t max_cs = static_cast<int64_t>(kStageSmemTarget / bytes_per_channel);
  if (max_cs < 4) max_cs = static_cast<int64_t>(kStageSmemHard / bytes_per_channel);
  if (max_cs > Cg) max_cs = Cg;
  int64_t slices = (2 * static_cast<int64_t>(sm_count()) + B - 1) / B;          // fill the machine about twice
  const int64_t need = (Cg + max_cs - 1) / max_cs;
  if (slices < need) slices = need;
  if (slices > Cg) slices = Cg;
  int64_t cs = (Cg + slices - 1) / slices;
  const int64_t cs4 = (cs + 3) & ~int64_t(3);                                   // multiples of 4 keep the rows 16-byte aligned
  if (static_cast<size_t>(cs4) * bytes_per_channel <= kStageSmemHard) cs = cs4;
  if (static_cast<size_t>(cs) * bytes_per_channel > kStageSmemHard) return false;
  *cs_out = static_cast<int>(cs);
  *slices_out = static_cast<int>((Cg + cs - 1) / cs);
  return true;
}

// Shared-memory layout of both staged kernels: [slab floats, rounded up to 4][lookup table ints].
// The lookup tables replace every per-element division by the kernel geometry (the kernels are instruction bound):
//   im2col: one entry per matrix column j of the slice = slab offset of (c, kh*dh, kw*dw) | kh*dh << 16 | kw*dw << 24
//   col2im: th[h * KH + kh] = oh or -1, tw[w * KW + kw] = ow or -1 (the output position that tap (kh, kw) maps h / w to)
__host__ __device__ inline int stage_round4(int n) { return (n + 3) & ~3; }

template <bool kVec>
__global__ void __launch_bounds__(kStageThreads)
im2col_staged_kernel(const float* __restrict__ x, float* __restrict__ col, bnn_conv2d_geom g, int cs, int vec_in) {
  extern __shared__ __align__(16) float s_stage[];
  const int b = blockIdx.x, c_lo = blockIdx.y * cs;
  const int cn = min(cs, g.Cg - c_lo);
  const int HW = g.H * g.W, KK = g.KH * g.KW, Kg = g.Cg * KK, P = g.OH * g.OW;
  const float* src = x + (static_cast<int64_t>(b) * g.C + g.c0 + c_lo) * HW;      // cn channels, contiguous in NCHW
  const int n_in = cn * HW;
  const int seg = cn * KK;                                   // the columns of every matrix row this block writes
  uint32_t* s_tab = reinterpret_cast<uint32_t*>(s_stage + stage_round4(cs * HW));
  if (vec_in) {
    for (int i = threadIdx.x * 4; i < n_in; i += 4 * kStageThreads)
      *reinterpret_cast<float4*>(s_stage + i) = __ldg(reinterpret_cast<const float4*>(src + i));
  } else {
    for (int i = threadIdx.x; i < n_in; i += kStageThreads) s_stage[i] = __ldg(src + i);
  }
  for (int j = threadIdx.x; j < stage_round4(seg); j += kStageThreads) {
    const int c = j / KK, r = j - c * KK;
    const int kh = r / g.KW, kw = r - kh * g.KW;
    const int khd = kh * g.dh, kwd = kw * g.dw;
    s_tab[j] = static_cast<uint32_t>((c * g.H + khd) * g.W + kwd) | (static_cast<uint32_t>(khd) << 16) |
               (static_cast<uint32_t>(kwd) << 24);
  }
  __syncthreads();
  float* dst = col + static_cast<int64_t>(b) * P * Kg + c_lo * KK;
  auto gather = [&](int ih0, int iw0, int base, uint32_t e) {
    const int ih = ih0 + static_cast<int>((e >> 16) & 0xffu), iw = iw0 + static_cast<int>(e >> 24);
    return (static_cast<unsigned>(ih) < static_cast<unsigned>(g.H) && static_cast<unsigned>(iw) < static_cast<unsigned>(g.W))
               ? s_stage[base + static_cast<int>(e & 0xffffu)] : 0.f;
  };
  if (kVec) {
    const int seg4 = seg >> 2;
    int m = threadIdx.x / seg4, j4 = threadIdx.x - m * seg4;          // one division, then incremental
    const int dm = kStageThreads / seg4, dj = kStageThreads - dm * seg4;
    for (; m < P;) {
      const int oh = m / g.OW, ow = m - oh * g.OW;
      const int ih0 = oh * g.sh - g.ph, iw0 = ow * g.sw - g.pw, base = ih0 * g.W + iw0;
      const uint4 e = *reinterpret_cast<const uint4*>(s_tab + 4 * j4);
      float4 v;
      v.x = gather(ih0, iw0, base, e.x); v.y = gather(ih0, iw0, base, e.y);
      v.z = gather(ih0, iw0, base, e.z); v.w = gather(ih0, iw0, base, e.w);
      *reinterpret_cast<float4*>(dst + static_cast<int64_t>(m) * Kg + 4 * j4) = v;
      m += dm; j4 += dj;
      if (j4 >= seg4) { j4 -= seg4; ++m; }
    }
  } else {
    for (int t = threadIdx.x; t < P * seg; t += kStageThreads) {
      const int m = t / seg, j = t - m * seg;
      const int oh = m / g.OW, ow = m - oh * g.OW;
      const int ih0 = oh * g.sh - g.ph, iw0 = ow * g.sw - g.pw;
      dst[static_cast<int64_t>(m) * Kg + j] = gather(ih0, iw0, ih0 * g.W + iw0, s_tab[j]);
    }
  }
}

template <bool kVec>
__global__ void __launch_bounds__(kStageThreads)
col2im_staged_kernel(const float* __restrict__ dcol, float* __restrict__ dx, bnn_conv2d_geom g, int cs, int accumulate) {
  extern __shared__ __align__(16) float s_stage[];
  const int b = blockIdx.x, c_lo = blockIdx.y * cs;
  const int cn = min(cs, g.Cg - c_lo);
  const int HW = g.H * g.W, KK = g.KH * g.KW, Kg = g.Cg * KK, P = g.OH * g.OW;
  const int seg = cn * KK, pitch = cs * KK;
  const float* src = dcol + static_cast<int64_t>(b) * P * Kg + c_lo * KK;
  int* s_th = reinterpret_cast<int*>(s_stage + stage_round4(P * pitch));
  int* s_tw = s_th + g.H * g.KH;
  if (kVec) {
    const int seg4 = seg >> 2;
    int m = threadIdx.x / seg4, j4 = threadIdx.x - m * seg4;
    const int dm = kStageThreads / seg4, dj = kStageThreads - dm * seg4;
    for (; m < P;) {
      *reinterpret_cast<float4*>(s_stage + m * pitch + 4 * j4) =
          __ldg(reinterpret_cast<const float4*>(src + static_cast<int64_t>(m) * Kg + 4 * j4));
      m += dm; j4 += dj;
      if (j4 >= seg4) { j4 -= seg4; ++m; }
    }
  } else {
    for (int t = threadIdx.x; t < P * seg; t += kStageThreads) {
      const int m = t / seg, j = t - m * seg;
      s_stage[m * pitch + j] = __ldg(src + static_cast<int64_t>(m) * Kg + j);
    }
  }
  for (int i = threadIdx.x; i < g.H * g.KH + g.W * g.KW; i += kStageThreads) {
    const bool is_h = i < g.H * g.KH;
    const int e = is_h ? i : i - g.H * g.KH;
    const int taps = is_h ? g.KH : g.KW, pad = is_h ? g.ph : g.pw, dil = is_h ? g.dh : g.dw;
    const int str = is_h ? g.sh : g.sw, lim = is_h ? g.OH : g.OW;
    const int pos = e / taps, tap = e - pos * taps;
    const int num = pos + pad - tap * dil;
    int o = -1;
    if (num >= 0 && num % str == 0 && num / str < lim) o = num / str;
    s_th[i] = o;                                        // s_tw follows s_th directly
  }
  __syncthreads();
  float* dst = dx + (static_cast<int64_t>(b) * g.C + g.c0 + c_lo) * HW;            // cn channels, contiguous in NCHW
  for (int t = threadIdx.x; t < cn * HW; t += kStageThreads) {
    const int c = t / HW, r = t - c * HW;
    const int h = r / g.W, w = r - h * g.W;
    const float* s_c = s_stage + c * KK;
    float acc = 0.f;                                    // same (kh, kw) order as the generic kernel
    for (int kh = 0; kh < g.KH; ++kh) {
      const int oh = s_th[h * g.KH + kh];
      if (oh < 0) continue;
      for (int kw = 0; kw < g.KW; ++kw) {
        const int ow = s_tw[w * g.KW + kw];
        if (ow < 0) continue;
        acc += s_c[(oh * g.OW + ow) * pitch + kh * g.KW + kw];
      }
    }
    dst[t] = accumulate ? dst[t] + acc : acc;
  }
}

template <typename Kernel>
int allow_stage_smem(Kernel kernel, SmemOptIn* done) {
  return allow_dynamic_smem(kernel, kStageSmemHard + 1024, done);      // + rounding of the two regions
}

int check_geom(const bnn_conv2d_geom* g) {
  BNN_REQUIRE(g != nullptr, BNN_ERR_BAD_ARGUMENT, "conv geometry is NULL");
  BNN_REQUIRE(g->B > 0 && g->C > 0 && g->H > 0 && g->W > 0 && g->Cg > 0 && g->c0 >= 0 &&
                  g->c0 + g->Cg <= g->C && g->KH > 0 && g->KW > 0 && g->OH > 0 && g->OW > 0 &&
                  g->sh > 0 && g->sw > 0 && g->dh > 0 && g->dw > 0 && g->ph >= 0 && g->pw >= 0,
              BNN_ERR_BAD_ARGUMENT, "invalid conv2d geometry");
  return BNN_OK;
}

}  // namespace
}  // namespace bnn

using namespace bnn;

extern "C" {

int bnn_stddev(const float* rho, float* sigma, int64_t numel, void* stream) {
  BNN_REQUIRE(numel >= 0, BNN_ERR_BAD_ARGUMENT, "numel < 0");
  if (numel == 0) return BNN_OK;
  BNN_REQUIRE(rho && sigma, BNN_ERR_BAD_ARGUMENT, "bnn_stddev: NULL pointer");
  int rc = check_device();
  if (rc != BNN_OK) return rc;
  const bool vec = aligned16(rho) && aligned16(sigma);
  stddev_kernel<<<grid_for(numel, kThreads * 4), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      rho, sigma, numel, vec);
  BNN_CUDA_OK(cudaGetLastError());
  return BNN_OK;
}

int bnn_materialize(const float* mu, const float* sigma, const float* eps_in, float* out,
                    float* eps_out, int64_t numel, int32_t S, uint32_t sample_begin,
                    const bnn_rng* rng, void* stream) {
  BNN_REQUIRE(numel >= 0 && S >= 0, BNN_ERR_BAD_ARGUMENT, "negative size");
  if (numel == 0 || S == 0) return BNN_OK;
  BNN_REQUIRE(mu && sigma && out && rng, BNN_ERR_BAD_ARGUMENT, "bnn_materialize: NULL pointer");
  BNN_REQUIRE(numel <= (int64_t(1) << 34), BNN_ERR_UNSUPPORTED, "tensor larger than 2^34 elements");
  int rc = check_device();
  if (rc != BNN_OK) return rc;
  const bool vec = (numel % 4 == 0) && aligned16(mu) && aligned16(sigma) && aligned16(out) &&
                   (eps_out == nullptr || aligned16(eps_out));
  const int64_t items = ((numel + 3) / 4) * S;
  materialize_kernel<<<grid_for(items, kThreads), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      mu, sigma, eps_in, out, eps_out, numel, S, sample_begin, *rng, vec);
  BNN_CUDA_OK(cudaGetLastError());
  return BNN_OK;
}

int bnn_bias_grad(bnn_view dy, int64_t dy_sample_stride, const float* rho_b, const float* eps_b,
                  float* dmu_b, float* drho_b, int32_t M, int32_t N, int32_t S,
                  uint32_t sample_begin, const bnn_rng* rng_b, void* stream) {
  BNN_REQUIRE(M >= 0 && N >= 0 && S >= 0, BNN_ERR_BAD_ARGUMENT, "negative size");
  if (M == 0 || N == 0 || S == 0) return BNN_OK;
  BNN_REQUIRE(dy.base && rho_b && dmu_b && drho_b && rng_b, BNN_ERR_BAD_ARGUMENT, "bnn_bias_grad: NULL pointer");
  BNN_REQUIRE(dy.P >= 1, BNN_ERR_BAD_ARGUMENT, "view P must be >= 1");
  BNN_REQUIRE(S <= 65535, BNN_ERR_UNSUPPORTED, "more than 65535 samples per launch");
  int rc = check_device();
  if (rc != BNN_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dy.P == 1) {
    const int col_blocks = (N + 31) / 32;
    int chunks = (2 * sm_count() + col_blocks * S - 1) / (col_blocks * S);     // fill the machine about twice
    const int max_chunks = (M + 63) / 64;                                      // at least 64 rows per block
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    dim3 grid(col_blocks, S, chunks);
    bias_grad_rowmajor_kernel<<<grid, kThreads, 0, st>>>(dy.base, dy.batch_stride, dy_sample_stride,
                                                        rho_b, eps_b, dmu_b, drho_b, M, N,
                                                        sample_begin, *rng_b);
  } else {
    dim3 grid(N, S);
    bias_grad_nchw_kernel<<<grid, kThreads, 0, st>>>(dy.base, dy.batch_stride, dy.P,
                                                    dy_sample_stride, rho_b, eps_b, dmu_b, drho_b,
                                                    M, N, sample_begin, *rng_b);
  }
  BNN_CUDA_OK(cudaGetLastError());
  return BNN_OK;
}

int bnn_im2col(const float* x, float* col, const bnn_conv2d_geom* g, void* stream) {
  int rc = check_geom(g);
  if (rc != BNN_OK) return rc;
  BNN_REQUIRE(x && col, BNN_ERR_BAD_ARGUMENT, "bnn_im2col: NULL pointer");
  rc = check_device();
  if (rc != BNN_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t HW = static_cast<int64_t>(g->H) * g->W, KK = static_cast<int64_t>(g->KH) * g->KW;
  const int64_t Kg = g->Cg * KK, P = static_cast<int64_t>(g->OH) * g->OW;
  int cs = 0, slices = 0;
  // table entry: slab offset < 2^16, dilated tap offsets < 2^8
  if (P * Kg < (int64_t(1) << 30) && static_cast<int64_t>(g->KH - 1) * g->dh < 256 && static_cast<int64_t>(g->KW - 1) * g->dw < 256 &&
      plan_slices(g->B, g->Cg, static_cast<size_t>(HW + KK) * 4, &cs, &slices) && slices <= 65535 &&
      static_cast<int64_t>(cs) * HW + static_cast<int64_t>(g->KH) * g->dh * g->W < 65536) {
    const bool vec = Kg % 4 == 0 && (static_cast<int64_t>(cs) * KK) % 4 == 0 && aligned16(col);
    const int vec_in = (static_cast<int64_t>(cs) * HW) % 4 == 0 && (static_cast<int64_t>(g->C) * HW) % 4 == 0 &&
                       (static_cast<int64_t>(g->c0) * HW) % 4 == 0 && (static_cast<int64_t>(g->Cg - (slices - 1) * cs) * HW) % 4 == 0 &&
                       aligned16(x);
    const size_t smem = (static_cast<size_t>(stage_round4(static_cast<int>(cs * HW))) + stage_round4(static_cast<int>(cs * KK))) * 4;
    const dim3 grid(g->B, slices);
    static SmemOptIn attr_vec, attr_scalar;
    if (vec) {
      rc = allow_stage_smem(im2col_staged_kernel<true>, &attr_vec);
      if (rc != BNN_OK) return rc;
      im2col_staged_kernel<true><<<grid, kStageThreads, smem, st>>>(x, col, *g, cs, vec_in);
    } else {
      rc = allow_stage_smem(im2col_staged_kernel<false>, &attr_scalar);
      if (rc != BNN_OK) return rc;
      im2col_staged_kernel<false><<<grid, kStageThreads, smem, st>>>(x, col, *g, cs, vec_in);
    }
  } else {
    const int64_t total = static_cast<int64_t>(g->B) * P * Kg;
    im2col_kernel<<<grid_for(total, kThreads), kThreads, 0, st>>>(x, col, *g);
  }
  BNN_CUDA_OK(cudaGetLastError());
  return BNN_OK;
}

int bnn_col2im(const float* dcol, float* dx, const bnn_conv2d_geom* g, int32_t accumulate, void* stream) {
  int rc = check_geom(g);
  if (rc != BNN_OK) return rc;
  BNN_REQUIRE(dcol && dx, BNN_ERR_BAD_ARGUMENT, "bnn_col2im: NULL pointer");
  rc = check_device();
  if (rc != BNN_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t HW = static_cast<int64_t>(g->H) * g->W, KK = static_cast<int64_t>(g->KH) * g->KW;
  const int64_t Kg = g->Cg * KK, P = static_cast<int64_t>(g->OH) * g->OW;
  int cs = 0, slices = 0;
  if (static_cast<int64_t>(g->Cg) * HW < (int64_t(1) << 30) &&
      static_cast<int64_t>(g->H) * g->KH + static_cast<int64_t>(g->W) * g->KW <= 2048 &&
      plan_slices(g->B, g->Cg, static_cast<size_t>(P * KK) * 4, &cs, &slices) && slices <= 65535 &&
      static_cast<size_t>(cs) * P * KK * 4 + 8192 + 16 <= kStageSmemHard) {
    const bool vec = Kg % 4 == 0 && (static_cast<int64_t>(cs) * KK) % 4 == 0 && aligned16(dcol);
    const size_t smem = (static_cast<size_t>(stage_round4(static_cast<int>(cs * P * KK))) + g->H * g->KH + g->W * g->KW) * 4;
    const dim3 grid(g->B, slices);
    static SmemOptIn attr_vec, attr_scalar;
    if (vec) {
      rc = allow_stage_smem(col2im_staged_kernel<true>, &attr_vec);
      if (rc != BNN_OK) return rc;
      col2im_staged_kernel<true><<<grid, kStageThreads, smem, st>>>(dcol, dx, *g, cs, accumulate);
    } else {
      rc = allow_stage_smem(col2im_staged_kernel<false>, &attr_scalar);
      if (rc != BNN_OK) return rc;
      col2im_staged_kernel<false><<<grid, kStageThreads, smem, st>>>(dcol, dx, *g, cs, accumulate);
    }
  } else {
    const int64_t total = static_cast<int64_t>(g->B) * g->Cg * HW;
    col2im_kernel<<<grid_for(total, kThreads), kThreads, 0, st>>>(dcol, dx, *g, accumulate);
  }
  BNN_CUDA_OK(cudaGetLastError());
  return BNN_OK;
}

}  // extern "C"
