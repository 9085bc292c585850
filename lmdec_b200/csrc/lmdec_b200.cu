// C-ABI implementation (include/lmdec_b200.h): tile-store kernels, operand quantisation, the tcgen05 product, and the
// small fp64 reductions of the power iteration.  sm_100a only; there is no CPU path in this library.
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <atomic>
#include <vector>
#include <utility>
#include "../../include/lmdec_b200.h"
#include "geno_mma.cuh"
#include "dense_mma.cuh"

namespace {
thread_local char g_err[512] = "";
int fail(const char* fmt, const char* a = "") {
  snprintf(g_err, sizeof(g_err), fmt, a);
  return 1;
}
std::atomic<long long> g_launches{0};
int check_launch(const char* what, int n_kernels = 1) {
  g_launches.fetch_add(n_kernels, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return 2;
  }
  return 0;
}
inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
}  // namespace

// Per-caller context (include/lmdec_b200.h): every tunable and the kernel-timing events live here, not in globals.
struct lmdec_ctx {
  int n_sm = 0;
  int sm_reserve = 0;        // SMs left free for concurrently running collectives (row-sharded runs)
  int pair_mode = 1;         // 0 disables the two-tiles-per-B-stage schedule (A/B tests)
  int pair_min_chunks = 64;  // shortest contraction (256-wide chunks) that pairs
  int issuer_groups = 2;     // single-tile work items: issuer pairs taking alternate chunks (1 = one pair)
  int producer_groups = 2;   // 2: two producer groups on alternate tile-chunks when there is no missing data
  int producer_groups_missing = 1;   // the same with missing data (needs a decoded-A ring of >= 3 stages to pay)
  int acc_sets = 0;          // accumulator sets of single-tile work items: 0 = choose, 1, 2
  int b_resident = 1;        // short unsplit contractions keep the whole operand image in shared memory
  int rr = 1;                // no missing data: round-robin whole-chunk issuers, one accumulator per tile
  bool timing = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> events;
};
static const lmdec_ctx kDefaultCtx{};

namespace lmdec {

// ------------------------------------------------------------------------------------------------ tile store
__host__ __device__ inline size_t tile_offset(int64_t mp, int64_t kc, int64_t n_kc) {
  return ((size_t)mp * n_kc + kc) * 8192;
}

template <typename T>
__device__ __forceinline__ uint32_t code_of(T v, int* bad) {
  if (v == (T)0) return 0;
  if (v == (T)1) return 1;
  if (v == (T)2) return 2;
  if (v != v) return 3;  // NaN
  *bad = 1;
  return 3;
}
template <>
__device__ __forceinline__ uint32_t code_of<int8_t>(int8_t v, int* bad) {
  if (v >= 0 && v <= 2) return (uint32_t)v;
  if (v < 0 || v == 3) return 3;
  *bad = 1;
  return 3;
}
template <>
__device__ __forceinline__ uint32_t code_of<uint8_t>(uint8_t v, int* bad) {
  if (v <= 3) return v;
  if (v == 255) return 3;
  *bad = 1;
  return 3;
}

// one thread = one 16-byte unit (64 codes) of the store
template <typename T>
__global__ void pack_kernel(const T* __restrict__ src, int64_t stride_m, int64_t stride_kd, int64_t block_m,
                            int64_t block_kd, int64_t m0, int64_t kd0, uint8_t* __restrict__ store, int64_t n_kc,
                            int64_t mp0, int64_t kc0, int64_t n_mp_blk, int64_t n_kc_blk, int32_t* bad_count) {
  const int64_t unit = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n_units = n_mp_blk * n_kc_blk * 512;
  if (unit >= n_units) return;
  const int t = (int)(unit & 127);
  const int piece = (int)((unit >> 7) & 3);
  const int64_t tile = unit >> 9;
  const int64_t kc = kc0 + tile % n_kc_blk, mp = mp0 + tile / n_kc_blk;
  const int64_t m = mp * 128 + t - m0;           // row inside the block
  const int64_t kd = kc * 256 + piece * 64 - kd0;  // first contraction index inside the block
  uint32_t w[4] = {0, 0, 0, 0};
  int bad = 0;
  if (m >= 0 && m < block_m) {
    const T* row = src + m * stride_m;
#pragma unroll 4
    for (int e = 0; e < 64; ++e) {
      const int64_t g = kd + e;
      if (g >= 0 && g < block_kd) w[e >> 4] |= code_of<T>(row[g * stride_kd], &bad) << (2 * (e & 15));
    }
  }
  uint4 v = make_uint4(w[0], w[1], w[2], w[3]);
  *reinterpret_cast<uint4*>(store + tile_offset(mp, kc, n_kc) + piece * 2048 + t * 16) = v;
  if (bad) atomicAdd(bad_count, 1);
}

__global__ void unpack_kernel(const uint8_t* __restrict__ store, int64_t M, int64_t Kd, int64_t n_kc,
                              int8_t* __restrict__ dst, int64_t ld, int64_t n_units) {
  const int64_t unit = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (unit >= n_units) return;
  const int t = (int)(unit & 127);
  const int piece = (int)((unit >> 7) & 3);
  const int64_t tile = unit >> 9;
  const int64_t kc = tile % n_kc, mp = tile / n_kc;
  const int64_t m = mp * 128 + t;
  if (m >= M) return;
  const uint4 v = *reinterpret_cast<const uint4*>(store + tile_offset(mp, kc, n_kc) + piece * 2048 + t * 16);
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  for (int e = 0; e < 64; ++e) {
    const int64_t g = kc * 256 + piece * 64 + e;
    if (g < Kd) {
      const uint32_t c = (w[e >> 4] >> (2 * (e & 15))) & 3u;
      dst[m * ld + g] = c == 3 ? (int8_t)-1 : (int8_t)c;
    }
  }
}

// counts per store row; one thread per row, coalesced 16-byte units across the warp
__global__ void code_counts_kernel(const uint8_t* __restrict__ store, int64_t M, int64_t n_kc, int64_t kd_limit,
                                   long long* __restrict__ counts) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const int64_t mp = m >> 7;
  const int t = (int)(m & 127);
  const int64_t full_units = kd_limit / 64;  // 16-byte units entirely below the limit
  long long c1 = 0, c2 = 0, c3 = 0;
  for (int64_t u = 0; u * 64 < kd_limit; ++u) {
    const int64_t kc = u >> 2;
    const int piece = (int)(u & 3);
    uint4 v = *reinterpret_cast<const uint4*>(store + tile_offset(mp, kc, n_kc) + piece * 2048 + t * 16);
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
    if (u >= full_units) {  // ragged tail: mask codes at or above the limit
      const int keep = (int)(kd_limit - u * 64);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = keep - 16 * i;
        if (k <= 0)
          w[i] = 0;
        else if (k < 16)
          w[i] &= (1u << (2 * k)) - 1u;
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t lo = w[i] & 0x55555555u, hi = (w[i] >> 1) & 0x55555555u;
      c3 += __popc(lo & hi);
      c1 += __popc(lo & ~hi);
      c2 += __popc(hi & ~lo);
    }
  }
  counts[m * 4 + 0] = kd_limit - c1 - c2 - c3;
  counts[m * 4 + 1] = c1;
  counts[m * 4 + 2] = c2;
  counts[m * 4 + 3] = c3;
}

// .bed (variant-major, PLINK codes) -> SNP-major tile store (M = variants, Kd = samples)
__global__ void recode_bed_kernel(const uint8_t* __restrict__ bed, int64_t n_variants, int64_t n_samples,
                                  int64_t bytes_per_variant, uint8_t* __restrict__ store, int64_t n_kc,
                                  int64_t n_units) {
  const int64_t unit = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (unit >= n_units) return;
  const int t = (int)(unit & 127);
  const int piece = (int)((unit >> 7) & 3);
  const int64_t tile = unit >> 9;
  const int64_t kc = tile % n_kc, mp = tile / n_kc;
  const int64_t j = mp * 128 + t;
  uint32_t w[4] = {0, 0, 0, 0};
  if (j < n_variants) {
    const uint8_t* row = bed + j * bytes_per_variant;
    const int64_t b0 = kc * 64 + piece * 16;
    for (int b = 0; b < 16; ++b) {
      if (b0 + b < bytes_per_variant) {
        uint32_t x = row[b0 + b];
        // PLINK 00->0, 01->missing(3), 10->1, 11->2 :  hi' = lo, lo' = lo ^ hi  per 2-bit field
        const uint32_t lo = x & 0x55u, hi = (x >> 1) & 0x55u;
        uint32_t y = ((lo) << 1) | (lo ^ hi);
        // codes of samples >= n_samples (byte padding) -> 0
        const int64_t s0 = (b0 + b) * 4;
        for (int e = 0; e < 4; ++e)
          if (s0 + e >= n_samples) y &= ~(3u << (2 * e));
        w[b >> 2] |= y << (8 * (b & 3));
      }
    }
  }
  *reinterpret_cast<uint4*>(store + tile_offset(mp, kc, n_kc) + piece * 2048 + t * 16) =
      make_uint4(w[0], w[1], w[2], w[3]);
}

// transpose: dst (Kd x M) unit (row g of dst, 64 codes along m) gathered from src; one thread per dst unit
__global__ void store_transpose_kernel(const uint8_t* __restrict__ src, int64_t M, int64_t Kd, int64_t src_kc,
                                       uint8_t* __restrict__ dst, int64_t dst_kc, int64_t n_units) {
  const int64_t unit = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (unit >= n_units) return;
  const int t = (int)(unit & 127);
  const int piece = (int)((unit >> 7) & 3);
  const int64_t tile = unit >> 9;
  const int64_t kc = tile % dst_kc, mp = tile / dst_kc;
  const int64_t g = mp * 128 + t;          // dst row = src contraction index
  const int64_t m_first = kc * 256 + piece * 64;  // dst contraction = src row
  uint32_t w[4] = {0, 0, 0, 0};
  if (g < Kd) {
    const int64_t s_kc = g >> 8;
    const int s_b = (int)((g & 255) >> 2), s_sh = 2 * (int)(g & 3);
    for (int e = 0; e < 64; ++e) {
      const int64_t m = m_first + e;
      if (m < M) {
        const uint8_t byte = src[tile_offset(m >> 7, s_kc, src_kc) + (s_b >> 4) * 2048 + (m & 127) * 16 + (s_b & 15)];
        w[e >> 4] |= (uint32_t)((byte >> s_sh) & 3u) << (2 * (e & 15));
      }
    }
  }
  *reinterpret_cast<uint4*>(dst + tile_offset(mp, kc, dst_kc) + piece * 2048 + t * 16) =
      make_uint4(w[0], w[1], w[2], w[3]);
}

// ------------------------------------------------------------------------------------------------ operand planes
// one thread = 16 contraction positions (K-block kb, half h) x one operand column k: it reads its 16 values once and
// writes the P plane units of image 1 (and of image 2).  Unit (kb, h, n) sits at ((kb*2 + h)*N + n)*16 bytes; byte e of
// a unit is MMA contraction index 16h + e, which the producers' decode maps to genotype position
// 16h + (e>>2) + 4*(e&3) of the K-block (geno_mma.cuh decode16).  Column n = k*P + plane.
// Fixed point: q = round(x mult) (|q| < 2^(8P-2) by the choice of mult), balanced base-256 digits d_i in [-128, 127].
// All P digits of q at once: the bytes of q + sum_i 128*256^i are d_i + 128, so (q + bias) ^ bias holds digit i in
// byte i.  Four rows' digit words are then transposed into one 32-bit word per plane with byte permutes.
template <int P>
__device__ __forceinline__ uint32_t digit_bytes(int q) {
  constexpr uint32_t kBias = P == 1 ? 0x80u : P == 2 ? 0x8080u : P == 3 ? 0x808080u : 0x80808080u;
  return ((uint32_t)q + kBias) ^ kBias;
}
template <int P>
__device__ __forceinline__ void planes_of_four(const uint32_t* t, uint32_t (&w)[4]) {
  const uint32_t a = __byte_perm(t[0], t[1], 0x5140), b = __byte_perm(t[2], t[3], 0x5140);
  w[0] = __byte_perm(a, b, 0x5410);
  if (P > 1) w[1] = __byte_perm(a, b, 0x7632);
  if (P > 2) {
    const uint32_t c = __byte_perm(t[0], t[1], 0x7362), d = __byte_perm(t[2], t[3], 0x7362);
    w[2] = __byte_perm(c, d, 0x5410);
    if (P > 3) w[3] = __byte_perm(c, d, 0x7632);
  }
}

template <int P, typename TB, bool kFull>
__device__ __forceinline__ void quantize_rows(const TB* __restrict__ bp, int64_t ldb, int64_t g0, int64_t kd_limit,
                                              double mk, const double* __restrict__ row1,
                                              const double* __restrict__ row2, bool two, uint32_t (&t1)[16],
                                              uint32_t (&t2)[16]) {
  // q = round(x) for |x| < 2^31 is the low word of x + 1.5*2^52 (one DADD, round-to-nearest-even like rint)
  constexpr double kMagic = 6755399441055744.0;
  // largest magnitude P balanced base-256 digits can hold: 127 * (256^P - 1) / 255
  constexpr int kQmax = P == 1 ? 127 : P == 2 ? 32639 : P == 3 ? 8355711 : 2139062143;
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    const int r = (e >> 2) + 4 * (e & 3);
    if (kFull || g0 + r < kd_limit) {
      double x = (double)bp[r * ldb] * mk;
      if (row1) x *= row1[g0 + r];
      const int q1 = __double2loint(x + kMagic);
      t1[e] = digit_bytes<P>(q1);
      if (two) {
        // the product kernel multiplies RAW codes (3 at a missing entry) by image 1: image 2 carries q2 - 3 q1
        const int q2 = __double2loint(x * row2[g0 + r] + kMagic) - 3 * q1;
        t2[e] = digit_bytes<P>(min(max(q2, -kQmax), kQmax));
      }
    }
  }
}

// grid: block b takes groups [b*gpb, (b+1)*gpb), thread (tg, k) = threadIdx.x / kq, % kq; kq = N / P columns per group
// (k < K real, K <= k < kq zero padding); the thread of the last column also zeroes columns [kq*P, N).
template <int P, typename TB>
__global__ void __launch_bounds__(256) quantize_planes_kernel(const TB* __restrict__ B, int64_t ldb,
                                                              int64_t kd_limit, int K, int N,
                                                              const double* __restrict__ mult,
                                                              const double* __restrict__ row1,
                                                              const double* __restrict__ row2,
                                                              int8_t* __restrict__ image1, int8_t* __restrict__ image2,
                                                              int64_t n_groups, int kq, int gpb) {
  const int tg = (int)threadIdx.x / kq;
  const int k = (int)threadIdx.x - tg * kq;
  const int64_t grp = (int64_t)blockIdx.x * gpb + tg;   // grp = kb*2 + h
  if (tg >= gpb || grp >= n_groups) return;
  uint32_t t1[16], t2[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) t1[e] = t2[e] = 0;
  const bool two = image2 != nullptr;
  if (k < K) {
    const double mk = mult[k];
    const int64_t g0 = grp * 16;
    const TB* bp = B + g0 * ldb + k;
    if (g0 + 16 <= kd_limit) quantize_rows<P, TB, true>(bp, ldb, g0, kd_limit, mk, row1, row2, two, t1, t2);
    else if (g0 < kd_limit) quantize_rows<P, TB, false>(bp, ldb, g0, kd_limit, mk, row1, row2, two, t1, t2);
  }
  const size_t base = ((size_t)grp * N + (size_t)k * P) * 16;
  uint32_t w1[4][4], w2[4][4];   // [row quad][plane]
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    planes_of_four<P>(t1 + 4 * j, w1[j]);
    if (two) planes_of_four<P>(t2 + 4 * j, w2[j]);
  }
#pragma unroll
  for (int pl = 0; pl < P; ++pl) {
    *reinterpret_cast<uint4*>(image1 + base + pl * 16) = make_uint4(w1[0][pl], w1[1][pl], w1[2][pl], w1[3][pl]);
    if (two) *reinterpret_cast<uint4*>(image2 + base + pl * 16) = make_uint4(w2[0][pl], w2[1][pl], w2[2][pl], w2[3][pl]);
  }
  if (k == kq - 1) {
    // columns [kq*P, N) are not covered by the groups when P does not divide N: zero them
    for (int n = kq * P; n < N; ++n) {
      const size_t off = ((size_t)grp * N + n) * 16;
      *reinterpret_cast<uint4*>(image1 + off) = make_uint4(0, 0, 0, 0);
      if (two) *reinterpret_cast<uint4*>(image2 + off) = make_uint4(0, 0, 0, 0);
    }
  }
}

__global__ void colabsmax_kernel(const double* __restrict__ B, int64_t ldb, int64_t kd_limit, int K,
                                 const double* __restrict__ row1, const double* __restrict__ row2,
                                 unsigned long long* __restrict__ out, unsigned long long* __restrict__ out2) {
  const int k = threadIdx.x % K, ry = threadIdx.x / K, rpb = blockDim.x / K;
  if (ry >= rpb) return;
  double m1 = 0.0, m2 = 0.0;
  for (int64_t r = (int64_t)blockIdx.x * rpb + ry; r < kd_limit; r += (int64_t)gridDim.x * rpb) {
    double v = fabs(B[r * ldb + k]);
    if (row1) v *= fabs(row1[r]);
    if (v != v) v = INFINITY;  // NaN poisons the scale so the caller can raise LinAlgError
    m1 = fmax(m1, v);
    if (row2) m2 = fmax(m2, v * fabs(row2[r]));
  }
  atomicMax(out + k, (unsigned long long)__double_as_longlong(m1));
  if (row2) atomicMax(out2 + k, (unsigned long long)__double_as_longlong(m2));
}

// Column statistics of a float64 operand in one pass: max |B r1|, max |B r1 r2| and sum_r w_r B_rk per column.
// Deterministic: per-block partials (fixed row -> block map) reduced in block order by colstats_finish_kernel, which
// also derives the fixed-point multiplier, its inverse and the rank-1 shift used by the finalize kernel.
constexpr int kStatsMaxBlocks = 592;
template <typename TB>
__global__ void __launch_bounds__(256) colstats_partial_kernel(const TB* __restrict__ B, int64_t ldb,
                                                               int64_t kd_limit, int K, const double* __restrict__ row1,
                                                               const double* __restrict__ row2,
                                                               const double* __restrict__ w, int use_w,
                                                               double* __restrict__ partial) {
  extern __shared__ double sm[];
  const int k = threadIdx.x % K, ry = threadIdx.x / K, rpb = blockDim.x / K;
  double m1 = 0.0, m2 = 0.0, ws = 0.0;
  if (ry < rpb) {
    const int64_t stride = (int64_t)gridDim.x * rpb;
    constexpr int U = 8;  // independent loads in flight per thread
    for (int64_t r0 = (int64_t)blockIdx.x * rpb + ry; r0 < kd_limit; r0 += stride * U) {
      double b[U], a1[U], a2[U], ww[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t r = r0 + u * stride;
        const bool in = r < kd_limit;
        b[u] = in ? (double)B[r * ldb + k] : 0.0;
        a1[u] = (in && row1) ? row1[r] : 1.0;
        a2[u] = (in && row2) ? row2[r] : 0.0;
        ww[u] = (in && w) ? w[r] : 1.0;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        double v = fabs(b[u]) * fabs(a1[u]);
        if (v != v) v = INFINITY;  // NaN poisons the scale: the Gram check on the host side raises LinAlgError
        m1 = fmax(m1, v);
        m2 = fmax(m2, v * fabs(a2[u]));
        if (use_w) ws += ww[u] * b[u];
      }
    }
  }
  double* s1 = sm;
  double* s2 = sm + blockDim.x;
  double* s3 = sm + 2 * blockDim.x;
  s1[threadIdx.x] = m1;
  s2[threadIdx.x] = m2;
  s3[threadIdx.x] = ws;
  __syncthreads();
  if (threadIdx.x < K) {
    for (int j = 1; j < rpb; ++j) {
      m1 = fmax(m1, s1[j * K + k]);
      m2 = fmax(m2, s2[j * K + k]);
      ws += s3[j * K + k];
    }
    double* out = partial + (size_t)blockIdx.x * 3 * K;
    out[k] = m1;
    out[K + k] = m2;
    out[2 * K + k] = ws;
  }
}
// stats layout: [0,K) mult, [K,2K) 1/mult, [2K,3K) shift = shift_sign * wsum, [3K,4K) amax, [4K,5K) amax2 (informative:
// only max(amax, amax2) enters mult; the fused producers report that maximum as amax and 0 as amax2)
// 1024 threads: thread (slice, k) folds blocks slice, slice+S, ... ; thread k then folds the slices in order.
__global__ void __launch_bounds__(1024) colstats_finish_kernel(const double* __restrict__ partial, int n_blocks, int K,
                                                               double lim, double shift_sign,
                                                               double* __restrict__ stats) {
  extern __shared__ double sm[];
  const int n_sl = blockDim.x / K;
  const int k = threadIdx.x % K, sl = threadIdx.x / K;
  double m1 = 0.0, m2 = 0.0, ws = 0.0;
  if (sl < n_sl) {
    for (int b = sl; b < n_blocks; b += n_sl) {
      const double* in = partial + (size_t)b * 3 * K;
      m1 = fmax(m1, in[k]);
      m2 = fmax(m2, in[K + k]);
      ws += in[2 * K + k];
    }
    sm[(sl * 3 + 0) * K + k] = m1;
    sm[(sl * 3 + 1) * K + k] = m2;
    sm[(sl * 3 + 2) * K + k] = ws;
  }
  __syncthreads();
  if (threadIdx.x < K) {
    for (int j = 1; j < n_sl; ++j) {
      m1 = fmax(m1, sm[(j * 3 + 0) * K + k]);
      m2 = fmax(m2, sm[(j * 3 + 1) * K + k]);
      ws += sm[(j * 3 + 2) * K + k];
    }
    const double s_ = fmax(m1, m2);
    const double mult = s_ > 0.0 ? lim / s_ : 1.0;  // a zero column quantises to zero with any multiplier
    stats[k] = mult;
    stats[K + k] = 1.0 / mult;
    stats[2 * K + k] = shift_sign * ws;
    stats[3 * K + k] = m1;
    stats[4 * K + k] = m2;
  }
}

template <typename TO>
__global__ void finalize_kernel(const long long* __restrict__ acc0, const long long* __restrict__ acc1, int64_t M, int K,
                                const double* __restrict__ colscale, const double* __restrict__ rowa,
                                const double* __restrict__ rowscale, const double* __restrict__ rowshift,
                                const double* __restrict__ colshift, TO* __restrict__ out, int64_t ldo) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * K) return;
  const int64_t m = idx / K;
  const int k = (int)(idx - m * K);
  double v = (double)acc0[idx];
  if (acc1) v += rowa[m] * (double)acc1[idx];
  v *= colscale[k];
  if (rowscale) v *= rowscale[m];
  if (colshift) v += (rowshift ? rowshift[m] : 1.0) * colshift[k];
  out[m * ldo + k] = (TO)v;
}

// ------------------------------------------------------------------------------------------------ fp64 reductions
constexpr int kGramRows = 128;     // rows per pass staged in smem
constexpr int kGramMaxBlocks = 592;
// stage 1: block b reduces rows {b, b+grid, ...}*kGramRows into work[b][K*K + K]
__global__ void __launch_bounds__(256) gram_partial_kernel(const double* __restrict__ X, int64_t ldx,
                                                           const double* __restrict__ Y, int64_t ldy, int64_t m, int K,
                                                           double* __restrict__ work) {
  extern __shared__ double sm[];
  double* sx = sm;                     // [kGramRows][K]
  double* sy = Y ? sm + kGramRows * K : sm;
  const int n_out = K * K + K;
  // each thread owns outputs o = tid, tid+256, ... (at most 4 for K <= 31; general loop otherwise)
  const int tid = threadIdx.x;
  constexpr int kMaxOwn = 68;  // supports K <= 128: (128*128+128)/256 = 64.5
  double acc[kMaxOwn];
  const int n_own = (n_out - tid + 255) / 256;
  for (int i = 0; i < kMaxOwn; ++i) acc[i] = 0.0;
  for (int64_t r0 = (int64_t)blockIdx.x * kGramRows; r0 < m; r0 += (int64_t)gridDim.x * kGramRows) {
    const int rows = (int)min((int64_t)kGramRows, m - r0);
    __syncthreads();
    for (int i = tid; i < rows * K; i += 256) {
      const int r = i / K, c = i - r * K;
      sx[i] = X[(r0 + r) * ldx + c];
      if (Y) sy[i] = Y[(r0 + r) * ldy + c];
    }
    __syncthreads();
    for (int i = 0; i < n_own; ++i) {
      const int o = tid + i * 256;
      double a = acc[i];
      if (o < K * K) {
        const int e = o / K, f = o - e * K;
        for (int r = 0; r < rows; ++r) a += sx[r * K + e] * sy[r * K + f];
      } else {
        const int e = o - K * K;
        for (int r = 0; r < rows; ++r) a += sx[r * K + e];
      }
      acc[i] = a;
    }
  }
  for (int i = 0; i < n_own; ++i) work[(size_t)blockIdx.x * n_out + tid + i * 256] = acc[i];
}
__global__ void gram_final_kernel(const double* __restrict__ work, int n_blocks, int K, double* __restrict__ G,
                                  double* __restrict__ colsum) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_out = K * K + K;
  if (o >= n_out) return;
  double a = 0.0;
  for (int b = 0; b < n_blocks; ++b) a += work[(size_t)b * n_out + o];
  if (o < K * K)
    G[o] = a;
  else if (colsum)
    colsum[o - K * K] = a;
}

// the same sum, ADDED into every rank's copy of G through the NVSwitch multicast address of a symmetric K x K buffer
// (row-sharded CholeskyQR2: the K x K all-reduce of the north-star partitioning without a collective call)
__global__ void gram_final_reduce_kernel(const double* __restrict__ work, int n_blocks, int K, double* __restrict__ G_mc) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= K * K) return;
  double a = 0.0;
  for (int b = 0; b < n_blocks; ++b) a += work[(size_t)b * (K * K + K) + o];
  asm volatile("multimem.red.relaxed.sys.global.add.f64 [%0], %1;" ::"l"(G_mc + o), "d"(a) : "memory");
  __threadfence_system();
}

// K x K Cholesky whitening in one block (float64, shared memory): G = R^T R with G scaled to unit diagonal first.
//   r_inv [K,K] and r [K,K] row-major (upper triangular), flag = 1 when G is non-finite, has a non-positive diagonal or
//   a pivot below rank_tol (the caller then takes the rank-revealing host path).  Replaces the K x K step of dask tsqr
//   (lmdec/decomp/iter_methods.py:382) without a device->host round trip.
// `n_partials` > 0: G is not a matrix but the per-block partial Gram matrices of gram_partial_kernel (stride
// K*K + K); they are summed here in block order (deterministic) and the sum is also written to g_out.
__global__ void __launch_bounds__(256) chol_whiten_kernel(const double* __restrict__ G, int K, double rank_tol,
                                                          double* __restrict__ r_inv, double* __restrict__ r,
                                                          int* __restrict__ flag, int n_partials,
                                                          double* __restrict__ g_out) {
  extern __shared__ double sm[];
  double* L = sm;            // [K][K] lower factor of the scaled Gram matrix
  double* Li = sm + K * K;   // [K][K] inverse of L (lower)
  double* d = sm + 2 * K * K;
  __shared__ int bad;
  const int t = threadIdx.x;
  if (t == 0) bad = 0;
  if (n_partials > 0) {
    const int stride = K * K + K;
    for (int i = t; i < K * K; i += blockDim.x) {
      // four loads in flight; the partial sums are always combined in the same order (deterministic)
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      int b = 0;
      for (; b + 4 <= n_partials; b += 4) {
        a0 += G[(size_t)b * stride + i];
        a1 += G[(size_t)(b + 1) * stride + i];
        a2 += G[(size_t)(b + 2) * stride + i];
        a3 += G[(size_t)(b + 3) * stride + i];
      }
      for (; b < n_partials; ++b) a0 += G[(size_t)b * stride + i];
      g_out[i] = (a0 + a1) + (a2 + a3);
    }
    __syncthreads();
    G = g_out;
  }
  __syncthreads();
  if (t < K) {
    const double g = G[t * K + t];
    if (!(g > 0.0) || !isfinite(g)) bad = 1;
    d[t] = (g > 0.0 && isfinite(g)) ? sqrt(g) : 1.0;
  }
  __syncthreads();
  for (int i = t; i < K * K; i += blockDim.x) {
    const int a = i / K, b = i - a * K;
    const double v = 0.5 * (G[a * K + b] + G[b * K + a]) / (d[a] * d[b]);
    if (!isfinite(v)) bad = 1;
    L[i] = v;
    Li[i] = 0.0;
  }
  __syncthreads();
  // right-looking Cholesky, the whole block on every elimination step: the pivot column is scaled, then the trailing
  // (K - j - 1)^2 / 2 elements are updated in parallel (K = 110: 0.33 ms with one thread per row, ~20 us this way)
  const int nt = blockDim.x;
  for (int j = 0; j < K; ++j) {
    const double piv = L[j * K + j];
    const double s = sqrt(piv > rank_tol ? piv : rank_tol);
    __syncthreads();                         // everybody has read the pivot before it is overwritten
    if (t == 0) {
      if (!(piv >= rank_tol)) bad = 1;
      L[j * K + j] = s;
    }
    for (int r = j + 1 + t; r < K; r += nt) L[r * K + j] /= s;
    __syncthreads();
    const int m = K - j - 1;
    for (int idx = t; idx < m * m; idx += nt) {
      const int r = j + 1 + idx / m, c = j + 1 + idx % m;
      if (c <= r) L[r * K + c] -= L[r * K + j] * L[c * K + j];
    }
    __syncthreads();
  }
  // inverse of the lower factor, right-looking: row j of Li is final after its scaling, then it is eliminated from the
  // rows below (rank-1 update of (K - j - 1) x (j + 1) elements in parallel).  Li starts as the identity.
  for (int i = t; i < K; i += nt) Li[i * K + i] = 1.0;
  __syncthreads();
  for (int j = 0; j < K; ++j) {
    const double dj = 1.0 / L[j * K + j];
    for (int c = t; c <= j; c += nt) Li[j * K + c] *= dj;
    __syncthreads();
    const int w = j + 1, m = K - j - 1;
    for (int idx = t; idx < m * w; idx += nt) {
      const int r = j + 1 + idx / w, c = idx % w;
      Li[r * K + c] -= L[r * K + j] * Li[j * K + c];
    }
    __syncthreads();
  }
  // G = D L L^T D  =>  r = L^T D (upper),  r_inv = D^-1 L^-T (upper)
  for (int i = t; i < K * K; i += blockDim.x) {
    const int a = i / K, b = i - a * K;  // element (a, b)
    r[i] = (b >= a) ? L[b * K + a] * d[b] : 0.0;
    r_inv[i] = (b >= a) ? Li[b * K + a] / d[a] : 0.0;
  }
  if (t == 0) *flag = bad;
}


// ---- single-launch CholeskyQR2 for small iterates (the n x K side of the power iteration)
constexpr int kCqrThreads = 512;
#ifndef LMDEC_CQR_CLUSTER
#define LMDEC_CQR_CLUSTER 8
#endif
constexpr int kCqrCluster = LMDEC_CQR_CLUSTER;   // CTAs of one cluster share the row range; Gram partials meet through DSMEM
constexpr int kCqrSliceDoubles = 8192;  // per-CTA row slice kept in shared memory (x2: input and first-pass output)

// Cholesky whitening of the K x K Gram matrix G (shared memory) by the whole block, one matrix element per thread
// (two when K*K > 512) held in a register; each elimination step costs one __syncthreads:
//   G -> unit-diagonal scaling -> right-looking Cholesky -> right-looking triangular inverse
//   -> r = (L diag(d))^T, r_inv = r^-1 (both upper, shared memory).
// Same algorithm and failure flag as chol_whiten_kernel (pivot below rank_tol, non-finite or non-positive diagonal).
// scratch: Lw [K*K] working columns, Lf [K*K] final factor, vec [4*K] (d | 1/L_jj | two row buffers).
__device__ int block_chol(const double* __restrict__ G, int K, double rank_tol, double* __restrict__ Lw,
                          double* __restrict__ Lf, double* __restrict__ vec, double* __restrict__ r_inv,
                          double* __restrict__ r) {
  const int t = threadIdx.x, kk = K * K;
  double* dvec = vec;
  double* invd = vec + K;
  double* xrow = vec + 2 * K;
  int bad = 0;
  if (t < K) {
    const double dg = G[t * K + t];
    const bool ok = dg > 0.0 && isfinite(dg);
    dvec[t] = ok ? sqrt(dg) : 1.0;
    if (!ok) bad = 1;
  }
  __syncthreads();
  int ei[2], ec[2];
  bool act[2];
  double v[2], lf[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int e = t + u * kCqrThreads;
    act[u] = e < kk;
    ei[u] = act[u] ? e / K : 0;
    ec[u] = act[u] ? e - ei[u] * K : 0;
    v[u] = 0.0;
    lf[u] = 0.0;
    if (act[u]) {
      v[u] = 0.5 * (G[ei[u] * K + ec[u]] + G[ec[u] * K + ei[u]]) / (dvec[ei[u]] * dvec[ec[u]]);
      if (!isfinite(v[u])) bad = 1;
      Lw[e] = v[u];
    }
  }
  __syncthreads();
  for (int j = 0; j < K; ++j) {
    const double piv = Lw[j * K + j];
    if (!(piv >= rank_tol)) bad = 1;
    const double pc = piv > rank_tol ? piv : rank_tol;
    const double x = rsqrt(pc);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (!act[u]) continue;
      const int i = ei[u], c = ec[u];
      if (c == j) {
        lf[u] = i > j ? v[u] * x : (i == j ? pc * x : 0.0);
        if (i == j) invd[j] = x;
      } else if (c > j && i >= c) {
        v[u] -= (Lw[i * K + j] * x) * (Lw[c * K + j] * x);
        if (c == j + 1) Lw[i * K + c] = v[u];     // the next step's column
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 2; ++u)
    if (act[u]) Lf[t + u * kCqrThreads] = lf[u];
  __syncthreads();
  // X = L^-1: element (i, c) accumulates delta_ic - sum_{m < i} L[i][m] X[m][c]; row m is final once steps < m ran
  double acc[2], xf[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    acc[u] = (act[u] && ei[u] == ec[u]) ? 1.0 : 0.0;
    xf[u] = 0.0;
  }
  for (int m = 0; m < K; ++m) {
    double* xr = xrow + (m & 1) * K;
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (act[u] && ei[u] == m) {
        xf[u] = acc[u] * invd[m];
        xr[ec[u]] = xf[u];
      }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (act[u] && ei[u] > m) acc[u] -= Lf[ei[u] * K + m] * xr[ec[u]];
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    if (!act[u]) continue;
    const int i = ei[u], c = ec[u];
    r[c * K + i] = (c <= i) ? lf[u] * dvec[i] : 0.0;         // r[a][b] = L[b][a] d[b]
    r_inv[c * K + i] = (c <= i) ? xf[u] / dvec[c] : 0.0;     // r_inv[a][b] = X[b][a] / d[a]
  }
  return __syncthreads_or(bad);
}

// float64 tensor-core step D[8x8] += A[8x4] B[4x8]: lane holds A[lane>>2][lane&3], B[lane&3][lane>>2] and
// D[lane>>2][2*(lane&3) + {0,1}]
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

constexpr int kCqrWarps = kCqrThreads / 32;
constexpr int kCqrPaccDoubles = kCqrWarps * 64;   // one 8x8 partial tile per warp

// Gram matrix of this CTA's row slice (shared memory, [rows][K]) -> part[K*K].  The 8x8 tiles of the upper triangle go
// to warps (tile, row split); each warp runs the slice through DMMA (k = 4 rows per step, two accumulator pairs in
// flight) and the splits are folded in a fixed order.
__device__ void cqr_gram_local(const double* __restrict__ sl, int rows, int K, double* __restrict__ pacc,
                               double* __restrict__ part) {
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int nt = (K + 7) >> 3, ntile = nt * (nt + 1) / 2;
  const int nsplit = kCqrWarps / ntile;            // >= 1 (nt <= 4 -> ntile <= 10)
  const int split = warp / ntile;
  if (split < nsplit) {
    int ta = 0, rem = warp - split * ntile;
    while (rem >= nt - ta) { rem -= nt - ta; ++ta; }
    const int tb = ta + rem;
    const int per = ((rows + nsplit - 1) / nsplit + 3) & ~3;
    const int lo = split * per, hi = min(rows, lo + per);
    const int j = lane & 3, i = lane >> 2;
    const int ca = ta * 8 + i, cb = tb * 8 + i;
    const bool oka = ca < K, okb = cb < K;
    double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
    int r = lo + j;
    for (; r - j + 4 < hi; r += 8) {
      const double a0 = oka ? sl[r * K + ca] : 0.0, b0 = okb ? sl[r * K + cb] : 0.0;
      const bool in1 = r + 4 < hi;
      const double a1 = (oka && in1) ? sl[(r + 4) * K + ca] : 0.0, b1 = (okb && in1) ? sl[(r + 4) * K + cb] : 0.0;
      dmma884(c0, c1, a0, b0);
      dmma884(d0, d1, a1, b1);
    }
    if (r - j < hi) {
      const bool in0 = r < hi;
      dmma884(c0, c1, (oka && in0) ? sl[r * K + ca] : 0.0, (okb && in0) ? sl[r * K + cb] : 0.0);
    }
    double* o = pacc + (size_t)warp * 64 + i * 8 + 2 * j;
    o[0] = c0 + d0;
    o[1] = c1 + d1;
  }
  __syncthreads();
  for (int e = t; e < K * K; e += kCqrThreads) {
    int a = e / K, b = e - a * K;
    if (a > b) { const int w = a; a = b; b = w; }
    const int ta = a >> 3, tb = b >> 3;
    const int idx = ta * nt - ta * (ta - 1) / 2 + (tb - ta);
    double acc = 0.0;
    for (int s_ = 0; s_ < nsplit; ++s_) acc += pacc[(size_t)(s_ * ntile + idx) * 64 + (a & 7) * 8 + (b & 7)];
    part[e] = acc;
  }
}

// out[r][b] = sum_{a <= b} in[r][a] r_inv[a][b] by DMMA: warp -> 8-row tiles, k = 4 columns of `in` per step; r_inv is
// upper triangular, so steps beyond the column tile are skipped.
__device__ void cqr_apply(const double* __restrict__ in, int rows, int K, const double* __restrict__ r_inv,
                          double* __restrict__ out, int64_t ldo) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = lane & 3, i = lane >> 2;
  const int nt = (K + 7) >> 3, mt = (rows + 7) >> 3;
  for (int tb = 0; tb < nt; ++tb) {
    const int n0 = tb * 8;
    const int ks = min((K + 3) >> 2, 2 * tb + 2);       // 4s <= n0 + 7
    double bf[8];
#pragma unroll
    for (int s_ = 0; s_ < 8; ++s_) {
      const int a = 4 * s_ + j, b = n0 + i;
      bf[s_] = (s_ < ks && a < K && b < K) ? r_inv[a * K + b] : 0.0;
    }
    for (int tm = warp; tm < mt; tm += kCqrWarps) {
      const int r = tm * 8 + i;
      const bool okr = r < rows;
      const double* row = in + r * K;
      double c0 = 0.0, c1 = 0.0;
#pragma unroll
      for (int s_ = 0; s_ < 8; ++s_) {
        if (s_ < ks) {
          const int a = 4 * s_ + j;
          dmma884(c0, c1, (okr && a < K) ? row[a] : 0.0, bf[s_]);
        }
      }
      const int b = n0 + 2 * j;
      if (okr) {
        if (b < K) out[(int64_t)r * ldo + b] = c0;
        if (b + 1 < K) out[(int64_t)r * ldo + b + 1] = c1;
      }
    }
  }
}

// CholeskyQR2 of a small iterate X [n][K] in ONE launch of one 8-CTA cluster: each CTA keeps its row slice in shared
// memory for both passes, the K x K Gram partials are summed through distributed shared memory in a fixed order
// (every CTA gets the same bits), every CTA factorises redundantly and applies r^-1 to its own rows.
// packed = [r2 r1 | flag1 | flag2 | max|Q1^T Q1 - I|] as qr_pack_kernel writes it.
__global__ void __cluster_dims__(kCqrCluster, 1, 1) __launch_bounds__(kCqrThreads)
    cholqr2_small_kernel(const double* __restrict__ X, int64_t ldx, int64_t n, int K, int per, double rank_tol,
                         double* __restrict__ Q, int64_t ldq, double* __restrict__ packed,
                         double* __restrict__ stats_work, double stats_lim) {
  extern __shared__ double cqr_sm[];
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int t = threadIdx.x;
  const int kk = K * K;
  double* buf0 = cqr_sm;                  // [per][K] input slice
  double* buf1 = buf0 + (size_t)per * K;  // [per][K] first-pass output
  double* part1 = buf1 + (size_t)per * K; // Gram partial of pass 1 (read by the peers)
  double* part2 = part1 + kk;
  double* pacc = part2 + kk;              // [warps][8][8] partial Gram tiles
  double* G = pacc + kCqrPaccDoubles;
  double* Lsm = G + kk;
  double* Lf = Lsm + kk;
  double* vec = Lf + kk;                  // [4][K]
  double* ri = vec + 4 * K;
  double* r1 = ri + kk;
  double* r2 = r1 + kk;
  __shared__ int flags[2];
  __shared__ double red[kCqrThreads / 32];
  const int64_t row0 = (int64_t)rank * per;
  const int rows = (int)max((int64_t)0, min((int64_t)per, n - row0));
  if (t == 0) LMDEC_STAMP(7, 0);
  for (int i = t; i < rows * K; i += kCqrThreads) {
    const int r = i / K, c = i - r * K;
    buf0[i] = X[(row0 + r) * ldx + c];
  }
  __syncthreads();
  if (t == 0) LMDEC_STAMP(7, 1);
  for (int pass = 0; pass < 2; ++pass) {
    double* part = pass ? part2 : part1;
    cqr_gram_local(pass ? buf1 : buf0, rows, K, pacc, part);
    if (t == 0) LMDEC_STAMP(7, 2 + 8 * pass);
    cluster.sync();
    if (t == 0) LMDEC_STAMP(7, 3 + 8 * pass);
    for (int i = t; i < kk; i += kCqrThreads) {
      double acc = 0.0;
      for (int c = 0; c < kCqrCluster; ++c) acc += cluster.map_shared_rank(part, c)[i];
      G[i] = acc;
    }
    __syncthreads();
    if (t == 0) LMDEC_STAMP(7, 4 + 8 * pass);
    if (pass == 1) {   // deviation of the second Gram matrix from the identity (health check of the first pass)
      double dev = 0.0;
      for (int i = t; i < kk; i += kCqrThreads) {
        const double g = G[i] - ((i / K) == (i % K) ? 1.0 : 0.0);
        dev = fmax(dev, g == g ? fabs(g) : INFINITY);
      }
      for (int o = 16; o > 0; o >>= 1) dev = fmax(dev, __shfl_xor_sync(0xffffffffu, dev, o));
      if ((t & 31) == 0) red[t >> 5] = dev;
    }
    const int f = block_chol(G, K, rank_tol, Lsm, Lf, vec, ri, pass ? r2 : r1);
    if (t == 0) flags[pass] = f;
    if (t == 0) LMDEC_STAMP(7, 5 + 8 * pass);
    if (pass == 0) cqr_apply(buf0, rows, K, ri, buf1, K);
    else cqr_apply(buf1, rows, K, ri, Q + row0 * ldq, ldq);
    __syncthreads();
    if (t == 0) LMDEC_STAMP(7, 6 + 8 * pass);
  }
  if (stats_work) {
    // column statistics of Q as the operand of the A^T q that follows (what colstats_partial_kernel computes for it:
    // max |q|, -, sum q per column), one partial row per CTA in the operator's work buffer
    const int n_sl = kCqrThreads / K;
    const int k = t % K, sl = t / K;
    if (sl < n_sl) {
      double m1 = 0.0, ws = 0.0;
      for (int r = sl; r < rows; r += n_sl) {
        const double y = Q[(row0 + r) * ldq + k];
        double v = fabs(y);
        if (v != v) v = INFINITY;
        m1 = fmax(m1, v);
        ws += y;
      }
      pacc[sl * K + k] = m1;
      pacc[kCqrThreads + sl * K + k] = ws;
    }
    __syncthreads();
    if (t < K) {
      double m1 = 0.0, ws = 0.0;
      for (int j = 0; j < n_sl; ++j) {
        m1 = fmax(m1, pacc[j * K + t]);
        ws += pacc[kCqrThreads + j * K + t];
      }
      double* out = stats_work + 5 * K + (size_t)rank * 2 * K;
      out[t] = m1;
      out[K + t] = ws;
    }
  }
  if (rank == 0) {
    for (int i = t; i < kk; i += kCqrThreads) {
      const int a = i / K, b = i - a * K;
      double acc = 0.0;
      for (int c = a; c <= b; ++c) acc += r2[a * K + c] * r1[c * K + b];
      packed[i] = acc;
    }
    if (t == 0) {
      double m = 0.0;
      for (int i = 0; i < kCqrThreads / 32; ++i) m = fmax(m, red[i]);
      packed[kk] = (double)flags[0];
      packed[kk + 1] = (double)flags[1];
      packed[kk + 2] = m;
    }
  }
  cluster.sync();   // peers may still be reading this CTA's part2
  if (stats_work && rank == 0 && t < K) {
    // fold the CTAs' rows into the final statistics (what colstats_finish_kernel writes)
    double m1 = 0.0, ws = 0.0;
    for (int c = 0; c < kCqrCluster; ++c) {
      const double* in = stats_work + 5 * K + (size_t)c * 2 * K;
      m1 = fmax(m1, __ldcg(in + t));
      ws += __ldcg(in + K + t);
    }
    const double mult = m1 > 0.0 ? stats_lim / m1 : 1.0;
    stats_work[t] = mult;
    stats_work[K + t] = 1.0 / mult;
    stats_work[2 * K + t] = -ws;
    stats_work[3 * K + t] = m1;
    stats_work[4 * K + t] = 0.0;
  }
  if (t == 0) LMDEC_STAMP(7, 16);
}

// in-place capable variant of cqr_apply: every warp reads its 8-row tile completely (all column tiles) before it
// writes it, so `out` may alias `in` (ldo == K)
__device__ void cqr_apply_rows(const double* in, int rows, int K, const double* __restrict__ r_inv, double* out,
                               int64_t ldo) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = lane & 3, i = lane >> 2;
  const int nt = (K + 7) >> 3, mt = (rows + 7) >> 3;
  for (int tm = warp; tm < mt; tm += kCqrWarps) {
    const int r = tm * 8 + i;
    const bool okr = r < rows;
    const double* row = in + (size_t)r * K;
    double c[4][2];
#pragma unroll
    for (int tb = 0; tb < 4; ++tb) {
      c[tb][0] = c[tb][1] = 0.0;
      if (tb < nt) {
        const int b = tb * 8 + i;
        const int ks = min((K + 3) >> 2, 2 * tb + 2);       // r_inv is upper triangular: rows a <= 8 tb + 7 only
#pragma unroll
        for (int s_ = 0; s_ < 8; ++s_) {
          if (s_ < ks) {
            const int a = 4 * s_ + j;
            dmma884(c[tb][0], c[tb][1], (okr && a < K) ? row[a] : 0.0, (a < K && b < K) ? r_inv[a * K + b] : 0.0);
          }
        }
      }
    }
    __syncwarp();
#pragma unroll
    for (int tb = 0; tb < 4; ++tb) {
      const int b = tb * 8 + 2 * j;
      if (tb < nt && okr) {
        if (b < K) out[(int64_t)r * ldo + b] = c[tb][0];
        if (b + 1 < K) out[(int64_t)r * ldo + b + 1] = c[tb][1];
      }
    }
    __syncwarp();
  }
}

constexpr int kGridSliceDoubles = 16384;   // per-CTA row slice (one buffer, the first pass is applied in place)

// CholeskyQR2 of a mid-size iterate (K <= 32, n*K <= SMs * 16384) in ONE cooperative launch over all SMs: the row slice of
// a CTA stays in shared memory for both passes; per pass the K x K Gram partials go through global memory -- element e is
// summed over the CTAs (fixed order) by CTA e % grid, then broadcast -- with two grid syncs.  Same algorithm, outputs and
// failure flags as cholqr2_small_kernel.  gwork: [grid][K*K] partials | [2][K*K] sums | [grid][2K] statistics rows.
__global__ void __launch_bounds__(kCqrThreads)
    cholqr2_grid_kernel(const double* __restrict__ X, int64_t ldx, int64_t n, int K, int per, double rank_tol,
                        double* __restrict__ Q, int64_t ldq, double* __restrict__ packed, double* __restrict__ gwork,
                        double* __restrict__ stats_work, double stats_lim) {
  extern __shared__ double cqr_sm[];
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  const int rank = blockIdx.x, G_ = gridDim.x;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int kk = K * K;
  double* buf = cqr_sm;                     // [per][K] row slice: X, then Q1 (in place)
  double* part = buf + (size_t)per * K;
  double* pacc = part + kk;                 // [warps][8][8] partial Gram tiles
  double* Gm = pacc + kCqrPaccDoubles;
  double* Lsm = Gm + kk;
  double* Lf = Lsm + kk;
  double* vec = Lf + kk;                    // [4][K]
  double* ri = vec + 4 * K;
  double* r1 = ri + kk;
  double* r2 = r1 + kk;
  double* gpart = gwork;
  double* gsum = gwork + (size_t)G_ * kk;
  double* gstat = gsum + 2 * kk;
  __shared__ int flags[2];
  __shared__ double red[kCqrThreads / 32];
  const int64_t row0 = (int64_t)rank * per;
  const int rows = (int)max((int64_t)0, min((int64_t)per, n - row0));
  for (int i = t; i < rows * K; i += kCqrThreads) {
    const int r = i / K, c = i - r * K;
    buf[i] = X[(row0 + r) * ldx + c];
  }
  __syncthreads();
  for (int pass = 0; pass < 2; ++pass) {
    cqr_gram_local(buf, rows, K, pacc, part);
    __syncthreads();
    for (int i = t; i < kk; i += kCqrThreads) gpart[(size_t)rank * kk + i] = part[i];
    grid.sync();
    // element e of the Gram matrix: warp (e / grid) of CTA e % grid sums the grid's partials, lane-strided then butterfly
    for (int e = rank + warp * G_; e < kk; e += kCqrWarps * G_) {
      double acc = 0.0;
      for (int b = lane; b < G_; b += 32) acc += __ldcg(gpart + (size_t)b * kk + e);
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) gsum[pass * kk + e] = acc;
    }
    grid.sync();
    for (int i = t; i < kk; i += kCqrThreads) Gm[i] = __ldcg(gsum + pass * kk + i);
    __syncthreads();
    if (pass == 1) {   // deviation of the second Gram matrix from the identity (health check of the first pass)
      double dev = 0.0;
      for (int i = t; i < kk; i += kCqrThreads) {
        const double g = Gm[i] - ((i / K) == (i % K) ? 1.0 : 0.0);
        dev = fmax(dev, g == g ? fabs(g) : INFINITY);
      }
      for (int o = 16; o > 0; o >>= 1) dev = fmax(dev, __shfl_xor_sync(0xffffffffu, dev, o));
      if (lane == 0) red[warp] = dev;
    }
    const int f = block_chol(Gm, K, rank_tol, Lsm, Lf, vec, ri, pass ? r2 : r1);
    if (t == 0) flags[pass] = f;
    if (pass == 0) cqr_apply_rows(buf, rows, K, ri, buf, K);
    else cqr_apply_rows(buf, rows, K, ri, Q + row0 * ldq, ldq);
    __syncthreads();
  }
  if (stats_work) {
    // column statistics of Q for the A^T q that follows (see cholqr2_small_kernel): one row per CTA, CTA 0 folds them
    const int n_sl = kCqrThreads / K;
    const int k = t % K, sl = t / K;
    if (sl < n_sl) {
      double m1 = 0.0, ws = 0.0;
      for (int r = sl; r < rows; r += n_sl) {
        const double y = Q[(row0 + r) * ldq + k];
        double v = fabs(y);
        if (v != v) v = INFINITY;
        m1 = fmax(m1, v);
        ws += y;
      }
      pacc[sl * K + k] = m1;
      pacc[kCqrThreads + sl * K + k] = ws;
    }
    __syncthreads();
    if (t < K) {
      double m1 = 0.0, ws = 0.0;
      for (int j = 0; j < n_sl; ++j) {
        m1 = fmax(m1, pacc[j * K + t]);
        ws += pacc[kCqrThreads + j * K + t];
      }
      gstat[(size_t)rank * 2 * K + t] = m1;
      gstat[(size_t)rank * 2 * K + K + t] = ws;
    }
    grid.sync();
    if (rank == 0) {
      const int k2 = t % K, sl2 = t / K;
      if (sl2 < n_sl) {
        double m1 = 0.0, ws = 0.0;
        for (int b = sl2; b < G_; b += n_sl) {
          m1 = fmax(m1, __ldcg(gstat + (size_t)b * 2 * K + k2));
          ws += __ldcg(gstat + (size_t)b * 2 * K + K + k2);
        }
        pacc[sl2 * K + k2] = m1;
        pacc[kCqrThreads + sl2 * K + k2] = ws;
      }
      __syncthreads();
      if (t < K) {
        double m1 = 0.0, ws = 0.0;
        for (int j = 0; j < n_sl; ++j) {
          m1 = fmax(m1, pacc[j * K + t]);
          ws += pacc[kCqrThreads + j * K + t];
        }
        const double mult = m1 > 0.0 ? stats_lim / m1 : 1.0;
        stats_work[t] = mult;
        stats_work[K + t] = 1.0 / mult;
        stats_work[2 * K + t] = -ws;
        stats_work[3 * K + t] = m1;
        stats_work[4 * K + t] = 0.0;
      }
    }
  }
  if (rank == 0) {
    for (int i = t; i < kk; i += kCqrThreads) {
      const int a = i / K, b = i - a * K;
      double acc = 0.0;
      for (int c = a; c <= b; ++c) acc += r2[a * K + c] * r1[c * K + b];
      packed[i] = acc;
    }
    if (t == 0) {
      double m = 0.0;
      for (int i = 0; i < kCqrThreads / 32; ++i) m = fmax(m, red[i]);
      packed[kk] = (double)flags[0];
      packed[kk + 1] = (double)flags[1];
      packed[kk + 2] = m;
    }
  }
}

__global__ void qvals_kernel(const double* __restrict__ si, const double* __restrict__ sj, int k, int scale,
                             double* __restrict__ out) {
  double d = 0, a = 0, b = 0;
  for (int i = threadIdx.x; i < k; i += 32) {
    const double x = si[i], y = sj[i];
    d += (x - y) * (x - y);
    a += x * x;
    b += y * y;
  }
  for (int o = 16; o > 0; o >>= 1) {
    d += __shfl_xor_sync(0xffffffffu, d, o);
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if (threadIdx.x == 0) {
    double acc = sqrt(d);
    if (scale) acc /= (sqrt(a) + sqrt(b)) / 2;
    out[0] = acc;
  }
}

__global__ void __launch_bounds__(256) sqdiff_partial_kernel(const double* __restrict__ A, int64_t lda,
                                                             const double* __restrict__ B, int64_t ldb,
                                                             const double* __restrict__ s, int64_t m, int K,
                                                             double* __restrict__ work) {
  __shared__ double red[8];
  double a = 0.0;
  const int64_t total = m * K;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int64_t r = i / K;
    const int k = (int)(i - r * K);
    const double d = A[r * lda + k] - B[r * ldb + k] * (s ? s[k] : 1.0);
    a += d * d;
  }
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int i = 0; i < 8; ++i) t += red[i];
    work[blockIdx.x] = t;
  }
}
__global__ void sum_final_kernel(const double* __restrict__ work, int n, double* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double t = 0;
    for (int i = 0; i < n; ++i) t += work[i];
    out[0] = t;
  }
}

}  // namespace lmdec

// ================================================================================================ C ABI
using namespace lmdec;

extern "C" {

const char* lmdec_last_error(void) { return g_err; }
int lmdec_abi_version(void) { return 3; }   // 3: + lmdec_dense_*, lmdec_operator_apply_reduce, lmdec_gram_reduce
#ifndef LMDEC_SRC_HASH
#define LMDEC_SRC_HASH "unknown"
#endif
static const char kSrcHashMarker[] = "LMDEC_SRC_HASH=" LMDEC_SRC_HASH;   // also found by scanning the file (_build.py)
const char* lmdec_source_hash(void) { return kSrcHashMarker + 15; }
long long lmdec_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int64_t lmdec_store_bytes(int64_t M, int64_t Kd) { return cdiv(M, 128) * cdiv(Kd, 256) * 8192; }
int64_t lmdec_image_bytes(int64_t Kd, int N) { return cdiv(Kd, 256) * 8 * (int64_t)N * 32; }

int lmdec_pack_2bit(const void* src, int dtype, int64_t stride_m, int64_t stride_kd, int64_t block_m, int64_t block_kd,
                    int64_t m0, int64_t kd0, uint8_t* store, int64_t M, int64_t Kd, int32_t* bad_count, void* stream) {
  if (m0 % 128 || kd0 % 256) return fail("lmdec_pack_2bit: m0 must be a multiple of 128 and kd0 of 256");
  if (m0 + block_m > M || kd0 + block_kd > Kd || block_m <= 0 || block_kd <= 0)
    return fail("lmdec_pack_2bit: block outside the store");
  const int64_t n_kc = cdiv(Kd, 256);
  const int64_t mp0 = m0 / 128, kc0 = kd0 / 256;
  const int64_t n_mp_blk = cdiv(m0 + block_m, 128) - mp0, n_kc_blk = cdiv(kd0 + block_kd, 256) - kc0;
  const int64_t n_units = n_mp_blk * n_kc_blk * 512;
  const int threads = 256;
  const unsigned blocks = (unsigned)cdiv(n_units, threads);
  cudaStream_t st = (cudaStream_t)stream;
#define LMDEC_PACK(T)                                                                                              \
  pack_kernel<T><<<blocks, threads, 0, st>>>((const T*)src, stride_m, stride_kd, block_m, block_kd, m0, kd0, store, \
                                             n_kc, mp0, kc0, n_mp_blk, n_kc_blk, bad_count)
  switch (dtype) {
    case LMDEC_I8: LMDEC_PACK(int8_t); break;
    case LMDEC_U8: LMDEC_PACK(uint8_t); break;
    case LMDEC_F32: LMDEC_PACK(float); break;
    case LMDEC_F64: LMDEC_PACK(double); break;
    default: return fail("lmdec_pack_2bit: unknown dtype");
  }
#undef LMDEC_PACK
  return check_launch("lmdec_pack_2bit");
}

int lmdec_unpack_2bit(const uint8_t* store, int64_t M, int64_t Kd, int8_t* dst, int64_t ld, void* stream) {
  const int64_t n_kc = cdiv(Kd, 256), n_units = cdiv(M, 128) * n_kc * 512;
  unpack_kernel<<<(unsigned)cdiv(n_units, 256), 256, 0, (cudaStream_t)stream>>>(store, M, Kd, n_kc, dst, ld, n_units);
  return check_launch("lmdec_unpack_2bit");
}

int lmdec_recode_bed(const uint8_t* bed, int64_t n_variants, int64_t n_samples, uint8_t* store, void* stream) {
  const int64_t n_kc = cdiv(n_samples, 256), n_units = cdiv(n_variants, 128) * n_kc * 512;
  recode_bed_kernel<<<(unsigned)cdiv(n_units, 256), 256, 0, (cudaStream_t)stream>>>(
      bed, n_variants, n_samples, cdiv(n_samples, 4), store, n_kc, n_units);
  return check_launch("lmdec_recode_bed");
}

int lmdec_store_transpose(const uint8_t* src, int64_t M, int64_t Kd, uint8_t* dst, void* stream) {
  const int64_t src_kc = cdiv(Kd, 256), dst_kc = cdiv(M, 256), n_units = cdiv(Kd, 128) * dst_kc * 512;
  store_transpose_kernel<<<(unsigned)cdiv(n_units, 256), 256, 0, (cudaStream_t)stream>>>(src, M, Kd, src_kc, dst, dst_kc,
                                                                                        n_units);
  return check_launch("lmdec_store_transpose");
}

int lmdec_code_counts(const uint8_t* store, int64_t M, int64_t Kd, int64_t kd_limit, int64_t* counts, void* stream) {
  if (kd_limit < 0 || kd_limit > Kd) return fail("lmdec_code_counts: kd_limit outside the store");
  code_counts_kernel<<<(unsigned)cdiv(M, 128), 128, 0, (cudaStream_t)stream>>>(store, M, cdiv(Kd, 256), kd_limit,
                                                                              (long long*)counts);
  return check_launch("lmdec_code_counts");
}

}  // extern "C"

template <typename TB>
static int quantize_planes_t(const TB* B, int64_t ldb, int64_t Kd, int64_t kd_limit, int K, int P, int N,
                             const double* mult, const double* row1, const double* row2, int8_t* image1,
                             int8_t* image2, void* stream) {
  if (P < 1 || P > 4 || K < 1 || N % 16 || N < 16 || N > 128 || K * P > N)
    return fail("lmdec_quantize_planes: need 1<=P<=4, N%16==0, 16<=N<=128, K*P<=N");
  if (image2 && !row2) return fail("lmdec_quantize_planes: image2 needs row2");
  if (kd_limit > Kd) return fail("lmdec_quantize_planes: kd_limit > Kd");
  const int64_t n_groups = cdiv(Kd, 256) * 8 * 2;
  const int kq = N / P;
  const int gpb = 256 / kq;          // groups per block (kq <= 128)
  const unsigned blocks = (unsigned)cdiv(n_groups, gpb);
  cudaStream_t st_ = (cudaStream_t)stream;
#define LMDEC_Q(PP)                                                                                              \
  quantize_planes_kernel<PP, TB><<<blocks, 256, 0, st_>>>(B, ldb, kd_limit, K, N, mult, row1, row2, image1, image2, \
                                                          n_groups, kq, gpb)
  switch (P) {
    case 1: LMDEC_Q(1); break;
    case 2: LMDEC_Q(2); break;
    case 3: LMDEC_Q(3); break;
    default: LMDEC_Q(4); break;
  }
#undef LMDEC_Q
  return check_launch("lmdec_quantize_planes");
}

extern "C" {

int lmdec_quantize_planes(const double* B, int64_t ldb, int64_t Kd, int64_t kd_limit, int K, int P, int N,
                          const double* mult, const double* row1, const double* row2, int8_t* image1, int8_t* image2,
                          void* stream) {
  return quantize_planes_t<double>(B, ldb, Kd, kd_limit, K, P, N, mult, row1, row2, image1, image2, stream);
}

int lmdec_colabsmax(const double* B, int64_t ldb, int64_t kd_limit, int K, const double* row1, const double* row2,
                    double* out, double* out2, void* stream) {
  if (K < 1 || K > 256) return fail("lmdec_colabsmax: K out of range");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(out, 0, K * sizeof(double), st);
  if (row2) cudaMemsetAsync(out2, 0, K * sizeof(double), st);
  const int rpb = 256 / K > 0 ? 256 / K : 1;
  const int threads = K > 256 ? K : 256;
  int64_t blocks = cdiv(kd_limit, rpb);
  if (blocks > 1184) blocks = 1184;
  if (blocks < 1) blocks = 1;
  colabsmax_kernel<<<(unsigned)blocks, threads, 0, st>>>(B, ldb, kd_limit, K, row1, row2, (unsigned long long*)out,
                                                         (unsigned long long*)out2);
  return check_launch("lmdec_colabsmax");
}

}  // extern "C"

// fz (optional): finalize fused into the epilogue when the product is not split; *fused reports whether it was
static int launch_geno_mma(lmdec_ctx* ctx_in, const uint8_t* store, int64_t M, int64_t Kd, int64_t m_limit, int64_t kd_limit,
                           int has_missing, int mode, const int8_t* image1, const int8_t* image2, int K, int P, int N,
                           int64_t* out0, int64_t* out1, int splits, void* stream, const FusedFinalize* fz,
                           bool* fused, bool* stats_emitted = nullptr) {
  if (N % 16 || N < 16 || N > 128 || K * P > N || P < 1 || P > 4) return fail("lmdec_geno_matmul: bad K/P/N");
  if (m_limit <= 0 || m_limit > M || kd_limit <= 0 || kd_limit > Kd) return fail("lmdec_geno_matmul: bad limits");
  if (mode == 1 && (!has_missing || !out1)) return fail("lmdec_geno_matmul: mode 1 needs has_missing and out1");
  if (!image1 || !out0) return fail("lmdec_geno_matmul: null operand");
  const lmdec_ctx& ctx = ctx_in ? *ctx_in : kDefaultCtx;
  static thread_local bool attr_set = false;
  int n_sm = ctx.n_sm;
  if (!n_sm) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (ctx_in) ctx_in->n_sm = n_sm;
  }
  GenoMmaParams p;
  memset(&p, 0, sizeof(p));
  p.store = store;
  p.store_kc = cdiv(Kd, 256);
  p.m_limit = m_limit;
  p.m_tiles = (int)cdiv(m_limit, 128);
  p.n_chunks = (int)cdiv(kd_limit, 256);
  p.has_missing = has_missing ? 1 : 0;
  p.mode = mode;
  p.n_img = (has_missing && mode == 0 && image2) ? 2 : 1;
  p.image1 = image1;
  p.image2 = image2;
  p.K = K;
  p.P = P;
  p.N = N;
  p.out0 = (long long*)out0;
  p.out1 = (long long*)out1;
  // pair mode: two m-tiles share every staged chunk of the operand image (halves the dominant L2 -> SM traffic).  Only
  // for long contractions (the accumulators are not double-buffered across work items) and when TMEM has room.
  p.pair = kPairNone;
  if (p.m_tiles >= 2 && p.n_chunks >= ctx.pair_min_chunks && ctx.pair_mode) {
    if (has_missing && N <= 64) p.pair = kPairM;
    else if (!has_missing && N <= 64) p.pair = kPairE;
    else if (!has_missing) p.pair = kPairT;
  }
  p.m_units = p.pair ? (p.m_tiles + 1) / 2 : p.m_tiles;
  p.issuer_groups = ctx.issuer_groups;
  // TMEM budget (512 columns): accumulators first, the rest is the decoded-A ring.  Short contractions alternate two
  // accumulator sets so the epilogue overlaps the next tile; long ones (epilogue negligible) give the columns to a
  // deeper A ring instead.  Without missing data (N <= 64) the round-robin issuer schedule keeps ONE accumulator per
  // tile (geno_mma.cuh, issuer_loop_rr).
  p.rr = (!has_missing && N <= 64 && ctx.rr && kIssuerWarps == 4) ? 1 : 0;
  p.one_acc = (p.pair == kPairT || p.rr) ? 1 : 0;
  {
    const int acc_w = N <= 64 ? 64 : 128;
    int acc_cols;
    if (p.pair) {
      p.acc_sets = 1;
      acc_cols = (p.one_acc ? 2 : 4) * acc_w;
    } else {
      p.acc_sets = (N <= 64 && p.n_chunks < ctx.pair_min_chunks) ? 2 : 1;
      if (ctx.acc_sets == 1 || (ctx.acc_sets == 2 && N <= 64)) p.acc_sets = ctx.acc_sets;
      acc_cols = p.acc_sets * (p.one_acc ? 1 : 2) * acc_w;
    }
    p.a_ring_col0 = acc_cols;
    p.a_stages = (512 - acc_cols) / (has_missing ? 128 : 64);
    if (p.a_stages > kMaxAStages) p.a_stages = kMaxAStages;
  }
  // split-K: minimise the makespan  ceil(work / SMs) * (chunks per split + per-item overhead)  over the split count;
  // the int32 accumulators cap a split at 256 chunks (value bytes are 64*code: 64 * 3 * 127 * 65536 < 2^31)
  int want = splits;
  const int min_splits = (int)cdiv(p.n_chunks, 256);
  if (want <= 0) {
    const int max_s = p.n_chunks / 4 > 1 ? (p.n_chunks / 4 < 256 ? p.n_chunks / 4 : 256) : 1;
    long long best = -1;
    for (int s_ = min_splits > 1 ? min_splits : 1; s_ <= max_s; ++s_) {
      const long long waves = cdiv((long long)p.m_units * s_, n_sm - ctx.sm_reserve > 8 ? n_sm - ctx.sm_reserve : 8);
      const long long cost = waves * ((p.pair ? 2 : 1) * cdiv(p.n_chunks, s_) + 6) + (s_ > 1 ? 2 : 0);
      if (best < 0 || cost < best) {
        best = cost;
        want = s_;
      }
    }
    if (want <= 0) want = 1;
  }
  if (want < min_splits) want = min_splits;
  if (want > p.n_chunks) want = p.n_chunks;
  const int cps = (int)cdiv(p.n_chunks, want);
  p.splits = (int)cdiv(p.n_chunks, cps);
  p.use_atomics = p.splits > 1;
  if (fused) *fused = false;
  if (fz && !p.use_atomics) {
    p.fz = *fz;
    p.fuse = 1;
    if (fz->stats_partial) {
      if (K <= 32 && fz->stats_final) p.fuse = 2;   // epilogue also emits the output's column statistics
      else p.fz.stats_partial = nullptr;
    }
    if (fused) *fused = true;
  }
  if (stats_emitted) *stats_emitted = p.fuse == 2;
  if (fz && fz->mc_out && !p.fuse) return fail("lmdec_operator_apply_reduce: the contraction is split, nothing to fuse into") + 2;
  const size_t stage_bytes = (size_t)8 * N * 32 * p.n_img;
  const size_t budget = 223 * 1024 - 512;  // 227 KB per CTA minus the static barriers / column constants (4 KB), slack
  // shared memory behind the rings: the float32 epilogue's staging area (two output tiles, per-row factors, fold
  // scratch: epi32_stage_bytes), or -- float64 output with fused statistics -- the per-thread accumulators
  // [column slot][2][256] doubles of the round-1 epilogue
  size_t scratch = 0;
  p.epi_staged = 0;
  if (p.fuse && p.fz.out_f32 && LMDEC_EPI_F32 == 2) {
    scratch = (epi32_stage_bytes(K) + 15) / 16 * 16;
    if (scratch + (size_t)2 * kTileBytes + 2 * stage_bytes <= budget) p.epi_staged = 1;
    else scratch = 0;
  }
  if (p.fuse && p.fz.mc_out && !p.epi_staged)
    return fail("lmdec_operator_apply_reduce: no room for the staged float32 epilogue") + 2;
  if (p.fuse == 2 && !p.epi_staged) {
    const int split = kEpilogueWarps / 4;
    int cols = 0;
    for (int g = 0; g * 8 < K; ++g) {
      const int hi = (g / split) * 8 + (K - 8 * g < 8 ? K - 8 * g : 8);
      if (hi > cols) cols = hi;
    }
    p.stats_cols = cols;
    scratch = (size_t)cols * 2 * kEpilogueWarps * 32 * sizeof(double);
    const size_t fold = (size_t)3 * kEpilogueWarps * 32 * sizeof(double);   // the last CTA's fold reuses it: [slices][3][K]
    if (scratch < fold) scratch = fold;
    if (scratch + (size_t)kMaxTileStages * kTileBytes + 2 * stage_bytes > budget) {   // no room: stand-alone pass
      scratch = 0;
      p.fuse = 1;
      p.fz.stats_partial = nullptr;
      if (stats_emitted) *stats_emitted = false;
    }
  }
  int tile_stages = kMaxTileStages, b_stages = 0;
  // Short, unsplit contractions (A^T q with a few thousand samples): the WHOLE operand image stays in shared memory
  // for the life of the CTA.  Every m-tile otherwise re-reads it from L2 -- 16 KB of image per 8 KB of genotypes, i.e.
  // two thirds of the kernel's L2 -> SM traffic, which at ~42 B/cycle/SM is what bounds it (profiles/r02).
  p.b_resident = 0;
  const size_t image_bytes = (size_t)p.n_chunks * stage_bytes;
  if (ctx.b_resident && !p.rr && !p.pair && p.splits == 1 && image_bytes + scratch + 4 * (size_t)kTileBytes <= budget) {
    p.b_resident = 1;
    tile_stages = (int)((budget - scratch - image_bytes) / kTileBytes);
    if (tile_stages > kMaxTileStages) tile_stages = kMaxTileStages;
    b_stages = p.n_chunks;      // (only sizes the region; the resident image has no ring)
  } else {
    // the operand ring keeps its full depth first; tiles beyond 8 in flight only from what is left
    for (; tile_stages > 8; tile_stages -= 2) {
      b_stages = (int)((budget - scratch - (size_t)tile_stages * kTileBytes) / stage_bytes);
      if (b_stages >= kMaxBStages) break;
    }
    for (; tile_stages >= 2; tile_stages /= 2) {
      b_stages = (int)((budget - scratch - (size_t)tile_stages * kTileBytes) / stage_bytes);
      if (b_stages >= 2) break;
    }
    if (b_stages > kMaxBStages) b_stages = kMaxBStages;
    if (b_stages < 2 || tile_stages < 2) return fail("lmdec_geno_matmul: shared memory budget");
  }
  p.stats_off = (int)((size_t)tile_stages * kTileBytes + b_stages * stage_bytes);
  p.b_stages = b_stages;
  p.tile_stages = tile_stages;
  const size_t smem = (size_t)tile_stages * kTileBytes + b_stages * stage_bytes + scratch;
  cudaStream_t st = (cudaStream_t)stream;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(geno_mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 223 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(geno_mma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 223 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(geno_mma_kernel<(kProducerWarps >= 16 ? 4 : 2)>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               223 * 1024);
    if (e != cudaSuccess) return fail("lmdec_geno_matmul: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  if (p.use_atomics) {
    cudaMemsetAsync(out0, 0, (size_t)m_limit * K * 8, st);
    if (mode == 1) cudaMemsetAsync(out1, 0, (size_t)m_limit * K * 8, st);
  }
  const int n_work = p.m_units * p.splits;
  const int sm_avail = n_sm - ctx.sm_reserve > 8 ? n_sm - ctx.sm_reserve : 8;
  const int grid = n_work < sm_avail ? n_work : sm_avail;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (ctx.timing) {
    cudaEventCreate(&ev0);
    cudaEventCreate(&ev1);
    cudaEventRecord(ev0, st);
  }
  // producer grouping (see geno_mma.cuh): two groups without missing data, one with
  int groups = has_missing ? ctx.producer_groups_missing : ctx.producer_groups;
#ifndef LMDEC_GROUPS_ANY_STAGES
  if (has_missing && p.a_stages < 3) groups = 1;
#endif
  if (groups > p.a_stages) groups = p.a_stages >= 2 ? 2 : 1;   // a group owns at least one stage of the decoded-A ring
  if (groups == 4 && kProducerWarps < 16) groups = 2;
  if (groups == 4)
    geno_mma_kernel<(kProducerWarps >= 16 ? 4 : 2)><<<grid, kThreads, smem, st>>>(p);
  else if (groups == 2)
    geno_mma_kernel<2><<<grid, kThreads, smem, st>>>(p);
  else
    geno_mma_kernel<1><<<grid, kThreads, smem, st>>>(p);
  if (ctx.timing) {
    cudaEventRecord(ev1, st);
    ctx_in->events.emplace_back(ev0, ev1);
  }
  return check_launch("lmdec_geno_matmul");
}

extern "C" {

int lmdec_geno_matmul(lmdec_ctx* ctx, const uint8_t* store, int64_t M, int64_t Kd, int64_t m_limit, int64_t kd_limit,
                      int has_missing, int mode, const int8_t* image1, const int8_t* image2, int K, int P, int N,
                      int64_t* out0, int64_t* out1, int splits, void* stream) {
  return launch_geno_mma(ctx, store, M, Kd, m_limit, kd_limit, has_missing, mode, image1, image2, K, P, N, out0, out1,
                         splits, stream, nullptr, nullptr);
}

int lmdec_finalize(const int64_t* acc0, const int64_t* acc1, int64_t M, int K, const double* colscale,
                   const double* rowa, const double* rowscale, const double* rowshift, const double* colshift,
                   double* out, int64_t ldo, void* stream) {
  if (acc1 && !rowa) return fail("lmdec_finalize: acc1 needs rowa");
  finalize_kernel<double><<<(unsigned)cdiv(M * K, 256), 256, 0, (cudaStream_t)stream>>>(
      (const long long*)acc0, (const long long*)acc1, M, K, colscale, rowa, rowscale, rowshift, colshift, out, ldo);
  return check_launch("lmdec_finalize");
}

static int gram_blocks(int64_t m) {
  int64_t b = cdiv(m, kGramRows);
  if (b > kGramMaxBlocks) b = kGramMaxBlocks;
  if (b < 1) b = 1;
  return (int)b;
}
int64_t lmdec_gram_work_doubles(int64_t m, int K) { return (int64_t)gram_blocks(m) * (K * K + K); }

int lmdec_gram(const double* X, int64_t ldx, const double* Y, int64_t ldy, int64_t m, int K, double* G, double* colsum,
               double* work, void* stream) {
  if (K < 1 || K > 128) return fail("lmdec_gram: K out of range (1..128)");
  const int nb = gram_blocks(m);
  const size_t smem = (size_t)kGramRows * K * sizeof(double) * (Y ? 2 : 1);
  cudaStream_t st = (cudaStream_t)stream;
  static thread_local bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(gram_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    attr_set = true;
  }
  gram_partial_kernel<<<nb, 256, smem, st>>>(X, ldx, Y, ldy, m, K, work);
  gram_final_kernel<<<(unsigned)cdiv(K * K + K, 128), 128, 0, st>>>(work, nb, K, G, colsum);
  return check_launch("lmdec_gram", 2);
}

int lmdec_gram_reduce(const double* X, int64_t ldx, int64_t m, int K, double* G_multicast, double* work, void* stream) {
  if (K < 1 || K > 128) return fail("lmdec_gram_reduce: K out of range (1..128)");
  if (!G_multicast) return fail("lmdec_gram_reduce: null multicast address");
  if (m <= 0) return 0;                  // a rank without rows in this prefix adds nothing
  const int nb = gram_blocks(m);
  cudaStream_t st = (cudaStream_t)stream;
  static thread_local bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(gram_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    attr_set = true;
  }
  gram_partial_kernel<<<nb, 256, (size_t)kGramRows * K * sizeof(double), st>>>(X, ldx, nullptr, 0, m, K, work);
  gram_final_reduce_kernel<<<(unsigned)cdiv(K * K, 128), 128, 0, st>>>(work, nb, K, G_multicast);
  return check_launch("lmdec_gram_reduce", 2);
}

int lmdec_qvals(const double* si, const double* sj, int k, int scale, double* out, void* stream) {
  qvals_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(si, sj, k, scale, out);
  return check_launch("lmdec_qvals");
}

int lmdec_sqdiff(const double* A, int64_t lda, const double* B, int64_t ldb_, const double* s, int64_t m, int K,
                 double* out, double* work, void* stream) {
  int64_t nb = cdiv(m * K, 256 * 8);
  if (nb > 592) nb = 592;
  if (nb < 1) nb = 1;
  cudaStream_t st = (cudaStream_t)stream;
  sqdiff_partial_kernel<<<(unsigned)nb, 256, 0, st>>>(A, lda, B, ldb_, s, m, K, work);
  sum_final_kernel<<<1, 32, 0, st>>>(work, (int)nb, out);
  return check_launch("lmdec_sqdiff", 2);
}

// [0, 5K) statistics | [5K, 5K + 592*3K) partial rows | one more double: the ticket counter of the epilogue's last-CTA
// fold (flag 16), which must be zero before the first call and is left zero by every launch
int64_t lmdec_apply_work_doubles(int K) { return (int64_t)5 * K + (int64_t)kStatsMaxBlocks * 3 * K + 1; }

// column statistics of a float64 / float32 operand: partial rows, then the fold into stats (multiplier for `lim`, its
// inverse, rank-1 shift, maxima)
static int launch_colstats(const void* B, bool b_f32, int64_t ldb, int64_t kd_limit, int K, const double* row1,
                           const double* row2, const double* w, int use_w, double lim, double* stats, double* partial,
                           cudaStream_t st) {
  const int rpb = 256 / K > 0 ? 256 / K : 1;
  int64_t blocks = cdiv(kd_limit, (int64_t)rpb * 8);
  if (blocks > kStatsMaxBlocks) blocks = kStatsMaxBlocks;
  if (blocks < 1) blocks = 1;
  const int threads = 256;
  if (b_f32)
    colstats_partial_kernel<float><<<(unsigned)blocks, threads, 3 * threads * sizeof(double), st>>>(
        (const float*)B, ldb, kd_limit, K, row1, row2, w, use_w, partial);
  else
    colstats_partial_kernel<double><<<(unsigned)blocks, threads, 3 * threads * sizeof(double), st>>>(
        (const double*)B, ldb, kd_limit, K, row1, row2, w, use_w, partial);
  colstats_finish_kernel<<<1, 1024, (size_t)(1024 / K) * 3 * K * sizeof(double), st>>>(partial, (int)blocks, K, lim, -1.0,
                                                                                          stats);
  return check_launch("lmdec_operator_apply (colstats)", 2);
}

}  // extern "C"

static int operator_apply_impl(lmdec_ctx* ctx, const uint8_t* store, int64_t M, int64_t Kd, int64_t m_limit,
                               int64_t kd_limit, int has_missing, int transpose_flags, const void* B, int64_t ldb, int K,
                               int P, const double* scale, const double* impute, const double* center_w, double* work,
                               int8_t* image1, int8_t* image2, int64_t* acc0, int64_t* acc1, void* out, int64_t ldo,
                               void* stream, float* mc_out) {
  if (K < 1 || K > 128 || P < 1 || P > 4) return fail("lmdec_operator_apply: bad K/P");
  const int transpose = transpose_flags & 1;
  const bool reuse_operand = (transpose_flags & 2) != 0;  // stats + plane images of B are already in work / image1/2
  const bool b_f32 = (transpose_flags & 4) != 0;          // B is float32 (else float64)
  const bool out_f32 = (transpose_flags & 8) != 0;        // out is float32 (else float64)
  const bool emit_stats = (transpose_flags & 16) != 0;    // A^T y only: leave out's column statistics in work
  const bool stats_ready = (transpose_flags & 32) != 0;   // B's column statistics are already in work
  const int N = (K * P + 15) / 16 * 16;
  if (N > 128) return fail("lmdec_operator_apply: K*P needs more than 128 tensor-core columns");
  if (has_missing && !impute) return fail("lmdec_operator_apply: has_missing needs the imputation vector");
  if (has_missing && ((!transpose && !image2) || (transpose && !acc1)))
    return fail("lmdec_operator_apply: has_missing needs image2 (A.dot) / acc1 (A.T.dot)");
  if ((emit_stats && !transpose) || (stats_ready && reuse_operand))
    return fail("lmdec_operator_apply: flag 16 goes with A.T.dot, flag 32 excludes flag 2");
  cudaStream_t st = (cudaStream_t)stream;
  double* stats = work;
  double* partial = work + 5 * K;
  const bool two = has_missing && !transpose;
  // ---- column statistics -> fixed-point multiplier and rank-1 shift
  const double lim = ldexp(1.0, 8 * P - (two ? 3 : 2));
  int rc = 0;
  if (!reuse_operand) {
    if (!stats_ready) {
      rc = launch_colstats(B, b_f32, ldb, kd_limit, K, transpose ? nullptr : scale, two ? impute : nullptr,
                           transpose ? nullptr : center_w, center_w != nullptr, lim, stats, partial, st);
      if (rc) return rc;
    }
    // ---- int8 planes
    rc = b_f32 ? quantize_planes_t<float>((const float*)B, ldb, Kd, kd_limit, K, P, N, stats,
                                          transpose ? nullptr : scale, two ? impute : nullptr, image1,
                                          two ? image2 : nullptr, stream)
               : quantize_planes_t<double>((const double*)B, ldb, Kd, kd_limit, K, P, N, stats,
                                           transpose ? nullptr : scale, two ? impute : nullptr, image1,
                                           two ? image2 : nullptr, stream);
    if (rc) return rc;
  }
  // ---- tensor-core product, with the diagonal / rank-1 terms fused into its epilogue when it is not split
  unsigned int* ticket = reinterpret_cast<unsigned int*>(work + lmdec_apply_work_doubles(K) - 1);
  // the following A.dot quantises `out` with has_missing ? 8P-3 : 8P-2 bits below its column maximum
  const double lim_next = ldexp(1.0, 8 * P - (has_missing ? 3 : 2));
  FusedFinalize fz;
  fz.colscale = stats + K;
  fz.colshift = center_w ? stats + 2 * K : nullptr;
  fz.rowa = (transpose && has_missing) ? impute : nullptr;
  fz.rowscale = transpose ? scale : nullptr;
  fz.rowshift = transpose ? center_w : nullptr;
  fz.out = out;
  fz.ldo = ldo;
  fz.out_f32 = out_f32 ? 1 : 0;
  fz.stats_partial = emit_stats ? partial : nullptr;
  fz.stats_final = emit_stats ? stats : nullptr;    // (the kernel has read its own colscale / colshift by then)
  fz.stats_lim = lim_next;
  fz.stats_ticket = ticket;
  fz.mc_out = mc_out;
  if (mc_out && (!transpose || !out_f32 || emit_stats || ldo != K || (K & 3) || (reinterpret_cast<uintptr_t>(out) & 15) ||
                 (reinterpret_cast<uintptr_t>(mc_out) & 15)))
    return fail("lmdec_operator_apply_reduce: needs A^T y, float32 output, ldo == K, K % 4 == 0, 16-byte alignment");
  bool fused = false, emitted = false;
  rc = launch_geno_mma(ctx, store, M, Kd, m_limit, kd_limit, has_missing, (transpose && has_missing) ? 1 : 0, image1,
                       two ? image2 : nullptr, K, P, N, acc0, (transpose && has_missing) ? acc1 : nullptr,
                       mc_out ? 1 : 0, stream, &fz, &fused, &emitted);
  if (rc) return rc;
  if (!fused) {
    if (!out_f32)
      rc = lmdec_finalize(acc0, fz.rowa ? acc1 : nullptr, m_limit, K, fz.colscale, fz.rowa, fz.rowscale, fz.rowshift,
                          fz.colshift, (double*)out, ldo, stream);
    else {
      finalize_kernel<float><<<(unsigned)cdiv(m_limit * K, 256), 256, 0, st>>>(
          (const long long*)acc0, (const long long*)(fz.rowa ? acc1 : nullptr), m_limit, K, fz.colscale, fz.rowa,
          fz.rowscale, fz.rowshift, fz.colshift, (float*)out, ldo);
      rc = check_launch("lmdec_operator_apply (finalize)");
    }
    if (rc) return rc;
  }
  if (emit_stats && !emitted)
    // the epilogue could not take the statistics along (split contraction or K > 32): one pass over out
    rc = launch_colstats(out, out_f32, ldo, m_limit, K, scale, has_missing ? impute : nullptr, center_w,
                         center_w != nullptr, lim_next, stats, partial, st);
  return rc;
}

extern "C" {

int lmdec_operator_apply(lmdec_ctx* ctx, const uint8_t* store, int64_t M, int64_t Kd, int64_t m_limit,
                         int64_t kd_limit, int has_missing, int transpose_flags, const void* B, int64_t ldb, int K, int P,
                         const double* scale, const double* impute, const double* center_w, double* work,
                         int8_t* image1, int8_t* image2, int64_t* acc0, int64_t* acc1, void* out, int64_t ldo,
                         void* stream) {
  return operator_apply_impl(ctx, store, M, Kd, m_limit, kd_limit, has_missing, transpose_flags, B, ldb, K, P, scale,
                             impute, center_w, work, image1, image2, acc0, acc1, out, ldo, stream, nullptr);
}

int lmdec_operator_apply_reduce(lmdec_ctx* ctx, const uint8_t* store, int64_t M, int64_t Kd, int64_t m_limit,
                                int64_t kd_limit, int has_missing, int transpose_flags, const void* B, int64_t ldb, int K,
                                int P, const double* scale, const double* impute, const double* center_w, double* work,
                                int8_t* image1, int8_t* image2, int64_t* acc0, int64_t* acc1, float* out,
                                float* out_multicast, void* stream) {
  if (!out_multicast) return fail("lmdec_operator_apply_reduce: null multicast address");
  return operator_apply_impl(ctx, store, M, Kd, m_limit, kd_limit, has_missing, transpose_flags, B, ldb, K, P, scale,
                             impute, center_w, work, image1, image2, acc0, acc1, out, K, stream, out_multicast);
}

__global__ void scale_inplace_kernel(double* __restrict__ x, int64_t m, int K, int64_t ld, double a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m * K) x[(i / K) * ld + i % K] *= a;
}

// One application of the Gram operator, the whole step of the power iteration (lmdec/decomp/iter_methods.py:383-385):
//   t = A^T q  (p x K, float32)      x = A t * inv_factor  (n_limit x K, float64)
// Two passes over the two orientations of the store, every kernel of both enqueued by this one call: operand statistics
// of q (skipped with flag 1), planes of q, A^T q with the statistics of t reduced in its epilogue, planes of t, A t with
// the diagonal / rank-1 terms in its epilogue, the division by the factor.
int lmdec_apass(lmdec_ctx* ctx, const uint8_t* sample_store, const uint8_t* snp_store, int64_t n, int64_t p,
                int64_t n_limit, int has_missing, const void* q, int64_t ldq, int q_flags, int K, int P,
                const double* scale, const double* impute, const double* center_w, double* work, int8_t* image_q,
                int8_t* image1, int8_t* image2, int64_t* acc0, int64_t* acc1, float* t_out, double* x_out, int64_t ldx,
                double inv_factor, void* stream) {
  if (!sample_store || !snp_store || !q || !t_out || !x_out) return fail("lmdec_apass: null operand");
  if (n_limit <= 0 || n_limit > n) return fail("lmdec_apass: bad row limit");
  const int q_f32 = (q_flags & 2) ? 4 : 0;
  const int q_ready = (q_flags & 1) ? 32 : 0;
  // A^T q: transpose (1) | float32 out (8) | leave the statistics of t in work (16)
  int rc = lmdec_operator_apply(ctx, snp_store, p, n, p, n_limit, has_missing, 1 | 8 | 16 | q_f32 | q_ready, q, ldq, K, P,
                                scale, impute, center_w, work, image_q, nullptr, acc0, acc1, t_out, K, stream);
  if (rc) return rc;
  // A t: float32 operand (4) | its statistics are in work (32)
  rc = lmdec_operator_apply(ctx, sample_store, n, p, n_limit, p, has_missing, 4 | 32, t_out, K, K, P, scale, impute,
                            center_w, work, image1, image2, acc0, acc1, x_out, ldx, stream);
  if (rc) return rc;
  if (inv_factor != 1.0) {
    scale_inplace_kernel<<<(unsigned)cdiv(n_limit * K, 256), 256, 0, (cudaStream_t)stream>>>(x_out, n_limit, K, ldx,
                                                                                           inv_factor);
    rc = check_launch("lmdec_apass (factor)");
  }
  return rc;
}

int lmdec_chol_whiten(const double* G, int K, double rank_tol, double* r_inv, double* r, int* flag, void* stream) {
  if (K < 1 || K > 128) return fail("lmdec_chol_whiten: K out of range (1..128)");
  const size_t smem = (size_t)(2 * K * K + K) * sizeof(double);
  static thread_local bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(chol_whiten_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    attr_set = true;
  }
  chol_whiten_kernel<<<1, 256, smem, (cudaStream_t)stream>>>(G, K, rank_tol, r_inv, r, flag, 0, nullptr);
  return check_launch("lmdec_chol_whiten");
}

// packed[0 .. K*K) = r2 r1,  packed[K*K] = flag1, [K*K+1] = flag2, [K*K+2] = max |G2 - I|: everything the host needs
// from one CholeskyQR2 in a single contiguous buffer (one device->host copy per iterate).
__global__ void qr_pack_kernel(const double* __restrict__ r1, const double* __restrict__ r2,
                               const double* __restrict__ g2, const int* __restrict__ f1, const int* __restrict__ f2,
                               int K, double* __restrict__ packed) {
  __shared__ double red[128];
  extern __shared__ double qsm[];        // r1 | r2 staged in shared memory when they fit (K <= 64)
  const int t = threadIdx.x;
  if (K <= 64) {
    for (int i = t; i < K * K; i += blockDim.x) {
      qsm[i] = r1[i];
      qsm[K * K + i] = r2[i];
    }
    __syncthreads();
    r1 = qsm;
    r2 = qsm + K * K;
  }
  double dev = 0.0;
  for (int i = t; i < K * K; i += blockDim.x) {
    const int a = i / K, b = i - a * K;
    double acc = 0.0;
    for (int c = a; c <= b; ++c) acc += r2[a * K + c] * r1[c * K + b];  // both factors are upper triangular
    packed[i] = acc;
    const double g = g2[i] - (a == b ? 1.0 : 0.0);
    dev = fmax(dev, g == g ? fabs(g) : INFINITY);
  }
  red[t] = dev;
  __syncthreads();
  if (t == 0) {
    for (int i = 1; i < blockDim.x; ++i) dev = fmax(dev, red[i]);
    packed[K * K] = (double)*f1;
    packed[K * K + 1] = (double)*f2;
    packed[K * K + 2] = dev;
  }
}

int lmdec_gram_chol(const double* X, int64_t ldx, int64_t m, int K, double rank_tol, double* G, double* r_inv,
                    double* r, int* flag, double* work, void* stream) {
  if (K < 1 || K > 128) return fail("lmdec_gram_chol: K out of range (1..128)");
  const int nb = gram_blocks(m);
  cudaStream_t st = (cudaStream_t)stream;
  static thread_local bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(gram_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(chol_whiten_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    attr_set = true;
  }
  gram_partial_kernel<<<nb, 256, (size_t)kGramRows * K * sizeof(double), st>>>(X, ldx, nullptr, 0, m, K, work);
  if ((int64_t)nb * K * K <= (1 << 18)) {
    // few partials: the factorisation block sums them itself (one launch less)
    chol_whiten_kernel<<<1, 256, (size_t)(2 * K * K + K) * sizeof(double), st>>>(work, K, rank_tol, r_inv, r, flag, nb, G);
    return check_launch("lmdec_gram_chol", 2);
  }
  // wide iterates (configs[4]: K = 110, a thousand partial matrices = 100 MB): one thread per matrix element sums them
  // in block order across the whole GPU; a single block doing it took 8 ms per call
  gram_final_kernel<<<(unsigned)cdiv(K * K + K, 128), 128, 0, st>>>(work, nb, K, G, nullptr);
  chol_whiten_kernel<<<1, 256, (size_t)(2 * K * K + K) * sizeof(double), st>>>(G, K, rank_tol, r_inv, r, flag, 0, nullptr);
  return check_launch("lmdec_gram_chol", 3);
}

int64_t lmdec_cholqr2_small_max_elems(void) { return (int64_t)kCqrCluster * kCqrSliceDoubles; }

int lmdec_cholqr2_small(const double* X, int64_t ldx, int64_t n, int K, double rank_tol, double* Q, int64_t ldq,
                        double* packed, double* stats_work, double stats_lim, void* stream) {
  if (K < 1 || K > 32) return fail("lmdec_cholqr2_small: K out of range (1..32)");
  if (n < 1 || Q == X) return fail("lmdec_cholqr2_small: needs n >= 1 and an output distinct from the input");
  const int64_t per = (n + kCqrCluster - 1) / kCqrCluster;
  if (per * K > kCqrSliceDoubles) return fail("lmdec_cholqr2_small: iterate too large (see lmdec_cholqr2_small_max_elems)");
  const size_t smem = ((size_t)2 * per * K + 9 * K * K + 4 * K + kCqrPaccDoubles) * sizeof(double);
  static thread_local bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(cholqr2_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024);
    if (kCqrCluster > 8) cudaFuncSetAttribute(cholqr2_small_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    attr_set = true;
  }
  cholqr2_small_kernel<<<kCqrCluster, kCqrThreads, smem, (cudaStream_t)stream>>>(X, ldx, n, K, (int)per, rank_tol, Q,
                                                                                 ldq, packed, stats_work, stats_lim);
  return check_launch("lmdec_cholqr2_small");
}

static int sm_count() {
  static thread_local int n_sm = -1;
  if (n_sm < 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      n_sm = 0;
  }
  return n_sm;
}

int64_t lmdec_cholqr2_grid_max_elems(void) { return (int64_t)sm_count() * kGridSliceDoubles; }
int64_t lmdec_cholqr2_grid_work_doubles(int K) { return (int64_t)(sm_count() + 2) * K * K + (int64_t)sm_count() * 2 * K; }

int lmdec_cholqr2_grid(const double* X, int64_t ldx, int64_t n, int K, double rank_tol, double* Q, int64_t ldq,
                       double* packed, double* gwork, double* stats_work, double stats_lim, void* stream) {
  if (K < 1 || K > 32) return fail("lmdec_cholqr2_grid: K out of range (1..32)");
  if (n < 1 || Q == X || !gwork) return fail("lmdec_cholqr2_grid: needs n >= 1, an output distinct from the input, gwork");
  const int n_sm = sm_count();
  if (n_sm < 1) return fail("lmdec_cholqr2_grid: no CUDA device");
  int64_t grid = cdiv(n, 8);
  if (grid > n_sm) grid = n_sm;
  int64_t per = (cdiv(n, grid) + 7) / 8 * 8;      // 8-row DMMA tiles never straddle CTAs
  grid = cdiv(n, per);
  if (per * K > kGridSliceDoubles) return fail("lmdec_cholqr2_grid: iterate too large (see lmdec_cholqr2_grid_max_elems)");
  const size_t smem = ((size_t)per * K + 9 * K * K + 4 * K + kCqrPaccDoubles) * sizeof(double);
  static thread_local bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(cholqr2_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024);
    attr_set = true;
  }
  int per_i = (int)per;
  void* args[] = {(void*)&X, (void*)&ldx, (void*)&n, (void*)&K, (void*)&per_i, (void*)&rank_tol, (void*)&Q,
                  (void*)&ldq, (void*)&packed, (void*)&gwork, (void*)&stats_work, (void*)&stats_lim};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)cholqr2_grid_kernel, dim3((unsigned)grid), dim3(kCqrThreads),
                                              args, smem, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail("lmdec_cholqr2_grid: cooperative launch: %s", cudaGetErrorString(e));
  return check_launch("lmdec_cholqr2_grid");
}

int lmdec_qr_pack(const double* r1, const double* r2, const double* g2, const int* flag1, const int* flag2, int K,
                  double* packed, void* stream) {
  if (K < 1 || K > 128) return fail("lmdec_qr_pack: K out of range (1..128)");
  const size_t smem = K <= 64 ? (size_t)2 * K * K * sizeof(double) : 0;
  static thread_local bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(qr_pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024 + 1024);
    attr_set = true;
  }
  qr_pack_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(r1, r2, g2, flag1, flag2, K, packed);
  return check_launch("lmdec_qr_pack");
}

// ---- CUDA-event timing of the dominant kernel (bench.py's roofline): events are recorded on the launching stream
#ifdef LMDEC_DEBUG_HANG
int lmdec_hang_read(unsigned* host_out) {   // [0] = count, then (block, warp, barrier address, parity) x 64
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(host_out, g_hang_n, sizeof(unsigned));
  cudaMemcpyFromSymbol(host_out + 4 + 256, g_prog, sizeof(unsigned) * 148 * 8);
  return (int)cudaMemcpyFromSymbol(host_out + 4, g_hang, sizeof(uint4) * 64);
}
#endif
#ifdef LMDEC_TRACE
int lmdec_trace_read(long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, g_trace, sizeof(long long) * 12 * kTraceChunks);
}
// wait-time accounting of CTA 0 (see LMDEC_WAITACC); reset != 0 zeroes it after the read
int lmdec_wait_read(long long* host_out, int reset) {
  int rc = (int)cudaMemcpyFromSymbol(host_out, g_wait, sizeof(long long) * 16);
  if (reset) {
    long long z[16] = {0};
    cudaMemcpyToSymbol(g_wait, z, sizeof(z));
  }
  return rc;
}
#endif

lmdec_ctx* lmdec_ctx_create(void) { return new lmdec_ctx(); }

static void ctx_drop_events(lmdec_ctx* ctx) {
  for (auto& e : ctx->events) {
    cudaEventDestroy(e.first);
    cudaEventDestroy(e.second);
  }
  ctx->events.clear();
}

void lmdec_ctx_destroy(lmdec_ctx* ctx) {
  if (!ctx) return;
  ctx_drop_events(ctx);
  delete ctx;
}

int lmdec_ctx_set(lmdec_ctx* ctx, int knob, int value) {
  if (!ctx) return fail("lmdec_ctx_set: null context");
  switch (knob) {
    case LMDEC_KNOB_SM_RESERVE: ctx->sm_reserve = value < 0 ? 0 : value; return 0;
    case LMDEC_KNOB_PAIR_MODE: ctx->pair_mode = value != 0; return 0;
    case LMDEC_KNOB_PAIR_MIN_CHUNKS: ctx->pair_min_chunks = value < 2 ? 2 : value; return 0;
    case LMDEC_KNOB_ISSUER_GROUPS: ctx->issuer_groups = value == 1 ? 1 : 2; return 0;
    case LMDEC_KNOB_PRODUCER_GROUPS: ctx->producer_groups = value == 1 ? 1 : (value == 4 ? 4 : 2); return 0;
    case LMDEC_KNOB_PRODUCER_GROUPS_MISSING: ctx->producer_groups_missing = value == 2 ? 2 : (value == 4 ? 4 : 1); return 0;
    case LMDEC_KNOB_ACC_SETS: ctx->acc_sets = value == 1 || value == 2 ? value : 0; return 0;
    case LMDEC_KNOB_B_RESIDENT: ctx->b_resident = value != 0; return 0;
    case LMDEC_KNOB_ROUND_ROBIN: ctx->rr = value != 0; return 0;
  }
  return fail("lmdec_ctx_set: unknown knob");
}

int lmdec_ctx_get(const lmdec_ctx* ctx, int knob) {
  const lmdec_ctx& c = ctx ? *ctx : kDefaultCtx;
  switch (knob) {
    case LMDEC_KNOB_SM_RESERVE: return c.sm_reserve;
    case LMDEC_KNOB_PAIR_MODE: return c.pair_mode;
    case LMDEC_KNOB_PAIR_MIN_CHUNKS: return c.pair_min_chunks;
    case LMDEC_KNOB_ISSUER_GROUPS: return c.issuer_groups;
    case LMDEC_KNOB_PRODUCER_GROUPS: return c.producer_groups;
    case LMDEC_KNOB_PRODUCER_GROUPS_MISSING: return c.producer_groups_missing;
    case LMDEC_KNOB_ACC_SETS: return c.acc_sets;
    case LMDEC_KNOB_B_RESIDENT: return c.b_resident;
    case LMDEC_KNOB_ROUND_ROBIN: return c.rr;
  }
  return -1;
}

void lmdec_ctx_timing(lmdec_ctx* ctx, int on) {
  if (!ctx) return;
  ctx_drop_events(ctx);
  ctx->timing = on != 0;
}
int lmdec_ctx_timing_read(lmdec_ctx* ctx, double* total_ms, int64_t* launches) {
  if (!ctx) return fail("lmdec_ctx_timing_read: null context");
  double tot = 0;
  for (auto& e : ctx->events) {
    cudaEventSynchronize(e.second);
    float ms = 0;
    cudaEventElapsedTime(&ms, e.first, e.second);
    tot += ms;
  }
  *total_ms = tot;
  *launches = (int64_t)ctx->events.size();
  return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// Dense float32 operands (dense_mma.cuh): tile store, TF32 hi / lo operand images, the 3xTF32 tcgen05 product
// ---------------------------------------------------------------------------------------------------------------
extern "C" {

int64_t lmdec_dense_store_bytes(int64_t M, int64_t Kd) { return cdiv(M, 128) * cdiv(Kd, kDChunkElems) * (int64_t)kDTileBytes; }
int64_t lmdec_dense_image_bytes(int64_t Kd, int N) { return cdiv(Kd, kDChunkElems) * kDKb * (int64_t)N * 32; }

int lmdec_dense_pack(const void* src, int dtype, int64_t rows, int64_t cols, int64_t ld, int transpose, float* store,
                     void* stream) {
  if (!src || !store || rows <= 0 || cols <= 0 || ld < cols) return fail("lmdec_dense_pack: bad arguments");
  const int64_t M = transpose ? cols : rows, Kd = transpose ? rows : cols;
  const int64_t n_kc = cdiv(Kd, kDChunkElems), n_units = cdiv(M, 128) * n_kc * 1024;
  const unsigned blocks = (unsigned)(cdiv(n_units, 256) < (1 << 20) ? cdiv(n_units, 256) : (1 << 20));
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == 0)
    dense_pack_kernel<float><<<blocks, 256, 0, st>>>((const float*)src, rows, cols, ld, transpose, store, M, Kd, n_kc, n_units);
  else if (dtype == 1)
    dense_pack_kernel<double><<<blocks, 256, 0, st>>>((const double*)src, rows, cols, ld, transpose, store, M, Kd, n_kc,
                                                      n_units);
  else
    return fail("lmdec_dense_pack: dtype must be 0 (float32) or 1 (float64)");
  return check_launch("lmdec_dense_pack");
}

int lmdec_dense_unpack(const float* store, int64_t M, int64_t Kd, float* dst, int64_t ld, void* stream) {
  if (!store || !dst || M <= 0 || Kd <= 0 || ld < Kd) return fail("lmdec_dense_unpack: bad arguments");
  const int64_t n_kc = cdiv(Kd, kDChunkElems), n_units = cdiv(M, 128) * n_kc * 1024;
  const unsigned blocks = (unsigned)(cdiv(n_units, 256) < (1 << 20) ? cdiv(n_units, 256) : (1 << 20));
  dense_unpack_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(store, M, Kd, n_kc, dst, ld, n_units);
  return check_launch("lmdec_dense_unpack");
}

int lmdec_dense_matmul(lmdec_ctx* ctx_in, const float* store, int64_t M, int64_t Kd, int64_t m_limit, int64_t kd_limit,
                       const void* B, int b_dtype, int64_t ldb, int K, float* image_hi, float* image_lo, double* out,
                       int64_t ldo, int splits, void* stream) {
  if (!store || !B || !image_hi || !image_lo || !out) return fail("lmdec_dense_matmul: null argument");
  if (K < 1 || K > 128 || ldb < K || ldo < K) return fail("lmdec_dense_matmul: 1 <= K <= 128 operand columns per call");
  if (m_limit <= 0 || m_limit > M || kd_limit <= 0 || kd_limit > Kd) return fail("lmdec_dense_matmul: bad limits");
  const lmdec_ctx& ctx = ctx_in ? *ctx_in : kDefaultCtx;
  int n_sm = ctx.n_sm;
  if (!n_sm) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (ctx_in) ctx_in->n_sm = n_sm;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int N = (K + 15) / 16 * 16;
  DenseMmaParams p;
  memset(&p, 0, sizeof(p));
  p.store = (const uint8_t*)store;
  p.store_kc = cdiv(Kd, kDChunkElems);
  p.m_limit = m_limit;
  p.m_tiles = (int)cdiv(m_limit, 128);
  p.n_chunks = (int)cdiv(kd_limit, kDChunkElems);
  p.image_hi = (const uint8_t*)image_hi;
  p.image_lo = (const uint8_t*)image_lo;
  p.K = K;
  p.N = N;
  p.out = out;
  p.ldo = ldo;
  // operand images (rows beyond kd_limit are zero: a row prefix of the contraction needs no masking of the store)
  {
    const int64_t n_units = (int64_t)p.n_chunks * (kDChunkElems / 4) * N;
    const unsigned blocks = (unsigned)(cdiv(n_units, 256) < (1 << 20) ? cdiv(n_units, 256) : (1 << 20));
    if (b_dtype == 0)
      dense_image_kernel<float><<<blocks, 256, 0, st>>>((const float*)B, ldb, kd_limit, K, N, image_hi, image_lo, n_units);
    else if (b_dtype == 1)
      dense_image_kernel<double><<<blocks, 256, 0, st>>>((const double*)B, ldb, kd_limit, K, N, image_hi, image_lo, n_units);
    else
      return fail("lmdec_dense_matmul: b_dtype must be 0 (float32) or 1 (float64)");
  }
  // contraction ranges: enough work items to fill the SMs several times over when there are few m-tiles (A^T y of a
  // tall matrix), and never more than 2,048 chunks (65,536 elements) summed in one float32 accumulator
  int want = splits;
  if (want <= 0) {
    want = 1;
    if (p.m_tiles < 4 * n_sm) want = (int)cdiv(8 * (int64_t)n_sm, p.m_tiles);
    if (want > p.n_chunks / 16) want = p.n_chunks / 16 > 1 ? p.n_chunks / 16 : 1;
  }
  const int min_splits = (int)cdiv(p.n_chunks, 2048);
  if (want < min_splits) want = min_splits;
  if (want > p.n_chunks) want = p.n_chunks;
  const int cps = (int)cdiv(p.n_chunks, want);
  p.splits = (int)cdiv(p.n_chunks, cps);
  if (p.splits > 1) {
    if (ldo == K) cudaMemsetAsync(out, 0, (size_t)m_limit * K * sizeof(double), st);
    else cudaMemset2DAsync(out, (size_t)ldo * sizeof(double), 0, (size_t)K * sizeof(double), (size_t)m_limit, st);
  }
  const size_t stage_bytes = (size_t)2 * kDKb * N * 32;
  const size_t budget = 223 * 1024;
  int tile_stages = (int)((budget - kDBStages * stage_bytes) / kDTileBytes);
  if (tile_stages > kDMaxTileStages) tile_stages = kDMaxTileStages;
  if (tile_stages < 2) return fail("lmdec_dense_matmul: shared memory budget");
  p.tile_stages = tile_stages;
  const size_t smem = (size_t)tile_stages * kDTileBytes + kDBStages * stage_bytes;
  static thread_local bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(dense_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 223 * 1024);
    if (e != cudaSuccess) return fail("lmdec_dense_matmul: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int n_work = p.m_tiles * p.splits;
  const int grid = n_work < n_sm ? n_work : n_sm;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (ctx.timing) {
    cudaEventCreate(&ev0);
    cudaEventCreate(&ev1);
    cudaEventRecord(ev0, st);
  }
  dense_mma_kernel<<<grid, kDThreads, smem, st>>>(p);
  if (ctx.timing) {
    cudaEventRecord(ev1, st);
    ctx_in->events.emplace_back(ev0, ev1);
  }
  return check_launch("lmdec_dense_matmul", 2);
}

}  // extern "C"
