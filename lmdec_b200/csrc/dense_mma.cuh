// Dense float32 matrix times a narrow operand on the sm_100a tensor cores: the product behind `ChainedArray.dot` /
// `.T.dot` for dense float members (reference lmdec/array/chained.py:169-177 -> dask chunk matmuls; BASELINE configs[4]:
// 1M x 4k float, K = 110).
//
//   out[m, k] = sum_kd A[m, kd] * B[kd, k]          A float32 in a pre-tiled store, B [Kd, K] float64 / float32
//
// The product is HBM-bound by its algorithmic bytes (4 bytes of A per 2K flops: K = 110 -> 55 flop/B), so the kernel is
// built to stream A exactly once at HBM speed while the tensor cores do error-compensated TF32 ("3xTF32"):
//   A = A_hi + A_lo,  B = B_hi + B_lo   (hi = rounded to the 10 mantissa bits TF32 keeps, lo = the exact float32 remainder)
//   out = A_hi B_hi + A_lo B_hi + A_hi B_lo                       (dropped: A_lo B_lo, 2^-22 relative)
// accumulated in float32 in TMEM.  The tensor core adds into its accumulator with round-toward-zero (measured,
// tools/dense_err_scan.py: a same-sign sum loses 2^-25 of its value per MMA, -5.5e-6 after 512 contraction elements and
// growing linearly), so an accumulator only ever sums kDFlushChunks chunks (256 elements: -2.7e-6 worst case); the
// epilogue warps add these partial tiles up in registers with round-to-nearest float32 adds and write once per work
// item.  Three MMAs per 32-byte K-block
// of N = 112 columns take 3 x 56 cycles per 128 x 8 elements = 168 cycles per 4 KB of A; HBM delivers 4 KB per SM every
// 183 cycles (6.55 TB/s over 148 SMs at 1.965 GHz): the two ceilings meet, the tensor pipe is not the limiter.
//
// Same skeleton as geno_mma.cuh (one persistent CTA per SM, warp-specialised, A operand in TMEM, B image in shared
// memory), one schedule:
//   warps 0-7   epilogue (two per TMEM lane quarter, half of the columns each): tcgen05.ld of a partial accumulator
//               every kDFlushChunks chunks, running sums in registers, float64 store (atomicAdd for split contractions)
//   warps 8-15  producers: read their row's 64 bytes of the float32 tile from the smem ring, split hi / lo in registers
//               (two integer operations + one FADD per element), tcgen05.st both into the TMEM operand ring
//   warp  16    tile loader: cp.async.bulk of whole 16 KB store tiles (128 rows x 32 elements) into the smem ring
//   warps 17-18 MMA issuers (one elected lane each) on ALTERNATE chunks, 12 MMAs per chunk into the same accumulator.
//               A single issuer spends 60 % of its time issuing (the issue blocks on the tensor pipe) and 30 % in its
//               barrier waits, during which the pipe runs dry (profiles/r02/dense_waits.txt); with two, one waits for
//               its operands while the other issues.  They hand over in strict turn (one mbarrier hop per chunk), so
//               the order of the float32 additions -- and with it every bit of the result -- is fixed.
//   warp  19    B loader: cp.async.bulk of the hi / lo operand images (K-major, no-swizzle core-matrix order)
// A K-block of kind::tf32 is 8 elements = 32 bytes: byte for byte the operand geometry of kind::i8 (32 int8), so the
// TMEM addressing and the shared-memory descriptors are the ones geno_mma.cuh uses.
#pragma once
#include "umma.cuh"

namespace lmdec {

constexpr int kDEpilogueWarps = 8, kDProducerWarps = 8;
#ifndef LMDEC_DENSE_ISSUERS
#define LMDEC_DENSE_ISSUERS 2
#endif
constexpr int kDIssuers = LMDEC_DENSE_ISSUERS;            // 1 or 2
constexpr int kDEpilogueWarp0 = 0, kDProducerWarp0 = 8, kDTileLoaderWarp = 16, kDIssuerWarp = 17,
              kDBLoaderWarp = kDIssuerWarp + kDIssuers;
constexpr int kDThreads = (kDBLoaderWarp + 1 + 3) / 4 * 4 * 32;   // whole warpgroups (setmaxnreg is per warpgroup)
#ifndef LMDEC_DENSE_FLUSH_CHUNKS
#define LMDEC_DENSE_FLUSH_CHUNKS 8
#endif
constexpr int kDFlushChunks = LMDEC_DENSE_FLUSH_CHUNKS;   // chunks summed by the tensor core before the epilogue takes the partial tile
constexpr int kDKb = 4;                       // K-blocks (8 elements) per chunk: 32 contraction elements, 128 B per row
constexpr int kDChunkElems = kDKb * 8;
constexpr uint32_t kDTileBytes = 128 * kDChunkElems * 4;   // 16 KB: 8 pieces of [128 rows][16 bytes]
constexpr int kDMaxTileStages = 6, kDBStages = 4, kDAStages = 4;
constexpr uint32_t kDAccCols = 128;           // TMEM columns per accumulator set (N <= 128 float32 columns)
constexpr uint32_t kDARingCol0 = 2 * kDAccCols, kDAStageCols = 64;   // hi: 32 columns, lo: 32 columns
// mbarrier slots
constexpr int kDBarFullT = 0, kDBarEmptyT = 6, kDBarFullA = 12, kDBarEmptyA = 16, kDBarFullB = 20, kDBarEmptyB = 24,
              kDBarAccFull = 28, kDBarAccEmpty = 30, kDBarTurn = 32, kDNumBars = 34;

struct DenseMmaParams {
  const uint8_t* store;
  long long store_kc;      // chunks per row panel in the store
  long long m_limit;       // rows produced
  int m_tiles;
  int n_chunks;            // contraction chunks to run: ceil(kd_limit / 32)
  int splits;              // contraction ranges (work items per m-tile); > 1 => atomicAdd into a zeroed output
  const uint8_t* image_hi;
  const uint8_t* image_lo;
  int K, N;
  double* out;
  long long ldo;
  int tile_stages;
};

// Instruction descriptor for kind::tf32: A, B = tf32 (format 2), D = f32 (format 1), K-major operands.
__host__ __device__ constexpr uint32_t make_idesc_tf32(uint32_t M, uint32_t N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// D[tmem,f32] (+)= A[tmem,tf32] * B[smem,tf32]   (A: lane = row m, 32-bit column c = contraction element c)
__device__ __forceinline__ void umma_ts_tf32(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// LMDEC_DENSE_PROFILE: CTA 0 prints the cycles each role spent inside its barrier waits (tools/time_dense.py output)
#ifdef LMDEC_DENSE_PROFILE
#define DWAIT(acc, stmt) do { const long long t0_ = clock64(); stmt; acc += clock64() - t0_; } while (0)
#define DWAIT_DECL(...) long long __VA_ARGS__
#define DWAIT_PRINT(...) do { if (blockIdx.x == 0 && lane == 0) printf(__VA_ARGS__); } while (0)
#else
#define DWAIT(acc, stmt) do { stmt; } while (0)
#define DWAIT_DECL(...) do { } while (0)
#define DWAIT_PRINT(...) do { } while (0)
#endif

__device__ __forceinline__ uint4 dense_lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

__global__ void __launch_bounds__(kDThreads, 1) dense_mma_kernel(const DenseMmaParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[kDNumBars];
  __shared__ uint32_t tmem_base_s;
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t full_t = bar0 + 8 * kDBarFullT, empty_t = bar0 + 8 * kDBarEmptyT;
  const uint32_t full_a = bar0 + 8 * kDBarFullA, empty_a = bar0 + 8 * kDBarEmptyA;
  const uint32_t full_b = bar0 + 8 * kDBarFullB, empty_b = bar0 + 8 * kDBarEmptyB;
  const uint32_t acc_full = bar0 + 8 * kDBarAccFull, acc_empty = bar0 + 8 * kDBarAccEmpty;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.N;
  const uint32_t img_bytes = (uint32_t)kDKb * N * 32;      // one image (hi or lo) of one chunk
  const uint32_t stage_bytes = 2 * img_bytes;
  const uint32_t smem_t_a = smem_u32(smem);
  const uint32_t smem_b_a = smem_t_a + (uint32_t)p.tile_stages * kDTileBytes;

  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  if (tid == 32) {
    for (int i = 0; i < kDMaxTileStages; ++i) {
      mbar_init(&bars[kDBarFullT + i], 1);
      mbar_init(&bars[kDBarEmptyT + i], kDProducerWarps);
    }
    for (int i = 0; i < kDAStages; ++i) {
      mbar_init(&bars[kDBarFullA + i], kDProducerWarps);
      mbar_init(&bars[kDBarEmptyA + i], 1);
    }
    for (int i = 0; i < kDBStages; ++i) {
      mbar_init(&bars[kDBarFullB + i], 1);
      mbar_init(&bars[kDBarEmptyB + i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[kDBarAccFull + i], kDIssuers);
      mbar_init(&bars[kDBarAccEmpty + i], kDEpilogueWarps);
      mbar_init(&bars[kDBarTurn + i], 1);
    }
    mbar_fence_init();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tbase = tmem_base_s;

  // work item w -> (contraction range sp, m-tile mt): consecutive items are the m-tiles of ONE contraction range, so the
  // CTAs running at the same time read the same chunks of the operand image (one HBM read, the rest L2 hits)
  const int n_work = p.m_tiles * p.splits;
  const int chunks_per_split = (p.n_chunks + p.splits - 1) / p.splits;

  if (warp >= kDProducerWarp0 && warp < kDProducerWarp0 + kDProducerWarps) {
    // ------------------------------------------------------------------ producers: float32 -> (hi, lo) in TMEM
    reg_dec<72>();
    const int pw = warp - kDProducerWarp0;
    const int q = pw & 3, half = pw >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tbase + ((uint32_t)(q * 32) << 16) + kDARingCol0 + 16u * half;
    const uint32_t my_off = (uint32_t)(4 * half) * 2048 + (uint32_t)row * 16;   // pieces 4 half .. 4 half + 3
    uint32_t a_stage = 0, a_phase = 0, t_stage = 0, t_phase = 0;
    DWAIT_DECL(w_tile = 0, w_astage = 0, w_st = 0, t_begin = clock64());
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
      const int sp = w / p.m_tiles;
      const int c0 = sp * chunks_per_split;
      const int c1 = min(c0 + chunks_per_split, p.n_chunks);
      for (int c = c0; c < c1; ++c) {
        DWAIT(w_tile, mbar_wait_a(full_t + 8 * t_stage, t_phase));
        const uint32_t src = smem_t_a + t_stage * kDTileBytes + my_off;
        uint4 raw[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) raw[i] = dense_lds128(src + i * 2048);
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t x[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            // hi = x rounded to NEAREST at the 10 mantissa bits kind::tf32 reads (a carry into the exponent is the
            // correct rounding), so lo = x - hi is signed, 12 bits, and its own truncation by the tensor core
            // (11 bits) is unbiased: truncating hi instead leaves a one-sided 2^-23 error per term that accumulates
            const uint32_t h = (x[e] + 0x1000u) & 0xFFFFE000u;
            hi[4 * i + e] = h;
            lo[4 * i + e] = __float_as_uint(__uint_as_float(x[e]) - __uint_as_float(h));   // exact in float32
          }
        }
        DWAIT(w_astage, mbar_wait_a(empty_a + 8 * a_stage, a_phase ^ 1));
        tc_fence_after_sync();
        const uint32_t a_addr = lane_addr + a_stage * kDAStageCols;
        tmem_st16(a_addr, hi);
        tmem_st16(a_addr + 32, lo);
        DWAIT(w_st, tmem_wait_st());
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive_a(full_a + 8 * a_stage);
          mbar_arrive_a(empty_t + 8 * t_stage);
        }
        if (++a_stage == (uint32_t)kDAStages) {
          a_stage = 0;
          a_phase ^= 1;
        }
        if (++t_stage == (uint32_t)p.tile_stages) {
          t_stage = 0;
          t_phase ^= 1;
        }
      }
    }
    if (pw == 0) DWAIT_PRINT("producer 0: total %lld  wait tile %lld  wait A stage %lld  wait::st %lld\n",
                             clock64() - t_begin, w_tile, w_astage, w_st);
  } else if (warp >= kDTileLoaderWarp) {
   reg_dec<40>();                    // warps 16-19: one warpgroup of single-thread roles
   if (warp == kDTileLoaderWarp) {
    // ------------------------------------------------------------------ store-tile loader
    if (lane == 0) {
      uint32_t t_stage = 0, t_phase = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        const int mt = w % p.m_tiles, sp = w / p.m_tiles;
        const int c0 = sp * chunks_per_split;
        const int c1 = min(c0 + chunks_per_split, p.n_chunks);
        const uint8_t* tile = p.store + ((size_t)mt * p.store_kc) * kDTileBytes;
        for (int c = c0; c < c1; ++c) {
          mbar_wait_relaxed_a(empty_t + 8 * t_stage, t_phase ^ 1);
          mbar_arrive_expect_tx_a(full_t + 8 * t_stage, kDTileBytes);
          bulk_g2s_a(smem_t_a + t_stage * kDTileBytes, tile + (size_t)c * kDTileBytes, kDTileBytes, full_t + 8 * t_stage);
          if (++t_stage == (uint32_t)p.tile_stages) {
            t_stage = 0;
            t_phase ^= 1;
          }
        }
      }
    }
  } else if (warp == kDBLoaderWarp) {
    // ------------------------------------------------------------------ operand-image loader
    if (lane == 0) {
      uint32_t b_stage = 0, b_phase = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        const int sp = w / p.m_tiles;
        const int c0 = sp * chunks_per_split;
        const int c1 = min(c0 + chunks_per_split, p.n_chunks);
        for (int c = c0; c < c1; ++c) {
          mbar_wait_relaxed_a(empty_b + 8 * b_stage, b_phase ^ 1);
          const uint32_t dst = smem_b_a + b_stage * stage_bytes;
          mbar_arrive_expect_tx_a(full_b + 8 * b_stage, stage_bytes);
          bulk_g2s_a(dst, p.image_hi + (size_t)c * img_bytes, img_bytes, full_b + 8 * b_stage);
          bulk_g2s_a(dst + img_bytes, p.image_lo + (size_t)c * img_bytes, img_bytes, full_b + 8 * b_stage);
          if (++b_stage == (uint32_t)kDBStages) {
            b_stage = 0;
            b_phase ^= 1;
          }
        }
      }
    }
  } else if (warp >= kDIssuerWarp && warp < kDIssuerWarp + kDIssuers) {
    // ------------------------------------------------------------------ MMA issuers
    const uint32_t me = (uint32_t)(warp - kDIssuerWarp);
    const uint32_t turn_mine = bar0 + 8 * (kDBarTurn + me), turn_other = bar0 + 8 * (kDBarTurn + (me ^ 1u));
    const uint32_t idesc = make_idesc_tf32(128, (uint32_t)N);
    const uint64_t bdesc_ring = make_smem_desc(smem_b_a, (uint32_t)N * 16, 128);
    const uint32_t kb_step = (uint32_t)(2 * N);          // one K-block of the image, in 16-byte units
    uint32_t a_stage = 0, a_phase = 0, b_stage = 0, b_phase = 0, sub_no = 0;
    uint32_t gc = 0;                                     // running chunk count of this CTA: issuer gc % kDIssuers owns it
    DWAIT_DECL(w_acc = 0, w_b = 0, w_a = 0, w_turn = 0, w_issue = 0, t_begin = clock64());
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
      const int sp = w / p.m_tiles;
      const int c0 = sp * chunks_per_split;
      const int c1 = min(c0 + chunks_per_split, p.n_chunks);
      for (int g0 = c0; g0 < c1; g0 += kDFlushChunks, ++sub_no) {      // one partial accumulator per kDFlushChunks chunks
        const int g1 = min(g0 + kDFlushChunks, c1);
        const uint32_t set = sub_no & 1, set_phase = (sub_no >> 1) & 1;
        DWAIT(w_acc, mbar_wait_a(acc_empty + 8 * set, set_phase ^ 1));
        tc_fence_after_sync();
        const uint32_t d_addr = tbase + set * kDAccCols;
        bool issued = false;
        for (int c = g0; c < g1; ++c, ++gc) {
          if (kDIssuers == 1 || (gc & 1u) == me) {
            DWAIT(w_b, mbar_wait_a(full_b + 8 * b_stage, b_phase));
            DWAIT(w_a, mbar_wait_a(full_a + 8 * a_stage, a_phase));
            if (kDIssuers == 2 && gc > 0) {
              // my k-th chunk (k = gc / 2) goes after the other issuer's previous one: its arrival number k (me = 1)
              // or k - 1 (me = 0) on my turn barrier
              const uint32_t k = gc >> 1;
              DWAIT(w_turn, mbar_wait_a(turn_mine, (me ? k : k - 1) & 1u));
            }
            tc_fence_after_sync();
#ifdef LMDEC_DENSE_PROFILE
            const long long ti_ = clock64();
#endif
            if (elect_one()) {
              const uint32_t a_hi = tbase + kDARingCol0 + a_stage * kDAStageCols, a_lo = a_hi + 32;
              const uint64_t b_hi = bdesc_ring + (uint64_t)((b_stage * stage_bytes) >> 4);
              const uint64_t b_lo = b_hi + (uint64_t)(img_bytes >> 4);
#pragma unroll
              for (int kb = 0; kb < kDKb; ++kb) {
#if defined(LMDEC_DENSE_ABL_MMA1)      // ablation (timing only, wrong results): one of the three MMAs
                umma_ts_tf32(d_addr, a_hi + kb * 8, b_hi + kb * kb_step, idesc, (c == g0 && kb == 0) ? 0u : 1u);
#elif defined(LMDEC_DENSE_ABL_NOMMA)   // ablation (timing only): barriers and commits, no tensor-core work
#else
                // small terms first: they are added while the accumulator is still small
                umma_ts_tf32(d_addr, a_lo + kb * 8, b_hi + kb * kb_step, idesc, (c == g0 && kb == 0) ? 0u : 1u);
                umma_ts_tf32(d_addr, a_hi + kb * 8, b_lo + kb * kb_step, idesc, 1u);
                umma_ts_tf32(d_addr, a_hi + kb * 8, b_hi + kb * kb_step, idesc, 1u);
#endif
              }
              umma_commit_a(empty_a + 8 * a_stage);
              umma_commit_a(empty_b + 8 * b_stage);
              if (kDIssuers == 2) {
                tc_fence_before_sync();
                mbar_arrive_a(turn_other);          // the other issuer may queue the next chunk behind this one
              }
            }
            __syncwarp();
            issued = true;
#ifdef LMDEC_DENSE_PROFILE
            w_issue += clock64() - ti_;
#endif
          }
          if (++a_stage == (uint32_t)kDAStages) {
            a_stage = 0;
            a_phase ^= 1;
          }
          if (++b_stage == (uint32_t)kDBStages) {
            b_stage = 0;
            b_phase ^= 1;
          }
        }
        // the partial tile is complete when the MMAs of BOTH issuers are: each commits behind its own last chunk
        if (elect_one()) {
          if (issued) umma_commit_a(acc_full + 8 * set);
          else mbar_arrive_a(acc_full + 8 * set);
        }
        __syncwarp();
      }
    }
    DWAIT_PRINT("issuer %u: total %lld  wait acc %lld  wait B %lld  wait A %lld  wait turn %lld  issue+commit %lld\n", me,
                clock64() - t_begin, w_acc, w_b, w_a, w_turn, w_issue);
   }
  } else if (warp >= kDEpilogueWarp0 && warp < kDEpilogueWarp0 + kDEpilogueWarps) {
    // ------------------------------------------------------------------ epilogue: float32 accumulator -> float64 out
    reg_inc<144>();   // (144 - 96) x 256 threads <= what the other warpgroups gave back: (96 - 72) x 256 + (96 - 40) x 128
    const int e = warp - kDEpilogueWarp0;
    const int quarter = e & 3, colhalf = e >> 2;
    const int n8 = N / 16;                          // groups of 8 columns owned by this thread (N / 2 columns)
    const int col0 = colhalf * (N / 2);
    uint32_t sub_no = 0;
    const bool atomics = p.splits > 1;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
      const int mt = w % p.m_tiles, sp = w / p.m_tiles;
      const int c0 = sp * chunks_per_split;
      const int c1 = min(c0 + chunks_per_split, p.n_chunks);
      float sum[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) sum[i] = 0.f;
      for (int g0 = c0; g0 < c1; g0 += kDFlushChunks, ++sub_no) {
        const uint32_t set = sub_no & 1, set_phase = (sub_no >> 1) & 1;
        mbar_wait_relaxed_a(acc_full + 8 * set, set_phase);
        tc_fence_after_sync();
        const uint32_t t_addr = tbase + ((uint32_t)(quarter * 32) << 16) + set * kDAccCols + col0;
        // two round trips of up to 32 columns each; the set goes back to the issuers as soon as the second has landed
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t r[32];
#pragma unroll
          for (int g = 0; g < 4; ++g)
            if (4 * h + g < n8) tmem_ld8p(t_addr + (4 * h + g) * 8, r + g * 8);
          tmem_wait_ld();
          if (h == 1) {
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive_a(acc_empty + 8 * set);
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (4 * h + g < n8) {
#pragma unroll
              for (int j = 0; j < 8; ++j) sum[(4 * h + g) * 8 + j] += __uint_as_float(r[g * 8 + j]);   // round-to-nearest adds
            }
          }
        }
      }
      const long long row = (long long)mt * 128 + quarter * 32 + lane;
      if (row < p.m_limit) {
        double* orow = p.out + row * p.ldo;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          if (g < n8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int k = col0 + g * 8 + j;
              if (k < p.K) {
                if (atomics) atomicAdd(orow + k, (double)sum[g * 8 + j]);
                else orow[k] = (double)sum[g * 8 + j];
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tbase);
}

// ---------------------------------------------------------------------------------------------------------------
// HBM-bound helpers of the dense path
// ---------------------------------------------------------------------------------------------------------------
// Row-major float matrix -> tile store.  Tile (mp, kc) = rows [128 mp, +128) x contraction [32 kc, +32) = 16 KB, tiles
// mp-major; inside a tile the 16-byte piece j (elements 4j .. 4j+3) of row t sits at j * 2048 + t * 16 (piece-major:
// one bulk copy moves a tile, a warp reading "its rows' piece" reads 512 contiguous bytes).  `transpose`: the store
// holds src^T (M = src columns, contraction = src rows).  One thread per (tile, row t, piece j); padding is zero.
template <typename T>
__global__ void dense_pack_kernel(const T* __restrict__ src, long long rows, long long cols, long long ld, int transpose,
                                  float* __restrict__ store, long long M, long long Kd, long long n_kc, long long n_units) {
  for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < n_units;
       u += (long long)gridDim.x * blockDim.x) {
    const long long tile = u >> 10;
    const int r = (int)(u & 1023);
    // transposed: lanes run along t (src columns: coalesced reads, 512 contiguous bytes written);
    // direct: lanes run along j (32 lanes = 4 rows x 128 contiguous bytes read)
    const int t = transpose ? (r & 127) : (r >> 3);
    const int j = transpose ? (r >> 7) : (r & 7);
    const long long mp = tile / n_kc, kc = tile % n_kc;
    const long long m = mp * 128 + t, kd0 = kc * 32 + 4 * j;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m < M) {
      float x[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const long long kd = kd0 + e;
        x[e] = kd < Kd ? (float)(transpose ? src[kd * ld + m] : src[m * ld + kd]) : 0.f;
      }
      v = make_float4(x[0], x[1], x[2], x[3]);
    }
    *reinterpret_cast<float4*>(store + tile * (kDTileBytes / 4) + (long long)j * 512 + t * 4) = v;
  }
}

// tile store -> row-major float32 (bit-exact round trip for the tests)
__global__ void dense_unpack_kernel(const float* __restrict__ store, long long M, long long Kd, long long n_kc,
                                    float* __restrict__ dst, long long ld, long long n_units) {
  for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < n_units;
       u += (long long)gridDim.x * blockDim.x) {
    const long long tile = u >> 10;
    const int r = (int)(u & 1023), t = r >> 3, j = r & 7;
    const long long mp = tile / n_kc, kc = tile % n_kc;
    const long long m = mp * 128 + t, kd0 = kc * 32 + 4 * j;
    if (m >= M) continue;
    const float4 v = *reinterpret_cast<const float4*>(store + tile * (kDTileBytes / 4) + (long long)j * 512 + t * 4);
    const float x[4] = {v.x, v.y, v.z, v.w};
    for (int e = 0; e < 4; ++e)
      if (kd0 + e < Kd) dst[m * ld + kd0 + e] = x[e];
  }
}

// Operand B [Kd, K] (row-major, ld) -> the two TF32 images (hi, lo) in the tensor core's K-major no-swizzle order:
// K-block kb (8 contraction elements), half h (4 elements = 16 bytes), column n:  byte (kb * 2 + h) * 16 N + n * 16.
// Rows beyond kd_limit and columns beyond K are zero.  One thread per (kb, h, n): reads 4 rows of B at column n
// (coalesced along n), writes 16 bytes to each image.
template <typename T>
__global__ void dense_image_kernel(const T* __restrict__ B, long long ldb, long long kd_limit, int K, int N,
                                   float* __restrict__ img_hi, float* __restrict__ img_lo, long long n_units) {
  for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < n_units;
       u += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(u % N);
    const long long g = u / N;         // group of 4 contraction elements
    float h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const long long kd = g * 4 + e;
      // float64 operands: hi + lo carries 21 of the 53 mantissa bits; the float32 rounding of b happens first
      const float b = (n < K && kd < kd_limit) ? (float)B[kd * ldb + n] : 0.f;
      const float bh = __uint_as_float((__float_as_uint(b) + 0x1000u) & 0xFFFFE000u);   // round to nearest (see producers)
      h[e] = bh;
      l[e] = b - bh;
    }
    *reinterpret_cast<float4*>(img_hi + u * 4) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4*>(img_lo + u * 4) = make_float4(l[0], l[1], l[2], l[3]);
  }
}

}  // namespace lmdec
