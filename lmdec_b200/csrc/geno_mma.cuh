// The hot product of the decomposition path: a 2-bit genotype tile store times an int8-plane operand on the
// sm_100a tensor cores (tcgen05.mma kind::i8, accumulators and the decoded A operand in TMEM).
//
//   out[m, k] = sum_kd value(m,kd) * q1[kd,k]  (+ miss(m,kd) * q2[kd,k])          exact, int64
//
// Replaces the chunked int8->f64 matmuls under PowerMethod._solution_step (reference
// lmdec/decomp/iter_methods.py:381-386 -> lmdec/array/stacked.py:172-217, chained.py:169-177).
//
// CTA = 11 warps, one CTA per SM, persistent over work items (m-tile, k-split):
//   warps 0-7  producers: coalesced LDG of the packed tile (1 row per thread), 2-bit -> uint8 decode in registers,
//              tcgen05.st straight into the TMEM A ring; afterwards the epilogue (TMEM -> smem -> int64 combine)
//   warp  8    B loader: one thread, cp.async.bulk of the pre-laid-out plane image into a smem ring
//   warps 9,10 MMA issuers: one thread each.  A narrow-N tcgen05.mma is limited by the issuing thread (~45-54 cycles
//              per instruction, profiles/r01/umma_probe4_issuers.log); two issuers reach the N/2-cycle math rate.
//              Without missing data they take alternate K-blocks; with missing data issuer 0 runs the value plane and
//              issuer 1 the indicator plane.  Each owns one accumulator.
#pragma once
#include "umma.cuh"

namespace lmdec {

constexpr int kProducerWarps = 8;
constexpr int kLoaderWarp = 8;
constexpr int kIssuerWarp0 = 9;
constexpr int kThreads = 11 * 32;
constexpr int kKbPerChunk = 8;             // 8 K-blocks of 32 = one 256-wide contraction chunk = one store tile
constexpr uint32_t kAccCols = 128;         // accumulator i at TMEM column i*128
constexpr uint32_t kARingCol0 = 256;       // A ring: columns [256, 512)
constexpr int kMaxBStages = 4;

struct GenoMmaParams {
  const uint8_t* store;
  long long store_kc;      // chunks per row panel in the store (tile stride)
  long long m_limit;       // rows produced
  int m_tiles;
  int n_chunks;            // contraction chunks to run: ceil(kd_limit / 256)
  int splits;
  int has_missing;
  int mode;                // 0: out0 = acc0 + acc1 ; 1: out0 = acc0, out1 = acc1
  int n_img;               // B images staged per chunk (1 or 2)
  const int8_t* image1;
  const int8_t* image2;
  int K, P, N;
  long long* out0;
  long long* out1;
  int b_stages;
  int use_atomics;
};

__device__ __forceinline__ uint4 ldg_stream16(const uint8_t* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// 16 codes in r -> 4 words of 4 uint8 each; word k holds codes k, k+4, k+8, k+12 (the B image is permuted to match)
template <bool kMiss>
__device__ __forceinline__ void decode16(uint32_t r, uint32_t* v, uint32_t* m) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t x = (r >> (2 * k)) & 0x03030303u;
    if (kMiss) {
      const uint32_t both = x & (x >> 1) & 0x01010101u;  // code 3 = missing
      v[k] = x ^ (both * 3u);
      m[k] = both;
    } else {
      v[k] = x;
    }
  }
}

template <bool kMiss>
__device__ __forceinline__ void produce_chunk(const uint4 (&raw)[2], uint32_t a_stage_addr, int half) {
  // this thread's 32 bytes = K-blocks 4*half .. 4*half+3 of the chunk; raw[i] = 64 codes = 2 K-blocks
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const uint32_t regs[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      uint32_t v[8], m[8];
      decode16<kMiss>(regs[2 * j], v, m);
      decode16<kMiss>(regs[2 * j + 1], v + 4, m + 4);
      const int kb = 4 * half + 2 * i + j;
      tmem_st8(a_stage_addr + kb * 8, v);
      if (kMiss) tmem_st8(a_stage_addr + 64 + kb * 8, m);
    }
  }
}

__global__ void __launch_bounds__(kThreads, 1) geno_mma_kernel(const GenoMmaParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full_a[4], empty_a[4], full_b[kMaxBStages], empty_b[kMaxBStages], acc_full, acc_empty;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.N;
  const uint32_t img_bytes = (uint32_t)kKbPerChunk * N * 32;
  const uint32_t stage_bytes = img_bytes * p.n_img;
  uint8_t* smem_b = smem;
  int* stage_out = reinterpret_cast<int*>(smem + (size_t)p.b_stages * stage_bytes);  // [128][N+1] int32
  const int a_stages = p.has_missing ? 2 : 4;
  const uint32_t a_stage_cols = p.has_missing ? 128 : 64;

  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  if (tid == 32) {
    for (int i = 0; i < 4; ++i) {
      mbar_init(&full_a[i], kProducerWarps);
      mbar_init(&empty_a[i], 2);
    }
    for (int i = 0; i < kMaxBStages; ++i) {
      mbar_init(&full_b[i], 1);
      mbar_init(&empty_b[i], 2);
    }
    mbar_init(&acc_full, 2);
    mbar_init(&acc_empty, kProducerWarps);
    mbar_fence_init();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tbase = tmem_base_s;

  const int n_work = p.m_tiles * p.splits;
  const int chunks_per_split = (p.n_chunks + p.splits - 1) / p.splits;

  if (warp < kProducerWarps) {
    // ------------------------------------------------------------------ producers + epilogue
    const int q = warp & 3, half = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tbase + ((uint32_t)(q * 32) << 16);
    uint32_t a_stage = 0, a_phase = 0, acc_phase = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
      const int mt = w % p.m_tiles, sp = w / p.m_tiles;
      const int c0 = sp * chunks_per_split;
      const int c1 = min(c0 + chunks_per_split, p.n_chunks);
      const uint8_t* tile0 = p.store + ((size_t)mt * p.store_kc) * 8192 + (size_t)(2 * half) * 2048 + (size_t)row * 16;
      uint4 raw[2], nxt[2];
      if (c0 < c1) {
        raw[0] = ldg_stream16(tile0 + (size_t)c0 * 8192);
        raw[1] = ldg_stream16(tile0 + (size_t)c0 * 8192 + 2048);
      }
      for (int c = c0; c < c1; ++c) {
        if (c + 1 < c1) {
          nxt[0] = ldg_stream16(tile0 + (size_t)(c + 1) * 8192);
          nxt[1] = ldg_stream16(tile0 + (size_t)(c + 1) * 8192 + 2048);
        }
        mbar_wait(&empty_a[a_stage], a_phase ^ 1);
        tc_fence_after_sync();
        const uint32_t a_addr = lane_addr + kARingCol0 + a_stage * a_stage_cols;
        if (p.has_missing)
          produce_chunk<true>(raw, a_addr, half);
        else
          produce_chunk<false>(raw, a_addr, half);
        tmem_wait_st();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_a[a_stage]);
        if (++a_stage == (uint32_t)a_stages) {
          a_stage = 0;
          a_phase ^= 1;
        }
        raw[0] = nxt[0];
        raw[1] = nxt[1];
      }
      // ---- epilogue: TMEM -> smem staging -> exact int64 plane combine -> global
      mbar_wait(&acc_full, acc_phase);
      acc_phase ^= 1;
      tc_fence_after_sync();
      const int ld_s = N + 1;
      const int n_pass = (p.mode == 1) ? 2 : 1;
      for (int pass = 0; pass < n_pass; ++pass) {
        // columns [half*N/2, (half+1)*N/2) of this lane quarter, 8 at a time (N is a multiple of 16)
        for (int cb = half * (N / 2); cb < (half + 1) * (N / 2); cb += 8) {
          uint32_t r0[8], r1[8];
          if (p.mode == 1) {
            tmem_ld8(lane_addr + pass * kAccCols + cb, r0);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 8; ++i) stage_out[row * ld_s + cb + i] = (int)r0[i];
          } else {
            tmem_ld8(lane_addr + cb, r0);
            tmem_ld8(lane_addr + kAccCols + cb, r1);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 8; ++i) stage_out[row * ld_s + cb + i] = (int)r0[i] + (int)r1[i];
          }
        }
        if (pass == n_pass - 1) {
          // all TMEM reads of this tile are done: hand the accumulators back to the issuers
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty);
        }
        named_bar_sync(1, kProducerWarps * 32);
        long long* out = pass == 0 ? p.out0 : p.out1;
        const int K = p.K, P = p.P;
        const long long m0 = (long long)mt * 128;
        for (int idx = tid; idx < 128 * K; idx += kProducerWarps * 32) {
          const int r = idx / K, k = idx - r * K;
          if (m0 + r < p.m_limit) {
            long long acc = 0;
            for (int pl = P - 1; pl >= 0; --pl) acc = acc * 256 + (long long)stage_out[r * ld_s + pl * K + k];
            long long* dst = out + (m0 + r) * K + k;
            if (p.use_atomics)
              atomicAdd(reinterpret_cast<unsigned long long*>(dst), (unsigned long long)acc);
            else
              *dst = acc;
          }
        }
        named_bar_sync(1, kProducerWarps * 32);
      }
    }
  } else if (warp == kLoaderWarp) {
    // ------------------------------------------------------------------ B image loader
    if (lane == 0) {
      uint32_t b_stage = 0, b_phase = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        const int sp = w / p.m_tiles;
        const int c0 = sp * chunks_per_split;
        const int c1 = min(c0 + chunks_per_split, p.n_chunks);
        for (int c = c0; c < c1; ++c) {
          mbar_wait(&empty_b[b_stage], b_phase ^ 1);
          uint8_t* dst = smem_b + (size_t)b_stage * stage_bytes;
          mbar_arrive_expect_tx(&full_b[b_stage], stage_bytes);
          bulk_g2s(dst, p.image1 + (size_t)c * img_bytes, img_bytes, &full_b[b_stage]);
          if (p.n_img == 2) bulk_g2s(dst + img_bytes, p.image2 + (size_t)c * img_bytes, img_bytes, &full_b[b_stage]);
          if (++b_stage == (uint32_t)p.b_stages) {
            b_stage = 0;
            b_phase ^= 1;
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ MMA issuers
    if (lane == 0) {
      const int me = warp - kIssuerWarp0;
      const uint32_t idesc = make_idesc_i8(128, (uint32_t)N, 0, 1);
      const uint32_t lbo = (uint32_t)N * 16;
      const uint32_t d_addr = tbase + me * kAccCols;
      const uint32_t plane = p.has_missing ? (uint32_t)me : 0u;
      const uint32_t img = (p.has_missing && p.n_img == 2) ? (uint32_t)me : 0u;
      uint32_t a_stage = 0, a_phase = 0, b_stage = 0, b_phase = 0, acc_phase = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        const int sp = w / p.m_tiles;
        const int c0 = sp * chunks_per_split;
        const int c1 = min(c0 + chunks_per_split, p.n_chunks);
        mbar_wait(&acc_empty, acc_phase ^ 1);
        acc_phase ^= 1;
        tc_fence_after_sync();
        uint32_t accumulate = 0;
        for (int c = c0; c < c1; ++c) {
          mbar_wait(&full_a[a_stage], a_phase);
          mbar_wait(&full_b[b_stage], b_phase);
          tc_fence_after_sync();
          const uint32_t a_addr = tbase + kARingCol0 + a_stage * a_stage_cols + plane * 64;
          const uint32_t b_addr = smem_u32(smem_b + (size_t)b_stage * stage_bytes + (size_t)img * img_bytes);
#pragma unroll
          for (int j = 0; j < kKbPerChunk; ++j) {
            if (!p.has_missing && (j & 1) != me) continue;
            const uint64_t bdesc = make_smem_desc(b_addr + (uint32_t)j * N * 32, lbo, 128);
            umma_ts_i8(d_addr, a_addr + j * 8, bdesc, idesc, accumulate);
            accumulate = 1;
          }
          umma_commit(&empty_a[a_stage]);
          umma_commit(&empty_b[b_stage]);
          if (++a_stage == (uint32_t)a_stages) {
            a_stage = 0;
            a_phase ^= 1;
          }
          if (++b_stage == (uint32_t)p.b_stages) {
            b_stage = 0;
            b_phase ^= 1;
          }
        }
        if (accumulate == 0) {
          // (cannot happen: every work item has >= 1 chunk) keep the protocol alive anyway
        }
        umma_commit(&acc_full);
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tbase);
}

}  // namespace lmdec
