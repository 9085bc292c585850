// The hot product of the decomposition path: a 2-bit genotype tile store times an int8-plane operand on the
// sm_100a tensor cores (tcgen05.mma kind::i8, accumulators and the decoded A operand in TMEM).
//
//   out[m, k] = sum_kd value(m,kd) * q1[kd,k]  (+ miss(m,kd) * q2[kd,k])          exact, int64
// (the value plane carries raw codes, 3 at a missing entry; image 2 holds q2 - 3 q1 to compensate, see lmdec_b200.h)
//
// Replaces the chunked int8->f64 matmuls under PowerMethod._solution_step (reference
// lmdec/decomp/iter_methods.py:381-386 -> lmdec/array/stacked.py:172-217, chained.py:169-177).
//
// CTA = 22 warps, one CTA per SM, persistent over work items (m-tile or m-tile pair, k-split).  Every role streams
// across work-item boundaries, so HBM latency is paid once per launch, not once per tile:
//   warps 0-7   producers: read their row of the packed tile from the smem ring (conflict-free 16-byte pieces), decode
//               2-bit -> uint8 in registers, tcgen05.st straight into the TMEM A ring
//   warp  8     tile loader: one thread, cp.async.bulk (TMA engine) of whole 8 KB store tiles into an 8-deep smem ring
//   warps 9-12  MMA issuers: one elected lane each.  A narrow-N tcgen05.mma is limited by the issuing thread (~45-54
//               cycles per instruction, profiles/r01/umma_probe4_issuers.log), so the MMAs of a chunk are spread over
//               several issuers, each owning one accumulator: an issuer pair = (value plane, indicator plane) with
//               missing data or (even, odd K-blocks) without; pair mode M: 4 issuers = (tile, plane); single tiles: two
//               issuer pairs on alternate chunks (the serial per-chunk barrier round trips of an issuing thread, not
//               the tensor pipe, set the period: profiles/r01/pipeline_trace.txt).
//   warps 13-20 epilogue (two per TMEM lane quarter, splitting the column groups): TMEM -> registers -> exact plane
//               combine -> global (finalized float, int64 store, or int64 atomicAdd for split-K).  Short contractions
//               alternate two accumulator sets, so the epilogue of tile i overlaps the MMAs of tile i+1.  For A^T q the
//               epilogue also reduces the column statistics of its output for the A t that follows (kOutFusedStats).
//   warp  21    B loader: one thread, cp.async.bulk of the pre-laid-out plane image into its own smem ring
// Pair mode (long contractions): two m-tiles are processed against every staged chunk of the operand image, halving the
// image's L2 -> SM traffic (it is re-read by every m-tile and dominates the product's on-chip traffic).
#pragma once
#include "umma.cuh"

namespace lmdec {

// 8 producer warps in kGroups groups (kernel template parameter):
//   1 group : warps 0-3 decode K-blocks 0-3 and warps 4-7 K-blocks 4-7 of EVERY tile-chunk (32 bytes of a row per thread)
//   2 groups: warps 0-3 and warps 4-7 take ALTERNATE tile-chunks, all 8 K-blocks each (64 bytes of a row per thread).
// The per-iteration latency chain of a producer (barrier waits, fences, arrives: ~500 cycles,
// profiles/r01/pipeline_trace.txt) is then paid once per 64 bytes instead of once per 32.  Two groups win without
// missing data (deep A ring: 61,000 x 800,000 K=30 -7 %); with missing data the ring has two stages, each group would
// own one and serialise with its MMAs (+17 % on A t), so one group is used there.  (A run-time switch inside one
// kernel cost 6 %: the producer loop is that sensitive to extra instructions.)
#ifndef LMDEC_PRODUCER_WARPS
#define LMDEC_PRODUCER_WARPS 8
#endif
constexpr int kProducerWarps = LMDEC_PRODUCER_WARPS;          // 8, or 16 (more decode warps in flight per scheduler)
#ifndef LMDEC_ISSUER_WARPS
#define LMDEC_ISSUER_WARPS 4
#endif
constexpr int kIssuerWarps = LMDEC_ISSUER_WARPS;              // 2, or 4 (pair mode M: one issuer per (tile, plane))
#ifndef LMDEC_EPILOGUE_WARPS
#define LMDEC_EPILOGUE_WARPS 8
#endif
constexpr int kEpilogueWarps = LMDEC_EPILOGUE_WARPS;          // 4, or 8 (two warps per quarter split the column groups)
// Warp ids by role.  The warp scheduler of an SM sub-partition prefers the HIGHEST warp id among its eligible warps
// (B300_MICROARCH.md, "multi-warp arbiter"), so the latency-critical single-thread roles (loaders, MMA issuers) get the
// highest ids, the decode producers the middle ones, and the epilogue warps -- bulk work that only has to finish within
// the next tile's MMAs, and that polls a barrier most of the time -- the lowest.  (LMDEC_ROLE_ORDER=0: the round-1
// order, producers lowest.)  Producer and epilogue ranges start at multiples of 4: warp % 4 is the TMEM lane quarter.
#ifndef LMDEC_ROLE_ORDER
#define LMDEC_ROLE_ORDER 1
#endif
#if LMDEC_ROLE_ORDER == 1
constexpr int kEpilogueWarp0 = 0;
constexpr int kProducerWarp0 = kEpilogueWarps;
constexpr int kTileLoaderWarp = kProducerWarp0 + kProducerWarps;
constexpr int kIssuerWarp0 = kTileLoaderWarp + 1;
constexpr int kBLoaderWarp = kIssuerWarp0 + kIssuerWarps;
#else
constexpr int kProducerWarp0 = 0;
constexpr int kTileLoaderWarp = kProducerWarps;
constexpr int kIssuerWarp0 = kProducerWarps + 1;
constexpr int kEpilogueWarp0 = kProducerWarps + 1 + kIssuerWarps;  // consecutive warps cover the 4 TMEM lane quarters
constexpr int kBLoaderWarp = kEpilogueWarp0 + kEpilogueWarps;
#endif
static_assert(kProducerWarp0 % 4 == 0 && kEpilogueWarps % 4 == 0, "TMEM lane quarters");  // (epilogue: any 4 consecutive warps)
// LMDEC_GENO_REGS = 1: setmaxnreg moves registers from the single-thread warps (loaders, issuers) to the epilogue warps,
// whose finalize variants otherwise spill; needs whole warpgroups (the block is padded to a multiple of 4 warps)
#ifndef LMDEC_GENO_REGS
#define LMDEC_GENO_REGS 0
#endif
#if LMDEC_GENO_REGS
static_assert(LMDEC_ROLE_ORDER == 1 && kEpilogueWarps == 8 && kProducerWarps == 8 && kIssuerWarps == 4, "warpgroup layout");
constexpr int kThreads = ((kProducerWarps + kEpilogueWarps + kIssuerWarps + 2 + 3) / 4) * 4 * 32;
#else
constexpr int kThreads = (kProducerWarps + kEpilogueWarps + kIssuerWarps + 2) * 32;
#endif
constexpr int kKbPerChunk = 8;             // 8 K-blocks of 32 = one 256-wide contraction chunk = one store tile
// TMEM: accumulators from column 0, the decoded-A ring behind them up to column 512 (p.a_ring_col0, p.a_stages)
constexpr int kMaxBStages = 4;
#ifndef LMDEC_TILE_STAGES
#define LMDEC_TILE_STAGES 8
#endif
constexpr int kMaxTileStages = LMDEC_TILE_STAGES;   // packed tiles in flight per SM (8 KB each)
constexpr uint32_t kTileBytes = 8192;
// mbarrier slots (8 bytes each) in one shared array, addressed by 32-bit shared-window addresses
constexpr int kMaxAStages = 8;
constexpr int kBarFullT = 0, kBarEmptyT = kMaxTileStages, kBarFullA = 2 * kMaxTileStages, kBarEmptyA = kBarFullA + 8,
              kBarFullB = kBarEmptyA + 8, kBarEmptyB = kBarFullB + 4, kBarAccFull = kBarEmptyB + 4,
              kBarAccEmpty = kBarAccFull + 2, kBarFirstDone = kBarAccEmpty + 2, kNumBars = kBarFirstDone + 4;

// out[m,k] = (acc0 + rowa[m]*acc1) * colscale[k] * rowscale[m] + rowshift[m] * colshift[k]   (NULL = absent / 1)
struct FusedFinalize {
  const double* colscale;
  const double* colshift;
  const double* rowa;
  const double* rowscale;
  const double* rowshift;
  void* out;      // float64, or float32 when out_f32
  long long ldo;
  int out_f32;
  // optional (K <= 32): column statistics of the OUTPUT taken as the operand of the following A.dot -- per-CTA partials
  // [gridDim.x][3][K] = max|out rowscale|, max|out rowscale rowa|, sum rowshift out (the layout colstats_partial_kernel
  // writes); the last CTA to finish folds them into stats_final (multiplier for stats_lim, its inverse, shift, maxima:
  // what colstats_finish_kernel writes), so the next product needs no pass over its operand and no finish launch
  double* stats_partial;
  double* stats_final;
  double stats_lim;
  unsigned int* stats_ticket;   // device counter, zero between launches
  // optional (float32 output, staged epilogue): NVSwitch multicast alias of `out` in a symmetric allocation shared by the
  // ranks of a row-sharded run.  The epilogue then ADDS its tile into every rank's replica (multimem.red) instead of
  // storing it: the all-reduce of t = A^T q happens tile by tile while the tensor cores work on the next tile, and the
  // separate NCCL all-reduce behind the kernel disappears (out must be zero on every rank before the launch)
  float* mc_out;
};

#ifdef LMDEC_TRACE
// debug timeline: CTA 0 records clock64() stamps of its first kTraceChunks chunks (tools/trace_mma.py prints them)
constexpr int kTraceChunks = 64;
__device__ long long g_trace[12][kTraceChunks];
#define LMDEC_STAMP(row, idx) do { if (blockIdx.x == 0 && (idx) < kTraceChunks) g_trace[row][idx] = clock64(); } while (0)
// cycles a role spends inside its barrier waits, summed over the kernel (CTA 0, one warp per role; tools/trace_mma.py):
// the role that never waits is the one that paces the pipeline
__device__ long long g_wait[16];
#define LMDEC_WAITACC(slot, cond, stmt) do { if (blockIdx.x == 0 && (cond)) { long long t0_ = clock64(); stmt; \
    wacc_[slot] += clock64() - t0_; } else { stmt; } } while (0)
#define LMDEC_WAITDECL long long wacc_[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}
#define LMDEC_WAITOUT(slot, cond) do { if (blockIdx.x == 0 && (cond)) g_wait[slot] = wacc_[slot]; } while (0)
#else
#define LMDEC_STAMP(row, idx) do { } while (0)
#define LMDEC_WAITACC(slot, cond, stmt) do { stmt; } while (0)
#define LMDEC_WAITDECL do { } while (0)
#define LMDEC_WAITOUT(slot, cond) do { } while (0)
#endif

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
  int tile_stages;
  int use_atomics;
  int acc_sets;            // accumulator sets alternated between consecutive work items (1 or 2; 1 in pair mode)
  int a_ring_col0;         // first TMEM column of the decoded-A ring (= columns taken by the accumulators)
  int a_stages;            // stages of the decoded-A ring: (512 - a_ring_col0) / (128 with missing data, else 64)
  int pair;                // 0: one m-tile per work item; 1/2/3: TWO m-tiles share every B stage (kPairM/E/T)
  int m_units;             // work items per split: m_tiles, or ceil(m_tiles / 2) in pair mode
  int stats_off;           // byte offset of the statistics accumulators in dynamic shared memory (fuse == 2)
  int stats_cols;          // column slots per epilogue thread
  int rr;                  // no missing data: 4 issuers take whole chunks round-robin, ONE accumulator per tile (issuer_loop_rr)
  int one_acc;             // one accumulator per tile (pair mode T, rr): the epilogue reads acc0 only
  int b_resident;          // the whole operand image stays in shared memory (short, unsplit contractions): no B ring
  int epi_staged;          // float32 output: the staged epilogue (epilogue_loop32) has its staging area at stats_off
  int issuer_groups;       // 2: chunk-interleaved issuer pairs (single-tile work items)
  int fuse;                // write the finalized float64 result from the epilogue (never with atomics)
  FusedFinalize fz;
};

// 16 codes in r -> 4 words of 4 uint8 each; word k holds codes k, k+4, k+8, k+12 (the B image is permuted to match).
// The decode is bound by the ALU pipe (LOP3/SHF issue every other cycle per scheduler), so it is written with LEFT
// shifts only -- ptxas places those on the FMA pipe as IMAD.SHL -- and leaves each code in the TOP two bits of its byte:
//   value byte  = 64 * code       (code 3 = missing is NOT cleared: the epilogue / the second image subtract 3 * miss)
//   miss  byte  = 128 * [code==3]
// The accumulators are therefore exact multiples of 64 / 128 and are shifted back in the epilogue.
constexpr int kValueShift = 6, kMissShift = 7;
template <bool kMiss>
__device__ __forceinline__ void decode16(uint32_t r, uint32_t* v, uint32_t* m) {
#ifdef LMDEC_ABL_NODECODE   // ablation (timing only, wrong results): the producers skip the 2-bit -> byte arithmetic
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    v[k] = r;
    if (kMiss) m[k] = r;
  }
  return;
#endif
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t x = (k == 3 ? r : r * (1u << (6 - 2 * k))) & 0xC0C0C0C0u;
    v[k] = x;
    if (kMiss) m[k] = x & (x * 2u) & 0x80808080u;
  }
}

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

// Width of the tcgen05.st that moves a thread's decoded bytes into the TMEM operand ring: 8, 16 or 32 columns (32-bit
// registers) per instruction.  The producers' store phase (~390 cycles per chunk for 64 KB: profiles/r02) is what paces
// the kernel; fewer, wider stores move the same bytes with less per-instruction overhead.
#ifndef LMDEC_STTM_W
#define LMDEC_STTM_W 16
#endif
template <bool kMiss, int KB>
__device__ __forceinline__ void produce_chunk(const uint4 (&raw)[KB / 2], uint32_t a_stage_addr, int kb0) {
  // this thread's bytes = K-blocks kb0 .. kb0 + KB of the chunk; raw[i] = 64 codes = 2 K-blocks
#if LMDEC_STTM_W == 32
#pragma unroll
  for (int h = 0; h < KB / 4; ++h) {       // 4 K-blocks = 32 columns per plane per store
    uint32_t v[32], m[32];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const uint4 rw = raw[2 * h + i];
      const uint32_t regs[4] = {rw.x, rw.y, rw.z, rw.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) decode16<kMiss>(regs[j], v + 16 * i + 4 * j, m + 16 * i + 4 * j);
    }
    tmem_st32(a_stage_addr + (kb0 + 4 * h) * 8, v);
    if (kMiss) tmem_st32(a_stage_addr + 64 + (kb0 + 4 * h) * 8, m);
  }
#elif LMDEC_STTM_W == 16
#pragma unroll
  for (int i = 0; i < KB / 2; ++i) {       // 2 K-blocks = 16 columns per plane per store
    const uint32_t regs[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
    uint32_t v[16], m[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) decode16<kMiss>(regs[j], v + 4 * j, m + 4 * j);
    tmem_st16(a_stage_addr + (kb0 + 2 * i) * 8, v);
    if (kMiss) tmem_st16(a_stage_addr + 64 + (kb0 + 2 * i) * 8, m);
  }
#else
#pragma unroll
  for (int i = 0; i < KB / 2; ++i) {
    const uint32_t regs[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      uint32_t v[8], m[8];
      decode16<kMiss>(regs[2 * j], v, m);
      decode16<kMiss>(regs[2 * j + 1], v + 4, m + 4);
      const int kb = kb0 + 2 * i + j;
#ifdef LMDEC_ABL_NOSTTM     // ablation (timing only): no TMEM stores of the decoded operand
      if (v[0] == 0x12345678u && m[7] == 0x9abcdef1u) tmem_st8(a_stage_addr + kb * 8, v);
#else
      tmem_st8(a_stage_addr + kb * 8, v);
      if (kMiss) tmem_st8(a_stage_addr + 64 + kb * 8, m);
#endif
    }
  }
#endif
}

// Epilogue of one warp (one TMEM lane quarter, thread = output row).  Accumulator column n = k*P + plane, so the P
// planes of 8 consecutive outputs are 8*P consecutive columns: everything stays in registers, indices are compile-time.
// MODE 0: out0 = acc0 + acc1;  MODE 1: out0 = acc0 - 3 acc1 (value plane carried raw codes), out1 = acc1.
// OUT  0: int64 store, 1: int64 atomicAdd (split-K), 2: finalized float64 (fz_col = colscale[K] | colshift[K] in smem).
//      3: as 2, plus the column statistics of the finalized output (FusedFinalize::stats_partial).
//      4 / 5: as 2 / 3 with float32 output, the whole finalize on the fp32 pipe (the steady-state A^T q of the iteration).
enum { kOutStore = 0, kOutAtomic = 1, kOutFused = 2, kOutFusedStats = 3, kOutFused32 = 4, kOutFusedStats32 = 5 };
// Statistics accumulators (kOutFusedStats): every epilogue thread (= output row of a tile) keeps its own running
// max / sum per column in shared memory, [column slot][2][kEpilogueWarps * 32] doubles, folded once per kernel.
// kOutFusedStats32: [column slot][kStatThreads] float64 sums, then [column slot][kStatThreads] float32 maxima kept as
// their bit patterns (non-negative floats order like unsigned integers, and a NaN sorts above +inf, so it poisons
// the scale exactly like the `v != v` test of the float64 path).
constexpr int kStatThreads = kEpilogueWarps * 32;
// Pair modes.  The operand image is the dominant L2 -> SM traffic (every m-tile re-reads all of it); processing two
// m-tiles against one staged chunk halves it.  M: missing data, issuer i runs plane i of both tiles (4 accumulators of
// 64 columns); E: no missing data, issuer i takes K-blocks of parity i of both tiles (4 x 64); T: no missing data and
// N > 64, issuer i owns tile i (2 accumulators of 128 columns).
enum { kPairNone = 0, kPairM = 1, kPairE = 2, kPairT = 3 };

template <int P>
__device__ __forceinline__ double planes_to_double(const uint32_t* r) {
  double a = (double)(int)r[P - 1];   // planes are int32 and the combination is < 2^53: exact in float64
#pragma unroll
  for (int pl = P - 2; pl >= 0; --pl) a = fma(a, 256.0, (double)(int)r[pl]);
  return a;
}
// float32 plane combine (float32 outputs): each int32 plane rounds to 24 bits, i.e. an error of 2^-24 of the LARGEST
// plane term -- below the float32 rounding of the output itself and far below the 2^-22 operand quantisation.
template <int P>
__device__ __forceinline__ float planes_to_float(const uint32_t* r) {
  float a = (float)(int)r[P - 1];
#pragma unroll
  for (int pl = P - 2; pl >= 0; --pl) a = fmaf(a, 256.0f, (float)(int)r[pl]);
  return a;
}
template <int P>
__device__ __forceinline__ long long planes_to_int64(const uint32_t* r) {
  long long a = (long long)(int)r[P - 1];
#pragma unroll
  for (int pl = P - 2; pl >= 0; --pl) a = a * 256 + (long long)(int)r[pl];
  return a;
}

template <int P, int MODE, int OUT>
__device__ __forceinline__ void epilogue_loop(const GenoMmaParams& p, uint32_t tbase, uint32_t acc_full,
                                              uint32_t acc_empty, int warp, int lane, int acc_sets, uint32_t acc_w,
                                              const double* fz_col, double* stats_acc) {
  const float* fz_col32 = reinterpret_cast<const float*>(fz_col + 256);   // float32 copies behind the float64 ones
  const int q = warp & 3;  // TMEM lane quarter this warp may touch
  const int row = q * 32 + lane;
  const uint32_t lane_addr = tbase + ((uint32_t)(q * 32) << 16);
  const int K = p.K;
  const int n_groups = (K + 7) / 8;
  constexpr int kSplit = kEpilogueWarps / 4;                  // warps per lane quarter
  const int g_first = (warp - kEpilogueWarp0) / 4;            // this warp takes column groups g_first, g_first + kSplit, ..
  const int n_work = p.m_units * p.splits;
  const int shift1 = p.has_missing ? kMissShift : kValueShift;
  const int tiles_per_item = p.pair ? 2 : 1;
  // running column statistics (kOutFusedStats): this thread's own accumulators, no atomics, fixed order
  double* my_stats = stats_acc + (warp - kEpilogueWarp0) * 32 + lane;
  if (OUT == kOutFusedStats) {
    for (int i = 0; i < p.stats_cols * 2; ++i) my_stats[i * kStatThreads] = 0.0;
  }
  uint32_t* my_max32 = reinterpret_cast<uint32_t*>(stats_acc + (size_t)p.stats_cols * kStatThreads) +
                       (warp - kEpilogueWarp0) * 32 + lane;
  if (OUT == kOutFusedStats32) {
    for (int i = 0; i < p.stats_cols; ++i) {
      my_stats[i * kStatThreads] = 0.0;
      my_max32[i * kStatThreads] = 0u;
    }
  }
  uint32_t tile_no = 0;
  for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++tile_no) {
    const int unit = w % p.m_units;
    // accumulator hand-shake: per accumulator set when single tiles alternate between two sets, one for a pair
    const uint32_t set = p.pair ? 0u : tile_no % acc_sets;
    const uint32_t set_phase = p.pair ? (tile_no & 1) : ((tile_no / acc_sets) & 1);
    mbar_wait_relaxed_a(acc_full + 8 * set, set_phase);
    if (warp == kEpilogueWarp0 && lane == 0) LMDEC_STAMP(9, (int)tile_no);    // accumulators seen full
    tc_fence_after_sync();
    for (int t = 0; t < tiles_per_item; ++t) {
      const int mt = p.pair ? 2 * unit + t : unit;
      const uint32_t slot = p.pair ? (uint32_t)t : set;
      const uint32_t acc0 = lane_addr + (p.one_acc ? slot : slot * 2) * acc_w, acc1 = acc0 + acc_w;
      const long long m = (long long)mt * 128 + row;
      const bool live = mt < p.m_tiles && m < p.m_limit;
      long long* dst0 = p.out0 + m * K;
      long long* dst1 = p.out1 + m * K;
      double row_a = 0.0, row_scale = 1.0, row_shift = 1.0;
      double* fdst = nullptr;
      float* fdst32 = nullptr;
      if (OUT >= kOutFused && live) {
        if (p.fz.rowa) row_a = __ldg(p.fz.rowa + m);
        if (p.fz.rowscale) row_scale = __ldg(p.fz.rowscale + m);
        if (p.fz.rowshift) row_shift = __ldg(p.fz.rowshift + m);
        fdst = (double*)p.fz.out + m * p.fz.ldo;
        fdst32 = (float*)p.fz.out + m * p.fz.ldo;
      }
      const double row_a3 = (row_a - 3.0) * row_scale;  // MODE 1: out = rs*(acc0 - 3 acc1) + rs*ra*acc1
      const float rs32 = (float)row_scale, ra32 = (float)row_a3, rsh32 = (float)row_shift;
      if (g_first >= n_groups && t == tiles_per_item - 1) {
        // no column group for this warp (K <= 8 with two warps per quarter): it still takes part in the hand-back
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(acc_empty + 8 * set);
      }
      for (int g = g_first; g < n_groups; g += kSplit) {
        uint32_t r0[8 * P], r1[8 * P];
        const uint32_t col = (uint32_t)g * 8 * P;
#pragma unroll
        for (int i = 0; i < P; ++i) tmem_ld8p(acc0 + col + i * 8, r0 + 8 * i);
        if (p.one_acc) {
#pragma unroll
          for (int i = 0; i < 8 * P; ++i) r1[i] = 0;  // one accumulator per tile
        } else {
#pragma unroll
          for (int i = 0; i < P; ++i) tmem_ld8p(acc1 + col + i * 8, r1 + 8 * i);
        }
        tmem_wait_ld();
        if (g + kSplit >= n_groups && t == tiles_per_item - 1) {
          // this warp's TMEM reads of the work item are done: hand the accumulators back to the issuers
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive_a(acc_empty + 8 * set);
        }
        if (!live) continue;
#ifdef LMDEC_ABL_NOEPI      // ablation (timing only): TMEM loads and the hand-back, no finalize / stores
        if (r0[0] != 0x12345678u || r1[5] != 0x9abcdef0u) continue;
#endif
        // undo the decode's byte scaling (exact: the sums are multiples of 64 / 128); the float32 finalize folds the
        // two powers of two into its row factors instead
        if (OUT != kOutFused32 && OUT != kOutFusedStats32)
#pragma unroll
        for (int i = 0; i < 8 * P; ++i) {
          const int a0 = (int)r0[i] >> kValueShift;
          const int a1 = (int)r1[i] >> shift1;
          if (MODE == 1) {
            r0[i] = (uint32_t)(OUT >= kOutFused ? a0 : a0 - 3 * a1);
            r1[i] = (uint32_t)a1;
          } else {
            r0[i] = (uint32_t)(a0 + a1);
          }
        }
        const int k_hi = min(8, K - g * 8);
        if (OUT == kOutFusedStats32 || OUT == kOutFused32) {
          // float32 output: plane combine, finalize and the running maximum on the fp32 / integer pipes; only the
          // centring sum (sum_r w_r y_r, a sum of 1e6 terms with cancellation behind it) stays in float64
          const float sc32 = fabsf(rs32) * fmaxf(fabsf((float)row_a), 1.0f);
          const float c0 = rs32 * (1.0f / (1 << kValueShift));
          const float c1 = MODE == 1 ? ra32 * (1.0f / (1 << kMissShift)) : rs32 / (float)(1 << shift1);
          double* ssum = my_stats + (size_t)(((g - g_first) / kSplit) * 8) * kStatThreads;
          uint32_t* smax = my_max32 + (size_t)(((g - g_first) / kSplit) * 8) * kStatThreads;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            if (kk < k_hi) {
              const int k = g * 8 + kk;
              const float v = fmaf(planes_to_float<P>(r1 + kk * P), c1, planes_to_float<P>(r0 + kk * P) * c0);
              const float y32 = fmaf(v, fz_col32[k], rsh32 * fz_col32[128 + k]);
              fdst32[k] = y32;
              if (OUT == kOutFusedStats32) {
                const uint32_t bits = __float_as_uint(fabsf(y32) * sc32);   // |y| d max(1, mu) >= 0 (or NaN)
                if (bits > smax[kk * kStatThreads]) smax[kk * kStatThreads] = bits;
                ssum[kk * kStatThreads] = fma(row_shift, (double)y32, ssum[kk * kStatThreads]);
              }
            }
          }
        } else if (OUT == kOutFusedStats) {
          // what colstats_partial_kernel computes for the next A.dot: max(|y| d, |y| d mu) -- only the larger of its two
          // maxima sets the fixed-point scale, and |y| d max(1, mu) is bit for bit that larger one -- and sum w y
          const double abs_scale = fabs(row_scale), abs_a1 = fmax(fabs(row_a), 1.0);
          double* acc = my_stats + (size_t)(((g - g_first) / kSplit) * 8 * 2) * kStatThreads;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            if (kk < k_hi) {
              const int k = g * 8 + kk;
              double v = (double)planes_to_int64<P>(r0 + kk * P) * row_scale;
              if (MODE == 1) v = fma((double)planes_to_int64<P>(r1 + kk * P), row_a3, v);
              double y = fma(v, fz_col[k], row_shift * fz_col[128 + k]);
              if (p.fz.out_f32) {
                const float y32 = (float)y;   // the statistics describe the value the next product will read
                fdst32[k] = y32;
                y = (double)y32;
              } else {
                fdst[k] = y;
              }
              double* a = acc + kk * 2 * kStatThreads;
              double v1 = fabs(y) * abs_scale;
              if (v1 != v1) v1 = INFINITY;   // NaN poisons the scale, as in the stand-alone pass
              v1 *= abs_a1;
              if (v1 > a[0]) a[0] = v1;
              a[kStatThreads] = fma(row_shift, y, a[kStatThreads]);
            }
          }
        } else {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          if (kk < k_hi) {
            const int k = g * 8 + kk;
            if (OUT == kOutFused) {
              // integer plane combine (ALU pipe) and ONE int64 -> float64 conversion per accumulator: the fp64 /
              // conversion pipe is the scarce one in this epilogue (stall_math in profiles/r01)
              if (p.fz.out_f32) {
                // float32 output: the whole finalize runs on the fp32 pipe (one rounding more than the store itself)
                float v = (float)planes_to_int64<P>(r0 + kk * P) * rs32;
                if (MODE == 1) v = fmaf((float)planes_to_int64<P>(r1 + kk * P), ra32, v);
                fdst32[k] = fmaf(v, (float)fz_col[k], rsh32 * (float)fz_col[128 + k]);
              } else {
                double v = (double)planes_to_int64<P>(r0 + kk * P) * row_scale;
                if (MODE == 1) v = fma((double)planes_to_int64<P>(r1 + kk * P), row_a3, v);
                fdst[k] = fma(v, fz_col[k], row_shift * fz_col[128 + k]);
              }
            } else {
              const long long acc = planes_to_int64<P>(r0 + kk * P);
              if (OUT == kOutAtomic)
                atomicAdd(reinterpret_cast<unsigned long long*>(dst0 + k), (unsigned long long)acc);
              else
                dst0[k] = acc;
              if (MODE == 1) {
                const long long am = planes_to_int64<P>(r1 + kk * P);
                if (OUT == kOutAtomic)
                  atomicAdd(reinterpret_cast<unsigned long long*>(dst1 + k), (unsigned long long)am);
                else
                  dst1[k] = am;
              }
            }
          }
        }
        }
      }
    }
    if (warp == kEpilogueWarp0 && lane == 0) LMDEC_STAMP(10, (int)tile_no);   // tile finalized
  }
  if (OUT == kOutFusedStats || OUT == kOutFusedStats32) {
    // fold the rows in a fixed order: one partial row per CTA, the layout colstats_finish_kernel reads
    asm volatile("bar.sync 1, %0;" ::"n"(kStatThreads) : "memory");
    const int k = (warp - kEpilogueWarp0) * 32 + lane;
    if (k < K) {
      const int g = k >> 3;
      double m1 = 0.0, ws = 0.0;
      if (OUT == kOutFusedStats32) {
        const size_t off = (size_t)((g / kSplit) * 8 + (k & 7)) * kStatThreads + (g % kSplit) * 128;
        const double* bsum = stats_acc + off;
        const uint32_t* bmax = reinterpret_cast<const uint32_t*>(stats_acc + (size_t)p.stats_cols * kStatThreads) + off;
        uint32_t mx = 0u;
        for (int r = 0; r < 128; ++r) {
          mx = max(mx, bmax[r]);
          ws += bsum[r];
        }
        m1 = mx > 0x7f800000u ? (double)INFINITY : (double)__uint_as_float(mx);   // NaN sorts above +inf
      } else {
        const double* base = stats_acc + (size_t)(((g / kSplit) * 8 + (k & 7)) * 2) * kStatThreads + (g % kSplit) * 128;
        for (int r = 0; r < 128; ++r) {
          m1 = fmax(m1, base[r]);
          ws += base[kStatThreads + r];
        }
      }
      double* out = p.fz.stats_partial + (size_t)blockIdx.x * 3 * K;
      out[k] = m1;        // the larger of the two maxima of the stand-alone pass
      out[K + k] = 0.0;
      out[2 * K + k] = ws;
    }
    // the last CTA to get here folds all partial rows (block order) into the final statistics
    __threadfence();
    asm volatile("bar.sync 1, %0;" ::"n"(kStatThreads) : "memory");
    if (k == 0) stats_acc[0] = atomicAdd(p.fz.stats_ticket, 1u) == gridDim.x - 1 ? 1.0 : 0.0;
    asm volatile("bar.sync 1, %0;" ::"n"(kStatThreads) : "memory");
    if (stats_acc[0] != 0.0) {
      __threadfence();
      // thread (slice, column) folds the rows slice, slice + n_sl, ...; thread `column` then folds the slices in order
      asm volatile("bar.sync 1, %0;" ::"n"(kStatThreads) : "memory");
      const int n_sl = kStatThreads / K, kc = k % K, sl = k / K;
      if (sl < n_sl) {
        double m1 = 0.0, m2 = 0.0, ws = 0.0;
        for (unsigned b = sl; b < gridDim.x; b += n_sl) {
          const double* in = p.fz.stats_partial + (size_t)b * 3 * K;
          m1 = fmax(m1, __ldcg(in + kc));
          m2 = fmax(m2, __ldcg(in + K + kc));
          ws += __ldcg(in + 2 * K + kc);
        }
        stats_acc[(sl * 3 + 0) * K + kc] = m1;
        stats_acc[(sl * 3 + 1) * K + kc] = m2;
        stats_acc[(sl * 3 + 2) * K + kc] = ws;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kStatThreads) : "memory");
      if (k < K) {
        double m1 = 0.0, m2 = 0.0, ws = 0.0;
        for (int j = 0; j < n_sl; ++j) {
          m1 = fmax(m1, stats_acc[(j * 3 + 0) * K + k]);
          m2 = fmax(m2, stats_acc[(j * 3 + 1) * K + k]);
          ws += stats_acc[(j * 3 + 2) * K + k];
        }
        const double s_ = fmax(m1, m2);
        const double mult = s_ > 0.0 ? p.fz.stats_lim / s_ : 1.0;   // as colstats_finish_kernel
        double* st = p.fz.stats_final;
        st[k] = mult;
        st[K + k] = 1.0 / mult;
        st[2 * K + k] = -ws;
        st[3 * K + k] = m1;
        st[4 * K + k] = m2;
      }
      if (k == 0) *p.fz.stats_ticket = 0u;
    }
  }
}

#ifndef LMDEC_EPI_F32
#define LMDEC_EPI_F32 2   // 0: float64 finalize for every output; 1: round-2 interim fp32 epilogue; 2: staged fp32 epilogue
#endif

// ---------------------------------------------------------------------------------------------------------------------
// Epilogue of the float32-output products (the steady-state A^T q of the iteration), round 2.
// Measured on the round-1 epilogue (profiles/r02/trace_round2.txt): 8,860 cycles per 128 x 20 tile -- it, not the MMAs
// (5,120 cycles), set the tile period -- spent on 4-byte stores scattered over 128 rows and on per-thread
// read-modify-write statistics in shared memory.  Now:
//   main phase   TMEM -> registers -> float32 finalize -> the tile's [rows][K] float32 image in one of two shared-memory
//                staging buffers; the accumulators go back to the issuers as soon as the last tcgen05.ld has landed;
//   post phase   the staged tile is contiguous in global memory (128 x K floats): ONE bulk copy (TMA engine) stores it;
//                the column statistics of the tile are reduced from shared memory into two registers per thread
//                (lane = column, warp w takes rows w, w + 8, ...), folded once at the end of the kernel.
// Staging area (ONE tile: the next tile's main phase starts a whole tile of MMAs later, so a second buffer would buy
// nothing and shared memory is better spent on the operand image / the tile ring):
//   y[128][K] f32 (at least 3 * kStatThreads f64: the last CTA's fold reuses it) | row_sc[128] f32 | row_w[128] f64 |
//   fold_mx[8][32] u32 | fold_sm[8][32] f64
__host__ __device__ constexpr size_t epi32_y_bytes(int K) {
  return (size_t)128 * K * 4 > 3 * (size_t)kStatThreads * 8 ? (size_t)128 * K * 4 : 3 * (size_t)kStatThreads * 8;
}
__host__ __device__ constexpr size_t epi32_stage_bytes(int K) {
  return epi32_y_bytes(K) + 128 * 4 + 128 * 8 + 8 * 32 * 4 + 8 * 32 * 8;
}
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, uint32_t smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_src), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void multimem_red_add_v4(float* mc_addr, const float4& v) {
  asm volatile("multimem.red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void stats_last_cta_fold(const GenoMmaParams& p, int k, int K, double* scratch);

template <int P, int MODE, bool kStats>
__device__ __forceinline__ void epilogue_loop32(const GenoMmaParams& p, uint32_t tbase, uint32_t acc_full,
                                                uint32_t acc_empty, int warp, int lane, int acc_sets, uint32_t acc_w,
                                                const double* fz_col, uint8_t* stage) {
  const float* fz_col32 = reinterpret_cast<const float*>(fz_col + 256);
  const int K = p.K;
  float* sy = reinterpret_cast<float*>(stage);
  float* row_sc = reinterpret_cast<float*>(stage + epi32_y_bytes(K));
  double* row_w = reinterpret_cast<double*>(stage + epi32_y_bytes(K) + 512);
  uint32_t* fold_mx = reinterpret_cast<uint32_t*>(stage + epi32_y_bytes(K) + 1536);
  double* fold_sm = reinterpret_cast<double*>(stage + epi32_y_bytes(K) + 2560);
  double* fold_scratch = reinterpret_cast<double*>(stage);   // (the staged tile is dead by then)
  const int ew = warp - kEpilogueWarp0;      // 0 .. kEpilogueWarps-1
  const int et = ew * 32 + lane;             // 0 .. kStatThreads-1
  const int q = warp & 3;
  const int row = q * 32 + lane;
  const uint32_t lane_addr = tbase + ((uint32_t)(q * 32) << 16);
  const int n_groups = (K + 7) / 8;
  constexpr int kSplit = kEpilogueWarps / 4;
  const int g_first = ew / 4;
  const int n_work = p.m_units * p.splits;
  const int shift1 = p.has_missing ? kMissShift : kValueShift;
  const int tiles_per_item = p.pair ? 2 : 1;
  const bool bulk_ok = p.fz.ldo == K && (reinterpret_cast<uintptr_t>(p.fz.out) & 15) == 0;
  LMDEC_WAITDECL;
  uint32_t mx = 0u;      // running column statistics of this thread's (column, row class): max as float bits, sum
  double sm = 0.0;
  uint32_t tile_no = 0;
  for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++tile_no) {
    const int unit = w % p.m_units;
    const uint32_t set = p.pair ? 0u : tile_no % acc_sets;
    const uint32_t set_phase = p.pair ? (tile_no & 1) : ((tile_no / acc_sets) & 1);
    LMDEC_WAITACC(0, ew == 0 && lane == 0, mbar_wait_relaxed_a(acc_full + 8 * set, set_phase));
    if (ew == 0 && lane == 0) LMDEC_STAMP(9, (int)tile_no);
    tc_fence_after_sync();
    for (int t = 0; t < tiles_per_item; ++t) {
      const int mt = p.pair ? 2 * unit + t : unit;
      const uint32_t slot = p.pair ? (uint32_t)t : set;
      const uint32_t acc0 = lane_addr + (p.one_acc ? slot : slot * 2) * acc_w, acc1 = acc0 + acc_w;
      const long long m0 = (long long)mt * 128;
      const long long m = m0 + row;
      const int live_rows = mt < p.m_tiles ? (int)min((long long)128, p.m_limit - m0) : 0;
      const bool live = row < live_rows;
      double row_a = 0.0, row_scale = 1.0, row_shift = 1.0;
      if (live) {
        if (p.fz.rowa) row_a = __ldg(p.fz.rowa + m);
        if (p.fz.rowscale) row_scale = __ldg(p.fz.rowscale + m);
        if (p.fz.rowshift) row_shift = __ldg(p.fz.rowshift + m);
      }
      const float rs32 = (float)row_scale, rsh32 = (float)row_shift;
      const float c0 = rs32 * (1.0f / (1 << kValueShift));   // the decode's byte scaling folded into the row factors
      const float c1 = MODE == 1 ? (float)((row_a - 3.0) * row_scale) * (1.0f / (1 << kMissShift))
                                 : rs32 / (float)(1 << shift1);
      if (g_first == 0) {   // one warp per lane quarter publishes the per-row factors of the statistics
        row_sc[row] = live ? fabsf(rs32) * fmaxf(fabsf((float)row_a), 1.0f) : 0.0f;
        row_w[row] = live ? row_shift : 0.0;
      }
      if (g_first >= n_groups && t == tiles_per_item - 1) {
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(acc_empty + 8 * set);
      }
      for (int g = g_first; g < n_groups; g += kSplit) {
        uint32_t r0[8 * P], r1[8 * P];
        const uint32_t col = (uint32_t)g * 8 * P;
#pragma unroll
        for (int i = 0; i < P; ++i) tmem_ld8p(acc0 + col + i * 8, r0 + 8 * i);
        if (p.one_acc) {
#pragma unroll
          for (int i = 0; i < 8 * P; ++i) r1[i] = 0;
        } else {
#pragma unroll
          for (int i = 0; i < P; ++i) tmem_ld8p(acc1 + col + i * 8, r1 + 8 * i);
        }
        tmem_wait_ld();
        if (g + kSplit >= n_groups && t == tiles_per_item - 1) {
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive_a(acc_empty + 8 * set);
        }
#ifdef LMDEC_ABL_NOEPI
        if (r0[0] != 0x12345678u || r1[5] != 0x9abcdef0u) continue;
#endif
        if (!live) continue;
        const int k_hi = min(8, K - g * 8);
        float* dst = sy + (size_t)row * K + g * 8;
        float y[8];
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const int k = g * 8 + kk;
          const float v = fmaf(planes_to_float<P>(r1 + kk * P), c1, planes_to_float<P>(r0 + kk * P) * c0);
          y[kk] = kk < k_hi ? fmaf(v, fz_col32[k], rsh32 * fz_col32[128 + k]) : 0.0f;
        }
        if ((K & 3) == 0) {      // rows are 16-byte aligned in the staging buffer: two 128-bit stores per group
          if (k_hi > 0) *reinterpret_cast<float4*>(dst) = make_float4(y[0], y[1], y[2], y[3]);
          if (k_hi > 4) *reinterpret_cast<float4*>(dst + 4) = make_float4(y[4], y[5], y[6], y[7]);
        } else {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            if (kk < k_hi) dst[kk] = y[kk];
        }
      }
      // ---- post phase: staged tile -> global (one bulk copy), column statistics from the staged tile
      const uint32_t bytes = (uint32_t)live_rows * K * 4;
      const bool bulk = bulk_ok && (bytes & 15) == 0;
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, %0;" ::"n"(kStatThreads) : "memory");
      if (live_rows > 0) {
        float* gdst = (float*)p.fz.out + m0 * p.fz.ldo;
        if (p.fz.mc_out) {
          // fused all-reduce: the staged tile is added into every rank's copy of t through the NVSwitch multicast
          // address (in-switch reduction), 16 bytes per instruction; ldo == K and K % 4 == 0 are checked at launch
          float* mdst = p.fz.mc_out + m0 * K;
          const float4* s4 = reinterpret_cast<const float4*>(sy);
          for (int i = et; i < live_rows * K / 4; i += kStatThreads) multimem_red_add_v4(mdst + 4 * i, s4[i]);
        } else if (bulk) {
          if (et == 0) bulk_s2g(gdst, smem_u32(sy), bytes);
        } else {
          for (int i = et; i < live_rows * K; i += kStatThreads) gdst[(long long)(i / K) * p.fz.ldo + i % K] = sy[i];
        }
        if (kStats && lane < K) {
          for (int r = ew; r < live_rows; r += kEpilogueWarps) {
            const float yv = sy[r * K + lane];
            mx = max(mx, __float_as_uint(fabsf(yv) * row_sc[r]));   // |y| d max(1, mu) >= 0; a NaN sorts above +inf
            sm = fma(row_w[r], (double)yv, sm);
          }
        }
      }
      // the staging buffer is rewritten by the next tile: the bulk store must have read it, the statistics too
      if (et == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      asm volatile("bar.sync 1, %0;" ::"n"(kStatThreads) : "memory");
      if (ew == 0 && lane == 0) LMDEC_STAMP(10, (int)tile_no);
    }
  }
  LMDEC_WAITOUT(0, ew == 0 && lane == 0);
  if (et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  if (p.fz.mc_out) __threadfence_system();    // the reductions sent to the peers are ordered before the kernel's end
  if (kStats) {
    fold_mx[ew * 32 + lane] = mx;
    fold_sm[ew * 32 + lane] = sm;
    asm volatile("bar.sync 1, %0;" ::"n"(kStatThreads) : "memory");
    const int k = et;
    if (k < K) {
      uint32_t m = 0u;
      double ws = 0.0;
      for (int j = 0; j < kEpilogueWarps; ++j) {    // fixed order
        m = max(m, fold_mx[j * 32 + k]);
        ws += fold_sm[j * 32 + k];
      }
      double* out = p.fz.stats_partial + (size_t)blockIdx.x * 3 * K;
      out[k] = m > 0x7f800000u ? (double)INFINITY : (double)__uint_as_float(m);
      out[K + k] = 0.0;
      out[2 * K + k] = ws;
    }
    stats_last_cta_fold(p, k, K, fold_scratch);
  }
}

// the last CTA to get here folds all per-CTA partial rows (block order) into the final statistics
__device__ __forceinline__ void stats_last_cta_fold(const GenoMmaParams& p, int k, int K, double* scratch) {
  __threadfence();
  asm volatile("bar.sync 1, %0;" ::"n"(kStatThreads) : "memory");
  if (k == 0) scratch[0] = atomicAdd(p.fz.stats_ticket, 1u) == gridDim.x - 1 ? 1.0 : 0.0;
  asm volatile("bar.sync 1, %0;" ::"n"(kStatThreads) : "memory");
  if (scratch[0] != 0.0) {
    __threadfence();
    // thread (slice, column) folds the rows slice, slice + n_sl, ...; thread `column` then folds the slices in order
    asm volatile("bar.sync 1, %0;" ::"n"(kStatThreads) : "memory");
    const int n_sl = kStatThreads / K, kc = k % K, sl = k / K;
    if (sl < n_sl) {
      double m1 = 0.0, m2 = 0.0, ws = 0.0;
      for (unsigned b = sl; b < gridDim.x; b += n_sl) {
        const double* in = p.fz.stats_partial + (size_t)b * 3 * K;
        m1 = fmax(m1, __ldcg(in + kc));
        m2 = fmax(m2, __ldcg(in + K + kc));
        ws += __ldcg(in + 2 * K + kc);
      }
      scratch[(sl * 3 + 0) * K + kc] = m1;
      scratch[(sl * 3 + 1) * K + kc] = m2;
      scratch[(sl * 3 + 2) * K + kc] = ws;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kStatThreads) : "memory");
    if (k < K) {
      double m1 = 0.0, m2 = 0.0, ws = 0.0;
      for (int j = 0; j < n_sl; ++j) {
        m1 = fmax(m1, scratch[(j * 3 + 0) * K + k]);
        m2 = fmax(m2, scratch[(j * 3 + 1) * K + k]);
        ws += scratch[(j * 3 + 2) * K + k];
      }
      const double s_ = fmax(m1, m2);
      const double mult = s_ > 0.0 ? p.fz.stats_lim / s_ : 1.0;   // as colstats_finish_kernel
      double* st = p.fz.stats_final;
      st[k] = mult;
      st[K + k] = 1.0 / mult;
      st[2 * K + k] = -ws;
      st[3 * K + k] = m1;
      st[4 * K + k] = m2;
    }
    if (k == 0) *p.fz.stats_ticket = 0u;
  }
}

template <int P>
__device__ __forceinline__ void epilogue_dispatch(const GenoMmaParams& p, uint32_t tbase, uint32_t acc_full,
                                                  uint32_t acc_empty, int warp, int lane, int acc_sets, uint32_t acc_w,
                                                  const double* fz_col, double* stats_acc) {
#define LMDEC_EPI(MODE, OUT) \
  epilogue_loop<P, MODE, OUT>(p, tbase, acc_full, acc_empty, warp, lane, acc_sets, acc_w, fz_col, stats_acc)
  const bool f32 = LMDEC_EPI_F32 && p.fuse && p.fz.out_f32;
  if (f32 && p.epi_staged) {
    uint8_t* stage = reinterpret_cast<uint8_t*>(stats_acc);
    if (p.mode == 1) {
      if (p.fuse == 2) epilogue_loop32<P, 1, true>(p, tbase, acc_full, acc_empty, warp, lane, acc_sets, acc_w, fz_col, stage);
      else epilogue_loop32<P, 1, false>(p, tbase, acc_full, acc_empty, warp, lane, acc_sets, acc_w, fz_col, stage);
    } else {
      if (p.fuse == 2) epilogue_loop32<P, 0, true>(p, tbase, acc_full, acc_empty, warp, lane, acc_sets, acc_w, fz_col, stage);
      else epilogue_loop32<P, 0, false>(p, tbase, acc_full, acc_empty, warp, lane, acc_sets, acc_w, fz_col, stage);
    }
    return;
  }
  if (p.mode == 1) {
    if (p.fuse == 2 && f32) LMDEC_EPI(1, kOutFusedStats32);
    else if (p.fuse == 2) LMDEC_EPI(1, kOutFusedStats);
    else if (f32) LMDEC_EPI(1, kOutFused32);
    else if (p.fuse) LMDEC_EPI(1, kOutFused);
    else if (p.use_atomics) LMDEC_EPI(1, kOutAtomic);
    else LMDEC_EPI(1, kOutStore);
  } else {
    if (p.fuse == 2 && f32) LMDEC_EPI(0, kOutFusedStats32);
    else if (p.fuse == 2) LMDEC_EPI(0, kOutFusedStats);
    else if (f32) LMDEC_EPI(0, kOutFused32);
    else if (p.fuse) LMDEC_EPI(0, kOutFused);
    else if (p.use_atomics) LMDEC_EPI(0, kOutAtomic);
    else LMDEC_EPI(0, kOutStore);
  }
#undef LMDEC_EPI
}

// MMA issuer warp.  The whole warp runs the loop (uniform control flow keeps the operand arithmetic on the uniform
// datapath); one elected lane issues.  With a plain `if (lane == 0)` ptxas wraps every UTCIMMA operand in an
// ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall (~100 cycles per MMA, profiles/r01).
//   kPair: schedule (kPairNone / M / E / T);  kMiss: value + indicator planes (else K-block parity per issuer);
//   kIlv: single tiles, issuer pair `issuer >> 1` takes the chunks of that parity within a work item.
template <int kPair, bool kMiss, bool kIlv, bool kRes = false>
__device__ __forceinline__ void issuer_loop(const GenoMmaParams& p, int issuer, int lane, uint32_t tbase, uint32_t bar0,
                                            uint32_t smem_b_a, int n_work, int chunks_per_split) {
  constexpr bool quad = kIssuerWarps == 4 && kPair == kPairM;
  constexpr int tiles_per_item = kPair ? 2 : 1;
  constexpr bool all_kb = kMiss || kPair == kPairT;   // this issuer runs all 8 K-blocks of a chunk (else parity me)
  const uint32_t empty_a = bar0 + 8 * kBarEmptyA, full_a = bar0 + 8 * kBarFullA;
  const uint32_t full_b = bar0 + 8 * kBarFullB, empty_b = bar0 + 8 * kBarEmptyB;
  const uint32_t acc_full = bar0 + 8 * kBarAccFull, acc_empty = bar0 + 8 * kBarAccEmpty;
  const uint32_t first_done = bar0 + 8 * kBarFirstDone;
  const int N = p.N;
  const uint32_t img_bytes = (uint32_t)kKbPerChunk * N * 32;
  const uint32_t stage_bytes = img_bytes * p.n_img;
  const uint32_t a_stage_cols = kMiss ? 128 : 64;
  const uint32_t acc_w = N <= 64 ? 64 : 128;
  const int acc_sets = p.acc_sets;
  const uint32_t a_stages = (uint32_t)p.a_stages, b_stages = (uint32_t)p.b_stages;
  const int me = issuer & 1;            // plane (missing data) or K-block parity
  const int hi = issuer >> 1;           // quad: the tile of the pair this issuer owns; kIlv: its chunk parity
  const uint32_t idesc = make_idesc_i8(128, (uint32_t)N, 0, 1);
  const uint32_t lbo = (uint32_t)N * 16;
  const uint32_t plane = kMiss ? (uint32_t)me : 0u;
  const uint32_t img = (kMiss && p.n_img == 2) ? (uint32_t)me : 0u;
  // descriptor of the B ring base; stage / image / K-block offsets are added to its 16-byte address field
  const uint64_t bdesc_ring = make_smem_desc(smem_b_a, lbo, 128) + (uint64_t)((img * img_bytes) >> 4);
  const uint32_t a_ring = tbase + (uint32_t)p.a_ring_col0 + plane * 64;
  const uint32_t kb_step = (uint32_t)(2 * N);  // one K-block of the image, in 16-byte units
  LMDEC_WAITDECL;
  uint32_t a_stage = 0, a_phase = 0, b_stage = 0, b_phase = 0;
  uint32_t tile_no = 0;
  int tr = 0;
  (void)tr;
  for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++tile_no) {
    const int sp = w / p.m_units;
    const int c0 = sp * chunks_per_split;
    const int c1 = min(c0 + chunks_per_split, p.n_chunks);
    const uint32_t set = kPair ? 0u : tile_no % acc_sets;
    const uint32_t set_phase = kPair ? (tile_no & 1) : ((tile_no / acc_sets) & 1);
    LMDEC_WAITACC(1, issuer == 0 && lane == 0, mbar_wait_a(acc_empty + 8 * set, set_phase ^ 1));
    tc_fence_after_sync();
    uint32_t accumulate[2] = {0, 0};
    // interleaved pairs: pair 0 owns the first chunk of the work item and its first MMA overwrites the accumulator;
    // pair 1 may only add to it after that MMA has completed (first_done, one barrier per accumulator)
    bool ordered = !kIlv || hi == 0;
    if (kIlv && hi == 1) accumulate[0] = 1;
    for (int c = c0; c < c1; ++c) {
      const bool my_chunk = !kIlv || ((c - c0) & 1) == hi;
      uint64_t bdesc0 = 0;
      if (kRes) {
        // resident image: chunk c of the (unsplit) contraction sits at c * stage_bytes, loaded once per kernel
        if (tile_no == 0 && c == c0) mbar_wait_a(full_b, 0);
        bdesc0 = bdesc_ring + (uint64_t)(((uint32_t)c * stage_bytes) >> 4);
      } else if (my_chunk) {
        LMDEC_WAITACC(2, issuer == 0 && lane == 0, mbar_wait_a(full_b + 8 * b_stage, b_phase));
        if (issuer == 0 && lane == 0) LMDEC_STAMP(6, tr);   // B stage full seen by issuer 0
        bdesc0 = bdesc_ring + (uint64_t)((b_stage * stage_bytes) >> 4);
      }
#pragma unroll
      for (int t = 0; t < tiles_per_item; ++t) {
        const bool mine = my_chunk && (quad ? (t == hi) : (kPair != kPairT || t == me));
        if (mine) {
          LMDEC_WAITACC(3, issuer == 0 && lane == 0, mbar_wait_a(full_a + 8 * a_stage, a_phase));
          if (issuer == 0 && lane == 0) LMDEC_STAMP(4, tr);   // A stage full seen by issuer 0
          tc_fence_after_sync();
          const uint32_t slot = kPair ? (uint32_t)t : set;
          const uint32_t d_addr = tbase + (kPair == kPairT ? slot : slot * 2 + me) * acc_w;
          const uint32_t a_addr = a_ring + a_stage * a_stage_cols;
          if (kIlv && !ordered) {
            LMDEC_WAITACC(4, issuer == 2 && lane == 0, mbar_wait_a(first_done + 8 * (set * 2 + me), set_phase));
            tc_fence_after_sync();
            ordered = true;
          }
#ifdef LMDEC_ABL_NOMMA      // ablation (timing only): barriers and commits, no tensor-core work
          if (elect_one()) {
            if (kIlv && !accumulate[t]) umma_commit_a(first_done + 8 * (set * 2 + me));
            umma_commit_a(empty_a + 8 * a_stage);
          }
          if (false) {
#else
          if (elect_one()) {
#endif
            if (all_kb) {
              umma_ts_i8(d_addr, a_addr, bdesc0, idesc, accumulate[t]);
              if (kIlv && !accumulate[t]) umma_commit_a(first_done + 8 * (set * 2 + me));
#pragma unroll
              for (int j = 1; j < kKbPerChunk; ++j)
                umma_ts_i8(d_addr, a_addr + j * 8, bdesc0 + j * kb_step, idesc, 1u);
            } else {
              umma_ts_i8(d_addr, a_addr + me * 8, bdesc0 + me * kb_step, idesc, accumulate[t]);
              if (kIlv && !accumulate[t]) umma_commit_a(first_done + 8 * (set * 2 + me));
#pragma unroll
              for (int j = 1; j < kKbPerChunk / 2; ++j)
                umma_ts_i8(d_addr, a_addr + (2 * j + me) * 8, bdesc0 + (2 * j + me) * kb_step, idesc, 1u);
            }
            umma_commit_a(empty_a + 8 * a_stage);
          }
          __syncwarp();
          if (issuer == 0 && lane == 0) LMDEC_STAMP(5, tr);   // MMAs issued + committed
          if (issuer == 0) ++tr;
          accumulate[t] = 1;
        }
        if (++a_stage == a_stages) {
          a_stage = 0;
          a_phase ^= 1;
        }
      }
      if (!kRes) {
        if (my_chunk) {
          if (elect_one()) umma_commit_a(empty_b + 8 * b_stage);  // after every MMA that reads this B stage
          __syncwarp();
          if (issuer == 0 && lane == 0) LMDEC_STAMP(8, tr - 1);   // B stage released
        }
        if (++b_stage == b_stages) {
          b_stage = 0;
          b_phase ^= 1;
        }
      }
    }
    if (elect_one()) umma_commit_a(acc_full + 8 * set);
    __syncwarp();
    if (issuer == 0 && lane == 0) LMDEC_STAMP(7, (int)tile_no);   // work item's MMAs issued
  }
  LMDEC_WAITOUT(1, issuer == 0 && lane == 0);
  LMDEC_WAITOUT(2, issuer == 0 && lane == 0);
  LMDEC_WAITOUT(3, issuer == 0 && lane == 0);
  LMDEC_WAITOUT(4, issuer == 2 && lane == 0);
}


// Round-robin issuers for data WITHOUT missing codes (one plane): issuer i owns the tile-chunks i, i + 4, ... of a work
// item (a tile-chunk = one 128 x 256 block of one m-tile; pair mode streams tile 0 / tile 1 alternately) and issues all 8
// K-blocks of them into ONE accumulator per tile.  Measured with the wait accounting (profiles/r02/waits_*.txt): with the
// K-block-parity split each issuer made a full set of barrier round trips (~8 per chunk, ~100 cycles each) for only 4
// MMAs and was busy 85 % of the time -- the issuers, not the tensor pipe (36 % busy) or the producers, paced the
// no-missing shapes.  Whole tile-chunks halve the barrier work per MMA, four issuers overlap what is left, and the freed
// accumulator columns deepen the decoded-operand ring (6 stages).
// The first MMA of a tile overwrites its accumulator: the issuers that do not own the tile's first tile-chunk wait until
// that one's operand stage has been released (= its MMAs are complete) before they add to the accumulator.
// (A first version let one issuer run both tiles of a chunk: 16 MMAs and 3-5 tcgen05.commit in one burst hung on some
// SMs, with the first commits never arriving; bursts of 8 MMAs + 2 commits, as everywhere else in this file, are fine.)
template <int kTiles>
__device__ __forceinline__ void issuer_loop_rr(const GenoMmaParams& p, int issuer, int lane, uint32_t tbase,
                                               uint32_t bar0, uint32_t smem_b_a, int n_work, int chunks_per_split) {
  const uint32_t empty_a = bar0 + 8 * kBarEmptyA, full_a = bar0 + 8 * kBarFullA;
  const uint32_t full_b = bar0 + 8 * kBarFullB, empty_b = bar0 + 8 * kBarEmptyB;
  const uint32_t acc_full = bar0 + 8 * kBarAccFull, acc_empty = bar0 + 8 * kBarAccEmpty;
  const int N = p.N;
  const uint32_t img_bytes = (uint32_t)kKbPerChunk * N * 32;
  const uint32_t stage_bytes = img_bytes * p.n_img;
  const uint32_t acc_w = N <= 64 ? 64 : 128;
  const int acc_sets = p.acc_sets;
  const uint32_t a_stages = (uint32_t)p.a_stages, b_stages = (uint32_t)p.b_stages;
  const uint32_t idesc = make_idesc_i8(128, (uint32_t)N, 0, 1);
  const uint64_t bdesc_ring = make_smem_desc(smem_b_a, (uint32_t)N * 16, 128);
  const uint32_t a_ring = tbase + (uint32_t)p.a_ring_col0;
  const uint32_t kb_step = (uint32_t)(2 * N);
  LMDEC_WAITDECL;
  uint32_t a_stage = 0, a_phase = 0, b_stage = 0, b_phase = 0;
  uint32_t tile_no = 0;
  for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++tile_no) {
    const int sp = w / p.m_units;
    const int c0 = sp * chunks_per_split;
    const int c1 = min(c0 + chunks_per_split, p.n_chunks);
    const uint32_t set = kTiles == 2 ? 0u : tile_no % acc_sets;
    const uint32_t set_phase = kTiles == 2 ? (tile_no & 1) : ((tile_no / acc_sets) & 1);
    LMDEC_WAITACC(1, issuer == 0 && lane == 0, mbar_wait_a(acc_empty + 8 * set, set_phase ^ 1));
    tc_fence_after_sync();
    const uint32_t first_stage = a_stage, first_phase = a_phase;   // ring position of the item's first tile-chunk
    bool ordered = issuer < kTiles;      // issuer t owns the first tile-chunk of tile t
    uint32_t tc = 0;                     // tile-chunk index within the work item
    for (int c = c0; c < c1; ++c) {
#pragma unroll
      for (int t = 0; t < kTiles; ++t, ++tc) {
        if ((tc & 3) == (uint32_t)issuer) {
          LMDEC_WAITACC(2, issuer == 0 && lane == 0, mbar_wait_a(full_b + 8 * b_stage, b_phase));
          const uint64_t bdesc0 = bdesc_ring + (uint64_t)((b_stage * stage_bytes) >> 4);
          LMDEC_WAITACC(3, issuer == 0 && lane == 0, mbar_wait_a(full_a + 8 * a_stage, a_phase));
          tc_fence_after_sync();
          const uint32_t d_addr = tbase + (set * kTiles + t) * acc_w;
          const uint32_t a_addr = a_ring + a_stage * 64;
          if (!ordered) {
            uint32_t fs = first_stage + t, fp = first_phase;
            if (fs >= a_stages) {
              fs -= a_stages;
              fp ^= 1;
            }
            LMDEC_WAITACC(4, issuer == 2 && lane == 0, mbar_wait_a(empty_a + 8 * fs, fp));
            tc_fence_after_sync();
            ordered = true;
          }
          if (elect_one()) {
            umma_ts_i8(d_addr, a_addr, bdesc0, idesc, c == c0 ? 0u : 1u);   // (c == c0 only for the tile's first owner)
#pragma unroll
            for (int j = 1; j < kKbPerChunk; ++j) umma_ts_i8(d_addr, a_addr + j * 8, bdesc0 + j * kb_step, idesc, 1u);
            umma_commit_a(empty_a + 8 * a_stage);
            umma_commit_a(empty_b + 8 * b_stage);      // (count kTiles: one per tile-chunk that read this B stage)
          }
          __syncwarp();
        }
        if (++a_stage == a_stages) {
          a_stage = 0;
          a_phase ^= 1;
        }
      }
      if (++b_stage == b_stages) {
        b_stage = 0;
        b_phase ^= 1;
      }
    }
    if (elect_one()) umma_commit_a(acc_full + 8 * set);
    __syncwarp();
  }
  LMDEC_WAITOUT(1, issuer == 0 && lane == 0);
  LMDEC_WAITOUT(2, issuer == 0 && lane == 0);
  LMDEC_WAITOUT(3, issuer == 0 && lane == 0);
  LMDEC_WAITOUT(4, issuer == 2 && lane == 0);
}

template <int kGroups>
__global__ void __launch_bounds__(kThreads, 1) geno_mma_kernel(const GenoMmaParams p) {
  constexpr int kGroupWarps = kProducerWarps / kGroups;   // producer warps per tile-chunk
  constexpr int kKb = 32 / kGroupWarps;                   // K-blocks of a chunk decoded by one producer thread
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[kNumBars];  // see the kBar* offsets: all rings' full / empty barriers
  __shared__ uint32_t tmem_base_s;
  __shared__ double fz_col[256 + 128];  // fused finalize: colscale[k] at [k], colshift[k] at [128 + k]; then 256 float32 copies
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t full_t = bar0 + 8 * kBarFullT, empty_t = bar0 + 8 * kBarEmptyT;   // packed tiles (smem ring)
  const uint32_t full_a = bar0 + 8 * kBarFullA, empty_a = bar0 + 8 * kBarEmptyA;   // decoded A (TMEM ring)
  const uint32_t full_b = bar0 + 8 * kBarFullB, empty_b = bar0 + 8 * kBarEmptyB;   // B image (smem ring)
  const uint32_t acc_full = bar0 + 8 * kBarAccFull, acc_empty = bar0 + 8 * kBarAccEmpty;  // accumulator sets
  // (kBarFirstDone: [set][accumulator], its overwriting MMA has completed -- interleaved issuer pairs)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.N;
  const uint32_t img_bytes = (uint32_t)kKbPerChunk * N * 32;
  const uint32_t stage_bytes = img_bytes * p.n_img;
  uint8_t* smem_t = smem;                                              // tile ring
  uint8_t* smem_b = smem + (size_t)p.tile_stages * kTileBytes;         // B ring
  const uint32_t smem_t_a = smem_u32(smem_t), smem_b_a = smem_u32(smem_b);
  const int a_stages = p.a_stages;
  // pair mode M with 4 issuer warps: one issuer per (tile, plane) accumulator, so each thread issues 8 instead of 16
  // MMAs per staged operand chunk (a single thread sustains one narrow MMA per ~54 cycles)
  const bool quad = kIssuerWarps == 4 && p.pair == kPairM;
  // single tiles: two issuer pairs take alternate chunks, so the per-chunk barrier round trips of an issuing thread
  // (~900 cycles of waits and commits around ~400 of MMA issue, profiles/r01/pipeline_trace.txt) overlap
  const bool interleave = kIssuerWarps == 4 && !p.pair && p.issuer_groups == 2;
  const int n_issuers = (quad || interleave || p.rr) ? 4 : 2;
  const int tiles_per_item_init = p.pair ? 2 : 1;
  const uint32_t kARingCol0 = (uint32_t)p.a_ring_col0;
  const uint32_t a_stage_cols = p.has_missing ? 128 : 64;
  const int acc_sets = p.acc_sets;
  const uint32_t acc_w = N <= 64 ? 64 : 128;

  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  if (p.fuse && tid >= 64 && tid < 64 + 128) {
    const int k = tid - 64;
    fz_col[k] = k < p.K ? p.fz.colscale[k] : 0.0;
    fz_col[128 + k] = (k < p.K && p.fz.colshift) ? p.fz.colshift[k] : 0.0;
    float* c32 = reinterpret_cast<float*>(fz_col + 256);
    c32[k] = (float)fz_col[k];
    c32[128 + k] = (float)fz_col[128 + k];
  }
  if (tid == 32) {
    for (int i = 0; i < kMaxTileStages; ++i) {
      mbar_init(&bars[kBarFullT + i], 1);
      mbar_init(&bars[kBarEmptyT + i], kGroupWarps);
    }
    for (int i = 0; i < kMaxAStages; ++i) {
      mbar_init(&bars[kBarFullA + i], kGroupWarps);
      mbar_init(&bars[kBarEmptyA + i], (p.pair == kPairT || p.rr) ? 1 : 2);
    }
    for (int i = 0; i < kMaxBStages; ++i) {
      mbar_init(&bars[kBarFullB + i], 1);
      mbar_init(&bars[kBarEmptyB + i], p.rr ? tiles_per_item_init : (quad ? 4 : 2));
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[kBarAccFull + i], n_issuers);
      mbar_init(&bars[kBarAccEmpty + i], kEpilogueWarps);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&bars[kBarFirstDone + i], 1);
    mbar_fence_init();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tbase = tmem_base_s;
#ifdef LMDEC_DEBUG_HANG
  if (blockIdx.x == 0 && tid == 0) printf("bars@%u rr=%d pair=%d acc_sets=%d a_stages=%d b_stages=%d n_chunks=%d splits=%d m_units=%d\n", bar0, p.rr, p.pair, p.acc_sets, p.a_stages, p.b_stages, p.n_chunks, p.splits, p.m_units);
#endif
#ifdef LMDEC_TRACE
  if (blockIdx.x == 0 && tid == 0) g_wait[14] = clock64();
#endif

  const int n_work = p.m_units * p.splits;
  const int chunks_per_split = (p.n_chunks + p.splits - 1) / p.splits;
  const int tiles_per_item = p.pair ? 2 : 1;

  if (warp >= kProducerWarp0 && warp < kProducerWarp0 + kProducerWarps) {
    // ------------------------------------------------------------------ producers
    const int pw = warp - kProducerWarp0;
    const int group = pw / kGroupWarps, lw = pw % kGroupWarps;
    const int q = lw & 3, half = lw >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tbase + ((uint32_t)(q * 32) << 16);
    const uint32_t my_off = (uint32_t)((kKb / 2) * half) * 2048 + (uint32_t)row * 16;
    uint32_t gsc = 0;   // running tile-chunk count: with two groups, group g takes those of parity g
    uint32_t a_stage = 0, a_phase = 0, t_stage = 0, t_phase = 0;
    int tr = 0;
    const bool tracer = pw == 0 && lane == 0;
    LMDEC_WAITDECL;
    (void)tr; (void)tracer;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
      const int sp = w / p.m_units;
      const int c0 = sp * chunks_per_split;
      const int c1 = min(c0 + chunks_per_split, p.n_chunks);
      // in pair mode the tile stream alternates between the two m-tiles of the work item, chunk by chunk
      for (int sc = (c1 - c0) * tiles_per_item; sc > 0; --sc, ++gsc) {
        if (kGroups > 1 && (gsc % (uint32_t)kGroups) != (uint32_t)group) {   // the other group's tile-chunk: step the ring positions
          if (++a_stage == (uint32_t)a_stages) {
            a_stage = 0;
            a_phase ^= 1;
          }
          if (++t_stage == (uint32_t)p.tile_stages) {
            t_stage = 0;
            t_phase ^= 1;
          }
          continue;
        }
#ifndef LMDEC_ABL_NOTILESYNC
        LMDEC_WAITACC(5, tracer, mbar_wait_a(full_t + 8 * t_stage, t_phase));
#endif
        if (tracer) LMDEC_STAMP(0, tr);   // tile arrived
        const uint32_t src = smem_t_a + t_stage * kTileBytes + my_off;
        uint4 raw[kKb / 2];
#ifdef LMDEC_ABL_NOLDS
#pragma unroll
        for (int i = 0; i < kKb / 2; ++i) raw[i] = make_uint4(src, src + 1, src + 2, src + 3);
#else
#pragma unroll
        for (int i = 0; i < kKb / 2; ++i) raw[i] = lds128(src + i * 2048);
#endif
        LMDEC_WAITACC(6, tracer, mbar_wait_a(empty_a + 8 * a_stage, a_phase ^ 1));
        if (tracer) LMDEC_STAMP(1, tr);   // A stage free
#ifndef LMDEC_ABL_NOFENCE
        tc_fence_after_sync();
#endif
        const uint32_t a_addr = lane_addr + kARingCol0 + a_stage * a_stage_cols;
        if (p.has_missing)
          produce_chunk<true, kKb>(raw, a_addr, kKb * half);
        else
          produce_chunk<false, kKb>(raw, a_addr, kKb * half);
        if (tracer) LMDEC_STAMP(2, tr);   // decode + tcgen05.st issued
#ifndef LMDEC_ABL_NOFENCE
        tmem_wait_st();
#endif
        if (tracer) LMDEC_STAMP(3, tr);   // stores complete
        ++tr;
#ifndef LMDEC_ABL_NOFENCE
        tc_fence_before_sync();
#endif
        __syncwarp();
        if (lane == 0) {
          mbar_arrive_a(full_a + 8 * a_stage);
#ifndef LMDEC_ABL_NOTILESYNC
          mbar_arrive_a(empty_t + 8 * t_stage);  // raw[] is consumed: the tile slot can be refilled
#endif
        }
        if (++a_stage == (uint32_t)a_stages) {
          a_stage = 0;
          a_phase ^= 1;
        }
        if (++t_stage == (uint32_t)p.tile_stages) {
          t_stage = 0;
          t_phase ^= 1;
        }
      }
    }
    LMDEC_WAITOUT(5, tracer);
    LMDEC_WAITOUT(6, tracer);
#if LMDEC_GENO_REGS
  } else if (warp >= kTileLoaderWarp) {
   reg_dec<40>();          // warps 16-23: loaders, issuers, two idle warps
   if (warp == kTileLoaderWarp) {
#else
  } else if (warp == kTileLoaderWarp) {
#endif
    // ------------------------------------------------------------------ packed-tile loader (TMA bulk copies)
#ifdef LMDEC_ABL_NOTILESYNC
    if (false) {
#else
    if (lane == 0) {
#endif
      uint32_t t_stage = 0, t_phase = 0;
      int ld_no = 0;
      (void)ld_no;
      LMDEC_WAITDECL;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        const int unit = w % p.m_units, sp = w / p.m_units;
        const int c0 = sp * chunks_per_split;
        const int c1 = min(c0 + chunks_per_split, p.n_chunks);
        // a pair whose second tile does not exist (odd m_tiles) streams the first tile twice; the epilogue drops it
        const int mt0 = p.pair ? 2 * unit : unit;
        const int mt1 = p.pair ? min(2 * unit + 1, p.m_tiles - 1) : unit;
        const uint8_t* tile_a = p.store + ((size_t)mt0 * p.store_kc) * kTileBytes;
        const uint8_t* tile_b = p.store + ((size_t)mt1 * p.store_kc) * kTileBytes;
        for (int c = c0; c < c1; ++c) {
          for (int t = 0; t < tiles_per_item; ++t) {
            LMDEC_WAITACC(7, true, mbar_wait_relaxed_a(empty_t + 8 * t_stage, t_phase ^ 1));
            LMDEC_STAMP(11, ld_no++);                                   // tile slot free, copy issued
#ifdef LMDEC_ABL_NOLOAD     // ablation (timing only): the packed-tile ring cycles without any HBM traffic
            mbar_arrive_a(full_t + 8 * t_stage);
#else
            mbar_arrive_expect_tx_a(full_t + 8 * t_stage, kTileBytes);
            bulk_g2s_a(smem_t_a + t_stage * kTileBytes, (t ? tile_b : tile_a) + (size_t)c * kTileBytes, kTileBytes,
                       full_t + 8 * t_stage);
#endif
            if (++t_stage == (uint32_t)p.tile_stages) {
              t_stage = 0;
              t_phase ^= 1;
            }
          }
        }
      }
      LMDEC_WAITOUT(7, true);
    }
  } else if (warp == kBLoaderWarp) {
    LMDEC_WAITDECL;
    // ------------------------------------------------------------------ B image loader
    if (lane == 0 && p.b_resident) {
      // the whole image once: every work item of this CTA multiplies by the same n_chunks * stage_bytes
      mbar_arrive_expect_tx_a(full_b, (uint32_t)p.n_chunks * stage_bytes);
      for (int c = 0; c < p.n_chunks; ++c) {
        const uint32_t dst = smem_b_a + (uint32_t)c * stage_bytes;
        bulk_g2s_a(dst, p.image1 + (size_t)c * img_bytes, img_bytes, full_b);
        if (p.n_img == 2) bulk_g2s_a(dst + img_bytes, p.image2 + (size_t)c * img_bytes, img_bytes, full_b);
      }
    } else if (lane == 0) {
      uint32_t b_stage = 0, b_phase = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        const int sp = w / p.m_units;
        const int c0 = sp * chunks_per_split;
        const int c1 = min(c0 + chunks_per_split, p.n_chunks);
        for (int c = c0; c < c1; ++c) {
          LMDEC_WAITACC(8, true, mbar_wait_relaxed_a(empty_b + 8 * b_stage, b_phase ^ 1));
          const uint32_t dst = smem_b_a + b_stage * stage_bytes;
          mbar_arrive_expect_tx_a(full_b + 8 * b_stage, stage_bytes);
          bulk_g2s_a(dst, p.image1 + (size_t)c * img_bytes, img_bytes, full_b + 8 * b_stage);
          if (p.n_img == 2)
            bulk_g2s_a(dst + img_bytes, p.image2 + (size_t)c * img_bytes, img_bytes, full_b + 8 * b_stage);
          if (++b_stage == (uint32_t)p.b_stages) {
            b_stage = 0;
            b_phase ^= 1;
          }
        }
      }
    }
    LMDEC_WAITOUT(8, lane == 0);
  } else if (warp >= kIssuerWarp0 && warp < kIssuerWarp0 + kIssuerWarps) {
    // ------------------------------------------------------------------ MMA issuers
    if (warp - kIssuerWarp0 < n_issuers) {
      const int issuer = warp - kIssuerWarp0;
#define LMDEC_ISSUE(PAIR, MISS, ILV) \
  issuer_loop<PAIR, MISS, ILV>(p, issuer, lane, tbase, bar0, smem_b_a, n_work, chunks_per_split)
      // one instance per schedule: the issuing thread is the serial bottleneck of the pipeline, every mode test left in
      // its per-chunk path costs throughput
      if (p.rr) {
        if (p.pair) issuer_loop_rr<2>(p, issuer, lane, tbase, bar0, smem_b_a, n_work, chunks_per_split);
        else issuer_loop_rr<1>(p, issuer, lane, tbase, bar0, smem_b_a, n_work, chunks_per_split);
      }
      else if (p.pair == kPairM) LMDEC_ISSUE(kPairM, true, false);
      else if (p.pair == kPairE) LMDEC_ISSUE(kPairE, false, false);
      else if (p.pair == kPairT) LMDEC_ISSUE(kPairT, false, false);
      else if (p.b_resident) {
#define LMDEC_ISSUE_RES(MISS, ILV) \
  issuer_loop<kPairNone, MISS, ILV, true>(p, issuer, lane, tbase, bar0, smem_b_a, n_work, chunks_per_split)
        if (p.has_missing) { if (interleave) LMDEC_ISSUE_RES(true, true); else LMDEC_ISSUE_RES(true, false); }
        else { if (interleave) LMDEC_ISSUE_RES(false, true); else LMDEC_ISSUE_RES(false, false); }
#undef LMDEC_ISSUE_RES
      }
      else if (p.has_missing) { if (interleave) LMDEC_ISSUE(kPairNone, true, true); else LMDEC_ISSUE(kPairNone, true, false); }
      else { if (interleave) LMDEC_ISSUE(kPairNone, false, true); else LMDEC_ISSUE(kPairNone, false, false); }
#undef LMDEC_ISSUE
    }
#if LMDEC_GENO_REGS
   }
  } else if (warp >= kEpilogueWarp0 && warp < kEpilogueWarp0 + kEpilogueWarps) {
    reg_inc<120>();         // (120 - 80) x 256 threads = what warps 16-23 gave back: (80 - 40) x 256
#else
  } else if (warp >= kEpilogueWarp0 && warp < kEpilogueWarp0 + kEpilogueWarps) {
#endif
    // ------------------------------------------------------------------ epilogue
    double* stats_acc = reinterpret_cast<double*>(smem + p.stats_off);
    switch (p.P) {
      case 1: epilogue_dispatch<1>(p, tbase, acc_full, acc_empty, warp, lane, acc_sets, acc_w, fz_col, stats_acc); break;
      case 2: epilogue_dispatch<2>(p, tbase, acc_full, acc_empty, warp, lane, acc_sets, acc_w, fz_col, stats_acc); break;
      case 3: epilogue_dispatch<3>(p, tbase, acc_full, acc_empty, warp, lane, acc_sets, acc_w, fz_col, stats_acc); break;
      default: epilogue_dispatch<4>(p, tbase, acc_full, acc_empty, warp, lane, acc_sets, acc_w, fz_col, stats_acc); break;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
#ifdef LMDEC_TRACE
  if (blockIdx.x == 0 && tid == 0) g_wait[15] = clock64() - g_wait[14];
#endif
  if (warp == 0) tmem_dealloc<512>(tbase);
}

}  // namespace lmdec
