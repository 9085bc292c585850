/* lmdec_b200 -- C ABI of the B200-native decomposition hot path.
 *
 * The reference (TNonet/lmdec) is pure Python: it has no FFI of its own.  The boundary its hot path sits behind is
 * the duck-typed operator protocol  .dot(x) / .T.dot(y)  that PowerMethod._solution_step calls
 * (lmdec/decomp/iter_methods.py:381-386) on the operator make_snp_array builds (lmdec/decomp/utils.py:195-289).
 * Every entry point below replaces one of those reference call sites; the citation names it.  INTEGRATION.md shows the
 * ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in _host; the caller (Python/PyTorch) owns all memory;
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it;
 *   - return value 0 = ok, nonzero = error; lmdec_last_error() returns a message for the calling thread;
 *   - no mutable global state: tunables and the kernel-timing events live in an opaque lmdec_ctx the caller owns
 *     (one per thread / stream; NULL = defaults); the only process-wide objects are the thread-local error string and
 *     an atomic launch counter.  Thread-compatible.
 *
 * Genotype store ("tile store"): 2-bit dosage codes, 4 per byte, LSB first: 0,1,2 = dosage, 3 = missing (NaN).
 * A PLINK .bed 2-bit field (hi lo): 00 hom-A1, 01 missing, 10 het, 11 hom-A2 (pandas_plink reads these as 0, NaN, 1, 2)
 * maps to the store code (hi' lo') = (lo, lo ^ hi); see lmdec_recode_bed.
 * A store holds a logical  M x Kd  matrix (M = output rows of a product, Kd = the contraction axis) as
 *   tile(mp, kc) = 8192 bytes for rows [128*mp, +128) x contraction [256*kc, +256), tiles ordered mp-major;
 *   inside a tile: byte (t, b) of row t's 64 packed bytes sits at  (b/16)*2048 + t*16 + b%16
 * so that one warp reads 512 contiguous bytes per 16-byte piece.  Padding codes are 0.
 * A genotype matrix is kept as TWO stores: sample-major (M = samples, feeds A.dot) and SNP-major (M = SNPs, the
 * .bed orientation, feeds A.T.dot).
 */
#ifndef LMDEC_B200_H
#define LMDEC_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define LMDEC_TILE_ROWS 128
#define LMDEC_TILE_KD 256
#define LMDEC_TILE_BYTES 8192

enum lmdec_dtype { LMDEC_I8 = 0, LMDEC_U8 = 1, LMDEC_F32 = 2, LMDEC_F64 = 3 };

const char* lmdec_last_error(void);
int lmdec_abi_version(void);
/* content hash of the CUDA sources this binary was built from (the Python loader rebuilds a stale library) */
const char* lmdec_source_hash(void);

/* Caller-owned context for the product kernel: tunables (below) and CUDA-event timing of its launches. */
typedef struct lmdec_ctx lmdec_ctx;
lmdec_ctx* lmdec_ctx_create(void);
void lmdec_ctx_destroy(lmdec_ctx* ctx);
enum lmdec_knob {
  LMDEC_KNOB_SM_RESERVE = 0,      /* SMs left out of the persistent grid so a collective on another stream can run; 0 */
  LMDEC_KNOB_PAIR_MODE = 1,       /* 1: long contractions process two 128-row tiles per staged operand chunk; 0: one */
  LMDEC_KNOB_PAIR_MIN_CHUNKS = 2, /* shortest contraction, in 256-wide chunks, that uses pair mode; 64 */
  LMDEC_KNOB_ISSUER_GROUPS = 3,   /* 2: two MMA issuer pairs take alternate chunks of single-tile work items; 1: one */
  LMDEC_KNOB_PRODUCER_GROUPS = 4, /* 2: two decode groups on alternate tile-chunks when no code is missing; 1: one */
  LMDEC_KNOB_PRODUCER_GROUPS_MISSING = 5, /* the same when codes are missing (only with a decoded-operand ring >= 3 deep) */
  LMDEC_KNOB_ACC_SETS = 6,        /* accumulator sets of single-tile work items: 0 = choose, 1 (deeper operand ring), 2 */
  LMDEC_KNOB_B_RESIDENT = 7,      /* 1: short unsplit contractions keep the whole operand image in shared memory; 0: ring */
  LMDEC_KNOB_ROUND_ROBIN = 8      /* 1: without missing codes 4 MMA issuers take whole chunks round-robin; 0: K-block parity */
};
int lmdec_ctx_set(lmdec_ctx* ctx, int knob, int value);
int lmdec_ctx_get(const lmdec_ctx* ctx, int knob);
/* CUDA-event timing of the product kernel's launches made through ctx, on their launching stream (bench.py's roofline) */
void lmdec_ctx_timing(lmdec_ctx* ctx, int on);
int lmdec_ctx_timing_read(lmdec_ctx* ctx, double* total_ms_host, int64_t* launches_host);
/* number of CUDA kernels this library has launched in the process (bench.py reports the delta per timed region) */
long long lmdec_launch_count(void);

/* bytes of a tile store holding a logical M x Kd matrix */
int64_t lmdec_store_bytes(int64_t M, int64_t Kd);

/* Pack a dense block into a tile store.  Replaces  fill_array + astype(int8)  and the NaN scan of mask_imputation
 * (lmdec/decomp/utils.py:15-30, 167-192, 272).  Element (m, kd) of the block is  src[m*stride_m + kd*stride_kd]
 * (element strides), m in [0,block_m), kd in [0,block_kd); it is stored at logical position (m0+m, kd0+kd) of a store
 * of logical size M x Kd.  m0 must be a multiple of 128 and kd0 of 256.  Values 0,1,2 are kept; NaN (float types) and
 * negative / 3 (integer types) become the missing code; anything else increments *bad_count (int32, device).  */
int lmdec_pack_2bit(const void* src, int dtype, int64_t stride_m, int64_t stride_kd, int64_t block_m, int64_t block_kd,
                    int64_t m0, int64_t kd0, uint8_t* store, int64_t M, int64_t Kd, int32_t* bad_count, void* stream);

/* Inverse of the above for tests: dst[m*ld + kd] (int8) = 0,1,2 or -1 for missing. */
int lmdec_unpack_2bit(const uint8_t* store, int64_t M, int64_t Kd, int8_t* dst, int64_t ld, void* stream);

/* PLINK .bed bytes (variant-major, rows padded to whole bytes, no 3-byte magic) -> SNP-major tile store.
 * Replaces pandas_plink.read_plink as called by load_plink_array (lmdec/array/io.py:16-73).  */
int lmdec_recode_bed(const uint8_t* bed, int64_t n_variants, int64_t n_samples, uint8_t* store, void* stream);

/* Transpose one tile store into the other orientation (M x Kd -> Kd x M), codes untouched. */
int lmdec_store_transpose(const uint8_t* src, int64_t M, int64_t Kd, uint8_t* dst, void* stream);

/* Per-row code counts of a store: counts[m*4 + c] (int64) = #{kd < kd_limit : code(m,kd) == c}.  With the SNP-major
 * store this is the exact integer content of get_array_moments' nanmean / nanstd (lmdec/decomp/utils.py:64-82):
 * mean_j = (c1 + 2 c2) / (n - c3).  */
int lmdec_code_counts(const uint8_t* store, int64_t M, int64_t Kd, int64_t kd_limit, int64_t* counts, void* stream);

/* Quantise a float64 operand  B[kd, k] * row1[kd] (* row2[kd])  into P signed-byte planes laid out as the tensor-core
 * B image: image[kd/32][N*32] in the no-swizzle K-major core-matrix order, column n = k*P + plane,
 * value = round(B * mult[k]) written as balanced base-256 digits.  rows >= kd_limit and columns >= P*K are zero.
 * row1/row2 may be NULL.  With q1 = round(B*row1*mult) and q2 = round(B*row1*row2*mult) (the mean-imputation operand
 * mu o D T  of StackedArray.dot's sparse-mask term, lmdec/array/stacked.py:204-217), image1 holds q1 and image2
 * (optional) holds  q2 - 3*q1 : lmdec_geno_matmul multiplies the RAW 2-bit codes (3 at a missing entry) by image1, so
 * the indicator plane times image2 restores  value.q1 + miss.q2  exactly.  Keep |q2 - 3 q1| below
 * 127*(256^P - 1)/255 (values are clamped there): with row2 in [0, 2] that means |q1| <= 2^(8P-3).  */
int lmdec_quantize_planes(const double* B, int64_t ldb, int64_t Kd, int64_t kd_limit, int K, int P, int N,
                          const double* mult, const double* row1, const double* row2, int8_t* image1, int8_t* image2,
                          void* stream);
int64_t lmdec_image_bytes(int64_t Kd, int N);

/* column-wise max |B[kd,k]*row1[kd](*row2[kd])| over kd < kd_limit -> out[k] (and out2[k] with row2), float64 */
int lmdec_colabsmax(const double* B, int64_t ldb, int64_t kd_limit, int K, const double* row1, const double* row2,
                    double* out, double* out2, void* stream);

/* The hot product:  out[m, k] = sum_kd  value(m,kd) * q1[kd,k]  (+ miss(m,kd) * q2[kd,k])   exactly, in int64,
 * where value is the dosage (missing -> 0), miss the missing indicator and q the integer the planes encode.
 * Replaces the chunked int8->f64 matmuls of StackedArray.dot / ChainedArray.dot (lmdec/array/stacked.py:172-217,
 * lmdec/array/chained.py:169-177) under PowerMethod._solution_step (lmdec/decomp/iter_methods.py:381-386) and
 * ScaledCenterArray.sym_mat_mult / dot (lmdec/array/scaled_legacy.py:379-399, 526-542).
 *   store, M, Kd       the tile store (logical size);  m_limit <= M rows are produced, kd_limit <= Kd is contracted
 *   has_missing        0: every code is 0..2 (one MMA per tile);  1: also run the indicator plane
 *   mode               0: out0 = value.q1 + miss.q2 (image2 as written by lmdec_quantize_planes; without image2
 *                         and with has_missing the result is code.q1, i.e. value.q1 + 3 miss.q1)
 *                      1: out0 = value.q1 ; out1 = miss.q1      (needs has_missing)
 *   splits             split-K factor; 0 = choose; >1 accumulates with int64 atomics (outputs are zeroed here);
 *                      a split never spans more than 65,536 contraction elements (int32 accumulators)
 * out0/out1: int64 [m_limit, K] row-major.  */
int lmdec_geno_matmul(lmdec_ctx* ctx, const uint8_t* store, int64_t M, int64_t Kd, int64_t m_limit, int64_t kd_limit,
                      int has_missing, int mode, const int8_t* image1, const int8_t* image2, int K, int P, int N,
                      int64_t* out0, int64_t* out1, int splits, void* stream);

/* out[m,k] = (acc0[m,k] + rowa[m]*acc1[m,k]) * colscale[k] * rowscale[m] + rowshift[m]*colshift[k]
 * (any of acc1/rowa, rowscale, rowshift/colshift may be NULL).  The diag / rank-1 terms of the operator algebra:
 * diag_dot (lmdec/array/matrix_ops.py:158-198), the (p,) mean term of StackedArray.dot (stacked.py:207-215),
 * ScaledCenterArray._center_x/_scale_x (scaled_legacy.py:452-524).  */
int lmdec_finalize(const int64_t* acc0, const int64_t* acc1, int64_t M, int K, const double* colscale,
                   const double* rowa, const double* rowscale, const double* rowshift, const double* colshift,
                   double* out, int64_t ldo, void* stream);

/* G[K,K] = X^T X and colsum[K] of a float64 X[m, K] (ld = ldx): the Gram matrix CholeskyQR2 factors in place of
 * dask tsqr (lmdec/decomp/iter_methods.py:382, lmdec/array/matrix_ops.py:72).  Deterministic two-stage reduction;
 * work must hold lmdec_gram_work_doubles(m, K) doubles.  Also serves subspace_dist's Vi^T Vj (metrics.py:81) via Y. */
int64_t lmdec_gram_work_doubles(int64_t m, int K);
int lmdec_gram(const double* X, int64_t ldx, const double* Y, int64_t ldy, int64_t m, int K, double* G, double* colsum,
               double* work, void* stream);

/* q_value_converge (lmdec/array/metrics.py:219-235): out[0] = ||si-sj||_2 / ((||si||+||sj||)/2) (scale != 0). */
/* Gram matrix of a row shard ADDED into every rank's copy of a symmetric K x K buffer through its NVSwitch multicast
 * address (`multimem.red.add.f64`): the "k x k Gram allreduce" of the north-star's CholeskyQR2 without a collective call
 * (reference: the stacked-R step of dask tsqr, lmdec/decomp/iter_methods.py:382).  The buffer is zero on every rank
 * before any rank launches and complete after one cross-rank barrier; m = 0 (a rank without rows) launches nothing. */
int lmdec_gram_reduce(const double* X, int64_t ldx, int64_t m, int K, double* G_multicast, double* work, void* stream);
int lmdec_qvals(const double* si, const double* sj, int k, int scale, double* out, void* stream);
/* rmse_k's Frobenius part (metrics.py:161): out[0] = sum_{m,k} (A[m,k] - B[m,k]*s[k])^2  (deterministic). */
int lmdec_sqdiff(const double* A, int64_t lda, const double* B, int64_t ldb_, const double* s, int64_t m, int K,
                 double* out, double* work /* >= 592 doubles */, void* stream);

/* One application of the genotype operator, all kernels enqueued back to back (no host work in between):
 *   transpose = 0:  out[m_limit, K] = A B        A = ((G + M) - 1 mu^T) D,  B is [Kd = p, K]
 *   transpose = 1:  out[m_limit, K] = A^T B      store is the SNP-major one (M = p),  B is [Kd = n, K]
 * i.e. `.dot` / `.T.dot` of the operator make_snp_array builds (lmdec/decomp/utils.py:280-287) as called by
 * PowerMethod._solution_step (lmdec/decomp/iter_methods.py:383), v_init (init_methods.py:33), subspace_to_V
 * (matrix_ops.py:93) and rmse_k (metrics.py:157).
 *   scale     D = 1/std per SNP, or NULL;  impute  mu per SNP (needed when has_missing);
 *   center_w  NULL = no centring; else (mu o D) per SNP  [transpose = 1 multiplies it by sum_i B_ik, transpose = 0
 *             contracts it with B]
 *   work      lmdec_apply_work_doubles(K) doubles;  image1/2, acc0/1: caller-owned scratch (lmdec_image_bytes(Kd, N)
 *             with N = 16*ceil(K*P/16);  int64 [m_limit, K])
 * transpose_flags: bit 0 = transpose; bit 1 = the statistics and plane images of B left in work / image1/2 by the
 * previous call are reused (several output row ranges of the same product, e.g. to pipeline an all-reduce);
 * bit 2 = B is float32 (else float64); bit 3 = out is float32 (else float64).  The p x K side of the iteration is
 * carried in float32 (24-bit mantissa > the 22-bit fixed point it is quantised to), the n x K side in float64.
 * bit 4 (transpose = 1 only) = also leave in `work` the column statistics (fixed-point multiplier, rank-1 shift) of `out`
 * taken as the operand of the A.dot that follows in the power iteration (x = A (A^T q), iter_methods.py:383): the
 * tensor-core epilogue reduces them while it writes `out` and the last CTA folds them (K <= 32, unsplit contraction;
 * otherwise one extra pass over out; the last double of `work` is the ticket counter of that fold: zero-fill `work` once
 * before the first call with this bit, every launch leaves it zero);  bit 5 = the column
 * statistics of B are already in `work` -- left there by the previous A^T product with bit 4 (transpose = 0: B is all
 * m_limit rows of that `out`, unmodified) or by lmdec_cholqr2_small (transpose = 1: B is its Q) -- and the pass over B
 * for them is skipped.
 * Steps: column statistics of B (max for the fixed-point scale, weighted column sum for the rank-1 term), plane
 * quantisation, lmdec_geno_matmul, lmdec_finalize.  */
int64_t lmdec_apply_work_doubles(int K);
int lmdec_operator_apply(lmdec_ctx* ctx, const uint8_t* store, int64_t M, int64_t Kd, int64_t m_limit, int64_t kd_limit,
                         int has_missing, int transpose_flags, const void* B, int64_t ldb, int K, int P,
                         const double* scale, const double* impute, const double* center_w, double* work,
                         int8_t* image1, int8_t* image2, int64_t* acc0, int64_t* acc1, void* out, int64_t ldo,
                         void* stream);

/* A^T y fused with its all-reduce over sample-row shards (the "p x k allreduce of A^T.Y" of the north-star partitioning;
 * reference: the same `A.T.dot(q)` of iter_methods.py:383, whose row chunks dask sums).  As lmdec_operator_apply with
 * transpose and float32 output, but the epilogue of the tensor-core kernel ADDS every finished 128 x K tile into all
 * ranks' copies of `out` through `out_multicast`, the NVSwitch multicast address of the same symmetric allocation
 * (`multimem.red.add.f32`): the reduction overlaps the product tile by tile and no collective follows the kernel.
 * Contract: `out` (this rank's replica, leading dimension K, K % 4 == 0) is zero on EVERY rank before ANY rank launches,
 * and is complete on a rank once every rank's launch has finished (one cross-rank barrier; the host side uses the signal
 * pads of torch's symmetric memory).  Returns 3, having launched no product, when the reduction cannot be fused (the
 * contraction would have to be split, or shared memory has no room for the staged epilogue): fall back to
 * lmdec_operator_apply + an all-reduce.  The sum order over ranks is not fixed: float32 round-off differs run to run. */
int lmdec_operator_apply_reduce(lmdec_ctx* ctx, const uint8_t* store, int64_t M, int64_t Kd, int64_t m_limit,
                                int64_t kd_limit, int has_missing, int transpose_flags, const void* B, int64_t ldb, int K,
                                int P, const double* scale, const double* impute, const double* center_w, double* work,
                                int8_t* image1, int8_t* image2, int64_t* acc0, int64_t* acc1, float* out,
                                float* out_multicast, void* stream);

/* The whole step of the power iteration in ONE call (SURVEY 8b `lmdec_apass`): the Gram operator applied to an
 * orthonormal block,   t = A^T q  (p x K float32, kept: V and the v-subspace / rmse metrics reuse it)   and
 * x = A t * inv_factor  (n_limit x K float64) -- `x = A.dot(A.T.dot(q)) / factor` of PowerMethod._solution_step
 * (lmdec/decomp/iter_methods.py:383-385; legacy: ScaledCenterArray.sym_mat_mult, scaled_legacy.py:379-399).
 * Both orientations of the store are passed (sample-major n x p for A t, SNP-major p x n for A^T q); n_limit <= n is the
 * row prefix in use (SuccessiveBatchedPowerMethod).  q_flags: bit 0 = q's column statistics are already in `work` (left
 * by lmdec_cholqr2_small / _grid), bit 1 = q is float32.  Scratch as for lmdec_operator_apply; image_q holds
 * lmdec_image_bytes(n, N), image1 / image2 lmdec_image_bytes(p, N).  Everything is enqueued on `stream`, nothing
 * returns to the host in between.  (A is still read twice per step, once per orientation: with 3 int8 planes and an
 * indicator plane the step is bound by the tensor-core / decode work, not by the 2 x n p / 4 bytes -- DESIGN.md 3.) */
int lmdec_apass(lmdec_ctx* ctx, const uint8_t* sample_store, const uint8_t* snp_store, int64_t n, int64_t p,
                int64_t n_limit, int has_missing, const void* q, int64_t ldq, int q_flags, int K, int P,
                const double* scale, const double* impute, const double* center_w, double* work, int8_t* image_q,
                int8_t* image1, int8_t* image2, int64_t* acc0, int64_t* acc1, float* t_out, double* x_out, int64_t ldx,
                double inv_factor, void* stream);

/* K x K Cholesky whitening on the device (one block, float64): G = r^T r, r upper triangular, r_inv = r^-1.
 * *flag (int32, device) = 1 when G is not numerically positive definite (pivot of the unit-diagonal-scaled matrix below
 * rank_tol, or non-finite input): the host then takes its rank-revealing path.  The K x K step of CholeskyQR2, in
 * place of dask tsqr's stacked-R QR (lmdec/decomp/iter_methods.py:382).  */
int lmdec_chol_whiten(const double* G, int K, double rank_tol, double* r_inv, double* r, int* flag, void* stream);

/* Single-GPU fast path of one CholeskyQR pass: Gram matrix of X (deterministic partials, summed inside the
 * factorisation kernel) and its Cholesky whitening, two launches.  G receives X^T X.  work: lmdec_gram_work_doubles. */
int lmdec_gram_chol(const double* X, int64_t ldx, int64_t m, int K, double rank_tol, double* G, double* r_inv,
                    double* r, int* flag, double* work, void* stream);
/* Whole CholeskyQR2 of a small local iterate (K <= 32, n*K <= lmdec_cholqr2_small_max_elems()) in one launch of one
 * 8-CTA cluster (row slices stay in shared memory, Gram partials meet through distributed shared memory; replaces the
 * reference's tsqr of the n x K iterate, lmdec/decomp/iter_methods.py:478-479): Q = orthonormalised X,
 * packed = [r2 r1 (K*K) | flag1 | flag2 | max|Q1^T Q1 - I|].  stats_work (optional): the `work` buffer of the
 * lmdec_operator_apply call that will multiply A^T by Q next, stats_lim = 2^(8P-2) for its P planes; the kernel leaves
 * Q's column statistics there and that call may pass flag 32.  */
int64_t lmdec_cholqr2_small_max_elems(void);
int lmdec_cholqr2_small(const double* X, int64_t ldx, int64_t n, int K, double rank_tol, double* Q, int64_t ldq,
                        double* packed, double* stats_work, double stats_lim, void* stream);
/* The same CholeskyQR2 for mid-size local iterates (K <= 32, n*K <= lmdec_cholqr2_grid_max_elems() = SMs * 16384) in one
 * COOPERATIVE launch over all SMs (row slices in shared memory, Gram partials through global memory, grid-wide syncs).
 * gwork: lmdec_cholqr2_grid_work_doubles(K) doubles of scratch.  Outputs as lmdec_cholqr2_small.  */
int64_t lmdec_cholqr2_grid_max_elems(void);
int64_t lmdec_cholqr2_grid_work_doubles(int K);
int lmdec_cholqr2_grid(const double* X, int64_t ldx, int64_t n, int K, double rank_tol, double* Q, int64_t ldq,
                       double* packed, double* gwork, double* stats_work, double stats_lim, void* stream);
/* packed[0..K*K) = r2 r1, then flag1, flag2, max|g2 - I|: what the host reads back from one CholeskyQR2. */
int lmdec_qr_pack(const double* r1, const double* r2, const double* g2, const int* flag1, const int* flag2, int K,
                  double* packed, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Dense float operands: the member products of `ChainedArray.dot` / `.T.dot` for dense float matrices
 * (lmdec/array/chained.py:169-177, 159-167 -> dask chunk matmuls `np.dot(chunk, x)`; BASELINE configs[4]).
 * The matrix lives in HBM as float32 in a pre-tiled store (tiles of 128 rows x 32 contraction elements, 16 KB, one
 * bulk copy each); one store per orientation, as for genotypes.  The product runs on the tensor cores as
 * error-compensated TF32 (A_hi B_hi + A_lo B_hi + A_hi B_lo, float32 accumulation): float32-level accuracy.
 * ------------------------------------------------------------------------------------------------------------- */
int64_t lmdec_dense_store_bytes(int64_t M, int64_t Kd);
int64_t lmdec_dense_image_bytes(int64_t Kd, int N);            /* N = 16 * ceil(K / 16); one image (hi or lo) */
/* src: row-major [rows, cols] (leading dimension ld, elements), dtype 0 = float32, 1 = float64 (rounded to float32).
 * transpose = 0: store of src (M = rows, contraction = cols); 1: store of src^T (M = cols, contraction = rows). */
int lmdec_dense_pack(const void* src, int dtype, int64_t rows, int64_t cols, int64_t ld, int transpose, float* store,
                     void* stream);
int lmdec_dense_unpack(const float* store, int64_t M, int64_t Kd, float* dst, int64_t ld, void* stream);
/* out[m_limit, K] (float64, leading dimension ldo) = store[0:m_limit, 0:kd_limit] * B[0:kd_limit, 0:K].
 * B: row-major, leading dimension ldb, b_dtype 0 = float32, 1 = float64; K <= 128 columns per call.
 * image_hi / image_lo: scratch of lmdec_dense_image_bytes(kd_limit, N) bytes each.  splits = 0: chosen by the library
 * (contraction ranges accumulate with float64 atomicAdd into the zeroed output).  Row prefixes (m_limit, kd_limit) are
 * the `array[0:m, :]` views of SuccessiveBatchedPowerMethod (iter_methods.py:557,573). */
int lmdec_dense_matmul(lmdec_ctx* ctx, const float* store, int64_t M, int64_t Kd, int64_t m_limit, int64_t kd_limit,
                       const void* B, int b_dtype, int64_t ldb, int K, float* image_hi, float* image_lo, double* out,
                       int64_t ldo, int splits, void* stream);


#ifdef __cplusplus
}
#endif
#endif
