// Probe 2: why did probe 1 see ~146 cycles per narrow-N MMA?  Separates
//   (a) instruction-issue overhead (descriptor building in the loop) from
//   (b) a dependency latency between MMAs that accumulate into the SAME TMEM accumulator,
// by round-robining over NACC independent accumulators with precomputed descriptors.
// Also: tcgen05.st x8/x16/x32 throughput with 4 and 8 warps.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "../lmdec_b200/csrc/umma.cuh"
using namespace lmdec;
#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e = (x);                                                             \
    if (e != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(2);                                                                       \
    }                                                                                \
  } while (0)

template <int N, int NACC, bool A_TMEM>
__global__ void __launch_bounds__(128, 1) mma_rate(int reps, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (16384 + 4 * N * 32) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tbase = tmem_base_s;
  {
    uint32_t r[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int c = 0; c < 64; ++c) tmem_st8(tbase + ((uint32_t)(warp * 32) << 16) + c * 8, r);
    tmem_wait_st();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (tid == 0) {
    constexpr uint32_t idesc = make_idesc_f16(128, N);
    uint64_t bd[4], ad[4];
    uint32_t at[4];
    uint8_t* sB = smem + 16384;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      bd[c] = make_smem_desc(smem_u32(sB + c * N * 32), (N / 8) * 128, 128);
      ad[c] = make_smem_desc(smem_u32(smem + c * 4096), 2048, 128);
      at[c] = tbase + 448 + c * 8;
    }
    constexpr int ACC_STRIDE = (N <= 64) ? 64 : N;
    long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll
      for (int j = 0; j < NACC; ++j) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          // consecutive MMAs go to DIFFERENT accumulators when NACC > 1
          const uint32_t d = tbase + ((c + j) % NACC) * ACC_STRIDE;
          if (A_TMEM)
            umma_ts_f16(d, at[c], bd[c], idesc, 1);
          else
            umma_ss_f16(d, ad[c], bd[c], idesc, 1);
        }
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tbase);
}

template <int N, int NACC, bool A_TMEM>
void run_rate(long long* dCyc, int sms) {
  const int smem = 16384 + 4 * 256 * 32;
  CK(cudaFuncSetAttribute(mma_rate<N, NACC, A_TMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int reps = 1024 / NACC;
  mma_rate<N, NACC, A_TMEM><<<1, 128, smem>>>(reps, dCyc);
  CK(cudaDeviceSynchronize());
  long long cyc;
  CK(cudaMemcpy(&cyc, dCyc, 8, cudaMemcpyDeviceToHost));
  const double n_mma = (double)reps * NACC * 4;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int reps2 = 65536 / NACC;
  mma_rate<N, NACC, A_TMEM><<<sms, 128, smem>>>(reps2, dCyc);
  CK(cudaEventRecord(e0));
  mma_rate<N, NACC, A_TMEM><<<sms, 128, smem>>>(reps2, dCyc);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  double flops = 2.0 * 128 * N * 16 * (double)reps2 * NACC * 4 * sms;
  printf("rate N=%3d NACC=%d A=%s : 1CTA %.1f cyc/MMA (ideal %d) | %d CTAs: %.3f ms -> %.0f TFLOP/s executed\n", N, NACC,
         A_TMEM ? "tmem" : "smem", (double)cyc / n_mma, N / 2, sms, ms, flops / ms * 1e-9);
}

template <int X, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) st_rate(int iters, long long* cycles) {
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tbase = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 256;
  uint32_t r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = tid * 8 + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint32_t col = (it * X) & 255;
    if (X == 8) {
      tmem_st8(tbase + col, r);
    } else if (X == 16) {
      asm volatile(
          "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(
              tbase + col),
          "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
          : "memory");
    } else {
      asm volatile(
          "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
          "{%1,%2,%3,%4,%5,%6,%7,%8,%1,%2,%3,%4,%5,%6,%7,%8,%1,%2,%3,%4,%5,%6,%7,%8,%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(
              tbase + col),
          "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
          : "memory");
    }
  }
  tmem_wait_st();
  __syncthreads();
  long long t1 = clock64();
  if (tid == 0) cycles[0] = t1 - t0;
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base_s);
}

template <int X, int WARPS>
void run_st(long long* dCyc) {
  st_rate<X, WARPS><<<1, WARPS * 32>>>(4096, dCyc);
  CK(cudaDeviceSynchronize());
  long long cyc;
  CK(cudaMemcpy(&cyc, dCyc, 8, cudaMemcpyDeviceToHost));
  double bytes = 4096.0 * WARPS * 32 * X * 4;
  printf("tcgen05.st x%-2d %d warps: %.1f B/cycle\n", X, WARPS, bytes / cyc);
}

int main() {
  CK(cudaSetDevice(0));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  long long* dCyc;
  CK(cudaMalloc(&dCyc, 1024 * 8));
  const int sms = prop.multiProcessorCount;
  run_rate<32, 1, true>(dCyc, sms);
  run_rate<32, 2, true>(dCyc, sms);
  run_rate<32, 4, true>(dCyc, sms);
  run_rate<48, 1, true>(dCyc, sms);
  run_rate<48, 2, true>(dCyc, sms);
  run_rate<48, 4, true>(dCyc, sms);
  run_rate<64, 1, true>(dCyc, sms);
  run_rate<64, 2, true>(dCyc, sms);
  run_rate<64, 4, true>(dCyc, sms);
  run_rate<96, 4, true>(dCyc, sms);
  run_rate<128, 2, true>(dCyc, sms);
  run_rate<256, 1, true>(dCyc, sms);
  run_rate<32, 4, false>(dCyc, sms);
  run_rate<48, 4, false>(dCyc, sms);
  run_rate<64, 4, false>(dCyc, sms);
  run_rate<128, 2, false>(dCyc, sms);
  run_rate<256, 1, false>(dCyc, sms);
  run_st<8, 4>(dCyc);
  run_st<16, 4>(dCyc);
  run_st<32, 4>(dCyc);
  run_st<8, 8>(dCyc);
  run_st<32, 8>(dCyc);
  printf("probe2 done\n");
  return 0;
}
