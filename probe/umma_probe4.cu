// Probe 4: is the ~44-cycle floor of a narrow-N tcgen05.mma an ISSUE-side limit (one thread) or a datapath limit?
//   W issuer threads (lane 0 of warps 0..W-1) each stream kind::i8 128xNx32 MMAs into their own accumulator.
//   If cycles/MMA (aggregate) halves with W=2, the floor is per-issuer and a 2-issuer kernel doubles narrow-N rate.
#include <cstdio>
#include <cstdlib>
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

template <int N, int W, int M>
__global__ void __launch_bounds__(128, 1) issue_rate(int reps, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t tmem_base_s;
  __shared__ long long tstart[4], tstop[4];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (4 * 256 * 32) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1);
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
  if (lane == 0 && warp < W) {
    constexpr uint32_t idesc = make_idesc_i8(M, N, 0, 1);
    uint64_t bd[4];
    uint32_t at[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      bd[c] = make_smem_desc(smem_u32(smem + c * N * 32), (N / 8) * 128, 128);
      at[c] = tbase + 448 + c * 8;
    }
    const uint32_t d = tbase + warp * 96;
    long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll
      for (int c = 0; c < 4; ++c) umma_ts_i8(d, at[c], bd[c], idesc, 1);
    }
    umma_commit(&bar[warp]);
    mbar_wait(&bar[warp], 0);
    long long t1 = clock64();
    tstart[warp] = t0;
    tstop[warp] = t1;
  }
  __syncthreads();
  if (tid == 0) {
    long long a = tstart[0], b = tstop[0];
    for (int w = 1; w < W; ++w) {
      a = min(a, tstart[w]);
      b = max(b, tstop[w]);
    }
    cycles[blockIdx.x] = b - a;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tbase);
}

template <int N, int W, int M>
void run(long long* dCyc) {
  const int smem = 4 * 256 * 32;
  CK(cudaFuncSetAttribute(issue_rate<N, W, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int reps = 2048;
  issue_rate<N, W, M><<<1, 128, smem>>>(reps, dCyc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("N=%d W=%d M=%d: CUDA ERROR %s\n", N, W, M, cudaGetErrorString(e));
    exit(3);
  }
  long long cyc;
  CK(cudaMemcpy(&cyc, dCyc, 8, cudaMemcpyDeviceToHost));
  const double n_mma = (double)reps * 4 * W;
  printf("i8 M=%3d N=%3d issuers=%d : %.1f cycles per MMA aggregate (%.1f per issuer); N/2 = %d\n", M, N, W,
         (double)cyc / n_mma, (double)cyc / (reps * 4.0), N / 2);
}

int main() {
  CK(cudaSetDevice(0));
  long long* dCyc;
  CK(cudaMalloc(&dCyc, 1024 * 8));
  run<32, 1, 128>(dCyc);
  run<32, 2, 128>(dCyc);
  run<32, 4, 128>(dCyc);
  run<80, 1, 128>(dCyc);
  run<80, 2, 128>(dCyc);
  run<80, 4, 128>(dCyc);
  run<96, 2, 128>(dCyc);
  run<32, 1, 64>(dCyc);
  run<80, 1, 64>(dCyc);
  run<80, 2, 64>(dCyc);
  printf("probe4 done\n");
  return 0;
}
