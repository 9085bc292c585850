// Probe 3: kind::i8 (uint8 A in TMEM x int8 B in smem -> int32), K = 32 per instruction.
//   correctness of the assumed layouts + cycles per MMA for N = 32..256.
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

// A: [128][K] uint8, B: [N][K] int8, out [128][N] int32.  K = nchunks*32
template <int N>
__global__ void __launch_bounds__(128, 1)
i8_kernel(int nchunks, int reps, const uint8_t* __restrict__ A, const int8_t* __restrict__ B, int* __restrict__ out,
          long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int K = nchunks * 32;
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  // B tile per chunk: N x 32 bytes, K-major no-swizzle: off(n,k) = (k/16)*LBO + (n/8)*128 + (n%8)*16 + k%16
  const uint32_t lbo = (N / 8) * 128;
  for (int idx = tid; idx < N * K; idx += 128) {
    int n = idx / K, k = idx % K, c = k / 32, kk = k % 32;
    smem[c * N * 32 + (kk / 16) * lbo + (n / 8) * 128 + (n % 8) * 16 + (kk % 16)] = (uint8_t)B[idx];
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tbase = tmem_base_s;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  const uint32_t a_col = 256;
  for (int c = 0; c < nchunks; ++c) {
    uint32_t r[8];
    const uint8_t* arow = A + (size_t)tid * K + c * 32;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      r[i] = arow[4 * i] | (arow[4 * i + 1] << 8) | (arow[4 * i + 2] << 16) | ((uint32_t)arow[4 * i + 3] << 24);
    tmem_st8(tbase + lane_base + a_col + c * 8, r);
  }
  tmem_wait_st();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (tid == 0) {
    constexpr uint32_t idesc = make_idesc_i8(128, N, 0, 1);
    uint64_t bd[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) bd[c] = make_smem_desc(smem_u32(smem + c * N * 32), lbo, 128);
    long long t0 = clock64();
    if (reps == 1) {
      for (int c = 0; c < nchunks; ++c) umma_ts_i8(tbase, tbase + a_col + c * 8, bd[c], idesc, c > 0 ? 1u : 0u);
    } else {
      for (int rep = 0; rep < reps; ++rep) {
#pragma unroll
        for (int c = 0; c < 4; ++c) umma_ts_i8(tbase, tbase + a_col + c * 8, bd[c], idesc, 1u);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (cycles) cycles[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  tc_fence_after_sync();
  if (blockIdx.x == 0 && out) {
    for (int n0 = 0; n0 < N; n0 += 8) {
      uint32_t r[8];
      tmem_ld8(tbase + lane_base + n0, r);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 8; ++i) out[(size_t)tid * N + n0 + i] = (int)r[i];
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tbase);
}

template <int N>
void run(const uint8_t* dA, const int8_t* dB, int* dOut, long long* dCyc, const std::vector<uint8_t>& hA,
         const std::vector<int8_t>& hB, int sms) {
  const int K = 128, nchunks = 4;
  const int smem = nchunks * N * 32;
  CK(cudaFuncSetAttribute(i8_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  i8_kernel<N><<<1, 128, smem>>>(nchunks, 1, dA, dB, dOut, dCyc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("i8 N=%d: CUDA ERROR %s\n", N, cudaGetErrorString(e));
    exit(3);
  }
  std::vector<int> hOut(128 * N);
  CK(cudaMemcpy(hOut.data(), dOut, hOut.size() * 4, cudaMemcpyDeviceToHost));
  long long bad = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      long long ref = 0;
      for (int k = 0; k < K; ++k) ref += (long long)hA[m * K + k] * (long long)hB[n * K + k];
      if (ref != hOut[m * N + n]) ++bad;
    }
  i8_kernel<N><<<1, 128, smem>>>(nchunks, 512, dA, dB, nullptr, dCyc);
  CK(cudaDeviceSynchronize());
  long long cyc;
  CK(cudaMemcpy(&cyc, dCyc, 8, cudaMemcpyDeviceToHost));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  i8_kernel<N><<<sms, 128, smem>>>(nchunks, 16384, dA, dB, nullptr, dCyc);
  CK(cudaEventRecord(e0));
  i8_kernel<N><<<sms, 128, smem>>>(nchunks, 16384, dA, dB, nullptr, dCyc);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  double ops = 2.0 * 128 * N * 32 * 4 * 16384.0 * sms;
  printf("i8 N=%3d : mismatches %lld / %d | 1CTA %.1f cyc/MMA (128xNx32; N/2 = %d) | %d CTAs %.3f ms -> %.0f TOP/s\n", N,
         bad, 128 * N, (double)cyc / (4 * 512), N / 2, sms, ms, ops / ms * 1e-9);
}

int main() {
  CK(cudaSetDevice(0));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int K = 128;
  std::vector<uint8_t> hA(128 * K);
  std::vector<int8_t> hB(256 * K);
  srand(7);
  for (auto& a : hA) a = rand() % 3;
  for (auto& b : hB) b = (int8_t)((rand() % 256) - 128);
  uint8_t* dA;
  int8_t* dB;
  int* dOut;
  long long* dCyc;
  CK(cudaMalloc(&dA, hA.size()));
  CK(cudaMalloc(&dB, hB.size()));
  CK(cudaMalloc(&dOut, 128 * 256 * 4));
  CK(cudaMalloc(&dCyc, 1024 * 8));
  CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
  const int sms = prop.multiProcessorCount;
  run<32>(dA, dB, dOut, dCyc, hA, hB, sms);
  run<64>(dA, dB, dOut, dCyc, hA, hB, sms);
  run<80>(dA, dB, dOut, dCyc, hA, hB, sms);
  run<96>(dA, dB, dOut, dCyc, hA, hB, sms);
  run<128>(dA, dB, dOut, dCyc, hA, hB, sms);
  run<256>(dA, dB, dOut, dCyc, hA, hB, sms);
  printf("probe3 done\n");
  return 0;
}
