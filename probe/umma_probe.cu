// Hardware probe for the sm_100a tensor path used by lmdec_b200 (not part of the product).
// Answers, on a real B200:
//   1. is the TMEM A-operand layout (lane = row, 32-bit column c = k 2c,2c+1) what we assume?
//   2. which of LBO/SBO is the K-direction / MN-direction stride for a SWIZZLE_NONE K-major smem operand?
//   3. are fp16 SUBNORMAL A values (genotype code << 9 == code * 2^-15) honoured by tcgen05.mma?
//   4. how many cycles does one 128xNx16 MMA take with A from TMEM vs from shared memory, for narrow N?
//   5. tcgen05.st throughput.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o probe/umma_probe probe/umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#include "../lmdec_b200/csrc/umma.cuh"

using namespace lmdec;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e = (x);                                                               \
    if (e != cudaSuccess) {                                                            \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);   \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

struct ProbeArgs {
  int N;        // MMA N (multiple of 16)
  int nchunks;  // K chunks of 16
  int a_mode;   // 0 = A in TMEM, 1 = A in smem K-major, 2 = A in smem MN-major
  int swap;     // 0: desc.LBO = K-direction stride, desc.SBO = MN-direction stride ; 1: swapped
  int repeat;   // MMA repetitions (timing)
};

constexpr int kMaxChunks = 4;
constexpr int kMaxN = 256;

__global__ void __launch_bounds__(128, 1)
probe_kernel(ProbeArgs args, const __half* __restrict__ A,  // [128][K]
             const __half* __restrict__ B,                    // [N][K]
             float* __restrict__ out,                         // [128][N]
             long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int N = args.N, K = args.nchunks * 16;
  uint8_t* sA = smem;                                  // nchunks * 4096
  uint8_t* sB = smem + kMaxChunks * 4096;              // nchunks * N*32

  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  // strides of the canonical no-swizzle layout used by the FILL code
  const uint32_t a_kstride = 128 * 16;  // bytes between the two K-halves (k/8) of one chunk: 16 row-groups * 128 B
  const uint32_t a_mstride = 128;       // bytes between 8-row groups
  const uint32_t b_kstride = (N / 8) * 128;
  const uint32_t b_nstride = 128;
  // fill A (smem) for SS modes
  for (int idx = tid; idx < 128 * K; idx += 128) {
    int m = idx / K, k = idx % K;
    int c = k / 16, kk = k % 16;
    uint32_t off;
    if (args.a_mode == 2) {
      // MN-major: 8 mn-elements contiguous (16 B), k rows 16 B apart, mn-groups SBO apart, k-groups LBO apart
      off = c * 4096 + (kk / 8) * a_kstride + (m / 8) * a_mstride + (kk % 8) * 16 + (m % 8) * 2;
    } else {
      off = c * 4096 + (kk / 8) * a_kstride + (m / 8) * a_mstride + (m % 8) * 16 + (kk % 8) * 2;
    }
    *reinterpret_cast<__half*>(sA + off) = A[idx];
  }
  for (int idx = tid; idx < N * K; idx += 128) {
    int n = idx / K, k = idx % K;
    int c = k / 16, kk = k % 16;
    uint32_t off = c * (N * 32) + (kk / 8) * b_kstride + (n / 8) * b_nstride + (n % 8) * 16 + (kk % 8) * 2;
    *reinterpret_cast<__half*>(sB + off) = B[idx];
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tbase = tmem_base_s;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  const uint32_t d_col = 0, a_col = 256;
  if (args.a_mode == 0) {
    for (int c = 0; c < args.nchunks; ++c) {
      uint32_t r[8];
      const __half* arow = A + (size_t)tid * K + c * 16;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        __half2 h = __halves2half2(arow[2 * i], arow[2 * i + 1]);
        r[i] = *reinterpret_cast<uint32_t*>(&h);
      }
      tmem_st8(tbase + lane_base + a_col + c * 8, r);
    }
    tmem_wait_st();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();

  if (tid == 0) {
    const uint32_t idesc = make_idesc_f16(128, N, args.a_mode == 2 ? 1 : 0, 0);
    long long t0 = clock64();
    for (int rep = 0; rep < args.repeat; ++rep) {
      for (int c = 0; c < args.nchunks; ++c) {
        uint32_t lbo = args.swap ? b_nstride : b_kstride, sbo = args.swap ? b_kstride : b_nstride;
        uint64_t bdesc = make_smem_desc(smem_u32(sB + c * (N * 32)), lbo, sbo);
        uint32_t acc = (rep > 0 || c > 0) ? 1u : 0u;
        if (args.a_mode == 0) {
          umma_ts_f16(tbase + d_col, tbase + a_col + c * 8, bdesc, idesc, acc);
        } else {
          uint32_t albo = args.swap ? a_mstride : a_kstride, asbo = args.swap ? a_kstride : a_mstride;
          uint64_t adesc = make_smem_desc(smem_u32(sA + c * 4096), albo, asbo);
          umma_ss_f16(tbase + d_col, adesc, bdesc, idesc, acc);
        }
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (cycles) cycles[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  tc_fence_after_sync();
  if (blockIdx.x == 0) {
    for (int n0 = 0; n0 < N; n0 += 8) {
      uint32_t r[8];
      tmem_ld8(tbase + lane_base + d_col + n0, r);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 8; ++i) out[(size_t)tid * N + n0 + i] = __uint_as_float(r[i]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tbase);
}

// tcgen05.st throughput: 4 warps, `iters` x (x8 store) each
__global__ void __launch_bounds__(128, 1) st_kernel(int iters, long long* cycles, int with_ld) {
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tbase = tmem_base_s + ((uint32_t)(warp * 32) << 16);
  uint32_t r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = tid * 8 + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    tmem_st8(tbase + ((it * 8) & 511), r);
    if (with_ld) {
      uint32_t q[8];
      tmem_ld8(tbase + (((it + 17) * 8) & 511), q);
      tmem_wait_ld();
      r[0] ^= q[0];
    }
  }
  tmem_wait_st();
  __syncthreads();
  long long t1 = clock64();
  if (tid == 0) cycles[0] = t1 - t0;
  if (r[0] == 0x12345678) cycles[1] = 1;
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base_s);
}

static float h2f(__half h) { return __half2float(h); }

int main() {
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  printf("device: %s, SMs %d, cc %d.%d\n", prop.name, prop.multiProcessorCount, prop.major, prop.minor);
  const int smem_bytes = kMaxChunks * 4096 + kMaxChunks * kMaxN * 32;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));

  const int K = 64;
  std::vector<__half> hA(128 * K), hB(kMaxN * K);
  __half *dA, *dB;
  float* dOut;
  long long* dCyc;
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMalloc(&dOut, 128 * kMaxN * 4));
  CK(cudaMalloc(&dCyc, 1024 * 8));
  std::vector<float> hOut(128 * kMaxN);

  // ---------------- correctness ----------------
  for (int subnormal = 0; subnormal < 2; ++subnormal) {
    srand(1234);
    for (int i = 0; i < 128 * K; ++i) {
      int c = rand() % 3;
      if (subnormal) {
        unsigned short bits = (unsigned short)(c << 9);  // fp16 subnormal: c * 2^-15
        __half h;
        memcpy(&h, &bits, 2);
        hA[i] = h;
      } else {
        hA[i] = __float2half((float)((rand() % 7) - 3));
      }
    }
    for (int i = 0; i < kMaxN * K; ++i) hB[i] = __float2half((float)((rand() % 9) - 4) * (subnormal ? 64.f : 1.f));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    for (int N : {16, 48, 64}) {
      // B as [N][K]
      std::vector<__half> hBn(N * K);
      for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) hBn[n * K + k] = hB[n * K + k];
      CK(cudaMemcpy(dB, hBn.data(), hBn.size() * 2, cudaMemcpyHostToDevice));
      for (int a_mode = 0; a_mode < 3; ++a_mode) {
        for (int swap = 0; swap < 2; ++swap) {
          ProbeArgs args{N, K / 16, a_mode, swap, 1};
          CK(cudaMemset(dOut, 0xFF, 128 * kMaxN * 4));
          probe_kernel<<<1, 128, smem_bytes>>>(args, dA, dB, dOut, nullptr);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) {
            printf("correctness sub=%d N=%d a_mode=%d swap=%d: CUDA ERROR %s\n", subnormal, N, a_mode, swap,
                   cudaGetErrorString(e));
            return 3;
          }
          CK(cudaMemcpy(hOut.data(), dOut, 128 * N * 4, cudaMemcpyDeviceToHost));
          double maxerr = 0, maxref = 0;
          for (int m = 0; m < 128; ++m)
            for (int n = 0; n < N; ++n) {
              double ref = 0;
              for (int k = 0; k < K; ++k) ref += (double)h2f(hA[m * K + k]) * (double)h2f(hBn[n * K + k]);
              double err = fabs(ref - (double)hOut[m * N + n]);
              if (!(err <= 1e30)) err = 1e30;
              maxerr = fmax(maxerr, err);
              maxref = fmax(maxref, fabs(ref));
            }
          printf("correctness subnormalA=%d N=%3d a_mode=%d(%s) swap=%d : max_abs_err=%.6g (max |ref| %.6g) %s\n",
                 subnormal, N, a_mode, a_mode == 0 ? "tmem" : (a_mode == 1 ? "smemK" : "smemMN"), swap, maxerr, maxref,
                 maxerr <= 1e-6 * fmax(maxref, 1e-30) ? "MATCH" : "mismatch");
        }
      }
    }
  }

  // ---------------- timing: one CTA ----------------
  memset(hB.data(), 0, hB.size() * 2);
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  for (int a_mode = 0; a_mode < 2; ++a_mode) {
    for (int N : {16, 32, 48, 64, 96, 128, 256}) {
      ProbeArgs args{N, 4, a_mode, 0, 256};
      probe_kernel<<<1, 128, smem_bytes>>>(args, dA, dB, dOut, dCyc);
      CK(cudaDeviceSynchronize());
      long long cyc;
      CK(cudaMemcpy(&cyc, dCyc, 8, cudaMemcpyDeviceToHost));
      printf("timing 1CTA a_mode=%d N=%3d : %.1f cycles/MMA (128xNx16), ideal N/2 = %d\n", a_mode, N,
             (double)cyc / (4 * 256), N / 2);
    }
  }
  // ---------------- timing: all SMs ----------------
  {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int a_mode = 0; a_mode < 2; ++a_mode) {
      for (int N : {32, 48, 64, 256}) {
        ProbeArgs args{N, 4, a_mode, 0, 8192};
        probe_kernel<<<prop.multiProcessorCount, 128, smem_bytes>>>(args, dA, dB, dOut, dCyc);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        probe_kernel<<<prop.multiProcessorCount, 128, smem_bytes>>>(args, dA, dB, dOut, dCyc);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        double flops = 2.0 * 128 * N * 16 * 4 * 8192.0 * prop.multiProcessorCount;
        printf("timing %dCTA a_mode=%d N=%3d : %.3f ms, %.1f TFLOP/s executed\n", prop.multiProcessorCount, a_mode, N,
               ms, flops / ms * 1e-9);
      }
    }
  }
  // ---------------- tcgen05.st throughput ----------------
  for (int with_ld = 0; with_ld < 2; ++with_ld) {
    st_kernel<<<1, 128>>>(4096, dCyc, with_ld);
    CK(cudaDeviceSynchronize());
    long long cyc;
    CK(cudaMemcpy(&cyc, dCyc, 8, cudaMemcpyDeviceToHost));
    printf("tcgen05.st x8 (4 warps, 4 KB per round%s): %.1f cycles per round -> %.1f B/cycle\n",
           with_ld ? ", + dependent ld" : "", (double)cyc / 4096, 4096.0 * 4096 / cyc);
  }
  printf("probe done\n");
  return 0;
}
