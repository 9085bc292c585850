// Micro-probe: latency of one mbarrier "hop" (arrive by one warp -> the waiting warp resumes) inside a 704-thread CTA,
// as a function of the wait flavour and of how many OTHER warps are blocked on barriers at the same time.
// The product kernel's decoded-operand ring makes two such hops per stage round trip (profiles/r02/).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe/mbar_hop_probe probe/mbar_hop_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void commit_arrive(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
template <int MODE>
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  if (MODE == 0)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity), "r"(0x00100000u) : "memory");
  else if (MODE == 1)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
template <int MODE>
__device__ __forceinline__ void wait(uint32_t bar, uint32_t parity) {
  while (!try_wait<MODE>(bar, parity)) {
    if (MODE == 3) __nanosleep(20);
  }
}

// warp 0 <-> warp 1 ping-pong; warps 2 .. 2+n_blocked-1 sit blocked on their own barrier until the end
template <int MODE, bool COMMIT>
__global__ void __launch_bounds__(704, 1) hop_kernel(int iters, int n_blocked, int one_lane, long long* out) {
  __shared__ uint64_t bars[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t b0 = smem_u32(bars);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 32; ++i) mbar_init(b0 + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint32_t A = b0, B = b0 + 8;
  if (warp == 0) {
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (lane == 0) mbar_arrive(A);
      if (!one_lane || lane == 0) wait<MODE>(B, i & 1);
      __syncwarp();
    }
    long long t1 = clock64();
    if (lane == 0) {
      out[blockIdx.x] = t1 - t0;
      for (int w = 0; w < n_blocked; ++w) mbar_arrive(b0 + 8 * (2 + w));
    }
  } else if (warp == 1) {
    for (int i = 0; i < iters; ++i) {
      if (!one_lane || lane == 0) wait<MODE>(A, i & 1);
      __syncwarp();
      if (lane == 0) {
        if (COMMIT) commit_arrive(B);
        else mbar_arrive(B);
      }
    }
  } else if (warp - 2 < n_blocked) {
    wait<MODE>(b0 + 8 * warp, 0);
  }
}

template <int MODE, bool COMMIT>
void run(const char* name, long long* d_out) {
  const int iters = 2000;
  for (int one_lane = 0; one_lane < 2; ++one_lane)
    for (int nb : {0, 4, 12, 20}) {
      hop_kernel<MODE, COMMIT><<<148, 704>>>(iters, nb, one_lane, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
      long long h[148];
      cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
      double avg = 0;
      for (int i = 0; i < 148; ++i) avg += (double)h[i];
      avg /= 148.0 * iters;
      printf("%-34s %s blocked warps=%2d : %7.1f cycles per round trip (2 hops)\n", name,
             one_lane ? "lane-0 waits " : "warp waits   ", nb, avg);
    }
}

// Ring skeleton: NP producer warps fill a stage (arrive on full[s], count NP), NI issuer warps consume it (arrive or
// tcgen05.commit on empty[s], count NI); S stages.  Cycles per stage hand-over, nothing else in the loops.
template <bool COMMIT, bool ELECT>
__global__ void __launch_bounds__(704, 1) ring_kernel(int iters, int NP, int NI, int S, int named, long long* out) {
  __shared__ uint64_t bars[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t b0 = smem_u32(bars);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) {
      mbar_init(b0 + 8 * i, named ? 1 : NP);          // full[i]
      mbar_init(b0 + 64 + 8 * i, NI);                  // empty[i]
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  long long t0 = clock64();
  if (warp < NP) {
    uint32_t s = 0, ph = 0;
    for (int i = 0; i < iters; ++i) {
      wait<0>(b0 + 64 + 8 * s, ph ^ 1);
      if (named) {
        // rendezvous of the producer warps on a hardware named barrier, then ONE mbarrier arrive
        asm volatile("bar.sync 1, %0;" ::"r"(NP * 32) : "memory");
        if (warp == 0 && lane == 0) mbar_arrive(b0 + 8 * s);
      } else {
        __syncwarp();
        if (lane == 0) mbar_arrive(b0 + 8 * s);
      }
      if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
    }
  } else if (warp >= 16 && warp < 16 + NI) {
    uint32_t s = 0, ph = 0;
    for (int i = 0; i < iters; ++i) {
      wait<0>(b0 + 8 * s, ph);
      if (ELECT) {
        uint32_t pred;
        asm volatile("{\n\t.reg .pred P1;\n\t.reg .b32 rx;\n\telect.sync rx|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
        if (pred) { if (COMMIT) commit_arrive(b0 + 64 + 8 * s); else mbar_arrive(b0 + 64 + 8 * s); }
        __syncwarp();
      } else if (lane == 0) {
        if (COMMIT) commit_arrive(b0 + 64 + 8 * s); else mbar_arrive(b0 + 64 + 8 * s);
      }
      if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

template <bool COMMIT, bool ELECT>
void run_ring(const char* name, long long* d_out) {
  const int iters = 4000;
  for (int named = 0; named < 2; ++named)
    for (int S : {2, 3, 4})
      for (int NI : {1, 2})
        for (int NP : {1, 4, 8}) {
          if (named && NP == 1) continue;
          ring_kernel<COMMIT, ELECT><<<148, 704>>>(iters, NP, NI, S, named, d_out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
          long long h[148];
          cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
          double avg = 0;
          for (int i = 0; i < 148; ++i) avg += (double)h[i];
          avg /= 148.0 * iters;
          printf("ring %-28s %s stages=%d issuers=%d producers=%d : %7.1f cycles per stage hand-over\n", name,
                 named ? "bar.sync+1 arrive" : "arrive per warp  ", S, NI, NP, avg);
        }
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 148 * sizeof(long long));
  if (getenv("HOP_ALL")) {
    run<0, false>("try_wait + 1ms hint, arrive", d_out);
    run<1, false>("try_wait, arrive", d_out);
    run<2, false>("test_wait spin, arrive", d_out);
    run<3, false>("test_wait + nanosleep(20), arrive", d_out);
    run<0, true>("try_wait + 1ms hint, tcgen05.commit", d_out);
    run<1, true>("try_wait, tcgen05.commit", d_out);
    run<2, true>("test_wait spin, tcgen05.commit", d_out);
  }
  run_ring<false, false>("mbarrier.arrive, lane 0", d_out);
  run_ring<true, true>("tcgen05.commit, elect.sync", d_out);
  printf("hop probe done\n");
  return 0;
}
