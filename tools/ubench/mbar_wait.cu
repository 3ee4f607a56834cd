// Microbenchmark: cost of waiting on an mbarrier whose phase has ALREADY completed, for 512 threads at once
// (the chain kernel's epilogue does this 3-4 times per 32-column chunk).
//   mode 0: every thread try_wait      mode 1: lane 0 try_wait + __syncwarp      mode 2: every thread test_wait
//   mode 3: no wait at all (loop overhead)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "common.cuh"
using namespace dln;

__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

__global__ void k(int mode, int iters, long long* out, int nthreads_waiting) {
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  __syncthreads();
  if (threadIdx.x == 0) mbar_arrive(&bar);      // phase 0 completes: waits on parity 0 succeed immediately
  __syncthreads();
  long long t0 = clock64();
  if ((int)threadIdx.x < nthreads_waiting) {
    for (int i = 0; i < iters; ++i) {
      if (mode == 0) { while (!mbar_try_wait(&bar, 0)) {} }
      else if (mode == 1) { if ((threadIdx.x & 31) == 0) { while (!mbar_try_wait(&bar, 0)) {} } __syncwarp(); }
      else if (mode == 2) { while (!mbar_test_wait(&bar, 0)) {} }
      __syncwarp();
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

int main() {
  long long* out; cudaMalloc(&out, 148 * 8);
  const int iters = 1000;
  const char* names[4] = {"all threads try_wait", "lane 0 try_wait + syncwarp", "all threads test_wait", "no wait"};
  for (int nt : {512, 128, 32})
    for (int mode = 0; mode < 4; ++mode) {
      for (int rep = 0; rep < 2; ++rep) { k<<<148, 512>>>(mode, iters, out, nt); cudaDeviceSynchronize(); }
      long long c; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
      printf("%3d waiting threads, %-28s: %7.1f cycles per wait round (%s)\n", nt, names[mode], (double)c / iters, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
