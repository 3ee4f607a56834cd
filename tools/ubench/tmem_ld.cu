// Microbenchmark: tcgen05.ld throughput per SM for different warp counts / shapes.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "common.cuh"
using namespace dln;

__device__ __forceinline__ void ld64(uint32_t taddr, uint32_t (&r)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
      "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]),
        "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]),
        "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
      : "r"(taddr) : "memory");
}

template <int MODE>   // 0: x32 + wait each; 1: 2 x32 in flight; 2: x64 + wait
__global__ void k(int nwarps, int iters, long long* out, uint32_t* sink) {
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&tbase, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t t = tbase + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  if (warp < nwarps) {
    for (int i = 0; i < iters; ++i) {
      if (MODE == 0) {
        uint32_t v[32]; tmem_ld32(t + ((i * 32) & 255), v); tmem_ld_wait();
        acc ^= v[0] ^ v[31];
      } else if (MODE == 1) {
        uint32_t a[32], b[32]; tmem_ld32(t + ((i * 64) & 255), a); tmem_ld32(t + ((i * 64 + 32) & 255), b); tmem_ld_wait();
        acc ^= a[0] ^ b[31];
      } else {
        uint32_t v[64]; ld64(t + ((i * 64) & 255), v); tmem_ld_wait();
        acc ^= v[0] ^ v[63];
      }
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

int main() {
  long long* out; uint32_t* sink;
  cudaMalloc(&out, 1024 * 8); cudaMalloc(&sink, 148 * 512 * 4);
  const int iters = 2000;
  for (int mode = 0; mode < 3; ++mode)
    for (int nw : {1, 2, 4, 8, 16}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, 512>>>(nw, iters, out, sink);
        if (mode == 1) k<1><<<148, 512>>>(nw, iters, out, sink);
        if (mode == 2) k<2><<<148, 512>>>(nw, iters, out, sink);
        cudaDeviceSynchronize();
      }
      long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
      const double bytes = (double)nw * iters * (mode == 0 ? 4096 : 8192);
      printf("mode %d warps %2d: %lld cycles, %.1f cyc/iter, %.1f B/cycle/SM  (%s)\n", mode, nw, h, (double)h / iters, bytes / h, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
