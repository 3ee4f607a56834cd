// Microbenchmark: cp.async.bulk L2->smem streaming throughput per SM (3-stage ring of 32 KB stages over an
// L2-resident 1.2 MB buffer), for different numbers of active CTAs, plus the smem->global direction.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "common.cuh"
using namespace dln;
constexpr int STAGE = 32768, NST = 3;

__global__ void __launch_bounds__(128, 1) k_load(const uint8_t* src, size_t bytes, int passes, long long* out, int stage_bytes) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + NST * STAGE);
  uint64_t* empty = full + NST;
  if (threadIdx.x == 0) { for (int i = 0; i < NST; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); } mbar_fence_init(); }
  __syncthreads();
  const int n = (int)(bytes / stage_bytes) * passes;
  long long t0 = clock64();
  if (threadIdx.x == 0) {          // producer
    uint32_t st = 0, ph = 0;
    for (int i = 0; i < n; ++i) {
      mbar_wait(&empty[st], ph ^ 1);
      mbar_expect_tx(&full[st], stage_bytes);
      bulk_g2s(smem + st * STAGE, src + (size_t)(i % (bytes / stage_bytes)) * stage_bytes, stage_bytes, &full[st]);
      if (++st == NST) st = 0, ph ^= 1;
    }
  } else if (threadIdx.x == 32) {  // consumer: just releases
    uint32_t st = 0, ph = 0;
    for (int i = 0; i < n; ++i) {
      mbar_wait(&full[st], ph);
      mbar_arrive(&empty[st]);
      if (++st == NST) st = 0, ph ^= 1;
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

__global__ void __launch_bounds__(128, 1) k_store(uint8_t* dst, int n, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  __syncthreads();
  long long t0 = clock64();
  if (threadIdx.x == 0) {
    for (int i = 0; i < n; ++i) {
      bulk_s2g(dst + ((size_t)blockIdx.x * n + i) * 16384, smem + (i & 3) * 16384, 16384);
      bulk_commit();
      if (i >= 3) asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
    }
    bulk_wait_all0();
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

int main() {
  const size_t bytes = 1179648;   // 36 stages of 32 KB
  uint8_t* src; long long* out; uint8_t* dst;
  cudaMalloc(&src, bytes); cudaMemset(src, 1, bytes); cudaMalloc(&out, 1024 * 8);
  cudaMalloc(&dst, (size_t)148 * 400 * 16384);
  const int smem = NST * STAGE + 64 + 1024;
  cudaFuncSetAttribute(k_load, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k_store, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 16384 + 1024);
  for (int sb : {32768, 16384})
  for (int grid : {1, 8, 37, 74, 148}) {
    const int passes = 20;
    for (int rep = 0; rep < 2; ++rep) { k_load<<<grid, 128, smem>>>(src, bytes, passes, out, sb); cudaDeviceSynchronize(); }
    long long h[148]; cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("load  stage %5d grid %3d: %lld cycles  %.1f B/clk/SM  %.0f B/clk chip  (%s)\n", sb, grid, mx, (double)bytes * passes / mx,
           (double)bytes * passes / mx * grid, cudaGetErrorString(cudaGetLastError()));
  }
  for (int grid : {1, 37, 148}) {
    const int n = 400;
    for (int rep = 0; rep < 2; ++rep) { k_store<<<grid, 128, 4 * 16384 + 1024>>>(dst, n, out); cudaDeviceSynchronize(); }
    long long h[148]; cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("store grid %3d: %lld cycles  %.1f B/clk/SM  %.0f B/clk chip  (%s)\n", grid, mx, (double)n * 16384 / mx, (double)n * 16384 / mx * grid,
           cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
