// Microbenchmark + known-answer test: tcgen05.mma with the A operand read from TENSOR MEMORY (written there by
// tcgen05.st, the way an epilogue would hand its bf16 activations to the next layer) against A read from shared
// memory, with and without a concurrent stream of bulk global->shared copies competing for the shared-memory port.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -I depth-lidar-nerf_b200/csrc tools/ubench/tmem_a.cu -o gpurun_out/tmem_a
//
// D[128 x 256] = A[128 x 256] * B[256 x 256]^T with small integers (exact in bf16 / fp32).
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "common.cuh"
using namespace dln;

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
      "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
      "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__host__ __device__ inline float a_val(int m, int k) { return (float)(((m + 2 * k) % 7) - 3); }
__host__ __device__ inline float b_val(int n, int k) { return (float)(((3 * n + k) % 5) - 2); }

constexpr int kBSlab = 32768, kASlab = 16384, kScratch = 32768;

// mode bit 0: A from TMEM (else shared memory); bit 1: concurrent bulk global->shared copies;
// bit 2: the layer issued as two N=128 halves (B rows [0,128) then [128,256): +16 KB inside each 32 KB K slab)
__global__ void __launch_bounds__(192, 1) k(int mode, int reps, float* out, long long* cyc, const uint8_t* gsrc) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* sB = smem;                    // 4 x 32 KB
  uint8_t* sA = smem + 4 * kBSlab;       // 4 x 16 KB
  uint8_t* scratch = sA + 4 * kASlab;    // 32 KB
  __shared__ uint64_t bar_mma, bar_ld;
  __shared__ uint32_t tbase_s;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool ts = mode & 1, loads = mode & 2, halves = mode & 4;
  if (threadIdx.x == 0) {
    mbar_init(&bar_mma, 1);
    mbar_init(&bar_ld, 1);
    mbar_fence_init();
    stop = 0;
  }
  if (warp == 4) {
    tmem_alloc(&tbase_s, 512);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < 256 * 256; i += blockDim.x) {
    const int n = i >> 8, kq = i & 255;
    *reinterpret_cast<__nv_bfloat16*>(sB + (kq >> 6) * kBSlab + slab_off(n, kq & 63)) = __float2bfloat16(b_val(n, kq));
  }
  for (int i = threadIdx.x; i < 128 * 256; i += blockDim.x) {
    const int m = i >> 8, kq = i & 255;
    *reinterpret_cast<__nv_bfloat16*>(sA + (kq >> 6) * kASlab + slab_off(m, kq & 63)) = __float2bfloat16(a_val(m, kq));
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tbase_s;
  if (warp < 4) {  // A -> tensor memory columns [0, 128): lane = row m, column c holds (A[m][2c], A[m][2c+1])
    const int m = warp * 32 + lane;
    const uint32_t t = tbase + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = pack_bf16(a_val(m, 2 * (c0 + i)), a_val(m, 2 * (c0 + i) + 1));
      tmem_st32(t + c0, v);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 4) {
    const uint64_t desc_k = umma_desc_sw128(0, 16, 1024);
    const uint32_t idesc = umma_idesc_bf16(128, 256, 0, 0);
    const uint32_t d_tmem = tbase + 256;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (elect_one()) {
        if (!halves) {
#pragma unroll
        for (int ks = 0; ks < 16; ++ks) {
          const uint64_t bd = (desc_k | (uint64_t)(smem_u32(sB + (ks >> 2) * kBSlab) >> 4)) + 2 * (ks & 3);
          if (ts) {
            umma_bf16_ts(d_tmem, tbase + ks * 8, bd, idesc, ks != 0);
          } else {
            const uint64_t ad = (desc_k | (uint64_t)(smem_u32(sA + (ks >> 2) * kASlab) >> 4)) + 2 * (ks & 3);
            umma_bf16(d_tmem, ad, bd, idesc, ks != 0);
          }
        }
        } else {
          const uint32_t idesc_h = umma_idesc_bf16(128, 128, 0, 0);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int ks = 0; ks < 16; ++ks) {
              const uint64_t bd = (desc_k | (uint64_t)(smem_u32(sB + (ks >> 2) * kBSlab + h * 16384) >> 4)) + 2 * (ks & 3);
              if (ts) {
                umma_bf16_ts(d_tmem + 128 * h, tbase + ks * 8, bd, idesc_h, ks != 0);
              } else {
                const uint64_t ad = (desc_k | (uint64_t)(smem_u32(sA + (ks >> 2) * kASlab) >> 4)) + 2 * (ks & 3);
                umma_bf16(d_tmem + 128 * h, ad, bd, idesc_h, ks != 0);
              }
            }
          }
        }
        umma_commit(&bar_mma);
      }
      __syncwarp();
      mbar_wait(&bar_mma, r & 1);
    }
    long long t1 = clock64();
    if (lane == 0) cyc[blockIdx.x] = t1 - t0, stop = 1;
    tc_fence_before();
  } else if (warp == 5 && loads && lane == 0) {
    // 32 KB bulk copies back to back into the scratch area while the MMAs run
    uint32_t ph = 0;
    while (!stop) {
      mbar_expect_tx(&bar_ld, kScratch);
      bulk_g2s(scratch, gsrc + (size_t)blockIdx.x * kScratch, kScratch, &bar_ld);
      mbar_wait(&bar_ld, ph);
      ph ^= 1;
    }
  }
  __syncthreads();
  tc_fence_after();
  if (warp < 4) {
    const int m = warp * 32 + lane;
    const uint32_t t = tbase + ((uint32_t)(warp * 32) << 16) + 256;
    for (int c0 = 0; c0 < 256; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(t + c0, v);
      tmem_ld_wait();
      if (blockIdx.x == 0)
        for (int i = 0; i < 32; ++i) out[m * 256 + c0 + i] = __uint_as_float(v[i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tbase, 512);
}

int main() {
  float* out;
  long long* cyc;
  uint8_t* gsrc;
  cudaMalloc(&out, 128 * 256 * 4);
  cudaMalloc(&cyc, 148 * 8);
  cudaMalloc(&gsrc, 148 * kScratch);
  cudaMemset(gsrc, 0, 148 * kScratch);
  const size_t smem = 4 * kBSlab + 4 * kASlab + kScratch + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  static float h[128 * 256];
  const int reps = 2000;
  const char* names[8] = {"A smem", "A tmem", "A smem + bulk loads", "A tmem + bulk loads", "A smem, 2 x N128", "A tmem, 2 x N128", "A smem + loads, 2 x N128", "A tmem + loads, 2 x N128"};
  for (int mode = 0; mode < 8; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      k<<<148, 192, smem>>>(mode, reps, out, cyc, gsrc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("mode %d: %s\n", mode, cudaGetErrorString(e));
        return 1;
      }
    }
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    int bad = 0;
    double maxerr = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 256; ++n) {
        double ref = 0;
        for (int kk = 0; kk < 256; ++kk) ref += (double)a_val(m, kk) * b_val(n, kk);
        const double err = fabs(ref - h[m * 256 + n]);
        if (err > maxerr) maxerr = err;
        if (err > 1e-3 && bad++ < 4) printf("   mismatch m=%d n=%d got %f want %f\n", m, n, h[m * 256 + n], ref);
      }
    printf("%-22s: %7.1f cycles per M128 N256 K16 of work (one layer per commit, 148 SMs), KAT %s (max err %.3g, %d bad)\n",
           names[mode], (double)c / reps / 16.0, bad ? "FAIL" : "ok", maxerr, bad);
  }
  return 0;
}
