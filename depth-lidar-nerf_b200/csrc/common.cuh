// Shared device helpers for the sm_100a kernels: warp primitives, mbarrier / bulk-copy (TMA engine) /
// tcgen05 (5th-gen tensor core + TMEM) PTX wrappers.  Everything here is inline PTX written for
// sm_100a; there is no fallback for other architectures.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <type_traits>

#define DLN_OK 0
#define DLN_EINVAL (-1)

#define DLN_CHECK_ARG(cond) \
  do {                      \
    if (!(cond)) return DLN_EINVAL; \
  } while (0)

static inline int dln_launch_status() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? DLN_OK : (int)e;
}

// Per-device "already configured" flags (cudaFuncSetAttribute and the SM count belong to a device / context, not to
// the process): slot `which` of the current device.
static inline bool& dln_device_flag(int which) {
  static bool flags[64][8] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  return flags[dev & 63][which & 7];
}
static inline int dln_sm_count() {
  static int sms[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  int& s = sms[dev & 63];
  if (!s) {
    cudaDeviceGetAttribute(&s, cudaDevAttrMultiProcessorCount, dev);
    if (s <= 0) s = 148;
  }
  return s;
}

namespace dln {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// torch.linspace(0, 1, n)[i] in fp32, same two-sided formula as ATen's CPU/CUDA kernels
// (start + step*i below the midpoint, end - step*(n-1-i) above) so that u / t_vals match bit for bit.
__device__ __forceinline__ float linspace01(int i, int n) {
  if (n <= 1) return 0.f;
  const float step = __fdiv_rn(1.0f, (float)(n - 1));
  // ATen evaluates both branches as one fused multiply-add (single rounding)
  return (i < n / 2) ? __fmul_rn(step, (float)i) : __fmaf_rn(-step, (float)(n - 1 - i), 1.0f);
}

// ------------------------------------------------------------------ counter-based random numbers
// Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3"): the draws the reference takes from
// torch.rand / torch.randn (run_nerf.py:585, run_nerf_helpers.py:509, :565) are generated inside the consuming
// kernels instead of being written to and re-read from HBM.  key = the 64-bit seed, counter = (block index of the
// element, 64-bit word w) with w = state[1] + offset: state[1] is a device-resident step counter (so a captured
// CUDA graph draws fresh numbers on every replay), `offset` identifies the tensor within the step.
struct RngRef {
  const unsigned long long* state;   // device {seed, base}; null = no in-kernel generation
  unsigned long long offset;
};
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u, k.y += 0xBB67AE85u;
  }
  return c;
}
struct RngKey {
  uint2 key;
  uint32_t w_lo, w_hi;
};
__device__ __forceinline__ RngKey rng_key(const RngRef& r) {
  RngKey k;
  const unsigned long long seed = __ldg(r.state), w = __ldg(r.state + 1) + r.offset;
  k.key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  k.w_lo = (uint32_t)w, k.w_hi = (uint32_t)(w >> 32);
  return k;
}
__device__ __forceinline__ uint4 rng_block(const RngKey& k, unsigned long long block) {
  return philox4x32_10(make_uint4((uint32_t)block, (uint32_t)(block >> 32), k.w_lo, k.w_hi), k.key);
}
// 32 random bits -> uniform in [0, 1) with 24 bits (the grid torch.rand draws fp32 from)
__device__ __forceinline__ float rng_uniform(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-08f; }
// Box-Muller on (x0, x1): two independent N(0,1) values.  u1 in (0, 1], angle in [-pi, pi).
__device__ __forceinline__ void rng_normal2(uint32_t x0, uint32_t x1, float& n0, float& n1) {
  const float u1 = (float)((x0 >> 8) + 1u) * 5.9604644775390625e-08f;
  const float r = sqrtf(-2.0f * __logf(u1));
  const float th = fmaf((float)(x1 >> 8), 3.7450704e-07f /* 2 pi / 2^24 */, -3.14159265358979f);
  float sn, cs;
  __sincosf(th, &sn, &cs);
  n0 = r * cs, n1 = r * sn;
}
// K consecutive elements e0 .. e0+K-1 of a random tensor: element e is component (e & 3) of Philox block (e >> 2);
// normal tensors turn components (0,1) and (2,3) into Box-Muller pairs.
template <int K, bool NORMAL>
__device__ __forceinline__ void rng_fill(const RngKey& key, unsigned long long e0, float (&out)[K]) {
  unsigned long long cur = ~0ull;
  float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const unsigned long long e = e0 + k, b = e >> 2;
    if (b != cur) {
      const uint4 x = rng_block(key, b);
      if (NORMAL) {
        rng_normal2(x.x, x.y, v0, v1);
        rng_normal2(x.z, x.w, v2, v3);
      } else {
        v0 = rng_uniform(x.x), v1 = rng_uniform(x.y), v2 = rng_uniform(x.z), v3 = rng_uniform(x.w);
      }
      cur = b;
    }
    const int j = (int)(e & 3);
    out[k] = j == 0 ? v0 : j == 1 ? v1 : j == 2 ? v2 : v3;
  }
}

// ------------------------------------------------------------------ stratified depths (run_nerf.py:571-593)
// z_i before the jitter: near (1 - t_i) + far t_i, or the lindisp form; t = linspace(0, 1, S)
__device__ __forceinline__ float base_z(float nr, float fr, int i, int S, int lindisp) {
  const float t = linspace01(i, S);
  if (!lindisp) return __fadd_rn(__fmul_rn(nr, __fsub_rn(1.f, t)), __fmul_rn(fr, t));
  return __fdiv_rn(1.f, __fadd_rn(__fmul_rn(__fdiv_rn(1.f, nr), __fsub_rn(1.f, t)), __fmul_rn(__fdiv_rn(1.f, fr), t)));
}
// Sample i of a ray with bounds (nr, fr): bin centre (jitter == false) or lower + (upper - lower) * t_rand with
// mids / upper / lower as the reference forms them.  One definition for the stand-alone kernel (render_kernels.cu) and for
// the MLP chain's tile prologue (mlp_chain2.cu), so the two routes give the same bits.
__device__ __forceinline__ float stratified_z_point(float nr, float fr, int i, int S, int lindisp, bool jitter, float t_rand) {
  const float zi = base_z(nr, fr, i, S, lindisp);
  if (!jitter) return zi;
  const float zl = i > 0 ? base_z(nr, fr, i - 1, S, lindisp) : zi;
  const float zr = i < S - 1 ? base_z(nr, fr, i + 1, S, lindisp) : zi;
  const float lower = i > 0 ? __fmul_rn(0.5f, __fadd_rn(zi, zl)) : zi;
  const float upper = i < S - 1 ? __fmul_rn(0.5f, __fadd_rn(zr, zi)) : zi;
  return __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), t_rand));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// one lane of a fully converged warp (keeps the surrounding code on the uniform datapath)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (-> launch error reported to the caller) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) {
      printf("dlnerf: mbarrier timeout block %d thread %d bar %p parity %u\n", blockIdx.x, threadIdx.x,
             (void*)bar, parity);
      __trap();
    }
  }
}

// ------------------------------------------------------------------ bulk async copies (TMA engine, 1-D)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read1() {   // all but the most recent bulk group have finished reading smem
  asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy writes to smem -> visible to the async proxy (UMMA operand reads, bulk S->G)
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ------------------------------------------------------------------ tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_out, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_out)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 inputs, fp32 accumulate, issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same, with the A operand ([128 x 16] bf16) read from TENSOR MEMORY: lane = row, 8 consecutive 32-bit columns
// starting at a_tmem, column c holding elements (2c, 2c+1) (validated by tools/ubench/tmem_a.cu).
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once all previously issued tcgen05.mma of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets lane (base_lane+i), columns c..c+31.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
        "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
        "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// registers -> 32 lanes x 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// Orders every later use of r[] after the preceding tmem_ld_wait when another load is already in flight
// (volatile asm statements keep their relative order; this one "redefines" the registers).
__device__ __forceinline__ void tmem_ld_pin16(uint32_t (&r)[16]) {
  asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                    "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]),
                    "+r"(r[15]));
}

__device__ __forceinline__ void tmem_ld_pin32(uint32_t (&r)[32]) {
  asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                    "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]),
                    "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]),
                    "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]),
                    "+r"(r[29]), "+r"(r[30]), "+r"(r[31]));
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout): start>>4 [0,14),
// LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout_type [61,64) (2 = SWIZZLE_128B).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor for kind::f16 (cute::UMMA::InstrDescriptor): c_format F32 (1<<4), a/b BF16
// (1<<7, 1<<10), a_major bit 15, b_major bit 16 (0 = K-major, 1 = MN-major), N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Byte offset of element (row, col) inside one "slab": a [rows x 64] bf16 panel stored as 128-byte rows,
// 8-row groups of 1024 B, 16-byte chunks XOR-swizzled by (row % 8)  (the SWIZZLE_128B canonical layout).
// The same image serves as a K-major operand (row = M/N index, col = K) and as an MN-major operand
// (row = K index, col = M/N).
__host__ __device__ __forceinline__ uint32_t slab_off(uint32_t row, uint32_t col) {
  return (row >> 3) * 1024u + (row & 7u) * 128u + ((((col >> 3) ^ row) & 7u) << 4) + ((col & 7u) << 1);
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

}  // namespace dln
