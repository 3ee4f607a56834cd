// Semantic head of the NeRF module (run_nerf_helpers.py:107-111, :126-127) and the per-ray semantic logits of
// raw2outputs (:586-589), plus the cross-entropy of the training loop (run_nerf.py:1541-1548).
//
// The head is two Linear layers WITHOUT an activation behind feature_linear, which has none either, so per point
//     semantic = W_s2 (W_s1 (W_f h + b_f) + b_s1) + b_s2 = Sw h + sc,   Sw = W_s2 W_s1 W_f  [K x 256],
// with h the last hidden activation of the trunk -- the very slabs the MLP chain already keeps for wgrad.  And the
// reference sums the logits UNWEIGHTED over the samples of a ray (:589), so per ray
//     sem_preds = Sw (sum_s h_s) + S sc.
// The head therefore never enters the tensor-core chain: one HBM-bound pass over the kept activations (512 B per
// point) forms Hsum = sum_s h_s per ray, and everything else is [rays x 256] x [256 x K] work:
//     forward   sem_head_fwd : Hsum, sem_preds                       (S = 1 gives per-point logits for `raw`)
//     backward  sem_head_bwd : G = dsem Sw (added to dH of the last trunk layer by the dgrad chain's epilogue),
//                              dSw += dsem^T Hsum, dsc += S sum dsem
//     fold / unfold          : Sw, sc from the six parameter tensors after a weight update; dSw, dsc back into
//                              their gradients (and into feature_linear's) after a backward pass.
// K <= 32 classes (KITTI-360: 19).
#include "common.cuh"
#include "../../include/dlnerf_b200.h"
#include <math.h>
#include <stdlib.h>

namespace {

constexpr int kW = 256;       // trunk width
constexpr int kHid = 128;     // hidden width of the semantic head (W/2)
constexpr int kMaxK = DLN_SEM_MAX_CLASSES;

// ------------------------------------------------------------------------------------------------
// fold: A = W_s1 W_f [128 x 256], a = W_s1 b_f + b_s1 [128];  Sw = W_s2 A [K x 256], sc = W_s2 a + b_s2 [K]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sem_fold_a_kernel(float* __restrict__ flat, DlnSemOffsets o) {
  __shared__ float wrow[kW];
  __shared__ float red[8];
  const int i = blockIdx.x, j = threadIdx.x;
  wrow[j] = flat[o.w_s1 + (long long)i * kW + j];
  __syncthreads();
  const float* wf = flat + o.w_f;
  float acc = 0.f;
#pragma unroll 8
  for (int k = 0; k < kW; ++k) acc = fmaf(wrow[k], wf[k * kW + j], acc);
  flat[o.A + (long long)i * kW + j] = acc;
  const float b = dln::warp_sum(wrow[j] * flat[o.b_f + j]);
  if ((j & 31) == 0) red[j >> 5] = b;
  __syncthreads();
  if (j == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    flat[o.a + i] = t + flat[o.b_s1 + i];
  }
}

__global__ void __launch_bounds__(256) sem_fold_s_kernel(float* __restrict__ flat, DlnSemOffsets o) {
  __shared__ float wrow[kHid];
  __shared__ float red[8];
  const int k = blockIdx.x, j = threadIdx.x;
  if (j < kHid) wrow[j] = flat[o.w_s2 + (long long)k * kHid + j];
  __syncthreads();
  const float* A = flat + o.A;
  float acc = 0.f;
#pragma unroll 8
  for (int i = 0; i < kHid; ++i) acc = fmaf(wrow[i], A[i * kW + j], acc);
  flat[o.Sw + (long long)k * kW + j] = acc;
  const float b = dln::warp_sum(j < kHid ? wrow[j] * flat[o.a + j] : 0.f);
  if ((j & 31) == 0) red[j >> 5] = b;
  __syncthreads();
  if (j == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    flat[o.sc + k] = t + flat[o.b_s2 + k];
  }
}

// ------------------------------------------------------------------------------------------------
// unfold (after one backward call; g[o.Sw], g[o.sc] hold dSw, dsc of that call):
//   launch 1: dW_s2[k][i] += sum_j dSw[k][j] A[i][j] + dsc[k] a[i];  db_s2 += dsc
//             dA[i][j] = sum_k W_s2[k][i] dSw[k][j]  -> g[o.A];      da[i] = sum_k W_s2[k][i] dsc[k] -> g[o.a]
//   launch 2: dW_s1[i][m] += sum_j dA[i][j] W_f[m][j] + da[i] b_f[m];  db_s1 += da
//             dW_f[m][j] += sum_i W_s1[i][m] dA[i][j];                db_f[m] += sum_i W_s1[i][m] da[i]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sem_unfold1_kernel(const float* __restrict__ flat, float* __restrict__ g,
                                                          DlnSemOffsets o) {
  __shared__ float sh[kW];
  const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int K = o.K;
  if (b < K) {                    // row k of dW_s2: a warp per i, lanes over j
    const int k = b;
    sh[t] = g[o.Sw + (long long)k * kW + t];
    __syncthreads();
    const float dk = g[o.sc + k];
    for (int i = warp; i < kHid; i += 8) {
      const float* Ai = flat + o.A + (long long)i * kW;
      float acc = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) acc = fmaf(sh[q * 32 + lane], __ldg(Ai + q * 32 + lane), acc);
      acc = dln::warp_sum(acc);
      if (lane == 0) g[o.w_s2 + (long long)k * kHid + i] += acc + dk * flat[o.a + i];
    }
    if (t == 0) g[o.b_s2 + k] += dk;
  } else if (b < K + kHid) {      // row i of dA
    const int i = b - K;
    if (t < K) sh[t] = flat[o.w_s2 + (long long)t * kHid + i];
    __syncthreads();
    float acc = 0.f, da = 0.f;
    for (int k = 0; k < K; ++k) {
      acc = fmaf(sh[k], g[o.Sw + (long long)k * kW + t], acc);
      da = fmaf(sh[k], g[o.sc + k], da);
    }
    g[o.A + (long long)i * kW + t] = acc;
    if (t == 0) g[o.a + i] = da;
  }
}

__global__ void __launch_bounds__(256) sem_unfold2_kernel(const float* __restrict__ flat, float* __restrict__ g,
                                                          DlnSemOffsets o) {
  __shared__ float sh[kW];
  __shared__ float out[kW];
  const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (b < kHid) {                 // row i of dW_s1: (W_f dA_i)[m] + da_i b_f[m]; a warp per m, lanes over j
    const int i = b;
    sh[t] = g[o.A + (long long)i * kW + t];
    __syncthreads();
    const float dai = g[o.a + i];
#pragma unroll 4
    for (int m = warp; m < kW; m += 8) {
      const float* wf = flat + o.w_f + (long long)m * kW;
      float acc = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) acc = fmaf(sh[q * 32 + lane], __ldg(wf + q * 32 + lane), acc);
      acc = dln::warp_sum(acc);
      if (lane == 0) out[m] = acc;
    }
    __syncthreads();
    g[o.w_s1 + (long long)i * kW + t] += out[t] + dai * __ldg(flat + o.b_f + t);
    if (t == 0) g[o.b_s1 + i] += dai;
  } else if (b < kHid + kW) {     // row m of dW_f: sum_i W_s1[i][m] dA[i][:]
    const int m = b - kHid;
    if (t < kHid) sh[t] = flat[o.w_s1 + (long long)t * kW + m];
    __syncthreads();
    float acc = 0.f;
#pragma unroll 16
    for (int i = 0; i < kHid; ++i) acc = fmaf(sh[i], g[o.A + (long long)i * kW + t], acc);
    g[o.w_f + (long long)m * kW + t] += acc;
  } else {                        // db_f
    if (t < kHid) sh[t] = g[o.a + t];
    __syncthreads();
    float acc = 0.f;
#pragma unroll 16
    for (int i = 0; i < kHid; ++i) acc = fmaf(__ldg(flat + o.w_s1 + (long long)i * kW + t), sh[i], acc);
    g[o.b_f + t] += acc;
  }
}

// ------------------------------------------------------------------------------------------------
// forward: a warp per group of S consecutive points (a ray; S = 1: a point).  Lane l owns features [8l, 8l+8):
// one 16-byte chunk of the point's 128-byte row in slab l/8 of the kept activation (SWIZZLE_128B image, see
// DESIGN.md section 2), i.e. a warp reads the four rows of a point as 4 x 128 contiguous bytes.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void acc_bf16x8(float (&h)[8], const uint4 v) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[2 * i] += __uint_as_float(w[i] << 16);
    h[2 * i + 1] += __uint_as_float(w[i] & 0xffff0000u);
  }
}

// WPR warps share a group: each sums a quarter of the samples (4x the loads in flight per ray, which is what a
// latency-bound stream of 16-byte loads needs), the partial sums meet in shared memory.
template <int WPR>
__global__ void __launch_bounds__(256)
    sem_head_fwd_kernel(const uint8_t* __restrict__ stash, int fwd_slots, int h_slot, long long P, int S,
                        const float* __restrict__ Sw, const float* __restrict__ sc, int K,
                        float* __restrict__ hsum, float* __restrict__ out, int out_ld, long long n_groups) {
  constexpr int GPB = 8 / WPR;                       // groups per block
  __shared__ __align__(16) float part[WPR > 1 ? 8 : 1][kW];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int gib = wib / WPR, sub = wib % WPR;        // group in block, share of the group's samples
  const size_t tile_bytes = (size_t)fwd_slots * DLN_SLAB_BYTES;
  const uint8_t* base = stash + (size_t)(h_slot + (lane >> 3)) * DLN_SLAB_BYTES;
  const uint32_t chunk = lane & 7;
  const int per = (S + WPR - 1) / WPR;
  for (long long gb = (long long)blockIdx.x * GPB; gb < n_groups; gb += (long long)gridDim.x * GPB) {   // block-uniform
    const long long g = gb + gib;
    float h[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const long long q0 = g * S;
    long long q1 = q0 + S;
    if (q1 > P) q1 = P;
    long long p0 = q0 + (long long)sub * per, p1 = p0 + per;
    if (p1 > q1) p1 = q1;
    if (g >= n_groups) p1 = p0;
    auto addr = [&](long long p) {
      const uint32_t r = (uint32_t)(p & (DLN_TILE_ROWS - 1));
      return reinterpret_cast<const uint4*>(base + (size_t)(p >> 7) * tile_bytes + (r >> 3) * 1024u + (r & 7u) * 128u +
                                            (((chunk ^ r) & 7u) << 4));
    };
    long long p = p0;
    for (; p < p1 && (p & 7); ++p) acc_bf16x8(h, __ldg(addr(p)));
    // 8 consecutive points from a multiple of 8 are the 8 rows of one 1 KB swizzle atom: row i of the atom keeps this
    // lane's chunk at i * 128 + ((chunk ^ i) << 4) -- constant offsets, so the eight 16-byte loads go out back to back
    for (; p + 8 <= p1; p += 8) {
      const uint8_t* atom = base + (size_t)(p >> 7) * tile_bytes + (((uint32_t)p & 127u) >> 3) * 1024u;
      uint4 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = __ldg(reinterpret_cast<const uint4*>(atom + i * 128 + (((chunk ^ i) & 7u) << 4)));
#pragma unroll
      for (int i = 0; i < 8; ++i) acc_bf16x8(h, v[i]);
    }
    for (; p < p1; ++p) acc_bf16x8(h, __ldg(addr(p)));
    if (WPR > 1) {
      float4* ps = reinterpret_cast<float4*>(&part[wib][8 * lane]);
      ps[0] = make_float4(h[0], h[1], h[2], h[3]);
      ps[1] = make_float4(h[4], h[5], h[6], h[7]);
      __syncthreads();
      // every warp of the group takes the total: the K logits are then split over the WPR warps
      float4 a = *reinterpret_cast<const float4*>(&part[gib * WPR][8 * lane]);
      float4 b = *reinterpret_cast<const float4*>(&part[gib * WPR][8 * lane + 4]);
#pragma unroll
      for (int w = 1; w < WPR; ++w) {
        const float4 c = *reinterpret_cast<const float4*>(&part[gib * WPR + w][8 * lane]);
        const float4 d = *reinterpret_cast<const float4*>(&part[gib * WPR + w][8 * lane + 4]);
        a.x += c.x, a.y += c.y, a.z += c.z, a.w += c.w, b.x += d.x, b.y += d.y, b.z += d.z, b.w += d.w;
      }
      h[0] = a.x, h[1] = a.y, h[2] = a.z, h[3] = a.w, h[4] = b.x, h[5] = b.y, h[6] = b.z, h[7] = b.w;
      __syncthreads();                       // part[] is rewritten by the next round
    }
    if (g >= n_groups) continue;
    if (hsum != nullptr && sub == 0) {
      float4* hs = reinterpret_cast<float4*>(hsum + (size_t)g * kW + 8 * lane);
      hs[0] = make_float4(h[0], h[1], h[2], h[3]);
      hs[1] = make_float4(h[4], h[5], h[6], h[7]);
    }
    if (out != nullptr) {
      const float cnt = (float)(q1 - q0);
      for (int k = sub; k < K; k += WPR) {
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(Sw + (size_t)k * kW + 8 * lane));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(Sw + (size_t)k * kW + 8 * lane + 4));
        float acc = h[0] * w0.x;
        acc = fmaf(h[1], w0.y, acc), acc = fmaf(h[2], w0.z, acc), acc = fmaf(h[3], w0.w, acc);
        acc = fmaf(h[4], w1.x, acc), acc = fmaf(h[5], w1.y, acc), acc = fmaf(h[6], w1.z, acc), acc = fmaf(h[7], w1.w, acc);
        acc = dln::warp_sum(acc);
        if (lane == 0) out[(size_t)g * out_ld + k] = acc + cnt * __ldg(sc + k);
      }
    }
  }
}

// Rays with S % 32 == 0 (every shipped config: 64 coarse / 128 fine samples): the 32-row blocks of a ray are 4 KB
// contiguous in each of the four slabs, so they are staged with bulk copies (TMA engine) instead of per-lane 16-byte
// loads -- a warp keeps two 16 KB blocks in flight without holding a single register for them.  Four warps per CTA,
// one CTA per SM (128 KB of staging), each warp walks its own stream of (ray, block) items: wait for the block,
// add its 32 rows (conflict-free swizzled LDS.128), hand the buffer back to the copy engine, and at the end of a ray
// write Hsum and the K logits while the next ray's blocks are already landing.
constexpr int kBulkWarps = 4;                    // measured: 7 warps (224 KB of staging) are 3 % slower than 4
constexpr int kBulkBlockBytes = 4 * 4096;            // 32 rows x 128 B in each of the 4 slabs
constexpr size_t kBulkSmemBytes = (size_t)kBulkWarps * 2 * kBulkBlockBytes + kBulkWarps * 2 * sizeof(uint64_t) + 128;

__global__ void __launch_bounds__(kBulkWarps * 32, 1)
    sem_head_fwd_bulk_kernel(const uint8_t* __restrict__ stash, int fwd_slots, int h_slot, int S,
                             const float* __restrict__ Sw, const float* __restrict__ sc, int K,
                             float* __restrict__ hsum, float* __restrict__ out, int out_ld, long long n_groups) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((128u - (dln::smem_u32(smem_raw) & 127u)) & 127u);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  uint8_t* buf = smem + (size_t)wib * 2 * kBulkBlockBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kBulkWarps * 2 * kBulkBlockBytes) + wib * 2;
  if (lane == 0) {
    dln::mbar_init(&bars[0], 1), dln::mbar_init(&bars[1], 1);
    dln::mbar_fence_init();
  }
  __syncthreads();
  const long long total_warps = (long long)gridDim.x * kBulkWarps;
  const long long gw = (long long)blockIdx.x * kBulkWarps + wib;
  const int n = S >> 5;                                            // blocks per ray
  const long long my_rays = gw < n_groups ? (n_groups - gw + total_warps - 1) / total_warps : 0;
  const long long n_items = my_rays * n;
  const size_t tile_bytes = (size_t)fwd_slots * DLN_SLAB_BYTES;
  auto issue = [&](long long t, int b) {                           // lane 0 only
    const long long ray = gw + (t / n) * total_warps;
    const long long p0 = ray * S + (t % n) * 32;
    const uint8_t* src = stash + (size_t)(p0 >> 7) * tile_bytes + (size_t)h_slot * DLN_SLAB_BYTES +
                         (((uint32_t)p0 & 127u) >> 3) * 1024u;
    dln::mbar_expect_tx(&bars[b], kBulkBlockBytes);
#pragma unroll
    for (int sl = 0; sl < 4; ++sl)
      dln::bulk_g2s(buf + b * kBulkBlockBytes + sl * 4096, src + (size_t)sl * DLN_SLAB_BYTES, 4096, &bars[b]);
  };
  if (lane == 0) {
    if (n_items > 0) issue(0, 0);
    if (n_items > 1) issue(1, 1);
  }
  uint32_t phase = 0;                                              // bit b = parity to wait for on buffer b
  float h[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const uint32_t chunk = lane & 7;
  for (long long t = 0; t < n_items; ++t) {
    const int b = (int)(t & 1);
    dln::mbar_wait(&bars[b], (phase >> b) & 1u);
    phase ^= 1u << b;
    const uint8_t* rows = buf + b * kBulkBlockBytes + (lane >> 3) * 4096;
#pragma unroll 8
    for (int row = 0; row < 32; ++row)
      acc_bf16x8(h, *reinterpret_cast<const uint4*>(rows + row * 128 + (((chunk ^ (uint32_t)row) & 7u) << 4)));
    __syncwarp();                                                  // every lane has read buffer b
    if (lane == 0 && t + 2 < n_items) issue(t + 2, b);
    if ((t % n) != n - 1) continue;
    const long long g = gw + (t / n) * total_warps;                // the ray is complete
    if (hsum != nullptr) {
      float4* hs = reinterpret_cast<float4*>(hsum + (size_t)g * kW + 8 * lane);
      hs[0] = make_float4(h[0], h[1], h[2], h[3]);
      hs[1] = make_float4(h[4], h[5], h[6], h[7]);
    }
    if (out != nullptr) {
      const float cnt = (float)S;
      for (int k = 0; k < K; ++k) {
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(Sw + (size_t)k * kW + 8 * lane));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(Sw + (size_t)k * kW + 8 * lane + 4));
        float acc = h[0] * w0.x;
        acc = fmaf(h[1], w0.y, acc), acc = fmaf(h[2], w0.z, acc), acc = fmaf(h[3], w0.w, acc);
        acc = fmaf(h[4], w1.x, acc), acc = fmaf(h[5], w1.y, acc), acc = fmaf(h[6], w1.z, acc), acc = fmaf(h[7], w1.w, acc);
        acc = dln::warp_sum(acc);
        if (lane == 0) out[(size_t)g * out_ld + k] = acc + cnt * __ldg(sc + k);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] = 0.f;
  }
}

// Per-point logits (S = 1, no Hsum wanted: the semantic columns of a returned `raw`).  A warp per point would spend
// K shuffle reductions per point; here a 256-thread block takes a 128-point tile, thread = (point, half of the
// class PAIRS), the point's 512-byte row comes in 16-byte chunks (neighbouring threads read neighbouring 128-byte
// rows of the slab image) and Sw sits in shared memory interleaved by class pair, {Sw[2q][f], Sw[2q+1][f]}, so one
// packed fma.rn.f32x2 (two fp32 FMAs per issue slot) advances both classes of a pair: 19 x 256 MACs per point at
// ~1.2 issue slots per 2 MACs.
constexpr int kPtPairs = kMaxK / 4;      // class pairs per thread (half of the at most kMaxK / 2 pairs)

__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void fma2(unsigned long long& acc, unsigned long long a, unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}

__global__ void __launch_bounds__(256)
    sem_point_logits_kernel(const uint8_t* __restrict__ stash, int fwd_slots, int h_slot, long long P,
                            const float* __restrict__ Sw, const float* __restrict__ sc, int K,
                            float* __restrict__ out, int out_ld, long long n_tiles) {
  __shared__ __align__(16) float2 sw2[(kMaxK / 2) * kW];      // 32 KB
  __shared__ float sb[kMaxK];
  const int n_pairs = (K + 1) / 2;
  for (int i = threadIdx.x; i < n_pairs * kW; i += 256) {
    const int q = i / kW, f = i % kW;
    sw2[i] = make_float2(__ldg(Sw + (size_t)(2 * q) * kW + f), 2 * q + 1 < K ? __ldg(Sw + (size_t)(2 * q + 1) * kW + f) : 0.f);
  }
  if ((int)threadIdx.x < kMaxK) sb[threadIdx.x] = (int)threadIdx.x < K ? __ldg(sc + threadIdx.x) : 0.f;
  __syncthreads();
  const uint32_t r = threadIdx.x & 127u;
  const int half = threadIdx.x >> 7;                 // warp-uniform
  const int q0 = half ? (n_pairs + 1) / 2 : 0;
  const int qn = half ? n_pairs - q0 : (n_pairs + 1) / 2;        // <= kPtPairs
  const size_t tile_bytes = (size_t)fwd_slots * DLN_SLAB_BYTES;
  const uint32_t row_off = (r >> 3) * 1024u + (r & 7u) * 128u;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long p = tile * DLN_TILE_ROWS + r;
    const uint8_t* base = stash + (size_t)tile * tile_bytes + (size_t)h_slot * DLN_SLAB_BYTES + row_off;
    unsigned long long acc[kPtPairs];
#pragma unroll
    for (int q = 0; q < kPtPairs; ++q) acc[q] = 0ull;
#pragma unroll 2
    for (int c = 0; c < 32; ++c) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(base + (size_t)(c >> 3) * DLN_SLAB_BYTES +
                                                           ((((uint32_t)c ^ r) & 7u) << 4)));
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
      unsigned long long hh[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float lo = __uint_as_float(w[i] << 16), hi = __uint_as_float(w[i] & 0xffff0000u);
        hh[2 * i] = pack2(lo, lo), hh[2 * i + 1] = pack2(hi, hi);
      }
#pragma unroll
      for (int q = 0; q < kPtPairs; ++q)
        if (q < qn) {
          const ulonglong2* wp = reinterpret_cast<const ulonglong2*>(sw2 + (q0 + q) * kW + 8 * c);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const ulonglong2 ww = wp[i];            // features 8c + 2i, 8c + 2i + 1 of the class pair
            fma2(acc[q], hh[2 * i], ww.x);
            fma2(acc[q], hh[2 * i + 1], ww.y);
          }
        }
    }
    if (p < P) {
#pragma unroll
      for (int q = 0; q < kPtPairs; ++q)
        if (q < qn) {
          const int k = 2 * (q0 + q);
          float lo, hi;
          asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[q]));
          out[(size_t)p * out_ld + k] = lo + sb[k];
          if (k + 1 < K) out[(size_t)p * out_ld + k + 1] = hi + sb[k + 1];
        }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward: thread f of a block owns feature f; the block walks a contiguous range of groups.
//   G[g][f]  = sum_k dsem[g][k] Sw[k][f]          (what the dgrad chain adds to dH of the last trunk layer)
//   dSw[k][f] += sum_g dsem[g][k] Hsum[g][f],   dsc[k] += sum_g cnt_g dsem[g][k]
// ------------------------------------------------------------------------------------------------
constexpr int kBwdGroupsPerStage = 32;

__global__ void __launch_bounds__(256)
    sem_head_bwd_kernel(const float* __restrict__ dsem, int dsem_ld, const float* __restrict__ hsum,
                        const float* __restrict__ Sw, int K, long long n_groups, long long P, int S,
                        float* __restrict__ G, float* __restrict__ dSw, float* __restrict__ dsc) {
  __shared__ float ds[kBwdGroupsPerStage][kMaxK];
  const int f = threadIdx.x;
  float sw[kMaxK], acc[kMaxK];
#pragma unroll
  for (int k = 0; k < kMaxK; ++k) {
    sw[k] = k < K ? __ldg(Sw + (size_t)k * kW + f) : 0.f;
    acc[k] = 0.f;
  }
  float bsum = 0.f;                     // thread k < K: sum_g cnt_g dsem[g][k]
  const long long per = (n_groups + gridDim.x - 1) / gridDim.x;
  const long long g0 = (long long)blockIdx.x * per;
  const long long g1 = g0 + per < n_groups ? g0 + per : n_groups;
  for (long long gb = g0; gb < g1; gb += kBwdGroupsPerStage) {
    const int n = (int)(g1 - gb < kBwdGroupsPerStage ? g1 - gb : kBwdGroupsPerStage);
    __syncthreads();
    for (int idx = f; idx < n * kMaxK; idx += 256) {
      const int gi = idx / kMaxK, k = idx % kMaxK;
      ds[gi][k] = k < K ? __ldg(dsem + (size_t)(gb + gi) * dsem_ld + k) : 0.f;
    }
    __syncthreads();
    for (int gi = 0; gi < n; ++gi) {
      const long long g = gb + gi;
      const float h = __ldg(hsum + (size_t)g * kW + f);
      float gv = 0.f;
#pragma unroll
      for (int k = 0; k < kMaxK; ++k) {
        const float d = ds[gi][k];
        gv = fmaf(d, sw[k], gv);
        acc[k] = fmaf(d, h, acc[k]);
      }
      G[(size_t)g * kW + f] = gv;
      if (f < K) {
        long long cnt = P - g * S;
        cnt = cnt < S ? cnt : S;
        bsum = fmaf((float)cnt, ds[gi][f], bsum);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < kMaxK; ++k)
    if (k < K && acc[k] != 0.f) atomicAdd(dSw + (size_t)k * kW + f, acc[k]);
  if (f < K && bsum != 0.f) atomicAdd(dsc + f, bsum);
}

// ------------------------------------------------------------------------------------------------
// cross-entropy of the per-ray logits against class indices (F.cross_entropy, mean reduction, run_nerf.py:1542):
// loss_sum += sum_{n < n_rgb} (logsumexp(x_n) - x_n[t_n]);  dsem[n] = coef (softmax(x_n) - onehot(t_n)), 0 for the
// rays behind n_rgb (the depth rays carry no semantic target, :1458-1460).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    sem_ce_kernel(const float* __restrict__ logits, int ld, const long long* __restrict__ target, int n_rgb, int N,
                  int K, float coef, float* __restrict__ dsem, float* __restrict__ loss_sum) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  float loss = 0.f;
  if (n < N) {
    float* d = dsem + (size_t)n * K;
    if (n < n_rgb) {
      const float* x = logits + (size_t)n * ld;
      const int t = (int)target[n];
      float m = -INFINITY;
      for (int k = 0; k < K; ++k) m = fmaxf(m, x[k]);
      float s = 0.f;
      for (int k = 0; k < K; ++k) s += expf(x[k] - m);
      const float lse = m + logf(s);
      const bool ok = t >= 0 && t < K;       // anything else contributes nothing (cross_entropy would have raised)
      loss = ok ? lse - x[t] : 0.f;
      const float inv = 1.f / s;
      for (int k = 0; k < K; ++k) d[k] = ok ? coef * (expf(x[k] - m) * inv - (k == t ? 1.f : 0.f)) : 0.f;
    } else {
      for (int k = 0; k < K; ++k) d[k] = 0.f;
    }
  }
  loss = dln::warp_sum(loss);
  if ((threadIdx.x & 31) == 0 && loss != 0.f) atomicAdd(loss_sum, loss);
}

// ------------------------------------------------------------------------------------------------
// generic route: torch.sum(raw[..., c0:], -2) of raw2outputs (:589) and its broadcast backward
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    sample_sum_kernel(const float* __restrict__ raw, int C, int c0, int N, int S, float* __restrict__ out) {
  const int Kc = C - c0;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)N * Kc) return;
  const int n = (int)(idx / Kc), k = (int)(idx % Kc);
  const float* p = raw + (size_t)n * S * C + c0 + k;
  float a = 0.f;
  for (int s = 0; s < S; ++s) a += p[(size_t)s * C];
  out[idx] = a;
}

__global__ void __launch_bounds__(256)
    sample_sum_bwd_kernel(const float* __restrict__ g, int C, int c0, int N, int S, float* __restrict__ d_raw) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)N * S * C) return;
  const int c = (int)(idx % C);
  const long long n = idx / ((long long)S * C);
  d_raw[idx] = c >= c0 ? g[n * (C - c0) + c - c0] : 0.f;
}

long long n_groups_exact(long long P, int S) { return P / S; }

bool offsets_ok(const DlnSemOffsets* o) {
  return o && o->K >= 1 && o->K <= kMaxK && o->w_f >= 0 && o->b_f >= 0 && o->w_s1 >= 0 && o->b_s1 >= 0 &&
         o->w_s2 >= 0 && o->b_s2 >= 0 && o->A >= 0 && o->a >= 0 && o->Sw >= 0 && o->sc >= 0 && (o->Sw & 3) == 0;
}

}  // namespace

extern "C" {

int dln_sem_fold(float* params_flat, const DlnSemOffsets* off, void* stream) {
  DLN_CHECK_ARG(params_flat && offsets_ok(off));
  sem_fold_a_kernel<<<kHid, 256, 0, (cudaStream_t)stream>>>(params_flat, *off);
  sem_fold_s_kernel<<<off->K, 256, 0, (cudaStream_t)stream>>>(params_flat, *off);
  return dln_launch_status();
}

int dln_sem_unfold_grads(const float* params_flat, float* grads_flat, const DlnSemOffsets* off, void* stream) {
  DLN_CHECK_ARG(params_flat && grads_flat && offsets_ok(off));
  sem_unfold1_kernel<<<off->K + kHid, 256, 0, (cudaStream_t)stream>>>(params_flat, grads_flat, *off);
  sem_unfold2_kernel<<<kHid + kW + 1, 256, 0, (cudaStream_t)stream>>>(params_flat, grads_flat, *off);
  return dln_launch_status();
}

int dln_sem_head_fwd(const void* stash_fwd, int fwd_slots, int h_slot, long long P, int S, const float* params_flat,
                     const DlnSemOffsets* off, float* hsum, float* out, int out_ld, void* stream) {
  DLN_CHECK_ARG(stash_fwd && params_flat && offsets_ok(off) && P >= 0 && S >= 1 && fwd_slots >= 4 && h_slot >= 0 &&
                h_slot + 4 <= fwd_slots);
  DLN_CHECK_ARG((hsum || out) && (!out || out_ld >= off->K));
  DLN_CHECK_ARG((reinterpret_cast<uintptr_t>(stash_fwd) & 15) == 0 && (reinterpret_cast<uintptr_t>(hsum) & 15) == 0);
  if (P == 0) return DLN_OK;
  if (S == 1 && hsum == nullptr) {       // per-point logits only: tile-per-block kernel
    const long long n_tiles = (P + DLN_TILE_ROWS - 1) / DLN_TILE_ROWS;
    const long long nb = n_tiles > 148 * 4 ? 148 * 4 : n_tiles;
    sem_point_logits_kernel<<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint8_t*>(stash_fwd), fwd_slots, h_slot, P, params_flat + off->Sw, params_flat + off->sc,
        off->K, out, out_ld, n_tiles);
    return dln_launch_status();
  }
  const long long n_groups = (P + S - 1) / S;
  static const bool no_bulk = getenv("DLN_SEM_NO_BULK") != nullptr;      // A/B switch for the per-lane-load kernels
  if (S % 32 == 0 && P == n_groups_exact(P, S) * S && !no_bulk) {         // whole rays of 32-row blocks
    bool& attr_set = dln_device_flag(2);
    if (!attr_set) {
      cudaError_t e = cudaFuncSetAttribute(sem_head_fwd_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)kBulkSmemBytes);
      if (e != cudaSuccess) return (int)e;
      attr_set = true;
    }
    const long long ng = P / S;
    long long nb = (ng + kBulkWarps - 1) / kBulkWarps;
    nb = nb > 148 ? 148 : nb;
    sem_head_fwd_bulk_kernel<<<(unsigned)nb, kBulkWarps * 32, kBulkSmemBytes, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint8_t*>(stash_fwd), fwd_slots, h_slot, S, params_flat + off->Sw, params_flat + off->sc,
        off->K, hsum, out, out_ld, ng);
    return dln_launch_status();
  }
  const bool split = S >= 32;            // rays: four warps per ray; points / tiny groups: a warp per group
  const int gpb = split ? 2 : 8;
  long long blocks = (n_groups + gpb - 1) / gpb;
  blocks = blocks > 148 * 16 ? 148 * 16 : blocks;
  if (split)
    sem_head_fwd_kernel<4><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint8_t*>(stash_fwd), fwd_slots, h_slot, P, S, params_flat + off->Sw,
        params_flat + off->sc, off->K, hsum, out, out_ld, n_groups);
  else
    sem_head_fwd_kernel<1><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint8_t*>(stash_fwd), fwd_slots, h_slot, P, S, params_flat + off->Sw,
        params_flat + off->sc, off->K, hsum, out, out_ld, n_groups);
  return dln_launch_status();
}

int dln_sem_head_bwd(const float* dsem, int dsem_ld, const float* hsum, long long P, int S, const float* params_flat,
                     float* grads_flat, const DlnSemOffsets* off, float* G, void* stream) {
  DLN_CHECK_ARG(dsem && hsum && params_flat && grads_flat && G && offsets_ok(off) && P >= 0 && S >= 1 &&
                dsem_ld >= off->K);
  if (P == 0) return DLN_OK;
  const long long n_groups = (P + S - 1) / S;
  long long blocks = (n_groups + kBwdGroupsPerStage - 1) / kBwdGroupsPerStage;
  blocks = blocks > 148 * 2 ? 148 * 2 : blocks;
  sem_head_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      dsem, dsem_ld, hsum, params_flat + off->Sw, off->K, n_groups, P, S, G, grads_flat + off->Sw, grads_flat + off->sc);
  return dln_launch_status();
}

int dln_sample_sum(const float* raw, int C, int c0, int N, int S, float* out, void* stream) {
  DLN_CHECK_ARG(raw && out && C > c0 && c0 >= 0 && N >= 0 && S >= 1);
  if (N == 0) return DLN_OK;
  const long long n = (long long)N * (C - c0);
  sample_sum_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(raw, C, c0, N, S, out);
  return dln_launch_status();
}

int dln_sample_sum_bwd(const float* g, int C, int c0, int N, int S, float* d_raw, void* stream) {
  DLN_CHECK_ARG(g && d_raw && C > c0 && c0 >= 0 && N >= 0 && S >= 1);
  if (N == 0) return DLN_OK;
  const long long n = (long long)N * S * C;
  sample_sum_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(g, C, c0, N, S, d_raw);
  return dln_launch_status();
}

int dln_sem_ce_loss(const float* logits, int ld, const long long* target, int n_rgb, int N, int K, float coef,
                    float* dsem, float* loss_sum, void* stream) {
  DLN_CHECK_ARG(logits && dsem && loss_sum && N >= 0 && n_rgb >= 0 && n_rgb <= N && K >= 1 && K <= kMaxK && ld >= K);
  DLN_CHECK_ARG(n_rgb == 0 || target);
  if (N == 0) return DLN_OK;
  sem_ce_kernel<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(logits, ld, target, n_rgb, N, K, coef, dsem, loss_sum);
  return dln_launch_status();
}

}  // extern "C"
