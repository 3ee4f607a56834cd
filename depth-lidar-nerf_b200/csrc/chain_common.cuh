// Device helpers shared by the chain kernels (mlp_kernels.cu: one tile per CTA, activations through tensor memory;
// mlp_chain2.cu: CTA pairs, two tiles in flight per CTA, activations through shared memory): positional encoding of a
// row quarter, swizzled slab stores, and the per-chunk epilogue arithmetic (bias / ReLU / masks / heads).
#pragma once
#include "common.cuh"
#include "../../include/dlnerf_b200.h"
#include <math.h>

namespace dln {

constexpr int kSlab = DLN_SLAB_BYTES;

// ---------------------------------------------------------------------------------------------
// row helpers
// ---------------------------------------------------------------------------------------------
// Columns [16q, 16q+16) of gamma(v) = [v, sin(2^f v), cos(2^f v)]_f (entries past 3+6L are zero): the four
// epilogue warpgroups encode one quarter of a row each.  A quarter touches at most 4 frequencies; sincosf is
// evaluated once per coordinate at the lowest of them and the higher octaves follow by the double-angle recurrence
// (at most 3 steps, abs. error < 1e-6).
template <int q>
__device__ __forceinline__ void encode_quarter_t(float x, float y, float z, int L, float (&e)[16]) {
  constexpr int c0 = 16 * q;
  constexpr int f0 = c0 >= 3 ? (c0 - 3) / 6 : 0;        // lowest frequency index with a column in the quarter
  const float sc = exp2f((float)f0);
  float s[3], c[3];
  sincosf(x * sc, &s[0], &c[0]);
  sincosf(y * sc, &s[1], &c[1]);
  sincosf(z * sc, &s[2], &c[2]);
  const float v[3] = {x, y, z};
  const int ncol = 3 + 6 * L;
#pragma unroll
  for (int i = 0; i < 16; ++i) e[i] = 0.f;
#pragma unroll
  for (int df = 0; df < 4; ++df) {                       // frequencies f0 .. f0+3 cover any 16-column window
    const int base = 3 + 6 * (f0 + df) - c0;             // column (relative to the quarter) of sin(2^f x); compile-time
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (i == base + k && c0 + i < ncol) e[i] = s[k];
        if (i == base + 3 + k && c0 + i < ncol) e[i] = c[k];
      }
      const float s2 = 2.f * s[k] * c[k], c2 = 1.f - 2.f * s[k] * s[k];
      s[k] = s2, c[k] = c2;
    }
  }
  if (q == 0) e[0] = x, e[1] = y, e[2] = z;
  (void)v;
}
__device__ __forceinline__ void encode_quarter(float x, float y, float z, int L, int q, float (&e)[16]) {
  switch (q) {      // q is warpgroup-uniform: the column positions inside a quarter become compile-time constants
    case 0: encode_quarter_t<0>(x, y, z, L, e); break;
    case 1: encode_quarter_t<1>(x, y, z, L, e); break;
    case 2: encode_quarter_t<2>(x, y, z, L, e); break;
    default: encode_quarter_t<3>(x, y, z, L, e); break;
  }
}
// write columns [16q, 16q+16) of row r (two 16-byte swizzle chunks)
__device__ __forceinline__ void store_quarter(uint8_t* slab, int r, int q, const uint32_t (&pk)[8]) {
  uint8_t* row = slab + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
  for (int h = 0; h < 2; ++h)
    *reinterpret_cast<uint4*>(row + (((2 * q + h) ^ (r & 7)) << 4)) = make_uint4(pk[4 * h], pk[4 * h + 1], pk[4 * h + 2], pk[4 * h + 3]);
}

// write one 64-wide row (bf16) of a slab; `gslab` (optional) is the same slab image in the global stash
__device__ __forceinline__ void store_row64(uint8_t* slab, uint8_t* gslab, int r, const float (&e)[64]) {
  const int rowoff = (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
  for (int ch = 0; ch < 8; ++ch) {
    uint4 v;
    v.x = pack_bf16(e[8 * ch + 0], e[8 * ch + 1]);
    v.y = pack_bf16(e[8 * ch + 2], e[8 * ch + 3]);
    v.z = pack_bf16(e[8 * ch + 4], e[8 * ch + 5]);
    v.w = pack_bf16(e[8 * ch + 6], e[8 * ch + 7]);
    const int off = rowoff + ((ch ^ (r & 7)) << 4);
    *reinterpret_cast<uint4*>(slab + off) = v;
    if (gslab) *reinterpret_cast<uint4*>(gslab + off) = v;
  }
}

__device__ __forceinline__ uint32_t pack_bf16_relu(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));   // relu fused into the conversion
  return r;
}

// 32 fp32 values -> 16 packed bf16x2 words (optionally relu'd on the fly)
template <bool kRelu>
__device__ __forceinline__ void pack32(const float (&f)[32], uint32_t (&pk)[16]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) pk[i] = kRelu ? pack_bf16_relu(f[2 * i], f[2 * i + 1]) : pack_bf16(f[2 * i], f[2 * i + 1]);
}
// write the 32 consecutive columns [cb, cb+32) of row r (16 packed words) into the activation slabs
__device__ __forceinline__ void store_packed32(uint8_t* act, int r, int cb, const uint32_t (&pk)[16]) {
  uint8_t* row = act + (cb >> 6) * kSlab + (r >> 3) * 1024 + (r & 7) * 128;
  const int ch0 = (cb & 63) >> 3;
#pragma unroll
  for (int q = 0; q < 4; ++q)
    *reinterpret_cast<uint4*>(row + (((ch0 + q) ^ (r & 7)) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
}
// Global-memory variants for the direct stash: the two 16-byte swizzle chunks (2k, 2k+1) of a row are adjacent,
// whatever the XOR with (r & 7) does to their order, so 16 columns are ONE aligned 32-byte sector = one 256-bit store.
__device__ __forceinline__ void st_global_256(void* p, const uint32_t* lo4, const uint32_t* hi4) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(lo4[0]), "r"(lo4[1]),
               "r"(lo4[2]), "r"(lo4[3]), "r"(hi4[0]), "r"(hi4[1]), "r"(hi4[2]), "r"(hi4[3])
               : "memory");
}
// columns [cb, cb+16) of row r of a slab image in global memory (pk = 8 packed words)
__device__ __forceinline__ void store_global16(uint8_t* gslab_base, int r, int cb, const uint32_t* pk) {
  uint8_t* row = gslab_base + (cb >> 6) * kSlab + (r >> 3) * 1024 + (r & 7) * 128;
  const int c_even = ((cb & 63) >> 3) ^ (r & 7);          // swizzled position of the first chunk
  uint8_t* sector = row + ((c_even & ~1) << 4);
  if (c_even & 1) st_global_256(sector, pk + 4, pk);      // odd position: the first chunk is the upper half
  else st_global_256(sector, pk, pk + 4);
}
__device__ __forceinline__ void store_global32(uint8_t* gslab_base, int r, int cb, const uint32_t (&pk)[16]) {
  store_global16(gslab_base, r, cb, pk);
  store_global16(gslab_base, r, cb + 16, pk + 8);
}

// two fp32 adds in one instruction (FADD2)
__device__ __forceinline__ void add2(float& a0, float& a1, float b0, float b1) {
  unsigned long long a, b;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(a) : "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(a));
}

// Column ownership inside a tile: every epilogue thread owns one row and two 32-column chunks, chunk h (0/1) of
// warpgroup g being columns [128h + 32g, +32), i.e. slab 2h + (g>>1), half g&1.  Two warpgroups therefore
// finish slabs 0 and 1 (the first K slabs of the next layer) together before anybody starts on slabs 2 and 3,
// and a 128-wide output keeps all four warpgroups busy.  relu-mask word h of the thread's uint2 covers chunk h.
//
// One 32-column chunk of an epilogue: TMEM values -> (+bias | +dsigma*head) -> relu / mask -> 16 packed bf16x2 words.
template <int EPI, bool kBiasLdg = false>      // kBiasLdg: the bias vector lives in global memory (read-only path) instead of smem
__device__ __forceinline__ void epi_chunk(const uint32_t (&v)[32], const float* __restrict__ bias,
                                          const float* __restrict__ hw, int n_out, int nheads, float dsig,
                                          uint32_t mw_in, uint32_t& mw_out, float (&hacc)[5], uint32_t (&pk)[16],
                                          int cb, const float* __restrict__ semrow) {
  constexpr bool kFwd = EPI <= DLN_EPI_RELU_OUT;
  constexpr bool kRelu = EPI == DLN_EPI_RELU || EPI == DLN_EPI_RELU_SIGMA || EPI == DLN_EPI_RELU_RGB || EPI == DLN_EPI_RELU_OUT;
  float f[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
  if (kFwd) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 b = kBiasLdg ? __ldg(reinterpret_cast<const float4*>(bias + cb + 4 * q))
                                : *reinterpret_cast<const float4*>(bias + cb + 4 * q);   // smem broadcast
      add2(f[4 * q], f[4 * q + 1], b.x, b.y);
      add2(f[4 * q + 2], f[4 * q + 3], b.z, b.w);
    }
  }
  if (EPI == DLN_EPI_BWD_MASK_SIGMA) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 h = __ldg(reinterpret_cast<const float4*>(hw + cb + 4 * q));
      f[4 * q] += dsig * h.x, f[4 * q + 1] += dsig * h.y, f[4 * q + 2] += dsig * h.z, f[4 * q + 3] += dsig * h.w;
    }
    if (semrow != nullptr) {      // semantic head: dH += dsem Sw, one fp32 row per ray (dln_sem_head_bwd)
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 h = __ldg(reinterpret_cast<const float4*>(semrow + cb + 4 * q));
        add2(f[4 * q], f[4 * q + 1], h.x, h.y);
        add2(f[4 * q + 2], f[4 * q + 3], h.z, h.w);
      }
    }
  }
  if (kRelu) {
    // sign bits -> mask word with one funnel shift per element (bit i <-> column cb+i); relu' := (x >= +0)
    // (four independent 8-long chains instead of one 32-long dependent chain)
    uint32_t n0 = 0, n1 = 0, n2 = 0, n3 = 0;
#pragma unroll
    for (int i = 7; i >= 0; --i) {
      n0 = __funnelshift_l(__float_as_uint(f[i]), n0, 1);
      n1 = __funnelshift_l(__float_as_uint(f[8 + i]), n1, 1);
      n2 = __funnelshift_l(__float_as_uint(f[16 + i]), n2, 1);
      n3 = __funnelshift_l(__float_as_uint(f[24 + i]), n3, 1);
    }
    mw_out = ~(n0 | (n1 << 8) | (n2 << 16) | (n3 << 24));
  }
  if (EPI >= DLN_EPI_BWD_MASK) {
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = (mw_in & (1u << i)) ? f[i] : 0.f;
  }
  if (kFwd && nheads > 0) {
    // heads (alpha / rgb / output_linear) act on the fp32 relu'd activations
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
#pragma unroll
    for (int h = 0; h < 5; ++h)
      if (h < nheads) {
        float a = 0.f;
        const float* w = hw + h * n_out + cb;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 w4 = __ldg(reinterpret_cast<const float4*>(w + 4 * q));
          a += f[4 * q] * w4.x + f[4 * q + 1] * w4.y + f[4 * q + 2] * w4.z + f[4 * q + 3] * w4.w;
        }
        hacc[h] += a;
      }
  }
  pack32<kRelu>(f, pk);
}

}  // namespace dln
