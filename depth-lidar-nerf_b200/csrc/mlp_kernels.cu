// NeRF MLP (run_nerf_helpers.py:77-145) on the Blackwell tensor cores.
//
// Three kernels, all tcgen05 (UMMA, fp32 accumulators in TMEM) with bf16 operands staged in shared
// memory in the SWIZZLE_128B canonical layout and moved by the bulk-copy (TMA) engine:
//
//  chain_kernel   the fused per-tile layer chain.  One persistent CTA per SM walks 128-point tiles; the
//                 activations of the tile stay in shared memory from the first layer to the last, the
//                 weights of every layer are streamed from L2 through a 3-stage ring of 32 KB stages,
//                 and the epilogue of layer l (TMEM -> bias/ReLU -> bf16 -> smem) overlaps the MMAs of
//                 layer l+1 slab by slab (two TMEM accumulators).  The same machine runs the forward
//                 pass (prologue = stratified point + positional encoding computed in-kernel, heads =
//                 alpha / rgb on CUDA cores in the epilogue) and the dgrad pass (prologue = d raw ->
//                 d hidden through rgb_linear, epilogue = ReLU mask from 1-bit masks).
//  wgrad_kernel   dW += dZ^T * X over all points, both operands read back from the slab stashes the
//                 chain kernels wrote, as MN-major UMMA operands; split over the points across CTAs.
//  pack_kernel    fp32 master weights -> bf16 swizzled weight stages.
//
// Warp roles in chain_kernel (384 threads): warp 0 weight producer, warp 1 MMA issuer (+TMEM owner),
// warp 2 stash writer, warps 4..11 epilogue (two warpgroups, each owns half of the output columns).
#include "common.cuh"
#include "../../include/dlnerf_b200.h"
#include <math.h>

using namespace dln;

namespace {

constexpr int kThreads = 384;
constexpr int kEpiWarp0 = 4;
constexpr int kEpiThreads = 256;
constexpr int kNumStages = 3;
constexpr int kStageBytes = 32768;
constexpr int kSlab = DLN_SLAB_BYTES;
constexpr int kNumSlabs = 6;  // 0..3 activations, 4 encoded position / d_raw, 5 encoded direction
constexpr int kMaxHeadFloats = 1280;

struct ChainSmall {
  uint64_t w_full[kNumStages], w_empty[kNumStages];
  uint64_t a_ready[kNumSlabs], s_free[kNumSlabs];
  uint64_t acc_full[2];
  uint32_t tmem_base;
  uint32_t pad_;
  float bias[2][256];
  float heads[kMaxHeadFloats];
  float part[2][5][128];
};

constexpr size_t kChainSmemBytes = (size_t)kNumSlabs * kSlab + (size_t)kNumStages * kStageBytes + sizeof(ChainSmall) + 1024;

__device__ __forceinline__ void named_bar_epi() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// slabs produced by the tile prologue
__device__ __forceinline__ uint32_t prologue_mask(const DlnChainProgram& p) {
  if (!p.backward) return 0x30u;                 // encoded position + direction
  return p.use_viewdirs ? 0x13u : 0x1Fu;         // d_raw slab + dZ of the first backward layer
}
__device__ __forceinline__ uint32_t step_out_mask(const DlnChainStep& s) { return s.n_out == 256 ? 0xFu : 0x3u; }

// ---------------------------------------------------------------------------------------------
// row helpers
// ---------------------------------------------------------------------------------------------
// gamma(v) for one 3-vector into e[0..63]; entries past 3+6L are zero.  sincosf once per coordinate,
// higher octaves by the double-angle recurrence (abs. error <= 2^L * 1e-7, far below bf16 resolution).
__device__ __forceinline__ void encode_row(float x, float y, float z, int L, float (&e)[64]) {
#pragma unroll
  for (int i = 0; i < 64; ++i) e[i] = 0.f;
  e[0] = x, e[1] = y, e[2] = z;
  float s[3], c[3];
  sincosf(x, &s[0], &c[0]);
  sincosf(y, &s[1], &c[1]);
  sincosf(z, &s[2], &c[2]);
#pragma unroll
  for (int f = 0; f < 10; ++f) {
    if (f < L) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        e[3 + 6 * f + k] = s[k];
        e[3 + 6 * f + 3 + k] = c[k];
        const float s2 = 2.f * s[k] * c[k];
        const float c2 = 1.f - 2.f * s[k] * s[k];
        s[k] = s2, c[k] = c2;
      }
    }
  }
}

// write one 64-wide row (bf16) of a slab
__device__ __forceinline__ void store_row64(uint8_t* slab, int r, const float (&e)[64]) {
  uint8_t* row = slab + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
  for (int ch = 0; ch < 8; ++ch) {
    uint4 v;
    v.x = pack_bf16(e[8 * ch + 0], e[8 * ch + 1]);
    v.y = pack_bf16(e[8 * ch + 2], e[8 * ch + 3]);
    v.z = pack_bf16(e[8 * ch + 4], e[8 * ch + 5]);
    v.w = pack_bf16(e[8 * ch + 6], e[8 * ch + 7]);
    *reinterpret_cast<uint4*>(row + ((ch ^ (r & 7)) << 4)) = v;
  }
}

// write 32 consecutive columns [cb, cb+32) of row r into the activation slabs
__device__ __forceinline__ void store_cols32(uint8_t* act, int r, int cb, const float (&f)[32]) {
  uint8_t* row = act + (cb >> 6) * kSlab + (r >> 3) * 1024 + (r & 7) * 128;
  const int ch0 = (cb & 63) >> 3;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 v;
    v.x = pack_bf16(f[8 * q + 0], f[8 * q + 1]);
    v.y = pack_bf16(f[8 * q + 2], f[8 * q + 3]);
    v.z = pack_bf16(f[8 * q + 4], f[8 * q + 5]);
    v.w = pack_bf16(f[8 * q + 6], f[8 * q + 7]);
    *reinterpret_cast<uint4*>(row + (((ch0 + q) ^ (r & 7)) << 4)) = v;
  }
}

struct ProdTrack {
  uint32_t par = 0, any = 0;  // per-slab parity of the production count / "produced at least once"
};

// ---------------------------------------------------------------------------------------------
// the chain kernel
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1)
    chain_kernel(const __grid_constant__ DlnChainProgram prog, const __grid_constant__ DlnChainArgs args,
                 const long long n_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* slabs = smem;                              // 6 x 16 KB
  uint8_t* wring = smem + kNumSlabs * kSlab;          // 3 x 32 KB
  ChainSmall* sm = reinterpret_cast<ChainSmall*>(wring + kNumStages * kStageBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool keep = args.stash != nullptr && prog.stash_slots > 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kNumStages; ++i) mbar_init(&sm->w_full[i], 1), mbar_init(&sm->w_empty[i], 1);
    for (int i = 0; i < kNumSlabs; ++i) mbar_init(&sm->a_ready[i], 128), mbar_init(&sm->s_free[i], 1);
    mbar_init(&sm->acc_full[0], 1), mbar_init(&sm->acc_full[1], 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(&sm->tmem_base, 512);
    tmem_relinquish();
  }
  // head vectors -> smem: [prologue heads (bwd)] then the heads of every step, in step order
  if (warp >= kEpiWarp0) {
    const int t = threadIdx.x - kEpiWarp0 * 32;
    int base = 0;
    if (prog.backward) {
      const int n = (prog.use_viewdirs ? 3 * 128 : prog.out_ch * 256);
      for (int i = t; i < n; i += kEpiThreads) sm->heads[i] = args.fblob[prog.pro_head_off + i];
      base = n;
    }
    for (int s = 0; s < prog.n_steps; ++s) {
      const int n = prog.steps[s].n_heads * prog.steps[s].n_out;
      for (int i = t; i < n; i += kEpiThreads) sm->heads[base + i] = args.fblob[prog.steps[s].head_off + i];
      base += n;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm->tmem_base;

  if (warp == 0) {
    // ===================================================== weight producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      const uint8_t* wb = reinterpret_cast<const uint8_t*>(args.wblob);
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int s = 0; s < prog.n_steps; ++s) {
          const DlnChainStep& st = prog.steps[s];
          const uint32_t bytes = (uint32_t)st.n_out * 128u;
          for (int j = 0; j < st.nk; ++j) {
            mbar_wait(&sm->w_empty[stage], phase ^ 1);
            mbar_expect_tx(&sm->w_full[stage], bytes);
            bulk_g2s(wring + stage * kStageBytes, wb + st.w_off + (size_t)j * bytes, bytes, &sm->w_full[stage]);
            if (++stage == kNumStages) stage = 0, phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, gstep = 0;
      uint32_t par = 0;  // parity of the production count per slab
      const uint32_t pmask = prologue_mask(prog);
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        par ^= pmask;
        for (int s = 0; s < prog.n_steps; ++s, ++gstep) {
          const DlnChainStep& st = prog.steps[s];
          const uint32_t d_tmem = tmem_base + (gstep & 1) * 256;
          const uint32_t idesc = umma_idesc_bf16(128, st.n_out, 0, 0);
          for (int j = 0; j < st.nk; ++j) {
            const int slab = st.kslab[j];
            mbar_wait(&sm->a_ready[slab], ((par >> slab) & 1) ^ 1);   // last production completed
            mbar_wait(&sm->w_full[stage], phase);
            tc_fence_after();
            const uint32_t a_base = smem_u32(slabs + slab * kSlab);
            const uint32_t b_base = smem_u32(wring + stage * kStageBytes);
            for (int k = 0; k < st.kcnt[j]; ++k) {
              umma_bf16(d_tmem, umma_desc_sw128(a_base + k * 32, 16, 1024), umma_desc_sw128(b_base + k * 32, 16, 1024),
                        idesc, (j | k) != 0);
            }
            umma_commit(&sm->w_empty[stage]);
            if (++stage == kNumStages) stage = 0, phase ^= 1;
          }
          umma_commit(&sm->acc_full[gstep & 1]);
          par ^= step_out_mask(st);
        }
        // Parity waits are only sound while a waiter is never two phases ahead of the barrier.  No MMA
        // consumes the slabs the LAST step produces (they only go to the stash), so observe them here
        // before waiting for the next tile's productions of the same slabs.
        const uint32_t om = step_out_mask(prog.steps[prog.n_steps - 1]);
        for (int slab = 0; slab < 4; ++slab)
          if ((om >> slab) & 1) mbar_wait(&sm->a_ready[slab], ((par >> slab) & 1) ^ 1);
      }
    }
  } else if (warp == 2) {
    // ===================================================== stash writer (training only)
    if (lane == 0 && keep) {
      uint32_t par = 0;
      const uint32_t pmask = prologue_mask(prog);
      uint8_t* stash = reinterpret_cast<uint8_t*>(args.stash);
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        uint8_t* tbase = stash + (size_t)tile * prog.stash_slots * kSlab;
        auto handle = [&](int slab, int slot) {
          mbar_wait(&sm->a_ready[slab], (par >> slab) & 1);
          if (slot >= 0) {
            bulk_s2g(tbase + (size_t)slot * kSlab, slabs + slab * kSlab, kSlab);
            bulk_commit();
            bulk_wait_read0();
          }
          mbar_arrive(&sm->s_free[slab]);
          par ^= 1u << slab;
        };
        // prologue productions: aux slabs go to slots 0 (and 1), activation slabs to pro_slot + slab
        for (int slab = 0; slab < kNumSlabs; ++slab)
          if ((pmask >> slab) & 1) handle(slab, slab >= 4 ? slab - 4 : prog.pro_slot + slab);
        for (int s = 0; s < prog.n_steps; ++s) {
          const DlnChainStep& st = prog.steps[s];
          const uint32_t om = step_out_mask(st);
          const int order[4] = {0, 2, 1, 3};
          for (int i = 0; i < 4; ++i) {
            const int slab = order[i];
            if ((om >> slab) & 1) handle(slab, st.stash_slot >= 0 ? st.stash_slot + slab : -1);
          }
        }
      }
      bulk_wait_all0();
    }
  } else if (warp >= kEpiWarp0) {
    // ===================================================== prologue + epilogue warps
    const int et = threadIdx.x - kEpiWarp0 * 32;   // 0..255
    const int g = et >> 7;                         // warpgroup: column half
    const int r = ((warp & 3) << 5) | lane;        // tile row == TMEM lane
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    ProdTrack pt;
    const uint32_t pmask = prologue_mask(prog);
    uint32_t gstep = 0;
    auto begin_produce = [&](int slab) {
      if (keep && ((pt.any >> slab) & 1)) mbar_wait(&sm->s_free[slab], ((pt.par >> slab) & 1) ^ 1);
    };
    auto end_produce = [&](int slab) {
      fence_async_smem();
      tc_fence_before();
      mbar_arrive(&sm->a_ready[slab]);
    };
    // offsets of the head vectors in smem
    int head_base[DLN_MAX_STEPS];
    {
      int base = prog.backward ? (prog.use_viewdirs ? 3 * 128 : prog.out_ch * 256) : 0;
      for (int s = 0; s < DLN_MAX_STEPS; ++s) {
        head_base[s] = base;
        if (s < prog.n_steps) base += prog.steps[s].n_heads * prog.steps[s].n_out;
      }
    }

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long p = tile * DLN_TILE_ROWS + r;
      const bool valid = p < args.P;
      float dsig = 0.f;
      // ------------------------------------------------------------------ prologue
      if (!prog.backward) {
        float e[64];
        const int slab = 4 + g;
        if (args.x == nullptr) {
          float vx = 0.f, vy = 0.f, vz = 0.f;
          if (valid) {
            const long long ray = p / args.S;
            const float* rp = args.rays + (size_t)ray * args.ray_stride;
            if (g == 0) {
              const float zz = args.z[p];
              vx = rp[0] + rp[3] * zz, vy = rp[1] + rp[4] * zz, vz = rp[2] + rp[5] * zz;
            } else if (prog.use_viewdirs) {
              vx = rp[args.vd_col], vy = rp[args.vd_col + 1], vz = rp[args.vd_col + 2];
            }
          }
          encode_row(vx, vy, vz, g == 0 ? prog.L_pts : prog.L_dir, e);
          if (!valid || (g == 1 && !prog.use_viewdirs)) {
#pragma unroll
            for (int i = 0; i < 64; ++i) e[i] = 0.f;
          }
        } else {
          const int n_pts = 3 + 6 * prog.L_pts, n_dir = prog.use_viewdirs ? 3 + 6 * prog.L_dir : 0;
          const int n = g == 0 ? n_pts : n_dir;
          const float* xp = args.x + (size_t)(valid ? p : 0) * args.x_ld + (g == 0 ? 0 : n_pts);
#pragma unroll
          for (int i = 0; i < 64; ++i) e[i] = (valid && i < n) ? xp[i] : 0.f;
        }
        begin_produce(slab);
        store_row64(slabs + slab * kSlab, r, e);
        end_produce(slab);
      } else {
        // d raw row -> dZ of the first backward layer (through rgb_linear / output_linear) + d_raw slab
        float dr[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) dr[j] = (valid && j < prog.out_ch) ? args.d_out[(size_t)p * prog.out_ch + j] : 0.f;
        dsig = dr[3];
        const int nh = prog.use_viewdirs ? 3 : prog.out_ch;
        const int width = prog.use_viewdirs ? 128 : 256;
        const int ncols = width / 2, col0 = g * ncols;
        const uint4 mw = reinterpret_cast<const uint4*>(args.masks)[(((size_t)prog.pro_mask_slot * n_tiles + tile) * 2 + g) * 128 + r];
        const uint32_t mwords[4] = {mw.x, mw.y, mw.z, mw.w};
        for (int c = 0; c < ncols / 32; ++c) {
          const int cb = col0 + 32 * c;
          if ((cb & 63) == 0) begin_produce(cb >> 6);
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float v = 0.f;
#pragma unroll
            for (int j = 0; j < 5; ++j)
              if (j < nh) v += dr[j] * sm->heads[j * width + cb + i];
            f[i] = ((mwords[c] >> i) & 1u) ? v : 0.f;
          }
          store_cols32(slabs, r, cb, f);
          if ((cb & 63) == 32) end_produce(cb >> 6);
        }
        if (g == 0) {
          float e[64];
#pragma unroll
          for (int i = 0; i < 64; ++i) e[i] = 0.f;
#pragma unroll
          for (int j = 0; j < 5; ++j) e[j] = dr[j];
          begin_produce(4);
          store_row64(slabs + 4 * kSlab, r, e);
          end_produce(4);
        }
      }
      pt.par ^= pmask, pt.any |= pmask;

      // ------------------------------------------------------------------ layer epilogues
      for (int s = 0; s < prog.n_steps; ++s, ++gstep) {
        const DlnChainStep& st = prog.steps[s];
        const int epi = st.epi;
        const bool has_bias = epi <= DLN_EPI_RELU_OUT;
        float* bias = sm->bias[gstep & 1];
        if (et < st.n_out) bias[et] = has_bias ? args.fblob[st.bias_off + et] : 0.f;
        uint4 mw = make_uint4(0, 0, 0, 0);
        const size_t mask_idx = (((size_t)(st.mask_slot < 0 ? 0 : st.mask_slot) * n_tiles + tile) * 2 + g) * 128 + r;
        if (epi >= DLN_EPI_BWD_MASK && st.mask_slot >= 0) mw = reinterpret_cast<const uint4*>(args.masks)[mask_idx];
        named_bar_epi();
        mbar_wait(&sm->acc_full[gstep & 1], (gstep >> 1) & 1);
        tc_fence_after();

        const int ncols = st.n_out / 2, col0 = g * ncols;
        const uint32_t t_acc = tmem_base + (gstep & 1) * 256 + lane_addr;
        const bool relu = (epi == DLN_EPI_RELU || epi == DLN_EPI_RELU_SIGMA || epi == DLN_EPI_RELU_RGB || epi == DLN_EPI_RELU_OUT);
        const int nheads = (epi <= DLN_EPI_RELU_OUT) ? st.n_heads : 0;
        const float* hw = sm->heads + head_base[s];
        float hacc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        uint32_t mwords_in[4] = {mw.x, mw.y, mw.z, mw.w};
        uint32_t mwords_out[4] = {0, 0, 0, 0};
        for (int c = 0; c < ncols / 32; ++c) {
          const int cb = col0 + 32 * c;
          uint32_t v[32];
          tmem_ld32(t_acc + cb, v);
          tmem_ld_wait();
          if ((cb & 63) == 0) begin_produce(cb >> 6);
          float f[32];
          uint32_t mo = 0;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float x = __uint_as_float(v[i]) + bias[cb + i];
            if (epi == DLN_EPI_BWD_MASK_SIGMA) x += dsig * hw[cb + i];
            if (relu) {
              mo |= (x > 0.f ? 1u : 0u) << i;
              x = fmaxf(x, 0.f);
            }
            if (epi >= DLN_EPI_BWD_MASK) x = ((mwords_in[c] >> i) & 1u) ? x : 0.f;
            f[i] = x;
          }
          mwords_out[c] = mo;
          if (nheads > 0) {
#pragma unroll
            for (int h = 0; h < 5; ++h)
              if (h < nheads) {
                float a = 0.f;
#pragma unroll
                for (int i = 0; i < 32; ++i) a += f[i] * hw[h * st.n_out + cb + i];
                hacc[h] += a;
              }
          }
          store_cols32(slabs, r, cb, f);
          if ((cb & 63) == 32) end_produce(cb >> 6);
        }
        if (relu && st.mask_slot >= 0 && args.masks != nullptr)
          reinterpret_cast<uint4*>(args.masks)[mask_idx] = make_uint4(mwords_out[0], mwords_out[1], mwords_out[2], mwords_out[3]);
        pt.par ^= step_out_mask(st), pt.any |= step_out_mask(st);

        if (epi == DLN_EPI_RELU_SIGMA) {
          sm->part[g][4][r] = hacc[0] + (g == 0 ? args.fblob[st.head_bias_off] : 0.f);
        } else if (epi == DLN_EPI_RELU_RGB || epi == DLN_EPI_RELU_OUT) {
#pragma unroll
          for (int h = 0; h < 5; ++h)
            if (h < nheads) sm->part[g][h][r] = hacc[h];
          named_bar_epi();
          if (g == 0 && valid) {
            float o[5];
#pragma unroll
            for (int h = 0; h < 5; ++h)
              o[h] = (h < nheads) ? sm->part[0][h][r] + sm->part[1][h][r] + args.fblob[st.head_bias_off + h] : 0.f;
            if (epi == DLN_EPI_RELU_RGB) o[3] = sm->part[0][4][r] + sm->part[1][4][r];
            float* op = args.out + (size_t)p * prog.out_ch;
            if (prog.out_ch == 4) {
              *reinterpret_cast<float4*>(op) = make_float4(o[0], o[1], o[2], o[3]);
            } else {
              for (int j = 0; j < prog.out_ch; ++j) op[j] = o[j];
            }
          }
        }
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------
// wgrad kernel
// ---------------------------------------------------------------------------------------------
constexpr int kWgThreads = 256;       // warp 0 producer, warp 1 MMA, warps 4..7 bias sums + epilogue
constexpr int kWgStages = 3;
constexpr int kHalfSlab = kSlab / 2;  // 64 points x 64 features
constexpr int kWgStageBytes = 8 * kHalfSlab;  // 4 A + 4 B half-slabs = 64 KB

struct WgradSmall {
  uint64_t full[kWgStages], empty[kWgStages];
  uint64_t acc_full;
  uint32_t tmem_base;
};
constexpr size_t kWgradSmemBytes = (size_t)kWgStages * kWgStageBytes + sizeof(WgradSmall) + 1024;

__global__ void __launch_bounds__(kWgThreads, 1)
    wgrad_kernel(const DlnWgradItem* __restrict__ items, int splits, const uint8_t* __restrict__ stash_fwd,
                 int fwd_slots, const uint8_t* __restrict__ stash_bwd, int bwd_slots, long long n_tiles,
                 float* __restrict__ grads) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  WgradSmall* sm = reinterpret_cast<WgradSmall*>(smem + kWgStages * kWgStageBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const DlnWgradItem it = items[blockIdx.x / splits];
  const int split = blockIdx.x % splits;
  const long long per = (n_tiles + splits - 1) / splits;
  const long long t0 = split * per, t1 = (t0 + per < n_tiles) ? t0 + per : n_tiles;
  if (t0 >= t1) return;
  const long long n_stages = (t1 - t0) * 2;
  const int nh = it.a_nslab == 4 ? 2 : 1;
  const int N = it.b_nslab * 64;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kWgStages; ++i) mbar_init(&sm->full[i], 1), mbar_init(&sm->empty[i], 129);
    mbar_init(&sm->acc_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(&sm->tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      const uint8_t* sa = it.a_bwd_stash ? stash_bwd : stash_fwd;
      const int sa_slots = it.a_bwd_stash ? bwd_slots : fwd_slots;
      const uint8_t* sb = it.b_from_bwd ? stash_bwd : stash_fwd;
      const int sb_slots = it.b_from_bwd ? bwd_slots : fwd_slots;
      for (long long q = 0; q < n_stages; ++q) {
        const long long tile = t0 + (q >> 1);
        const int half = (int)(q & 1);
        uint8_t* buf = smem + stage * kWgStageBytes;
        mbar_wait(&sm->empty[stage], phase ^ 1);
        mbar_expect_tx(&sm->full[stage], (uint32_t)(it.a_nslab + it.b_nslab) * kHalfSlab);
        for (int i = 0; i < it.a_nslab; ++i)
          bulk_g2s(buf + i * kHalfSlab, sa + ((size_t)tile * sa_slots + it.a_slot + i) * kSlab + half * kHalfSlab,
                   kHalfSlab, &sm->full[stage]);
        for (int i = 0; i < it.b_nslab; ++i)
          bulk_g2s(buf + (4 + i) * kHalfSlab, sb + ((size_t)tile * sb_slots + it.b_slot + i) * kSlab + half * kHalfSlab,
                   kHalfSlab, &sm->full[stage]);
        if (++stage == kWgStages) stage = 0, phase ^= 1;
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      const uint32_t idesc = umma_idesc_bf16(128, N, 1, 1);
      for (long long q = 0; q < n_stages; ++q) {
        mbar_wait(&sm->full[stage], phase);
        tc_fence_after();
        const uint32_t abuf = smem_u32(smem + stage * kWgStageBytes);
        const uint32_t bbuf = abuf + 4 * kHalfSlab;
        for (int kk = 0; kk < 4; ++kk) {          // 16 points per MMA
          const uint64_t bdesc = umma_desc_sw128(bbuf + kk * 2048, kHalfSlab, 1024);
          for (int h = 0; h < nh; ++h) {
            const uint64_t adesc = umma_desc_sw128(abuf + h * 2 * kHalfSlab + kk * 2048, kHalfSlab, 1024);
            umma_bf16(tmem_base + h * 256, adesc, bdesc, idesc, (q | kk) != 0);
          }
        }
        umma_commit(&sm->empty[stage]);
        if (++stage == kWgStages) stage = 0, phase ^= 1;
      }
      umma_commit(&sm->acc_full);
    }
  } else if (warp >= 4) {
    const int t = threadIdx.x - 128;     // 0..127
    const int r = t;                     // accumulator row within a 128-row half == TMEM lane
    // ---- bias gradient: column sums of the A slabs, two adjacent features per thread
    uint32_t stage = 0, phase = 0;
    float b0 = 0.f, b1 = 0.f;
    const int feat = 2 * t;
    const bool do_bias = it.db_off >= 0 && feat < it.a_nslab * 64;
    for (long long q = 0; q < n_stages; ++q) {
      mbar_wait(&sm->full[stage], phase);
      if (do_bias) {
        const uint8_t* a = smem + stage * kWgStageBytes + (feat >> 6) * kHalfSlab;
#pragma unroll 8
        for (int row = 0; row < 64; ++row) {
          const uint32_t w = *reinterpret_cast<const uint32_t*>(a + slab_off(row, feat & 63));
          b0 += __uint_as_float(w << 16);
          b1 += __uint_as_float(w & 0xffff0000u);
        }
      }
      mbar_arrive(&sm->empty[stage]);
      if (++stage == kWgStages) stage = 0, phase ^= 1;
    }
    if (do_bias) {
      float* db = grads + it.db_off;
      if (feat >= it.db_col_off && feat < it.db_col_off + it.db_n) atomicAdd(db + feat - it.db_col_off, b0);
      if (feat + 1 >= it.db_col_off && feat + 1 < it.db_col_off + it.db_n) atomicAdd(db + feat + 1 - it.db_col_off, b1);
    }
    // ---- dW: TMEM -> global (fp32 atomics; several CTAs own the same item)
    mbar_wait(&sm->acc_full, 0);
    tc_fence_after();
    float* dw = grads + it.dw_off;
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    for (int h = 0; h < nh; ++h) {
      const int dst_row = h * 128 + r - it.row_off;
      const bool row_ok = dst_row >= 0 && dst_row < it.n_rows;
      for (int c = 0; c < N / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_base + h * 256 + c * 32 + lane_addr, v);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int col = c * 32 + i;
            if (col < it.n_cols) atomicAdd(dw + (size_t)dst_row * it.ld + it.col_off + col, __uint_as_float(v[i]));
          }
        }
      }
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------
// weight packing
// ---------------------------------------------------------------------------------------------
__global__ void pack_kernel(const float* __restrict__ params, const DlnPackJob* __restrict__ jobs, uint8_t* __restrict__ blob) {
  const DlnPackJob jb = jobs[blockIdx.x];
  const float* W = params + jb.src_off;
  uint8_t* dst = blob + jb.dst_off;
  for (int idx = threadIdx.x; idx < jb.n_rows * 8; idx += blockDim.x) {
    const int n = idx >> 3, ch = idx & 7;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = ch * 8 + i;
      float x = 0.f;
      if (n < jb.n_valid && k < jb.k_valid)
        x = jb.transposed ? W[(size_t)(jb.row0 + k) * jb.ld + jb.col0 + n] : W[(size_t)(jb.row0 + n) * jb.ld + jb.col0 + k];
      v[i] = x;
    }
    uint4 o;
    o.x = pack_bf16(v[0], v[1]), o.y = pack_bf16(v[2], v[3]), o.z = pack_bf16(v[4], v[5]), o.w = pack_bf16(v[6], v[7]);
    *reinterpret_cast<uint4*>(dst + (n >> 3) * 1024 + (n & 7) * 128 + ((ch ^ (n & 7)) << 4)) = o;
  }
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

int dln_mlp_chain(const DlnChainProgram* prog, const DlnChainArgs* args, int num_sms, void* stream) {
  DLN_CHECK_ARG(prog && args && num_sms > 0);
  DLN_CHECK_ARG(prog->n_steps >= 1 && prog->n_steps <= DLN_MAX_STEPS);
  DLN_CHECK_ARG(prog->out_ch >= 1 && prog->out_ch <= 5);
  DLN_CHECK_ARG(prog->L_pts >= 0 && prog->L_pts <= 10 && prog->L_dir >= 0 && prog->L_dir <= 10);
  DLN_CHECK_ARG(args->P >= 0);
  if (args->P == 0) return DLN_OK;
  DLN_CHECK_ARG(args->wblob && args->fblob);
  int head_floats = prog->backward ? (prog->use_viewdirs ? 3 * 128 : prog->out_ch * 256) : 0;
  for (int s = 0; s < prog->n_steps; ++s) {
    const DlnChainStep& st = prog->steps[s];
    DLN_CHECK_ARG(st.n_out == 256 || st.n_out == 128);
    DLN_CHECK_ARG(st.nk >= 1 && st.nk <= DLN_MAX_KSLABS && st.n_heads <= 5);
    for (int j = 0; j < st.nk; ++j) DLN_CHECK_ARG(st.kslab[j] < kNumSlabs && st.kcnt[j] >= 1 && st.kcnt[j] <= 4);
    DLN_CHECK_ARG((st.w_off & 1023u) == 0);
    head_floats += st.n_heads * st.n_out;
  }
  DLN_CHECK_ARG(head_floats <= kMaxHeadFloats);
  if (prog->backward) {
    DLN_CHECK_ARG(args->d_out && args->masks);
  } else {
    DLN_CHECK_ARG(args->out);
    DLN_CHECK_ARG(args->x || (args->rays && args->z && args->S >= 1 && args->ray_stride >= 6));
  }
  if (args->stash && prog->stash_slots > 0) DLN_CHECK_ARG((reinterpret_cast<uintptr_t>(args->stash) & 15) == 0);
  if (args->P == 0) return DLN_OK;
  const long long n_tiles = (args->P + DLN_TILE_ROWS - 1) / DLN_TILE_ROWS;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kChainSmemBytes);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const unsigned grid = (unsigned)(n_tiles < num_sms ? n_tiles : num_sms);
  chain_kernel<<<grid, kThreads, kChainSmemBytes, (cudaStream_t)stream>>>(*prog, *args, n_tiles);
  return dln_launch_status();
}

int dln_mlp_wgrad(const DlnWgradItem* items_dev, int n_items, int splits, const void* stash_fwd, int fwd_slots,
                  const void* stash_bwd, int bwd_slots, long long n_tiles, float* grads_flat, void* stream) {
  DLN_CHECK_ARG(items_dev && n_items >= 1 && splits >= 1 && stash_fwd && stash_bwd && grads_flat && n_tiles >= 0);
  if (n_tiles == 0) return DLN_OK;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWgradSmemBytes);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  wgrad_kernel<<<(unsigned)(n_items * splits), kWgThreads, kWgradSmemBytes, (cudaStream_t)stream>>>(
      items_dev, splits, reinterpret_cast<const uint8_t*>(stash_fwd), fwd_slots,
      reinterpret_cast<const uint8_t*>(stash_bwd), bwd_slots, n_tiles, grads_flat);
  return dln_launch_status();
}

int dln_mlp_pack_weights(const float* params_flat, const DlnPackJob* jobs_dev, int n_jobs, void* wblob, void* stream) {
  DLN_CHECK_ARG(params_flat && jobs_dev && wblob && n_jobs >= 1);
  pack_kernel<<<n_jobs, 256, 0, (cudaStream_t)stream>>>(params_flat, jobs_dev, reinterpret_cast<uint8_t*>(wblob));
  return dln_launch_status();
}

int dln_abi_sizes(int* out) {
  DLN_CHECK_ARG(out);
  out[0] = (int)sizeof(DlnChainStep), out[1] = (int)sizeof(DlnChainProgram), out[2] = (int)sizeof(DlnChainArgs);
  out[3] = (int)sizeof(DlnWgradItem), out[4] = (int)sizeof(DlnPackJob);
  return DLN_OK;
}

const char* dln_build_info(void) { return "dlnerf_b200 sm_100a tcgen05/TMEM + bulk-copy (TMA) weights"; }

}  // extern "C"
