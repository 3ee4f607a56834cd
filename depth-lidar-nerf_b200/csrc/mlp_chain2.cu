// chain2_kernel: the fused per-tile layer chain of the NeRF MLP (run_nerf_helpers.py:113-145 and its autograd dgrad)
// for CTA PAIRS with TWO 128-point tiles in flight per CTA.
//
// Why (round-1 profile of the one-tile kernel, mlp_kernels.cu): with one tile per SM the MMAs of layer l+1 wait for the
// epilogue of layer l, so the tensor pipe idles ~65 % of the time; a second tile in flight did not fit next to a
// whole-layer weight ring (128 KB) in 227 KB of shared memory.  Here
//   * `tcgen05.mma.cta_group::2` (M = 256: rows 0..127 from this CTA, 128..255 from its cluster peer) lets each CTA stage
//     only HALF of every weight matrix (n_out/2 rows: 64 KB per 256x256 layer), and the two tiles of a CTA share it, so
//     the weight traffic per tile and SM drops 4x and a whole layer fits a 4 x 16 KB ring;
//   * each CTA keeps two tiles (slots X, Y): activations as K-major SWIZZLE_128B slab images in shared memory (4 x 16 KB
//     per slot = the A operand of the next layer AND the staging buffer of the stash copy -- one st.shared serves both),
//     one 256-column fp32 accumulator per slot in tensor memory (2 x 256 = all 512 columns);
//   * the MMA warp of the leader CTA alternates X(l), Y(l), X(l+1), ... : the epilogue of X(l) runs under the MMAs of
//     Y(l).  Operands are handed over per (slot, step) in two halves (activation slabs 0, 1 first), one mbarrier phase
//     each; remote arrivals use the default CTA-scope semantics behind fence.proxy.async (the cluster-scope forms cost
//     a MEMBAR + ERRBAR per arrival and an L1 invalidation per wait, see mbar_arrive_cluster).
//
// Bound: without the activation stash the epilogue's instruction issue, ~0.72 of the burst bf16
// peak for the D=8 forward; with it (training) the HBM write-only rate -- the kernel is a pure write stream.
//
// Per-CTA shared memory: A[2][4] slabs 128 KB | AUX[2] 32 KB (encoded position -> encoded direction, or d_raw; scratch
// of the head sums at the end of a tile) | weight ring 4 x 16 KB | barriers.  Biases and head weights are read through
// the read-only path (no room to stage them).
//
// Warp roles (640 threads, both CTAs): warp 0 weight producer (its half of every K slab), warp 1 MMA issuer (leader CTA
// only) + TMEM owner, warp 2 stash lane, warp 3 weight-arrival helper (tells the leader that this CTA's half has
// landed), warps 4..19 epilogue: thread = row, chunk c of warpgroup g = columns [128 c + 32 g, + 32) of both slots, so the
// slabs fill in order.  The service warpgroup hands registers to the epilogue warps (setmaxnreg 24 / 112).
// Fused into the tile prologue: stratified sampling (coarse pass), o + d z, positional encoding.
//
// Same program / argument structs, stash images and ReLU-mask layout as the one-tile kernel, so wgrad_kernel, the host
// plans and every test are shared (dln_mlp_chain picks the kernel).
#include "chain_common.cuh"

using namespace dln;

namespace {

#ifndef DLN_CHAIN2_WG
#define DLN_CHAIN2_WG 4
#endif
// Epilogue warpgroups.  4 (default): 20 warps launched with 96 registers each; the four service warps (one warpgroup)
// then hand registers over with setmaxnreg (24 / 112: 128 x 24 + 512 x 112 = 60 416 of the 61 440 the CTA was launched
// with -- the pool is the CTA's launch allocation, 40 / 120 hangs at the inc), issued INSIDE the role branches so that
// the register budget of each role's code is the one that dominates it (issued ahead of the branches ptxas holds all
// code to the minimum).  2: 11 warps with 168 registers (measured: forward without stash 0.61 ms against 0.52 ms).
constexpr int k2WG = DLN_CHAIN2_WG;
constexpr bool k2Rebalance = k2WG == 4;
constexpr int k2EpiWarp0 = k2WG == 4 ? 4 : 3;
constexpr int kFastChunks = 256 / 32 / k2WG;     // 32-column chunks per thread in a 256-wide step
constexpr int k2Threads = (k2EpiWarp0 + 4 * k2WG) * 32;
constexpr int k2Stages = 4;
constexpr int k2StageBytes = 16384;     // one K slab of this CTA's half of a weight matrix: (n_out / 2) rows x 128 B
constexpr int k2TraceSteps = 8;         // traced (round, step) pairs; 16 events each
constexpr uint32_t k2TraceStep0 = 10;   // first traced (round, step) index: steady state of the second round

struct Chain2Small {
  uint64_t w_full[k2Stages];            // local: this CTA's half of the stage has landed (transaction bytes)
  uint64_t w_empty[k2Stages];           // local: the pair's MMAs have read the stage (multicast commit)
  uint64_t w_ready[k2Stages];           // leader's copy is the live one: both CTAs' helpers arrive (count 2)
  uint64_t a_ready[2][2];               // leader's copy: [slot][0: activation slabs 0, 1 | 1: slabs 2, 3 + aux] of the slot's next
                                        // step written in BOTH CTAs (one arrival per epilogue warp of the pair)
  uint64_t acc_full[2];                 // local: the slot's accumulator is complete (multicast commit)
  uint64_t s_ready[2][2];               // local: [slot][0 activation slabs | 1 aux slab] staged for the stash lane (16 warps)
  uint64_t s_free[2][2];                // local: the stash copies have read them (1)
  uint32_t tmem_base;
  uint32_t pad_;
  uint16_t tr[k2TraceSteps * 16];
};
static_assert(sizeof(Chain2Small) <= 2048, "barrier block too large");

constexpr size_t kChain2SmemBytes = (size_t)(8 + 2) * kSlab + (size_t)k2Stages * k2StageBytes + 2048 + 1024;
static_assert(kChain2SmemBytes <= 232448, "chain2 kernel shared memory exceeds the 227 KB per-CTA limit");

// ------------------------------------------------------------------ cluster / pair PTX
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
// Remote arrival with the default semantics (release at CTA scope).  The cluster-scope forms cost far more than they
// look: `mbarrier.arrive.release.cluster` compiles to MEMBAR.ALL + ERRBAR in front of the arrival (it waits for every
// outstanding load / store of the thread, 8 % of all stall samples of the first version) and
// `mbarrier.try_wait.acquire.cluster` to a CCTL.IVALL after every successful wait, which empties the L1 and with it the
// bias vectors and head weights the epilogue reads through the read-only path.  What the hand-offs need is ordering
// between generic-proxy shared-memory writes and the async proxy, which fence.proxy.async in front of the arrival gives.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Bounded wait without the printf of common.cuh's mbar_wait (a protocol bug still traps): ~20 call sites with their
// argument set-up would otherwise sit in the instruction stream of every role.
__device__ __forceinline__ void mbar_wait2(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity))
    if (++spins > (1u << 26)) __trap();
}
// wait on a LOCAL barrier whose arrivals may come from the peer CTA (or from the pair's multicast commits)
__device__ __forceinline__ void mbar_wait_cl(uint64_t* bar, uint32_t parity) { mbar_wait2(bar, parity); }
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_out, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_out)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem, 256 rows over the pair] (+)= A[smem, 128 rows per CTA] * B[smem, n_out/2 rows per CTA]^T
__device__ __forceinline__ void umma2_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at this shared-memory offset in BOTH CTAs once all MMAs issued so far have completed
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)), "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void named_bar_epi2() { asm volatile("bar.sync 1, %0;" ::"n"(128 * k2WG) : "memory"); }

// index of the 32-bit ReLU-mask word that covers columns [col0, col0+32) (col0 % 32 == 0) of row r: the one-tile
// kernel's layout [mask_slot][tile][warpgroup = (col0 % 128) / 32][row] uint2, component col0 / 128
__device__ __forceinline__ size_t mask_word(int mask_slot, long long n_tiles, long long tile, int col0, int r) {
  return (((((size_t)mask_slot * n_tiles + tile) * 4 + ((col0 & 127) >> 5)) * 128 + r) << 1) + (col0 >> 7);
}

__device__ __forceinline__ void trace2(Chain2Small* sm, const long long* trace, uint32_t gstep, int ev) {
#ifndef DLN_CHAIN2_TRACE      // the timeline costs registers and branches in the hot loop: built in only on request
  (void)sm, (void)trace, (void)gstep, (void)ev;
  return;
#endif
  if (trace != nullptr && blockIdx.x == 0 && gstep - k2TraceStep0 < (uint32_t)k2TraceSteps) {
    uint32_t c;
    asm volatile("mov.u32 %0, %%clock;" : "=r"(c));
    sm->tr[(gstep - k2TraceStep0) * 16 + ev] = (uint16_t)(c >> 3) | 1u;
  }
}

// One 32-column chunk of an epilogue: accumulator values -> (+bias | +dsigma * head row, both PRELOADED in `b` so the
// loads are in flight before the accumulator arrives) -> relu / mask -> 16 packed bf16x2 words.
template <int EPI>
__device__ __forceinline__ void epi2_chunk(const uint32_t (&v)[32], const float4 (&b)[8], const float* __restrict__ hw,
                                           int n_out, int nheads, float dsig, uint32_t mw_in, uint32_t& mw_out,
                                           float (&hacc)[5], uint32_t (&pk)[16], int cb, const float* __restrict__ semrow,
                                           bool want_mask = true) {
  constexpr bool kFwd = EPI <= DLN_EPI_RELU_OUT;
  constexpr bool kRelu = EPI == DLN_EPI_RELU || EPI == DLN_EPI_RELU_SIGMA || EPI == DLN_EPI_RELU_RGB || EPI == DLN_EPI_RELU_OUT;
  float f[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
  if (kFwd) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      add2(f[4 * q], f[4 * q + 1], b[q].x, b[q].y);
      add2(f[4 * q + 2], f[4 * q + 3], b[q].z, b[q].w);
    }
  }
  if (EPI == DLN_EPI_BWD_MASK_SIGMA) {
#pragma unroll
    for (int q = 0; q < 8; ++q)
      f[4 * q] += dsig * b[q].x, f[4 * q + 1] += dsig * b[q].y, f[4 * q + 2] += dsig * b[q].z, f[4 * q + 3] += dsig * b[q].w;
    if (semrow != nullptr) {      // semantic head: dH += dsem Sw, one fp32 row per ray (dln_sem_head_bwd)
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 h = __ldg(reinterpret_cast<const float4*>(semrow + cb + 4 * q));
        add2(f[4 * q], f[4 * q + 1], h.x, h.y);
        add2(f[4 * q + 2], f[4 * q + 3], h.z, h.w);
      }
    }
  }
  if (kRelu && want_mask) {     // (an inference pass keeps no masks: ~40 of a chunk's ~90 instructions)
    uint32_t n0 = 0, n1 = 0, n2 = 0, n3 = 0;      // sign bits -> mask word, one funnel shift per element
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
  // heads: fixed counts for the sigma (1) and rgb (3) steps, run-time count for output_linear (and 0 for plain ReLU)
  constexpr int kHeads = EPI == DLN_EPI_RELU_SIGMA ? 1 : EPI == DLN_EPI_RELU_RGB ? 3 : EPI == DLN_EPI_RELU_OUT ? 5 : 0;
  if (kHeads > 0 && (EPI != DLN_EPI_RELU_OUT || nheads > 0)) {
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
#pragma unroll
    for (int h = 0; h < kHeads; ++h)
      if (EPI != DLN_EPI_RELU_OUT || h < nheads) {
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
// 32 consecutive columns (16 packed words) of one row into a slab image, 32-bit shared addressing: `row_addr` = slab base
// + 128-byte row + ((row & 7) << 4), so chunk c of the row sits at row_addr ^ (c << 4)
__device__ __forceinline__ void sts_packed32(uint32_t row_addr, int ch0, const uint32_t (&pk)[16]) {
#pragma unroll
  for (int q = 0; q < 4; ++q)
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_addr ^ (uint32_t)((ch0 + q) << 4)), "r"(pk[4 * q]),
                 "r"(pk[4 * q + 1]), "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3])
                 : "memory");
}

// ---------------------------------------------------------------------------------------------
// kNarrow: the plan has zero-padded layers (netwidth < 256).  A separate instantiation, so that the register allocation
// of the full-width kernels is exactly what it was without that code (at 112 registers the epilogue is on the edge).
template <bool kBwd, bool kSem, bool kNarrow = false>
__global__ void __launch_bounds__(k2Threads, 1)
    chain2_kernel(const __grid_constant__ DlnChainProgram prog, const __grid_constant__ DlnChainArgs args,
                  const long long n_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* abuf = smem;                                  // [slot][4] activation slabs
  uint8_t* aux = smem + 8 * kSlab;                       // [slot] aux slab
  uint8_t* wring = smem + 10 * kSlab;                    // k2Stages x 16 KB
  Chain2Small* sm = reinterpret_cast<Chain2Small*>(wring + k2Stages * k2StageBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const long long n_pairs = gridDim.x >> 1, pair = blockIdx.x >> 1;
  const long long t4 = (n_tiles + 3) / 4;                  // groups of four tiles (the last one may be partly padding)
  const long long n_rounds = t4 > pair ? (t4 - pair + n_pairs - 1) / n_pairs : 0;
  const int n_steps = prog.n_steps;
  const bool keep = args.stash != nullptr && prog.stash_slots > 0;
  const int reload_step = (!kBwd && prog.use_viewdirs) ? prog.reload_step : -1;
  // tiles of round r: 4 consecutive tiles per pair, two per CTA
  auto tile_of = [&](long long r, int slot) -> long long { return (r * n_pairs + pair) * 4 + rank * 2 + slot; };

  if (threadIdx.x == 0) {
    for (int i = 0; i < k2Stages; ++i) mbar_init(&sm->w_full[i], 1), mbar_init(&sm->w_empty[i], 1), mbar_init(&sm->w_ready[i], 2);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sm->a_ready[i][0], 8 * k2WG), mbar_init(&sm->a_ready[i][1], 8 * k2WG);
      mbar_init(&sm->acc_full[i], 1);
      for (int k = 0; k < 2; ++k) mbar_init(&sm->s_ready[i][k], 4 * k2WG), mbar_init(&sm->s_free[i][k], 1);
    }
    mbar_fence_init();
  }
  if (args.trace != nullptr && blockIdx.x == 0)
    for (int i = threadIdx.x; i < k2TraceSteps * 16; i += blockDim.x) sm->tr[i] = 0;
  if (warp == 1) {
    tmem_alloc2(&sm->tmem_base, 512);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync_all();            // both CTAs' barriers are initialised before anybody arrives remotely
  tc_fence_after();
  const uint32_t tmem_base = sm->tmem_base;

  // weight-arrival helper: local landing of a stage -> the leader's w_ready.  A lane of its own warp where there is one
  // (it sits on the critical path of the first pass over a layer's weights: the ring holds exactly one layer, so a
  // stage is refilled only ~2 k cycles before it is needed again).
  auto helper_loop = [&]() {
    uint32_t L = 0;
    uint32_t remote[k2Stages];
#pragma unroll
    for (int i = 0; i < k2Stages; ++i) remote[i] = mapa_u32(smem_u32(&sm->w_ready[i]), 0);
    for (long long r = 0; r < n_rounds; ++r)
      for (int s = 0; s < n_steps; ++s) {
        const int nk = prog.steps[s].nk;
        const int n = nk <= k2Stages ? nk : 2 * nk;
        for (int j = 0; j < n; ++j, ++L) {
          const uint32_t stage = L % k2Stages, ph = (L / k2Stages) & 1;
          mbar_wait2(&sm->w_full[stage], ph);
          mbar_arrive_cluster(stage == 0 ? remote[0] : stage == 1 ? remote[1] : stage == 2 ? remote[2] : remote[3]);
        }
      }
  };

  if (warp < k2EpiWarp0) {
  if (k2Rebalance) asm volatile("setmaxnreg.dec.sync.aligned.u32 24;" ::: "memory");
  if (warp == 0) {
    // ===================================================== weight producer: this CTA's half (n_out/2 rows) of every K slab.
    // Steps with <= k2Stages slabs keep their weights resident for both slots (loaded once per step); longer steps
    // (the skip layer: 5 slabs) stream them once per slot.
    if (lane == 0) {
      uint32_t L = 0;
      const uint8_t* wb = reinterpret_cast<const uint8_t*>(args.wblob);
      for (long long r = 0; r < n_rounds; ++r)
        for (int s = 0; s < n_steps; ++s) {
          const DlnChainStep& st = prog.steps[s];
          const uint32_t bytes = (uint32_t)st.n_out * 64u;
          const int passes = st.nk <= k2Stages ? 1 : 2;
          for (int pass = 0; pass < passes; ++pass)
            for (int j = 0; j < st.nk; ++j, ++L) {
              const uint32_t stage = L % k2Stages, ph = (L / k2Stages) & 1;
              mbar_wait_cl(&sm->w_empty[stage], ph ^ 1);
              mbar_expect_tx(&sm->w_full[stage], bytes);
              bulk_g2s(wring + stage * k2StageBytes, wb + st.w_off + (size_t)j * (2 * bytes) + (size_t)rank * bytes, bytes,
                       &sm->w_full[stage]);
            }
        }
    } else if (lane == 1 && k2EpiWarp0 == 3) {
      helper_loop();       // no spare service warp in the 11-warp layout: second lane of the producer warp
    }
  } else if (warp == 3 && k2EpiWarp0 == 4) {
    if (lane == 0) helper_loop();
  } else if (warp == 1) {
    // ===================================================== MMA issuer (leader CTA): X(s), Y(s), X(s+1), ...
    if (rank == 0) {
      uint32_t L = 0, nev = 0;
      const uint64_t desc_k = umma_desc_sw128(0, 16, 1024);      // K-major SWIZZLE_128B, SBO 1024 B
      const uint32_t a_addr0 = smem_u32(abuf), aux_addr0 = smem_u32(aux), ring_addr0 = smem_u32(wring);
      for (long long r = 0; r < n_rounds; ++r)
        for (int s = 0; s < n_steps; ++s, ++nev) {
          const DlnChainStep& st = prog.steps[s];
          const int nk = st.nk;
          const bool reuse = nk <= k2Stages;
          const uint32_t idesc = umma_idesc_bf16(256, st.n_out, 0, 0);
#pragma unroll 1
          for (int slot = 0; slot < 2; ++slot) {
            // operands arrive in two halves (activation slabs 0, 1 first): the MMAs on K slabs 0, 1 of the next step start
            // while the epilogue still converts the columns of slabs 2, 3
            bool seen_lo = false, seen_hi = false;
            const uint32_t d_tmem = tmem_base + slot * 256;
            const uint32_t Lb = (reuse || slot == 0) ? L : L + nk;
            const bool first_use = !(reuse && slot == 1), release = !reuse || slot == 1;
            for (int j = 0; j < nk; ++j) {
              const uint32_t l = Lb + j, stage = l % k2Stages, ph = (l / k2Stages) & 1;
              if (first_use) {
                mbar_wait_cl(&sm->w_ready[stage], ph);
                tc_fence_after();
              }
              if (j == 0) trace2(sm, args.trace, nev, slot * 8 + 0);
              const int slab = st.kslab[j], kc = st.kcnt[j];
              if (slab < 2 && !seen_lo) {
                mbar_wait_cl(&sm->a_ready[slot][0], nev & 1);
                tc_fence_after();
                seen_lo = true;
              } else if (slab >= 2 && !seen_hi) {
                mbar_wait_cl(&sm->a_ready[slot][1], nev & 1);
                tc_fence_after();
                seen_hi = true;
              }
              const uint64_t ad = desc_k | (uint64_t)((slab < 4 ? a_addr0 + (slot * 4 + slab) * kSlab : aux_addr0 + slot * kSlab) >> 4);
              const uint64_t bd = desc_k | (uint64_t)((ring_addr0 + stage * k2StageBytes) >> 4);
              if (elect_one()) {
                umma2_bf16(d_tmem, ad, bd, idesc, j != 0);
                if (kc > 1) umma2_bf16(d_tmem, ad + 2, bd + 2, idesc, 1);
                if (kc > 2) umma2_bf16(d_tmem, ad + 4, bd + 4, idesc, 1);
                if (kc > 3) umma2_bf16(d_tmem, ad + 6, bd + 6, idesc, 1);
                if (release) umma2_commit_mc(&sm->w_empty[stage]);
              }
              __syncwarp();
            }
            if (elect_one()) umma2_commit_mc(&sm->acc_full[slot]);
            __syncwarp();
            // both halves complete exactly once per (slot, step): observe the one this step did not read
            if (!seen_lo) mbar_wait_cl(&sm->a_ready[slot][0], nev & 1);
            if (!seen_hi) mbar_wait_cl(&sm->a_ready[slot][1], nev & 1);
            trace2(sm, args.trace, nev, slot * 8 + 1);
          }
          L += reuse ? nk : 2 * nk;
        }
    }
  } else if (warp == 2) {
    // ===================================================== stash lane (training): walks the productions of this CTA in
    // the order the epilogue warps make them and copies each slab image to the global stash (bulk smem -> global).
    if (lane == 0 && keep) {
      uint8_t* stash = reinterpret_cast<uint8_t*>(args.stash);
      uint32_t cnt[2][2] = {{0, 0}, {0, 0}};
      auto event = [&](int slot, int kind, long long tile, int slot0, int nslab) {
        mbar_wait2(&sm->s_ready[slot][kind], cnt[slot][kind] & 1);
        ++cnt[slot][kind];
        if (slot0 >= 0 && tile < n_tiles) {
          uint8_t* dst = stash + ((size_t)tile * prog.stash_slots + slot0) * kSlab;
          const uint8_t* src = kind == 0 ? abuf + (size_t)slot * 4 * kSlab : aux + (size_t)slot * kSlab;
          for (int i = 0; i < nslab; ++i) bulk_s2g(dst + (size_t)i * kSlab, src + (size_t)i * kSlab, kSlab);
          bulk_commit();
          // Release the buffer as soon as ITS copies have read shared memory.  (Releasing it when the next event's copies
          // are issued, as the one-tile kernel does, ties the release of X to the production of Y -- and the epilogue
          // of X(s+1) starts right after Y(s) is produced, so it found its buffer still held: 2-3 k cycles per step.)
          bulk_wait_read0();
        }
        mbar_arrive(&sm->s_free[slot][kind]);
      };
      auto prologue_events = [&](int slot, long long tile) {
        if (!kBwd) {
          event(slot, 1, tile, 0, 1);                                     // encoded position -> slot 0
        } else {
          event(slot, 0, tile, prog.pro_slot, prog.use_viewdirs ? 2 : 4); // dZ of the first backward layer
          event(slot, 1, tile, 0, 1);                                     // d_raw slab -> slot 0
        }
      };
      for (long long r = 0; r < n_rounds; ++r) {
        if (r == 0)
          for (int slot = 0; slot < 2; ++slot) prologue_events(slot, tile_of(0, slot));
        for (int s = 0; s < n_steps; ++s) {
          const DlnChainStep& st = prog.steps[s];
          const bool last = s == n_steps - 1;
          for (int slot = 0; slot < 2; ++slot) {
            const long long tile = tile_of(r, slot);
            if (!last || st.stash_slot >= 0) event(slot, 0, tile, st.stash_slot, st.n_out >> 6);
            if (s == reload_step) event(slot, 1, tile, 1, 1);             // encoded direction -> slot 1
            if (last && r + 1 < n_rounds) prologue_events(slot, tile_of(r + 1, slot));
          }
        }
      }
      bulk_wait_all0();
    }
  }
  } else {
    if (k2Rebalance) asm volatile("setmaxnreg.inc.sync.aligned.u32 112;" ::: "memory");
    // ===================================================== prologue + epilogue warps: k2WG warpgroups, warpgroup g owns
    // the columns [g n_out/k2WG, (g+1) n_out/k2WG) of both slots, thread = row
    const int et = threadIdx.x - k2EpiWarp0 * 32;
    const int g = et >> 7;                          // warpgroup
    const int r = ((warp & 3) << 5) | lane;         // tile row == TMEM lane
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t ready_remote0 = mapa_u32(smem_u32(&sm->a_ready[0][0]), 0), ready_remote1 = mapa_u32(smem_u32(&sm->a_ready[1][0]), 0);
    // (the second half's barrier of a slot follows the first: + 8 bytes)
    const uint32_t abuf_addr = smem_u32(abuf);
    const uint32_t row_sw = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((r & 7) << 4));
    uint32_t cntA0 = 0, cntA1 = 0, cntX0 = 0, cntX1 = 0;      // productions of A[slot] / AUX[slot] so far
    uint32_t nev = 0;                                          // (round, step) counter: parity of acc_full
    float dsig0 = 0.f, dsig1 = 0.f, sig0 = 0.f, sig1 = 0.f;    // per-slot state carried across the steps of a tile
    const float* sem0 = nullptr;
    const float* sem1 = nullptr;

    auto begin_A = [&](int slot) {
      const uint32_t c = slot ? cntA1 : cntA0;
      if (keep && c > 0) mbar_wait2(&sm->s_free[slot][0], (c - 1) & 1);
    };
    auto end_A = [&](int slot) {           // the smem image is complete for the stash lane
      if (slot) ++cntA1; else ++cntA0;
      if (keep) {
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm->s_ready[slot][0]);
      }
    };
    // end_A + hand-off with ONE proxy fence: the activation slabs are complete for the stash lane and for the MMAs
    // (`lo_too`: the first half has not been handed over yet)
    auto end_A_and_ready = [&](int slot, bool lo_too) {
      if (slot) ++cntA1; else ++cntA0;
      tc_fence_before();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (keep) mbar_arrive(&sm->s_ready[slot][0]);
        const uint32_t a = slot ? ready_remote1 : ready_remote0;
        if (lo_too) mbar_arrive_cluster(a);
        mbar_arrive_cluster(a + 8);
      }
    };
    // first half (slabs 0, 1) of the slot's next operands is in place and ALL accumulator reads of this thread are done
    auto arrive_lo = [&](int slot) {
      tc_fence_before();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(slot ? ready_remote1 : ready_remote0);
    };
    auto begin_X = [&](int slot) {
      const uint32_t c = slot ? cntX1 : cntX0;
      if (keep && c > 0) mbar_wait2(&sm->s_free[slot][1], (c - 1) & 1);
    };
    auto end_X = [&](int slot) {
      if (slot) ++cntX1; else ++cntX0;
      if (keep) {
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm->s_ready[slot][1]);
      }
    };
    // everything this warp wrote for the slot's next MMA step (or prologue) is in place; its accumulator reads are done
    auto arrive_ready = [&](int slot) {
      tc_fence_before();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        const uint32_t a = slot ? ready_remote1 : ready_remote0;
        mbar_arrive_cluster(a);
        mbar_arrive_cluster(a + 8);
      }
    };
    // word of the ReLU-mask block of (mask_slot, tile) that covers columns [col0, col0 + 32) of this thread's row
    auto mask_idx = [&](int col0) -> int { return (((col0 & 127) >> 5) * 128 + r) * 2 + (col0 >> 7); };
    // shared address of this thread's row in the slab that holds column col0 of the slot (swizzle phase folded in)
    auto row_addr_of = [&](int slot, int col0) -> uint32_t { return abuf_addr + (uint32_t)(slot * 4 + (col0 >> 6)) * kSlab + row_sw; };
    // quarter q (columns [16q, 16q+16)) of one encoded row (which = 0 position / 1 direction) of point p
    auto encoded_quarter = [&](long long p, bool valid, int which, int q, uint32_t (&pk)[8]) {
      float e[16];
      if (16 * q >= 3 + 6 * (which == 0 ? prog.L_pts : prog.L_dir)) {
        // quarter beyond the encoding's width (the direction encoding has 27 columns: quarters 2, 3 are padding): no
        // loads, no sincosf -- these warps share their schedulers with the warps that do encode
#pragma unroll
        for (int i = 0; i < 8; ++i) pk[i] = 0u;
        return;
      }
      if (args.x == nullptr) {
        float vx = 0.f, vy = 0.f, vz = 0.f;
        if (valid) {
          const long long ray = (unsigned)p / (unsigned)args.S;      // P < 2^31 (checked by the host side)
          const float* rp = args.rays + (size_t)ray * args.ray_stride;
          if (which == 0) {
            float zz;
            if (args.z_gen != nullptr) {
              // fused stratified sampling (run_nerf.py:571-593): the depth of this point is computed here -- every
              // warpgroup of the row gets the same bits -- and the owner of the first quarter writes it out
              const int i = (int)((unsigned)p - (unsigned)ray * (unsigned)args.S);
              float tr[1] = {0.f};
              if (args.z_rng_state != nullptr)
                rng_fill<1, false>(rng_key(RngRef{args.z_rng_state, args.z_rng_offset}), (unsigned long long)p, tr);
              zz = stratified_z_point(rp[6], rp[7], i, args.S, args.z_lindisp, args.z_rng_state != nullptr, tr[0]);
              if (q == 0) args.z_gen[p] = zz;
            } else {
              zz = args.z[p];
            }
            vx = rp[0] + rp[3] * zz, vy = rp[1] + rp[4] * zz, vz = rp[2] + rp[5] * zz;
          } else {
            vx = rp[args.vd_col], vy = rp[args.vd_col + 1], vz = rp[args.vd_col + 2];
          }
        }
        encode_quarter(vx, vy, vz, which == 0 ? prog.L_pts : prog.L_dir, q, e);
        if (!valid) {
#pragma unroll
          for (int i = 0; i < 16; ++i) e[i] = 0.f;
        }
      } else {
        const int n_pts = 3 + 6 * prog.L_pts, n_dir = 3 + 6 * prog.L_dir;
        const int n = which == 0 ? n_pts : n_dir;
        const float* xp = args.x + (size_t)(valid ? p : 0) * args.x_ld + (which == 0 ? 0 : n_pts);
#pragma unroll
        for (int i = 0; i < 16; ++i) e[i] = (valid && 16 * q + i < n) ? xp[16 * q + i] : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) pk[i] = pack_bf16(e[2 * i], e[2 * i + 1]);
    };
    // this warpgroup's 4 / k2WG quarters of the encoded row -> aux slab of the slot
    auto produce_enc = [&](int slot, long long p, bool valid, int which) {
      begin_X(slot);
#pragma unroll 1
      for (int qq = 0; qq < 4 / k2WG; ++qq) {
        const int q = g * (4 / k2WG) + qq;
        uint32_t pk[8];
        encoded_quarter(p, valid, which, q, pk);
        store_quarter(aux + (size_t)slot * kSlab, r, q, pk);
      }
      end_X(slot);
    };

    // tile prologue of a slot: the operands of step 0 (forward: encoded position in AUX; backward: d raw -> dZ of the
    // first backward layer through rgb_linear / output_linear in A, d_raw slab in AUX for the stash)
    auto prologue = [&](int slot, long long tile) {
      const long long p = tile * DLN_TILE_ROWS + r;
      const bool valid = p < args.P;
      if (!kBwd) {
        produce_enc(slot, p, valid, 0);
      } else {
        float dr[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) dr[j] = (valid && j < prog.out_ch) ? args.d_out[(size_t)p * prog.out_ch + j] : 0.f;
        if (slot) dsig1 = dr[3]; else dsig0 = dr[3];
        const int nh = prog.use_viewdirs ? 3 : prog.out_ch;
        const int width = prog.use_viewdirs ? 128 : 256;
        const int pvalid = (kNarrow && prog.pro_valid > 0) ? prog.pro_valid : width;     // head rows are pvalid wide; the rest is padding
        const int cpw = width / k2WG;
        if (kSem) {
          const float* sp = valid ? args.sem_g + (size_t)((unsigned)p / (unsigned)args.sem_g_div) * 256 : nullptr;
          if (slot) sem1 = sp; else sem0 = sp;
          if (sp != nullptr) {
            for (int c = 0; c < 256 / k2WG; c += 32) asm volatile("prefetch.global.L1 [%0];" ::"l"(sp + g * (256 / k2WG) + c));
          }
        }
        const float* ph = args.fblob + prog.pro_head_off;
        const uint32_t* mp = reinterpret_cast<const uint32_t*>(args.masks) + ((size_t)prog.pro_mask_slot * n_tiles + (tile < n_tiles ? tile : 0)) * 1024;
        begin_A(slot);
#pragma unroll 1
        for (int c = 0; c < (cpw >> 5); ++c) {
          const int col0 = g * cpw + 32 * c;
          const uint32_t mk = (tile < n_tiles && col0 < pvalid) ? mp[mask_idx(col0)] : 0u;
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = 0.f;
#pragma unroll
          for (int j = 0; j < 5; ++j)
            if (j < nh && col0 < pvalid) {
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 w4 = __ldg(reinterpret_cast<const float4*>(ph + j * pvalid + col0 + 4 * q));
                f[4 * q] += dr[j] * w4.x, f[4 * q + 1] += dr[j] * w4.y, f[4 * q + 2] += dr[j] * w4.z, f[4 * q + 3] += dr[j] * w4.w;
              }
            }
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = (mk & (1u << i)) ? f[i] : 0.f;
          uint32_t pk[16];
          pack32<false>(f, pk);
          sts_packed32(row_addr_of(slot, col0), (col0 & 63) >> 3, pk);
        }
        end_A(slot);
        if (keep) {                      // the d_raw slab only feeds the stash (head wgrad items)
          begin_X(slot);
          if (g == k2WG - 1) {
            float e[64];
#pragma unroll
            for (int i = 0; i < 64; ++i) e[i] = 0.f;
#pragma unroll
            for (int j = 0; j < 5; ++j) e[j] = dr[j];
            store_row64(aux + (size_t)slot * kSlab, nullptr, r, e);
          }
          end_X(slot);
        }
      }
    };

    if (n_rounds > 0) {
      for (int slot = 0; slot < 2; ++slot) {
        prologue(slot, tile_of(0, slot));
        arrive_ready(slot);
      }
    }
    for (long long rnd = 0; rnd < n_rounds; ++rnd) {
      // per-round invariants of the two slots (kept out of the step loop: the epilogue is instruction-issue bound)
      const long long tile0 = tile_of(rnd, 0), tile1 = tile0 + 1;
      const bool ok0 = tile0 < n_tiles, ok1 = tile1 < n_tiles;
      uint32_t* const mtile0 = reinterpret_cast<uint32_t*>(args.masks) + (size_t)(ok0 ? tile0 : 0) * 1024;
      uint32_t* const mtile1 = reinterpret_cast<uint32_t*>(args.masks) + (size_t)(ok1 ? tile1 : 0) * 1024;
      for (int s = 0; s < n_steps; ++s, ++nev) {
        const DlnChainStep& st = prog.steps[s];
        const int epi = st.epi, n_out = st.n_out, mask_slot = st.mask_slot;
        const int n_valid = (kNarrow && st.n_valid32) ? 32 * st.n_valid32 : n_out;     // netwidth < 256: the columns beyond are padding
        const bool last = s == n_steps - 1;
        const float* hw = args.fblob + st.head_off;
        const int nheads = (epi <= DLN_EPI_RELU_OUT) ? st.n_heads : 0;
        const bool relu = (epi == DLN_EPI_RELU || epi == DLN_EPI_RELU_SIGMA || epi == DLN_EPI_RELU_RGB || epi == DLN_EPI_RELU_OUT);
        // Columns of a thread: chunk c = [32 k2WG c + 32 g, + 32) -- chunk c of ALL warpgroups together covers 32 k2WG
        // consecutive columns, so the activation slabs fill up in order (slabs 0, 1 by the first half of the chunks)
        const int nch = n_out / (32 * k2WG);           // 32-column chunks per thread: 1, 2 or 4
        auto col_of = [&](int c) -> int { return c * 32 * k2WG + 32 * g; };
        const bool write_a = !last || (keep && st.stash_slot >= 0);
        // bias (forward) or alpha head row (dgrad sigma step): added per column, preloaded ahead of the accumulator
        const float* const brow = kBwd ? hw : args.fblob + st.bias_off;
        const bool need_b = !kBwd || epi == DLN_EPI_BWD_MASK_SIGMA;
        const bool use_mi = kBwd && epi >= DLN_EPI_BWD_MASK && mask_slot >= 0;
        const size_t mask_step_off = (size_t)(mask_slot < 0 ? 0 : mask_slot) * (size_t)n_tiles * 1024;
        // steps with a compile-time specialised epilogue: the plans of every shipped configuration consist of these
        const bool fast = (kNarrow && n_valid != n_out) ? false : kBwd ? n_out == 256 && (epi == DLN_EPI_BWD_MASK || epi == DLN_EPI_BWD_MASK_SIGMA) && mask_slot >= 0
                               : (n_out == 256 && ((epi == DLN_EPI_RELU && nheads == 0) || (epi == DLN_EPI_RELU_SIGMA && nheads == 1))) ||
                                     (n_out == 128 && epi == DLN_EPI_RELU_RGB && nheads == 3 && kFastChunks >= 2);
#pragma unroll 1
        for (int slot = 0; slot < 2; ++slot) {
          const long long tile = slot ? tile1 : tile0;
          const long long p = tile * DLN_TILE_ROWS + r;
          const bool valid = p < args.P, tile_ok = slot ? ok1 : ok0;
          const float dsig = slot ? dsig1 : dsig0;
          const float* semrow = kSem ? (slot ? sem1 : sem0) : nullptr;
          uint32_t* const mblock = (slot ? mtile1 : mtile0) + mask_step_off;
          const bool mask_in = use_mi && tile_ok, mask_out = relu && mask_slot >= 0 && args.masks != nullptr && tile_ok;
          const uint32_t t_acc = tmem_base + slot * 256 + lane_addr;
          float hacc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
          bool simple = false;
          // Early prologue (forward): the next tile's positions are encoded into registers BEFORE this (last) step's
          // accumulator is waited for -- the coarse network's epilogue warps are idle there (its per-slot MMA -> epilogue
          // chain is the bound), so the ~4 k cycles of sincosf per tile leave the critical path.  Stored after the heads.
          uint32_t epk[4 / k2WG][8];
          const bool pre = !kBwd && last && rnd + 1 < n_rounds;
          if (pre) {
            const long long pn = (tile + 4 * n_pairs) * DLN_TILE_ROWS + r;
#pragma unroll
            for (int qq = 0; qq < 4 / k2WG; ++qq) encoded_quarter(pn, pn < args.P, 0, g * (4 / k2WG) + qq, epk[qq]);
          }
          float4 bq[8];
          auto load_b = [&](int c) {
#pragma unroll
            for (int q = 0; q < 8; ++q) bq[q] = __ldg(reinterpret_cast<const float4*>(brow + col_of(c) + 4 * q));
          };

          if (fast) {
            // ---------------------------------------------------------------- common steps (256 outputs, no heads): the chunk
            // pipeline fully unrolled with the chunk count and the epilogue type fixed at compile time -- the next chunk's
            // accumulator columns and bias are in flight while the current chunk is converted.  (With run-time chunk
            // counts and flags the loop carried ~350 instructions of control flow per thread and step next to its ~300
            // of arithmetic, and the epilogue is instruction-issue bound.)
            auto fast_path = [&](auto epi_tag, auto nch_tag) {
              constexpr int E = decltype(epi_tag)::value;
              constexpr int NCH = decltype(nch_tag)::value;
              constexpr bool kNeedB = E <= DLN_EPI_RELU_OUT || E == DLN_EPI_BWD_MASK_SIGMA;
              constexpr bool kMaskIn = E >= DLN_EPI_BWD_MASK, kMaskOut = E <= DLN_EPI_RELU_OUT && E != DLN_EPI_LINEAR;
              uint32_t mi[NCH], mo[NCH];
#pragma unroll
              for (int c = 0; c < NCH; ++c) mi[c] = 0, mo[c] = 0;
              if (kMaskIn && tile_ok) {
#pragma unroll
                for (int c = 0; c < NCH; ++c) mi[c] = mblock[mask_idx(col_of(c))];
              }
              if (kNeedB) load_b(0);
              mbar_wait_cl(&sm->acc_full[slot], nev & 1);
              tc_fence_after();
              if (et == 0) trace2(sm, args.trace, nev, slot * 8 + 2);
              // With two chunks per thread BOTH are loaded up front: once they have landed nothing of this thread reads the
              // accumulator any more, and the first half of the operands can be handed over after chunk 0.
              constexpr bool kEarly = NCH == 2;
              uint32_t va[32], vb[32], pk[16];
              tmem_ld32(t_acc + col_of(0), va);
              if (kEarly) tmem_ld32(t_acc + col_of(1), vb);
              if (write_a) begin_A(slot);
              tmem_ld_wait();
              tmem_ld_pin32(va);
              if (kEarly) tmem_ld_pin32(vb);
              simple = write_a && !last && s != reload_step;
              if (et == 0) trace2(sm, args.trace, nev, slot * 8 + 5);
#pragma unroll
              for (int c = 0; c < NCH; ++c) {
                uint32_t(&cur)[32] = (c & 1) ? vb : va;
                uint32_t(&nxt)[32] = (c & 1) ? va : vb;
                if (!kEarly && c + 1 < NCH) tmem_ld32(t_acc + col_of(c + 1), nxt);
                const int c0 = col_of(c);
                epi2_chunk<E>(cur, bq, hw, NCH * 32 * k2WG, 0, dsig, mi[c], mo[c], hacc, pk, c0, semrow, mask_out);
                if (c == 0 && et == 0) trace2(sm, args.trace, nev, slot * 8 + 6);
                if (c + 1 < NCH && kNeedB) load_b(c + 1);
                if (write_a) sts_packed32(row_addr_of(slot, c0), (c0 & 63) >> 3, pk);
                if (!kEarly && c + 1 < NCH) {
                  tmem_ld_wait();
                  tmem_ld_pin32(nxt);
                }
                if (kEarly && c == 0 && simple) arrive_lo(slot);
                if (c == 0 && et == 0) trace2(sm, args.trace, nev, slot * 8 + 7);
              }
              // plain step: one fence + one arrival hands the slabs to the stash lane and the (rest to the) MMA warp
              if (simple) end_A_and_ready(slot, !kEarly);
              else if (write_a) end_A(slot);
              if (et == 0) trace2(sm, args.trace, nev, slot * 8 + 3);
              // the ReLU masks go out AFTER the hand-off: nothing in this kernel reads them
              if (kMaskOut && mask_out) {
#pragma unroll
                for (int c = 0; c < NCH; ++c) mblock[mask_idx(col_of(c))] = mo[c];
              }
            };
            using Wide = std::integral_constant<int, kFastChunks>;
            using Half = std::integral_constant<int, kFastChunks / 2>;
            if (!kBwd) {
              if (epi == DLN_EPI_RELU) fast_path(std::integral_constant<int, DLN_EPI_RELU>{}, Wide{});
              else if (epi == DLN_EPI_RELU_SIGMA) fast_path(std::integral_constant<int, DLN_EPI_RELU_SIGMA>{}, Wide{});
              else fast_path(std::integral_constant<int, DLN_EPI_RELU_RGB>{}, Half{});
            } else {
              if (epi == DLN_EPI_BWD_MASK) fast_path(std::integral_constant<int, DLN_EPI_BWD_MASK>{}, Wide{});
              else fast_path(std::integral_constant<int, DLN_EPI_BWD_MASK_SIGMA>{}, Wide{});
            }
          } else {
            // ---------------------------------------------------------------- head steps (sigma / rgb / output_linear) and the
            // unfolded plan's linear / copy steps: one chunk at a time, every epilogue variant behind one call site
            mbar_wait_cl(&sm->acc_full[slot], nev & 1);
            tc_fence_after();
            if (et == 0) trace2(sm, args.trace, nev, slot * 8 + 2);
            if (write_a) begin_A(slot);
#pragma unroll 1
            for (int c = 0; c < nch; ++c) {
              const int c0 = col_of(c);
              uint32_t v[32], pk[16], mo = 0;
              if (kNarrow && c0 >= n_valid) {      // padding columns of a narrow layer: zeros, no bias, no head, mask 0
#pragma unroll
                for (int i = 0; i < 16; ++i) pk[i] = 0u;
                if (write_a) sts_packed32(row_addr_of(slot, c0), (c0 & 63) >> 3, pk);
                if (mask_out) mblock[mask_idx(c0)] = 0u;
                continue;
              }
              tmem_ld32(t_acc + c0, v);
              if (need_b) load_b(c);
              const uint32_t mi = mask_in ? mblock[mask_idx(c0)] : 0u;
              tmem_ld_wait();
              tmem_ld_pin32(v);
              if (!kBwd) {
                if (epi == DLN_EPI_LINEAR) epi2_chunk<DLN_EPI_LINEAR>(v, bq, hw, n_valid, nheads, dsig, mi, mo, hacc, pk, c0, semrow);
                else epi2_chunk<DLN_EPI_RELU_OUT>(v, bq, hw, n_valid, nheads, dsig, mi, mo, hacc, pk, c0, semrow);
              } else {
                if (epi == DLN_EPI_BWD_COPY) epi2_chunk<DLN_EPI_BWD_COPY>(v, bq, hw, n_valid, nheads, dsig, mi, mo, hacc, pk, c0, semrow);
                else if (epi == DLN_EPI_BWD_MASK) epi2_chunk<DLN_EPI_BWD_MASK>(v, bq, hw, n_valid, nheads, dsig, mi, mo, hacc, pk, c0, semrow);
                else epi2_chunk<DLN_EPI_BWD_MASK_SIGMA>(v, bq, hw, n_valid, nheads, dsig, mi, mo, hacc, pk, c0, semrow);
              }
              if (write_a) sts_packed32(row_addr_of(slot, c0), (c0 & 63) >> 3, pk);
              if (mask_out) mblock[mask_idx(c0)] = mo;
            }
            if (write_a) end_A(slot);
            if (et == 0) trace2(sm, args.trace, nev, slot * 8 + 3);
          }

          if (s == reload_step) {
            // every MMA that reads the encoded position has completed (acc_full of this step): the aux slab takes the
            // encoded view direction for the views layer
            produce_enc(slot, p, valid, 1);
          }

          // heads: per-warpgroup partial sums meet in the aux slab (dead once the last step's MMAs have completed)
          if (epi == DLN_EPI_RELU_SIGMA) {
            const float sg = hacc[0] + (g == 0 ? args.fblob[st.head_bias_off] : 0.f);
            if (slot) sig1 = sg; else sig0 = sg;
          } else if (epi == DLN_EPI_RELU_RGB || epi == DLN_EPI_RELU_OUT) {
            float* part = reinterpret_cast<float*>(aux + (size_t)slot * kSlab);      // [k2WG warpgroups][4][128]
            begin_X(slot);                                                               // its last stash copy has read it
            auto total = [&](int h) {
              float t = part[h * 128 + r];
#pragma unroll
              for (int w = 1; w < k2WG; ++w) t += part[(w * 4 + h) * 128 + r];
              return t;
            };
            if (epi == DLN_EPI_RELU_RGB) {
#pragma unroll
              for (int h = 0; h < 3; ++h) part[(g * 4 + h) * 128 + r] = hacc[h];
              part[(g * 4 + 3) * 128 + r] = slot ? sig1 : sig0;
              named_bar_epi2();
              if (g == 0 && valid) {
                float o[4];
#pragma unroll
                for (int h = 0; h < 4; ++h) o[h] = total(h);
#pragma unroll
                for (int h = 0; h < 3; ++h) o[h] += args.fblob[st.head_bias_off + h];
                *reinterpret_cast<float4*>(args.out + (size_t)p * 4) = make_float4(o[0], o[1], o[2], o[3]);
              }
              named_bar_epi2();
            } else {
              for (int h0 = 0; h0 < nheads; h0 += 4) {
#pragma unroll
                for (int h = 0; h < 4; ++h)
                  if (h0 + h < nheads) part[(g * 4 + h) * 128 + r] = (h0 == 0) ? hacc[h] : hacc[4];
                named_bar_epi2();
                if (g == 0 && valid) {
                  for (int h = 0; h < 4 && h0 + h < nheads; ++h)
                    args.out[(size_t)p * prog.out_ch + h0 + h] = total(h) + args.fblob[st.head_bias_off + h0 + h];
                }
                named_bar_epi2();
              }
            }
          }

          if (!last) {
            if (!simple) arrive_ready(slot);
          } else if (rnd + 1 < n_rounds) {
            if (pre) {
              begin_X(slot);
#pragma unroll
              for (int qq = 0; qq < 4 / k2WG; ++qq) store_quarter(aux + (size_t)slot * kSlab, r, g * (4 / k2WG) + qq, epk[qq]);
              end_X(slot);
            } else {
              prologue(slot, tile_of(rnd + 1, slot));
            }
            arrive_ready(slot);
          }
          if (et == 0) trace2(sm, args.trace, nev, slot * 8 + 4);
        }
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  cluster_sync_all();            // nobody leaves (or frees tensor memory) while the peer may still touch this CTA
  tc_fence_after();
  if (warp == 1) tmem_dealloc2(tmem_base, 512);
  if (args.trace != nullptr && blockIdx.x == 0 && threadIdx.x < k2TraceSteps * 16) args.trace[threadIdx.x] = sm->tr[threadIdx.x];
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// launcher (called by dln_mlp_chain in mlp_kernels.cu)
// ---------------------------------------------------------------------------------------------
int dln_chain2_launch(const DlnChainProgram* prog, const DlnChainArgs* args, int num_sms, long long n_tiles, cudaStream_t stream) {
  bool& attr_set = dln_device_flag(3);
  if (!attr_set) {
    const void* fns[5] = {(const void*)chain2_kernel<false, false>, (const void*)chain2_kernel<true, false>,
                          (const void*)chain2_kernel<true, true>, (const void*)chain2_kernel<false, false, true>,
                          (const void*)chain2_kernel<true, false, true>};
    for (const void* f : fns) {
      cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kChain2SmemBytes);
      if (e != cudaSuccess) return (int)e;
    }
    attr_set = true;
  }
  bool narrow = prog->pro_valid != 0;
  for (int s = 0; s < prog->n_steps; ++s)
    narrow = narrow || (prog->steps[s].n_valid32 != 0 && prog->steps[s].n_valid32 * 32 != prog->steps[s].n_out);
  if (narrow && prog->backward && args->sem_g) return DLN_EINVAL;      // the semantic head is built for netwidth 256
  const long long want_pairs = (n_tiles + 3) / 4;
  long long pairs = num_sms / 2;
  if (pairs < 1) pairs = 1;
  if (pairs > want_pairs) pairs = want_pairs;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * pairs)), cfg.blockDim = dim3(k2Threads), cfg.dynamicSmemBytes = kChain2SmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
  cfg.attrs = at, cfg.numAttrs = 1;
  cudaError_t e;
  if (narrow && prog->backward) e = cudaLaunchKernelEx(&cfg, chain2_kernel<true, false, true>, *prog, *args, n_tiles);
  else if (narrow) e = cudaLaunchKernelEx(&cfg, chain2_kernel<false, false, true>, *prog, *args, n_tiles);
  else if (prog->backward && args->sem_g) e = cudaLaunchKernelEx(&cfg, chain2_kernel<true, true>, *prog, *args, n_tiles);
  else if (prog->backward) e = cudaLaunchKernelEx(&cfg, chain2_kernel<true, false>, *prog, *args, n_tiles);
  else e = cudaLaunchKernelEx(&cfg, chain2_kernel<false, false>, *prog, *args, n_tiles);
  if (e != cudaSuccess) return (int)e;
  return dln_launch_status();
}
