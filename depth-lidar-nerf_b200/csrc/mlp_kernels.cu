// NeRF MLP (run_nerf_helpers.py:77-145) on the Blackwell tensor cores.
//
// Three kernels, all tcgen05 (UMMA, fp32 accumulators in TMEM), bf16 operands, weights moved by the bulk-copy (TMA)
// engine as SWIZZLE_128B canonical shared-memory images:
//
//  chain_kernel   the fused per-tile layer chain.  One persistent CTA per SM walks 128-point tiles.  The weights
//                 of every layer are streamed from L2 through a ring of four 32 KB stages (= one whole 256x256
//                 layer, refilled while the epilogue runs).  The epilogue of layer l (TMEM -> bias / ReLU ->
//                 bf16) writes the activations back IN PLACE into tensor memory, where the MMAs of layer l+1
//                 read them as their A operand (TS form) while accumulating into the other of the two TMEM
//                 buffers; shared-memory images of the activations exist only as staging for the stash copies
//                 (training).  The same machine runs the forward pass (positions + positional encoding computed
//                 in-kernel by all four epilogue warpgroups, heads = alpha / rgb on CUDA cores) and the dgrad pass
//                 (prologue = d raw -> d hidden through rgb_linear, epilogue = 1-bit ReLU masks).
//  wgrad_kernel   dW += dZ^T * X over all points, both operands read back from the slab stashes the chain
//                 kernels wrote, as MN-major UMMA operands; split over the points across CTAs.
//  pack_kernel    fp32 master weights -> bf16 swizzled weight stages.
//
// Warp roles in chain_kernel (640 threads): warp 0 weight producer, warp 1 MMA issuer (+TMEM owner), warp 2 stash
// lane, warp 3 weight-arrival helper, warps 4..19 epilogue (four warpgroups; warpgroup g owns the columns
// [128 h + 32 g, +32), h = 0, 1, of every layer output).
#include "chain_common.cuh"
#include <stdlib.h>

using namespace dln;

namespace {

constexpr int kThreads = 640;
constexpr int kEpiWarp0 = 4;
constexpr int kNumStages = 4;      // the ring holds one whole 256x256 layer: it is refilled during the epilogue
constexpr int kStageBytes = 32768;
constexpr int kGrpStages = 2;      // weight arrival is handed to the MMA thread per group of this many stages: with
                                   // whole-layer groups (4) the first MMAs of a layer waited for the refill of the stage
                                   // the PREVIOUS layer released last (~900 cycles per layer in the smem timeline)
constexpr int kNumSlabs = 5;       // 0..3 activations, 4 encoded position -> encoded direction (fwd) / d_raw (bwd)
constexpr bool kEarlyPrologue = true;  // forward: encode the next tile's positions under the last layer's MMAs
#ifndef DLN_DIRECT_STASH
#define DLN_DIRECT_STASH 0
#endif
constexpr bool kDirectStash = DLN_DIRECT_STASH != 0;  // epilogue threads write the stash images straight to global memory instead of
                                      // staging them in smem for bulk copies: parity-green; 50 % slower with 16-byte stores, still 13 % slower (fwd D=8
                                      // 1.00 vs 0.88 ms) with one 256-bit store per 32-byte sector -- kept selectable
constexpr int kMaxBiasFloats = 2432;   // 9 x 256 + 128: netdepth <= 8 with view directions, <= 9 without

struct ChainSmall {
  uint64_t w_full[kNumStages], w_empty[kNumStages];
  uint64_t a_ready[kNumSlabs], s_free[kNumSlabs];
  uint64_t s_ready[4];             // smem staging image of activation slab i complete (epilogue -> stash lane)
  uint64_t acc_full[2];            // accumulator buffer complete (tcgen05.commit of the step's last MMA)
  uint64_t grp_full[kNumStages];   // weights of a group of <=4 stages have landed (helper -> MMA thread); rotating,
                                   // so a parity wait can never alias: at most kNumStages groups are ever in flight
  uint32_t tmem_base;
  uint32_t pad_;
  alignas(16) float bias[kMaxBiasFloats];   // bias vectors of all steps, packed back to back (forward only)
  float part[4][4][128];                    // per-warpgroup partial head sums: rgb 0..2 (or out 0..3), sigma 3
  uint16_t tr[128];                         // debug timeline (see trace_ev): clock / 8, 16 bits
};

constexpr size_t kChainSmemBytes = (size_t)kNumSlabs * kSlab + (size_t)kNumStages * kStageBytes + sizeof(ChainSmall) + 1024;
static_assert(kChainSmemBytes <= 232448, "chain kernel shared memory exceeds the 227 KB per-CTA limit");

__device__ __forceinline__ void named_bar_epi() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

// slabs produced by the tile prologue
__device__ __forceinline__ uint32_t prologue_mask(const DlnChainProgram& p) {
  if (!p.backward) return 0x10u;                 // encoded position
  return p.use_viewdirs ? 0x13u : 0x1Fu;         // d_raw slab + dZ of the first backward layer
}
__device__ __forceinline__ uint32_t step_out_mask(const DlnChainStep& s) { return s.n_out == 256 ? 0xFu : 0x3u; }

// debug timeline of CTA 0: %clock (in units of 8 cycles, 16 bits) at 8 events per step for kTraceSteps consecutive steps starting at gstep kTraceStep0
// (whole tiles, tile boundaries included), kept in shared memory (a clock read + one st.shared per event, so the
// traced schedule is the real one) and copied to args.trace when the kernel ends.  Events: 0 MMA warp enters the
// step, 1 first K slab's operands ready (first MMA issued right after), 2 last MMA + commit issued; epilogue thread 0
// (warpgroup 0): 3 accumulator seen complete, 4 chunk 0 handed to the MMA warp, 5 chunk 1 handed over, 6 stash
// staging done / step left, 7 tile prologue or head epilogue done (first / last step of a tile).
constexpr uint32_t kTraceStep0 = 8, kTraceSteps = 16;
__device__ __forceinline__ void trace_ev(ChainSmall* sm, const long long* trace, int role, uint32_t gstep, int ev) {
  if (trace != nullptr && blockIdx.x == 0 && gstep - kTraceStep0 < kTraceSteps && (role != 0 || (threadIdx.x & 31) == 0)) {
    uint32_t c;
    asm volatile("mov.u32 %0, %%clock;" : "=r"(c));
    sm->tr[(gstep - kTraceStep0) * 8 + ev] = (uint16_t)(c >> 3) | 1u;      // never 0: 0 = event not reached
  }
}

struct ProdTrack {
  uint32_t par = 0, any = 0;  // per-slab parity of the production count / "produced at least once"
};

// ---------------------------------------------------------------------------------------------
// the chain kernel
// ---------------------------------------------------------------------------------------------
template <bool kBwd, bool kSem = false>      // kSem: dgrad with the semantic head's per-ray term (args.sem_g)
__global__ void __launch_bounds__(kThreads, 1)
    chain_kernel(const __grid_constant__ DlnChainProgram prog, const __grid_constant__ DlnChainArgs args,
                 const long long n_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* slabs = smem;                              // 5 x 16 KB
  uint8_t* wring = smem + kNumSlabs * kSlab;          // 4 x 32 KB
  ChainSmall* sm = reinterpret_cast<ChainSmall*>(wring + kNumStages * kStageBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool keep = args.stash != nullptr && prog.stash_slots > 0;
  // forward: the encoded-direction rows replace the encoded position in slab 4 after step `reload_step`
  const int reload_step = (!kBwd && prog.use_viewdirs) ? prog.reload_step : -1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kNumStages; ++i) mbar_init(&sm->w_full[i], 1), mbar_init(&sm->w_empty[i], 1);
    for (int i = 0; i < kNumSlabs; ++i) mbar_init(&sm->a_ready[i], i < 4 ? 8 : (kBwd ? 4 : 16)), mbar_init(&sm->s_free[i], 1);   // one arrival per producing warp
    for (int i = 0; i < 2; ++i) mbar_init(&sm->acc_full[i], 1);
    for (int i = 0; i < 4; ++i) mbar_init(&sm->s_ready[i], 8);
    for (int i = 0; i < kNumStages; ++i) mbar_init(&sm->grp_full[i], 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(&sm->tmem_base, 512);
    tmem_relinquish();
  }
  if (!kBwd && warp >= kEpiWarp0) {       // all bias vectors -> smem, packed in step order
    const int t = threadIdx.x - kEpiWarp0 * 32;
    int base = 0;
    for (int s = 0; s < prog.n_steps; ++s) {
      if (t < prog.steps[s].n_out) sm->bias[base + t] = args.fblob[prog.steps[s].bias_off + t];
      base += prog.steps[s].n_out;
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
    // ===================================================== MMA issuer: the whole warp walks the (warp-uniform) program
    // so that descriptors and addresses live in uniform registers; one elected lane issues tcgen05.mma / commit
    {
      uint32_t stage = 0, gstep = 0, grp = 0;
      uint32_t par = 0;  // parity of the production count per slab
      uint32_t catch_mask = 0, catch_par = 0;
      const uint32_t pmask = prologue_mask(prog);
      const uint64_t desc_k = umma_desc_sw128(0, 16, 1024);      // K-major SWIZZLE_128B, SBO 1024 B
      const uint32_t slab_addr0 = smem_u32(slabs), ring_addr0 = smem_u32(wring);
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        par ^= pmask;
        for (int s = 0; s < prog.n_steps; ++s, ++gstep) {
          const DlnChainStep& st = prog.steps[s];
          if (s == 1 && catch_mask) {      // deferred observation of the previous tile's last productions
            for (int slab = 0; slab < 4; ++slab)
              if ((catch_mask >> slab) & 1) mbar_wait(&sm->a_ready[slab], ((catch_par >> slab) & 1) ^ 1);
            catch_mask = 0;
          }
          const uint32_t idesc = umma_idesc_bf16(128, st.n_out, 0, 0);
          const uint32_t d_tmem = tmem_base + (gstep & 1) * 256;
          trace_ev(sm, args.trace, 0, gstep, 0);
          for (int j = 0; j < st.nk; ++j) {
            const int slab = st.kslab[j];
            mbar_wait(&sm->a_ready[slab], ((par >> slab) & 1) ^ 1);   // last production completed
            if (j == 0) trace_ev(sm, args.trace, 0, gstep, 1);
            if ((j % kGrpStages) == 0) {   // the helper thread has seen w_full of this group of stages
              mbar_wait(&sm->grp_full[grp & (kNumStages - 1)], (grp / kNumStages) & 1);
              ++grp;
            }
            tc_fence_after();
            // descriptors differ only in the 14-bit start-address field: +2 (= 32 B) per K=16 step
            const uint64_t bd = desc_k | (uint64_t)((ring_addr0 + stage * kStageBytes) >> 4);
            const int kc = st.kcnt[j];
            if (slab < 4) {
              // activations of the previous step: bf16 pairs in tensor memory, written in place over the
              // accumulator buffer the previous step used -- each 32-feature group [32q, +32) packed into the
              // first 16 of its own 32 accumulator columns, so K step kk of slab j starts at column
              // 64 j + 32 (kk >> 1) + 8 (kk & 1)
              const uint32_t at = tmem_base + ((gstep + 1) & 1) * 256 + slab * 64;
              if (elect_one()) {
                umma_bf16_ts(d_tmem, at, bd, idesc, j != 0);
                if (kc > 1) umma_bf16_ts(d_tmem, at + 8, bd + 2, idesc, 1);
                if (kc > 2) umma_bf16_ts(d_tmem, at + 32, bd + 4, idesc, 1);
                if (kc > 3) umma_bf16_ts(d_tmem, at + 40, bd + 6, idesc, 1);
                umma_commit(&sm->w_empty[stage]);
              }
            } else {
              // encoded position / direction (fwd) or d_raw (bwd): a shared-memory slab
              const uint64_t ad = desc_k | (uint64_t)((slab_addr0 + slab * kSlab) >> 4);
              if (elect_one()) {
                umma_bf16(d_tmem, ad, bd, idesc, j != 0);
                if (kc > 1) umma_bf16(d_tmem, ad + 2, bd + 2, idesc, 1);
                if (kc > 2) umma_bf16(d_tmem, ad + 4, bd + 4, idesc, 1);
                if (kc > 3) umma_bf16(d_tmem, ad + 6, bd + 6, idesc, 1);
                umma_commit(&sm->w_empty[stage]);
              }
            }
            __syncwarp();
            if (++stage == kNumStages) stage = 0;
          }
          if (elect_one()) umma_commit(&sm->acc_full[gstep & 1]);
          __syncwarp();
          trace_ev(sm, args.trace, 0, gstep, 2);
          par ^= step_out_mask(st);
          if (s == reload_step) par ^= 0x10u;
        }
        // Parity waits are only sound while a waiter is never two phases ahead of the barrier.  No MMA
        // consumes the slabs the LAST step produces (they only go to the stash), so observe them here
        // before waiting for the next tile's productions of the same slabs.
        // Forward chain: the next tile's first layer reads the encoded position only (slab 4, produced early by
        // warpgroup 0 while this tile's last MMAs ran), so it is issued BEFORE these waits and overlaps the last
        // epilogue; the waits then happen ahead of the next tile's step 1 with the parities captured here.
        const uint32_t om = step_out_mask(prog.steps[prog.n_steps - 1]);
        if (kEarlyPrologue && !kBwd) {
          catch_mask = om, catch_par = par;
        } else {
          for (int slab = 0; slab < 4; ++slab)
            if ((om >> slab) & 1) mbar_wait(&sm->a_ready[slab], ((par >> slab) & 1) ^ 1);
        }
      }
    }
  } else if (warp == 2) {
    // ===================================================== stash writer (training only): ONE lane walks the slab
    // productions in program order and copies each to the global stash with a bulk smem -> global copy.  (Five
    // lanes spinning on five different barriers in one warp serialise each other: the divergent paths are only
    // switched every few hundred cycles, which showed up as late s_free arrivals in the epilogue.)  The copies
    // still overlap: slab X is released (s_free) once the NEXT copy has been issued and X's group has finished
    // reading shared memory; the tail of a tile is flushed so the next tile's prologue never waits on it.
    if (lane == 0 && keep && !kDirectStash) {
      uint32_t par = 0, sgstep = 0, pending_step = 0;
      int pending = -1;
      const uint32_t pmask = prologue_mask(prog);
      uint8_t* stash = reinterpret_cast<uint8_t*>(args.stash);
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        uint8_t* tbase = stash + (size_t)tile * prog.stash_slots * kSlab;
        auto flush = [&]() {
          if (pending >= 0) {
            bulk_wait_read0();
            mbar_arrive(&sm->s_free[pending]);
            pending = -1;
          }
        };
        auto handle = [&](int slab, int slot) {
          if (pending == slab) flush();
          mbar_wait(slab < 4 ? &sm->s_ready[slab] : &sm->a_ready[slab], (par >> slab) & 1u);
          par ^= 1u << slab;
          if (slot >= 0) {
            bulk_s2g(tbase + (size_t)slot * kSlab, slabs + slab * kSlab, kSlab);
            bulk_commit();
            if (pending >= 0) {
              bulk_wait_read1();
              mbar_arrive(&sm->s_free[pending]);
            }
            pending = slab;
            pending_step = sgstep;
          } else {
            mbar_arrive(&sm->s_free[slab]);
          }
        };
        for (int slab = 0; slab < kNumSlabs; ++slab)
          if ((pmask >> slab) & 1) handle(slab, slab == 4 ? 0 : prog.pro_slot + slab);
        for (int s = 0; s < prog.n_steps; ++s, ++sgstep) {
          const DlnChainStep& st = prog.steps[s];
          const uint32_t om = step_out_mask(st);
          for (int slab = 0; slab < 4; ++slab)
            if ((om >> slab) & 1) handle(slab, st.stash_slot >= 0 ? st.stash_slot + slab : -1);
          if (s == reload_step) handle(4, 1);                    // encoded direction -> slot 1
          // The last slab of a step is released as soon as ITS copy has read shared memory, not when the next
          // step's first slab arrives: otherwise the warpgroups that own it (2, 3) start their next staging only
          // after warpgroups 0, 1 have finished theirs (smem timeline: ~1000 cycles of lag per layer).
          flush();
        }
        flush();
      }
      bulk_wait_all0();
    }
  } else if (warp == 3) {
    // ===================================================== weight-arrival helper: takes the per-stage
    // mbarrier latency off the MMA thread's critical path (one grp_full wait per <=4 stages instead)
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, grp = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int s = 0; s < prog.n_steps; ++s) {
          const int nk = prog.steps[s].nk;
          for (int j = 0; j < nk; ++j) {
            mbar_wait(&sm->w_full[stage], phase);
            if (++stage == kNumStages) stage = 0, phase ^= 1;
            if ((j % kGrpStages) == kGrpStages - 1 || j == nk - 1) mbar_arrive(&sm->grp_full[grp++ & (kNumStages - 1)]);
          }
        }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================================================== prologue + epilogue warps
    const int et = threadIdx.x - kEpiWarp0 * 32;   // 0..511
    const int g = et >> 7;                         // warpgroup 0..3: owns slab g of the output
    const int r = ((warp & 3) << 5) | lane;        // tile row == TMEM lane
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    ProdTrack pt;
    const uint32_t pmask = prologue_mask(prog);
    uint32_t gstep = 0;
    bool enc_done = false;      // the encoded positions of the current tile were produced during the previous tile
    // Activation slabs 0..3 live in TENSOR MEMORY for the next layer's MMAs; their shared-memory images exist only
    // as the staging buffer of the stash copies (training), so without a stash they are not written at all.
    auto begin_produce = [&](int slab) {
      if (!kDirectStash && keep && ((pt.any >> slab) & 1)) mbar_wait(&sm->s_free[slab], ((pt.par >> slab) & 1) ^ 1);
    };
    auto end_produce = [&](int slab) {   // slab 4 (shared memory only): one arrival per warp
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm->a_ready[slab]);
    };
    // Activation slab: the MMA warp is released as soon as the tensor-memory write has landed; the shared-memory
    // staging image for the stash copy (wait for the previous copy, 4 x st.shared, proxy fence) comes AFTER that
    // arrival and signals the stash lane on its own barrier -- it is off the epilogue -> MMA critical path.
    auto hand_off = [&](int slab) {
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm->a_ready[slab]);
    };
    auto stage_out = [&](int slab, int cb, const uint32_t (&pk)[16], uint8_t* gdirect) {
      if (keep) {
        if (kDirectStash) {
          if (gdirect != nullptr) store_global32(gdirect, r, cb, pk);
        } else {
          begin_produce(slab);
          store_packed32(slabs, r, cb, pk);
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm->s_ready[slab]);
        }
      }
    };
    auto publish = [&](int slab, int cb, const uint32_t (&pk)[16], uint8_t* gdirect) {
      hand_off(slab);
      stage_out(slab, cb, pk, gdirect);
    };
    // columns [16g, 16g+16) of one encoded row (which = 0 position / 1 direction) of point p, packed to bf16 pairs;
    // fused (rays, z) or pre-encoded (x) input.  All four warpgroups take part: a quarter of the row each.
    auto encoded_quarter = [&](long long p, bool valid, int which, uint32_t (&pk)[8]) {
      float e[16];
      if (args.x == nullptr) {
        float vx = 0.f, vy = 0.f, vz = 0.f;
        if (valid) {
          const long long ray = (unsigned)p / (unsigned)args.S;      // P < 2^31 (checked by the host side)
          const float* rp = args.rays + (size_t)ray * args.ray_stride;
          if (which == 0) {
            const float zz = args.z[p];
            vx = rp[0] + rp[3] * zz, vy = rp[1] + rp[4] * zz, vz = rp[2] + rp[5] * zz;
          } else {
            vx = rp[args.vd_col], vy = rp[args.vd_col + 1], vz = rp[args.vd_col + 2];
          }
        }
        encode_quarter(vx, vy, vz, which == 0 ? prog.L_pts : prog.L_dir, g, e);
        if (!valid) {
#pragma unroll
          for (int i = 0; i < 16; ++i) e[i] = 0.f;
        }
      } else {
        const int n_pts = 3 + 6 * prog.L_pts, n_dir = 3 + 6 * prog.L_dir;
        const int n = which == 0 ? n_pts : n_dir;
        const float* xp = args.x + (size_t)(valid ? p : 0) * args.x_ld + (which == 0 ? 0 : n_pts);
#pragma unroll
        for (int i = 0; i < 16; ++i) e[i] = (valid && 16 * g + i < n) ? xp[16 * g + i] : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) pk[i] = pack_bf16(e[2 * i], e[2 * i + 1]);
    };
    auto store_enc = [&](const uint32_t (&pk)[8], uint8_t* gslab) {   // slab 4 <- this warpgroup's quarter (+ direct stash)
      begin_produce(4);
      store_quarter(slabs + 4 * kSlab, r, g, pk);
      if (gslab != nullptr) store_global16(gslab, r, 16 * g, pk);
      end_produce(4);
    };

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long p = tile * DLN_TILE_ROWS + r;
      const bool valid = p < args.P;
      float dsig = 0.f;
      const float* const semrow =
          (kBwd && kSem && valid) ? args.sem_g + (size_t)((unsigned)p / (unsigned)args.sem_g_div) * 256 : nullptr;   // padded rows: dZ = 0
      if (kBwd && kSem && semrow != nullptr) {
        // the row is read once per tile, one step into the chain: start the two 128-byte lines this thread will need
        // on their way now so the epilogue of that step does not wait for L2
        asm volatile("prefetch.global.L1 [%0];" ::"l"(semrow + 32 * g));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(semrow + 128 + 32 * g));
      }
      uint8_t* const gtile = keep ? reinterpret_cast<uint8_t*>(args.stash) + (size_t)tile * prog.stash_slots * kSlab : nullptr;
      auto gslot = [&](int slot) -> uint8_t* { return (kDirectStash && keep) ? gtile + (size_t)slot * kSlab : nullptr; };
      // ------------------------------------------------------------------ prologue
      if (!kBwd) {
        if (!enc_done) {                // first tile of this CTA (later tiles: produced early, see the last step)
          uint32_t epk[8];
          encoded_quarter(p, valid, 0, epk);
          store_enc(epk, gslot(0));
        }
        enc_done = false;
      } else {
        // d raw row -> dZ of the first backward layer (through rgb_linear / output_linear) + d_raw slab
        float dr[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) dr[j] = (valid && j < prog.out_ch) ? args.d_out[(size_t)p * prog.out_ch + j] : 0.f;
        dsig = dr[3];
        const int nh = prog.use_viewdirs ? 3 : prog.out_ch;
        const int width = prog.use_viewdirs ? 128 : 256;
        const float* ph = args.fblob + prog.pro_head_off;
        {
          const uint2 mw = reinterpret_cast<const uint2*>(args.masks)[(((size_t)prog.pro_mask_slot * n_tiles + tile) * 4 + g) * 128 + r];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (128 * h < width) {
              const int cb = 128 * h + 32 * g;
              const uint32_t mk = h ? mw.y : mw.x;
              float f[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) f[i] = 0.f;
#pragma unroll
              for (int j = 0; j < 5; ++j)
                if (j < nh) {
#pragma unroll
                  for (int q = 0; q < 8; ++q) {
                    const float4 w4 = __ldg(reinterpret_cast<const float4*>(ph + j * width + cb + 4 * q));
                    f[4 * q] += dr[j] * w4.x, f[4 * q + 1] += dr[j] * w4.y, f[4 * q + 2] += dr[j] * w4.z, f[4 * q + 3] += dr[j] * w4.w;
                  }
                }
#pragma unroll
              for (int i = 0; i < 32; ++i) f[i] = (mk & (1u << i)) ? f[i] : 0.f;
              uint32_t pk[16];
              pack32<false>(f, pk);
              // dZ of the first backward layer -> tensor memory, in the buffer the previous step (last step of the
              // previous tile) accumulated in, over columns only this thread reads
              tmem_st16(tmem_base + ((gstep + 1) & 1) * 256 + lane_addr + cb, pk);
              publish(cb >> 6, cb, pk, keep ? gtile + (size_t)prog.pro_slot * kSlab : nullptr);
            }
          }
        }
        if (g == (prog.use_viewdirs ? 2 : 0)) {
          float e[64];
#pragma unroll
          for (int i = 0; i < 64; ++i) e[i] = 0.f;
#pragma unroll
          for (int j = 0; j < 5; ++j) e[j] = dr[j];
          begin_produce(4);
          store_row64(slabs + 4 * kSlab, gslot(0), r, e);
          end_produce(4);
        }
      }
      pt.par ^= pmask, pt.any |= pmask;

      // ------------------------------------------------------------------ layer epilogues
      int bias_base = 0;
      for (int s = 0; s < prog.n_steps; ++s, ++gstep) {
        const DlnChainStep& st = prog.steps[s];
        const int epi = st.epi;
        const float* bias = sm->bias + bias_base;
        bias_base += st.n_out;
        // The tile prologue's inputs are L2 hits at best (~700 cycles per dependent load, on the serial path of the
        // epilogue threads at a tile boundary): two steps ahead of their use the lines are started towards L1.
        if (s == (prog.n_steps > 2 ? prog.n_steps - 3 : 0) && tile + gridDim.x < n_tiles) {
          const long long pn = (tile + gridDim.x) * DLN_TILE_ROWS + r;
          if (pn < args.P) {
            if (!kBwd) {
              if (args.x == nullptr) {
                asm volatile("prefetch.global.L1 [%0];" ::"l"(args.z + pn));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(args.rays + (size_t)((unsigned)pn / (unsigned)args.S) * args.ray_stride));
              }
            } else {
              asm volatile("prefetch.global.L1 [%0];" ::"l"(args.d_out + (size_t)pn * prog.out_ch));
            }
          }
        }
        if (!kBwd && args.x == nullptr && s + 2 == reload_step && valid)      // this tile's view direction (read at reload_step)
          asm volatile("prefetch.global.L1 [%0];" ::"l"(args.rays + (size_t)((unsigned)p / (unsigned)args.S) * args.ray_stride + args.vd_col));
        uint2 mw = make_uint2(0, 0);
        const size_t mask_idx = (((size_t)(st.mask_slot < 0 ? 0 : st.mask_slot) * n_tiles + tile) * 4 + g) * 128 + r;
        if (epi >= DLN_EPI_BWD_MASK && st.mask_slot >= 0) mw = reinterpret_cast<const uint2*>(args.masks)[mask_idx];
        const int trole = (et == 0) ? 1 : -1;
        // Early prologue (forward): while this tile's LAST layer is still in the tensor pipe, warpgroup 0 encodes
        // the next tile's positions into registers; slab 4 (encoded direction, read by this layer) is free as soon
        // as acc_full fires, the row is stored and the next tile's layer 0 runs under this layer's epilogue.
        uint32_t epk[8];
        const long long next_tile = tile + gridDim.x;
        const bool early = kEarlyPrologue && !kBwd && s == prog.n_steps - 1 && next_tile < n_tiles && prog.n_steps > 1;
        if (early) {
          const long long pn = next_tile * DLN_TILE_ROWS + r;
          encoded_quarter(pn, pn < args.P, 0, epk);
        }
        mbar_wait(&sm->acc_full[gstep & 1], (gstep >> 1) & 1);
        tc_fence_after();
        if (trole > 0) trace_ev(sm, args.trace, trole, gstep, 3);
        if (early) {
          store_enc(epk, (kDirectStash && keep) ? reinterpret_cast<uint8_t*>(args.stash) + (size_t)next_tile * prog.stash_slots * kSlab : nullptr);
          enc_done = true;
        }

        const uint32_t t_acc = tmem_base + (gstep & 1) * 256 + lane_addr;
        const bool relu = (epi == DLN_EPI_RELU || epi == DLN_EPI_RELU_SIGMA || epi == DLN_EPI_RELU_RGB || epi == DLN_EPI_RELU_OUT);
        const int nheads = (epi <= DLN_EPI_RELU_OUT) ? st.n_heads : 0;
        const float* hw = args.fblob + st.head_off;
        float hacc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        uint32_t mo0 = 0, mo1 = 0;
        {
          auto run = [&](auto tag, const uint32_t(&v)[32], uint32_t mi, uint32_t& mo, int cb, uint32_t (&pk)[16]) {
            constexpr int E = decltype(tag)::value;
            epi_chunk<E>(v, bias, hw, st.n_out, nheads, dsig, mi, mo, hacc, pk, cb, semrow);
          };
          auto dispatch = [&](const uint32_t(&v)[32], uint32_t mi, uint32_t& mo, int cb, uint32_t (&pk)[16]) {
            if (!kBwd) {
              if (epi == DLN_EPI_LINEAR) run(std::integral_constant<int, DLN_EPI_LINEAR>{}, v, mi, mo, cb, pk);
              else run(std::integral_constant<int, DLN_EPI_RELU_OUT>{}, v, mi, mo, cb, pk);     // relu (+ heads when nheads > 0)
            } else {
              if (epi == DLN_EPI_BWD_COPY) run(std::integral_constant<int, DLN_EPI_BWD_COPY>{}, v, mi, mo, cb, pk);
              else if (epi == DLN_EPI_BWD_MASK) run(std::integral_constant<int, DLN_EPI_BWD_MASK>{}, v, mi, mo, cb, pk);
              else run(std::integral_constant<int, DLN_EPI_BWD_MASK_SIGMA>{}, v, mi, mo, cb, pk);
            }
          };
          uint32_t v[32], pk[16];
          const bool wide = st.n_out == 256;
          const int cb0 = 32 * g, cb1 = 128 + 32 * g;     // chunk h: features [128h + 32g, +32)
          tmem_ld32(t_acc + cb0, v);
          tmem_ld_wait();
          // The bf16 outputs go back IN PLACE into tensor memory for the next layer's MMAs: features [cb, cb+32) are
          // packed into columns [cb, cb+16) of this thread's lane -- columns nobody else reads, so no cross-warp
          // ordering is needed.
          dispatch(v, mw.x, mo0, cb0, pk);
          uint8_t* const gdirect = (keep && st.stash_slot >= 0) ? gtile + (size_t)st.stash_slot * kSlab : nullptr;
          tmem_st16(t_acc + cb0, pk);
          if (wide) {
            // Both hand-offs to the MMA warp come before any stash staging, and chunk 1's accumulator columns are on
            // their way to registers while chunk 0's store drains: the next layer's K slabs 2, 3 are released ~1000
            // cycles earlier than with load -> convert -> store -> stage per chunk.  (One branch diamond on purpose:
            // with the wide-only steps as separate `if`s ptxas spilled 350 bytes per thread in the forward kernel.)
            tmem_ld32(t_acc + cb1, v);
            hand_off(cb0 >> 6);
            if (trole > 0) trace_ev(sm, args.trace, trole, gstep, 4);
            tmem_ld_wait();
            tmem_ld_pin32(v);
            dispatch(v, mw.y, mo1, cb1, pk);
            tmem_st16(t_acc + cb1, pk);
            hand_off(cb1 >> 6);
            if (trole > 0) trace_ev(sm, args.trace, trole, gstep, 5);
            // staging for the stash copies, off the MMA critical path: chunk 0 read back from its tensor-memory
            // columns (the MMAs only read them), chunk 1 from the registers it was packed in
            if (keep) {
              uint32_t(&p0)[16] = *reinterpret_cast<uint32_t(*)[16]>(v);
              tmem_ld16(t_acc + cb0, p0);
              tmem_ld_wait();
              tmem_ld_pin16(p0);
              stage_out(cb0 >> 6, cb0, p0, gdirect);     // slabs 0, 1 first: the stash lane copies in slab order
              stage_out(cb1 >> 6, cb1, pk, gdirect);
            }
          } else {
            publish(cb0 >> 6, cb0, pk, gdirect);
          }
          if (trole > 0) trace_ev(sm, args.trace, trole, gstep, 6);
          if (relu && st.mask_slot >= 0 && args.masks != nullptr)
            reinterpret_cast<uint2*>(args.masks)[mask_idx] = make_uint2(mo0, mo1);
        }
        pt.par ^= step_out_mask(st), pt.any |= step_out_mask(st);

        if (s == reload_step) {
          // every MMA that reads the encoded position has completed (acc_full of this step): slab 4 is
          // overwritten with the encoded view direction for the views layer
          {
            uint32_t dpk[8];
            encoded_quarter(p, valid, 1, dpk);
            store_enc(dpk, gslot(1));
          }
          pt.par ^= 0x10u;
        }

        if (epi == DLN_EPI_RELU_SIGMA) {
          sm->part[g][3][r] = hacc[0] + (g == 0 ? args.fblob[st.head_bias_off] : 0.f);
        } else if (epi == DLN_EPI_RELU_RGB) {
#pragma unroll
          for (int h = 0; h < 3; ++h) sm->part[g][h][r] = hacc[h];       // idle warpgroups contribute 0
          named_bar_epi();
          if (g == 0 && valid) {
            float o[4];
#pragma unroll
            for (int h = 0; h < 3; ++h)
              o[h] = (sm->part[0][h][r] + sm->part[1][h][r]) + (sm->part[2][h][r] + sm->part[3][h][r]) +
                     args.fblob[st.head_bias_off + h];
            o[3] = (sm->part[0][3][r] + sm->part[1][3][r]) + (sm->part[2][3][r] + sm->part[3][3][r]);
            *reinterpret_cast<float4*>(args.out + (size_t)p * 4) = make_float4(o[0], o[1], o[2], o[3]);
          }
          named_bar_epi();      // part[] may be rewritten by the next tile only after everybody has read it
        } else if (epi == DLN_EPI_RELU_OUT) {
          // output_linear: up to 5 heads through the 4-deep partial-sum buffer, in two rounds
          for (int h0 = 0; h0 < nheads; h0 += 4) {
#pragma unroll
            for (int h = 0; h < 4; ++h)
              if (h0 + h < nheads) sm->part[g][h][r] = (h0 == 0) ? hacc[h] : hacc[4];
            named_bar_epi();
            if (g == 0 && valid) {
              for (int h = 0; h < 4 && h0 + h < nheads; ++h)
                args.out[(size_t)p * prog.out_ch + h0 + h] =
                    (sm->part[0][h][r] + sm->part[1][h][r]) + (sm->part[2][h][r] + sm->part[3][h][r]) +
                    args.fblob[st.head_bias_off + h0 + h];
            }
            named_bar_epi();
          }
        }
        if (et == 0) trace_ev(sm, args.trace, 1, gstep, 7);
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
  if (args.trace != nullptr && blockIdx.x == 0 && threadIdx.x < 128) args.trace[threadIdx.x] = sm->tr[threadIdx.x];
}

// ---------------------------------------------------------------------------------------------
// wgrad kernel
// ---------------------------------------------------------------------------------------------
constexpr int kWgThreads = 256;       // warp 0 producer, warp 1 MMA, warps 4..7 bias sums + epilogue
constexpr int kWgStages = 3;
constexpr int kHalfSlab = kSlab / 2;  // 64 points x 64 features
constexpr int kWgStageBytes = 8 * kHalfSlab;  // 4 A + 4 B half-slabs = 64 KB
constexpr int kWgPartialFloats = 256 * 256 + 256;   // per-CTA block of the deterministic reduction: dW accumulator + bias sums

struct WgradSmall {
  uint64_t full[kWgStages], empty[kWgStages];
  uint64_t acc_full;
  uint32_t tmem_base;
};
constexpr size_t kWgradSmemBytes = (size_t)kWgStages * kWgStageBytes + sizeof(WgradSmall) + 1024;

__global__ void __launch_bounds__(kWgThreads, 1)
    wgrad_kernel(const DlnWgradItem* __restrict__ items, int splits, const uint8_t* __restrict__ stash_fwd,
                 int fwd_slots, const uint8_t* __restrict__ stash_bwd, int bwd_slots, long long n_tiles,
                 float* __restrict__ grads, float* __restrict__ partial) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  WgradSmall* sm = reinterpret_cast<WgradSmall*>(smem + kWgStages * kWgStageBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // item-major grid with EQUAL split counts: CTA k of every item covers the same tile range (see the host side)
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
      const uint64_t desc_mn = umma_desc_sw128(0, kHalfSlab, 1024);   // MN-major: LBO = slab pitch, SBO = 8-row group
      const uint32_t smem0 = smem_u32(smem);
      for (long long q = 0; q < n_stages; ++q) {
        mbar_wait(&sm->full[stage], phase);
        tc_fence_after();
        const uint32_t abuf = smem0 + stage * kWgStageBytes;
        const uint64_t ad = desc_mn | (uint64_t)(abuf >> 4);
        const uint64_t bd = desc_mn | (uint64_t)((abuf + 4 * kHalfSlab) >> 4);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {          // 16 points (2048 B = 128 descriptor units) per MMA
          umma_bf16(tmem_base, ad + 128 * kk, bd + 128 * kk, idesc, (q | kk) != 0);
          if (nh == 2) umma_bf16(tmem_base + 256, ad + 128 * kk + (2 * kHalfSlab >> 4), bd + 128 * kk, idesc, (q | kk) != 0);
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
    // Deterministic mode (partial != null): this CTA's share of dW / db goes to its own block of `partial` with plain
    // stores -- [256 accumulator rows][256 columns] then [256 bias sums] -- and wgrad_reduce_kernel adds the blocks of
    // an item in split order.  Otherwise fp32 atomics straight into the gradient buffer (run-to-run order varies).
    float* const pblk = partial ? partial + (size_t)blockIdx.x * kWgPartialFloats : nullptr;
    if (pblk) {
      *reinterpret_cast<float2*>(pblk + 65536 + feat) = make_float2(do_bias ? b0 : 0.f, do_bias ? b1 : 0.f);
    } else if (do_bias) {
      float* db = grads + it.db_off;
      if (feat >= it.db_col_off && feat < it.db_col_off + it.db_n) atomicAdd(db + feat - it.db_col_off, b0);
      if (feat + 1 >= it.db_col_off && feat + 1 < it.db_col_off + it.db_n) atomicAdd(db + feat + 1 - it.db_col_off, b1);
    }
    // ---- dW: TMEM -> global
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
        if (pblk) {
          float4* dst = reinterpret_cast<float4*>(pblk + (size_t)(h * 128 + r) * 256 + c * 32);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            dst[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                                 __uint_as_float(v[4 * q + 3]));
        } else if (row_ok) {
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

// Second pass of the deterministic reduction: grid (16-row groups, items); thread = column.  The blocks of the splits
// of an item are added in split order (fixed summation order -> bit-reproducible gradients), skipping the splits that
// had no tiles, and the sum is ADDED to the gradient buffer (it accumulates across ray chunks; items never overlap).
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const DlnWgradItem* __restrict__ items, int splits, long long n_tiles,
                                                            const float* __restrict__ partial, float* __restrict__ grads) {
  const DlnWgradItem it = items[blockIdx.y];
  const long long per = (n_tiles + splits - 1) / splits;
  const int live = (int)((n_tiles + per - 1) / per);                 // splits [0, live) had work
  const float* blk = partial + (size_t)blockIdx.y * splits * kWgPartialFloats;
  const int col = threadIdx.x;
  for (int i = 0; i < 16; ++i) {
    const int row = blockIdx.x * 16 + i;
    if (row >= it.n_rows || col >= it.n_cols) continue;
    float acc = 0.f;
    for (int s = 0; s < live; ++s) acc += blk[(size_t)s * kWgPartialFloats + (size_t)(it.row_off + row) * 256 + col];
    grads[it.dw_off + (size_t)row * it.ld + it.col_off + col] += acc;
  }
  if (blockIdx.x == 0 && it.db_off >= 0 && col < it.db_n) {
    float acc = 0.f;
    for (int s = 0; s < live; ++s) acc += blk[(size_t)s * kWgPartialFloats + 65536 + it.db_col_off + col];
    grads[it.db_off + col] += acc;
  }
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
int dln_chain2_launch(const DlnChainProgram* prog, const DlnChainArgs* args, int num_sms, long long n_tiles, cudaStream_t stream);

// DLN_CHAIN=1 selects the one-tile-per-CTA kernel of this file, anything else the CTA-pair kernel (mlp_chain2.cu)
static int chain_variant() {
  static int v = 0;
  if (!v) {
    const char* e = getenv("DLN_CHAIN");
    v = (e && e[0] == '1') ? 1 : 2;
  }
  return v;
}

extern "C" {

int dln_mlp_chain(const DlnChainProgram* prog, const DlnChainArgs* args, int num_sms, void* stream) {
  DLN_CHECK_ARG(prog && args && num_sms > 0);
  DLN_CHECK_ARG(prog->n_steps >= 1 && prog->n_steps <= DLN_MAX_STEPS);
  DLN_CHECK_ARG(prog->out_ch >= 1 && prog->out_ch <= 5);
  DLN_CHECK_ARG(prog->L_pts >= 0 && prog->L_pts <= 10 && prog->L_dir >= 0 && prog->L_dir <= 10);
  DLN_CHECK_ARG(args->P >= 0 && args->P < (1ll << 31));
  if (args->P == 0) return DLN_OK;
  DLN_CHECK_ARG(args->wblob && args->fblob);
  int bias_floats = 0;
  for (int s = 0; s < prog->n_steps; ++s) {
    const DlnChainStep& st = prog->steps[s];
    DLN_CHECK_ARG(st.n_out == 256 || st.n_out == 128);
    DLN_CHECK_ARG(st.nk >= 1 && st.nk <= DLN_MAX_KSLABS && st.n_heads <= 5);
    for (int j = 0; j < st.nk; ++j) DLN_CHECK_ARG(st.kslab[j] < kNumSlabs && st.kcnt[j] >= 1 && st.kcnt[j] <= 4);
    DLN_CHECK_ARG((st.w_off & 1023u) == 0);
    bias_floats += st.n_out;
  }
  DLN_CHECK_ARG(bias_floats <= kMaxBiasFloats);
  DLN_CHECK_ARG(prog->reload_step < prog->n_steps);
  if (prog->backward) {
    DLN_CHECK_ARG(args->d_out && args->masks);
    DLN_CHECK_ARG(!args->sem_g || (args->sem_g_div >= 1 && (reinterpret_cast<uintptr_t>(args->sem_g) & 15) == 0));
  } else {
    DLN_CHECK_ARG(args->out);
    DLN_CHECK_ARG(args->x || (args->rays && (args->z || args->z_gen) && args->S >= 1 && args->ray_stride >= 6));
    DLN_CHECK_ARG(!args->z_gen || (!args->x && args->ray_stride >= 8 && chain_variant() == 2));   // CTA-pair kernel only
  }
  if (args->stash && prog->stash_slots > 0) DLN_CHECK_ARG((reinterpret_cast<uintptr_t>(args->stash) & 15) == 0);
  if (args->P == 0) return DLN_OK;
  const long long n_tiles = (args->P + DLN_TILE_ROWS - 1) / DLN_TILE_ROWS;
  bool narrow = prog->pro_valid != 0;
  for (int s = 0; s < prog->n_steps; ++s) {
    const int nv = prog->steps[s].n_valid32;
    DLN_CHECK_ARG(nv * 32 <= prog->steps[s].n_out);
    narrow = narrow || (nv != 0 && nv * 32 != prog->steps[s].n_out);
  }
  if (chain_variant() == 2 || narrow) return dln_chain2_launch(prog, args, num_sms, n_tiles, (cudaStream_t)stream);
  bool& attr_set = dln_device_flag(0);     // the attribute is per device (context), not per process
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(chain_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kChainSmemBytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(chain_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kChainSmemBytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(chain_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kChainSmemBytes);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const unsigned grid = (unsigned)(n_tiles < num_sms ? n_tiles : num_sms);
  if (prog->backward && args->sem_g)
    chain_kernel<true, true><<<grid, kThreads, kChainSmemBytes, (cudaStream_t)stream>>>(*prog, *args, n_tiles);
  else if (prog->backward)
    chain_kernel<true><<<grid, kThreads, kChainSmemBytes, (cudaStream_t)stream>>>(*prog, *args, n_tiles);
  else
    chain_kernel<false><<<grid, kThreads, kChainSmemBytes, (cudaStream_t)stream>>>(*prog, *args, n_tiles);
  return dln_launch_status();
}

int dln_mlp_wgrad(const DlnWgradItem* items_dev, int n_items, int splits, const void* stash_fwd, int fwd_slots,
                  const void* stash_bwd, int bwd_slots, long long n_tiles, float* grads_flat, float* partial,
                  void* stream) {
  DLN_CHECK_ARG(items_dev && n_items >= 1 && splits >= 1 && stash_fwd && stash_bwd && grads_flat && n_tiles >= 0);
  DLN_CHECK_ARG(!partial || (reinterpret_cast<uintptr_t>(partial) & 15) == 0);
  if (n_tiles == 0) return DLN_OK;
  bool& attr_set = dln_device_flag(1);
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWgradSmemBytes);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  wgrad_kernel<<<(unsigned)(n_items * splits), kWgThreads, kWgradSmemBytes, (cudaStream_t)stream>>>(
      items_dev, splits, reinterpret_cast<const uint8_t*>(stash_fwd), fwd_slots,
      reinterpret_cast<const uint8_t*>(stash_bwd), bwd_slots, n_tiles, grads_flat, partial);
  if (partial) {
    int st = dln_launch_status();
    if (st != DLN_OK) return st;
    wgrad_reduce_kernel<<<dim3(16, (unsigned)n_items), 256, 0, (cudaStream_t)stream>>>(items_dev, splits, n_tiles, partial,
                                                                                        grads_flat);
  }
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
  out[3] = (int)sizeof(DlnWgradItem), out[4] = (int)sizeof(DlnPackJob), out[5] = (int)sizeof(DlnSemOffsets);
  return DLN_OK;
}

const char* dln_build_info(void) { return "dlnerf_b200 sm_100a tcgen05/TMEM + bulk-copy (TMA) weights"; }

}  // extern "C"
