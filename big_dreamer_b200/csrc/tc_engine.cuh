// Persistent row-tile engine of the tensor-core path.
//
// One CTA owns a tile of 128 rows (start states / candidates) for ALL time steps, so the
// recurrent state never leaves the SM: activations live in shared memory as 16-bit KM8 operand
// tiles, accumulators in TMEM, and only the reference's output tensors go to HBM.  Weights
// (~1.6 MB packed, L2-resident) are streamed through a ring of shared-memory stages by bulk TMA.
//
// Warp roles (320 threads):  warp 0 = weight producer (bulk TMA), warp 1 = MMA issuer
// (tcgen05.mma, one lane) and TMEM owner, warps 2..9 = epilogue (tcgen05.ld -> bias is already
// in the accumulator via the constant-1 column -> activation / gate / sampling math -> next
// layer's operand tile + global outputs).
//
// The work of one time step is a host-built PROGRAM: a list of GEMMs grouped into PHASES, each
// phase ending in one epilogue.  All three roles walk the same program, which makes the
// producer / issuer / epilogue hand-offs structurally consistent:
//     w_full[s] / w_empty[s]   producer <-> issuer, one ring stage = one K-block of the weight image
//     acc_full[Gm & 3]         issuer -> epilogue, tcgen05.commit after the last GEMM of phase Gm
//     epi_done[Ge & 7]         epilogue -> issuer, 256 arrivals when (sub-)epilogue Ge is done;
//                              every GEMM names the completion it depends on (dep_back)
// A phase names how far back its dependency is (dep_back = 1: previous phase; 2: the one before,
// which lets independent phases -- the GRU's N-slices -- overlap MMA with the previous epilogue
// using the two TMEM halves).
//
// Column-split mode (nranks = 2 or 4, small row counts): the CTAs of a thread-block cluster share
// ONE row tile.  Every rank runs its own program over a slice of each layer's output columns
// (its own weight rows, so the weight stream and the epilogue work are divided by nranks) and
// writes its slice of the next operand tile into every rank's shared memory.  The barriers then
// become cluster-wide: tcgen05.commit multicasts to acc_full of every rank (an epilogue may only
// overwrite operand columns once NO rank's MMAs still read them) and every epilogue warp arrives
// on epi_done of every rank.  Narrow phases (actor / prior / head outputs) are replicated.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_pack.cuh"

namespace bd {
namespace tc {

constexpr int kTileRows = 128;
constexpr int kThreads = 320;
constexpr int kEpiThreads = 256;
// rollout kernel (round 2): kParts2 column parts of 4 epilogue warps (one per TMEM quadrant).  4 parts = 18 warps:
// five on two of the SM's sub-partitions, hence 96 registers per thread; 3 parts = 14 warps at 128 registers.
#ifndef BD_ROLLOUT_PARTS
#define BD_ROLLOUT_PARTS 3
#endif
constexpr int kParts2 = BD_ROLLOUT_PARTS;
constexpr int kEpiThreads2 = 128 * kParts2;
constexpr int kThreads2 = 64 + kEpiThreads2;
constexpr int kMaxGemms = 64, kMaxPhases = 40;
constexpr int kMaxRanks = 4;            // CTAs per cluster in column-split mode (1 = off)
constexpr uint32_t kLboA = kTileRows * 16;   // bytes between 8-column groups of an activation tile

enum TileId : uint8_t { TILE_BCUR = 0, TILE_BNXT = 1, TILE_SA = 2, TILE_H = 3, TILE_H2 = 4,
                       TILE_SLAB0 = 5, TILE_SLAB1 = 6, TILE_D2 = 7 };
enum EpiKind : uint8_t {
  EPI_ACT_H = 1,      // act(D) -> H tile
  EPI_ACTOR_OUT = 2,  // action sample + entropy
  EPI_GRU = 3,        // GRU gates for one N-slice -> b'
  EPI_PRIOR_OUT = 4,  // mean / std / sampled state
  EPI_HEAD_OUT = 5,   // scalar head output (reward / value) of the fused rollout
  EPI_STORE_OUT = 6   // plain fp32 store of n_valid output columns (MLP forward)
};

struct Gemm {
  uint32_t w_off;     // element offset of the packed weight image (rows Np, cols Kp)
  uint16_t Np, Kp;    // MMA N (mult of 16, <= 256), K (mult of 16)
  uint16_t a_k0;      // first K column inside the A tile (mult of 8)
  uint16_t d_col;     // TMEM column of the accumulator
  uint8_t a_tile;     // TileId
  uint8_t accumulate; // 1: continue a running sum in D; 2: only from the second time step on
  uint16_t kc;        // K columns per weight-ring stage (mult of 16): Np*kc*2 <= stage bytes
  uint8_t dep_back;   // this GEMM may start once epilogue completion (Ge_at_phase_start - dep_back) is in
  uint8_t pad;
};
struct Phase {
  uint8_t g0, ng;     // GEMM range
  uint8_t epi;        // EpiKind
  uint8_t dep_back;   // 1 or 2 (see header comment)
  uint16_t n_valid;   // valid output columns
  uint16_t Np;        // accumulator columns (per gate for EPI_GRU)
  uint16_t Kp_out;    // columns of the operand tile this epilogue must (re)write
  uint16_t d_col;     // TMEM column of the accumulator(s)
  uint16_t aux0;      // EPI_GRU: first belief column of the slice; EPI_HEAD_OUT: head index
  uint8_t out_tile;   // TileId written by EPI_ACT_H
  uint8_t pad;
  uint8_t n_sub;      // sub-epilogues (1 or 2): an ACT epilogue may publish columns [0, split) early
  uint8_t pad2;
  uint16_t split;     // first column of the second sub-epilogue (multiple of 32)
  uint16_t col0;      // column-split mode: first (global) output column of this rank; Kp_out is its end
  uint16_t pad3;
};
struct Program {
  int n_gemms, n_phases;
  Gemm g[kMaxGemms];
  Phase p[kMaxPhases];
};

struct SmemPlan {
  uint32_t off_tile[8];   // indexed by TileId (byte offsets from the 1024-aligned base)
  uint32_t off_ring, stage_bytes, nstage;
  uint32_t total;
};

// Per-(tile, step) global inputs of the epilogues that the producer warp pulls into L2 one step
// ahead (prefetch.global.L2), so the epilogues see L2 latency instead of first-touch DRAM latency.
struct PrefetchPlan {
  int n;                     // number of ranges (0 = off)
  int reverse;               // 1: the kernel walks time backwards (step index i -> t = T-1-i)
  const char* base[6];
  long long step_stride[6];  // bytes between consecutive t
  long long tile_stride[6];  // bytes between consecutive tiles
  unsigned int bytes[6];     // bytes per (tile, step)
};
__device__ __forceinline__ void prefetch_step(const PrefetchPlan& pf, long long tile, int t) {
  const int lane = threadIdx.x & 31;
  for (int r = 0; r < pf.n; ++r) {
    const char* p = pf.base[r] + t * pf.step_stride[r] + tile * pf.tile_stride[r];
    for (unsigned int o = lane * 128u; o < pf.bytes[r]; o += 32u * 128u)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p + o));
  }
}

// debug counters: fire-and-forget reductions (a load-add-store would stall the in-order issuer on the load)
__device__ __forceinline__ void prof_add(long long* p, long long v) {
  atomicAdd(reinterpret_cast<unsigned long long*>(p), static_cast<unsigned long long>(v));
}

// The actor's layer-0 input of a rollout step IS the belief tile [b_t | 1] and the first columns of the [s_t ; a]
// tile as they sit in shared memory when the step starts -- exactly the wgrad operand images (x0b / x0s) the batched
// actor backward would otherwise rebuild from the fp32 outputs (a serial, latency-bound tile prologue: ~17 K cycles
// of its 57 K per tile).  The issuer warp copies them out per (step, tile) with two bulk shared -> global copies.
struct X0Save {
  uint16_t* x0b;             // [(t * ntiles + tile)][128 x kp_b]; null = off
  uint16_t* x0s;             // [(t * ntiles + tile)][128 x kp_s]
  uint32_t bytes_b, bytes_s; // bytes per tile image
};

struct RolloutArgs {
  Program prog[kMaxRanks];   // one program per cluster rank (column-split mode); [0] when nranks == 1
  int nranks;
  int ws;                    // weight-share cluster size (1, 2 or 4; nranks == 1 only), set by the launcher
  SmemPlan sm;
  const uint16_t* wpack;
  long long N;            // rows
  int T;
  int Be, S, A, Hi, J;
  int Kp_b, Kp_sa, Kp_h;  // operand tile widths
  int act;
  float min_std;
  bd_actor_cfg cfg;
  const float *prev_state, *prev_belief, *eps_a, *eps_e, *eps_s;
  float *beliefs, *states, *means, *stds, *entropy, *actions, *actor_raw, *dent;
  float *head_out[2];     // optional fused heads: (T,N) reward / value
  const float* ext_actions;   // CEM / TransitionModel.forward: actions given, no actor
  long long* prof;            // optional (debug): per-phase cycle counters of CTA 0, see tc_imagine.cu
  int dbg;                    // debug experiments of the PROF build (BD_TC_DBG): 1 = epilogues skip their work, 2 = no weight copies
  float* mlp_out;             // EPI_STORE_OUT target (rows, n_valid)
  // saved for the tensor-core BPTT (16-bit tile images, see tc_bptt.cuh); null = do not save
  uint16_t *sv_gate, *sv_xa, *sv_ha;
  int kb_sv, kh_sv;
  uint16_t* sv_mlp[BD_MAX_LAYERS];   // MLP forward / the rollout's actor: image of hidden layer l
  int kact_sv;                       // its columns (0: the phase's own Kp_out -- not in column-split mode, where Kp_out is the rank's slice end)
  // fused Dreamer rollout (imagine_and_returns): act'(h) images of the two heads' hidden layers per
  // (t, tile) for the fused backward (index head * BD_MAX_LAYERS + layer; null = do not save), and the
  // lambda-return tail (src/dreamer.py:447-471 with bootstrap = value[-1], :329-335) over head_out[0/1]
  uint16_t* sv_hd[2 * BD_MAX_LAYERS];
  int kh_hd;                         // columns of those images
  float* returns;                    // (T,N); null = no tail
  float lr_disc, lr_lam, lr_oml;     // discount, lambda, 1 - lambda as fp32 (rounded like the standalone kernel)
  int has_b1;                 // 0: single belief tile (MLP forward), 1: ping-pong
  // CEM: rows are (batch row, local candidate); start latents are per batch row and the state
  // noise is indexed by the GLOBAL candidate (src/planner.py:37-39, 53-65)
  int cem_cl, cem_c, cem_c0;  // local candidates per batch row (0 = off), global candidates, first
  PrefetchPlan pf;            // next tile's / step's epilogue inputs (producer-warp L2 prefetch)
  X0Save x0;                  // the actor's layer-0 input images for the batched actor backward (see X0Save)
};

// fast-math activations for the 16-bit path (the result is rounded to 10 / 7 mantissa bits anyway)
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float tc_act(int act, float x) {
  switch (act) {
    case BD_ACT_ELU: return x > 0.f ? x : __expf(x) - 1.f;
    case BD_ACT_RELU: return fmaxf(x, 0.f);
    case BD_ACT_TANH: return fast_tanh(x);
    default: return x;
  }
}

template <int FMT>
__device__ __forceinline__ void store16(uint8_t* tile, int row, int col0, const float* v) {
  uint32_t pk[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) pk[j] = Half16<FMT>::pack2(v[2 * j], v[2 * j + 1]);
  uint8_t* p = tile + km8_offset(kTileRows, row, col0);
  *reinterpret_cast<uint4*>(p) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  *reinterpret_cast<uint4*>(p + kLboA) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
}
template <int FMT>
__device__ __forceinline__ void store1(uint8_t* tile, int row, int col, float v) {
  *reinterpret_cast<uint16_t*>(tile + km8_offset(kTileRows, row, col)) = Half16<FMT>::cvt(v);
}

// ---------------------------------------------------------------------------------------------
// Shared skeleton of every kernel on this path: barrier setup, the weight-producer warp and the
// MMA-issuer warp.  Kernels differ only in their tile initialisation and epilogues.
struct EngineShared {
  uint64_t w_full[8], w_empty[8], acc_full[4], epi_done[8];
  uint64_t w_peer[8];       // CTA-pair mode, leader only: the peer's half of ring stage i has landed (relay_role)
  uint32_t tmem_holder;
};

// The program tables travel as kernel parameters (constant bank), but every role indexes them
// dynamically once per GEMM / phase; an indexed constant load that misses the small constant cache
// costs an L2 round trip (measured: ~570 cycles per GEMM on the issuer's critical path).  Each CTA
// therefore keeps its own copy in shared memory and all roles read that.
__device__ __forceinline__ void stage_program(Program& dst, const Program& src) {
  static_assert(sizeof(Program) % 4 == 0, "Program must be a whole number of words");
  const uint32_t* s = reinterpret_cast<const uint32_t*>(&src);
  uint32_t* d = reinterpret_cast<uint32_t*>(&dst);
  for (uint32_t i = threadIdx.x; i < sizeof(Program) / 4; i += blockDim.x) d[i] = s[i];
}

// ws > 1 (weight-share cluster, see producer_role): a ring stage is refilled by all ws producers of the
// cluster, so it is free only when the issuers of all ws CTAs have consumed it.
//
// CTA-pair mode (pair2, clusters of 2; see issuer_role<.., PAIR2>): every CTA owns its OWN row tile, the leader (cluster
// rank 0) issues tcgen05.mma.cta_group::2 for both, and every weight-ring stage holds HALF of the weight rows in
// each CTA -- the weight bytes a CTA pulls from L2 per row tile are halved.  Barriers: w_full (local copy landed),
// w_peer (leader: the peer's half landed), w_empty / acc_full (the leader's commits, multicast to both CTAs),
// epi_done (leader: one arrival per epilogue warp of BOTH CTAs).
// (PAIR2 is a template parameter: a kernel that merely CONTAINS cta_group::2 instructions can no longer be launched
// without a cluster -- "cluster misconfiguration" -- so the pair variants are separate instantiations)
template <bool PAIR2 = false>
__device__ __forceinline__ uint32_t engine_setup(EngineShared& sh, uint32_t nstage, uint32_t R = 1,
                                                 uint32_t epi_threads = kEpiThreads, uint32_t ws = 1) {
  constexpr bool pair2 = PAIR2;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    for (uint32_t i = 0; i < nstage; ++i) {
      mbar_init(&sh.w_full[i], 1); mbar_init(&sh.w_empty[i], ws);
      if (pair2) mbar_init(&sh.w_peer[i], 1);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&sh.acc_full[i], pair2 ? 1 : R);
    // epi_threads == kEpiThreads2: the 16-warp kernels arrive once per warp; the 8-warp kernels once per thread
    for (int i = 0; i < 8; ++i)
      mbar_init(&sh.epi_done[i], pair2 ? (epi_threads / 32) * 2
                                       : ((R == 1 && epi_threads == kEpiThreads) ? epi_threads : (epi_threads / 32) * R));
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (PAIR2) tmem_alloc_pair<512>(&sh.tmem_holder);
    else tmem_alloc<512>(&sh.tmem_holder);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (R > 1 || ws > 1 || pair2) cluster_sync_all();   // peers' barriers are initialised before anyone arrives remotely
  return sh.tmem_holder;
}

// The whole warp walks the program converged with warp-uniform waits (mbar_wait_u), so every
// address stays in uniform registers; the copy and its expect_tx are single instructions predicated
// on the elected lane.
//
// Weight-share clusters (ws = 2 or 4 CTAs, each with its OWN row tile, all walking the same program in
// lockstep): every weight stage is read from L2 once per cluster -- producer k copies the k-th 1/ws of the
// stage with .multicast::cluster into the ring of every CTA of the cluster.  With one CTA per SM streaming
// all ~1 MB of weights per time step for a single 128-row tile, the L2 -> SM weight stream (148 x 1 MB per
// step) and the latency it queues up at the ring's depth were what the issuer waited for (measured: ~20 B/clk
// per SM = 4.1 K cycles per 208 x 208 layer against 1.35 K of MMA).
__device__ __forceinline__ void producer_role(const Program& P, const SmemPlan& sm,
                                              const uint16_t* wpack, long long ntiles, int T,
                                              uint8_t* smem, EngineShared& sh,
                                              const PrefetchPlan* pf = nullptr, uint32_t R = 1, uint32_t ws = 1,
                                              bool dbg_no_copy = false, int pair_rank = -1) {
  const uint32_t wrank = ws > 1 ? (uint32_t)blockIdx.x % ws : 0u;
  const uint16_t wmask = (uint16_t)((1u << ws) - 1u);
  const uint32_t ring = smem_u32(smem) + sm.off_ring;
  const uint32_t bar_full = smem_u32(&sh.w_full[0]), bar_empty = smem_u32(&sh.w_empty[0]);
  uint32_t nstage = sm.nstage, stage_bytes = sm.stage_bytes;
  asm volatile("" : "+r"(nstage), "+r"(stage_bytes));     // (kept in registers, see issuer_role)
  uint32_t st = 0, ph = 0;
  const long long tile0 = blockIdx.x / R, tstride = gridDim.x / R;
  if (pair_rank >= 0) {
    // CTA-pair mode (R = 2; ntiles = tile pairs): this CTA's half of every stage -- rows [Np/2 rank, +Np/2) of the
    // weight image, packed as two half images (PackJob::split2) so a K range of a half is one contiguous copy
    for (long long tile = tile0; tile < ntiles; tile += tstride)
      for (int t = 0; t < T; ++t) {
        if (pf && pf->n) {
          const long long mine = tile * 2 + pair_rank;       // (prefetch plans are per row tile)
          if (t == 0 && tile == tile0) prefetch_step(*pf, mine, pf->reverse ? T - 1 : 0);
          if (t + 1 < T) prefetch_step(*pf, mine, pf->reverse ? T - 2 - t : t + 1);
          else if (tile + tstride < ntiles) prefetch_step(*pf, mine + 2 * tstride, pf->reverse ? T - 1 : 0);
        }
        for (int gi = 0; gi < P.n_gemms; ++gi) {
          const Gemm g = P.g[gi];
          const uint32_t nh = g.Np / 2u;
          const uint16_t* src = wpack + g.w_off + (size_t)pair_rank * nh * g.Kp;
          for (int k0 = 0; k0 < g.Kp; k0 += g.kc) {
            const int kc = min((int)g.kc, g.Kp - k0);
            mbar_wait_u(bar_empty + st * 8, ph ^ 1);
            tma_bulk_g2s_elect(ring + st * stage_bytes, src + (size_t)k0 * nh, nh * kc * 2u, bar_full + st * 8);
            if (++st == nstage) { st = 0; ph ^= 1; }
          }
        }
      }
    return;
  }
  for (long long tile = tile0; tile < ntiles; tile += tstride)
    for (int t = 0; t < T; ++t) {
      if (pf && pf->n) {      // inputs of the NEXT step (and of step 0 when a tile starts)
        if (t == 0 && tile == tile0) prefetch_step(*pf, tile, pf->reverse ? T - 1 : 0);
        if (t + 1 < T) prefetch_step(*pf, tile, pf->reverse ? T - 2 - t : t + 1);
        else if (tile + tstride < ntiles) prefetch_step(*pf, tile + tstride, pf->reverse ? T - 1 : 0);
      }
      for (int gi = 0; gi < P.n_gemms; ++gi) {
        const Gemm g = P.g[gi];
        const uint16_t* src = wpack + g.w_off;
        for (int k0 = 0; k0 < g.Kp; k0 += g.kc) {
          const int kc = min((int)g.kc, g.Kp - k0);
          const uint32_t bytes = (uint32_t)g.Np * kc * 2;
          mbar_wait_u(bar_empty + st * 8, ph ^ 1);
          if (dbg_no_copy) {
            if (elect_one()) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_full + st * 8) : "memory");
            __syncwarp();
          } else if (ws == 1) {
            tma_bulk_g2s_elect(ring + st * stage_bytes, src + (size_t)k0 * g.Np, bytes, bar_full + st * 8);
          } else {      // (bytes is a multiple of 512: Np and kc are multiples of 16)
            const uint32_t slice = bytes / ws;
            tma_bulk_g2s_mc_elect(ring + st * stage_bytes + wrank * slice,
                                  reinterpret_cast<const char*>(src + (size_t)k0 * g.Np) + wrank * slice, slice,
                                  bar_full + st * 8, bytes, wmask);
          }
          if (++st == nstage) { st = 0; ph ^= 1; }
        }
      }
    }
}

// Ge counts epilogue completions (incl. the per-tile init pseudo-phase, which has no MMAs); Gm
// counts phases with MMAs.  Each indexes its own barrier ring so generations stay in step.
//
// The whole warp walks the program converged: every lane executes the barrier waits and the
// descriptor arithmetic on warp-uniform values, and each tcgen05.mma / tcgen05.commit is a single
// instruction predicated on the elected lane (umma_f16_elect).  Wrapping the MMAs of a ring stage
// in an `if (elected)` region instead costs ~300 cycles per region (operands are moved into
// uniform registers again) -- measured with scripts/mmabench2.py: 216 vs 79 cycles per MMA.
template <int FMT, bool PROF>
__device__ __forceinline__ void issuer_role(const Program& P, const SmemPlan& sm, long long ntiles,
                                            int T, uint8_t* smem, EngineShared& sh,
                                            uint32_t tmem_base, long long* prof, uint32_t R = 1,
                                            uint32_t ws = 1, bool dbg_no_ring = false,
                                            const X0Save* x0 = nullptr) {
  const int lane = threadIdx.x & 31;
  const bool x0_on = x0 != nullptr && x0->x0b != nullptr && (R == 1 || blockIdx.x % R == 0);
  const long long x0_tiles = (long long)ntiles;
  const uint16_t wmask = (uint16_t)((1u << ws) - 1u);
  uint32_t nstage = sm.nstage, stage_bytes = sm.stage_bytes;
  // (opaque to the compiler: otherwise both are re-read from the constant bank on every ring stage)
  asm volatile("" : "+r"(nstage), "+r"(stage_bytes));
  uint32_t st = 0, wph = 0, Ge = 0, Gm = 0;
  uint32_t waited = 0xFFFFFFFFu;   // highest epilogue-completion index already waited for (-1: none)
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t ring_addr = smem_base + sm.off_ring;
  const uint32_t bar_w_full = smem_u32(&sh.w_full[0]), bar_w_empty = smem_u32(&sh.w_empty[0]);
  // the ring slot of the next stage as running values (no multiply, no constant-bank reload per stage): its
  // descriptor address field, and the byte offset of its barrier pair
  const uint32_t ring16 = (ring_addr >> 4) & 0x3FFFu, stage16 = stage_bytes >> 4;
  uint32_t slot16 = ring16, bar_off = 0;
  const uint32_t bar_acc_full = smem_u32(&sh.acc_full[0]), bar_epi_done = smem_u32(&sh.epi_done[0]);
  // descriptor high word: SBO = 128 B, version 1, no swizzle; low word: start >> 4 | LBO >> 4 << 16
  const uint32_t desc_hi32 = (uint32_t)(make_smem_desc(0, 0, 128) >> 32);
  // per-stage clock reads cost ~70 cycles each on this warp's critical path: only on request
  // (BD_TC_PROF=2); the default debug build times whole phases and dependency waits only
  const bool fine = PROF && prof != nullptr && prof[39 * 8 + 7] != 0;
  for (long long tile = blockIdx.x / R; tile < ntiles; tile += gridDim.x / R) {
    ++Ge;  // the tile-initialisation pseudo-phase (epilogue only)
    for (int t = 0; t < T; ++t) {
      const uint32_t par = (uint32_t)t & 1u;
      for (int pi = 0; pi < P.n_phases; ++pi) {
        const Phase ph = P.p[pi];
        long long c1 = 0, wsum = 0, dsum = 0, msum = 0, csum = 0;
        if (PROF) c1 = clock64();
        for (int gi = ph.g0; gi < ph.g0 + ph.ng; ++gi) {
          const Gemm g = P.g[gi];
          {   // per-GEMM dependency (sub-epilogue pipelining); older completions are implied
            const uint32_t D = Ge - g.dep_back;
            if ((int)(D - waited) > 0) {
              long long d0 = 0;
              if (PROF) d0 = clock64();
              if (R > 1) {     // peers wrote their column slices of the operand tiles (generic proxy)
                mbar_wait_cluster_u(bar_epi_done + (D & 7) * 8, (D >> 3) & 1);
                fence_proxy_async_all();
              } else {
                mbar_wait_u(bar_epi_done + (D & 7) * 8, (D >> 3) & 1);
              }
              tc_fence_after_sync();
              waited = D;
              if (PROF) dsum += clock64() - d0;
            }
          }
          uint32_t tile_id = g.a_tile;
          if (tile_id < 2) tile_id ^= par;
          // (in a cluster launch the shared-window address carries the CTA rank in its upper bits:
          // the descriptor takes only the 18-bit offset)
          uint32_t a_lo = ((uint32_t)(kLboA >> 4) << 16) |
                          ((((smem_base + sm.off_tile[tile_id]) >> 4) & 0x3FFFu) + (uint32_t)(g.a_k0 >> 3) * (kLboA >> 4));
          const uint32_t idesc = make_idesc_f16(FMT, kTileRows, g.Np);
          const uint32_t lbo_b16 = (uint32_t)g.Np;                // (Np * 16 bytes) >> 4
          const uint32_t b_lbo = lbo_b16 << 16;
          const uint32_t d_tmem = tmem_base + g.d_col;
          uint32_t acc = g.accumulate == 2 ? (t > 0 ? 1u : 0u) : g.accumulate;
          for (int k0 = 0; k0 < g.Kp; k0 += g.kc) {
            const int kc = min((int)g.kc, g.Kp - k0);
            long long w0 = 0;
            if (PROF && fine) w0 = clock64();
            // (no tcgen05.fence here: the weights arrive through the async proxy and their mbarrier
            // completion orders them before the MMAs; a fence::after_thread_sync per ring stage measured
            // ~220 cycles of issue stall each -- it is only needed after the epilogue hand-offs above)
            if (!(PROF && dbg_no_ring)) mbar_wait_u(bar_w_full + bar_off, wph);
            if (PROF && fine) wsum += clock64() - w0;
            uint32_t b_lo = b_lbo | slot16;
            long long m0 = 0;
            if (PROF && fine) m0 = clock64();
#pragma unroll 1
            for (int ks = 0; ks < kc; ks += 16) {
              // one K=16 step = two 8-column groups: A advances 2*kLboA bytes, B 2*lbo_b bytes (both >> 4)
              umma_f16_u32(d_tmem, a_lo, b_lo, desc_hi32, idesc, acc);
              acc = 1;
              a_lo += 2 * (kLboA >> 4);
              b_lo += 2 * lbo_b16;
            }
            if (PROF && fine) msum += clock64() - m0;
            if (PROF && dbg_no_ring) {}
            else if (ws == 1) umma_commit_elect(bar_w_empty + bar_off);
            else umma_commit_mc_elect(bar_w_empty + bar_off, wmask);     // the stage is free in every CTA of the cluster
            if (PROF && fine) csum += clock64() - m0;
            slot16 += stage16; bar_off += 8;
            if (++st == nstage) { st = 0; wph ^= 1; slot16 = ring16; bar_off = 0; }
          }
        }
        if (x0_on && pi <= 1) {
          // phase 0's dependency wait saw the tiles of this step complete (and published to the async proxy by the
          // epilogues' fence); the copies read shared memory while the first phases' MMAs run, and their reads are
          // waited for one phase later -- long before anything rewrites either tile
          if (lane == 0) {
            if (pi == 0) {
              const size_t img = (size_t)t * x0_tiles + (size_t)tile;
              tma_bulk_s2g(reinterpret_cast<char*>(x0->x0b) + img * x0->bytes_b,
                           smem_base + sm.off_tile[TILE_BCUR ^ par], x0->bytes_b);
              tma_bulk_s2g(reinterpret_cast<char*>(x0->x0s) + img * x0->bytes_s,
                           smem_base + sm.off_tile[TILE_SA], x0->bytes_s);
              bulk_commit();
            } else {
              bulk_wait_read();
            }
          }
          __syncwarp();
        }
        if (R > 1) umma_commit_mc_elect(bar_acc_full + (Gm & 3) * 8, (uint16_t)((1u << R) - 1));
        else umma_commit_elect(bar_acc_full + (Gm & 3) * 8);
        if (PROF && blockIdx.x == 0 && lane == 0) {
          prof_add(&prof[pi * 8 + 0], dsum);                          // issuer: wait for the dependency epilogue(s)
          prof_add(&prof[pi * 8 + 1], wsum);                          // issuer: wait for weight stages
          prof_add(&prof[pi * 8 + 2], clock64() - c1 - wsum - dsum);  // issuer: issue time
          prof_add(&prof[pi * 8 + 7], msum);                          // of which: inside the tcgen05.mma loops
          prof_add(&prof[(20 + pi) * 8 + 0], csum - msum);            //           the per-stage tcgen05.commit
        }
        ++Gm;
        Ge += ph.n_sub;
      }
    }
  }
  if (x0_on && lane == 0) bulk_wait_all();     // the images are in global memory before the CTA retires
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// CTA-pair mode (clusters of 2, engine_setup(pair2 = true)).  Both CTAs walk the same program over their OWN row
// tiles (pair p: tiles 2p and 2p + 1); `npairs` is the loop bound.
//   leader (cluster rank 0), warp 1: issuer_pair_role -- the ONE issuer of the pair: M = 256 MMAs
//       (tcgen05.mma.cta_group::2) whose A operand is each CTA's own tile at the same shared-memory offset, whose B
//       operand is the two half stages, and whose accumulators land at the same TMEM address in both CTAs; commits
//       are multicast, so both producers see w_empty and both epilogues see acc_full.
//   peer (rank 1), warp 1: relay_role -- forwards "my half of stage s has landed" to the leader's w_peer[s].
// The epilogue warps of both CTAs arrive on the LEADER's epi_done (release at cluster scope after a proxy fence: the
// pair's MMAs read both CTAs' operand tiles through the async proxy).
template <int FMT>
__device__ __forceinline__ void issuer_pair_role(const Program& P, const SmemPlan& sm, long long npairs, int T,
                                                 uint8_t* smem, EngineShared& sh, uint32_t tmem_base) {
  uint32_t nstage = sm.nstage, stage_bytes = sm.stage_bytes;
  asm volatile("" : "+r"(nstage), "+r"(stage_bytes));
  uint32_t st = 0, wph = 0, Ge = 0, Gm = 0;
  uint32_t waited = 0xFFFFFFFFu;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t ring_addr = smem_base + sm.off_ring;
  const uint32_t bar_w_full = smem_u32(&sh.w_full[0]), bar_w_empty = smem_u32(&sh.w_empty[0]);
  const uint32_t bar_w_peer = smem_u32(&sh.w_peer[0]);
  const uint32_t bar_acc_full = smem_u32(&sh.acc_full[0]), bar_epi_done = smem_u32(&sh.epi_done[0]);
  const uint64_t desc_hi = make_smem_desc(0, 0, 128);
  for (long long pr = blockIdx.x / 2; pr < npairs; pr += gridDim.x / 2) {
    ++Ge;  // the tile-initialisation pseudo-phase
    for (int t = 0; t < T; ++t) {
      const uint32_t par = (uint32_t)t & 1u;
      for (int pi = 0; pi < P.n_phases; ++pi) {
        const Phase ph = P.p[pi];
        for (int gi = ph.g0; gi < ph.g0 + ph.ng; ++gi) {
          const Gemm g = P.g[gi];
          {
            const uint32_t D = Ge - g.dep_back;
            if ((int)(D - waited) > 0) {
              mbar_wait_cluster_u(bar_epi_done + (D & 7) * 8, (D >> 3) & 1);
              fence_proxy_async_all();
              tc_fence_after_sync();
              waited = D;
            }
          }
          uint32_t tile_id = g.a_tile;
          if (tile_id < 2) tile_id ^= par;
          uint64_t a_desc = desc_hi | ((uint64_t)(kLboA >> 4) << 16) |
                            (uint64_t)((((smem_base + sm.off_tile[tile_id]) >> 4) & 0x3FFFu) +
                                       (uint32_t)(g.a_k0 >> 3) * (kLboA >> 4));
          const uint32_t idesc = make_idesc_f16(FMT, 2 * kTileRows, g.Np);
          const uint32_t lbo_b = (uint32_t)(g.Np / 2) * 16;       // a half stage holds Np / 2 weight rows
          const uint32_t d_tmem = tmem_base + g.d_col;
          uint32_t acc = g.accumulate == 2 ? (t > 0 ? 1u : 0u) : g.accumulate;
          for (int k0 = 0; k0 < g.Kp; k0 += g.kc) {
            const int kc = min((int)g.kc, g.Kp - k0);
            mbar_wait_u(bar_w_full + st * 8, wph);
            mbar_wait_cluster_u(bar_w_peer + st * 8, wph);
            uint64_t b_desc = desc_hi | ((uint64_t)(lbo_b >> 4) << 16) |
                              (uint64_t)(((ring_addr + st * stage_bytes) >> 4) & 0x3FFFu);
            for (int ks = 0; ks < kc; ks += 16) {
              umma_f16_pair_u(d_tmem, a_desc, b_desc, idesc, acc);
              acc = 1;
              a_desc += 2 * (kLboA >> 4);
              b_desc += 2 * (lbo_b >> 4);
            }
            umma_commit_pair_mc_elect(bar_w_empty + st * 8, (uint16_t)3);    // the stage is free in both CTAs
            if (++st == nstage) { st = 0; wph ^= 1; }
          }
        }
        umma_commit_pair_mc_elect(bar_acc_full + (Gm & 3) * 8, (uint16_t)3);
        ++Gm;
        Ge += ph.n_sub;
      }
    }
  }
  __syncwarp();
}
// peer CTA, warp 1: the same walk over the ring stages, forwarding each local completion to the leader
__device__ __forceinline__ void relay_role(const Program& P, const SmemPlan& sm, long long npairs, int T,
                                           EngineShared& sh) {
  uint32_t nstage = sm.nstage;
  asm volatile("" : "+r"(nstage));
  uint32_t st = 0, wph = 0;
  const uint32_t bar_w_full = smem_u32(&sh.w_full[0]);
  const uint32_t peer0 = mapa_u32(smem_u32(&sh.w_peer[0]), 0);      // the leader's w_peer[0]
  const int lane = threadIdx.x & 31;
  for (long long pr = blockIdx.x / 2; pr < npairs; pr += gridDim.x / 2)
    for (int t = 0; t < T; ++t)
      for (int gi = 0; gi < P.n_gemms; ++gi) {
        const Gemm g = P.g[gi];
        for (int k0 = 0; k0 < g.Kp; k0 += g.kc) {
          mbar_wait_u(bar_w_full + st * 8, wph);
          if (lane == 0) mbar_arrive_cluster(peer0 + st * 8);
          __syncwarp();
          if (++st == nstage) { st = 0; wph ^= 1; }
        }
      }
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float fast_exp(float x) {   // single MUFU.EX2, flush-to-zero
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
  return y;
}
template <int ACT>
__device__ __forceinline__ float tc_act_t(float x) {
  if (ACT == BD_ACT_ELU) return fmaxf(x, fast_exp(fminf(x, 0.f)) - 1.f);   // e^min(x,0) - 1 >= x for x < 0, = 0 for x >= 0
  if (ACT == BD_ACT_RELU) return fmaxf(x, 0.f);
  if (ACT == BD_ACT_TANH) return fast_tanh(x);
  return x;
}
template <int ACT>
__device__ __forceinline__ float tc_dact_from_out(float y) {
  if (ACT == BD_ACT_ELU) return fminf(y, 0.f) + 1.f;      // y > 0 ? 1 : y + 1
  if (ACT == BD_ACT_RELU) return y > 0.f ? 1.f : 0.f;
  if (ACT == BD_ACT_TANH) return 1.f - y * y;
  return 1.f;
}
// sigmoid through the single-MUFU tanh: s(x) = 0.5 tanh(x/2) + 0.5
__device__ __forceinline__ float sigmoid_via_tanh(float x) { return fmaf(0.5f, fast_tanh(0.5f * x), 0.5f); }

template <int FMT>
__device__ __forceinline__ void store8(uint8_t* p, const float* v) {
  *reinterpret_cast<uint4*>(p) =
      make_uint4(Half16<FMT>::pack2(v[0], v[1]), Half16<FMT>::pack2(v[2], v[3]),
                 Half16<FMT>::pack2(v[4], v[5]), Half16<FMT>::pack2(v[6], v[7]));
}

// ---------------------------------------------------------------------------------------------
// Activation of an accumulator pair -> packed 16-bit pair.  fp16 ELU: the exponential stays in fp32
// (e^x - 1 formed on a 10-bit exponential loses the small negative outputs to cancellation: measured
// 2e-2 instead of 7e-3 on the c3 CEM action) but the scale and the -1 run as packed fp32x2
// instructions, so a pair costs 9 instructions instead of 11.
// ---------------------------------------------------------------------------------------------
template <int FMT, int ACT>
__device__ __forceinline__ uint32_t act_pack2(float a, float b) {
  if (ACT == BD_ACT_ELU) {
    // max(x, 2^(min(x log2 e, 0)) - 1)
    uint64_t xv, tv;
    asm("mov.b64 %0, {%1, %2};" : "=l"(xv) : "f"(a), "f"(b));
    asm("mul.f32x2 %0, %1, %2;" : "=l"(tv) : "l"(xv), "l"(0x3FB8AA3B3FB8AA3Bull));   // log2 e
    float t0, t1;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(t0), "=f"(t1) : "l"(tv));
    float e0, e1;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fminf(t0, 0.f)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fminf(t1, 0.f)));
    uint64_t ev;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ev) : "f"(e0), "f"(e1));
    asm("add.f32x2 %0, %1, %2;" : "=l"(ev) : "l"(ev), "l"(0xBF800000BF800000ull));   // - 1
    asm("mov.b64 {%0, %1}, %2;" : "=f"(e0), "=f"(e1) : "l"(ev));
    return Half16<FMT>::pack2(fmaxf(a, e0), fmaxf(b, e1));
  }
  if (FMT == 0 && ACT == BD_ACT_RELU) {
    uint32_t x = Half16<0>::pack2(a, b), y;
    asm("max.f16x2 %0, %1, %2;" : "=r"(y) : "r"(x), "r"(0u));
    return y;
  }
  return Half16<FMT>::pack2(tc_act_t<ACT>(a), tc_act_t<ACT>(b));
}
// act'(x) from the packed activation output y
template <int FMT, int ACT>
__device__ __forceinline__ uint32_t dact_pack2(uint32_t y) {
  if (FMT == 0) {
    uint32_t d;
    if (ACT == BD_ACT_ELU) {            // y > 0 ? 1 : y + 1  =  min(y, 0) + 1
      asm("{\n\t.reg .b32 t;\n\tmin.f16x2 t, %1, %2;\n\tadd.f16x2 %0, t, %3;\n\t}"
          : "=r"(d) : "r"(y), "r"(0u), "r"(0x3C003C00u));
    } else if (ACT == BD_ACT_RELU) {
      asm("set.gt.f16x2.f16x2 %0, %1, %2;" : "=r"(d) : "r"(y), "r"(0u));
    } else if (ACT == BD_ACT_TANH) {    // 1 - y^2
      asm("{\n\t.reg .b32 t;\n\tneg.f16x2 t, %1;\n\tfma.rn.f16x2 %0, t, %1, %2;\n\t}"
          : "=r"(d) : "r"(y), "r"(0x3C003C00u));
    } else {
      d = 0x3C003C00u;
    }
    return d;
  }
  const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&y);
  const float2 f = __bfloat1622float2(h);
  return Half16<FMT>::pack2(tc_dact_from_out<ACT>(f.x), tc_dact_from_out<ACT>(f.y));
}
// The five saved planes of a GRU gate column (c_r, c_z, c_n, c_nr, z; see EPI_GRU) for 8 columns.
template <int FMT>
__device__ __forceinline__ void gru_coeff_planes(const float* r, const float* z, const float* hn,
                                                 const float* omn2, const float* hmn, uint4* pl) {
  uint32_t cr[4], cz[4], cn[4], cnr[4], zz[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (FMT == 0) {
      const uint32_t R2 = Half16<0>::pack2(r[2 * j], r[2 * j + 1]), Z2 = Half16<0>::pack2(z[2 * j], z[2 * j + 1]);
      const uint32_t HN2 = Half16<0>::pack2(hn[2 * j], hn[2 * j + 1]);
      const uint32_t ON2 = Half16<0>::pack2(omn2[2 * j], omn2[2 * j + 1]);
      const uint32_t HM2 = Half16<0>::pack2(hmn[2 * j], hmn[2 * j + 1]);
      asm("{\n\t.reg .b32 omz, omr, t;\n\t"
          "sub.f16x2 omz, %5, %6;\n\t"          // 1 - z
          "mul.f16x2 %2, omz, %8;\n\t"          // c_n = (1 - z)(1 - n^2)
          "sub.f16x2 omr, %5, %7;\n\t"          // 1 - r
          "mul.f16x2 t, %2, %9;\n\t"
          "mul.f16x2 t, t, %7;\n\t"
          "mul.f16x2 %0, t, omr;\n\t"           // c_r = c_n hn r (1 - r)
          "mul.f16x2 %3, %2, %7;\n\t"           // c_nr = c_n r
          "mul.f16x2 t, %10, %6;\n\t"
          "mul.f16x2 %1, t, omz;\n\t"           // c_z = (h - n) z (1 - z)
          "mov.b32 %4, %6;\n\t}"
          : "=&r"(cr[j]), "=&r"(cz[j]), "=&r"(cn[j]), "=&r"(cnr[j]), "=&r"(zz[j])
          : "r"(0x3C003C00u), "r"(Z2), "r"(R2), "r"(ON2), "r"(HN2), "r"(HM2));
    } else {
      float c_r[2], c_z[2], c_n[2], c_nr[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int i = 2 * j + k;
        c_n[k] = (1.f - z[i]) * omn2[i];
        c_r[k] = c_n[k] * hn[i] * r[i] * (1.f - r[i]);
        c_nr[k] = c_n[k] * r[i];
        c_z[k] = hmn[i] * z[i] * (1.f - z[i]);
      }
      cr[j] = Half16<FMT>::pack2(c_r[0], c_r[1]); cz[j] = Half16<FMT>::pack2(c_z[0], c_z[1]);
      cn[j] = Half16<FMT>::pack2(c_n[0], c_n[1]); cnr[j] = Half16<FMT>::pack2(c_nr[0], c_nr[1]);
      zz[j] = Half16<FMT>::pack2(z[2 * j], z[2 * j + 1]);
    }
  }
  pl[0] = make_uint4(cr[0], cr[1], cr[2], cr[3]);
  pl[1] = make_uint4(cz[0], cz[1], cz[2], cz[3]);
  pl[2] = make_uint4(cn[0], cn[1], cn[2], cn[3]);
  pl[3] = make_uint4(cnr[0], cnr[1], cnr[2], cnr[3]);
  pl[4] = make_uint4(zz[0], zz[1], zz[2], zz[3]);
}

template <int FMT, int ACT, bool WITH_ACTOR, bool PROF, bool CLUSTER>
__global__ void __launch_bounds__(kThreads2, 1) rollout_fwd_kernel(const __grid_constant__ RolloutArgs A_) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const RolloutArgs& a = A_;
  uint8_t* smem = smem_raw;
  __shared__ EngineShared sh;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform: the role code stays on the uniform datapath
  const uint32_t R = CLUSTER ? (uint32_t)a.nranks : 1u;   // compile-time 1 keeps the single-CTA path lean
  const uint32_t rank = R > 1 ? (uint32_t)blockIdx.x % R : 0u;   // = %cluster_ctarank for (R,1,1) clusters; provably uniform
  __shared__ Program sprog;
  stage_program(sprog, CLUSTER ? a.prog[rank] : a.prog[0]);
  // (weight-share clusters -- multicast weight stages for CTAs with their own row tiles -- measured no gain and are
  // compiled out: a run-time cluster size costs the issuer a constant load and a branch per ring stage)
  constexpr uint32_t WS = 1u;
  const uint32_t tmem_base = engine_setup(sh, a.sm.nstage, R, kEpiThreads2, WS);
  uint64_t* const acc_full = sh.acc_full;
  uint64_t* const epi_done = sh.epi_done;
  long long prof_c0 = 0, prof_g0 = 0;
  if (PROF && tid == 0) {
    prof_c0 = clock64();
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(prof_g0));
  }

  const long long ntiles = (a.N + kTileRows - 1) / kTileRows;
  const Program& P = sprog;
  const long long tile0 = blockIdx.x / R, tstride = gridDim.x / R;

  if (warp == 0) {
    if (!(PROF && (a.dbg & 4)))
    producer_role(P, a.sm, a.wpack, ntiles, a.T, smem, sh, &a.pf, R, WS, PROF && (a.dbg & 2));
  } else if (warp == 1) {
    issuer_role<FMT, PROF>(P, a.sm, ntiles, a.T, smem, sh, tmem_base, a.prof, R, WS, (a.dbg & 4) != 0,
                           WITH_ACTOR ? &a.x0 : nullptr);
  } else {
    // =========================================================== epilogue warps
    // 16 warps: TMEM quadrant q = warp % 4 (a warp reaches lanes [32 q, 32 q + 32) only), column part
    // (warp - 2) / 4 in 0..3.  Four warps per scheduler hide the TMEM-load / MUFU latencies that two could
    // not (the epilogues, not the MMAs, bound this engine); chunks are 16 columns (activations) or 8
    // (GRU / prior output) so the whole kernel fits the 112 registers 576 threads leave.
    const int q = warp & 3, part = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t rowoff = (uint32_t)((row >> 3) * 128 + (row & 7) * 16);   // this row inside a KM8 column group
    const int etid = tid - 64;
    const int Be = a.Be, S = a.S, Ad = a.A;
    const bool be4 = (Be & 3) == 0;
    const bool wr0 = rank == 0;       // replicated phases: only rank 0 writes the global outputs
    // 16 B of an operand tile (8 columns of one row, already packed): local store, mirrored into every
    // peer's shared memory in column-split mode
    auto put16 = [&](uint8_t* p, const uint4 u) {
      if (R == 1) {
        *reinterpret_cast<uint4*>(p) = u;
      } else {
        const uint32_t la = smem_u32(p);
        for (uint32_t k = 0; k < R; ++k) st_cluster_v4(mapa_u32(la, k), u);
      }
    };
    // publish epilogue completion Ge: one arrival per warp (512 thread arrivals on one mbarrier serialise in
    // shared memory), on every rank in column-split mode
    auto epi_arrive = [&](uint32_t ge) {
      if (R == 1) {
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&epi_done[ge & 7]);
      } else {
        fence_proxy_async_all();
        __syncwarp();
        if (lane < (int)R) mbar_arrive_cluster(mapa_u32(smem_u32(&epi_done[ge & 7]), (uint32_t)lane));
      }
    };
    uint32_t Ge = 0, Gm = 0;
    for (long long tile = tile0; tile < ntiles; tile += tstride) {
      const long long grow = tile * kTileRows + row;
      const bool rvalid = grow < a.N;
      // ---------------- tile initialisation: B0 <- prev_belief | 1, B1 <- 0 | 1, SA <- prev_state | . | 1
      {
        uint8_t* B0 = smem + a.sm.off_tile[0];
        uint8_t* B1 = smem + a.sm.off_tile[1];
        uint8_t* SA = smem + a.sm.off_tile[2];
        uint8_t* H = smem + a.sm.off_tile[3];
        // one 8-column group (16 B of the operand tile) per item; consecutive threads take
        // consecutive rows of the same group -> conflict-free 16 B shared stores
        const int gb = a.Kp_b >> 3;
        if (be4 && (reinterpret_cast<uintptr_t>(a.prev_belief) & 15) == 0) {
          // 16-byte loads, four items (8 columns of one row each) requested per thread before the first is used:
          // one item at a time with scalar loads left ~9 dependent memory round trips in this prologue (~17 K
          // cycles per tile -- a third of a whole MLP-forward tile)
          for (int i0 = etid; i0 < kTileRows * gb; i0 += 4 * kEpiThreads2) {
            float4 lo[4], hi[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int i = i0 + u * kEpiThreads2;
              const int kg = i / kTileRows, r = i - kg * kTileRows;
              const long long gr = tile * kTileRows + r;
              const long long sr = a.cem_cl ? gr / a.cem_cl : gr;     // CEM: latents are per batch row
              lo[u] = hi[u] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (i < kTileRows * gb && gr < a.N) {
                const float* src = a.prev_belief + sr * Be + kg * 8;
                if (kg * 8 + 4 <= Be) lo[u] = *reinterpret_cast<const float4*>(src);
                if (kg * 8 + 8 <= Be) hi[u] = *reinterpret_cast<const float4*>(src + 4);
              }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int i = i0 + u * kEpiThreads2;
              if (i >= kTileRows * gb) break;
              const int kg = i / kTileRows, r = i - kg * kTileRows;
              float v[8] = {lo[u].x, lo[u].y, lo[u].z, lo[u].w, hi[u].x, hi[u].y, hi[u].z, hi[u].w}, z[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {           // (Be is a multiple of 4: the ones column starts a 4-group)
                const int k = kg * 8 + j;
                if (k == Be) v[j] = 1.f;
                z[j] = (k == Be) ? 1.f : 0.f;
              }
              store8<FMT>(B0 + kg * kLboA + r * 16, v);
              if (a.has_b1) store8<FMT>(B1 + kg * kLboA + r * 16, z);
            }
          }
        } else {
        for (int i = etid; i < kTileRows * gb; i += kEpiThreads2) {
          const int kg = i / kTileRows, r = i - kg * kTileRows;
          const long long gr = tile * kTileRows + r;
          const long long sr = a.cem_cl ? gr / a.cem_cl : gr;     // CEM: latents are per batch row
          float v[8], z[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int k = kg * 8 + j;
            v[j] = (k < Be) ? ((gr < a.N) ? a.prev_belief[sr * Be + k] : 0.f) : (k == Be ? 1.f : 0.f);
            z[j] = (k == Be) ? 1.f : 0.f;
          }
          store8<FMT>(B0 + kg * kLboA + r * 16, v);
          if (a.has_b1) store8<FMT>(B1 + kg * kLboA + r * 16, z);
        }
        }
        const int gs = a.Kp_sa >> 3;
        for (int i = etid; i < kTileRows * gs; i += kEpiThreads2) {
          const int kg = i / kTileRows, r = i - kg * kTileRows;
          const long long gr = tile * kTileRows + r;
          const long long sr = a.cem_cl ? gr / a.cem_cl : gr;
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int k = kg * 8 + j;
            float x = 0.f;
            if (k < S) x = (gr < a.N && a.prev_state) ? a.prev_state[sr * S + k] : 0.f;
            else if (k == S + Ad) x = 1.f;
            else if (!WITH_ACTOR && k < S + Ad && gr < a.N && a.ext_actions) x = a.ext_actions[gr * Ad + (k - S)];
            v[j] = x;
          }
          store8<FMT>(SA + kg * kLboA + r * 16, v);
        }
        const int gh = a.Kp_h >> 3;
        for (int i = etid; i < kTileRows * gh; i += kEpiThreads2) {
          const int kg = i / kTileRows, r = i - kg * kTileRows;
          *reinterpret_cast<uint4*>(H + kg * kLboA + r * 16) = make_uint4(0, 0, 0, 0);
        }
        epi_arrive(Ge);
        ++Ge;
      }
      for (int t = 0; t < a.T; ++t) {
        const int par = t & 1;
        uint8_t* Bnxt = smem + a.sm.off_tile[1 ^ par];
        uint8_t* SAt = smem + a.sm.off_tile[2];
        const long long orow = (long long)t * a.N + grow;        // row in (T,N,.) outputs
        for (int pi = 0; pi < P.n_phases; ++pi) {
          const Phase ph = P.p[pi];
          long long e0 = 0, e1 = 0;
          if (PROF) e0 = clock64();
          const uint32_t tacc = trow + ph.d_col;
          switch ((PROF && (a.dbg & 1)) ? 0 : ph.epi) {
            case EPI_ACT_H: {
              // What this layer also leaves in HBM for a backward pass: 1 = the hidden activation itself
              // (MLP forward; padded rows zero: the image doubles as the wgrad operand), 2 = act'(output)
              // (embed x, prior h, fused heads; padding needs no mask: every consumer multiplies it into an
              // accumulator whose padded rows / columns are exactly zero)
              int smode = 0, skp = 0;
              uint16_t* simg = nullptr;
              if (ph.aux0 >= 3 && ph.aux0 < 3 + BD_MAX_LAYERS) {
                if (a.sv_mlp[ph.aux0 - 3]) {
                  smode = 1; skp = a.kact_sv ? a.kact_sv : ph.Kp_out;
                  // (per (t, tile): the MLP forward has T = 1; the rollout saves the actor's hidden layers per step)
                  simg = a.sv_mlp[ph.aux0 - 3] + ((size_t)t * ntiles + tile) * kTileRows * skp + row * 8;
                }
              } else if (ph.aux0 == 1 || ph.aux0 == 2) {
                if (a.sv_xa) {
                  smode = 2; skp = ph.aux0 == 1 ? a.kb_sv : a.kh_sv;
                  simg = (ph.aux0 == 1 ? a.sv_xa : a.sv_ha) + ((size_t)t * ntiles + tile) * kTileRows * skp + row * 8;
                }
              } else if (ph.aux0 >= 16) {
                if (a.sv_hd[ph.aux0 - 16]) {
                  smode = 2; skp = a.kh_hd;
                  simg = a.sv_hd[ph.aux0 - 16] + ((size_t)t * ntiles + tile) * kTileRows * skp + row * 8;
                }
              }
              mbar_wait(&acc_full[Gm & 3], (Gm >> 2) & 1);
              tc_fence_after_sync();
              if (PROF) e1 = clock64();
              // (a belief tile as output = the ping-pong half that is free at this point: the fused
              // heads' second hidden tile)
              const int otile = ph.out_tile < 2 ? (ph.out_tile ^ par) : ph.out_tile;
              uint8_t* out = smem + a.sm.off_tile[otile] + rowoff;
              const int nv = ph.n_valid;
              for (int sub = 0; sub < ph.n_sub; ++sub) {
                // this rank's columns are [col0, Kp_out) (col0 = 0 unless column-split); the
                // accumulator holds them at TMEM columns [0, Np)
                const int c_lo = sub == 0 ? ph.col0 : ph.split;
                const int c_hi = (sub == ph.n_sub - 1) ? ph.Kp_out : ph.split;
                if (sub > 0) {   // publish the first column range: the next layer's K-slab 0 may start
                  tc_fence_before_sync();
                  epi_arrive(Ge);
                  ++Ge;
                }
                // 16-column chunks, the next chunk's TMEM load in flight while this one is processed
                float v[2][16];
                const int c0 = c_lo + part * 16;
                auto load = [&](int c, float* dst) {
                  const int ca = c - ph.col0;           // accumulator column
                  if (ca < ph.Np) {
                    tmem_ld16(tacc + ca, dst);
                  } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) dst[j] = 0.f;
                  }
                };
                if (c0 < c_hi) load(c0, v[0]);
#pragma unroll
                for (int it = 0; it < (16 + kParts2 - 1) / kParts2; ++it) {
                  const int c = c0 + it * (16 * kParts2);
                  if (c < c_hi) {
                    tmem_ld_wait();
                    if (c + 16 * kParts2 < c_hi) load(c + 16 * kParts2, v[(it + 1) & 1]);
                    uint32_t y[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) y[j] = act_pack2<FMT, ACT>(v[it & 1][2 * j], v[it & 1][2 * j + 1]);
                    if (c + 16 > nv) {   // the chunk holding the constant-1 (bias) column and zero padding
#pragma unroll
                      for (int j = 0; j < 8; ++j) {
                        const int col = c + 2 * j;
                        if (col >= nv) y[j] = (y[j] & 0xFFFF0000u) | (col == nv ? Half16<FMT>::kOne : 0u);
                        if (col + 1 >= nv) y[j] = (y[j] & 0x0000FFFFu) | ((col + 1 == nv ? Half16<FMT>::kOne : 0u) << 16);
                      }
                    }
                    uint8_t* p = out + (c >> 3) * kLboA;
                    put16(p, make_uint4(y[0], y[1], y[2], y[3]));
                    put16(p + kLboA, make_uint4(y[4], y[5], y[6], y[7]));
                    if (smode == 1) {          // the hidden image itself (rows past N: zeros)
                      uint16_t* gi = simg + (size_t)(c >> 3) * kTileRows * 8;
                      const uint32_t m = rvalid ? 0xFFFFFFFFu : 0u;
                      *reinterpret_cast<uint4*>(gi) = make_uint4(y[0] & m, y[1] & m, y[2] & m, y[3] & m);
                      *reinterpret_cast<uint4*>(gi + (size_t)kTileRows * 8) = make_uint4(y[4] & m, y[5] & m, y[6] & m, y[7] & m);
                    } else if (smode == 2 && c < skp) {
                      uint32_t dy[8];
#pragma unroll
                      for (int j = 0; j < 8; ++j) dy[j] = dact_pack2<FMT, ACT>(y[j]);
                      uint16_t* gi = simg + (size_t)(c >> 3) * kTileRows * 8;
                      *reinterpret_cast<uint4*>(gi) = make_uint4(dy[0], dy[1], dy[2], dy[3]);
                      if (c + 8 < skp)
                        *reinterpret_cast<uint4*>(gi + (size_t)kTileRows * 8) = make_uint4(dy[4], dy[5], dy[6], dy[7]);
                    }
                  }
                }
              }
            } break;
            case EPI_GRU: {
              // 8-column chunks: c = 8 (part + kParts2 it) (a slice is at most 64 columns wide)
              const int n0 = ph.aux0, Ns = ph.Np;
              const float* bold = (t == 0) ? a.prev_belief + (a.cem_cl ? grow / a.cem_cl : grow) * Be
                                           : a.beliefs + ((long long)(t - 1) * a.N + grow) * Be;
              float* bnew = a.beliefs + orow * Be;
              // previous belief h_t for this thread's columns.  fp16 mode: the 16-bit copy in the operand tile
              // (one 16-byte shared-memory load per chunk, no L2 round trip on the recurrence's critical path;
              // the carried state is thereby rounded to fp16 once per step -- 2^-12 relative, below the
              // operand rounding the GEMMs see anyway).  bf16 mode (7-bit mantissa): the fp32 master copy
              // from the previous step's output, requested before the accumulator wait.
              const uint8_t* Bcur = smem + a.sm.off_tile[par] + rowoff;
              constexpr int kGruIts = (8 + kParts2 - 1) / kParts2;
              float hb[kGruIts][8];
              if (FMT != 0) {
#pragma unroll
                for (int it = 0; it < kGruIts; ++it) {
                  const int col = n0 + (part + kParts2 * it) * 8;
#pragma unroll
                  for (int j4 = 0; j4 < 2; ++j4) {
                    const int cc = col + j4 * 4;
                    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (rvalid && (part + kParts2 * it) * 8 < Ns) {
                      if (be4 && cc + 3 < Be) x = *reinterpret_cast<const float4*>(bold + cc);
                      else {
                        if (cc < Be) x.x = bold[cc];
                        if (cc + 1 < Be) x.y = bold[cc + 1];
                        if (cc + 2 < Be) x.z = bold[cc + 2];
                        if (cc + 3 < Be) x.w = bold[cc + 3];
                      }
                    }
                    hb[it][j4 * 4] = x.x; hb[it][j4 * 4 + 1] = x.y; hb[it][j4 * 4 + 2] = x.z; hb[it][j4 * 4 + 3] = x.w;
                  }
                }
              }
              mbar_wait(&acc_full[Gm & 3], (Gm >> 2) & 1);
              tc_fence_after_sync();
              if (PROF) e1 = clock64();
#pragma unroll
              for (int it = 0; it < kGruIts; ++it) {
                const int c = (part + kParts2 * it) * 8;
                if (c < Ns) {
                  float r_[8], z_[8], in_[8], hn_[8], o[8];
                  tmem_ld8(tacc + c, in_);              // accumulator columns: IN | R | Z | HN
                  tmem_ld8(tacc + Ns + c, r_);
                  tmem_ld8(tacc + 2 * Ns + c, z_);
                  tmem_ld8(tacc + 3 * Ns + c, hn_);
                  const int col0 = n0 + c;
                  if (FMT == 0) {
                    const uint4 hu = *reinterpret_cast<const uint4*>(Bcur + (col0 >> 3) * kLboA);
                    const uint32_t hw[4] = {hu.x, hu.y, hu.z, hu.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hw[j]));
                      hb[it][2 * j] = f.x; hb[it][2 * j + 1] = f.y;
                    }
                  }
                  tmem_ld_wait();
                  float hmn[8], omn2[8];
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    const float r = sigmoid_via_tanh(r_[j]);
                    const float z = sigmoid_via_tanh(z_[j]);
                    const float n = fast_tanh(fmaf(r, hn_[j], in_[j]));
                    hmn[j] = hb[it][j] - n;
                    o[j] = fmaf(z, hmn[j], n);                 // (1-z) n + z h
                    omn2[j] = fmaf(-n, n, 1.f);                // 1 - n^2
                    r_[j] = r; z_[j] = z;
                  }
                  if (a.sv_gate && col0 < a.kb_sv && !(PROF && (a.dbg & 64))) {
                    // backward coefficients of this gate column per unit of dL/db' (SURVEY A.1 / tc_bptt.cuh):
                    //   c_n = (1-z)(1-n^2), c_r = c_n hn r (1-r), c_nr = c_n r, c_z = (h-n) z (1-z), and z for the
                    // carry.  The cancellation-prone factors (h - n, 1 - n^2) are formed in fp32 above; the
                    // products run on packed 16-bit pairs (their results are stored at that precision anyway).
                    uint4 pl[5];
                    gru_coeff_planes<FMT>(r_, z_, hn_, omn2, hmn, pl);
                    uint16_t* img = a.sv_gate + ((size_t)t * ntiles + tile) * 5 * kTileRows * a.kb_sv +
                                    (size_t)(col0 >> 3) * kTileRows * 8 + row * 8;
                    const size_t plane = (size_t)kTileRows * a.kb_sv;
#pragma unroll
                    for (int k = 0; k < 5; ++k) *reinterpret_cast<uint4*>(img + k * plane) = pl[k];
                  }
                  if (rvalid && !(PROF && (a.dbg & 128))) {
#pragma unroll
                    for (int j4 = 0; j4 < 2; ++j4) {
                      const int col = col0 + j4 * 4;
                      if (be4 && col + 3 < Be)
                        *reinterpret_cast<float4*>(bnew + col) = make_float4(o[j4 * 4], o[j4 * 4 + 1], o[j4 * 4 + 2], o[j4 * 4 + 3]);
                      else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) if (col + j < Be) bnew[col + j] = o[j4 * 4 + j];
                      }
                    }
                  }
                  if (col0 + 8 > Be) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) if (col0 + j >= Be) o[j] = (col0 + j == Be) ? 1.f : 0.f;
                  }
                  if (col0 < a.Kp_b)
                    put16(Bnxt + (col0 >> 3) * kLboA + rowoff,
                          make_uint4(Half16<FMT>::pack2(o[0], o[1]), Half16<FMT>::pack2(o[2], o[3]),
                                     Half16<FMT>::pack2(o[4], o[5]), Half16<FMT>::pack2(o[6], o[7])));
                }
              }
            } break;
            case EPI_PRIOR_OUT: {
              const int Sp = ph.Np;    // mean at [0,Sp), raw std at [Sp, 2Sp)
              // row of this thread in the state-noise tensor
              const long long erow = a.cem_cl
                  ? (long long)t * ((a.N / a.cem_cl) * a.cem_c) + (grow / a.cem_cl) * a.cem_c + a.cem_c0 + grow % a.cem_cl
                  : orow;
              // 8-column chunks cc = 8 part + 32 it; the first chunk's noise is requested before the wait
              float eps[8];
              const int cc0 = part * 8;
#pragma unroll
              for (int j = 0; j < 8; ++j) eps[j] = (rvalid && cc0 + j < S && !(PROF && (a.dbg & 32))) ? a.eps_s[erow * S + cc0 + j] : 0.f;
              mbar_wait(&acc_full[Gm & 3], (Gm >> 2) & 1);
              tc_fence_after_sync();
              if (PROF) e1 = clock64();
              // The (mean, std, state) rows of a tile are one contiguous block of each fp32 output tensor.  A thread
              // owns 8 consecutive columns of its row, so direct stores are 8-byte pieces at a 4 S-byte row stride:
              // 32 sectors per warp instruction, 12 instructions per chunk -- measured 6.5 K cycles per time step of
              // store issue on the recurrence's critical path (the next step's actor waits for this epilogue).
              // Instead the three tiles are staged row-major in the H tile (dead here: its last reader was this
              // phase's MMA, its next writer is a later epilogue of these same warps) with conflict-free 8-byte
              // shared stores and copied out by all epilogue threads as contiguous 16-byte stores.
              const bool out_cta = a.means != nullptr && wr0 && !(PROF && (a.dbg & 16));     // uniform per CTA
              const bool s2 = (S & 1) == 0;      // rows are 8-byte aligned
              const bool staged = out_cta && s2 && (3 * S * 4 <= a.Kp_h * 2);
              const bool want_out = out_cta && rvalid && !staged;
              float* stage = reinterpret_cast<float*>(smem + a.sm.off_tile[TILE_H]);
              for (int cc = cc0; cc < Sp; cc += 8 * kParts2) {
                if (cc != cc0) {
#pragma unroll
                  for (int j = 0; j < 8; ++j) eps[j] = (rvalid && cc + j < S) ? a.eps_s[erow * S + cc + j] : 0.f;
                }
                float m_[8], s_[8], st[8];
                tmem_ld8(tacc + cc, m_);
                tmem_ld8(tacc + Sp + cc, s_);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  // softplus(x) = max(x,0) + log(1 + exp(-|x|)) with single-MUFU exp / log
                  s_[j] = fmaxf(s_[j], 0.f) + __logf(1.f + fast_exp(-fabsf(s_[j]))) + a.min_std;
                  st[j] = fmaf(s_[j], eps[j], m_[j]);
                }
                if (cc + 8 <= S) {
                  *reinterpret_cast<uint4*>(SAt + (cc >> 3) * kLboA + rowoff) =
                      make_uint4(Half16<FMT>::pack2(st[0], st[1]), Half16<FMT>::pack2(st[2], st[3]),
                                 Half16<FMT>::pack2(st[4], st[5]), Half16<FMT>::pack2(st[6], st[7]));
                } else {
#pragma unroll
                  for (int j = 0; j < 8; ++j) if (cc + j < S) store1<FMT>(SAt, row, cc + j, st[j]);
                }
                if (staged) {
                  float* sm_ = stage + row * S + cc;
#pragma unroll
                  for (int j = 0; j < 8; j += 2) {
                    if (cc + j + 1 < S) {
                      *reinterpret_cast<float2*>(sm_ + j) = make_float2(m_[j], m_[j + 1]);
                      *reinterpret_cast<float2*>(sm_ + kTileRows * S + j) = make_float2(s_[j], s_[j + 1]);
                      *reinterpret_cast<float2*>(sm_ + 2 * kTileRows * S + j) = make_float2(st[j], st[j + 1]);
                    }
                  }
                } else if (want_out) {
                  float* om = a.means + orow * S + cc;
                  float* od = a.stds + orow * S + cc;
                  float* os = a.states + orow * S + cc;
#pragma unroll
                  for (int j = 0; j < 8; j += 2) {
                    if (s2 && cc + j + 1 < S) {
                      *reinterpret_cast<float2*>(om + j) = make_float2(m_[j], m_[j + 1]);
                      *reinterpret_cast<float2*>(od + j) = make_float2(s_[j], s_[j + 1]);
                      *reinterpret_cast<float2*>(os + j) = make_float2(st[j], st[j + 1]);
                    } else {
                      if (cc + j < S) { om[j] = m_[j]; od[j] = s_[j]; os[j] = st[j]; }
                      if (cc + j + 1 < S) { om[j + 1] = m_[j + 1]; od[j + 1] = s_[j + 1]; os[j + 1] = st[j + 1]; }
                    }
                  }
                }
              }
              if (staged) {      // uniform per CTA: all epilogue threads take part
                asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads2) : "memory");
                const long long row0 = tile * kTileRows;
                const int nrows = (int)((a.N - row0) < kTileRows ? (a.N - row0) : kTileRows);
                const int nfl = nrows * S;                                  // floats per output (even)
                const long long g0 = ((long long)t * a.N + row0) * S;
                float* outs[3] = {a.means, a.stds, a.states};
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                  float* dst = outs[k] + g0;
                  const float* src = stage + k * kTileRows * S;
                  if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (nfl & 3) == 0) {
                    for (int i = etid * 4; i < nfl; i += kEpiThreads2 * 4)
                      *reinterpret_cast<float4*>(dst + i) = *reinterpret_cast<const float4*>(src + i);
                  } else {
                    for (int i = etid * 2; i < nfl; i += kEpiThreads2 * 2)
                      *reinterpret_cast<float2*>(dst + i) = *reinterpret_cast<const float2*>(src + i);
                  }
                }
                // (the H tile's next writer must not overtake the copy-out of a slower warp)
                asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads2) : "memory");
              }
              if (!WITH_ACTOR && part == 0 && t + 1 < a.T && a.ext_actions) {   // next step's given action -> [s ; a] tile
                for (int j = 0; j < Ad; ++j)
                  store1<FMT>(SAt, row, S + j, rvalid ? a.ext_actions[((long long)(t + 1) * a.N + grow) * Ad + j] : 0.f);
              }
            } break;
            case EPI_ACTOR_OUT: {
              // (src/models.py:513-516, src/dreamer.py:435-443) the action noise is requested before the wait;
              // single-MUFU tanh / exp / log: the 16-bit modes round the action to 10 / 7 mantissa bits anyway
              float ea[16];
              if (WITH_ACTOR && part == 0) {
#pragma unroll
                for (int j = 0; j < 16; ++j) ea[j] = (j < Ad && rvalid) ? a.eps_a[orow * Ad + j] : 0.f;
              }
              mbar_wait(&acc_full[Gm & 3], (Gm >> 2) & 1);
              tc_fence_after_sync();
              if (PROF) e1 = clock64();
              if (WITH_ACTOR && part == 0) {
                const int Ap = ph.Np;
                float m_[16], s_[16];
                tmem_ld16(tacc, m_);
                tmem_ld16(tacc + Ap, s_);
                tmem_ld_wait();
                const float inv_ms = 1.f / a.cfg.mean_scale;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  if (j < Ad) {
                    const float mean = a.cfg.mean_scale * fast_tanh(m_[j] * inv_ms);
                    const float xs = s_[j] + a.cfg.raw_init_std;
                    const float sd = (xs > 20.f ? xs : fmaxf(xs, 0.f) + __logf(1.f + fast_exp(-fabsf(xs)))) + a.cfg.min_std;
                    const float act = fast_tanh(fmaf(ea[j], sd, mean));
                    store1<FMT>(SAt, row, S + j, act);
                    if (rvalid && wr0) {
                      a.actions[orow * Ad + j] = act;
                      a.actor_raw[orow * 2 * Ad + j] = m_[j];
                      a.actor_raw[orow * 2 * Ad + Ad + j] = s_[j];
                    }
                  }
                }
              }
            } break;
            case EPI_HEAD_OUT: {
              mbar_wait(&acc_full[Gm & 3], (Gm >> 2) & 1);
              tc_fence_after_sync();
              if (PROF) e1 = clock64();
              if (part == 0) {
                float v[8];
                tmem_ld8(tacc, v);
                tmem_ld_wait();
                if (rvalid && wr0 && a.head_out[ph.aux0]) a.head_out[ph.aux0][orow] = v[0];
              }
            } break;
            case EPI_STORE_OUT: {
              mbar_wait(&acc_full[Gm & 3], (Gm >> 2) & 1);
              tc_fence_after_sync();
              if (PROF) e1 = clock64();
              const int nv = ph.n_valid;
              for (int c = part * 16; c < ph.Np; c += 16 * kParts2) {
                float v[16];
                tmem_ld16(tacc + c, v);
                tmem_ld_wait();
                if (rvalid && wr0) {
#pragma unroll
                  for (int j = 0; j < 16; ++j)
                    if (c + j < nv) a.mlp_out[orow * nv + c + j] = v[j];
                }
              }
            } break;
            default: {
              mbar_wait(&acc_full[Gm & 3], (Gm >> 2) & 1);
              tc_fence_after_sync();
              if (PROF) e1 = clock64();
              for (int sub = 1; sub < ph.n_sub; ++sub) {   // (debug skip of an ACT epilogue: keep the completion count)
                tc_fence_before_sync();
                epi_arrive(Ge);
                ++Ge;
              }
            } break;
          }
          tc_fence_before_sync();
          epi_arrive(Ge);
          if (PROF && blockIdx.x == 0 && lane == 0 && (warp == 2 || warp == 6)) {
            const int o = pi * 8 + (warp == 2 ? 3 : 5);
            prof_add(&a.prof[o], e1 - e0);                          // epilogue: wait for the accumulator
            prof_add(&a.prof[o + 1], clock64() - e1);               // epilogue: work
          }
          ++Ge;
          ++Gm;
        }
      }
      // lambda-return tail of the fused Dreamer rollout: this thread wrote reward / value of its row
      // for every step (EPI_HEAD_OUT, part 0), so program order makes them visible to it.  Same
      // arithmetic (separate fp32 roundings, no FMA contraction) as f32::lambda_return_fwd_kernel.
      if (a.returns != nullptr && part == 0 && rvalid && wr0) {
        const float* rw = a.head_out[0];
        const float* vl = a.head_out[1];
        float last = vl[(long long)(a.T - 1) * a.N + grow];      // bootstrap = value[-1]
        float nextv = last;
        const float dl = __fmul_rn(a.lr_disc, a.lr_lam);
        for (int t = a.T - 1; t >= 0; --t) {
          const long long o = (long long)t * a.N + grow;
          const float inp = __fadd_rn(rw[o], __fmul_rn(__fmul_rn(a.lr_disc, nextv), a.lr_oml));
          last = __fadd_rn(inp, __fmul_rn(dl, last));
          a.returns[o] = last;
          nextv = vl[o];
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (PROF && tid == 0) {
    long long g1;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g1));
    if (blockIdx.x == 0) {
      prof_add(&a.prof[39 * 8 + 0], clock64() - prof_c0);   // whole-kernel SM cycles of CTA 0
      prof_add(&a.prof[39 * 8 + 1], g1 - prof_g0);          // same interval in ns
    }
    if (blockIdx.x < 160) {                        // per-CTA start / end timestamps
      uint32_t smid;
      asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
      a.prof[40 * 8 + blockIdx.x * 3 + 0] = prof_g0;
      a.prof[40 * 8 + blockIdx.x * 3 + 1] = g1;
      a.prof[40 * 8 + blockIdx.x * 3 + 2] = smid;
    }
  }
  if (R > 1 || WS > 1) cluster_sync_all();   // no rank leaves while peers may still write to / arrive on it
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ---------------------------------------------------------------------------------------------
// J-sample Monte-Carlo entropy of the tanh-Normal policy and its gradient wrt (mean, std)
// (src/models.py:725-733, 656-673), computed after the rollout from the saved raw actor outputs:
// it does not feed the recurrence, so it runs as a plain fully-parallel kernel over (t, row).
// SPLIT adjacent lanes share a row and take every SPLIT-th sample; their sums meet by shuffle.  One thread
// per row left c2 (35 000 row-steps) with 35 000 threads walking 100 dependent-latency MUFU chains: 58 us,
// 35 us with SPLIT = 8.  At large row counts the machine is full anyway and the strided noise reads of a
// split cost more than they give (2^17 rows: 0.66 ms unsplit, 1.15 ms with 8), so the host picks SPLIT.
// Per-(row, action) constants of the sample loop and one Monte-Carlo sample.
struct EntRow { float mean, sd, inv_var, inv_sd, nhiv, lp_const; };
__device__ __forceinline__ EntRow ent_row(const bd_actor_cfg& cfg, float raw_mean, float raw_std) {
  EntRow p;
  p.mean = cfg.mean_scale * tanhf(raw_mean / cfg.mean_scale);
  p.sd = softplusf_(raw_std + cfg.raw_init_std) + cfg.min_std;
  p.inv_var = 1.f / (p.sd * p.sd); p.inv_sd = 1.f / p.sd;
  p.nhiv = -0.5f * p.inv_var; p.lp_const = -logf(p.sd) - 0.9189385332046727f;    // - log sd - log sqrt(2 pi)
  return p;
}
// Branch-free, three MUFU operations per sample (tanhf's two code paths diverged per lane and, with the two
// logarithms of 1 +- y, cost ~110 instructions per sample): with u = e^{-2x},
//   y = tanh x = 2 / (1 + u) - 1,   atanh(y) = x,   log|d tanh / dx| = 2 (ln 2 - x - ln(1 + u))
// (src/models.py:668-673 evaluates exactly this expression at x' = atanh(y)); no cancellation in 1 - y.
// Clamped samples (|tanh x| > kClamp, |x| > ~8.7): the reference inverts the CLAMPED y (src/models.py:656-666), so
// x' = +-atanh(kClamp) (kXc) and the log-det is the constant at that point (kLadjC).
__device__ __forceinline__ void ent_sample(const EntRow& p, float e, float kXc, float kLadjC, float& lp_sum,
                                           float& dm_sum, float& ds_sum) {
  const float kClamp = 0.99999997f, kLog2 = 0.6931471805599453f;
  const float x = fmaf(e, p.sd, p.mean);
  float u, r, l2w;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(u) : "f"(x * -2.8853900817779268f));     // -2 log2(e)
  const float w = 1.f + u;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(w));
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2w) : "f"(w));
  const float y = fmaf(2.f, r, -1.f);
  const bool clamped = fabsf(y) > kClamp;
  const float gate = clamped ? 0.f : 1.f;
  const float yc = clamped ? copysignf(kClamp, y) : y;
  // (explicit roundings: both entropy kernels inline this function and must agree bit for bit -- rows are
  // independent of how a batch is split, tests/test_gpu_properties.py)
  const float d = clamped ? __fsub_rn(copysignf(kXc, y), p.mean) : __fmul_rn(e, p.sd);
  const float ladj = clamped ? kLadjC : fmaf(-2.f, fmaf(l2w, kLog2, x), 2.f * kLog2);
  const float dd = __fmul_rn(d, d), div = __fmul_rn(d, p.inv_var);
  lp_sum = __fadd_rn(lp_sum, __fsub_rn(fmaf(dd, p.nhiv, p.lp_const), ladj));     // log N(x'; mean, sd) - log|d tanh / dx|
  const float dlp = fmaf(2.f, yc, -div);
  dm_sum = __fadd_rn(dm_sum, fmaf(gate, dlp, div));
  ds_sum = __fadd_rn(ds_sum, fmaf(__fmul_rn(gate, e), dlp, fmaf(__fmul_rn(dd, p.inv_var), p.inv_sd, -p.inv_sd)));
}

template <int SPLIT>
static __global__ void actor_entropy_kernel(const float* __restrict__ raw, const float* __restrict__ eps_e,
                                     bd_actor_cfg cfg, long long N, int A,
                                     float* __restrict__ entropy, float* __restrict__ dent) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int q = (int)(gid & (SPLIT - 1));
  const long long n_ = gid / SPLIT;
  const bool valid = n_ < N;
  const long long n = valid ? n_ : N - 1;          // no early exit: the shuffles below are warp-wide
  const int t = blockIdx.y;
  const long long r = (long long)t * N + n;
  const float kClamp = 0.99999997f;
  const float kXc = 0.5f * (logf(1.f + kClamp) - logf(1.f - kClamp));
  const float kLadjC = logf(1.f + kClamp) + logf(1.f - kClamp);        // = log(1 - kClamp^2)
  const int J = cfg.entropy_samples;
  float ent_acc = 0.f;
  for (int a = 0; a < A; ++a) {
    const EntRow p = ent_row(cfg, raw[r * 2 * A + a], raw[r * 2 * A + A + a]);
    const float* ee = eps_e + ((long long)t * J * N + n) * A + a;
    float lp_sum = 0.f, dm_sum = 0.f, ds_sum = 0.f;
#pragma unroll 4
    for (int j = q; j < J; j += SPLIT) ent_sample(p, ee[(long long)j * N * A], kXc, kLadjC, lp_sum, dm_sum, ds_sum);
#pragma unroll
    for (int o = 1; o < SPLIT; o <<= 1) {
      lp_sum += __shfl_xor_sync(0xffffffffu, lp_sum, o);
      dm_sum += __shfl_xor_sync(0xffffffffu, dm_sum, o);
      ds_sum += __shfl_xor_sync(0xffffffffu, ds_sum, o);
    }
    ent_acc += lp_sum;
    if (valid && q == 0) {
      dent[r * 2 * A + a] = -dm_sum / (float)J;
      dent[r * 2 * A + A + a] = -ds_sum / (float)J;
    }
  }
  if (valid && q == 0) entropy[r] = -ent_acc / (float)J;
}

// Large row counts, one action dimension, N a multiple of 4: a thread owns FOUR consecutive rows and reads their
// samples as one 16-byte load per j -- a block then reads 2 KB runs of the (T, J, N) noise tensor instead of 512-byte
// runs at a 4 N-byte stride (the scalar kernel streamed it at 1.1 TB/s: 0.66 ms at 2^17 start states).
static __global__ void __launch_bounds__(128) actor_entropy_vec4_kernel(const float* __restrict__ raw,
                                                                        const float* __restrict__ eps_e,
                                                                        bd_actor_cfg cfg, long long N,
                                                                        float* __restrict__ entropy,
                                                                        float* __restrict__ dent) {
  const long long n = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (n >= N) return;
  const int t = blockIdx.y;
  const long long r = (long long)t * N + n;
  const float kClamp = 0.99999997f;
  const float kXc = 0.5f * (logf(1.f + kClamp) - logf(1.f - kClamp));
  const float kLadjC = logf(1.f + kClamp) + logf(1.f - kClamp);
  const int J = cfg.entropy_samples;
  const float4 ra = *reinterpret_cast<const float4*>(raw + r * 2), rb = *reinterpret_cast<const float4*>(raw + r * 2 + 4);
  const EntRow p0 = ent_row(cfg, ra.x, ra.y), p1 = ent_row(cfg, ra.z, ra.w);
  const EntRow p2 = ent_row(cfg, rb.x, rb.y), p3 = ent_row(cfg, rb.z, rb.w);
  float lp[4] = {0.f, 0.f, 0.f, 0.f}, dm[4] = {0.f, 0.f, 0.f, 0.f}, ds[4] = {0.f, 0.f, 0.f, 0.f};
  const float* ee = eps_e + (long long)t * J * N + n;
#pragma unroll 2
  for (int j = 0; j < J; ++j) {
    const float4 e = *reinterpret_cast<const float4*>(ee + (long long)j * N);
    ent_sample(p0, e.x, kXc, kLadjC, lp[0], dm[0], ds[0]);
    ent_sample(p1, e.y, kXc, kLadjC, lp[1], dm[1], ds[1]);
    ent_sample(p2, e.z, kXc, kLadjC, lp[2], dm[2], ds[2]);
    ent_sample(p3, e.w, kXc, kLadjC, lp[3], dm[3], ds[3]);
  }
  const float fj = (float)J;      // (divisions, like the one-row kernel: the two must round alike)
  *reinterpret_cast<float4*>(entropy + r) = make_float4(-lp[0] / fj, -lp[1] / fj, -lp[2] / fj, -lp[3] / fj);
  *reinterpret_cast<float4*>(dent + r * 2) = make_float4(-dm[0] / fj, -ds[0] / fj, -dm[1] / fj, -ds[1] / fj);
  *reinterpret_cast<float4*>(dent + r * 2 + 4) = make_float4(-dm[2] / fj, -ds[2] / fj, -dm[3] / fj, -ds[3] / fj);
}

}  // namespace tc
}  // namespace bd
