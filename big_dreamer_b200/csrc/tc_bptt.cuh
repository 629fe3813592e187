// Tensor-core BPTT of the imagination rollout (the dgrad chain of Dreamer's actor loss through
// prior -> GRU -> embed, src/dreamer.py:363 over the graph built at :213-227).  Transition weights
// are frozen there (FreezeParameters, src/dreamer.py:313), so this kernel computes data gradients
// only: dL/d(action_t) -> dL/d(raw actor outputs) for every step (the actor's own backward is a
// batched MLP backward, tc_mlp_bwd.cuh) and optionally dL/d(prev_state), dL/d(prev_belief).
//
// Same engine as the forward (one CTA = 128 rows for all T steps, reverse time).  The forward pass
// saved, per step and 128-row tile, 16-bit images of
//     gate coefficients  c_r, c_z, c_n, c_nr, z   (5 planes x Be)   [d pre-gates per unit dL/db']
//     act'(x) of the embed output and act'(h) of the prior hidden layer
// so nothing is recomputed here.  Per step (t = T-1 .. 0), with G = dL/d b_{t+1} (total):
//     p0  d_pre2 = [d mu ; d raw_sigma] from (carry d s_{t+1}, upstream g_*)          -> D2 tile
//     p1  DH  = d_pre2 W_p2 ;            d_h = DH . act'(h)                            -> H tile
//     p2  ACC_B (+)= d_h W_p1            (ACC_B already holds  d_gh(t+1) W_hh )
//         gate stage A, slice 0:  G = ACC_B + g_beliefs[t]  -> scratch; slab <- d_gi; ACC_B[slice] <- z(t) G
//         (the carry z.G of the GRU's direct path is written back into ACC_B with tcgen05.st, so the W_hh
//          GEMMs of stage B accumulate onto it and the next step's ACC_B already contains it)
//     p3..  DX (+)= slab W_ih[slice]  |  next gate stage A / B (B re-reads G, slab <- d_gh)
//     ..    ACC_B' (+)= slab W_hh[slice]
//     p10 d_pre_x = DX . act'(x)                                                       -> H tile
//     p11 DSA = d_pre_x W_sa  -> d s_t (stays in TMEM for the next step's p0), d a_t -> d raw_t
// TMEM: ACC_B at column 0, DH / DX / DSA share the region at column 256.
#pragma once
#include "tc_engine.cuh"
#include "tc_mlp_bwd.cuh"

namespace bd {
namespace tc {

enum BpttEpi : uint8_t {
  EPI_P_DPRE2 = 1,
  EPI_P_MULSAVED = 2,   // D . saved act' image -> H tile   (aux0: 2 = prior hidden, 1 = embed x)
  EPI_P_GATE = 3,       // gate stage (aux1 = 0: pass A, 1: pass B); aux0 = first column of the slice
  EPI_P_DSA = 4,
  EPI_P_HEAD_DY = 5     // fused heads: d h_last = d out . w_out . act'(h_last) -> H tile (no MMA); aux0 = head
};

struct BpttArgs {
  Program prog;
  SmemPlan sm;
  const uint16_t* wpack;
  long long N;
  int T;
  long long* prof;
  int Be, S, A, Hi, Kb, Kh, Sp, Ksa;
  float min_std;
  bd_actor_cfg cfg;
  const uint16_t *sv_gate, *sv_xa, *sv_ha;
  const float *stds, *eps_s, *eps_a, *actions, *actor_raw, *dent;
  const float *g_beliefs, *g_states, *g_means, *g_stds, *g_entropy;
  const float* gbt;              // g_beliefs re-laid per (t, tile) as [col/4][row][4] (gb_tile_kernel), or null
  float* d_raw;
  float *d_prev_state, *d_prev_belief;
  float *scr_carry, *scr_gtot;   // per-CTA scratch, [gridDim.x][128][Kb] fp32
  const unsigned int* amax_bits;
  PrefetchPlan pf;
  // fused heads (imagine_and_returns): the lambda-return adjoint runs as a per-tile prologue, the heads'
  // dgrad chains as extra phases at the start of every step (their dX accumulates straight into ACC_B
  // and the carried d s in TMEM).  n_heads = 0: plain imagine BPTT.
  int n_heads, kh_hd;
  const uint16_t* sv_hd[2 * BD_MAX_LAYERS];   // act' images of the heads' hidden layers (forward-saved)
  const float* w_out[2];                      // last layer of each head: (1, hidden)
  int hd_last;                                // index of the last hidden layer (n_layers - 2)
  int hd_nvalid;                              // its width (columns of w_out)
  const float *g_returns, *g_reward, *g_value;   // upstream (T,N), each optional
  float lr_disc, lr_lam;
  float* scr_drv;                             // per-CTA [T][2][128]: d reward, d value of the tile's rows
};

// Upstream belief gradients (T,N,Be) row-major -> per (t, tile) images [col/4][row][4] (zero padded),
// the layout of the kernel's own fp32 scratch: with one thread per row, row-major reads touch 32
// different 128-byte lines per warp instruction; this layout makes them 512 contiguous bytes.
// Also folds in this tensor's contribution to the abs-max that sets the gradient scale.
static __global__ void gb_tile_kernel(const float* __restrict__ g, long long N, int Be, int Kb,
                                      long long ntiles, float* __restrict__ out, unsigned int* amax) {
  __shared__ float sm_[kTileRows][65];
  const long long tile = blockIdx.x;
  const int t = blockIdx.y, p = blockIdx.z;
  const long long row0 = tile * kTileRows;
  float m = 0.f;
  for (int i = threadIdx.x; i < kTileRows * 64; i += blockDim.x) {
    const int r = i >> 6, c = i & 63, col = p * 64 + c;
    const float v = (row0 + r < N && col < Be) ? g[((long long)t * N + row0 + r) * Be + col] : 0.f;
    sm_[r][c] = v;
    const float av = fabsf(v);
    if (av < 3.0e38f) m = fmaxf(m, av);      // ignore inf / nan, as absmax_kernel does
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(amax, __float_as_uint(m));
  __syncthreads();
  const int ngroups = min(16, (Kb - p * 64) >> 2);
  float* o = out + ((long long)t * ntiles + tile) * kTileRows * Kb;
  for (int i = threadIdx.x; i < ngroups * kTileRows; i += blockDim.x) {
    const int cg = i >> 7, r = i & 127;
    *reinterpret_cast<float4*>(o + ((size_t)(p * 16 + cg) * kTileRows + r) * 4) =
        make_float4(sm_[r][cg * 4], sm_[r][cg * 4 + 1], sm_[r][cg * 4 + 2], sm_[r][cg * 4 + 3]);
  }
}

// NPARTS = column parts of the epilogue (4 warps each, one per TMEM quadrant): 4 -> 18 warps, 5 on two of the
// SM's sub-partitions, i.e. 96 registers per thread; 3 -> 14 warps at 128 registers.
template <int FMT, bool PROF, int NPARTS>
__global__ void __launch_bounds__(64 + NPARTS * 128, 1) bptt_kernel(const __grid_constant__ BpttArgs A_) {
  static_assert(NPARTS == 3 || NPARTS == 4, "per-warp epilogue arrivals: engine_setup counts threads at 256");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const BpttArgs& a = A_;
  uint8_t* smem = smem_raw;
  __shared__ EngineShared sh;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform: the role code stays on the uniform datapath
  __shared__ Program sprog;
  stage_program(sprog, a.prog);
  // fused heads: the output layers' weight rows (every thread multiplies its row by all of them in
  // EPI_P_HEAD_DY) staged once per CTA; read back as broadcast 16-byte shared loads
  __shared__ __align__(16) float s_wout[2][256];
  if (a.n_heads) {
    for (int i = tid; i < 2 * 256; i += blockDim.x) {
      const int k = i >> 8, c = i & 255;
      s_wout[k][c] = (k < a.n_heads && a.w_out[k] && c < a.hd_nvalid) ? a.w_out[k][c] : 0.f;
    }
  }
  const uint32_t tmem_base = engine_setup(sh, a.sm.nstage, 1, NPARTS * 128);     // (per-warp arrivals; its __syncthreads also publishes s_wout)
  const long long ntiles = (a.N + kTileRows - 1) / kTileRows;
  const Program& P = sprog;

  if (warp == 0) {
    producer_role(P, a.sm, a.wpack, ntiles, a.T, smem, sh, &a.pf);
  } else if (warp == 1) {
    issuer_role<FMT, PROF>(P, a.sm, ntiles, a.T, smem, sh, tmem_base, a.prof);
  } else {
    // 16 epilogue warps: TMEM quadrant q = warp % 4, column part (warp - 2) / 4 in 0..3; chunks of 16 columns
    // (image products) or 8 (gate stages, d_pre2) keep the kernel inside the 112 registers 576 threads leave
    const int q = warp & 3, part = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    // one arrival per warp on the epilogue-completion barrier
    auto epi_arrive = [&](uint32_t ge) {
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh.epi_done[ge & 7]);
    };
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t rowoff = (row >> 3) * 128 + (row & 7) * 16;
    const int Be = a.Be, S = a.S, Ad = a.A, Kb = a.Kb;
    uint8_t* Ht = smem + a.sm.off_tile[TILE_H];
    uint8_t* D2 = smem + a.sm.off_tile[TILE_D2];
    // per-CTA fp32 scratch, stored as [col/4][row][4] so a warp's float4 accesses are contiguous;
    // SIDX(col) (col a multiple of 4) is this thread's float4 slot
    float* gtot = a.scr_gtot + (size_t)blockIdx.x * kTileRows * Kb + row * 4;
#define SIDX(col) ((size_t)((col) >> 2) * (kTileRows * 4))
    float inv_scale;
    const float scale = grad_scale(a.amax_bits, &inv_scale, a.n_heads ? 5 : 0);
    uint32_t Ge = 0, Gm = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const long long grow = tile * kTileRows + row;
      const bool rvalid = grow < a.N;
      if (a.n_heads && part == 0) {
        // adjoint of lambda_return with bootstrap = value[-1] (f32::lambda_return_bwd_kernel), forward in
        // time: G[t] = g_ret[t] + disc lam G[t-1]; d r[t] = G[t]; d v[t+1] += disc (1-lam) G[t];
        // d v[T-1] += disc G[T-1] (the bootstrap enters as next value and as the initial return)
        float* drv = a.scr_drv + (size_t)blockIdx.x * a.T * 2 * kTileRows + row;
        const long long lrow = rvalid ? grow : 0;
        float G = 0.f, dv_next = 0.f;
        for (int t = 0; t < a.T; ++t) {
          const long long o = (long long)t * a.N + lrow;
          G = (a.g_returns ? a.g_returns[o] : 0.f) + a.lr_disc * a.lr_lam * G;
          float dv = dv_next + (a.g_value ? a.g_value[o] : 0.f);
          if (t == a.T - 1) dv += a.lr_disc * G;
          drv[(size_t)(t * 2 + 0) * kTileRows] = rvalid ? G + (a.g_reward ? a.g_reward[o] : 0.f) : 0.f;
          drv[(size_t)(t * 2 + 1) * kTileRows] = rvalid ? dv : 0.f;
          dv_next = a.lr_disc * (1.f - a.lr_lam) * G;
        }
      }
      if (a.n_heads) asm volatile("bar.sync 1, %0;" ::"n"(NPARTS * 128) : "memory");   // every column part reads the row's d r / d v
      epi_arrive(Ge);   // nothing else to initialise per tile
      ++Ge;
      for (int i = 0; i < a.T; ++i) {
        const int t = a.T - 1 - i;
        const long long orow = (long long)t * a.N + grow;
        const size_t tl = (size_t)t * ntiles + tile;          // (t, tile) index of the saved images
        for (int pi = 0; pi < P.n_phases; ++pi) {
          const Phase ph = P.p[pi];
          long long e0 = 0, e1 = 0;
          if (PROF) e0 = clock64();
          // Gate stages after the first do not read their own phase's accumulator: they only need
          // the slab they overwrite to be free, i.e. the PREVIOUS phase's MMAs done (Kp_out = 1).
          const uint32_t Gw = Gm - ((ph.epi == EPI_P_GATE && ph.Kp_out) ? 1u : 0u);
#define BD_WAIT_ACC()                                         \
  do {                                                        \
    mbar_wait(&sh.acc_full[Gw & 3], (Gw >> 2) & 1);           \
    tc_fence_after_sync();                                    \
    if (PROF) e1 = clock64();                                 \
  } while (0)
          switch (ph.epi) {
            case EPI_P_DPRE2: {
              const int Sp = a.Sp;
              // branch-free clamped loads of everything that does not depend on TMEM; the first
              // chunk's loads are issued before the accumulator wait
              const long long lrow = rvalid ? orow : (long long)t * a.N;       // any valid row
              // With cs = carried d s_{t+1}:  d mu = cs + A,  d raw_sigma = cs C + D, where
              //   A = (g_s + g_mu) k,  C = eps f,  D = (g_s k eps + g_sigma k) f,  f = softplus'(raw) = 1 - e^-(sigma - min_std)
              // (k = gradient scale) -- three values per element instead of five live across the accumulator wait
              float A_[8], C_[8], D_[8];
              auto load_chunk = [&](int c) {
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                  float sd[2], ep[2], gs[2] = {0.f, 0.f}, gm[2] = {0.f, 0.f}, gd[2] = {0.f, 0.f};
                  if ((S & 1) == 0) {     // rows are 8-byte aligned: 64-bit loads, clamped inside the row
                    const long long o = lrow * S + min(c + j, S - 2);
                    const float2 x0 = *reinterpret_cast<const float2*>(a.stds + o);
                    const float2 x1 = *reinterpret_cast<const float2*>(a.eps_s + o);
                    sd[0] = x0.x; sd[1] = x0.y; ep[0] = x1.x; ep[1] = x1.y;
                    if (a.g_states) { const float2 y = *reinterpret_cast<const float2*>(a.g_states + o); gs[0] = y.x; gs[1] = y.y; }
                    if (a.g_means) { const float2 y = *reinterpret_cast<const float2*>(a.g_means + o); gm[0] = y.x; gm[1] = y.y; }
                    if (a.g_stds) { const float2 y = *reinterpret_cast<const float2*>(a.g_stds + o); gd[0] = y.x; gd[1] = y.y; }
                  } else {
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                      const long long o = lrow * S + min(c + j + k, S - 1);
                      sd[k] = a.stds[o]; ep[k] = a.eps_s[o];
                      if (a.g_states) gs[k] = a.g_states[o];
                      if (a.g_means) gm[k] = a.g_means[o];
                      if (a.g_stds) gd[k] = a.g_stds[o];
                    }
                  }
#pragma unroll
                  for (int k = 0; k < 2; ++k) {
                    const float f = 1.f - fast_exp(-(sd[k] - a.min_std));
                    const float gss = gs[k] * scale;
                    A_[j + k] = gss + gm[k] * scale;
                    C_[j + k] = ep[k] * f;
                    D_[j + k] = (gss * ep[k] + gd[k] * scale) * f;
                  }
                }
              };
              if (part * 8 < Sp) load_chunk(part * 8);
              BD_WAIT_ACC();
              for (int c = part * 8; c < Sp; c += NPARTS * 8) {
                if (c != part * 8) load_chunk(c);
                float cs[8], m_[8], s_[8];
                if (i > 0 || a.n_heads) {
                  tmem_ld8(trow + 256 + c, cs);     // d s_{t+1} left by the previous step's DSA (+ the heads' d s_t)
                  tmem_ld_wait();
                } else {
#pragma unroll
                  for (int j = 0; j < 8; ++j) cs[j] = 0.f;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const bool ok = (c + j < S) && rvalid;
                  m_[j] = ok ? cs[j] + A_[j] : 0.f;
                  s_[j] = ok ? fmaf(cs[j], C_[j], D_[j]) : 0.f;
                }
                store8<FMT>(D2 + (c >> 3) * kLboA + rowoff, m_);
                store8<FMT>(D2 + ((Sp + c) >> 3) * kLboA + rowoff, s_);
              }
            } break;
            case EPI_P_MULSAVED: {
              const int kp = ph.aux0 == 1 ? Kb : (ph.aux0 == 2 ? a.Kh : a.kh_hd);
              const uint16_t* img = (ph.aux0 == 1 ? a.sv_xa : (ph.aux0 == 2 ? a.sv_ha : a.sv_hd[ph.aux0 - 16])) +
                                    tl * kTileRows * kp + row * 8;
              const uint32_t tacc = trow + ph.d_col;
              // 16-column chunks c = 16 (part + NPARTS k) (kp is a multiple of 16); the image pieces of two or three
              // chunks are in flight at any time, the first ones requested before the accumulator wait (one chunk
              // ahead left a full L2 / HBM round trip exposed per chunk: ~5 K cycles per phase)
              // (a rotating set of named registers: an array indexed by the chunk number ends up in local memory)
              const int cstep = NPARTS * 16;
              auto ld0 = [&](int c) { return c < kp ? *reinterpret_cast<const uint4*>(img + (size_t)(c >> 3) * kTileRows * 8) : make_uint4(0, 0, 0, 0); };
              auto ld1 = [&](int c) { return c < kp ? *reinterpret_cast<const uint4*>(img + (size_t)((c >> 3) + 1) * kTileRows * 8) : make_uint4(0, 0, 0, 0); };
              const int cfirst = part * 16;
              uint4 a0 = ld0(cfirst), a1 = ld1(cfirst);
              uint4 b0 = ld0(cfirst + cstep), b1 = ld1(cfirst + cstep);
              uint4 d0 = make_uint4(0, 0, 0, 0), d1 = d0;
              if (NPARTS == 3) { d0 = ld0(cfirst + 2 * cstep); d1 = ld1(cfirst + 2 * cstep); }
              BD_WAIT_ACC();
              for (int c = cfirst; c < kp; c += cstep) {
                float v[16];
                tmem_ld16(tacc + c, v);
                const uint4 hu[2] = {a0, a1};
                a0 = b0; a1 = b1;
                if (NPARTS == 3) {
                  b0 = d0; b1 = d1;
                  d0 = ld0(c + 3 * cstep); d1 = ld1(c + 3 * cstep);
                } else {
                  b0 = ld0(c + 2 * cstep); b1 = ld1(c + 2 * cstep);
                }
                tmem_ld_wait();
#pragma unroll
                for (int g8 = 0; g8 < 2; ++g8) {
                  float h[8];
                  unpack8<FMT>(hu[g8], h);
#pragma unroll
                  for (int j = 0; j < 8; ++j) v[g8 * 8 + j] *= h[j];
                  store8<FMT>(Ht + ((c >> 3) + g8) * kLboA + rowoff, v + g8 * 8);
                }
              }
            } break;
            case EPI_P_GATE: {
              // 8-column chunks: c = 8 (part + NPARTS it) (a slice is at most 64 columns wide)
              constexpr int kIts = (8 + NPARTS - 1) / NPARTS;
              const int n0 = ph.aux0, Ns = ph.Np;
              const bool passB = ph.pad != 0;
              uint8_t* slab = smem + a.sm.off_tile[ph.out_tile];
              const uint16_t* gimg = a.sv_gate + tl * 5 * kTileRows * Kb + row * 8;
              const size_t plane = (size_t)kTileRows * Kb;
              // pass A: the upstream belief gradient of this thread's columns (both chunks), requested
              // raw before the accumulator wait; masked / scaled where it is added.  Nothing to fetch when the
              // loss reaches the beliefs only through this kernel's own heads (fused step: g_beliefs == null).
              const bool have_gb = !passB && (a.g_beliefs != nullptr);
              float4 gb4[2];
              auto load_gb = [&](int it) {
                const float* gbrow = !a.gbt ? a.g_beliefs + (rvalid ? orow : (long long)t * a.N) * Be : nullptr;
                const float* gbt = a.gbt ? a.gbt + tl * kTileRows * Kb + row * 4 : nullptr;
                const bool vec = ((Be & 3) == 0);
                const int c = (part + NPARTS * it) * 8;
                const int col0 = n0 + min(c, Ns - 8);
#pragma unroll
                for (int j4 = 0; j4 < 2; ++j4) {
                  const int cb = col0 + j4 * 4;
                  if (gbt) {
                    gb4[j4] = *reinterpret_cast<const float4*>(gbt + SIDX(cb));     // zero padded
                  } else if (vec) {
                    gb4[j4] = *reinterpret_cast<const float4*>(gbrow + min(cb, Be - 4));
                  } else {
                    gb4[j4].x = gbrow[min(cb, Be - 1)]; gb4[j4].y = gbrow[min(cb + 1, Be - 1)];
                    gb4[j4].z = gbrow[min(cb + 2, Be - 1)]; gb4[j4].w = gbrow[min(cb + 3, Be - 1)];
                  }
                }
              };
              long long q0 = 0, q1 = 0, q2 = 0;
              if (PROF) q0 = clock64();
              if (have_gb) load_gb(0);
              if (PROF) q1 = clock64();
              const int pl3 = passB ? 3 : 2;
              uint4 cf[4];
              auto load_planes = [&](int c) {
                const uint16_t* gp = gimg + (size_t)((n0 + c) >> 3) * kTileRows * 8;
                cf[0] = *reinterpret_cast<const uint4*>(gp + 0 * plane);
                cf[1] = *reinterpret_cast<const uint4*>(gp + 1 * plane);
                cf[2] = *reinterpret_cast<const uint4*>(gp + pl3 * plane);
                if (!passB) cf[3] = *reinterpret_cast<const uint4*>(gp + 4 * plane);
              };
              float Gb[8];
              if (part * 8 < Ns) {
                load_planes(part * 8);
                if (passB) {
#pragma unroll
                  for (int j4 = 0; j4 < 2; ++j4) {
                    const float4 g4 = *reinterpret_cast<const float4*>(gtot + SIDX(n0 + part * 8 + j4 * 4));
                    Gb[j4 * 4] = g4.x; Gb[j4 * 4 + 1] = g4.y; Gb[j4 * 4 + 2] = g4.z; Gb[j4 * 4 + 3] = g4.w;
                  }
                }
              }
              if (PROF) q2 = clock64();
              BD_WAIT_ACC();
              if (PROF && blockIdx.x == 0 && lane == 0 && warp == 2) {
                prof_add(&a.prof[(20 + pi) * 8 + 0], q1 - q0);    // upstream-gradient / carry fetch + combine
                prof_add(&a.prof[(20 + pi) * 8 + 1], q2 - q1);    // coefficient-plane loads issued
                prof_add(&a.prof[(20 + pi) * 8 + 2], e1 - q2);    // accumulator wait proper
              }
#pragma unroll
              for (int it = 0; it < kIts; ++it) {
                const int c = (part + NPARTS * it) * 8;
                if (c < Ns) {
                  const int col0 = n0 + c;
                  float G[8];
                  if (it > 0) {
                    load_planes(c);
                    if (have_gb) load_gb(it);
                  }
                  if (!passB) {
                    tmem_ld8(trow + col0, G);           // ACC_B = d_gh(t+1) W_hh + d_h W_p1
                    tmem_ld_wait();
                    if (have_gb) {
                      const float sc = rvalid ? scale : 0.f;
#pragma unroll
                      for (int j4 = 0; j4 < 2; ++j4) {
                        const int cb = col0 + j4 * 4;       // (columns past Be: padding of the last slice)
                        const float4 gb = gb4[j4];
                        G[j4 * 4] += cb < Be ? gb.x * sc : 0.f; G[j4 * 4 + 1] += cb + 1 < Be ? gb.y * sc : 0.f;
                        G[j4 * 4 + 2] += cb + 2 < Be ? gb.z * sc : 0.f; G[j4 * 4 + 3] += cb + 3 < Be ? gb.w * sc : 0.f;
                      }
                    }
#pragma unroll
                    for (int j4 = 0; j4 < 2; ++j4) {
                      *reinterpret_cast<float4*>(gtot + SIDX(col0 + j4 * 4)) =
                          make_float4(G[j4 * 4], G[j4 * 4 + 1], G[j4 * 4 + 2], G[j4 * 4 + 3]);
                    }
                  } else if (it == 0) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) G[j] = Gb[j];
                  } else {
#pragma unroll
                    for (int j4 = 0; j4 < 2; ++j4) {
                      const float4 g4 = *reinterpret_cast<const float4*>(gtot + SIDX(col0 + j4 * 4));
                      G[j4 * 4] = g4.x; G[j4 * 4 + 1] = g4.y; G[j4 * 4 + 2] = g4.z; G[j4 * 4 + 3] = g4.w;
                    }
                  }
#pragma unroll
                  for (int gt = 0; gt < 3; ++gt) {
                    float o[8], c8[8];
                    unpack8<FMT>(cf[gt], c8);
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[j] = G[j] * c8[j];
                    store8<FMT>(slab + ((gt * Ns + c) >> 3) * kLboA + rowoff, o);
                  }
                  if (!passB) {
                    // carry of the GRU's direct path for the next (earlier) step, z . G, back into this slice's
                    // ACC_B columns: the W_hh GEMMs of stage B accumulate onto it
                    float zf[8];
                    uint32_t cz[8];
                    unpack8<FMT>(cf[3], zf);
#pragma unroll
                    for (int j = 0; j < 8; ++j) cz[j] = __float_as_uint(G[j] * zf[j]);
                    tmem_st8(trow + col0, cz);
                  }
                }
              }
              if (!passB) tmem_st_wait();
            } break;
            case EPI_P_HEAD_DY: {
              // d h_last[row, c] = d out[t, row] . w_out[c] . act'(h_last)[row, c]   (the scalar output layer's
              // dgrad is a rank-1 product: no MMA); operands carry the gradient scale
              const int k = ph.aux0, kp = a.kh_hd;
              const uint16_t* img = a.sv_hd[k * BD_MAX_LAYERS + a.hd_last] + tl * kTileRows * kp + row * 8;
              const float* wo = s_wout[k];
              const float dout = a.scr_drv[((size_t)blockIdx.x * a.T + t) * 2 * kTileRows + (size_t)k * kTileRows + row] * scale;
              const int cstep = NPARTS * 16;
              auto ld0 = [&](int c) { return c < kp ? *reinterpret_cast<const uint4*>(img + (size_t)(c >> 3) * kTileRows * 8) : make_uint4(0, 0, 0, 0); };
              auto ld1 = [&](int c) { return c < kp ? *reinterpret_cast<const uint4*>(img + (size_t)((c >> 3) + 1) * kTileRows * 8) : make_uint4(0, 0, 0, 0); };
              const int cfirst = part * 16;
              uint4 a0 = ld0(cfirst), a1 = ld1(cfirst);
              uint4 b0 = ld0(cfirst + cstep), b1 = ld1(cfirst + cstep);
              uint4 d0 = make_uint4(0, 0, 0, 0), d1 = d0;
              if (NPARTS == 3) { d0 = ld0(cfirst + 2 * cstep); d1 = ld1(cfirst + 2 * cstep); }
              BD_WAIT_ACC();             // the previous phase's MMAs (which read the H tile) are done
              for (int c = cfirst; c < kp; c += cstep) {
                const uint4 hu[2] = {a0, a1};
                a0 = b0; a1 = b1;
                if (NPARTS == 3) {
                  b0 = d0; b1 = d1;
                  d0 = ld0(c + 3 * cstep); d1 = ld1(c + 3 * cstep);
                } else {
                  b0 = ld0(c + 2 * cstep); b1 = ld1(c + 2 * cstep);
                }
#pragma unroll
                for (int g8 = 0; g8 < 2; ++g8) {
                  float h[8], o[8];
                  unpack8<FMT>(hu[g8], h);
                  const float4 w0 = *reinterpret_cast<const float4*>(wo + c + g8 * 8);      // (zero past n_valid)
                  const float4 w1 = *reinterpret_cast<const float4*>(wo + c + g8 * 8 + 4);
                  const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                  for (int j = 0; j < 8; ++j) o[j] = dout * wv[j] * h[j];
                  store8<FMT>(Ht + ((c >> 3) + g8) * kLboA + rowoff, o);
                }
              }
            } break;
            case EPI_P_DSA: {
              // A = 1 (the common case): the per-row inputs of the action gradient are loaded before
              // the accumulator wait; larger action spaces load them in the loop below
              float pm = 0.f, ps = 0.f, pa = 0.f, pdm = 0.f, pds = 0.f, pe = 0.f, pge = 0.f;
              const bool pre = part == 0 && rvalid && Ad == 1;
              if (pre) {
                pge = a.g_entropy ? a.g_entropy[orow] : 0.f;
                pm = a.actor_raw[orow * 2]; ps = a.actor_raw[orow * 2 + 1];
                pa = a.actions[orow]; pe = a.eps_a[orow];
                pdm = a.dent[orow * 2]; pds = a.dent[orow * 2 + 1];
              }
              BD_WAIT_ACC();
              const uint32_t tacc = trow + 256;
              if (part == 0) {
                // d a_t sits at accumulator columns [S, S + A): at most two 16-column pieces (A <= 16)
                const int c0 = (S >> 4) << 4;
                float vv[16], vw[16];
                tmem_ld16(tacc + c0, vv);
                const bool two = (S + Ad > c0 + 16) && (c0 + 16 < a.Ksa);      // warp-uniform
                if (two) tmem_ld16(tacc + c0 + 16, vw);
                tmem_ld_wait();
                if (rvalid) {
                  const float ge = (pre ? pge : (a.g_entropy ? a.g_entropy[orow] : 0.f)) * scale;
                  for (int j = 0; j < Ad; ++j) {
                    const int idx = S - c0 + j;
                    float da = 0.f;
#pragma unroll
                    for (int k = 0; k < 16; ++k) if (k == idx) da = vv[k];
                    if (two) {
#pragma unroll
                      for (int k = 0; k < 16; ++k) if (k + 16 == idx) da = vw[k];
                    }
                    const long long oa = orow * Ad + j;
                    const float m_raw = pre ? pm : a.actor_raw[orow * 2 * Ad + j];
                    const float s_raw = pre ? ps : a.actor_raw[orow * 2 * Ad + Ad + j];
                    const float act = pre ? pa : a.actions[oa];
                    const float dy = da * (1.f - act * act);
                    const float dmean = dy + ge * (pre ? pdm : a.dent[orow * 2 * Ad + j]);
                    const float dsd = dy * (pre ? pe : a.eps_a[oa]) + ge * (pre ? pds : a.dent[orow * 2 * Ad + Ad + j]);
                    const float th = tanhf(m_raw / a.cfg.mean_scale);
                    a.d_raw[orow * 2 * Ad + j] = dmean * (1.f - th * th) * inv_scale;
                    a.d_raw[orow * 2 * Ad + Ad + j] = dsd * softplus_gradf_(s_raw + a.cfg.raw_init_std) * inv_scale;
                  }
                }
              }
              if (t == 0) {   // gradients wrt the start latents
                if (a.d_prev_state) {
                  for (int c = part * 16; c < a.Sp; c += NPARTS * 16) {
                    float v[16];
                    tmem_ld16(tacc + c, v);            // warp-collective: never under a per-lane branch
                    tmem_ld_wait();
                    if (rvalid) {
#pragma unroll
                      for (int j = 0; j < 16; ++j) if (c + j < S) a.d_prev_state[grow * S + c + j] = v[j] * inv_scale;
                    }
                  }
                }
                if (a.d_prev_belief) {
                  for (int c = part * 16; c < Kb; c += NPARTS * 16) {
                    float v[16];
                    tmem_ld16(trow + c, v);            // ACC_B = d_gh(0) W_hh + z(0) G(0)
                    tmem_ld_wait();
                    if (rvalid) {
#pragma unroll
                      for (int j = 0; j < 16; ++j)
                        if (c + j < Be) a.d_prev_belief[grow * Be + c + j] = v[j] * inv_scale;
                    }
                  }
                }
              }
            } break;
            default: BD_WAIT_ACC(); break;
          }
#undef BD_WAIT_ACC
          tc_fence_before_sync();
          epi_arrive(Ge);
          if (PROF && blockIdx.x == 0 && lane == 0 && (warp == 2 || warp == 6)) {
            const int o = pi * 8 + (warp == 2 ? 3 : 5);
            prof_add(&a.prof[o], e1 - e0);
            prof_add(&a.prof[o + 1], clock64() - e1);
          }
          ++Ge;
          ++Gm;
        }
      }
    }
  }
#undef SIDX
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

}  // namespace tc
}  // namespace bd
