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
//     w_full[s] / w_empty[s]   producer <-> issuer, one ring stage = one K-block of 32
//     acc_full[Gm & 3]         issuer -> epilogue, tcgen05.commit after the last GEMM of phase Gm
//     epi_done[Ge & 3]         epilogue -> issuer, 256 arrivals when epilogue Ge is done
// A phase names how far back its dependency is (dep_back = 1: previous phase; 2: the one before,
// which lets independent phases -- the GRU's N-slices -- overlap MMA with the previous epilogue
// using the two TMEM halves).
#pragma once
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_pack.cuh"

namespace bd {
namespace tc {

constexpr int kTileRows = 128;
constexpr int kThreads = 320;
constexpr int kEpiThreads = 256;
constexpr int kMaxGemms = 64, kMaxPhases = 40;
constexpr uint32_t kLboA = kTileRows * 16;   // bytes between 8-column groups of an activation tile

enum TileId : uint8_t { TILE_BCUR = 0, TILE_BNXT = 1, TILE_SA = 2, TILE_H = 3, TILE_H2 = 4 };
enum EpiKind : uint8_t {
  EPI_ACT_H = 1,      // act(D) -> H tile
  EPI_ACTOR_OUT = 2,  // action sample + entropy
  EPI_GRU = 3,        // GRU gates for one N-slice -> b'
  EPI_PRIOR_OUT = 4,  // mean / std / sampled state
  EPI_HEAD_OUT = 5    // scalar head output (reward / value)
};

struct Gemm {
  uint32_t w_off;     // element offset of the packed weight image (rows Np, cols Kp)
  uint16_t Np, Kp;    // MMA N (mult of 16, <= 256), K (mult of 16)
  uint16_t a_k0;      // first K column inside the A tile (mult of 8)
  uint16_t d_col;     // TMEM column of the accumulator
  uint8_t a_tile;     // TileId
  uint8_t accumulate; // 1: continue a running sum in D
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
};
struct Program {
  int n_gemms, n_phases;
  Gemm g[kMaxGemms];
  Phase p[kMaxPhases];
};

struct SmemPlan {
  uint32_t off_tile[5];   // B0, B1, SA, H, H2 (byte offsets from the 1024-aligned base)
  uint32_t off_ring, stage_bytes, nstage;
  uint32_t total;
};

struct RolloutArgs {
  Program prog;
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
template <int FMT, bool WITH_ACTOR>
__global__ void __launch_bounds__(kThreads, 1) rollout_fwd_kernel(const __grid_constant__ RolloutArgs A_) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const RolloutArgs& a = A_;
  uint8_t* smem = smem_raw;
  __shared__ uint64_t w_full[8], w_empty[8], acc_full[4], epi_done[4];
  __shared__ uint32_t tmem_holder;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nstage = a.sm.nstage;
  if (tid == 0) {
    for (int i = 0; i < nstage; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&epi_done[i], kEpiThreads); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(&tmem_holder);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_holder;

  const long long ntiles = (a.N + kTileRows - 1) / kTileRows;
  const Program& P = a.prog;
  uint8_t* ring = smem + a.sm.off_ring;

  if (warp == 0) {
    // =========================================================== weight producer
    if (lane == 0) {
      uint32_t cnt = 0;
      for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
        for (int t = 0; t < a.T; ++t)
          for (int gi = 0; gi < P.n_gemms; ++gi) {
            const Gemm g = P.g[gi];
            for (int k0 = 0; k0 < g.Kp; k0 += 32) {
              const int kc = min(32, g.Kp - k0);
              const uint32_t bytes = (uint32_t)g.Np * kc * 2;
              const uint32_t st = cnt % nstage, ph = (cnt / nstage) & 1;
              mbar_wait(&w_empty[st], ph ^ 1);
              mbar_expect_tx(&w_full[st], bytes);
              tma_bulk_g2s(ring + st * a.sm.stage_bytes, a.wpack + g.w_off + (size_t)k0 * g.Np, bytes,
                           &w_full[st]);
              ++cnt;
            }
          }
    }
  } else if (warp == 1) {
    // =========================================================== MMA issuer
    if (lane == 0) {
      // Ge counts epilogue completions (incl. the per-tile init pseudo-phase, which has no MMAs);
      // Gm counts phases with MMAs.  Each indexes its own barrier ring so generations stay in step.
      uint32_t cnt = 0, Ge = 0, Gm = 0;
      for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        ++Ge;  // the tile-initialisation pseudo-phase (epilogue only)
        for (int t = 0; t < a.T; ++t) {
          const int par = t & 1;
          for (int pi = 0; pi < P.n_phases; ++pi) {
            const Phase ph = P.p[pi];
            {
              const uint32_t D = Ge - ph.dep_back;
              mbar_wait(&epi_done[D & 3], (D >> 2) & 1);
              tc_fence_after_sync();
            }
            for (int gi = ph.g0; gi < ph.g0 + ph.ng; ++gi) {
              const Gemm g = P.g[gi];
              uint32_t tile_id = g.a_tile;
              if (tile_id < 2) tile_id ^= par;
              const uint32_t a_base = smem_u32(smem + a.sm.off_tile[tile_id]) + (g.a_k0 >> 3) * kLboA;
              const uint32_t idesc = make_idesc_f16(FMT, kTileRows, g.Np);
              const uint32_t lbo_b = (uint32_t)g.Np * 16;
              for (int k0 = 0; k0 < g.Kp; k0 += 32) {
                const int kc = min(32, g.Kp - k0);
                const uint32_t st = cnt % nstage, wph = (cnt / nstage) & 1;
                mbar_wait(&w_full[st], wph);
                tc_fence_after_sync();
                const uint32_t b_base = smem_u32(ring + st * a.sm.stage_bytes);
                for (int ks = 0; ks < kc; ks += 16) {
                  const uint64_t ad = make_smem_desc(a_base + ((k0 + ks) >> 3) * kLboA, kLboA, 128);
                  const uint64_t bd_ = make_smem_desc(b_base + (ks >> 3) * lbo_b, lbo_b, 128);
                  umma_f16(tmem_base + g.d_col, ad, bd_, idesc, (g.accumulate | (k0 + ks)) ? 1u : 0u);
                }
                umma_commit(&w_empty[st]);
                ++cnt;
              }
            }
            umma_commit(&acc_full[Gm & 3]);
            ++Gm;
            ++Ge;
          }
        }
      }
    }
  } else {
    // =========================================================== epilogue warps
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const int etid = tid - 64;
    const int Be = a.Be, S = a.S, Ad = a.A;
    uint32_t Ge = 0, Gm = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const long long grow = tile * kTileRows + row;
      const bool rvalid = grow < a.N;
      // ---------------- tile initialisation: B0 <- prev_belief | 1, B1 <- 0 | 1, SA <- prev_state | . | 1
      {
        uint8_t* B0 = smem + a.sm.off_tile[0];
        uint8_t* B1 = smem + a.sm.off_tile[1];
        uint8_t* SA = smem + a.sm.off_tile[2];
        uint8_t* H = smem + a.sm.off_tile[3];
        for (int i = etid; i < kTileRows * a.Kp_b; i += kEpiThreads) {
          const int r = i / a.Kp_b, k = i - r * a.Kp_b;
          const long long gr = tile * kTileRows + r;
          float v = (k < Be) ? ((gr < a.N) ? a.prev_belief[gr * Be + k] : 0.f) : (k == Be ? 1.f : 0.f);
          store1<FMT>(B0, r, k, v);
          store1<FMT>(B1, r, k, k == Be ? 1.f : 0.f);
        }
        for (int i = etid; i < kTileRows * a.Kp_sa; i += kEpiThreads) {
          const int r = i / a.Kp_sa, k = i - r * a.Kp_sa;
          const long long gr = tile * kTileRows + r;
          float v = (k < S) ? ((gr < a.N) ? a.prev_state[gr * S + k] : 0.f) : (k == S + Ad ? 1.f : 0.f);
          store1<FMT>(SA, r, k, v);
        }
        for (int i = etid; i < kTileRows * a.Kp_h; i += kEpiThreads) {
          const int r = i / a.Kp_h, k = i - r * a.Kp_h;
          store1<FMT>(H, r, k, 0.f);
        }
        fence_proxy_async_smem();
        mbar_arrive(&epi_done[Ge & 3]);
        ++Ge;
      }
      for (int t = 0; t < a.T; ++t) {
        const int par = t & 1;
        uint8_t* Bnxt = smem + a.sm.off_tile[1 ^ par];
        uint8_t* SAt = smem + a.sm.off_tile[2];
        const long long orow = (long long)t * a.N + grow;        // row in (T,N,.) outputs
        for (int pi = 0; pi < P.n_phases; ++pi) {
          const Phase ph = P.p[pi];
          mbar_wait(&acc_full[Gm & 3], (Gm >> 2) & 1);
          tc_fence_after_sync();
          const uint32_t tacc = trow + ph.d_col;
          switch (ph.epi) {
            case EPI_ACT_H: {
              uint8_t* out = smem + a.sm.off_tile[ph.out_tile];
              for (int c = half * 16; c < ph.Kp_out; c += 32) {
                float v[16];
                if (c < ph.Np) {
                  tmem_ld16(tacc + c, v);
                  tmem_ld_wait();
                } else {
#pragma unroll
                  for (int j = 0; j < 16; ++j) v[j] = 0.f;
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const int col = c + j;
                  v[j] = col < ph.n_valid ? tc_act(a.act, v[j]) : (col == ph.n_valid ? 1.f : 0.f);
                }
                store16<FMT>(out, row, c, v);
              }
            } break;
            case EPI_GRU: {
              const int n0 = ph.aux0, Ns = ph.Np;
              const float* bold = (t == 0) ? a.prev_belief + grow * Be
                                           : a.beliefs + ((long long)(t - 1) * a.N + grow) * Be;
              float* bnew = a.beliefs + orow * Be;
              for (int c = half * 16; c < Ns; c += 32) {
                float r_[16], z_[16], in_[16], hn_[16], o[16];
                tmem_ld16(tacc + c, r_);
                tmem_ld16(tacc + Ns + c, z_);
                tmem_ld16(tacc + 2 * Ns + c, in_);
                tmem_ld16(tacc + 3 * Ns + c, hn_);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const int col = n0 + c + j;
                  float val = (col == Be) ? 1.f : 0.f;
                  if (col < Be) {
                    const float h = rvalid ? bold[col] : 0.f;
                    const float r = fast_sigmoid(r_[j]);
                    const float z = fast_sigmoid(z_[j]);
                    const float n = fast_tanh(in_[j] + r * hn_[j]);
                    val = (1.f - z) * n + z * h;
                    if (rvalid) bnew[col] = val;
                  }
                  o[j] = val;
                }
                if (n0 + c < a.Kp_b) store16<FMT>(Bnxt, row, n0 + c, o);
              }
            } break;
            case EPI_PRIOR_OUT: {
              if (half == 0) {
                const int Sp = ph.Np;    // mean at [0,Sp), raw std at [Sp, 2Sp)
                for (int c = 0; c < Sp; c += 16) {
                  float m_[16], s_[16];
                  tmem_ld16(tacc + c, m_);
                  tmem_ld16(tacc + Sp + c, s_);
                  tmem_ld_wait();
#pragma unroll
                  for (int j = 0; j < 16; ++j) {
                    const int col = c + j;
                    if (col < S) {
                      const float sd = softplusf_(s_[j]) + a.min_std;
                      const float e = rvalid ? a.eps_s[orow * S + col] : 0.f;
                      const float st = m_[j] + sd * e;
                      if (rvalid) {
                        a.means[orow * S + col] = m_[j];
                        a.stds[orow * S + col] = sd;
                        a.states[orow * S + col] = st;
                      }
                      store1<FMT>(SAt, row, col, st);
                    }
                  }
                }
              }
            } break;
            case EPI_ACTOR_OUT: {
              if (WITH_ACTOR && half == 0) {
                const int Ap = ph.Np;
                float m_[16], s_[16];
                tmem_ld16(tacc, m_);
                tmem_ld16(tacc + Ap, s_);
                tmem_ld_wait();
                const float kClamp = 0.99999997f, kLogSqrt2Pi = 0.9189385332046727f,
                            kLog2 = 0.6931471805599453f;
                const int J = a.J;
                float ent_acc = 0.f;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  if (j < Ad) {
                    const float mean = a.cfg.mean_scale * tanhf(m_[j] / a.cfg.mean_scale);
                    const float sd = softplusf_(s_[j] + a.cfg.raw_init_std) + a.cfg.min_std;
                    const float ea = rvalid ? a.eps_a[orow * Ad + j] : 0.f;
                    const float act = tanhf(mean + ea * sd);
                    store1<FMT>(SAt, row, S + j, act);
                    float lp_sum = 0.f, dm_sum = 0.f, ds_sum = 0.f;
                    if (rvalid) {
                      a.actions[orow * Ad + j] = act;
                      a.actor_raw[orow * 2 * Ad + j] = m_[j];
                      a.actor_raw[orow * 2 * Ad + Ad + j] = s_[j];
                      const float var2 = 2.f * sd * sd, log_sd = logf(sd), inv_var = 1.f / (sd * sd);
                      const float* ee = a.eps_e + ((long long)t * J * a.N + grow) * Ad + j;
                      for (int s = 0; s < J; ++s) {
                        const float e = ee[(long long)s * a.N * Ad];
                        const float y = tanhf(mean + e * sd);
                        const float yc = fminf(fmaxf(y, -kClamp), kClamp);
                        const float gate = (yc == y) ? 1.f : 0.f;
                        const float xh = 0.5f * logf((1.f + yc) / (1.f - yc));
                        const float d = xh - mean;
                        lp_sum += -(d * d) / var2 - log_sd - kLogSqrt2Pi -
                                  2.f * (kLog2 - xh - softplusf_(-2.f * xh));
                        const float dlp = -d * inv_var + 2.f * tanhf(xh);
                        dm_sum += d * inv_var + gate * dlp;
                        ds_sum += d * d * inv_var / sd - 1.f / sd + gate * e * dlp;
                      }
                      a.dent[orow * 2 * Ad + j] = -dm_sum / (float)J;
                      a.dent[orow * 2 * Ad + Ad + j] = -ds_sum / (float)J;
                    }
                    ent_acc += lp_sum;
                  }
                }
                if (rvalid) a.entropy[orow] = -ent_acc / (float)J;
              }
            } break;
            case EPI_HEAD_OUT: {
              if (half == 0) {
                float v[16];
                tmem_ld16(tacc, v);
                tmem_ld_wait();
                if (rvalid && a.head_out[ph.aux0]) a.head_out[ph.aux0][orow] = v[0];
              }
            } break;
            default: break;
          }
          tc_fence_before_sync();
          fence_proxy_async_smem();
          mbar_arrive(&epi_done[Ge & 3]);
          ++Ge;
          ++Gm;
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

}  // namespace tc
}  // namespace bd
