// Host side of the tensor-core rollout: builds the per-step PROGRAM (GEMM / phase tables), the
// weight-packing jobs and the shared-memory plan from the model dimensions, then launches the
// pack kernel and the persistent rollout kernel (tc_engine.cuh).
#include "api_internal.h"
#include "tc_engine.cuh"
#include "tc_builder.cuh"
#include "tc_bptt.cuh"
#include <stdlib.h>

namespace bd {
namespace tc {

// (+ the actor's hidden activations, up to kSavedActorLayers images of (T x tiles) x 128 x Kact: the batched actor
// backward reads them instead of recomputing the actor's forward pass)
constexpr int kSavedActorLayers = 4;
// (+ the actor's layer-0 input images x0b = [b_t | 1], x0s = s_t per (t, tile), copied out of the rollout's operand
// tiles: see X0Save)
struct SavedLayout { size_t off_gate, off_xa, off_ha, off_act, act_bytes, off_x0b, off_x0s, total; int Kb, Kh, Kact, Kx0b, Kx0s; };
static SavedLayout saved_layout(const bd_rssm& r, int T, long long N) {
  SavedLayout s;
  s.Kb = r16(r.belief_size); s.Kh = r16(r.hidden_size);
  const size_t tiles = (size_t)((N + kTileRows - 1) / kTileRows) * T;
  s.off_gate = 0;
  s.off_xa = s.off_gate + tiles * 5 * kTileRows * s.Kb * 2;
  s.off_ha = s.off_xa + tiles * kTileRows * s.Kb * 2;
  s.off_act = s.off_ha + tiles * kTileRows * s.Kh * 2;
  s.Kact = r16(r.hidden_size + 1);
  s.act_bytes = tiles * kTileRows * s.Kact * 2;          // one hidden layer
  s.Kx0b = r16(r.belief_size + 1); s.Kx0s = r16(r.state_size);
  s.off_x0b = s.off_act + kSavedActorLayers * s.act_bytes;
  s.off_x0s = s.off_x0b + tiles * kTileRows * s.Kx0b * 2;
  s.total = s.off_x0s + tiles * kTileRows * s.Kx0s * 2;
  return s;
}
size_t imagine_saved_bytes(const bd_rssm& r, int T, long long N) { return saved_layout(r, T, N).total + 256; }
// the actor's saved hidden images inside tc_saved (layout of mlp_backward's `saved`: layer after layer), or null
// when the actor has more hidden layers than the buffer holds / the debug switch BD_NO_ACTOR_SAVE is set
const void* imagine_saved_actor(const bd_rssm& r, const bd_mlp& actor, int T, long long N, const void* tc_saved) {
  // (small row counts run latency-bound in column-split clusters, where the extra stores cost the rollout part of
  // what the backward saves -- measured at 2 500 rows with the layer-0 images also coming from the rollout:
  // rollout +0.03 ms, actor backward -0.05 ms)
  const char* e = getenv("BD_ACTOR_SAVE_MIN_ROWS");      // (read per call: the tests switch it)
  const long long min_rows = e ? atoll(e) : 2048;
  if (!tc_saved || actor.n_layers - 1 > kSavedActorLayers || N < min_rows || getenv("BD_NO_ACTOR_SAVE")) return nullptr;
  return static_cast<const char*>(tc_saved) + saved_layout(r, T, N).off_act;
}
// the actor's layer-0 input images next to them (x0b images of all (t, tile), then the x0s images), or null
const void* imagine_saved_actor_x0(const bd_rssm& r, const bd_mlp& actor, int T, long long N, const void* tc_saved,
                                   const void** x0s) {
  if (!imagine_saved_actor(r, actor, T, N, tc_saved) || getenv("BD_NO_X0_SAVE")) return nullptr;
  // (the rollout's [s ; a] tile must hold the state image's columns: r16(S) <= r16(S + A + 1) always)
  const SavedLayout sl = saved_layout(r, T, N);
  if (x0s) *x0s = static_cast<const char*>(tc_saved) + sl.off_x0s;
  return static_cast<const char*>(tc_saved) + sl.off_x0b;
}

// Do the operand tiles of the rollout engine (two belief tiles, [s;a], one hidden tile) plus at
// least two weight-ring stages fit the 227 KB of shared memory?  (Same arithmetic as plan_smem.)
static bool rollout_tiles_fit(const bd_rssm& r, int extra_hidden) {
  const int Kp_b = r16(r.belief_size + 1), Kp_sa = r16(r.state_size + r.action_size + 1);
  const int Kp_h = max(max(r16(r.hidden_size + 1), Kp_b), r16(extra_hidden + 1));
  const size_t tiles = (size_t)kTileRows * 2 * (2 * Kp_b + Kp_sa + Kp_h);
  const int widest = max(max(r16(r.hidden_size), r16(r.belief_size)), max(192, r16(extra_hidden)));
  const size_t stage = ((size_t)widest * 64 + 1023) & ~size_t(1023);
  return tiles + 2 * stage <= (size_t)227 * 1024 - 4096;
}

bool imagine_supported(const bd_rssm& r, const bd_mlp& actor, int precision) {
  if (precision != BD_PREC_FP16 && precision != BD_PREC_BF16) return false;
  if (!(r.activation == BD_ACT_ELU || r.activation == BD_ACT_RELU || r.activation == BD_ACT_TANH ||
        r.activation == BD_ACT_IDENTITY)) return false;
  if (actor.activation != r.activation) return false;
  if (r.belief_size + 1 > 256 || r.hidden_size + 1 > 256 || r.state_size > 128 || r.action_size > 16)
    return false;
  if (actor.n_layers < 2) return false;
  for (int l = 0; l + 1 < actor.n_layers; ++l)
    if (actor.layer[l].out_features != r.hidden_size) return false;
  return rollout_tiles_fit(r, 0);
}

size_t imagine_pack_bytes(const bd_rssm& r, const bd_mlp& actor) {
  // generous upper bound: every weight padded to multiples of 16 in both dims (+ bias column)
  size_t e = 0;
  auto img = [&](int n, int k) { e += (size_t)r16(n) * r16(k + 1); };
  for (int l = 0; l < actor.n_layers; ++l) img(actor.layer[l].out_features + 16, actor.layer[l].in_features + 16);
  img(r.belief_size, r.state_size + r.action_size);
  e += (size_t)8 * 6 * 64 * r16(r.belief_size + 1);       // GRU slices (<= 8 slices of 64)
  img(r.hidden_size, r.belief_size);
  img(2 * r.state_size + 32, r.hidden_size);
  // narrow output layers are replicated per cluster rank in column-split mode
  e += (size_t)kMaxRanks * (2 * 16 + 2 * r16(r.state_size) + 32) * r16(max(r.hidden_size, r.belief_size) + 1);
  return e * 2 + 4096;
}

static bool plan_smem(int Kp_b, int Kp_sa, int Kp_h, uint32_t stage, SmemPlan& sm, bool has_b1 = true) {
  uint32_t off = 0;
  auto take = [&](uint32_t bytes) { uint32_t o = off; off += (bytes + 1023) & ~1023u; return o; };
  sm.off_tile[0] = take(kTileRows * Kp_b * 2);
  sm.off_tile[1] = (has_b1 && !getenv("BD_DBG_ALIAS")) ? take(kTileRows * Kp_b * 2) : sm.off_tile[0];
  sm.off_tile[2] = take(kTileRows * Kp_sa * 2);
  sm.off_tile[3] = take(kTileRows * Kp_h * 2);
  sm.off_tile[4] = sm.off_tile[3];
  sm.stage_bytes = align_stage(stage);
  sm.off_ring = off;
  const uint32_t budget = 227 * 1024 - 4096;   // static barriers + program tables + alignment slack
  if (off + 2 * sm.stage_bytes > budget) return false;
  sm.nstage = min(8u, (budget - off) / sm.stage_bytes);
  sm.total = off + sm.nstage * sm.stage_bytes + 1024;
  return true;
}


// ---------------------------------------------------------------------------------------------
// Column-split (cluster) mode helpers: rank `rank` of R owns 16-column groups
// [G*rank/R, G*(rank+1)/R) of a G-group output.
// ---------------------------------------------------------------------------------------------
static inline void split_cols(int groups, int rank, int R, int& c0, int& c1) {
  c0 = groups * rank / R * 16;
  c1 = groups * (rank + 1) / R * 16;
}
// every rank must own at least one accumulator group of an n-wide layer written to a Kp_out tile
static bool split_ok_act(int n, int Kp_out, int R) {
  const int G = Kp_out / 16;
  if (G < R) return false;
  for (int k = 0; k < R; ++k) {
    int c0, c1;
    split_cols(G, k, R, c0, c1);
    if (c1 <= c0 || c0 >= r16(n)) return false;
  }
  return true;
}
static bool split_ok_gru(int Be, int R) {
  const int G = r16(Be) / 16;
  const int ns = ((G + R - 1) / R + 3) / 4;
  return G / R >= ns && G / R >= 1;
}
// cluster size for `ntiles` row tiles: split the columns of a tile over 2 or 4 SMs while that
// still leaves every cluster its own SMs (BD_TC_CLUSTER=1|2|4 forces a size, for tests)
static int pick_ranks(long long ntiles, const int* act_n, const int* act_kp, int nact, int Be) {
  int want = ntiles <= 32 ? 4 : (ntiles <= 70 ? 2 : 1);
  if (const char* e = getenv("BD_TC_CLUSTER")) {
    const int v = atoi(e);
    if (v == 1 || v == 2 || v == 4) want = v;
  }
  for (int R = want; R > 1; R >>= 1) {
    bool ok = Be <= 0 || split_ok_gru(Be, R);
    for (int i = 0; i < nact && ok; ++i) ok = split_ok_act(act_n[i], act_kp[i], R);
    if (ok) return R;
  }
  return 1;
}

struct ActSrc { const float* w; int ld, src_c0, len, Kp, a_tile; const float* bias; int bias_k; };
// act(sum_i A_i W_i^T + b) -> operand tile out_tile (Kp_out columns).  R == 1: the whole layer,
// its first GEMM chained on the previous phase's early publication when chain_split > 0 (returns
// this phase's own split).  R > 1: this rank's column slice of the layer (weight rows [c0, c1)).
static int add_act_phase(Builder& b, int rank, int R, const ActSrc* src, int nsrc, int n, int Kp_out,
                         int aux0, int out_tile, int chain_split, int dep_back = 1, bool allow_split = true) {
  const int Np = r16(n);
  int c0 = 0, c1 = Kp_out;
  if (R > 1) split_cols(Kp_out / 16, rank, R, c0, c1);
  const int Nr = max(0, min(c1, Np) - c0);
  const int nvr = max(0, min(n - c0, Nr));
  const int d = b.dcol();
  for (int i = 0; i < nsrc && Nr > 0; ++i) {
    const ActSrc& x = src[i];
    uint32_t off = b.add_pack(x.w, x.ld, c0, nvr, Nr, x.Kp, x.src_c0, x.len, x.bias, x.bias_k);
    if (i == 0 && R == 1 && chain_split > 0) b.chain_gemm(off, Nr, x.Kp, x.a_tile, d, chain_split);
    else b.add_gemm(off, Nr, x.Kp, x.a_tile, 0, d, i > 0 ? 1 : 0);
  }
  b.end_phase(EPI_ACT_H, dep_back, n, Nr, c1, d, aux0, out_tile);
  if (b.ok) b.prog.p[b.prog.n_phases - 1].col0 = (uint16_t)c0;
  return (R == 1 && allow_split) ? b.split_last_phase() : 0;
}
// GRUCell in N-slices of <= 64 belief columns: accumulators IN | R | Z | HN per slice.  Stacked
// gate images: 3 GEMMs per slice instead of 6, so the A tiles (x, h) are fetched from shared
// memory half as often.  R > 1: only this rank's belief columns, in evenly sized slices.
static void add_gru_phases(Builder& b, const bd_rssm& r, int rank, int R, int Kp_x, int Kp_b, int sp_x) {
  const int Be = r.belief_size;
  int n0s[8], nss[8], ns = 0;
  if (R == 1) {
    for (int n0 = 0; n0 < Be && ns < 8; n0 += 64, ++ns) { n0s[ns] = n0; nss[ns] = r16(min(64, Be - n0)); }
  } else {
    const int G = r16(Be) / 16;
    ns = ((G + R - 1) / R + 3) / 4;
    const int g0 = G * rank / R, ng = G * (rank + 1) / R - g0;
    for (int sl = 0; sl < ns; ++sl) {
      const int ga = g0 + ng * sl / ns, gb = g0 + ng * (sl + 1) / ns;
      n0s[sl] = ga * 16; nss[sl] = (gb - ga) * 16;
    }
  }
  for (int slice = 0; slice < ns; ++slice) {
    const int n0 = n0s[slice], Ns = nss[slice];
    const int nv = max(0, min(Ns, Be - n0));
    int d = b.dcol();
    if (Ns > 0) {
      PackSeg rx[3] = {{0, 2 * Be + n0, nv}, {Ns, n0, nv}, {2 * Ns, Be + n0, nv}};      // W_in, W_ir, W_iz
      PackSeg rh[2] = {{0, n0, nv}, {Ns, Be + n0, nv}};                                 // W_hr, W_hz
      uint32_t wx = b.add_pack_rows(r.w_ih, Be, 3, rx, 3 * Ns, Kp_x, 0, Be, r.b_ih, Be);
      uint32_t wh = b.add_pack_rows(r.w_hh, Be, 2, rh, 2 * Ns, Kp_b, 0, Be, r.b_hh, Be);
      uint32_t wn = b.add_pack(r.w_hh, Be, 2 * Be + n0, nv, Ns, Kp_b, 0, Be, r.b_hh, Be);
      if (slice == 0) b.chain_gemm(wx, 3 * Ns, Kp_x, TILE_H, d, sp_x);
      else b.add_gemm(wx, 3 * Ns, Kp_x, TILE_H, 0, d, 0);          // [IN | R | Z] = x W_i*^T + b_i*
      b.add_gemm(wh, 2 * Ns, Kp_b, TILE_BCUR, 0, d + Ns, 1);       // [R | Z]    += h W_h{r,z}^T + b_h{r,z}
      b.add_gemm(wn, Ns, Kp_b, TILE_BCUR, 0, d + 3 * Ns, 0);       // HN          = h W_hn^T + b_hn
    }
    b.end_phase(EPI_GRU, slice == 0 ? 1 : 2, nv, Ns, 0, d, n0, TILE_BNXT);
  }
}
// embed -> GRU -> prior of one transition step (shared by imagine, CEM, TransitionModel.forward)
static void add_transition_phases(Builder& b, const bd_rssm& r, int rank, int R, int Kp_b, int Kp_sa,
                                  int Kp_hid, int Kp_x, bool save) {
  const int Be = r.belief_size, Hi = r.hidden_size, S = r.state_size, A = r.action_size;
  const int Sp = r16(S);
  // x = act(W_sa [s ; a] + b)
  ActSrc se{r.embed.w, S + A, 0, S + A, Kp_sa, TILE_SA, r.embed.b, S + A};
  const int sp_x = add_act_phase(b, rank, R, &se, 1, Be, Kp_x, save ? 1 : 0, TILE_H, 0);   // aux0 = 1: save act'(x)
  add_gru_phases(b, r, rank, R, Kp_x, Kp_b, sp_x);
  // prior: h = act(W_p1 b' + b), (mean | raw std) = W_p2 h + b   (output phase replicated on every rank)
  ActSrc sp1{r.prior1.w, Be, 0, Be, Kp_b, TILE_BNXT, r.prior1.b, Be};
  const int sp_h = add_act_phase(b, rank, R, &sp1, 1, Hi, Kp_hid, save ? 2 : 0, TILE_H, 0);  // aux0 = 2: save act'(h)
  uint32_t wm = b.add_pack(r.prior2.w, Hi, 0, S, Sp, Kp_hid, 0, Hi, r.prior2.b, Hi);
  uint32_t wsd = b.add_pack(r.prior2.w, Hi, S, S, Sp, Kp_hid, 0, Hi, r.prior2.b, Hi);
  const int d = b.dcol();
  b.chain_gemm(wm, Sp, Kp_hid, TILE_H, d, sp_h);
  b.chain_gemm(wsd, Sp, Kp_hid, TILE_H, d + Sp, sp_h);
  b.end_phase(EPI_PRIOR_OUT, 1, 2 * S, Sp, 0, d, 0, TILE_SA);
}

static void set_programs(RolloutArgs& ra, Builder& b, int R) {
  ra.nranks = R;
  if (R == 1) {
    b.finalize_blocks(ra.sm.stage_bytes);
    ra.prog[0] = b.prog;
  } else {
    for (int k = 0; k < R; ++k) {
      Builder::finalize_blocks_of(b.ranks[k], ra.sm.stage_bytes);
      ra.prog[k] = b.ranks[k];
    }
  }
}

// kernel instantiations live in tc_rollout_f{0,1}{a,n}.cu (compiled in parallel)
int launch_rollout_f0a(int act, bool prof, unsigned grid, const RolloutArgs& ra, cudaStream_t s);
int launch_rollout_f0n(int act, bool prof, unsigned grid, const RolloutArgs& ra, cudaStream_t s);
int launch_rollout_f1a(int act, bool prof, unsigned grid, const RolloutArgs& ra, cudaStream_t s);
int launch_rollout_f1n(int act, bool prof, unsigned grid, const RolloutArgs& ra, cudaStream_t s);
static int launch_rollout(int fmt, int act, bool with_actor, bool prof, unsigned grid,
                          const RolloutArgs& ra, cudaStream_t s) {
  if (fmt == 0) return with_actor ? launch_rollout_f0a(act, prof, grid, ra, s) : launch_rollout_f0n(act, prof, grid, ra, s);
  return with_actor ? launch_rollout_f1a(act, prof, grid, ra, s) : launch_rollout_f1n(act, prof, grid, ra, s);
}

// ---------------------------------------------------------------------------------------------
// Fused Dreamer rollout (imagine_and_returns, SURVEY 8b level L2): the reward and value heads
// (src/dreamer.py:321-322) ride in the per-step program and lambda_return (:329-335) is the kernel's
// tail.  The two heads are independent chains: their layers alternate (reward L_k, value L_k,
// reward L_k+1, ...), each depending on the epilogue TWO phases back, so the MMAs of one head run
// under the epilogue of the other (the reward chain lives in the H tile, the value chain in the
// belief ping-pong half that is free between the GRU of this step and the GRU of the next).
// ---------------------------------------------------------------------------------------------
struct HeadsFwd {
  const bd_mlp* head[2];        // reward, value
  float* out[2];                // (T,N) each
  float* returns;               // (T,N)
  double discount, lambda_;
  void* saved;                  // act' images for the fused backward (heads_saved_bytes), or null
};
static int head_hidden(const bd_mlp& m) { return m.layer[0].out_features; }
bool heads_supported(const bd_rssm& r, const bd_mlp& reward, const bd_mlp& value) {
  const bd_mlp* hs[2] = {&reward, &value};
  if (reward.n_layers != value.n_layers) return false;
  for (const bd_mlp* m : hs) {
    if (m->n_layers < 2 || m->n_layers > BD_MAX_LAYERS || m->activation != r.activation) return false;
    if (m->layer[0].in_features != r.belief_size + r.state_size) return false;
    if (m->layer[m->n_layers - 1].out_features != 1) return false;
    const int hh = head_hidden(*m);
    for (int l = 0; l + 1 < m->n_layers; ++l)
      if (m->layer[l].out_features != hh || (l > 0 && m->layer[l].in_features != hh)) return false;
    // hidden tile = a belief tile; backward: one 208-column accumulator region beside ACC_B / d_s
    if (r16(hh + 1) > r16(r.belief_size + 1) || r16(hh) > 208) return false;
  }
  if (r16(r.belief_size) > 208 || r16(r.state_size + r.action_size) > 48) return false;
  // program tables: imagine (<= 12 phases / ~31 GEMMs at 4 GRU slices) + 2 n_layers phases
  const int slices = (r.belief_size + 63) / 64;
  if (8 + slices + 2 * reward.n_layers > kMaxPhases) return false;
  if (16 + 4 * slices + 2 * (reward.n_layers + 1) > kMaxGemms) return false;
  return 13 + 3 * slices + 2 * (reward.n_layers + 1) <= kMaxPackJobs;
}
size_t heads_pack_bytes(const bd_mlp& reward, const bd_mlp& value) {
  return 2 * mlp_pack_bytes(reward) + 2 * mlp_pack_bytes(value) + (size_t)kMaxRanks * 2 * 16 * 272 * 2;
}
struct HeadsSaved { size_t off[2][BD_MAX_LAYERS], total; int Kh; };
static HeadsSaved heads_saved_layout(const bd_mlp& reward, int T, long long N) {
  HeadsSaved h{};
  h.Kh = r16(head_hidden(reward));
  const size_t per = (size_t)((N + kTileRows - 1) / kTileRows) * T * kTileRows * h.Kh * 2;
  size_t o = 0;
  for (int k = 0; k < 2; ++k)
    for (int l = 0; l + 1 < reward.n_layers; ++l) { h.off[k][l] = o; o += per; }
  h.total = o;
  return h;
}
size_t heads_saved_bytes(const bd_mlp& reward, int T, long long N) { return heads_saved_layout(reward, T, N).total + 256; }

// in_tile / hid1_tile: where the (belief | 1) input sits and which tile the second head's chain lives in -- the rollout
// reads the NEW belief (TILE_BNXT) and parks the value chain in the old one; the stand-alone pair forward
// (heads_pair_forward) reads TILE_BCUR and uses TILE_BNXT.  hidden_images: the hidden layers leave their activations
// as sv_mlp images (aux0 = 3 + k (L - 1) + l, the layout bd_mlp_backward reads) instead of act' images (sv_hd).
static void add_head_phases(Builder& b, const bd_rssm& r, const HeadsFwd& hd, int rank, int R, int Kp_b, int Ks,
                            int in_tile = TILE_BNXT, int hid1_tile = TILE_BCUR, bool hidden_images = false) {
  const int Be = r.belief_size, S = r.state_size;
  const int L = hd.head[0]->n_layers;
  for (int l = 0; l < L; ++l)
    for (int k = 0; k < 2; ++k) {
      const bd_mlp& m = *hd.head[k];
      const bd_linear& Lr = m.layer[l];
      const int tile = k == 0 ? TILE_H : hid1_tile;        // this head's hidden tile
      const int dep = (l == 0 && k == 0) ? 1 : 2;           // the same head's previous layer is two phases back
      const int n = Lr.out_features;
      if (l + 1 < L) {
        const int aux = !hd.saved ? 0 : (hidden_images ? 3 + k * (L - 1) + l : 16 + k * BD_MAX_LAYERS + l);
        if (l == 0) {
          ActSrc s0[2] = {{Lr.w, Be + S, 0, Be, Kp_b, in_tile, Lr.b, Be},
                          {Lr.w, Be + S, Be, S, Ks, TILE_SA, nullptr, -1}};
          add_act_phase(b, rank, R, s0, 2, n, r16(n + 1), aux, tile, 0, dep, false);
        } else {
          const int kin = Lr.in_features;
          ActSrc sl{Lr.w, kin, 0, kin, r16(kin + 1), tile, Lr.b, kin};
          add_act_phase(b, rank, R, &sl, 1, n, r16(n + 1), aux, tile, 0, dep, false);
        }
      } else {   // scalar output: replicated on every rank
        const int kin = Lr.in_features, Kp = r16(kin + 1);
        const int d = b.dcol();
        uint32_t w = b.add_pack(Lr.w, kin, 0, 1, 16, Kp, 0, kin, Lr.b, kin);
        b.add_gemm(w, 16, Kp, tile, 0, d, 0);
        b.end_phase(EPI_HEAD_OUT, dep, 1, 16, 0, d, k, tile);
      }
    }
}

static int imagine_forward_impl(const bd_imagine_args* a, const HeadsFwd* hd, void* ws, size_t ws_bytes,
                                int precision, bd_stream_t stream) {
  const bd_rssm& r = a->rssm;
  const bd_mlp& ac = a->actor;
  if (!imagine_supported(r, ac, precision))
    BD_FAIL(BD_ERR_UNSUPPORTED, "tensor-core imagine_forward: sizes/activation not supported "
                                "(Be,Hi <= 255, S <= 128, A <= 16, ELU/ReLU/Tanh/Identity)");
  const int Be = r.belief_size, Hi = r.hidden_size, S = r.state_size, A = r.action_size;
  const int Kp_b = r16(Be + 1), Kp_sa = r16(S + A + 1), Kp_hid = r16(Hi + 1), Kp_x = r16(Be + 1);
  const int Ks = r16(S);
  int Kp_h = max(Kp_hid, Kp_x);
  if (hd) Kp_h = max(Kp_h, r16(head_hidden(*hd->head[0]) + 1));
  const int Nh = r16(Hi), Nb = r16(Be), Ap = 16, Sp = r16(S);
  const long long ntiles = (a->N + kTileRows - 1) / kTileRows;
  const int act_n[2] = {Hi, Be}, act_kp[2] = {Kp_hid, Kp_x};
  int R = pick_ranks(ntiles, act_n, act_kp, 2, Be);
  Builder b;
  for (;; R = 1, b = Builder()) {     // a column-split build that overflows the tables falls back to R = 1
  for (int rank = 0; rank < R; ++rank) {
    // ---- actor (src/models.py:506-517): L0 on [b ; s], hidden layers, output (mean | std)
    const bd_linear& L0 = ac.layer[0];
    ActSrc s0[2] = {{L0.w, Be + S, 0, Be, Kp_b, TILE_BCUR, L0.b, Be},
                    {L0.w, Be + S, Be, S, Ks, TILE_SA, nullptr, -1}};
    // (aux0 = 3 + l: the hidden activation image goes to sv_mlp[l] when the backward will want it)
    int sp = add_act_phase(b, rank, R, s0, 2, Hi, Kp_hid, 3, TILE_H, 0);
    for (int l = 1; l + 1 < ac.n_layers; ++l) {
      const bd_linear& L = ac.layer[l];
      ActSrc sl{L.w, Hi, 0, Hi, Kp_hid, TILE_H, L.b, Hi};
      sp = add_act_phase(b, rank, R, &sl, 1, Hi, Kp_hid, 3 + l, TILE_H, sp);
    }
    {   // output layer: narrow, replicated on every rank in column-split mode
      const bd_linear& Lo = ac.layer[ac.n_layers - 1];
      uint32_t wm = b.add_pack(Lo.w, Hi, 0, A, Ap, Kp_hid, 0, Hi, Lo.b, Hi);
      uint32_t wsd = b.add_pack(Lo.w, Hi, A, A, Ap, Kp_hid, 0, Hi, Lo.b, Hi);
      const int d = b.dcol();
      b.chain_gemm(wm, Ap, Kp_hid, TILE_H, d, sp);
      b.chain_gemm(wsd, Ap, Kp_hid, TILE_H, d + Ap, sp);
      b.end_phase(EPI_ACTOR_OUT, 1, 2 * A, Ap, 0, d, 0, TILE_SA);
    }
    add_transition_phases(b, r, rank, R, Kp_b, Kp_sa, Kp_hid, Kp_x, true);
    if (hd) add_head_phases(b, r, *hd, rank, R, Kp_b, Ks);
    if (R > 1) b.end_rank(rank);
  }
  if (b.ok || R == 1) break;
  }
  if (!b.ok) BD_FAIL(BD_ERR_UNSUPPORTED, "tensor-core imagine_forward: program too large");
  const size_t pack_bytes = (size_t)b.w_elems * 2;
  if (pack_bytes > ws_bytes)
    BD_FAIL(BD_ERR_WORKSPACE, "tensor-core imagine_forward: workspace %zu < %zu", ws_bytes, pack_bytes);

  RolloutArgs ra{};
  if (!plan_smem(Kp_b, Kp_sa, Kp_h, b.max_stage, ra.sm))
    BD_FAIL(BD_ERR_UNSUPPORTED, "tensor-core imagine_forward: tiles do not fit shared memory");
  set_programs(ra, b, R);
  ra.wpack = static_cast<const uint16_t*>(ws);
  ra.N = a->N; ra.T = a->T; ra.Be = Be; ra.S = S; ra.A = A; ra.Hi = Hi; ra.J = a->actor_cfg.entropy_samples;
  ra.Kp_b = Kp_b; ra.Kp_sa = Kp_sa; ra.Kp_h = Kp_h; ra.act = r.activation; ra.min_std = r.min_std_dev;
  ra.cfg = a->actor_cfg;
  ra.prev_state = a->prev_state; ra.prev_belief = a->prev_belief;
  ra.eps_a = a->eps_a; ra.eps_e = a->eps_e; ra.eps_s = a->eps_s;
  ra.beliefs = a->beliefs; ra.states = a->states; ra.means = a->means; ra.stds = a->stds;
  ra.entropy = a->entropy; ra.actions = a->actions; ra.actor_raw = a->actor_raw; ra.dent = a->dent;
  ra.has_b1 = 1;
  if (hd) {
    ra.head_out[0] = hd->out[0]; ra.head_out[1] = hd->out[1]; ra.returns = hd->returns;
    ra.lr_disc = (float)hd->discount; ra.lr_lam = (float)hd->lambda_; ra.lr_oml = (float)(1.0 - hd->lambda_);
    if (hd->saved) {
      const HeadsSaved hs = heads_saved_layout(*hd->head[0], a->T, a->N);
      ra.kh_hd = hs.Kh;
      for (int k = 0; k < 2; ++k)
        for (int l = 0; l + 1 < hd->head[0]->n_layers; ++l)
          ra.sv_hd[k * BD_MAX_LAYERS + l] = reinterpret_cast<uint16_t*>(static_cast<char*>(hd->saved) + hs.off[k][l]);
    }
  }
  if (a->tc_saved && !getenv("BD_DBG_NOSAVE")) {
    SavedLayout sl = saved_layout(r, a->T, a->N);
    char* sb = static_cast<char*>(a->tc_saved);
    ra.sv_gate = reinterpret_cast<uint16_t*>(sb + sl.off_gate);
    ra.sv_xa = reinterpret_cast<uint16_t*>(sb + sl.off_xa);
    ra.sv_ha = reinterpret_cast<uint16_t*>(sb + sl.off_ha);
    ra.kb_sv = sl.Kb; ra.kh_sv = sl.Kh;
    if (imagine_saved_actor(r, ac, a->T, a->N, a->tc_saved)) {
      ra.kact_sv = sl.Kact;
      for (int l = 0; l + 1 < ac.n_layers; ++l)
        ra.sv_mlp[l] = reinterpret_cast<uint16_t*>(sb + sl.off_act + (size_t)l * sl.act_bytes);
      if (imagine_saved_actor_x0(r, ac, a->T, a->N, a->tc_saved, nullptr)) {
        ra.x0.x0b = reinterpret_cast<uint16_t*>(sb + sl.off_x0b);
        ra.x0.x0s = reinterpret_cast<uint16_t*>(sb + sl.off_x0s);
        ra.x0.bytes_b = (uint32_t)kTileRows * sl.Kx0b * 2;
        ra.x0.bytes_s = (uint32_t)kTileRows * sl.Kx0s * 2;
      }
    }
  }

  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // debug: BD_TC_PROF=1 appends cycle counters after the packed weights (scripts/prof_fwd.py)
  const bool prof = getenv("BD_TC_PROF") && pack_bytes + 8192 + kMaxPhases * 64 + 160 * 24 <= ws_bytes;
  if (prof) {
    ra.prof = reinterpret_cast<long long*>(static_cast<char*>(ws) + ((pack_bytes + 4095) & ~size_t(4095)));
    cudaMemsetAsync(ra.prof, 0, kMaxPhases * 64 + 160 * 24, s);
    if (getenv("BD_TC_PROF")[0] == '2') {     // per-stage counters as well (they perturb the issuer)
      const long long one = 1;
      cudaMemcpyAsync(ra.prof + 39 * 8 + 7, &one, sizeof(one), cudaMemcpyHostToDevice, s);
    }
    if (const char* e = getenv("BD_TC_DBG")) ra.dbg = atoi(e);
  }
  long long max_img = 0;
  for (int i = 0; i < b.pack.njobs; ++i)
    max_img = max(max_img, (long long)b.pack.job[i].Np * b.pack.job[i].Kp);
  long long pgx = (max_img + 255) / 256;
  if (pgx > 64) pgx = 64;
  dim3 pgrid((unsigned)pgx, (unsigned)b.pack.njobs);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned grid = (unsigned)(ntiles < sms ? ntiles : sms);
  const int fmt = precision == BD_PREC_FP16 ? 0 : 1;
  if (fmt == 0) pack_weights_kernel<0><<<pgrid, 256, 0, s>>>(b.pack, static_cast<uint16_t*>(ws));
  else pack_weights_kernel<1><<<pgrid, 256, 0, s>>>(b.pack, static_cast<uint16_t*>(ws));
  BD_CUDA_LAUNCH_CHECK();
  {
    ProfScope ps(BD_PROF_ROLLOUT_FWD, s);
    BD_TRY(launch_rollout(fmt, r.activation, true, prof, grid, ra, s));
  }
  ProfScope pe(BD_PROF_ENTROPY, s);
  // entropy + its gradient wrt (mean, std): independent of the recurrence -> separate parallel pass
  if ((long long)a->T * a->N <= (1 << 16)) {        // few row-steps: 8 lanes per row
    dim3 egrid((unsigned)((a->N * 8 + 127) / 128), (unsigned)a->T);
    actor_entropy_kernel<8><<<egrid, 128, 0, s>>>(a->actor_raw, a->eps_e, a->actor_cfg, a->N, A, a->entropy, a->dent);
  } else if (A == 1 && (a->N & 3) == 0 && !getenv("BD_ENT_SCALAR") &&
             ((reinterpret_cast<uintptr_t>(a->actor_raw) | reinterpret_cast<uintptr_t>(a->eps_e) |
               reinterpret_cast<uintptr_t>(a->entropy) | reinterpret_cast<uintptr_t>(a->dent)) & 15) == 0) {
    dim3 egrid((unsigned)((a->N / 4 + 127) / 128), (unsigned)a->T);     // four consecutive rows per thread
    actor_entropy_vec4_kernel<<<egrid, 128, 0, s>>>(a->actor_raw, a->eps_e, a->actor_cfg, a->N, a->entropy, a->dent);
  } else {
    dim3 egrid((unsigned)((a->N + 127) / 128), (unsigned)a->T);
    actor_entropy_kernel<1><<<egrid, 128, 0, s>>>(a->actor_raw, a->eps_e, a->actor_cfg, a->N, A, a->entropy, a->dent);
  }
  BD_CUDA_LAUNCH_CHECK();
  return BD_OK;
}

int imagine_forward(const bd_imagine_args* a, void* ws, size_t ws_bytes, int precision,
                    bd_stream_t stream) {
  return imagine_forward_impl(a, nullptr, ws, ws_bytes, precision, stream);
}
int imagine_returns_forward(const bd_imagine_args* a, const bd_mlp* reward, const bd_mlp* value,
                            double discount, double lambda_, float* reward_out, float* value_out,
                            float* returns, void* heads_saved, void* ws, size_t ws_bytes, int precision,
                            bd_stream_t stream) {
  if (!imagine_supported(a->rssm, a->actor, precision) || !heads_supported(a->rssm, *reward, *value))
    BD_FAIL(BD_ERR_UNSUPPORTED, "fused imagine + heads + lambda_return: configuration not supported by the "
                                "tensor-core rollout (use the piecewise entries)");
  HeadsFwd hd{};
  hd.head[0] = reward; hd.head[1] = value; hd.out[0] = reward_out; hd.out[1] = value_out;
  hd.returns = returns; hd.discount = discount; hd.lambda_ = lambda_; hd.saved = heads_saved;
  return imagine_forward_impl(a, &hd, ws, ws_bytes, precision, stream);
}

bool cem_supported(const bd_rssm& r, const bd_mlp& reward, int precision) {
  if (precision != BD_PREC_FP16 && precision != BD_PREC_BF16) return false;
  if (!(r.activation == BD_ACT_ELU || r.activation == BD_ACT_RELU || r.activation == BD_ACT_TANH ||
        r.activation == BD_ACT_IDENTITY) || reward.activation != r.activation) return false;
  if (r.belief_size + 1 > 256 || r.hidden_size + 1 > 256 || r.state_size > 128 || r.action_size > 16) return false;
  if (reward.n_layers < 2 || reward.layer[reward.n_layers - 1].out_features != 1) return false;
  int widest = 0;
  for (int l = 0; l + 1 < reward.n_layers; ++l) {
    if (reward.layer[l].out_features + 1 > 256) return false;
    widest = max(widest, reward.layer[l].out_features);
  }
  if (!rollout_tiles_fit(r, widest)) return false;
  // the 5-slice GRU + head programs must fit the job tables
  return (r.belief_size + 63) / 64 * 6 + 2 * reward.n_layers + 6 <= kMaxPackJobs;
}
size_t cem_tc_workspace_bytes(const bd_rssm& r, const bd_mlp& reward, long long rows, int H) {
  // packed transition (+ head) images | the deferred head's own pack | beliefs | means, stds, states | rewards
  return imagine_pack_bytes(r, reward) + 2 * mlp_pack_bytes(reward) +
         ((size_t)H * rows * (r.belief_size + 3 * r.state_size) + (size_t)H * rows) * sizeof(float) + 16384;
}

int mlp_forward(const bd_mlp* m, const float* x1, int k1, const float* x2, int k2, int64_t rows,
                float* y, void* ws, size_t ws_bytes, int precision, bd_stream_t stream, void* saved);

// Candidate evaluation of one CEM iteration on the rollout engine: actions are given; a->actions must already hold
// the sampled local actions (H, B, Cl, A).
// The reward head does not feed the recurrence, so it does not have to sit on the per-step serial chain of a plan
// that is 120 strictly dependent steps long: by default the rollout only leaves (beliefs, states) of every step
// (fp32, (H, rows, .)) and ONE batched MLP forward over all H * rows latents computes the rewards afterwards
// (src/planner.py:68-72 evaluates the reward model on the stacked latents in the same way).  The rollout's step
// shrinks from 8 to 4 dependent phases (column-split clusters of 4: embed, GRU, prior hidden, prior output).
// Used while the row tiles leave SMs idle (<= 70 tiles); BD_CEM_FUSED_HEAD=1 / 0 forces the head inside the rollout
// (every step: + one phase per head layer) / the batched pass.
int cem_rollout(const bd_cem_eval_args* a, void* ws, size_t ws_bytes, int precision, float* rew_out,
                bd_stream_t stream, bool weights_packed) {
  const bd_rssm& r = a->rssm;
  const bd_mlp& hd = a->reward;
  const int Be = r.belief_size, Hi = r.hidden_size, S = r.state_size, A = r.action_size;
  const int Kp_b = r16(Be + 1), Kp_sa = r16(S + A + 1), Kp_hid = r16(Hi + 1), Kp_x = r16(Be + 1), Ks = r16(S);
  int Kp_h = max(Kp_hid, Kp_x);
  const int Cl = a->c_end - a->c_begin;
  const long long rows = (long long)a->B * Cl;
  const long long ntiles = (rows + kTileRows - 1) / kTileRows;
  // (measured: 1 000 candidates = 8 row tiles: 4.22 -> 2.77 ms per plan; 10 000 candidates = 79 tiles: 4.38 -> 4.61 ms --
  // once the row tiles fill the GPU the step is no longer pure latency and the batched pass costs more than it saves)
  bool defer_head = ntiles <= 70;
  if (const char* e = getenv("BD_CEM_FUSED_HEAD")) defer_head = atoi(e) == 0;
  defer_head = defer_head && mlp_supported(hd, Be, S, precision);
  int act_n[2 + BD_MAX_LAYERS] = {Hi, Be}, act_kp[2 + BD_MAX_LAYERS] = {Kp_hid, Kp_x}, nact = 2;
  for (int l = 0; l + 1 < hd.n_layers && !defer_head; ++l) {
    act_n[nact] = hd.layer[l].out_features; act_kp[nact] = r16(hd.layer[l].out_features + 1); ++nact;
    Kp_h = max(Kp_h, r16(hd.layer[l].out_features + 1));
  }
  int R = pick_ranks(ntiles, act_n, act_kp, nact, Be);
  Builder b;
  for (;; R = 1, b = Builder()) {
  for (int rank = 0; rank < R; ++rank) {
    add_transition_phases(b, r, rank, R, Kp_b, Kp_sa, Kp_hid, Kp_x, false);
    // reward head on (b', s'): DenseModel (src/planner.py:68-72)
    int sp_hd = 0;
    for (int l = 0; l < hd.n_layers && !defer_head; ++l) {
      const bd_linear& L = hd.layer[l];
      const bool last = (l == hd.n_layers - 1);
      const int n = L.out_features;
      if (!last) {
        if (l == 0) {
          ActSrc s0[2] = {{L.w, Be + S, 0, Be, Kp_b, TILE_BNXT, L.b, Be},
                          {L.w, Be + S, Be, S, Ks, TILE_SA, nullptr, -1}};
          sp_hd = add_act_phase(b, rank, R, s0, 2, n, r16(n + 1), 0, TILE_H, 0);
        } else {
          const int kin = L.in_features;
          ActSrc sl{L.w, kin, 0, kin, r16(kin + 1), TILE_H, L.b, kin};
          sp_hd = add_act_phase(b, rank, R, &sl, 1, n, r16(n + 1), 0, TILE_H, sp_hd);
        }
      } else {   // scalar output: replicated
        const int kin = L.in_features, Kp = r16(kin + 1), Np = r16(n);
        const int d = b.dcol();
        uint32_t w = b.add_pack(L.w, kin, 0, n, Np, Kp, 0, kin, L.b, kin);
        b.chain_gemm(w, Np, Kp, TILE_H, d, sp_hd);
        b.end_phase(EPI_HEAD_OUT, 1, 1, Np, 0, d, 0, TILE_H);
      }
    }
    if (R > 1) b.end_rank(rank);
  }
  if (b.ok || R == 1) break;
  }
  if (!b.ok) BD_FAIL(BD_ERR_UNSUPPORTED, "tensor-core CEM: program too large");
  char* base = static_cast<char*>(ws);
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p_ = base + off; off += (bytes + 255) & ~size_t(255); return p_; };
  uint16_t* wpack = reinterpret_cast<uint16_t*>(take((size_t)b.w_elems * 2));
  float* beliefs = reinterpret_cast<float*>(take((size_t)a->H * rows * Be * sizeof(float)));
  float *means = nullptr, *stds = nullptr, *states = nullptr;
  char* head_ws = nullptr;
  const size_t head_ws_bytes = mlp_pack_bytes(hd) + 4096;
  if (defer_head) {
    means = reinterpret_cast<float*>(take((size_t)a->H * rows * S * sizeof(float)));
    stds = reinterpret_cast<float*>(take((size_t)a->H * rows * S * sizeof(float)));
    states = reinterpret_cast<float*>(take((size_t)a->H * rows * S * sizeof(float)));
    head_ws = take(head_ws_bytes);
  }
  if (off > ws_bytes) BD_FAIL(BD_ERR_WORKSPACE, "tensor-core CEM: workspace %zu < %zu", ws_bytes, off);

  RolloutArgs ra{};
  if (!plan_smem(Kp_b, Kp_sa, Kp_h, b.max_stage, ra.sm))
    BD_FAIL(BD_ERR_UNSUPPORTED, "tensor-core CEM: tiles do not fit shared memory");
  set_programs(ra, b, R);
  ra.wpack = wpack;
  ra.N = rows; ra.T = a->H; ra.Be = Be; ra.S = S; ra.A = A; ra.Hi = Hi; ra.J = 0;
  ra.Kp_b = Kp_b; ra.Kp_sa = Kp_sa; ra.Kp_h = Kp_h; ra.act = r.activation; ra.min_std = r.min_std_dev;
  ra.prev_state = a->state; ra.prev_belief = a->belief; ra.eps_s = a->eps_s;
  ra.beliefs = beliefs; ra.states = states; ra.means = means; ra.stds = stds;
  ra.head_out[0] = defer_head ? nullptr : rew_out; ra.ext_actions = a->actions; ra.has_b1 = 1;
  ra.cem_cl = Cl; ra.cem_c = a->C; ra.cem_c0 = a->c_begin;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  long long max_img = 0;
  for (int i = 0; i < b.pack.njobs; ++i)
    max_img = max(max_img, (long long)b.pack.job[i].Np * b.pack.job[i].Kp);
  long long pgx = (max_img + 255) / 256;
  if (pgx > 64) pgx = 64;
  dim3 pgrid((unsigned)pgx, (unsigned)b.pack.njobs);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned grid = (unsigned)(ntiles < sms ? ntiles : sms);
  const int fmt = precision == BD_PREC_FP16 ? 0 : 1;
  if (!weights_packed) {     // bd_cem_plan: iterations after the first reuse the images at `wpack`
    if (fmt == 0) pack_weights_kernel<0><<<pgrid, 256, 0, s>>>(b.pack, wpack);
    else pack_weights_kernel<1><<<pgrid, 256, 0, s>>>(b.pack, wpack);
    BD_CUDA_LAUNCH_CHECK();
  }
  {
    ProfScope ps(BD_PROF_ROLLOUT_FWD, s);
    BD_TRY(launch_rollout(fmt, r.activation, false, false, grid, ra, s));
  }
  if (!defer_head) return BD_OK;
  // rewards of all (step, candidate) latents in one batched pass: rew_out[t * rows + n]
  return mlp_forward(&hd, beliefs, Be, states, S, (int64_t)a->H * rows, rew_out, head_ws, head_ws_bytes, precision,
                     stream, nullptr);
}

// TransitionModel.forward, prior-only mode without a nonterminal mask (src/models.py:239-260), on the
// rollout engine: actions are given, outputs are the reference's four (L,B,.) tensors.
bool transition_supported(const bd_transition_args& a, int precision) {
  if (precision != BD_PREC_FP16 && precision != BD_PREC_BF16) return false;
  if (a.embeddings || a.nonterminals) return false;
  const bd_rssm& r = a.rssm;
  if (!(r.activation == BD_ACT_ELU || r.activation == BD_ACT_RELU || r.activation == BD_ACT_TANH ||
        r.activation == BD_ACT_IDENTITY)) return false;
  return r.belief_size + 1 <= 256 && r.hidden_size + 1 <= 256 && r.state_size <= 128 && r.action_size <= 16 &&
         rollout_tiles_fit(r, 0);
}
int transition_forward(const bd_transition_args* a, void* ws, size_t ws_bytes, int precision,
                       bd_stream_t stream) {
  const bd_rssm& r = a->rssm;
  const int Be = r.belief_size, Hi = r.hidden_size, S = r.state_size, A = r.action_size;
  const int Kp_b = r16(Be + 1), Kp_sa = r16(S + A + 1), Kp_hid = r16(Hi + 1), Kp_x = r16(Be + 1);
  const long long ntiles = (a->B + kTileRows - 1) / kTileRows;
  const int act_n[2] = {Hi, Be}, act_kp[2] = {Kp_hid, Kp_x};
  const int R = pick_ranks(ntiles, act_n, act_kp, 2, Be);
  Builder b;
  for (int rank = 0; rank < R; ++rank) {
    add_transition_phases(b, r, rank, R, Kp_b, Kp_sa, Kp_hid, Kp_x, false);
    if (R > 1) b.end_rank(rank);
  }
  if (!b.ok) BD_FAIL(BD_ERR_UNSUPPORTED, "tensor-core transition_forward: program too large");
  if ((size_t)b.w_elems * 2 > ws_bytes)
    BD_FAIL(BD_ERR_WORKSPACE, "tensor-core transition_forward: workspace %zu too small", ws_bytes);
  RolloutArgs ra{};
  if (!plan_smem(Kp_b, Kp_sa, max(Kp_hid, Kp_x), b.max_stage, ra.sm))
    BD_FAIL(BD_ERR_UNSUPPORTED, "tensor-core transition_forward: tiles do not fit shared memory");
  set_programs(ra, b, R);
  ra.wpack = static_cast<const uint16_t*>(ws);
  ra.N = a->B; ra.T = a->L; ra.Be = Be; ra.S = S; ra.A = A; ra.Hi = Hi;
  ra.Kp_b = Kp_b; ra.Kp_sa = Kp_sa; ra.Kp_h = max(Kp_hid, Kp_x); ra.act = r.activation;
  ra.min_std = r.min_std_dev;
  ra.prev_state = a->init_state; ra.prev_belief = a->init_belief; ra.eps_s = a->eps_prior;
  ra.beliefs = a->beliefs; ra.states = a->prior_states; ra.means = a->prior_means; ra.stds = a->prior_stds;
  ra.ext_actions = a->actions; ra.has_b1 = 1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  long long max_img = 0;
  for (int i = 0; i < b.pack.njobs; ++i)
    max_img = max(max_img, (long long)b.pack.job[i].Np * b.pack.job[i].Kp);
  long long pgx = (max_img + 255) / 256;
  if (pgx > 64) pgx = 64;
  dim3 pgrid((unsigned)pgx, (unsigned)b.pack.njobs);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned grid = (unsigned)(ntiles < sms ? ntiles : sms);
  const int fmt = precision == BD_PREC_FP16 ? 0 : 1;
  if (fmt == 0) pack_weights_kernel<0><<<pgrid, 256, 0, s>>>(b.pack, static_cast<uint16_t*>(ws));
  else pack_weights_kernel<1><<<pgrid, 256, 0, s>>>(b.pack, static_cast<uint16_t*>(ws));
  BD_CUDA_LAUNCH_CHECK();
  ProfScope ps(BD_PROF_ROLLOUT_FWD, s);
  return launch_rollout(fmt, r.activation, false, false, grid, ra, s);
}

// ---------------------------------------------------------------------------------------------
// BPTT of the imagination rollout on the tensor-core engine (tc_bptt.cuh)
// ---------------------------------------------------------------------------------------------
static uint32_t add_pack_T(Builder& b, const float* w, int ld, int n_valid, int Np, int Kp, int nseg,
                           const PackSeg* segs, int col0 = 0) {
  // packed(n, k) = w[src_row(k), n]: the dgrad (transposed) image of an nn.Linear weight
  if (b.pack.njobs >= kMaxPackJobs) { b.ok = false; return 0; }
  PackJob& j = b.pack.job[b.pack.njobs++];
  j = PackJob{};
  // (transposed jobs: image row n <- source COLUMN col0 + n)
  j.w = w; j.bias = nullptr; j.dst_off = b.w_elems; j.ld = ld; j.row0 = col0; j.N = n_valid; j.Np = Np;
  j.Kp = Kp; j.bias_k = -1; j.nseg = nseg; j.transpose = 1;
  for (int i = 0; i < nseg; ++i) j.seg[i] = segs[i];
  const uint32_t off = (uint32_t)b.w_elems;
  b.w_elems += (long long)Np * Kp;
  return off;
}

size_t bptt_workspace_bytes(const bd_rssm& r, int T, long long N) {
  const int Kb = r16(r.belief_size);
  const size_t gbt = (size_t)T * ((N + kTileRows - 1) / kTileRows) * kTileRows * Kb * 4 + 256;   // tiled g_beliefs
  size_t pack = (size_t)r16(r.hidden_size) * 2 * r16(r.state_size) + (size_t)Kb * r16(r.hidden_size) +
                (size_t)2 * Kb * 3 * (r.belief_size + 64) + (size_t)r16(r.state_size + r.action_size) * Kb;
  return pack * 2 + 4096 + (size_t)2 * 160 * kTileRows * Kb * 4 + 65536 + gbt;
}

struct HeadsBwd {
  const bd_mlp* head[2];
  const void* saved;
  const float *g_returns, *g_reward, *g_value;
  double discount, lambda_;
};
size_t heads_bwd_workspace_bytes(const bd_rssm& r, const bd_mlp& reward, int T, long long N) {
  (void)r; (void)N;
  const size_t kh = r16(head_hidden(reward));
  size_t pack = 2 * ((size_t)(reward.n_layers - 2) * kh * kh + kh * 256 + kh * 48);   // W^T images of both heads
  return pack * 2 + (size_t)160 * T * 2 * kTileRows * sizeof(float) + 8192;
}

static int imagine_bptt_impl(const bd_imagine_bwd_args* a, const HeadsBwd* hb, float* d_raw, void* ws,
                             size_t ws_bytes, int precision, bd_stream_t stream) {
  const bd_imagine_args& f = a->fwd;
  const bd_rssm& r = f.rssm;
  const int Be = r.belief_size, Hi = r.hidden_size, S = r.state_size, A = r.action_size;
  const int Kb = r16(Be), Kh = r16(Hi), Sp = r16(S), Ksa = r16(S + A);
  if (!f.tc_saved) BD_FAIL(BD_ERR_BAD_ARG, "imagine_bptt: forward did not save tensor-core state");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Builder b;
  const int hL = hb ? hb->head[0]->n_layers : 0;
  const int Khd = hb ? r16(head_hidden(*hb->head[0])) : 0;
  if (hb) {
    // fused heads: for head k, d h_last (rank-1, no MMA) -> dgrad through the hidden layers (D . act'(h) from the
    // forward's images) -> dX = d h_0 W_0, whose belief part accumulates into ACC_B and whose state part into the
    // carried d s (TMEM columns [256, 256 + Sp)); the chain's own accumulator sits at column 304.  The dX GEMMs of
    // head k ride in the phase whose epilogue starts the next chain (its accumulator wait keeps the H tile safe).
    for (int k = 0; k < 2; ++k) {
      const bd_mlp& m = *hb->head[k];
      b.end_phase(EPI_P_HEAD_DY, 1, head_hidden(m), 0, 0, 0, k, TILE_H);
      for (int l = hL - 2; l >= 1; --l) {       // d h_{l-1} = (d h_l W_l) . act'(h_{l-1})
        PackSeg sg[1] = {{0, 0, m.layer[l].out_features}};
        uint32_t w = add_pack_T(b, m.layer[l].w, m.layer[l].in_features, m.layer[l].in_features, Khd, Khd, 1, sg);
        b.add_gemm(w, Khd, Khd, TILE_H, 0, 304, 0);
        b.end_phase(EPI_P_MULSAVED, 1, m.layer[l].in_features, Khd, 0, 304, 16 + k * BD_MAX_LAYERS + (l - 1), TILE_H);
      }
      PackSeg sg[1] = {{0, 0, m.layer[0].out_features}};
      uint32_t wb = add_pack_T(b, m.layer[0].w, Be + S, Be, Kb, Khd, 1, sg, 0);
      uint32_t wsx = add_pack_T(b, m.layer[0].w, Be + S, S, Sp, Khd, 1, sg, Be);
      b.add_gemm(wb, Kb, Khd, TILE_H, 0, 0, k == 0 ? 2 : 1);        // ACC_B (+)= d h_0 W_0[:, :Be]
      b.add_gemm(wsx, Sp, Khd, TILE_H, 0, 256, k == 0 ? 2 : 1);     // d s   (+)= d h_0 W_0[:, Be:]
      // (closed by the next end_phase: head 1's HEAD_DY, or p0 below)
    }
  }
  // p0: d_pre2 (no MMA of its own)
  b.end_phase(EPI_P_DPRE2, 1, 2 * S, 0, 0, 0, 0, TILE_D2);
  // p1: DH = d_pre2 * W_p2  (K = [mean rows | std rows], N = Hi)
  {
    PackSeg sg[2] = {{0, 0, S}, {Sp, S, S}};
    uint32_t w = add_pack_T(b, r.prior2.w, Hi, Hi, Kh, 2 * Sp, 2, sg);
    b.add_gemm(w, Kh, 2 * Sp, TILE_D2, 0, 256, 0);
    b.end_phase(EPI_P_MULSAVED, 1, Hi, Kh, 0, 256, 2, TILE_H);
  }
  // p2: ACC_B (+)= d_h * W_p1 ; then gate stage A slice 0
  int slab = 0;
  {
    PackSeg sg[1] = {{0, 0, Hi}};
    uint32_t w = add_pack_T(b, r.prior1.w, Be, Be, Kb, Kh, 1, sg);
    b.add_gemm(w, Kb, Kh, TILE_H, 0, 0, hb ? 1 : 2);     // (with fused heads ACC_B already holds their d b_t)
  }
  int nsl = 0, n0s[8], nss[8], nvs[8];
  for (int n0 = 0; n0 < Be; n0 += 64) { n0s[nsl] = n0; nvs[nsl] = min(64, Be - n0); nss[nsl] = r16(nvs[nsl]); ++nsl; }
  auto gate_phase = [&](int sl, bool passB) {
    b.end_phase(EPI_P_GATE, 1, nvs[sl], nss[sl], 0, 0, n0s[sl], slab ? TILE_SLAB1 : TILE_SLAB0);
    b.prog.p[b.prog.n_phases - 1].pad = passB ? 1 : 0;
    // every gate stage except the very first may start as soon as the previous phase's MMAs are done
    b.prog.p[b.prog.n_phases - 1].Kp_out = (passB || sl > 0) ? 1 : 0;
  };
  gate_phase(0, false);
  for (int pass = 0; pass < 2; ++pass) {
    const float* wsrc = pass == 0 ? r.w_ih : r.w_hh;
    const int dcol = pass == 0 ? 256 : 0;
    for (int sl = 0; sl < nsl; ++sl) {
      const int Ns = nss[sl], nv = nvs[sl], n0 = n0s[sl];
      PackSeg sg[3] = {{0, n0, nv}, {Ns, Be + n0, nv}, {2 * Ns, 2 * Be + n0, nv}};
      uint32_t w = add_pack_T(b, wsrc, Be, Be, Kb, 3 * Ns, 3, sg);
      // (stage B accumulates from its first slice on: ACC_B holds the carry z . G written by stage A)
      b.add_gemm(w, Kb, 3 * Ns, slab ? TILE_SLAB1 : TILE_SLAB0, 0, dcol, (pass == 1 || sl > 0) ? 1 : 0);
      slab ^= 1;
      if (pass == 0 && sl + 1 < nsl) gate_phase(sl + 1, false);
      else if (pass == 0) gate_phase(0, true);
      else if (sl + 1 < nsl) gate_phase(sl + 1, true);
      else b.end_phase(EPI_P_MULSAVED, 1, Be, Kb, 0, 256, 1, TILE_H);
    }
  }
  {
    PackSeg sg[1] = {{0, 0, Be}};
    uint32_t w = add_pack_T(b, r.embed.w, S + A, S + A, Ksa, Kb, 1, sg);
    b.add_gemm(w, Ksa, Kb, TILE_H, 0, 256, 0);
    b.end_phase(EPI_P_DSA, 1, S + A, Ksa, 0, 256, 0, TILE_H);
  }
  if (!b.ok) BD_FAIL(BD_ERR_UNSUPPORTED, "tensor-core BPTT: program too large");

  BpttArgs ba{};
  {
    SmemPlan& sm = ba.sm;
    uint32_t o = 0;
    auto tk = [&](uint32_t bytes) { uint32_t r_ = o; o += (bytes + 1023) & ~1023u; return r_; };
    for (int i = 0; i < 8; ++i) sm.off_tile[i] = 0;
    sm.off_tile[TILE_D2] = tk(kTileRows * 2 * Sp * 2);
    sm.off_tile[TILE_H] = tk(kTileRows * max(Kb, Kh) * 2);
    sm.off_tile[TILE_SLAB0] = tk(kTileRows * 192 * 2);
    sm.off_tile[TILE_SLAB1] = tk(kTileRows * 192 * 2);
    sm.stage_bytes = align_stage(b.max_stage);
    sm.off_ring = o;
    const uint32_t budget = 227 * 1024 - 4096;
    if (o + 2 * sm.stage_bytes > budget) BD_FAIL(BD_ERR_UNSUPPORTED, "tensor-core BPTT: tiles do not fit shared memory");
    sm.nstage = min(8u, (budget - o) / sm.stage_bytes);
    sm.total = o + sm.nstage * sm.stage_bytes + 1024;
  }
  b.finalize_blocks(ba.sm.stage_bytes);
  ba.prog = b.prog;

  const long long ntiles = (f.N + kTileRows - 1) / kTileRows;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned grid = (unsigned)(ntiles < sms ? ntiles : sms);

  // workspace: packed weights | amax | per-CTA scratch
  char* base = static_cast<char*>(ws);
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p_ = base + off; off += (bytes + 255) & ~size_t(255); return p_; };
  uint16_t* wpack = reinterpret_cast<uint16_t*>(take((size_t)b.w_elems * 2));
  unsigned int* amax = reinterpret_cast<unsigned int*>(take(256));
  float* scr_carry = reinterpret_cast<float*>(take((size_t)grid * kTileRows * Kb * 4));
  float* scr_gtot = reinterpret_cast<float*>(take((size_t)grid * kTileRows * Kb * 4));
  // tiled copy of the upstream belief gradients: pays off while the tensor stays L2-resident (an extra
  // HBM round trip of T*N*Be*4 bytes costs more than the coalescing saves beyond that)
  const bool tile_gb = a->g_beliefs && (size_t)f.T * f.N * Be * 4 <= ((size_t)48 << 20);
  float* gbt = tile_gb ? reinterpret_cast<float*>(take((size_t)f.T * ntiles * kTileRows * Kb * 4)) : nullptr;
  float* scr_drv = hb ? reinterpret_cast<float*>(take((size_t)grid * f.T * 2 * kTileRows * 4)) : nullptr;
  if (off > ws_bytes) BD_FAIL(BD_ERR_WORKSPACE, "tensor-core BPTT: workspace %zu < %zu", ws_bytes, off);
  cudaMemsetAsync(amax, 0, 256, s);
  {
    const float* gs[8] = {a->g_beliefs, a->g_states, a->g_means, a->g_stds, a->g_entropy,
                          hb ? hb->g_returns : nullptr, hb ? hb->g_reward : nullptr, hb ? hb->g_value : nullptr};
    const long long tn = (long long)f.T * f.N;
    const long long ns[8] = {tn * Be, tn * S, tn * S, tn * S, tn, tn, tn, tn};
    if (gbt) {     // upstream belief gradients in the kernel's tile layout (coalesced gate-stage reads)
      gb_tile_kernel<<<dim3((unsigned)ntiles, (unsigned)f.T, (unsigned)((Kb + 63) / 64)), 256, 0, s>>>(
          a->g_beliefs, f.N, Be, Kb, ntiles, gbt, amax);
      BD_CUDA_LAUNCH_CHECK();
    }
    AbsmaxJobs jobs{};
    int nj = 0;
    long long nmax = 0;
    for (int i = 0; i < 8; ++i) {
      if (!gs[i] || (i == 0 && gbt)) continue;      // g_beliefs: folded into gb_tile_kernel
      jobs.x[nj] = gs[i]; jobs.n[nj] = ns[i]; ++nj;
      nmax = max(nmax, ns[i]);
    }
    if (nj) {      // one launch for all upstream-gradient tensors
      long long g = (nmax + 255) / 256;
      if (g > 592) g = 592;
      absmax_multi_kernel<<<dim3((unsigned)(g < 1 ? 1 : g), (unsigned)nj), 256, 0, s>>>(jobs, amax);
      BD_CUDA_LAUNCH_CHECK();
    }
  }
  const int fmt = precision == BD_PREC_FP16 ? 0 : 1;
  {
    long long max_img = 0;
    for (int i = 0; i < b.pack.njobs; ++i)
      max_img = max(max_img, (long long)b.pack.job[i].Np * b.pack.job[i].Kp);
    long long pgx = (max_img + 255) / 256;
    if (pgx > 64) pgx = 64;
    dim3 pgrid((unsigned)pgx, (unsigned)b.pack.njobs);
    if (fmt == 0) pack_weights_kernel<0><<<pgrid, 256, 0, s>>>(b.pack, wpack);
    else pack_weights_kernel<1><<<pgrid, 256, 0, s>>>(b.pack, wpack);
    BD_CUDA_LAUNCH_CHECK();
  }
  SavedLayout sl = saved_layout(r, f.T, f.N);
  const char* sb = static_cast<const char*>(f.tc_saved);
  ba.wpack = wpack; ba.N = f.N; ba.T = f.T; ba.prof = nullptr;
  ba.Be = Be; ba.S = S; ba.A = A; ba.Hi = Hi; ba.Kb = Kb; ba.Kh = Kh; ba.Sp = Sp; ba.Ksa = Ksa;
  ba.min_std = r.min_std_dev; ba.cfg = f.actor_cfg;
  ba.sv_gate = reinterpret_cast<const uint16_t*>(sb + sl.off_gate);
  ba.sv_xa = reinterpret_cast<const uint16_t*>(sb + sl.off_xa);
  ba.sv_ha = reinterpret_cast<const uint16_t*>(sb + sl.off_ha);
  ba.stds = f.stds; ba.eps_s = f.eps_s; ba.eps_a = f.eps_a; ba.actions = f.actions;
  ba.actor_raw = f.actor_raw; ba.dent = f.dent;
  ba.g_beliefs = a->g_beliefs; ba.g_states = a->g_states; ba.g_means = a->g_means; ba.g_stds = a->g_stds;
  ba.g_entropy = a->g_entropy;
  ba.d_raw = d_raw; ba.d_prev_state = a->d_prev_state; ba.d_prev_belief = a->d_prev_belief;
  ba.scr_carry = scr_carry; ba.scr_gtot = scr_gtot; ba.amax_bits = amax;
  ba.gbt = gbt;
  if (hb) {
    // the lambda-return adjoint sums up to T upstream terms per row: leave headroom in the fp16 operands
    ba.n_heads = 2; ba.kh_hd = Khd; ba.hd_last = hL - 2; ba.hd_nvalid = head_hidden(*hb->head[0]);
    const HeadsSaved hs = heads_saved_layout(*hb->head[0], f.T, f.N);
    for (int k = 0; k < 2; ++k) {
      ba.w_out[k] = hb->head[k]->layer[hL - 1].w;
      for (int l = 0; l + 1 < hL; ++l)
        ba.sv_hd[k * BD_MAX_LAYERS + l] = reinterpret_cast<const uint16_t*>(static_cast<const char*>(hb->saved) + hs.off[k][l]);
    }
    ba.g_returns = hb->g_returns; ba.g_reward = hb->g_reward; ba.g_value = hb->g_value;
    ba.lr_disc = (float)hb->discount; ba.lr_lam = (float)hb->lambda_;
    ba.scr_drv = scr_drv;
  }
  {
    // next-step inputs the producer warp prefetches into L2 (per tile and time step)
    PrefetchPlan& pf = ba.pf;
    pf.n = 0; pf.reverse = 1;
    auto add = [&](const void* p, long long step_stride, long long tile_stride, size_t bytes) {
      if (!p || pf.n >= 6) return;
      pf.base[pf.n] = static_cast<const char*>(p); pf.step_stride[pf.n] = step_stride;
      pf.tile_stride[pf.n] = tile_stride; pf.bytes[pf.n] = (unsigned int)bytes; ++pf.n;
    };
    const long long nt = ntiles;
    add(ba.sv_gate, nt * 5LL * kTileRows * Kb * 2, 5LL * kTileRows * Kb * 2, (size_t)5 * kTileRows * Kb * 2);
    add(ba.sv_xa, nt * (long long)kTileRows * Kb * 2, (long long)kTileRows * Kb * 2, (size_t)kTileRows * Kb * 2);
    add(ba.sv_ha, nt * (long long)kTileRows * Kh * 2, (long long)kTileRows * Kh * 2, (size_t)kTileRows * Kh * 2);
    if (gbt) add(gbt, nt * (long long)kTileRows * Kb * 4, (long long)kTileRows * Kb * 4, (size_t)kTileRows * Kb * 4);
    else add(a->g_beliefs, f.N * (long long)Be * 4, (long long)kTileRows * Be * 4, (size_t)kTileRows * Be * 4);
    add(a->g_states, f.N * (long long)S * 4, (long long)kTileRows * S * 4, (size_t)kTileRows * S * 4);
    add(f.stds, f.N * (long long)S * 4, (long long)kTileRows * S * 4, (size_t)kTileRows * S * 4);
  }
  ProfScope ps(BD_PROF_BPTT, s);
  if (getenv("BD_TC_PROF") && fmt == 0 && off + kMaxPhases * 64 + 4096 <= ws_bytes) {
    // debug: per-phase cycle counters of CTA 0 (printed by scripts/prof_bptt.py)
    ba.prof = reinterpret_cast<long long*>(base + ((off + 4095) & ~size_t(4095)));
    cudaMemsetAsync(ba.prof, 0, kMaxPhases * 64, s);
    set_smem_attr(bptt_kernel<0, true, 3>, ba.sm.total);
    bptt_kernel<0, true, 3><<<grid, 64 + 3 * 128, ba.sm.total, s>>>(ba);
    static long long* host_prof = nullptr;
    if (!host_prof) cudaMallocHost(&host_prof, kMaxPhases * 64);
    cudaMemcpyAsync(host_prof, ba.prof, kMaxPhases * 64, cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    fprintf(stderr, "bptt phase   iss_dep iss_wwait iss_issue | epi0_wait epi0_work | epi1_wait epi1_work  (cycles/step, CTA0)\n");
    long long tot[7] = {0};
    for (int pi = 0; pi < b.prog.n_phases; ++pi) {
      fprintf(stderr, "p%-2d epi=%d ", pi, (int)b.prog.p[pi].epi);
      for (int k = 0; k < 7; ++k) { fprintf(stderr, "%9lld", host_prof[pi * 8 + k] / f.T); tot[k] += host_prof[pi * 8 + k] / f.T; }
      if (pi < 20) fprintf(stderr, "   | pre %lld planes %lld accwait %lld", host_prof[(20 + pi) * 8] / f.T,
                           host_prof[(20 + pi) * 8 + 1] / f.T, host_prof[(20 + pi) * 8 + 2] / f.T);
      fprintf(stderr, "\n");
    }
    fprintf(stderr, "total     ");
    for (int k = 0; k < 7; ++k) fprintf(stderr, "%9lld", tot[k]);
    fprintf(stderr, "\n");
  } else {
    static const int nparts = getenv("BD_BPTT_PARTS") ? atoi(getenv("BD_BPTT_PARTS")) : 3;   // (4: A/B timing)
#define BD_LAUNCH_BPTT(F, NP)                                                      \
  do {                                                                             \
    set_smem_attr(bptt_kernel<F, false, NP>, ba.sm.total);                         \
    bptt_kernel<F, false, NP><<<grid, 64 + NP * 128, ba.sm.total, s>>>(ba);        \
  } while (0)
    if (fmt == 0) { if (nparts == 3) BD_LAUNCH_BPTT(0, 3); else BD_LAUNCH_BPTT(0, 4); }
    else { if (nparts == 3) BD_LAUNCH_BPTT(1, 3); else BD_LAUNCH_BPTT(1, 4); }
#undef BD_LAUNCH_BPTT
  }
  BD_CUDA_LAUNCH_CHECK();
  return BD_OK;
}

int imagine_bptt(const bd_imagine_bwd_args* a, float* d_raw, void* ws, size_t ws_bytes, int precision,
                 bd_stream_t stream) {
  return imagine_bptt_impl(a, nullptr, d_raw, ws, ws_bytes, precision, stream);
}
int imagine_returns_bptt(const bd_imagine_returns_bwd_args* a, float* d_raw, void* ws, size_t ws_bytes,
                         int precision, bd_stream_t stream) {
  bd_imagine_bwd_args ia{};
  ia.fwd = a->fwd.img;
  ia.g_beliefs = a->g_beliefs; ia.g_states = a->g_states; ia.g_means = a->g_means; ia.g_stds = a->g_stds;
  ia.g_entropy = a->g_entropy; ia.d_prev_state = a->d_prev_state; ia.d_prev_belief = a->d_prev_belief;
  for (int l = 0; l < BD_MAX_LAYERS; ++l) { ia.actor_dw[l] = a->actor_dw[l]; ia.actor_db[l] = a->actor_db[l]; }
  HeadsBwd hb{};
  hb.head[0] = &a->fwd.reward; hb.head[1] = &a->fwd.value; hb.saved = a->fwd.heads_saved;
  hb.g_returns = a->g_returns; hb.g_reward = a->g_reward; hb.g_value = a->g_value;
  hb.discount = a->fwd.discount; hb.lambda_ = a->fwd.lambda_;
  return imagine_bptt_impl(&ia, &hb, d_raw, ws, ws_bytes, precision, stream);
}

// ---------------------------------------------------------------------------------------------
// DenseModel forward (src/models.py:393-408) on the tensor-core engine: T = 1, no recurrence.
// ---------------------------------------------------------------------------------------------
bool mlp_supported(const bd_mlp& m, int k1, int k2, int precision) {
  if (precision != BD_PREC_FP16 && precision != BD_PREC_BF16) return false;
  if (!(m.activation == BD_ACT_ELU || m.activation == BD_ACT_RELU || m.activation == BD_ACT_TANH ||
        m.activation == BD_ACT_IDENTITY)) return false;
  if (k1 + 1 > 256 || k2 > 240 || m.n_layers < 1) return false;
  for (int l = 0; l < m.n_layers; ++l)
    if (m.layer[l].out_features + 1 > 256) return false;
  return true;
}
size_t mlp_pack_bytes(const bd_mlp& m) {
  size_t e = 0;
  for (int l = 0; l < m.n_layers; ++l)
    e += (size_t)r16(m.layer[l].out_features) * (r16(m.layer[l].in_features + 1) + 16);
  return e * 2 + 4096;
}

size_t mlp_saved_bytes(const bd_mlp& m, int64_t rows) {
  size_t per_tile = 0;
  for (int l = 0; l + 1 < m.n_layers; ++l) per_tile += (size_t)kTileRows * r16(m.layer[l].out_features + 1) * 2;
  return (size_t)((rows + kTileRows - 1) / kTileRows) * per_tile + 256;
}

int mlp_forward(const bd_mlp* m, const float* x1, int k1, const float* x2, int k2, int64_t rows,
                float* y, void* ws, size_t ws_bytes, int precision, bd_stream_t stream, void* saved) {
  if (!mlp_supported(*m, k1, k2, precision))
    BD_FAIL(BD_ERR_UNSUPPORTED, "tensor-core mlp_forward: sizes/activation not supported");
  Builder b;
  const int Kp_b = r16(k1 + 1), Ks = k2 > 0 ? r16(k2) : 0;
  int Kp_h = 16, sp_m = 0;
  for (int l = 0; l < m->n_layers; ++l) {
    const bd_linear& L = m->layer[l];
    const bool last = (l == m->n_layers - 1);
    const int n = L.out_features, Np = r16(n);
    const int d = b.dcol();
    if (l == 0) {
      uint32_t wb = b.add_pack(L.w, k1 + k2, 0, n, Np, Kp_b, 0, k1, L.b, k1);
      b.add_gemm(wb, Np, Kp_b, TILE_BCUR, 0, d, 0);
      if (k2 > 0) {
        uint32_t wsx = b.add_pack(L.w, k1 + k2, 0, n, Np, Ks, k1, k2, nullptr, -1);
        b.add_gemm(wsx, Np, Ks, TILE_SA, 0, d, 1);
      }
    } else {
      const int kin = L.in_features, Kp = r16(kin + 1);
      uint32_t w = b.add_pack(L.w, kin, 0, n, Np, Kp, 0, kin, L.b, kin);
      b.chain_gemm(w, Np, Kp, TILE_H, d, sp_m);
    }
    if (last) b.end_phase(EPI_STORE_OUT, 1, n, Np, 0, d, 0, TILE_H);
    else {
      b.end_phase(EPI_ACT_H, 1, n, Np, r16(n + 1), d, saved ? 3 + l : 0, TILE_H);
      sp_m = b.split_last_phase();
      Kp_h = max(Kp_h, r16(n + 1));
    }
  }
  if (!b.ok) BD_FAIL(BD_ERR_UNSUPPORTED, "tensor-core mlp_forward: program too large");
  const size_t pack_bytes = (size_t)b.w_elems * 2;
  if (pack_bytes > ws_bytes)
    BD_FAIL(BD_ERR_WORKSPACE, "tensor-core mlp_forward: workspace %zu < %zu", ws_bytes, pack_bytes);
  RolloutArgs ra{};
  if (!plan_smem(Kp_b, max(Ks, 16), Kp_h, b.max_stage, ra.sm, false))
    BD_FAIL(BD_ERR_UNSUPPORTED, "tensor-core mlp_forward: tiles do not fit shared memory");
  set_programs(ra, b, 1);
  ra.wpack = static_cast<const uint16_t*>(ws);
  ra.N = rows; ra.T = 1; ra.Be = k1; ra.S = k2; ra.A = 0; ra.Hi = 0; ra.J = 0;
  ra.Kp_b = Kp_b; ra.Kp_sa = max(Ks, 16); ra.Kp_h = Kp_h; ra.act = m->activation;
  ra.prev_belief = x1; ra.prev_state = x2; ra.mlp_out = y; ra.has_b1 = 0;
  {   // next tile's input rows -> L2 while the current tile computes
    PrefetchPlan& pf = ra.pf;
    pf.n = 0; pf.reverse = 0;
    pf.base[0] = reinterpret_cast<const char*>(x1); pf.step_stride[0] = 0;
    pf.tile_stride[0] = (long long)kTileRows * k1 * 4; pf.bytes[0] = (unsigned)(kTileRows * k1 * 4); pf.n = 1;
    if (x2 && k2 > 0) {
      pf.base[1] = reinterpret_cast<const char*>(x2); pf.step_stride[1] = 0;
      pf.tile_stride[1] = (long long)kTileRows * k2 * 4; pf.bytes[1] = (unsigned)(kTileRows * k2 * 4); pf.n = 2;
    }
  }
  if (saved) {   // hidden images, layer after layer, each [tiles][128 x Kp_l] (same layout mlp_backward reads)
    const size_t tiles = (size_t)((rows + kTileRows - 1) / kTileRows);
    char* sb = static_cast<char*>(saved);
    for (int l = 0; l + 1 < m->n_layers; ++l) {
      ra.sv_mlp[l] = reinterpret_cast<uint16_t*>(sb);
      sb += tiles * kTileRows * r16(m->layer[l].out_features + 1) * 2;
    }
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  long long max_img = 0;
  for (int i = 0; i < b.pack.njobs; ++i)
    max_img = max(max_img, (long long)b.pack.job[i].Np * b.pack.job[i].Kp);
  long long pgx = (max_img + 255) / 256;
  if (pgx > 64) pgx = 64;
  dim3 pgrid((unsigned)pgx, (unsigned)b.pack.njobs);
  const long long ntiles = (rows + kTileRows - 1) / kTileRows;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned grid = (unsigned)(ntiles < sms ? ntiles : sms);
  const int fmt = precision == BD_PREC_FP16 ? 0 : 1;
  if (fmt == 0) pack_weights_kernel<0><<<pgrid, 256, 0, s>>>(b.pack, static_cast<uint16_t*>(ws));
  else pack_weights_kernel<1><<<pgrid, 256, 0, s>>>(b.pack, static_cast<uint16_t*>(ws));
  BD_CUDA_LAUNCH_CHECK();
  ProfScope ps(BD_PROF_MLP_FWD, s);
  return launch_rollout(fmt, m->activation, false, false, grid, ra, s);
}

// ---------------------------------------------------------------------------------------------
// Two heads on the same rows in ONE launch (the reward and value models of Dreamer's behaviour step on the T*N
// imagined latents, src/dreamer.py:321-322, when the step runs piecewise): one tile prologue for both, and the two
// chains interleaved phase by phase exactly as in the fused rollout (every phase depends on the epilogue two phases
// back), so the MMAs of one head run under the epilogue of the other.  Hidden activations are left as the images
// bd_mlp_backward reads (one buffer per head, bd_mlp_saved_bytes each).
// ---------------------------------------------------------------------------------------------
bool heads_pair_supported(const bd_mlp& reward, const bd_mlp& value, int k1, int k2, int precision) {
  const bd_mlp* hs[2] = {&reward, &value};
  if (k2 <= 0 || reward.n_layers != value.n_layers || reward.activation != value.activation) return false;
  if (reward.n_layers < 2 || 2 * (reward.n_layers - 1) > BD_MAX_LAYERS) return false;
  for (const bd_mlp* m : hs) {
    if (!mlp_supported(*m, k1, k2, precision)) return false;
    if (m->layer[0].in_features != k1 + k2 || m->layer[m->n_layers - 1].out_features != 1) return false;
    const int hh = head_hidden(*m);
    for (int l = 0; l + 1 < m->n_layers; ++l)
      if (m->layer[l].out_features != hh || (l > 0 && m->layer[l].in_features != hh)) return false;
    if (r16(hh + 1) > r16(k1 + 1)) return false;            // the second chain's hidden tile is a belief tile
  }
  if (4 * reward.n_layers + 2 > kMaxGemms || 2 * reward.n_layers > kMaxPhases) return false;
  bd_rssm r{};
  r.belief_size = k1; r.state_size = k2; r.action_size = 0; r.hidden_size = max(head_hidden(reward), head_hidden(value));
  return rollout_tiles_fit(r, r.hidden_size);
}
size_t heads_pair_pack_bytes(const bd_mlp& reward, const bd_mlp& value) {
  return mlp_pack_bytes(reward) + mlp_pack_bytes(value) + 2 * 16 * 272 * 2 + 4096;
}
int heads_pair_forward(const bd_mlp* reward, const bd_mlp* value, const float* x1, int k1, const float* x2, int k2,
                       int64_t rows, float* y_reward, float* y_value, void* saved_reward, void* saved_value,
                       void* ws, size_t ws_bytes, int precision, bd_stream_t stream) {
  if (!heads_pair_supported(*reward, *value, k1, k2, precision))
    BD_FAIL(BD_ERR_UNSUPPORTED, "tensor-core heads_pair_forward: configuration not supported");
  bd_rssm r{};
  r.belief_size = k1; r.state_size = k2;
  HeadsFwd hd{};
  hd.head[0] = reward; hd.head[1] = value;
  hd.saved = (saved_reward && saved_value) ? saved_reward : nullptr;      // (only tested for null below)
  const int Kp_b = r16(k1 + 1), Ks = r16(k2);
  const int Kp_h = max(r16(head_hidden(*reward) + 1), r16(head_hidden(*value) + 1));
  const int L = reward->n_layers;
  Builder b;
  add_head_phases(b, r, hd, 0, 1, Kp_b, Ks, TILE_BCUR, TILE_BNXT, true);
  if (!b.ok) BD_FAIL(BD_ERR_UNSUPPORTED, "tensor-core heads_pair_forward: program too large");
  const size_t pack_bytes = (size_t)b.w_elems * 2;
  if (pack_bytes > ws_bytes)
    BD_FAIL(BD_ERR_WORKSPACE, "tensor-core heads_pair_forward: workspace %zu < %zu", ws_bytes, pack_bytes);
  RolloutArgs ra{};
  if (!plan_smem(Kp_b, max(Ks, 16), Kp_h, b.max_stage, ra.sm, true))
    BD_FAIL(BD_ERR_UNSUPPORTED, "tensor-core heads_pair_forward: tiles do not fit shared memory");
  set_programs(ra, b, 1);
  ra.wpack = static_cast<const uint16_t*>(ws);
  ra.N = rows; ra.T = 1; ra.Be = k1; ra.S = k2; ra.A = 0; ra.Hi = 0; ra.J = 0;
  ra.Kp_b = Kp_b; ra.Kp_sa = max(Ks, 16); ra.Kp_h = Kp_h; ra.act = reward->activation;
  ra.prev_belief = x1; ra.prev_state = x2; ra.has_b1 = 1;
  ra.head_out[0] = y_reward; ra.head_out[1] = y_value;
  {
    PrefetchPlan& pf = ra.pf;
    pf.reverse = 0;
    pf.base[0] = reinterpret_cast<const char*>(x1); pf.step_stride[0] = 0;
    pf.tile_stride[0] = (long long)kTileRows * k1 * 4; pf.bytes[0] = (unsigned)(kTileRows * k1 * 4);
    pf.base[1] = reinterpret_cast<const char*>(x2); pf.step_stride[1] = 0;
    pf.tile_stride[1] = (long long)kTileRows * k2 * 4; pf.bytes[1] = (unsigned)(kTileRows * k2 * 4);
    pf.n = 2;
  }
  if (hd.saved) {   // per head: hidden images layer after layer, [tiles][128 x Kp_l] (what mlp_backward reads)
    const size_t tiles = (size_t)((rows + kTileRows - 1) / kTileRows);
    void* sv[2] = {saved_reward, saved_value};
    for (int k = 0; k < 2; ++k) {
      char* sb = static_cast<char*>(sv[k]);
      const bd_mlp& m = *hd.head[k];
      for (int l = 0; l + 1 < L; ++l) {
        ra.sv_mlp[k * (L - 1) + l] = reinterpret_cast<uint16_t*>(sb);
        sb += tiles * kTileRows * r16(m.layer[l].out_features + 1) * 2;
      }
    }
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  long long max_img = 0;
  for (int i = 0; i < b.pack.njobs; ++i)
    max_img = max(max_img, (long long)b.pack.job[i].Np * b.pack.job[i].Kp);
  long long pgx = (max_img + 255) / 256;
  if (pgx > 64) pgx = 64;
  dim3 pgrid((unsigned)pgx, (unsigned)b.pack.njobs);
  const long long ntiles = (rows + kTileRows - 1) / kTileRows;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned grid = (unsigned)(ntiles < sms ? ntiles : sms);
  const int fmt = precision == BD_PREC_FP16 ? 0 : 1;
  if (fmt == 0) pack_weights_kernel<0><<<pgrid, 256, 0, s>>>(b.pack, static_cast<uint16_t*>(ws));
  else pack_weights_kernel<1><<<pgrid, 256, 0, s>>>(b.pack, static_cast<uint16_t*>(ws));
  BD_CUDA_LAUNCH_CHECK();
  ProfScope ps(BD_PROF_MLP_FWD, s);
  return launch_rollout(fmt, reward->activation, false, false, grid, ra, s);
}

}  // namespace tc
}  // namespace bd
