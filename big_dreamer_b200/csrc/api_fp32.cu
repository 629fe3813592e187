// fp32 check-mode implementations behind the C ABI (include/bd_b200.h): host-side
// orchestration of the FFMA kernels in fp32_kernels.cuh.  Rows are processed in chunks
// sized to the caller's workspace; all rows are independent.
#include <map>
#include <mutex>
#include "api_internal.h"
#include "fp32_kernels.cuh"
#include "observe_persist.cuh"
#include <stdlib.h>

namespace bd {
namespace f32 {

static inline cudaStream_t S(bd_stream_t s) { return static_cast<cudaStream_t>(s); }

// -------------------------------------------------------------------------------------
// MLP
// -------------------------------------------------------------------------------------
int mlp_max_width(const bd_mlp& m) {
  int w = 1;
  for (int l = 0; l < m.n_layers; ++l) w = max(w, m.layer[l].out_features);
  return w;
}

// forward keeping hidden activations: hid[l] (rows, out_l) for l < n_layers-1
static int mlp_forward_rows(const bd_mlp& m, const float* x1, int k1, long long ld1,
                            const float* x2, int k2, long long ld2, int rows, float* const* hid,
                            float* y, long long ldy, cudaStream_t s) {
  const float* in = nullptr;
  int in_w = 0;
  for (int l = 0; l < m.n_layers; ++l) {
    const bd_linear& L = m.layer[l];
    const bool last = (l == m.n_layers - 1);
    float* out = last ? y : hid[l];
    long long ldo = last ? ldy : L.out_features;
    int act = last ? BD_ACT_IDENTITY : m.activation;
    if (l == 0) BD_TRY(linear_fwd(L, act, x1, k1, ld1, x2, k2, ld2, nullptr, rows, out, ldo, s));
    else BD_TRY(linear_fwd(L, act, in, in_w, in_w, nullptr, 0, 0, nullptr, rows, out, ldo, s));
    in = out;
    in_w = L.out_features;
  }
  return BD_OK;
}

// backward given all hidden activations.  dy (rows,out) with leading dim lddy.
// dbuf0/dbuf1: (rows, maxw) ping-pong.  dw/db accumulate; dx1/dx2 overwritten (or += if beta).
static int mlp_backward_rows(const bd_mlp& m, const float* x1, int k1, long long ld1,
                             const float* x2, int k2, long long ld2, int rows,
                             float* const* hid, const float* dy, long long lddy, float* dbuf0,
                             float* dbuf1, float* const* dw, float* const* db, float* dx1,
                             long long lddx1, float* dx2, long long lddx2, cudaStream_t s) {
  const float* dcur = dy;
  long long ldd = lddy;
  float* bufs[2] = {dbuf0, dbuf1};
  int which = 0;
  for (int l = m.n_layers - 1; l >= 0; --l) {
    const bd_linear& L = m.layer[l];
    if (db && db[l]) BD_TRY(bias_grad(dcur, L.out_features, ldd, rows, db[l], s));
    if (dw && dw[l]) {
      if (l == 0) {
        BD_TRY(linear_wgrad(dcur, L.out_features, ldd, x1, k1, ld1, rows, dw[l], L.in_features, 0, s));
        if (k2 > 0)
          BD_TRY(linear_wgrad(dcur, L.out_features, ldd, x2, k2, ld2, rows, dw[l], L.in_features, k1, s));
      } else {
        BD_TRY(linear_wgrad(dcur, L.out_features, ldd, hid[l - 1], L.in_features, L.in_features,
                            rows, dw[l], L.in_features, 0, s));
      }
    }
    if (l > 0) {
      float* dprev = bufs[which];
      which ^= 1;
      BD_TRY(linear_dgrad(dcur, L.out_features, ldd, L.w, L.in_features, 0, L.in_features, rows,
                          dprev, L.in_features, m.activation, hid[l - 1], L.in_features, 0, s));
      dcur = dprev;
      ldd = L.in_features;
    } else {
      if (dx1) BD_TRY(linear_dgrad(dcur, L.out_features, ldd, L.w, L.in_features, 0, k1, rows, dx1,
                                   lddx1, BD_ACT_IDENTITY, nullptr, 0, 0, s));
      if (dx2 && k2 > 0)
        BD_TRY(linear_dgrad(dcur, L.out_features, ldd, L.w, L.in_features, k1, k2, rows, dx2,
                            lddx2, BD_ACT_IDENTITY, nullptr, 0, 0, s));
    }
  }
  return BD_OK;
}

static size_t mlp_row_floats(const bd_mlp& m, bool backward) {
  size_t w = mlp_max_width(m);
  size_t hid = 0;
  for (int l = 0; l + 1 < m.n_layers; ++l) hid += m.layer[l].out_features;
  return backward ? hid + 2 * w : 2 * w;
}

size_t mlp_workspace_bytes(const bd_mlp* m, int64_t rows, int backward) {
  int64_t r = rows < kMaxChunkRows ? rows : kMaxChunkRows;
  return (size_t)(r > 0 ? r : 1) * mlp_row_floats(*m, backward != 0) * sizeof(float) + kSlackBytes;
}

int check_mlp(const bd_mlp& m, int in_features) {
  BD_CHECK_ARG(m.n_layers >= 1 && m.n_layers <= BD_MAX_LAYERS, "mlp: n_layers=%d out of range", m.n_layers);
  BD_CHECK_ARG(valid_act(m.activation), "mlp: unsupported activation %d", m.activation);
  int in = in_features;
  for (int l = 0; l < m.n_layers; ++l) {
    BD_CHECK_ARG(m.layer[l].w && m.layer[l].b, "mlp: layer %d has null parameters", l);
    BD_CHECK_ARG(m.layer[l].in_features == in, "mlp: layer %d in_features %d != %d", l,
                 m.layer[l].in_features, in);
    in = m.layer[l].out_features;
  }
  return BD_OK;
}

static int chunk_rows_for(size_t ws_bytes, size_t row_floats, int64_t rows, int* out) {
  size_t per = row_floats * sizeof(float);
  int64_t fit = (int64_t)((ws_bytes > kSlackBytes ? ws_bytes - kSlackBytes : 0) / (per ? per : 1));
  fit = fit / 64 * 64 > 0 ? fit / 64 * 64 : fit;   // keep chunks tile-aligned when possible
  if (fit > kMaxChunkRows) fit = kMaxChunkRows;
  if (fit > rows) fit = rows;
  if (fit < 1) BD_FAIL(BD_ERR_WORKSPACE, "workspace too small: %zu bytes, need >= %zu per row",
                       ws_bytes, per);
  *out = (int)fit;
  return BD_OK;
}

int mlp_forward(const bd_mlp* m, const float* x1, int k1, const float* x2, int k2, int64_t rows,
                float* y, void* ws, size_t ws_bytes, bd_stream_t stream) {
  BD_TRY(check_mlp(*m, k1 + k2));
  if (rows == 0) return BD_OK;
  const int out = m->layer[m->n_layers - 1].out_features;
  int chunk;
  BD_TRY(chunk_rows_for(ws_bytes, mlp_row_floats(*m, false), rows, &chunk));
  const int w = mlp_max_width(*m);
  for (int64_t r0 = 0; r0 < rows; r0 += chunk) {
    int nr = (int)((rows - r0) < chunk ? (rows - r0) : chunk);
    Arena ar(ws, ws_bytes);
    float* a = ar.f32((size_t)chunk * w);
    float* b = ar.f32((size_t)chunk * w);
    if (!ar.ok()) BD_FAIL(BD_ERR_WORKSPACE, "mlp_forward: workspace carve failed");
    float* hid[BD_MAX_LAYERS];
    for (int l = 0; l < BD_MAX_LAYERS; ++l) hid[l] = (l & 1) ? b : a;
    BD_TRY(mlp_forward_rows(*m, x1 + r0 * k1, k1, k1, x2 ? x2 + r0 * k2 : nullptr, k2, k2, nr, hid,
                            y + r0 * out, out, S(stream)));
  }
  return BD_OK;
}

int mlp_backward(const bd_mlp* m, const bd_mlp_bwd_args* a, void* ws, size_t ws_bytes,
                 bd_stream_t stream) {
  BD_TRY(check_mlp(*m, a->k1 + a->k2));
  if (a->rows == 0) return BD_OK;
  const int out = m->layer[m->n_layers - 1].out_features;
  int chunk;
  BD_TRY(chunk_rows_for(ws_bytes, mlp_row_floats(*m, true), a->rows, &chunk));
  const int w = mlp_max_width(*m);
  for (int64_t r0 = 0; r0 < a->rows; r0 += chunk) {
    int nr = (int)((a->rows - r0) < chunk ? (a->rows - r0) : chunk);
    Arena ar(ws, ws_bytes);
    float* hid[BD_MAX_LAYERS] = {nullptr};
    for (int l = 0; l + 1 < m->n_layers; ++l) hid[l] = ar.f32((size_t)chunk * m->layer[l].out_features);
    float* d0 = ar.f32((size_t)chunk * w);
    float* d1 = ar.f32((size_t)chunk * w);
    if (!ar.ok()) BD_FAIL(BD_ERR_WORKSPACE, "mlp_backward: workspace carve failed");
    const float* x1 = a->x1 + r0 * a->k1;
    const float* x2 = a->x2 ? a->x2 + r0 * a->k2 : nullptr;
    // recompute hidden activations (the last layer's output is not needed: write it to d0)
    BD_TRY(mlp_forward_rows(*m, x1, a->k1, a->k1, x2, a->k2, a->k2, nr, hid, d0, out, S(stream)));
    BD_TRY(mlp_backward_rows(*m, x1, a->k1, a->k1, x2, a->k2, a->k2, nr, hid, a->dy + r0 * out, out,
                             d0, d1, a->dw, a->db, a->dx1 ? a->dx1 + r0 * a->k1 : nullptr, a->k1,
                             a->dx2 ? a->dx2 + r0 * a->k2 : nullptr, a->k2, S(stream)));
  }
  return BD_OK;
}

// -------------------------------------------------------------------------------------
// lambda_return
// -------------------------------------------------------------------------------------
int lambda_return_forward(const float* reward, const float* value, const float* bootstrap, int T,
                          int64_t N, double discount, double lambda_, float* returns,
                          bd_stream_t stream) {
  BD_CHECK_ARG(T >= 1 && N >= 0, "lambda_return: bad T/N");
  if (N == 0) return BD_OK;
  lambda_return_fwd_kernel<<<grid1d(N), 256, 0, S(stream)>>>(
      reward, value, bootstrap, T, N, (float)discount, (float)lambda_, (float)(1.0 - lambda_), returns);
  BD_CUDA_LAUNCH_CHECK();
  return BD_OK;
}
int lambda_return_backward(const float* d_returns, int T, int64_t N, double discount,
                           double lambda_, float* d_reward, float* d_value, float* d_bootstrap,
                           bd_stream_t stream) {
  BD_CHECK_ARG(T >= 1 && N >= 0, "lambda_return: bad T/N");
  if (N == 0) return BD_OK;
  lambda_return_bwd_kernel<<<grid1d(N), 256, 0, S(stream)>>>(d_returns, T, N, (float)discount,
                                                             (float)lambda_, d_reward, d_value,
                                                             d_bootstrap);
  BD_CUDA_LAUNCH_CHECK();
  return BD_OK;
}

// -------------------------------------------------------------------------------------
// KL loss of the dynamics update (src/planet.py:288-308, src/dreamer.py:111-146)
// -------------------------------------------------------------------------------------
int kl_loss_forward(const float* post_mean, const float* post_std, const float* prior_mean,
                    const float* prior_std, int64_t rows, int Sz, const float* free_nats, double balance,
                    float* div, float* loss, bd_stream_t stream) {
  BD_CHECK_ARG(rows >= 1 && Sz >= 1, "kl_loss: bad rows/S");
  BD_CHECK_ARG(post_mean && post_std && prior_mean && prior_std && free_nats && div && loss, "kl_loss: null pointer");
  BD_CHECK_ARG(balance < 0.0 || balance <= 1.0, "kl_loss: kl_balance must be -1 or in [0, 1]");
  kl_rows_kernel<<<grid1d(rows * 32), 256, 0, S(stream)>>>(post_mean, post_std, prior_mean, prior_std, rows, Sz, div);
  BD_CUDA_LAUNCH_CHECK();
  kl_finish_kernel<<<1, 1024, 0, S(stream)>>>(div, rows, Sz, free_nats, balance >= 0.0 ? 1 : 0, loss);
  BD_CUDA_LAUNCH_CHECK();
  return BD_OK;
}
int kl_loss_backward(const float* post_mean, const float* post_std, const float* prior_mean,
                     const float* prior_std, int64_t rows, int Sz, const float* free_nats, double balance,
                     const float* div, const float* loss, const float* g_loss, float* d_post_mean,
                     float* d_post_std, float* d_prior_mean, float* d_prior_std, bd_stream_t stream) {
  BD_CHECK_ARG(rows >= 1 && Sz >= 1, "kl_loss: bad rows/S");
  BD_CHECK_ARG(post_mean && post_std && prior_mean && prior_std && free_nats && div && loss && g_loss,
               "kl_loss backward: null pointer");
  kl_bwd_kernel<<<grid1d(rows * Sz), 256, 0, S(stream)>>>(post_mean, post_std, prior_mean, prior_std, div, loss,
                                                         free_nats, g_loss, rows, Sz, (float)balance, d_post_mean,
                                                         d_post_std, d_prior_mean, d_prior_std);
  BD_CUDA_LAUNCH_CHECK();
  return BD_OK;
}

// -------------------------------------------------------------------------------------
// One RSSM transition step, forward and backward-with-recompute
// -------------------------------------------------------------------------------------
int check_rssm(const bd_rssm& r, bool need_post) {
  BD_CHECK_ARG(r.belief_size > 0 && r.state_size > 0 && r.action_size > 0 && r.hidden_size > 0,
               "rssm: non-positive size");
  BD_CHECK_ARG(valid_act(r.activation), "rssm: unsupported activation %d", r.activation);
  BD_CHECK_ARG(r.embed.w && r.embed.b && r.w_ih && r.w_hh && r.b_ih && r.b_hh && r.prior1.w &&
               r.prior1.b && r.prior2.w && r.prior2.b, "rssm: null parameter pointer");
  BD_CHECK_ARG(r.embed.in_features == r.state_size + r.action_size &&
               r.embed.out_features == r.belief_size, "rssm: embed shape");
  BD_CHECK_ARG(r.prior1.in_features == r.belief_size && r.prior1.out_features == r.hidden_size &&
               r.prior2.in_features == r.hidden_size && r.prior2.out_features == 2 * r.state_size,
               "rssm: prior shape");
  if (need_post) {
    BD_CHECK_ARG(r.post1.w && r.post1.b && r.post2.w && r.post2.b, "rssm: null posterior parameters");
    BD_CHECK_ARG(r.post1.in_features == r.belief_size + r.embedding_size &&
                 r.post1.out_features == r.hidden_size && r.post2.in_features == r.hidden_size &&
                 r.post2.out_features == 2 * r.state_size, "rssm: posterior shape");
  }
  return BD_OK;
}

struct StepBuf {          // per-chunk scratch of one transition step
  float *x, *gi, *gh, *h, *pre, *hq, *preq;                       // forward / recompute
  float *d_pre, *d_h, *Gb, *d_gi, *d_gh, *carry_b, *dx, *dsa, *carry_s, *d_preq, *d_hq;  // backward
};
static size_t step_row_floats(const bd_rssm& r, bool observe, bool backward) {
  size_t Be = r.belief_size, Hi = r.hidden_size, S = r.state_size, A = r.action_size;
  size_t f = Be + 6 * Be + Hi + 2 * S + (observe ? Hi + 2 * S : 0);
  if (backward) f += 2 * S + Hi + Be + 6 * Be + Be + Be + (S + A) + S + (observe ? 2 * S + Hi : 0);
  return f;
}
static bool carve_step(Arena& ar, const bd_rssm& r, size_t rows, bool observe, bool backward,
                       StepBuf& b) {
  size_t Be = r.belief_size, Hi = r.hidden_size, S = r.state_size, A = r.action_size;
  b = StepBuf{};
  b.x = ar.f32(rows * Be); b.gi = ar.f32(rows * 3 * Be); b.gh = ar.f32(rows * 3 * Be);
  b.h = ar.f32(rows * Hi); b.pre = ar.f32(rows * 2 * S);
  if (observe) { b.hq = ar.f32(rows * Hi); b.preq = ar.f32(rows * 2 * S); }
  if (backward) {
    b.d_pre = ar.f32(rows * 2 * S); b.d_h = ar.f32(rows * Hi); b.Gb = ar.f32(rows * Be);
    b.d_gi = ar.f32(rows * 3 * Be); b.d_gh = ar.f32(rows * 3 * Be); b.carry_b = ar.f32(rows * Be);
    b.dx = ar.f32(rows * Be); b.dsa = ar.f32(rows * (S + A)); b.carry_s = ar.f32(rows * S);
    if (observe) { b.d_preq = ar.f32(rows * 2 * S); b.d_hq = ar.f32(rows * Hi); }
  }
  return ar.ok();
}

// forward of embed -> GRU -> prior (-> posterior).  Pointers are already offset to the chunk.
static int step_forward(const bd_rssm& r, const StepBuf& w, int rows, const float* s_prev,
                        const float* nonterm, const float* action, const float* b_prev,
                        const float* eps_prior, const float* emb, const float* eps_post,
                        float* b_new, float* prior_s, float* prior_m, float* prior_sd,
                        float* post_s, float* post_m, float* post_sd, cudaStream_t s) {
  const int Be = r.belief_size, Hi = r.hidden_size, Sz = r.state_size, A = r.action_size,
            E = r.embedding_size;
  // hidden = act(W_sa [s * nonterminal ; a] + b)                       src/models.py:241-251
  BD_TRY(linear_fwd(r.embed, r.activation, s_prev, Sz, Sz, action, A, A, nonterm, rows, w.x, Be, s));
  // GRUCell                                                              src/models.py:252
  BD_TRY(matmul_nt_bias(w.x, Be, Be, r.w_ih, r.b_ih, 3 * Be, rows, w.gi, 3 * Be, s));
  BD_TRY(matmul_nt_bias(b_prev, Be, Be, r.w_hh, r.b_hh, 3 * Be, rows, w.gh, 3 * Be, s));
  gru_gate_fwd_kernel<<<grid1d((long long)rows * Be), 256, 0, s>>>(w.gi, w.gh, b_prev, b_new,
                                                                    (long long)rows * Be, Be);
  BD_CUDA_LAUNCH_CHECK();
  // prior                                                                src/models.py:256, 70-73
  BD_TRY(linear_fwd(r.prior1, r.activation, b_new, Be, Be, nullptr, 0, 0, nullptr, rows, w.h, Hi, s));
  BD_TRY(linear_fwd(r.prior2, BD_ACT_IDENTITY, w.h, Hi, Hi, nullptr, 0, 0, nullptr, rows, w.pre, 2 * Sz, s));
  if (prior_s) {
    belief_sample_fwd_kernel<<<grid1d((long long)rows * Sz), 256, 0, s>>>(
        w.pre, eps_prior, r.min_std_dev, prior_s, prior_m, prior_sd, (long long)rows * Sz, Sz);
    BD_CUDA_LAUNCH_CHECK();
  }
  if (emb) {  // posterior                                                src/models.py:262-269
    BD_TRY(linear_fwd(r.post1, r.activation, b_new, Be, Be, emb, E, E, nullptr, rows, w.hq, Hi, s));
    BD_TRY(linear_fwd(r.post2, BD_ACT_IDENTITY, w.hq, Hi, Hi, nullptr, 0, 0, nullptr, rows, w.preq, 2 * Sz, s));
    if (post_s) {
      belief_sample_fwd_kernel<<<grid1d((long long)rows * Sz), 256, 0, s>>>(
          w.preq, eps_post, r.min_std_dev, post_s, post_m, post_sd, (long long)rows * Sz, Sz);
      BD_CUDA_LAUNCH_CHECK();
    }
  }
  return BD_OK;
}

// backward of one step.  On entry w.Gb holds dL/d b_new (total), and the state gradients are
// given as up to two addends each.  On exit: w.carry_b = dL/d b_prev, w.dsa = dL/d [s*nt ; a]
// (rows, S+A).  Recomputes the forward internals first.  Parameter grads accumulate into G.
static int step_backward(const bd_rssm& r, const StepBuf& w, int rows, const float* s_prev,
                         const float* nonterm, const float* action, const float* b_prev,
                         const float* b_new, const float* eps_prior, const float* emb,
                         const float* eps_post, const float* g_prior_s, const float* g_prior_s2,
                         const float* g_prior_m, const float* g_prior_sd, const float* g_post_s,
                         const float* g_post_s2, const float* g_post_m, const float* g_post_sd,
                         float* d_emb, const bd_rssm_grads* G, cudaStream_t s) {
  const int Be = r.belief_size, Hi = r.hidden_size, Sz = r.state_size, A = r.action_size,
            E = r.embedding_size;
  const long long nS = (long long)rows * Sz, nB = (long long)rows * Be;
  // ---- recompute (b_new is read from the saved output; it is bit-identical)
  BD_TRY(linear_fwd(r.embed, r.activation, s_prev, Sz, Sz, action, A, A, nonterm, rows, w.x, Be, s));
  BD_TRY(matmul_nt_bias(w.x, Be, Be, r.w_ih, r.b_ih, 3 * Be, rows, w.gi, 3 * Be, s));
  BD_TRY(matmul_nt_bias(b_prev, Be, Be, r.w_hh, r.b_hh, 3 * Be, rows, w.gh, 3 * Be, s));
  BD_TRY(linear_fwd(r.prior1, r.activation, b_new, Be, Be, nullptr, 0, 0, nullptr, rows, w.h, Hi, s));
  BD_TRY(linear_fwd(r.prior2, BD_ACT_IDENTITY, w.h, Hi, Hi, nullptr, 0, 0, nullptr, rows, w.pre, 2 * Sz, s));
  // ---- posterior
  if (emb) {
    BD_TRY(linear_fwd(r.post1, r.activation, b_new, Be, Be, emb, E, E, nullptr, rows, w.hq, Hi, s));
    BD_TRY(linear_fwd(r.post2, BD_ACT_IDENTITY, w.hq, Hi, Hi, nullptr, 0, 0, nullptr, rows, w.preq, 2 * Sz, s));
    belief_sample_bwd_kernel<<<grid1d(nS), 256, 0, s>>>(w.preq, eps_post, g_post_s, g_post_s2,
                                                        g_post_m, g_post_sd, w.d_preq, nS, Sz);
    BD_CUDA_LAUNCH_CHECK();
    if (G && G->post2_b) BD_TRY(bias_grad(w.d_preq, 2 * Sz, 2 * Sz, rows, G->post2_b, s));
    if (G && G->post2_w) BD_TRY(linear_wgrad(w.d_preq, 2 * Sz, 2 * Sz, w.hq, Hi, Hi, rows, G->post2_w, Hi, 0, s));
    BD_TRY(linear_dgrad(w.d_preq, 2 * Sz, 2 * Sz, r.post2.w, Hi, 0, Hi, rows, w.d_hq, Hi,
                        r.activation, w.hq, Hi, 0, s));
    if (G && G->post1_b) BD_TRY(bias_grad(w.d_hq, Hi, Hi, rows, G->post1_b, s));
    if (G && G->post1_w) {
      BD_TRY(linear_wgrad(w.d_hq, Hi, Hi, b_new, Be, Be, rows, G->post1_w, Be + E, 0, s));
      BD_TRY(linear_wgrad(w.d_hq, Hi, Hi, emb, E, E, rows, G->post1_w, Be + E, Be, s));
    }
    BD_TRY(linear_dgrad(w.d_hq, Hi, Hi, r.post1.w, Be + E, 0, Be, rows, w.Gb, Be, BD_ACT_IDENTITY,
                        nullptr, 0, 1, s));
    if (d_emb) BD_TRY(linear_dgrad(w.d_hq, Hi, Hi, r.post1.w, Be + E, Be, E, rows, d_emb, E,
                                   BD_ACT_IDENTITY, nullptr, 0, 0, s));
  }
  // ---- prior
  belief_sample_bwd_kernel<<<grid1d(nS), 256, 0, s>>>(w.pre, eps_prior, g_prior_s, g_prior_s2,
                                                      g_prior_m, g_prior_sd, w.d_pre, nS, Sz);
  BD_CUDA_LAUNCH_CHECK();
  if (G && G->prior2_b) BD_TRY(bias_grad(w.d_pre, 2 * Sz, 2 * Sz, rows, G->prior2_b, s));
  if (G && G->prior2_w) BD_TRY(linear_wgrad(w.d_pre, 2 * Sz, 2 * Sz, w.h, Hi, Hi, rows, G->prior2_w, Hi, 0, s));
  BD_TRY(linear_dgrad(w.d_pre, 2 * Sz, 2 * Sz, r.prior2.w, Hi, 0, Hi, rows, w.d_h, Hi, r.activation,
                      w.h, Hi, 0, s));
  if (G && G->prior1_b) BD_TRY(bias_grad(w.d_h, Hi, Hi, rows, G->prior1_b, s));
  if (G && G->prior1_w) BD_TRY(linear_wgrad(w.d_h, Hi, Hi, b_new, Be, Be, rows, G->prior1_w, Be, 0, s));
  BD_TRY(linear_dgrad(w.d_h, Hi, Hi, r.prior1.w, Be, 0, Be, rows, w.Gb, Be, BD_ACT_IDENTITY, nullptr,
                      0, 1, s));
  // ---- GRU gates
  gru_gate_bwd_kernel<<<grid1d(nB), 256, 0, s>>>(w.gi, w.gh, b_prev, w.Gb, w.d_gi, w.d_gh,
                                                 w.carry_b, nB, Be);
  BD_CUDA_LAUNCH_CHECK();
  if (G && G->b_ih) BD_TRY(bias_grad(w.d_gi, 3 * Be, 3 * Be, rows, G->b_ih, s));
  if (G && G->b_hh) BD_TRY(bias_grad(w.d_gh, 3 * Be, 3 * Be, rows, G->b_hh, s));
  if (G && G->w_ih) BD_TRY(linear_wgrad(w.d_gi, 3 * Be, 3 * Be, w.x, Be, Be, rows, G->w_ih, Be, 0, s));
  if (G && G->w_hh) BD_TRY(linear_wgrad(w.d_gh, 3 * Be, 3 * Be, b_prev, Be, Be, rows, G->w_hh, Be, 0, s));
  BD_TRY(linear_dgrad(w.d_gi, 3 * Be, 3 * Be, r.w_ih, Be, 0, Be, rows, w.dx, Be, r.activation, w.x,
                      Be, 0, s));
  BD_TRY(linear_dgrad(w.d_gh, 3 * Be, 3 * Be, r.w_hh, Be, 0, Be, rows, w.carry_b, Be,
                      BD_ACT_IDENTITY, nullptr, 0, 1, s));
  // ---- embed
  if (G && G->embed_b) BD_TRY(bias_grad(w.dx, Be, Be, rows, G->embed_b, s));
  if (G && G->embed_w) {
    // X1 = s_prev * nonterminal: wgrad needs the masked state; build it in w.carry_s scratch
    const float* sx = s_prev;
    if (nonterm) {
      slice_cols_kernel<<<grid1d(nS), 256, 0, s>>>(s_prev, Sz, 0, Sz, nonterm, w.carry_s, rows);
      BD_CUDA_LAUNCH_CHECK();
      sx = w.carry_s;
    }
    BD_TRY(linear_wgrad(w.dx, Be, Be, sx, Sz, Sz, rows, G->embed_w, Sz + A, 0, s));
    BD_TRY(linear_wgrad(w.dx, Be, Be, action, A, A, rows, G->embed_w, Sz + A, Sz, s));
  }
  BD_TRY(linear_dgrad(w.dx, Be, Be, r.embed.w, Sz + A, 0, Sz + A, rows, w.dsa, Sz + A,
                      BD_ACT_IDENTITY, nullptr, 0, 0, s));
  // carry_s = d s_prev = dsa[:, :S] * nonterminal
  slice_cols_kernel<<<grid1d(nS), 256, 0, s>>>(w.dsa, Sz + A, 0, Sz, nonterm, w.carry_s, rows);
  BD_CUDA_LAUNCH_CHECK();
  return BD_OK;
}

// -------------------------------------------------------------------------------------
// TransitionModel.forward / backward
// -------------------------------------------------------------------------------------
// Weight / bias gradients of TransitionModel.forward batched over time: instead of 10 rank-B wgrad
// GEMMs + 7 column sums PER STEP (B = 50 rows in the observe pass: 17 tiny launches x 49 steps), the
// backward keeps every step's wgrad operands (L x B rows each) and runs ONE GEMM / column sum per
// parameter after the time loop.  Used when the whole batch is one chunk and the buffers are small.
constexpr size_t kTimeBatchMaxBytes = (size_t)512 << 20;
static size_t time_batch_row_floats(const bd_rssm& r, bool observe) {
  size_t Be = r.belief_size, Hi = r.hidden_size, S = r.state_size;
  // x, h, d_pre, d_h, d_gi, d_gh, dx, masked previous state (+ hq, d_preq, d_hq)
  return Be + Hi + 2 * S + Hi + 3 * Be + 3 * Be + Be + S + (observe ? Hi + 2 * S + Hi : 0);
}
static size_t time_batch_bytes(const bd_rssm& r, int L, int64_t B, bool observe) {
  if (B <= 0 || B > kMaxChunkRows) return 0;
  size_t b = (size_t)L * B * time_batch_row_floats(r, observe) * sizeof(float);
  return b <= kTimeBatchMaxBytes ? b + 4096 : 0;
}

// the wgrad operands of every step, time-major (L*B rows each)
struct TimeBatch {
  float *tx, *th, *tdpre, *tdh, *tdgi, *tdgh, *tdx, *tsx, *thq, *tdpreq, *tdhq;
};
// one GEMM / column sum per parameter over all L*B rows (+ d embeddings)
static int time_batched_param_grads(const bd_rssm& r, const bd_transition_args& f,
                                    const bd_transition_bwd_args& a, const TimeBatch& tb, cudaStream_t s) {
  const bd_rssm_grads& G = a.grads;
  const long long Be = r.belief_size, Sz = r.state_size, A = r.action_size, E = r.embedding_size,
                  Hi = r.hidden_size, B = f.B;
  const bool observe = f.embeddings != nullptr;
  const int n = (int)((size_t)f.L * B), nB = (int)B;
  if (observe) {
    if (G.post2_b) BD_TRY(bias_grad(tb.tdpreq, 2 * Sz, 2 * Sz, n, G.post2_b, s));
    if (G.post2_w) BD_TRY(linear_wgrad(tb.tdpreq, 2 * Sz, 2 * Sz, tb.thq, Hi, Hi, n, G.post2_w, Hi, 0, s));
    if (G.post1_b) BD_TRY(bias_grad(tb.tdhq, Hi, Hi, n, G.post1_b, s));
    if (G.post1_w) {
      BD_TRY(linear_wgrad(tb.tdhq, Hi, Hi, f.beliefs, Be, Be, n, G.post1_w, Be + E, 0, s));
      BD_TRY(linear_wgrad(tb.tdhq, Hi, Hi, f.embeddings, E, E, n, G.post1_w, Be + E, Be, s));
    }
    if (a.d_embeddings)
      BD_TRY(linear_dgrad(tb.tdhq, Hi, Hi, r.post1.w, Be + E, Be, E, n, a.d_embeddings, E,
                          BD_ACT_IDENTITY, nullptr, 0, 0, s));
  }
  if (G.prior2_b) BD_TRY(bias_grad(tb.tdpre, 2 * Sz, 2 * Sz, n, G.prior2_b, s));
  if (G.prior2_w) BD_TRY(linear_wgrad(tb.tdpre, 2 * Sz, 2 * Sz, tb.th, Hi, Hi, n, G.prior2_w, Hi, 0, s));
  if (G.prior1_b) BD_TRY(bias_grad(tb.tdh, Hi, Hi, n, G.prior1_b, s));
  if (G.prior1_w) BD_TRY(linear_wgrad(tb.tdh, Hi, Hi, f.beliefs, Be, Be, n, G.prior1_w, Be, 0, s));
  if (G.b_ih) BD_TRY(bias_grad(tb.tdgi, 3 * Be, 3 * Be, n, G.b_ih, s));
  if (G.b_hh) BD_TRY(bias_grad(tb.tdgh, 3 * Be, 3 * Be, n, G.b_hh, s));
  if (G.w_ih) BD_TRY(linear_wgrad(tb.tdgi, 3 * Be, 3 * Be, tb.tx, Be, Be, n, G.w_ih, Be, 0, s));
  if (G.w_hh) {      // h_{t-1}: init_belief for t = 0, beliefs[t-1] after
    BD_TRY(linear_wgrad(tb.tdgh, 3 * Be, 3 * Be, f.init_belief, Be, Be, nB, G.w_hh, Be, 0, s));
    if (f.L > 1)
      BD_TRY(linear_wgrad(tb.tdgh + (size_t)B * 3 * Be, 3 * Be, 3 * Be, f.beliefs, Be, Be, n - nB, G.w_hh, Be, 0, s));
  }
  if (G.embed_b) BD_TRY(bias_grad(tb.tdx, Be, Be, n, G.embed_b, s));
  if (G.embed_w) {
    BD_TRY(linear_wgrad(tb.tdx, Be, Be, tb.tsx, Sz, Sz, n, G.embed_w, Sz + A, 0, s));
    BD_TRY(linear_wgrad(tb.tdx, Be, Be, f.actions, A, A, n, G.embed_w, Sz + A, Sz, s));
  }
  return BD_OK;
}


// -------------------------------------------------------------------------------------
// Observe pass on the persistent cluster kernels (observe_persist.cuh)
// -------------------------------------------------------------------------------------
constexpr int64_t kPersistMaxRows = 8 * obs::kR;      // up to 8 clusters of 16 CTAs
static bool persist_shape_ok(const bd_rssm& r, int L, int64_t B) {
  const int Be = r.belief_size, Hi = r.hidden_size, S = r.state_size, A = r.action_size;
  // (L = 1, the acting path's posterior step: the cluster kernels' per-call set-up -- per-CTA weight packing, a
  // 16-CTA cluster launch -- costs more than the handful of per-step launches: 133 vs 119 us per acting step)
  return L >= 2 && B >= 1 && B <= kPersistMaxRows && Be <= 256 && Hi <= 256 && S <= 128 && S + A <= obs::kMaxSmallK &&
         2 * S <= obs::kMaxSmallK;
}
static int obs_prof_flag() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("BD_OBS_PROF"); v = e ? atoi(e) : 0; }
  return v;
}
struct PersistOps { obs::OpDesc d[obs::OP_COUNT]; };
static PersistOps persist_ops(const bd_rssm& r, bool tc = false) {
  const int Be = r.belief_size, Hi = r.hidden_size, S = r.state_size, A = r.action_size;
  const int Bep = (Be + 3) & ~3;
  PersistOps o;
  o.d[obs::OP_EMB] = obs::make_op(S + A, (Be + 3) / 4);
  o.d[obs::OP_GRU] = obs::make_op(2 * Be, Be);
  o.d[obs::OP_Q1F] = obs::make_op(Be, (Hi + 3) / 4);
  o.d[obs::OP_Q2F] = obs::make_op(Hi, S);
  o.d[obs::OP_B1] = obs::make_op(2 * S, (Hi + 3) / 4);
  o.d[obs::OP_B2] = obs::make_op(Hi, (Be + 3) / 4);
  o.d[obs::OP_B3] = obs::make_op(4 * Be, 2 * Bep / 4);
  o.d[obs::OP_B4] = obs::make_op(Be, (S + A + 3) / 4);
  if (tc && (Be & 3) == 0) {      // TF32 mma.sync for the two big contractions (K % 8 == 0, 32 or 64 slots)
    for (int i : {(int)obs::OP_GRU, (int)obs::OP_B3})
      if (o.d[i].WP == 8 || o.d[i].WP == 16) { o.d[i].tc = 1; o.d[i].KG = 4; }
  }
  return o;
}
static size_t persist_ws_bytes(const bd_rssm& r, int L, int64_t B, bool backward) {
  const size_t Be = r.belief_size, Hi = r.hidden_size, S = r.state_size;
  const size_t n = (size_t)L * B, nch = (size_t)((B + obs::kR - 1) / obs::kR);
  const PersistOps o = persist_ops(r);
  size_t bytes = 0;
  auto add = [&](size_t floats) { bytes += pad256(floats); };
  if (!backward) {
    add(n * Hi); add(n * Hi); add(n * 2 * S);
    for (int i = obs::OP_EMB; i <= obs::OP_Q2F; ++i) add(obs::op_floats(o.d[i]));
    add(nch * obs::fwd_scratch_floats((int)Be, (int)Hi));
  } else {
    add(n * Be); add(n * Hi); add(n * 2 * S); add(n * Hi); add(n * 3 * Be); add(n * 3 * Be); add(n * Be);
    add(n * S); add(n * Hi); add(n * 2 * S); add(n * Hi);                       // TimeBatch
    add(n * 3 * Be); add(n * 3 * Be); add(n * 2 * S); add(n * 2 * S); add(n * Be);   // gi gh pre preq Gtot
    for (int i = obs::OP_B1; i <= obs::OP_B4; ++i) add(obs::op_floats(o.d[i]));
    add(nch * obs::bwd_scratch_floats((int)Be, (int)Hi, (int)S));
  }
  return bytes + 4096;
}
// Can a cluster of 16 CTAs with this much shared memory be scheduled?  (asked once per kernel)
template <typename Kern>
static bool persist_launchable(Kern kern) {
  // asked once per (device, kernel): function attributes and cluster occupancy are per device
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, int> cache;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  int& cached = cache.emplace(std::make_pair(dev, reinterpret_cast<const void*>(kern)), -1).first->second;
  if (cached >= 0) return cached != 0;
  const char* env = getenv("BD_OBS_PERSIST");
  if (env && env[0] == '0') { cached = 0; return false; }
  cached = 0;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, obs::kSmemBytes) != cudaSuccess ||
      cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(obs::kC); cfg.blockDim = dim3(obs::kThreads); cfg.dynamicSmemBytes = obs::kSmemBytes;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = obs::kC; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int nclusters = 0;
  if (cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return false; }
  cached = nclusters >= 1 ? 1 : 0;
  return cached != 0;
}
template <typename Kern, typename Args>
static int persist_launch(Kern kern, const Args& args, int64_t B, cudaStream_t s) {
  cudaLaunchConfig_t cfg = {};
  const unsigned nch = (unsigned)((B + obs::kR - 1) / obs::kR);
  cfg.gridDim = dim3(obs::kC * nch); cfg.blockDim = dim3(obs::kThreads);
  cfg.dynamicSmemBytes = obs::kSmemBytes; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = obs::kC; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, args);
  ++bd::g_launch_count;
  if (e != cudaSuccess) BD_FAIL(BD_ERR_CUDA, "observe persistent kernel launch failed: %s", cudaGetErrorString(e));
  return BD_OK;
}
static int persist_pack(const bd_rssm& r, PersistOps& o, int first, int count, Arena& ar, cudaStream_t s) {
  obs::PackArgs p{};
  int kmax = 1;
  for (int i = 0; i < obs::OP_COUNT; ++i) { p.K[i] = o.d[i].K; p.NJ[i] = o.d[i].NJ; p.Wc[i] = o.d[i].Wc; }
  for (int i = first; i < first + count; ++i) {
    p.dst[i] = ar.f32(obs::op_floats(o.d[i]));
    o.d[i].w = p.dst[i];
    kmax = max(kmax, o.d[i].K);
  }
  if (!ar.ok()) BD_FAIL(BD_ERR_WORKSPACE, "observe (persistent): workspace too small for packed weights");
  p.first = first; p.count = count;
  p.Be = r.belief_size; p.Bep = (r.belief_size + 3) & ~3; p.Hi = r.hidden_size; p.S = r.state_size;
  p.A = r.action_size; p.E = r.embedding_size;
  p.w_sa = r.embed.w; p.w_ih = r.w_ih; p.w_hh = r.w_hh; p.w_q1 = r.post1.w; p.w_q2 = r.post2.w;
  dim3 grid((unsigned)(((size_t)obs::kC * kmax * obs::kRowFloats + 255) / 256), count);
  obs::pack_ops_kernel<<<grid, 256, 0, s>>>(p);
  BD_CUDA_LAUNCH_CHECK();
  return BD_OK;
}

static int observe_forward_persist(const bd_transition_args* a, void* ws, size_t ws_bytes, bool tc, cudaStream_t s) {
  const bd_rssm& r = a->rssm;
  const long long Be = r.belief_size, Sz = r.state_size, E = r.embedding_size, Hi = r.hidden_size, B = a->B;
  const size_t n = (size_t)a->L * B, nch = (size_t)((B + obs::kR - 1) / obs::kR);
  Tf32Scope tf32_gemms(tc);      // 16-bit modes: the batched GEMMs below on TF32 mma.sync
  Arena ar(ws, ws_bytes);
  float* PE = ar.f32(n * Hi);
  float* h = ar.f32(n * Hi);
  float* pre = ar.f32(n * 2 * Sz);
  PersistOps o = persist_ops(r, tc);
  BD_TRY(persist_pack(r, o, obs::OP_EMB, 4, ar, s));
  float* scratch = ar.f32(nch * obs::fwd_scratch_floats((int)Be, (int)Hi));
  if (!ar.ok()) BD_FAIL(BD_ERR_WORKSPACE, "observe forward (persistent): workspace too small");
  // embedding half of the posterior's first layer for every step: PE = emb W_q1[:, Be:]^T + b_q1
  {
    GemmArgs g;
    g.A1 = a->embeddings; g.lda1 = E; g.K1 = (int)E; g.B = r.post1.w + Be; g.ldb = Be + E;
    g.C = PE; g.ldc = Hi; g.bias = r.post1.b; g.M = (int)n; g.N = (int)Hi; g.act = BD_ACT_IDENTITY;
    BD_TRY((launch_gemm<false, true, EPI_BIAS_ACT>(g, s)));
  }
  obs::FwdArgs k{};
  k.emb = o.d[obs::OP_EMB]; k.gru = o.d[obs::OP_GRU]; k.q1 = o.d[obs::OP_Q1F]; k.q2 = o.d[obs::OP_Q2F];
  k.L = a->L; k.B = B; k.Be = (int)Be; k.Hi = (int)Hi; k.S = (int)Sz; k.A = r.action_size; k.act = r.activation;
  k.min_std = r.min_std_dev;
  k.init_state = a->init_state; k.init_belief = a->init_belief; k.actions = a->actions;
  k.nonterm = a->nonterminals; k.eps_post = a->eps_post; k.PE = PE;
  k.b_sa = r.embed.b; k.b_ih = r.b_ih; k.b_hh = r.b_hh; k.b_q2 = r.post2.b;
  k.beliefs = a->beliefs; k.post_s = a->post_states; k.post_m = a->post_means; k.post_sd = a->post_stds;
  k.scratch = scratch;
  k.prof = obs_prof_flag();
  BD_TRY(persist_launch(obs::observe_fwd_kernel, k, B, s));
  // prior branch: not fed back in observe mode -> batched over all L*B beliefs (src/models.py:256)
  BD_TRY(linear_fwd(r.prior1, r.activation, a->beliefs, (int)Be, Be, nullptr, 0, 0, nullptr, (int)n, h, Hi, s));
  BD_TRY(linear_fwd(r.prior2, BD_ACT_IDENTITY, h, (int)Hi, Hi, nullptr, 0, 0, nullptr, (int)n, pre, 2 * Sz, s));
  belief_sample_fwd_kernel<<<grid1d((long long)n * Sz), 256, 0, s>>>(
      pre, a->eps_prior, r.min_std_dev, a->prior_states, a->prior_means, a->prior_stds, (long long)n * Sz, (int)Sz);
  BD_CUDA_LAUNCH_CHECK();
  return BD_OK;
}

static int observe_backward_persist(const bd_transition_bwd_args* a, void* ws, size_t ws_bytes, bool tc, cudaStream_t s) {
  const bd_transition_args& f = a->fwd;
  const bd_rssm& r = f.rssm;
  const long long Be = r.belief_size, Sz = r.state_size, A = r.action_size, E = r.embedding_size,
                  Hi = r.hidden_size, B = f.B;
  const int L = f.L;
  const size_t n = (size_t)L * B, nch = (size_t)((B + obs::kR - 1) / obs::kR);
  const int ni = (int)n, nB = (int)B;
  Tf32Scope tf32_gemms(tc);
  Arena ar(ws, ws_bytes);
  TimeBatch tb;
  tb.tx = ar.f32(n * Be); tb.th = ar.f32(n * Hi); tb.tdpre = ar.f32(n * 2 * Sz); tb.tdh = ar.f32(n * Hi);
  tb.tdgi = ar.f32(n * 3 * Be); tb.tdgh = ar.f32(n * 3 * Be); tb.tdx = ar.f32(n * Be); tb.tsx = ar.f32(n * Sz);
  tb.thq = ar.f32(n * Hi); tb.tdpreq = ar.f32(n * 2 * Sz); tb.tdhq = ar.f32(n * Hi);
  float* gi = ar.f32(n * 3 * Be);
  float* gh = ar.f32(n * 3 * Be);
  float* pre = ar.f32(n * 2 * Sz);
  float* preq = ar.f32(n * 2 * Sz);
  float* Gtot = ar.f32(n * Be);
  PersistOps o = persist_ops(r, tc);
  BD_TRY(persist_pack(r, o, obs::OP_B1, 4, ar, s));
  float* scratch = ar.f32(nch * obs::bwd_scratch_floats((int)Be, (int)Hi, (int)Sz));
  if (!ar.ok()) BD_FAIL(BD_ERR_WORKSPACE, "observe backward (persistent): workspace too small");
  // ---- recompute every step's internals from the saved outputs, batched over time
  slice_cols_kernel<<<grid1d(B * Sz), 256, 0, s>>>(f.init_state, Sz, 0, (int)Sz, f.nonterminals, tb.tsx, B);
  BD_CUDA_LAUNCH_CHECK();
  if (L > 1) {
    slice_cols_kernel<<<grid1d((long long)(n - B) * Sz), 256, 0, s>>>(
        f.post_states, Sz, 0, (int)Sz, f.nonterminals ? f.nonterminals + B : nullptr, tb.tsx + B * Sz,
        (long long)(n - B));
    BD_CUDA_LAUNCH_CHECK();
  }
  BD_TRY(linear_fwd(r.embed, r.activation, tb.tsx, (int)Sz, Sz, f.actions, (int)A, A, nullptr, ni, tb.tx, Be, s));
  BD_TRY(matmul_nt_bias(tb.tx, (int)Be, Be, r.w_ih, r.b_ih, (int)(3 * Be), ni, gi, 3 * Be, s));
  BD_TRY(matmul_nt_bias(f.init_belief, (int)Be, Be, r.w_hh, r.b_hh, (int)(3 * Be), nB, gh, 3 * Be, s));
  if (L > 1)
    BD_TRY(matmul_nt_bias(f.beliefs, (int)Be, Be, r.w_hh, r.b_hh, (int)(3 * Be), ni - nB, gh + B * 3 * Be, 3 * Be, s));
  BD_TRY(linear_fwd(r.prior1, r.activation, f.beliefs, (int)Be, Be, nullptr, 0, 0, nullptr, ni, tb.th, Hi, s));
  BD_TRY(linear_fwd(r.prior2, BD_ACT_IDENTITY, tb.th, (int)Hi, Hi, nullptr, 0, 0, nullptr, ni, pre, 2 * Sz, s));
  BD_TRY(linear_fwd(r.post1, r.activation, f.beliefs, (int)Be, Be, f.embeddings, (int)E, E, nullptr, ni, tb.thq, Hi, s));
  BD_TRY(linear_fwd(r.post2, BD_ACT_IDENTITY, tb.thq, (int)Hi, Hi, nullptr, 0, 0, nullptr, ni, preq, 2 * Sz, s));
  // ---- prior branch backward (no recurrence in observe mode): d b_t += d_h W_p1
  belief_sample_bwd_kernel<<<grid1d((long long)n * Sz), 256, 0, s>>>(
      pre, f.eps_prior, a->g_prior_states, nullptr, a->g_prior_means, a->g_prior_stds, tb.tdpre, (long long)n * Sz, (int)Sz);
  BD_CUDA_LAUNCH_CHECK();
  BD_TRY(linear_dgrad(tb.tdpre, (int)(2 * Sz), 2 * Sz, r.prior2.w, (int)Hi, 0, (int)Hi, ni, tb.tdh, Hi, r.activation,
                      tb.th, Hi, 0, s));
  if (a->g_beliefs) cudaMemcpyAsync(Gtot, a->g_beliefs, n * Be * sizeof(float), cudaMemcpyDeviceToDevice, s);
  else cudaMemsetAsync(Gtot, 0, n * Be * sizeof(float), s);
  BD_TRY(linear_dgrad(tb.tdh, (int)Hi, Hi, r.prior1.w, (int)Be, 0, (int)Be, ni, Gtot, Be, BD_ACT_IDENTITY, nullptr, 0, 1, s));
  // ---- the recurrence
  obs::BwdArgs k{};
  k.b1 = o.d[obs::OP_B1]; k.b2 = o.d[obs::OP_B2]; k.b3 = o.d[obs::OP_B3]; k.b4 = o.d[obs::OP_B4];
  k.L = L; k.B = B; k.Be = (int)Be; k.Bep = ((int)Be + 3) & ~3; k.Hi = (int)Hi; k.S = (int)Sz; k.A = (int)A;
  k.act = r.activation;
  k.nonterm = f.nonterminals; k.eps_post = f.eps_post;
  k.g_post_s = a->g_post_states; k.g_post_m = a->g_post_means; k.g_post_sd = a->g_post_stds;
  k.preq = preq; k.hq = tb.thq; k.x = tb.tx; k.gi = gi; k.gh = gh; k.init_belief = f.init_belief;
  k.beliefs = f.beliefs; k.Gtot = Gtot;
  k.tdpreq = tb.tdpreq; k.tdhq = tb.tdhq; k.tdgi = tb.tdgi; k.tdgh = tb.tdgh; k.tdx = tb.tdx;
  k.d_actions = a->d_actions; k.d_init_state = a->d_init_state; k.d_init_belief = a->d_init_belief;
  k.scratch = scratch;
  k.vec_h = ((reinterpret_cast<uintptr_t>(f.init_belief) | reinterpret_cast<uintptr_t>(f.beliefs)) & 15) == 0;
  k.prof = obs_prof_flag();
  BD_TRY(persist_launch(obs::observe_bwd_kernel, k, B, s));
  // ---- parameter gradients and d embeddings, batched over time
  return time_batched_param_grads(r, f, *a, tb, s);
}
int actor_act(const float* raw, const float* eps, const bd_actor_cfg* cfg, int64_t rows, int A,
              int deterministic, float* action, bd_stream_t stream) {
  BD_CHECK_ARG(raw && eps && cfg && action, "actor_act: null pointer");
  BD_CHECK_ARG(rows >= 0 && A >= 1, "actor_act: bad rows / action size");
  BD_CHECK_ARG(!deterministic || cfg->entropy_samples >= 1, "actor_act: SampleDist.mode needs >= 1 sample");
  if (rows == 0) return BD_OK;
  const int wpb = 4;
  actor_act_kernel<<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, S(stream)>>>(raw, eps, *cfg, rows, A,
                                                                                 deterministic, action);
  BD_CUDA_LAUNCH_CHECK();
  return BD_OK;
}
int value_loss(const float* value, const float* target, const float* weight, int64_t n, float* loss,
               float* d_value, void* ws, size_t ws_bytes, bd_stream_t stream) {
  BD_CHECK_ARG(value && target && loss, "value_loss: null pointer");
  BD_CHECK_ARG(n >= 1, "value_loss: n must be >= 1");
  BD_CHECK_ARG(ws && ws_bytes >= kValueLossBlocks * sizeof(float), "value_loss: workspace too small");
  long long g = (n + 255) / 256;
  if (g > kValueLossBlocks) g = kValueLossBlocks;
  float* partial = static_cast<float*>(ws);
  value_loss_kernel<<<(unsigned)g, 256, 0, S(stream)>>>(value, target, weight, n, d_value, partial);
  BD_CUDA_LAUNCH_CHECK();
  value_loss_finish_kernel<<<1, 1024, 0, S(stream)>>>(partial, (int)g, n, loss);
  BD_CUDA_LAUNCH_CHECK();
  return BD_OK;
}

size_t transition_workspace_bytes(const bd_rssm* r, int L, int64_t B, int observe, int backward) {
  int64_t rows = B < kMaxChunkRows ? B : kMaxChunkRows;
  size_t step = (size_t)(rows > 0 ? rows : 1) * step_row_floats(*r, observe != 0, backward != 0) * sizeof(float);
  size_t total = step + (backward ? time_batch_bytes(*r, L, B, observe != 0) : 0) + kSlackBytes;
  if (observe && persist_shape_ok(*r, L, B)) {
    const size_t p = persist_ws_bytes(*r, L, B, backward != 0) + kSlackBytes;
    if (p > total) total = p;
  }
  return total;
}

static int check_transition(const bd_transition_args& a) {
  const bool observe = a.embeddings != nullptr;
  BD_TRY(check_rssm(a.rssm, observe));
  BD_CHECK_ARG(a.L >= 1 && a.B >= 0, "transition: bad L/B");
  BD_CHECK_ARG(a.init_state && a.init_belief && a.actions && a.eps_prior, "transition: null input");
  BD_CHECK_ARG(a.beliefs && a.prior_states && a.prior_means && a.prior_stds, "transition: null output");
  if (observe) BD_CHECK_ARG(a.eps_post && a.post_states && a.post_means && a.post_stds,
                            "transition: observe mode needs eps_post and posterior outputs");
  return BD_OK;
}

int transition_forward(const bd_transition_args* a, void* ws, size_t ws_bytes, bd_stream_t stream, int precision) {
  BD_TRY(check_transition(*a));
  if (a->B == 0) return BD_OK;
  const bd_rssm& r = a->rssm;
  const bool observe = a->embeddings != nullptr;
  const long long Be = r.belief_size, Sz = r.state_size, A = r.action_size, E = r.embedding_size, B = a->B;
  if (observe && persist_shape_ok(r, a->L, B) && ws_bytes >= persist_ws_bytes(r, a->L, B, false) &&
      persist_launchable(obs::observe_fwd_kernel))
    return observe_forward_persist(a, ws, ws_bytes, precision != BD_PREC_FP32, S(stream));
  int chunk;
  BD_TRY(chunk_rows_for(ws_bytes, step_row_floats(r, observe, false), B, &chunk));
  for (long long r0 = 0; r0 < B; r0 += chunk) {
    int nr = (int)((B - r0) < chunk ? (B - r0) : chunk);
    Arena ar(ws, ws_bytes);
    StepBuf w;
    if (!carve_step(ar, r, chunk, observe, false, w)) BD_FAIL(BD_ERR_WORKSPACE, "transition_forward: workspace");
    for (int t = 0; t < a->L; ++t) {
      const long long o = (long long)t * B + r0, op = (long long)(t - 1) * B + r0;
      const float* b_prev = t == 0 ? a->init_belief + r0 * Be : a->beliefs + op * Be;
      const float* s_prev = t == 0 ? a->init_state + r0 * Sz
                                   : (observe ? a->post_states : a->prior_states) + op * Sz;
      BD_TRY(step_forward(r, w, nr, s_prev, a->nonterminals ? a->nonterminals + o : nullptr,
                          a->actions + o * A, b_prev, a->eps_prior + o * Sz,
                          observe ? a->embeddings + o * E : nullptr,
                          observe ? a->eps_post + o * Sz : nullptr, a->beliefs + o * Be,
                          a->prior_states + o * Sz, a->prior_means + o * Sz, a->prior_stds + o * Sz,
                          observe ? a->post_states + o * Sz : nullptr,
                          observe ? a->post_means + o * Sz : nullptr,
                          observe ? a->post_stds + o * Sz : nullptr, S(stream)));
    }
  }
  return BD_OK;
}

int transition_backward(const bd_transition_bwd_args* a, void* ws, size_t ws_bytes,
                        bd_stream_t stream, int precision) {
  const bd_transition_args& f = a->fwd;
  BD_TRY(check_transition(f));
  if (f.B == 0) return BD_OK;
  const bd_rssm& r = f.rssm;
  const bool observe = f.embeddings != nullptr;
  const long long Be = r.belief_size, Sz = r.state_size, A = r.action_size, E = r.embedding_size, B = f.B;
  cudaStream_t s = S(stream);
  const long long Hi = r.hidden_size;
  if (observe && persist_shape_ok(r, f.L, B) && ws_bytes >= persist_ws_bytes(r, f.L, B, true) &&
      persist_launchable(obs::observe_bwd_kernel))
    return observe_backward_persist(a, ws, ws_bytes, precision != BD_PREC_FP32, s);
  const bd_rssm_grads& G = a->grads;
  const bool any_wgrad = G.embed_w || G.embed_b || G.w_ih || G.w_hh || G.b_ih || G.b_hh || G.prior1_w ||
                         G.prior1_b || G.prior2_w || G.prior2_b || G.post1_w || G.post1_b || G.post2_w ||
                         G.post2_b;
  const size_t tb_bytes = time_batch_bytes(r, f.L, B, observe);
  const size_t step_bytes = (size_t)B * step_row_floats(r, observe, true) * sizeof(float);
  // time-batched weight gradients: the whole batch is one chunk and the workspace holds the operands
  const bool batched = any_wgrad && tb_bytes > 0 && ws_bytes >= step_bytes + tb_bytes + kSlackBytes;
  int chunk;
  BD_TRY(chunk_rows_for(batched ? ws_bytes - tb_bytes : ws_bytes, step_row_floats(r, observe, true), B, &chunk));
  for (long long r0 = 0; r0 < B; r0 += chunk) {
    int nr = (int)((B - r0) < chunk ? (B - r0) : chunk);
    Arena ar(ws, ws_bytes);
    StepBuf w;
    if (!carve_step(ar, r, chunk, observe, true, w)) BD_FAIL(BD_ERR_WORKSPACE, "transition_backward: workspace");
    // per-step slices (t * B rows) of the wgrad operands
    float *tx = nullptr, *th = nullptr, *tdpre = nullptr, *tdh = nullptr, *tdgi = nullptr, *tdgh = nullptr,
          *tdx = nullptr, *tsx = nullptr, *thq = nullptr, *tdpreq = nullptr, *tdhq = nullptr;
    const size_t LB = (size_t)f.L * B;
    if (batched) {
      tx = ar.f32(LB * Be); th = ar.f32(LB * Hi); tdpre = ar.f32(LB * 2 * Sz); tdh = ar.f32(LB * Hi);
      tdgi = ar.f32(LB * 3 * Be); tdgh = ar.f32(LB * 3 * Be); tdx = ar.f32(LB * Be); tsx = ar.f32(LB * Sz);
      if (observe) { thq = ar.f32(LB * Hi); tdpreq = ar.f32(LB * 2 * Sz); tdhq = ar.f32(LB * Hi); }
      if (!ar.ok() || chunk != B) BD_FAIL(BD_ERR_WORKSPACE, "transition_backward: time-batch workspace");
    }
    bool have_carry = false;
    for (int t = f.L - 1; t >= 0; --t) {
      const long long o = (long long)t * B + r0, op = (long long)(t - 1) * B + r0;
      const float* b_prev = t == 0 ? f.init_belief + r0 * Be : f.beliefs + op * Be;
      const float* s_prev = t == 0 ? f.init_state + r0 * Sz
                                   : (observe ? f.post_states : f.prior_states) + op * Sz;
      // Gb = g_beliefs[t] + carry_b
      add2_kernel<<<grid1d((long long)nr * Be), 256, 0, s>>>(
          a->g_beliefs ? a->g_beliefs + o * Be : nullptr, have_carry ? w.carry_b : nullptr, w.Gb,
          (long long)nr * Be);
      BD_CUDA_LAUNCH_CHECK();
      const float* carry_s = have_carry ? w.carry_s : nullptr;
      StepBuf wt = w;
      if (batched) {     // this step's wgrad operands land in their time-major slices
        const size_t tb = (size_t)t * B;
        wt.x = tx + tb * Be; wt.h = th + tb * Hi; wt.d_pre = tdpre + tb * 2 * Sz; wt.d_h = tdh + tb * Hi;
        wt.d_gi = tdgi + tb * 3 * Be; wt.d_gh = tdgh + tb * 3 * Be; wt.dx = tdx + tb * Be;
        if (observe) { wt.hq = thq + tb * Hi; wt.d_preq = tdpreq + tb * 2 * Sz; wt.d_hq = tdhq + tb * Hi; }
        if (G.embed_w) {   // masked previous state = the embed layer's input (src/models.py:241-251)
          slice_cols_kernel<<<grid1d((long long)nr * Sz), 256, 0, s>>>(
              s_prev, Sz, 0, (int)Sz, f.nonterminals ? f.nonterminals + o : nullptr, tsx + tb * Sz, nr);
          BD_CUDA_LAUNCH_CHECK();
        }
      }
      // the carried state gradient belongs to whichever sample fed the next step; it must be
      // consumed before step_backward overwrites w.carry_s, so stage it in w.dsa[:, :S]... the
      // sample-bwd kernels run before anything writes carry_s/dsa, so passing it directly is safe.
      BD_TRY(step_backward(
          r, wt, nr, s_prev, f.nonterminals ? f.nonterminals + o : nullptr, f.actions + o * A, b_prev,
          f.beliefs + o * Be, f.eps_prior + o * Sz, observe ? f.embeddings + o * E : nullptr,
          observe ? f.eps_post + o * Sz : nullptr,
          a->g_prior_states ? a->g_prior_states + o * Sz : nullptr, observe ? nullptr : carry_s,
          a->g_prior_means ? a->g_prior_means + o * Sz : nullptr,
          a->g_prior_stds ? a->g_prior_stds + o * Sz : nullptr,
          (observe && a->g_post_states) ? a->g_post_states + o * Sz : nullptr,
          observe ? carry_s : nullptr,
          (observe && a->g_post_means) ? a->g_post_means + o * Sz : nullptr,
          (observe && a->g_post_stds) ? a->g_post_stds + o * Sz : nullptr,
          (observe && a->d_embeddings && !batched) ? a->d_embeddings + o * E : nullptr,
          batched ? nullptr : &a->grads, s));
      if (a->d_actions) {
        slice_cols_kernel<<<grid1d((long long)nr * A), 256, 0, s>>>(w.dsa, Sz + A, (int)Sz, (int)A, nullptr,
                                                                    a->d_actions + o * A, nr);
        BD_CUDA_LAUNCH_CHECK();
      }
      have_carry = true;
    }
    if (batched) {     // one GEMM / column sum per parameter over all L*B rows
      TimeBatch tb{tx, th, tdpre, tdh, tdgi, tdgh, tdx, tsx, thq, tdpreq, tdhq};
      BD_TRY(time_batched_param_grads(r, f, *a, tb, s));
    }
    if (a->d_init_belief)
      cudaMemcpyAsync(a->d_init_belief + r0 * Be, w.carry_b, (size_t)nr * Be * sizeof(float),
                      cudaMemcpyDeviceToDevice, s);
    if (a->d_init_state)
      cudaMemcpyAsync(a->d_init_state + r0 * Sz, w.carry_s, (size_t)nr * Sz * sizeof(float),
                      cudaMemcpyDeviceToDevice, s);
  }
  return BD_OK;
}

// -------------------------------------------------------------------------------------
// Dreamer.imagine_ahead forward / backward (src/dreamer.py:178-237)
// -------------------------------------------------------------------------------------
static size_t imagine_row_floats(const bd_rssm& r, const bd_mlp& actor, bool backward) {
  size_t f = step_row_floats(r, false, backward) + mlp_row_floats(actor, backward);
  if (backward) f += 2 * r.action_size;
  return f;
}
size_t imagine_workspace_bytes(const bd_rssm* r, const bd_mlp* actor, int T, int64_t N, int backward) {
  (void)T;
  int64_t rows = N < kMaxChunkRows ? N : kMaxChunkRows;
  return (size_t)(rows > 0 ? rows : 1) * imagine_row_floats(*r, *actor, backward != 0) * sizeof(float) + kSlackBytes;
}

int check_imagine(const bd_imagine_args& a) {
  BD_TRY(check_rssm(a.rssm, false));
  BD_TRY(check_mlp(a.actor, a.rssm.belief_size + a.rssm.state_size));
  BD_CHECK_ARG(a.actor.layer[a.actor.n_layers - 1].out_features == 2 * a.rssm.action_size,
               "imagine: actor output must be 2*action_size");
  BD_CHECK_ARG(a.T >= 1 && a.N >= 0, "imagine: bad T/N");
  BD_CHECK_ARG(a.actor_cfg.entropy_samples >= 1, "imagine: entropy_samples must be >= 1");
  BD_CHECK_ARG(a.prev_state && a.prev_belief && a.eps_a && a.eps_e && a.eps_s, "imagine: null input");
  BD_CHECK_ARG(a.beliefs && a.states && a.means && a.stds && a.entropy && a.actions &&
               a.actor_raw && a.dent, "imagine: null output");
  return BD_OK;
}

int imagine_forward(const bd_imagine_args* a, void* ws, size_t ws_bytes, bd_stream_t stream) {
  BD_TRY(check_imagine(*a));
  if (a->N == 0) return BD_OK;
  const bd_rssm& r = a->rssm;
  const long long Be = r.belief_size, Sz = r.state_size, A = r.action_size, N = a->N;
  const int J = a->actor_cfg.entropy_samples;
  cudaStream_t s = S(stream);
  int chunk;
  BD_TRY(chunk_rows_for(ws_bytes, imagine_row_floats(r, a->actor, false), N, &chunk));
  const int aw = mlp_max_width(a->actor);
  for (long long r0 = 0; r0 < N; r0 += chunk) {
    int nr = (int)((N - r0) < chunk ? (N - r0) : chunk);
    Arena ar(ws, ws_bytes);
    StepBuf w;
    bool ok = carve_step(ar, r, chunk, false, false, w);
    float* ha = ar.f32((size_t)chunk * aw);
    float* hb = ar.f32((size_t)chunk * aw);
    if (!ok || !ar.ok()) BD_FAIL(BD_ERR_WORKSPACE, "imagine_forward: workspace");
    float* hid[BD_MAX_LAYERS];
    for (int l = 0; l < BD_MAX_LAYERS; ++l) hid[l] = (l & 1) ? hb : ha;
    for (int t = 0; t < a->T; ++t) {
      const long long o = (long long)t * N + r0, op = (long long)(t - 1) * N + r0;
      const float* b_prev = t == 0 ? a->prev_belief + r0 * Be : a->beliefs + op * Be;
      const float* s_prev = t == 0 ? a->prev_state + r0 * Sz : a->states + op * Sz;
      // actor on (b_t, s_t)                                  src/dreamer.py:215, src/models.py:511
      BD_TRY(mlp_forward_rows(a->actor, b_prev, (int)Be, Be, s_prev, (int)Sz, Sz, nr, hid,
                              a->actor_raw + o * 2 * A, 2 * A, s));
      actor_head_fwd_kernel<<<grid1d(nr, 128), 128, 0, s>>>(
          a->actor_raw + o * 2 * A, a->eps_a + o * A, a->eps_e + (long long)t * J * N * A,
          a->actor_cfg, r0, N, nr, (int)A, a->actions + o * A, a->entropy + o, a->dent + o * 2 * A);
      BD_CUDA_LAUNCH_CHECK();
      BD_TRY(step_forward(r, w, nr, s_prev, nullptr, a->actions + o * A, b_prev, a->eps_s + o * Sz,
                          nullptr, nullptr, a->beliefs + o * Be, a->states + o * Sz,
                          a->means + o * Sz, a->stds + o * Sz, nullptr, nullptr, nullptr, s));
    }
  }
  return BD_OK;
}

int imagine_backward(const bd_imagine_bwd_args* a, void* ws, size_t ws_bytes, bd_stream_t stream) {
  return imagine_backward_ex(a, ws, ws_bytes, stream, nullptr);
}

// d_raw_all != nullptr: write the gradient wrt the raw actor outputs of every step to
// d_raw_all (T,N,2A) and leave the actor's own backward to the caller (tensor-core path).
int imagine_backward_ex(const bd_imagine_bwd_args* a, void* ws, size_t ws_bytes, bd_stream_t stream,
                        float* d_raw_all) {
  const bd_imagine_args& f = a->fwd;
  BD_TRY(check_imagine(f));
  if (f.N == 0) return BD_OK;
  const bd_rssm& r = f.rssm;
  const bd_mlp& actor = f.actor;
  const long long Be = r.belief_size, Sz = r.state_size, A = r.action_size, N = f.N;
  cudaStream_t s = S(stream);
  int chunk;
  BD_TRY(chunk_rows_for(ws_bytes, imagine_row_floats(r, actor, true), N, &chunk));
  const int aw = mlp_max_width(actor);
  bool want_actor = false;
  for (int l = 0; l < actor.n_layers; ++l) want_actor |= (a->actor_dw[l] || a->actor_db[l]);
  if (d_raw_all) want_actor = true;
  for (long long r0 = 0; r0 < N; r0 += chunk) {
    int nr = (int)((N - r0) < chunk ? (N - r0) : chunk);
    Arena ar(ws, ws_bytes);
    StepBuf w;
    bool ok = carve_step(ar, r, chunk, false, true, w);
    float* hid[BD_MAX_LAYERS] = {nullptr};
    for (int l = 0; l + 1 < actor.n_layers; ++l) hid[l] = ar.f32((size_t)chunk * actor.layer[l].out_features);
    float* d0 = ar.f32((size_t)chunk * aw);
    float* d1 = ar.f32((size_t)chunk * aw);
    float* d_raw = ar.f32((size_t)chunk * 2 * A);
    if (!ok || !ar.ok()) BD_FAIL(BD_ERR_WORKSPACE, "imagine_backward: workspace");
    bool have_carry = false;
    for (int t = f.T - 1; t >= 0; --t) {
      const long long o = (long long)t * N + r0, op = (long long)(t - 1) * N + r0;
      const float* b_prev = t == 0 ? f.prev_belief + r0 * Be : f.beliefs + op * Be;
      const float* s_prev = t == 0 ? f.prev_state + r0 * Sz : f.states + op * Sz;
      add2_kernel<<<grid1d((long long)nr * Be), 256, 0, s>>>(
          a->g_beliefs ? a->g_beliefs + o * Be : nullptr, have_carry ? w.carry_b : nullptr, w.Gb,
          (long long)nr * Be);
      BD_CUDA_LAUNCH_CHECK();
      BD_TRY(step_backward(r, w, nr, s_prev, nullptr, f.actions + o * A, b_prev, f.beliefs + o * Be,
                           f.eps_s + o * Sz, nullptr, nullptr,
                           a->g_states ? a->g_states + o * Sz : nullptr,
                           have_carry ? w.carry_s : nullptr,
                           a->g_means ? a->g_means + o * Sz : nullptr,
                           a->g_stds ? a->g_stds + o * Sz : nullptr, nullptr, nullptr, nullptr,
                           nullptr, nullptr, nullptr, s));
      have_carry = true;
      if (want_actor) {
        // d raw actor outputs from d action (= dsa[:, S:]) and the entropy cotangent
        actor_head_bwd_kernel<<<grid1d((long long)nr * A), 256, 0, s>>>(
            f.actor_raw + o * 2 * A, f.eps_a + o * A, f.actions + o * A, f.dent + o * 2 * A,
            w.dsa + Sz, Sz + A, a->g_entropy ? a->g_entropy + o : nullptr, f.actor_cfg, nr, (int)A,
            d_raw_all ? d_raw_all + o * 2 * A : d_raw);
        BD_CUDA_LAUNCH_CHECK();
        if (d_raw_all) continue;
        // actor inputs are detached (src/dreamer.py:215): recompute hiddens, wgrad only
        BD_TRY(mlp_forward_rows(actor, b_prev, (int)Be, Be, s_prev, (int)Sz, Sz, nr, hid, d0, 2 * A, s));
        BD_TRY(mlp_backward_rows(actor, b_prev, (int)Be, Be, s_prev, (int)Sz, Sz, nr, hid, d_raw,
                                 2 * A, d0, d1, a->actor_dw, a->actor_db, nullptr, 0, nullptr, 0, s));
      }
    }
    if (a->d_prev_belief)
      cudaMemcpyAsync(a->d_prev_belief + r0 * Be, w.carry_b, (size_t)nr * Be * sizeof(float),
                      cudaMemcpyDeviceToDevice, s);
    if (a->d_prev_state)
      cudaMemcpyAsync(a->d_prev_state + r0 * Sz, w.carry_s, (size_t)nr * Sz * sizeof(float),
                      cudaMemcpyDeviceToDevice, s);
  }
  return BD_OK;
}

// -------------------------------------------------------------------------------------
// CEM (src/planner.py:28-90)
// -------------------------------------------------------------------------------------
static size_t cem_row_floats(const bd_rssm& r, const bd_mlp& reward, int H) {
  // per local candidate row: step scratch + head scratch + double-buffered belief/state +
  // per-step rewards
  return step_row_floats(r, false, false) + mlp_row_floats(reward, false) +
         2 * (size_t)(r.belief_size + r.state_size) + (size_t)H + 3 * r.state_size;
}
size_t cem_workspace_bytes(const bd_rssm* r, const bd_mlp* reward, int B, int C_local, int H) {
  size_t f = (size_t)B * C_local * cem_row_floats(*r, *reward, H) * sizeof(float) + kSlackBytes;
  size_t t = tc::cem_tc_workspace_bytes(*r, *reward, (long long)B * C_local, H) +
             (size_t)B * C_local * H * sizeof(float) + kSlackBytes;
  return f > t ? f : t;
}

int cem_evaluate(const bd_cem_eval_args* a, void* ws, size_t ws_bytes, bd_stream_t stream,
                 int precision, bool weights_packed) {
  const bd_rssm& r = a->rssm;
  BD_TRY(check_rssm(r, false));
  BD_TRY(check_mlp(a->reward, r.belief_size + r.state_size));
  BD_CHECK_ARG(a->reward.layer[a->reward.n_layers - 1].out_features == 1, "cem: reward head must output 1");
  BD_CHECK_ARG(a->B >= 1 && a->C >= 1 && a->H >= 1 && a->c_begin >= 0 && a->c_end <= a->C &&
               a->c_begin < a->c_end, "cem: bad B/C/H or candidate range");
  BD_CHECK_ARG(a->belief && a->state && a->action_mean && a->action_std && a->eps_act && a->eps_s &&
               a->actions && a->returns, "cem: null pointer");
  cudaStream_t s = S(stream);
  const int Be = r.belief_size, Sz = r.state_size, A = r.action_size, H = a->H, B = a->B, C = a->C;
  const int Cl = a->c_end - a->c_begin;
  const long long rows = (long long)B * Cl;
  BD_CHECK_ARG(rows <= kMaxChunkRows * 4LL, "cem: too many candidate rows for one call");
  if (precision != BD_PREC_FP32 && tc::cem_supported(r, a->reward, precision)) {
    // tensor-core path: sample the actions, then one persistent rollout with the reward head fused
    Arena ar(ws, ws_bytes);
    float* rew = ar.f32((size_t)rows * H);
    if (!ar.ok()) BD_FAIL(BD_ERR_WORKSPACE, "cem_evaluate: workspace too small");
    cem_sample_kernel<<<grid1d((long long)H * rows * A), 256, 0, s>>>(
        a->action_mean, a->action_std, a->eps_act, H, B, C, a->c_begin, Cl, A, a->actions);
    BD_CUDA_LAUNCH_CHECK();
    BD_TRY(tc::cem_rollout(a, ar.base + ar.off, ar.cap - ar.off, precision, rew, stream, weights_packed));
    cem_sum_rewards_kernel<<<grid1d(rows), 256, 0, s>>>(rew, H, rows, a->returns);
    BD_CUDA_LAUNCH_CHECK();
    return BD_OK;
  }
  Arena ar(ws, ws_bytes);
  StepBuf w;
  bool ok = carve_step(ar, r, rows, false, false, w);
  const int hw = mlp_max_width(a->reward);
  float* ha = ar.f32(rows * hw);
  float* hb = ar.f32(rows * hw);
  float* bel[2] = {ar.f32(rows * Be), ar.f32(rows * Be)};
  float* sta[2] = {ar.f32(rows * Sz), ar.f32(rows * Sz)};
  float* rew = ar.f32(rows * H);
  float* mean_s = ar.f32(rows * Sz);
  float* std_s = ar.f32(rows * Sz);
  float* eps_l = ar.f32(rows * Sz);
  if (!ok || !ar.ok()) BD_FAIL(BD_ERR_WORKSPACE, "cem_evaluate: workspace too small (%zu bytes)", ws_bytes);
  float* hid[BD_MAX_LAYERS];
  for (int l = 0; l < BD_MAX_LAYERS; ++l) hid[l] = (l & 1) ? hb : ha;

  cem_expand_kernel<<<grid1d(rows * Be), 256, 0, s>>>(a->belief, B, Cl, Be, bel[0]);
  BD_CUDA_LAUNCH_CHECK();
  cem_expand_kernel<<<grid1d(rows * Sz), 256, 0, s>>>(a->state, B, Cl, Sz, sta[0]);
  BD_CUDA_LAUNCH_CHECK();
  cem_sample_kernel<<<grid1d((long long)H * rows * A), 256, 0, s>>>(
      a->action_mean, a->action_std, a->eps_act, H, B, C, a->c_begin, Cl, A, a->actions);
  BD_CUDA_LAUNCH_CHECK();
  for (int h = 0; h < H; ++h) {
    const int cur = h & 1, nxt = cur ^ 1;
    // local slice of eps_s[h]: global row b*C + c  ->  local row b*Cl + (c - c_begin)
    const float* eps_h = a->eps_s + (long long)h * B * C * Sz;
    const float* eps_use = eps_h;
    if (Cl != C) {
      for (int b = 0; b < B; ++b)
        cudaMemcpyAsync(eps_l + (long long)b * Cl * Sz, eps_h + ((long long)b * C + a->c_begin) * Sz,
                        (size_t)Cl * Sz * sizeof(float), cudaMemcpyDeviceToDevice, s);
      eps_use = eps_l;
    }
    BD_TRY(step_forward(r, w, (int)rows, sta[cur], nullptr, a->actions + (long long)h * rows * A,
                        bel[cur], eps_use, nullptr, nullptr, bel[nxt], sta[nxt], mean_s, std_s,
                        nullptr, nullptr, nullptr, s));
    BD_TRY(mlp_forward_rows(a->reward, bel[nxt], Be, Be, sta[nxt], Sz, Sz, (int)rows, hid,
                            rew + (long long)h * rows, 1, s));
  }
  cem_sum_rewards_kernel<<<grid1d(rows), 256, 0, s>>>(rew, H, rows, a->returns);
  BD_CUDA_LAUNCH_CHECK();
  return BD_OK;
}

int cem_refit(const float* returns, const float* actions, int B, int C, int K, int H, int A,
              int64_t* topk_idx, float* action_mean, float* action_std, bd_stream_t stream) {
  BD_CHECK_ARG(B >= 1 && C >= 1 && K >= 1 && K <= C && H >= 1 && A >= 1, "cem_refit: bad sizes");
  BD_CHECK_ARG(returns && actions && action_mean && action_std, "cem_refit: null pointer");
  size_t P = 1;
  while (P < (size_t)C) P <<= 1;
  size_t smem = P * sizeof(unsigned long long) + (size_t)C * sizeof(int) + (size_t)K * sizeof(int);
  BD_CHECK_ARG(smem <= 200 * 1024, "cem_refit: candidates=%d too large for one CTA", C);
  set_smem_attr(cem_refit_kernel, 200 * 1024);     // per (device, kernel), grows only
  cem_refit_kernel<<<B, CEM_REFIT_THREADS, smem, S(stream)>>>(
      returns, actions, B, C, K, H, A, reinterpret_cast<long long*>(topk_idx), action_mean, action_std);
  BD_CUDA_LAUNCH_CHECK();
  return BD_OK;
}

__global__ void fill_kernel(float* p, float v, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

size_t cem_plan_workspace_bytes(const bd_rssm* r, const bd_mlp* reward, int B, int C, int K, int H) {
  (void)K;
  size_t A = r->action_size;
  return cem_workspace_bytes(r, reward, B, C, H) +
         ((size_t)H * B * C * A + (size_t)B * C + 2 * (size_t)H * B * A) * sizeof(float) + kSlackBytes;
}

int cem_plan(const bd_cem_plan_args* a, void* ws, size_t ws_bytes, bd_stream_t stream, int precision) {
  const bd_rssm& r = a->rssm;
  BD_CHECK_ARG(a->iters >= 1 && a->K >= 1 && a->K <= a->C, "cem_plan: bad iters/K");
  BD_CHECK_ARG(a->action_out && a->eps_act && a->eps_s, "cem_plan: null pointer");
  cudaStream_t s = S(stream);
  const int A = r.action_size, B = a->B, C = a->C, H = a->H;
  Arena ar(ws, ws_bytes);
  float* actions = ar.f32((size_t)H * B * C * A);
  float* returns = ar.f32((size_t)B * C);
  float* mean = ar.f32((size_t)H * B * A);
  float* stdv = ar.f32((size_t)H * B * A);
  if (!ar.ok()) BD_FAIL(BD_ERR_WORKSPACE, "cem_plan: workspace too small");
  void* sub = ar.base + ar.off;
  size_t sub_bytes = ar.cap - ar.off;
  const long long nm = (long long)H * B * A;
  fill_kernel<<<grid1d(nm), 256, 0, s>>>(mean, 0.f, nm);    // src/planner.py:42-47
  BD_CUDA_LAUNCH_CHECK();
  fill_kernel<<<grid1d(nm), 256, 0, s>>>(stdv, 1.f, nm);
  BD_CUDA_LAUNCH_CHECK();
  for (int it = 0; it < a->iters; ++it) {
    bd_cem_eval_args e;
    e.rssm = r; e.reward = a->reward; e.B = B; e.C = C; e.H = H; e.c_begin = 0; e.c_end = C;
    e.belief = a->belief; e.state = a->state; e.action_mean = mean; e.action_std = stdv;
    e.eps_act = a->eps_act + (long long)it * H * B * C * A;
    e.eps_s = a->eps_s + (long long)it * H * B * C * r.state_size;
    e.actions = actions; e.returns = returns;
    // the weights do not change inside a plan: their 16-bit images are packed by the first iteration only
    BD_TRY(cem_evaluate(&e, sub, sub_bytes, stream, precision, it > 0));
    if (a->returns_trace)
      cudaMemcpyAsync(a->returns_trace + (long long)it * B * C, returns, (size_t)B * C * sizeof(float),
                      cudaMemcpyDeviceToDevice, s);
    BD_TRY(cem_refit(returns, actions, B, C, a->K, H, A,
                     a->topk_trace ? a->topk_trace + (long long)it * B * a->K : nullptr, mean, stdv,
                     stream));
  }
  // first action mean                                                    src/planner.py:90
  cudaMemcpyAsync(a->action_out, mean, (size_t)B * A * sizeof(float), cudaMemcpyDeviceToDevice, s);
  return BD_OK;
}

}  // namespace f32
}  // namespace bd
