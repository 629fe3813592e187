// extern "C" surface of libbd_b200.so (include/bd_b200.h): argument checks that are common to
// all arithmetic modes and dispatch on bd_precision.  No CPU fallback anywhere: a precision
// that this build does not implement returns BD_ERR_UNSUPPORTED.
#include <atomic>
#include <map>
#include <mutex>
#include <unordered_map>
#include <stdlib.h>
#include "api_internal.h"

namespace bd {
static thread_local char g_err[1024] = "";
std::atomic<unsigned long long> g_launch_count{0};

// function attributes are per DEVICE: the cache is keyed by (device, kernel)
void grow_smem_attr(const void* kernel, int bytes) {
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, int> cur;
  static const bool always_max = [] { const char* e = getenv("BD_SMEM_ATTR"); return e && e[0] == 'm'; }();
  if (always_max) bytes = kMaxOptinSmem;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  int& c = cur[{dev, kernel}];
  if (bytes > c) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    c = bytes;
  }
}
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace bd

#include <vector>
namespace bd {
namespace {
struct ProfState {
  std::mutex mu;                 // autograd runs backward on another host thread
  bool on = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev[BD_PROF_COUNT];
  std::vector<cudaEvent_t> pool;
  cudaEvent_t open_start[BD_PROF_COUNT] = {};
} g_prof;
cudaEvent_t prof_event() {
  if (!g_prof.pool.empty()) { cudaEvent_t e = g_prof.pool.back(); g_prof.pool.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
}  // namespace
bool prof_enabled() { return g_prof.on; }
void prof_begin(int k, cudaStream_t s) {
  if (!g_prof.on || k < 0 || k >= BD_PROF_COUNT) return;
  std::lock_guard<std::mutex> lock(g_prof.mu);
  cudaEvent_t e = prof_event();
  cudaEventRecord(e, s);
  g_prof.open_start[k] = e;
}
void prof_end(int k, cudaStream_t s) {
  if (!g_prof.on || k < 0 || k >= BD_PROF_COUNT) return;
  std::lock_guard<std::mutex> lock(g_prof.mu);
  if (!g_prof.open_start[k]) return;
  cudaEvent_t e = prof_event();
  cudaEventRecord(e, s);
  g_prof.ev[k].push_back({g_prof.open_start[k], e});
  g_prof.open_start[k] = nullptr;
}
}  // namespace bd

using namespace bd;

#define BD_NEED(p, what)                                           \
  do {                                                             \
    if (!(p)) BD_FAIL(BD_ERR_BAD_ARG, "%s: null %s", __func__, what); \
  } while (0)

// Entry points that have no tensor-core implementation (yet) run their fp32 CUDA kernels when a
// 16-bit mode is requested: higher precision, same device path, never a CPU fallback.
#define BD_ONLY_FP32(precision)                                                             \
  do {                                                                                      \
    if ((precision) != BD_PREC_FP32 && (precision) != BD_PREC_FP16 &&                        \
        (precision) != BD_PREC_BF16)                                                        \
      BD_FAIL(BD_ERR_UNSUPPORTED, "%s: precision %d is not implemented for this entry point", \
              __func__, (int)(precision));                                                  \
  } while (0)

extern "C" {

int bd_version(void) { return BD_ABI_VERSION; }
const char* bd_last_error(void) { return bd::g_err; }
unsigned long long bd_launch_count(void) { return bd::g_launch_count.load(); }
void bd_prof_enable(int on) { bd::g_prof.on = on != 0; }
int bd_prof_read(int kernel, float* ms_total, int* launches) {
  if (kernel < 0 || kernel >= BD_PROF_COUNT) BD_FAIL(BD_ERR_BAD_ARG, "bd_prof_read: bad kernel id");
  float tot = 0.f;
  int n = 0;
  std::lock_guard<std::mutex> lock(bd::g_prof.mu);
  for (auto& pr : bd::g_prof.ev[kernel]) {
    float ms = 0.f;
    cudaEventSynchronize(pr.second);
    if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) { tot += ms; ++n; }
    bd::g_prof.pool.push_back(pr.first);
    bd::g_prof.pool.push_back(pr.second);
  }
  bd::g_prof.ev[kernel].clear();
  if (ms_total) *ms_total = tot;
  if (launches) *launches = n;
  return BD_OK;
}
int bd_precision_supported(int precision) {
  return (precision == BD_PREC_FP32 || precision == BD_PREC_FP16 || precision == BD_PREC_BF16) ? 1 : 0;
}

size_t bd_mlp_workspace_bytes(const bd_mlp* m, int64_t rows, int backward) {
  if (!m) return 0;
  size_t f = f32::mlp_workspace_bytes(m, rows, backward), t = tc::mlp_pack_bytes(*m);
  if (backward && m->n_layers >= 1) {
    const int kin = m->layer[0].in_features;
    size_t tb = tc::mlp_backward_workspace_bytes(*m, kin, 0, rows);
    if (tb > t) t = tb;
  }
  return f > t ? f : t;
}
int bd_mlp_forward(const bd_mlp* m, const float* x1, int k1, const float* x2, int k2, int64_t rows,
                   float* y, void* ws, size_t ws_bytes, int precision, bd_stream_t stream) {
  BD_NEED(m, "mlp");
  if (rows == 0) return BD_OK;
  BD_NEED(x1, "x1"); BD_NEED(y, "y"); BD_NEED(ws, "workspace");
  BD_CHECK_ARG(k1 > 0 && k2 >= 0 && (k2 == 0 || x2), "bd_mlp_forward: bad k1/k2/x2");
  BD_ONLY_FP32(precision);
  BD_TRY(f32::check_mlp(*m, k1 + k2));
  if (tc::mlp_supported(*m, k1, k2, precision))
    return tc::mlp_forward(m, x1, k1, x2, k2, rows, y, ws, ws_bytes, precision, stream);
  return f32::mlp_forward(m, x1, k1, x2, k2, rows, y, ws, ws_bytes, stream);
}
size_t bd_mlp_saved_bytes(const bd_mlp* m, int k1, int k2, int64_t rows, int precision) {
  if (!m || !tc::mlp_backward_supported(*m, k1, k2, precision)) return 0;
  return tc::mlp_saved_bytes(*m, rows);
}
int bd_mlp_forward_save(const bd_mlp* m, const float* x1, int k1, const float* x2, int k2,
                        int64_t rows, float* y, void* saved, void* ws, size_t ws_bytes,
                        int precision, bd_stream_t stream) {
  BD_NEED(m, "mlp");
  if (rows == 0) return BD_OK;
  BD_NEED(x1, "x1"); BD_NEED(y, "y"); BD_NEED(ws, "workspace");
  BD_CHECK_ARG(k1 > 0 && k2 >= 0 && (k2 == 0 || x2), "bd_mlp_forward_save: bad k1/k2/x2");
  BD_ONLY_FP32(precision);
  BD_TRY(f32::check_mlp(*m, k1 + k2));
  if (tc::mlp_supported(*m, k1, k2, precision))
    return tc::mlp_forward(m, x1, k1, x2, k2, rows, y, ws, ws_bytes, precision, stream,
                           tc::mlp_backward_supported(*m, k1, k2, precision) ? saved : nullptr);
  return f32::mlp_forward(m, x1, k1, x2, k2, rows, y, ws, ws_bytes, stream);
}
int bd_mlp_backward(const bd_mlp* m, const bd_mlp_bwd_args* a, void* ws, size_t ws_bytes,
                    int precision, bd_stream_t stream) {
  BD_NEED(m, "mlp"); BD_NEED(a, "args");
  if (a->rows == 0) return BD_OK;
  BD_NEED(ws, "workspace"); BD_NEED(a->x1, "x1"); BD_NEED(a->dy, "dy");
  BD_CHECK_ARG(a->k1 > 0 && a->k2 >= 0 && (a->k2 == 0 || a->x2), "bd_mlp_backward: bad k1/k2/x2");
  BD_ONLY_FP32(precision);
  BD_TRY(f32::check_mlp(*m, a->k1 + a->k2));
  if (tc::mlp_backward_supported(*m, a->k1, a->k2, precision))
    return tc::mlp_backward(m, a, ws, ws_bytes, precision, stream);
  return f32::mlp_backward(m, a, ws, ws_bytes, stream);
}

int bd_heads_forward_supported(const bd_mlp* reward, const bd_mlp* value, int k1, int k2, int precision) {
  if (!(reward && value) || k1 <= 0 || k2 <= 0) return 0;
  if (f32::check_mlp(*reward, k1 + k2) != BD_OK || f32::check_mlp(*value, k1 + k2) != BD_OK) return 0;
  return (tc::heads_pair_supported(*reward, *value, k1, k2, precision) &&
          tc::mlp_backward_supported(*reward, k1, k2, precision) &&
          tc::mlp_backward_supported(*value, k1, k2, precision)) ? 1 : 0;
}
size_t bd_heads_forward_workspace_bytes(const bd_mlp* reward, const bd_mlp* value) {
  return (reward && value) ? tc::heads_pair_pack_bytes(*reward, *value) + 65536 : 0;
}
int bd_heads_forward(const bd_mlp* reward, const bd_mlp* value, const float* x1, int k1, const float* x2,
                     int k2, int64_t rows, float* y_reward, float* y_value, void* saved_reward,
                     void* saved_value, void* ws, size_t ws_bytes, int precision, bd_stream_t stream) {
  BD_NEED(reward, "reward"); BD_NEED(value, "value");
  if (rows == 0) return BD_OK;
  BD_NEED(x1, "x1"); BD_NEED(x2, "x2"); BD_NEED(y_reward, "y_reward"); BD_NEED(y_value, "y_value");
  BD_NEED(ws, "workspace");
  BD_CHECK_ARG(k1 > 0 && k2 > 0, "bd_heads_forward: bad k1/k2");
  BD_TRY(f32::check_mlp(*reward, k1 + k2));
  BD_TRY(f32::check_mlp(*value, k1 + k2));
  if (!bd_heads_forward_supported(reward, value, k1, k2, precision))
    BD_FAIL(BD_ERR_UNSUPPORTED, "bd_heads_forward: configuration not supported (see bd_heads_forward_supported)");
  return tc::heads_pair_forward(reward, value, x1, k1, x2, k2, rows, y_reward, y_value, saved_reward, saved_value,
                                ws, ws_bytes, precision, stream);
}

size_t bd_heads_backward_workspace_bytes(const bd_mlp* reward, const bd_mlp* value, int k1, int k2) {
  return (reward && value) ? tc::heads_pair_backward_workspace_bytes(*reward, *value, k1, k2) : 0;
}
int bd_heads_backward(const bd_mlp* reward, const bd_mlp* value, int k1, int k2, int64_t rows,
                      const float* dy_reward, const float* dy_value, const void* saved_reward,
                      const void* saved_value, float* dx1, float* dx2, void* ws, size_t ws_bytes,
                      int precision, bd_stream_t stream) {
  BD_NEED(reward, "reward"); BD_NEED(value, "value");
  if (rows == 0) return BD_OK;
  BD_NEED(dy_reward, "dy_reward"); BD_NEED(dy_value, "dy_value"); BD_NEED(ws, "workspace");
  BD_NEED(saved_reward, "saved_reward"); BD_NEED(saved_value, "saved_value");
  BD_CHECK_ARG(dx1 || dx2, "bd_heads_backward: no output requested");
  BD_CHECK_ARG(k1 > 0 && k2 > 0, "bd_heads_backward: bad k1/k2");
  if (!bd_heads_forward_supported(reward, value, k1, k2, precision))
    BD_FAIL(BD_ERR_UNSUPPORTED, "bd_heads_backward: configuration not supported (see bd_heads_forward_supported)");
  return tc::heads_pair_backward(reward, value, k1, k2, rows, dy_reward, dy_value, saved_reward, saved_value, dx1,
                                 dx2, ws, ws_bytes, precision, stream);
}

int bd_lambda_return_forward(const float* reward, const float* value, const float* bootstrap, int T,
                             int64_t N, double discount, double lambda_, float* returns,
                             bd_stream_t stream) {
  BD_NEED(reward, "reward"); BD_NEED(value, "value"); BD_NEED(bootstrap, "bootstrap");
  BD_NEED(returns, "returns");
  return f32::lambda_return_forward(reward, value, bootstrap, T, N, discount, lambda_, returns, stream);
}
int bd_lambda_return_backward(const float* d_returns, int T, int64_t N, double discount,
                              double lambda_, float* d_reward, float* d_value, float* d_bootstrap,
                              bd_stream_t stream) {
  BD_NEED(d_returns, "d_returns");
  return f32::lambda_return_backward(d_returns, T, N, discount, lambda_, d_reward, d_value,
                                     d_bootstrap, stream);
}

int bd_kl_loss_forward(const float* post_mean, const float* post_std, const float* prior_mean,
                       const float* prior_std, int64_t rows, int S, const float* free_nats,
                       double balance, float* div, float* loss, bd_stream_t stream) {
  return f32::kl_loss_forward(post_mean, post_std, prior_mean, prior_std, rows, S, free_nats, balance, div,
                              loss, stream);
}
int bd_kl_loss_backward(const float* post_mean, const float* post_std, const float* prior_mean,
                        const float* prior_std, int64_t rows, int S, const float* free_nats,
                        double balance, const float* div, const float* loss, const float* g_loss,
                        float* d_post_mean, float* d_post_std, float* d_prior_mean,
                        float* d_prior_std, bd_stream_t stream) {
  return f32::kl_loss_backward(post_mean, post_std, prior_mean, prior_std, rows, S, free_nats, balance, div,
                               loss, g_loss, d_post_mean, d_post_std, d_prior_mean, d_prior_std, stream);
}

int bd_value_loss(const float* value, const float* target, const float* weight, int64_t n, float* loss,
                  float* d_value, void* ws, size_t ws_bytes, bd_stream_t stream) {
  return f32::value_loss(value, target, weight, n, loss, d_value, ws, ws_bytes, stream);
}

int bd_actor_act(const float* raw, const float* eps, const bd_actor_cfg* cfg, int64_t rows, int action_size,
                 int deterministic, float* action, bd_stream_t stream) {
  return f32::actor_act(raw, eps, cfg, rows, action_size, deterministic, action, stream);
}

size_t bd_transition_workspace_bytes(const bd_rssm* r, int L, int64_t B, int observe, int backward) {
  if (!r) return 0;
  size_t f = f32::transition_workspace_bytes(r, L, B, observe, backward);
  bd_mlp none{};
  size_t t = tc::imagine_pack_bytes(*r, none);
  return f > t ? f : t;
}
int bd_transition_forward(const bd_transition_args* a, void* ws, size_t ws_bytes, int precision,
                          bd_stream_t stream) {
  BD_NEED(a, "args"); BD_NEED(ws, "workspace");
  BD_ONLY_FP32(precision);
  BD_TRY(f32::check_rssm(a->rssm, a->embeddings != nullptr));
  BD_CHECK_ARG(a->L >= 1 && a->B >= 0, "bd_transition_forward: bad L/B");
  if (a->B == 0) return BD_OK;
  if (tc::transition_supported(*a, precision) && a->init_state && a->init_belief &&
      a->actions && a->eps_prior && a->beliefs && a->prior_states && a->prior_means && a->prior_stds)
    return tc::transition_forward(a, ws, ws_bytes, precision, stream);
  return f32::transition_forward(a, ws, ws_bytes, stream, precision);
}
int bd_transition_backward(const bd_transition_bwd_args* a, void* ws, size_t ws_bytes, int precision,
                           bd_stream_t stream) {
  BD_NEED(a, "args"); BD_NEED(ws, "workspace");
  BD_ONLY_FP32(precision);
  return f32::transition_backward(a, ws, ws_bytes, stream, precision);
}

size_t bd_imagine_workspace_bytes(const bd_rssm* r, const bd_mlp* actor, int T, int64_t N,
                                  int backward) {
  if (!(r && actor)) return 0;
  size_t f = f32::imagine_workspace_bytes(r, actor, T, N, backward);
  size_t t = tc::imagine_pack_bytes(*r, *actor);
  if (backward) {   // tensor-core modes: d_raw (T,N,2A) + the larger of the BPTT / actor-MLP workspaces
    size_t mb = tc::mlp_backward_workspace_bytes(*actor, r->belief_size, r->state_size, (int64_t)T * N);
    size_t draw = ((size_t)T * N * 2 * r->action_size * sizeof(float) + 255) & ~size_t(255);
    size_t bp = tc::bptt_workspace_bytes(*r, T, N);
    if (bp > mb) mb = bp;
    t = draw + (mb > f ? mb : f);
  }
  return f > t ? f : t;
}
size_t bd_imagine_saved_bytes(const bd_rssm* r, int T, int64_t N, int precision) {
  if (!r || (precision != BD_PREC_FP16 && precision != BD_PREC_BF16)) return 0;
  return tc::imagine_saved_bytes(*r, T, N);
}
int bd_imagine_forward(const bd_imagine_args* a, void* ws, size_t ws_bytes, int precision,
                       bd_stream_t stream) {
  BD_NEED(a, "args"); BD_NEED(ws, "workspace");
  BD_ONLY_FP32(precision);
  BD_TRY(f32::check_imagine(*a));
  if (a->N == 0) return BD_OK;
  // 16-bit modes: the tensor-core rollout when the configuration fits it, else the fp32 kernels
  // (higher precision, same device path; the backward makes the same decision)
  if ((precision == BD_PREC_FP16 || precision == BD_PREC_BF16) &&
      tc::imagine_supported(a->rssm, a->actor, precision))
    return tc::imagine_forward(a, ws, ws_bytes, precision, stream);
  BD_ONLY_FP32(precision);
  return f32::imagine_forward(a, ws, ws_bytes, stream);
}
int bd_imagine_backward(const bd_imagine_bwd_args* a, void* ws, size_t ws_bytes, int precision,
                        bd_stream_t stream) {
  BD_NEED(a, "args"); BD_NEED(ws, "workspace");
  BD_ONLY_FP32(precision);
  const bd_imagine_args& f = a->fwd;
  BD_TRY(f32::check_imagine(f));
  if (f.N == 0) return BD_OK;
  bool want_actor = false;
  for (int l = 0; l < f.actor.n_layers; ++l) want_actor |= (a->actor_dw[l] || a->actor_db[l]);
  if (precision != BD_PREC_FP32 && want_actor && f.T >= 1 && f.N > 0 &&
      tc::mlp_backward_supported(f.actor, f.rssm.belief_size, f.rssm.state_size, precision)) {
    // BPTT dgrad chain (fp32 kernels) -> d_raw for every step; then the actor's recompute + dgrad +
    // wgrad as one batched tensor-core MLP backward
    const int A = f.rssm.action_size, Be = f.rssm.belief_size, S = f.rssm.state_size;
    const size_t draw_bytes = ((size_t)f.T * f.N * 2 * A * sizeof(float) + 255) & ~size_t(255);
    BD_CHECK_ARG(ws_bytes > draw_bytes + 65536, "bd_imagine_backward: workspace too small");
    float* d_raw = static_cast<float*>(ws);
    void* rest = static_cast<char*>(ws) + draw_bytes;
    const size_t rest_bytes = ws_bytes - draw_bytes;
    if (f.tc_saved && tc::imagine_supported(f.rssm, f.actor, precision))
      BD_TRY(tc::imagine_bptt(a, d_raw, rest, rest_bytes, precision, stream));
    else
      BD_TRY(f32::imagine_backward_ex(a, rest, rest_bytes, stream, d_raw));
    bd_mlp_bwd_args m{};
    m.k1 = Be; m.k2 = S;
    for (int l = 0; l < f.actor.n_layers; ++l) { m.dw[l] = a->actor_dw[l]; m.db[l] = a->actor_db[l]; }
    // ONE batched backward over all T*N rows: step 0 reads (prev_belief, prev_state), steps 1.. read
    // (beliefs, states)[t-1] -- a two-segment input (two passes cost a second, 20-CTA launch chain)
    m.x1 = f.prev_belief; m.x2 = f.prev_state; m.rows = (int64_t)f.T * f.N; m.dy = d_raw;
    // (hidden activations saved per (t, tile) by the tensor-core forward: no recompute, step-wise tiling)
    m.saved = tc::imagine_saved_actor(f.rssm, f.actor, f.T, f.N, f.tc_saved);
    // (and its layer-0 input images, copied out of the rollout's operand tiles)
    const void* x0s = nullptr;
    const void* x0b = tc::imagine_saved_actor_x0(f.rssm, f.actor, f.T, f.N, f.tc_saved, &x0s);
    BD_TRY(tc::mlp_backward(&f.actor, &m, rest, rest_bytes, precision, stream, f.beliefs, f.states, f.N,
                            m.saved ? f.N : 0, x0b, x0s));
    return BD_OK;
  }
  return f32::imagine_backward(a, ws, ws_bytes, stream);
}

int bd_imagine_returns_supported(const bd_rssm* r, const bd_mlp* actor, const bd_mlp* reward,
                                 const bd_mlp* value, int precision) {
  if (!(r && actor && reward && value)) return 0;
  return (tc::imagine_supported(*r, *actor, precision) && tc::heads_supported(*r, *reward, *value)) ? 1 : 0;
}
size_t bd_imagine_returns_workspace_bytes(const bd_imagine_returns_args* a, int backward) {
  if (!a) return 0;
  return bd_imagine_workspace_bytes(&a->img.rssm, &a->img.actor, a->img.T, a->img.N, backward) +
         tc::heads_pack_bytes(a->reward, a->value) + (backward ? tc::heads_bwd_workspace_bytes(a->img.rssm, a->reward, a->img.T, a->img.N) : 0) + 65536;
}
size_t bd_imagine_returns_saved_bytes(const bd_imagine_returns_args* a) {
  return a ? tc::heads_saved_bytes(a->reward, a->img.T, a->img.N) : 0;
}
int bd_imagine_returns_forward(const bd_imagine_returns_args* a, void* ws, size_t ws_bytes, int precision,
                               bd_stream_t stream) {
  BD_NEED(a, "args"); BD_NEED(ws, "workspace");
  BD_ONLY_FP32(precision);
  BD_TRY(f32::check_imagine(a->img));
  const int kin = a->img.rssm.belief_size + a->img.rssm.state_size;
  BD_TRY(f32::check_mlp(a->reward, kin));
  BD_TRY(f32::check_mlp(a->value, kin));
  BD_CHECK_ARG(a->reward_out && a->value_out && a->returns, "bd_imagine_returns_forward: null output");
  if (a->img.N == 0) return BD_OK;
  return tc::imagine_returns_forward(&a->img, &a->reward, &a->value, a->discount, a->lambda_, a->reward_out,
                                     a->value_out, a->returns, a->heads_saved, ws, ws_bytes, precision, stream);
}
int bd_imagine_returns_backward(const bd_imagine_returns_bwd_args* a, void* ws, size_t ws_bytes,
                                int precision, bd_stream_t stream) {
  BD_NEED(a, "args"); BD_NEED(ws, "workspace");
  BD_ONLY_FP32(precision);
  const bd_imagine_args& f = a->fwd.img;
  BD_TRY(f32::check_imagine(f));
  if (f.N == 0) return BD_OK;
  BD_CHECK_ARG(f.tc_saved && a->fwd.heads_saved, "bd_imagine_returns_backward: the forward did not save its state");
  if (!bd_imagine_returns_supported(&f.rssm, &f.actor, &a->fwd.reward, &a->fwd.value, precision))
    BD_FAIL(BD_ERR_UNSUPPORTED, "bd_imagine_returns_backward: configuration not supported");
  // d_raw (T,N,2A) | rest: fused BPTT (lambda-return adjoint + heads dgrad + recurrence), then the actor's
  // batched backward with weight gradients
  const int A = f.rssm.action_size, Be = f.rssm.belief_size, S = f.rssm.state_size;
  const size_t draw_bytes = ((size_t)f.T * f.N * 2 * A * sizeof(float) + 255) & ~size_t(255);
  BD_CHECK_ARG(ws_bytes > draw_bytes + 65536, "bd_imagine_returns_backward: workspace too small");
  float* d_raw = static_cast<float*>(ws);
  void* rest = static_cast<char*>(ws) + draw_bytes;
  const size_t rest_bytes = ws_bytes - draw_bytes;
  BD_TRY(tc::imagine_returns_bptt(a, d_raw, rest, rest_bytes, precision, stream));
  bool want_actor = false;
  for (int l = 0; l < f.actor.n_layers; ++l) want_actor |= (a->actor_dw[l] || a->actor_db[l]);
  if (!want_actor) return BD_OK;
  bd_mlp_bwd_args m{};
  m.k1 = Be; m.k2 = S;
  for (int l = 0; l < f.actor.n_layers; ++l) { m.dw[l] = a->actor_dw[l]; m.db[l] = a->actor_db[l]; }
  m.x1 = f.prev_belief; m.x2 = f.prev_state; m.rows = (int64_t)f.T * f.N; m.dy = d_raw;
  m.saved = tc::imagine_saved_actor(f.rssm, f.actor, f.T, f.N, f.tc_saved);
  const void* x0s = nullptr;
  const void* x0b = tc::imagine_saved_actor_x0(f.rssm, f.actor, f.T, f.N, f.tc_saved, &x0s);
  return tc::mlp_backward(&f.actor, &m, rest, rest_bytes, precision, stream, f.beliefs, f.states, f.N,
                          m.saved ? f.N : 0, x0b, x0s);
}

size_t bd_cem_workspace_bytes(const bd_rssm* r, const bd_mlp* reward, int B, int C_local, int H) {
  return (r && reward) ? f32::cem_workspace_bytes(r, reward, B, C_local, H) : 0;
}
int bd_cem_evaluate(const bd_cem_eval_args* a, void* ws, size_t ws_bytes, int precision,
                    bd_stream_t stream) {
  BD_NEED(a, "args"); BD_NEED(ws, "workspace");
  BD_ONLY_FP32(precision);
  return f32::cem_evaluate(a, ws, ws_bytes, stream, precision);
}
int bd_cem_refit(const float* returns, const float* actions, int B, int C, int K, int H, int A,
                 int64_t* topk_idx, float* action_mean, float* action_std, bd_stream_t stream) {
  return f32::cem_refit(returns, actions, B, C, K, H, A, topk_idx, action_mean, action_std, stream);
}
size_t bd_cem_plan_workspace_bytes(const bd_rssm* r, const bd_mlp* reward, int B, int C, int K, int H) {
  return (r && reward) ? f32::cem_plan_workspace_bytes(r, reward, B, C, K, H) : 0;
}
int bd_cem_plan(const bd_cem_plan_args* a, void* ws, size_t ws_bytes, int precision,
                bd_stream_t stream) {
  BD_NEED(a, "args"); BD_NEED(ws, "workspace");
  BD_ONLY_FP32(precision);
  return f32::cem_plan(a, ws, ws_bytes, stream, precision);
}

}  // extern "C"
