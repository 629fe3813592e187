// Internal C++ entry points behind the C ABI, one namespace per arithmetic mode.
#pragma once
#include "common.cuh"

namespace bd {

constexpr int64_t kMaxChunkRows = 65536;   // rows processed per pass of the check-mode path
constexpr size_t kSlackBytes = 65536;      // alignment slack for workspace carving

namespace f32 {
// argument validation shared by every arithmetic mode (run in api.cu BEFORE the precision dispatch)
int check_mlp(const bd_mlp& m, int in_features);
int check_rssm(const bd_rssm& r, bool need_post);
int check_imagine(const bd_imagine_args& a);
size_t mlp_workspace_bytes(const bd_mlp* m, int64_t rows, int backward);
int mlp_forward(const bd_mlp* m, const float* x1, int k1, const float* x2, int k2, int64_t rows,
                float* y, void* ws, size_t ws_bytes, bd_stream_t stream);
int mlp_backward(const bd_mlp* m, const bd_mlp_bwd_args* a, void* ws, size_t ws_bytes,
                 bd_stream_t stream);
int lambda_return_forward(const float* reward, const float* value, const float* bootstrap, int T,
                          int64_t N, double discount, double lambda_, float* returns,
                          bd_stream_t stream);
int lambda_return_backward(const float* d_returns, int T, int64_t N, double discount,
                           double lambda_, float* d_reward, float* d_value, float* d_bootstrap,
                           bd_stream_t stream);
int kl_loss_forward(const float* post_mean, const float* post_std, const float* prior_mean,
                    const float* prior_std, int64_t rows, int S, const float* free_nats, double balance,
                    float* div, float* loss, bd_stream_t stream);
int kl_loss_backward(const float* post_mean, const float* post_std, const float* prior_mean,
                     const float* prior_std, int64_t rows, int S, const float* free_nats, double balance,
                     const float* div, const float* loss, const float* g_loss, float* d_post_mean,
                     float* d_post_std, float* d_prior_mean, float* d_prior_std, bd_stream_t stream);
int value_loss(const float* value, const float* target, const float* weight, int64_t n, float* loss,
               float* d_value, void* ws, size_t ws_bytes, bd_stream_t stream);
int actor_act(const float* raw, const float* eps, const bd_actor_cfg* cfg, int64_t rows, int A,
              int deterministic, float* action, bd_stream_t stream);
size_t transition_workspace_bytes(const bd_rssm* r, int L, int64_t B, int observe, int backward);
int transition_forward(const bd_transition_args* a, void* ws, size_t ws_bytes, bd_stream_t stream,
                       int precision = BD_PREC_FP32);
int transition_backward(const bd_transition_bwd_args* a, void* ws, size_t ws_bytes,
                        bd_stream_t stream, int precision = BD_PREC_FP32);
size_t imagine_workspace_bytes(const bd_rssm* r, const bd_mlp* actor, int T, int64_t N, int backward);
int imagine_forward(const bd_imagine_args* a, void* ws, size_t ws_bytes, bd_stream_t stream);
int imagine_backward(const bd_imagine_bwd_args* a, void* ws, size_t ws_bytes, bd_stream_t stream);
int imagine_backward_ex(const bd_imagine_bwd_args* a, void* ws, size_t ws_bytes, bd_stream_t stream,
                        float* d_raw_all);
size_t cem_workspace_bytes(const bd_rssm* r, const bd_mlp* reward, int B, int C_local, int H);
int cem_evaluate(const bd_cem_eval_args* a, void* ws, size_t ws_bytes, bd_stream_t stream,
                 int precision = BD_PREC_FP32, bool weights_packed = false);
int cem_refit(const float* returns, const float* actions, int B, int C, int K, int H, int A,
              int64_t* topk_idx, float* action_mean, float* action_std, bd_stream_t stream);
size_t cem_plan_workspace_bytes(const bd_rssm* r, const bd_mlp* reward, int B, int C, int K, int H);
int cem_plan(const bd_cem_plan_args* a, void* ws, size_t ws_bytes, bd_stream_t stream,
             int precision = BD_PREC_FP32);
}  // namespace f32

namespace tc {   // tensor-core (tcgen05) path: 16-bit operands, fp32 accumulation and state
bool imagine_supported(const bd_rssm& r, const bd_mlp& actor, int precision);
size_t imagine_pack_bytes(const bd_rssm& r, const bd_mlp& actor);
int imagine_forward(const bd_imagine_args* a, void* ws, size_t ws_bytes, int precision,
                    bd_stream_t stream);
size_t imagine_saved_bytes(const bd_rssm& r, int T, long long N);
const void* imagine_saved_actor(const bd_rssm& r, const bd_mlp& actor, int T, long long N, const void* tc_saved);
bool heads_pair_supported(const bd_mlp& reward, const bd_mlp& value, int k1, int k2, int precision);
size_t heads_pair_pack_bytes(const bd_mlp& reward, const bd_mlp& value);
int heads_pair_forward(const bd_mlp* reward, const bd_mlp* value, const float* x1, int k1, const float* x2, int k2,
                       int64_t rows, float* y_reward, float* y_value, void* saved_reward, void* saved_value,
                       void* ws, size_t ws_bytes, int precision, bd_stream_t stream);
size_t heads_pair_backward_workspace_bytes(const bd_mlp& reward, const bd_mlp& value, int k1, int k2);
int heads_pair_backward(const bd_mlp* reward, const bd_mlp* value, int k1, int k2, int64_t rows,
                        const float* dy_reward, const float* dy_value, const void* saved_reward,
                        const void* saved_value, float* dx1, float* dx2, void* ws, size_t ws_bytes, int precision,
                        bd_stream_t stream);
const void* imagine_saved_actor_x0(const bd_rssm& r, const bd_mlp& actor, int T, long long N, const void* tc_saved,
                                   const void** x0s);
// fused imagine + reward/value heads + lambda_return (SURVEY 8b level L2)
bool heads_supported(const bd_rssm& r, const bd_mlp& reward, const bd_mlp& value);
size_t heads_pack_bytes(const bd_mlp& reward, const bd_mlp& value);
size_t heads_saved_bytes(const bd_mlp& reward, int T, long long N);
size_t heads_bwd_workspace_bytes(const bd_rssm& r, const bd_mlp& reward, int T, long long N);
int imagine_returns_bptt(const bd_imagine_returns_bwd_args* a, float* d_raw, void* ws, size_t ws_bytes,
                         int precision, bd_stream_t stream);
int imagine_returns_forward(const bd_imagine_args* a, const bd_mlp* reward, const bd_mlp* value,
                            double discount, double lambda_, float* reward_out, float* value_out,
                            float* returns, void* heads_saved, void* ws, size_t ws_bytes, int precision,
                            bd_stream_t stream);
size_t bptt_workspace_bytes(const bd_rssm& r, int T, long long N);
int imagine_bptt(const bd_imagine_bwd_args* a, float* d_raw, void* ws, size_t ws_bytes, int precision,
                 bd_stream_t stream);
bool transition_supported(const bd_transition_args& a, int precision);
int transition_forward(const bd_transition_args* a, void* ws, size_t ws_bytes, int precision,
                       bd_stream_t stream);
bool cem_supported(const bd_rssm& r, const bd_mlp& reward, int precision);
size_t cem_tc_workspace_bytes(const bd_rssm& r, const bd_mlp& reward, long long rows, int H);
int cem_rollout(const bd_cem_eval_args* a, void* ws, size_t ws_bytes, int precision, float* rew_out,
                bd_stream_t stream, bool weights_packed = false);
bool mlp_supported(const bd_mlp& m, int k1, int k2, int precision);
size_t mlp_pack_bytes(const bd_mlp& m);
int mlp_forward(const bd_mlp* m, const float* x1, int k1, const float* x2, int k2, int64_t rows,
                float* y, void* ws, size_t ws_bytes, int precision, bd_stream_t stream,
                void* saved = nullptr);
size_t mlp_saved_bytes(const bd_mlp& m, int64_t rows);
bool mlp_backward_supported(const bd_mlp& m, int k1, int k2, int precision);
size_t mlp_backward_workspace_bytes(const bd_mlp& m, int k1, int k2, int64_t rows);
int mlp_backward(const bd_mlp* m, const bd_mlp_bwd_args* a, void* ws, size_t ws_bytes, int precision,
                 bd_stream_t stream, const float* x1b = nullptr, const float* x2b = nullptr, int64_t split = -1,
                 int64_t seg_rows = 0, const void* x0b_img = nullptr, const void* x0s_img = nullptr);
}  // namespace tc

}  // namespace bd
