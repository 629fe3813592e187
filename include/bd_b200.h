/* bd_b200.h -- C ABI of libbd_b200.so: the B200 (sm_100a) implementation of
 * big-dreamer's RSSM latent-dynamics hot path.
 *
 * The reference (jgsimard/big-dreamer) is pure Python/PyTorch: it has no FFI of
 * its own.  Its "operator interface" for this path is the set of Python
 * callables below; each entry point replaces the arithmetic of one of them and
 * is what a ctypes binding on the reference side would bind:
 *
 *   bd_mlp_forward / bd_mlp_backward        DenseModel.forward     src/models.py:393-408
 *                                           (build_mlp              src/utils.py:368-404)
 *   bd_transition_forward / _backward       TransitionModel.forward src/models.py:190-299
 *   bd_imagine_forward / _backward          Dreamer.imagine_ahead   src/dreamer.py:178-237
 *                                           (+ get_action           src/dreamer.py:429-444,
 *                                              ActorModel.forward   src/models.py:506-517,
 *                                              SampleDist.entropy   src/models.py:725-733)
 *   bd_mlp_forward_save (+ bd_mlp_saved_bytes, bd_imagine_saved_bytes): the same forwards, also
 *                                           keeping 16-bit images for the tensor-core backward
 *                                           (what autograd's saved tensors are in the reference)
 *   bd_lambda_return_forward / _backward    lambda_return           src/dreamer.py:447-471
 *   bd_kl_loss_forward / _backward          Planet/Dreamer._kl_loss src/planet.py:288-308,
 *                                                                   src/dreamer.py:111-146
 *   bd_cem_evaluate / bd_cem_refit / bd_cem_plan
 *                                           MPCPlanner.forward      src/planner.py:28-90
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous row-major fp32 unless noted;
 *     the caller (PyTorch) owns all memory, the library never allocates or frees;
 *   - weights are the reference's own parameter tensors: Linear weight (out,in),
 *     GRUCell weight_ih / weight_hh (3*Be, Be) with gate order r,z,n;
 *   - every Gaussian draw of the reference is an explicit noise input;
 *   - calls only enqueue work on `stream` (a cudaStream_t) and never synchronise;
 *   - return value: 0 on success, otherwise a bd_status; bd_last_error() holds
 *     a message for the calling thread.  There is no CPU fallback.
 */
#ifndef BD_B200_H
#define BD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BD_ABI_VERSION 1
#define BD_MAX_LAYERS 8

typedef void* bd_stream_t; /* cudaStream_t */

enum bd_status {
  BD_OK = 0,
  BD_ERR_BAD_ARG = 1,
  BD_ERR_UNSUPPORTED = 2,
  BD_ERR_WORKSPACE = 3,
  BD_ERR_CUDA = 4
};

/* nn.<name> accepted for dense_activation_function (src/utils.py:391-394) */
enum bd_activation {
  BD_ACT_IDENTITY = 0,
  BD_ACT_ELU = 1,
  BD_ACT_RELU = 2,
  BD_ACT_TANH = 3,
  BD_ACT_SIGMOID = 4
};

/* arithmetic of the contractions */
enum bd_precision {
  BD_PREC_FP32 = 0, /* check mode: FFMA, fp32 everywhere              */
  BD_PREC_BF16 = 1, /* tcgen05 kind::f16 (bf16 operands, fp32 accum)  */
  BD_PREC_TF32 = 2, /* tcgen05 kind::tf32 (reserved)                  */
  BD_PREC_FP16 = 3  /* tcgen05 kind::f16 (fp16 operands: 10-bit mantissa, fp32 accum) */
};

typedef struct {
  const float* w; /* (out_features, in_features) */
  const float* b; /* (out_features)              */
  int in_features;
  int out_features;
} bd_linear;

/* build_mlp: [Linear, act] * (n_layers-1) + Linear (+ Identity) */
typedef struct {
  int n_layers;   /* number of Linear layers (DenseModel default: 5) */
  int activation; /* bd_activation of the hidden layers              */
  bd_linear layer[BD_MAX_LAYERS];
} bd_mlp;

/* TransitionModel parameters (src/models.py:149-167) */
typedef struct {
  int belief_size, state_size, action_size, hidden_size, embedding_size;
  int activation; /* bd_activation */
  float min_std_dev;
  bd_linear embed;  /* fc_embed_state_action.0 : (Be, S+A)       */
  const float* w_ih; /* rnn.weight_ih (3Be, Be)                  */
  const float* w_hh; /* rnn.weight_hh (3Be, Be)                  */
  const float* b_ih; /* rnn.bias_ih   (3Be)                      */
  const float* b_hh; /* rnn.bias_hh   (3Be)                      */
  bd_linear prior1, prior2; /* belief_prior.model.0 / .2         */
  bd_linear post1, post2;   /* belief_posterior.model.0 / .2 (w may be NULL when unused) */
} bd_rssm;

/* gradients of the TransitionModel parameters; every pointer is optional (NULL =
 * not needed) and is ACCUMULATED into (+=): the caller zero-fills. */
typedef struct {
  float *embed_w, *embed_b, *w_ih, *w_hh, *b_ih, *b_hh;
  float *prior1_w, *prior1_b, *prior2_w, *prior2_b;
  float *post1_w, *post1_b, *post2_w, *post2_b;
} bd_rssm_grads;

/* ActorModel squashing constants (src/models.py:499-516) */
typedef struct {
  float mean_scale;   /* _mean_scale  = 5                       */
  float raw_init_std; /* log(exp(init_std) - 1), init_std = 5   */
  float min_std;      /* _min_std     = 1e-4                    */
  int entropy_samples; /* SampleDist(samples=100)               */
} bd_actor_cfg;

int bd_version(void);
const char* bd_last_error(void);
/* number of CUDA kernels this library has launched in this process (monotonic) */
unsigned long long bd_launch_count(void);
/* 1 if the named precision is implemented by this build for these sizes */
int bd_precision_supported(int precision);

/* ------------------------------------------------------------------ MLP ---- */
/* y (rows, out) = MLP([x1 (rows,k1) ; x2 (rows,k2)]); x2 may be NULL with k2=0.
 * DenseModel.forward(belief, state) / DenseModel.forward(x). */
size_t bd_mlp_workspace_bytes(const bd_mlp* m, int64_t rows, int backward);
int bd_mlp_forward(const bd_mlp* m, const float* x1, int k1, const float* x2, int k2,
                   int64_t rows, float* y, void* ws, size_t ws_bytes, int precision,
                   bd_stream_t stream);

typedef struct {
  const float* x1; int k1;
  const float* x2; int k2;
  int64_t rows;
  const float* dy;            /* (rows, out)                                   */
  float* dx1;                 /* (rows,k1) optional, overwritten               */
  float* dx2;                 /* (rows,k2) optional, overwritten               */
  float* dw[BD_MAX_LAYERS];   /* optional, accumulated (+=)                    */
  float* db[BD_MAX_LAYERS];   /* optional, accumulated (+=)                    */
  const void* saved;          /* optional: buffer written by bd_mlp_forward_save (tensor-core modes) */
} bd_mlp_bwd_args;
/* Tensor-core modes: the forward can leave 16-bit images of its hidden activations in a caller
 * buffer of bd_mlp_saved_bytes() bytes; the backward then skips their recomputation.  Without it
 * (saved == NULL, or fp32 mode) hidden activations are recomputed from the inputs. */
size_t bd_mlp_saved_bytes(const bd_mlp* m, int k1, int k2, int64_t rows, int precision);
int bd_mlp_forward_save(const bd_mlp* m, const float* x1, int k1, const float* x2, int k2,
                        int64_t rows, float* y, void* saved, void* ws, size_t ws_bytes,
                        int precision, bd_stream_t stream);
int bd_mlp_backward(const bd_mlp* m, const bd_mlp_bwd_args* a, void* ws, size_t ws_bytes,
                    int precision, bd_stream_t stream);
/* Two scalar heads on the same rows in one call: reward_model(beliefs, states) and value_model(beliefs, states)
 * of Dreamer's behaviour step (src/dreamer.py:321-322; DenseModel.forward, src/models.py:393-408).  Tensor-core
 * modes only (bd_heads_forward_supported: same depth / activation / hidden width, one output each): one launch,
 * one tile prologue, the two chains interleaved so the MMAs of one head run under the epilogue of the other.
 * y_reward, y_value: (rows); saved_reward / saved_value: optional buffers of bd_mlp_saved_bytes() each, in the
 * layout bd_mlp_backward reads -- the backward of each head is a plain bd_mlp_backward call.
 * ws: bd_heads_forward_workspace_bytes(). */
int bd_heads_forward_supported(const bd_mlp* reward, const bd_mlp* value, int k1, int k2, int precision);
size_t bd_heads_forward_workspace_bytes(const bd_mlp* reward, const bd_mlp* value);
int bd_heads_forward(const bd_mlp* reward, const bd_mlp* value, const float* x1, int k1, const float* x2,
                     int k2, int64_t rows, float* y_reward, float* y_value, void* saved_reward,
                     void* saved_value, void* ws, size_t ws_bytes, int precision, bd_stream_t stream);
/* Input gradients of that pair, heads frozen (FreezeParameters(value_model / reward_model), src/dreamer.py:320):
 *   dx1 (rows,k1), dx2 (rows,k2) = d(reward)/dx . dy_reward + d(value)/dx . dy_value   (overwritten)
 * from the buffers bd_heads_forward saved; dy_reward, dy_value: (rows).  One launch: the two dgrad chains are
 * interleaved and their last GEMMs accumulate into one dX accumulator.  Weight gradients: bd_mlp_backward per head.
 * ws: bd_heads_backward_workspace_bytes(). */
size_t bd_heads_backward_workspace_bytes(const bd_mlp* reward, const bd_mlp* value, int k1, int k2);
int bd_heads_backward(const bd_mlp* reward, const bd_mlp* value, int k1, int k2, int64_t rows,
                      const float* dy_reward, const float* dy_value, const void* saved_reward,
                      const void* saved_value, float* dx1, float* dx2, void* ws, size_t ws_bytes,
                      int precision, bd_stream_t stream);

/* -------------------------------------------------------- lambda_return ---- */
/* reward, value, returns: (T, N); bootstrap: (N).  src/dreamer.py:447-471.
 * discount / lambda_ are the Python doubles the reference passes; they are rounded to
 * fp32 exactly where torch rounds them, so the forward is bit-exact. */
int bd_lambda_return_forward(const float* reward, const float* value, const float* bootstrap,
                             int T, int64_t N, double discount, double lambda_, float* returns,
                             bd_stream_t stream);
/* d_reward, d_value (T,N) and d_bootstrap (N) are overwritten; any may be NULL */
int bd_lambda_return_backward(const float* d_returns, int T, int64_t N, double discount,
                              double lambda_, float* d_reward, float* d_value, float* d_bootstrap,
                              bd_stream_t stream);

/* ------------------------------------------------------------- KL loss ---- */
/* KL(posterior || prior) of the dynamics update, summed over the S latent dimensions, with the
 * free-nats floor: Planet._kl_loss (src/planet.py:288-308) and Dreamer._kl_loss
 * (src/dreamer.py:111-146).  All four parameter tensors are (rows, S) with rows = L*B.
 *   balance < 0 : loss = mean_rows max(sum_S KL, free_nats)                (kl_balance == -1)
 *   balance >= 0: loss = balance * max(mean KL(sg(post) || prior), free_nats)
 *                      + (1 - balance) * max(mean KL(post || sg(prior)), free_nats)
 * free_nats: device pointer to one float (the reference keeps it as a (1,) tensor).
 * div (rows) and loss (2 floats: the loss, and the mean element KL of the balanced form) are
 * outputs the backward needs again.  g_loss: device pointer to dL/dloss (one float).
 * Any of the four gradient outputs may be NULL. */
int bd_kl_loss_forward(const float* post_mean, const float* post_std, const float* prior_mean,
                       const float* prior_std, int64_t rows, int S, const float* free_nats,
                       double balance, float* div, float* loss, bd_stream_t stream);
int bd_kl_loss_backward(const float* post_mean, const float* post_std, const float* prior_mean,
                        const float* prior_std, int64_t rows, int S, const float* free_nats,
                        double balance, const float* div, const float* loss, const float* g_loss,
                        float* d_post_mean, float* d_post_std, float* d_prior_mean,
                        float* d_prior_std, bd_stream_t stream);

/* ------------------------------------- critic regression loss (value update) ---- */
/* Dreamer.train_step's value loss, src/dreamer.py:380-385:
 *   loss = -mean( weight * Normal(value, 1).log_prob(target) ),  weight = cumulated discount
 *   (use_discount) or NULL for 1.  value / target / weight / d_value hold n elements;
 *   loss is one float; d_value (optional) receives d loss / d value.  ws: >= 4 KB.
 * bd.value_update (the critic forward bd_mlp_forward_save, this loss and bd_mlp_backward with
 * weight gradients) is the whole update block behind one host call. */
int bd_value_loss(const float* value, const float* target, const float* weight, int64_t n,
                  float* loss, float* d_value, void* ws, size_t ws_bytes, bd_stream_t stream);

/* ------------------------------------------------- acting: Dreamer.get_action ---- */
/* The action of Dreamer.get_action (src/dreamer.py:429-444) as Planet.update_belief_and_act uses it
 * (src/planet.py:388-390: the entropy is discarded): raw (rows, 2A) = the actor MLP's output
 * (bd_mlp_forward on ActorModel.model), squashed as ActorModel.forward does (src/models.py:513-516).
 *   deterministic == 0: action = tanh(mean + eps * std), eps (rows, A)                 (dist.rsample)
 *   deterministic == 1: SampleDist.mode (src/models.py:707-723): eps (cfg->entropy_samples, rows, A);
 *     the sample with the largest tanh-Normal log-probability is the action.
 * action (rows, A). */
int bd_actor_act(const float* raw, const float* eps, const bd_actor_cfg* cfg, int64_t rows, int action_size,
                 int deterministic, float* action, bd_stream_t stream);

/* ----------------------------------------------- TransitionModel.forward ---- */
typedef struct {
  bd_rssm rssm;
  int L;      /* number of transitions = actions.size(0)       */
  int64_t B;  /* rows                                          */
  const float* init_state;   /* (B,S)                          */
  const float* init_belief;  /* (B,Be)                         */
  const float* actions;      /* (L,B,A)                        */
  const float* embeddings;   /* (L,B,E) or NULL: prior-only    */
  const float* nonterminals; /* (L,B,1) or NULL                */
  const float* eps_prior;    /* (L,B,S)                        */
  const float* eps_post;     /* (L,B,S) (observe mode)         */
  /* outputs, each (L,B,.) */
  float *beliefs, *prior_states, *prior_means, *prior_stds;
  float *post_states, *post_means, *post_stds; /* observe mode */
} bd_transition_args;

size_t bd_transition_workspace_bytes(const bd_rssm* r, int L, int64_t B, int observe, int backward);
int bd_transition_forward(const bd_transition_args* a, void* ws, size_t ws_bytes, int precision,
                          bd_stream_t stream);

typedef struct {
  bd_transition_args fwd;    /* same inputs and the forward's outputs              */
  /* upstream gradients, each (L,B,.) or NULL */
  const float *g_beliefs, *g_prior_states, *g_prior_means, *g_prior_stds;
  const float *g_post_states, *g_post_means, *g_post_stds;
  /* input gradients (optional, overwritten) */
  float *d_init_state, *d_init_belief, *d_actions, *d_embeddings;
  bd_rssm_grads grads;       /* parameter gradients (optional, accumulated)        */
} bd_transition_bwd_args;
int bd_transition_backward(const bd_transition_bwd_args* a, void* ws, size_t ws_bytes,
                           int precision, bd_stream_t stream);

/* ------------------------------------------------- Dreamer.imagine_ahead ---- */
typedef struct {
  bd_rssm rssm;
  bd_mlp actor;            /* ActorModel.model: (Be+S) -> Hi x n -> 2A */
  bd_actor_cfg actor_cfg;
  int T;                   /* planning_horizon - 1 transitions */
  int64_t N;               /* start states (rows)              */
  const float* prev_state;  /* (N,S)                           */
  const float* prev_belief; /* (N,Be)                          */
  const float* eps_a;       /* (T,N,A)   action rsample noise  */
  const float* eps_e;       /* (T,J,N,A) entropy noise, J = entropy_samples */
  const float* eps_s;       /* (T,N,S)   prior state noise     */
  /* reference outputs */
  float *beliefs, *states, *means, *stds; /* (T,N,.)           */
  float* entropy;                          /* (T,N)            */
  /* saved for backward (caller-allocated) */
  float* actions;    /* (T,N,A)                                 */
  float* actor_raw;  /* (T,N,2A) raw actor outputs              */
  float* dent;       /* (T,N,2A) d entropy / d (mean, std)      */
  /* tensor-core modes only: opaque buffer of bd_imagine_saved_bytes() bytes in which the forward
   * leaves 16-bit images of the GRU gate coefficients and activation derivatives for the backward
   * (NULL: not saved; the backward then runs the fp32 recompute kernels) */
  void* tc_saved;
} bd_imagine_args;

size_t bd_imagine_saved_bytes(const bd_rssm* r, int T, int64_t N, int precision);

size_t bd_imagine_workspace_bytes(const bd_rssm* r, const bd_mlp* actor, int T, int64_t N,
                                  int backward);
int bd_imagine_forward(const bd_imagine_args* a, void* ws, size_t ws_bytes, int precision,
                       bd_stream_t stream);

typedef struct {
  bd_imagine_args fwd;     /* inputs + everything the forward wrote               */
  /* upstream gradients (T,N,.) / (T,N); any may be NULL */
  const float *g_beliefs, *g_states, *g_means, *g_stds, *g_entropy;
  /* optional input gradients (overwritten) */
  float *d_prev_state, *d_prev_belief;
  /* actor parameter gradients (optional, accumulated) */
  float* actor_dw[BD_MAX_LAYERS];
  float* actor_db[BD_MAX_LAYERS];
} bd_imagine_bwd_args;
int bd_imagine_backward(const bd_imagine_bwd_args* a, void* ws, size_t ws_bytes, int precision,
                        bd_stream_t stream);

/* ------------------------- imagine_ahead + reward/value heads + lambda_return ---- */
/* The behaviour-learning block of Dreamer.train_step, src/dreamer.py:313-335, as ONE forward and ONE
 * backward call ("imagine_and_returns", SURVEY.md 8b level L2):
 *     beliefs, states, (means, stds), entropy = imagine_ahead(prev_state, prev_belief)
 *     reward = reward_model(beliefs, states); value = value_model(beliefs, states)
 *     returns = lambda_return(reward, value, bootstrap=value[-1], discount, lambda_)
 * In the tensor-core modes the heads ride in the per-step program of the persistent rollout kernel
 * (their epilogues feed the next layer from shared memory: (T,N,Be+S) is never re-read from HBM) and
 * the lambda-return recursion is the kernel's tail; the backward runs the lambda-return adjoint and
 * the heads' dgrad chain inside the BPTT kernel.  Head weights are constants here (the reference
 * evaluates them under FreezeParameters, src/dreamer.py:320): no head weight gradients.
 * bd_imagine_returns_supported() == 0 -> use the piecewise entry points (same results). */
typedef struct {
  bd_imagine_args img;     /* as for bd_imagine_forward (tc_saved required when a backward follows) */
  bd_mlp reward, value;    /* DenseModel heads on [belief ; state], one output each               */
  double discount, lambda_;
  float* reward_out;       /* (T,N)                                                               */
  float* value_out;        /* (T,N)                                                               */
  float* returns;          /* (T,N)                                                               */
  void* heads_saved;       /* bd_imagine_returns_saved_bytes() bytes, or NULL (no backward)       */
} bd_imagine_returns_args;
int bd_imagine_returns_supported(const bd_rssm* r, const bd_mlp* actor, const bd_mlp* reward,
                                 const bd_mlp* value, int precision);
size_t bd_imagine_returns_workspace_bytes(const bd_imagine_returns_args* a, int backward);
size_t bd_imagine_returns_saved_bytes(const bd_imagine_returns_args* a);
int bd_imagine_returns_forward(const bd_imagine_returns_args* a, void* ws, size_t ws_bytes, int precision,
                               bd_stream_t stream);
typedef struct {
  bd_imagine_returns_args fwd;   /* the forward call's arguments (outputs filled, saved buffers intact) */
  /* upstream gradients, each optional: (T,N,Be) (T,N,S) (T,N,S) (T,N,S) (T,N) | (T,N) (T,N) (T,N) */
  const float *g_beliefs, *g_states, *g_means, *g_stds, *g_entropy;
  const float *g_reward, *g_value, *g_returns;
  float* d_prev_state;           /* (N,S)  optional */
  float* d_prev_belief;          /* (N,Be) optional */
  float* actor_dw[BD_MAX_LAYERS];   /* optional, accumulated (+=) */
  float* actor_db[BD_MAX_LAYERS];
} bd_imagine_returns_bwd_args;
int bd_imagine_returns_backward(const bd_imagine_returns_bwd_args* a, void* ws, size_t ws_bytes,
                                int precision, bd_stream_t stream);

/* ------------------------------------------------------------- CEM planner ---- */
typedef struct {
  bd_rssm rssm;
  bd_mlp reward;
  int B;       /* batch rows                                          */
  int C;       /* candidates (global)                                 */
  int H;       /* planning_horizon transitions                        */
  int c_begin; /* this rank evaluates candidates [c_begin, c_end)     */
  int c_end;
  const float* belief;      /* (B,Be)                                 */
  const float* state;       /* (B,S)                                  */
  const float* action_mean; /* (H,B,A)                                */
  const float* action_std;  /* (H,B,A)                                */
  const float* eps_act;     /* (H,B,C,A) this iteration's noise       */
  const float* eps_s;       /* (H,B*C,S) this iteration's noise       */
  float* actions;           /* (H,B,Cl,A) out: sampled local actions, Cl=c_end-c_begin */
  float* returns;           /* (B,Cl)     out: sum of rewards         */
} bd_cem_eval_args;
size_t bd_cem_workspace_bytes(const bd_rssm* r, const bd_mlp* reward, int B, int C_local, int H);
/* one CEM iteration, candidate evaluation: src/planner.py:53-72 */
int bd_cem_evaluate(const bd_cem_eval_args* a, void* ws, size_t ws_bytes, int precision,
                    bd_stream_t stream);

/* elite selection + refit: src/planner.py:74-87.  returns (B,C), actions (H,B,C,A).
 * topk_idx (B,K) int64, candidate indices in [0,C), ascending.  mean/std (H,B,A). */
int bd_cem_refit(const float* returns, const float* actions, int B, int C, int K, int H, int A,
                 int64_t* topk_idx, float* action_mean, float* action_std, bd_stream_t stream);

typedef struct {
  bd_rssm rssm;
  bd_mlp reward;
  int B, C, K, H, iters;
  const float* belief;  /* (B,Be)                  */
  const float* state;   /* (B,S)                   */
  const float* eps_act; /* (iters,H,B,C,A)         */
  const float* eps_s;   /* (iters,H,B*C,S)         */
  float* action_out;    /* (B,A) first action mean */
  /* optional per-iteration trace */
  float* returns_trace;  /* (iters,B,C) or NULL    */
  int64_t* topk_trace;   /* (iters,B,K) or NULL    */
} bd_cem_plan_args;
size_t bd_cem_plan_workspace_bytes(const bd_rssm* r, const bd_mlp* reward, int B, int C, int K,
                                   int H);
/* whole MPCPlanner.forward on one GPU */
int bd_cem_plan(const bd_cem_plan_args* a, void* ws, size_t ws_bytes, int precision,
                bd_stream_t stream);

/* ---------------------------------------------------------------- profiling ---- */
/* Optional CUDA-event timing of the library's own main kernels, recorded on the stream each
 * kernel is launched on.  bd_prof_enable(1) starts collecting; bd_prof_read synchronises the
 * recorded events and returns the accumulated milliseconds and launch count of one kernel id,
 * then clears it.  Used by bench.py for the roofline block; off by default (no overhead). */
enum bd_prof_kernel {
  BD_PROF_ROLLOUT_FWD = 0, /* persistent tcgen05 rollout (imagine / transition / CEM)   */
  BD_PROF_MLP_FWD = 1,     /* tcgen05 MLP forward (DenseModel)                            */
  BD_PROF_BPTT = 2,        /* tcgen05 reverse-time dgrad chain                            */
  BD_PROF_MLP_BWD = 3,     /* tcgen05 MLP recompute + dgrad chain                         */
  BD_PROF_WGRAD = 4,       /* streaming tcgen05 wgrad                                     */
  BD_PROF_ENTROPY = 5,     /* Monte-Carlo policy entropy                                  */
  BD_PROF_COUNT = 6
};
void bd_prof_enable(int on);
int bd_prof_read(int kernel, float* ms_total, int* launches);


#ifdef __cplusplus
}
#endif
#endif /* BD_B200_H */
