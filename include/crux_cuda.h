/* crux_cuda.h -- C ABI of libcrux_cuda.so, the B200 (sm_100a) actor-learner hot path
 * behind the Crux.jl solver surface.
 *
 * The reference (sisl/Crux.jl @ d1b6ab5) has no FFI: its boundary is Julia dispatch on
 * the names exported in src/Crux.jl:34-63.  Every entry point below cites the reference
 * function it replaces (paths relative to the reference root).  A Julia shim binds these
 * with `ccall`; tests bind them with Python ctypes.
 *
 * Conventions
 *  - every function returns an int32 status (CRUX_OK == 0); `crux_last_error` gives text.
 *    Nothing throws or longjmps across the boundary.
 *  - handles are opaque pointers; one handle is not thread-safe, distinct handles are.
 *  - all device work is stream-ordered on the context's stream (crux_ctx_set_stream); raw
 *    pointers are caller-owned device pointers unless a parameter name ends in `_host`.
 *  - arrays are batch-major `[batch][features]`, i.e. the memory order of the reference's
 *    column-major `[features, batch]` arrays (src/devices.jl:23-34).  Rollout columns are
 *    `[T][N]` (row t*N+e): env stream e is one reference `Sampler`.
 *  - Dense parameters are flat float32, per layer `W` in Julia memory order
 *    (column-major [out,in] == row-major [in][out]) followed by `b[out]`: Flux.params order.
 *  - indices cross the ABI 0-based (the Julia shim adds 1).
 *  - `*_in` pointers that may be NULL are parity hooks: when given, the kernel consumes
 *    caller-chosen noise / permutations / sample indices instead of the device Philox RNG.
 */
#ifndef CRUX_CUDA_H
#define CRUX_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRUX_ABI_VERSION 1

enum {
  CRUX_OK = 0,
  CRUX_ERR_INVALID = 1, /* bad argument (the reference's @assert / MethodError) */
  CRUX_ERR_CUDA = 2,    /* CUDA runtime failure */
  CRUX_ERR_NAN = 3,     /* training.jl:20 `error("NaN detected!")`, sampler.jl:270 @assert !isnan(A) */
  CRUX_ERR_OOM = 4,
  CRUX_ERR_NCCL = 5,
  CRUX_ERR_STATE = 6
};

enum { CRUX_ACT_IDENTITY = 0, CRUX_ACT_TANH = 1, CRUX_ACT_RELU = 2 };
enum { CRUX_U8 = 0, CRUX_F32 = 1, CRUX_I32 = 2, CRUX_I64 = 3 };

typedef struct crux_ctx crux_ctx;
typedef struct crux_mlp crux_mlp;
typedef struct crux_gaussian crux_gaussian;
typedef struct crux_buffer crux_buffer;

/* ------------------------------------------------------------------ context / devices
 * replaces src/devices.jl:1-21 (device/gpucall/cpucall/mdcall): data stays on the device. */
/* stream arguments are cudaStream_t values.  NULL asks the context to create (and own) a non-blocking stream;
 * the legacy default stream is spelled CRUX_STREAM_LEGACY (== cudaStreamLegacy). */
#define CRUX_STREAM_LEGACY ((void *)0x1)
int32_t crux_abi_version(void);
int32_t crux_ctx_create(int32_t device, void *stream /* cudaStream_t or NULL: own stream */, crux_ctx **out);
int32_t crux_ctx_destroy(crux_ctx *ctx);
int32_t crux_ctx_set_stream(crux_ctx *ctx, void *stream);
int32_t crux_ctx_stream(crux_ctx *ctx, void **stream_out);
int32_t crux_ctx_sync(crux_ctx *ctx);
const char *crux_last_error(crux_ctx *ctx /* NULL: last error of a failed crux_ctx_create */);
/* number of kernels this context has launched since creation (bench `gpu_launches`) */
int32_t crux_ctx_launch_count(crux_ctx *ctx, int64_t *out);
/* Opt-in device timing per kernel family (bench.py roofline): between begin and end every launch of the families below is
 * bracketed by a CUDA event pair on the context stream; end synchronises and returns summed milliseconds and launch counts.
 * families: 0 fused minibatch (forward+loss+backward), 1 partial reduction, 2 norm/record/Adam, 3 fused forward,
 *           4 GAE/returns scan, 5 synthetic env step.  ms_out_host / count_out_host have 8 entries. */
int32_t crux_ctx_timing_begin(crux_ctx *ctx);
int32_t crux_ctx_timing_end(crux_ctx *ctx, float *ms_out_host, int32_t *count_out_host);
/* sticky device-side error flag (NaN in gradients / advantages); reading synchronises. */
int32_t crux_ctx_check(crux_ctx *ctx);

int32_t crux_dev_alloc(crux_ctx *ctx, size_t bytes, void **out);
int32_t crux_dev_free(crux_ctx *ctx, void *ptr);
int32_t crux_pinned_alloc(crux_ctx *ctx, size_t bytes, void **out_host);
int32_t crux_pinned_free(crux_ctx *ctx, void *ptr_host);
int32_t crux_memcpy_h2d(crux_ctx *ctx, void *dst, const void *src_host, size_t bytes); /* async */
int32_t crux_memcpy_d2h(crux_ctx *ctx, void *dst_host, const void *src, size_t bytes); /* async */
int32_t crux_memcpy_d2d(crux_ctx *ctx, void *dst, const void *src, size_t bytes);
int32_t crux_memset(crux_ctx *ctx, void *dst, int32_t byte, size_t bytes);
/* CUDA-graph capture of a fixed launch sequence on the context stream */
int32_t crux_graph_begin(crux_ctx *ctx);
int32_t crux_graph_end(crux_ctx *ctx, void **graph_exec_out);
int32_t crux_graph_launch(crux_ctx *ctx, void *graph_exec);
int32_t crux_graph_destroy(crux_ctx *ctx, void *graph_exec);

/* ------------------------------------------------------------------ networks
 * crux_mlp == a Flux `Chain(Dense...)` held by ContinuousNetwork / DiscreteNetwork
 * (src/policies.jl:68-98,104-157) plus its Adam state (src/training.jl:3, Flux Adam). */
int32_t crux_mlp_create(crux_ctx *ctx, int32_t n_layers, const int32_t *dims /* n_layers+1 */,
                        const int32_t *acts /* n_layers */, crux_mlp **out);
int32_t crux_mlp_destroy(crux_mlp *mlp);
int32_t crux_mlp_num_params(crux_mlp *mlp, int64_t *out);
int32_t crux_mlp_set_params(crux_mlp *mlp, const float *flat_host);
int32_t crux_mlp_get_params(crux_mlp *mlp, float *flat_host); /* synchronises */
int32_t crux_mlp_params_ptr(crux_mlp *mlp, float **dev_out);
int32_t crux_mlp_grads_ptr(crux_mlp *mlp, float **dev_out);   /* gradient of the last backward */
/* Adam(η, (β1,β2), ϵ) with Flux's Float64 scalars; resets the moments (training.jl:3) */
int32_t crux_mlp_set_adam(crux_mlp *mlp, double eta, double beta1, double beta2, double eps);
/* value(π, s)  policies.jl:94,120 */
int32_t crux_mlp_forward(crux_mlp *mlp, const float *x, int64_t B, float *y);
/* value(V, sp_i) for every row of a [T][N] rollout (the second value call of fill_gae!, sampler.jl:268), given v_s = value(V, s):
 * wherever sp[t][e] equals s[t+1][e] bit for bit (every transition that is not followed by a reset) the result is v_s[t+1][e] --
 * identical to evaluating the network on the same input -- and only the remaining tiles run the network. */
int32_t crux_value_next(crux_mlp *mlp, const float *sp, const float *s, const float *v_s, int64_t T, int64_t N, float *v_sp);
/* value(π, s, a) = network(vcat(s, a))  policies.jl:96 */
int32_t crux_mlp_forward_sa(crux_mlp *mlp, const float *s, int32_t sdim, const float *a, int32_t adim,
                            int64_t B, float *y);
/* copyto!(to, from) policies.jl:61-65 and polyak_average!(to, from, τ) policies.jl:48-59 */
int32_t crux_mlp_copy(crux_mlp *to, crux_mlp *from);
int32_t crux_mlp_polyak(crux_mlp *to, crux_mlp *from, float tau);
/* one `train!` (training.jl:15-25) of Flux.mse(net(x), y) -- the PPO/A2C critic loss (ppo.jl:60);
 * info_out_host[0..1] = loss, grad_norm (synchronises).  Mainly a parity hook. */
int32_t crux_mlp_train_mse(crux_mlp *mlp, const float *x, const float *y, int64_t B, float *info_out_host);

/* GaussianPolicy(μ::ContinuousNetwork, logΣ::AbstractArray) policies.jl:315-350 and
 * SquashedGaussianPolicy policies.jl:355-400.
 *   log_sigma_host != NULL : state-independent logΣ vector (ConstantLayer), trainable
 *   log_sigma_host == NULL : `mu` has 2*adim outputs = [μ | logΣ] heads on a shared trunk
 *                            (examples/rl/half_cheetah_mujoco.jl:37-43)
 * The policy borrows `mu` (it must outlive the policy). */
int32_t crux_gaussian_create(crux_ctx *ctx, crux_mlp *mu, int32_t adim, const float *log_sigma_host,
                             int32_t squashed, float ascale, crux_gaussian **out);
int32_t crux_gaussian_destroy(crux_gaussian *pol);
/* A DiscreteNetwork ACTOR (policies.jl:104-157) for the on-policy updates: ppo_loss / a2c_loss / reinforce_loss (rl/ppo.jl:4-21, a2c.jl:4-16,
 * reinforce.jl:4-13) are generic over the policy through logpdf = categorical_logpdf (:135: log(sum(softmax(net(s)) .* a_onehot))) and
 * entropy (:152-155: -sum(p .* log.(p .+ eps(Float32))), e_loss = -mean over the minibatch) -- the reference's cartpole PPO / A2C / REINFORCE
 * examples (examples/rl/cartpole.jl:8-9).  The handle is accepted by crux_ppo_update / crux_ppo_update_async only (a = one-hot rows [n][n_actions],
 * logprob as written by crux_discrete_explore); the crux_gaussian_* / rollout entry points reject it.  Destroy with crux_gaussian_destroy. */
int32_t crux_categorical_create(crux_ctx *ctx, crux_mlp *logits, int32_t n_actions, crux_gaussian **out);
int32_t crux_gaussian_log_sigma_ptr(crux_gaussian *pol, float **dev_out);
/* exploration(π, s): a = ε·σ + μ (tanh-squashed if squashed), logprob.  eps_in NULL => Philox(seed, ctr). */
int32_t crux_gaussian_explore(crux_gaussian *pol, const float *s, int64_t B, const float *eps_in,
                              uint64_t seed, uint64_t ctr, float *a_out, float *logp_out);
/* action(π, s): μ(s) (ascale·tanh μ if squashed)  policies.jl:331,372 */
int32_t crux_gaussian_action(crux_gaussian *pol, const float *s, int64_t B, float *a_out);
/* logpdf(π, s, a) policies.jl:346,396 ; entropy(π, s) :348 (scalar -> out[0]) / :398 ([B]) */
int32_t crux_gaussian_logpdf(crux_gaussian *pol, const float *s, const float *a, int64_t B, float *out);
int32_t crux_gaussian_entropy(crux_gaussian *pol, const float *s, int64_t B, float *out);

/* DiscreteNetwork policies.jl:104-157.  q is [B][nA] from crux_mlp_forward. */
/* action: argmax index (first max wins) :124 */
int32_t crux_discrete_argmax(crux_ctx *ctx, const float *q, int64_t B, int32_t nA, int32_t *a_idx);
/* exploration(::DiscreteNetwork) :137-142 : softmax -> inverse-CDF sample (u_in NULL => Philox), logprob */
int32_t crux_discrete_explore(crux_ctx *ctx, const float *q, int64_t B, int32_t nA, const double *u_in,
                              uint64_t seed, uint64_t ctr, int32_t *a_idx, float *logp);
/* logpdf :144-150 (a one-hot [B][nA] float) and entropy :152-155 */
int32_t crux_discrete_logpdf(crux_ctx *ctx, const float *q, const float *a_onehot, int64_t B, int32_t nA, float *out);
int32_t crux_discrete_entropy(crux_ctx *ctx, const float *q, int64_t B, int32_t nA, float *out);
/* The same three with logit_conversion = softmax(value(π, s) ./ α), the conversion SoftQ installs (rl/softq.jl:48);
 * α = 1 is the default conversion (policies.jl:108) and bit-identical to the entry points above. */
int32_t crux_discrete_explore_t(crux_ctx *ctx, const float *q, int64_t B, int32_t nA, float alpha, const double *u_in,
                                uint64_t seed, uint64_t ctr, int32_t *a_idx, float *logp);
int32_t crux_discrete_logpdf_t(crux_ctx *ctx, const float *q, const float *a_onehot, int64_t B, int32_t nA, float alpha,
                               float *out);
int32_t crux_discrete_entropy_t(crux_ctx *ctx, const float *q, int64_t B, int32_t nA, float alpha, float *out);
/* ϵ-greedy exploration(::MixedPolicy) :474-494 over B env streams: with prob eps a uniform action
 * (u_in[2B]: [coin, pick] pairs, NULL => Philox) else argmax; writes index, one-hot row and logprob. */
int32_t crux_discrete_eps_greedy(crux_ctx *ctx, const float *q, int64_t B, int32_t nA, double eps,
                                 const double *u_in, uint64_t seed, uint64_t ctr, int32_t *a_idx,
                                 float *a_onehot /* nullable */, float *logp /* nullable */);

/* ------------------------------------------------------------------ hot path (i): Sampler.steps!
 * One vector step of src/sampler.jl:71-137 for N env streams on a Gaussian actor-critic:
 * reads obs[N][sdim] (already tovec-normalised, spaces.jl:24-25), writes a, logprob, V(s).
 * critic may be NULL (then v_out is not written). */
int32_t crux_rollout_step(crux_gaussian *actor, crux_mlp *critic, const float *obs, int64_t N,
                          const float *eps_in, uint64_t seed, uint64_t ctr, float *a_out,
                          float *logp_out, float *v_out);
/* The same hot path for a HOST environment: the whole steps! loop (sampler.jl:139-155) of T vector steps in one call.
 * The environment (the user's POMDPs.jl model) is reached through two callbacks, which a Julia host passes as @cfunction
 * pointers:  step  = @gen(:sp,:r)(mdp, s, a) + isterminal for the streams [e0, e1) (host arrays in pinned memory; the
 *                    pointers address the full [N] arrays, rows e0..e1-1 are read / written),
 *            reset = rand(initialstate(mdp)) + convert_s for the listed streams (sampler.jl:31-43).
 * With the fused policy shapes a vector step is issued as two half ranges so that the device forward of one half overlaps
 * the host env step of the other; results do not depend on the split (noise streams are keyed by the stream id).
 * obs_pinned [N][sdim] (pinned host memory, in/out) holds the current observation of every stream (already tovec'ed);
 * episode_length [N] (host, in/out) is Sampler.episode_length.  Columns are device pointers to rows [T*N] (row t*N+e).
 * Exploration noise comes from the device Philox stream (seed, ctr0 + t). */
typedef void (*crux_env_step_fn)(void *user, int32_t e0, int32_t e1, const float *a, float *sp, float *r, uint8_t *done);
typedef void (*crux_env_reset_fn)(void *user, const int32_t *idx, int32_t n_idx, float *obs_out);
typedef struct crux_rollout_cols {
  float *s, *a, *sp, *r;
  uint8_t *done, *episode_end;
  float *logprob; /* nullable */
} crux_rollout_cols;
int32_t crux_rollout_host(crux_gaussian *actor, int64_t N, int32_t T, int32_t max_steps, int32_t reset_at_end,
                          crux_env_step_fn step, crux_env_reset_fn reset, void *user, float *obs_pinned,
                          int32_t *episode_length, const crux_rollout_cols *cols, uint64_t seed, uint64_t ctr0);
/* tovec(o, ContinuousSpace) = (o - μ)/σ  spaces.jl:24-25, utils.jl:42 (in place when out == x) */
int32_t crux_normalize_obs(crux_ctx *ctx, const float *x, int64_t n, float mu, float sigma, float *out);

/* ------------------------------------------------------------------ hot path (ii): advantages / targets
 * fill_gae! (sampler.jl:262-273) + fill_returns! (:275-281) for every episode range of every env
 * stream of a [T][N] rollout, as one segmented reverse scan.  adv or ret may be NULL. */
int32_t crux_fill_gae_returns(crux_ctx *ctx, const float *r, const uint8_t *done, const uint8_t *episode_end,
                              const float *v_s, const float *v_sp, int64_t T, int64_t N, float gamma,
                              float lambda, float *adv, float *ret);
/* whiten(v) utils.jl:41-42 in place (Bessel std, no epsilon); ppo.jl:61, a2c.jl:48.
 * With NCCL initialised the statistics are all-reduced so every rank whitens identically. */
int32_t crux_whiten(crux_ctx *ctx, float *x, int64_t n);
/* dqn_target rl/dqn.jl:4-6 : y = r + γ(1-done)·max_a Q⁻(sp) */
int32_t crux_dqn_target(crux_ctx *ctx, const float *r, const uint8_t *done, const float *q_sp, int64_t B,
                        int32_t nA, float gamma, float *y);
/* softq_target rl/softq.jl:13-17 : y = r + γ(1-done)·soft_value(sp), soft_value = α·logsumexp(Q⁻(sp)/α) (:8) */
int32_t crux_softq_target(crux_ctx *ctx, const float *r, const uint8_t *done, const float *q_sp, int64_t B,
                          int32_t nA, float gamma, float alpha, float *y);
/* sac_target rl/sac.jl:4-9 : y = r + γ(1-done)(min(Q1⁻,Q2⁻) - e^{logα}·logp) */
int32_t crux_sac_target(crux_ctx *ctx, const float *r, const uint8_t *done, const float *q1, const float *q2,
                        const float *logp, int64_t B, float gamma, const float *log_alpha_dev, float *y);
/* td_error utils.jl:112 : |Q(s,a) - y| ; Q(s,a) for a DiscreteNetwork = Σ Q·onehot (policies.jl:122) */
int32_t crux_td_error(crux_ctx *ctx, const float *q_sa, const float *y, int64_t B, float *out);
int32_t crux_discrete_q_sa(crux_ctx *ctx, const float *q, const float *a_onehot, int64_t B, int32_t nA, float *out);

/* ------------------------------------------------------------------ hot path (iii): minibatch updates */
typedef struct crux_ppo_hp {
  float eps_clip;   /* 𝒫[:ϵ]  ppo.jl:42 */
  float lambda_p;   /* 𝒫[:λp] */
  float lambda_e;   /* 𝒫[:λe] */
  float target_kl;  /* early stop on the last minibatch's KL (ppo.jl:59); +inf disables */
  int32_t a2c;      /* 1: a2c_loss (rl/a2c.jl:4-16) instead of ppo_loss */
  int32_t actor_epochs, actor_batch;   /* a_opt TrainingParams (training.jl:1-11) */
  int32_t critic_epochs, critic_batch; /* c_opt; critic_epochs == 0 skips the critic */
  int64_t actor_max_batches, critic_max_batches; /* <= 0 : Inf */
} crux_ppo_hp;

#define CRUX_PPO_INFO_STRIDE 8
/* per-minibatch info record (float32): written for every minibatch actually trained */
enum { CRUX_PPO_LOSS = 0, CRUX_PPO_GRAD_NORM = 1, CRUX_PPO_ENTROPY = 2, CRUX_PPO_KL = 3,
       CRUX_PPO_CLIP_FRAC = 4, CRUX_PPO_AVG_ADV = 5, CRUX_PPO_AVG_RET = 6, CRUX_PPO_VALID = 7 };

/* policy_gradient_training (src/model_free/on_policy.jl:56-78) = batch_train!(actor) then
 * batch_train!(critic) (src/training.jl:28-55) over the n rows of a rollout buffer.
 *   order_actor / order_critic : [epochs][n] int32 row orders (the buffer order after that epoch's
 *     shuffle!, experience_buffer.jl:118-124); NULL => device-generated random permutations.
 *   info_actor_host  [actor_epochs * ceil(n/actor_batch)][CRUX_PPO_INFO_STRIDE]
 *   info_critic_host [critic_epochs * ceil(n/critic_batch)][CRUX_PPO_INFO_STRIDE] (loss, grad_norm, valid)
 * Early stopping is evaluated on the device (no host sync between minibatches); minibatches after
 * the stop have VALID == 0.  Returns CRUX_ERR_NAN if a gradient norm was NaN (training.jl:20). */
int32_t crux_ppo_update(crux_gaussian *actor, crux_mlp *critic, const float *s, const float *a,
                        const float *logprob, const float *advantage, const float *ret, int64_t n,
                        const crux_ppo_hp *hp, const int32_t *order_actor, const int32_t *order_critic,
                        uint64_t seed, float *info_actor_host, float *info_critic_host);
/* same, but asynchronous: info stays in device memory (crux_ppo_info_ptrs) until the caller reads it */
int32_t crux_ppo_update_async(crux_gaussian *actor, crux_mlp *critic, const float *s, const float *a,
                              const float *logprob, const float *advantage, const float *ret, int64_t n,
                              const crux_ppo_hp *hp, const int32_t *order_actor, const int32_t *order_critic,
                              uint64_t seed);
int32_t crux_ppo_info_ptrs(crux_gaussian *actor, float **info_actor_dev, float **info_critic_dev);

/* LagrangePPO (rl/ppo.jl:133-214): policy_gradient_training (on_policy.jl:56-78) with lagrange_ppo_loss (ppo.jl:70-131) for the
 * actor, mse(V(s), return) for the critic and mse(Vc(s), cost_return) for the cost critic (:207-208), in that order.
 * The loss evaluates a PID controller on the minibatch's average episode cost sum(cost)/sum(episode_end) (:79-106) ONCE per
 * minibatch and scales the clipped cost-advantage surrogate with the resulting penalty:
 *   loss = (λp·p_loss + λe·e_loss + penalty·mean(max(r·Ac, clamp(r, 1-ϵ, 1+ϵ)·Ac))) / (1 + penalty).
 * pid_state_dev: device float[5] = {I, smooth_Δ, smooth_Jc, Jc_prev, penalty}, zero-initialised by the caller, persists
 * across updates (the reference keeps them in 𝒫).  Runs on the layer-by-layer engine (the fused kernels cover ppo/a2c/mse).
 * info_lagrange_host [actor minibatches][8] = penalty, cur_cost, prop_term, deriv_term, integral term, λp·p_loss, cost_loss, valid;
 * info_cost_host [cost minibatches][8] like info_critic_host.  Several ranks: sum(cost) and sum(episode_end) of the minibatch are summed
 * over ranks before the PID step (the cost estimate of the UNION minibatch, identical state on every rank), gradients as in crux_ppo_update. */
typedef struct crux_lagrange_hp {
  float target_cost;  /* 𝒫[:target_cost] ppo.jl:146 */
  float penalty_max;  /* Inf32 */
  float Ki_max, Ki, Kp, Kd;
  double ema_alpha;   /* Float64 0.95 in the reference */
  int32_t cost_epochs;
  int64_t cost_batch;
  int64_t cost_max_batches; /* 0 = Inf */
} crux_lagrange_hp;
int32_t crux_lagrange_ppo_update(crux_gaussian *actor, crux_mlp *critic, crux_mlp *cost_critic, const float *s,
                                 const float *a, const float *logprob, const float *advantage, const float *ret,
                                 const float *cost, const float *cost_advantage, const float *cost_return,
                                 const uint8_t *episode_end, int64_t n, const crux_ppo_hp *hp,
                                 const crux_lagrange_hp *lhp, float *pid_state_dev, const int32_t *order_actor,
                                 const int32_t *order_critic, const int32_t *order_cost, uint64_t seed,
                                 float *info_actor_host, float *info_critic_host, float *info_lagrange_host,
                                 float *info_cost_host);

/* One DQN critic `train!` (off_policy.jl:91-93 with td_loss utils.jl:76-87 on a DiscreteNetwork):
 * loss = agg((Σ Q(s)·a_onehot - y)²), agg = mean or weighted_mean(weight) (utils.jl:47).
 * info_out_host[0..2] = loss, grad_norm, Qavg (synchronises; NULL skips the readback). */
int32_t crux_dqn_train(crux_mlp *q, const float *s, const float *a_onehot, const float *y,
                       const float *weight /* nullable */, int64_t B, float *info_out_host);

/* Pixel-DQN network (examples/rl/atari.jl:8):
 *   Chain(x -> x ./ 255f0, Conv((k1,k1), C => c1, relu, stride = s1), Conv((k2,k2), c1 => c2, relu, stride = s2), flatten,
 *         Dense(F, hidden, relu), Dense(hidden, nA))                F = c2 * OH2 * OW2
 * Observations are [B][C][H][W] with W fastest (the memory of a Flux WHCN array), uint8 (s_is_u8 = 1: the replay buffer keeps pixels as
 * bytes, the conversion and the 1/255 scale are fused into the first layer's operand fetch) or float32.  Flux's Conv is a true convolution
 * (flipped kernel), no padding, dilation 1.  Flat parameters in Flux.params order and memory: W1 [kw,kh,ci,co] | b1 | W2 | b2 | Dense
 * W [out,in] | b | ...  scale255 = 0 drops the leading scale layer.  Channels per conv layer <= 32. */
typedef struct crux_convq crux_convq;
int32_t crux_convq_create(crux_ctx *ctx, int32_t C, int32_t H, int32_t W, int32_t scale255, int32_t k1, int32_t s1, int32_t c1,
                          int32_t k2, int32_t s2, int32_t c2, int32_t hidden, int32_t nA, crux_convq **out);
int32_t crux_convq_destroy(crux_convq *net);
int32_t crux_convq_num_params(crux_convq *net, int64_t *out);
int32_t crux_convq_shape(crux_convq *net, int32_t *flatten_out, int32_t *oh1, int32_t *ow1, int32_t *oh2, int32_t *ow2);
int32_t crux_convq_set_params(crux_convq *net, const float *flat_host);
int32_t crux_convq_get_params(crux_convq *net, float *flat_host); /* synchronises */
int32_t crux_convq_grads(crux_convq *net, float *flat_host);      /* gradient of the last train step (parity hook; synchronises) */
/* Flux.Optimiser(ClipValue(clip_value), Adam(eta, (beta1, beta2), eps)) (atari.jl:10); clip_value = 0: plain Adam.  Resets the moments. */
int32_t crux_convq_set_adam(crux_convq *net, double eta, double beta1, double beta2, double eps, float clip_value);
/* value(π, s): Q values [B][nA] */
int32_t crux_convq_forward(crux_convq *net, const void *s, int32_t s_is_u8, int64_t B, float *q_out);
int32_t crux_convq_copy(crux_convq *to, crux_convq *from);
int32_t crux_convq_polyak(crux_convq *to, crux_convq *from, float tau);
/* One DQN critic `train!` like crux_dqn_train on the pixel network: forward, td_loss head, backward through the head and both
 * convolutions, ||g||, [ClipValue] + Adam.  info_out_host[0..2] = loss, grad_norm, Qavg (NULL skips the readback).  Single rank. */
int32_t crux_convq_dqn_train(crux_convq *net, const void *s, int32_t s_is_u8, const float *a_onehot, const float *y,
                             const float *weight /* nullable */, int64_t B, float *info_out_host);

/* One SAC value_training epoch (off_policy.jl:66-111 with rl/sac.jl:4-52) on a sampled minibatch:
 * target -> temperature step -> double-Q critic step -> actor step -> polyak of the targets.
 *   eps_target / eps_temp / eps_actor : [B][adim] noise for the three exploration() draws (NULL => Philox)
 *   log_alpha_dev : 1 float, trained by Adam(alpha_eta) with moments alpha_state_dev[2] + betas on host side
 * info_out_host[0..7] = temp_loss, critic_loss, critic_grad_norm, actor_loss, actor_grad_norm, entropy, Q1avg, Q2avg
 * Several ranks (crux_nccl_init): every rank passes its own minibatch of B rows (per-rank replay shards); the losses are
 * means over world * B rows, the gradients of every optimiser and the temperature mean are summed over ranks before the step, so
 * parameters, targets and log α stay bit-identical on every rank; device noise streams differ per rank; info values are rank-local. */
typedef struct crux_sac_state crux_sac_state;
int32_t crux_sac_create(crux_gaussian *actor, crux_mlp *q1, crux_mlp *q2, crux_mlp *q1_target, crux_mlp *q2_target,
                        float log_alpha, float h_target, double alpha_eta, float tau, crux_sac_state **out);
int32_t crux_sac_destroy(crux_sac_state *st);
int32_t crux_sac_log_alpha(crux_sac_state *st, float *out_host);
int32_t crux_sac_train(crux_sac_state *st, const float *s, const float *a, const float *sp, const float *r,
                       const uint8_t *done, int64_t B, float gamma, const float *eps_target,
                       const float *eps_temp, const float *eps_actor, uint64_t seed, uint64_t ctr,
                       float *y_out /* nullable [B] */, float *info_out_host /* nullable */);

/* exploration(::GaussianNoiseExplorationPolicy) policies.jl:499-514 applied in place to a = action(π_on, s) [B][A]:
 * a = clamp(a + clamp(randn·σ(i), ϵ_min, ϵ_max), a_min, a_max).  a_min / a_max: host vectors of n_min / n_max entries
 * (0 => unbounded, 1 => broadcast, A => per dimension).  eps_in [B][A] nullable => Philox. */
int32_t crux_noise_explore(crux_ctx *ctx, float *a, int64_t B, int32_t A, float sigma, float eps_min, float eps_max,
                           const float *a_min, int32_t n_min, const float *a_max, int32_t n_max,
                           const float *eps_in, uint64_t seed, uint64_t ctr);

/* One DDPG / TD3 value_training epoch (off_policy.jl:66-111) on a sampled minibatch:
 *   target  ddpg_target rl/ddpg.jl:6-8, smoothed_ddpg_target :14-17 (smooth = 1), td3_target rl/td3.jl:4-7 (two critics):
 *           a' = action(π⁻, sp) [clamp(a' + clamp(σ·ε, ϵ_min, ϵ_max), a_min, a_max)], y = r + γ(1-done)·min_k Q_k⁻(sp, a')
 *   critic  td_loss (utils.jl:76-87) or double_Q_loss (:89-96), skipped when train_critic == 0 (c_opt.update_every)
 *   actor   ddpg_actor_loss rl/ddpg.jl:25 / td3_actor_loss rl/td3.jl:12: -mean(Q1(s, μ(s))), then polyak τ of the whole π⁻
 *           (actor and critics, off_policy.jl:55,100); both skipped when train_actor == 0 (a_opt.update_every)
 * q2 / q2_target NULL => DDPG.  info_out_host[0..7] = -, critic_loss, critic_grad_norm, actor_loss, actor_grad_norm, -, Q1avg, Q2avg
 * Several ranks: like crux_sac_train (means over world * B rows, gradients summed over ranks before each optimiser step). */
typedef struct crux_ddpg_state crux_ddpg_state;
int32_t crux_ddpg_create(crux_mlp *actor, crux_mlp *actor_target, crux_mlp *q1, crux_mlp *q1_target,
                         crux_mlp *q2 /* nullable */, crux_mlp *q2_target /* nullable */, float tau, crux_ddpg_state **out);
int32_t crux_ddpg_destroy(crux_ddpg_state *st);
int32_t crux_ddpg_train(crux_ddpg_state *st, const float *s, const float *a, const float *sp, const float *r,
                        const uint8_t *done, int64_t B, float gamma, int32_t smooth, float sigma, float eps_min,
                        float eps_max, const float *a_min, int32_t n_min, const float *a_max, int32_t n_max,
                        const float *eps_smooth /* nullable [B][A] */, uint64_t seed, uint64_t ctr,
                        int32_t train_critic, int32_t train_actor, float *y_out /* nullable [B] */,
                        float *info_out_host /* nullable */);

/* ------------------------------------------------------------------ ExperienceBuffer (src/experience_buffer.jl)
 * Device-resident structure-of-arrays ring buffer.  Column ids are caller-chosen small integers
 * (the shim maps Symbols :s,:a,:sp,:r,:done,:episode_end,... to ids). */
typedef struct crux_col_desc {
  int32_t id;      /* 0..63 */
  int32_t dtype;   /* CRUX_U8 / CRUX_F32 / CRUX_I32 / CRUX_I64 */
  int64_t rowlen;  /* elements per row (prod(dim(space))) */
  double init;     /* fill value (mdp_data: zeros / ones for weights, experience_buffer.jl:4-35) */
} crux_col_desc;

int32_t crux_buffer_create(crux_ctx *ctx, int64_t capacity, int32_t n_cols, const crux_col_desc *cols,
                           int32_t prioritized, float alpha, crux_buffer **out);
int32_t crux_buffer_destroy(crux_buffer *buf);
/* b.data[k] : zero-copy device pointer of a whole column (b[:key] is its first `elements` rows, :173) */
int32_t crux_buffer_col(crux_buffer *buf, int32_t col_id, void **dev_ptr, int64_t *rowlen, int32_t *dtype);
/* elements, next_ind (0-based), total_count, capacity  (experience_buffer.jl:53-60,183-185) */
int32_t crux_buffer_state(crux_buffer *buf, int64_t *elements, int64_t *next_ind, int64_t *total_count,
                          int64_t *capacity);
int32_t crux_buffer_clear(crux_buffer *buf); /* clear! :97-104 */
/* push!(b, data; ids) :232-259.  col_ptrs[i] is the source column for col_ids[i] (host memory if
 * src_on_host, else device); ids (nullable, device int32 if !src_on_host else host) gathers source rows.
 * New rows of a prioritized buffer get max_priority.  first_index_out: 0-based ring position of row 0. */
int32_t crux_buffer_push(crux_buffer *buf, int64_t n_rows, int32_t n_src_cols, const int32_t *col_ids,
                         const void *const *col_ptrs, int32_t src_on_host, const int32_t *ids,
                         int64_t *first_index_out);
/* push!(target, source, ids = ids) for two device buffers (uniform_sample! :317-321 after the draw) */
int32_t crux_buffer_push_from(crux_buffer *target, crux_buffer *source, int64_t n_rows,
                              const int32_t *ids_dev /* nullable: rows 0..n-1 */);
/* get_last_N_indices :223-229 -> host int64[min(N,len)] 0-based; returns count in n_out */
int32_t crux_buffer_last_n_indices(crux_buffer *buf, int64_t N, int64_t *out_host, int64_t *n_out);
/* uniform_sample! :317-321 : ids_in_host (0-based, nullable => Philox) -> target.indices, rows gathered */
int32_t crux_buffer_sample_uniform(crux_buffer *target, crux_buffer *source, int64_t B, const int32_t *ids_in_host,
                                   uint64_t seed, uint64_t ctr);
/* prioritized_sample! :324-349 : device prefix-sum of priorities (cached until priorities change),
 * stratified searchsortedfirst with Float64 thresholds (u_in_host[B] nullable => Philox), IS weights written
 * into the source's weight column (col id weight_col) and rows gathered into target. */
int32_t crux_buffer_sample_prioritized(crux_buffer *target, crux_buffer *source, int64_t B, float beta,
                                       int32_t weight_col, const double *u_in_host, uint64_t seed, uint64_t ctr);
/* target.indices of the last sample (device int32[B]) */
int32_t crux_buffer_indices(crux_buffer *buf, int32_t **idx_dev, int64_t *n);
/* update_priorities!(b, I, v) :290-301 : I device int32 (0-based), v device float32 */
int32_t crux_buffer_update_priorities(crux_buffer *buf, const int32_t *idx_dev, const float *v_dev, int64_t n);
/* priority state: device pointers to priorities[capacity] and the cached prefix sum; host copies of
 * max_priority / min_priority (synchronises) */
int32_t crux_buffer_priorities(crux_buffer *buf, float **prs_dev, float **cumsum_dev, float *max_p, float *min_p);
/* split_batches(N, fracs) :126-131 (host integer arithmetic) */
int32_t crux_split_batches(int64_t N, const double *fracs, int32_t n_fracs, int64_t *out);
/* generic row gather: dst[i] = src[idx[i]] for rows of `rowbytes` bytes (minibatch(), :170) */
int32_t crux_gather_rows(crux_ctx *ctx, void *dst, const void *src, const int32_t *idx_dev, int64_t n,
                         int64_t rowbytes);

/* ------------------------------------------------------------------ synthetic device-side MDP (bench `value` leg)
 * "LinQuad" MDP of SURVEY 8d stepped on the device: s' = clip(A s + B tanh(a) + 0.01 ξ, -10, 10), ...
 * A_host [sdim][sdim], B_host [sdim][adim] row-major.  Handles reset (done or t >= max_steps):
 * writes sp (pre-reset), r, done, episode_end and the next obs (post-reset) in one kernel. */
typedef struct crux_linquad crux_linquad;
int32_t crux_linquad_create(crux_ctx *ctx, int32_t sdim, int32_t adim, const float *A_host, const float *B_host,
                            int64_t n_env, int32_t max_steps, uint64_t seed, crux_linquad **out);
int32_t crux_linquad_destroy(crux_linquad *env);
int32_t crux_linquad_reset(crux_linquad *env, float *obs_out);
int32_t crux_linquad_step(crux_linquad *env, const float *obs, const float *a, float *sp, float *r, uint8_t *done,
                          uint8_t *episode_end, float *next_obs, int32_t force_end /* last step of a rollout */);

/* The whole steps! loop (sampler.jl:139-155) of T vector steps for the device env in ONE persistent launch: policy forward,
 * Gaussian sample + logprob, transition, episode bookkeeping and reset, with the observation tile resident in shared memory.
 * Bit-identical to T x (crux_rollout_step + crux_linquad_step).  obs_io [N][sdim] (device) is the current observation of every
 * stream (in/out); columns are device rows [T*N]; noise = device Philox (seed, ctr0 + t). */
int32_t crux_linquad_rollout(crux_linquad *env, crux_gaussian *actor, int32_t T, int32_t force_end_last, float *obs_io,
                             const crux_rollout_cols *cols, uint64_t seed, uint64_t ctr0);

/* ------------------------------------------------------------------ multi-GPU (one rank per GPU)
 * NCCL is resolved at run time (dlopen libnccl.so.2).  After crux_nccl_init every gradient produced by
 * crux_*_update / crux_*_train is summed over ranks before Adam and minibatch means use the global count. */
int32_t crux_nccl_unique_id(uint8_t *id_out_host /* 128 bytes */);
int32_t crux_nccl_init(crux_ctx *ctx, int32_t rank, int32_t world, const uint8_t *id_host);
int32_t crux_nccl_allreduce_f32(crux_ctx *ctx, float *buf, int64_t n); /* in-place sum */
/* one-shot peer all-reduce over NVLink (CUDA IPC): exchange handles through the host. */
int32_t crux_peer_handle(crux_ctx *ctx, uint8_t *handle_out_host /* 64 bytes */, int64_t max_floats);
int32_t crux_peer_init(crux_ctx *ctx, int32_t rank, int32_t world, const uint8_t *handles_host /* world*64 */);
/* back to NCCL for every exchange: call on ALL ranks when any rank's crux_peer_handle / crux_peer_init failed */
int32_t crux_peer_disable(crux_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* CRUX_CUDA_H */
