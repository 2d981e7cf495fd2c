// Internal Gaussian policy object.  Not part of the ABI.
#pragma once
#include "mlp.cuh"

#define CRUX_MAX_ADIM 64
#define CRUX_GRAD_TAIL 128  // floats appended to an actor's gradient vector: [logΣ grad (64) | info sums (64)]

struct crux_gaussian {
  crux_ctx *ctx = nullptr;
  crux_mlp *mu = nullptr;       // borrowed
  int adim = 0;
  bool squashed = false;
  float ascale = 1.0f;
  bool head_mode = false;       // mu has 2*adim outputs: [mu | log_sigma]
  bool categorical = false;     // crux_categorical_create: `mu` is the logits network of a DiscreteNetwork actor, adim = number of actions
                                // (accepted by the on-policy updates only: a = one-hot rows, logprob = categorical_logpdf)
  // state-independent log-sigma (ConstantLayer): parameter, gradient (aliases mu->grads tail), Adam moments
  float *log_sigma = nullptr;
  float *ls_m = nullptr, *ls_v = nullptr;
  // PPO workspace
  float *mb = nullptr;          // gathered minibatch columns
  size_t mb_bytes = 0;
  int32_t *order = nullptr;     // device-generated permutations (actor epochs)
  size_t order_bytes = 0;
  int32_t *order2 = nullptr;    // ... critic epochs (they may run concurrently on the side stream)
  size_t order2_bytes = 0;
  float *info_actor = nullptr, *info_critic = nullptr;
  size_t info_actor_bytes = 0, info_critic_bytes = 0;
  int *ctl = nullptr;           // [0]=skip (actor early stop), [1]=pending stop
  double *partials = nullptr;   // head partial sums
};

// gradient tail accessors (mlp->grads is allocated with CRUX_GRAD_TAIL extra floats)
static inline float *tail_ls_grad(crux_mlp *m) { return m->grads + m->n_params; }
static inline float *tail_sums(crux_mlp *m) { return m->grads + m->n_params + 64; }

// ---- fused PPO update (ppo_fused.cu); *handled == 0 -> the generic engine in ppo.cu runs
int ppo_update_fused(crux_gaussian *actor, crux_mlp *critic, const float *s, const float *a, const float *logprob, const float *advantage,
                     const float *ret, int64_t n, const crux_ppo_hp *hp, const int32_t *order_actor, const int32_t *order_critic,
                     uint64_t seed, int *handled);
int ppo_fill_order(crux_ctx *ctx, int32_t *out, int64_t n, uint64_t seed, uint32_t epoch);
int ppo_fill_orders(crux_ctx *ctx, int32_t *out, int64_t n, uint64_t seed, uint32_t epoch0, int epochs);
int ppo_ensure_bytes(crux_ctx *ctx, void **p, size_t *have, size_t need);
