// Policies of the hot path: GaussianPolicy / SquashedGaussianPolicy heads (src/policies.jl:315-400),
// DiscreteNetwork helpers (:104-157), eps-greedy MixedPolicy (:474-494) and the batched rollout step
// that replaces one `step!` per env (src/sampler.jl:71-137).
#include "policy.cuh"

namespace {

#define LOG_SQRT_2PI 0.9189385332046727f
#define ENT_CONST 1.4189385332046727f

__device__ __forceinline__ float softplus_f(float x) {  // NNlib: log1p(exp(-|x|)) + relu(x)
  return log1pf(expf(-fabsf(x))) + fmaxf(x, 0.f);
}

// one thread per row.  net: [B][ld] with mu at [0,A) and (head mode) log-sigma at [A,2A)
// mode 0: explore (eps given or Philox) ; mode 1: logpdf of given action ; mode 2: greedy action
__global__ void gaussian_head_kernel(const float *__restrict__ net, int ld, const float *__restrict__ ls_const,
                                     int A, int squashed, float ascale, int mode, const float *__restrict__ eps_in,
                                     const float *__restrict__ a_in, uint64_t seed, uint64_t ctr, int64_t B,
                                     float *__restrict__ a_out, float *__restrict__ logp_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const float *row = net + i * ld;
  float logp = 0.f;
  float nrm[4];
  for (int j = 0; j < A; ++j) {
    const float mu = row[j];
    const float ls = ls_const ? ls_const[j] : row[A + j];
    if (mode == 2) {
      a_out[i * A + j] = squashed ? ascale * tanhf(mu) : mu;  // policies.jl:331,372
      continue;
    }
    const float sigma = squashed ? expf(fminf(fmaxf(ls, -5.f), 2.f)) : expf(ls);  // :374-380 / :340
    const float var = sigma * sigma;
    float a;  // pre-tanh action for the squashed policy
    if (mode == 0) {
      float e;
      if (eps_in) e = eps_in[i * A + j];
      else {
        if ((j & 3) == 0) {
          const Philox4 p = philox4x32_10(seed, ctr, (uint64_t)i * ((A + 3) / 4) + (j >> 2));
          box_muller(p.x, p.y, nrm[0], nrm[1]);
          box_muller(p.z, p.w, nrm[2], nrm[3]);
        }
        e = nrm[j & 3];
      }
      a = e * sigma + mu;  // :342,392
      a_out[i * A + j] = squashed ? ascale * tanhf(a) : a;
    } else {
      a = a_in[i * A + j];
      if (squashed) a = atanhf(fminf(fmaxf(a / ascale, -1.0f + 1.0e-5f), 1.0f - 1.0e-5f));  // :396
    }
    const float d = a - mu;
    float t = -(d * d) / (2.f * var) - LOG_SQRT_2PI - ls;  // :335 / :385 (unclamped logΣ)
    if (squashed) t -= 2.f * (0.6931471805599453f - a - softplus_f(-2.f * a));
    logp += t;
  }
  if (mode != 2 && logp_out) logp_out[i] = logp;
}

__global__ void gaussian_entropy_kernel(const float *__restrict__ net, int ld, const float *__restrict__ ls_const, int A,
                                        int64_t B, float *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (ls_const) {  // policies.jl:348: scalar, constant not scaled by A
    if (i == 0) {
      float s = 0.f;
      for (int j = 0; j < A; ++j) s += ls_const[j];
      out[0] = ENT_CONST + s;
    }
    return;
  }
  if (i >= B) return;
  float s = 0.f;
  for (int j = 0; j < A; ++j) s += net[i * ld + A + j];
  out[i] = ENT_CONST + s;  // :398
}

// ---------------------------------------------------------------- discrete
// NNlib softmax (max-subtracted) of value(π, s) ./ α: α = 1 is the default logit_conversion (policies.jl:108), any other α the
// SoftQ one (rl/softq.jl:48).  x / 1.f == x bit for bit.
__device__ __forceinline__ void softmax_row(const float *q, int nA, float alpha, float *p) {
  float m = q[0] / alpha;
  for (int a = 1; a < nA; ++a) m = fmaxf(m, q[a] / alpha);
  float s = 0.f;
  for (int a = 0; a < nA; ++a) { p[a] = expf(q[a] / alpha - m); s += p[a]; }
  for (int a = 0; a < nA; ++a) p[a] /= s;
}
#define MAX_NA 64
__global__ void discrete_argmax_kernel(const float *__restrict__ q, int64_t B, int nA, int32_t *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  int best = 0; float bv = q[i * nA];
  for (int a = 1; a < nA; ++a) { const float v = q[i * nA + a]; if (v > bv) { bv = v; best = a; } }
  out[i] = best;
}
__global__ void discrete_explore_kernel(const float *__restrict__ q, int64_t B, int nA, float alpha, const double *__restrict__ u_in,
                                        uint64_t seed, uint64_t ctr, int32_t *__restrict__ a_idx, float *__restrict__ logp) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  float p[MAX_NA];
  softmax_row(q + i * nA, nA, alpha, p);
  double u;
  if (u_in) u = u_in[i];
  else { const Philox4 r = philox4x32_10(seed, ctr, (uint64_t)i); u = u64_to_unit(r.x, r.y); }
  double c = 0.0; int k = 0;
  for (int a = 0; a < nA; ++a) { c += (double)p[a]; if (c < u) k = a + 1; }
  k = min(k, nA - 1);
  a_idx[i] = k;
  if (logp) logp[i] = logf(p[k]);  // categorical_logpdf :135
}
__global__ void discrete_logpdf_kernel(const float *__restrict__ q, const float *__restrict__ oh, int64_t B, int nA, float alpha,
                                       float *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  float p[MAX_NA];
  softmax_row(q + i * nA, nA, alpha, p);
  float s = 0.f;
  for (int a = 0; a < nA; ++a) s += p[a] * oh[i * nA + a];
  out[i] = logf(s);
}
__global__ void discrete_entropy_kernel(const float *__restrict__ q, int64_t B, int nA, float alpha, float *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  float p[MAX_NA];
  softmax_row(q + i * nA, nA, alpha, p);
  float s = 0.f;
  for (int a = 0; a < nA; ++a) s += p[a] * logf(p[a] + 1.1920929e-07f);  // eps(Float32) :154
  out[i] = -s;
}
__global__ void eps_greedy_kernel(const float *__restrict__ q, int64_t B, int nA, double eps, const double *__restrict__ u_in,
                                  uint64_t seed, uint64_t ctr, int32_t *__restrict__ a_idx, float *__restrict__ a_oh,
                                  float *__restrict__ logp) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  double coin, pick;
  if (u_in) { coin = u_in[2 * i]; pick = u_in[2 * i + 1]; }
  else { const Philox4 r = philox4x32_10(seed, ctr, (uint64_t)i); coin = u64_to_unit(r.x, r.y); pick = u64_to_unit(r.z, r.w); }
  int k;
  if (coin < eps) {  // policies.jl:476 rand() < ϵ -> uniform action
    k = min((int)(pick * nA), nA - 1);
  } else {
    k = 0; float bv = q[i * nA];
    for (int a = 1; a < nA; ++a) { const float v = q[i * nA + a]; if (v > bv) { bv = v; k = a; } }
  }
  a_idx[i] = k;
  if (a_oh) for (int a = 0; a < nA; ++a) a_oh[i * nA + a] = (a == k) ? 1.f : 0.f;
  if (logp) logp[i] = (float)log(eps * exp(log(1.0 / (double)nA)) + (1.0 - eps));  // :485-493
}

}  // namespace

extern "C" {

int32_t crux_gaussian_create(crux_ctx *ctx, crux_mlp *mu, int32_t adim, const float *log_sigma_host, int32_t squashed,
                             float ascale, crux_gaussian **out) {
  if (!ctx) return CRUX_ERR_INVALID;
  CRUX_REQUIRE(ctx, mu && out, "crux_gaussian_create: NULL argument");
  CRUX_REQUIRE(ctx, adim >= 1 && adim <= CRUX_MAX_ADIM, "crux_gaussian_create: action dim must be 1..64");
  const int outw = mu->dims[mu->n_layers];
  if (log_sigma_host) CRUX_REQUIRE(ctx, outw == adim, "crux_gaussian_create: mu output width != action dim");
  else CRUX_REQUIRE(ctx, outw == 2 * adim, "crux_gaussian_create: head mode needs 2*adim outputs [mu | logΣ]");
  crux_gaussian *p = new crux_gaussian();
  p->ctx = ctx; p->mu = mu; p->adim = adim; p->squashed = squashed != 0; p->ascale = ascale;
  p->head_mode = log_sigma_host == nullptr;
  if (cudaMalloc((void **)&p->ctl, 4 * sizeof(int)) != cudaSuccess || cudaMalloc((void **)&p->partials, 4096 * sizeof(double)) != cudaSuccess) {
    delete p; return crux_set_err(ctx, CRUX_ERR_OOM, "crux_gaussian_create: cudaMalloc");
  }
  cudaMemsetAsync(p->ctl, 0, 4 * sizeof(int), ctx->stream);
  if (log_sigma_host) {
    if (cudaMalloc((void **)&p->log_sigma, 3 * CRUX_MAX_ADIM * sizeof(float)) != cudaSuccess) { delete p; return crux_set_err(ctx, CRUX_ERR_OOM, "crux_gaussian_create: cudaMalloc"); }
    p->ls_m = p->log_sigma + CRUX_MAX_ADIM; p->ls_v = p->log_sigma + 2 * CRUX_MAX_ADIM;
    cudaMemsetAsync(p->log_sigma, 0, 3 * CRUX_MAX_ADIM * sizeof(float), ctx->stream);
    cudaMemcpyAsync(p->log_sigma, log_sigma_host, adim * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
  }
  *out = p;
  return CRUX_OK;
}

// DiscreteNetwork actor (policies.jl:104-157) for the on-policy updates: ppo_loss / a2c_loss / reinforce_loss are generic over the policy
// through logpdf (categorical_logpdf :135: log(sum(softmax(net(s)) .* a_onehot))) and entropy (:152-155: -sum(p log(p + eps(Float32)))).
int32_t crux_categorical_create(crux_ctx *ctx, crux_mlp *logits, int32_t n_actions, crux_gaussian **out) {
  if (!ctx) return CRUX_ERR_INVALID;
  CRUX_REQUIRE(ctx, logits && out, "crux_categorical_create: NULL argument");
  CRUX_REQUIRE(ctx, n_actions >= 2 && n_actions <= CRUX_MAX_ADIM, "crux_categorical_create: 2..64 actions");
  CRUX_REQUIRE(ctx, logits->dims[logits->n_layers] == n_actions, "crux_categorical_create: network output width != number of actions");
  crux_gaussian *p = new crux_gaussian();
  p->ctx = ctx; p->mu = logits; p->adim = n_actions; p->categorical = true;
  if (cudaMalloc((void **)&p->ctl, 4 * sizeof(int)) != cudaSuccess || cudaMalloc((void **)&p->partials, 4096 * sizeof(double)) != cudaSuccess) {
    delete p; return crux_set_err(ctx, CRUX_ERR_OOM, "crux_categorical_create: cudaMalloc");
  }
  cudaMemsetAsync(p->ctl, 0, 4 * sizeof(int), ctx->stream);
  *out = p;
  return CRUX_OK;
}

int32_t crux_gaussian_destroy(crux_gaussian *p) {
  if (!p) return CRUX_OK;
  cudaStreamSynchronize(p->ctx->stream);
  if (p->log_sigma) cudaFree(p->log_sigma);
  if (p->mb) cudaFree(p->mb);
  if (p->order) cudaFree(p->order);
  if (p->order2) cudaFree(p->order2);
  if (p->info_actor) cudaFree(p->info_actor);
  if (p->info_critic) cudaFree(p->info_critic);
  if (p->ctl) cudaFree(p->ctl);
  if (p->partials) cudaFree(p->partials);
  delete p;
  return CRUX_OK;
}

int32_t crux_gaussian_log_sigma_ptr(crux_gaussian *p, float **dev_out) {
  if (!p || !dev_out) return CRUX_ERR_INVALID;
  *dev_out = p->log_sigma;
  return CRUX_OK;
}

static int gaussian_head(crux_gaussian *p, const float *s, int64_t B, int mode, const float *eps_in, const float *a_in,
                         uint64_t seed, uint64_t ctr, float *a_out, float *logp_out) {
  crux_ctx *ctx = p->ctx;
  CRUX_REQUIRE(ctx, !p->categorical, "gaussian: the handle is a categorical (DiscreteNetwork) actor: use the crux_discrete_* entry points");
  CRUX_REQUIRE(ctx, B >= 0, "gaussian: negative batch");
  if (B == 0) return CRUX_OK;
  CRUX_REQUIRE(ctx, s, "gaussian: NULL states");
  int rc = mlp_forward_keep(p->mu, s, B, nullptr);
  if (rc) return rc;
  const int L = p->mu->n_layers;
  gaussian_head_kernel<<<(unsigned)cdiv(B, 128), 128, 0, ctx->stream>>>(p->mu->act[L], p->mu->dims[L], p->log_sigma, p->adim,
                                                                        p->squashed ? 1 : 0, p->ascale, mode, eps_in, a_in, seed,
                                                                        ctr, B, a_out, logp_out);
  CRUX_LAUNCHED(ctx);
  return CRUX_OK;
}

int32_t crux_gaussian_explore(crux_gaussian *p, const float *s, int64_t B, const float *eps_in, uint64_t seed, uint64_t ctr,
                              float *a_out, float *logp_out) {
  if (!p) return CRUX_ERR_INVALID;
  CRUX_REQUIRE(p->ctx, a_out, "crux_gaussian_explore: NULL a_out");
  return gaussian_head(p, s, B, 0, eps_in, nullptr, seed, ctr, a_out, logp_out);
}
int32_t crux_gaussian_action(crux_gaussian *p, const float *s, int64_t B, float *a_out) {
  if (!p) return CRUX_ERR_INVALID;
  CRUX_REQUIRE(p->ctx, a_out, "crux_gaussian_action: NULL a_out");
  return gaussian_head(p, s, B, 2, nullptr, nullptr, 0, 0, a_out, nullptr);
}
int32_t crux_gaussian_logpdf(crux_gaussian *p, const float *s, const float *a, int64_t B, float *out) {
  if (!p) return CRUX_ERR_INVALID;
  CRUX_REQUIRE(p->ctx, a && out, "crux_gaussian_logpdf: NULL pointer");
  return gaussian_head(p, s, B, 1, nullptr, a, 0, 0, nullptr, out);
}
int32_t crux_gaussian_entropy(crux_gaussian *p, const float *s, int64_t B, float *out) {
  if (!p) return CRUX_ERR_INVALID;
  crux_ctx *ctx = p->ctx;
  CRUX_REQUIRE(ctx, out, "crux_gaussian_entropy: NULL out");
  if (p->head_mode) {
    if (B <= 0) return CRUX_OK;
    int rc = mlp_forward_keep(p->mu, s, B, nullptr);
    if (rc) return rc;
  }
  const int L = p->mu->n_layers;
  const int64_t nb = p->head_mode ? cdiv(B, 128) : 1;
  gaussian_entropy_kernel<<<(unsigned)nb, 128, 0, ctx->stream>>>(p->head_mode ? p->mu->act[L] : nullptr, p->mu->dims[L], p->log_sigma,
                                                                 p->adim, B, out);
  CRUX_LAUNCHED(ctx);
  return CRUX_OK;
}

int32_t crux_rollout_step_fused(crux_gaussian *actor, crux_mlp *critic, const float *obs, int64_t N, const float *eps_in,
                                uint64_t seed, uint64_t ctr, float *a_out, float *logp_out, float *v_out, int *handled, int64_t row0);
// rows [row0, row0 + N) of a larger vector step: device noise streams stay keyed by the absolute stream id
int32_t crux_rollout_step_rows(crux_gaussian *actor, const float *obs, int64_t N, int64_t row0, uint64_t seed, uint64_t ctr, float *a_out,
                               float *logp_out) {
  int handled = 0;
  int rc = crux_rollout_step_fused(actor, nullptr, obs, N, nullptr, seed, ctr, a_out, logp_out, nullptr, &handled, row0);
  if (rc) return rc;
  if (!handled) return crux_set_err(actor->ctx, CRUX_ERR_STATE, "crux_rollout_step_rows: only the fused policy shapes support split vector steps");
  return CRUX_OK;
}

int32_t crux_rollout_step(crux_gaussian *actor, crux_mlp *critic, const float *obs, int64_t N, const float *eps_in,
                          uint64_t seed, uint64_t ctr, float *a_out, float *logp_out, float *v_out) {
  if (!actor) return CRUX_ERR_INVALID;
  crux_ctx *ctx = actor->ctx;
  CRUX_REQUIRE(ctx, N >= 0, "crux_rollout_step: negative N");
  if (N == 0) return CRUX_OK;
  CRUX_REQUIRE(ctx, obs && a_out, "crux_rollout_step: NULL pointer");
  int handled = 0;
  int rc = crux_rollout_step_fused(actor, critic, obs, N, eps_in, seed, ctr, a_out, logp_out, v_out, &handled, 0);
  if (rc || handled) return rc;
  rc = gaussian_head(actor, obs, N, 0, eps_in, nullptr, seed, ctr, a_out, logp_out);
  if (rc) return rc;
  if (critic && v_out) {
    CRUX_REQUIRE(ctx, critic->dims[critic->n_layers] == 1, "crux_rollout_step: critic must have one output");
    rc = mlp_forward_out(critic, obs, N, v_out);
  }
  return rc;
}

// ---------------------------------------------------------------- discrete ABI
#define DISCRETE_GUARD(ctx, nA) CRUX_REQUIRE(ctx, (nA) >= 1 && (nA) <= MAX_NA, "discrete: 1..64 actions supported")

int32_t crux_discrete_argmax(crux_ctx *ctx, const float *q, int64_t B, int32_t nA, int32_t *a_idx) {
  if (!ctx) return CRUX_ERR_INVALID;
  DISCRETE_GUARD(ctx, nA);
  if (B <= 0) return CRUX_OK;
  discrete_argmax_kernel<<<(unsigned)cdiv(B, 128), 128, 0, ctx->stream>>>(q, B, nA, a_idx);
  CRUX_LAUNCHED(ctx);
  return CRUX_OK;
}
int32_t crux_discrete_explore_t(crux_ctx *ctx, const float *q, int64_t B, int32_t nA, float alpha, const double *u_in, uint64_t seed,
                                uint64_t ctr, int32_t *a_idx, float *logp) {
  if (!ctx) return CRUX_ERR_INVALID;
  DISCRETE_GUARD(ctx, nA);
  CRUX_REQUIRE(ctx, alpha > 0.f, "crux_discrete_explore: temperature must be positive");
  if (B <= 0) return CRUX_OK;
  discrete_explore_kernel<<<(unsigned)cdiv(B, 128), 128, 0, ctx->stream>>>(q, B, nA, alpha, u_in, seed, ctr, a_idx, logp);
  CRUX_LAUNCHED(ctx);
  return CRUX_OK;
}
int32_t crux_discrete_explore(crux_ctx *ctx, const float *q, int64_t B, int32_t nA, const double *u_in, uint64_t seed,
                              uint64_t ctr, int32_t *a_idx, float *logp) {
  return crux_discrete_explore_t(ctx, q, B, nA, 1.f, u_in, seed, ctr, a_idx, logp);
}
int32_t crux_discrete_logpdf_t(crux_ctx *ctx, const float *q, const float *a_onehot, int64_t B, int32_t nA, float alpha, float *out) {
  if (!ctx) return CRUX_ERR_INVALID;
  DISCRETE_GUARD(ctx, nA);
  CRUX_REQUIRE(ctx, alpha > 0.f, "crux_discrete_logpdf: temperature must be positive");
  if (B <= 0) return CRUX_OK;
  discrete_logpdf_kernel<<<(unsigned)cdiv(B, 128), 128, 0, ctx->stream>>>(q, a_onehot, B, nA, alpha, out);
  CRUX_LAUNCHED(ctx);
  return CRUX_OK;
}
int32_t crux_discrete_logpdf(crux_ctx *ctx, const float *q, const float *a_onehot, int64_t B, int32_t nA, float *out) {
  return crux_discrete_logpdf_t(ctx, q, a_onehot, B, nA, 1.f, out);
}
int32_t crux_discrete_entropy_t(crux_ctx *ctx, const float *q, int64_t B, int32_t nA, float alpha, float *out) {
  if (!ctx) return CRUX_ERR_INVALID;
  DISCRETE_GUARD(ctx, nA);
  CRUX_REQUIRE(ctx, alpha > 0.f, "crux_discrete_entropy: temperature must be positive");
  if (B <= 0) return CRUX_OK;
  discrete_entropy_kernel<<<(unsigned)cdiv(B, 128), 128, 0, ctx->stream>>>(q, B, nA, alpha, out);
  CRUX_LAUNCHED(ctx);
  return CRUX_OK;
}
int32_t crux_discrete_entropy(crux_ctx *ctx, const float *q, int64_t B, int32_t nA, float *out) {
  return crux_discrete_entropy_t(ctx, q, B, nA, 1.f, out);
}
int32_t crux_discrete_eps_greedy(crux_ctx *ctx, const float *q, int64_t B, int32_t nA, double eps, const double *u_in,
                                 uint64_t seed, uint64_t ctr, int32_t *a_idx, float *a_onehot, float *logp) {
  if (!ctx) return CRUX_ERR_INVALID;
  DISCRETE_GUARD(ctx, nA);
  if (B <= 0) return CRUX_OK;
  eps_greedy_kernel<<<(unsigned)cdiv(B, 128), 128, 0, ctx->stream>>>(q, B, nA, eps, u_in, seed, ctr, a_idx, a_onehot, logp);
  CRUX_LAUNCHED(ctx);
  return CRUX_OK;
}

}  // extern "C"
