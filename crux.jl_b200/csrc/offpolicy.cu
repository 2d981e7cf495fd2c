// Off-policy updates (src/model_free/off_policy.jl:66-111): the DQN critic step (td_loss, utils.jl:76-87),
// one SAC value_training epoch (rl/sac.jl:4-52): target -> temperature -> double-Q critic -> actor -> polyak,
// and one DDPG / TD3 epoch (rl/ddpg.jl:6-25, rl/td3.jl:4-12): (smoothed) target -> critic(s) -> deterministic actor -> polyak.
#include "policy.cuh"
#include "conv.cuh"

namespace {

#define LOG_SQRT_2PI 0.9189385332046727f

__device__ __forceinline__ float softplus_f(float x) { return log1pf(expf(-fabsf(x))) + fmaxf(x, 0.f); }

__device__ __forceinline__ double block_sum_d(double v, double *sh /*32*/) {
  v = warp_sum_d(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x < 32) {
    r = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
    r = warp_sum_d(r);
  }
  __syncthreads();
  return r;  // valid on thread 0
}

// DQN: Q(s,a) = sum_a Q(s)[a]*onehot[a]; loss = agg((Q(s,a)-y)^2); part[b] = {sum e*w, sum Qsa}
__global__ void dqn_head_kernel(const float *__restrict__ q, const float *__restrict__ oh, const float *__restrict__ y,
                                const float *__restrict__ w, int64_t B, int nA, float inv_bg, float *__restrict__ dq,
                                double *__restrict__ part) {
  __shared__ double sh[32];
  double se = 0.0, sq = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) {
    float qsa = 0.f;
    for (int a = 0; a < nA; ++a) qsa += q[i * nA + a] * oh[i * nA + a];
    const float d = qsa - y[i];
    const float wi = w ? w[i] : 1.f;
    se += (double)(d * d * wi); sq += (double)qsa;
    const float g = 2.f * d * wi * inv_bg;
    for (int a = 0; a < nA; ++a) dq[i * nA + a] = g * oh[i * nA + a];
  }
  double r = block_sum_d(se, sh);
  if (threadIdx.x == 0) part[2 * blockIdx.x] = r;
  r = block_sum_d(sq, sh);
  if (threadIdx.x == 0) part[2 * blockIdx.x + 1] = r;
}
__global__ void dqn_finalize_kernel(const double *__restrict__ part, int nb, float *__restrict__ sums) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < nb; ++i) { a += part[2 * i]; b += part[2 * i + 1]; }
    sums[0] = (float)a; sums[1] = (float)b;
  }
}
__global__ void dqn_record_kernel(const float *__restrict__ sums, float cnt_local, int world, float *__restrict__ info) {
  const float cnt = cnt_local * (float)world;
  info[0] = sums[0] / cnt;  // loss
  info[2] = sums[1] / cnt;  // Qavg (utils.jl:81)
}

// ---- SAC ---------------------------------------------------------------------------------------------
// squashed-Gaussian sample from net [B][2A] = [mu | logΣ]: a (tanh-squashed, scaled), logprob, optional noise out
__global__ void sac_sample_kernel(const float *__restrict__ net, int A, float ascale, const float *__restrict__ eps_in, uint64_t seed,
                                  uint64_t ctr, int64_t B, float *__restrict__ a_out, float *__restrict__ logp_out,
                                  float *__restrict__ eps_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  float logp = 0.f, nrm[4];
  for (int j = 0; j < A; ++j) {
    const float mu = net[i * 2 * A + j], ls = net[i * 2 * A + A + j];
    const float sg = expf(fminf(fmaxf(ls, -5.f), 2.f));
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
    if (eps_out) eps_out[i * A + j] = e;
    const float ap = e * sg + mu;
    const float d = ap - mu;
    logp += -(d * d) / (2.f * (sg * sg)) - LOG_SQRT_2PI - ls - 2.f * (0.6931471805599453f - ap - softplus_f(-2.f * ap));
    if (a_out) a_out[i * A + j] = ascale * tanhf(ap);
  }
  if (logp_out) logp_out[i] = logp;
}

// temperature: loss = -mean(exp(logα)(logp + H)); single block: reduce, Flux Adam on the 1-element array.
// st = {m, v} float, step counter int. info[0] = temp_loss
// mode 0: the whole step on one rank.  Several ranks: mode 1 leaves this rank's sum of (logp + H) in sum_io[0], the caller all-reduces
// it, mode 2 applies the step with the mean over all world * B rows (identical on every rank).
__global__ void sac_temp_kernel(const float *__restrict__ logp, int64_t B, float h_target, float *__restrict__ log_alpha,
                                float *__restrict__ st, int *__restrict__ step, double eta, float *__restrict__ info,
                                unsigned int *__restrict__ err_flags, int mode, int world, float *__restrict__ sum_io) {
  __shared__ double sh[32];
  double s = 0.0;
  if (mode != 2)
    for (int64_t i = threadIdx.x; i < B; i += blockDim.x) s += (double)(logp[i] + h_target);
  double tot = block_sum_d(s, sh);
  if (threadIdx.x == 0) {
    if (mode == 1) { sum_io[0] = (float)tot; return; }
    if (mode == 2) tot = (double)sum_io[0];
    const float alpha = expf(log_alpha[0]);
    const float mean_t = (float)(tot / ((double)B * (double)world));
    const float loss = -(alpha * mean_t);
    const float g = -(alpha * mean_t);  // d/dlogα of -mean(exp(logα) t) = -exp(logα) mean(t)
    info[0] = loss;
    if (isnan(g)) { atomicOr(err_flags, CRUX_FLAG_NAN); return; }
    const double b1 = 0.9, b2 = 0.999, eps = 1e-8;
    const int t = ++(*step);
    const float mt = (float)(b1 * (double)st[0] + (1.0 - b1) * (double)g);
    const float vt = (float)(b2 * (double)st[1] + (1.0 - b2) * (double)g * (double)g);
    st[0] = mt; st[1] = vt;
    const float delta = (float)((double)mt / (1.0 - pow(b1, (double)t)) / (sqrt((double)vt / (1.0 - pow(b2, (double)t))) + eps) * eta);
    log_alpha[0] = log_alpha[0] - delta;
  }
}

// double_Q_loss (utils.jl:89-96): .5(mse(Q1,y)+mse(Q2,y)); part[b] = {sum e1, sum e2, sum q1, sum q2}
__global__ void sac_critic_head_kernel(const float *__restrict__ q1, const float *__restrict__ q2, const float *__restrict__ y, int64_t B,
                                       float inv_b, float *__restrict__ dq1, float *__restrict__ dq2, double *__restrict__ part) {
  __shared__ double sh[32];
  double e1 = 0.0, e2 = 0.0, s1 = 0.0, s2 = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) {
    const float d1 = q1[i] - y[i], d2 = q2[i] - y[i];
    e1 += (double)(d1 * d1); e2 += (double)(d2 * d2); s1 += (double)q1[i]; s2 += (double)q2[i];
    dq1[i] = 0.5f * 2.f * d1 * inv_b; dq2[i] = 0.5f * 2.f * d2 * inv_b;
  }
  double r;
  r = block_sum_d(e1, sh); if (threadIdx.x == 0) part[4 * blockIdx.x + 0] = r;
  r = block_sum_d(e2, sh); if (threadIdx.x == 0) part[4 * blockIdx.x + 1] = r;
  r = block_sum_d(s1, sh); if (threadIdx.x == 0) part[4 * blockIdx.x + 2] = r;
  r = block_sum_d(s2, sh); if (threadIdx.x == 0) part[4 * blockIdx.x + 3] = r;
}
__global__ void sac_critic_record_kernel(const double *__restrict__ part, int nb, int64_t B, float *__restrict__ info) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double a = 0, b = 0, c = 0, d = 0;
    for (int i = 0; i < nb; ++i) { a += part[4 * i]; b += part[4 * i + 1]; c += part[4 * i + 2]; d += part[4 * i + 3]; }
    const float l1 = (float)(a / (double)B), l2 = (float)(b / (double)B);
    info[1] = 0.5f * (l1 + l2);
    info[6] = (float)(c / (double)B); info[7] = (float)(d / (double)B);
  }
}

// actor loss forward part: which critic is the min (ties keep Q1), dQ seeds, loss partials {sum (α logp - qmin), sum logp}
__global__ void sac_actor_seed_kernel(const float *__restrict__ q1, const float *__restrict__ q2, const float *__restrict__ logp,
                                      const float *__restrict__ log_alpha, int64_t B, float inv_b, float *__restrict__ dq1,
                                      float *__restrict__ dq2, double *__restrict__ part) {
  __shared__ double sh[32];
  const float alpha = expf(log_alpha[0]);
  double sl = 0.0, sp = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) {
    const bool second = q2[i] < q1[i];
    const float qm = second ? q2[i] : q1[i];
    sl += (double)(alpha * logp[i] - qm); sp += (double)logp[i];
    dq1[i] = second ? 0.f : -inv_b; dq2[i] = second ? -inv_b : 0.f;
  }
  double r;
  r = block_sum_d(sl, sh); if (threadIdx.x == 0) part[2 * blockIdx.x] = r;
  r = block_sum_d(sp, sh); if (threadIdx.x == 0) part[2 * blockIdx.x + 1] = r;
}
__global__ void sac_actor_record_kernel(const double *__restrict__ part, int nb, int64_t B, float *__restrict__ info) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double a = 0, b = 0;
    for (int i = 0; i < nb; ++i) { a += part[2 * i]; b += part[2 * i + 1]; }
    info[3] = (float)(a / (double)B);       // actor loss
    info[5] = (float)(-(b / (double)B));    // entropy = -mean(logprob)  sac.jl:37
  }
}
// dL/dnet for the actor: net = [mu | logΣ], a_pre = eps*sigma + mu.
//   dL/da_pre_j = (α/B) 2 tanh(a_pre_j) + (dL/da_j) ascale (1 - tanh²(a_pre_j)),  dL/da = dcat1[sdim+j] + dcat2[sdim+j]
//   dL/dmu_j = dL/da_pre_j ; dL/dlogΣ_j = dL/da_pre_j eps_j sigma_j [-5<=logΣ<=2] - α/B
// (the Gaussian quadratic term (a_pre-mu)²/2σ² is constant under the reparameterisation: its two gradient paths cancel)
__global__ void sac_actor_bwd_head_kernel(const float *__restrict__ net, const float *__restrict__ eps, const float *__restrict__ dcat1,
                                          const float *__restrict__ dcat2, int sdim, int A, float ascale,
                                          const float *__restrict__ log_alpha, int64_t B, float inv_b, float *__restrict__ dnet) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const float alpha = expf(log_alpha[0]);
  const int ld = sdim + A;
  for (int j = 0; j < A; ++j) {
    const float mu = net[i * 2 * A + j], ls = net[i * 2 * A + A + j];
    const bool inside = ls >= -5.f && ls <= 2.f;
    const float sg = expf(fminf(fmaxf(ls, -5.f), 2.f));
    const float e = eps[i * A + j];
    const float ap = e * sg + mu;
    const float th = tanhf(ap);
    const float da = dcat1[i * ld + sdim + j] + dcat2[i * ld + sdim + j];
    const float g = alpha * inv_b * 2.f * th + da * ascale * (1.f - th * th);
    dnet[i * 2 * A + j] = g;
    dnet[i * 2 * A + A + j] = (inside ? g * e * sg : 0.f) - alpha * inv_b;
  }
}

// ---- DDPG / TD3 -----------------------------------------------------------------------------------------
// exploration(::GaussianNoiseExplorationPolicy) policies.jl:510-514 on an action batch that action(π_on, s) already filled:
//   a = clamp(a + clamp(randn * σ(i), ϵ_min, ϵ_max), a_min, a_max); a_min / a_max broadcast when they have one entry.
struct NoiseBounds { float lo[32], hi[32]; int n_lo, n_hi; };
__global__ void noise_clamp_kernel(float *__restrict__ a, int64_t B, int A, float sigma, float eps_min, float eps_max, NoiseBounds nb,
                                   const float *__restrict__ eps_in, uint64_t seed, uint64_t ctr) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  float nrm[4];
  for (int j = 0; j < A; ++j) {
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
    const float n = fminf(fmaxf(e * sigma, eps_min), eps_max);
    const float lo = nb.lo[nb.n_lo == 1 ? 0 : j], hi = nb.hi[nb.n_hi == 1 ? 0 : j];
    a[i * A + j] = fminf(fmaxf(a[i * A + j] + n, lo), hi);
  }
}
// ddpg_target rl/ddpg.jl:6-8 / td3_target rl/td3.jl:4-7: y = r + γ(1-done)·Q⁻(sp, a') or ·min(Q1⁻, Q2⁻)(sp, a')
__global__ void ddpg_target_kernel(const float *__restrict__ r, const uint8_t *__restrict__ done, const float *__restrict__ q1,
                                   const float *__restrict__ q2, int64_t B, float gamma, float *__restrict__ y) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const float q = q2 ? fminf(q1[i], q2[i]) : q1[i];
  y[i] = r[i] + (gamma * (1.0f - (float)done[i])) * q;
}
// td_loss utils.jl:76-87 on a ContinuousNetwork critic: mse(Q(s,a), y); part[b] = {sum e, sum q}
__global__ void q_mse_head_kernel(const float *__restrict__ q, const float *__restrict__ y, int64_t B, float inv_b, float *__restrict__ dq,
                                  double *__restrict__ part) {
  __shared__ double sh[32];
  double e = 0.0, sq = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = q[i] - y[i];
    e += (double)(d * d); sq += (double)q[i];
    dq[i] = 2.f * d * inv_b;
  }
  double r;
  r = block_sum_d(e, sh); if (threadIdx.x == 0) part[2 * blockIdx.x] = r;
  r = block_sum_d(sq, sh); if (threadIdx.x == 0) part[2 * blockIdx.x + 1] = r;
}
__global__ void q_mse_record_kernel(const double *__restrict__ part, int nb, int64_t B, float *__restrict__ info) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double a = 0, b = 0;
    for (int i = 0; i < nb; ++i) { a += part[2 * i]; b += part[2 * i + 1]; }
    info[1] = (float)(a / (double)B);
    info[6] = (float)(b / (double)B);
  }
}
// ddpg_actor_loss rl/ddpg.jl:25 / td3_actor_loss rl/td3.jl:12: -mean(Q(s, μ(s))): dL/dQ = -1/B; part[b] = sum q
__global__ void ddpg_actor_seed_kernel(const float *__restrict__ q, int64_t B, float inv_b, float *__restrict__ dq, double *__restrict__ part) {
  __shared__ double sh[32];
  double sq = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) {
    sq += (double)q[i];
    dq[i] = -inv_b;
  }
  const double r = block_sum_d(sq, sh);
  if (threadIdx.x == 0) part[blockIdx.x] = r;
}
__global__ void ddpg_actor_record_kernel(const double *__restrict__ part, int nb, int64_t B, float *__restrict__ info) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double a = 0;
    for (int i = 0; i < nb; ++i) a += part[i];
    info[3] = (float)(-(a / (double)B));
  }
}
// dL/da = the action columns of dL/dvcat(s, a)
__global__ void slice_action_grad_kernel(const float *__restrict__ dcat, int sdim, int A, int64_t B, float *__restrict__ da) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B * A; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / A; const int j = (int)(i % A);
    da[i] = dcat[b * (sdim + A) + sdim + j];
  }
}

__global__ void concat2_kernel(const float *__restrict__ s, int sd, const float *__restrict__ a, int ad, int64_t B, float *__restrict__ out) {
  const int d = sd + ad;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B * d; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / d; const int c = (int)(i % d);
    out[i] = c < sd ? s[b * sd + c] : a[b * ad + (c - sd)];
  }
}

}  // namespace

struct crux_sac_state {
  crux_ctx *ctx = nullptr;
  crux_gaussian *actor = nullptr;
  crux_mlp *q1 = nullptr, *q2 = nullptr, *q1t = nullptr, *q2t = nullptr;
  float h_target = 0.f, tau = 0.005f;
  double alpha_eta = (double)3e-4f;
  float *log_alpha = nullptr;   // device: [0] log α, [1..2] Adam m, v
  int *alpha_step = nullptr;
  float *info = nullptr;        // device [8]
  float *ws = nullptr;          // workspace
  size_t ws_bytes = 0;
  double *part = nullptr;
};

struct crux_ddpg_state {
  crux_ctx *ctx = nullptr;
  crux_mlp *actor = nullptr, *actor_t = nullptr;
  crux_mlp *q1 = nullptr, *q2 = nullptr, *q1t = nullptr, *q2t = nullptr;   // q2 == nullptr: DDPG (one critic)
  float tau = 0.005f;
  float *info = nullptr;        // device [8]
  float *ws = nullptr;
  size_t ws_bytes = 0;
  double *part = nullptr;
};

extern "C" {

int32_t crux_dqn_train(crux_mlp *q, const float *s, const float *a_onehot, const float *y, const float *weight, int64_t B,
                       float *info_out_host) {
  if (!q) return CRUX_ERR_INVALID;
  crux_ctx *ctx = q->ctx;
  CRUX_REQUIRE(ctx, B >= 1 && s && a_onehot && y, "crux_dqn_train: bad arguments");
  const int L = q->n_layers, nA = q->dims[L];
  int rc = mlp_forward_keep(q, s, B, nullptr); if (rc) return rc;
  const int nb = (int)i64min(cdiv(B, 256), 256);
  char *sc = (char *)crux_scratch(ctx, 3, 512 * sizeof(double) + 64);
  if (!sc) return CRUX_ERR_OOM;
  float *info_dev = (float *)sc;
  double *part = (double *)(sc + 64);
  const float inv_bg = 1.0f / ((float)B * (float)ctx->world);
  dqn_head_kernel<<<nb, 256, 0, ctx->stream>>>(q->act[L], a_onehot, y, weight, B, nA, inv_bg, q->dz[L], part);
  CRUX_LAUNCHED(ctx);
  float *sums = q->grads + q->n_params + 64;
  dqn_finalize_kernel<<<1, 32, 0, ctx->stream>>>(part, nb, sums);
  CRUX_LAUNCHED(ctx);
  rc = mlp_backward(q, s, B, q->dz[L], false, false, true, nullptr); if (rc) return rc;
  if (ctx->world > 1) { rc = grads_allreduce(ctx, q->grads, q->n_params + CRUX_GRAD_TAIL); if (rc) return rc; }
  dqn_record_kernel<<<1, 1, 0, ctx->stream>>>(sums, (float)B, ctx->world, info_dev);
  CRUX_LAUNCHED(ctx);
  rc = mlp_adam_step(q, info_dev + 1, nullptr); if (rc) return rc;
  if (info_out_host) {
    CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(info_out_host, info_dev, 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    return crux_ctx_check(ctx);
  }
  return CRUX_OK;
}

int32_t crux_convq_dqn_train(crux_convq *net, const void *s, int32_t s_is_u8, const float *a_onehot, const float *y, const float *weight, int64_t B,
                             float *info_out_host) {
  if (!net) return CRUX_ERR_INVALID;
  crux_ctx *ctx = net->ctx;
  CRUX_REQUIRE(ctx, B >= 1 && s && a_onehot && y, "crux_convq_dqn_train: bad arguments");
  CRUX_REQUIRE(ctx, ctx->world == 1, "crux_convq_dqn_train: single rank only (the conv gradients are not exchanged yet)");
  crux_mlp *h = net->head;
  const int L = h->n_layers, nA = h->dims[L];
  int rc = convq_forward_keep(net, s, s_is_u8, B); if (rc) return rc;
  const int nb = (int)i64min(cdiv(B, 256), 256);
  char *sc = (char *)crux_scratch(ctx, 3, 512 * sizeof(double) + 64);
  if (!sc) return CRUX_ERR_OOM;
  float *info_dev = (float *)sc;
  double *part = (double *)(sc + 64);
  dqn_head_kernel<<<nb, 256, 0, ctx->stream>>>(h->act[L], a_onehot, y, weight, B, nA, 1.0f / (float)B, h->dz[L], part);
  CRUX_LAUNCHED(ctx);
  float *sums = h->grads + h->n_params + 64;
  dqn_finalize_kernel<<<1, 32, 0, ctx->stream>>>(part, nb, sums);
  CRUX_LAUNCHED(ctx);
  rc = convq_backward(net, s, s_is_u8, B, h->dz[L]); if (rc) return rc;
  dqn_record_kernel<<<1, 1, 0, ctx->stream>>>(sums, (float)B, 1, info_dev);
  CRUX_LAUNCHED(ctx);
  rc = convq_adam_step(net, info_dev + 1); if (rc) return rc;
  if (info_out_host) {
    CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(info_out_host, info_dev, 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    return crux_ctx_check(ctx);
  }
  return CRUX_OK;
}

int32_t crux_sac_create(crux_gaussian *actor, crux_mlp *q1, crux_mlp *q2, crux_mlp *q1_target, crux_mlp *q2_target, float log_alpha,
                        float h_target, double alpha_eta, float tau, crux_sac_state **out) {
  if (!actor) return CRUX_ERR_INVALID;
  crux_ctx *ctx = actor->ctx;
  CRUX_REQUIRE(ctx, q1 && q2 && q1_target && q2_target && out, "crux_sac_create: NULL argument");
  CRUX_REQUIRE(ctx, actor->squashed && actor->head_mode, "crux_sac_create: actor must be a SquashedGaussianPolicy with [mu|logΣ] heads");
  const int sdim = actor->mu->dims[0], A = actor->adim;
  CRUX_REQUIRE(ctx, q1->dims[0] == sdim + A && q2->dims[0] == sdim + A && q1->dims[q1->n_layers] == 1 && q2->dims[q2->n_layers] == 1,
               "crux_sac_create: critics must map vcat(s,a) -> 1");
  CRUX_REQUIRE(ctx, q1->n_params == q1_target->n_params && q2->n_params == q2_target->n_params, "crux_sac_create: target shape mismatch");
  crux_sac_state *st = new crux_sac_state();
  st->ctx = ctx; st->actor = actor; st->q1 = q1; st->q2 = q2; st->q1t = q1_target; st->q2t = q2_target;
  st->h_target = h_target; st->tau = tau; st->alpha_eta = alpha_eta;
  if (cudaMalloc((void **)&st->log_alpha, 4 * sizeof(float)) != cudaSuccess || cudaMalloc((void **)&st->alpha_step, sizeof(int)) != cudaSuccess ||
      cudaMalloc((void **)&st->info, 8 * sizeof(float)) != cudaSuccess || cudaMalloc((void **)&st->part, 4096 * sizeof(double)) != cudaSuccess) {
    crux_sac_destroy(st);
    return crux_set_err(ctx, CRUX_ERR_OOM, "crux_sac_create: cudaMalloc");
  }
  const float init[4] = {log_alpha, 0.f, 0.f, 0.f};
  cudaMemcpyAsync(st->log_alpha, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream);
  cudaMemsetAsync(st->alpha_step, 0, sizeof(int), ctx->stream);
  cudaMemsetAsync(st->info, 0, 8 * sizeof(float), ctx->stream);
  cudaStreamSynchronize(ctx->stream);
  *out = st;
  return CRUX_OK;
}

int32_t crux_sac_destroy(crux_sac_state *st) {
  if (!st) return CRUX_OK;
  cudaStreamSynchronize(st->ctx->stream);
  if (st->log_alpha) cudaFree(st->log_alpha);
  if (st->alpha_step) cudaFree(st->alpha_step);
  if (st->info) cudaFree(st->info);
  if (st->ws) cudaFree(st->ws);
  if (st->part) cudaFree(st->part);
  delete st;
  return CRUX_OK;
}

int32_t crux_sac_log_alpha(crux_sac_state *st, float *out_host) {
  if (!st || !out_host) return CRUX_ERR_INVALID;
  CRUX_CHECK_CUDA(st->ctx, cudaMemcpyAsync(out_host, st->log_alpha, sizeof(float), cudaMemcpyDeviceToHost, st->ctx->stream));
  CRUX_CHECK_CUDA(st->ctx, cudaStreamSynchronize(st->ctx->stream));
  return CRUX_OK;
}

int32_t crux_sac_train(crux_sac_state *st, const float *s, const float *a, const float *sp, const float *r, const uint8_t *done,
                       int64_t B, float gamma, const float *eps_target, const float *eps_temp, const float *eps_actor, uint64_t seed,
                       uint64_t ctr, float *y_out, float *info_out_host) {
  if (!st) return CRUX_ERR_INVALID;
  crux_ctx *ctx = st->ctx;
  CRUX_REQUIRE(ctx, B >= 1 && s && a && sp && r && done, "crux_sac_train: bad arguments");
  // Several ranks (replicas with per-rank replay shards): every loss is a mean over the world * B rows of all ranks -- the heads
  // scale by 1 / (world B), the gradients of the critics and of the actor are summed over ranks before their optimiser steps, the
  // temperature step uses the all-reduced sum; parameters, targets and log α stay identical on every rank.  Device noise streams are
  // decorrelated per rank.  Info values are this rank's means.
  const int world = ctx->world;
  if (world > 1) seed ^= 0x9E3779B97F4A7C15ULL * (uint64_t)(ctx->rank + 1);
  crux_gaussian *pol = st->actor;
  crux_mlp *net = pol->mu;
  const int sdim = net->dims[0], A = pol->adim, ld = sdim + A, La = net->n_layers;
  // workspace: ap[B][A], logp[B], cat[B][ld], q1v[B], q2v[B], y[B], eps[B][A], dcat1[B][ld], dcat2[B][ld], dq1[B], dq2[B]
  const size_t nfl = (size_t)B * (A + 1 + ld + 3 + A + 2 * ld + 2);
  if (st->ws_bytes < nfl * sizeof(float)) {
    CRUX_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (st->ws) cudaFree(st->ws);
    st->ws = nullptr; st->ws_bytes = 0;
    if (cudaMalloc((void **)&st->ws, nfl * sizeof(float) * 5 / 4) != cudaSuccess) return crux_set_err(ctx, CRUX_ERR_OOM, "crux_sac_train: workspace");
    st->ws_bytes = nfl * sizeof(float) * 5 / 4;
  }
  float *ap = st->ws, *logp = ap + (size_t)B * A, *cat = logp + B, *q1v = cat + (size_t)B * ld, *q2v = q1v + B, *y = q2v + B,
        *epsb = y + B, *dcat1 = epsb + (size_t)B * A, *dcat2 = dcat1 + (size_t)B * ld, *dq1 = dcat2 + (size_t)B * ld, *dq2 = dq1 + B;
  const unsigned rb = (unsigned)cdiv(B, 128);
  const unsigned eb = (unsigned)i64min(cdiv(B * ld, 256), (int64_t)ctx->num_sms * 8);
  const int nb = (int)i64min(cdiv(B, 256), 256);
  const float inv_b = 1.0f / ((float)B * (float)world);
  int rc;

  // 1. y = sac_target (rl/sac.jl:4-9): a' ~ online actor(sp), target critics
  rc = mlp_forward_keep(net, sp, B, nullptr); if (rc) return rc;
  sac_sample_kernel<<<rb, 128, 0, ctx->stream>>>(net->act[La], A, pol->ascale, eps_target, seed, ctr, B, ap, logp, nullptr);
  CRUX_LAUNCHED(ctx);
  concat2_kernel<<<eb, 256, 0, ctx->stream>>>(sp, sdim, ap, A, B, cat);
  CRUX_LAUNCHED(ctx);
  rc = mlp_forward_out(st->q1t, cat, B, q1v); if (rc) return rc;
  rc = mlp_forward_out(st->q2t, cat, B, q2v); if (rc) return rc;
  rc = crux_sac_target(ctx, r, done, q1v, q2v, logp, B, gamma, st->log_alpha, y); if (rc) return rc;
  if (y_out) CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(y_out, y, (size_t)B * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));

  // 2. temperature step (off_policy.jl:86-88 with sac_temp_loss rl/sac.jl:45-52): fresh noise on s
  rc = mlp_forward_keep(net, s, B, nullptr); if (rc) return rc;
  sac_sample_kernel<<<rb, 128, 0, ctx->stream>>>(net->act[La], A, pol->ascale, eps_temp, seed, ctr + 1, B, nullptr, logp, nullptr);
  CRUX_LAUNCHED(ctx);
  if (world == 1) {
    sac_temp_kernel<<<1, 256, 0, ctx->stream>>>(logp, B, st->h_target, st->log_alpha, st->log_alpha + 1, st->alpha_step, st->alpha_eta,
                                                st->info, ctx->flags_dev, 0, 1, nullptr);
    CRUX_LAUNCHED(ctx);
  } else {
    sac_temp_kernel<<<1, 256, 0, ctx->stream>>>(logp, B, st->h_target, st->log_alpha, st->log_alpha + 1, st->alpha_step, st->alpha_eta,
                                                st->info, ctx->flags_dev, 1, world, st->log_alpha + 3);
    CRUX_LAUNCHED(ctx);
    rc = grads_allreduce(ctx, st->log_alpha + 3, 1); if (rc) return rc;
    sac_temp_kernel<<<1, 32, 0, ctx->stream>>>(logp, B, st->h_target, st->log_alpha, st->log_alpha + 1, st->alpha_step, st->alpha_eta,
                                               st->info, ctx->flags_dev, 2, world, st->log_alpha + 3);
    CRUX_LAUNCHED(ctx);
  }

  // 3. critic step (off_policy.jl:91-93, double_Q_loss utils.jl:89-96): one optimiser over (Q1, Q2)
  concat2_kernel<<<eb, 256, 0, ctx->stream>>>(s, sdim, a, A, B, cat);
  CRUX_LAUNCHED(ctx);
  rc = mlp_forward_keep(st->q1, cat, B, nullptr); if (rc) return rc;
  rc = mlp_forward_keep(st->q2, cat, B, nullptr); if (rc) return rc;
  sac_critic_head_kernel<<<nb, 256, 0, ctx->stream>>>(st->q1->act[st->q1->n_layers], st->q2->act[st->q2->n_layers], y, B, inv_b, dq1, dq2,
                                                     st->part);
  CRUX_LAUNCHED(ctx);
  sac_critic_record_kernel<<<1, 32, 0, ctx->stream>>>(st->part, nb, B, st->info);
  CRUX_LAUNCHED(ctx);
  rc = mlp_backward(st->q1, cat, B, dq1, false, false, true, nullptr); if (rc) return rc;
  rc = mlp_backward(st->q2, cat, B, dq2, false, false, true, nullptr); if (rc) return rc;
  rc = grads_allreduce(ctx, st->q1->grads, st->q1->n_params); if (rc) return rc;
  rc = grads_allreduce(ctx, st->q2->grads, st->q2->n_params); if (rc) return rc;
  {
    AdamSegs segs;
    segs.n = 2;
    segs.s[0] = AdamSeg{st->q1->params, st->q1->grads, st->q1->m, st->q1->v, st->q1->n_params};
    segs.s[1] = AdamSeg{st->q2->params, st->q2->grads, st->q2->m, st->q2->v, st->q2->n_params};
    rc = adam_step_segments(ctx, segs, st->q1->eta, st->q1->beta1, st->q1->beta2, st->q1->eps, st->q1->step_dev, st->info + 2, nullptr,
                            st->q1->norm_part);
    if (rc) return rc;
  }

  // 4. actor step (off_policy.jl:96-98, sac_actor_loss rl/sac.jl:34-40): fresh noise, updated critics and temperature
  sac_sample_kernel<<<rb, 128, 0, ctx->stream>>>(net->act[La], A, pol->ascale, eps_actor, seed, ctr + 2, B, ap, logp, epsb);
  CRUX_LAUNCHED(ctx);
  concat2_kernel<<<eb, 256, 0, ctx->stream>>>(s, sdim, ap, A, B, cat);
  CRUX_LAUNCHED(ctx);
  rc = mlp_forward_keep(st->q1, cat, B, nullptr); if (rc) return rc;
  rc = mlp_forward_keep(st->q2, cat, B, nullptr); if (rc) return rc;
  sac_actor_seed_kernel<<<nb, 256, 0, ctx->stream>>>(st->q1->act[st->q1->n_layers], st->q2->act[st->q2->n_layers], logp, st->log_alpha, B,
                                                    inv_b, dq1, dq2, st->part);
  CRUX_LAUNCHED(ctx);
  sac_actor_record_kernel<<<1, 32, 0, ctx->stream>>>(st->part, nb, B, st->info);
  CRUX_LAUNCHED(ctx);
  rc = mlp_backward(st->q1, cat, B, dq1, true, false, false, nullptr); if (rc) return rc;
  rc = mlp_backward(st->q2, cat, B, dq2, true, false, false, nullptr); if (rc) return rc;
  (void)dcat1; (void)dcat2;
  sac_actor_bwd_head_kernel<<<rb, 128, 0, ctx->stream>>>(net->act[La], epsb, st->q1->dz[0], st->q2->dz[0], sdim, A, pol->ascale,
                                                        st->log_alpha, B, inv_b, net->dz[La]);
  CRUX_LAUNCHED(ctx);
  rc = mlp_backward(net, s, B, net->dz[La], false, false, true, nullptr); if (rc) return rc;
  rc = grads_allreduce(ctx, net->grads, net->n_params); if (rc) return rc;
  rc = mlp_adam_step(net, st->info + 4, nullptr); if (rc) return rc;

  // 5. target_update (off_policy.jl:100): polyak τ on the critics (the actor copy inside π⁻ is never read)
  rc = crux_mlp_polyak(st->q1t, st->q1, st->tau); if (rc) return rc;
  rc = crux_mlp_polyak(st->q2t, st->q2, st->tau); if (rc) return rc;

  if (info_out_host) {
    CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(info_out_host, st->info, 8 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    return crux_ctx_check(ctx);
  }
  return CRUX_OK;
}

static int fill_bounds(crux_ctx *ctx, NoiseBounds &nb, const float *a_min, int32_t n_min, const float *a_max, int32_t n_max, int A) {
  if ((n_min != 0 && n_min != 1 && n_min != A) || (n_max != 0 && n_max != 1 && n_max != A) || A > 32)
    return crux_set_err(ctx, CRUX_ERR_INVALID, "noise bounds: a_min / a_max need 0, 1 or adim entries (adim <= 32)");
  nb.n_lo = n_min ? n_min : 1; nb.n_hi = n_max ? n_max : 1;
  nb.lo[0] = -INFINITY; nb.hi[0] = INFINITY;
  for (int j = 0; j < n_min; ++j) nb.lo[j] = a_min[j];
  for (int j = 0; j < n_max; ++j) nb.hi[j] = a_max[j];
  return CRUX_OK;
}

int32_t crux_noise_explore(crux_ctx *ctx, float *a, int64_t B, int32_t A, float sigma, float eps_min, float eps_max, const float *a_min,
                           int32_t n_min, const float *a_max, int32_t n_max, const float *eps_in, uint64_t seed, uint64_t ctr) {
  if (!ctx) return CRUX_ERR_INVALID;
  CRUX_REQUIRE(ctx, a && B >= 0 && A >= 1, "crux_noise_explore: bad arguments");
  NoiseBounds nb;
  int rc = fill_bounds(ctx, nb, a_min, n_min, a_max, n_max, A); if (rc) return rc;
  if (B == 0) return CRUX_OK;
  noise_clamp_kernel<<<(unsigned)cdiv(B, 128), 128, 0, ctx->stream>>>(a, B, A, sigma, eps_min, eps_max, nb, eps_in, seed, ctr);
  CRUX_LAUNCHED(ctx);
  return CRUX_OK;
}

int32_t crux_ddpg_create(crux_mlp *actor, crux_mlp *actor_target, crux_mlp *q1, crux_mlp *q1_target, crux_mlp *q2, crux_mlp *q2_target,
                         float tau, crux_ddpg_state **out) {
  if (!actor) return CRUX_ERR_INVALID;
  crux_ctx *ctx = actor->ctx;
  CRUX_REQUIRE(ctx, actor_target && q1 && q1_target && out && ((q2 == nullptr) == (q2_target == nullptr)), "crux_ddpg_create: NULL argument");
  const int sdim = actor->dims[0], A = actor->dims[actor->n_layers];
  CRUX_REQUIRE(ctx, q1->dims[0] == sdim + A && q1->dims[q1->n_layers] == 1 && (!q2 || (q2->dims[0] == sdim + A && q2->dims[q2->n_layers] == 1)),
               "crux_ddpg_create: critics must map vcat(s,a) -> 1");
  CRUX_REQUIRE(ctx, actor->n_params == actor_target->n_params && q1->n_params == q1_target->n_params && (!q2 || q2->n_params == q2_target->n_params),
               "crux_ddpg_create: target shape mismatch");
  crux_ddpg_state *st = new crux_ddpg_state();
  st->ctx = ctx; st->actor = actor; st->actor_t = actor_target; st->q1 = q1; st->q1t = q1_target; st->q2 = q2; st->q2t = q2_target; st->tau = tau;
  if (cudaMalloc((void **)&st->info, 8 * sizeof(float)) != cudaSuccess || cudaMalloc((void **)&st->part, 4096 * sizeof(double)) != cudaSuccess) {
    crux_ddpg_destroy(st);
    return crux_set_err(ctx, CRUX_ERR_OOM, "crux_ddpg_create: cudaMalloc");
  }
  cudaMemsetAsync(st->info, 0, 8 * sizeof(float), ctx->stream);
  cudaStreamSynchronize(ctx->stream);
  *out = st;
  return CRUX_OK;
}

int32_t crux_ddpg_destroy(crux_ddpg_state *st) {
  if (!st) return CRUX_OK;
  cudaStreamSynchronize(st->ctx->stream);
  if (st->info) cudaFree(st->info);
  if (st->ws) cudaFree(st->ws);
  if (st->part) cudaFree(st->part);
  delete st;
  return CRUX_OK;
}

int32_t crux_ddpg_train(crux_ddpg_state *st, const float *s, const float *a, const float *sp, const float *r, const uint8_t *done, int64_t B,
                        float gamma, int32_t smooth, float sigma, float eps_min, float eps_max, const float *a_min, int32_t n_min,
                        const float *a_max, int32_t n_max, const float *eps_smooth, uint64_t seed, uint64_t ctr, int32_t train_critic,
                        int32_t train_actor, float *y_out, float *info_out_host) {
  if (!st) return CRUX_ERR_INVALID;
  crux_ctx *ctx = st->ctx;
  CRUX_REQUIRE(ctx, B >= 1 && s && a && sp && r && done, "crux_ddpg_train: bad arguments");
  // several ranks: like crux_sac_train -- means over world * B rows, gradients summed over ranks before every optimiser step
  const int world = ctx->world;
  if (world > 1) seed ^= 0x9E3779B97F4A7C15ULL * (uint64_t)(ctx->rank + 1);
  crux_mlp *act = st->actor;
  const int sdim = act->dims[0], La = act->n_layers, A = act->dims[La], ld = sdim + A;
  NoiseBounds nb;
  int rc = fill_bounds(ctx, nb, a_min, n_min, a_max, n_max, A); if (rc) return rc;
  // workspace: ap[B][A], cat[B][ld], q1v[B], q2v[B], y[B], dq1[B], dq2[B], da[B][A]
  const size_t nfl = (size_t)B * (A + ld + 5 + A);
  if (st->ws_bytes < nfl * sizeof(float)) {
    CRUX_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (st->ws) cudaFree(st->ws);
    st->ws = nullptr; st->ws_bytes = 0;
    if (cudaMalloc((void **)&st->ws, nfl * sizeof(float) * 5 / 4) != cudaSuccess) return crux_set_err(ctx, CRUX_ERR_OOM, "crux_ddpg_train: workspace");
    st->ws_bytes = nfl * sizeof(float) * 5 / 4;
  }
  float *ap = st->ws, *cat = ap + (size_t)B * A, *q1v = cat + (size_t)B * ld, *q2v = q1v + B, *y = q2v + B, *dq1 = y + B, *dq2 = dq1 + B,
        *da = dq2 + B;
  const unsigned rb = (unsigned)cdiv(B, 128);
  const unsigned eb = (unsigned)i64min(cdiv(B * ld, 256), (int64_t)ctx->num_sms * 8);
  const int nb_ = (int)i64min(cdiv(B, 256), 256);
  const float inv_b = 1.0f / ((float)B * (float)world);

  // 1. target (off_policy.jl:80): a' = action(π⁻, sp) [+ clipped smoothing noise, then the action clamp], y from the TARGET critic(s)
  rc = mlp_forward_out(st->actor_t, sp, B, ap); if (rc) return rc;
  if (smooth) {
    noise_clamp_kernel<<<rb, 128, 0, ctx->stream>>>(ap, B, A, sigma, eps_min, eps_max, nb, eps_smooth, seed, ctr);
    CRUX_LAUNCHED(ctx);
  }
  concat2_kernel<<<eb, 256, 0, ctx->stream>>>(sp, sdim, ap, A, B, cat);
  CRUX_LAUNCHED(ctx);
  rc = mlp_forward_out(st->q1t, cat, B, q1v); if (rc) return rc;
  if (st->q2) { rc = mlp_forward_out(st->q2t, cat, B, q2v); if (rc) return rc; }
  ddpg_target_kernel<<<(unsigned)cdiv(B, 256), 256, 0, ctx->stream>>>(r, done, q1v, st->q2 ? q2v : nullptr, B, gamma, y);
  CRUX_LAUNCHED(ctx);
  if (y_out) CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(y_out, y, (size_t)B * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));

  // 2. critic step (off_policy.jl:91-93): td_loss on one critic (DDPG) or double_Q_loss with one optimiser over (Q1, Q2) (TD3)
  if (train_critic) {
    concat2_kernel<<<eb, 256, 0, ctx->stream>>>(s, sdim, a, A, B, cat);
    CRUX_LAUNCHED(ctx);
    rc = mlp_forward_keep(st->q1, cat, B, nullptr); if (rc) return rc;
    if (st->q2) {
      rc = mlp_forward_keep(st->q2, cat, B, nullptr); if (rc) return rc;
      sac_critic_head_kernel<<<nb_, 256, 0, ctx->stream>>>(st->q1->act[st->q1->n_layers], st->q2->act[st->q2->n_layers], y, B, inv_b, dq1, dq2,
                                                          st->part);
      CRUX_LAUNCHED(ctx);
      sac_critic_record_kernel<<<1, 32, 0, ctx->stream>>>(st->part, nb_, B, st->info);
      CRUX_LAUNCHED(ctx);
      rc = mlp_backward(st->q1, cat, B, dq1, false, false, true, nullptr); if (rc) return rc;
      rc = mlp_backward(st->q2, cat, B, dq2, false, false, true, nullptr); if (rc) return rc;
      rc = grads_allreduce(ctx, st->q1->grads, st->q1->n_params); if (rc) return rc;
      rc = grads_allreduce(ctx, st->q2->grads, st->q2->n_params); if (rc) return rc;
      AdamSegs segs;
      segs.n = 2;
      segs.s[0] = AdamSeg{st->q1->params, st->q1->grads, st->q1->m, st->q1->v, st->q1->n_params};
      segs.s[1] = AdamSeg{st->q2->params, st->q2->grads, st->q2->m, st->q2->v, st->q2->n_params};
      rc = adam_step_segments(ctx, segs, st->q1->eta, st->q1->beta1, st->q1->beta2, st->q1->eps, st->q1->step_dev, st->info + 2, nullptr,
                              st->q1->norm_part);
      if (rc) return rc;
    } else {
      q_mse_head_kernel<<<nb_, 256, 0, ctx->stream>>>(st->q1->act[st->q1->n_layers], y, B, inv_b, dq1, st->part);
      CRUX_LAUNCHED(ctx);
      q_mse_record_kernel<<<1, 32, 0, ctx->stream>>>(st->part, nb_, B, st->info);
      CRUX_LAUNCHED(ctx);
      rc = mlp_backward(st->q1, cat, B, dq1, false, false, true, nullptr); if (rc) return rc;
      rc = grads_allreduce(ctx, st->q1->grads, st->q1->n_params); if (rc) return rc;
      rc = mlp_adam_step(st->q1, st->info + 2, nullptr); if (rc) return rc;
    }
  }

  // 3. actor step (off_policy.jl:96-101): -mean(Q1(s, μ(s))) through the (already updated, frozen) first critic, then the target update
  if (train_actor) {
    rc = mlp_forward_keep(act, s, B, nullptr); if (rc) return rc;
    concat2_kernel<<<eb, 256, 0, ctx->stream>>>(s, sdim, act->act[La], A, B, cat);
    CRUX_LAUNCHED(ctx);
    rc = mlp_forward_keep(st->q1, cat, B, nullptr); if (rc) return rc;
    ddpg_actor_seed_kernel<<<nb_, 256, 0, ctx->stream>>>(st->q1->act[st->q1->n_layers], B, inv_b, dq1, st->part);
    CRUX_LAUNCHED(ctx);
    ddpg_actor_record_kernel<<<1, 32, 0, ctx->stream>>>(st->part, nb_, B, st->info);
    CRUX_LAUNCHED(ctx);
    rc = mlp_backward(st->q1, cat, B, dq1, true, false, false, nullptr); if (rc) return rc;
    slice_action_grad_kernel<<<(unsigned)i64min(cdiv(B * A, 256), (int64_t)ctx->num_sms * 8), 256, 0, ctx->stream>>>(st->q1->dz[0], sdim, A, B, da);
    CRUX_LAUNCHED(ctx);
    rc = mlp_backward(act, s, B, da, false, false, true, nullptr); if (rc) return rc;
    rc = grads_allreduce(ctx, act->grads, act->n_params); if (rc) return rc;
    rc = mlp_adam_step(act, st->info + 4, nullptr); if (rc) return rc;
    // target_update(π⁻, π) = polyak_average!(π⁻, π, τ) over every parameter of the policy (off_policy.jl:55,100)
    rc = crux_mlp_polyak(st->actor_t, act, st->tau); if (rc) return rc;
    rc = crux_mlp_polyak(st->q1t, st->q1, st->tau); if (rc) return rc;
    if (st->q2) { rc = crux_mlp_polyak(st->q2t, st->q2, st->tau); if (rc) return rc; }
  }

  if (info_out_host) {
    CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(info_out_host, st->info, 8 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    return crux_ctx_check(ctx);
  }
  return CRUX_OK;
}

}  // extern "C"
