// Hot path (iii): policy_gradient_training (src/model_free/on_policy.jl:56-78) =
// batch_train!(actor, ppo_loss | a2c_loss) then batch_train!(critic, mse) (src/training.jl:28-55),
// as one stream-ordered launch sequence with no host synchronisation between minibatches:
// the KL early stop (rl/ppo.jl:59) is evaluated on the device and later kernels become no-ops.
//
// `shuffle!` (experience_buffer.jl:118-124) never moves data here: each epoch is an index order
// consumed by the gather of the minibatch.
#include "policy.cuh"

namespace {

#define LOG_SQRT_2PI 0.9189385332046727f
#define ENT_CONST 1.4189385332046727f

// ---- device-side random permutation: 4-round Feistel network with cycle walking ------------------
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
// blockIdx.y = epoch offset: the row orders of several epochs are written behind each other ([epochs][n]) by one launch
__global__ void perm_fill_kernel(int32_t *__restrict__ out, int64_t n, int half_bits, uint64_t seed, uint32_t epoch) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out += (int64_t)blockIdx.y * n;
  epoch += blockIdx.y;
  const uint32_t mask = (1u << half_bits) - 1u;
  uint64_t x = (uint64_t)i;
  do {
    uint32_t l = (uint32_t)(x >> half_bits) & mask, r = (uint32_t)x & mask;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t f = mix32(r ^ mix32((uint32_t)seed + 0x9E3779B9u * (k + 1)) ^ mix32((uint32_t)(seed >> 32) + epoch * 0x85EBCA6Bu + k)) & mask;
      const uint32_t nl = r; r = l ^ f; l = nl;
    }
    x = ((uint64_t)l << half_bits) | r;
  } while (x >= (uint64_t)n);
  out[i] = (int32_t)x;
}

// ctl[0] = skip, ctl[1] = pending stop.  One thread, first kernel of every actor minibatch.
__global__ void advance_kernel(int *__restrict__ ctl, float *__restrict__ rec) {
  if (ctl[1]) ctl[0] = 1;
  rec[CRUX_PPO_VALID] = ctl[0] ? 0.f : 1.f;
}
__global__ void reset_ctl_kernel(int *__restrict__ ctl) { ctl[0] = 0; ctl[1] = 0; }

// gather a minibatch: dst columns [bm][d] <- src[idx[i]][d]
struct GatherCols { const float *src[6]; float *dst[6]; int dim[6]; int n; };
__global__ void gather_cols_kernel(GatherCols g, const int32_t *__restrict__ idx, int64_t bm, const int *__restrict__ skip) {
  if (skip && *skip) return;
  const int c = blockIdx.y;
  const int d = g.dim[c];
  const float *src = g.src[c];
  float *dst = g.dst[c];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < bm * d; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / d; const int col = (int)(i - row * d);
    dst[i] = src[(int64_t)idx[row] * d + col];
  }
}

// ppo_loss / a2c_loss head (rl/ppo.jl:4-21, rl/a2c.jl:4-16) for a state-independent-logΣ GaussianPolicy.
// One thread per sample.  Writes dL/dmu and block partial sums:
//  part[b][0..7] = sum surr|logp*A, sum (old-new), clip count, sum adv, sum ret, 0, sum cost surrogate, 0 ; part[b][8+j] = d/dlogΣ_j
#define HEAD_STRIDE (8 + CRUX_MAX_ADIM)
__global__ void __launch_bounds__(128)
ppo_head_kernel(const float *__restrict__ mu, const float *__restrict__ a, const float *__restrict__ old_logp,
                const float *__restrict__ adv, const float *__restrict__ ret, const float *__restrict__ ls, int A,
                int64_t bm, float inv_bg, float eps_clip, float lambda_p, int a2c, float *__restrict__ dmu,
                double *__restrict__ part, const int *__restrict__ skip, const float *__restrict__ cadv,
                const float *__restrict__ penalty) {
  if (skip && *skip) return;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float s_obj = 0.f, s_kl = 0.f, s_clip = 0.f, s_adv = 0.f, s_ret = 0.f, s_cobj = 0.f;
  // lagrange_ppo_loss rl/ppo.jl:70-131: (λp·p_loss + λe·e_loss + penalty·mean(max(r·Ac, clamp(r)·Ac))) / (1 + penalty)
  const float pen = penalty ? penalty[0] : 0.f;
  const float lscale = 1.f / (1.f + pen);
  float dls[CRUX_MAX_ADIM];
#pragma unroll 8
  for (int j = 0; j < CRUX_MAX_ADIM; ++j) dls[j] = 0.f;
  if (i < bm) {
    float logp = 0.f;
    for (int j = 0; j < A; ++j) {
      const float sg = expf(ls[j]);
      const float d = a[i * A + j] - mu[i * A + j];
      logp += -(d * d) / (2.f * (sg * sg)) - LOG_SQRT_2PI - ls[j];
    }
    const float Ai = adv[i], old = old_logp[i];
    float dlogp;
    if (a2c) {
      s_obj = logp * Ai;
      dlogp = -lambda_p * inv_bg * Ai;
    } else {
      const float r = expf(logp - old);
      const float lo = 1.f - eps_clip, hi = 1.f + eps_clip;
      const float x = r * Ai, y = fminf(fmaxf(r, lo), hi) * Ai;
      const bool first = !(y < x);  // min(x, y) keeps x on ties (Base.min)
      s_obj = first ? x : y;
      dlogp = first ? -lambda_p * inv_bg * x : 0.f;  // the clamped branch is only taken outside [lo, hi]: zero slope
      s_clip = (r > hi || r < lo) ? 1.f : 0.f;
      if (cadv) {
        const float Ac = cadv[i];
        const float xc = r * Ac, yc = fminf(fmaxf(r, lo), hi) * Ac;
        const bool takex = xc > yc;                 // max(x, y): the clamped branch on ties (same slope inside [lo, hi])
        s_cobj = takex ? xc : yc;
        const bool inside = r >= lo && r <= hi;
        dlogp += pen * inv_bg * ((takex || inside) ? xc : 0.f);
        dlogp *= lscale;
      }
    }
    s_kl = old - logp; s_adv = Ai; s_ret = ret ? ret[i] : 0.f;
    for (int j = 0; j < A; ++j) {
      const float sg = expf(ls[j]);
      const float var = sg * sg;
      const float d = a[i * A + j] - mu[i * A + j];
      dmu[i * A + j] = dlogp * d / var;
      dls[j] = dlogp * (d * d / var - 1.f);
    }
  }
  // block reduction (128 threads = 4 warps) in double
  __shared__ double sh[4][HEAD_STRIDE];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double v;
  v = warp_sum_d((double)s_obj); if (lane == 0) sh[w][0] = v;
  v = warp_sum_d((double)s_kl); if (lane == 0) sh[w][1] = v;
  v = warp_sum_d((double)s_clip); if (lane == 0) sh[w][2] = v;
  v = warp_sum_d((double)s_adv); if (lane == 0) sh[w][3] = v;
  v = warp_sum_d((double)s_ret); if (lane == 0) sh[w][4] = v;
  v = warp_sum_d((double)s_cobj); if (lane == 0) sh[w][6] = v;
  for (int j = 0; j < A; ++j) { v = warp_sum_d((double)dls[j]); if (lane == 0) sh[w][8 + j] = v; }
  __syncthreads();
  for (int k = threadIdx.x; k < 8 + A; k += blockDim.x) {
    if (k == 5 || k == 7) { part[(int64_t)blockIdx.x * HEAD_STRIDE + k] = 0.0; continue; }
    part[(int64_t)blockIdx.x * HEAD_STRIDE + k] = sh[0][k] + sh[1][k] + sh[2][k] + sh[3][k];
  }
}

// The same losses for a DiscreteNetwork actor (policies.jl:104-157): p = softmax(net(s)) (logits, :110,133), logpdf = categorical_logpdf
// (:135: log(sum(p .* a_onehot))), entropy = -sum(p .* log.(p .+ eps(Float32))) per sample (:152-155), e_loss = -mean(entropy).
// One thread per sample; writes dL/dz for the nA network outputs.  part[b][7] = sum of the per-sample entropies.
__global__ void __launch_bounds__(128)
ppo_head_cat_kernel(const float *__restrict__ z, const float *__restrict__ a, const float *__restrict__ old_logp,
                    const float *__restrict__ adv, const float *__restrict__ ret, int nA, int64_t bm, float inv_bg, float eps_clip,
                    float lambda_p, float lambda_e, int a2c, float *__restrict__ dz, double *__restrict__ part, const int *__restrict__ skip) {
  if (skip && *skip) return;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float s_obj = 0.f, s_kl = 0.f, s_clip = 0.f, s_adv = 0.f, s_ret = 0.f, s_ent = 0.f;
  if (i < bm) {
    const float *zi = z + i * nA, *ai = a + i * nA;
    float m = zi[0];
    for (int j = 1; j < nA; ++j) m = fmaxf(m, zi[j]);
    float S = 0.f;
    for (int j = 0; j < nA; ++j) S += expf(zi[j] - m);
    float q = 0.f, H = 0.f, gp = 0.f;   // q = sum(p a); gp = sum_i g_i p_i with g_i = dH/dp_i
    for (int j = 0; j < nA; ++j) {
      const float p = expf(zi[j] - m) / S;
      const float lg = logf(p + 1.1920929e-07f);
      q += p * ai[j];
      H -= p * lg;
      gp += -(lg + p / (p + 1.1920929e-07f)) * p;
    }
    const float logp = logf(q);
    const float Ai = adv[i], old = old_logp[i];
    float dlogp;
    if (a2c) {
      s_obj = logp * Ai;
      dlogp = -lambda_p * inv_bg * Ai;
    } else {
      const float r = expf(logp - old);
      const float lo = 1.f - eps_clip, hi = 1.f + eps_clip;
      const float x = r * Ai, y = fminf(fmaxf(r, lo), hi) * Ai;
      const bool first = !(y < x);  // min(x, y) keeps x on ties (Base.min)
      s_obj = first ? x : y;
      dlogp = first ? -lambda_p * inv_bg * x : 0.f;
      s_clip = (r > hi || r < lo) ? 1.f : 0.f;
    }
    s_kl = old - logp; s_adv = Ai; s_ret = ret ? ret[i] : 0.f; s_ent = H;
    const float de = -lambda_e * inv_bg;   // d(λe e_loss)/dH_i
    for (int k = 0; k < nA; ++k) {
      const float p = expf(zi[k] - m) / S;
      const float g = -(logf(p + 1.1920929e-07f) + p / (p + 1.1920929e-07f));
      dz[i * nA + k] = dlogp * p * (ai[k] - q) / q + de * p * (g - gp);
    }
  }
  __shared__ double sh[4][8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double v;
  v = warp_sum_d((double)s_obj); if (lane == 0) sh[w][0] = v;
  v = warp_sum_d((double)s_kl); if (lane == 0) sh[w][1] = v;
  v = warp_sum_d((double)s_clip); if (lane == 0) sh[w][2] = v;
  v = warp_sum_d((double)s_adv); if (lane == 0) sh[w][3] = v;
  v = warp_sum_d((double)s_ret); if (lane == 0) sh[w][4] = v;
  v = warp_sum_d((double)s_ent); if (lane == 0) sh[w][7] = v;
  __syncthreads();
  if (threadIdx.x < 8) {
    const int k = threadIdx.x;
    part[(int64_t)blockIdx.x * HEAD_STRIDE + k] = (k == 5 || k == 6) ? 0.0 : sh[0][k] + sh[1][k] + sh[2][k] + sh[3][k];
  }
}

// one block: reduce head partials -> gradient tail (logΣ gradient, sums).  sums: [obj, kl, clip, adv, ret, count]
// launched with 8 x 32 threads: lane = entry k, warp w adds blocks w, w + 8, ... (eight loads in flight per lane), the eight warp sums are
// combined in warp order: fixed order, bit-reproducible.  (One thread per entry walking every block in turn took 20 us for 256 blocks.)
__global__ void ppo_finalize_kernel(const double *__restrict__ part, int nblocks, int A, int64_t bm, float lambda_e,
                                    float *__restrict__ ls_grad, float *__restrict__ sums, const int *__restrict__ skip) {
  if (skip && *skip) return;
  __shared__ double sh[8][32];
  const int k = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int k0 = 0; k0 < 8 + A; k0 += 32) {   // (A <= 64: at most three rounds)
    const int kk = k0 + k;
    double s = 0.0;
    if (kk < 8 + A) {
#pragma unroll 8
      for (int b = w; b < nblocks; b += 8) s += part[(int64_t)b * HEAD_STRIDE + kk];
    }
    sh[w][k] = s;
    __syncthreads();
    if (w == 0 && kk < 8 + A) {
      s = 0.0;
#pragma unroll
      for (int q = 0; q < 8; ++q) s += sh[q][k];
      if (kk < 5 || kk == 6 || kk == 7) sums[kk] = (float)s;   // [7]: sum of per-sample entropies (categorical actor; 0 otherwise)
      else if (kk == 5) sums[5] = (float)bm;
      else ls_grad[kk - 8] = (float)s;  // entropy term added after the all-reduce (record kernel)
    }
    __syncthreads();
  }
  (void)lambda_e;
}

// after the (optional) all-reduce: info record, entropy gradient, early-stop vote.  One thread.
__global__ void ppo_record_kernel(const float *__restrict__ sums, float *__restrict__ ls_grad, const float *__restrict__ ls,
                                  int A, float lambda_p, float lambda_e, float target_kl, int a2c, int world,
                                  float *__restrict__ rec, int *__restrict__ ctl, const float *__restrict__ penalty,
                                  float *__restrict__ lrec, int categorical) {
  if (ctl[0]) return;
  const float cnt = sums[5];
  float sls = 0.f;
  for (int j = 0; j < A; ++j) sls += ls[j];
  const float entropy = categorical ? sums[7] / cnt : ENT_CONST + sls;     // policies.jl:152-155 (mean over the minibatch) / :348
  const float p_loss = -(sums[0] / cnt);
  const float e_loss = -entropy;
  if (penalty) {                             // lagrange_ppo_loss rl/ppo.jl:108-131
    const float pen = penalty[0];
    const float cost_loss = pen * (sums[6] / cnt);
    rec[CRUX_PPO_LOSS] = (lambda_p * p_loss + lambda_e * e_loss + cost_loss) / (1.f + pen);
    lrec[5] = lambda_p * p_loss; lrec[6] = cost_loss;
    lambda_e = lambda_e / (1.f + pen);       // the entropy term's share of the gradient below
  } else
  rec[CRUX_PPO_LOSS] = lambda_p * p_loss + lambda_e * e_loss;
  rec[CRUX_PPO_ENTROPY] = entropy;
  const float kl = sums[1] / cnt;
  rec[CRUX_PPO_KL] = kl;
  rec[CRUX_PPO_CLIP_FRAC] = a2c ? 0.f : sums[2] / cnt;
  rec[CRUX_PPO_AVG_ADV] = sums[3] / cnt;
  rec[CRUX_PPO_AVG_RET] = sums[4] / cnt;
  // d(λe * e_loss)/dlogΣ_j = -λe on every rank's replica: not summed over ranks
  for (int j = 0; j < A; ++j) ls_grad[j] += -lambda_e;
  if (kl > target_kl) ctl[1] = 1;  // rl/ppo.jl:59 checked after this minibatch's update (training.jl:46)
  (void)world;
}

// The PID penalty update inside lagrange_ppo_loss (rl/ppo.jl:79-106), once per loss evaluation = once per minibatch:
//   Jc = sum(cost) / sum(episode_end) over the minibatch rows; Δ = Jc - target; I = clamp(I + Ki Δ, 0, Ki_max);
//   smooth_Δ, smooth_Jc: EMA with the Float64 α rounded to the Float32 state; ∂ = max(0, smooth_Jc - Jc_prev);
//   penalty = clamp(Kp smooth_Δ + I + Kd ∂, 0, penalty_max).   state = {I, smooth_Δ, smooth_Jc, Jc_prev}, one block.
struct LagrangePid { float target_cost, penalty_max, Ki_max, Ki, Kp, Kd; double ema_alpha; };
// the PID step itself (rl/ppo.jl:79-106) from the minibatch sums c = sum(cost), e = sum(episode_end): one thread
__device__ __forceinline__ void lagrange_pid_step(float c, float e, const LagrangePid &h, float *__restrict__ state, float *__restrict__ penalty,
                                                  float *__restrict__ lrec) {
  const float Jc = c / e;
  const float d = Jc - h.target_cost;
  const float I = fminf(fmaxf(state[0] + h.Ki * d, 0.f), h.Ki_max);
  const float sd = (float)(h.ema_alpha * (double)state[1] + (1.0 - h.ema_alpha) * (double)d);
  const float sj = (float)(h.ema_alpha * (double)state[2] + (1.0 - h.ema_alpha) * (double)Jc);
  const float der = fmaxf(0.f, sj - state[3]);
  state[0] = I; state[1] = sd; state[2] = sj; state[3] = sj;
  const float pen = fminf(fmaxf(h.Kp * sd + I + h.Kd * der, 0.f), h.penalty_max);
  penalty[0] = pen;
  lrec[0] = pen; lrec[1] = Jc; lrec[2] = h.Kp * sd; lrec[3] = der; lrec[4] = I; lrec[7] = 1.f;
}
// One rank: sums and PID step in one launch.  Several ranks (sums_out != NULL): this rank's sums are written out, all-reduced like the
// gradient, and lagrange_pid_apply_kernel runs the identical PID step on every rank -- the cost estimate is that of the UNION minibatch.
__global__ void lagrange_pid_kernel(const float *__restrict__ cost, const uint8_t *__restrict__ episode_end, const int32_t *__restrict__ idx,
                                    int64_t bm, LagrangePid h, float *__restrict__ state, float *__restrict__ penalty,
                                    float *__restrict__ lrec, const int *__restrict__ skip, float *__restrict__ sums_out) {
  if (skip && *skip) { if (sums_out && threadIdx.x < 2) sums_out[threadIdx.x] = 0.f; return; }
  __shared__ double sh[2][32];
  double c = 0.0, e = 0.0;
  for (int64_t i = threadIdx.x; i < bm; i += blockDim.x) { const int64_t r = idx[i]; c += (double)cost[r]; e += (double)episode_end[r]; }
  c = warp_sum_d(c); e = warp_sum_d(e);
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = c; sh[1][threadIdx.x >> 5] = e; }
  __syncthreads();
  if (threadIdx.x == 0) {
    c = 0.0; e = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { c += sh[0][w]; e += sh[1][w]; }
    if (sums_out) { sums_out[0] = (float)c; sums_out[1] = (float)e; }
    else lagrange_pid_step((float)c, (float)e, h, state, penalty, lrec);
  }
}
__global__ void lagrange_pid_apply_kernel(const float *__restrict__ sums, LagrangePid h, float *__restrict__ state, float *__restrict__ penalty,
                                          float *__restrict__ lrec, const int *__restrict__ skip) {
  if (skip && *skip) return;
  lagrange_pid_step(sums[0], sums[1], h, state, penalty, lrec);
}

// critic: mse head, dz = 2 (V - R) / Bg, block partial sums of squared error
__global__ void critic_head_kernel(const float *__restrict__ v, const float *__restrict__ ret, int64_t bm, float inv_bg,
                                   float *__restrict__ dv, double *__restrict__ part) {
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < bm; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = v[i] - ret[i];
    s += (double)(d * d);
    dv[i] = 2.f * d * inv_bg;
  }
  __shared__ double sh[32];
  s = warp_sum_d(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
    s = warp_sum_d(s);
    if (threadIdx.x == 0) part[blockIdx.x] = s;
  }
}
__global__ void critic_finalize_kernel(const double *__restrict__ part, int n, int64_t bm, float *__restrict__ sums) {   // one warp
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 32) s += part[i];
  s = warp_sum_d(s);   // fixed shuffle tree: bit-reproducible
  if (threadIdx.x == 0) { sums[0] = (float)s; sums[5] = (float)bm; }
}
__global__ void critic_record_kernel(const float *__restrict__ sums, float *__restrict__ rec) {
  rec[CRUX_PPO_LOSS] = sums[0] / sums[5];
  rec[CRUX_PPO_VALID] = 1.f;
}

int ensure_bytes(crux_ctx *ctx, void **p, size_t *have, size_t need) {
  if (*have >= need) return CRUX_OK;
  CRUX_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (*p) cudaFree(*p);
  *p = nullptr; *have = 0;
  if (cudaMalloc(p, need + need / 4 + 256) != cudaSuccess) return crux_set_err(ctx, CRUX_ERR_OOM, "cudaMalloc(%zu)", need);
  *have = need + need / 4 + 256;
  return CRUX_OK;
}

}  // namespace

// helpers shared with the fused implementation (ppo_fused.cu)
int ppo_fill_order(crux_ctx *ctx, int32_t *out, int64_t n, uint64_t seed, uint32_t epoch) {
  int half_bits = 1;
  while ((1ll << (2 * half_bits)) < n) ++half_bits;
  perm_fill_kernel<<<(unsigned)cdiv(n, 256), 256, 0, ctx->stream>>>(out, n, half_bits, seed, epoch);
  CRUX_LAUNCHED(ctx);
  return CRUX_OK;
}
// the row orders of `epochs` consecutive epochs, [epochs][n], in ONE launch (epoch e of the result == ppo_fill_order(..., epoch0 + e))
int ppo_fill_orders(crux_ctx *ctx, int32_t *out, int64_t n, uint64_t seed, uint32_t epoch0, int epochs) {
  if (epochs <= 0) return CRUX_OK;
  int half_bits = 1;
  while ((1ll << (2 * half_bits)) < n) ++half_bits;
  perm_fill_kernel<<<dim3((unsigned)cdiv(n, 256), (unsigned)epochs), 256, 0, ctx->stream>>>(out, n, half_bits, seed, epoch0);
  CRUX_LAUNCHED(ctx);
  return CRUX_OK;
}
int ppo_ensure_bytes(crux_ctx *ctx, void **p, size_t *have, size_t need) { return ensure_bytes(ctx, p, have, need); }

// LagrangePPO (rl/ppo.jl:133-214): the cost critic Vc, the :cost / :cost_advantage / :cost_return columns and the PID state
struct LagrangeArgs {
  crux_mlp *cost_critic;
  const float *cost, *cost_adv, *cost_ret;
  const uint8_t *episode_end;
  LagrangePid pid;
  float *state;                 // device float[5]: I, smooth_Δ, smooth_Jc, Jc_prev, penalty (persists across updates)
  int cost_epochs; int64_t cost_batch, cost_max_batches;
  const int32_t *order_cost;
  float *info_l, *info_cost;    // device records, CRUX_PPO_INFO_STRIDE floats per minibatch
  float *sums;                  // device float[4] (16-byte aligned): this minibatch's {sum(cost), sum(episode_end)} for the all-reduce over ranks
};

// batch_train!(V, opt, 𝒫, 𝒟) with Flux.mse(value(V, s), target) (ppo.jl:60, :208): the critic and the cost critic
static int value_epochs(crux_gaussian *actor, crux_mlp *net, const float *s, const float *target, int64_t n, int epochs, int64_t batch,
                        int64_t max_batches, const int32_t *order_in, uint64_t seed, int half_bits, float *mb_s, float *mb_ret, float *info) {
  crux_ctx *ctx = actor->ctx;
  const int sdim = net->dims[0];
  const int64_t nmb = cdiv(n, batch);
  int64_t total = 0;
  const int64_t maxb = max_batches > 0 ? max_batches : INT64_MAX;
  int rc;
  for (int e = 0; e < epochs && total < maxb; ++e) {
    const int32_t *order;
    if (order_in) order = order_in + (int64_t)e * n;
    else {
      perm_fill_kernel<<<(unsigned)cdiv(n, 256), 256, 0, ctx->stream>>>(actor->order, n, half_bits, seed, (uint32_t)e);
      CRUX_LAUNCHED(ctx);
      order = actor->order;
    }
    const int Lc = net->n_layers;
    for (int64_t mbi = 0; mbi < nmb && total < maxb; ++mbi, ++total) {
      const int64_t off = mbi * batch, bm = i64min(batch, n - off);
      float *rec = info + ((int64_t)e * nmb + mbi) * CRUX_PPO_INFO_STRIDE;
      const float inv_bg = 1.0f / ((float)bm * (float)ctx->world);
      {
        GatherCols g;
        g.n = 2;
        g.src[0] = s; g.dst[0] = mb_s; g.dim[0] = sdim;
        g.src[1] = target; g.dst[1] = mb_ret; g.dim[1] = 1;
        dim3 gg((unsigned)i64min(cdiv(bm * sdim, 256), 1024), 2);
        gather_cols_kernel<<<gg, 256, 0, ctx->stream>>>(g, order + off, bm, nullptr);
        CRUX_LAUNCHED(ctx);
        rc = mlp_forward_keep(net, mb_s, bm, nullptr); if (rc) return rc;
        const int hb = (int)i64min(cdiv(bm, 256), 512);
        critic_head_kernel<<<hb, 256, 0, ctx->stream>>>(net->act[Lc], mb_ret, bm, inv_bg, net->dz[Lc], actor->partials);
        CRUX_LAUNCHED(ctx);
        critic_finalize_kernel<<<1, 32, 0, ctx->stream>>>(actor->partials, hb, bm, tail_sums(net));
        CRUX_LAUNCHED(ctx);
        rc = mlp_backward(net, mb_s, bm, net->dz[Lc], false, false, true, nullptr); if (rc) return rc;
      }
      if (ctx->world > 1) { rc = grads_allreduce(ctx, net->grads, net->n_params + CRUX_GRAD_TAIL); if (rc) return rc; }
      critic_record_kernel<<<1, 1, 0, ctx->stream>>>(tail_sums(net), rec);
      CRUX_LAUNCHED(ctx);
      rc = mlp_adam_step(net, rec + CRUX_PPO_GRAD_NORM, nullptr); if (rc) return rc;
    }
  }
  return CRUX_OK;
}

static int ppo_update_impl(crux_gaussian *actor, crux_mlp *critic, const float *s, const float *a, const float *logprob,
                           const float *advantage, const float *ret, int64_t n, const crux_ppo_hp *hp,
                           const int32_t *order_actor, const int32_t *order_critic, uint64_t seed, const LagrangeArgs *lg = nullptr) {
  crux_ctx *ctx = actor->ctx;
  CRUX_REQUIRE(ctx, hp, "crux_ppo_update: NULL hyper-parameters");
  CRUX_REQUIRE(ctx, n >= 1, "crux_ppo_update: empty buffer");
  CRUX_REQUIRE(ctx, s && a && logprob && advantage, "crux_ppo_update: NULL column");
  CRUX_REQUIRE(ctx, !actor->head_mode && !actor->squashed, "crux_ppo_update: needs GaussianPolicy with a logΣ vector (ppo.jl examples) or a categorical actor");
  const bool cat = actor->categorical;   // DiscreteNetwork actor: a = one-hot rows [n][nA], logprob = categorical_logpdf
  CRUX_REQUIRE(ctx, !(cat && lg), "crux_lagrange_ppo_update: Gaussian actors only");
  CRUX_REQUIRE(ctx, hp->actor_batch >= 1 && hp->actor_epochs >= 0, "crux_ppo_update: bad actor batch/epochs");
  CRUX_REQUIRE(ctx, n < (1ll << 31), "crux_ppo_update: n must fit int32 indices");
  crux_mlp *mu = actor->mu;
  const int sdim = mu->dims[0], A = actor->adim;
  const int64_t nmb_a = cdiv(n, hp->actor_batch);
  const int64_t nmb_c = (critic && hp->critic_epochs > 0) ? cdiv(n, hp->critic_batch) : 0;
  if (nmb_c) {
    CRUX_REQUIRE(ctx, ret, "crux_ppo_update: critic training needs the :return column");
    CRUX_REQUIRE(ctx, hp->critic_batch >= 1, "crux_ppo_update: bad critic batch");
    CRUX_REQUIRE(ctx, critic->dims[0] == sdim && critic->dims[critic->n_layers] == 1, "crux_ppo_update: critic shape");
  }
  int rc;
  if (!lg) {   // the fused kernels cover ppo_loss / a2c_loss / mse; lagrange_ppo_loss runs on the layer-by-layer engine below
    int handled = 0;
    rc = ppo_update_fused(actor, nmb_c ? critic : nullptr, s, a, logprob, advantage, ret, n, hp, order_actor, order_critic, seed, &handled);
    if (rc || handled) return rc;
  }
  // workspaces
  const int64_t bmax = i64max(i64max(i64min(n, hp->actor_batch), nmb_c ? i64min(n, hp->critic_batch) : 0),
                              (lg && lg->cost_epochs > 0) ? i64min(n, lg->cost_batch) : 0);
  const size_t mb_floats = (size_t)bmax * (sdim + A + 4);
  rc = ensure_bytes(ctx, (void **)&actor->mb, &actor->mb_bytes, mb_floats * sizeof(float)); if (rc) return rc;
  const size_t ia = (size_t)i64max(1, hp->actor_epochs * nmb_a) * CRUX_PPO_INFO_STRIDE * sizeof(float);
  const size_t ic = (size_t)i64max(1, (int64_t)hp->critic_epochs * nmb_c) * CRUX_PPO_INFO_STRIDE * sizeof(float);
  rc = ensure_bytes(ctx, (void **)&actor->info_actor, &actor->info_actor_bytes, ia); if (rc) return rc;
  rc = ensure_bytes(ctx, (void **)&actor->info_critic, &actor->info_critic_bytes, ic); if (rc) return rc;
  CRUX_CHECK_CUDA(ctx, cudaMemsetAsync(actor->info_actor, 0, ia, ctx->stream));
  CRUX_CHECK_CUDA(ctx, cudaMemsetAsync(actor->info_critic, 0, ic, ctx->stream));
  if (!order_actor || (nmb_c && !order_critic)) {
    rc = ensure_bytes(ctx, (void **)&actor->order, &actor->order_bytes, (size_t)n * sizeof(int32_t)); if (rc) return rc;
  }
  rc = mlp_ensure_workspace(mu, bmax); if (rc) return rc;
  if (nmb_c) { rc = mlp_ensure_workspace(critic, bmax); if (rc) return rc; }
  if (lg && lg->cost_epochs > 0) { rc = mlp_ensure_workspace(lg->cost_critic, bmax); if (rc) return rc; }
  int half_bits = 1;
  while ((1ll << (2 * half_bits)) < n) ++half_bits;

  float *mb_s = actor->mb, *mb_a = mb_s + (size_t)bmax * sdim, *mb_lp = mb_a + (size_t)bmax * A, *mb_adv = mb_lp + bmax,
        *mb_ret = mb_adv + bmax, *mb_cadv = mb_ret + bmax;
  int *skip = actor->ctl;
  reset_ctl_kernel<<<1, 1, 0, ctx->stream>>>(actor->ctl);
  CRUX_LAUNCHED(ctx);
  const int L = mu->n_layers;
  const float inv_world = 1.0f / (float)ctx->world;

  // ---------------- actor: batch_train!(actor(π), a_opt, 𝒫, 𝒟)  on_policy.jl:65
  int64_t total = 0;
  const int64_t maxb_a = hp->actor_max_batches > 0 ? hp->actor_max_batches : INT64_MAX;
  for (int e = 0; e < hp->actor_epochs && total < maxb_a; ++e) {
    const int32_t *order;
    if (order_actor) order = order_actor + (int64_t)e * n;
    else {
      perm_fill_kernel<<<(unsigned)cdiv(n, 256), 256, 0, ctx->stream>>>(actor->order, n, half_bits, seed, (uint32_t)e);
      CRUX_LAUNCHED(ctx);
      order = actor->order;
    }
    for (int64_t mbi = 0; mbi < nmb_a && total < maxb_a; ++mbi, ++total) {
      const int64_t off = mbi * hp->actor_batch, bm = i64min(hp->actor_batch, n - off);
      float *rec = actor->info_actor + ((int64_t)e * nmb_a + mbi) * CRUX_PPO_INFO_STRIDE;
      advance_kernel<<<1, 1, 0, ctx->stream>>>(actor->ctl, rec);
      CRUX_LAUNCHED(ctx);
      const float inv_bg = 1.0f / ((float)bm * (float)ctx->world);
      {
        GatherCols g;
        g.n = 5;
        g.src[0] = s; g.dst[0] = mb_s; g.dim[0] = sdim;
        g.src[1] = a; g.dst[1] = mb_a; g.dim[1] = A;
        g.src[2] = logprob; g.dst[2] = mb_lp; g.dim[2] = 1;
        g.src[3] = advantage; g.dst[3] = mb_adv; g.dim[3] = 1;
        g.src[4] = ret ? ret : advantage; g.dst[4] = mb_ret; g.dim[4] = 1;
        if (lg) { g.n = 6; g.src[5] = lg->cost_adv; g.dst[5] = mb_cadv; g.dim[5] = 1; }
        dim3 gg((unsigned)i64min(cdiv(bm * sdim, 256), 1024), g.n);
        gather_cols_kernel<<<gg, 256, 0, ctx->stream>>>(g, order + off, bm, skip);
        CRUX_LAUNCHED(ctx);
        if (lg) {   // the PID step of the loss evaluation (rl/ppo.jl:79-106) on this minibatch's rows
          float *lrec = lg->info_l + ((int64_t)e * nmb_a + mbi) * CRUX_PPO_INFO_STRIDE;
          lagrange_pid_kernel<<<1, 1024, 0, ctx->stream>>>(lg->cost, lg->episode_end, order + off, bm, lg->pid, lg->state, lg->state + 4, lrec, skip,
                                                         ctx->world > 1 ? lg->sums : nullptr);
          CRUX_LAUNCHED(ctx);
          if (ctx->world > 1) {
            rc = grads_allreduce(ctx, lg->sums, 4); if (rc) return rc;
            lagrange_pid_apply_kernel<<<1, 1, 0, ctx->stream>>>(lg->sums, lg->pid, lg->state, lg->state + 4, lrec, skip);
            CRUX_LAUNCHED(ctx);
          }
        }
        rc = mlp_forward_keep(mu, mb_s, bm, skip); if (rc) return rc;
        const int hb = (int)cdiv(bm, 128);
        // head partials live in scratch slot 4 (hb blocks x HEAD_STRIDE doubles)
        double *part = (double *)crux_scratch(ctx, 4, (size_t)hb * HEAD_STRIDE * sizeof(double));
        if (!part) return CRUX_ERR_OOM;
        if (cat)
          ppo_head_cat_kernel<<<hb, 128, 0, ctx->stream>>>(mu->act[L], mb_a, mb_lp, mb_adv, ret ? mb_ret : nullptr, A, bm, inv_bg, hp->eps_clip,
                                                           hp->lambda_p, hp->lambda_e, hp->a2c, mu->dz[L], part, skip);
        else
        ppo_head_kernel<<<hb, 128, 0, ctx->stream>>>(mu->act[L], mb_a, mb_lp, mb_adv, ret ? mb_ret : nullptr, actor->log_sigma, A, bm,
                                                     inv_bg, hp->eps_clip, hp->lambda_p, hp->a2c, mu->dz[L], part, skip,
                                                     lg ? mb_cadv : nullptr, lg ? lg->state + 4 : nullptr);
        CRUX_LAUNCHED(ctx);
        ppo_finalize_kernel<<<1, 256, 0, ctx->stream>>>(part, hb, cat ? 0 : A, bm, hp->lambda_e, tail_ls_grad(mu), tail_sums(mu), skip);
        CRUX_LAUNCHED(ctx);
        rc = mlp_backward(mu, mb_s, bm, mu->dz[L], false, false, true, skip); if (rc) return rc;
      }
      if (ctx->world > 1) { rc = grads_allreduce(ctx, mu->grads, mu->n_params + CRUX_GRAD_TAIL); if (rc) return rc; }
      ppo_record_kernel<<<1, 1, 0, ctx->stream>>>(tail_sums(mu), tail_ls_grad(mu), actor->log_sigma, cat ? 0 : A, hp->lambda_p, hp->lambda_e,
                                                  hp->target_kl, hp->a2c, ctx->world, rec, actor->ctl, lg ? lg->state + 4 : nullptr,
                                                  lg ? lg->info_l + ((int64_t)e * nmb_a + mbi) * CRUX_PPO_INFO_STRIDE : nullptr, cat ? 1 : 0);
      CRUX_LAUNCHED(ctx);
      AdamSegs segs;
      segs.n = cat ? 1 : 2;   // a categorical actor has no logΣ vector
      segs.s[0] = AdamSeg{mu->params, mu->grads, mu->m, mu->v, mu->n_params};
      segs.s[1] = AdamSeg{actor->log_sigma, tail_ls_grad(mu), actor->ls_m, actor->ls_v, (int64_t)A};
      rc = adam_step_segments(ctx, segs, mu->eta, mu->beta1, mu->beta2, mu->eps, mu->step_dev, rec + CRUX_PPO_GRAD_NORM, skip,
                              mu->norm_part);
      if (rc) return rc;
    }
  }
  (void)inv_world;

  // ---------------- critic: batch_train!(critic(π), c_opt, 𝒫, 𝒟)  on_policy.jl:68-70
  if (nmb_c) {
    rc = value_epochs(actor, critic, s, ret, n, hp->critic_epochs, hp->critic_batch, hp->critic_max_batches, order_critic,
                      seed ^ 0xC2B2AE3D27D4EB4FULL, half_bits, mb_s, mb_ret, actor->info_critic);
    if (rc) return rc;
  }
  // ---------------- cost critic: batch_train!(𝒮.Vc, cost_opt, 𝒫, 𝒟)  on_policy.jl:73-75 with mse(Vc(s), cost_return) (ppo.jl:208)
  if (lg && lg->cost_epochs > 0) {
    rc = value_epochs(actor, lg->cost_critic, s, lg->cost_ret, n, lg->cost_epochs, lg->cost_batch, lg->cost_max_batches, lg->order_cost,
                      seed ^ 0x9E3779B97F4A7C15ULL, half_bits, mb_s, mb_ret, lg->info_cost);
    if (rc) return rc;
  }
  return CRUX_OK;
}

extern "C" {

int32_t crux_ppo_update_async(crux_gaussian *actor, crux_mlp *critic, const float *s, const float *a, const float *logprob,
                              const float *advantage, const float *ret, int64_t n, const crux_ppo_hp *hp,
                              const int32_t *order_actor, const int32_t *order_critic, uint64_t seed) {
  if (!actor) return CRUX_ERR_INVALID;
  return ppo_update_impl(actor, critic, s, a, logprob, advantage, ret, n, hp, order_actor, order_critic, seed);
}

int32_t crux_lagrange_ppo_update(crux_gaussian *actor, crux_mlp *critic, crux_mlp *cost_critic, const float *s, const float *a,
                                 const float *logprob, const float *advantage, const float *ret, const float *cost,
                                 const float *cost_advantage, const float *cost_return, const uint8_t *episode_end, int64_t n,
                                 const crux_ppo_hp *hp, const crux_lagrange_hp *lhp, float *pid_state_dev, const int32_t *order_actor,
                                 const int32_t *order_critic, const int32_t *order_cost, uint64_t seed, float *info_actor_host,
                                 float *info_critic_host, float *info_lagrange_host, float *info_cost_host) {
  if (!actor) return CRUX_ERR_INVALID;
  crux_ctx *ctx = actor->ctx;
  CRUX_REQUIRE(ctx, hp && lhp && cost && cost_advantage && episode_end && pid_state_dev && n >= 1, "crux_lagrange_ppo_update: NULL argument");
  CRUX_REQUIRE(ctx, !hp->a2c, "crux_lagrange_ppo_update: the Lagrange loss extends ppo_loss");
  const int64_t nmb_a = cdiv(n, hp->actor_batch);
  const bool train_cost = cost_critic && lhp->cost_epochs > 0;
  if (train_cost) {
    CRUX_REQUIRE(ctx, cost_return && lhp->cost_batch >= 1, "crux_lagrange_ppo_update: cost critic training needs :cost_return and a batch size");
    CRUX_REQUIRE(ctx, cost_critic->dims[0] == actor->mu->dims[0] && cost_critic->dims[cost_critic->n_layers] == 1,
                 "crux_lagrange_ppo_update: cost critic shape");
  }
  const int64_t nmb_k = train_cost ? cdiv(n, lhp->cost_batch) : 0;
  const size_t il = (size_t)i64max(1, hp->actor_epochs * nmb_a) * CRUX_PPO_INFO_STRIDE * sizeof(float);
  const size_t ik = (size_t)i64max(1, (int64_t)lhp->cost_epochs * nmb_k) * CRUX_PPO_INFO_STRIDE * sizeof(float);
  float *info = (float *)crux_scratch(ctx, 5, il + ik + 64);
  if (!info) return CRUX_ERR_OOM;
  CRUX_CHECK_CUDA(ctx, cudaMemsetAsync(info, 0, il + ik + 64, ctx->stream));
  LagrangeArgs lg;
  lg.cost_critic = cost_critic; lg.cost = cost; lg.cost_adv = cost_advantage; lg.cost_ret = cost_return; lg.episode_end = episode_end;
  lg.pid = LagrangePid{lhp->target_cost, lhp->penalty_max, lhp->Ki_max, lhp->Ki, lhp->Kp, lhp->Kd, lhp->ema_alpha};
  lg.state = pid_state_dev;
  lg.cost_epochs = train_cost ? lhp->cost_epochs : 0; lg.cost_batch = lhp->cost_batch; lg.cost_max_batches = lhp->cost_max_batches;
  lg.order_cost = order_cost;
  lg.info_l = info; lg.info_cost = (float *)((char *)info + il); lg.sums = (float *)((char *)info + il + ik);
  int rc = ppo_update_impl(actor, critic, s, a, logprob, advantage, ret, n, hp, order_actor, order_critic, seed, &lg);
  if (rc) return rc;
  const int64_t nmb_c = (critic && hp->critic_epochs > 0) ? cdiv(n, hp->critic_batch) : 0;
  if (info_actor_host && hp->actor_epochs > 0)
    CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(info_actor_host, actor->info_actor, il, cudaMemcpyDeviceToHost, ctx->stream));
  if (info_lagrange_host && hp->actor_epochs > 0)
    CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(info_lagrange_host, lg.info_l, il, cudaMemcpyDeviceToHost, ctx->stream));
  if (info_critic_host && nmb_c)
    CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(info_critic_host, actor->info_critic,
                                         (size_t)hp->critic_epochs * nmb_c * CRUX_PPO_INFO_STRIDE * sizeof(float),
                                         cudaMemcpyDeviceToHost, ctx->stream));
  if (info_cost_host && nmb_k)
    CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(info_cost_host, lg.info_cost, ik, cudaMemcpyDeviceToHost, ctx->stream));
  return crux_ctx_check(ctx);
}

int32_t crux_ppo_info_ptrs(crux_gaussian *actor, float **info_actor_dev, float **info_critic_dev) {
  if (!actor) return CRUX_ERR_INVALID;
  if (info_actor_dev) *info_actor_dev = actor->info_actor;
  if (info_critic_dev) *info_critic_dev = actor->info_critic;
  return CRUX_OK;
}

int32_t crux_ppo_update(crux_gaussian *actor, crux_mlp *critic, const float *s, const float *a, const float *logprob,
                        const float *advantage, const float *ret, int64_t n, const crux_ppo_hp *hp,
                        const int32_t *order_actor, const int32_t *order_critic, uint64_t seed, float *info_actor_host,
                        float *info_critic_host) {
  if (!actor) return CRUX_ERR_INVALID;
  crux_ctx *ctx = actor->ctx;
  int rc = ppo_update_impl(actor, critic, s, a, logprob, advantage, ret, n, hp, order_actor, order_critic, seed);
  if (rc) return rc;
  const int64_t nmb_a = cdiv(n, hp->actor_batch);
  const int64_t nmb_c = (critic && hp->critic_epochs > 0) ? cdiv(n, hp->critic_batch) : 0;
  if (info_actor_host && hp->actor_epochs > 0)
    CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(info_actor_host, actor->info_actor,
                                         (size_t)hp->actor_epochs * nmb_a * CRUX_PPO_INFO_STRIDE * sizeof(float),
                                         cudaMemcpyDeviceToHost, ctx->stream));
  if (info_critic_host && nmb_c)
    CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(info_critic_host, actor->info_critic,
                                         (size_t)hp->critic_epochs * nmb_c * CRUX_PPO_INFO_STRIDE * sizeof(float),
                                         cudaMemcpyDeviceToHost, ctx->stream));
  return crux_ctx_check(ctx);  // synchronises; CRUX_ERR_NAN if a gradient norm was NaN (training.jl:20)
}

}  // extern "C"
