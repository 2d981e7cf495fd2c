// Hot path (ii): advantages, returns, whitening and TD targets.
//
// fill_gae! / fill_returns! (src/sampler.jl:262-281) are reverse-time linear recurrences per
// episode.  Over a [T][N] rollout (row t*N+e) both are segmented scans of affine maps
//      A_t = delta_t + c_t A_{t+1},  c_t = episode_end_t ? 0 : lambda*gamma
//      R_t = r_t     + g_t R_{t+1},  g_t = episode_end_t ? 0 : gamma
// with delta_t = r_t + (1-done_t) gamma V(sp_t) - V(s_t).  HBM-bound: 22 B per transition
// (r 4, done 1, episode_end 1, V(s) 4, V(sp) 4 in; adv 4, ret 4 out), read and written once.
//
// Mapping: a warp covers 32 adjacent env streams (128 B coalesced rows); a thread keeps L
// consecutive time steps of its stream in registers; the W warps of a CTA cover a chunk of W*L
// steps and exchange carries through shared memory; chunks of the same env tile live in different
// CTAs and exchange their (coefficient, offset) aggregates through global memory with a
// decoupled look-back over the later chunks only, so the grid fills all SMs even when N is small.
// Blocks are ordered chunk-major, latest chunk first: a CTA only ever waits on CTAs with a lower
// block index (dispatched earlier, no deadlock), and those were dispatched a whole tile-sweep
// earlier, so their aggregates are normally already published.  The look-back itself is parallel:
// warp k of the CTA polls and fetches the aggregate of later chunk k (one L2 round trip in total
// instead of one per predecessor), then the composition runs out of shared memory.
#include "common.cuh"

namespace {

struct Agg { float pa, la, pr, lr; };  // A_in -> la + pa*A_in ; R_in -> lr + pr*R_in

template <int L>
__global__ void __launch_bounds__(512, 2)
gae_returns_kernel(const float *__restrict__ r, const uint8_t *__restrict__ done,
                   const uint8_t *__restrict__ ee, const float *__restrict__ vs,
                   const float *__restrict__ vsp, int64_t T, int64_t N, float gamma, float lambda,
                   float *__restrict__ adv, float *__restrict__ ret, int n_chunks, int chunk_len,
                   float4 *__restrict__ agg_g, unsigned int *__restrict__ flag_g,
                   unsigned int *__restrict__ err_flags) {
  const int lane = threadIdx.x, w = threadIdx.y, W = blockDim.y;
  const int chunk = n_chunks - 1 - (int)blockIdx.y;  // grid (tiles, chunks): chunk-major dispatch, later chunks first
  const int tile = blockIdx.x;
  const int64_t e = (int64_t)tile * 32 + lane;
  const bool live = e < N;
  const int64_t c_lo = (int64_t)chunk * chunk_len;
  const int steps = (int)min((int64_t)chunk_len, T - c_lo);   // steps of this chunk
  const int s0 = w * L;                                       // this thread's steps: [s0, s0+L) clipped to `steps`
  const float c = lambda * gamma;
  // 64-bit base once, 32-bit offsets per step (chunk_len * N < 2^31 is checked on the host)
  const int64_t base = c_lo * N + (live ? e : N - 1);
  const float *rp = r + base, *vsb = vs + base, *vspb = vsp + base;
  const uint8_t *dnb = done + base, *eeb = ee + base;
  const unsigned int Nu = (unsigned int)N;

  // Issue ALL loads of this thread's L steps back to back (no control dependence), so every warp keeps 5*L cache lines
  // in flight.  Out-of-range steps / lanes read a clamped valid address and are masked when they are used.
  float dl[L], rr[L];
  unsigned int cut = 0;  // bit i: episode_end at step s0+i
  {
    float va[L], vb[L];
    unsigned int dn[L], en[L];
#pragma unroll
    for (int i = 0; i < L; ++i) {
      // unconditional loads from clamped (always valid) addresses: plain address arithmetic + LDG, nothing to branch on
      const unsigned int off = (unsigned int)min(s0 + i, steps - 1) * Nu;
      rr[i] = __ldcs(rp + off);
      va[i] = __ldcs(vsb + off);
      vb[i] = __ldcs(vspb + off);
      dn[i] = (unsigned int)__ldcs(dnb + off);
      en[i] = (unsigned int)__ldcs(eeb + off);
    }
#pragma unroll
    for (int i = 0; i < L; ++i) {
      const float nd = dn[i] ? 0.f : gamma;                   // (1 - done) * gamma
      if (en[i]) cut |= 1u << i;
      dl[i] = (rr[i] + nd * vb[i]) - va[i];
    }
  }
  // local aggregate of this thread's L steps.  Padded steps (beyond `steps`) must be the identity, not a cut:
  Agg g = {1.f, 0.f, 1.f, 0.f};
#pragma unroll
  for (int i = L - 1; i >= 0; --i) {
    if (s0 + i < steps) {
      const bool k = (cut >> i) & 1u;
      const float ca = k ? 0.f : c, cg = k ? 0.f : gamma;
      g.la = fmaf(ca, g.la, dl[i]); g.pa *= ca;
      g.lr = fmaf(cg, g.lr, rr[i]); g.pr *= cg;
    }
  }
  __shared__ float4 s_agg[16][32];
  __shared__ float2 s_carry[16][32];
  __shared__ float4 s_look[64][32];
  s_agg[w][lane] = make_float4(g.pa, g.la, g.pr, g.lr);
  __syncthreads();

  // warp 0 publishes this chunk's aggregate for the earlier chunks BEFORE anybody waits on later chunks
  if (w == 0 && n_chunks > 1 && chunk > 0) {
    Agg q = {1.f, 0.f, 1.f, 0.f};
    for (int ww = W - 1; ww >= 0; --ww) {
      const float4 x = s_agg[ww][lane];
      q.la = fmaf(x.x, q.la, x.y); q.pa *= x.x;
      q.lr = fmaf(x.z, q.lr, x.w); q.pr *= x.z;
    }
    agg_g[((int64_t)tile * n_chunks + chunk) * 32 + lane] = make_float4(q.pa, q.la, q.pr, q.lr);
    __threadfence();
    __syncwarp();
    if (lane == 0) atomicExch(flag_g + (int64_t)tile * n_chunks + chunk, 1u);
  }
  // parallel look-back fetch: warp ww handles later chunks chunk+1+ww, chunk+1+ww+W, ... -> shared memory
  const int n_later = n_chunks - 1 - chunk;
  for (int q = w; q < n_later; q += W) {
    const int k = chunk + 1 + q;
    if (lane == 0) {
      const volatile unsigned int *f = flag_g + (int64_t)tile * n_chunks + k;
      while (*f == 0u) { __nanosleep(32); }
    }
    __syncwarp();
    __threadfence();
    s_look[q][lane] = __ldcg(agg_g + ((int64_t)tile * n_chunks + k) * 32 + lane);
  }
  __syncthreads();

  // warp 0: compose the carry into the END of this chunk from the later chunks, then walk the warps of this CTA from the
  // latest to the earliest, leaving each warp's incoming carry in shared memory.
  if (w == 0) {
    float Ain = 0.f, Rin = 0.f;
    for (int q = n_later - 1; q >= 0; --q) {  // latest chunk first
      const float4 x = s_look[q][lane];
      Ain = fmaf(x.x, Ain, x.y);
      Rin = fmaf(x.z, Rin, x.w);
    }
    for (int ww = W - 1; ww >= 0; --ww) {
      s_carry[ww][lane] = make_float2(Ain, Rin);
      const float4 x = s_agg[ww][lane];
      Ain = fmaf(x.x, Ain, x.y);
      Rin = fmaf(x.z, Rin, x.w);
    }
  }
  __syncthreads();
  // final pass over the registers with the true carry
  const float2 cin = s_carry[w][lane];
  float A = cin.x, R = cin.y;
  bool bad = false;
  float *ap = adv ? adv + base : nullptr, *rtp = ret ? ret + base : nullptr;
#pragma unroll
  for (int i = L - 1; i >= 0; --i) {
    if (live && s0 + i < steps) {
      const bool k = (cut >> i) & 1u;
      A = fmaf(k ? 0.f : c, A, dl[i]);
      R = fmaf(k ? 0.f : gamma, R, rr[i]);
      const unsigned int off = (unsigned int)(s0 + i) * Nu;
      if (ap) __stcs(ap + off, A);
      if (rtp) __stcs(rtp + off, R);
      bad |= isnan(A);
    }
  }
  if (bad) atomicOr(err_flags, CRUX_FLAG_NAN);  // sampler.jl:270 @assert !isnan(A)
}

// ---------------------------------------------------------------- whiten (utils.jl:41-42)
__global__ void moments_kernel(const float *__restrict__ x, int64_t n, double *__restrict__ part) {
  double s = 0.0, s2 = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = (double)x[i];
    s += v; s2 += v * v;
  }
  __shared__ double sh[2][32];
  s = warp_sum_d(s); s2 = warp_sum_d(s2);
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s; sh[1][threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int nw = blockDim.x >> 5;
    s = threadIdx.x < nw ? sh[0][threadIdx.x] : 0.0;
    s2 = threadIdx.x < nw ? sh[1][threadIdx.x] : 0.0;
    s = warp_sum_d(s); s2 = warp_sum_d(s2);
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = s; part[2 * blockIdx.x + 1] = s2; }
  }
}
// stats[0]=sum, stats[1]=sumsq, stats[2]=count (float for the NCCL sum; exact below 2^24 per rank... use double pairs)
__global__ void moments_final_kernel(const double *__restrict__ part, int nparts, int64_t n, float *__restrict__ stats3) {
  // one warp, fixed order: lane l sums partials l, l+32, ... then a butterfly (deterministic for a given nparts)
  double s = 0.0, s2 = 0.0;
  for (int i = threadIdx.x; i < nparts; i += 32) { s += part[2 * i]; s2 += part[2 * i + 1]; }
  s = warp_sum_d(s); s2 = warp_sum_d(s2);
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    // shift-free f32 triple is too lossy for the variance: keep hi/lo splits of the doubles
    float sh = (float)s, sl = (float)(s - (double)sh);
    float qh = (float)s2, ql = (float)(s2 - (double)qh);
    stats3[0] = sh; stats3[1] = sl; stats3[2] = qh; stats3[3] = ql; stats3[4] = (float)n; stats3[5] = 0.f;
  }
}
__global__ void whiten_apply_kernel(float *__restrict__ x, int64_t n, const float *__restrict__ stats) {
  const double s = (double)stats[0] + (double)stats[1];
  const double s2 = (double)stats[2] + (double)stats[3];
  const double cnt = (double)stats[4];
  const double mean = s / cnt;
  double var = (s2 - cnt * mean * mean) / (cnt - 1.0);  // Bessel
  const float mu = (float)mean, sd = (float)sqrt(var > 0.0 ? var : 0.0);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    x[i] = (x[i] - mu) / sd;
}

// ---------------------------------------------------------------- TD targets
__global__ void dqn_target_kernel(const float *__restrict__ r, const uint8_t *__restrict__ done,
                                  const float *__restrict__ q, int64_t B, int nA, float gamma, float *__restrict__ y) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  float m = q[i * nA];
  for (int a = 1; a < nA; ++a) m = fmaxf(m, q[i * nA + a]);
  y[i] = r[i] + (gamma * (1.0f - (float)done[i])) * m;  // rl/dqn.jl:5 (left-to-right)
}
// softq_target rl/softq.jl:13-17 with soft_value :8: α .* logsumexp(Q⁻(sp) ./ α) (NNlib logsumexp: max + log(Σ exp(x - max)))
__global__ void softq_target_kernel(const float *__restrict__ r, const uint8_t *__restrict__ done,
                                    const float *__restrict__ q, int64_t B, int nA, float gamma, float alpha, float *__restrict__ y) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  float m = q[i * nA] / alpha;
  for (int a = 1; a < nA; ++a) m = fmaxf(m, q[i * nA + a] / alpha);
  float s = 0.f;
  for (int a = 0; a < nA; ++a) s += expf(q[i * nA + a] / alpha - m);
  const float v = alpha * (m + logf(s));
  y[i] = r[i] + (gamma * (1.0f - (float)done[i])) * v;
}
__global__ void sac_target_kernel(const float *__restrict__ r, const uint8_t *__restrict__ done,
                                  const float *__restrict__ q1, const float *__restrict__ q2,
                                  const float *__restrict__ logp, int64_t B, float gamma,
                                  const float *__restrict__ log_alpha, float *__restrict__ y) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const float alpha = expf(log_alpha[0]);
  y[i] = r[i] + (gamma * (1.0f - (float)done[i])) * (fminf(q1[i], q2[i]) - alpha * logp[i]);  // rl/sac.jl:7
}
__global__ void td_error_kernel(const float *__restrict__ q, const float *__restrict__ y, int64_t B, float *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) out[i] = fabsf(q[i] - y[i]);
}
__global__ void discrete_q_sa_kernel(const float *__restrict__ q, const float *__restrict__ oh, int64_t B, int nA, float *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  float s = 0.f;
  for (int a = 0; a < nA; ++a) s += q[i * nA + a] * oh[i * nA + a];  // policies.jl:122
  out[i] = s;
}
__global__ void normalize_kernel(const float *__restrict__ x, int64_t n, float mu, float sigma, float *__restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (x[i] - mu) / sigma;
}

}  // namespace

// gae_tma.cu: TMA-fed streaming scan for large rollouts (*handled = 0 when the shape is not eligible)
int gae_tma_launch(crux_ctx *ctx, const float *r, const uint8_t *done, const uint8_t *ee, const float *vs, const float *vsp, int64_t T, int64_t N,
                   float gamma, float lambda, float *adv, float *ret, int *handled);

extern "C" {

int32_t crux_fill_gae_returns(crux_ctx *ctx, const float *r, const uint8_t *done, const uint8_t *episode_end,
                              const float *v_s, const float *v_sp, int64_t T, int64_t N, float gamma,
                              float lambda, float *adv, float *ret) {
  if (!ctx) return CRUX_ERR_INVALID;
  CRUX_REQUIRE(ctx, T >= 0 && N >= 0, "crux_fill_gae_returns: negative shape");
  if (T == 0 || N == 0) return CRUX_OK;
  CRUX_REQUIRE(ctx, r && done && episode_end && v_s && v_sp, "crux_fill_gae_returns: NULL input column");
  CRUX_REQUIRE(ctx, adv || ret, "crux_fill_gae_returns: both outputs NULL");
  {
    int handled = 0;
    const int rc = gae_tma_launch(ctx, r, done, episode_end, v_s, v_sp, T, N, gamma, lambda, adv, ret, &handled);
    if (rc || handled) return rc;
  }
  const int64_t tiles = cdiv(N, 32);
  // steps per thread L, warps per CTA W -> chunk of W*L steps
  int L, W;
  if (T <= 32) { L = 4; W = (int)cdiv(T, 4); }
  else if (T <= 128) { L = 8; W = (int)cdiv(T, 8); }
  else { L = 8; W = 16; }  // chunk of 128 steps; 512 threads x <= 64 registers -> 2 CTAs (1024 threads) per SM
  int chunk_len = L * W;
  int n_chunks = (int)cdiv(T, chunk_len);
  CRUX_REQUIRE(ctx, n_chunks <= 65, "crux_fill_gae_returns: T > 8320 steps per rollout is not supported");
  CRUX_REQUIRE(ctx, (int64_t)chunk_len * N < (1ll << 31), "crux_fill_gae_returns: N too large for 32-bit in-chunk offsets");
  CRUX_REQUIRE(ctx, tiles * n_chunks < (1ll << 31), "crux_fill_gae_returns: grid too large");
  float4 *agg = nullptr;
  unsigned int *flags = nullptr;
  if (n_chunks > 1) {
    const size_t agg_bytes = (size_t)tiles * n_chunks * 32 * sizeof(float4);
    const size_t flag_bytes = (size_t)tiles * n_chunks * sizeof(unsigned int);
    char *p = (char *)crux_scratch(ctx, 0, agg_bytes + flag_bytes);
    if (!p) return CRUX_ERR_OOM;
    agg = (float4 *)p;
    flags = (unsigned int *)(p + agg_bytes);
    CRUX_CHECK_CUDA(ctx, cudaMemsetAsync(flags, 0, flag_bytes, ctx->stream));
  }
  dim3 block(32, W), grid((unsigned)tiles, (unsigned)n_chunks);
#define GAE_LAUNCH(LL)                                                                                   \
  gae_returns_kernel<LL><<<grid, block, 0, ctx->stream>>>(r, done, episode_end, v_s, v_sp, T, N, gamma,  \
                                                          lambda, adv, ret, n_chunks, chunk_len, agg,    \
                                                          flags, ctx->flags_dev)
  { CruxTimed timed(ctx, CRUX_T_GAE);
  if (L == 4) GAE_LAUNCH(4); else GAE_LAUNCH(8); }
#undef GAE_LAUNCH
  CRUX_LAUNCHED(ctx);
  return CRUX_OK;
}

int32_t crux_nccl_allreduce_f32(crux_ctx *ctx, float *buf, int64_t n);

int32_t crux_whiten(crux_ctx *ctx, float *x, int64_t n) {
  if (!ctx) return CRUX_ERR_INVALID;
  CRUX_REQUIRE(ctx, n >= 0, "crux_whiten: negative n");
  if (n == 0) return CRUX_OK;
  CRUX_REQUIRE(ctx, x, "crux_whiten: NULL");
  const int threads = 256;
  const int blocks = (int)i64min(cdiv(n, threads * 4), (int64_t)ctx->num_sms * 4);
  char *p = (char *)crux_scratch(ctx, 1, (size_t)blocks * 2 * sizeof(double) + 64);
  if (!p) return CRUX_ERR_OOM;
  float *stats = (float *)p;
  double *part = (double *)(p + 64);
  moments_kernel<<<blocks, threads, 0, ctx->stream>>>(x, n, part);
  CRUX_LAUNCHED(ctx);
  moments_final_kernel<<<1, 32, 0, ctx->stream>>>(part, blocks, n, stats);
  CRUX_LAUNCHED(ctx);
  if (ctx->world > 1) {
    int rc = crux_nccl_allreduce_f32(ctx, stats, 6);
    if (rc) return rc;
  }
  whiten_apply_kernel<<<(int)i64min(cdiv(n, threads), (int64_t)ctx->num_sms * 8), threads, 0, ctx->stream>>>(x, n, stats);
  CRUX_LAUNCHED(ctx);
  return CRUX_OK;
}

int32_t crux_dqn_target(crux_ctx *ctx, const float *r, const uint8_t *done, const float *q_sp, int64_t B,
                        int32_t nA, float gamma, float *y) {
  if (!ctx) return CRUX_ERR_INVALID;
  CRUX_REQUIRE(ctx, B >= 0 && nA >= 1, "crux_dqn_target: bad shape");
  if (B == 0) return CRUX_OK;
  dqn_target_kernel<<<(unsigned)cdiv(B, 256), 256, 0, ctx->stream>>>(r, done, q_sp, B, nA, gamma, y);
  CRUX_LAUNCHED(ctx);
  return CRUX_OK;
}

int32_t crux_softq_target(crux_ctx *ctx, const float *r, const uint8_t *done, const float *q_sp, int64_t B,
                          int32_t nA, float gamma, float alpha, float *y) {
  if (!ctx) return CRUX_ERR_INVALID;
  CRUX_REQUIRE(ctx, B >= 0 && nA >= 1 && alpha > 0.f, "crux_softq_target: bad shape or temperature");
  if (B == 0) return CRUX_OK;
  softq_target_kernel<<<(unsigned)cdiv(B, 256), 256, 0, ctx->stream>>>(r, done, q_sp, B, nA, gamma, alpha, y);
  CRUX_LAUNCHED(ctx);
  return CRUX_OK;
}

int32_t crux_sac_target(crux_ctx *ctx, const float *r, const uint8_t *done, const float *q1, const float *q2,
                        const float *logp, int64_t B, float gamma, const float *log_alpha_dev, float *y) {
  if (!ctx) return CRUX_ERR_INVALID;
  if (B <= 0) return CRUX_OK;
  sac_target_kernel<<<(unsigned)cdiv(B, 256), 256, 0, ctx->stream>>>(r, done, q1, q2, logp, B, gamma, log_alpha_dev, y);
  CRUX_LAUNCHED(ctx);
  return CRUX_OK;
}

int32_t crux_td_error(crux_ctx *ctx, const float *q_sa, const float *y, int64_t B, float *out) {
  if (!ctx) return CRUX_ERR_INVALID;
  if (B <= 0) return CRUX_OK;
  td_error_kernel<<<(unsigned)cdiv(B, 256), 256, 0, ctx->stream>>>(q_sa, y, B, out);
  CRUX_LAUNCHED(ctx);
  return CRUX_OK;
}

int32_t crux_discrete_q_sa(crux_ctx *ctx, const float *q, const float *a_onehot, int64_t B, int32_t nA, float *out) {
  if (!ctx) return CRUX_ERR_INVALID;
  if (B <= 0) return CRUX_OK;
  discrete_q_sa_kernel<<<(unsigned)cdiv(B, 256), 256, 0, ctx->stream>>>(q, a_onehot, B, nA, out);
  CRUX_LAUNCHED(ctx);
  return CRUX_OK;
}

int32_t crux_normalize_obs(crux_ctx *ctx, const float *x, int64_t n, float mu, float sigma, float *out) {
  if (!ctx) return CRUX_ERR_INVALID;
  if (n <= 0) return CRUX_OK;
  normalize_kernel<<<(int)i64min(cdiv(n, 256), (int64_t)ctx->num_sms * 8), 256, 0, ctx->stream>>>(x, n, mu, sigma, out);
  CRUX_LAUNCHED(ctx);
  return CRUX_OK;
}

}  // extern "C"
