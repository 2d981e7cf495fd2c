// Device-side synthetic MDP ("LinQuad") for the HBM-resident leg of the benchmark.
// Not part of the reference: it stands in for a user's POMDPs.jl model so that the sampler hot path
// (src/sampler.jl:71-137: gen -> isterminal -> write row -> terminate_episode!/reset) can be timed
// without the host in the loop.  The CPU twin lives in crux.jl_b200/envs.py.
#include "env.cuh"

namespace {

__device__ __forceinline__ void linquad_s0(uint64_t seed, uint64_t tick, int64_t e, int sdim, float *out) {
  for (int k = 0; k < sdim; k += 4) {
    const Philox4 p = philox4x32_10(seed ^ 0x5851F42D4C957F2DULL, tick, (uint64_t)e * 16 + (k >> 2));
    const uint32_t u[4] = {p.x, p.y, p.z, p.w};
    for (int q = 0; q < 4 && k + q < sdim; ++q) out[k + q] = (u32_to_unit_open(u[q]) * 2.f - 1.f) * 0.1f;  // U(-0.1, 0.1)
  }
}

// the last block to finish advances the tick: every block has read it by then
__device__ __forceinline__ void tick_advance(unsigned long long *tick_dev) {
  __shared__ bool last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(tick_dev + 1, 1ULL) == (unsigned long long)gridDim.x - 1ULL;
  __syncthreads();
  if (last && threadIdx.x == 0) { tick_dev[0] += 1ULL; tick_dev[1] = 0ULL; __threadfence(); }
}

__global__ void linquad_reset_kernel(float *__restrict__ obs, int32_t *__restrict__ ep_len, int64_t n, int sdim, uint64_t seed,
                                     unsigned long long *__restrict__ tick_dev) {
  const uint64_t tick = *(volatile unsigned long long *)tick_dev;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) {
    float s0[LQ_MAX_S];
    linquad_s0(seed, tick, e, sdim, s0);
    for (int k = 0; k < sdim; ++k) obs[e * sdim + k] = s0[k];
    ep_len[e] = 0;
  }
  tick_advance(tick_dev);
}

// 8 lanes per env stream: lane q owns the 4 state dimensions 4q..4q+3 (one Philox draw = the 4 noise values of those
// dimensions, same counters as before), the |s'|^2 partial sums are combined with width-8 shuffles.
__global__ void __launch_bounds__(256)
linquad_step_kernel(const float *__restrict__ Am, const float *__restrict__ Bm, const float *__restrict__ obs,
                    const float *__restrict__ act, int64_t n, int sdim, int adim, int max_steps, uint64_t seed,
                    unsigned long long *__restrict__ tick_dev, int force_end, float *__restrict__ sp_out, float *__restrict__ r_out, uint8_t *__restrict__ done_out,
                    uint8_t *__restrict__ end_out, float *__restrict__ next_obs, int32_t *__restrict__ ep_len) {
  __shared__ float sA[LQ_MAX_S * LQ_MAX_S], sB[LQ_MAX_S * LQ_MAX_A];
  for (int i = threadIdx.x; i < sdim * sdim; i += blockDim.x) sA[i] = Am[i];
  for (int i = threadIdx.x; i < sdim * adim; i += blockDim.x) sB[i] = Bm[i];
  const uint64_t tick = *(volatile unsigned long long *)tick_dev;
  __syncthreads();
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t e = gid >> 3;
  const int q = (int)(gid & 7);
  const bool live = e < n;
  const int64_t ee = live ? e : n - 1;  // clamp: every lane takes part in the shuffles
  float s[LQ_MAX_S], ta[LQ_MAX_A];
#pragma unroll 4
  for (int k = 0; k < sdim; ++k) s[k] = obs[ee * sdim + k];
  const int len = ep_len[ee] + 1;  // sampler.jl:130 (read by all 8 lanes BEFORE the full-mask shuffles; lane 0 writes after them)
  float a2 = 0.f;
  for (int j = 0; j < adim; ++j) { const float a = act[ee * adim + j]; a2 = fmaf(a, a, a2); ta[j] = tanhf(a); }
  float sp[4] = {0.f, 0.f, 0.f, 0.f}, n2 = 0.f;
  const int k0 = 4 * q;
  if (k0 < sdim) {
    const Philox4 p = philox4x32_10(seed, tick, (uint64_t)ee * 16 + q);
    float xi[4];
    box_muller(p.x, p.y, xi[0], xi[1]);
    box_muller(p.z, p.w, xi[2], xi[3]);
    for (int i = 0; i < 4 && k0 + i < sdim; ++i) {
      const int kk = k0 + i;
      float v = 0.f;
      for (int j = 0; j < sdim; ++j) v = fmaf(sA[kk * sdim + j], s[j], v);
      for (int j = 0; j < adim; ++j) v = fmaf(sB[kk * adim + j], ta[j], v);
      v = fmaf(0.01f, xi[i], v);
      v = fminf(fmaxf(v, -10.f), 10.f);
      sp[i] = v; n2 = fmaf(v, v, n2);
    }
  }
  // |s'|^2 over the 8 lanes of the stream (fixed butterfly order); lane 0 holds s'_1
  n2 += __shfl_xor_sync(0xffffffffu, n2, 1, 8);
  n2 += __shfl_xor_sync(0xffffffffu, n2, 2, 8);
  n2 += __shfl_xor_sync(0xffffffffu, n2, 4, 8);
  const float sp0 = __shfl_sync(0xffffffffu, sp[0], 0, 8);
  if (live) {
    const float r = 1.f - n2 / (float)sdim - 0.1f * a2 / (float)adim;
    const bool done = fabsf(sp0) > 5.f;
    const bool end = done || len >= max_steps || force_end;  // :131 and the forced terminate of steps!(reset=true) :148
    float nx[4] = {sp[0], sp[1], sp[2], sp[3]};
    if (end && k0 < sdim) {  // reset_sampler! :31-43 (same counters as linquad_s0)
      const Philox4 p = philox4x32_10(seed ^ 0x5851F42D4C957F2DULL, tick + 0x100000000ULL, (uint64_t)e * 16 + q);
      const uint32_t u[4] = {p.x, p.y, p.z, p.w};
      for (int i = 0; i < 4; ++i) nx[i] = (u32_to_unit_open(u[i]) * 2.f - 1.f) * 0.1f;
    }
    for (int i = 0; i < 4 && k0 + i < sdim; ++i) {
      sp_out[e * sdim + k0 + i] = sp[i];
      next_obs[e * sdim + k0 + i] = nx[i];
    }
    if (q == 0) {
      r_out[e] = r; done_out[e] = done ? 1 : 0; end_out[e] = end ? 1 : 0;
      ep_len[e] = end ? 0 : len;
    }
  }
  tick_advance(tick_dev);
}

}  // namespace

extern "C" {

int32_t crux_linquad_create(crux_ctx *ctx, int32_t sdim, int32_t adim, const float *A_host, const float *B_host, int64_t n_env,
                            int32_t max_steps, uint64_t seed, crux_linquad **out) {
  if (!ctx) return CRUX_ERR_INVALID;
  CRUX_REQUIRE(ctx, out && A_host && B_host, "crux_linquad_create: NULL argument");
  CRUX_REQUIRE(ctx, sdim >= 1 && sdim <= LQ_MAX_S && adim >= 1 && adim <= LQ_MAX_A, "crux_linquad_create: sdim<=32, adim<=16");
  CRUX_REQUIRE(ctx, n_env >= 1 && max_steps >= 1, "crux_linquad_create: bad n_env/max_steps");
  crux_linquad *env = new crux_linquad();
  env->ctx = ctx; env->sdim = sdim; env->adim = adim; env->n_env = n_env; env->max_steps = max_steps; env->seed = seed;
  if (cudaMalloc((void **)&env->A, sdim * sdim * sizeof(float)) != cudaSuccess || cudaMalloc((void **)&env->B, sdim * adim * sizeof(float)) != cudaSuccess ||
      cudaMalloc((void **)&env->ep_len, n_env * sizeof(int32_t)) != cudaSuccess ||
      cudaMalloc((void **)&env->tick, 2 * sizeof(unsigned long long)) != cudaSuccess) {
    crux_linquad_destroy(env);
    return crux_set_err(ctx, CRUX_ERR_OOM, "crux_linquad_create: cudaMalloc");
  }
  cudaMemcpyAsync(env->A, A_host, sdim * sdim * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
  cudaMemcpyAsync(env->B, B_host, sdim * adim * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
  cudaMemsetAsync(env->ep_len, 0, n_env * sizeof(int32_t), ctx->stream);
  cudaMemsetAsync(env->tick, 0, 2 * sizeof(unsigned long long), ctx->stream);
  cudaStreamSynchronize(ctx->stream);
  *out = env;
  return CRUX_OK;
}

int32_t crux_linquad_destroy(crux_linquad *env) {
  if (!env) return CRUX_OK;
  cudaStreamSynchronize(env->ctx->stream);
  if (env->A) cudaFree(env->A);
  if (env->B) cudaFree(env->B);
  if (env->ep_len) cudaFree(env->ep_len);
  if (env->tick) cudaFree(env->tick);
  delete env;
  return CRUX_OK;
}

int32_t crux_linquad_reset(crux_linquad *env, float *obs_out) {
  if (!env || !obs_out) return CRUX_ERR_INVALID;
  crux_ctx *ctx = env->ctx;
  linquad_reset_kernel<<<(unsigned)cdiv(env->n_env, 128), 128, 0, ctx->stream>>>(obs_out, env->ep_len, env->n_env, env->sdim, env->seed,
                                                                                env->tick);
  CRUX_LAUNCHED(ctx);
  return CRUX_OK;
}

int32_t crux_linquad_step(crux_linquad *env, const float *obs, const float *a, float *sp, float *r, uint8_t *done, uint8_t *episode_end,
                          float *next_obs, int32_t force_end) {
  if (!env) return CRUX_ERR_INVALID;
  crux_ctx *ctx = env->ctx;
  CRUX_REQUIRE(ctx, obs && a && sp && r && done && episode_end && next_obs, "crux_linquad_step: NULL pointer");
  CruxTimed timed(ctx, CRUX_T_ENV);
  linquad_step_kernel<<<(unsigned)cdiv(env->n_env * 8, 256), 256, 0, ctx->stream>>>(env->A, env->B, obs, a, env->n_env, env->sdim, env->adim,
                                                                               env->max_steps, env->seed, env->tick, force_end, sp, r,
                                                                               done, episode_end, next_obs, env->ep_len);
  CRUX_LAUNCHED(ctx);
  return CRUX_OK;
}

}  // extern "C"
