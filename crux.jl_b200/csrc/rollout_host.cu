// Hot path (i) for HOST environments: the whole `steps!` loop (src/sampler.jl:139-155, step! :71-137, terminate_episode!
// :53-69) for N env streams in one C call.  The environment (the user's POMDPs.jl model in the reference) is reached through
// two callbacks -- step (@gen(:sp,:r) + isterminal for every stream) and reset (rand(initialstate) for the listed streams) --
// which a Julia host passes as @cfunction pointers; everything else (pinned staging, H2D/D2H, the fused policy forward,
// episode bookkeeping) stays on this side of the ABI, so no interpreter sits between two vector steps.
//
// Per vector step t:  obs (pinned) --H2D--> s[t]   fused_forward (a[t], logprob[t])   a[t] --D2H--> pinned   sync
//                     step callback -> sp, r, done (pinned)   episode_end = done | len >= max_steps | (reset_at_end & last step)
//                     reset callback for ended streams -> next obs     sp, r, done, episode_end --H2D--> row t
#include "policy.cuh"
#include <vector>

struct HostRolloutStage {
  float *a = nullptr, *sp = nullptr, *r = nullptr, *robs = nullptr;
  uint8_t *done = nullptr, *ee = nullptr;
  int64_t N = 0; int sdim = 0, adim = 0;
  std::vector<int32_t> idx;
};

static HostRolloutStage g_stage;  // one staging set per process (contexts are one per process / GPU)

static int ensure_stage(crux_ctx *ctx, int64_t N, int sdim, int adim) {
  HostRolloutStage &S = g_stage;
  if (S.N >= N && S.sdim == sdim && S.adim == adim) return CRUX_OK;
  CRUX_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (S.a) { cudaFreeHost(S.a); cudaFreeHost(S.sp); cudaFreeHost(S.r); cudaFreeHost(S.robs); cudaFreeHost(S.done); cudaFreeHost(S.ee); }
  S = HostRolloutStage();
  CRUX_CHECK_CUDA(ctx, cudaMallocHost((void **)&S.a, (size_t)N * adim * sizeof(float)));
  CRUX_CHECK_CUDA(ctx, cudaMallocHost((void **)&S.sp, (size_t)N * sdim * sizeof(float)));
  CRUX_CHECK_CUDA(ctx, cudaMallocHost((void **)&S.robs, (size_t)N * sdim * sizeof(float)));
  CRUX_CHECK_CUDA(ctx, cudaMallocHost((void **)&S.r, (size_t)N * sizeof(float)));
  CRUX_CHECK_CUDA(ctx, cudaMallocHost((void **)&S.done, (size_t)N));
  CRUX_CHECK_CUDA(ctx, cudaMallocHost((void **)&S.ee, (size_t)N));
  S.N = N; S.sdim = sdim; S.adim = adim;
  S.idx.reserve((size_t)N);
  return CRUX_OK;
}

extern "C" int32_t crux_rollout_host(crux_gaussian *actor, int64_t N, int32_t T, int32_t max_steps, int32_t reset_at_end,
                                     crux_env_step_fn step, crux_env_reset_fn reset, void *user, float *obs_pinned,
                                     int32_t *episode_length, const crux_rollout_cols *cols, uint64_t seed, uint64_t ctr0) {
  if (!actor) return CRUX_ERR_INVALID;
  crux_ctx *ctx = actor->ctx;
  CRUX_REQUIRE(ctx, N >= 1 && T >= 1 && max_steps >= 1, "crux_rollout_host: bad N/T/max_steps");
  CRUX_REQUIRE(ctx, step && reset && obs_pinned && episode_length && cols, "crux_rollout_host: NULL argument");
  CRUX_REQUIRE(ctx, cols->s && cols->a && cols->sp && cols->r && cols->done && cols->episode_end, "crux_rollout_host: NULL column");
  const int sdim = actor->mu->dims[0], adim = actor->adim;
  int rc = ensure_stage(ctx, N, sdim, adim);
  if (rc) return rc;
  HostRolloutStage &S = g_stage;
  cudaStream_t st = ctx->stream;
  const size_t ob = (size_t)N * sdim * sizeof(float), ab = (size_t)N * adim * sizeof(float);
  for (int t = 0; t < T; ++t) {
    const int64_t row = (int64_t)t * N;
    float *s_t = cols->s + row * sdim, *a_t = cols->a + row * adim;
    CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(s_t, obs_pinned, ob, cudaMemcpyHostToDevice, st));           // svec of every stream
    rc = crux_rollout_step(actor, nullptr, s_t, N, nullptr, seed, ctr0 + (uint64_t)t, a_t, cols->logprob ? cols->logprob + row : nullptr, nullptr);
    if (rc) return rc;
    CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(S.a, a_t, ab, cudaMemcpyDeviceToHost, st));                  // the env needs the action
    CRUX_CHECK_CUDA(ctx, cudaStreamSynchronize(st));  // also: the copies queued from the staging buffers at step t-1 are done
    step(user, S.a, S.sp, S.r, S.done);                                                               // @gen(:sp,:r), isterminal
    S.idx.clear();
    const bool force = reset_at_end && t == T - 1;                                                     // steps!(reset=true) :148
    for (int64_t e = 0; e < N; ++e) {
      const int32_t len = ++episode_length[e];                                                         // sampler.jl:130
      const bool end = S.done[e] || len >= max_steps || force;
      S.ee[e] = end ? 1 : 0;
      if (end) { S.idx.push_back((int32_t)e); episode_length[e] = 0; }
    }
    CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(cols->sp + row * sdim, S.sp, ob, cudaMemcpyHostToDevice, st));
    CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(cols->r + row, S.r, (size_t)N * sizeof(float), cudaMemcpyHostToDevice, st));
    CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(cols->done + row, S.done, (size_t)N, cudaMemcpyHostToDevice, st));
    CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(cols->episode_end + row, S.ee, (size_t)N, cudaMemcpyHostToDevice, st));
    // next observation: sp, or a fresh initial state where the episode ended (terminate_episode! -> reset_sampler!)
    memcpy(obs_pinned, S.sp, ob);
    if (!S.idx.empty()) {
      reset(user, S.idx.data(), (int32_t)S.idx.size(), S.robs);
      for (size_t q = 0; q < S.idx.size(); ++q) memcpy(obs_pinned + (size_t)S.idx[q] * sdim, S.robs + q * sdim, sizeof(float) * sdim);
    }
  }
  CRUX_CHECK_CUDA(ctx, cudaStreamSynchronize(st));
  return CRUX_OK;
}
