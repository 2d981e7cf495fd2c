// Hot path (i) for HOST environments: the whole `steps!` loop (src/sampler.jl:139-155, step! :71-137, terminate_episode!
// :53-69) for N env streams in one C call.  The environment (the user's POMDPs.jl model in the reference) is reached through
// two callbacks -- step (@gen(:sp,:r) + isterminal for a range of streams) and reset (rand(initialstate) for the listed streams)
// -- which a Julia host passes as @cfunction pointers; everything else (pinned staging, H2D/D2H, the fused policy forward,
// episode bookkeeping) stays on this side of the ABI, so no interpreter sits between two vector steps.
//
// The N streams are split into two halves that leapfrog: while the host steps the env streams of one half, the device runs the
// policy forward (and the copies) of the other half.  Per half g and vector step t, all on the context stream:
//     obs_g (pinned) --H2D--> s[t] rows of g      fused_forward (a, logprob rows of g)      a rows --D2H--> pinned      event E_g
//     wait E_g    step callback(g) -> sp, r, done (pinned)    episode_end = done | len >= max_steps | (reset_at_end & last step)
//     reset callback for ended streams -> next obs_g      sp, r, done, episode_end rows of g --H2D--> row t
// Exploration noise is keyed by (seed, ctr0 + t, absolute stream id): the result does not depend on the split.
#include "policy.cuh"
#include <vector>

extern "C" int32_t crux_rollout_step_rows(crux_gaussian *actor, const float *obs, int64_t N, int64_t row0, uint64_t seed, uint64_t ctr,
                                          float *a_out, float *logp_out);
bool fused_rows_supported(const crux_gaussian *actor);

struct HostRolloutStage {
  float *a = nullptr, *sp = nullptr, *r = nullptr, *robs = nullptr;
  uint8_t *done = nullptr, *ee = nullptr;
  int64_t N = 0; int sdim = 0, adim = 0;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  std::vector<int32_t> idx;
};

static HostRolloutStage g_stage;  // one staging set per process (contexts are one per process / GPU)

static int ensure_stage(crux_ctx *ctx, int64_t N, int sdim, int adim) {
  HostRolloutStage &S = g_stage;
  if (S.N >= N && S.sdim == sdim && S.adim == adim) return CRUX_OK;
  CRUX_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (S.a) { cudaFreeHost(S.a); cudaFreeHost(S.sp); cudaFreeHost(S.r); cudaFreeHost(S.robs); cudaFreeHost(S.done); cudaFreeHost(S.ee); }
  cudaEvent_t e0 = S.ev[0], e1 = S.ev[1];
  S = HostRolloutStage();
  S.ev[0] = e0; S.ev[1] = e1;
  CRUX_CHECK_CUDA(ctx, cudaMallocHost((void **)&S.a, (size_t)N * adim * sizeof(float)));
  CRUX_CHECK_CUDA(ctx, cudaMallocHost((void **)&S.sp, (size_t)N * sdim * sizeof(float)));
  CRUX_CHECK_CUDA(ctx, cudaMallocHost((void **)&S.robs, (size_t)N * sdim * sizeof(float)));
  CRUX_CHECK_CUDA(ctx, cudaMallocHost((void **)&S.r, (size_t)N * sizeof(float)));
  CRUX_CHECK_CUDA(ctx, cudaMallocHost((void **)&S.done, (size_t)N));
  CRUX_CHECK_CUDA(ctx, cudaMallocHost((void **)&S.ee, (size_t)N));
  for (int g = 0; g < 2; ++g)
    if (!S.ev[g]) CRUX_CHECK_CUDA(ctx, cudaEventCreateWithFlags(&S.ev[g], cudaEventDisableTiming));
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
  // two leapfrogging halves when the policy supports split vector steps and the halves are worth a launch each
  const int G = (fused_rows_supported(actor) && N >= 512) ? 2 : 1;
  const int64_t lo[2] = {0, G == 2 ? N / 2 : N}, hi[2] = {G == 2 ? N / 2 : N, N};

  auto enqueue_forward = [&](int g, int t) -> int {  // obs_g -> s[t]; policy forward; action back to the host; event
    const int64_t row = (int64_t)t * N + lo[g], n = hi[g] - lo[g];
    float *s_t = cols->s + row * sdim, *a_t = cols->a + row * adim;
    CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(s_t, obs_pinned + lo[g] * sdim, (size_t)n * sdim * sizeof(float), cudaMemcpyHostToDevice, st));
    float *lp = cols->logprob ? cols->logprob + row : nullptr;
    int rc2 = G == 2 ? crux_rollout_step_rows(actor, s_t, n, lo[g], seed, ctr0 + (uint64_t)t, a_t, lp)
                     : crux_rollout_step(actor, nullptr, s_t, n, nullptr, seed, ctr0 + (uint64_t)t, a_t, lp, nullptr);
    if (rc2) return rc2;
    CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(S.a + lo[g] * adim, a_t, (size_t)n * adim * sizeof(float), cudaMemcpyDeviceToHost, st));
    CRUX_CHECK_CUDA(ctx, cudaEventRecord(S.ev[g], st));
    return CRUX_OK;
  };

  for (int g = 0; g < G; ++g) { rc = enqueue_forward(g, 0); if (rc) return rc; }
  for (int t = 0; t < T; ++t) {
    const bool force = reset_at_end && t == T - 1;                                                     // steps!(reset=true) :148
    for (int g = 0; g < G; ++g) {
      const int64_t e0 = lo[g], e1 = hi[g], n = e1 - e0, row = (int64_t)t * N + e0;
      CRUX_CHECK_CUDA(ctx, cudaEventSynchronize(S.ev[g]));   // actions of half g are on the host; its earlier H2Ds are done too
      step(user, (int32_t)e0, (int32_t)e1, S.a, S.sp, S.r, S.done);                                    // @gen(:sp,:r), isterminal
      S.idx.clear();
      for (int64_t e = e0; e < e1; ++e) {
        const int32_t len = ++episode_length[e];                                                       // sampler.jl:130
        const bool end = S.done[e] || len >= max_steps || force;
        S.ee[e] = end ? 1 : 0;
        if (end) { S.idx.push_back((int32_t)e); episode_length[e] = 0; }
      }
      CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(cols->sp + row * sdim, S.sp + e0 * sdim, (size_t)n * sdim * sizeof(float), cudaMemcpyHostToDevice, st));
      CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(cols->r + row, S.r + e0, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, st));
      CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(cols->done + row, S.done + e0, (size_t)n, cudaMemcpyHostToDevice, st));
      CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(cols->episode_end + row, S.ee + e0, (size_t)n, cudaMemcpyHostToDevice, st));
      // next observation: sp, or a fresh initial state where the episode ended (terminate_episode! -> reset_sampler!)
      memcpy(obs_pinned + e0 * sdim, S.sp + e0 * sdim, (size_t)n * sdim * sizeof(float));
      if (!S.idx.empty()) {
        reset(user, S.idx.data(), (int32_t)S.idx.size(), S.robs);
        for (size_t q = 0; q < S.idx.size(); ++q) memcpy(obs_pinned + (size_t)S.idx[q] * sdim, S.robs + q * sdim, sizeof(float) * sdim);
      }
      // the forward of this half for the next vector step runs on the device while the host steps the other half
      if (t + 1 < T) { rc = enqueue_forward(g, t + 1); if (rc) return rc; }
    }
  }
  CRUX_CHECK_CUDA(ctx, cudaStreamSynchronize(st));
  return CRUX_OK;
}
