// Hot path (i) for HOST environments: the whole `steps!` loop (src/sampler.jl:139-155, step! :71-137, terminate_episode!
// :53-69) for N env streams in one C call.  The environment (the user's POMDPs.jl model in the reference) is reached through
// two callbacks -- step (@gen(:sp,:r) + isterminal for a range of streams) and reset (rand(initialstate) for the listed streams)
// -- which a Julia host passes as @cfunction pointers; everything else (pinned staging, H2D/D2H, the fused policy forward,
// episode bookkeeping) stays on this side of the ABI, so no interpreter sits between two vector steps.
//
// The N streams are split into two halves that leapfrog: while the host steps the env streams of one half, the device runs the
// policy forward of the other half.  Per half g and vector step t, all on the context stream:
//     ONE launch: the forward kernel reads obs_g straight from pinned host memory (PCIe), stores it as the s[t] rows of g, computes
//                 a / logprob and writes the actions to the device column AND to pinned host memory        event E_g
//     wait E_g    step callback(g) -> sp, r, done (pinned, full-rollout staging)    episode_end = done | len >= max_steps | forced
//     reset callback for ended streams -> next obs_g
// The transition columns (sp, r, done, episode_end) are not needed on the device before the rollout ends: they are uploaded in
// chunks of UPLOAD_STEPS whole vector steps (4 copy calls per chunk instead of 8 per vector step).
// Exploration noise is keyed by (seed, ctr0 + t, absolute stream id): the result does not depend on the split.
#include "policy.cuh"
#include <vector>
#include <chrono>
#include <stdlib.h>

extern "C" int32_t crux_rollout_step_rows_mapped(crux_gaussian *actor, const float *obs_pinned, int64_t N, int64_t row0, uint64_t seed, uint64_t ctr,
                                                 float *s_dev, float *a_dev, float *a_pinned, float *logp_dev, const uint8_t *reset_flag_pinned,
                                                 const float *reset_obs_pinned);
bool fused_rows_supported(const crux_gaussian *actor);

struct HostRolloutStage {
  float *a = nullptr, *sp = nullptr, *r = nullptr, *robs = nullptr;
  uint8_t *done = nullptr, *ee = nullptr;
  int64_t N = 0, T = 0; int sdim = 0, adim = 0;   // a: [N][adim]; sp, r, done, ee: [T][N] (full rollout)
  cudaEvent_t ev[2] = {nullptr, nullptr};
  std::vector<int32_t> idx;
};

static HostRolloutStage g_stage;  // one staging set per process (contexts are one per process / GPU)

static int ensure_stage(crux_ctx *ctx, int64_t N, int64_t T, int sdim, int adim) {
  HostRolloutStage &S = g_stage;
  if (S.N == N && S.T >= T && S.sdim == sdim && S.adim == adim) return CRUX_OK;
  CRUX_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (S.a) { cudaFreeHost(S.a); cudaFreeHost(S.sp); cudaFreeHost(S.r); cudaFreeHost(S.robs); cudaFreeHost(S.done); cudaFreeHost(S.ee); }
  cudaEvent_t e0 = S.ev[0], e1 = S.ev[1];
  S = HostRolloutStage();
  S.ev[0] = e0; S.ev[1] = e1;
  CRUX_CHECK_CUDA(ctx, cudaMallocHost((void **)&S.a, (size_t)N * adim * sizeof(float)));
  CRUX_CHECK_CUDA(ctx, cudaMallocHost((void **)&S.sp, (size_t)T * N * sdim * sizeof(float)));
  CRUX_CHECK_CUDA(ctx, cudaMallocHost((void **)&S.robs, (size_t)N * sdim * sizeof(float)));
  CRUX_CHECK_CUDA(ctx, cudaMallocHost((void **)&S.r, (size_t)T * N * sizeof(float)));
  CRUX_CHECK_CUDA(ctx, cudaMallocHost((void **)&S.done, (size_t)T * N));
  CRUX_CHECK_CUDA(ctx, cudaMallocHost((void **)&S.ee, (size_t)T * N));
  for (int g = 0; g < 2; ++g)
    if (!S.ev[g]) CRUX_CHECK_CUDA(ctx, cudaEventCreateWithFlags(&S.ev[g], cudaEventDisableTiming));
  S.N = N; S.T = T; S.sdim = sdim; S.adim = adim;
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
  int rc = ensure_stage(ctx, N, T, sdim, adim);
  if (rc) return rc;
  HostRolloutStage &S = g_stage;
  cudaStream_t st = ctx->stream;
  // two leapfrogging halves when the policy supports split vector steps and the halves are worth a launch each
  const int G = (fused_rows_supported(actor) && N >= 512) ? 2 : 1;
  const int64_t lo[2] = {0, G == 2 ? N / 2 : N}, hi[2] = {G == 2 ? N / 2 : N, N};

  constexpr int UPLOAD_STEPS = 8;
  const bool mapped = G == 2;   // the fused policy shapes: the kernel does its own PCIe reads / writes
  // The host copies s' into the observation buffer the next forward reads (one memcpy per half step, ~0.2 ms per rollout).  Letting the
  // kernel read the s' rows where the env's worker threads left them (CRUX_ROLLOUT_DIRECT=1: staging rows + an episode_end flag per
  // row selecting the reset state) measured 30x SLOWER on the PCIe side (8.0 ms instead of 0.26 ms of event waits per 32-step
  // rollout, scripts/e2e_ab.py).  A second experiment had the env write into a ONE-step pinned buffer reused every step (the kernel
  // also stored the rows as the sp column, no upload): still 4.9 ms of waits.  So most of the cost is the device reading lines that
  // were last written by the env's 16 worker cores (the rest: a rotating 9 MB buffer instead of 278 KB); the main thread's copy
  // gathers them through one core first and the device then reads 278 KB that one core wrote.
  static const bool copy_obs = getenv("CRUX_ROLLOUT_DIRECT") == nullptr;
  auto enqueue_forward = [&](int g, int t) -> int {  // obs_g -> s[t]; policy forward; action back to the host; event
    const int64_t row = (int64_t)t * N + lo[g], n = hi[g] - lo[g];
    float *s_t = cols->s + row * sdim, *a_t = cols->a + row * adim;
    float *lp = cols->logprob ? cols->logprob + row : nullptr;
    if (mapped) {
      // step 0 reads the current observations; every later step reads the s' rows the env wrote for the previous vector step straight
      // from the rollout staging, except the streams whose episode ended there (episode_end flag): those read the reset state
      const bool direct = t > 0 && !copy_obs;
      const float *x = direct ? S.sp + ((size_t)(t - 1) * N + lo[g]) * sdim : obs_pinned + lo[g] * sdim;
      const uint8_t *flag = direct ? S.ee + (size_t)(t - 1) * N + lo[g] : nullptr;
      int rc2 = crux_rollout_step_rows_mapped(actor, x, n, lo[g], seed, ctr0 + (uint64_t)t, s_t, a_t, S.a + lo[g] * adim, lp, flag,
                                              obs_pinned + lo[g] * sdim);
      if (rc2) return rc2;
    } else {
      CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(s_t, obs_pinned + lo[g] * sdim, (size_t)n * sdim * sizeof(float), cudaMemcpyHostToDevice, st));
      int rc2 = crux_rollout_step(actor, nullptr, s_t, n, nullptr, seed, ctr0 + (uint64_t)t, a_t, lp, nullptr);
      if (rc2) return rc2;
      CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(S.a + lo[g] * adim, a_t, (size_t)n * adim * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    CRUX_CHECK_CUDA(ctx, cudaEventRecord(S.ev[g], st));
    return CRUX_OK;
  };
  auto upload = [&](int t0, int t1) -> int {   // transition columns of the vector steps [t0, t1): contiguous rows in the [T][N] layout
    const int64_t row = (int64_t)t0 * N, n = (int64_t)(t1 - t0) * N;
    CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(cols->sp + row * sdim, S.sp + row * sdim, (size_t)n * sdim * sizeof(float), cudaMemcpyHostToDevice, st));
    CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(cols->r + row, S.r + row, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, st));
    CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(cols->done + row, S.done + row, (size_t)n, cudaMemcpyHostToDevice, st));
    CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(cols->episode_end + row, S.ee + row, (size_t)n, cudaMemcpyHostToDevice, st));
    return CRUX_OK;
  };

  // CRUX_ROLLOUT_PROFILE=1: host wall-clock split of the loop (printed to stderr at the end of the call)
  static const bool prof = getenv("CRUX_ROLLOUT_PROFILE") != nullptr;
  double t_wait = 0, t_step = 0, t_book = 0, t_enq = 0;
  auto now = [] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  for (int g = 0; g < G; ++g) { rc = enqueue_forward(g, 0); if (rc) return rc; }
  int uploaded = 0;
  for (int t = 0; t < T; ++t) {
    const bool force = reset_at_end && t == T - 1;                                                     // steps!(reset=true) :148
    float *sp_t = S.sp + (size_t)t * N * sdim, *r_t = S.r + (size_t)t * N;
    uint8_t *done_t = S.done + (size_t)t * N, *ee_t = S.ee + (size_t)t * N;
    for (int g = 0; g < G; ++g) {
      const int64_t e0 = lo[g], e1 = hi[g], n = e1 - e0;
      const double q0 = prof ? now() : 0;
      CRUX_CHECK_CUDA(ctx, cudaEventSynchronize(S.ev[g]));   // actions of half g are on the host (and obs_g has been read)
      const double q1 = prof ? now() : 0;
      step(user, (int32_t)e0, (int32_t)e1, S.a, sp_t, r_t, done_t);                                    // @gen(:sp,:r), isterminal
      const double q2 = prof ? now() : 0;
      S.idx.clear();
      {   // branch-free pass (vectorises); the index list of ended streams is only built when there is one
        int32_t *len_p = episode_length + e0;
        const uint8_t *dn_p = done_t + e0;
        uint8_t *ee_p = ee_t + e0;
        const int32_t ms = max_steps, f = force ? 1 : 0;
        int any = 0;
        for (int64_t q = 0; q < n; ++q) {
          const int32_t len = len_p[q] + 1;                                                            // sampler.jl:130
          const int32_t end = (dn_p[q] != 0) | (len >= ms) | f;
          ee_p[q] = (uint8_t)end;
          len_p[q] = end ? 0 : len;
          any |= end;
        }
        if (any)
          for (int64_t e = e0; e < e1; ++e)
            if (ee_t[e]) S.idx.push_back((int32_t)e);
      }
      // next observation: sp, or a fresh initial state where the episode ended (terminate_episode! -> reset_sampler!).  With the
      // mapped forward the kernel picks sp / reset rows itself; obs_pinned is brought up to date once, after the last vector step
      if (!mapped || copy_obs || t + 1 == T) memcpy(obs_pinned + e0 * sdim, sp_t + e0 * sdim, (size_t)n * sdim * sizeof(float));
      if (!S.idx.empty()) {
        reset(user, S.idx.data(), (int32_t)S.idx.size(), S.robs);
        for (size_t q = 0; q < S.idx.size(); ++q) memcpy(obs_pinned + (size_t)S.idx[q] * sdim, S.robs + q * sdim, sizeof(float) * sdim);
      }
      // the forward of this half for the next vector step runs on the device while the host steps the other half
      const double q3 = prof ? now() : 0;
      if (t + 1 < T) { rc = enqueue_forward(g, t + 1); if (rc) return rc; }
      if (prof) { const double q4 = now(); t_wait += q1 - q0; t_step += q2 - q1; t_book += q3 - q2; t_enq += q4 - q3; }
    }
    if (t + 1 - uploaded >= UPLOAD_STEPS || t + 1 == T) { rc = upload(uploaded, t + 1); if (rc) return rc; uploaded = t + 1; }
  }
  CRUX_CHECK_CUDA(ctx, cudaStreamSynchronize(st));
  if (prof) fprintf(stderr, "crux_rollout_host: T=%d N=%lld  wait %.0f us  env step %.0f us  bookkeeping+memcpy %.0f us  enqueue %.0f us\n", (int)T,
                    (long long)N, t_wait, t_step, t_book, t_enq);
  return CRUX_OK;
}
