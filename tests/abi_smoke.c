/* abi_smoke.c -- drives rollout -> values -> GAE -> whiten -> PPO update -> replay buffer through include/crux_cuda.h from plain C,
 * exactly the way the Julia `ccall` signatures of julia/CruxB200.jl/src/abi.jl do (same argument order, same struct layouts).
 * Test infrastructure: built by __graft_entry__.build() (gcc, links libcrux_cuda.so), run by tests/test_abi_symbols.py (layout
 * mode, no GPU) and tests/test_gpu_abi_c.py (full mode; the Python test repeats the same sequence through ctypes and compares).
 *
 *   abi_smoke layout          prints sizeof / offsetof of every struct that crosses the ABI (checked against ctypes and the
 *                             numbers hard-coded in the Julia package's test)
 *   abi_smoke run N T seed    one PPO iteration on the device LinQuad env; prints checksums of the rollout columns, advantages,
 *                             the update info records and the updated parameters as "key value" lines
 */
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "crux_cuda.h"

#define CHECK(call)                                                                                                  \
  do {                                                                                                               \
    int32_t rc_ = (call);                                                                                            \
    if (rc_ != CRUX_OK) {                                                                                            \
      fprintf(stderr, "%s failed: status %d: %s\n", #call, (int)rc_, crux_last_error(ctx));                          \
      return 1;                                                                                                      \
    }                                                                                                                \
  } while (0)

static int layout(void) {
  printf("crux_ppo_hp %zu eps_clip %zu target_kl %zu a2c %zu actor_epochs %zu actor_batch %zu critic_epochs %zu critic_batch %zu actor_max_batches %zu "
         "critic_max_batches %zu\n",
         sizeof(crux_ppo_hp), offsetof(crux_ppo_hp, eps_clip), offsetof(crux_ppo_hp, target_kl), offsetof(crux_ppo_hp, a2c),
         offsetof(crux_ppo_hp, actor_epochs), offsetof(crux_ppo_hp, actor_batch), offsetof(crux_ppo_hp, critic_epochs),
         offsetof(crux_ppo_hp, critic_batch), offsetof(crux_ppo_hp, actor_max_batches), offsetof(crux_ppo_hp, critic_max_batches));
  printf("crux_lagrange_hp %zu target_cost %zu Kd %zu ema_alpha %zu cost_epochs %zu cost_batch %zu cost_max_batches %zu\n", sizeof(crux_lagrange_hp),
         offsetof(crux_lagrange_hp, target_cost), offsetof(crux_lagrange_hp, Kd), offsetof(crux_lagrange_hp, ema_alpha),
         offsetof(crux_lagrange_hp, cost_epochs), offsetof(crux_lagrange_hp, cost_batch), offsetof(crux_lagrange_hp, cost_max_batches));
  printf("crux_rollout_cols %zu s %zu a %zu sp %zu r %zu done %zu episode_end %zu logprob %zu\n", sizeof(crux_rollout_cols),
         offsetof(crux_rollout_cols, s), offsetof(crux_rollout_cols, a), offsetof(crux_rollout_cols, sp), offsetof(crux_rollout_cols, r),
         offsetof(crux_rollout_cols, done), offsetof(crux_rollout_cols, episode_end), offsetof(crux_rollout_cols, logprob));
  printf("crux_col_desc %zu id %zu dtype %zu rowlen %zu init %zu\n", sizeof(crux_col_desc), offsetof(crux_col_desc, id),
         offsetof(crux_col_desc, dtype), offsetof(crux_col_desc, rowlen), offsetof(crux_col_desc, init));
  printf("abi_version %d\n", (int)crux_abi_version());
  return 0;
}

/* deterministic parameter / matrix fill shared with the Python twin: x_k = sin(0.37 k + phase) * scale */
static void fill(float *x, int64_t n, double phase, double scale) {
  for (int64_t k = 0; k < n; ++k) x[k] = (float)(sin(0.37 * (double)k + phase) * scale);
}
static double checksum_f32(crux_ctx *ctx, const float *dev, int64_t n, float *host) {
  if (crux_memcpy_d2h(ctx, host, dev, (size_t)n * sizeof(float)) || crux_ctx_sync(ctx)) return NAN;
  double s = 0;
  for (int64_t k = 0; k < n; ++k) s += (double)host[k] * (double)((k % 7) + 1);
  return s;
}
static double checksum_u8(crux_ctx *ctx, const uint8_t *dev, int64_t n, uint8_t *host) {
  if (crux_memcpy_d2h(ctx, host, dev, (size_t)n) || crux_ctx_sync(ctx)) return NAN;
  double s = 0;
  for (int64_t k = 0; k < n; ++k) s += (double)host[k] * (double)((k % 5) + 1);
  return s;
}

static int run(int64_t N, int32_t T, uint64_t seed) {
  enum { SDIM = 17, ADIM = 6, HID = 64 };
  crux_ctx *ctx = NULL;
  CHECK(crux_ctx_create(0, NULL, &ctx));
  const int64_t n = N * T;

  /* actor 17-64-64-6 tanh + logSigma vector, critic 17-64-64-1 (Flux.params order: W [in][out] row-major, then b) */
  const int32_t dims_a[4] = {SDIM, HID, HID, ADIM}, dims_c[4] = {SDIM, HID, HID, 1}, acts[3] = {CRUX_ACT_TANH, CRUX_ACT_TANH, CRUX_ACT_IDENTITY};
  crux_mlp *mu = NULL, *V = NULL;
  CHECK(crux_mlp_create(ctx, 3, dims_a, acts, &mu));
  CHECK(crux_mlp_create(ctx, 3, dims_c, acts, &V));
  int64_t pa = 0, pc = 0;
  CHECK(crux_mlp_num_params(mu, &pa));
  CHECK(crux_mlp_num_params(V, &pc));
  float *flat = (float *)malloc((size_t)(pa > pc ? pa : pc) * sizeof(float));
  fill(flat, pa, 0.1, 0.2);
  CHECK(crux_mlp_set_params(mu, flat));
  fill(flat, pc, 1.3, 0.2);
  CHECK(crux_mlp_set_params(V, flat));
  CHECK(crux_mlp_set_adam(mu, 3e-4, 0.9, 0.999, 1e-8));
  CHECK(crux_mlp_set_adam(V, 3e-4, 0.9, 0.999, 1e-8));
  float ls[ADIM];
  for (int j = 0; j < ADIM; ++j) ls[j] = -0.5f;
  crux_gaussian *pi = NULL;
  CHECK(crux_gaussian_create(ctx, mu, ADIM, ls, 0, 1.0f, &pi));

  /* device LinQuad env */
  float A[SDIM * SDIM], B[SDIM * ADIM];
  fill(A, SDIM * SDIM, 0.5, 0.02);
  for (int i = 0; i < SDIM; ++i) A[i * SDIM + i] += 0.95f;
  fill(B, SDIM * ADIM, 2.1, 0.1);
  crux_linquad *env = NULL;
  CHECK(crux_linquad_create(ctx, SDIM, ADIM, A, B, N, 1000, seed + 7, &env));

  /* rollout buffer = an ExperienceBuffer of capacity n with the PPO columns; the rollout is written in place */
  enum { C_S = 0, C_A, C_SP, C_R, C_DONE, C_EE, C_LOGP, C_ADV, C_RET };
  const crux_col_desc cols[9] = {{C_S, CRUX_F32, SDIM, 0}, {C_A, CRUX_F32, ADIM, 0}, {C_SP, CRUX_F32, SDIM, 0}, {C_R, CRUX_F32, 1, 0}, {C_DONE, CRUX_U8, 1, 0},
                                 {C_EE, CRUX_U8, 1, 0},    {C_LOGP, CRUX_F32, 1, 0}, {C_ADV, CRUX_F32, 1, 0},   {C_RET, CRUX_F32, 1, 0}};
  crux_buffer *buf = NULL;
  CHECK(crux_buffer_create(ctx, n, 9, cols, 0, 0.6f, &buf));
  void *p[9];
  for (int k = 0; k < 9; ++k) {
    int64_t rowlen; int32_t dt;
    CHECK(crux_buffer_col(buf, k, &p[k], &rowlen, &dt));
    if (rowlen != cols[k].rowlen || dt != cols[k].dtype) { fprintf(stderr, "crux_buffer_col: column %d describes itself wrongly\n", k); return 1; }
  }
  crux_rollout_cols rc = {(float *)p[C_S], (float *)p[C_A], (float *)p[C_SP], (float *)p[C_R], (uint8_t *)p[C_DONE], (uint8_t *)p[C_EE], (float *)p[C_LOGP]};
  float *obs = NULL, *v_s = NULL, *v_sp = NULL;
  CHECK(crux_dev_alloc(ctx, (size_t)N * SDIM * sizeof(float), (void **)&obs));
  CHECK(crux_dev_alloc(ctx, (size_t)n * sizeof(float), (void **)&v_s));
  CHECK(crux_dev_alloc(ctx, (size_t)n * sizeof(float), (void **)&v_sp));
  CHECK(crux_linquad_reset(env, obs));

  /* steps!(reset=true): T vector steps in one launch; then value(V, s), value(V, sp), fill_gae! + fill_returns!, whiten */
  CHECK(crux_linquad_rollout(env, pi, T, 1, obs, &rc, seed, 0));
  CHECK(crux_buffer_push(buf, n, 0, NULL, NULL, 0, NULL, NULL)); /* zero columns: only the ring bookkeeping advances (in-place rollout) */
  CHECK(crux_mlp_forward(V, rc.s, n, v_s));
  CHECK(crux_value_next(V, rc.sp, rc.s, v_s, T, N, v_sp));
  CHECK(crux_fill_gae_returns(ctx, rc.r, rc.done, rc.episode_end, v_s, v_sp, T, N, 0.99f, 0.95f, (float *)p[C_ADV], (float *)p[C_RET]));
  float *host = (float *)malloc((size_t)n * SDIM * sizeof(float));
  printf("sum_s %.9e\n", checksum_f32(ctx, rc.s, n * SDIM, host));
  printf("sum_a %.9e\n", checksum_f32(ctx, rc.a, n * ADIM, host));
  printf("sum_r %.9e\n", checksum_f32(ctx, rc.r, n, host));
  printf("sum_logp %.9e\n", checksum_f32(ctx, rc.logprob, n, host));
  printf("sum_ee %.9e\n", checksum_u8(ctx, rc.episode_end, n, (uint8_t *)host));
  printf("sum_adv %.9e\n", checksum_f32(ctx, (float *)p[C_ADV], n, host));
  printf("sum_ret %.9e\n", checksum_f32(ctx, (float *)p[C_RET], n, host));
  CHECK(crux_whiten(ctx, (float *)p[C_ADV], n));

  /* policy_gradient_training: 2 epochs x minibatches of n/2 with caller-chosen row orders (order[e][i] = (i * 7 + e) mod n, 7 coprime to n) */
  crux_ppo_hp hp;
  memset(&hp, 0, sizeof(hp));
  hp.eps_clip = 0.2f; hp.lambda_p = 1.0f; hp.lambda_e = 0.1f; hp.target_kl = INFINITY; hp.a2c = 0;
  hp.actor_epochs = 2; hp.actor_batch = (int32_t)(n / 2); hp.critic_epochs = 2; hp.critic_batch = (int32_t)(n / 2);
  int32_t *order_h = (int32_t *)malloc((size_t)2 * n * sizeof(int32_t)), *order = NULL;
  for (int e = 0; e < 2; ++e)
    for (int64_t i = 0; i < n; ++i) order_h[e * n + i] = (int32_t)((i * 7 + e) % n);
  CHECK(crux_dev_alloc(ctx, (size_t)2 * n * sizeof(int32_t), (void **)&order));
  CHECK(crux_memcpy_h2d(ctx, order, order_h, (size_t)2 * n * sizeof(int32_t)));
  float ia[4 * CRUX_PPO_INFO_STRIDE], ic[4 * CRUX_PPO_INFO_STRIDE];
  CHECK(crux_ppo_update(pi, V, rc.s, rc.a, rc.logprob, (float *)p[C_ADV], (float *)p[C_RET], n, &hp, order, order, seed, ia, ic));
  for (int m = 0; m < 4; ++m)
    printf("actor_mb%d loss %.9e grad_norm %.9e entropy %.9e kl %.9e clip %.9e valid %g\n", m, ia[m * 8 + CRUX_PPO_LOSS], ia[m * 8 + CRUX_PPO_GRAD_NORM],
           ia[m * 8 + CRUX_PPO_ENTROPY], ia[m * 8 + CRUX_PPO_KL], ia[m * 8 + CRUX_PPO_CLIP_FRAC], ia[m * 8 + CRUX_PPO_VALID]);
  for (int m = 0; m < 4; ++m) printf("critic_mb%d loss %.9e grad_norm %.9e valid %g\n", m, ic[m * 8 + CRUX_PPO_LOSS], ic[m * 8 + CRUX_PPO_GRAD_NORM], ic[m * 8 + CRUX_PPO_VALID]);
  CHECK(crux_mlp_get_params(mu, flat));
  double s = 0;
  for (int64_t k = 0; k < pa; ++k) s += (double)flat[k] * (double)((k % 7) + 1);
  printf("sum_actor_params %.9e\n", s);
  CHECK(crux_mlp_get_params(V, flat));
  s = 0;
  for (int64_t k = 0; k < pc; ++k) s += (double)flat[k] * (double)((k % 7) + 1);
  printf("sum_critic_params %.9e\n", s);

  /* ring bookkeeping as the Julia ExperienceBuffer wrapper reads it */
  int64_t elements, next_ind, total, cap;
  CHECK(crux_buffer_state(buf, &elements, &next_ind, &total, &cap));
  printf("buffer elements %lld next_ind %lld total %lld capacity %lld\n", (long long)elements, (long long)next_ind, (long long)total, (long long)cap);
  CHECK(crux_ctx_check(ctx));
  int64_t launches = 0;
  CHECK(crux_ctx_launch_count(ctx, &launches));
  printf("launches %lld\n", (long long)launches);

  crux_dev_free(ctx, order); crux_dev_free(ctx, obs); crux_dev_free(ctx, v_s); crux_dev_free(ctx, v_sp);
  crux_buffer_destroy(buf); crux_linquad_destroy(env); crux_gaussian_destroy(pi); crux_mlp_destroy(mu); crux_mlp_destroy(V);
  crux_ctx_destroy(ctx);
  free(flat); free(host); free(order_h);
  return 0;
}

int main(int argc, char **argv) {
  if (argc >= 2 && strcmp(argv[1], "layout") == 0) return layout();
  if (argc >= 5 && strcmp(argv[1], "run") == 0) return run(atoll(argv[2]), atoi(argv[3]), strtoull(argv[4], NULL, 10));
  fprintf(stderr, "usage: %s layout | run N T seed\n", argv[0]);
  return 2;
}
