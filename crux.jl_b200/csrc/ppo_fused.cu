// Fused fast path for the PPO headline shapes (17-64-64-6 actor / 17-64-64-1 critic).
// Until a shape is handled here the generic layer-by-layer engine (mlp.cu / ppo.cu) runs.
#include "policy.cuh"

extern "C" int32_t crux_rollout_step_fused(crux_gaussian *actor, crux_mlp *critic, const float *obs, int64_t N, const float *eps_in,
                                           uint64_t seed, uint64_t ctr, float *a_out, float *logp_out, float *v_out, int *handled) {
  (void)actor; (void)critic; (void)obs; (void)N; (void)eps_in; (void)seed; (void)ctr; (void)a_out; (void)logp_out; (void)v_out;
  *handled = 0;
  return CRUX_OK;
}
