// Internal synthetic-env object.  Not part of the ABI.
#pragma once
#include "common.cuh"

#define LQ_MAX_S 32
#define LQ_MAX_A 16

struct crux_linquad {
  crux_ctx *ctx = nullptr;
  int sdim = 0, adim = 0, max_steps = 0;
  int64_t n_env = 0;
  uint64_t seed = 0;
  float *A = nullptr, *B = nullptr;  // device, row-major [sdim][sdim], [sdim][adim]
  int32_t *ep_len = nullptr;         // per-env episode length (sampler.episode_length)
  unsigned long long *tick = nullptr; // device: [0] global step counter -> Philox stream position, [1] finished-block counter
                                     // (device-resident so a captured CUDA graph draws fresh noise on every replay)
};

