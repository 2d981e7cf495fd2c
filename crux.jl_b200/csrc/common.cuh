// Shared internals of libcrux_cuda.so (sm_100a).  Not part of the ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <string>
#include <vector>
#include "../../include/crux_cuda.h"

#define CRUX_NUM_SMS_DEFAULT 148

// device-side sticky error flags
#define CRUX_FLAG_NAN 1u

struct crux_ctx {
  int device = 0;
  int num_sms = CRUX_NUM_SMS_DEFAULT;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaStream_t side_stream = nullptr;   // concurrent critic epochs of the fused PPO update (single GPU)
  cudaEvent_t side_fork = nullptr, side_done = nullptr;
  std::string err;
  int64_t launches = 0;
  unsigned int *flags_dev = nullptr;   // sticky device error flags
  unsigned int *flags_pinned = nullptr;
  // generic scratch (grown on demand, stream-ordered reuse)
  void *scratch[8] = {nullptr};
  size_t scratch_bytes[8] = {0};
  // TMA GAE path (gae_tma.cu): epoch-tagged carry flags + a never-reset work counter in scratch slot 7
  void *gae_scratch_seen = nullptr;
  unsigned int gae_epoch = 0, gae_ctr_base = 0;
  size_t gae_flag_cap = 0;
  // opt-in per-kernel-family device timing (bench.py roofline): event pairs recorded around selected launches
  bool timing = false;
  std::vector<cudaEvent_t> t_start, t_stop;
  std::vector<int> t_family;
  // multi-GPU
  int rank = 0, world = 1;
  void *nccl_comm = nullptr;
  void *nccl_comm_side = nullptr;       // second communicator (ncclCommSplit): collectives enqueued on side_stream (concurrent critic epochs)
  // peer (IPC) all-reduce state
  float *peer_recv = nullptr;            // [world][peer_cap] receive slots (local)
  unsigned long long *peer_flags = nullptr; // [world] arrival sequence numbers (local)
  float *peer_recv_remote[16] = {nullptr};
  unsigned long long *peer_flags_remote[16] = {nullptr};
  int64_t peer_cap = 0;
  unsigned long long *peer_seq_dev = nullptr;  // device-resident sequence number of the last completed peer exchange (local flags + 33):
                                               // kept on the device so that minibatches skipped by the device-side early stop do not advance it
  bool peer_ready = false;
  // LL (flag-in-data) gradient exchange fused into the reduce / Adam kernels: [2 networks][2 parities][16 ranks][peer_cap] 8-byte
  // words {float bits, sequence number} in every rank's peer allocation; device-resident per-network sequence numbers / tickets
  unsigned long long *peer_ll = nullptr;
  unsigned long long *peer_ll_remote[16] = {nullptr};
};

int crux_set_err(crux_ctx *ctx, int code, const char *fmt, ...);
void *crux_scratch(crux_ctx *ctx, int slot, size_t bytes);  // nullptr on failure (error set)

#define CRUX_CHECK_CUDA(ctx, call)                                                           \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess)                                                                  \
      return crux_set_err((ctx), CRUX_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call,   \
                          cudaGetErrorString(e__));                                          \
  } while (0)

#define CRUX_REQUIRE(ctx, cond, msg)                                                         \
  do {                                                                                       \
    if (!(cond)) return crux_set_err((ctx), CRUX_ERR_INVALID, "%s:%d %s", __FILE__, __LINE__, msg); \
  } while (0)

// count + check a kernel launch
#define CRUX_LAUNCHED(ctx)                                                                   \
  do {                                                                                       \
    (ctx)->launches++;                                                                       \
    cudaError_t e__ = cudaGetLastError();                                                    \
    if (e__ != cudaSuccess)                                                                  \
      return crux_set_err((ctx), CRUX_ERR_CUDA, "%s:%d launch: %s", __FILE__, __LINE__,      \
                          cudaGetErrorString(e__));                                          \
  } while (0)

// timing families (crux_ctx_timing_end)
enum { CRUX_T_MINIBATCH = 0, CRUX_T_REDUCE = 1, CRUX_T_ADAM = 2, CRUX_T_FORWARD = 3, CRUX_T_GAE = 4, CRUX_T_ENV = 5, CRUX_T_FAMILIES = 8 };
struct CruxTimed {  // RAII: records an event pair around a launch when timing is enabled
  crux_ctx *ctx; int slot;
  CruxTimed(crux_ctx *c, int family) : ctx(c), slot(-1) {
    if (!c->timing) return;
    cudaEvent_t a, b;
    if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
    c->t_start.push_back(a); c->t_stop.push_back(b); c->t_family.push_back(family);
    slot = (int)c->t_start.size() - 1;
    cudaEventRecord(a, c->stream);
  }
  ~CruxTimed() { if (slot >= 0) cudaEventRecord(ctx->t_stop[slot], ctx->stream); }
};

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t i64min(int64_t a, int64_t b) { return a < b ? a : b; }
static inline int64_t i64max(int64_t a, int64_t b) { return a > b ? a : b; }

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <int ACT>
__device__ __forceinline__ float act_fwd(float z) {
  if (ACT == CRUX_ACT_TANH) return tanhf(z);
  if (ACT == CRUX_ACT_RELU) return fmaxf(z, 0.0f);
  return z;
}
// derivative expressed through the OUTPUT y = act(z)
__device__ __forceinline__ float act_bwd_from_out(int act, float y) {
  if (act == CRUX_ACT_TANH) return 1.0f - y * y;
  if (act == CRUX_ACT_RELU) return y > 0.0f ? 1.0f : 0.0f;
  return 1.0f;
}
__device__ __forceinline__ float act_fwd_rt(int act, float z) {
  if (act == CRUX_ACT_TANH) return tanhf(z);
  if (act == CRUX_ACT_RELU) return fmaxf(z, 0.0f);
  return z;
}

// Philox4x32-10 counter RNG (production noise; parity runs pass noise through the ABI)
struct Philox4 { uint32_t x, y, z, w; };
__device__ __forceinline__ Philox4 philox4x32_10(uint64_t seed, uint64_t ctr_hi, uint64_t ctr_lo) {
  uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return Philox4{c0, c1, c2, c3};
}
__device__ __forceinline__ float u32_to_unit_open(uint32_t x) {  // (0,1]
  return ((float)(x >> 8) + 1.0f) * (1.0f / 16777216.0f);
}
__device__ __forceinline__ double u64_to_unit(uint32_t hi, uint32_t lo) {  // [0,1) 53-bit
  uint64_t v = (((uint64_t)hi << 32) | lo) >> 11;
  return (double)v * (1.0 / 9007199254740992.0);
}
// two standard normals from two uniforms (Box-Muller)
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float &n0, float &n1) {
  float u1 = u32_to_unit_open(a), u2 = u32_to_unit_open(b);
  float rad = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincospif(2.0f * u2, &s, &c);
  n0 = rad * c; n1 = rad * s;
}
#endif
