// Fused fast path for the PPO headline shapes: Chain(Dense(I,64,act), Dense(64,64,act), Dense(64,O)) with I <= 32, O <= 8
// (actor 17-64-64-6 + logΣ vector, critic 17-64-64-1 of BASELINE config[1]).  Other shapes run on the generic
// layer-by-layer engine (mlp.cu / ppo.cu).
//
//   fused_forward_kernel    hot path (i)  : one launch = obs tile -> 3 Dense layers -> Gaussian sample + logprob (actor) and
//                                           V(s) (critic) for every env stream of a vector step (sampler.jl:73), or value(V, s)
//                                           over a whole rollout column (the two critic passes that feed the GAE scan).
//   fused_minibatch_kernel  hot path (iii): one launch = gather minibatch rows by shuffled index -> forward -> ppo_loss /
//                                           a2c_loss / mse terms -> analytic backward -> per-CTA weight-gradient partials
//                                           (rl/ppo.jl:4-21, rl/a2c.jl:4-16, ppo.jl:60, training.jl:15-18).  No activation
//                                           ever touches HBM: per row only s, a, logprob, advantage, return are read.
//
// Data layout inside a CTA (256 threads, a tile of R = 64 rows): activations live TRANSPOSED in shared memory,
// actT[feature][row] with the row contiguous (stride R+4), so that every GEMM of the forward and backward pass is a
// register-tiled 4x4 outer product fed by two 128-bit shared loads per 16 FFMA:
//     forward      C^T[j][r]  = sum_k A^T[k][r] W[k][j]          thread = (4 rows, 4 cols)
//     data bwd     dA^T[k][r] = sum_j dC^T[j][r] W^T[j][k]        thread = (4 rows, 4 cols)   (W^T kept in smem too)
//     weight bwd   dW[k][j]  += sum_r A^T[k][r] dC^T[j][r]        thread = (4 k, 4 j) strided by 16, r vectorised by 4;
//                                                                  accumulators stay in registers across all tiles of the CTA
// The flat parameter vector (Flux.params order, W_l row-major [in][out]) is staged into shared memory with ONE TMA bulk
// copy (cp.async.bulk ... mbarrier::complete_tx) per CTA.
// Bound: fp32 FFMA (SIMT).  1e-5 parity with the fp32 reference excludes TF32/BF16 tensor-core MMA for these layers.
#include "policy.cuh"
#include <stdlib.h>

namespace {

constexpr int H = 64;        // hidden width
constexpr int R = 64;        // rows per tile
constexpr int LD = R + 4;    // row stride of transposed tiles (floats); keeps 128-bit alignment
constexpr int NT = 256;      // threads per CTA
constexpr int MAX_I = 32, MAX_O = 8;
constexpr int P_MAX = MAX_I * H + H + H * H + H + H * MAX_O + MAX_O;  // 6792 floats
constexpr int P_SMEM = (P_MAX + 3) / 4 * 4 + 8;

#define LOG_SQRT_2PI 0.9189385332046727f

// ---- shared memory carve-up (floats) ----------------------------------------------------------------------------------------
struct SmemMap {
  static constexpr int P = 0;                       // raw params
  static constexpr int W2T = P + P_SMEM;            // [64][64]  W2T[j][k] = W2[k][j]
  static constexpr int W3T = W2T + H * H;           // [8][64]   W3T[o][k] = W3[k][o]
  static constexpr int XT = W3T + MAX_O * H;        // [32][LD]
  static constexpr int H1T = XT + MAX_I * LD;       // [64][LD]
  static constexpr int H2T = H1T + H * LD;          // [64][LD]
  static constexpr int OT = H2T + H * LD;           // [8][LD]  outputs, then dL/dout
  static constexpr int AT = OT + MAX_O * LD;        // [8][LD]  stored actions
  static constexpr int LP = AT + MAX_O * LD;        // [LD] old logprob
  static constexpr int ADV = LP + LD;               // [LD]
  static constexpr int RET = ADV + LD;              // [LD]
  static constexpr int IDX = RET + LD;              // [R] ints
  static constexpr int RED = IDX + R;               // [8][24] reduction scratch
  static constexpr int MBAR = RED + 8 * 24;         // 2 floats = one 64-bit mbarrier (8-byte aligned: all offsets are even)
  static constexpr int TOTAL = MBAR + 2;
};
static_assert(SmemMap::MBAR % 2 == 0, "mbarrier must be 8-byte aligned");
static_assert(SmemMap::W2T % 4 == 0 && SmemMap::XT % 4 == 0 && SmemMap::H1T % 4 == 0 && SmemMap::OT % 4 == 0, "16-byte alignment");
constexpr size_t SMEM_BYTES = (size_t)SmemMap::TOTAL * sizeof(float);

struct NetDesc {
  const float *params;  // device, flat
  int I, O, act;        // dims [I, 64, 64, O]; hidden activation
  uint32_t bytes16;     // parameter bytes rounded up to 16 (the allocation is padded)
};
__device__ __forceinline__ int off_b1(int I) { return I * H; }
__device__ __forceinline__ int off_W2(int I) { return I * H + H; }
__device__ __forceinline__ int off_b2(int I) { return I * H + H + H * H; }
__device__ __forceinline__ int off_W3(int I) { return I * H + H + H * H + H; }
__device__ __forceinline__ int off_b3(int I, int O) { return I * H + H + H * H + H + H * O; }

// ---- TMA bulk copy of the parameter vector into shared memory ------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void stage_params(float *sm, const NetDesc &nd) {
  const uint32_t mbar = smem_u32(sm + SmemMap::MBAR);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(1) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(nd.bytes16) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm + SmemMap::P)),
                 "l"(nd.params), "r"(nd.bytes16), "r"(mbar)
                 : "memory");
  }
  uint32_t done = 0;
  while (!done) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0, 1, 0, p; }"
                 : "=r"(done)
                 : "r"(mbar), "r"(0)
                 : "memory");
  }
}

// transposed copies used by the data-backward GEMMs
__device__ __forceinline__ void build_transposes(float *sm, int I, int O) {
  const float *W2 = sm + SmemMap::P + off_W2(I), *W3 = sm + SmemMap::P + off_W3(I);
  for (int e = threadIdx.x; e < H * H; e += NT) {
    const int j = e >> 6, k = e & 63;
    sm[SmemMap::W2T + e] = W2[k * H + j];
  }
  for (int e = threadIdx.x; e < MAX_O * H; e += NT) {
    const int o = e >> 6, k = e & 63;
    sm[SmemMap::W3T + e] = o < O ? W3[k * O + o] : 0.f;
  }
}

// ---- register-tiled GEMM pieces --------------------------------------------------------------------------------------------
// C^T[j][r] = act(b[j] + sum_{k<K} A^T[k][r] W[k][j]);  thread = rows 4rg..4rg+3, cols 4jg..4jg+3
__device__ __forceinline__ void layer_fwd(const float *__restrict__ AT, int K, const float *__restrict__ W, const float *__restrict__ b,
                                          float *__restrict__ CT, int act) {
  const int rg = threadIdx.x >> 4, jg = threadIdx.x & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const float *ap = AT + 4 * rg, *wp = W + 4 * jg;
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const float4 a = *reinterpret_cast<const float4 *>(ap + k * LD);
    const float4 w = *reinterpret_cast<const float4 *>(wp + k * H);
    const float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
  }
  const float4 bb = *reinterpret_cast<const float4 *>(b + 4 * jg);
  const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float4 v;
    v.x = act_fwd_rt(act, acc[0][j] + bv[j]);
    v.y = act_fwd_rt(act, acc[1][j] + bv[j]);
    v.z = act_fwd_rt(act, acc[2][j] + bv[j]);
    v.w = act_fwd_rt(act, acc[3][j] + bv[j]);
    *reinterpret_cast<float4 *>(CT + (4 * jg + j) * LD + 4 * rg) = v;
  }
}

// out^T[o][r] = b3[o] + sum_k h2^T[k][r] W3[k][o];  thread = (row t&63, outputs og and og+4)
__device__ __forceinline__ void layer_out(const float *__restrict__ H2T, const float *__restrict__ W3, const float *__restrict__ b3, int O,
                                          float *__restrict__ OT) {
  const int r = threadIdx.x & 63, og = threadIdx.x >> 6;
  const bool v0 = og < O, v1 = og + 4 < O;
  if (!v0) return;
  float a0 = b3[og], a1 = v1 ? b3[og + 4] : 0.f;
#pragma unroll 8
  for (int k = 0; k < H; ++k) {
    const float h = H2T[k * LD + r];
    a0 = fmaf(h, W3[k * O + og], a0);
    if (v1) a1 = fmaf(h, W3[k * O + og + 4], a1);
  }
  OT[og * LD + r] = a0;
  if (v1) OT[(og + 4) * LD + r] = a1;
}

// dA^T[k][r] = act'(A^T[k][r]) * sum_{j<J} dC^T[j][r] WT[j][k]   in place over A^T;  thread = rows 4rg.., cols 4kg..
__device__ __forceinline__ void layer_bwd_data(const float *__restrict__ DCT, int J, const float *__restrict__ WT, float *__restrict__ AT, int act) {
  const int rg = threadIdx.x >> 4, kg = threadIdx.x & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const float *dp = DCT + 4 * rg, *wp = WT + 4 * kg;
#pragma unroll 4
  for (int j = 0; j < J; ++j) {
    const float4 d = *reinterpret_cast<const float4 *>(dp + j * LD);
    const float4 w = *reinterpret_cast<const float4 *>(wp + j * H);
    const float dv[4] = {d.x, d.y, d.z, d.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[i][c] = fmaf(dv[i], wv[c], acc[i][c]);
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float4 *p = reinterpret_cast<float4 *>(AT + (4 * kg + c) * LD + 4 * rg);
    const float4 y = *p;
    float4 v;
    v.x = acc[0][c] * act_bwd_from_out(act, y.x);
    v.y = acc[1][c] * act_bwd_from_out(act, y.y);
    v.z = acc[2][c] * act_bwd_from_out(act, y.z);
    v.w = acc[3][c] * act_bwd_from_out(act, y.w);
    *p = v;
  }
}

__device__ __forceinline__ float dot4(const float4 &a, const float4 &b, float acc) {
  acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
  return acc;
}

// ---- tile loads --------------------------------------------------------------------------------------------------------------
// x^T[i][r] = x[row(r)][i]; rows beyond n are zero.  idx (shared) holds the source row or -1.
__device__ __forceinline__ void load_rows_T(float *__restrict__ XT, const float *__restrict__ x, const int *__restrict__ sidx, int I) {
  for (int e = threadIdx.x; e < R * I; e += NT) {
    const int r = e / I, i = e - r * I;
    const int row = sidx[r];
    XT[i * LD + r] = row >= 0 ? __ldg(x + (int64_t)row * I + i) : 0.f;
  }
}

// =================================================================================================== forward kernel
struct FwdArgs {
  NetDesc net[2];            // [0] = first network, [1] = optional second network (critic) on blockIdx.y == 1
  int mode[2];               // 0: write outputs y[row][O]; 1: Gaussian explore head (a, logprob)
  const float *x;            // [B][I]
  int64_t B;
  float *y[2];               // mode 0 output / mode 1 action output
  float *logp;               // mode 1
  const float *ls;           // mode 1: logΣ vector
  const float *eps_in;       // mode 1: injected noise or NULL
  uint64_t seed, ctr;
};

__global__ void __launch_bounds__(NT, 2) fused_forward_kernel(FwdArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int which = blockIdx.y;
  const NetDesc nd = a.net[which];
  const int I = nd.I, O = nd.O;
  stage_params(sm, nd);
  int *sidx = reinterpret_cast<int *>(sm + SmemMap::IDX);
  const float *P = sm + SmemMap::P;
  const int64_t n_tiles = (a.B + R - 1) / R;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    __syncthreads();
    if (threadIdx.x < R) {
      const int64_t row = tile * R + threadIdx.x;
      sidx[threadIdx.x] = row < a.B ? (int)row : -1;
    }
    __syncthreads();
    load_rows_T(sm + SmemMap::XT, a.x, sidx, I);
    __syncthreads();
    layer_fwd(sm + SmemMap::XT, I, P, P + off_b1(I), sm + SmemMap::H1T, nd.act);
    __syncthreads();
    layer_fwd(sm + SmemMap::H1T, H, P + off_W2(I), P + off_b2(I), sm + SmemMap::H2T, nd.act);
    __syncthreads();
    layer_out(sm + SmemMap::H2T, P + off_W3(I), P + off_b3(I, O), O, sm + SmemMap::OT);
    __syncthreads();
    const float *OT = sm + SmemMap::OT;
    if (a.mode[which] == 0) {
      float *y = a.y[which];
      for (int e = threadIdx.x; e < R * O; e += NT) {
        const int r = e / O, o = e - r * O;
        if (sidx[r] >= 0) y[(int64_t)sidx[r] * O + o] = OT[o * LD + r];
      }
    } else if (threadIdx.x < R && sidx[threadIdx.x] >= 0) {
      // exploration(::GaussianPolicy) policies.jl:338-344 + gaussian_logpdf :333-336 (same expression order as policy.cu)
      const int r = threadIdx.x;
      const int64_t i = sidx[r];
      float logp = 0.f, nrm[4];
      for (int j = 0; j < O; ++j) {
        const float mu = OT[j * LD + r];
        const float ls = a.ls[j];
        const float sigma = expf(ls);
        const float var = sigma * sigma;
        float e;
        if (a.eps_in) e = a.eps_in[i * O + j];
        else {
          if ((j & 3) == 0) {
            const Philox4 p = philox4x32_10(a.seed, a.ctr, (uint64_t)i * ((O + 3) / 4) + (j >> 2));
            box_muller(p.x, p.y, nrm[0], nrm[1]);
            box_muller(p.z, p.w, nrm[2], nrm[3]);
          }
          e = nrm[j & 3];
        }
        const float act = e * sigma + mu;
        a.y[which][i * O + j] = act;
        const float d = act - mu;
        logp += -(d * d) / (2.f * var) - LOG_SQRT_2PI - ls;
      }
      if (a.logp) a.logp[i] = logp;
    }
  }
}

// =================================================================================================== minibatch kernel
struct MbArgs {
  NetDesc net;
  const float *s, *act, *logp_old, *adv, *ret;   // full columns (gathered by index)
  const int32_t *order;                           // minibatch row ids (order + offset) or NULL for identity
  int64_t bm;                                     // rows in this minibatch
  const float *ls;                                // actor: logΣ vector
  float inv_bg, eps_clip, lambda_p;
  int a2c;
  float *partials;                                // [gridDim.x][pstride]
  int pstride;                                    // >= n_params + 16
  int n_params;
  const int *skip;
};

// HEAD 0: ppo_loss / a2c_loss on a GaussianPolicy with a logΣ vector.  HEAD 1: Flux.mse(V(s), return).
template <int HEAD>
__global__ void __launch_bounds__(NT, 2) fused_minibatch_kernel(MbArgs a) {
  if (a.skip && *a.skip) return;
  extern __shared__ __align__(16) float sm[];
  const NetDesc nd = a.net;
  const int I = nd.I, O = nd.O, act = nd.act;
  const int t = threadIdx.x;
  stage_params(sm, nd);
  build_transposes(sm, I, O);
  int *sidx = reinterpret_cast<int *>(sm + SmemMap::IDX);
  const float *P = sm + SmemMap::P;
  float *XT = sm + SmemMap::XT, *H1T = sm + SmemMap::H1T, *H2T = sm + SmemMap::H2T, *OT = sm + SmemMap::OT, *AT = sm + SmemMap::AT;

  // per-CTA gradient accumulators (registers, live across all tiles)
  const int kg = t >> 4, jg = t & 15;  // weight-gradient patch: rows kg + 16a, cols jg + 16b
  float acc2[4][4], acc1[2][4], acc3[2] = {0.f, 0.f}, accb = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc2[i][j] = 0.f;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc1[i][j] = 0.f;
  // head sums of this thread's rows (threads 0..63)
  float s_obj = 0.f, s_kl = 0.f, s_clip = 0.f, s_adv = 0.f, s_ret = 0.f, dls[MAX_O];
#pragma unroll
  for (int j = 0; j < MAX_O; ++j) dls[j] = 0.f;

  const int64_t n_tiles = (a.bm + R - 1) / R;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    __syncthreads();
    if (t < R) {
      const int64_t row = tile * R + t;
      sidx[t] = row < a.bm ? (a.order ? a.order[row] : (int)row) : -1;
    }
    __syncthreads();
    load_rows_T(XT, a.s, sidx, I);
    if (HEAD == 0) {
      load_rows_T(AT, a.act, sidx, O);
      if (t < R) {
        const int row = sidx[t];
        sm[SmemMap::LP + t] = row >= 0 ? a.logp_old[row] : 0.f;
        sm[SmemMap::ADV + t] = row >= 0 ? a.adv[row] : 0.f;
        sm[SmemMap::RET + t] = (row >= 0 && a.ret) ? a.ret[row] : 0.f;
      }
    } else if (t < R) {
      const int row = sidx[t];
      sm[SmemMap::RET + t] = row >= 0 ? a.ret[row] : 0.f;
    }
    __syncthreads();
    // ---------------- forward
    layer_fwd(XT, I, P, P + off_b1(I), H1T, act);
    __syncthreads();
    layer_fwd(H1T, H, P + off_W2(I), P + off_b2(I), H2T, act);
    __syncthreads();
    layer_out(H2T, P + off_W3(I), P + off_b3(I, O), O, OT);
    __syncthreads();
    // ---------------- loss head: dL/dout (scaled by 1/B_global) replaces out^T
    if (t < R) {
      const bool live = sidx[t] >= 0;
      if (HEAD == 0) {
        float logp = 0.f;
        for (int j = 0; j < O; ++j) {
          const float sg = expf(a.ls[j]);
          const float d = AT[j * LD + t] - OT[j * LD + t];
          logp += -(d * d) / (2.f * (sg * sg)) - LOG_SQRT_2PI - a.ls[j];
        }
        const float Ai = sm[SmemMap::ADV + t], old = sm[SmemMap::LP + t];
        float dlogp = 0.f;
        if (live) {
          if (a.a2c) {
            s_obj += logp * Ai;
            dlogp = -a.lambda_p * a.inv_bg * Ai;
          } else {
            const float rt = expf(logp - old);
            const float lo = 1.f - a.eps_clip, hi = 1.f + a.eps_clip;
            const float x = rt * Ai, y = fminf(fmaxf(rt, lo), hi) * Ai;
            const bool first = !(y < x);  // min(x, y) keeps x on ties
            s_obj += first ? x : y;
            dlogp = first ? -a.lambda_p * a.inv_bg * x : 0.f;
            s_clip += (rt > hi || rt < lo) ? 1.f : 0.f;
          }
          s_kl += old - logp; s_adv += Ai; s_ret += sm[SmemMap::RET + t];
        }
        for (int j = 0; j < O; ++j) {
          const float sg = expf(a.ls[j]);
          const float var = sg * sg;
          const float d = AT[j * LD + t] - OT[j * LD + t];
          OT[j * LD + t] = dlogp * d / var;
          dls[j] += dlogp * (d * d / var - 1.f);
        }
      } else {
        const float d = OT[t] - sm[SmemMap::RET + t];
        if (live) s_obj += d * d;
        OT[t] = live ? 2.f * d * a.inv_bg : 0.f;
      }
    }
    __syncthreads();
    // ---------------- dW3 += h2^T dOut ; db3
    {
      const int k = t & 63, og = t >> 6;
      if (og < O) {
        const bool v1 = og + 4 < O;
#pragma unroll 4
        for (int r4 = 0; r4 < R / 4; ++r4) {
          const float4 h = *reinterpret_cast<const float4 *>(H2T + k * LD + 4 * r4);
          acc3[0] = dot4(h, *reinterpret_cast<const float4 *>(OT + og * LD + 4 * r4), acc3[0]);
          if (v1) acc3[1] = dot4(h, *reinterpret_cast<const float4 *>(OT + (og + 4) * LD + 4 * r4), acc3[1]);
        }
      }
      if (t >= 128 && t < 128 + O) {  // db3
        const int o = t - 128;
        for (int r4 = 0; r4 < R / 4; ++r4) {
          const float4 d = *reinterpret_cast<const float4 *>(OT + o * LD + 4 * r4);
          accb += (d.x + d.y) + (d.z + d.w);
        }
      }
    }
    __syncthreads();
    // ---------------- dz2^T in place over h2^T
    layer_bwd_data(OT, O, sm + SmemMap::W3T, H2T, act);
    __syncthreads();
    // ---------------- dW2 += h1^T dz2 ; db2
#pragma unroll 2
    for (int r4 = 0; r4 < R / 4; ++r4) {
      float4 hv[4], zv[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        hv[q] = *reinterpret_cast<const float4 *>(H1T + (kg + 16 * q) * LD + 4 * r4);
        zv[q] = *reinterpret_cast<const float4 *>(H2T + (jg + 16 * q) * LD + 4 * r4);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc2[i][j] = dot4(hv[i], zv[j], acc2[i][j]);
    }
    if (t < 64) {  // db2[t]
      for (int r4 = 0; r4 < R / 4; ++r4) {
        const float4 d = *reinterpret_cast<const float4 *>(H2T + t * LD + 4 * r4);
        accb += (d.x + d.y) + (d.z + d.w);
      }
    }
    __syncthreads();
    // ---------------- dz1^T in place over h1^T
    layer_bwd_data(H2T, H, sm + SmemMap::W2T, H1T, act);
    __syncthreads();
    // ---------------- dW1 += x^T dz1 ; db1
#pragma unroll 2
    for (int r4 = 0; r4 < R / 4; ++r4) {
      float4 zv[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) zv[q] = *reinterpret_cast<const float4 *>(H1T + (jg + 16 * q) * LD + 4 * r4);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int ii = kg + 16 * i;
        if (ii < I) {
          const float4 xv = *reinterpret_cast<const float4 *>(XT + ii * LD + 4 * r4);
#pragma unroll
          for (int j = 0; j < 4; ++j) acc1[i][j] = dot4(xv, zv[j], acc1[i][j]);
        }
      }
    }
    if (t >= 64 && t < 128) {  // db1[t - 64]
      for (int r4 = 0; r4 < R / 4; ++r4) {
        const float4 d = *reinterpret_cast<const float4 *>(H1T + (t - 64) * LD + 4 * r4);
        accb += (d.x + d.y) + (d.z + d.w);
      }
    }
  }

  // ---------------- publish this CTA's partial gradient
  float *out = a.partials + (int64_t)blockIdx.x * a.pstride;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int ii = kg + 16 * i;
    if (ii < I)
#pragma unroll
      for (int j = 0; j < 4; ++j) out[ii * H + jg + 16 * j] = acc1[i][j];
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) out[off_W2(I) + (kg + 16 * i) * H + jg + 16 * j] = acc2[i][j];
  {
    const int k = t & 63, og = t >> 6;
    if (og < O) out[off_W3(I) + k * O + og] = acc3[0];
    if (og + 4 < O) out[off_W3(I) + k * O + og + 4] = acc3[1];
  }
  if (t < 64) out[off_b2(I) + t] = accb;
  else if (t < 128) out[off_b1(I) + (t - 64)] = accb;
  else if (t < 128 + O) out[off_b3(I, O) + (t - 128)] = accb;
  // head sums: threads 0..63 = warps 0,1
  __syncthreads();
  float *red = sm + SmemMap::RED;
  if (t < 64) {
    const int lane = t & 31, w = t >> 5;
    float v;
    v = warp_sum(s_obj); if (lane == 0) red[w * 24 + 0] = v;
    v = warp_sum(s_kl); if (lane == 0) red[w * 24 + 1] = v;
    v = warp_sum(s_clip); if (lane == 0) red[w * 24 + 2] = v;
    v = warp_sum(s_adv); if (lane == 0) red[w * 24 + 3] = v;
    v = warp_sum(s_ret); if (lane == 0) red[w * 24 + 4] = v;
#pragma unroll
    for (int j = 0; j < MAX_O; ++j) { v = warp_sum(dls[j]); if (lane == 0) red[w * 24 + 8 + j] = v; }
  }
  __syncthreads();
  if (t < 16) {
    // tail layout: [n_params .. +8) = dlogΣ, [n_params+8 .. +16) = obj, kl, clip, adv, ret, 0, 0, 0
    const int src = t < 8 ? 8 + t : t - 8;
    float v = (src < 5 || src >= 8) ? red[src] + red[24 + src] : 0.f;
    out[a.n_params + t] = v;
  }
}

// sum the per-CTA partials (double accumulation, fixed order: bit-reproducible) -> gradient vector + tail
//   grads[p]                    p < n_params
//   grads[n_params + j]         j < 8  : dL/dlogΣ_j           (tail_ls_grad)
//   grads[n_params + 64 + q]    q < 5  : obj, kl, clip, adv, ret sums ; [5] = row count   (tail_sums)
__global__ void reduce_fused_partials_kernel(const float *__restrict__ partials, int nparts, int pstride, int n_params, float *__restrict__ grads,
                                             float count, const int *__restrict__ skip) {
  if (skip && *skip) return;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_params + 16) return;
  double s = 0.0;
  const float *src = partials + p;
#pragma unroll 4
  for (int c = 0; c < nparts; ++c) s += (double)src[(int64_t)c * pstride];
  if (p < n_params) grads[p] = (float)s;
  else if (p < n_params + 8) grads[p] = (float)s;
  else {
    const int q = p - n_params - 8;
    if (q < 5) grads[n_params + 64 + q] = (float)s;
    else if (q == 5) grads[n_params + 64 + 5] = count;
  }
}

bool fusable(const crux_mlp *m) {
  return m && m->n_layers == 3 && m->dims[1] == H && m->dims[2] == H && m->dims[0] >= 1 && m->dims[0] <= MAX_I && m->dims[3] >= 1 &&
         m->dims[3] <= MAX_O && m->acts[0] == m->acts[1] && (m->acts[0] == CRUX_ACT_TANH || m->acts[0] == CRUX_ACT_RELU) &&
         m->acts[2] == CRUX_ACT_IDENTITY;
}
NetDesc describe(const crux_mlp *m) {
  NetDesc nd;
  nd.params = m->params; nd.I = m->dims[0]; nd.O = m->dims[3]; nd.act = m->acts[0];
  nd.bytes16 = (uint32_t)(((size_t)m->n_params * sizeof(float) + 15) / 16 * 16);
  return nd;
}
int set_smem_attr(crux_ctx *ctx) {
  static bool done[3] = {false, false, false};
  if (!done[0]) { CRUX_CHECK_CUDA(ctx, cudaFuncSetAttribute(fused_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES)); done[0] = true; }
  if (!done[1]) { CRUX_CHECK_CUDA(ctx, cudaFuncSetAttribute(fused_minibatch_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES)); done[1] = true; }
  if (!done[2]) { CRUX_CHECK_CUDA(ctx, cudaFuncSetAttribute(fused_minibatch_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES)); done[2] = true; }
  return CRUX_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ entry points (internal)
// value(π, s) fast path for crux_mlp_forward
int mlp_forward_fused(crux_mlp *mlp, const float *x, int64_t B, float *y, int *handled) {
  *handled = 0;
  if (!fusable(mlp) || getenv("CRUX_NO_FUSED")) return CRUX_OK;
  crux_ctx *ctx = mlp->ctx;
  int rc = set_smem_attr(ctx); if (rc) return rc;
  FwdArgs a;
  memset(&a, 0, sizeof(a));
  a.net[0] = describe(mlp); a.mode[0] = 0; a.x = x; a.B = B; a.y[0] = y;
  const int64_t tiles = cdiv(B, R);
  dim3 grid((unsigned)i64min(tiles, (int64_t)ctx->num_sms * 2), 1);
  fused_forward_kernel<<<grid, NT, SMEM_BYTES, ctx->stream>>>(a);
  CRUX_LAUNCHED(ctx);
  *handled = 1;
  return CRUX_OK;
}

extern "C" int32_t crux_rollout_step_fused(crux_gaussian *actor, crux_mlp *critic, const float *obs, int64_t N, const float *eps_in,
                                           uint64_t seed, uint64_t ctr, float *a_out, float *logp_out, float *v_out, int *handled) {
  *handled = 0;
  if (getenv("CRUX_NO_FUSED")) return CRUX_OK;
  if (!fusable(actor->mu) || actor->head_mode || actor->squashed || actor->adim != actor->mu->dims[3]) return CRUX_OK;
  const bool with_critic = critic && v_out;
  if (with_critic && (!fusable(critic) || critic->dims[3] != 1 || critic->dims[0] != actor->mu->dims[0])) return CRUX_OK;
  crux_ctx *ctx = actor->ctx;
  int rc = set_smem_attr(ctx); if (rc) return rc;
  FwdArgs a;
  memset(&a, 0, sizeof(a));
  a.net[0] = describe(actor->mu); a.mode[0] = 1; a.y[0] = a_out; a.logp = logp_out; a.ls = actor->log_sigma; a.eps_in = eps_in;
  a.seed = seed; a.ctr = ctr; a.x = obs; a.B = N;
  if (with_critic) { a.net[1] = describe(critic); a.mode[1] = 0; a.y[1] = v_out; }
  const int64_t tiles = cdiv(N, R);
  dim3 grid((unsigned)i64min(tiles, (int64_t)ctx->num_sms * 2), with_critic ? 2 : 1);
  fused_forward_kernel<<<grid, NT, SMEM_BYTES, ctx->stream>>>(a);
  CRUX_LAUNCHED(ctx);
  *handled = 1;
  return CRUX_OK;
}

// one minibatch: forward + loss + backward -> mlp->grads (+ tail).  head: 0 actor (ppo/a2c), 1 critic (mse)
int fused_minibatch(crux_mlp *mlp, int head, const float *s, const float *act, const float *logp_old, const float *adv, const float *ret,
                    const int32_t *order, int64_t bm, const float *ls, float inv_bg, float eps_clip, float lambda_p, int a2c,
                    const int *skip, int *handled) {
  *handled = 0;
  if (!fusable(mlp) || getenv("CRUX_NO_FUSED")) return CRUX_OK;
  crux_ctx *ctx = mlp->ctx;
  int rc = set_smem_attr(ctx); if (rc) return rc;
  const int64_t tiles = cdiv(bm, R);
  const int grid = (int)i64min(tiles, (int64_t)ctx->num_sms * 2);
  const int pstride = (int)((mlp->n_params + 16 + 31) / 32 * 32);
  const size_t need = (size_t)grid * pstride * sizeof(float);
  if (need > mlp->partials_bytes) {
    CRUX_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (mlp->partials) cudaFree(mlp->partials);
    mlp->partials = nullptr; mlp->partials_bytes = 0;
    const size_t want = (size_t)ctx->num_sms * 2 * pstride * sizeof(float);
    if (cudaMalloc((void **)&mlp->partials, want > need ? want : need) != cudaSuccess) return crux_set_err(ctx, CRUX_ERR_OOM, "fused partials");
    mlp->partials_bytes = want > need ? want : need;
  }
  MbArgs a;
  memset(&a, 0, sizeof(a));
  a.net = describe(mlp); a.s = s; a.act = act; a.logp_old = logp_old; a.adv = adv; a.ret = ret; a.order = order; a.bm = bm; a.ls = ls;
  a.inv_bg = inv_bg; a.eps_clip = eps_clip; a.lambda_p = lambda_p; a.a2c = a2c; a.partials = mlp->partials; a.pstride = pstride;
  a.n_params = (int)mlp->n_params; a.skip = skip;
  if (head == 0) fused_minibatch_kernel<0><<<grid, NT, SMEM_BYTES, ctx->stream>>>(a);
  else fused_minibatch_kernel<1><<<grid, NT, SMEM_BYTES, ctx->stream>>>(a);
  CRUX_LAUNCHED(ctx);
  const int n_out = (int)mlp->n_params + 16;
  reduce_fused_partials_kernel<<<(n_out + 255) / 256, 256, 0, ctx->stream>>>(mlp->partials, grid, pstride, (int)mlp->n_params, mlp->grads,
                                                                          (float)bm, skip);
  CRUX_LAUNCHED(ctx);
  *handled = 1;
  return CRUX_OK;
}
