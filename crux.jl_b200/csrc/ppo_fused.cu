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
#include <cooperative_groups.h>
#include <stdlib.h>
#include <vector>

namespace {

constexpr int H = 64;        // hidden width
constexpr int R = 64;        // rows per tile (2 CTAs per SM); RB = 128-row tiles (1 CTA per SM, 8x4 register tiles) for big batches
constexpr int RB = 128;
constexpr int NT = 256;      // threads per CTA
constexpr int MAX_I = 32, MAX_O = 8;
constexpr int P_MAX = MAX_I * H + H + H * H + H + H * MAX_O + MAX_O;  // 6792 floats
constexpr int P_SMEM = (P_MAX + 3) / 4 * 4 + 8;

#define LOG_SQRT_2PI 0.9189385332046727f

// tanh(x) = 1 - 2/(exp(2x) + 1) on the SFU: ex2.approx.ftz + rcp.approx.ftz + one FMA (4 instructions; __expf / __fdividef expand to ~12
// with their range fix-ups, libdevice tanhf to ~30).  Absolute error ~2e-7 (the rounding of e + 1 and of the final subtraction: the RELATIVE
// error grows as x -> 0, which is why every forward tolerance carries an absolute term of 1e-6 or more: a deliberate divergence from the
// reference's tanh).  Saturates correctly for |x| large (ex2 -> +inf gives rcp -> 0, ex2 -> 0 gives 1 - 2 = -1).  The same
// formula as mb_t5.cuh's tanh_t5, so rollout, value passes and every minibatch-kernel variant evaluate the activation identically.
__device__ __forceinline__ float tanh_fast(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"((2.0f * x) * 1.4426950216293334961f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
  return fmaf(-2.0f, r, 1.0f);
}
__device__ __forceinline__ float act_fused(int act, float z) { return act == CRUX_ACT_TANH ? tanh_fast(z) : fmaxf(z, 0.0f); }

// ---- shared memory carve-up (floats), per tile height RT; row stride LD = RT + 4 keeps 128-bit alignment ----------------------
template <int RT>
struct SmemMapT {
  static constexpr int LD = RT + 4;
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
  static constexpr int IDX = RET + LD;              // [RT] ints
  static constexpr int RED = IDX + RT;              // [8][24] reduction scratch
  static constexpr int MBAR = RED + 8 * 24;         // 2 floats = one 64-bit mbarrier (8-byte aligned: all offsets are even)
  static constexpr int TOTAL = MBAR + 2;
  static constexpr size_t BYTES = (size_t)TOTAL * sizeof(float);
  static_assert(MBAR % 2 == 0, "mbarrier must be 8-byte aligned");
  static_assert(W2T % 4 == 0 && XT % 4 == 0 && H1T % 4 == 0 && OT % 4 == 0 && AT % 4 == 0, "16-byte alignment");
};
using SmemMap = SmemMapT<R>;   // offsets that do not depend on the tile height (P, W2T, W3T) are shared by all variants
constexpr size_t SMEM_BYTES = SmemMapT<R>::BYTES;

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

__device__ __forceinline__ void stage_params(float *sm, const NetDesc &nd, int mbar_off) {
  const uint32_t mbar = smem_u32(sm + mbar_off);
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

// transposed copies used by the data-backward GEMMs.  32x32 blocks with a diagonal skew: lane l handles column (l + i) & 31
// of row l, so both the read (W2[k][j]) and the write (W2T[j][k]) touch 32 distinct banks.
__device__ __forceinline__ void build_transposes(float *sm, int I, int O) {
  const float *W2 = sm + SmemMap::P + off_W2(I), *W3 = sm + SmemMap::P + off_W3(I);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;   // 8 warps: 4 blocks of 32x32, two warps per block (16 diagonals each)
  {
    const int blk = w >> 1, k0 = (blk >> 1) * 32, j0 = (blk & 1) * 32, i0 = (w & 1) * 16;
#pragma unroll 4
    for (int i = i0; i < i0 + 16; ++i) {
      const int k = k0 + lane, j = j0 + ((lane + i) & 31);
      sm[SmemMap::W2T + j * H + k] = W2[k * H + j];
    }
  }
  for (int e = threadIdx.x; e < MAX_O * H; e += NT) {
    const int o = e >> 6, k = e & 63;
    sm[SmemMap::W3T + e] = o < O ? W3[k * O + o] : 0.f;
  }
}

// ---- register-tiled GEMM pieces --------------------------------------------------------------------------------------------
// C^T[j][r] = act(b[j] + sum_{k<K} A^T[k][r] W[k][j]);  thread = rows PR*rg..PR*rg+PR-1, cols 4jg..4jg+3  (PR = RT/16: 4 or 8)
template <int RT>
__device__ __forceinline__ void layer_fwd(const float *__restrict__ AT, int K, const float *__restrict__ W, const float *__restrict__ b,
                                          float *__restrict__ CT, int act) {
  constexpr int LD = RT + 4, PR = RT / 16, NV = PR / 4;
  const int rg = threadIdx.x >> 4, jg = threadIdx.x & 15;
  float acc[PR][4];
#pragma unroll
  for (int i = 0; i < PR; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const float *ap = AT + PR * rg, *wp = W + 4 * jg;
  float4 a[NV], w = *reinterpret_cast<const float4 *>(wp);
#pragma unroll
  for (int v = 0; v < NV; ++v) a[v] = *reinterpret_cast<const float4 *>(ap + 4 * v);
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    // operands of step k+1 are requested before the FFMAs of step k (the last prefetch re-reads row K-1: harmless)
    const int kn = k + 1 < K ? k + 1 : k;
    float4 an[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) an[v] = *reinterpret_cast<const float4 *>(ap + kn * LD + 4 * v);
    const float4 wn = *reinterpret_cast<const float4 *>(wp + kn * H);
    const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const float av[4] = {a[v].x, a[v].y, a[v].z, a[v].w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[4 * v + i][j] = fmaf(av[i], wv[j], acc[4 * v + i][j]);
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) a[v] = an[v];
    w = wn;
  }
  const float4 bb = *reinterpret_cast<const float4 *>(b + 4 * jg);
  const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      float4 o;
      o.x = act_fused(act, acc[4 * v + 0][j] + bv[j]);
      o.y = act_fused(act, acc[4 * v + 1][j] + bv[j]);
      o.z = act_fused(act, acc[4 * v + 2][j] + bv[j]);
      o.w = act_fused(act, acc[4 * v + 3][j] + bv[j]);
      *reinterpret_cast<float4 *>(CT + (4 * jg + j) * LD + PR * rg + 4 * v) = o;
    }
}

// out^T[o][r] = b3[o] + sum_k h2^T[k][r] W3[k][o];  thread = (row t % RT, outputs og, og + G, ...), G = NT / RT groups
template <int RT>
__device__ __forceinline__ void layer_out(const float *__restrict__ H2T, const float *__restrict__ W3, const float *__restrict__ b3, int O,
                                          float *__restrict__ OT) {
  constexpr int LD = RT + 4, G = NT / RT, NO = (MAX_O + G - 1) / G;
  const int r = threadIdx.x % RT, og = threadIdx.x / RT;
  if (og >= O) return;
  float acc[NO];
#pragma unroll
  for (int q = 0; q < NO; ++q) acc[q] = og + q * G < O ? b3[og + q * G] : 0.f;
#pragma unroll 8
  for (int k = 0; k < H; ++k) {
    const float h = H2T[k * LD + r];
#pragma unroll
    for (int q = 0; q < NO; ++q)
      if (og + q * G < O) acc[q] = fmaf(h, W3[k * O + og + q * G], acc[q]);
  }
#pragma unroll
  for (int q = 0; q < NO; ++q)
    if (og + q * G < O) OT[(og + q * G) * LD + r] = acc[q];
}

// dA^T[k][r] = act'(A^T[k][r]) * sum_{j<J} dC^T[j][r] WT[j][k]   in place over A^T;  thread = rows PR*rg.., cols 4kg..
template <int RT>
__device__ __forceinline__ void layer_bwd_data(const float *__restrict__ DCT, int J, const float *__restrict__ WT, float *__restrict__ AT, int act) {
  constexpr int LD = RT + 4, PR = RT / 16, NV = PR / 4;
  const int rg = threadIdx.x >> 4, kg = threadIdx.x & 15;
  float acc[PR][4];
#pragma unroll
  for (int i = 0; i < PR; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const float *dp = DCT + PR * rg, *wp = WT + 4 * kg;
  float4 d[NV], w = *reinterpret_cast<const float4 *>(wp);
#pragma unroll
  for (int v = 0; v < NV; ++v) d[v] = *reinterpret_cast<const float4 *>(dp + 4 * v);
#pragma unroll 4
  for (int j = 0; j < J; ++j) {
    const int jn = j + 1 < J ? j + 1 : j;
    float4 dn[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) dn[v] = *reinterpret_cast<const float4 *>(dp + jn * LD + 4 * v);
    const float4 wn = *reinterpret_cast<const float4 *>(wp + jn * H);
    const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const float dv[4] = {d[v].x, d[v].y, d[v].z, d[v].w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[4 * v + i][c] = fmaf(dv[i], wv[c], acc[4 * v + i][c]);
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) d[v] = dn[v];
    w = wn;
  }
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      float4 *p = reinterpret_cast<float4 *>(AT + (4 * kg + c) * LD + PR * rg + 4 * v);
      const float4 y = *p;
      float4 o;
      o.x = acc[4 * v + 0][c] * act_bwd_from_out(act, y.x);
      o.y = acc[4 * v + 1][c] * act_bwd_from_out(act, y.y);
      o.z = acc[4 * v + 2][c] * act_bwd_from_out(act, y.z);
      o.w = acc[4 * v + 3][c] * act_bwd_from_out(act, y.w);
      *p = o;
    }
}

__device__ __forceinline__ float dot4(const float4 &a, const float4 &b, float acc) {
  acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
  return acc;
}

// ---- tensor-core building blocks (fused_minibatch_tc_kernel) -----------------------------------------------------------------
// 3xTF32 split accumulation on mma.sync.m16n8k8: x = hi + lo with hi = rna_tf32(x); acc += a_lo*b_hi + a_hi*b_lo + a_hi*b_hi keeps
// fp32-level accuracy (measured ~1e-6 relative, experiments/tcgen05_tf32_test.cu) where a single TF32 pass gives ~1e-3.
// tcgen05 was evaluated for these GEMMs (experiments/tcgen05_layouts_test.cu): with tf32 operands only K-major shared-memory
// tiles are accepted without the 128B/32B-base swizzle, so the weight-gradient GEMMs (contraction over rows) would need a
// second, transposed hi/lo copy of every activation -- that does not fit next to the row-major copies at M = 128 rows, and M = 64
// leaves half of the epilogue lanes idle.  The warp-level MMA reads the transposed activation tiles [feature][row] directly,
// both as A^T (contraction over features) and as A / B (contraction over rows).
// hi = x rounded to tf32 (10 explicit mantissa bits), nearest with ties away -- the value cvt.rna.tf32.f32 returns for every finite
// x, in two integer instructions (on sm_100a the cvt itself expands to five: add, |x| < inf test, select, mask, ...);
// lo = x - hi is exact.  Inf stays Inf (lo = NaN propagates like the FFMA kernel's Inf - Inf would).
__device__ __forceinline__ void split_tf32(float x, uint32_t &hi, uint32_t &lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma3(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], float b0, float b1) {
  uint32_t bh0, bl0, bh1, bl1;
  split_tf32(b0, bh0, bl0); split_tf32(b1, bh1, bl1);
  mma_tf32(c, al, bh0, bh1);   // small terms first
  mma_tf32(c, ah, bl0, bl1);
  mma_tf32(c, ah, bh0, bh1);
}
// Fragment coordinates of a lane: g = lane / 4, tq = lane % 4.
//   A (16x8)  a0 = A[g][tq]       a1 = A[g+8][tq]      a2 = A[g][tq+4]   a3 = A[g+8][tq+4]
//   B (8x8)   b0 = B[tq][g]       b1 = B[tq+4][g]
//   C (16x8)  c0 = C[g][2tq]      c1 = C[g][2tq+1]     c2 = C[g+8][2tq]  c3 = C[g+8][2tq+1]
// Weight operands are kept in shared memory in B-fragment order, one float2 {b0, b1} per (k-step, n-tile, lane).
//
// acc[q] = C[rows r0 + {g, g+8}][n-tile q] = sum over `ks_n` k-steps of A^T[k][row] * B[k][col]; A^T = transposed activation tile
// [k][LD], bp = fragment base of this warp's first n-tile (lane included), ks_stride = float2 per k-step.
template <int NQ>
__device__ __forceinline__ void mma_rows(const float *__restrict__ AT, int ks_n, int r0, const float2 *__restrict__ bp, int ks_stride,
                                         float (&acc)[NQ][4]) {
  constexpr int LD = R + 4;
  const int lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
#pragma unroll
  for (int q = 0; q < NQ; ++q)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[q][c] = 0.f;
  const float *ap = AT + tq * LD + r0 + g;
#pragma unroll 4
  for (int ks = 0; ks < ks_n; ++ks) {
    uint32_t ah[4], al[4];
    split_tf32(ap[(8 * ks) * LD], ah[0], al[0]);
    split_tf32(ap[(8 * ks) * LD + 8], ah[1], al[1]);
    split_tf32(ap[(8 * ks + 4) * LD], ah[2], al[2]);
    split_tf32(ap[(8 * ks + 4) * LD + 8], ah[3], al[3]);
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const float2 b = bp[ks * ks_stride + q * 32];
      mma3(acc[q], ah, al, b.x, b.y);
    }
  }
}
// acc[q] += dW[m0 + {g, g+8}][n0 + 8q + {2tq, 2tq+1}] = sum_{row<64} A^T[m][row] * D^T[n][row]   (both operands are transposed tiles)
template <int NQ, int ROWS = R>
__device__ __forceinline__ void mma_wgrad(const float *__restrict__ AT, int m0, const float *__restrict__ DT, int n0, float (&acc)[NQ][4]) {
  constexpr int LD = ROWS + 4;
  const int lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
  const float *ap = AT + (m0 + g) * LD + tq;
  const float *dp = DT + (n0 + g) * LD + tq;
#pragma unroll (ROWS == R ? 8 : 4)
  for (int ks = 0; ks < ROWS / 8; ++ks) {
    uint32_t ah[4], al[4];
    split_tf32(ap[8 * ks], ah[0], al[0]);
    split_tf32(ap[8 * LD + 8 * ks], ah[1], al[1]);
    split_tf32(ap[8 * ks + 4], ah[2], al[2]);
    split_tf32(ap[8 * LD + 8 * ks + 4], ah[3], al[3]);
#pragma unroll
    for (int q = 0; q < NQ; ++q) mma3(acc[q], ah, al, dp[(8 * q) * LD + 8 * ks], dp[(8 * q) * LD + 8 * ks + 4]);
  }
}

// ---- tile loads --------------------------------------------------------------------------------------------------------------
// x^T[i][r] = x[row(r)][i]; rows beyond n are zero.  idx (shared) holds the source row or -1.
template <int RT>
__device__ __forceinline__ void load_rows_T(float *__restrict__ XT, const float *__restrict__ x, const int *__restrict__ sidx, int I) {
  constexpr int LD = RT + 4;
  for (int e = threadIdx.x; e < RT * I; e += NT) {
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
  int64_t row0;              // mode 1: index of row 0 in the full vector step (noise streams are keyed by the absolute stream id)
  // host-environment rollouts (crux_rollout_host): x may be PINNED HOST memory read over PCIe by the kernel itself (no H2D copy
  // call); x_copy receives the rows on the device (the s column of the rollout) and y2 (pinned host) a second copy of the actions
  float *x_copy;
  float *y2;
  // rows whose x_flag byte (pinned host) is set read their input from x_reset instead of x: the next observation of a stream is
  // its s' row of the previous vector step unless the episode ended there, in which case it is the freshly reset state
  const uint8_t *x_flag;
  const float *x_reset;
};

// ---- small-batch variant pieces: 16-row tiles (4x more CTAs for a 4096-stream vector step), thread = (1 row, 4 cols)
constexpr int R16 = 16, LD16 = R16 + 4;
__device__ __forceinline__ void layer_fwd16(const float *__restrict__ AT, int K, const float *__restrict__ W, const float *__restrict__ b,
                                            float *__restrict__ CT, int act) {
  const int r = threadIdx.x >> 4, jg = threadIdx.x & 15;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const float *ap = AT + r, *wp = W + 4 * jg;
#pragma unroll 8
  for (int k = 0; k < K; ++k) {
    const float av = ap[k * LD16];
    const float4 w = *reinterpret_cast<const float4 *>(wp + k * H);
    acc[0] = fmaf(av, w.x, acc[0]); acc[1] = fmaf(av, w.y, acc[1]); acc[2] = fmaf(av, w.z, acc[2]); acc[3] = fmaf(av, w.w, acc[3]);
  }
  const float4 bb = *reinterpret_cast<const float4 *>(b + 4 * jg);
  CT[(4 * jg + 0) * LD16 + r] = act_fused(act, acc[0] + bb.x);
  CT[(4 * jg + 1) * LD16 + r] = act_fused(act, acc[1] + bb.y);
  CT[(4 * jg + 2) * LD16 + r] = act_fused(act, acc[2] + bb.z);
  CT[(4 * jg + 3) * LD16 + r] = act_fused(act, acc[3] + bb.w);
}
__device__ __forceinline__ void layer_out16(const float *__restrict__ H2T, const float *__restrict__ W3, const float *__restrict__ b3, int O,
                                            float *__restrict__ OT) {
  const int r = threadIdx.x & 15, o = threadIdx.x >> 4;  // 16 output slots >= MAX_O
  if (o >= O) return;
  float a0 = b3[o];
#pragma unroll 8
  for (int k = 0; k < H; ++k) a0 = fmaf(H2T[k * LD16 + r], W3[k * O + o], a0);
  OT[o * LD16 + r] = a0;
}

template <int RT>
__global__ void __launch_bounds__(NT, RT == RB ? 1 : 2) fused_forward_kernel(FwdArgs a) {
  constexpr int TS = RT == R16 ? R : RT;          // the 16-row variant lives in the 64-row carve-up (rows stride LD16)
  using M = SmemMapT<TS>;
  constexpr int LDT = RT == R16 ? LD16 : M::LD;
  extern __shared__ __align__(16) float sm[];
  const int which = blockIdx.y;
  const NetDesc nd = a.net[which];
  const int I = nd.I, O = nd.O;
  stage_params(sm, nd, M::MBAR);
  int *sidx = reinterpret_cast<int *>(sm + M::IDX);
  const float *P = sm + M::P;
  const int64_t n_tiles = (a.B + RT - 1) / RT;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    __syncthreads();
    if (threadIdx.x < RT) {
      const int64_t row = tile * RT + threadIdx.x;
      sidx[threadIdx.x] = row < a.B ? (int)row : -1;
    }
    __syncthreads();
    if (RT != R16) {
      load_rows_T<TS>(sm + M::XT, a.x, sidx, I);
      __syncthreads();
      layer_fwd<TS>(sm + M::XT, I, P, P + off_b1(I), sm + M::H1T, nd.act);
      __syncthreads();
      layer_fwd<TS>(sm + M::H1T, H, P + off_W2(I), P + off_b2(I), sm + M::H2T, nd.act);
      __syncthreads();
      layer_out<TS>(sm + M::H2T, P + off_W3(I), P + off_b3(I, O), O, sm + M::OT);
    } else {
      for (int e = threadIdx.x; e < RT * I; e += NT) {
        const int r = e / I, i = e - r * I;
        const int row = sidx[r];
        float v = 0.f;
        if (row >= 0) {
          if (a.x_copy) {   // x is pinned host memory written by the host since the last launch: bypass every cache
            v = __ldcv(a.x + (int64_t)row * I + i);
            if (a.x_flag && __ldcv(a.x_flag + row)) v = __ldcv(a.x_reset + (int64_t)row * I + i);
            if (which == 0) a.x_copy[(int64_t)row * I + i] = v;
          } else {
            v = __ldg(a.x + (int64_t)row * I + i);
          }
        }
        sm[M::XT + i * LDT + r] = v;
      }
      __syncthreads();
      layer_fwd16(sm + M::XT, I, P, P + off_b1(I), sm + M::H1T, nd.act);
      __syncthreads();
      layer_fwd16(sm + M::H1T, H, P + off_W2(I), P + off_b2(I), sm + M::H2T, nd.act);
      __syncthreads();
      layer_out16(sm + M::H2T, P + off_W3(I), P + off_b3(I, O), O, sm + M::OT);
    }
    __syncthreads();
    const float *OT = sm + M::OT;
    if (a.mode[which] == 0) {
      float *y = a.y[which];
      for (int e = threadIdx.x; e < RT * O; e += NT) {
        const int r = e / O, o = e - r * O;
        if (sidx[r] >= 0) y[(int64_t)sidx[r] * O + o] = OT[o * LDT + r];
      }
    } else if (threadIdx.x < RT && sidx[threadIdx.x] >= 0) {
      // exploration(::GaussianPolicy) policies.jl:338-344 + gaussian_logpdf :333-336 (same expression order as policy.cu)
      const int r = threadIdx.x;
      const int64_t i = sidx[r];
      float logp = 0.f, nrm[4];
      for (int j = 0; j < O; ++j) {
        const float mu = OT[j * LDT + r];
        const float ls = a.ls[j];
        const float sigma = expf(ls);
        const float var = sigma * sigma;
        float e;
        if (a.eps_in) e = a.eps_in[i * O + j];
        else {
          if ((j & 3) == 0) {
            const Philox4 p = philox4x32_10(a.seed, a.ctr, (uint64_t)(a.row0 + i) * ((O + 3) / 4) + (j >> 2));
            box_muller(p.x, p.y, nrm[0], nrm[1]);
            box_muller(p.z, p.w, nrm[2], nrm[3]);
          }
          e = nrm[j & 3];
        }
        const float act = e * sigma + mu;
        a.y[which][i * O + j] = act;
        if (a.y2) a.y2[i * O + j] = act;
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
  const int *ctl;   // ctl[1] = 1 + index of the minibatch after which training stopped (0: not stopped); NULL: never skip
  int mb;           // index of this minibatch in the update
  const float *planes;   // mb_t5.cuh: the network's weight planes (crux_mlp::frag in plane mode)
  long long *prof;       // CRUX_MB6_PROF=1: clock64 at the phase boundaries of CTA 0 (development aid)
  int *nan_flag;         // mb_t5.cuh: raised when a published partial holds a NaN (reduce_adam_kernel skips the update: training.jl:20)
  unsigned long long *trace;   // CRUX_MB6_TRACE=1: %globaltimer at the start / end of every CTA: [gridDim.x][2] (development aid)
};
// a minibatch is skipped when an EARLIER minibatch raised the stop flag (rl/ppo.jl:59 via training.jl:46,49)
__device__ __forceinline__ bool stopped(const int *ctl, int mb) { return ctl && ctl[1] != 0 && ctl[1] <= mb; }

// HEAD 0: ppo_loss / a2c_loss on a GaussianPolicy with a logΣ vector.  HEAD 1: Flux.mse(V(s), return).
// RT = 64 : 2 CTAs per SM, 4x4 register tiles (small and medium minibatches).
// RT = 128: 1 CTA per SM, 8x4 register tiles in the forward / data-backward GEMMs: a third fewer shared-memory wavefronts
//           per FFMA (the 64-row variant is shared-memory-bandwidth bound, profiles/).
// The all-FFMA kernel: the A/B reference of the tensor-core kernel below (CRUX_NO_MMA=1) and the 128-row variant (CRUX_RB=1).
template <int HEAD, int RT>
__global__ void __launch_bounds__(NT, RT == RB ? 1 : 2) fused_minibatch_kernel(MbArgs a) {
  if (stopped(a.ctl, a.mb)) return;
  using M = SmemMapT<RT>;
  constexpr int LD = M::LD;
  extern __shared__ __align__(16) float sm[];
  const NetDesc nd = a.net;
  const int I = nd.I, O = nd.O, act = nd.act;
  const int t = threadIdx.x;
  stage_params(sm, nd, M::MBAR);
  build_transposes(sm, I, O);
  int *sidx = reinterpret_cast<int *>(sm + M::IDX);
  const float *P = sm + M::P;
  float *XT = sm + M::XT, *H1T = sm + M::H1T, *H2T = sm + M::H2T, *OT = sm + M::OT, *AT = sm + M::AT;

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
  // head sums of this thread's rows (threads 0..RT-1)
  float s_obj = 0.f, s_kl = 0.f, s_clip = 0.f, s_adv = 0.f, s_ret = 0.f, dls[MAX_O];
#pragma unroll
  for (int j = 0; j < MAX_O; ++j) dls[j] = 0.f;

  const int64_t n_tiles = (a.bm + RT - 1) / RT;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    __syncthreads();
    if (t < RT) {
      const int64_t row = tile * RT + t;
      sidx[t] = row < a.bm ? (a.order ? a.order[row] : (int)row) : -1;
    }
    __syncthreads();
    load_rows_T<RT>(XT, a.s, sidx, I);
    if (HEAD == 0) {
      load_rows_T<RT>(AT, a.act, sidx, O);
      if (t < RT) {
        const int row = sidx[t];
        sm[M::LP + t] = row >= 0 ? a.logp_old[row] : 0.f;
        sm[M::ADV + t] = row >= 0 ? a.adv[row] : 0.f;
        sm[M::RET + t] = (row >= 0 && a.ret) ? a.ret[row] : 0.f;
      }
    } else if (t < RT) {
      const int row = sidx[t];
      sm[M::RET + t] = row >= 0 ? a.ret[row] : 0.f;
    }
    __syncthreads();
    // ---------------- forward
    layer_fwd<RT>(XT, I, P, P + off_b1(I), H1T, act);
    __syncthreads();
    layer_fwd<RT>(H1T, H, P + off_W2(I), P + off_b2(I), H2T, act);
    __syncthreads();
    layer_out<RT>(H2T, P + off_W3(I), P + off_b3(I, O), O, OT);
    __syncthreads();
    // ---------------- loss head: dL/dout (scaled by 1/B_global) replaces out^T
    if (t < RT) {
      const bool live = sidx[t] >= 0;
      if (HEAD == 0) {
        float logp = 0.f;
        for (int j = 0; j < O; ++j) {
          const float sg = expf(a.ls[j]);
          const float d = AT[j * LD + t] - OT[j * LD + t];
          logp += -(d * d) / (2.f * (sg * sg)) - LOG_SQRT_2PI - a.ls[j];
        }
        const float Ai = sm[M::ADV + t], old = sm[M::LP + t];
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
          s_kl += old - logp; s_adv += Ai; s_ret += sm[M::RET + t];
        }
        for (int j = 0; j < O; ++j) {
          const float sg = expf(a.ls[j]);
          const float var = sg * sg;
          const float d = AT[j * LD + t] - OT[j * LD + t];
          OT[j * LD + t] = dlogp * d / var;
          dls[j] += dlogp * (d * d / var - 1.f);
        }
      } else {
        const float d = OT[t] - sm[M::RET + t];
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
        for (int r4 = 0; r4 < RT / 4; ++r4) {
          const float4 h = *reinterpret_cast<const float4 *>(H2T + k * LD + 4 * r4);
          acc3[0] = dot4(h, *reinterpret_cast<const float4 *>(OT + og * LD + 4 * r4), acc3[0]);
          if (v1) acc3[1] = dot4(h, *reinterpret_cast<const float4 *>(OT + (og + 4) * LD + 4 * r4), acc3[1]);
        }
      }
      if (t >= 128 && t < 128 + O) {  // db3
        const int o = t - 128;
        for (int r4 = 0; r4 < RT / 4; ++r4) {
          const float4 d = *reinterpret_cast<const float4 *>(OT + o * LD + 4 * r4);
          accb += (d.x + d.y) + (d.z + d.w);
        }
      }
    }
    __syncthreads();
    // ---------------- dz2^T in place over h2^T
    layer_bwd_data<RT>(OT, O, sm + M::W3T, H2T, act);
    __syncthreads();
    // ---------------- dW2 += h1^T dz2 ; db2
#pragma unroll 2
    for (int r4 = 0; r4 < RT / 4; ++r4) {
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
      for (int r4 = 0; r4 < RT / 4; ++r4) {
        const float4 d = *reinterpret_cast<const float4 *>(H2T + t * LD + 4 * r4);
        accb += (d.x + d.y) + (d.z + d.w);
      }
    }
    __syncthreads();
    // ---------------- dz1^T in place over h1^T
    layer_bwd_data<RT>(H2T, H, sm + M::W2T, H1T, act);
    __syncthreads();
    // ---------------- dW1 += x^T dz1 ; db1
#pragma unroll 2
    for (int r4 = 0; r4 < RT / 4; ++r4) {
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
      for (int r4 = 0; r4 < RT / 4; ++r4) {
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
  // head sums: threads 0..RT-1 = the first RT/32 warps
  __syncthreads();
  float *red = sm + M::RED;
  if (t < RT) {
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
    float v = 0.f;
    if (src < 5 || src >= 8)
#pragma unroll
      for (int w = 0; w < RT / 32; ++w) v += red[w * 24 + src];
    out[a.n_params + t] = v;
  }
}

// =================================================================================================== tensor-core minibatch kernel
// Same contract as fused_minibatch_kernel<HEAD, 64> (same partials layout, same head arithmetic), with EVERY GEMM of the tile on
// the tensor cores (3xTF32 split accumulation, mma.sync m16n8k8; 8 warps, warp (wm, wn) = (w & 3, w >> 2)):
//   layer 1      z1[64 x 64]  = x[64 x I]    W1          K = I padded to a multiple of 8 (zero rows)      mma_rows<4>
//   layer 2      z2[64 x 64]  = h1           W2          K = 64                                           mma_rows<4>
//   output       o [64 x 8]   = h2           W3          N = O padded to 8 (zero columns), warps 0..3     mma_rows<1>
//   dz2          [64 x 64]    = dOut[64 x 8] W3^T        K = 8 (one k-step)                               mma_rows<4>
//   dz1          [64 x 64]    = dz2          W2^T        K = 64                                           mma_rows<4>
//   dW3 [64 x 8]  += h2^T dOut   (warps 0..3)   dW2 [64 x 64] += h1^T dz2   dW1 [32 x 64] += x^T dz1      mma_wgrad<1/4/2>
// Weight operands live in shared memory in B-fragment order (built once per CTA; W2's fragments replace W2 inside the staged
// parameter vector, in place).  Activations stay transposed [feature][row]; bias gradients and the loss head are the FFMA code.
struct TcMapP {   // tensor-core FORWARD kernel: raw parameters staged, fragments built in the kernel
  static constexpr int LD = R + 4;
  static constexpr int P = 0;                        // raw parameter vector (TMA destination); W2's slot is rewritten as B fragments of W2
  static constexpr int W2T = P + P_SMEM;             // [8 ks][8 nt][32] float2 : B fragments of W2^T
  static constexpr int W1F = W2T + H * H;            // [4 ks][8 nt][32] float2 : B fragments of W1 (rows >= I are zero)
  static constexpr int W3F = W1F + 4 * 8 * 32 * 2;   // [8 ks][32] float2       : B fragments of W3 (columns >= O are zero)
  static constexpr int W3TF = W3F + 8 * 32 * 2;      // [8 nt][32] float2       : B fragments of W3^T (k = o < 8; o >= O zero)
  static constexpr int XT = W3TF + 8 * 32 * 2;       // [32][LD]  rows >= I stay zero
  static constexpr int H1T = XT + MAX_I * LD;        // [64][LD]
  static constexpr int H2T = H1T + H * LD;           // [64][LD]
  static constexpr int OT = H2T + H * LD;            // [8][LD]   outputs, then dL/dout; rows >= O stay zero
  static constexpr int AT = OT + MAX_O * LD;         // [8][LD]   stored actions
  static constexpr int LP = AT + MAX_O * LD;
  static constexpr int ADV = LP + LD;
  static constexpr int RET = ADV + LD;
  static constexpr int IDX = RET + LD;               // [R] ints: source rows of the current tile (-1: padding row)
  static constexpr int IDX2 = IDX + R;               // [R] ints: source rows of this CTA's next tile
  static constexpr int RED = IDX2 + R;
  static constexpr int LSC = RED + 8 * 24;           // [8] logΣ_j, [8] σ_j²  (actor head constants)
  static constexpr int MBAR = LSC + 16;
  static constexpr int TOTAL = MBAR + 2;
  static constexpr size_t BYTES = (size_t)TOTAL * sizeof(float);
  static_assert(MBAR % 2 == 0 && W2T % 2 == 0 && W1F % 2 == 0 && W3F % 2 == 0 && W3TF % 2 == 0, "8-byte alignment");
  // gather staging of the NEXT tile (cp.async, row-major) lives in the h2^T region while that is dead (after dz1 of the current tile,
  // before layer 2 of the next one): x [R][I] | actions [R][O] | logprob, advantage, return [R] each
  static constexpr int SX = H2T, SA = SX + R * MAX_I, SH = SA + R * MAX_O;
  static_assert(SH + 3 * R <= OT, "gather staging must fit in the h2^T region");
};
// Fragment buffer of a network (global memory, crux_mlp::frag): everything the tensor-core minibatch kernel needs from the
// parameters, in the order it is staged into shared memory with ONE TMA bulk copy.  Built by build_frag_kernel at the start of a
// PPO update, then kept in step with `params` by the fused Adam kernel (frag_scatter).
struct Frag {
  static constexpr int W2F = 0;                    // [8 ks][8 nt][32] float2 : B fragments of W2      (b0 = B[8ks+t][8nt+g], b1 = B[8ks+t+4][8nt+g])
  static constexpr int W2T = W2F + H * H;          // [8 ks][8 nt][32] float2 : B fragments of W2^T
  static constexpr int W1F = W2T + H * H;          // [4 ks][8 nt][32] float2 : B fragments of W1      (rows >= I zero)
  static constexpr int W3F = W1F + 4 * 8 * 32 * 2; // [8 ks][32] float2       : B fragments of W3      (columns >= O zero)
  static constexpr int W3TF = W3F + 8 * 32 * 2;    // [8 nt][32] float2       : B fragments of W3^T    (k = o < 8; o >= O zero)
  static constexpr int B1 = W3TF + 8 * 32 * 2;     // [64]
  static constexpr int B2 = B1 + H;                // [64]
  static constexpr int B3 = B2 + H;                // [8]
  static constexpr int TOTAL = B3 + MAX_O;         // 11400 floats = 45600 bytes (a multiple of 16)
  static_assert((TOTAL * 4) % 16 == 0, "TMA bulk copies move multiples of 16 bytes");
};
// position of B[k][n] inside a fragment array with `nts` n-tiles per k-step
__host__ __device__ __forceinline__ int frag_pos(int k, int n, int nts) {
  return 2 * ((((k >> 3) * nts + (n >> 3)) * 32) + (n & 7) * 4 + (k & 3)) + ((k >> 2) & 1);
}
// parameter i of the flat vector (Flux order W1 b1 W2 b2 W3 b3) -> its copies in the fragment buffer
__device__ __forceinline__ void frag_scatter(float *__restrict__ frag, int I, int O, int i, float v) {
  if (i < I * H) { frag[Frag::W1F + frag_pos(i / H, i % H, 8)] = v; return; }          // W1[ii][j]: B = W1
  i -= I * H;
  if (i < H) { frag[Frag::B1 + i] = v; return; }
  i -= H;
  if (i < H * H) {
    const int k = i / H, j = i % H;
    frag[Frag::W2F + frag_pos(k, j, 8)] = v;                                             // B = W2   : B[k][j]
    frag[Frag::W2T + frag_pos(j, k, 8)] = v;                                             // B = W2^T : B[j][k]
    return;
  }
  i -= H * H;
  if (i < H) { frag[Frag::B2 + i] = v; return; }
  i -= H;
  if (i < H * O) {
    const int k = i / O, o = i % O;
    frag[Frag::W3F + frag_pos(k, o, 1)] = v;                                             // B = W3   : B[k][o], a single n-tile
    frag[Frag::W3TF + frag_pos(o, k, 8)] = v;                                            // B = W3^T : B[o][k], a single k-step
    return;
  }
  i -= H * O;
  if (i < O) frag[Frag::B3 + i] = v;
}
// one thread per fragment-buffer entry (gather form; padding entries are written as zeros)
__global__ void build_frag_kernel(const float *__restrict__ params, float *__restrict__ frag, int I, int O) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= Frag::TOTAL) return;
  const float *W1 = params, *b1 = W1 + I * H, *W2 = b1 + H, *b2 = W2 + H * H, *W3 = b2 + H, *b3 = W3 + H * O;
  float v = 0.f;
  if (e < Frag::B1) {
    const int base = e < Frag::W2T ? Frag::W2F : e < Frag::W1F ? Frag::W2T : e < Frag::W3F ? Frag::W1F : e < Frag::W3TF ? Frag::W3F : Frag::W3TF;
    const int nts = base == Frag::W3F ? 1 : 8;
    const int idx = e - base, half = idx & 1, f = idx >> 1, lane = f & 31, blk = f >> 5, nt = blk % nts, ks = blk / nts;
    const int k = 8 * ks + (lane & 3) + 4 * half, n = 8 * nt + (lane >> 2);   // the entry holds B[k][n]
    if (base == Frag::W2F) v = W2[k * H + n];
    else if (base == Frag::W2T) v = W2[n * H + k];
    else if (base == Frag::W1F) v = k < I ? W1[k * H + n] : 0.f;
    else if (base == Frag::W3F) v = n < O ? W3[k * O + n] : 0.f;          // B[k][o]
    else v = k < O ? W3[n * O + k] : 0.f;                                  // B[o][k] = W3[k][o]
  } else if (e < Frag::B2) v = b1[e - Frag::B1];
  else if (e < Frag::B3) v = b2[e - Frag::B2];
  else v = e - Frag::B3 < O ? b3[e - Frag::B3] : 0.f;
  frag[e] = v;
}

struct TcMap {
  static constexpr int LD = R + 4;
  static constexpr int FR = 0;                       // the staged fragment buffer (Frag layout)
  static constexpr int XT = FR + Frag::TOTAL;        // [32][LD]  rows >= I stay zero
  static constexpr int H1T = XT + MAX_I * LD;        // [64][LD]
  static constexpr int H2T = H1T + H * LD;           // [64][LD]
  static constexpr int OT = H2T + H * LD;            // [8][LD]   outputs, then dL/dout; rows >= O stay zero
  static constexpr int AT = OT + MAX_O * LD;         // [8][LD]   stored actions
  static constexpr int LP = AT + MAX_O * LD;
  static constexpr int ADV = LP + LD;
  static constexpr int RET = ADV + LD;
  static constexpr int IDX = RET + LD;               // [R] ints: source rows of the current tile (-1: padding row)
  static constexpr int IDX2 = IDX + R;               // [R] ints: source rows of this CTA's next tile
  static constexpr int RED = IDX2 + R;
  static constexpr int LSC = RED + 8 * 24;           // [8] logΣ_j, [8] σ_j²  (actor head constants)
  static constexpr int MBAR = LSC + 16;
  static constexpr int TOTAL = MBAR + 2;
  static constexpr size_t BYTES = (size_t)TOTAL * sizeof(float);
  static_assert(MBAR % 2 == 0 && XT % 4 == 0, "alignment");
  // gather staging of the NEXT tile (cp.async, row-major) lives in the h2^T region while that is dead (after dz1 of the current tile,
  // before layer 2 of the next one): x [R][I] | actions [R][O] | logprob, advantage, return [R] each
  static constexpr int SX = H2T, SA = SX + R * MAX_I, SH = SA + R * MAX_O;
  static_assert(SH + 3 * R <= OT, "gather staging must fit in the h2^T region");
};
__device__ __forceinline__ void cp_async4(float *dst_smem, const float *src, bool valid) {
  const int nbytes = valid ? 4 : 0;   // 0 source bytes: the destination is zero-filled, nothing is read
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(nbytes) : "memory");
}

template <int HEAD>
__global__ void __launch_bounds__(NT, 2) fused_minibatch_tc_kernel(MbArgs a) {
  const int stop_at = a.ctl ? a.ctl[1] : 0;   // requested now, tested after the prologue (a dependent global load off the critical path)
  using M = TcMap;
  constexpr int LD = M::LD;
  extern __shared__ __align__(16) float sm[];
  const NetDesc nd = a.net;
  const int I = nd.I, O = nd.O, act = nd.act;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5, g = lane >> 2, tq = lane & 3;
  const int r0 = 16 * (w & 3), wn = w >> 2;
  float *XT = sm + M::XT, *H1T = sm + M::H1T, *H2T = sm + M::H2T, *OT = sm + M::OT, *AT = sm + M::AT;
  int *sidx = reinterpret_cast<int *>(sm + M::IDX);
  int *sidx2 = reinterpret_cast<int *>(sm + M::IDX2);
  const uint32_t inv_I = (65536u + (uint32_t)I - 1u) / (uint32_t)I, inv_O = (65536u + (uint32_t)O - 1u) / (uint32_t)O;   // exact e / I for e < 2176
  // asynchronous gather of the rows listed in sidx2 into the staging area (no registers held, no wait here)
  auto issue_gather = [&]() {
    for (int e = t; e < R * I; e += NT) {
      const int r = (int)(((uint32_t)e * inv_I) >> 16), i = e - r * I;
      const int row = sidx2[r];
      cp_async4(sm + M::SX + e, a.s + (row >= 0 ? (int64_t)row * I + i : 0), row >= 0);
    }
    if (HEAD == 0)
      for (int e = t; e < R * O; e += NT) {
        const int r = (int)(((uint32_t)e * inv_O) >> 16), o = e - r * O;
        const int row = sidx2[r];
        cp_async4(sm + M::SA + e, a.act + (row >= 0 ? (int64_t)row * O + o : 0), row >= 0);
      }
    if (t < R) {
      const int row = sidx2[t];
      if (HEAD == 0) {
        cp_async4(sm + M::SH + t, a.logp_old + (row >= 0 ? row : 0), row >= 0);
        cp_async4(sm + M::SH + R + t, a.adv + (row >= 0 ? row : 0), row >= 0);
      }
      const bool has_ret = a.ret != nullptr;
      cp_async4(sm + M::SH + 2 * R + t, has_ret ? a.ret + (row >= 0 ? row : 0) : a.s, has_ret && row >= 0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  const int64_t n_tiles = (a.bm + R - 1) / R;
  auto tile_row = [&](int64_t tile) -> int {   // source row of this thread's row in `tile` (t < R)
    const int64_t row = tile * R + t;
    return (tile < n_tiles && row < a.bm) ? (a.order ? a.order[row] : (int)row) : -1;
  };
  // the first tile's rows start streaming in while the parameters are staged and the weight fragments are built;
  // zero the padding rows the GEMMs read (x^T rows >= I, out^T rows >= O)
  if (t < R) sidx2[t] = tile_row(blockIdx.x);
  for (int e = t; e < MAX_I * LD; e += NT) XT[e] = 0.f;
  for (int e = t; e < MAX_O * LD; e += NT) OT[e] = 0.f;
  __syncthreads();
  issue_gather();
  stage_params(sm, nd, M::MBAR);   // nd.params = the network's fragment buffer (Frag layout): weights arrive in B-fragment order
  const float *FRs = sm + M::FR;
  const float2 *W2F = reinterpret_cast<const float2 *>(FRs + Frag::W2F), *W2TF = reinterpret_cast<const float2 *>(FRs + Frag::W2T);
  const float2 *W1F = reinterpret_cast<const float2 *>(FRs + Frag::W1F), *W3F = reinterpret_cast<const float2 *>(FRs + Frag::W3F);
  const float2 *W3TF = reinterpret_cast<const float2 *>(FRs + Frag::W3TF);
  if (stop_at != 0 && stop_at <= a.mb) {   // an EARLIER minibatch raised the KL stop flag (rl/ppo.jl:59): nothing to do
    asm volatile("cp.async.wait_all;" ::: "memory");
    return;
  }
  const int ks1 = (I + 7) >> 3;

  // per-CTA gradient accumulators (registers, live across all tiles), in C-fragment layout
  float acc2[4][4], acc1[2][4], acc3[1][4], accb = 0.f;
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int e = 0; e < 4; ++e) { acc2[q][e] = 0.f; if (q < 2) acc1[q][e] = 0.f; if (q < 1) acc3[q][e] = 0.f; }
  // head sums: rows on the tq == 0 lanes of warps 0..3; dL/dlogΣ of the outputs {2tq, 2tq+1} on every lane of warps 0..3
  float s_obj = 0.f, s_kl = 0.f, s_clip = 0.f, s_adv = 0.f, s_ret = 0.f, dls[2] = {0.f, 0.f};
  if (HEAD == 0 && t < O) {
    const float ls = a.ls[t], sg = expf(ls);
    sm[M::LSC + t] = ls; sm[M::LSC + 8 + t] = sg * sg;
  }

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();   // this tile's staged rows are visible to everybody; the previous tile is done with every other buffer
    // next tile's source rows: requested now, stored to shared memory after layer 1 (the load latency hides behind it)
    const int nidx = t < R ? tile_row(tile + gridDim.x) : -1;
    for (int e = t; e < R * I; e += NT) {
      const int r = (int)(((uint32_t)e * inv_I) >> 16), i = e - r * I;
      XT[i * LD + r] = sm[M::SX + e];
    }
    if (HEAD == 0) {
      for (int e = t; e < R * O; e += NT) {
        const int r = (int)(((uint32_t)e * inv_O) >> 16), o = e - r * O;
        AT[o * LD + r] = sm[M::SA + e];
      }
      if (t < R) { sm[M::LP + t] = sm[M::SH + t]; sm[M::ADV + t] = sm[M::SH + R + t]; }
    }
    if (t < R) { sm[M::RET + t] = sm[M::SH + 2 * R + t]; sidx[t] = sidx2[t]; }
    __syncthreads();
    // ---------------- layer 1
    {
      float c[4][4];
      mma_rows<4>(XT, ks1, r0, W1F + (4 * wn) * 32 + lane, 8 * 32, c);
      const float *b1 = FRs + Frag::B1;
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = 32 * wn + 8 * q + 2 * tq + (e & 1), r = r0 + g + 8 * (e >> 1);
          H1T[j * LD + r] = act_fused(act, c[q][e] + b1[j]);
        }
      if (t < R) sidx2[t] = nidx;
    }
    __syncthreads();
    // ---------------- layer 2
    {
      float c[4][4];
      mma_rows<4>(H1T, 8, r0, W2F + (4 * wn) * 32 + lane, 8 * 32, c);
      const float *b2 = FRs + Frag::B2;
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = 32 * wn + 8 * q + 2 * tq + (e & 1), r = r0 + g + 8 * (e >> 1);
          H2T[j * LD + r] = act_fused(act, c[q][e] + b2[j]);
        }
    }
    __syncthreads();
    // ---------------- output layer + loss head on warps 0..3: one 16-row m-tile each, a single n-tile of 8 outputs.  The C fragment
    //                  puts the outputs {2tq, 2tq+1} of rows g and g+8 in lane (g, tq): the four lanes of a row hold its whole output
    //                  vector, so ppo_loss / a2c_loss / mse and dL/dout (scaled by 1/B_global) are evaluated straight from the
    //                  accumulators (two shuffles per row for logpdf) and only dL/dout goes to shared memory (out^T).
    if (w < 4) {
      float c[1][4];
      mma_rows<1>(H2T, 8, r0, W3F + lane, 32, c);
      const float *b3 = FRs + Frag::B3;
      const float *lsc = sm + M::LSC;
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int row = r0 + g + 8 * rr;
        const bool live = sidx[row] >= 0;
        if (HEAD == 0) {
          float d[2] = {0.f, 0.f}, part = 0.f;
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int o = 2 * tq + u;
            if (o < O) {
              d[u] = AT[o * LD + row] - (c[0][2 * rr + u] + b3[o]);
              part += -(d[u] * d[u]) / (2.f * lsc[8 + o]) - LOG_SQRT_2PI - lsc[o];
            }
          }
          part += __shfl_xor_sync(0xffffffffu, part, 1);
          const float logp = part + __shfl_xor_sync(0xffffffffu, part, 2);
          const float Ai = sm[M::ADV + row], old = sm[M::LP + row];
          float dlogp = 0.f;
          if (live) {
            float obj, clip = 0.f;
            if (a.a2c) {
              obj = logp * Ai;
              dlogp = -a.lambda_p * a.inv_bg * Ai;
            } else {
              const float rt = expf(logp - old);
              const float lo = 1.f - a.eps_clip, hi = 1.f + a.eps_clip;
              const float x = rt * Ai, y = fminf(fmaxf(rt, lo), hi) * Ai;
              const bool first = !(y < x);  // min(x, y) keeps x on ties
              obj = first ? x : y;
              dlogp = first ? -a.lambda_p * a.inv_bg * x : 0.f;
              clip = (rt > hi || rt < lo) ? 1.f : 0.f;
            }
            if (tq == 0) { s_obj += obj; s_clip += clip; s_kl += old - logp; s_adv += Ai; s_ret += sm[M::RET + row]; }   // once per row
          }
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int o = 2 * tq + u;
            if (o < O) {
              const float var = lsc[8 + o];
              OT[o * LD + row] = dlogp * d[u] / var;
              dls[u] += dlogp * (d[u] * d[u] / var - 1.f);
            }
          }
        } else if (tq == 0) {   // critic: a single output, held by the tq == 0 lane of the row
          const float d = (c[0][2 * rr] + b3[0]) - sm[M::RET + row];
          if (live) s_obj += d * d;
          OT[row] = live ? 2.f * d * a.inv_bg : 0.f;
        }
      }
    }
    __syncthreads();
    // ---------------- dW3 += h2^T dOut (warps 0..3) ; db3 ; dz2 = (dOut W3^T) .* act'(h2), accumulated in registers first
    {
      if (w < 4) mma_wgrad<1>(H2T, r0, OT, 0, acc3);
      if (t >= 128 && t < 128 + O) {
        const int o = t - 128;
        for (int r4 = 0; r4 < R / 4; ++r4) {
          const float4 d = *reinterpret_cast<const float4 *>(OT + o * LD + 4 * r4);
          accb += (d.x + d.y) + (d.z + d.w);
        }
      }
      float c[4][4];
      mma_rows<4>(OT, 1, r0, W3TF + (4 * wn) * 32 + lane, 0, c);
      __syncthreads();   // dW3 has read h2^T
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int k = 32 * wn + 8 * q + 2 * tq + (e & 1), r = r0 + g + 8 * (e >> 1);
          float *p = H2T + k * LD + r;
          *p = c[q][e] * act_bwd_from_out(act, *p);
        }
    }
    __syncthreads();
    // ---------------- dW2 += h1^T dz2 ; db2 ; dz1 = (dz2 W2^T) .* act'(h1), accumulated in registers first
    {
      mma_wgrad<4>(H1T, r0, H2T, 32 * wn, acc2);
      if (t < 64) {
        for (int r4 = 0; r4 < R / 4; ++r4) {
          const float4 d = *reinterpret_cast<const float4 *>(H2T + t * LD + 4 * r4);
          accb += (d.x + d.y) + (d.z + d.w);
        }
      }
      float c[4][4];
      mma_rows<4>(H2T, 8, r0, W2TF + (4 * wn) * 32 + lane, 8 * 32, c);
      __syncthreads();   // dW2 has read h1^T
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int k = 32 * wn + 8 * q + 2 * tq + (e & 1), r = r0 + g + 8 * (e >> 1);
          float *p = H1T + k * LD + r;
          *p = c[q][e] * act_bwd_from_out(act, *p);
        }
    }
    __syncthreads();
    // ---------------- dW1 += x^T dz1 (M = 32 input rows: warp = m-tile w & 1, n-tiles 2 (w >> 1) + {0, 1}) ; db1
    if (tile + gridDim.x < n_tiles) issue_gather();   // h2^T is dead until layer 2 of the next tile: stream the next rows into it
    mma_wgrad<2>(XT, 16 * (w & 1), H1T, 16 * (w >> 1), acc1);
    if (t >= 64 && t < 128) {
      for (int r4 = 0; r4 < R / 4; ++r4) {
        const float4 d = *reinterpret_cast<const float4 *>(H1T + (t - 64) * LD + 4 * r4);
        accb += (d.x + d.y) + (d.z + d.w);
      }
    }
  }

  // ---------------- publish this CTA's partial gradient (same layout as fused_minibatch_kernel)
  float *out = a.partials + (int64_t)blockIdx.x * a.pstride;
#pragma unroll
  for (int q = 0; q < 2; ++q)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int i = 16 * (w & 1) + g + 8 * (e >> 1), j = 16 * (w >> 1) + 8 * q + 2 * tq + (e & 1);
      if (i < I) out[i * H + j] = acc1[q][e];
    }
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int e = 0; e < 4; ++e) out[off_W2(I) + (r0 + g + 8 * (e >> 1)) * H + 32 * wn + 8 * q + 2 * tq + (e & 1)] = acc2[q][e];
  if (w < 4) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = r0 + g + 8 * (e >> 1), o = 2 * tq + (e & 1);
      if (o < O) out[off_W3(I) + k * O + o] = acc3[0][e];
    }
  }
  if (t < 64) out[off_b2(I) + t] = accb;
  else if (t < 128) out[off_b1(I) + (t - 64)] = accb;
  else if (t < 128 + O) out[off_b3(I, O) + (t - 128)] = accb;
  // head sums live in warps 0..3
  __syncthreads();
  float *red = sm + M::RED;
  if (w < 4) {
    float v;
    v = warp_sum(s_obj); if (lane == 0) red[w * 24 + 0] = v;
    v = warp_sum(s_kl); if (lane == 0) red[w * 24 + 1] = v;
    v = warp_sum(s_clip); if (lane == 0) red[w * 24 + 2] = v;
    v = warp_sum(s_adv); if (lane == 0) red[w * 24 + 3] = v;
    v = warp_sum(s_ret); if (lane == 0) red[w * 24 + 4] = v;
#pragma unroll
    for (int u = 0; u < 2; ++u) {   // sum over the rows (g) of the warp: lanes g == 0 end up with output 2tq + u
      v = dls[u];
      v += __shfl_xor_sync(0xffffffffu, v, 4); v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (g == 0) red[w * 24 + 8 + 2 * tq + u] = v;
    }
  }
  __syncthreads();
  if (t < 16) {
    // tail layout: [n_params .. +8) = dlogΣ, [n_params+8 .. +16) = obj, kl, clip, adv, ret, 0, 0, 0
    const int src = t < 8 ? 8 + t : t - 8;
    float v = 0.f;
    if (src < 5 || src >= 8)
#pragma unroll
      for (int ww = 0; ww < 4; ++ww) v += red[ww * 24 + src];
    out[a.n_params + t] = v;
  }
}

#include "fwd_tc5.cuh"

#include "mb_tc5.cuh"
#include "mb_t5.cuh"
#include "mb_persist.cuh"

// =================================================================================================== tensor-core forward kernel
// value(π, s) over a whole rollout column (the two critic passes that feed the GAE scan, policies.jl:94-98) on the building blocks of
// fused_minibatch_tc_kernel: contiguous 64-row tiles streamed in with 16-byte cp.async one tile ahead (staging = the W2^T slot, which
// a forward pass does not need), three MMA layers, outputs written straight from the accumulators.
struct FwdTcArgs {
  NetDesc net; const float *x; int64_t B; float *y;
  // value(V, sp) over a [T][N] rollout (crux_value_next): x = sp, x_alt = s + N rows, y_alt = V(s) + N rows.  A tile whose rows all
  // satisfy sp[row] == s[row + N] bit for bit (every transition that is not followed by a reset) copies V(s)[row + N] instead of
  // running the network; alt_rows = number of rows that have a successor row (B - N).  NULL x_alt: plain forward.
  const float *x_alt; const float *y_alt; int64_t alt_rows;
};

template <int DUMMY>
__global__ void __launch_bounds__(NT, 2) fused_forward_tc_kernel(FwdTcArgs a) {
  using M = TcMapP;
  constexpr int LD = M::LD;
  extern __shared__ __align__(16) float sm[];
  const NetDesc nd = a.net;
  const int I = nd.I, O = nd.O, act = nd.act;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5, g = lane >> 2, tq = lane & 3;
  const int r0 = 16 * (w & 3), wn = w >> 2;
  float *XT = sm + M::XT, *H1T = sm + M::H1T, *H2T = sm + M::H2T, *ST = sm + M::W2T;
  const int64_t n_tiles = (a.B + R - 1) / R;
  const int chunks = R * I / 4;   // 16-byte chunks of one tile (R * I is a multiple of 4)
  auto has_alt = [&](int64_t tile) { return a.x_alt != nullptr && (tile + 1) * R <= a.alt_rows; };   // every row of the tile has a successor row
  auto issue_tile = [&](int64_t tile) {
    const int64_t f0 = tile * R * I, f_end = a.B * I;   // first float of the tile, end of the column
    const bool alt = has_alt(tile);
    for (int c = t; c < chunks; c += NT) {
      const int64_t f = f0 + 4 * c;
      const int64_t left = f_end - f;
      const int nbytes = left >= 4 ? 16 : (left > 0 ? (int)left * 4 : 0);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(ST + 4 * c)), "l"(a.x + (nbytes ? f : 0)), "r"(nbytes) : "memory");
      if (alt) asm volatile("cp.async.cg.shared.global [%0], [%1], 16, 16;" ::"r"(smem_u32(ST + R * MAX_I + 4 * c)), "l"(a.x_alt + f) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  for (int e = t; e < MAX_I * LD; e += NT) XT[e] = 0.f;
  issue_tile(blockIdx.x);
  stage_params(sm, nd, M::MBAR);
  const float *P = sm + M::P;
  {  // weight operands in B-fragment order (W2's fragments replace W2 inside the staged parameter vector)
    float *W2 = sm + M::P + off_W2(I);
    const float *W1 = P, *W3 = P + off_W3(I);
    float2 f2[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int e = t + u * NT, l = e & 31, nt = (e >> 5) & 7, ks = e >> 8, gg = l >> 2, tt = l & 3;
      f2[u] = make_float2(W2[(8 * ks + tt) * H + 8 * nt + gg], W2[(8 * ks + tt + 4) * H + 8 * nt + gg]);
    }
    float2 *w1f = reinterpret_cast<float2 *>(sm + M::W1F);
    for (int e = t; e < 4 * 8 * 32; e += NT) {
      const int l = e & 31, nt = (e >> 5) & 7, ks = e >> 8, gg = l >> 2, tt = l & 3;
      const int i0 = 8 * ks + tt, i1 = i0 + 4, j = 8 * nt + gg;
      w1f[e] = make_float2(i0 < I ? W1[i0 * H + j] : 0.f, i1 < I ? W1[i1 * H + j] : 0.f);
    }
    {
      const int l = t & 31, q = t >> 5, gg = l >> 2, tt = l & 3;
      reinterpret_cast<float2 *>(sm + M::W3F)[t] = make_float2(gg < O ? W3[(8 * q + tt) * O + gg] : 0.f, gg < O ? W3[(8 * q + tt + 4) * O + gg] : 0.f);
    }
    __syncthreads();
    float2 *w2f = reinterpret_cast<float2 *>(W2);
#pragma unroll
    for (int u = 0; u < 8; ++u) w2f[t + u * NT] = f2[u];
  }
  const float2 *W2F = reinterpret_cast<const float2 *>(sm + M::P + off_W2(I));
  const float2 *W1F = reinterpret_cast<const float2 *>(sm + M::W1F), *W3F = reinterpret_cast<const float2 *>(sm + M::W3F);
  const int ks1 = (I + 7) >> 3;
  const uint32_t inv_I = (65536u + (uint32_t)I - 1u) / (uint32_t)I;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    bool same = has_alt(tile);
    if (same) {   // bitwise comparison of the staged sp tile with the staged s[t+1] tile
      int diff = 0;
      for (int e = t; e < R * I; e += NT) diff |= __float_as_int(ST[e]) != __float_as_int(ST[R * MAX_I + e]);
      same = __syncthreads_or(diff) == 0;
    }
    if (same) {   // uniform for the CTA
      if (t < R) a.y[tile * R + t] = a.y_alt[tile * R + t];   // O == 1 on this path (checked by the launcher)
      __syncthreads();   // the staging area is free again
      if (tile + gridDim.x < n_tiles) issue_tile(tile + gridDim.x);
      continue;
    }
    for (int e = t; e < R * I; e += NT) {
      const int r = (int)(((uint32_t)e * inv_I) >> 16), i = e - r * I;
      XT[i * LD + r] = ST[e];
    }
    __syncthreads();
    if (tile + gridDim.x < n_tiles) issue_tile(tile + gridDim.x);
    {
      float c[4][4];
      mma_rows<4>(XT, ks1, r0, W1F + (4 * wn) * 32 + lane, 8 * 32, c);
      const float *b1 = P + off_b1(I);
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = 32 * wn + 8 * q + 2 * tq + (e & 1), r = r0 + g + 8 * (e >> 1);
          H1T[j * LD + r] = act_fused(act, c[q][e] + b1[j]);
        }
    }
    __syncthreads();
    {
      float c[4][4];
      mma_rows<4>(H1T, 8, r0, W2F + (4 * wn) * 32 + lane, 8 * 32, c);
      const float *b2 = P + off_b2(I);
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = 32 * wn + 8 * q + 2 * tq + (e & 1), r = r0 + g + 8 * (e >> 1);
          H2T[j * LD + r] = act_fused(act, c[q][e] + b2[j]);
        }
    }
    __syncthreads();
    if (w < 4) {
      float c[1][4];
      mma_rows<1>(H2T, 8, r0, W3F + lane, 32, c);
      const float *b3 = P + off_b3(I, O);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int o = 2 * tq + (e & 1);
        const int64_t row = tile * R + r0 + g + 8 * (e >> 1);
        if (o < O && row < a.B) a.y[row * O + o] = c[0][e] + b3[o];
      }
    }
  }
}

// `train!` tail (training.jl:18-23) + the loss bookkeeping of the minibatch, one launch:
//   gnorm = ||all grads||_2 (every CTA recomputes it, identical order -> identical value), NaN -> sticky error flag and no update,
//   info record (block 0), KL early-stop vote (block 0), Flux Adam on this CTA's slice of the parameters.
// Flux's Adam keeps the β-powers as state and multiplies them by β after every step (βp .*= β, [3P] Optimise.Adam); `pow(β, t)` in
// double precision costs a single thread ~10 us on this part (hundreds of dependent FP64 instructions) and sat on the critical path of
// every update tail.  The powers of the last step are cached behind the gradient-norm partials: cache = {t, β1^t, β2^t}.
constexpr int BETA_CACHE = 1021;   // doubles [1021, 1024) of crux_mlp::norm_part
__device__ __forceinline__ void beta_pows(const double *cache, int t, double b1, double b2, double &p1, double &p2) {
  const int ct = (int)__ldcg(cache);
  if (ct == t) { p1 = __ldcg(cache + 1); p2 = __ldcg(cache + 2); }
  else if (ct == t - 1 && t > 1) { p1 = __ldcg(cache + 1) * b1; p2 = __ldcg(cache + 2) * b2; }
  else { p1 = pow(b1, (double)t); p2 = pow(b2, (double)t); }
}
__device__ __forceinline__ void beta_cache_store(double *cache, int t, double p1, double p2) { cache[1] = p1; cache[2] = p2; cache[0] = (double)t; }

struct AdamArgs {
  float *p, *g, *m, *v; int n;                       // network parameters
  float *ls, *ls_g, *ls_m, *ls_v; int A;             // actor only: logΣ vector (A == 0 for a critic)
  const float *sums;                                 // tail sums (after the optional all-reduce): obj, kl, clip, adv, ret, count
  double eta, b1, b2, eps;
  const int *step_dev;
  float lambda_p, lambda_e, target_kl;
  int a2c, head;                                     // head 0: actor (ppo/a2c), 1: critic (mse)
  float *rec;                                        // info record of this minibatch
  const double *norm_part; int n_norm_part;          // per-CTA sums of squares from the reduce kernel (NULL: recompute)
  const double *beta_cache;                          // {t, β1^t, β2^t} of the step the reduce kernel has just counted
  // fused gradient all-reduce over NVLink peer memory (LL protocol): the gradient is the rank-ordered sum of the 8-byte words
  // {value, sequence number} every rank's reduce kernel stored into THIS rank's receive region of this network (double-buffered by
  // the parity of the device-resident sequence number); the last CTA of the Adam kernel advances the sequence number
  int peer, world; int64_t peer_cap;
  const unsigned long long *ll_recv; unsigned long long *ll_seq; unsigned int *ll_ticket;
  int *ctl; int mb;
  unsigned int *err_flags;
  float *frag; int fI, fO;                           // fragment buffer of the network (NULL: none) and its input / output widths
  unsigned long long *trace;                         // development aid (CRUX_MB6_TRACE): [0] first CTA start, [1] last CTA end
  int frag_mode;                                     // 0: MMA B-fragment order (mma.sync kernel), 1: tcgen05 hi/lo weight planes (mb_t5.cuh)
};
__device__ __forceinline__ void adam_body(const AdamArgs &a, int block_rank, int n_blocks);
struct PeerOut {   // where the reduce kernel stores this rank's gradient for the fused all-reduce (enabled == 0: local only)
  unsigned long long *ll[16];          // every rank's LL region of this network: [2 parities][16 ranks][cap] words
  int world, rank, enabled; int64_t cap;
  const unsigned long long *seq_dev;   // sequence number of the last completed exchange of this network
};

// sum the per-CTA partials (double accumulation, fixed order: bit-reproducible) -> gradient vector + tail
//   grads[p]                    p < n_params
//   grads[n_params + j]         j < 8  : dL/dlogΣ_j           (tail_ls_grad)
//   grads[n_params + 64 + q]    q < 5  : obj, kl, clip, adv, ret sums ; [5] = row count   (tail_sums)
// CTA = 32 parameters x 32 warps; warp w sums partials w, w+32, ... (all loads of a thread are independent: one L2 round
// trip), the 32 sub-sums are combined in a fixed order.  The logΣ entries get the entropy term d(λe·e_loss)/dlogΣ = -λe here
// (scaled by 1/world so that the all-reduce sum restores it).  Each CTA also emits the sum of squares of its 32 finished
// gradient entries (used by the Adam kernel when there is no all-reduce in between).  Block 0 counts the optimiser step.
constexpr int RW = 32;  // warps per reduce CTA
// FUSE = 1 (opt-in CRUX_FUSE_ADAM) compiles the Adam tail into the kernel; the default instantiation stays at 32 registers so that
// two 1024-thread CTAs share an SM (with the tail inlined it needs 60 and the 179-CTA grid runs in two waves: 14 us instead of 9).
template <int FUSE>
__global__ void __launch_bounds__(RW * 32, FUSE ? 1 : 2) reduce_fused_partials_kernel(const float *__restrict__ partials, int nparts, int pstride, int n_params,
                                                                       float *__restrict__ grads, float count, float ls_shift, int n_ls,
                                                                       double *__restrict__ norm_part, int *__restrict__ step_dev,
                                                                       const int *__restrict__ ctl, int mb, unsigned int *__restrict__ ticket,
                                                                       AdamArgs adam, int fuse_adam, PeerOut peer) {
  if (stopped(ctl, mb)) return;
  unsigned long long pseq = 0ULL;
  if (peer.enabled) pseq = *(volatile const unsigned long long *)peer.seq_dev + 1ULL;   // advanced by the Adam kernel that consumes this exchange
  const int64_t pbase = ((int64_t)(pseq & 1ULL) * 16 + peer.rank) * peer.cap;
  auto emit = [&](int idx, float v) {   // local gradient entry + (fused all-reduce) one {value, sequence} word in every rank's receive region
    grads[idx] = v;
    if (peer.enabled) {
      const unsigned long long word = ((unsigned long long)(unsigned int)pseq << 32) | (unsigned long long)__float_as_uint(v);
      // single 8-byte stores: no fence, no flag.  (Parking the words in shared memory and letting warp q store to rank q measured
      // slower: 15.0 us against 13.4 us per launch on 2 GPUs.)
      for (int q = 0; q < peer.world; ++q) *(volatile unsigned long long *)(peer.ll[q] + pbase + idx) = word;
    }
  };
  __shared__ double sh[RW][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  // groups of 32 entries, grid-strided: the grid may be much smaller than the number of groups (several ranks: the kernel runs on
  // the few SMs the concurrent minibatch kernel of the other network leaves free); results do not depend on the grid size
  const int n_groups = (n_params + 16 + 31) / 32;
  for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const int p = grp * 32 + lane;
    double s = 0.0;
    if (p < n_params + 16) {
      const float *src = partials + p;
#pragma unroll 10
      for (int c = w; c < nparts; c += RW) s += (double)__ldcg(src + (int64_t)c * pstride);
    }
    sh[w][lane] = s;
    __syncthreads();
    if (w == 0) {
      double sq = 0.0;
      if (p < n_params + 16) {
        double t = 0.0;
#pragma unroll
        for (int q = 0; q < RW; ++q) t += sh[q][lane];
        if (p < n_params) { emit(p, (float)t); sq = (double)(float)t * (double)(float)t; }
        else if (p < n_params + 8) {
          const int j = p - n_params;
          const float g = (float)t + (j < n_ls ? ls_shift : 0.f);
          emit(p, g);
          if (j < n_ls) sq = (double)g * (double)g;
        } else {
          const int q = p - n_params - 8;
          if (q < 5) emit(n_params + 64 + q, (float)t);
          else if (q == 5) emit(n_params + 64 + 5, count);
        }
      }
      sq = warp_sum_d(sq);
      if (lane == 0) norm_part[grp] = sq;
    }
    __syncthreads();   // sh is reused by the next group
  }
  if (!FUSE || !fuse_adam) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      const int t = *step_dev + 1;
      double p1, p2;
      beta_pows(norm_part + BETA_CACHE, t, adam.b1, adam.b2, p1, p2);
      beta_cache_store(norm_part + BETA_CACHE, t, p1, p2);
      *step_dev = t;
    }
    return;
  }
  // single GPU: no all-reduce follows, so the LAST CTA to finish runs the norm / record / Adam tail right here (one launch less)
  __shared__ bool last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    last = atomicAdd(ticket, 1u) == gridDim.x - 1u;
    if (last) {
      *ticket = 0u;
      const int t = *step_dev + 1;
      double p1, p2;
      beta_pows(norm_part + BETA_CACHE, t, adam.b1, adam.b2, p1, p2);
      beta_cache_store(norm_part + BETA_CACHE, t, p1, p2);
      *step_dev = t;
    }
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  adam_body(adam, 0, 1);
}

// block_rank / n_blocks: the slice of the parameter vector this CTA updates (the norm and the record are computed by every CTA /
// by CTA 0).  Gradients are read with ld.global.cg: they may have been written by other CTAs of the same launch.
__device__ __forceinline__ void adam_body(const AdamArgs &a, int block_rank, int n_blocks) {
  __shared__ double sh[32];
  __shared__ double s_n2, s_c1, s_c2;
  const int nw = blockDim.x >> 5;
  const unsigned long long *slot0 = nullptr;
  unsigned int seq32 = 0u;
  if (a.peer) {
    const unsigned long long seq = *(volatile const unsigned long long *)a.ll_seq + 1ULL;   // the exchange the reduce kernels just sent
    seq32 = (unsigned int)seq;
    slot0 = a.ll_recv + (int64_t)(seq & 1ULL) * 16 * a.peer_cap;
  }
  // gradient entry i (network grads, then the tail): local, or the rank-ordered sum of the ranks' words (identical on every rank);
  // a word is valid once its upper half carries this exchange's sequence number
  auto G = [&](const float *local, int64_t i) -> float {
    if (!slot0) return __ldcg(local);
    float sum = 0.f;
    for (int q = 0; q < a.world; ++q) {
      const volatile unsigned long long *p = slot0 + (int64_t)q * a.peer_cap + i;
      unsigned long long x = *p;
      while ((unsigned int)(x >> 32) != seq32) x = *p;
      sum += __uint_as_float((unsigned int)x);
    }
    return sum;
  };
  const int64_t off_ls = a.n, off_sums = (int64_t)a.n + 64;
  // ---- gradient norm over (network grads, logΣ grads incl. the entropy term): from the reduce kernel's per-CTA sums of
  //      squares when nothing changed the gradient in between, else recomputed here (after an all-reduce)
  double s = 0.0;
  if (a.norm_part) {
    for (int i = threadIdx.x; i < a.n_norm_part; i += blockDim.x) s += __ldcg(a.norm_part + i);
  } else {
    for (int i = threadIdx.x; i < a.n; i += blockDim.x) { const double v = (double)G(a.g + i, i); s += v * v; }
    if (threadIdx.x < a.A) { const double v = (double)G(a.ls_g + threadIdx.x, off_ls + threadIdx.x); s += v * v; }
  }
  s = warp_sum_d(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int q = 0; q < nw; ++q) t += sh[q];
    s_n2 = t;
    const int step = *a.step_dev;  // counts this step (incremented by the reduce kernel)
    double p1, p2;
    beta_pows(a.beta_cache, step, a.b1, a.b2, p1, p2);
    s_c1 = 1.0 - p1;
    s_c2 = 1.0 - p2;
  }
  __syncthreads();
  const double n2 = s_n2;
  const bool bad = isnan(n2);
  // ---- info record + early-stop vote (first CTA, before any parameter changes)
  if (block_rank == 0 && threadIdx.x == 0) {
    const float cnt = G(a.sums + 5, off_sums + 5);
    if (a.head == 0) {
      float sls = 0.f;
      for (int j = 0; j < a.A; ++j) sls += a.ls[j];
      const float entropy = 1.4189385332046727f + sls;          // policies.jl:348
      const float p_loss = -(G(a.sums + 0, off_sums + 0) / cnt);
      a.rec[CRUX_PPO_LOSS] = a.lambda_p * p_loss + a.lambda_e * (-entropy);
      a.rec[CRUX_PPO_ENTROPY] = entropy;
      const float kl = G(a.sums + 1, off_sums + 1) / cnt;
      a.rec[CRUX_PPO_KL] = kl;
      a.rec[CRUX_PPO_CLIP_FRAC] = a.a2c ? 0.f : G(a.sums + 2, off_sums + 2) / cnt;
      a.rec[CRUX_PPO_AVG_ADV] = G(a.sums + 3, off_sums + 3) / cnt;
      a.rec[CRUX_PPO_AVG_RET] = G(a.sums + 4, off_sums + 4) / cnt;
      if (a.ctl && kl > a.target_kl) a.ctl[1] = a.mb + 1;         // this minibatch is still applied; later ones are skipped
    } else {
      a.rec[CRUX_PPO_LOSS] = G(a.sums + 0, off_sums + 0) / cnt;
    }
    a.rec[CRUX_PPO_GRAD_NORM] = (float)sqrt(n2);
    a.rec[CRUX_PPO_VALID] = 1.f;
    if (bad) atomicOr(a.err_flags, CRUX_FLAG_NAN);                // training.jl:20: error before Flux.update!
  }
  if (bad) return;
  __syncthreads();
  // ---- Flux Adam (float32 moments, Float64 scalars)
  const double c1 = s_c1, c2 = s_c2;
  for (int i = block_rank * blockDim.x + threadIdx.x; i < a.n; i += n_blocks * blockDim.x) {
    const float gf = G(a.g + i, i);
    if (slot0) a.g[i] = gf;   // keep the local gradient vector observable (crux_mlp_grads_ptr)
    const double g = (double)gf;
    const float mt = (float)(a.b1 * (double)a.m[i] + (1.0 - a.b1) * g);
    const float vt = (float)(a.b2 * (double)a.v[i] + (1.0 - a.b2) * g * g);
    a.m[i] = mt; a.v[i] = vt;
    const float pn = a.p[i] - (float)((double)mt / c1 / (sqrt((double)vt / c2) + a.eps) * a.eta);
    a.p[i] = pn;
    if (a.frag) { if (a.frag_mode) mb6::plane_scatter(a.frag, a.fI, a.fO, i, pn); else frag_scatter(a.frag, a.fI, a.fO, i, pn); }   // what the next minibatch kernel stages
  }
  if (block_rank == 0 && threadIdx.x < a.A) {
    const int i = threadIdx.x;
    const double g = (double)G(a.ls_g + i, off_ls + i);
    const float mt = (float)(a.b1 * (double)a.ls_m[i] + (1.0 - a.b1) * g);
    const float vt = (float)(a.b2 * (double)a.ls_v[i] + (1.0 - a.b2) * g * g);
    a.ls_m[i] = mt; a.ls_v[i] = vt;
    a.ls[i] = a.ls[i] - (float)((double)mt / c1 / (sqrt((double)vt / c2) + a.eps) * a.eta);
  }
}

__global__ void __launch_bounds__(256) fused_adam_kernel(AdamArgs a) {
  if (stopped(a.ctl, a.mb)) return;
  adam_body(a, blockIdx.x, gridDim.x);
  if (a.peer) {   // the LAST CTA (every other one has read the sequence number long ago) closes this network's exchange
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(a.ll_ticket, 1u) == gridDim.x - 1u) {
      *a.ll_ticket = 0u;
      *a.ll_seq = *a.ll_seq + 1ULL;
    }
  }
}

// The Adam kernel of the fused LL gradient exchange: thread i owns gradient entry i from the first poll to the parameter update.
//   1. poll the `world` 8-byte {value, sequence} words of entry i (all loads in flight at once, re-polled until each carries this
//      exchange's sequence number) and sum them in rank order -> g_i, identical on every rank
//   2. ||g||^2: block sums of squares (double, fixed order) -> norm_slots[block]; the last block to arrive publishes the exchange's
//      sequence number in `ready`, every block spins on it and adds the slots in block order (the CTAs of this small grid become
//      co-resident as the concurrent minibatch kernel retires CTAs; they wait for nothing but each other)
//   3. record / KL vote (block 0), Flux Adam on entry i, fragment scatter; block 0 closes the exchange (sequence number + 1)
// against adam_body's LL path (every CTA polling the whole vector for the norm): 27.9 us -> see profiles/r1_notes.md.
__global__ void __launch_bounds__(256) fused_adam_ll_kernel(AdamArgs a, double *__restrict__ norm_slots, unsigned int *__restrict__ ticket,
                                                            unsigned long long *__restrict__ ready) {
  if (stopped(a.ctl, a.mb)) return;
  __shared__ double sh[8];
  __shared__ double s_n2, s_c1, s_c2;
  __shared__ float s_tail[16];   // [0..8) dlogΣ, [8..14) obj, kl, clip, adv, ret, count
  const int tid = threadIdx.x;
  const unsigned long long seq = *(volatile const unsigned long long *)a.ll_seq + 1ULL;   // the exchange the reduce kernels just sent
  const unsigned int seq32 = (unsigned int)seq;
  const unsigned long long *slot0 = a.ll_recv + (int64_t)(seq & 1ULL) * 16 * a.peer_cap;
  auto poll = [&](int64_t idx) -> float {
    float sum = 0.f;
    for (int base = 0; base < a.world; base += 8) {
      unsigned long long w[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (base + u < a.world) w[u] = *(const volatile unsigned long long *)(slot0 + (int64_t)(base + u) * a.peer_cap + idx);
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (base + u < a.world) {
          while ((unsigned int)(w[u] >> 32) != seq32) w[u] = *(const volatile unsigned long long *)(slot0 + (int64_t)(base + u) * a.peer_cap + idx);
          sum += __uint_as_float((unsigned int)w[u]);
        }
    }
    return sum;
  };
  const int i = blockIdx.x * 256 + tid;
  const int64_t off_ls = a.n, off_sums = (int64_t)a.n + 64;
  const float gf = i < a.n ? poll(i) : 0.f;
  double sq = (double)gf * (double)gf;
  if (blockIdx.x == 0) {
    if (tid < a.A) { const float g = poll(off_ls + tid); s_tail[tid] = g; sq += (double)g * (double)g; }
    else if (tid >= 32 && tid < 38) s_tail[8 + tid - 32] = poll(off_sums + (tid - 32));
  }
  sq = warp_sum_d(sq);
  if ((tid & 31) == 0) sh[tid >> 5] = sq;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int q = 0; q < 8; ++q) t += sh[q];
    norm_slots[blockIdx.x] = t;
    __threadfence();
    if (atomicAdd(ticket, 1u) == gridDim.x - 1u) {
      *ticket = 0u;
      __threadfence();
      *(volatile unsigned long long *)ready = seq;
    }
    while (*(volatile unsigned long long *)ready != seq) { }
    __threadfence();
    t = 0.0;
    for (unsigned int b = 0; b < gridDim.x; ++b) t += __ldcg(norm_slots + b);
    s_n2 = t;
    const int step = *a.step_dev;  // counts this step (incremented by the reduce kernel)
    double p1, p2;
    beta_pows(a.beta_cache, step, a.b1, a.b2, p1, p2);
    s_c1 = 1.0 - p1;
    s_c2 = 1.0 - p2;
  }
  __syncthreads();
  const double n2 = s_n2;
  const bool bad = isnan(n2);
  if (blockIdx.x == 0 && tid == 0) {
    // every block has read the sequence number and all of its words (it published its slot afterwards): close the exchange
    *(volatile unsigned long long *)a.ll_seq = seq;
    const float cnt = s_tail[8 + 5];
    if (a.head == 0) {
      float sls = 0.f;
      for (int j = 0; j < a.A; ++j) sls += a.ls[j];
      const float entropy = 1.4189385332046727f + sls;          // policies.jl:348
      const float p_loss = -(s_tail[8 + 0] / cnt);
      a.rec[CRUX_PPO_LOSS] = a.lambda_p * p_loss + a.lambda_e * (-entropy);
      a.rec[CRUX_PPO_ENTROPY] = entropy;
      const float kl = s_tail[8 + 1] / cnt;
      a.rec[CRUX_PPO_KL] = kl;
      a.rec[CRUX_PPO_CLIP_FRAC] = a.a2c ? 0.f : s_tail[8 + 2] / cnt;
      a.rec[CRUX_PPO_AVG_ADV] = s_tail[8 + 3] / cnt;
      a.rec[CRUX_PPO_AVG_RET] = s_tail[8 + 4] / cnt;
      if (a.ctl && kl > a.target_kl) a.ctl[1] = a.mb + 1;         // this minibatch is still applied; later ones are skipped
    } else {
      a.rec[CRUX_PPO_LOSS] = s_tail[8 + 0] / cnt;
    }
    a.rec[CRUX_PPO_GRAD_NORM] = (float)sqrt(n2);
    a.rec[CRUX_PPO_VALID] = 1.f;
    if (bad) atomicOr(a.err_flags, CRUX_FLAG_NAN);                // training.jl:20: error before Flux.update!
  }
  if (bad) return;
  const double c1 = s_c1, c2 = s_c2;
  if (i < a.n) {
    a.g[i] = gf;   // keep the local gradient vector observable (crux_mlp_grads_ptr)
    const double g = (double)gf;
    const float mt = (float)(a.b1 * (double)a.m[i] + (1.0 - a.b1) * g);
    const float vt = (float)(a.b2 * (double)a.v[i] + (1.0 - a.b2) * g * g);
    a.m[i] = mt; a.v[i] = vt;
    const float pn = a.p[i] - (float)((double)mt / c1 / (sqrt((double)vt / c2) + a.eps) * a.eta);
    a.p[i] = pn;
    if (a.frag) { if (a.frag_mode) mb6::plane_scatter(a.frag, a.fI, a.fO, i, pn); else frag_scatter(a.frag, a.fI, a.fO, i, pn); }
  }
  if (blockIdx.x == 0 && tid < a.A) {
    const double g = (double)s_tail[tid];
    const float mt = (float)(a.b1 * (double)a.ls_m[tid] + (1.0 - a.b1) * g);
    const float vt = (float)(a.b2 * (double)a.ls_v[tid] + (1.0 - a.b2) * g * g);
    a.ls_m[tid] = mt; a.ls_v[tid] = vt;
    a.ls[tid] = a.ls[tid] - (float)((double)mt / c1 / (sqrt((double)vt / c2) + a.eps) * a.eta);
  }
}

// Single-GPU tail of a t5 minibatch in ONE launch: per-CTA partials -> gradient (double accumulation, fixed order) -> Flux Adam on the
// 32 entries the CTA has just finished -> plane scatter; the last CTA to finish (ticket) adds the per-CTA sums of squares in a fixed
// order and writes the info record / KL vote / NaN flag.  Why it matters: the minibatch kernel leaves no room on an SM for another
// CTA (640 threads x 96 registers), so the tail of one network cannot hide behind the other network's minibatch kernel -- the GPU
// idles for the whole reduce -> Adam chain (measured with %globaltimer, scripts/mb6_trace.py: 17 us per minibatch pair with two
// launches).  Adam itself does not need the gradient norm (train! only logs it, training.jl:18-23); the reference's "NaN -> error
// BEFORE the update" is kept through a NaN flag the minibatch kernel raises while publishing its partials (state[2]).
// state: [0] Adam steps applied, [1] ticket, [2] NaN seen in a partial.
// RWT = 4 warps: two 128-thread CTAs at <= 56 registers fit NEXT TO a resident minibatch CTA (640 threads x 80 registers, 157 KB), so that
// the tail of one network runs -- in a single wave -- under the other network's minibatch kernel instead of after it.
template <int RWT>
__global__ void __launch_bounds__(RWT * 32, 9) reduce_adam_kernel(const float *__restrict__ partials, int nparts, int pstride, int n_params,
                                                                 float *__restrict__ grads, float count, float ls_shift, int n_ls,
                                                                 double *__restrict__ norm_part, int *__restrict__ state, AdamArgs a, PeerOut peer) {
  static_assert(RWT == 4, "16 row subsets = 4 warps x 4 lane groups");
  asm volatile("griddepcontrol.wait;" ::: "memory");                // the minibatch kernel's partials (and its KL-stop flag) are complete
#define TAIL_TRACE(slot) do { if (a.trace) { unsigned long long gt_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_)); a.trace[slot] = gt_; } } while (0)
  if (blockIdx.x == 0 && threadIdx.x == 0) TAIL_TRACE(0);
  // Latency is what this kernel costs (it sits on its network's dependency chain, scripts/mb6_trace.py).  The CTA's whole slice of the
  // partials is requested at once (10 x 16 bytes per thread in flight); everything else it will need -- the optimiser state of its 32
  // entries, the step counter, the β-power cache, the stop flag -- is PREFETCHED into L1 first (no destination registers: at 56
  // registers per thread, loads would be spilled, and a spill store waits for its load) and read once the partial rows are summed.
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int n_groups = (n_params + 16 + 31) / 32;
  auto prefetch = [](const void *ptr) { asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr)); };
  if (w == 0 && blockIdx.x < n_groups) {   // warp 0 owns the finished entries
    const int p = blockIdx.x * 32 + lane;
    if (p < n_params) { prefetch(a.p + p); prefetch(a.m + p); prefetch(a.v + p); }
    else if (p < n_params + n_ls) { const int j = p - n_params; prefetch(a.ls + j); prefetch(a.ls_m + j); prefetch(a.ls_v + j); }
  } else if (w == 1 && lane < 3) {
    if (lane == 0) prefetch(state);
    else if (lane == 1) prefetch(norm_part + BETA_CACHE);
    else if (a.ctl) prefetch(a.ctl + 1);
  }
  // several GPUs: the gradient all-reduce is fused in with the LL (flag-in-data) protocol.  The CTA that finishes entry i stores it as one
  // 8-byte {value, sequence} word into the receive region of EVERY rank over NVLink and then polls the `world` words of the same entry
  // in its own region: rank-ordered sum (bit-identical on all ranks) -> Adam on that entry.  No fence, no flag, no collective launch.
  unsigned int pseq = 0u;   // low word of the sequence number (advanced by the last CTA of this kernel); bit 0 selects the receive buffer
  auto exchange = [&](int idx, float v) -> float {
    if (!peer.enabled) return v;
    const int64_t pbase = ((int64_t)(pseq & 1u) * 16 + peer.rank) * peer.cap;
    const unsigned long long word = ((unsigned long long)pseq << 32) | (unsigned long long)__float_as_uint(v);
    for (int q = 0; q < peer.world; ++q) *(volatile unsigned long long *)(peer.ll[q] + pbase + idx) = word;
    float sum = 0.f;
    const volatile unsigned long long *src = a.ll_recv + (int64_t)(pseq & 1u) * 16 * a.peer_cap + idx;
    for (int base = 0; base < peer.world; base += 8) {   // the words of 8 ranks in flight at once, summed in rank order
      unsigned long long wd[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (base + u < peer.world) wd[u] = src[(int64_t)(base + u) * a.peer_cap];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (base + u < peer.world) {
          while ((unsigned int)(wd[u] >> 32) != pseq) wd[u] = src[(int64_t)(base + u) * a.peer_cap];
          sum += __uint_as_float((unsigned int)wd[u]);
        }
    }
    return sum;
  };
  __shared__ double sh[RWT][33];
  __shared__ float4 stage[10][RWT * 32];   // this CTA's slice of the partials: [row k of the thread][thread]
  __shared__ double s_c1, s_c2;
  __shared__ int s_bad;
  __shared__ bool last;
  // partial rows: thread (w, sub = lane >> 3) adds rows 4w + sub, + 16, ... of the four entries 4 (lane & 7) .. + 3 -- 128-bit loads,
  // double accumulation in a fixed order (lane groups by a shuffle tree, then the four warps in order): bit-reproducible
  const int q4 = (lane & 7) * 4, sub = lane >> 3, r0 = w * 4 + sub;
  float pm = 0.f, pv = 0.f, pp = 0.f;
  for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const float *src = partials + (int64_t)grp * 32 + q4 + (int64_t)r0 * pstride;   // grp * 32 + 31 < pstride: whole float4s inside the row
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    const bool first = grp == (int)blockIdx.x;
    if (nparts <= 160) {
      // the slice lands in shared memory through cp.async: ten 16-byte requests per thread in flight without a single destination register
#pragma unroll
      for (int k = 0; k < 10; ++k)
        if (r0 + 16 * k < nparts)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(&stage[k][threadIdx.x])), "l"(src + (int64_t)(16 * k) * pstride) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (first) {   // everything is in flight: now the first dependent instructions
        const int stop_at = a.ctl ? *(volatile const int *)(a.ctl + 1) : 0;
        if (peer.enabled) pseq = (unsigned int)(*(volatile const unsigned long long *)peer.seq_dev + 1ULL);
        if (stop_at != 0 && stop_at <= a.mb) {   // an EARLIER minibatch raised the KL stop flag (uniform over the grid): nothing is written
          asm volatile("cp.async.wait_group 0;" ::: "memory");
          return;
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) TAIL_TRACE(2);
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      if (blockIdx.x == 0 && threadIdx.x == 0) TAIL_TRACE(7);
#pragma unroll
      for (int k = 0; k < 10; ++k)
        if (r0 + 16 * k < nparts) { const float4 v = stage[k][threadIdx.x]; s0 += (double)v.x; s1 += (double)v.y; s2 += (double)v.z; s3 += (double)v.w; }
    } else {   // (more partial rows than any current GPU has SMs)
      if (first) {
        const int stop_at = a.ctl ? *(volatile const int *)(a.ctl + 1) : 0;
        if (peer.enabled) pseq = (unsigned int)(*(volatile const unsigned long long *)peer.seq_dev + 1ULL);
        if (stop_at != 0 && stop_at <= a.mb) return;
      }
      for (int r = r0; r < nparts; r += 16) {
        const float4 v = __ldcg(reinterpret_cast<const float4 *>(src + (int64_t)(r - r0) * pstride));
        s0 += (double)v.x; s1 += (double)v.y; s2 += (double)v.z; s3 += (double)v.w;
      }
    }
    if (w == 0) {   // optimiser state of this group's entries (L1 hits for the CTA's first group)
      const int p = grp * 32 + lane;
      if (p < n_params) { pp = a.p[p]; pm = a.m[p]; pv = a.v[p]; }
      else if (p < n_params + n_ls) { const int j = p - n_params; pp = a.ls[j]; pm = a.ls_m[j]; pv = a.ls_v[j]; }
    } else if (w == 1 && first && lane == 0) {   // β^t of this step from the cache of the previous one (the cache is advanced by the last CTA, after every CTA has read it)
      const int step = *(volatile const int *)state + 1;
      const int nan_seen = *(volatile const int *)(state + 2);
      const volatile double *bc = norm_part + BETA_CACHE;
      const double bc0 = bc[0], bc1 = bc[1], bc2 = bc[2];
      double p1, p2;
      if ((int)bc0 == step) { p1 = bc1; p2 = bc2; }
      else if ((int)bc0 == step - 1 && step > 1) { p1 = bc1 * a.b1; p2 = bc2 * a.b2; }
      else { p1 = pow(a.b1, (double)step); p2 = pow(a.b2, (double)step); }
      s_c1 = 1.0 - p1; s_c2 = 1.0 - p2; s_bad = nan_seen;   // read by warp 0 behind the barrier below
    }
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o); s3 += __shfl_xor_sync(0xffffffffu, s3, o);
    }
    if (sub == 0) { sh[w][q4] = s0; sh[w][q4 + 1] = s1; sh[w][q4 + 2] = s2; sh[w][q4 + 3] = s3; }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) TAIL_TRACE(3);
    if (w == 0) {
      const int p = grp * 32 + lane;
      const double c1 = s_c1, c2 = s_c2;
      const bool bad = s_bad != 0;
      double sq = 0.0;
      if (p < n_params + 16) {
        double t = 0.0;
#pragma unroll
        for (int q = 0; q < RWT; ++q) t += sh[q][lane];
        float g = (float)t;
        if (p >= n_params && p < n_params + n_ls) g += ls_shift;   // d(λe e_loss)/dlogΣ = -λe, scaled by 1/world: the sum over ranks restores it
        // entries < n_params + 8: gradient (+ dlogΣ); then obj, kl, clip, adv, ret | count | sum(logΣ) BEFORE this update (published by CTA 0 of
        // the minibatch kernel; identical on every rank, not exchanged).  ONE exchange call site: the polling loop is inlined once.
        const int q = p - n_params - 8;
        const int slot = q < 0 ? p : n_params + 64 + q;
        if (q == 5) g = count;
        if (q < 6) g = exchange(slot, g);
        if (q < 7) grads[slot] = g;
        if (p < n_params + n_ls) sq = (double)g * (double)g;
        if (!bad && p < n_params + n_ls) {   // Flux Adam on this entry (float32 moments, Float64 scalars)
          float *ap, *am, *av;
          if (p < n_params) { ap = a.p + p; am = a.m + p; av = a.v + p; } else { const int j = p - n_params; ap = a.ls + j; am = a.ls_m + j; av = a.ls_v + j; }
          const double gd = (double)g;
          const float mt = (float)(a.b1 * (double)pm + (1.0 - a.b1) * gd);
          const float vt = (float)(a.b2 * (double)pv + (1.0 - a.b2) * gd * gd);
          *am = mt; *av = vt;
          const float pn = pp - (float)((double)mt / c1 / (sqrt((double)vt / c2) + a.eps) * a.eta);
          *ap = pn;
          if (a.frag && p < n_params) { if (a.frag_mode) mb6::plane_scatter(a.frag, a.fI, a.fO, p, pn); else frag_scatter(a.frag, a.fI, a.fO, p, pn); }
        }
      }
      sq = warp_sum_d(sq);
      if (lane == 0) norm_part[grp] = sq;
    }
    __syncthreads();   // sh is reused by the next group
  }
  if (blockIdx.x >= n_groups && a.ctl && *(volatile const int *)(a.ctl + 1) != 0 && *(volatile const int *)(a.ctl + 1) <= a.mb) return;   // (a CTA without a group)
  if (blockIdx.x == 0 && threadIdx.x == 0) TAIL_TRACE(4);
  // parameters, moments and planes of this CTA's entries are on their way: this network's next minibatch kernel may be scheduled.  It runs its
  // weight-independent prologue (barriers, tensor memory, the first tile's gather) under the record keeping below and reads the weights
  // behind its own griddepcontrol.wait, i.e. after this grid has completed.  (Triggered at the START of this kernel, the dependent took
  // over every SM the other network's minibatch kernel was about to get and idled there: 1.40 against 1.34 ms per iteration.)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(reinterpret_cast<unsigned int *>(state + 1), 1u) == gridDim.x - 1u;
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0) TAIL_TRACE(5);
  if (!last) return;
  if (threadIdx.x == 0) TAIL_TRACE(6);
  __threadfence();
  if (w == 0) {   // ||g||^2: lane l adds groups l, l + 32, ... in order, the 32 lane sums are combined by a fixed shuffle tree: bit-reproducible
    const float *sums = grads + n_params + 64;
    float sv[7];
#pragma unroll
    for (int q = 0; q < 7; ++q) sv[q] = lane == 0 ? __ldcg(sums + q) : 0.f;   // obj, kl, clip, adv, ret | rows of this minibatch over all ranks | sum(logΣ)
    double np[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) np[k] = lane + 32 * k < n_groups ? __ldcg(norm_part + lane + 32 * k) : 0.0;
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += np[k];
    for (int q = lane + 256; q < n_groups; q += 32) t += __ldcg(norm_part + q);
    t = warp_sum_d(t);
    if (lane == 0) {
      const int step = __ldcg(state) + 1;
      double c1 = s_c1, c2 = s_c2;
      int nan_seen = s_bad;
      if (blockIdx.x >= n_groups) {   // (a last CTA that owned no group has not derived the β powers)
        double p1, p2;
        beta_pows(norm_part + BETA_CACHE, step, a.b1, a.b2, p1, p2);
        c1 = 1.0 - p1; c2 = 1.0 - p2; nan_seen = __ldcg(state + 2);
      }
      beta_cache_store(norm_part + BETA_CACHE, step, 1.0 - c1, 1.0 - c2);
      state[0] = step; state[1] = 0; state[2] = 0;
      if (peer.enabled) *(volatile unsigned long long *)a.ll_seq = *(volatile const unsigned long long *)peer.seq_dev + 1ULL;   // every CTA has read the sequence number and all of its words: close the exchange
      const double n2 = t;
      const float cnt = sv[5];
      if (a.head == 0) {
        const float entropy = 1.4189385332046727f + sv[6];   // policies.jl:348, logΣ as the minibatch kernel saw it
        const float p_loss = -(sv[0] / cnt);
        a.rec[CRUX_PPO_LOSS] = a.lambda_p * p_loss + a.lambda_e * (-entropy);
        a.rec[CRUX_PPO_ENTROPY] = entropy;
        const float kl = sv[1] / cnt;
        a.rec[CRUX_PPO_KL] = kl;
        a.rec[CRUX_PPO_CLIP_FRAC] = a.a2c ? 0.f : sv[2] / cnt;
        a.rec[CRUX_PPO_AVG_ADV] = sv[3] / cnt;
        a.rec[CRUX_PPO_AVG_RET] = sv[4] / cnt;
        if (a.ctl && kl > a.target_kl) a.ctl[1] = a.mb + 1;         // this minibatch is still applied; later ones are skipped
      } else {
        a.rec[CRUX_PPO_LOSS] = sv[0] / cnt;
      }
      a.rec[CRUX_PPO_GRAD_NORM] = (float)sqrt(n2);
      a.rec[CRUX_PPO_VALID] = 1.f;
      if (nan_seen || isnan(n2)) atomicOr(a.err_flags, CRUX_FLAG_NAN);   // training.jl:20: error before Flux.update!
      TAIL_TRACE(1);
    }
  }
#undef TAIL_TRACE
}

__global__ void fused_ctl_reset_kernel(int *ctl) { ctl[0] = 0; ctl[1] = 0; }

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
  static bool done = false;
  if (done) return CRUX_OK;
#define SET_ATTR(fn, bytes) CRUX_CHECK_CUDA(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)))
  SET_ATTR(fused_forward_kernel<R>, SmemMapT<R>::BYTES);
  SET_ATTR(fused_forward_kernel<R16>, SmemMapT<R>::BYTES);
  SET_ATTR(fused_forward_kernel<RB>, SmemMapT<RB>::BYTES);
  SET_ATTR((fused_minibatch_kernel<0, R>), SmemMapT<R>::BYTES);
  SET_ATTR((fused_minibatch_kernel<1, R>), SmemMapT<R>::BYTES);
  SET_ATTR(fused_forward_tc_kernel<0>, TcMapP::BYTES);
  SET_ATTR(fused_minibatch_tc_kernel<0>, TcMap::BYTES);
  SET_ATTR(fused_minibatch_tc_kernel<1>, TcMap::BYTES);
  SET_ATTR((fused_minibatch_kernel<0, RB>), SmemMapT<RB>::BYTES);
  SET_ATTR((fused_minibatch_kernel<1, RB>), SmemMapT<RB>::BYTES);
#undef SET_ATTR
  done = true;
  return CRUX_OK;
}

// forward launch: 16-row tiles when 64-row tiles cannot fill the GPU, 128-row tiles (1 CTA/SM) when they fill it at least once
int launch_forward(crux_ctx *ctx, FwdArgs &a, int nets) {
  const int64_t B = a.B;
  CruxTimed timed(ctx, CRUX_T_FORWARD);
  if (cdiv(B, R) * nets < (int64_t)ctx->num_sms || a.x_copy) {   // (the mapped-input form is implemented by the 16-row variant)
    dim3 grid((unsigned)i64min(cdiv(B, R16), (int64_t)ctx->num_sms * 2), nets);
    fused_forward_kernel<R16><<<grid, NT, SmemMapT<R>::BYTES, ctx->stream>>>(a);
  } else if (cdiv(B, RB) * nets >= (int64_t)ctx->num_sms && getenv("CRUX_RB")) {  // measured slower than 2 x 64-row CTAs/SM (profiles/): opt-in
    dim3 grid((unsigned)i64min(cdiv(B, RB), (int64_t)ctx->num_sms), nets);
    fused_forward_kernel<RB><<<grid, NT, SmemMapT<RB>::BYTES, ctx->stream>>>(a);
  } else {
    dim3 grid((unsigned)i64min(cdiv(B, R), (int64_t)ctx->num_sms * 2), nets);
    fused_forward_kernel<R><<<grid, NT, SmemMapT<R>::BYTES, ctx->stream>>>(a);
  }
  CRUX_LAUNCHED(ctx);
  return CRUX_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ entry points (internal)
// value(π, s) fast path for crux_mlp_forward
static int forward_fused_impl(crux_mlp *mlp, const float *x, int64_t B, float *y, const float *x_alt, const float *y_alt, int64_t alt_rows, int *handled) {
  *handled = 0;
  if (!fusable(mlp) || getenv("CRUX_NO_FUSED")) return CRUX_OK;
  crux_ctx *ctx = mlp->ctx;
  int rc = set_smem_attr(ctx); if (rc) return rc;
  // whole-column plain forwards: tcgen05 + TMEM kernel (fwd_tc5.cuh), one 128-row tile per SM and round
  // (CRUX_FWD_TC5=0: the mma.sync kernel below; =1s: the variant that keeps the activations in shared memory)
  const char *tc5_env = getenv("CRUX_FWD_TC5") ? getenv("CRUX_FWD_TC5") : "1";   // read per call: tests switch variants
  const bool alt_ok = x_alt && y_alt && mlp->dims[3] == 1 && ((uintptr_t)x_alt & 15) == 0 && !getenv("CRUX_NO_VALUE_REUSE");
  if (tc5_env[0] == '1' && (!alt_ok || tc5_env[1] != 's') && mlp->dims[0] <= tc5::KX && mlp->dims[3] <= 8 && cdiv(B, tc5::TR) >= (int64_t)ctx->num_sms &&
      ((uintptr_t)x & 15) == 0 && !getenv("CRUX_NO_MMA")) {
    static bool attr = false;
    if (!attr) {
      CRUX_CHECK_CUDA(ctx, cudaFuncSetAttribute(tc5::forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc5::Map::TOTAL));
      CRUX_CHECK_CUDA(ctx, cudaFuncSetAttribute(tc5::forward_kernel_tmem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc5::Map2::TOTAL));
      attr = true;
    }
    tc5::Args f;
    memset(&f, 0, sizeof(f));
    f.net = describe(mlp); f.x = x; f.B = B; f.y = y;
    if (alt_ok) { f.x_alt = x_alt; f.y_alt = y_alt; f.alt_rows = alt_rows; }
    {
      CruxTimed timed(ctx, CRUX_T_FORWARD);
      if (tc5_env[1] == 's')   // "1s": activations through shared memory (variant 1)
        tc5::forward_kernel<<<(unsigned)i64min(cdiv(B, tc5::TR), (int64_t)ctx->num_sms), NT, tc5::Map::TOTAL, ctx->stream>>>(f);
      else                     // activations stay in tensor memory, two CTAs per SM
        tc5::forward_kernel_tmem<<<(unsigned)i64min(cdiv(B, tc5::TR), (int64_t)ctx->num_sms * 2), NT, tc5::Map2::TOTAL, ctx->stream>>>(f);
    }
    CRUX_LAUNCHED(ctx);
    *handled = 1;
    return CRUX_OK;
  }
  // whole-column passes (at least one 64-row tile per resident CTA) run on the tensor cores; small batches keep the FFMA tiles
  if (cdiv(B, R) >= (int64_t)ctx->num_sms * 2 && ((uintptr_t)x & 15) == 0 && !getenv("CRUX_NO_MMA")) {
    FwdTcArgs f;
    memset(&f, 0, sizeof(f));
    f.net = describe(mlp); f.x = x; f.B = B; f.y = y;
    if (x_alt && y_alt && mlp->dims[3] == 1 && ((uintptr_t)x_alt & 15) == 0 && !getenv("CRUX_NO_VALUE_REUSE")) { f.x_alt = x_alt; f.y_alt = y_alt; f.alt_rows = alt_rows; }
    {
      CruxTimed timed(ctx, CRUX_T_FORWARD);
      fused_forward_tc_kernel<0><<<(unsigned)i64min(cdiv(B, R), (int64_t)ctx->num_sms * 2), NT, TcMapP::BYTES, ctx->stream>>>(f);
    }
    CRUX_LAUNCHED(ctx);
    *handled = 1;
    return CRUX_OK;
  }
  FwdArgs a;
  memset(&a, 0, sizeof(a));
  a.net[0] = describe(mlp); a.mode[0] = 0; a.x = x; a.B = B; a.y[0] = y;
  rc = launch_forward(ctx, a, 1);
  if (rc) return rc;
  *handled = 1;
  return CRUX_OK;
}
int mlp_forward_fused(crux_mlp *mlp, const float *x, int64_t B, float *y, int *handled) {
  return forward_fused_impl(mlp, x, B, y, nullptr, nullptr, 0, handled);
}
// value(V, sp) over a [T][N] rollout given V(s): see FwdTcArgs.  *handled == 0 -> the caller runs a plain forward.
int mlp_value_next_fused(crux_mlp *mlp, const float *sp, const float *s, const float *v_s, int64_t T, int64_t N, float *v_sp, int *handled) {
  const int I = mlp->dims[0];
  return forward_fused_impl(mlp, sp, T * N, v_sp, T > 1 ? s + N * I : nullptr, T > 1 ? v_s + N : nullptr, (T - 1) * N, handled);
}

bool fused_rows_supported(const crux_gaussian *actor) {
  return !getenv("CRUX_NO_FUSED") && !actor->categorical && fusable(actor->mu) && !actor->head_mode && !actor->squashed && actor->adim == actor->mu->dims[3];
}

extern "C" int32_t crux_rollout_step_fused(crux_gaussian *actor, crux_mlp *critic, const float *obs, int64_t N, const float *eps_in,
                                           uint64_t seed, uint64_t ctr, float *a_out, float *logp_out, float *v_out, int *handled,
                                           int64_t row0) {
  *handled = 0;
  if (getenv("CRUX_NO_FUSED")) return CRUX_OK;
  if (actor->categorical || !fusable(actor->mu) || actor->head_mode || actor->squashed || actor->adim != actor->mu->dims[3]) return CRUX_OK;
  const bool with_critic = critic && v_out;
  if (with_critic && (!fusable(critic) || critic->dims[3] != 1 || critic->dims[0] != actor->mu->dims[0])) return CRUX_OK;
  crux_ctx *ctx = actor->ctx;
  int rc = set_smem_attr(ctx); if (rc) return rc;
  FwdArgs a;
  memset(&a, 0, sizeof(a));
  a.net[0] = describe(actor->mu); a.mode[0] = 1; a.y[0] = a_out; a.logp = logp_out; a.ls = actor->log_sigma; a.eps_in = eps_in;
  a.seed = seed; a.ctr = ctr; a.x = obs; a.B = N; a.row0 = row0;
  if (with_critic) { a.net[1] = describe(critic); a.mode[1] = 0; a.y[1] = v_out; }
  rc = launch_forward(ctx, a, with_critic ? 2 : 1);
  if (rc) return rc;
  *handled = 1;
  return CRUX_OK;
}

// rows [row0, row0 + N) of a vector step whose observations sit in PINNED HOST memory: the kernel reads them over PCIe, stores them
// into the device column s_dev, and writes the actions both to the device column and to pinned host memory -- one launch, no copy
// calls (crux_rollout_host).  Same arithmetic and noise streams as crux_rollout_step_rows.
extern "C" int32_t crux_rollout_step_rows_mapped(crux_gaussian *actor, const float *obs_pinned, int64_t N, int64_t row0, uint64_t seed, uint64_t ctr,
                                                 float *s_dev, float *a_dev, float *a_pinned, float *logp_dev, const uint8_t *reset_flag_pinned,
                                                 const float *reset_obs_pinned) {
  if (!actor) return CRUX_ERR_INVALID;
  crux_ctx *ctx = actor->ctx;
  CRUX_REQUIRE(ctx, fused_rows_supported(actor), "crux_rollout_step_rows_mapped: only the fused policy shapes support split vector steps");
  CRUX_REQUIRE(ctx, obs_pinned && s_dev && a_dev && a_pinned && N >= 1, "crux_rollout_step_rows_mapped: bad arguments");
  int rc = set_smem_attr(ctx); if (rc) return rc;
  FwdArgs a;
  memset(&a, 0, sizeof(a));
  a.net[0] = describe(actor->mu); a.mode[0] = 1; a.y[0] = a_dev; a.y2 = a_pinned; a.logp = logp_dev; a.ls = actor->log_sigma;
  a.seed = seed; a.ctr = ctr; a.x = obs_pinned; a.x_copy = s_dev; a.B = N; a.row0 = row0;
  a.x_flag = reset_flag_pinned; a.x_reset = reset_obs_pinned;
  return launch_forward(ctx, a, 1);
}

// Programmatic dependent launch: the kernel may start (its prologue up to griddepcontrol.wait) while the previous kernel of the stream
// is still running, once that kernel has executed griddepcontrol.launch_dependents (or finished).  Used for the
// minibatch -> tail -> minibatch chain of one network, whose launch gaps (3 - 4 us each) are otherwise exposed.
template <class... KArgs, class... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
static bool pdl_enabled() { static const bool on = getenv("CRUX_NO_PDL") == nullptr; return on; }

// CRUX_MB_KERNEL selects the minibatch kernel: "t5" (default: all GEMMs on tcgen05, mb_t5.cuh), "mma" (warp-level mma.sync 3xTF32),
// "tc5" (= CRUX_MB_TC5=1: row GEMMs on tcgen05, weight gradients on mma.sync), "ffma" (= CRUX_NO_MMA=1)
static const char *mb_kernel_env() { const char *env = getenv("CRUX_MB_KERNEL"); return env ? env : ""; }   // read per call: tests switch kernels
static bool no_mma() { return getenv("CRUX_NO_MMA") != nullptr || strcmp(mb_kernel_env(), "ffma") == 0; }
static const char *mb5_selected() { return strcmp(mb_kernel_env(), "tc5") == 0 ? "1" : getenv("CRUX_MB_TC5"); }
static bool minibatch_kernel_is_t5(const crux_mlp *mlp) {
  const char *mb5_env = mb5_selected();
  if (no_mma() || (mb5_env && mb5_env[0] == '1')) return false;
  if (mb_kernel_env()[0] && strcmp(mb_kernel_env(), "t5") != 0) return false;
  return mlp->dims[0] <= mb6::KX && mlp->dims[3] <= 8;
}

// one minibatch = 3 launches: fused forward/loss/backward -> partial reduction (+ step count) -> [all-reduce] -> norm/record/Adam
static int fused_minibatch(crux_gaussian *actor, crux_mlp *mlp, int head, const float *s, const float *act, const float *logp_old,
                           const float *adv, const float *ret, const int32_t *order, int64_t bm, const crux_ppo_hp *hp, float *rec,
                           int *ctl, int mb, bool pdl_ok = false) {
  crux_ctx *ctx = mlp->ctx;
  // 128-row tiles (1 CTA/SM, 8x4 register tiles) measured SLOWER than 2 x 64-row CTAs per SM (occupancy halves; profiles/): opt-in
  const bool big = cdiv(bm, RB) >= (int64_t)ctx->num_sms && getenv("CRUX_RB");
  // CRUX_MB_RESERVE_SMS=k leaves k SMs out of this kernel's grid and runs the tail's reduce kernel on 2k CTAs (for the concurrent
  // reduce / all-reduce / Adam tail of the other network on several ranks).  Measured on 2 x B200: k = 0: 132 M env-steps/s,
  // k = 4: 81 M, k = 8: 109 M, k = 16: 127 M -- the reduction needs its full grid more than it needs to start early -- so 0 it is.
  static const int reserve_env = getenv("CRUX_MB_RESERVE_SMS") ? atoi(getenv("CRUX_MB_RESERVE_SMS")) : 0;
  const int reserve = reserve_env > 0 ? reserve_env : 0;
  const int sms = (int)i64max(1, ctx->num_sms - reserve);
  const char *mb5_env0 = mb5_selected();
  const bool tc5_grid = !big && !no_mma() && mb5_env0 && mb5_env0[0] == '1' && mlp->dims[0] <= mb5::KX && mlp->dims[3] <= 8;
  const bool t5k = minibatch_kernel_is_t5(mlp) && !big;
  const int grid = big ? (int)i64min(cdiv(bm, RB), (int64_t)ctx->num_sms)
                       : t5k ? (int)i64min(cdiv(cdiv(bm, mb6::NR), 2), (int64_t)(getenv("CRUX_MB_HALF") ? ctx->num_sms / 2 : ctx->num_sms))
                       : tc5_grid ? (int)i64min(cdiv(bm, mb5::TR), (int64_t)ctx->num_sms) : (int)i64min(cdiv(bm, R), (int64_t)sms * 2);
  const int pstride = (int)((mlp->n_params + 16 + 31) / 32 * 32);
  const size_t need = (size_t)ctx->num_sms * 2 * pstride * sizeof(float);
  int rc = ppo_ensure_bytes(ctx, (void **)&mlp->partials, &mlp->partials_bytes, need);
  if (rc) return rc;
  const float inv_bg = 1.0f / ((float)bm * (float)ctx->world);
  MbArgs a;
  memset(&a, 0, sizeof(a));
  a.net = describe(mlp); a.s = s; a.act = act; a.logp_old = logp_old; a.adv = adv; a.ret = ret; a.order = order; a.bm = bm;
  a.ls = head == 0 ? actor->log_sigma : nullptr;
  a.inv_bg = inv_bg; a.eps_clip = hp->eps_clip; a.lambda_p = hp->lambda_p; a.a2c = hp->a2c; a.partials = mlp->partials; a.pstride = pstride;
  a.n_params = (int)mlp->n_params; a.ctl = ctl; a.mb = mb;
  // CRUX_MB_TC5=1: row GEMMs on tcgen05 with the activations in tensor memory (mb_tc5.cuh), one 512-thread CTA per SM, 128-row tiles
  const char *mb5_env = mb5_selected();
  const bool tc5k = !big && !no_mma() && mb5_env && mb5_env[0] == '1' && mlp->dims[0] <= mb5::KX && mlp->dims[3] <= 8;
  const bool tc = !big && !tc5k && !t5k && !no_mma() && mlp->frag;   // mma.sync kernel: stages the fragment buffer instead of the raw parameters
  if (tc) { a.net.params = mlp->frag; a.net.bytes16 = (uint32_t)(Frag::TOTAL * sizeof(float)); }
  if (t5k) { a.planes = mlp->frag; a.nan_flag = mlp->step_dev + 2; }
  {
  CruxTimed timed(ctx, CRUX_T_MINIBATCH);
  if (big) {
    if (head == 0) fused_minibatch_kernel<0, RB><<<grid, NT, SmemMapT<RB>::BYTES, ctx->stream>>>(a);
    else fused_minibatch_kernel<1, RB><<<grid, NT, SmemMapT<RB>::BYTES, ctx->stream>>>(a);
  } else if (t5k) {   // default: every GEMM on tcgen05, features on the TMEM lanes (mb_t5.cuh), one 256-thread CTA per SM, 64-row tiles
    static bool attr6 = false;
    if (!attr6) {
      CRUX_CHECK_CUDA(ctx, cudaFuncSetAttribute((mb6::minibatch_kernel<0, CRUX_ACT_TANH>), cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      CRUX_CHECK_CUDA(ctx, cudaFuncSetAttribute((mb6::minibatch_kernel<1, CRUX_ACT_TANH>), cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      CRUX_CHECK_CUDA(ctx, cudaFuncSetAttribute((mb6::minibatch_kernel<0, CRUX_ACT_TANH>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mb6::Map::TOTAL));
      CRUX_CHECK_CUDA(ctx, cudaFuncSetAttribute((mb6::minibatch_kernel<1, CRUX_ACT_TANH>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mb6::Map::TOTAL));
      CRUX_CHECK_CUDA(ctx, cudaFuncSetAttribute((mb6::minibatch_kernel<0, CRUX_ACT_RELU>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mb6::Map::TOTAL));
      CRUX_CHECK_CUDA(ctx, cudaFuncSetAttribute((mb6::minibatch_kernel<1, CRUX_ACT_RELU>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mb6::Map::TOTAL));
      attr6 = true;
    }
    static long long *prof_dev = nullptr;
    if (getenv("CRUX_MB6_PROF")) {
      if (!prof_dev) CRUX_CHECK_CUDA(ctx, cudaMalloc(&prof_dev, 128 * sizeof(long long)));
      CRUX_CHECK_CUDA(ctx, cudaMemsetAsync(prof_dev, 0, 128 * sizeof(long long), ctx->stream));
      a.prof = prof_dev;
    }
    static unsigned long long *trace_dev = nullptr;
    static int trace_n = 0;
    if (getenv("CRUX_MB6_TRACE")) {
      if (!trace_dev) { CRUX_CHECK_CUDA(ctx, cudaMalloc(&trace_dev, 256 * 160 * 2 * sizeof(unsigned long long))); CRUX_CHECK_CUDA(ctx, cudaMemset(trace_dev, 0, 256 * 160 * 2 * sizeof(unsigned long long))); }
      if (trace_n < 256) a.trace = trace_dev + (size_t)(trace_n++) * 160 * 2;
      if (trace_n == 192) {   // dump once: per launch the first / last CTA start and first / last CTA end, in us from the first start
        std::vector<unsigned long long> h(192 * 160 * 2);
        CRUX_CHECK_CUDA(ctx, cudaDeviceSynchronize());
        CRUX_CHECK_CUDA(ctx, cudaMemcpy(h.data(), trace_dev, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        unsigned long long t0 = ~0ULL;
        for (int L = 128; L < 191; ++L) for (int c = 0; c < 148; ++c) if (h[(L * 160 + c) * 2]) t0 = t0 < h[(L * 160 + c) * 2] ? t0 : h[(L * 160 + c) * 2];
        for (int L = 128; L < 191; ++L) {
          unsigned long long s0 = ~0ULL, s1 = 0, e0 = ~0ULL, e1 = 0; int n = 0;
          for (int c = 0; c < 148; ++c) { const unsigned long long a0 = h[(L * 160 + c) * 2], a1 = h[(L * 160 + c) * 2 + 1]; if (!a0) continue; ++n;
            s0 = s0 < a0 ? s0 : a0; s1 = s1 > a0 ? s1 : a0; e0 = e0 < a1 ? e0 : a1; e1 = e1 > a1 ? e1 : a1; }
          fprintf(stderr, "trace launch %3d ctas %3d: start %8.2f .. %8.2f  end %8.2f .. %8.2f us   tail %8.2f .. %8.2f\n", L, n, (s0 - t0) * 1e-3, (s1 - t0) * 1e-3, (e0 - t0) * 1e-3, (e1 - t0) * 1e-3,
                  (h[(L * 160 + 150) * 2] - t0) * 1e-3, (h[(L * 160 + 150) * 2 + 1] - t0) * 1e-3);
          const unsigned long long *tt = &h[(L * 160 + 150) * 2];   // tail phases relative to the tail's start: prologue, reduce, Adam, ticket (CTA 0), last CTA's ticket, end
          fprintf(stderr, "      tail phases (us from wait-return): prologue %.2f  reduce %.2f  adam %.2f  ticket %.2f | last CTA at %.2f  end %.2f\n", (tt[2] - tt[0]) * 1e-3,
                  (tt[3] - tt[0]) * 1e-3, (tt[4] - tt[0]) * 1e-3, (tt[5] - tt[0]) * 1e-3, (tt[6] - tt[0]) * 1e-3, (tt[1] - tt[0]) * 1e-3);
          fprintf(stderr, "      slice landed %.2f\n", (tt[7] - tt[0]) * 1e-3);
        }
      }
    }
    const bool tanh_act = mlp->acts[0] == CRUX_ACT_TANH;   // the activation is a compile-time parameter: branch-free epilogues
    // Programmatic dependent launch of the MINIBATCH kernel behind its network's previous tail kernel is opt-in (CRUX_PDL_MB=1).  Measured
    // (scripts/mb6_trace.py, bench.py): with the tail's trigger at its START the minibatch kernel took over every SM as soon as one freed
    // up and idled at griddepcontrol.wait while the OTHER network's minibatch kernel waited behind it (1.40 ms per iteration against
    // 1.34); with the trigger after the tail's Adam stores (where it is now) the prologue does overlap the tail's record keeping, but
    // the step is no faster (1.163 against 1.156 ms): the SMs the early CTAs hold are the ones the other network's kernel would use.
    static const bool pdl_mb = getenv("CRUX_PDL_MB") != nullptr;
    const bool pdl = pdl_enabled() && pdl_ok && pdl_mb;   // never for a network's first minibatch: the row orders come from the kernels right before it
    void (*kern)(MbArgs) = head == 0 ? (tanh_act ? mb6::minibatch_kernel<0, CRUX_ACT_TANH> : mb6::minibatch_kernel<0, CRUX_ACT_RELU>)
                                     : (tanh_act ? mb6::minibatch_kernel<1, CRUX_ACT_TANH> : mb6::minibatch_kernel<1, CRUX_ACT_RELU>);
    CRUX_CHECK_CUDA(ctx, launch_pdl(kern, dim3(grid), dim3(mb6::NTH), (size_t)mb6::Map::TOTAL, ctx->stream, pdl, a));
    if (a.prof) {
      long long h[128];
      CRUX_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      CRUX_CHECK_CUDA(ctx, cudaMemcpy(h, prof_dev, sizeof(h), cudaMemcpyDeviceToHost));
      fprintf(stderr, "mb6 prof head=%d bm=%lld:", head, (long long)bm);
      for (int k = 1; k < 64 && h[k]; ++k) fprintf(stderr, " %lld", h[k] - h[k - 1]);
      fprintf(stderr, "\nmb6 abs E:");
      for (int k = 0; k < 64 && h[k]; ++k) fprintf(stderr, " %lld", h[k] - h[0]);
      fprintf(stderr, "\nmb6 abs I:");
      for (int k = 64; k < 128 && h[k]; ++k) fprintf(stderr, " %lld", h[k] - h[0]);
      fprintf(stderr, "\n");
    }
  } else if (tc5k) {
    static bool attr5 = false;
    if (!attr5) {
      CRUX_CHECK_CUDA(ctx, cudaFuncSetAttribute(mb5::minibatch_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mb5::Map::TOTAL));
      CRUX_CHECK_CUDA(ctx, cudaFuncSetAttribute(mb5::minibatch_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mb5::Map::TOTAL));
      attr5 = true;
    }
    if (head == 0) mb5::minibatch_kernel<0><<<grid, mb5::NTH, mb5::Map::TOTAL, ctx->stream>>>(a);
    else mb5::minibatch_kernel<1><<<grid, mb5::NTH, mb5::Map::TOTAL, ctx->stream>>>(a);
  } else if (!tc) {   // all-FFMA variant (CRUX_NO_MMA=1: A/B reference for the tensor-core path)
    if (head == 0) fused_minibatch_kernel<0, R><<<grid, NT, SmemMapT<R>::BYTES, ctx->stream>>>(a);
    else fused_minibatch_kernel<1, R><<<grid, NT, SmemMapT<R>::BYTES, ctx->stream>>>(a);
  } else {
    if (head == 0) fused_minibatch_tc_kernel<0><<<grid, NT, TcMap::BYTES, ctx->stream>>>(a);
    else fused_minibatch_tc_kernel<1><<<grid, NT, TcMap::BYTES, ctx->stream>>>(a);
  }
  }
  CRUX_LAUNCHED(ctx);
  const int nparts = grid;
  const int n_out = (int)mlp->n_params + 16;
  const int rblocks = (n_out + 31) / 32;   // <= 1024 doubles of norm_part (n_params <= 6792)
  const int rgrid = reserve > 0 ? (int)i64min(rblocks, 2 * reserve) : rblocks;   // two 1024-thread CTAs per reserved SM
  const float ls_shift = head == 0 ? -hp->lambda_e / (float)ctx->world : 0.f;
  AdamArgs g;
  memset(&g, 0, sizeof(g));
  g.p = mlp->params; g.g = mlp->grads; g.m = mlp->m; g.v = mlp->v; g.n = (int)mlp->n_params;
  if (head == 0) { g.ls = actor->log_sigma; g.ls_g = tail_ls_grad(mlp); g.ls_m = actor->ls_m; g.ls_v = actor->ls_v; g.A = actor->adim; }
  g.lambda_e = hp->lambda_e;
  g.beta_cache = mlp->norm_part + BETA_CACHE;
  g.trace = a.trace ? a.trace + 2 * 150 : nullptr;
  g.sums = tail_sums(mlp); g.eta = mlp->eta; g.b1 = mlp->beta1; g.b2 = mlp->beta2; g.eps = mlp->eps; g.step_dev = mlp->step_dev;
  g.lambda_p = hp->lambda_p; g.target_kl = hp->target_kl; g.a2c = hp->a2c; g.head = head; g.rec = rec; g.ctl = ctl; g.mb = mb;
  g.err_flags = ctx->flags_dev;
  if (tc || t5k) { g.frag = mlp->frag; g.fI = mlp->dims[0]; g.fO = mlp->dims[3]; g.frag_mode = t5k ? 1 : 0; }
  // Running the Adam tail in the last CTA of the reduce kernel saves a launch but serialises 5.7k double-precision updates on
  // one SM: measured 20.5 us against 6.5 + 6.9 us for the two separate kernels (profiles/r1_notes.md) -> opt-in only.
  const int fuse_adam = (ctx->world == 1 && getenv("CRUX_FUSE_ADAM")) ? 1 : 0;
  if (ctx->world == 1) { g.norm_part = mlp->norm_part; g.n_norm_part = rblocks; }
  // multi-GPU: when the peer buffers are mapped (crux_peer_init) the gradient all-reduce is fused into these two kernels with the LL
  // (flag-in-data) protocol: the reduce kernel stores every entry as one 8-byte {value, sequence} word straight into every rank's
  // receive region over NVLink, the Adam kernel spins on the words as it sums them -- no fences, no flag round trip, no NCCL launch.
  // Each network (head) has its own region and sequence number, so the actor and critic exchanges can be in flight concurrently.
  // (The first, fence-based version of this fusion measured 22 + 28 us per minibatch against 10 + 17 (NCCL) + 9 us.)
  PeerOut po;
  memset(&po, 0, sizeof(po));
  const bool use_peer = ctx->world > 1 && ctx->peer_ready && ctx->peer_ll && mlp->n_params + CRUX_GRAD_TAIL <= ctx->peer_cap && !getenv("CRUX_NO_PEER_LL");
  if (use_peer) {
    const int64_t net_off = (int64_t)head * 32 * ctx->peer_cap;   // [network][parity][rank][cap]
    po.enabled = 1; po.world = ctx->world; po.rank = ctx->rank; po.cap = ctx->peer_cap; po.seq_dev = ctx->peer_flags + 40 + head;
    for (int q = 0; q < ctx->world; ++q) po.ll[q] = ctx->peer_ll_remote[q] + net_off;
    g.peer = 1; g.world = ctx->world; g.peer_cap = ctx->peer_cap; g.ll_recv = ctx->peer_ll + net_off;
    g.ll_seq = ctx->peer_flags + 40 + head; g.ll_ticket = reinterpret_cast<unsigned int *>(ctx->peer_flags + 48 + head);
  }
  static const bool no_fused_tail = getenv("CRUX_NO_FUSED_TAIL") != nullptr;
  if (t5k && (ctx->world == 1 || use_peer) && !no_fused_tail) {   // reduce [+ LL gradient exchange over NVLink] + Adam + record in ONE launch
    static bool carve = false;
    if (!carve) {   // same shared-memory carve-out as the minibatch kernel: CTAs of kernels with different carve-outs do not share an SM
      CRUX_CHECK_CUDA(ctx, cudaFuncSetAttribute(reduce_adam_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      carve = true;
    }
    CruxTimed timed(ctx, CRUX_T_REDUCE);
    CRUX_CHECK_CUDA(ctx, launch_pdl(reduce_adam_kernel<4>, dim3(rgrid), dim3(4 * 32), 0, ctx->stream, pdl_enabled(), (const float *)mlp->partials, nparts, pstride,
                                    (int)mlp->n_params, mlp->grads, (float)bm, ls_shift, head == 0 ? actor->adim : 0, mlp->norm_part, mlp->step_dev, g, po));
    CRUX_LAUNCHED(ctx);
    return CRUX_OK;
  }
  { CruxTimed timed(ctx, CRUX_T_REDUCE);
  if (fuse_adam)
    reduce_fused_partials_kernel<1><<<rgrid, RW * 32, 0, ctx->stream>>>(mlp->partials, nparts, pstride, (int)mlp->n_params, mlp->grads, (float)bm,
                                                                        ls_shift, head == 0 ? actor->adim : 0, mlp->norm_part, mlp->step_dev, ctl, mb,
                                                                        ctx->flags_dev + 2, g, fuse_adam, po);
  else
    reduce_fused_partials_kernel<0><<<rgrid, RW * 32, 0, ctx->stream>>>(mlp->partials, nparts, pstride, (int)mlp->n_params, mlp->grads, (float)bm,
                                                                        ls_shift, head == 0 ? actor->adim : 0, mlp->norm_part, mlp->step_dev, ctl, mb,
                                                                        ctx->flags_dev + 2, g, fuse_adam, po);
  }
  CRUX_LAUNCHED(ctx);
  if (!fuse_adam) {
    if (!use_peer) { rc = grads_allreduce(ctx, mlp->grads, mlp->n_params + CRUX_GRAD_TAIL); if (rc) return rc; }
    { CruxTimed timed(ctx, CRUX_T_ADAM);
    static const bool ll_old = getenv("CRUX_LL_ADAM_V1") != nullptr;   // A/B: every CTA polls the whole vector (adam_body)
    if (use_peer && !ll_old)
      fused_adam_ll_kernel<<<(g.n + 255) / 256, 256, 0, ctx->stream>>>(g, mlp->norm_part, reinterpret_cast<unsigned int *>(ctx->peer_flags + 48 + head),
                                                                       ctx->peer_flags + 56 + head);
    else
      fused_adam_kernel<<<(g.n + 255) / 256, 256, 0, ctx->stream>>>(g); }
    CRUX_LAUNCHED(ctx);
  }
  return CRUX_OK;
}

// batch_train! of one network as one cluster launch (mb_persist.cuh).  Row orders of ALL epochs are needed up front: the caller's, or
// device-generated permutations written behind each other.
static int persistent_epochs(crux_gaussian *actor, crux_mlp *mlp, int head, const float *s, const float *act, const float *logp_old, const float *adv,
                             const float *ret, int64_t n, const crux_ppo_hp *hp, const int32_t *order_in, uint64_t seed) {
  crux_ctx *ctx = mlp->ctx;
  const int epochs = head == 0 ? hp->actor_epochs : hp->critic_epochs, batch = head == 0 ? hp->actor_batch : hp->critic_batch;
  const int32_t *order = order_in;
  if (!order) {
    int32_t **buf = head == 0 ? &actor->order : &actor->order2;
    size_t *have = head == 0 ? &actor->order_bytes : &actor->order2_bytes;
    int rc = ppo_ensure_bytes(ctx, (void **)buf, have, (size_t)epochs * n * sizeof(int32_t)); if (rc) return rc;
    rc = ppo_fill_orders(ctx, *buf, n, seed, 0u, epochs); if (rc) return rc;
    order = *buf;
  }
  mbp::Args a;
  memset(&a, 0, sizeof(a));
  a.net = describe(mlp); a.s = s; a.act = act; a.logp_old = logp_old; a.adv = adv; a.ret = ret; a.order = order; a.n = n; a.batch = batch;
  a.epochs = epochs; a.max_batches = head == 0 ? hp->actor_max_batches : hp->critic_max_batches;
  if (head == 0) { a.ls = actor->log_sigma; a.ls_m = actor->ls_m; a.ls_v = actor->ls_v; a.ctl = actor->ctl; }
  a.m = mlp->m; a.v = mlp->v; a.step_dev = mlp->step_dev; a.beta_cache = mlp->norm_part + BETA_CACHE;
  a.eta = mlp->eta; a.b1 = mlp->beta1; a.b2 = mlp->beta2; a.eps = mlp->eps;
  a.eps_clip = hp->eps_clip; a.lambda_p = hp->lambda_p; a.lambda_e = hp->lambda_e; a.target_kl = hp->target_kl; a.a2c = hp->a2c;
  a.info = head == 0 ? actor->info_actor : actor->info_critic; a.err_flags = ctx->flags_dev; a.n_params = (int)mlp->n_params;
  static bool attr = false;
  if (!attr) {
    CRUX_CHECK_CUDA(ctx, cudaFuncSetAttribute(mbp::epoch_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mbp::Map::BYTES));
    CRUX_CHECK_CUDA(ctx, cudaFuncSetAttribute(mbp::epoch_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mbp::Map::BYTES));
    attr = true;
  }
  if (head == 0) mbp::epoch_kernel<0><<<mbp::C, NT, mbp::Map::BYTES, ctx->stream>>>(a);
  else mbp::epoch_kernel<1><<<mbp::C, NT, mbp::Map::BYTES, ctx->stream>>>(a);
  CRUX_LAUNCHED(ctx);
  return CRUX_OK;
}

// policy_gradient_training (on_policy.jl:56-78) for fusable shapes: all actor epochs, then all critic epochs
int ppo_update_fused(crux_gaussian *actor, crux_mlp *critic, const float *s, const float *a, const float *logprob, const float *advantage,
                     const float *ret, int64_t n, const crux_ppo_hp *hp, const int32_t *order_actor, const int32_t *order_critic,
                     uint64_t seed, int *handled) {
  *handled = 0;
  crux_mlp *mu = actor->mu;
  if (getenv("CRUX_NO_FUSED") || actor->categorical || !fusable(mu) || actor->head_mode || actor->squashed || actor->adim != mu->dims[3]) return CRUX_OK;
  if (critic && (!fusable(critic) || critic->dims[3] != 1)) return CRUX_OK;
  crux_ctx *ctx = actor->ctx;
  int rc = set_smem_attr(ctx); if (rc) return rc;
  const int64_t nmb_a = cdiv(n, hp->actor_batch);
  const int64_t nmb_c = critic ? cdiv(n, hp->critic_batch) : 0;
  const size_t ia = (size_t)i64max(1, hp->actor_epochs * nmb_a) * CRUX_PPO_INFO_STRIDE * sizeof(float);
  const size_t ic = (size_t)i64max(1, (int64_t)hp->critic_epochs * nmb_c) * CRUX_PPO_INFO_STRIDE * sizeof(float);
  rc = ppo_ensure_bytes(ctx, (void **)&actor->info_actor, &actor->info_actor_bytes, ia); if (rc) return rc;
  rc = ppo_ensure_bytes(ctx, (void **)&actor->info_critic, &actor->info_critic_bytes, ic); if (rc) return rc;
  CRUX_CHECK_CUDA(ctx, cudaMemsetAsync(actor->info_actor, 0, ia, ctx->stream));
  CRUX_CHECK_CUDA(ctx, cudaMemsetAsync(actor->info_critic, 0, ic, ctx->stream));
  // device-generated row orders: ALL epochs of both networks in two launches at the start of the update (the per-epoch launch used to sit
  // on each network's dependency chain in front of the epoch's first minibatch)
  if (!order_actor || (nmb_c && !order_critic)) {
    rc = ppo_ensure_bytes(ctx, (void **)&actor->order, &actor->order_bytes, (size_t)i64max(1, hp->actor_epochs) * n * sizeof(int32_t)); if (rc) return rc;
    rc = ppo_ensure_bytes(ctx, (void **)&actor->order2, &actor->order2_bytes, (size_t)i64max(1, hp->critic_epochs) * n * sizeof(int32_t)); if (rc) return rc;
  }
  fused_ctl_reset_kernel<<<1, 1, 0, ctx->stream>>>(actor->ctl);
  CRUX_LAUNCHED(ctx);
  if (!no_mma()) {   // weights in the order the minibatch kernel stages them with one TMA bulk copy (kept current by the fused Adam kernels):
                     // tcgen05 hi/lo planes (mb_t5.cuh) or MMA B-fragment order (mma.sync kernel)
    crux_mlp *nets[2] = {mu, critic};
    for (crux_mlp *m : nets) {
      if (!m) continue;
      const bool t5 = minibatch_kernel_is_t5(m);
      const size_t bytes = t5 ? (size_t)mb6::Map::PLANES : Frag::TOTAL * sizeof(float);
      rc = ppo_ensure_bytes(ctx, (void **)&m->frag, &m->frag_bytes, bytes > Frag::TOTAL * sizeof(float) ? bytes : Frag::TOTAL * sizeof(float)); if (rc) return rc;
      if (t5) {
        CRUX_CHECK_CUDA(ctx, cudaMemsetAsync(m->frag, 0, mb6::Map::PLANES, ctx->stream));
        mb6::build_planes_kernel<<<((int)m->n_params + 255) / 256, 256, 0, ctx->stream>>>(m->params, m->frag, m->dims[0], m->dims[3], (int)m->n_params);
      } else {
        build_frag_kernel<<<(Frag::TOTAL + 255) / 256, 256, 0, ctx->stream>>>(m->params, m->frag, m->dims[0], m->dims[3]);
      }
      CRUX_LAUNCHED(ctx);
    }
  }
  if (ctx->side_stream) {   // fork point for the concurrent critic epochs: the side stream sees the rollout / whitening / memsets above
    CRUX_CHECK_CUDA(ctx, cudaEventRecord(ctx->side_fork, ctx->stream));
    CRUX_CHECK_CUDA(ctx, cudaStreamWaitEvent(ctx->side_stream, ctx->side_fork, 0));
  }
  // Small minibatches (the reference-default batch_size = 128): the whole batch_train! of a network is ONE launch of a thread-block
  // cluster that keeps parameters and Adam state on chip (mb_persist.cuh) instead of 4 launches per 27 us train! step.
  const bool no_persist = getenv("CRUX_NO_PERSIST") != nullptr;   // read per call: tests compare both paths
  const bool persist_a = !no_persist && ctx->world == 1 && !ctx->timing && hp->actor_batch <= mbp::MAXB && hp->actor_epochs * nmb_a >= 8;
  const bool persist_c = !no_persist && ctx->world == 1 && !ctx->timing && critic && hp->critic_batch <= mbp::MAXB && hp->critic_epochs * nmb_c >= 8;
  if (!order_actor && !persist_a) { rc = ppo_fill_orders(ctx, actor->order, n, seed, 0u, hp->actor_epochs); if (rc) return rc; }
  if (nmb_c && !order_critic && !persist_c) { rc = ppo_fill_orders(ctx, actor->order2, n, seed ^ 0xC2B2AE3D27D4EB4FULL, 0u, hp->critic_epochs); if (rc) return rc; }
  if (persist_a) { rc = persistent_epochs(actor, mu, 0, s, a, logprob, advantage, ret, n, hp, order_actor, seed); if (rc) return rc; }
  int64_t total = 0;
  const int64_t maxb_a = hp->actor_max_batches > 0 ? hp->actor_max_batches : INT64_MAX;
  for (int e = 0; e < hp->actor_epochs && total < maxb_a && !persist_a; ++e) {
    const int32_t *order;
    order = (order_actor ? order_actor : actor->order) + (int64_t)e * n;
    for (int64_t mbi = 0; mbi < nmb_a && total < maxb_a; ++mbi, ++total) {
      const int64_t off = mbi * hp->actor_batch, bm = i64min(hp->actor_batch, n - off);
      float *rec = actor->info_actor + ((int64_t)e * nmb_a + mbi) * CRUX_PPO_INFO_STRIDE;
      rc = fused_minibatch(actor, mu, 0, s, a, logprob, advantage, ret, order + off, bm, hp, rec, actor->ctl, (int)total, total > 0);
      if (rc) return rc;
    }
  }
  // The critic epochs depend only on the buffer (not on the actor), so they are enqueued on a side stream and run CONCURRENTLY
  // with the actor epochs: the small reduce / all-reduce / Adam launches of one network hide behind the minibatch kernel of the
  // other.  Results are exactly those of the sequential order (separate parameters, optimisers, scratch buffers and flags).
  cudaStream_t main_stream = ctx->stream;
  // With several ranks the critic's gradient exchange uses its own LL region (or, without peer buffers, the second NCCL
  // communicator of nccl.cu).
  // per-network LL exchange regions: safe to run both networks at once -- but only if BOTH gradients fit the mapped regions (a network that
  // does not fit falls back to NCCL, and two streams must never share one communicator: ADVICE r1)
  const bool ll = ctx->peer_ready && ctx->peer_ll && !getenv("CRUX_NO_PEER_LL") && mu->n_params + CRUX_GRAD_TAIL <= ctx->peer_cap &&
                  (!critic || critic->n_params + CRUX_GRAD_TAIL <= ctx->peer_cap);
  const bool side = (ctx->world == 1 || ll || (ctx->nccl_comm_side && !ctx->peer_ready)) && nmb_c > 0 && hp->critic_epochs > 0 && !ctx->timing &&
                    !getenv("CRUX_NO_SIDE_STREAM") && ctx->side_stream;
  if (side) ctx->stream = ctx->side_stream;   // every launch helper below enqueues on ctx->stream
  if (persist_c) {
    rc = persistent_epochs(actor, critic, 1, s, nullptr, nullptr, nullptr, ret, n, hp, order_critic, seed ^ 0xC2B2AE3D27D4EB4FULL);
    if (rc) { ctx->stream = main_stream; return rc; }
  }
  total = 0;
  const int64_t maxb_c = hp->critic_max_batches > 0 ? hp->critic_max_batches : INT64_MAX;
  for (int e = 0; e < hp->critic_epochs && nmb_c && total < maxb_c && !persist_c; ++e) {
    const int32_t *order;
    order = (order_critic ? order_critic : actor->order2) + (int64_t)e * n;
    for (int64_t mbi = 0; mbi < nmb_c && total < maxb_c; ++mbi, ++total) {
      const int64_t off = mbi * hp->critic_batch, bm = i64min(hp->critic_batch, n - off);
      float *rec = actor->info_critic + ((int64_t)e * nmb_c + mbi) * CRUX_PPO_INFO_STRIDE;
      rc = fused_minibatch(actor, critic, 1, s, nullptr, nullptr, nullptr, ret, order + off, bm, hp, rec, nullptr, (int)total, total > 0);
      if (rc) { ctx->stream = main_stream; return rc; }
    }
  }
  if (side) {   // join: everything after the update (info readback, the next rollout) sees both networks updated
    ctx->stream = main_stream;
    CRUX_CHECK_CUDA(ctx, cudaEventRecord(ctx->side_done, ctx->side_stream));
    CRUX_CHECK_CUDA(ctx, cudaStreamWaitEvent(main_stream, ctx->side_done, 0));
  }
  *handled = 1;
  return CRUX_OK;
}

// =================================================================================================== persistent device rollout
// The whole `steps!` loop (sampler.jl:139-155) for a DEVICE environment in ONE launch: every CTA owns 16 env streams and
// walks all T vector steps -- policy forward (fused_forward<16> building blocks, parameters staged once by TMA), Gaussian
// sample + logprob, the LinQuad transition, episode bookkeeping and reset -- with the observation tile resident in shared
// memory between steps.  Bit-identical to T x (crux_rollout_step + crux_linquad_step): same noise counters, same FMA order.
#include "env.cuh"

namespace {

struct RolloutArgs {
  NetDesc net;
  const float *ls;
  const float *A, *B;           // env matrices [sdim][sdim], [sdim][adim]
  int64_t N; int T, adim, max_steps, force_end;
  int rows;                     // env streams per CTA (<= 16): chosen so that the CTAs spread evenly over 2 x SM-count slots
  uint64_t seed_pi, ctr0;       // exploration noise: (seed_pi, ctr0 + t, stream)
  uint64_t seed_env;            // env noise: (seed_env, tick0 + t, stream*16 + group)
  unsigned long long *tick_dev; // [0] tick, [1] finished-block counter
  int32_t *ep_len;
  float *obs_io;                // [N][sdim] current observation of every stream (in: step 0, out: after step T-1)
  float *s, *a, *sp, *r, *logp; uint8_t *done, *ee;   // rollout columns, rows [T*N]
};

__global__ void __launch_bounds__(NT, 2) rollout_linquad_kernel(RolloutArgs g) {
  using M = SmemMapT<R>;
  extern __shared__ __align__(16) float sm[];
  // the env state reuses regions of the carve-up that the 16-row forward does not touch
  float *sA = sm + M::W2T;                       // [32*32] env A
  float *sB = sm + M::W2T + LQ_MAX_S * LQ_MAX_S; // [32*16] env B
  float *aT = sm + M::AT;                        // [8][LD16] sampled actions
  float *taT = sm + M::AT + MAX_O * LD16;        // [8][LD16] tanh(a)
  const NetDesc nd = g.net;
  const int I = nd.I, O = nd.O, sdim = I, adim = g.adim;
  const int t = threadIdx.x;
  stage_params(sm, nd, M::MBAR);
  float *spT = sm + M::H2T + H * LD16;           // [32][LD16] s': after the 16-row H2T tile ([64][LD16] = 1280 of the 4352 floats reserved)
  int *s_len = reinterpret_cast<int *>(spT + LQ_MAX_S * LD16);  // [16] episode lengths
  int *s_end = s_len + 16;                                      // [16] end flags of the current step
  for (int i = t; i < sdim * sdim; i += NT) sA[i] = g.A[i];
  for (int i = t; i < sdim * adim; i += NT) sB[i] = g.B[i];
  const float *P = sm + M::P;
  float *XT = sm + M::XT, *H1T = sm + M::H1T, *H2T = sm + M::H2T, *OT = sm + M::OT;
  const int rows = g.rows;
  const int64_t e0 = (int64_t)blockIdx.x * rows;
  const int r = t >> 4, d = t & 15;               // env row, dimension slot (dims d and d + 16)
  const int64_t e = e0 + r;
  const bool live = r < rows && e < g.N;
  const unsigned long long tick0 = *(volatile unsigned long long *)g.tick_dev;
  // initial observation tile + episode lengths
  for (int q = t; q < R16 * sdim; q += NT) {
    const int rr = q / sdim, i = q - rr * sdim;
    XT[i * LD16 + rr] = (rr < rows && e0 + rr < g.N) ? g.obs_io[(e0 + rr) * sdim + i] : 0.f;
  }
  if (t < R16) s_len[t] = (t < rows && e0 + t < g.N) ? g.ep_len[e0 + t] : 0;
  __syncthreads();

  // Every phase of a vector step uses the same mapping -- lane (r, d) = (t >> 4, t & 15): env stream r of the tile, slot d -- and only
  // touches columns [*][r] of the shared-memory tiles, so a stream's whole step (forward, head, transition, bookkeeping) lives in
  // ONE half-warp: the phases are ordered by __syncwarp() and the T-step loop needs no block barrier at all.  (It had eight per step,
  // 21 % of the stall samples; removing them moved the launch from 255 to 253.5 us only: the warps wait on shared-memory loads of
  // the two 64-wide layers and on the dependent FMA chains instead -- profiles/r1_notes.md.)  Arithmetic per value is unchanged.
  const int n_steps = (t >> 5) * 2 < rows ? g.T : 0;   // a warp whose two streams are both beyond the tile's rows has nothing to do
  for (int step = 0; step < n_steps; ++step) {
    const int64_t row0 = (int64_t)step * g.N + e0;
    // s row of this step
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int kk = d + 16 * h;
      if (kk < sdim && live) g.s[(row0 + r) * sdim + kk] = XT[kk * LD16 + r];
    }
    // ---- policy forward + Gaussian head (identical to fused_forward_kernel<16>): thread (r, jg) computes 4 outputs of stream r
    layer_fwd16(XT, I, P, P + off_b1(I), H1T, nd.act);
    __syncwarp();
    layer_fwd16(H1T, H, P + off_W2(I), P + off_b2(I), H2T, nd.act);
    __syncwarp();
    if (d < O) {   // output layer: lane (r, o)
      const float *W3 = P + off_W3(I);
      float a0 = P[off_b3(I, O) + d];
#pragma unroll 8
      for (int k = 0; k < H; ++k) a0 = fmaf(H2T[k * LD16 + r], W3[k * O + d], a0);
      OT[d * LD16 + r] = a0;
    }
    __syncwarp();
    // Gaussian head, lane (stream, action-dimension pair): one Philox block + one Box-Muller yield both normals of the pair; same noise
    // counters and the same arithmetic per value as the per-step kernel; the logpdf terms replace mu in OT and are summed in order below
    {
      const int j0 = 2 * d;
      const int64_t i = e0 + r;   // absolute stream id == row index of the vector step
      if (j0 < O) {
        const Philox4 p = philox4x32_10(g.seed_pi, g.ctr0 + (uint64_t)step, (uint64_t)i * ((O + 3) / 4) + (j0 >> 2));
        float n0, n1;
        if ((j0 & 2) == 0) box_muller(p.x, p.y, n0, n1); else box_muller(p.z, p.w, n0, n1);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int j = j0 + u;
          if (j < O) {
            const float mu = OT[j * LD16 + r];
            const float ls = g.ls[j];
            const float sigma = expf(ls);
            const float var = sigma * sigma;
            const float ev = u ? n1 : n0;
            const float act = ev * sigma + mu;
            aT[j * LD16 + r] = act;
            if (live) g.a[(row0 + r) * O + j] = act;
            const float dd = act - mu;
            OT[j * LD16 + r] = -(dd * dd) / (2.f * var) - LOG_SQRT_2PI - ls;
          }
        }
      }
    }
    __syncwarp();
    // ---- env transition (identical arithmetic to linquad_step_kernel)
    if (d < adim) taT[d * LD16 + r] = tanhf(aT[d * LD16 + r]);
    if (d == 15 && g.logp) {
      float logp = 0.f;
      for (int j = 0; j < O; ++j) logp += OT[j * LD16 + r];
      if (live) g.logp[row0 + r] = logp;
    }
    __syncwarp();
    const unsigned long long tick = tick0 + (unsigned long long)step;
    {
      // lane (r, d) owns the dimension PAIR (2d, 2d + 1): both normals come from one Box-Muller of one Philox block, so a stream costs
      // ceil(sdim / 2) Philox + Box-Muller evaluations in ONE pass instead of sdim in two (same counters, same arithmetic per value)
      const int k0 = 2 * d;
      if (k0 < sdim) {
        const Philox4 p = philox4x32_10(g.seed_env, tick, (uint64_t)(live ? e : 0) * 16 + (k0 >> 2));
        float x0, x1;
        if ((k0 & 2) == 0) box_muller(p.x, p.y, x0, x1); else box_muller(p.z, p.w, x0, x1);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int kk = k0 + u;
          if (kk < sdim) {
            float v = 0.f;
            const float *ar = sA + kk * sdim, *br = sB + kk * adim;
#pragma unroll 4
            for (int j = 0; j < sdim; ++j) v = fmaf(ar[j], XT[j * LD16 + r], v);
#pragma unroll 2
            for (int j = 0; j < adim; ++j) v = fmaf(br[j], taT[j * LD16 + r], v);
            v = fmaf(0.01f, u ? x1 : x0, v);
            v = fminf(fmaxf(v, -10.f), 10.f);
            spT[kk * LD16 + r] = v;
          }
        }
      }
    }
    // |s'|^2 in the exact order of linquad_step_kernel: 8 sequential 4-dim partials (one lane each), combined below like the
    // xor-1/2/4 butterfly.  tanh(a)^T is dead: it holds the partials.
    __syncwarp();
    if (d < 8) {
      float n2 = 0.f;
      for (int i = 0; i < 4 && 4 * d + i < sdim; ++i) { const float v = spT[(4 * d + i) * LD16 + r]; n2 = fmaf(v, v, n2); }
      taT[d * LD16 + r] = n2;
    }
    __syncwarp();
    if (d == 0) {   // reward, termination and episode bookkeeping of stream r
      float pq[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) pq[q] = taT[q * LD16 + r];
      const float n2 = ((pq[0] + pq[1]) + (pq[2] + pq[3])) + ((pq[4] + pq[5]) + (pq[6] + pq[7]));
      float a2 = 0.f;
      for (int j = 0; j < adim; ++j) { const float av = aT[j * LD16 + r]; a2 = fmaf(av, av, a2); }
      const float rew = 1.f - n2 / (float)sdim - 0.1f * a2 / (float)adim;
      const bool dn = fabsf(spT[r]) > 5.f;
      const int len = s_len[r] + 1;
      const bool end = dn || len >= g.max_steps || (g.force_end && step == g.T - 1);
      s_len[r] = end ? 0 : len;
      s_end[r] = end ? 1 : 0;
      if (live) { g.r[row0 + r] = rew; g.done[row0 + r] = dn ? 1 : 0; g.ee[row0 + r] = end ? 1 : 0; }
    }
    __syncwarp();
    // ---- s' rows out; next observation tile: s', or a fresh initial state where the episode ended
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int kk = d + 16 * h;
      if (kk < sdim) {
        const float v = spT[kk * LD16 + r];
        if (live) g.sp[(row0 + r) * sdim + kk] = v;
        float nx = v;
        if (s_end[r]) {
          const Philox4 p = philox4x32_10(g.seed_env ^ 0x5851F42D4C957F2DULL, tick + 0x100000000ULL, (uint64_t)(live ? e : 0) * 16 + (kk >> 2));
          const uint32_t u[4] = {p.x, p.y, p.z, p.w};
          nx = (u32_to_unit_open(u[kk & 3]) * 2.f - 1.f) * 0.1f;
        }
        XT[kk * LD16 + r] = nx;
      }
    }
    __syncwarp();
  }
  __syncthreads();
  // current observation + episode lengths back to global
  for (int q = t; q < R16 * sdim; q += NT) {
    const int rr = q / sdim, i = q - rr * sdim;
    if (rr < rows && e0 + rr < g.N) g.obs_io[(e0 + rr) * sdim + i] = XT[i * LD16 + rr];
  }
  if (t < rows && e0 + t < g.N) g.ep_len[e0 + t] = s_len[t];
  // the last block to finish advances the env tick by T (every block has read it by then)
  __shared__ bool last;
  __threadfence();
  __syncthreads();
  if (t == 0) last = atomicAdd(g.tick_dev + 1, 1ULL) == (unsigned long long)gridDim.x - 1ULL;
  __syncthreads();
  if (last && t == 0) { g.tick_dev[0] += (unsigned long long)g.T; g.tick_dev[1] = 0ULL; __threadfence(); }
}


// ---- register-tiled variants: a WARP owns WS = 2 RPT env streams, a thread an RPT-stream x 4-feature tile of the two hidden layers ----
// The kernel above gives every stream a half-warp, thread = (1 stream, 4 features): per k one scalar + one 128-bit shared load feed
// 4 FFMA, and ncu shows what that costs -- 82 % of the issued instructions are not FFMAs, 50.6 M shared-memory wavefronts, short
// scoreboard 1.8 per issue (profiles/r2_ncu_summary.md).  Here lane (rg, jg) = (lane >> 4, lane & 15) owns streams RPT rg .. + RPT - 1
// of the warp's WS and features 4 jg .. 4 jg + 3: one RPT-wide and one 128-bit load feed 4 RPT FFMA.  The phases outside the two
// 64-wide layers use lane (rs, slot) = (lane % WS, lane / WS): stream rs, task slot 0 .. 32 / WS - 1.  A stream's whole step still
// lives in ONE warp (phases ordered by __syncwarp(), no block barrier in the T-step loop), every value is computed by the same
// instruction sequence as before (accumulation order over k, noise counters, libdevice calls), so the rollout stays bit-identical to
// T x (crux_rollout_step + crux_linquad_step).  CTA = 16 / WS warps = up to 16 streams (`rows`), two CTAs per SM.
template <int RPT> struct RowVec;
template <> struct RowVec<4> { using T = float4; };
template <> struct RowVec<2> { using T = float2; };
template <int RPT> __device__ __forceinline__ void ld_rows(const float *p, float (&v)[RPT]) {
  const typename RowVec<RPT>::T x = *reinterpret_cast<const typename RowVec<RPT>::T *>(p);
  const float *f = reinterpret_cast<const float *>(&x);
#pragma unroll
  for (int i = 0; i < RPT; ++i) v[i] = f[i];
}
template <int RPT> __device__ __forceinline__ void st_rows(float *p, const float (&v)[RPT]) {
  typename RowVec<RPT>::T x;
  float *f = reinterpret_cast<float *>(&x);
#pragma unroll
  for (int i = 0; i < RPT; ++i) f[i] = v[i];
  *reinterpret_cast<typename RowVec<RPT>::T *>(p) = x;
}
template <int RPT>
__device__ __forceinline__ void layer_fwd_rt(const float *__restrict__ AT, int K, const float *__restrict__ W, const float *__restrict__ b,
                                             float *__restrict__ CT, int act, int r0, int jg) {
  float acc[RPT][4];
#pragma unroll
  for (int i = 0; i < RPT; ++i)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;
  const float *ap = AT + r0, *wp = W + 4 * jg;
#pragma unroll 8
  for (int k = 0; k < K; ++k) {
    float a4[RPT];
    ld_rows<RPT>(ap + k * LD16, a4);
    const float4 w = *reinterpret_cast<const float4 *>(wp + k * H);
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      acc[i][0] = fmaf(a4[i], w.x, acc[i][0]); acc[i][1] = fmaf(a4[i], w.y, acc[i][1]);
      acc[i][2] = fmaf(a4[i], w.z, acc[i][2]); acc[i][3] = fmaf(a4[i], w.w, acc[i][3]);
    }
  }
  const float4 bb = *reinterpret_cast<const float4 *>(b + 4 * jg);
  const float b4[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float o[RPT];
#pragma unroll
    for (int i = 0; i < RPT; ++i) o[i] = act_fused(act, acc[i][c] + b4[c]);
    st_rows<RPT>(CT + (4 * jg + c) * LD16 + r0, o);
  }
}

template <int RPT>
__global__ void __launch_bounds__(32 * 16 / (2 * RPT), 2) rollout_linquad_kernel_rt(RolloutArgs g) {
  constexpr int WS = 2 * RPT, NS = 32 / WS, NTH = 32 * 16 / WS;   // streams per warp, task slots per stream, threads per CTA
  using M = SmemMapT<R>;
  extern __shared__ __align__(16) float sm[];
  float *sA = sm + M::W2T;                       // [32*32] env A   (regions of the carve-up the 16-row forward does not touch)
  float *sB = sm + M::W2T + LQ_MAX_S * LQ_MAX_S; // [32*16] env B
  float *aT = sm + M::AT;                        // [8][LD16] sampled actions
  float *taT = sm + M::AT + MAX_O * LD16;        // [8][LD16] tanh(a), later the |s'|^2 partials
  const NetDesc nd = g.net;
  const int I = nd.I, O = nd.O, sdim = I, adim = g.adim;
  const int t = threadIdx.x;
  stage_params(sm, nd, M::MBAR);
  float *spT = sm + M::H2T + H * LD16;           // [32][LD16] s'
  int *s_len = reinterpret_cast<int *>(spT + LQ_MAX_S * LD16);  // [16] episode lengths
  int *s_end = s_len + 16;                                      // [16] end flags of the current step
  float *nzT = spT + LQ_MAX_S * LD16 + 32;       // [32][LD16] env noise of the step
  for (int i = t; i < sdim * sdim; i += NTH) sA[i] = g.A[i];
  for (int i = t; i < sdim * adim; i += NTH) sB[i] = g.B[i];
  const float *P = sm + M::P;
  float *XT = sm + M::XT, *H1T = sm + M::H1T, *H2T = sm + M::H2T, *OT = sm + M::OT;
  const int rows = g.rows;
  const int64_t e0 = (int64_t)blockIdx.x * rows;
  const int lane = t & 31, rb = (t >> 5) * WS;    // first stream of this warp inside the tile
  const int r0 = rb + RPT * (lane >> 4), jg = lane & 15;
  const int rs = rb + (lane % WS), slot = lane / WS;
  const int64_t e = e0 + rs;
  const bool live = rs < rows && e < g.N;
  const int nlive = (int)max((int64_t)0, min((int64_t)min(WS, rows - rb), g.N - e0 - rb));   // live streams of this warp (a prefix of its WS)
  const unsigned long long tick0 = *(volatile unsigned long long *)g.tick_dev;
  const uint32_t inv_s = (65536u + (uint32_t)sdim - 1u) / (uint32_t)sdim, inv_o = (65536u + (uint32_t)O - 1u) / (uint32_t)O;   // q / d for q < 1024
  for (int q = t; q < R16 * sdim; q += NTH) {
    const int rr = q / sdim, i = q - rr * sdim;
    XT[i * LD16 + rr] = (rr < rows && e0 + rr < g.N) ? g.obs_io[(e0 + rr) * sdim + i] : 0.f;
  }
  if (t < R16) s_len[t] = (t < rows && e0 + t < g.N) ? g.ep_len[e0 + t] : 0;
  __syncthreads();

  const int n_steps = nlive > 0 ? g.T : 0;
  for (int step = 0; step < n_steps; ++step) {
    const int64_t row0 = (int64_t)step * g.N + e0 + rb;   // rollout row of this warp's first stream
    // s rows of this step: the warp's live streams are nlive * sdim contiguous floats
    for (int q = lane; q < nlive * sdim; q += 32) {
      const int rr = (int)(((uint32_t)q * inv_s) >> 16), kk = q - rr * sdim;
      g.s[row0 * sdim + q] = XT[kk * LD16 + rb + rr];
    }
    layer_fwd_rt<RPT>(XT, I, P, P + off_b1(I), H1T, nd.act, r0, jg);
    __syncwarp();
    layer_fwd_rt<RPT>(H1T, H, P + off_W2(I), P + off_b2(I), H2T, nd.act, r0, jg);
    __syncwarp();
    {   // output layer: lane (rs, slot) owns the outputs slot, slot + NS, ... (independent chains; bias first, then k ascending)
      const float *W3 = P + off_W3(I);
      constexpr int NO = (MAX_O + NS - 1) / NS;
      int oo[NO]; float a0[NO];
#pragma unroll
      for (int u = 0; u < NO; ++u) { oo[u] = slot + NS * u < O ? slot + NS * u : 0; a0[u] = P[off_b3(I, O) + oo[u]]; }
#pragma unroll 8
      for (int k = 0; k < H; ++k) {
        const float h = H2T[k * LD16 + rs];
#pragma unroll
        for (int u = 0; u < NO; ++u) a0[u] = fmaf(h, W3[k * O + oo[u]], a0[u]);
      }
#pragma unroll
      for (int u = 0; u < NO; ++u)
        if (slot + NS * u < O) OT[(slot + NS * u) * LD16 + rs] = a0[u];
    }
    __syncwarp();
    // Gaussian head, lane (stream, action-dimension pair): one Philox block + one Box-Muller yield both normals of the pair
    for (int j0 = 2 * slot; j0 < O; j0 += 2 * NS) {
      const Philox4 p = philox4x32_10(g.seed_pi, g.ctr0 + (uint64_t)step, (uint64_t)e * ((O + 3) / 4) + (j0 >> 2));
      float n0, n1;
      if ((j0 & 2) == 0) box_muller(p.x, p.y, n0, n1); else box_muller(p.z, p.w, n0, n1);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int j = j0 + u;
        if (j < O) {
          const float mu = OT[j * LD16 + rs];
          const float ls = g.ls[j];
          const float sigma = expf(ls);
          const float var = sigma * sigma;
          const float ev = u ? n1 : n0;
          const float act = ev * sigma + mu;
          aT[j * LD16 + rs] = act;
          const float dd = act - mu;
          OT[j * LD16 + rs] = -(dd * dd) / (2.f * var) - LOG_SQRT_2PI - ls;
        }
      }
    }
    __syncwarp();
    for (int q = lane; q < nlive * O; q += 32) {   // action rows out (contiguous)
      const int rr = (int)(((uint32_t)q * inv_o) >> 16), j = q - rr * O;
      g.a[row0 * O + q] = aT[j * LD16 + rb + rr];
    }
    // ---- env transition (identical arithmetic to linquad_step_kernel)
    for (int j = slot; j < adim; j += NS) taT[j * LD16 + rs] = tanhf(aT[j * LD16 + rs]);
    if (slot == NS - 1 && g.logp) {
      float logp = 0.f;
      for (int j = 0; j < O; ++j) logp += OT[j * LD16 + rs];
      if (live) g.logp[row0 - rb + rs] = logp;
    }
    const unsigned long long tick = tick0 + (unsigned long long)step;
    // env noise: lane (stream, pair slot) owns the dimension pairs slot, slot + NS, ... (one Philox block + one Box-Muller per pair)
    for (int k0 = 2 * slot; k0 < sdim; k0 += 2 * NS) {
      const Philox4 p = philox4x32_10(g.seed_env, tick, (uint64_t)(live ? e : 0) * 16 + (k0 >> 2));
      float x0, x1;
      if ((k0 & 2) == 0) box_muller(p.x, p.y, x0, x1); else box_muller(p.z, p.w, x0, x1);
      nzT[k0 * LD16 + rs] = x0;
      if (k0 + 1 < sdim) nzT[(k0 + 1) * LD16 + rs] = x1;
    }
    __syncwarp();
    // s'[kk] of RPT streams per lane: one scalar (matrix entry) + one RPT-wide load per RPT FFMA
    for (int kk = jg; kk < sdim; kk += 16) {
      float v[RPT];
#pragma unroll
      for (int i = 0; i < RPT; ++i) v[i] = 0.f;
      const float *ar = sA + kk * sdim, *br = sB + kk * adim;
#pragma unroll 4
      for (int j = 0; j < sdim; ++j) {
        float x[RPT];
        ld_rows<RPT>(XT + j * LD16 + r0, x);
        const float c = ar[j];
#pragma unroll
        for (int i = 0; i < RPT; ++i) v[i] = fmaf(c, x[i], v[i]);
      }
#pragma unroll 2
      for (int j = 0; j < adim; ++j) {
        float x[RPT];
        ld_rows<RPT>(taT + j * LD16 + r0, x);
        const float c = br[j];
#pragma unroll
        for (int i = 0; i < RPT; ++i) v[i] = fmaf(c, x[i], v[i]);
      }
      float nz[RPT];
      ld_rows<RPT>(nzT + kk * LD16 + r0, nz);
#pragma unroll
      for (int i = 0; i < RPT; ++i) v[i] = fminf(fmaxf(fmaf(0.01f, nz[i], v[i]), -10.f), 10.f);
      st_rows<RPT>(spT + kk * LD16 + r0, v);
    }
    __syncwarp();
    // |s'|^2 in the exact order of linquad_step_kernel: 8 sequential 4-dim partials, combined below like the xor-1/2/4 butterfly
    for (int q = slot; q < 8; q += NS) {
      float n2 = 0.f;
      for (int i = 0; i < 4 && 4 * q + i < sdim; ++i) { const float v = spT[(4 * q + i) * LD16 + rs]; n2 = fmaf(v, v, n2); }
      taT[q * LD16 + rs] = n2;
    }
    __syncwarp();
    if (slot == 0) {   // reward, termination and episode bookkeeping of stream rs
      float pq[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) pq[q] = taT[q * LD16 + rs];
      const float n2 = ((pq[0] + pq[1]) + (pq[2] + pq[3])) + ((pq[4] + pq[5]) + (pq[6] + pq[7]));
      float a2 = 0.f;
      for (int j = 0; j < adim; ++j) { const float av = aT[j * LD16 + rs]; a2 = fmaf(av, av, a2); }
      const float rew = 1.f - n2 / (float)sdim - 0.1f * a2 / (float)adim;
      const bool dn = fabsf(spT[rs]) > 5.f;
      const int len = s_len[rs] + 1;
      const bool end = dn || len >= g.max_steps || (g.force_end && step == g.T - 1);
      s_len[rs] = end ? 0 : len;
      s_end[rs] = end ? 1 : 0;
      if (live) { g.r[row0 - rb + rs] = rew; g.done[row0 - rb + rs] = dn ? 1 : 0; g.ee[row0 - rb + rs] = end ? 1 : 0; }
    }
    __syncwarp();
    // ---- s' rows out (contiguous); next observation tile: s', or a fresh initial state where the episode ended
    for (int q = lane; q < WS * sdim; q += 32) {
      const int rr = (int)(((uint32_t)q * inv_s) >> 16), kk = q - rr * sdim;
      const float v = spT[kk * LD16 + rb + rr];
      if (rr < nlive) g.sp[row0 * sdim + q] = v;
      XT[kk * LD16 + rb + rr] = v;
    }
    __syncwarp();
    if (s_end[rs]) {   // rare outside the forced end of the rollout: one Philox block yields 4 dimensions of the reset state
      for (int blk = slot; 4 * blk < sdim; blk += NS) {
        const Philox4 p = philox4x32_10(g.seed_env ^ 0x5851F42D4C957F2DULL, tick + 0x100000000ULL, (uint64_t)(live ? e : 0) * 16 + blk);
        const uint32_t u[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (4 * blk + c < sdim) XT[(4 * blk + c) * LD16 + rs] = (u32_to_unit_open(u[c]) * 2.f - 1.f) * 0.1f;
      }
    }
    __syncwarp();
  }
  __syncthreads();
  for (int q = t; q < R16 * sdim; q += NTH) {
    const int rr = q / sdim, i = q - rr * sdim;
    if (rr < rows && e0 + rr < g.N) g.obs_io[(e0 + rr) * sdim + i] = XT[i * LD16 + rr];
  }
  if (t < rows && e0 + t < g.N) g.ep_len[e0 + t] = s_len[t];
  // the last block to finish advances the env tick by T (every block has read it by then)
  __shared__ bool last;
  __threadfence();
  __syncthreads();
  if (t == 0) last = atomicAdd(g.tick_dev + 1, 1ULL) == (unsigned long long)gridDim.x - 1ULL;
  __syncthreads();
  if (last && t == 0) { g.tick_dev[0] += (unsigned long long)g.T; g.tick_dev[1] = 0ULL; __threadfence(); }
}

}  // namespace

extern "C" int32_t crux_linquad_rollout(crux_linquad *env, crux_gaussian *actor, int32_t T, int32_t force_end_last, float *obs_io,
                                        const crux_rollout_cols *cols, uint64_t seed, uint64_t ctr0) {
  if (!env || !actor) return CRUX_ERR_INVALID;
  crux_ctx *ctx = env->ctx;
  CRUX_REQUIRE(ctx, T >= 1 && obs_io && cols, "crux_linquad_rollout: bad arguments");
  CRUX_REQUIRE(ctx, cols->s && cols->a && cols->sp && cols->r && cols->done && cols->episode_end, "crux_linquad_rollout: NULL column");
  CRUX_REQUIRE(ctx, fused_rows_supported(actor), "crux_linquad_rollout: needs a GaussianPolicy with a fusable I-64-64-O network and a logΣ vector");
  CRUX_REQUIRE(ctx, actor->mu->dims[0] == env->sdim && actor->adim == env->adim && env->adim <= MAX_O, "crux_linquad_rollout: policy / env shapes differ");
  int rc = set_smem_attr(ctx); if (rc) return rc;
  static bool attr = false;
  if (!attr) {
    CRUX_CHECK_CUDA(ctx, cudaFuncSetAttribute(rollout_linquad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmemMapT<R>::BYTES));
    CRUX_CHECK_CUDA(ctx, cudaFuncSetAttribute(rollout_linquad_kernel_rt<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmemMapT<R>::BYTES));
    CRUX_CHECK_CUDA(ctx, cudaFuncSetAttribute(rollout_linquad_kernel_rt<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmemMapT<R>::BYTES));
    attr = true;
  }
  RolloutArgs g;
  memset(&g, 0, sizeof(g));
  g.net = describe(actor->mu); g.ls = actor->log_sigma; g.A = env->A; g.B = env->B; g.N = env->n_env; g.T = T; g.adim = env->adim;
  g.max_steps = env->max_steps; g.force_end = force_end_last; g.seed_pi = seed; g.ctr0 = ctr0; g.seed_env = env->seed; g.tick_dev = env->tick;
  g.ep_len = env->ep_len; g.obs_io = obs_io; g.s = cols->s; g.a = cols->a; g.sp = cols->sp; g.r = cols->r; g.logp = cols->logprob;
  g.done = cols->done; g.ee = cols->episode_end;
  // streams per CTA: 16 would give cdiv(N, 16) CTAs, e.g. 256 for 4096 streams = two CTAs on 108 SMs and one on 40; taking
  // ceil(N / (2 x SMs)) streams (14 -> 293 CTAs) loads every SM alike (the idle half-warps of a tile cost nothing)
  g.rows = (int)i64max(1, i64min(R16, cdiv(env->n_env, (int64_t)2 * ctx->num_sms)));
  {
    CruxTimed timed(ctx, CRUX_T_ENV);
    static const int rpt = getenv("CRUX_ROLLOUT_RPT") ? atoi(getenv("CRUX_ROLLOUT_RPT")) : 2;   // A/B: streams per thread (1 = the half-warp-per-stream kernel)
    const unsigned grid = (unsigned)cdiv(env->n_env, g.rows);
    if (rpt == 1) rollout_linquad_kernel<<<grid, NT, SmemMapT<R>::BYTES, ctx->stream>>>(g);
    else if (rpt == 4) rollout_linquad_kernel_rt<4><<<grid, 64, SmemMapT<R>::BYTES, ctx->stream>>>(g);
    else rollout_linquad_kernel_rt<2><<<grid, 128, SmemMapT<R>::BYTES, ctx->stream>>>(g);
  }
  CRUX_LAUNCHED(ctx);
  return CRUX_OK;
}
