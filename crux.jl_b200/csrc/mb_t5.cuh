// The PPO minibatch kernel with EVERY GEMM -- forward, data-backward AND the weight gradients -- on tcgen05, in the TRANSPOSED form:
// features on the 128 TMEM lanes, minibatch rows on the TMEM columns.
//
// Why transposed.  The weight gradients contract over ROWS.  With rows on the TMEM lanes (fwd_tc5.cuh / mb_tc5.cuh) that needs both
// operands transposed through shared memory as hi/lo planes (does not fit) or MN-major tf32 descriptors (only with swizzled
// layouts).  With FEATURES on the lanes (D^T[M = 64 features][N rows] = W[64][K] * X[rows][K]^T) every operand of every GEMM is
// K-major in its natural layout:
//     forward / data-backward : A = weight planes [out][in] (shared memory, built once per CTA), B = activation rows [row][feature]
//                               (the epilogue's store is contiguous across the lanes of a warp: lane = feature)
//     weight gradients        : A = an activation that already sits in TENSOR MEMORY as [feature lanes][row columns] -- exactly what the
//                               previous epilogue wrote back with tcgen05.st -- and B = the other factor as [feature][rows] (one float4
//                               per 4 rows and thread); the accumulators dW2^T / dW1 / dW3^T stay in TMEM across all tiles of the CTA
//   so h1^T, h2^T and dz1^T never leave tensor memory, and no activation is ever transposed.
//
// M = 64 accumulators occupy 16 of the 32 lanes of every TMEM quarter (lane = 32*(m/16) + dp + m%16, dp = 0 or 16:
// experiments/tc5_probe.cu), so a tile is TWO interleaved 32-row atoms (dp = 0 / 16) and the epilogue thread (warp quarter q, lane l)
// owns feature 16q + l%16 of atom l/16 -- all 32 lanes busy.  Measured instruction costs (experiments/tc5_probe{3,4}.cu, B200):
// 23.5 cycles per tcgen05.mma for M = 64, N <= 32 (N/2 above), *provided the MMA is issued from a warp-uniform branch with
// elect.sync*; issued under `if (threadIdx.x == 0)` the compiler wraps every MMA in an election loop and it costs 96 cycles.
//
// Per 64-row tile (256 threads; Z* / D* are TMEM column blocks of 32 = rows of an atom):
//   P0  staged rows -> S [row][24] hi/lo (B of L1) and S^T [24][rows] hi/lo (B of dW1)
//   L1  ZA  = W1 S^T                  E1  h1 = act(ZA + b1): hi -> ZA, lo -> H1L (A of dW2), [row][feature] hi/lo -> ACT (B of L2)
//   L2  ZB  = W2 h1                   E2  h2: hi -> ZB, lo -> H2L (A of dW3), ACT
//   L3  OUT[rows][8] = h2 W3^T        (row-major form: A = ACT, B = W3 [8][64])      head: loss terms, dOut -> DO [row][8], DO^T [8][rows]
//                                     E3  dh2 = dOut W3 (K <= 8: FP32 pipe), dz2 = dh2 .* act'(h2): ACT (B of G5), DZ^T [feature][rows] (B of dW2); db2
//   G5  ZD  = W2^T dz2     ||  dW3^T += h2^T dOut, dW2^T += h1^T dz2        E4  dz1 = ZD .* act'(h1): hi -> ZD, lo -> D1L (A of dW1); db1
//   G6  dW1 += dz1^T S                                   (completion awaited at the top of the next tile)
// 3xTF32 everywhere (lo*hi + hi*lo + hi*hi, fp32-level accuracy), same partial-gradient layout and head arithmetic as
// fused_minibatch_tc_kernel.  Included by ppo_fused.cu after mb_tc5.cuh.
#pragma once

namespace mb6 {

constexpr int NR = 32, KX = tc5::KX;   // a tile = one 32-row atom
// Two independent pipelines ("groups", g = 0 / 1) share a CTA and the weight planes; group g owns the TMEM data-path half dp = 16 g.
// warps [0, 16): epilogue (8 per group: TMEM quarter q, column half ch) | 16, 17: MMA issuer of group 0 / 1 | 18, 19: loader of group 0 / 1
constexpr int NEW = 8, NTE = NEW * 32, W_ISSUE = 2 * NEW, W_LOAD = 2 * NEW + 2, NTH = (2 * NEW + 4) * 32, NGT = (NEW + 2) * 32;
constexpr int NB_READY = 1, NB_PRO = 3, NB_EPI = 5;   // named barriers (+ g): epilogue -> issuer hand-off, cooperative first load, epilogue-only
constexpr int TMEM_COLS = 512;
// TMEM columns (all 512; a group uses the data-path half dp = 16 g of every block, the weight blocks are replicated in both halves):
//   ZA  z1 -> h1 hi | H1L h1 lo | ZB z2 -> h2 hi | H2L h2 lo, later dz1 lo (D1L) | ZC L3's output (8 columns), later dh1 -> dz1 hi (ZD)
//   DW2C / DW1C / DW3C weight-gradient accumulators | W2A / W2TA: W2 and W2^T as hi | lo A operands (A from tensor memory costs 16 cycles
//   per MMA against 24 from shared memory and takes 2 KB per MMA off the shared-memory pipe: experiments/tc5_probe5.cu)
constexpr uint32_t ZA = 0, H1L = 32, ZB = 64, H2L = 96, D1L = H2L, ZC = 128, ZD = ZC, OUTC = ZC, DW2C = 160, DW1C = 224, DW3C = 248;
constexpr uint32_t W2A_HI = 256, W2A_LO = 320, W2TA_HI = 384, W2TA_LO = 448;
constexpr int ACT_LBO = 144, ACT_SBO = 2320;   // padded K-major tile: the [row][feature] stores of a warp hit 32 distinct banks
constexpr int STG_SPLIT = 4 * ACT_SBO * 2 / 4;   // floats of a group's ACT buffer: the first piece of its staged partial gradient (the rest: its DZ^T buffer)

struct Map {   // bytes; every operand is a pair of planes hi | lo
  static constexpr int W1A = 0;                                  // A [64 j][24 i]   = W1(j,i)
  static constexpr int W3B = W1A + 2 * 64 * KX * 4;              // B [8 o][64 k]    = W3(o,k)
  static constexpr int PLANES = W3B + 2 * 8 * 64 * 4;            // bytes of the weight planes = crux_mlp::frag in plane mode (one TMA bulk copy)
  // per group g and buffer b (the loader prepares tile n+1 while tile n runs): S hi|lo [32 rows][24] (B of L1), S^T hi|lo [24][32 rows] (B of dW1)
  static constexpr int S = PLANES;
  static constexpr int S_PLANE = NR * KX * 4, ST_PLANE = 3 * 1024;
  static constexpr int IN_BUF = 2 * S_PLANE + 2 * ST_PLANE;      // bytes of one input buffer; buffer (g, b) at S + (2 g + b) IN_BUF
  static constexpr int ACT = S + 4 * IN_BUF;                     // per group [32 rows][64] padded, hi | lo (L3 reads 64 rows: the tail is don't-care)
  static constexpr int ACT_PLANE = 4 * ACT_SBO, ACT_G = 2 * ACT_PLANE;
  static constexpr int DZT = ACT + 2 * ACT_G;                    // per group B [64 o][32 rows] hi | lo   (L3's 64-row read of group 1's lo plane ends inside it)
  static constexpr int DZT_PLANE = 8 * 1024, DZT_G = 2 * DZT_PLANE;
  static constexpr int DO = DZT + 2 * DZT_G;                     // per group fp32 dOut [32 rows][8] (read back by E3: dh2 = dOut W3 on the FP32 pipe, K <= 8)
  static constexpr int DO_G = NR * 8 * 4;
  static constexpr int DOT = DO + 2 * DO_G;                      // per group B [8 o][32 rows] hi | lo
  static constexpr int DOT_PLANE = 1024, DOT_G = 2 * DOT_PLANE;
  static constexpr int SX = DOT + 2 * DOT_G;                     // per group gather staging: x [32][I <= 24] (loader-private)
  static constexpr int SX_G = NR * KX * 4;
  static constexpr int SA = SX + 2 * SX_G;                       // per (g, b) head inputs: actions [32][8] | logprob, advantage, return [32] each | rows [32] ints
  static constexpr int HEAD_BUF = NR * 8 * 4 + 3 * NR * 4 + NR * 4;
  static constexpr int BIAS = SA + 4 * HEAD_BUF;                 // b3[8] | logΣ[8] | 1/σ²[8]
  static constexpr int RED = BIAS + 24 * 4;                      // per group: [2 ch][64] db1 | [2 ch][64] db2 | [8][24] head scratch
  static constexpr int RED_G = (2 * 64 * 2 + 8 * 24) * 4;
  static constexpr int BAR = RED + 2 * RED_G;                    // bar_par | per group: bar_mma, bar_g[2], bar_free[2] | tmem slot
  static constexpr int TOTAL = BAR + 128;
  static_assert(ACT % 16 == 0 && DZT % 16 == 0 && DO % 16 == 0 && SX % 16 == 0 && BAR % 8 == 0, "alignment");
  static_assert(TOTAL <= 232448, "shared memory");
};

__device__ __forceinline__ int canon(int row, int k, int K) { return (row >> 3) * (32 * K) + (k >> 2) * 128 + (row & 7) * 16 + (k & 3) * 4; }
__device__ __forceinline__ int canon_act(int row, int k) { return (row >> 3) * ACT_SBO + (k >> 2) * ACT_LBO + (row & 7) * 16 + (k & 3) * 4; }
using tc5::elect_one;

// Weight planes in GLOBAL memory (crux_mlp::frag, Map::PLANES bytes, same offsets as the shared-memory map): built once at the start
// of an update (build_planes_kernel) and kept current by the Adam kernels, which scatter every updated parameter into its plane
// positions -- the minibatch kernel stages them with ONE TMA bulk copy.  (W1 as the A operand of L1, W3 as the B operand of L3.)
__device__ __forceinline__ void plane_put(unsigned char *planes, int base, int plane_bytes, int off, float v) {
  float hi, lo;
  tc5::split(v, hi, lo);
  *reinterpret_cast<float *>(planes + base + off) = hi;
  *reinterpret_cast<float *>(planes + base + plane_bytes + off) = lo;
}
__device__ __forceinline__ void plane_scatter(float *planes_f, int I, int O, int i, float v) {   // i = index in the flat parameter vector
  unsigned char *pl = reinterpret_cast<unsigned char *>(planes_f);
  if (i < I * H) { const int in = i >> 6, j = i & 63; plane_put(pl, Map::W1A, 64 * KX * 4, canon(j, in, KX), v); return; }
  const int e3 = i - off_W3(I);   // (W2 and W2^T are loaded into tensor memory from the parameter vector by every CTA)
  if (e3 >= 0 && e3 < H * O) {
    const int k = e3 / O, o = e3 - k * O;
    plane_put(pl, Map::W3B, 8 * 64 * 4, canon(o, k, 64), v);
  }
}
__global__ void build_planes_kernel(const float *__restrict__ params, float *__restrict__ planes, int I, int O, int n_params) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_params) plane_scatter(planes, I, O, i, params[i]);   // the padding (inputs >= I, outputs >= O) was zeroed by the memset before
}
// 3 passes x KS k-steps; descriptors advance by a_adv / b_adv (16-byte units) per k-step.  `acc` = 0: the first MMA overwrites D.
template <int KS>
__device__ __forceinline__ void gemm_ss(uint32_t d, uint64_t a_hi, uint64_t a_lo, uint32_t a_adv, uint64_t b_hi, uint64_t b_lo, uint32_t b_adv, uint32_t idesc,
                                        uint32_t acc) {
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) tc5::mma_tf32_ss(d, a_lo + (uint64_t)(a_adv * ks), b_hi + (uint64_t)(b_adv * ks), idesc, (ks || acc) ? 1u : 0u);
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) tc5::mma_tf32_ss(d, a_hi + (uint64_t)(a_adv * ks), b_lo + (uint64_t)(b_adv * ks), idesc, 1u);
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) tc5::mma_tf32_ss(d, a_hi + (uint64_t)(a_adv * ks), b_hi + (uint64_t)(b_adv * ks), idesc, 1u);
}
template <int KS>
__device__ __forceinline__ void gemm_ts(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint64_t b_hi, uint64_t b_lo, uint32_t b_adv, uint32_t idesc, uint32_t acc) {
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) tc5::mma_tf32_ts(d, a_lo + 8 * ks, b_hi + (uint64_t)(b_adv * ks), idesc, (ks || acc) ? 1u : 0u);
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) tc5::mma_tf32_ts(d, a_hi + 8 * ks, b_lo + (uint64_t)(b_adv * ks), idesc, 1u);
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) tc5::mma_tf32_ts(d, a_hi + 8 * ks, b_hi + (uint64_t)(b_adv * ks), idesc, 1u);
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

#define MB6_LD16(v, taddr)                                                                                                                               \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"                     \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),      \
                 "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])                                                                          \
               : "r"(taddr)                                                                                                                               \
               : "memory")
#define MB6_ST16(taddr, v)                                                                                                                               \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),       \
               "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),      \
               "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])                                                                                             \
               : "memory")
#define MB6_ST8(taddr, v)                                                                                                                                \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),    \
               "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])                                                                                               \
               : "memory")
#define MB6_LD8(v, taddr)                                                                                                                                \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"                                                            \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])                                           \
               : "r"(taddr)                                                                                                                               \
               : "memory")
// 16 data-path lanes x 256 bit, twice: thread t of the warp holds lanes t/4 (v0 v1 | v4 v5) and t/4 + 8 (v2 v3 | v6 v7), columns
// 2 (t%4) + {0, 1} (v0..v3) and 8 + 2 (t%4) + {0, 1} (v4..v7) of the 16 lanes x 16 columns at taddr -- all 32 threads work on ONE atom
#define MB6_LDX2(v, taddr)                                                                                                                               \
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"                                                           \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])                                           \
               : "r"(taddr)                                                                                                                               \
               : "memory")
#define MB6_STX2(taddr, v)                                                                                                                               \
  asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),   \
               "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])                                                                                               \
               : "memory")
#define MB6_WAIT_LD() asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")
#define MB6_WAIT_ST() asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory")
// gather + hi/lo split of one 32-row tile of group g into input buffer b, by the `NL` threads lt = 0 .. NL-1 (sync() orders them)
template <int HEAD, class Sync>
__device__ __forceinline__ void load_tile(unsigned char *smb, const MbArgs &a, int I, int O, int g, int b, int64_t atile, int lt, int NL, Sync sync) {
  float *SXp = reinterpret_cast<float *>(smb + Map::SX + g * Map::SX_G);
  float *SAp = reinterpret_cast<float *>(smb + Map::SA + (2 * g + b) * Map::HEAD_BUF), *SHp = SAp + NR * 8;
  int *sidx = reinterpret_cast<int *>(SHp + 3 * NR);
  const uint32_t inv_I = (65536u + (uint32_t)I - 1u) / (uint32_t)I, inv_O = (65536u + (uint32_t)O - 1u) / (uint32_t)O;
  for (int r = lt; r < NR; r += NL) {
    const int64_t row = atile * NR + r;
    sidx[r] = row < a.bm ? (a.order ? a.order[row] : (int)row) : -1;
  }
  sync();
  for (int e = lt; e < NR * I; e += NL) {
    const int r = (int)(((uint32_t)e * inv_I) >> 16), i = e - r * I;
    const int row = sidx[r];
    cp_async4(SXp + e, a.s + (row >= 0 ? (int64_t)row * I + i : 0), row >= 0);
  }
  if (HEAD == 0)
    for (int e = lt; e < NR * O; e += NL) {
      const int r = (int)(((uint32_t)e * inv_O) >> 16), o = e - r * O;
      const int row = sidx[r];
      cp_async4(SAp + e, a.act + (row >= 0 ? (int64_t)row * O + o : 0), row >= 0);
    }
  for (int r = lt; r < NR; r += NL) {
    const int row = sidx[r];
    if (HEAD == 0) {
      cp_async4(SHp + r, a.logp_old + (row >= 0 ? row : 0), row >= 0);
      cp_async4(SHp + NR + r, a.adv + (row >= 0 ? row : 0), row >= 0);
    }
    const bool has_ret = a.ret != nullptr;
    cp_async4(SHp + 2 * NR + r, has_ret ? a.ret + (row >= 0 ? row : 0) : a.s, has_ret && row >= 0);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_all;" ::: "memory");
  sync();
  unsigned char *Sb = smb + Map::S + (2 * g + b) * Map::IN_BUF, *STb = Sb + 2 * Map::S_PLANE;
  for (int e = lt; e < NR * I; e += NL) {   // x -> S [row][24] (B of L1) and S^T [24][32 rows] (B of dW1), hi/lo
    const int r = (int)(((uint32_t)e * inv_I) >> 16), i = e - r * I;
    float hi, lo;
    tc5::split(SXp[e], hi, lo);
    const int o1 = canon(r, i, KX), o2 = canon(i, r, NR);
    *reinterpret_cast<float *>(Sb + o1) = hi;
    *reinterpret_cast<float *>(Sb + Map::S_PLANE + o1) = lo;
    *reinterpret_cast<float *>(STb + o2) = hi;
    *reinterpret_cast<float *>(STb + Map::ST_PLANE + o2) = lo;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// tanh_fast of ppo_fused.cu (1 - 2 / (exp(2x) + 1) on the SFU) without the denormal-range fix-up of __expf: branch-free, so that the
// 8 independent element chains of an epilogue thread interleave (with the activation as a RUN-TIME switch every element was a
// branch and the MUFU latencies added up: ~100 cycles per element, measured as ~1000-cycle epilogues)
__device__ __forceinline__ float tanh_t5(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"((2.0f * x) * 1.4426950216293334961f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
  return fmaf(-2.0f, r, 1.0f);
}
template <int ACT> __device__ __forceinline__ float actf(float z) { return ACT == CRUX_ACT_TANH ? tanh_t5(z) : fmaxf(z, 0.0f); }
template <int ACT> __device__ __forceinline__ float dactf(float y) { return ACT == CRUX_ACT_TANH ? fmaf(-y, y, 1.0f) : (y > 0.0f ? 1.0f : 0.0f); }

// 80 registers per thread: 640 x 80 leaves 14 k registers (and 70 KB of shared memory) on the SM for the 256-thread CTAs of the OTHER
// network's update tail (reduce_adam_kernel), which would otherwise wait for this kernel to drain
template <int HEAD, int ACT>
__global__ void __maxnreg__(80) minibatch_kernel(MbArgs a) {
  extern __shared__ __align__(1024) unsigned char smb[];
  const NetDesc nd = a.net;
  const int I = nd.I, O = nd.O;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  // group of this warp and its index inside the group's 320 threads (epilogue 0..255 | issuer 256..287 | loader 288..319)
  const int g = w < 2 * NEW ? w / NEW : (w - 2 * NEW) & 1;
  const int gt = w < 2 * NEW ? t - g * NTE : (w < W_LOAD ? NTE + lane : NTE + 32 + lane);
  int prof_n = 0;
#define MB6_STAMP() do { if (a.prof && blockIdx.x == 0 && t == 0 && prof_n < 64) a.prof[prof_n++] = clock64(); } while (0)
#define MB6_ISTAMP() do { if (a.prof && blockIdx.x == 0 && w == W_ISSUE && lane == 0 && prof_n < 64) a.prof[64 + prof_n++] = clock64(); } while (0)
  MB6_STAMP();
  if (a.trace && t == 0) { unsigned long long gt_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_)); a.trace[2 * blockIdx.x] = gt_; }
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // this network's tail kernel may be scheduled (it waits for this grid to complete)
  const uint32_t bar_par = smem_u32(smb + Map::BAR);
  const uint32_t bar_mma = smem_u32(smb + Map::BAR + 8 + 48 * g), bar_g = bar_mma + 8, bar_free = bar_mma + 24, bar_dw3 = bar_mma + 40;   // bar_g[2], bar_free[2]
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smb + Map::BAR + 112);
  int *pipe_token = reinterpret_cast<int *>(smb + Map::BAR + 116);   // tensor-pipe token shared by the two issuer warps
  float *bias = reinterpret_cast<float *>(smb + Map::BIAS);   // b3[8] | logΣ[8] | 1/σ²[8]
  const int64_t n_tiles = (a.bm + NR - 1) / NR;
  const int64_t tile0 = 2 * (int64_t)blockIdx.x + g, tstride = 2 * (int64_t)gridDim.x;   // this group's tiles: tile0, tile0 + tstride, ...

  // ---- prologue (all warps): barriers, weight planes (ONE TMA bulk copy), TMEM, zero the K padding of the input buffers
  if (t == 0) {
    *pipe_token = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_par), "r"(1) : "memory");
    for (int gg = 0; gg < 2; ++gg) {
      const uint32_t bm_ = smem_u32(smb + Map::BAR + 8 + 48 * gg);
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bm_), "r"(1) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bm_ + 40), "r"(1) : "memory");
      for (int b = 0; b < 2; ++b) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bm_ + 8 + 8 * b), "r"(32) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bm_ + 24 + 8 * b), "r"(1) : "memory");
      }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (w == W_ISSUE) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int e = t; e < 4 * Map::IN_BUF / 4; e += NTH) reinterpret_cast<float *>(smb + Map::S)[e] = 0.f;   // K padding stays zero
  for (int e = t; e < (2 * Map::DO_G + 2 * Map::DOT_G) / 4; e += NTH) reinterpret_cast<float *>(smb + Map::DO)[e] = 0.f;   // outputs >= O stay zero
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  // Everything above and the first tile's gather are independent of the previous Adam step: with a programmatic dependent launch they
  // run while this network's previous tail kernel is still in flight.  Parameters, planes and the KL-stop flag are read behind the wait.
  if (w >= 2 * NEW && tile0 < n_tiles) {
    // the group's FIRST tile is loaded by its issuer + loader warps (64 threads) while the epilogue warps fill tensor memory with W2
    load_tile<HEAD>(smb, a, I, O, g, 0, tile0, gt - NTE, 64, [&]() { asm volatile("bar.sync %0, %1;" ::"r"(NB_PRO + g), "n"(64) : "memory"); });
    asm volatile("bar.sync %0, %1;" ::"r"(NB_PRO + g), "n"(64) : "memory");
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int stop_at = a.ctl ? a.ctl[1] : 0;
  const bool skip = stop_at != 0 && stop_at <= a.mb;   // an EARLIER minibatch raised the KL stop flag (rl/ppo.jl:59): nothing to do
  if (t == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_par), "r"((uint32_t)Map::PLANES) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smb)), "l"(a.planes),
                 "r"((uint32_t)Map::PLANES), "r"(bar_par)
                 : "memory");
  }
  if (t < 8) bias[t] = t < O ? __ldg(nd.params + off_b3(I, O) + t) : 0.f;
  if (HEAD == 0 && t >= 32 && t < 40) { const int j = t - 32; const float ls = j < O ? a.ls[j] : 0.f, sg = expf(ls); bias[8 + j] = ls; bias[16 + j] = 1.0f / (sg * sg); }
  if (skip) {
    tc5::mbar_wait(bar_par, 0);
    __syncthreads();
    if (w == W_ISSUE) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
    return;
  }
  if (w >= 2 * NEW) {
    if (w >= W_LOAD && tile0 < n_tiles) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_g) : "memory");   // buffer 0 of this group is ready
  } else {   // W2 (A of L2) and W2^T (A of G5) -> tensor memory, hi | lo, replicated in both data-path halves:
    // warp (quarter wq, part) owns lanes [32 wq, +32) (feature m = 16 wq + lane % 16 in either half) x 32 of the 128 K-columns
    const int wq = w & 3, part = w >> 2, m = 16 * wq + (lane & 15);
    const float *W2p = nd.params + off_W2(I);   // flat W2[k * 64 + j] = W2(out j, in k)
    uint32_t hi[32], lo[32];
    if (part < 2) {            // W2A[m = j][k]: k = 32 part .. + 31
#pragma unroll
      for (int c = 0; c < 32; ++c) { float h, l; tc5::split(__ldg(W2p + (32 * part + c) * H + m), h, l); hi[c] = __float_as_uint(h); lo[c] = __float_as_uint(l); }
    } else {                   // W2TA[m = k][j]: j = 32 (part - 2) .. + 31
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4) {
        const float4 x = __ldg(reinterpret_cast<const float4 *>(W2p + m * H + 32 * (part - 2)) + c4);
        float h, l;
        tc5::split(x.x, h, l); hi[4 * c4] = __float_as_uint(h); lo[4 * c4] = __float_as_uint(l);
        tc5::split(x.y, h, l); hi[4 * c4 + 1] = __float_as_uint(h); lo[4 * c4 + 1] = __float_as_uint(l);
        tc5::split(x.z, h, l); hi[4 * c4 + 2] = __float_as_uint(h); lo[4 * c4 + 2] = __float_as_uint(l);
        tc5::split(x.w, h, l); hi[4 * c4 + 3] = __float_as_uint(h); lo[4 * c4 + 3] = __float_as_uint(l);
      }
    }
    const uint32_t dst = tmem + ((uint32_t)(32 * wq) << 16) + (part < 2 ? W2A_HI + 32 * part : W2TA_HI + 32 * (part - 2));
    TC5_ST32(dst, hi);
    TC5_ST32(dst + 64, lo);   // the lo block follows the hi block
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  MB6_STAMP();   // end of the prologue
  const uint32_t tm_g = tmem + ((uint32_t)(16 * g) << 16);   // this group's data-path half

  if (w >= W_LOAD) {
    // =========================================================== LOADER warp of group g: one tile ahead of the pipeline
    int k = 1;
    for (int64_t tile = tile0 + tstride; tile < n_tiles; tile += tstride, ++k) {
      const int b = k & 1;
      if (k >= 2) tc5::mbar_wait(bar_free + 8 * b, (uint32_t)(((k >> 1) - 1) & 1));   // dW1 of the tile that used this buffer has completed
      load_tile<HEAD>(smb, a, I, O, g, b, tile, lane, 32, [&]() { __syncwarp(); });
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_g + 8 * b) : "memory");
    }
  } else if (w >= W_ISSUE) {
    // =========================================================== MMA ISSUER warp of group g (one elected lane issues; the warp stays converged)
    tc5::mbar_wait(bar_par, 0);   // weight planes have landed
    const uint32_t sb = smem_u32(smb);
    const uint64_t dW1A = tc5::make_desc(sb + Map::W1A, 128, 32 * KX), dW1A_lo = tc5::make_desc(sb + Map::W1A + 64 * KX * 4, 128, 32 * KX);
    const uint64_t dW3B = tc5::make_desc(sb + Map::W3B, 128, 32 * 64), dW3B_lo = tc5::make_desc(sb + Map::W3B + 8 * 64 * 4, 128, 32 * 64);
    const uint64_t dACT = tc5::make_desc(sb + Map::ACT + g * Map::ACT_G, ACT_LBO, ACT_SBO), dACT_lo = dACT + (uint64_t)(Map::ACT_PLANE >> 4);
    const uint64_t dDZT = tc5::make_desc(sb + Map::DZT + g * Map::DZT_G, 128, 32 * NR), dDZT_lo = dDZT + (uint64_t)(Map::DZT_PLANE >> 4);
    const uint64_t dDOT = tc5::make_desc(sb + Map::DOT + g * Map::DOT_G, 128, 32 * NR), dDOT_lo = dDOT + (uint64_t)(Map::DOT_PLANE >> 4);
    constexpr uint32_t ADV = 16, ADV_ACT = (2 * ACT_LBO) >> 4;           // one k-step = two core matrices along K
    const uint32_t id32 = tc5::make_idesc(64, 32), id8 = tc5::make_idesc(64, 8), id64 = tc5::make_idesc(64, 64), id24 = tc5::make_idesc(64, 24);
    uint32_t first = 0;   // 0 on this group's first tile: the dW accumulators are overwritten, then accumulated
    int k = 0;
    // The tensor pipe executes the two groups' MMAs strictly interleaved when both issue at once (measured: 53 instead of 24 cycles per
    // MMA for BOTH chains), which keeps the two pipelines phase-locked: both wait for their GEMMs, then both run their epilogues.  A
    // chain is therefore issued under a token: the first collision serialises the two chains and from then on one group's GEMM runs
    // under the other group's epilogue.  (The issue loop is back-pressured by the pipe, so holding the token while issuing = owning it.)
    auto pipe_acquire = [&]() { while (atomicCAS(pipe_token, 0, 1) != 0) { } };
    auto pipe_release = [&]() { atomicExch(pipe_token, 0); };
#define MB6_WAIT_READY()                                                       \
    asm volatile("bar.sync %0, %1;" ::"r"(NB_READY + g), "n"(NTE + 32) : "memory");  \
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory")
    auto in_desc = [&](int bb, uint64_t &dS, uint64_t &dS_lo, uint64_t &dST, uint64_t &dST_lo) {
      dS = tc5::make_desc(sb + Map::S + (2 * g + bb) * Map::IN_BUF, 128, 32 * KX); dS_lo = dS + (uint64_t)(Map::S_PLANE >> 4);
      dST = tc5::make_desc(sb + Map::S + (2 * g + bb) * Map::IN_BUF + 2 * Map::S_PLANE, 128, 32 * NR); dST_lo = dST + (uint64_t)(Map::ST_PLANE >> 4);
    };
    auto issue_L1 = [&](int kk) {   // L1 of this group's kk-th tile: ZA = W1 S^T (waits for the loader's buffer)
      uint64_t dS, dS_lo, dST, dST_lo;
      in_desc(kk & 1, dS, dS_lo, dST, dST_lo);
      tc5::mbar_wait(bar_g + 8 * (kk & 1), (uint32_t)((kk >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        pipe_acquire();
        gemm_ss<KX / 8>(tm_g + ZA, dW1A, dW1A_lo, ADV, dS, dS_lo, ADV, id32, 0u);
        commit(bar_mma);
        pipe_release();
      }
      __syncwarp();
    };
    if (tile0 < n_tiles) issue_L1(0);
    for (int64_t tile = tile0; tile < n_tiles; tile += tstride, ++k) {
      const int b = k & 1;
      uint64_t dS, dS_lo, dST, dST_lo;
      in_desc(b, dS, dS_lo, dST, dST_lo);
      MB6_WAIT_READY();    // E1 done
      MB6_ISTAMP();
      if (elect_one()) {   // L2: ZB = W2 h1
        pipe_acquire();
        gemm_ts<8>(tm_g + ZB, tm_g + W2A_HI, tm_g + W2A_LO, dACT, dACT_lo, ADV_ACT, id32, 0u);
        commit(bar_mma);
        pipe_release();
      }
      __syncwarp();
      MB6_ISTAMP();
      MB6_WAIT_READY();    // E2 done
      MB6_ISTAMP();
      if (elect_one()) {   // L3 (rows on the lanes, M = 64: rows 32.. of the A operand are don't-care): OUT[rows][8] = h2 W3^T
        pipe_acquire();
        gemm_ss<8>(tm_g + OUTC, dACT, dACT_lo, ADV_ACT, dW3B, dW3B_lo, ADV, id8, 0u);
        commit(bar_mma);
        pipe_release();
      }
      __syncwarp();
      MB6_ISTAMP();
      MB6_WAIT_READY();    // head + E3 done (dh2 = dOut W3 has K <= 8: FP32 pipe inside E3, no GEMM round trip)
      if (elect_one()) {   // G5: ZD = W2^T dz2, then (under E4) dW3^T += h2^T dOut, dW2^T += h1^T dz2
        pipe_acquire();
        gemm_ts<8>(tm_g + ZD, tm_g + W2TA_HI, tm_g + W2TA_LO, dACT, dACT_lo, ADV_ACT, id32, 0u);
        commit(bar_mma);
        pipe_release();   // the other group's critical chain may go first
        pipe_acquire();
        gemm_ts<NR / 8>(tm_g + DW3C, tm_g + ZB, tm_g + H2L, dDOT, dDOT_lo, ADV, id8, first);
        commit(bar_dw3);  // h2 lo has been read: E4 may store dz1 lo over it
        gemm_ts<NR / 8>(tm_g + DW2C, tm_g + ZA, tm_g + H1L, dDZT, dDZT_lo, ADV, id64, first);
        pipe_release();
      }
      __syncwarp();
      MB6_WAIT_READY();    // E4 done
      if (elect_one()) {   // G6: dW1 += dz1^T S; its completion frees this tile's input buffer for the loader
        pipe_acquire();
        gemm_ts<NR / 8>(tm_g + DW1C, tm_g + ZD, tm_g + D1L, dST, dST_lo, ADV, id24, first);
        commit(bar_free + 8 * b);
        pipe_release();
      }
      __syncwarp();
      first = 1u;
      if (tile + tstride < n_tiles) issue_L1(k + 1);   // the next tile's L1 runs behind dW1: its round trip is off the critical path
    }
    if (elect_one()) commit(bar_mma);   // every weight-gradient MMA of this group has completed: its epilogue warps read the accumulators
    __syncwarp();
#undef MB6_WAIT_READY
  } else {
    // =========================================================== EPILOGUE warps of group g
    // warp (q, ch) owns TMEM lanes [32q + 16g, +16) x columns [16ch, +16); thread: features fa, fa + 8, rows c0 + 8 rep + rr + {0, 1}
    const int we = w - g * NEW, q = we & 3, ch = we >> 2;
    const int fa = 16 * q + (lane >> 2), rr = 2 * (lane & 3), c0 = 16 * ch;
    const float b1a = __ldg(nd.params + off_b1(I) + fa), b1b = __ldg(nd.params + off_b1(I) + fa + 8);
    const float b2a = __ldg(nd.params + off_b2(I) + fa), b2b = __ldg(nd.params + off_b2(I) + fa + 8);
    float w3a[MAX_O], w3b[MAX_O];   // W3(o, fa), W3(o, fa + 8): dh2 = dOut W3 in E3
#pragma unroll
    for (int o = 0; o < MAX_O; ++o) {
      w3a[o] = o < O ? __ldg(nd.params + off_W3(I) + fa * O + o) : 0.f;
      w3b[o] = o < O ? __ldg(nd.params + off_W3(I) + (fa + 8) * O + o) : 0.f;
    }
    const uint32_t quad_addr = tmem + ((uint32_t)(32 * q) << 16);           // 32x32b accesses (head, publication)
    const uint32_t atom_addr = tm_g + ((uint32_t)(32 * q) << 16) + c0;      // 16x256b accesses of this warp's 16 lanes x 16 columns
    unsigned char *act_a = smb + Map::ACT + g * Map::ACT_G + canon_act(c0 + rr, fa);   // (row c0 + rr, feature fa); fa + 8: + 2 LBO; rows + 8: + SBO
    unsigned char *dzt_a = smb + Map::DZT + g * Map::DZT_G + canon(fa, c0 + rr, NR);   // (feature fa, row c0 + rr); fa + 8: + 1024; rows + 8: + 256
    uint32_t ph = 0;
    float db1a = 0.f, db1b = 0.f, db2a = 0.f, db2b = 0.f;   // bias gradients of features fa, fa + 8 over this thread's rows
    float s_obj = 0.f, s_kl = 0.f, s_clip = 0.f, s_adv = 0.f, s_ret = 0.f, dls[MAX_O], db3[MAX_O];
#pragma unroll
    for (int j = 0; j < MAX_O; ++j) { dls[j] = 0.f; db3[j] = 0.f; }
    bool any = false;
    // generic-proxy shared-memory stores + tcgen05.st results -> visible to the MMAs the issuer launches after the hand-off
#define MB6_READY()                                                     \
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        \
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");    \
    asm volatile("bar.arrive %0, %1;" ::"r"(NB_READY + g), "n"(NTE + 32) : "memory")
#define MB6_WAIT_MMA()                                                  \
    tc5::mbar_wait(bar_mma, ph); ph ^= 1;                               \
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory")
    // element e of the 8 a thread holds: feature fa + 8 ((e >> 1) & 1), row c0 + 8 (e >> 2) + rr + (e & 1)
#define MB6_ACT_OFF(e) (((e) >> 2) * ACT_SBO + (((e) >> 1) & 1) * 2 * ACT_LBO + ((e) & 1) * 16)
#define MB6_HIDDEN_EPILOGUE(ZCOL, LCOL, BA, BB)                                                                          \
    {                                                                                                                    \
      uint32_t v[8], l[8];                                                                                               \
      MB6_LDX2(v, atom_addr + ZCOL);                                                                                     \
      MB6_WAIT_LD();                                                                                                     \
      _Pragma("unroll") for (int e = 0; e < 8; ++e) {                                                                    \
        float hi, lo;                                                                                                    \
        tc5::split(actf<ACT>(__uint_as_float(v[e]) + (((e >> 1) & 1) ? BB : BA)), hi, lo);                          \
        v[e] = __float_as_uint(hi); l[e] = __float_as_uint(lo);                                                          \
        *reinterpret_cast<float *>(act_a + MB6_ACT_OFF(e)) = hi;                                                         \
        *reinterpret_cast<float *>(act_a + Map::ACT_PLANE + MB6_ACT_OFF(e)) = lo;                                        \
      }                                                                                                                  \
      MB6_STX2(atom_addr + ZCOL, v);                                                                                     \
      MB6_STX2(atom_addr + LCOL, l);                                                                                     \
      MB6_WAIT_ST();                                                                                                     \
    }
    int k = 0;
    for (int64_t tile = tile0; tile < n_tiles; tile += tstride, ++k) {
      const int b = k & 1;
      any = true;
      MB6_WAIT_MMA();
      MB6_STAMP();   // L1 done
      MB6_HIDDEN_EPILOGUE(ZA, H1L, b1a, b1b)
      MB6_READY();
      MB6_WAIT_MMA();
      MB6_STAMP();   // E1 + L2 done
      MB6_HIDDEN_EPILOGUE(ZB, H2L, b2a, b2b)
      MB6_READY();
      MB6_WAIT_MMA();
      MB6_STAMP();   // E2 + L3 done
      if (HEAD == 0) {
        // Actor loss head, spread over (row, action dimension) pairs.  With one thread per row the head was a 2 700-cycle serial stretch on
        // two half-warps per tile (the critic's takes 350: scripts/mb6_prof.py), four times per launch on the update's critical chain.
        // Phase A: the 32 threads that can address the rows' TMEM lanes park mu = out + b3 in the dOut rows [row][8].
        const float *SAp = reinterpret_cast<const float *>(smb + Map::SA + (2 * g + b) * Map::HEAD_BUF), *SHp = SAp + NR * 8;
        const int *sidx = reinterpret_cast<const int *>(SHp + 3 * NR);
        float *DOp = reinterpret_cast<float *>(smb + Map::DO + g * Map::DO_G);
        tc5::mbar_wait(bar_g + 8 * b, (uint32_t)((k >> 1) & 1));   // (long complete) the cp.async data of this tile is visible
        if (ch == 0 && q < 2) {
          uint32_t v[8];
          MB6_LD8(v, quad_addr + OUTC);
          MB6_WAIT_LD();
          if ((lane >> 4) == g) {
            const int row = 16 * q + (lane & 15);
            float4 *p = reinterpret_cast<float4 *>(DOp + row * 8);
            p[0] = make_float4(__uint_as_float(v[0]) + bias[0], __uint_as_float(v[1]) + bias[1], __uint_as_float(v[2]) + bias[2], __uint_as_float(v[3]) + bias[3]);
            p[1] = make_float4(__uint_as_float(v[4]) + bias[4], __uint_as_float(v[5]) + bias[5], __uint_as_float(v[6]) + bias[6], __uint_as_float(v[7]) + bias[7]);
          }
        }
        asm volatile("bar.sync %0, %1;" ::"r"(NB_EPI + g), "n"(NTE) : "memory");
        // Phase B: thread (row, j) = (te >> 3, te & 7) of the group's 256 epilogue threads; the 8 lanes of a row add their logpdf terms with
        // a shuffle tree, every lane evaluates the (cheap) ratio / clip logic of its row, lane j owns dOut_j, dlogΣ_j and db3_j.
        {
          const int te_ = t - g * NTE, hrow = te_ >> 3, hj = te_ & 7;
          const bool live = sidx[hrow] >= 0;
          float d = 0.f, ivar = 0.f, logp = 0.f;
          if (hj < O) {
            d = SAp[hrow * O + hj] - DOp[hrow * 8 + hj];
            ivar = bias[16 + hj];
            logp = -(d * d) * (0.5f * ivar) - LOG_SQRT_2PI - bias[8 + hj];
          }
          logp += __shfl_xor_sync(0xffffffffu, logp, 1);
          logp += __shfl_xor_sync(0xffffffffu, logp, 2);
          logp += __shfl_xor_sync(0xffffffffu, logp, 4);
          const float Ai = SHp[NR + hrow], old = SHp[hrow];
          float dlogp = 0.f;
          if (live) {
            float obj, clipv = 0.f;
            if (a.a2c) {
              obj = logp * Ai;
              dlogp = -a.lambda_p * a.inv_bg * Ai;
            } else {
              const float rt = expf(logp - old);
              const float lo = 1.f - a.eps_clip, hi = 1.f + a.eps_clip;
              const float x = rt * Ai, y = fminf(fmaxf(rt, lo), hi) * Ai;
              const bool firstb = !(y < x);  // min(x, y) keeps x on ties
              obj = firstb ? x : y;
              dlogp = firstb ? -a.lambda_p * a.inv_bg * x : 0.f;
              clipv = (rt > hi || rt < lo) ? 1.f : 0.f;
            }
            if (hj == 0) { s_obj += obj; s_clip += clipv; s_kl += old - logp; s_adv += Ai; s_ret += SHp[2 * NR + hrow]; }
          }
          const float dout = dlogp * d * ivar;             // 0 for j >= O (d = ivar = 0) and for padding rows (dlogp = 0)
          if (hj < O) dls[0] += dlogp * (d * d * ivar - 1.f);   // this thread's OWN action dimension hj (the accumulator slots are per thread)
          db3[0] += dout;
          __syncwarp();                                        // every lane of the row has read mu before lane j overwrites its slot
          DOp[hrow * 8 + hj] = dout;
          float hi_, lo_;
          tc5::split(dout, hi_, lo_);
          unsigned char *pt = smb + Map::DOT + g * Map::DOT_G + canon(0, hrow, NR) + 16 * hj;
          *reinterpret_cast<float *>(pt) = hi_;
          *reinterpret_cast<float *>(pt + Map::DOT_PLANE) = lo_;
        }
      } else
      if (ch == 0 && q < 2) {   // loss head (thread = row: the group's 16 lanes of warps q = 0, 1)
        const float *SAp = reinterpret_cast<const float *>(smb + Map::SA + (2 * g + b) * Map::HEAD_BUF), *SHp = SAp + NR * 8;
        const int *sidx = reinterpret_cast<const int *>(SHp + 3 * NR);
        tc5::mbar_wait(bar_g + 8 * b, (uint32_t)((k >> 1) & 1));   // (long complete) the cp.async data of this tile is visible
        uint32_t v[8];
        MB6_LD8(v, quad_addr + OUTC);
        MB6_WAIT_LD();
        if ((lane >> 4) == g) {
          const int row = 16 * q + (lane & 15);
          const bool live = sidx[row] >= 0;
          float dout[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) dout[j] = 0.f;
          if (HEAD == 0) {
            float d[8], logp = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (j < O) {
                d[j] = SAp[row * O + j] - (__uint_as_float(v[j]) + bias[j]);
                logp += -(d[j] * d[j]) * (0.5f * bias[16 + j]) - LOG_SQRT_2PI - bias[8 + j];
              }
            const float Ai = SHp[NR + row], old = SHp[row];
            float dlogp = 0.f;
            if (live) {
              if (a.a2c) {
                s_obj += logp * Ai;
                dlogp = -a.lambda_p * a.inv_bg * Ai;
              } else {
                const float rt = expf(logp - old);
                const float lo = 1.f - a.eps_clip, hi = 1.f + a.eps_clip;
                const float x = rt * Ai, y = fminf(fmaxf(rt, lo), hi) * Ai;
                const bool firstb = !(y < x);  // min(x, y) keeps x on ties
                s_obj += firstb ? x : y;
                dlogp = firstb ? -a.lambda_p * a.inv_bg * x : 0.f;
                s_clip += (rt > hi || rt < lo) ? 1.f : 0.f;
              }
              s_kl += old - logp; s_adv += Ai; s_ret += SHp[2 * NR + row];
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (j < O) {
                const float ivar = bias[16 + j];
                dout[j] = dlogp * d[j] * ivar;
                dls[j] += dlogp * (d[j] * d[j] * ivar - 1.f);
              }
          } else {
            const float d = (__uint_as_float(v[0]) + bias[0]) - SHp[2 * NR + row];
            if (live) s_obj += d * d;
            dout[0] = live ? 2.f * d * a.inv_bg : 0.f;
          }
          // dOut -> fp32 rows [row][8] (E3) and DO^T [8][rows] hi/lo (B of dW3)
          float hi8[8], lo8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) { db3[j] += dout[j]; tc5::split(dout[j], hi8[j], lo8[j]); }
          float4 *p = reinterpret_cast<float4 *>(smb + Map::DO + g * Map::DO_G + row * 32);
          p[0] = make_float4(dout[0], dout[1], dout[2], dout[3]);
          p[1] = make_float4(dout[4], dout[5], dout[6], dout[7]);
          unsigned char *pt = smb + Map::DOT + g * Map::DOT_G + canon(0, row, NR);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            *reinterpret_cast<float *>(pt + 16 * j) = hi8[j];
            *reinterpret_cast<float *>(pt + Map::DOT_PLANE + 16 * j) = lo8[j];
          }
        }
      }
      asm volatile("bar.sync %0, %1;" ::"r"(NB_EPI + g), "n"(NTE) : "memory");   // dOut rows of this tile are in shared memory
      MB6_STAMP();   // head done
      {   // E3: dh2 = dOut W3 (K <= 8, FP32 pipe), dz2 = dh2 .* act'(h2): [row][feature] -> ACT (B of G5), [feature][rows] -> DZ^T (B of dW2)
        uint32_t hh[8], ll[8];
        MB6_LDX2(hh, atom_addr + ZB);
        MB6_LDX2(ll, atom_addr + H2L);
        float v[8];
        const float *dop = reinterpret_cast<const float *>(smb + Map::DO + g * Map::DO_G) + (c0 + rr) * 8;
#pragma unroll
        for (int rp = 0; rp < 4; ++rp) {   // this thread's 4 rows: c0 + 8 (rp >> 1) + rr + (rp & 1)  -> elements e = 4 (rp >> 1) + (rp & 1) (+ 2 for fa + 8)
          const float4 d0 = *reinterpret_cast<const float4 *>(dop + ((rp >> 1) * 8 + (rp & 1)) * 8);
          float xa = d0.x * w3a[0], xb = d0.x * w3b[0];
          if (HEAD == 0) {
            const float4 d1 = *reinterpret_cast<const float4 *>(dop + ((rp >> 1) * 8 + (rp & 1)) * 8 + 4);
            xa = fmaf(d0.y, w3a[1], xa); xb = fmaf(d0.y, w3b[1], xb);
            xa = fmaf(d0.z, w3a[2], xa); xb = fmaf(d0.z, w3b[2], xb);
            xa = fmaf(d0.w, w3a[3], xa); xb = fmaf(d0.w, w3b[3], xb);
            xa = fmaf(d1.x, w3a[4], xa); xb = fmaf(d1.x, w3b[4], xb);
            xa = fmaf(d1.y, w3a[5], xa); xb = fmaf(d1.y, w3b[5], xb);
            xa = fmaf(d1.z, w3a[6], xa); xb = fmaf(d1.z, w3b[6], xb);
            xa = fmaf(d1.w, w3a[7], xa); xb = fmaf(d1.w, w3b[7], xb);
          }
          v[4 * (rp >> 1) + (rp & 1)] = xa;
          v[4 * (rp >> 1) + (rp & 1) + 2] = xb;
        }
        MB6_WAIT_LD();
        float hi8[8], lo8[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float h2 = __uint_as_float(hh[e]) + __uint_as_float(ll[e]);
          const float dz = v[e] * dactf<ACT>(h2);
          if ((e >> 1) & 1) db2b += dz; else db2a += dz;
          tc5::split(dz, hi8[e], lo8[e]);
          *reinterpret_cast<float *>(act_a + MB6_ACT_OFF(e)) = hi8[e];
          *reinterpret_cast<float *>(act_a + Map::ACT_PLANE + MB6_ACT_OFF(e)) = lo8[e];
        }
#pragma unroll
        for (int pr = 0; pr < 4; ++pr) {   // pairs (e, e + 1) = rows (r, r + 1) of one feature: 8 contiguous bytes of a core-matrix row
          const int off = ((pr >> 1) * 256) + ((pr & 1) * 1024);   // rows + 8: two chunks along K (256 B); feature + 8: next 8-feature group
          *reinterpret_cast<float2 *>(dzt_a + off) = make_float2(hi8[2 * pr], hi8[2 * pr + 1]);
          *reinterpret_cast<float2 *>(dzt_a + Map::DZT_PLANE + off) = make_float2(lo8[2 * pr], lo8[2 * pr + 1]);
        }
      }
      MB6_READY();
      MB6_WAIT_MMA();
      MB6_STAMP();   // E3 + G5 (dh1) done
      {   // E4: dz1 = dh1 .* act'(h1): hi -> ZD, lo -> D1L (A of dW1)
        uint32_t v[8], hh[8], ll[8];
        MB6_LDX2(v, atom_addr + ZD);
        MB6_LDX2(hh, atom_addr + ZA);
        MB6_LDX2(ll, atom_addr + H1L);
        MB6_WAIT_LD();
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float h1 = __uint_as_float(hh[e]) + __uint_as_float(ll[e]);
          const float dz = __uint_as_float(v[e]) * dactf<ACT>(h1);
          if ((e >> 1) & 1) db1b += dz; else db1a += dz;
          float hi, lo;
          tc5::split(dz, hi, lo);
          v[e] = __float_as_uint(hi); ll[e] = __float_as_uint(lo);
        }
        MB6_STX2(atom_addr + ZD, v);
        tc5::mbar_wait(bar_dw3, (uint32_t)(k & 1));   // dW3 (12 MMAs, issued right behind G5) has read h2 lo: D1L shares its columns
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        MB6_STX2(atom_addr + D1L, ll);
        MB6_WAIT_ST();
      }
      MB6_READY();
      MB6_STAMP();   // E4 done
    }
    MB6_WAIT_MMA();   // the issuer's final commit: all weight-gradient MMAs of this group have completed
    MB6_STAMP();

    // ---------------- publish this GROUP's partial gradient (layout of fused_minibatch_kernel): partial row 2 blockIdx + g.
    // 32x32b accesses of the whole quarter; the lanes of this group's data-path half hold its accumulators (lane -> feature 16 q + lane % 16)
    // The partial is staged in this group's own (now dead) ACT and DZ^T buffers; after the CTA-wide barrier below all threads add the two
    // groups' vectors and store ONE coalesced row per CTA (half the rows for the reduction that follows, no strided global stores).
    float *stg_a = reinterpret_cast<float *>(smb + Map::ACT + g * Map::ACT_G), *stg_b = reinterpret_cast<float *>(smb + Map::DZT + g * Map::DZT_G) - STG_SPLIT;
#define out(idx) (*((idx) < STG_SPLIT ? stg_a + (idx) : stg_b + (idx)))
    const bool mine = (lane >> 4) == g;
    const int f = 16 * q + (lane & 15);
    if (!any) {   // this group had no tile (odd tile count): publish zeros
      for (int e = t - g * NTE; e < a.n_params + 16; e += NTE) out(e) = 0.f;
    } else {
      {   // dW2^T [i = f lanes][o columns]: this warp's half of the columns, 32 contiguous floats per thread
        uint32_t v[16], v2[16];
        MB6_LD16(v, quad_addr + DW2C + 32 * ch);
        MB6_LD16(v2, quad_addr + DW2C + 32 * ch + 16);
        MB6_WAIT_LD();
        if (mine) {
          const int base = off_W2(I) + f * H + 32 * ch;   // multiple of 4, like STG_SPLIT: a float4 never straddles the two staging pieces
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            *reinterpret_cast<float4 *>(&out(base + 4 * j4)) = make_float4(__uint_as_float(v[4 * j4]), __uint_as_float(v[4 * j4 + 1]), __uint_as_float(v[4 * j4 + 2]), __uint_as_float(v[4 * j4 + 3]));
            *reinterpret_cast<float4 *>(&out(base + 16 + 4 * j4)) = make_float4(__uint_as_float(v2[4 * j4]), __uint_as_float(v2[4 * j4 + 1]), __uint_as_float(v2[4 * j4 + 2]), __uint_as_float(v2[4 * j4 + 3]));
          }
        }
      }
      if (ch == 0) {   // dW1 [o = f lanes][i columns]
        uint32_t v[16], v2[8];
        MB6_LD16(v, quad_addr + DW1C);
        MB6_LD8(v2, quad_addr + DW1C + 16);
        MB6_WAIT_LD();
        if (mine) {
#pragma unroll
          for (int j = 0; j < 16; ++j) if (j < I) out(j * H + f) = __uint_as_float(v[j]);
#pragma unroll
          for (int j = 0; j < 8; ++j) if (16 + j < I) out((16 + j) * H + f) = __uint_as_float(v2[j]);
        }
      } else {         // dW3^T [k = f lanes][o columns]
        uint32_t v[8];
        MB6_LD8(v, quad_addr + DW3C);
        MB6_WAIT_LD();
        if (mine) {
#pragma unroll
          for (int j = 0; j < 8; ++j) if (j < O) out(off_W3(I) + f * O + j) = __uint_as_float(v[j]);
        }
      }
    }
    float *red = reinterpret_cast<float *>(smb + Map::RED + g * Map::RED_G);   // [2 ch][64] db1 | [2 ch][64] db2 | [HRED_ROWS][24] head
    constexpr int HRED_ROWS = HEAD == 0 ? 8 : 2;   // actor: one row per epilogue warp; critic: the two row-owning warps
    // a feature's rows are spread over the 4 lanes with the same lane / 4 (and over the two column halves ch)
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
      db1a += __shfl_xor_sync(0xffffffffu, db1a, o); db1b += __shfl_xor_sync(0xffffffffu, db1b, o);
      db2a += __shfl_xor_sync(0xffffffffu, db2a, o); db2b += __shfl_xor_sync(0xffffffffu, db2b, o);
    }
    if ((lane & 3) == 0) { red[ch * 64 + fa] = db1a; red[ch * 64 + fa + 8] = db1b; red[128 + ch * 64 + fa] = db2a; red[128 + ch * 64 + fa + 8] = db2b; }
    float *hred = red + 256;
    if (HEAD == 0) {
      // (row, j) threads: s_* live in the j = 0 lanes, dls[0] / db3[0] are this thread's action dimension j = lane & 7 -> per-warp rows of
      // hred [8 warps][24]: [0..5) sums, [8 + j] dlogΣ_j, [16 + j] db3_j; added over the 8 warps in a fixed order below
      const int we_ = (t - g * NTE) >> 5;
      float v;
      v = warp_sum(s_obj); if (lane == 0) hred[we_ * 24 + 0] = v;
      v = warp_sum(s_kl); if (lane == 0) hred[we_ * 24 + 1] = v;
      v = warp_sum(s_clip); if (lane == 0) hred[we_ * 24 + 2] = v;
      v = warp_sum(s_adv); if (lane == 0) hred[we_ * 24 + 3] = v;
      v = warp_sum(s_ret); if (lane == 0) hred[we_ * 24 + 4] = v;
      float dl = dls[0], d3 = db3[0];
      dl += __shfl_xor_sync(0xffffffffu, dl, 8); dl += __shfl_xor_sync(0xffffffffu, dl, 16);
      d3 += __shfl_xor_sync(0xffffffffu, d3, 8); d3 += __shfl_xor_sync(0xffffffffu, d3, 16);
      if (lane < 8) { hred[we_ * 24 + 8 + lane] = dl; hred[we_ * 24 + 16 + lane] = d3; }
    } else
    if (ch == 0 && q < 2) {   // head sums: this group's rows live in 16 lanes of warps q = 0, 1 (the other lanes hold zeros)
      float v;
      v = warp_sum(s_obj); if (lane == 0) hred[q * 24 + 0] = v;
      v = warp_sum(s_kl); if (lane == 0) hred[q * 24 + 1] = v;
      v = warp_sum(s_clip); if (lane == 0) hred[q * 24 + 2] = v;
      v = warp_sum(s_adv); if (lane == 0) hred[q * 24 + 3] = v;
      v = warp_sum(s_ret); if (lane == 0) hred[q * 24 + 4] = v;
#pragma unroll
      for (int j = 0; j < MAX_O; ++j) {
        v = warp_sum(dls[j]); if (lane == 0) hred[q * 24 + 8 + j] = v;
        v = warp_sum(db3[j]); if (lane == 0) hred[q * 24 + 16 + j] = v;
      }
    }
    asm volatile("bar.sync %0, %1;" ::"r"(NB_EPI + g), "n"(NTE) : "memory");
    const int te = t - g * NTE;
    if (te < 64) { const float v = red[te] + red[64 + te]; out(off_b1(I) + te) = v; if (a.nan_flag && v != v) atomicOr(a.nan_flag, 1); }
    else if (te < 128) { const int j = te - 64; const float v = red[128 + j] + red[192 + j]; out(off_b2(I) + j) = v; if (a.nan_flag && v != v) atomicOr(a.nan_flag, 1); }
    else if (te < 128 + O) {
      const int o = te - 128;
      float v = 0.f;
#pragma unroll
      for (int r_ = 0; r_ < HRED_ROWS; ++r_) v += hred[r_ * 24 + 16 + o];
      out(off_b3(I, O) + o) = v;
    }
    else if (te >= 160 && te < 176) {
      // tail layout: [n_params .. +8) = dlogΣ, [n_params+8 .. +16) = obj, kl, clip, adv, ret, 0, 0, 0
      const int kk = te - 160, src = kk < 8 ? 8 + kk : kk - 8;
      float v = 0.f;
      if (src < 5 || (src >= 8 && src < 16)) {
#pragma unroll
        for (int r_ = 0; r_ < HRED_ROWS; ++r_) v += hred[r_ * 24 + src];
      }
      if (HEAD == 0 && kk == 14 && blockIdx.x == 0 && g == 0) {   // sum(logΣ) as this kernel saw it: the entropy of the info record (policies.jl:348)
        v = 0.f;
        for (int j = 0; j < O; ++j) v += bias[8 + j];
      }
      out(a.n_params + kk) = v;
      if (a.nan_flag && v != v) atomicOr(a.nan_flag, 1);
    }
    MB6_STAMP();   // partial gradient published
#undef MB6_READY
#undef MB6_WAIT_MMA
#undef MB6_HIDDEN_EPILOGUE
#undef MB6_ACT_OFF
#undef out
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  {   // group 0's vector + group 1's vector -> this CTA's partial row
    const float *a0 = reinterpret_cast<const float *>(smb + Map::ACT), *b0 = reinterpret_cast<const float *>(smb + Map::DZT) - STG_SPLIT;
    const float *a1 = reinterpret_cast<const float *>(smb + Map::ACT + Map::ACT_G), *b1 = reinterpret_cast<const float *>(smb + Map::DZT + Map::DZT_G) - STG_SPLIT;
    float *row = a.partials + (int64_t)blockIdx.x * a.pstride;
    for (int e = t; e < a.n_params + 16; e += NTH) row[e] = (e < STG_SPLIT ? a0[e] : b0[e]) + (e < STG_SPLIT ? a1[e] : b1[e]);
  }
  if (a.trace && t == 0) { unsigned long long gt_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_)); a.trace[2 * blockIdx.x + 1] = gt_; }
  if (w == W_ISSUE) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
#undef MB6_STAMP
#undef MB6_ISTAMP
}

#undef MB6_LD16
#undef MB6_ST16
#undef MB6_ST8
#undef MB6_LD8
#undef MB6_LDX2
#undef MB6_STX2
#undef MB6_WAIT_LD
#undef MB6_WAIT_ST

}  // namespace mb6
