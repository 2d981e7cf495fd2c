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
//   G4  ZC  = W3^T dOut^T  ||  dW3^T += h2^T dOut       E3  dz2 = ZC .* act'(h2): ACT (B of G5), DZ^T [feature][rows] (B of dW2); db2
//   G5  ZD  = W2^T dz2     ||  dW2^T += h1^T dz2        E4  dz1 = ZD .* act'(h1): hi -> ZD, lo -> D1L (A of dW1); db1
//   G6  dW1 += dz1^T S                                   (completion awaited at the top of the next tile)
// 3xTF32 everywhere (lo*hi + hi*lo + hi*hi, fp32-level accuracy), same partial-gradient layout and head arithmetic as
// fused_minibatch_tc_kernel.  Included by ppo_fused.cu after mb_tc5.cuh.
#pragma once

namespace mb6 {

constexpr int TRW = 64, NR = 32, NTH = 256, KX = tc5::KX;
constexpr int TMEM_COLS = 512;
constexpr uint32_t ZA = 0, H1L = 32, ZB = 64, H2L = 96, ZC = 128, ZD = 160, D1L = 192, OUTC = 224, DW2C = 256, DW1C = 320, DW3C = 352;
constexpr int ACT_LBO = 144, ACT_SBO = 2320;   // padded K-major tile: the [row][feature] stores of a warp hit 32 distinct banks

struct Map {   // bytes; every operand is a pair of planes hi | lo
  static constexpr int W1A = 0;                                  // A [64 j][24 i]
  static constexpr int W2A = W1A + 2 * 64 * KX * 4;              // A [64 j][64 k]   = W2(j,k)
  static constexpr int W2TA = W2A + 2 * 64 * 64 * 4;             // A [64 k][64 j]   = W2(j,k)
  static constexpr int W3B = W2TA + 2 * 64 * 64 * 4;             // B [8 o][64 k]    = W3(o,k)
  static constexpr int W3TA = W3B + 2 * 8 * 64 * 4;              // A [64 k][8 o]    = W3(o,k)
  static constexpr int S = W3TA + 2 * 64 * 8 * 4;                // B [64 rows][24]
  static constexpr int S_PLANE = TRW * KX * 4;
  static constexpr int ST = S + 2 * S_PLANE;                     // B per atom [24 i][32 rows]; plane = 2 atoms x 3072
  static constexpr int ST_ATOM = 3 * 1024, ST_PLANE = 2 * ST_ATOM;
  static constexpr int ACT = ST + 2 * ST_PLANE;                  // [64 rows][64] padded (prologue: raw parameters)
  static constexpr int ACT_PLANE = 8 * ACT_SBO;
  static constexpr int DZT = ACT + 2 * ACT_PLANE;                // B per atom [64 o][32 rows]
  static constexpr int DZT_ATOM = 8 * 1024, DZT_PLANE = 2 * DZT_ATOM;
  static constexpr int DO = DZT + 2 * DZT_PLANE;                 // B [64 rows][8]
  static constexpr int DO_PLANE = 8 * 256;
  static constexpr int DOT = DO + 2 * DO_PLANE;                  // B per atom [8 o][32 rows]
  static constexpr int DOT_ATOM = 1024, DOT_PLANE = 2 * DOT_ATOM;
  static constexpr int SX = DOT + 2 * DOT_PLANE;                 // gather staging: x [64][I <= 24]
  static constexpr int SA = SX + TRW * KX * 4;                   //                 actions [64][O <= 8]
  static constexpr int SH = SA + TRW * 8 * 4;                    //                 logprob | advantage | return [64] each
  static constexpr int IDX = SH + 3 * TRW * 4;                   // [64] ints
  static constexpr int BIAS = IDX + TRW * 4;                     // b3[8] | logΣ[8] | σ²[8]
  static constexpr int RED = BIAS + 24 * 4;                      // [2][64] bias-gradient scratch, [4][24] head scratch
  static constexpr int BAR = RED + (2 * 64 * 2 + 4 * 24) * 4;
  static constexpr int TOTAL = BAR + 32;
  static_assert(ACT % 16 == 0 && DZT % 16 == 0 && DO % 16 == 0 && SX % 16 == 0 && BAR % 8 == 0, "alignment");
  static_assert(2 * ACT_PLANE >= P_SMEM * 4, "the raw parameters are staged in the ACT planes");
};

__device__ __forceinline__ int canon(int row, int k, int K) { return (row >> 3) * (32 * K) + (k >> 2) * 128 + (row & 7) * 16 + (k & 3) * 4; }
__device__ __forceinline__ int canon_act(int row, int k) { return (row >> 3) * ACT_SBO + (k >> 2) * ACT_LBO + (row & 7) * 16 + (k & 3) * 4; }
using tc5::elect_one;
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
#define MB6_LD8(v, taddr)                                                                                                                                \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"                                                            \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])                                           \
               : "r"(taddr)                                                                                                                               \
               : "memory")
#define MB6_WAIT_LD() asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")
#define MB6_WAIT_ST() asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory")
// generic-proxy shared-memory stores + tcgen05.st results -> visible to the MMAs the elected thread issues after the barrier
#define MB6_HANDOFF()                                                   \
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          \
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");      \
  __syncthreads()
#define MB6_WAIT_MMA()                                                  \
  tc5::mbar_wait(bar_mma, ph); ph ^= 1;                                 \
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory")

template <int HEAD>
__global__ void __launch_bounds__(NTH, 1) minibatch_kernel(MbArgs a) {
  extern __shared__ __align__(1024) unsigned char smb[];
  const NetDesc nd = a.net;
  const int I = nd.I, O = nd.O, act = nd.act;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int q = w & 3, ch = w >> 2, atom = lane >> 4, f = 16 * q + (lane & 15);   // epilogue owner of feature f, atom rows [16ch, 16ch + 16)
  const int stop_at = a.ctl ? a.ctl[1] : 0;
  const uint32_t bar_mma = smem_u32(smb + Map::BAR), bar_par = smem_u32(smb + Map::BAR + 8);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smb + Map::BAR + 16);
  float *bias = reinterpret_cast<float *>(smb + Map::BIAS);   // b3[8] | logΣ[8] | σ²[8]
  float *SXp = reinterpret_cast<float *>(smb + Map::SX), *SAp = reinterpret_cast<float *>(smb + Map::SA), *SHp = reinterpret_cast<float *>(smb + Map::SH);
  int *sidx = reinterpret_cast<int *>(smb + Map::IDX);
  const int64_t n_tiles = (a.bm + TRW - 1) / TRW;
  const uint32_t inv_I = (65536u + (uint32_t)I - 1u) / (uint32_t)I, inv_O = (65536u + (uint32_t)O - 1u) / (uint32_t)O;

  auto tile_row = [&](int64_t tile, int r) -> int {
    const int64_t row = tile * TRW + r;
    return (tile < n_tiles && row < a.bm) ? (a.order ? a.order[row] : (int)row) : -1;
  };
  auto issue_gather = [&]() {   // rows listed in sidx -> staging (cp.async, zero fill for padding rows)
    for (int e = t; e < TRW * I; e += NTH) {
      const int r = (int)(((uint32_t)e * inv_I) >> 16), i = e - r * I;
      const int row = sidx[r];
      cp_async4(SXp + e, a.s + (row >= 0 ? (int64_t)row * I + i : 0), row >= 0);
    }
    if (HEAD == 0)
      for (int e = t; e < TRW * O; e += NTH) {
        const int r = (int)(((uint32_t)e * inv_O) >> 16), o = e - r * O;
        const int row = sidx[r];
        cp_async4(SAp + e, a.act + (row >= 0 ? (int64_t)row * O + o : 0), row >= 0);
      }
    if (t < TRW) {
      const int row = sidx[t];
      if (HEAD == 0) {
        cp_async4(SHp + t, a.logp_old + (row >= 0 ? row : 0), row >= 0);
        cp_async4(SHp + TRW + t, a.adv + (row >= 0 ? row : 0), row >= 0);
      }
      const bool has_ret = a.ret != nullptr;
      cp_async4(SHp + 2 * TRW + t, has_ret ? a.ret + (row >= 0 ? row : 0) : a.s, has_ret && row >= 0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  // ---- prologue: barriers, first gather, raw parameters (TMA bulk) -> ACT planes, TMEM, weight planes
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_mma), "r"(1) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_par), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (t < TRW) sidx[t] = tile_row(blockIdx.x, t);
  for (int e = t; e < (2 * Map::S_PLANE + 2 * Map::ST_PLANE) / 4; e += NTH) reinterpret_cast<float *>(smb + Map::S)[e] = 0.f;   // K padding stays zero
  for (int e = t; e < (2 * Map::DO_PLANE + 2 * Map::DOT_PLANE) / 4; e += NTH) reinterpret_cast<float *>(smb + Map::DO)[e] = 0.f;
  __syncthreads();
  issue_gather();
  if (t == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_par), "r"(nd.bytes16) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smb + Map::ACT)), "l"(nd.params),
                 "r"(nd.bytes16), "r"(bar_par)
                 : "memory");
  }
  if (w == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc5::mbar_wait(bar_par, 0);
  float b1f, b2f;
  {
    const float *P = reinterpret_cast<const float *>(smb + Map::ACT);
    const float *W1 = P, *b1 = P + off_b1(I), *W2 = P + off_W2(I), *b2 = P + off_b2(I), *W3 = P + off_W3(I), *b3 = P + off_b3(I, O);
    auto put = [&](int base, int plane_bytes, int off, float v) {
      float hi, lo;
      tc5::split(v, hi, lo);
      *reinterpret_cast<float *>(smb + base + off) = hi;
      *reinterpret_cast<float *>(smb + base + plane_bytes + off) = lo;
    };
    for (int e = t; e < 64 * KX; e += NTH) { const int j = e / KX, i = e - j * KX; put(Map::W1A, 64 * KX * 4, canon(j, i, KX), i < I ? W1[i * H + j] : 0.f); }
    for (int e = t; e < 64 * 64; e += NTH) {
      const int k = e >> 6, j = e & 63;   // flat W2[k*64 + j] = W2(out j, in k)
      put(Map::W2A, 64 * 64 * 4, canon(j, k, 64), W2[e]);
      put(Map::W2TA, 64 * 64 * 4, canon(k, j, 64), W2[e]);
    }
    for (int e = t; e < 8 * 64; e += NTH) {
      const int o = e >> 6, k = e & 63;
      const float v = o < O ? W3[k * O + o] : 0.f;
      put(Map::W3B, 8 * 64 * 4, canon(o, k, 64), v);
      put(Map::W3TA, 64 * 8 * 4, canon(k, o, 8), v);
    }
    b1f = b1[f]; b2f = b2[f];
    if (t < 8) bias[t] = t < O ? b3[t] : 0.f;
    if (HEAD == 0 && t < 8) { const float ls = t < O ? a.ls[t] : 0.f, sg = expf(ls); bias[8 + t] = ls; bias[16 + t] = sg * sg; }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  if (stop_at != 0 && stop_at <= a.mb) {   // an EARLIER minibatch raised the KL stop flag (rl/ppo.jl:59): nothing to do
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
    return;
  }
  // operand descriptors (K-major canonical, no swizzle): dense tiles LBO = 128, SBO = 32 K; the ACT tile is padded
  const uint32_t sb = smem_u32(smb);
  const uint64_t dW1A = tc5::make_desc(sb + Map::W1A, 128, 32 * KX), dW1A_lo = tc5::make_desc(sb + Map::W1A + 64 * KX * 4, 128, 32 * KX);
  const uint64_t dW2A = tc5::make_desc(sb + Map::W2A, 128, 32 * 64), dW2A_lo = tc5::make_desc(sb + Map::W2A + 64 * 64 * 4, 128, 32 * 64);
  const uint64_t dW2TA = tc5::make_desc(sb + Map::W2TA, 128, 32 * 64), dW2TA_lo = tc5::make_desc(sb + Map::W2TA + 64 * 64 * 4, 128, 32 * 64);
  const uint64_t dW3B = tc5::make_desc(sb + Map::W3B, 128, 32 * 64), dW3B_lo = tc5::make_desc(sb + Map::W3B + 8 * 64 * 4, 128, 32 * 64);
  const uint64_t dW3TA = tc5::make_desc(sb + Map::W3TA, 128, 32 * 8), dW3TA_lo = tc5::make_desc(sb + Map::W3TA + 64 * 8 * 4, 128, 32 * 8);
  const uint64_t dS = tc5::make_desc(sb + Map::S, 128, 32 * KX), dS_lo = tc5::make_desc(sb + Map::S + Map::S_PLANE, 128, 32 * KX);
  const uint64_t dST = tc5::make_desc(sb + Map::ST, 128, 32 * NR), dST_lo = tc5::make_desc(sb + Map::ST + Map::ST_PLANE, 128, 32 * NR);
  const uint64_t dACT = tc5::make_desc(sb + Map::ACT, ACT_LBO, ACT_SBO), dACT_lo = tc5::make_desc(sb + Map::ACT + Map::ACT_PLANE, ACT_LBO, ACT_SBO);
  const uint64_t dDZT = tc5::make_desc(sb + Map::DZT, 128, 32 * NR), dDZT_lo = tc5::make_desc(sb + Map::DZT + Map::DZT_PLANE, 128, 32 * NR);
  const uint64_t dDO = tc5::make_desc(sb + Map::DO, 128, 32 * 8), dDO_lo = tc5::make_desc(sb + Map::DO + Map::DO_PLANE, 128, 32 * 8);
  const uint64_t dDOT = tc5::make_desc(sb + Map::DOT, 128, 32 * NR), dDOT_lo = tc5::make_desc(sb + Map::DOT + Map::DOT_PLANE, 128, 32 * NR);
  constexpr uint32_t ADV = 16, ADV_ACT = (2 * ACT_LBO) >> 4;           // one k-step = two core matrices along K
  constexpr uint32_t S_ATOM = (4 * 32 * KX) >> 4, ACT_ATOM = (4 * ACT_SBO) >> 4, DO_ATOM = (4 * 32 * 8) >> 4;   // rows 32.. of a [64 rows][K] tile
  const uint32_t id32 = tc5::make_idesc(64, 32), id8 = tc5::make_idesc(64, 8), id64 = tc5::make_idesc(64, 64), id24 = tc5::make_idesc(64, 24);
  const uint32_t lane_addr = tmem + ((uint32_t)(32 * q) << 16);
  const uint32_t c0 = 16 * ch;                    // this thread's columns (rows of its atom) inside a 32-column block
  uint32_t ph = 0;
  uint32_t first = 0;                             // 0 on this CTA's first tile: the dW accumulators are overwritten, then accumulated
  bool pending = false;                           // dW1 of the previous tile not yet awaited

  float db1 = 0.f, db2 = 0.f;                     // bias gradients of feature f over this thread's rows
  float s_obj = 0.f, s_kl = 0.f, s_clip = 0.f, s_adv = 0.f, s_ret = 0.f, dls[MAX_O], db3[MAX_O];
#pragma unroll
  for (int j = 0; j < MAX_O; ++j) { dls[j] = 0.f; db3[j] = 0.f; }

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    asm volatile("cp.async.wait_all;" ::: "memory");
    if (pending) { MB6_WAIT_MMA(); }   // dW1 of the previous tile has read S^T
    __syncthreads();                   // staged rows visible
    const int nidx = t < TRW ? tile_row(tile + gridDim.x, t) : -1;
    // ---------------- P0: staged x rows -> S (B of L1) and S^T (B of dW1), hi/lo
    for (int e = t; e < TRW * I; e += NTH) {
      const int r = (int)(((uint32_t)e * inv_I) >> 16), i = e - r * I;
      float hi, lo;
      tc5::split(SXp[e], hi, lo);
      const int o1 = canon(r, i, KX), o2 = (r >> 5) * Map::ST_ATOM + canon(i, r & 31, NR);
      *reinterpret_cast<float *>(smb + Map::S + o1) = hi;
      *reinterpret_cast<float *>(smb + Map::S + Map::S_PLANE + o1) = lo;
      *reinterpret_cast<float *>(smb + Map::ST + o2) = hi;
      *reinterpret_cast<float *>(smb + Map::ST + Map::ST_PLANE + o2) = lo;
    }
    MB6_HANDOFF();
    // ---------------- L1: ZA[atom] = W1 S^T
    if (w == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
#pragma unroll
        for (uint32_t at = 0; at < 2; ++at)
          gemm_ss<KX / 8>(tmem + ((16u * at) << 16) + ZA, dW1A, dW1A_lo, ADV, dS + at * S_ATOM, dS_lo + at * S_ATOM, ADV, id32, 0u);
        commit(bar_mma);
      }
      __syncwarp();
    }
    MB6_WAIT_MMA();
#define MB6_HIDDEN_EPILOGUE(ZCOL, LCOL, BIASF)                                                                           \
    {                                                                                                                    \
      uint32_t v[16], l[16];                                                                                             \
      MB6_LD16(v, lane_addr + ZCOL + c0);                                                                                \
      MB6_WAIT_LD();                                                                                                     \
      unsigned char *ap = smb + Map::ACT + canon_act(32 * atom + (int)c0, f);                                           \
      _Pragma("unroll") for (int j = 0; j < 16; ++j) {                                                                   \
        float hi, lo;                                                                                                    \
        tc5::split(act_fused(act, __uint_as_float(v[j]) + BIASF), hi, lo);                                               \
        v[j] = __float_as_uint(hi); l[j] = __float_as_uint(lo);                                                          \
        const int ro = (j >> 3) * ACT_SBO + (j & 7) * 16;   /* c0 is a multiple of 16: row = 32 atom + c0 + j */        \
        *reinterpret_cast<float *>(ap + ro) = hi;                                                                        \
        *reinterpret_cast<float *>(ap + Map::ACT_PLANE + ro) = lo;                                                       \
      }                                                                                                                  \
      MB6_ST16(lane_addr + ZCOL + c0, v);                                                                                \
      MB6_ST16(lane_addr + LCOL + c0, l);                                                                                \
      MB6_WAIT_ST();                                                                                                     \
    }
    MB6_HIDDEN_EPILOGUE(ZA, H1L, b1f)
    MB6_HANDOFF();
    // ---------------- L2: ZB[atom] = W2 h1
    if (w == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
#pragma unroll
        for (uint32_t at = 0; at < 2; ++at)
          gemm_ss<8>(tmem + ((16u * at) << 16) + ZB, dW2A, dW2A_lo, ADV, dACT + at * ACT_ATOM, dACT_lo + at * ACT_ATOM, ADV_ACT, id32, 0u);
        commit(bar_mma);
      }
      __syncwarp();
    }
    MB6_WAIT_MMA();
    MB6_HIDDEN_EPILOGUE(ZB, H2L, b2f)
    MB6_HANDOFF();
    // ---------------- L3 (rows on the lanes): OUT[64 rows][8] = h2 W3^T, then the loss head (thread = row: 16 lanes of warps 0..3)
    if (w == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        gemm_ss<8>(tmem + OUTC, dACT, dACT_lo, ADV_ACT, dW3B, dW3B_lo, ADV, id8, 0u);
        commit(bar_mma);
      }
      __syncwarp();
    }
    MB6_WAIT_MMA();
    if (w < 4) {
      uint32_t v[8];
      MB6_LD8(v, lane_addr + OUTC);
      MB6_WAIT_LD();
      if (lane < 16) {
        const int row = 16 * q + lane;
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
              logp += -(d[j] * d[j]) / (2.f * bias[16 + j]) - LOG_SQRT_2PI - bias[8 + j];
            }
          const float Ai = SHp[TRW + row], old = SHp[row];
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
            s_kl += old - logp; s_adv += Ai; s_ret += SHp[2 * TRW + row];
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (j < O) {
              const float var = bias[16 + j];
              dout[j] = dlogp * d[j] / var;
              dls[j] += dlogp * (d[j] * d[j] / var - 1.f);
            }
        } else {
          const float d = (__uint_as_float(v[0]) + bias[0]) - SHp[2 * TRW + row];
          if (live) s_obj += d * d;
          dout[0] = live ? 2.f * d * a.inv_bg : 0.f;
        }
        // dOut -> DO [row][8] (B of G4) and DO^T [atom][8][rows] (B of dW3), hi/lo
        float hi8[8], lo8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { db3[j] += dout[j]; tc5::split(dout[j], hi8[j], lo8[j]); }
        unsigned char *p = smb + Map::DO + canon(row, 0, 8);
        *reinterpret_cast<float4 *>(p) = make_float4(hi8[0], hi8[1], hi8[2], hi8[3]);
        *reinterpret_cast<float4 *>(p + 128) = make_float4(hi8[4], hi8[5], hi8[6], hi8[7]);
        *reinterpret_cast<float4 *>(p + Map::DO_PLANE) = make_float4(lo8[0], lo8[1], lo8[2], lo8[3]);
        *reinterpret_cast<float4 *>(p + Map::DO_PLANE + 128) = make_float4(lo8[4], lo8[5], lo8[6], lo8[7]);
        unsigned char *pt = smb + Map::DOT + (row >> 5) * Map::DOT_ATOM + canon(0, row & 31, NR);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          *reinterpret_cast<float *>(pt + 16 * j) = hi8[j];
          *reinterpret_cast<float *>(pt + Map::DOT_PLANE + 16 * j) = lo8[j];
        }
      }
    }
    MB6_HANDOFF();
    // the staging area has been consumed: stream in this CTA's next tile while the backward half runs
    if (t < TRW) sidx[t] = nidx;
    // ---------------- G4: ZC[atom] = W3^T dOut^T  ||  dW3^T[atom] += h2^T dOut
    if (w == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
#pragma unroll
        for (uint32_t at = 0; at < 2; ++at) {
          const uint32_t dp = tmem + ((16u * at) << 16);
          gemm_ss<1>(dp + ZC, dW3TA, dW3TA_lo, ADV, dDO + at * DO_ATOM, dDO_lo + at * DO_ATOM, ADV, id32, 0u);
          gemm_ts<NR / 8>(dp + DW3C, dp + ZB, dp + H2L, dDOT + at * (Map::DOT_ATOM >> 4), dDOT_lo + at * (Map::DOT_ATOM >> 4), ADV, id8, first);
        }
        commit(bar_mma);
      }
      __syncwarp();
    }
    __syncthreads();   // sidx of the next tile is complete
    if (tile + gridDim.x < n_tiles) issue_gather();
    MB6_WAIT_MMA();
    {   // E3: dz2 = dh2 .* act'(h2)
      uint32_t v[16], hh[16], ll[16];
      MB6_LD16(v, lane_addr + ZC + c0);
      MB6_LD16(hh, lane_addr + ZB + c0);
      MB6_LD16(ll, lane_addr + H2L + c0);
      MB6_WAIT_LD();
      unsigned char *ap = smb + Map::ACT + canon_act(32 * atom + (int)c0, f);
      unsigned char *zp = smb + Map::DZT + atom * Map::DZT_ATOM + canon(f, (int)c0, NR);
      float hi16[16], lo16[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float h2 = __uint_as_float(hh[j]) + __uint_as_float(ll[j]);
        const float dz = __uint_as_float(v[j]) * act_bwd_from_out(act, h2);
        db2 += dz;
        tc5::split(dz, hi16[j], lo16[j]);
        const int ro = (j >> 3) * ACT_SBO + (j & 7) * 16;
        *reinterpret_cast<float *>(ap + ro) = hi16[j];
        *reinterpret_cast<float *>(ap + Map::ACT_PLANE + ro) = lo16[j];
      }
#pragma unroll
      for (int g4 = 0; g4 < 4; ++g4) {   // 4 consecutive rows = one 16-byte chunk of a core-matrix row; next chunk along K is 128 B away
        *reinterpret_cast<float4 *>(zp + 128 * g4) = make_float4(hi16[4 * g4], hi16[4 * g4 + 1], hi16[4 * g4 + 2], hi16[4 * g4 + 3]);
        *reinterpret_cast<float4 *>(zp + Map::DZT_PLANE + 128 * g4) = make_float4(lo16[4 * g4], lo16[4 * g4 + 1], lo16[4 * g4 + 2], lo16[4 * g4 + 3]);
      }
    }
    MB6_HANDOFF();
    // ---------------- G5: ZD[atom] = W2^T dz2  ||  dW2^T[atom] += h1^T dz2
    if (w == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
#pragma unroll
        for (uint32_t at = 0; at < 2; ++at) {
          const uint32_t dp = tmem + ((16u * at) << 16);
          gemm_ss<8>(dp + ZD, dW2TA, dW2TA_lo, ADV, dACT + at * ACT_ATOM, dACT_lo + at * ACT_ATOM, ADV_ACT, id32, 0u);
          gemm_ts<NR / 8>(dp + DW2C, dp + ZA, dp + H1L, dDZT + at * (Map::DZT_ATOM >> 4), dDZT_lo + at * (Map::DZT_ATOM >> 4), ADV, id64, first);
        }
        commit(bar_mma);
      }
      __syncwarp();
    }
    MB6_WAIT_MMA();
    {   // E4: dz1 = dh1 .* act'(h1): hi -> ZD, lo -> D1L (A of dW1)
      uint32_t v[16], hh[16], ll[16];
      MB6_LD16(v, lane_addr + ZD + c0);
      MB6_LD16(hh, lane_addr + ZA + c0);
      MB6_LD16(ll, lane_addr + H1L + c0);
      MB6_WAIT_LD();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float h1 = __uint_as_float(hh[j]) + __uint_as_float(ll[j]);
        const float dz = __uint_as_float(v[j]) * act_bwd_from_out(act, h1);
        db1 += dz;
        float hi, lo;
        tc5::split(dz, hi, lo);
        v[j] = __float_as_uint(hi); ll[j] = __float_as_uint(lo);
      }
      MB6_ST16(lane_addr + ZD + c0, v);
      MB6_ST16(lane_addr + D1L + c0, ll);
      MB6_WAIT_ST();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    // ---------------- G6: dW1[atom] += dz1^T S   (awaited at the top of the next tile / after the loop)
    if (w == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
#pragma unroll
        for (uint32_t at = 0; at < 2; ++at) {
          const uint32_t dp = tmem + ((16u * at) << 16);
          gemm_ts<NR / 8>(dp + DW1C, dp + ZD, dp + D1L, dST + at * (Map::ST_ATOM >> 4), dST_lo + at * (Map::ST_ATOM >> 4), ADV, id24, first);
        }
        commit(bar_mma);
      }
      __syncwarp();
    }
    pending = true;
    first = 1u;
  }
  if (pending) { MB6_WAIT_MMA(); }

  // ---------------- publish this CTA's partial gradient (layout of fused_minibatch_kernel); the two atoms' accumulators are summed here
  float *out = a.partials + (int64_t)blockIdx.x * a.pstride;
  if (first == 0u) {   // no tile (cannot happen with grid <= n_tiles, kept for safety): publish zeros
    for (int e = t; e < a.n_params + 16; e += NTH) out[e] = 0.f;
  } else {
    {   // dW2^T [i = f lanes][o columns]: this warp's half of the columns
      uint32_t v[16], v2[16];
      MB6_LD16(v, lane_addr + DW2C + 32 * ch);
      MB6_LD16(v2, lane_addr + DW2C + 32 * ch + 16);
      MB6_WAIT_LD();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float x = __uint_as_float(v[j]), y = __uint_as_float(v2[j]);
        x += __shfl_xor_sync(0xffffffffu, x, 16);
        y += __shfl_xor_sync(0xffffffffu, y, 16);
        if (lane < 16) { out[off_W2(I) + f * H + 32 * ch + j] = x; out[off_W2(I) + f * H + 32 * ch + 16 + j] = y; }
      }
    }
    if (ch == 0) {   // dW1 [o = f lanes][i columns]
      uint32_t v[16], v2[8];
      MB6_LD16(v, lane_addr + DW1C);
      MB6_LD8(v2, lane_addr + DW1C + 16);
      MB6_WAIT_LD();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float x = __uint_as_float(v[j]);
        x += __shfl_xor_sync(0xffffffffu, x, 16);
        if (lane < 16 && j < I) out[j * H + f] = x;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float x = __uint_as_float(v2[j]);
        x += __shfl_xor_sync(0xffffffffu, x, 16);
        if (lane < 16 && 16 + j < I) out[(16 + j) * H + f] = x;
      }
    } else {         // dW3^T [k = f lanes][o columns]
      uint32_t v[8];
      MB6_LD8(v, lane_addr + DW3C);
      MB6_WAIT_LD();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float x = __uint_as_float(v[j]);
        x += __shfl_xor_sync(0xffffffffu, x, 16);
        if (lane < 16 && j < O) out[off_W3(I) + f * O + j] = x;
      }
    }
  }
  float *red = reinterpret_cast<float *>(smb + Map::RED);   // [2 ch][64] db1 | [2 ch][64] db2 | [4][24] head
  db1 += __shfl_xor_sync(0xffffffffu, db1, 16);
  db2 += __shfl_xor_sync(0xffffffffu, db2, 16);
  if (lane < 16) { red[ch * 64 + f] = db1; red[128 + ch * 64 + f] = db2; }
  float *hred = red + 256;
  if (w < 4) {   // head sums: rows live in lanes 0..15 of warps 0..3 (the other lanes hold zeros)
    float v;
    v = warp_sum(s_obj); if (lane == 0) hred[w * 24 + 0] = v;
    v = warp_sum(s_kl); if (lane == 0) hred[w * 24 + 1] = v;
    v = warp_sum(s_clip); if (lane == 0) hred[w * 24 + 2] = v;
    v = warp_sum(s_adv); if (lane == 0) hred[w * 24 + 3] = v;
    v = warp_sum(s_ret); if (lane == 0) hred[w * 24 + 4] = v;
#pragma unroll
    for (int j = 0; j < MAX_O; ++j) {
      v = warp_sum(dls[j]); if (lane == 0) hred[w * 24 + 8 + j] = v;
      v = warp_sum(db3[j]); if (lane == 0) hred[w * 24 + 16 + j] = v;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (t < 64) out[off_b1(I) + t] = red[t] + red[64 + t];
  else if (t < 128) out[off_b2(I) + (t - 64)] = red[128 + (t - 64)] + red[192 + (t - 64)];
  else if (t < 128 + O) { const int o = t - 128; out[off_b3(I, O) + o] = (hred[16 + o] + hred[24 + 16 + o]) + (hred[48 + 16 + o] + hred[72 + 16 + o]); }
  else if (t >= 160 && t < 176) {
    // tail layout: [n_params .. +8) = dlogΣ, [n_params+8 .. +16) = obj, kl, clip, adv, ret, 0, 0, 0
    const int k = t - 160, src = k < 8 ? 8 + k : k - 8;
    float v = 0.f;
    if (src < 5 || (src >= 8 && src < 16))
#pragma unroll
      for (int ww = 0; ww < 4; ++ww) v += hred[ww * 24 + src];
    out[a.n_params + k] = v;
  }
  if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
#undef MB6_HIDDEN_EPILOGUE
}

#undef MB6_LD16
#undef MB6_ST16
#undef MB6_LD8
#undef MB6_WAIT_LD
#undef MB6_WAIT_ST
#undef MB6_HANDOFF
#undef MB6_WAIT_MMA

}  // namespace mb6
