// The PPO minibatch kernel with the row GEMMs on tcgen05 and the activations resident in TENSOR MEMORY.
//
// Same contract as fused_minibatch_tc_kernel (same partial-gradient layout, same head arithmetic).  A CTA (512 threads, one per SM)
// owns 128-row tiles; thread (row, cg) = (TMEM lane, 16-column group) is the epilogue owner of 16 features of one row.
//
//   GEMM (tcgen05.mma kind::tf32, 3xTF32: lo*hi + hi*lo + hi*hi; A from TMEM, B = canonical K-major hi/lo planes in shared memory)
//     G0  z1  [128 x 64] = x   W1      -> C1      E1: h1 = act(z1 + b1): hi -> C1, lo -> C2, fp32 transposed copy -> h1^T (smem)
//     G1  z2  [128 x 64] = h1  W2      -> C3      E2: h2: hi -> C3, lo -> C2, h2^T
//     G2  out [128 x 16] = h2  W3      -> C1      head (thread = row): loss terms, dL/dout -> dOut^T (smem)
//         dh2 = dOut W3^T has K <= 8: plain FFMA in E4 (cheaper than a tensor-core round trip)
//                                                 E4: dz2 = dh2 .* act'(h2)  (h2 = C3 + C2): hi -> C1, lo -> C2, dz2^T over h2^T
//     G4  dh1 [128 x 64] = dz2 W2^T    -> C3      E5: dz1 = dh1 .* act'(h1)  (h1 from h1^T): dz1^T over h1^T
//   The weight-gradient GEMMs contract over ROWS; tf32 operands cannot be transposed by the tcgen05 descriptors without the
//   128B/32B-base swizzle, so they run as warp-level MMAs (mma.sync 3xTF32, 16 warps) on the fp32 transposed tiles the epilogues
//   leave in shared memory -- CONCURRENTLY with the asynchronous tcgen05 GEMM of the same phase:
//     dW2 += h1^T dz2  (during G4)      dW3 += h2^T dOut (before E4)      dW1 += x^T dz1  (after E5)
//   TMEM columns (256 of 512): [0,24) x hi | [24,48) x lo | C1 [64,128) | C2 [128,192) | C3 [192,256)
//
// Included by ppo_fused.cu after fwd_tc5.cuh (uses tc5::make_desc / issue_gemm_ts / split / the TC5_* macros, MbArgs, mma_wgrad).
#pragma once

namespace mb5 {

constexpr int TR = 128, LD = TR + 4, NTH = 512, KX = tc5::KX, NOUT = tc5::NOUT;
constexpr int TMEM_COLS = 256;
constexpr uint32_t XH = 0, XL = 24, C1 = 64, C2 = 128, C3 = 192;

struct Map {   // bytes
  static constexpr int W1 = 0;                                   // [64 n][24 k] hi | lo
  static constexpr int W2F = W1 + 2 * 64 * KX * 4;               // B[n=j][k]  = W2[k][j]   hi | lo
  static constexpr int W2B = W2F + 2 * 64 * 64 * 4;              // B[n=k][kk=j] = W2[k][j] hi | lo
  static constexpr int W3F = W2B + 2 * 64 * 64 * 4;              // B[n=o (16)][k] = W3[k][o]
  static constexpr int W3N = W3F + 2 * NOUT * 64 * 4;            // fp32 [64][8]: W3[k][o] (o >= O zero), for dh2 = dOut W3^T in E4
  static constexpr int XT = W3N + 64 * 8 * 4;                    // fp32 [32][LD]   (rows >= I zero)
  static constexpr int H1T = XT + 32 * LD * 4;                   // fp32 [64][LD]   h1^T, later dz1^T       (prologue: raw parameters)
  static constexpr int H2T = H1T + 64 * LD * 4;                  // fp32 [64][LD]   h2^T, later dz2^T
  static constexpr int OT = H2T + 64 * LD * 4;                   // fp32 [8][LD]    dOut^T (rows >= O zero)
  static constexpr int SX = OT + 8 * LD * 4;                     // gather staging: x [128][I<=24]
  static constexpr int SA = SX + TR * KX * 4;                    //                 actions [128][O<=8]
  static constexpr int SH = SA + TR * 8 * 4;                     //                 logprob | advantage | return [128] each
  static constexpr int IDX = SH + 3 * TR * 4;                    // [128] ints: source rows of the staged tile
  static constexpr int BIAS = IDX + TR * 4;                      // b1[64] b2[64] b3[16]
  static constexpr int LSC = BIAS + (64 + 64 + 16) * 4;          // logΣ[8], σ²[8]
  static constexpr int RED = LSC + 16 * 4;                       // [4][24] head-sum scratch
  static constexpr int BAR = RED + 4 * 24 * 4;
  static constexpr int TOTAL = BAR + 32;
};

template <int HEAD>
__global__ void __launch_bounds__(NTH, 1) minibatch_kernel(MbArgs a) {
  extern __shared__ __align__(1024) unsigned char smb[];
  const NetDesc nd = a.net;
  const int I = nd.I, O = nd.O, act = nd.act;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5, g = lane >> 2, tq = lane & 3;
  const int stop_at = a.ctl ? a.ctl[1] : 0;
  const uint32_t bar_mma = smem_u32(smb + Map::BAR), bar_par = smem_u32(smb + Map::BAR + 8);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smb + Map::BAR + 16);
  float *bias = reinterpret_cast<float *>(smb + Map::BIAS), *lsc = reinterpret_cast<float *>(smb + Map::LSC);
  float *XT = reinterpret_cast<float *>(smb + Map::XT), *H1T = reinterpret_cast<float *>(smb + Map::H1T), *H2T = reinterpret_cast<float *>(smb + Map::H2T);
  float *OT = reinterpret_cast<float *>(smb + Map::OT);
  const float *SXp = reinterpret_cast<const float *>(smb + Map::SX), *SAp = reinterpret_cast<const float *>(smb + Map::SA);
  const float *SHp = reinterpret_cast<const float *>(smb + Map::SH);
  int *sidx = reinterpret_cast<int *>(smb + Map::IDX);
  const int64_t n_tiles = (a.bm + TR - 1) / TR;
  const uint32_t inv_I = (65536u + (uint32_t)I - 1u) / (uint32_t)I, inv_O = (65536u + (uint32_t)O - 1u) / (uint32_t)O;

  auto tile_row = [&](int64_t tile, int r) -> int {
    const int64_t row = tile * TR + r;
    return (tile < n_tiles && row < a.bm) ? (a.order ? a.order[row] : (int)row) : -1;
  };
  auto issue_gather = [&]() {   // rows listed in sidx -> staging (cp.async, zero fill for padding rows)
    for (int e = t; e < TR * I; e += NTH) {
      const int r = (int)(((uint32_t)e * inv_I) >> 16), i = e - r * I;
      const int row = sidx[r];
      cp_async4(const_cast<float *>(SXp) + e, a.s + (row >= 0 ? (int64_t)row * I + i : 0), row >= 0);
    }
    if (HEAD == 0)
      for (int e = t; e < TR * O; e += NTH) {
        const int r = (int)(((uint32_t)e * inv_O) >> 16), o = e - r * O;
        const int row = sidx[r];
        cp_async4(const_cast<float *>(SAp) + e, a.act + (row >= 0 ? (int64_t)row * O + o : 0), row >= 0);
      }
    if (t < TR) {
      const int row = sidx[t];
      if (HEAD == 0) {
        cp_async4(const_cast<float *>(SHp) + t, a.logp_old + (row >= 0 ? row : 0), row >= 0);
        cp_async4(const_cast<float *>(SHp) + TR + t, a.adv + (row >= 0 ? row : 0), row >= 0);
      }
      const bool has_ret = a.ret != nullptr;
      cp_async4(const_cast<float *>(SHp) + 2 * TR + t, has_ret ? a.ret + (row >= 0 ? row : 0) : a.s, has_ret && row >= 0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  // ---- prologue
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_mma), "r"(1) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_par), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (t < TR) sidx[t] = tile_row(blockIdx.x, t);
  for (int e = t; e < 32 * LD; e += NTH) XT[e] = 0.f;
  for (int e = t; e < 8 * LD; e += NTH) OT[e] = 0.f;
  __syncthreads();
  issue_gather();
  if (t == 0) {   // raw parameters -> the h1^T / h2^T area (dead until the first epilogue)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_par), "r"(nd.bytes16) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smb + Map::H1T)), "l"(nd.params),
                 "r"(nd.bytes16), "r"(bar_par)
                 : "memory");
  }
  if (w == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc5::mbar_wait(bar_par, 0);
  {
    const float *P = reinterpret_cast<const float *>(smb + Map::H1T);
    const float *W1 = P, *b1 = P + off_b1(I), *W2 = P + off_W2(I), *b2 = P + off_b2(I), *W3 = P + off_W3(I), *b3 = P + off_b3(I, O);
    auto put = [&](int base, int plane_bytes, int off, float v) {
      float hi, lo;
      tc5::split(v, hi, lo);
      *reinterpret_cast<float *>(smb + base + off) = hi;
      *reinterpret_cast<float *>(smb + base + plane_bytes + off) = lo;
    };
    for (int e = t; e < 64 * KX; e += NTH) { const int j = e / KX, i = e - j * KX; put(Map::W1, 64 * KX * 4, tc5::canon(j, i, KX), i < I ? W1[i * H + j] : 0.f); }
    for (int e = t; e < 64 * 64; e += NTH) {
      const int k = e >> 6, j = e & 63;
      put(Map::W2F, 64 * 64 * 4, tc5::canon(j, k, 64), W2[e]);      // B[n = j][k]
      put(Map::W2B, 64 * 64 * 4, tc5::canon(k, j, 64), W2[e]);      // B[n = k][kk = j]
    }
    for (int e = t; e < NOUT * 64; e += NTH) { const int o = e >> 6, k = e & 63; put(Map::W3F, NOUT * 64 * 4, tc5::canon(o, k, 64), o < O ? W3[k * O + o] : 0.f); }
    for (int e = t; e < 64 * 8; e += NTH) { const int k = e >> 3, o = e & 7; reinterpret_cast<float *>(smb + Map::W3N)[e] = o < O ? W3[k * O + o] : 0.f; }
    if (t < 64) { bias[t] = b1[t]; bias[64 + t] = b2[t]; }
    if (t < NOUT) bias[128 + t] = t < O ? b3[t] : 0.f;
    if (HEAD == 0 && t < O) { const float ls = a.ls[t], sg = expf(ls); lsc[t] = ls; lsc[8 + t] = sg * sg; }
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
  const uint32_t idesc64 = tc5::make_idesc(TR, 64), idesc16 = tc5::make_idesc(TR, NOUT);
  const uint32_t sW1 = smem_u32(smb + Map::W1), sW2F = smem_u32(smb + Map::W2F), sW2B = smem_u32(smb + Map::W2B);
  const uint32_t sW3F = smem_u32(smb + Map::W3F);
  const float *W3N = reinterpret_cast<const float *>(smb + Map::W3N);
  const int row = 32 * (w & 3) + lane, c0 = 16 * (w >> 2);                       // epilogue owner of features [c0, c0 + 16) of `row`
  const uint32_t lane_addr = tmem + ((uint32_t)(32 * (w & 3)) << 16);
  uint32_t ph = 0;

  // per-CTA gradient accumulators (C-fragment layout of the warp-level MMAs), live across tiles
  float acc2[2][4], acc1[1][4], acc3[1][4], accb = 0.f;
#pragma unroll
  for (int e = 0; e < 4; ++e) { acc2[0][e] = acc2[1][e] = 0.f; acc1[0][e] = 0.f; acc3[0][e] = 0.f; }
  float s_obj = 0.f, s_kl = 0.f, s_clip = 0.f, s_adv = 0.f, s_ret = 0.f, dls[MAX_O];
#pragma unroll
  for (int j = 0; j < MAX_O; ++j) dls[j] = 0.f;

#define MB5_LD16(v, taddr)                                                                                                                               \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"                     \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),      \
                 "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])                                                                          \
               : "r"(taddr)                                                                                                                               \
               : "memory")
#define MB5_ST16(taddr, v)                                                                                                                               \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),       \
               "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),      \
               "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])                                                                                             \
               : "memory")
#define MB5_HANDOFF()                                                   \
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");      \
  __syncthreads()
#define MB5_WAIT_MMA()                                                  \
  tc5::mbar_wait(bar_mma, ph); ph ^= 1;                                 \
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory")

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();   // staged rows visible; every MMA / warp-level GEMM of the previous tile is complete
    const int nidx = t < TR ? tile_row(tile + gridDim.x, t) : -1;   // next tile's source rows: requested now, stored after the head
    // ---------------- x row -> TMEM hi/lo (A operand of layer 1) and x^T (fp32, A operand of dW1)
    if ((w >> 2) < KX / 8) {   // thread (row, q): columns [8q, 8q + 8)
      const int q = w >> 2;
      const float *xr = SXp + row * I;
      uint32_t h8[8], l8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int i = 8 * q + j;
        const float x = i < I ? xr[i] : 0.f;
        float hi, lo;
        tc5::split(x, hi, lo);
        h8[j] = __float_as_uint(hi); l8[j] = __float_as_uint(lo);
        if (i < I) XT[i * LD + row] = x;
      }
      TC5_ST8(lane_addr + XH + 8 * q, h8);
      TC5_ST8(lane_addr + XL + 8 * q, l8);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    MB5_HANDOFF();
    // ---------------- layer 1
    if (w == 0) {   // warp-uniform branch + elect.sync: under `if (t == 0)` nvcc wraps every MMA in an election loop (96 instead of 24-48 cycles each)
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (tc5::elect_one()) tc5::issue_gemm_ts<KX / 8>(tmem + C1, tmem + XH, tmem + XL, sW1, 64 * KX * 4, 32 * KX, idesc64, bar_mma);
      __syncwarp();
    }
    MB5_WAIT_MMA();
    {
      uint32_t v[16], l[16];
      MB5_LD16(v, lane_addr + C1 + c0);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float h = act_fused(act, __uint_as_float(v[j]) + bias[c0 + j]);
        float hi, lo;
        tc5::split(h, hi, lo);
        v[j] = __float_as_uint(hi); l[j] = __float_as_uint(lo);
        H1T[(c0 + j) * LD + row] = h;
      }
      MB5_ST16(lane_addr + C1 + c0, v);
      MB5_ST16(lane_addr + C2 + c0, l);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    MB5_HANDOFF();
    // ---------------- layer 2
    if (w == 0) {   // warp-uniform branch + elect.sync: under `if (t == 0)` nvcc wraps every MMA in an election loop (96 instead of 24-48 cycles each)
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (tc5::elect_one()) tc5::issue_gemm_ts<8>(tmem + C3, tmem + C1, tmem + C2, sW2F, 64 * 64 * 4, 32 * 64, idesc64, bar_mma);
      __syncwarp();
    }
    MB5_WAIT_MMA();
    {
      uint32_t v[16], l[16];
      MB5_LD16(v, lane_addr + C3 + c0);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float h = act_fused(act, __uint_as_float(v[j]) + bias[64 + c0 + j]);
        float hi, lo;
        tc5::split(h, hi, lo);
        v[j] = __float_as_uint(hi); l[j] = __float_as_uint(lo);
        H2T[(c0 + j) * LD + row] = h;
      }
      MB5_ST16(lane_addr + C3 + c0, v);
      MB5_ST16(lane_addr + C2 + c0, l);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    MB5_HANDOFF();
    // ---------------- output layer + loss head (thread = row, warps 0..3)
    if (w == 0) {   // warp-uniform branch + elect.sync: under `if (t == 0)` nvcc wraps every MMA in an election loop (96 instead of 24-48 cycles each)
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (tc5::elect_one()) tc5::issue_gemm_ts<8>(tmem + C1, tmem + C3, tmem + C2, sW3F, NOUT * 64 * 4, 32 * 64, idesc16, bar_mma);
      __syncwarp();
    }
    MB5_WAIT_MMA();
    if (w < 4) {
      uint32_t v[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(lane_addr + C1)
                   : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const bool live = sidx[row] >= 0;
      float dout[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) dout[j] = 0.f;
      if (HEAD == 0) {
        float d[8], logp = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (j < O) {
            d[j] = SAp[row * O + j] - (__uint_as_float(v[j]) + bias[128 + j]);
            logp += -(d[j] * d[j]) / (2.f * lsc[8 + j]) - LOG_SQRT_2PI - lsc[j];
          }
        const float Ai = SHp[TR + row], old = SHp[row];
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
          s_kl += old - logp; s_adv += Ai; s_ret += SHp[2 * TR + row];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (j < O) {
            const float var = lsc[8 + j];
            dout[j] = dlogp * d[j] / var;
            dls[j] += dlogp * (d[j] * d[j] / var - 1.f);
          }
      } else {
        const float d = (__uint_as_float(v[0]) + bias[128]) - SHp[2 * TR + row];
        if (live) s_obj += d * d;
        dout[0] = live ? 2.f * d * a.inv_bg : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < O) OT[j * LD + row] = dout[j];
    }
    MB5_HANDOFF();
    // the staging area has been consumed: stream in this CTA's next tile while the backward half runs
    if (t < TR) sidx[t] = nidx;
    // ---------------- dW3 += h2^T dOut, db3 (warp-level MMA, warps 0..3)
    if (w < 4) mma_wgrad<1, TR>(H2T, 16 * w, OT, 0, acc3);
    if (t >= 128 && t < 128 + O) {
      const float *p = OT + (t - 128) * LD;
      for (int r4 = 0; r4 < TR / 4; ++r4) { const float4 d = *reinterpret_cast<const float4 *>(p + 4 * r4); accb += (d.x + d.y) + (d.z + d.w); }
    }
    __syncthreads();   // dW3 has read h2^T; sidx of the next tile is complete
    if (tile + gridDim.x < n_tiles) issue_gather();
    {   // dz2 = (dOut W3^T) .* act'(h2): K <= 8, fp32 FFMA; h2 = hi + lo from tensor memory
      uint32_t hh[16], ll[16];
      MB5_LD16(hh, lane_addr + C3 + c0);
      MB5_LD16(ll, lane_addr + C2 + c0);
      float dout[8];
#pragma unroll
      for (int o = 0; o < 8; ++o) dout[o] = OT[o * LD + row];
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float4 wa = *reinterpret_cast<const float4 *>(W3N + (c0 + j) * 8), wb = *reinterpret_cast<const float4 *>(W3N + (c0 + j) * 8 + 4);
        float dh = 0.f;
        dh = fmaf(dout[0], wa.x, dh); dh = fmaf(dout[1], wa.y, dh); dh = fmaf(dout[2], wa.z, dh); dh = fmaf(dout[3], wa.w, dh);
        dh = fmaf(dout[4], wb.x, dh); dh = fmaf(dout[5], wb.y, dh); dh = fmaf(dout[6], wb.z, dh); dh = fmaf(dout[7], wb.w, dh);
        const float h2 = __uint_as_float(hh[j]) + __uint_as_float(ll[j]);
        const float dz = dh * act_bwd_from_out(act, h2);
        float hi, lo;
        tc5::split(dz, hi, lo);
        hh[j] = __float_as_uint(hi); ll[j] = __float_as_uint(lo);
        H2T[(c0 + j) * LD + row] = dz;
      }
      MB5_ST16(lane_addr + C1 + c0, hh);
      MB5_ST16(lane_addr + C2 + c0, ll);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    MB5_HANDOFF();
    // ---------------- dh1 = dz2 W2^T (tcgen05)   ||   dW2 += h1^T dz2, db2 (warp-level MMA)
    if (w == 0) {   // warp-uniform branch + elect.sync: under `if (t == 0)` nvcc wraps every MMA in an election loop (96 instead of 24-48 cycles each)
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (tc5::elect_one()) tc5::issue_gemm_ts<8>(tmem + C3, tmem + C1, tmem + C2, sW2B, 64 * 64 * 4, 32 * 64, idesc64, bar_mma);
      __syncwarp();
    }
    mma_wgrad<2, TR>(H1T, 16 * (w & 3), H2T, 16 * (w >> 2), acc2);
    if (t < 64) {
      const float *p = H2T + t * LD;
      for (int r4 = 0; r4 < TR / 4; ++r4) { const float4 d = *reinterpret_cast<const float4 *>(p + 4 * r4); accb += (d.x + d.y) + (d.z + d.w); }
    }
    MB5_WAIT_MMA();
    __syncthreads();   // dW2 has read h1^T
    {
      uint32_t v[16];
      MB5_LD16(v, lane_addr + C3 + c0);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float *p = H1T + (c0 + j) * LD + row;
        *p = __uint_as_float(v[j]) * act_bwd_from_out(act, *p);
      }
    }
    MB5_HANDOFF();
    // ---------------- dW1 += x^T dz1 (M = 32: m-tile w & 1, n-tile w >> 1), db1
    mma_wgrad<1, TR>(XT, 16 * (w & 1), H1T, 8 * (w >> 1), acc1);
    if (t >= 64 && t < 128) {
      const float *p = H1T + (t - 64) * LD;
      for (int r4 = 0; r4 < TR / 4; ++r4) { const float4 d = *reinterpret_cast<const float4 *>(p + 4 * r4); accb += (d.x + d.y) + (d.z + d.w); }
    }
  }

  // ---------------- publish this CTA's partial gradient (layout of fused_minibatch_kernel)
  float *out = a.partials + (int64_t)blockIdx.x * a.pstride;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int i = 16 * (w & 1) + g + 8 * (e >> 1), j = 8 * (w >> 1) + 2 * tq + (e & 1);
    if (i < I) out[i * H + j] = acc1[0][e];
  }
#pragma unroll
  for (int q = 0; q < 2; ++q)
#pragma unroll
    for (int e = 0; e < 4; ++e) out[off_W2(I) + (16 * (w & 3) + g + 8 * (e >> 1)) * H + 16 * (w >> 2) + 8 * q + 2 * tq + (e & 1)] = acc2[q][e];
  if (w < 4) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = 16 * w + g + 8 * (e >> 1), o = 2 * tq + (e & 1);
      if (o < O) out[off_W3(I) + k * O + o] = acc3[0][e];
    }
  }
  if (t < 64) out[off_b2(I) + t] = accb;
  else if (t < 128) out[off_b1(I) + (t - 64)] = accb;
  else if (t < 128 + O) out[off_b3(I, O) + (t - 128)] = accb;
  // head sums live in warps 0..3 (thread = row)
  __syncthreads();
  float *red = reinterpret_cast<float *>(smb + Map::RED);
  if (w < 4) {
    float v;
    v = warp_sum(s_obj); if (lane == 0) red[w * 24 + 0] = v;
    v = warp_sum(s_kl); if (lane == 0) red[w * 24 + 1] = v;
    v = warp_sum(s_clip); if (lane == 0) red[w * 24 + 2] = v;
    v = warp_sum(s_adv); if (lane == 0) red[w * 24 + 3] = v;
    v = warp_sum(s_ret); if (lane == 0) red[w * 24 + 4] = v;
#pragma unroll
    for (int j = 0; j < MAX_O; ++j) { v = warp_sum(dls[j]); if (lane == 0) red[w * 24 + 8 + j] = v; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
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
  if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
#undef MB5_LD16
#undef MB5_ST16
#undef MB5_HANDOFF
#undef MB5_WAIT_MMA
}

}  // namespace mb5
