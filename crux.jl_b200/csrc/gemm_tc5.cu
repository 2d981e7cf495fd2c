// Generic fp32 GEMM of the layer-by-layer engine (mlp.cu) on the 5th-generation tensor cores: C = op(A) op(B) with 3xTF32 split
// accumulation (lo*hi + hi*lo + hi*hi: fp32-level accuracy, the 1e-5 parity bar) -- the Dense layers that the fused 17-64-64-X
// kernels do not cover: SAC / DDPG / TD3 256-wide actors and critics (BASELINE config[3]), the pixel-DQN head, any other Chain.
//
// Same operand conventions and epilogues as sgemm_kernel (mlp.cu), so it replaces that kernel call for call:
//     A(m, k) = TA ? A[k*lda + m] : A[m*lda + k]      B(k, n) = TB ? B[n*ldb + k] : B[k*ldb + n]
//     EPI_FWD       C = act(acc + bias[n])                                   Dense forward            (policies.jl:94-96)
//     EPI_BWD_DATA  C = acc * act'(yprev[m][n])                              data gradient            (Zygote pullback of a Dense)
//     EPI_PARTIAL   slab z of the split-K weight gradient [z][M + 1][N]: rows 0..M-1 = acc, row M = column sums of B (bias gradient)
// One CTA = a 128 x 64 tile of C: TMEM lanes = rows (M = 128 MMAs), 64 fp32 accumulator columns.  Per 32-wide k-tile all 256 threads
// fetch their elements of the A and B tiles through the accessors above (coalesced along whichever index is contiguous), split them
// x = hi + lo (hi = rna_tf32(x)) and store both halves as canonical no-swizzle K-major planes (core matrix = 8 rows x 16 B); one elected
// lane then issues 3 passes x 4 k-steps of tcgen05.mma.kind::tf32 (M = 128, N = 64, K = 8) and commits them to the stage's mbarrier, so the
// tensor pipe works on tile kt while the threads stage tile kt + 1 into the other buffer.  The epilogue reads the accumulator with
// tcgen05.ld (32 lanes x 32 columns per warp) and applies bias / activation / mask on the way to global memory.
// 96 KB of shared memory per CTA (2 stages x 2 planes x (128 + 64) x 32 floats): two CTAs per SM, 64 TMEM columns each.
#include "mlp.cuh"

namespace {

constexpr int BM = 128, BN = 64, BK = 32, NTH = 256, NST = 4;               // NST stages of one k-tile each
constexpr int A_PLANE = BM * BK * 4, B_PLANE = BN * BK * 4;                 // bytes
constexpr int STAGE = 2 * A_PLANE + 2 * B_PLANE;                            // A hi | A lo | B hi | B lo   (48 KB)
constexpr int SM_BAR = NST * STAGE;                                         // bar[NST] (stage free), bar_done, tmem slot
constexpr int SM_BSUM = SM_BAR + 64;                                        // [8][64] column-sum scratch (EPI_PARTIAL)
constexpr int SM_TOTAL = SM_BSUM + 8 * BN * 4;
constexpr int LDC_S = BN + 1;                                               // padded row of the C tile staged for coalesced stores
static_assert(BM * LDC_S * 4 <= STAGE, "the C tile is staged in stage 0");
enum { EPI_FWD = 0, EPI_BWD_DATA = 1, EPI_PARTIAL = 2 };

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ int canon(int row, int k) { return (row >> 3) * (32 * BK) + (k >> 2) * 128 + (row & 7) * 16 + (k & 3) * 4; }   // bytes
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {   // K-major, no swizzle: LBO = 128 B (cores adjacent along K), SBO = 32 BK B
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((128u >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(((32u * BK) >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {   // D = F32, A = B = TF32, both K-major
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(da),
               "l"(db), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void split(float x, float &hi, float &lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
  lo = x - hi;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
// asynchronous global -> shared copies; `bytes` of the source are read, the rest of the destination is zero-filled (edges of the matrices)
__device__ __forceinline__ void cp16(uint32_t dst, const void *src, int bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp4(uint32_t dst, const void *src, int bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}

// One operand tile (ROWS x BK, canonical K-major) of k-tile [k0, k0 + BK): element (row, k) = TR ? src[(k0 + k) * ld + r0 + row] : src[(r0 + row) * ld + k0 + k].
// The copies land DIRECTLY in the canonical layout: a 16-byte chunk of four consecutive k of one row is one core-matrix row, so a
// k-contiguous, 16-byte-aligned source needs one cp.async per chunk; every other source (row-contiguous = transposed, or rows that are
// not 16-byte aligned such as the 393-wide vcat(s, a) input of a critic) is copied element by element, coalesced along its contiguous index.
template <int ROWS, bool TR>
__device__ __forceinline__ void stage_tile(uint32_t dst, const float *__restrict__ src, int ld, int r0, int n_rows, int k0, int k_end, bool vec, int t) {
  if (!TR && vec) {
#pragma unroll
    for (int i = 0; i < ROWS * (BK / 4) / NTH; ++i) {
      const int c = t + i * NTH, row = c >> 3, kc = c & 7;      // 8 chunks per row: a warp reads 4 rows x 128 contiguous bytes
      const int gr = r0 + row, gk = k0 + 4 * kc;
      const int nb = (gr < n_rows && gk < k_end) ? min(16, 4 * (k_end - gk)) : 0;
      cp16(dst + canon(row, 4 * kc), nb ? (const void *)(src + (int64_t)gr * ld + gk) : (const void *)src, nb);
    }
  } else if (TR) {
    // consecutive threads walk the rows (contiguous in memory); thread t: row = t % ROWS, k = t / ROWS + (NTH / ROWS) i
    constexpr int KS = NTH / ROWS, NI = ROWS * BK / NTH;
    const int row = t % ROWS, kb = t / ROWS;
    const bool row_ok = r0 + row < n_rows;
    const float *g = src + (int64_t)(k0 + kb) * ld + r0 + row;
    const uint32_t d0 = dst + (row >> 3) * (32 * BK) + (row & 7) * 16;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int k = kb + KS * i;
      const bool ok = row_ok && k0 + k < k_end;
      cp4(d0 + (k >> 2) * 128 + (k & 3) * 4, ok ? (const void *)(g + (int64_t)(KS * i) * ld) : (const void *)src, ok ? 4 : 0);
    }
  } else {
    // k-contiguous rows that are not 16-byte aligned: thread t: k = t % 32, row = t / 32 + 8 i
    constexpr int NI = ROWS * BK / NTH;
    const int k = t & (BK - 1), rb = t >> 5;
    const bool k_ok = k0 + k < k_end;
    const float *g = src + (int64_t)(r0 + rb) * ld + k0 + k;
    const uint32_t d0 = dst + (k >> 2) * 128 + (k & 3) * 4 + rb * 16;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const bool ok = k_ok && r0 + rb + 8 * i < n_rows;
      cp4(d0 + i * (32 * BK), ok ? (const void *)(g + (int64_t)(8 * i) * ld) : (const void *)src, ok ? 4 : 0);
    }
  }
}

template <bool TA, bool TB, int EPI>
__global__ void __launch_bounds__(NTH, 1) gemm_tc5_kernel(const float *__restrict__ A, int lda, const float *__restrict__ Bm, int ldb, float *__restrict__ C,
                                                          int ldc, int M, int N, int K, const float *__restrict__ bias, int act,
                                                          const float *__restrict__ yprev, int prev_act, int k_per_slab, int bias_row,
                                                          const int *__restrict__ skip, int vec_a, int vec_b) {
  if (skip && *skip) return;
  extern __shared__ __align__(1024) unsigned char smb[];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
  int k_begin = 0, k_end = K;
  if (EPI == EPI_PARTIAL) { k_begin = blockIdx.z * k_per_slab; k_end = min(K, k_begin + k_per_slab); }
  const uint32_t sb = smem_u32(smb);
  const uint32_t bar0 = sb + SM_BAR, bar_done = bar0 + 8 * NST;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smb + SM_BAR + 8 * NST + 8);
  if (t == 0) {
    for (int i = 0; i <= NST; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8 * i), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (w == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  const int n_kt = (k_end - k_begin + BK - 1) / BK;
  // prologue of the copy pipeline: tiles 0 .. NST-2 in flight (one commit group per tile, empty past the end so that the counts stay uniform)
#pragma unroll
  for (int p = 0; p < NST - 1; ++p) {
    if (p < n_kt) {
      stage_tile<BM, TA>(sb + p * STAGE, A, lda, m0, M, k_begin + p * BK, k_end, vec_a != 0, t);
      stage_tile<BN, !TB>(sb + p * STAGE + 2 * A_PLANE, Bm, ldb, n0, N, k_begin + p * BK, k_end, vec_b != 0, t);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  const uint32_t idesc = make_idesc(BM, BN);
  const bool do_bias = (EPI == EPI_PARTIAL) && bias_row && blockIdx.y == 0;
  float bsum[2] = {0.f, 0.f};   // column sums of the B tile: chunk i of this thread always belongs to column n = 8 (t / 64 + 4 i) + t % 8

  for (int kt = 0; kt < n_kt; ++kt) {
    const int s = kt % NST;
    unsigned char *st = smb + s * STAGE;
    asm volatile("cp.async.wait_group %0;" ::"n"(NST - 2) : "memory");   // this thread's copies of tile kt have landed
    __syncthreads();                                                      // ... and everybody else's
    // x -> hi (in place) | lo: every thread walks the planes linearly in 16-byte chunks (conflict-free)
#pragma unroll
    for (int i = 0; i < A_PLANE / 16 / NTH; ++i) {
      float4 *ph = reinterpret_cast<float4 *>(st) + t + i * NTH;
      const float4 x = *ph;
      float4 h, l;
      split(x.x, h.x, l.x); split(x.y, h.y, l.y); split(x.z, h.z, l.z); split(x.w, h.w, l.w);
      *ph = h;
      *(ph + A_PLANE / 16) = l;
    }
#pragma unroll
    for (int i = 0; i < B_PLANE / 16 / NTH; ++i) {
      float4 *ph = reinterpret_cast<float4 *>(st + 2 * A_PLANE) + t + i * NTH;
      const float4 x = *ph;
      if (do_bias) bsum[i] += (x.x + x.y) + (x.z + x.w);
      float4 h, l;
      split(x.x, h.x, l.x); split(x.y, h.y, l.y); split(x.z, h.z, l.z); split(x.w, h.w, l.w);
      *ph = h;
      *(ph + B_PLANE / 16) = l;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor cores
    __syncthreads();
    if (w == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint32_t a_hi = sb + s * STAGE, a_lo = a_hi + A_PLANE, b_hi = a_hi + 2 * A_PLANE, b_lo = b_hi + B_PLANE;
#pragma unroll
        for (int ks = 0; ks < BK / 8; ++ks) mma_tf32_ss(tmem, make_desc(a_lo + ks * 256), make_desc(b_hi + ks * 256), idesc, (kt || ks) ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < BK / 8; ++ks) mma_tf32_ss(tmem, make_desc(a_hi + ks * 256), make_desc(b_lo + ks * 256), idesc, 1u);
#pragma unroll
        for (int ks = 0; ks < BK / 8; ++ks) mma_tf32_ss(tmem, make_desc(a_hi + ks * 256), make_desc(b_hi + ks * 256), idesc, 1u);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar0 + 8 * s) : "memory");
        if (kt == n_kt - 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_done) : "memory");
      }
      __syncwarp();
    }
    // refill: tile kt + NST - 1 goes into the stage tile kt - 1 used; its MMAs were committed one iteration ago
    const int nt = kt + NST - 1;
    if (nt < n_kt) {
      const int sn = nt % NST;
      if (kt >= 1) mbar_wait(bar0 + 8 * sn, (uint32_t)(((kt - 1) / NST) & 1));
      stage_tile<BM, TA>(sb + sn * STAGE, A, lda, m0, M, k_begin + nt * BK, k_end, vec_a != 0, t);
      stage_tile<BN, !TB>(sb + sn * STAGE + 2 * A_PLANE, Bm, ldb, n0, N, k_begin + nt * BK, k_end, vec_b != 0, t);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  // ---- epilogue: warp w reads TMEM lanes [32 (w & 3), +32) x columns [32 (w >> 2), +32), stages the C tile in shared memory (stage 0 is
  //      free once every MMA has completed) and the CTA writes it out row by row: coalesced stores, coalesced reads of the mask operand
  float *Cz = C;
  if (EPI == EPI_PARTIAL) Cz = C + (int64_t)blockIdx.z * (int64_t)(M + (bias_row ? 1 : 0)) * ldc;
  if (n_kt > 0) {
    mbar_wait(bar_done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[32];
    const int c0 = 32 * (w >> 2);
    const uint32_t taddr = tmem + ((uint32_t)(32 * (w & 3)) << 16) + (uint32_t)c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
        "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
          "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
          "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    float *sC = reinterpret_cast<float *>(smb);
    const int r = 32 * (w & 3) + lane;
#pragma unroll
    for (int j = 0; j < 32; ++j) sC[r * LDC_S + c0 + j] = __uint_as_float(v[j]);
    __syncthreads();
    for (int e = t; e < BM * BN; e += NTH) {
      const int row = e >> 6, col = e & 63;
      const int gm = m0 + row, gn = n0 + col;
      if (gm < M && gn < N) {
        float x = sC[row * LDC_S + col];
        if (EPI == EPI_FWD) x = act_fwd_rt(act, x + bias[gn]);
        if (EPI == EPI_BWD_DATA && yprev) x *= act_bwd_from_out(prev_act, yprev[(int64_t)gm * ldc + gn]);
        Cz[(int64_t)gm * ldc + gn] = x;
      }
    }
  } else if (EPI == EPI_PARTIAL) {   // empty slab: zeros
    for (int e = t; e < BM * BN; e += NTH) {
      const int gm = m0 + e / BN, gn = n0 + e % BN;
      if (gm < M && gn < N) Cz[(int64_t)gm * ldc + gn] = 0.f;
    }
  }
  if (do_bias) {   // bias gradient of this slab: column sums of B over k.  Chunk i of thread t is k-chunk (t % 64) / 8 of column 8 (t / 64 + 4 i) + t % 8
    float *sc = reinterpret_cast<float *>(smb + SM_BSUM);
#pragma unroll
    for (int i = 0; i < 2; ++i) sc[((t & 63) >> 3) * BN + 8 * ((t >> 6) + 4 * i) + (t & 7)] = bsum[i];
    __syncthreads();
    if (t < BN && n0 + t < N) {
      float sum = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) sum += sc[q * BN + t];   // fixed order: bit-reproducible
      Cz[(int64_t)M * ldc + n0 + t] = sum;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory");
}

template <bool TA, bool TB, int EPI>
int launch(crux_ctx *ctx, dim3 grid, const float *A, int lda, const float *B, int ldb, float *C, int ldc, int M, int N, int K, const float *bias, int act,
           const float *yprev, int prev_act, int k_per_slab, int bias_row, const int *skip) {
  static bool attr = false;
  if (!attr) {
    CRUX_CHECK_CUDA(ctx, cudaFuncSetAttribute(gemm_tc5_kernel<TA, TB, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
    attr = true;
  }
  // 16-byte copies need a k-contiguous operand whose rows start 16-byte aligned
  const int vec_a = (!TA && lda % 4 == 0 && ((uintptr_t)A & 15) == 0) ? 1 : 0;
  const int vec_b = (TB && ldb % 4 == 0 && ((uintptr_t)B & 15) == 0) ? 1 : 0;
  gemm_tc5_kernel<TA, TB, EPI><<<grid, NTH, SM_TOTAL, ctx->stream>>>(A, lda, B, ldb, C, ldc, M, N, K, bias, act, yprev, prev_act, k_per_slab, bias_row, skip,
                                                                     vec_a, vec_b);
  CRUX_LAUNCHED(ctx);
  return CRUX_OK;
}

}  // namespace

// Opt-in (CRUX_GEMM_TC5=1, read per call: A/B runs and tests switch it).  Measured on B200 (scripts/bench_offpolicy.py): parity-green
// through the whole generic-engine suite, but at the batch sizes of the off-policy configs the layer-by-layer engine is bound by its ~70
// small dependent launches per update, not by the GEMM pipe -- SAC 376/17/256-256 at B = 2048: 1.60 ms per update with this kernel (64 CTAs,
// element-wise staging through the operand accessors) against 1.17 ms with the 64 x 64 FFMA tiles (128 CTAs); pixel-DQN update 3.63 against
// 3.29 ms.  It pays only once the layers of a network are fused around it (what mb_t5.cuh does for the 64-wide PPO networks).
bool gemm_tc5_eligible(int64_t M, int64_t N, int64_t K) {
  const char *on = getenv("CRUX_GEMM_TC5");
  return on && on[0] == '1' && M >= 64 && N >= 16 && K >= 16;
}
// Dense forward: y[B][N] = act(x[B][K] W[K][N] + b)
int gemm_tc5_fwd(crux_ctx *ctx, const float *x, int K, const float *W, int N, float *y, int64_t B, const float *b, int act, const int *skip) {
  dim3 grid((unsigned)cdiv(N, BN), (unsigned)cdiv(B, BM), 1);
  return launch<false, false, EPI_FWD>(ctx, grid, x, K, W, N, y, N, (int)B, N, K, b, act, nullptr, 0, 0, 0, skip);
}
// data gradient: dx[B][K] = (dz[B][N] W^T) .* act'(yprev[B][K])   (yprev NULL: no mask)
int gemm_tc5_bwd_data(crux_ctx *ctx, const float *dz, int N, const float *W, float *dx, int K, int64_t B, const float *yprev, int prev_act, const int *skip) {
  dim3 grid((unsigned)cdiv(K, BN), (unsigned)cdiv(B, BM), 1);
  return launch<false, true, EPI_BWD_DATA>(ctx, grid, dz, N, W, N, dx, K, (int)B, K, N, nullptr, 0, yprev, prev_act, 0, 0, skip);
}
// weight-gradient partial slabs: part[z][K + 1][N] = x[rows of slab z]^T dz[rows of slab z]  (+ bias row: column sums of dz)
int gemm_tc5_wgrad(crux_ctx *ctx, const float *x, int K, const float *dz, int N, float *part, int64_t B, int slabs, int rows_per_slab, const int *skip) {
  dim3 grid((unsigned)cdiv(N, BN), (unsigned)cdiv(K, BM), (unsigned)slabs);
  return launch<true, false, EPI_PARTIAL>(ctx, grid, x, K, dz, N, part, N, K, N, (int)B, nullptr, 0, nullptr, 0, rows_per_slab, 1, skip);
}
