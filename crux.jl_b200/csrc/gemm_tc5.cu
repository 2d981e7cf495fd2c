// Generic fp32 GEMM of the layer-by-layer engine (mlp.cu) on the 5th-generation tensor cores: C = op(A) op(B) with 3xTF32 split
// accumulation (lo*hi + hi*lo + hi*hi: fp32-level accuracy, the 1e-5 parity bar) -- the Dense layers that the fused 17-64-64-X
// kernels do not cover: SAC / DDPG / TD3 256-wide actors and critics (BASELINE config[3]), the pixel-DQN head, any other Chain.
//
// Same operand conventions and epilogues as sgemm_kernel (mlp.cu), so it replaces that kernel call for call:
//     A(m, k) = TA ? A[k*lda + m] : A[m*lda + k]      B(k, n) = TB ? B[n*ldb + k] : B[k*ldb + n]
//     EPI_FWD       C = act(acc + bias[n])                                   Dense forward            (policies.jl:94-96)
//     EPI_BWD_DATA  C = acc * act'(yprev[m][n])                              data gradient            (Zygote pullback of a Dense)
//     EPI_PARTIAL   slab z of the split-K weight gradient [z][M + 1][N]: rows 0..M-1 = acc, row M = column sums of B (bias gradient)
// One CTA = a 128 x 64 tile of C: TMEM lanes = rows (M = 128 MMAs), 64 fp32 accumulator columns, 32-wide k-tiles.  Both operands are copied
// RAW into shared memory by 16-byte cp.async (element copies for unaligned rows); x = hi + lo (hi = rna_tf32(x)) happens in registers on
// the way to the tensor cores: the A operand into TENSOR MEMORY (tcgen05.st), the B operand into canonical no-swizzle K-major planes
// (core matrix = 8 rows x 16 B).  One warp only issues 3 passes x 4 k-steps of tcgen05.mma.kind::tf32 (M = 128, N = 64, K = 8) per k-tile.
// The epilogue reads the accumulator with tcgen05.ld, stages the C tile in shared memory and applies bias / activation / mask on the way
// to global memory with coalesced accesses.  Two variants: 4 stages, one CTA per SM (172 KB) for grids that leave SMs idle anyway; 2 stages,
// two CTAs per SM (86 KB, <= 92 registers) for grids of more than one wave.  Details and the measurements behind them: the comment on
// gemm_tc5_kernel below and profiles/r2_notes.md.
#include "mlp.cuh"

namespace {

constexpr int BM = 128, BN = 64, BK = 32, NW = 256, NTH = 288, NTB = 128;   // NW worker threads (warps 0..7) + the MMA issuer warp 8
constexpr int B_PLANE = BN * BK * 4;                                        // bytes
constexpr int LDA_S = BK + 4;                                               // padded row of the raw A tile [m][k] (floats): rows and chunks conflict-free
constexpr int A_RAW = BM * LDA_S * 4;                                       // 18 KB (the transposed form [k][m] needs 16 KB)
constexpr int B_RAW = BN * LDA_S * 4;                                       // raw B tile: [n][LDA_S] of a k-contiguous source (9 KB) or [k][BN] of an n-contiguous one (8 KB)
constexpr int OFF_A = 2 * B_PLANE, OFF_BR = OFF_A + A_RAW;
constexpr int STAGE = OFF_BR + B_RAW;                                       // B hi | B lo | A raw | B raw   (42 KB)
constexpr int C_STAGE = BM * (BN + 1) * 4;                                  // the C tile staged for coalesced stores (padded rows)
// NST stages of one k-tile each: 4 (172 KB, one CTA per SM) for grids that leave SMs idle anyway, 2 (86 KB, two CTAs per SM at <= 112 registers)
// for grids of several waves, where a second resident CTA hides the per-k-tile latencies of the first
template <int NST> struct Smem {
  static constexpr int B = C_STAGE > NST * STAGE ? C_STAGE : NST * STAGE;
  static constexpr int BAR = B;                                             // bar_free[2], bar_ready[2], bar_done, tmem slot
  static constexpr int BSUM = BAR + 64;                                     // [4][64] column-sum scratch (EPI_PARTIAL)
  static constexpr int TOTAL = BSUM + 4 * BN * 4;
};
constexpr int LDC_S = BN + 1;
// tensor memory: accumulator D [128 lanes][64 columns] | two A buffers, each hi [32 columns] | lo [32 columns]
constexpr uint32_t TM_D = 0, TM_A = 64, TM_COLS = 256;
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
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem), "r"(a_tmem),
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
#define G5_ST32(taddr, v)                                                                                                                              \
  asm volatile(                                                                                                                                          \
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, " \
      "%24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),                                                                                      \
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]),      \
      "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]),    \
      "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])                                                               \
      : "memory")

// ---- operand staging (all 256 worker threads copy; one k-tile = [k0, k0 + BK)) -------------------------------------------------------------------
// A raw tile.  !TA: element (m, k) = A[(m0 + m) * lda + k0 + k] -> shared [m][LDA_S]; aligned rows: 16-byte chunks, lanes = 8 rows x 4 chunks (whole
// 32-byte sectors in global memory, distinct bank groups in shared memory), else element copies.  TA: element (m, k) = A[(k0 + k) * lda + m0 + m]
// -> shared [k][BM] (the A-side thread m then reads a column: consecutive lanes, consecutive words); 16-byte chunks along m when aligned.
template <bool TA>
__device__ __forceinline__ void stage_a(uint32_t dst, const float *__restrict__ A, int lda, int m0, int M, int k0, int k_end, bool vec, int t) {
  if (!TA) {
    if (vec) {
      const int w = t >> 5, l = t & 31;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int u = w * 4 + i, row = (u >> 1) * 8 + (l & 7), kc = (u & 1) * 4 + (l >> 3);
        const int gr = m0 + row, gk = k0 + 4 * kc;
        const int nb = (gr < M && gk < k_end) ? min(16, 4 * (k_end - gk)) : 0;
        cp16(dst + (row * LDA_S + 4 * kc) * 4, nb ? (const void *)(A + (int64_t)gr * lda + gk) : (const void *)A, nb);
      }
    } else {           // unaligned k-contiguous rows: thread t: k = t % 32, row = t / 32 + 8 i
      const int k = t & (BK - 1), rb = t >> 5;
      const bool k_ok = k0 + k < k_end;
      const float *g = A + (int64_t)(m0 + rb) * lda + k0 + k;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int row = rb + 8 * i;
        const bool ok = k_ok && m0 + row < M;
        cp4(dst + (row * LDA_S + k) * 4, ok ? (const void *)(g + (int64_t)(8 * i) * lda) : (const void *)A, ok ? 4 : 0);
      }
    }
  } else {
    if (vec) {         // 32 k-rows x 32 chunks of 4 m: thread t: chunk mc = t % 32, k = t / 32 + 8 i
      const int mc = t & 31, kb = t >> 5;
      const int gm = m0 + 4 * mc;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = kb + 8 * i;
        const int nb = (k0 + k < k_end && gm < M) ? min(16, 4 * (M - gm)) : 0;
        cp16(dst + (k * BM + 4 * mc) * 4, nb ? (const void *)(A + (int64_t)(k0 + k) * lda + gm) : (const void *)A, nb);
      }
    } else {           // thread t: m = t % 128, k = t / 128 + 2 i
      const int row = t & 127, kb = t >> 7;
      const bool row_ok = m0 + row < M;
      const float *g = A + (int64_t)(k0 + kb) * lda + m0 + row;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int k = kb + 2 * i;
        const bool ok = row_ok && k0 + k < k_end;
        cp4(dst + (k * BM + row) * 4, ok ? (const void *)(g + (int64_t)(2 * i) * lda) : (const void *)A, ok ? 4 : 0);
      }
    }
  }
}
// B raw tile.  TB (k-contiguous source, element (n, k) = B[(n0 + n) * ldb + k0 + k]) -> shared [n][LDA_S]; !TB (n-contiguous source, element
// (n, k) = B[(k0 + k) * ldb + n0 + n]) -> shared [k][BN]; 16-byte chunks along the contiguous index when the rows are aligned.  The split pass
// writes the canonical K-major hi | lo planes from it (transposing in the !TB case), so a copy never waits for the tensor pipe.
template <bool TB>
__device__ __forceinline__ void stage_b(uint32_t stage, const float *__restrict__ Bm, int ldb, int n0, int N, int k0, int k_end, bool vec, int t) {
  if (TB) {
    const uint32_t dst = stage + OFF_BR;
    if (vec) {         // 64 rows x 8 chunks = 512 chunks, 2 per thread
      const int w = t >> 5, l = t & 31;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int u = w * 2 + i, row = (u >> 1) * 8 + (l & 7), kc = (u & 1) * 4 + (l >> 3);
        const int gr = n0 + row, gk = k0 + 4 * kc;
        const int nb = (gr < N && gk < k_end) ? min(16, 4 * (k_end - gk)) : 0;
        cp16(dst + (row * LDA_S + 4 * kc) * 4, nb ? (const void *)(Bm + (int64_t)gr * ldb + gk) : (const void *)Bm, nb);
      }
    } else {           // thread t: k = t % 32, row = t / 32 + 8 i
      const int k = t & (BK - 1), rb = t >> 5;
      const bool k_ok = k0 + k < k_end;
      const float *g = Bm + (int64_t)(n0 + rb) * ldb + k0 + k;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = rb + 8 * i;
        const bool ok = k_ok && n0 + row < N;
        cp4(dst + (row * LDA_S + k) * 4, ok ? (const void *)(g + (int64_t)(8 * i) * ldb) : (const void *)Bm, ok ? 4 : 0);
      }
    }
  } else {
    const uint32_t dst = stage + OFF_BR;
    if (vec) {         // 32 k-rows x 16 chunks of 4 n = 512 chunks, 2 per thread: thread t: chunk nc = t % 16, k = t / 16 + 16 i
      const int nc = t & 15, kb = t >> 4;
      const int gn = n0 + 4 * nc;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int k = kb + 16 * i;
        const int nb = (k0 + k < k_end && gn < N) ? min(16, 4 * (N - gn)) : 0;
        cp16(dst + (k * BN + 4 * nc) * 4, nb ? (const void *)(Bm + (int64_t)(k0 + k) * ldb + gn) : (const void *)Bm, nb);
      }
    } else {           // thread t: n = t % 64, k = t / 64 + 4 i
      const int n = t & 63, kb = t >> 6;
      const bool n_ok = n0 + n < N;
      const float *g = Bm + (int64_t)(k0 + kb) * ldb + n0 + n;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = kb + 4 * i;
        const bool ok = n_ok && k0 + k < k_end;
        cp4(dst + (k * BN + n) * 4, ok ? (const void *)(g + (int64_t)(4 * i) * ldb) : (const void *)Bm, ok ? 4 : 0);
      }
    }
  }
}

// One CTA = a 128 x 64 tile of C.  Shared-memory bandwidth and lock-step phases are what the first versions of this kernel ran out of: with
// both operands written hi | lo into shared memory and every thread walking copy -> split -> barrier -> MMA issue in turn, a 32-wide k-tile
// cost 2 300-2 700 cycles (copies 1 060, split 420, MMAs 400, barriers 400 -- strictly added up, CRUX_G5_DEBUG).  Now:
//   * every global -> shared copy is a 16-byte cp.async whenever the source allows it (4-stage pipeline, 256 worker threads);
//   * the A operand leaves shared memory only once: warps 0..3 (thread = row = TMEM lane) read their row of the RAW tile, split it and store
//     hi | lo into one of two A buffers in TENSOR MEMORY (tcgen05.st); warps 4..7 split (and, for n-contiguous sources, transpose) the B tile;
//   * warp 8 does nothing but wait for "operands of tile kt ready" (mbarrier, 256 arrivals) and issue 3 passes x 4 k-steps of
//     tcgen05.mma.kind::tf32 (A from tensor memory, B from shared memory): the workers never wait for the tensor pipe except through the
//     buffer-reuse barriers two tiles back.
template <bool TA, bool TB, int EPI, int NST>
__global__ void __launch_bounds__(NTH, NST == 2 ? 2 : 1) gemm_tc5_kernel(const float *__restrict__ A, int lda, const float *__restrict__ Bm, int ldb, float *__restrict__ C,
                                                          int ldc, int M, int N, int K, const float *__restrict__ bias, int act,
                                                          const float *__restrict__ yprev, int prev_act, int k_per_slab, int bias_row,
                                                          const int *__restrict__ skip, int vec_a, int vec_b) {
  // Programmatic dependent launch: the next kernel of the stream may be scheduled at once (onto SMs this grid leaves free -- 64 CTAs at the
  // off-policy batch sizes); this kernel's own on-chip prologue (barriers, tensor-memory allocation) runs under the tail of its predecessor,
  // and only then waits for the predecessor's results (griddepcontrol.wait below: everything before it touches no global memory).
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  extern __shared__ __align__(1024) unsigned char smb[];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
  int k_begin = 0, k_end = K;
  if (EPI == EPI_PARTIAL) { k_begin = blockIdx.z * k_per_slab; k_end = min(K, k_begin + k_per_slab); }
  const uint32_t sb = smem_u32(smb);
  constexpr int SM_BAR = Smem<NST>::BAR, SM_BSUM = Smem<NST>::BSUM;
  const uint32_t bar_free = sb + SM_BAR, bar_ready = bar_free + 16, bar_done = bar_free + 32;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smb + SM_BAR + 40);
  if (t == 0) {
    for (int i = 0; i < 2; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_free + 8 * i), "r"(1) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_ready + 8 * i), "r"(NW) : "memory");
    }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_done), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (w == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");   // the predecessor's activations / gradients / stop flag are complete and visible
  const bool skipped = skip && *skip;                   // uniform over the grid; the allocation above is still released below
  const int n_kt = skipped ? 0 : (k_end - k_begin + BK - 1) / BK;
  const bool worker = w < 8, a_side = w < 4;
  const int tb = t - NTB;               // B-side thread index (warps 4..7)
  if (worker) {   // copy pipeline: tiles 0 .. NST-2 in flight (one commit group per tile, empty past the end so that the counts stay uniform)
#pragma unroll
    for (int p = 0; p < NST - 1; ++p) {
      if (p < n_kt) {
        stage_a<TA>(sb + p * STAGE + OFF_A, A, lda, m0, M, k_begin + p * BK, k_end, vec_a != 0, t);
        stage_b<TB>(sb + p * STAGE, Bm, ldb, n0, N, k_begin + p * BK, k_end, vec_b != 0, t);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  const bool do_bias = (EPI == EPI_PARTIAL) && bias_row && blockIdx.y == 0 && !skipped;
  float bsum[2] = {0.f, 0.f};   // !TB split: thread (wb, l) owns columns l (even units) and 32 + l (odd units)

  if (!worker) {
    // ================================================================ MMA issuer (warp 8)
    const uint32_t idesc = make_idesc(BM, BN);
    for (int kt = 0; kt < n_kt; ++kt) {
      const int buf = kt & 1, s = kt % NST;
      mbar_wait(bar_ready + 8 * buf, (uint32_t)((kt >> 1) & 1));   // the A buffer and the B planes of tile kt are complete
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint32_t a_hi = tmem + TM_A + 64 * buf, a_lo = a_hi + 32;
        const uint64_t b_hi = make_desc(sb + s * STAGE), b_lo = make_desc(sb + s * STAGE + B_PLANE);
#pragma unroll
        for (int ks = 0; ks < BK / 8; ++ks) mma_tf32_ts(tmem + TM_D, a_lo + 8 * ks, b_hi + 16 * ks, idesc, (kt || ks) ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < BK / 8; ++ks) mma_tf32_ts(tmem + TM_D, a_hi + 8 * ks, b_lo + 16 * ks, idesc, 1u);
#pragma unroll
        for (int ks = 0; ks < BK / 8; ++ks) mma_tf32_ts(tmem + TM_D, a_hi + 8 * ks, b_hi + 16 * ks, idesc, 1u);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_free + 8 * buf) : "memory");
        if (kt == n_kt - 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_done) : "memory");
      }
      __syncwarp();
    }
  } else {
    // ================================================================ workers (warps 0..7)
    for (int kt = 0; kt < n_kt; ++kt) {
      const int buf = kt & 1, s = kt % NST;
      unsigned char *st = smb + s * STAGE;
      asm volatile("cp.async.wait_group %0;" ::"n"(NST - 2) : "memory");   // this thread's copies of tile kt have landed
      asm volatile("bar.sync 1, %0;" ::"n"(NW) : "memory");                // ... and every other worker's
      // refill first (asynchronous): tile kt + NST - 1 goes into the stage tile kt - 1 used.  Its raw tiles were read before the barrier
      // above; its B planes are free once the MMAs of tile kt - 1 have completed.
      const int nt = kt + NST - 1;
      if (nt < n_kt) {
        const uint32_t sn = sb + (nt % NST) * STAGE;
        stage_a<TA>(sn + OFF_A, A, lda, m0, M, k_begin + nt * BK, k_end, vec_a != 0, t);
        stage_b<TB>(sn, Bm, ldb, n0, N, k_begin + nt * BK, k_end, vec_b != 0, t);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (a_side) {   // thread = row = TMEM lane: its 32 k values -> hi | lo -> A buffer `buf` in tensor memory
        uint32_t hi[32], lo[32];
        if (!TA) {
          const float4 *row = reinterpret_cast<const float4 *>(st + OFF_A + t * LDA_S * 4);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 x = row[q];
            float h, l;
            split(x.x, h, l); hi[4 * q] = __float_as_uint(h); lo[4 * q] = __float_as_uint(l);
            split(x.y, h, l); hi[4 * q + 1] = __float_as_uint(h); lo[4 * q + 1] = __float_as_uint(l);
            split(x.z, h, l); hi[4 * q + 2] = __float_as_uint(h); lo[4 * q + 2] = __float_as_uint(l);
            split(x.w, h, l); hi[4 * q + 3] = __float_as_uint(h); lo[4 * q + 3] = __float_as_uint(l);
          }
        } else {
          const float *col = reinterpret_cast<const float *>(st + OFF_A) + t;
#pragma unroll
          for (int k = 0; k < 32; ++k) { float h, l; split(col[k * BM], h, l); hi[k] = __float_as_uint(h); lo[k] = __float_as_uint(l); }
        }
        if (kt >= 2) mbar_wait(bar_free + 8 * buf, (uint32_t)(((kt - 2) >> 1) & 1));   // the MMAs of tile kt - 2 have read this A buffer
        const uint32_t ta = tmem + ((uint32_t)(32 * w) << 16) + TM_A + 64 * buf;
        G5_ST32(ta, hi);
        G5_ST32(ta + 32, lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      } else if (TB) {   // raw [n][LDA_S] -> canonical hi | lo: lanes = 8 rows x 4 chunks, conflict-free 16-byte reads and stores
        if (kt >= 2) mbar_wait(bar_free + 8 * buf, (uint32_t)(((kt - 2) >> 1) & 1));   // (see the !TB branch)
        const int wb = tb >> 5;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int u = wb * 4 + i, n = (u >> 1) * 8 + (lane & 7), kc = (u & 1) * 4 + (lane >> 3);
          const float4 x = *reinterpret_cast<const float4 *>(st + OFF_BR + (n * LDA_S + 4 * kc) * 4);
          float4 h, l;
          split(x.x, h.x, l.x); split(x.y, h.y, l.y); split(x.z, h.z, l.z); split(x.w, h.w, l.w);
          float4 *ph = reinterpret_cast<float4 *>(st + canon(n, 4 * kc));
          *ph = h;
          *(ph + B_PLANE / 16) = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor cores
      } else {           // raw [k][n] -> canonical hi | lo: unit u = (n half, k chunk), lane = n: conflict-free reads and 16-byte stores
        // the planes of this stage were last read by the MMAs of tile kt - NST; waiting for tile kt - 2 (the newest completed-or-pending phase
        // of this barrier: an older phase must not be waited for by parity) covers it, the tensor pipe completes in order
        if (kt >= 2) mbar_wait(bar_free + 8 * buf, (uint32_t)(((kt - 2) >> 1) & 1));
        const float *raw = reinterpret_cast<const float *>(st + OFF_BR);
        const int wb = tb >> 5;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int u = wb * 4 + i, n = 32 * (u & 1) + lane, kc = u >> 1;
          float4 x;
          x.x = raw[(4 * kc) * BN + n]; x.y = raw[(4 * kc + 1) * BN + n]; x.z = raw[(4 * kc + 2) * BN + n]; x.w = raw[(4 * kc + 3) * BN + n];
          if (do_bias) bsum[i & 1] += (x.x + x.y) + (x.z + x.w);
          float4 h, l;
          split(x.x, h.x, l.x); split(x.y, h.y, l.y); split(x.z, h.z, l.z); split(x.w, h.w, l.w);
          float4 *ph = reinterpret_cast<float4 *>(st + canon(n, 4 * kc));
          *ph = h;
          *(ph + B_PLANE / 16) = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      }
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_ready + 8 * buf) : "memory");
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  // ---- epilogue: warp w reads TMEM lanes [32 (w & 3), +32) x columns [32 (w >> 2), +32), stages the C tile in shared memory (the B stages
  //      are free once every MMA has completed) and the CTA writes it out row by row: coalesced stores, coalesced reads of the mask operand
  float *Cz = C;
  if (EPI == EPI_PARTIAL) Cz = C + (int64_t)blockIdx.z * (int64_t)(M + (bias_row ? 1 : 0)) * ldc;
  if (n_kt > 0) {
    mbar_wait(bar_done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float *sC = reinterpret_cast<float *>(smb);
    if (worker) {
    uint32_t v[32];
    const int c0 = 32 * (w >> 2);
    const uint32_t taddr = tmem + ((uint32_t)(32 * (w & 3)) << 16) + TM_D + (uint32_t)c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
        "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
          "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
          "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const int r = 32 * (w & 3) + lane;
#pragma unroll
    for (int j = 0; j < 32; ++j) sC[r * LDC_S + c0 + j] = __uint_as_float(v[j]);
    }
    __syncthreads();
    // four elements per thread and round: the mask operand's loads (EPI_BWD_DATA) are all requested before the first one is used -- one
    // dependent global load per element made the data-gradient GEMM take twice its forward time
    for (int e0 = t; e0 < BM * BN; e0 += 4 * NTH) {
      float x[4], m[4];
      bool ok[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = e0 + u * NTH, row = e >> 6, col = e & 63;
        const int gr = m0 + row, gn = n0 + col;
        ok[u] = e < BM * BN && gr < M && gn < N;
        m[u] = 0.f;
        if (EPI == EPI_BWD_DATA && yprev && ok[u]) m[u] = __ldg(yprev + (int64_t)gr * ldc + gn);
        if (EPI == EPI_FWD && ok[u]) m[u] = __ldg(bias + gn);
        x[u] = ok[u] ? sC[row * LDC_S + col] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (!ok[u]) continue;
        const int e = e0 + u * NTH, row = e >> 6, col = e & 63;
        float v = x[u];
        if (EPI == EPI_FWD) v = act_fwd_rt(act, v + m[u]);
        if (EPI == EPI_BWD_DATA && yprev) v *= act_bwd_from_out(prev_act, m[u]);
        Cz[(int64_t)(m0 + row) * ldc + n0 + col] = v;
      }
    }
  } else if (EPI == EPI_PARTIAL && !skipped) {   // empty slab: zeros
    for (int e = t; e < BM * BN; e += NTH) {
      const int gr = m0 + e / BN, gn = n0 + e % BN;
      if (gr < M && gn < N) Cz[(int64_t)gr * ldc + gn] = 0.f;
    }
  }
  if (do_bias) {   // bias gradient of this slab: column sums of B over k (EPI_PARTIAL => !TB: B-side thread (wb, lane) holds columns lane and 32 + lane)
    float *sc = reinterpret_cast<float *>(smb + SM_BSUM);
    if (worker && !a_side) { sc[(tb >> 5) * BN + lane] = bsum[0]; sc[(tb >> 5) * BN + 32 + lane] = bsum[1]; }
    __syncthreads();
    if (t < BN && n0 + t < N) Cz[(int64_t)M * ldc + n0 + t] = (sc[t] + sc[BN + t]) + (sc[2 * BN + t] + sc[3 * BN + t]);   // fixed order: bit-reproducible
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (w == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TM_COLS) : "memory");
}

template <bool TA, bool TB, int EPI, int NST>
int launch_ns(crux_ctx *ctx, dim3 grid, const float *A, int lda, const float *B, int ldb, float *C, int ldc, int M, int N, int K, const float *bias, int act,
              const float *yprev, int prev_act, int k_per_slab, int bias_row, const int *skip) {
  static bool attr = false;
  if (!attr) {
    CRUX_CHECK_CUDA(ctx, cudaFuncSetAttribute(gemm_tc5_kernel<TA, TB, EPI, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<NST>::TOTAL));
    attr = true;
  }
  // 16-byte copies need rows that start 16-byte aligned (along k or along m / n, whichever is contiguous)
  const int vec_a = (lda % 4 == 0 && ((uintptr_t)A & 15) == 0) ? 1 : 0;
  const int vec_b = (ldb % 4 == 0 && ((uintptr_t)B & 15) == 0) ? 1 : 0;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = dim3(NTH); cfg.dynamicSmemBytes = Smem<NST>::TOTAL; cfg.stream = ctx->stream;
  cudaLaunchAttribute la[1];
  la[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  la[0].val.programmaticStreamSerializationAllowed = 1;
  static const bool pdl = getenv("CRUX_NO_PDL") == nullptr;
  cfg.attrs = la; cfg.numAttrs = pdl ? 1 : 0;
  CRUX_CHECK_CUDA(ctx, cudaLaunchKernelEx(&cfg, gemm_tc5_kernel<TA, TB, EPI, NST>, A, lda, B, ldb, C, ldc, M, N, K, bias, act, yprev, prev_act, k_per_slab, bias_row, skip,
                                          vec_a, vec_b));
  CRUX_LAUNCHED(ctx);
  return CRUX_OK;
}
template <bool TA, bool TB, int EPI>
int launch(crux_ctx *ctx, dim3 grid, const float *A, int lda, const float *B, int ldb, float *C, int ldc, int M, int N, int K, const float *bias, int act,
           const float *yprev, int prev_act, int k_per_slab, int bias_row, const int *skip) {
  // several waves of CTAs: two per SM (2 stages each) overlap each other's k-tile latencies; CRUX_G5_STAGES=2|4 forces a variant (A/B runs)
  const char *env = getenv("CRUX_G5_STAGES");
  const int64_t ctas = (int64_t)grid.x * grid.y * grid.z;
  const bool two = env ? env[0] == '2' : ctas > (int64_t)ctx->num_sms;
  if (two) return launch_ns<TA, TB, EPI, 2>(ctx, grid, A, lda, B, ldb, C, ldc, M, N, K, bias, act, yprev, prev_act, k_per_slab, bias_row, skip);
  return launch_ns<TA, TB, EPI, 4>(ctx, grid, A, lda, B, ldb, C, ldc, M, N, K, bias, act, yprev, prev_act, k_per_slab, bias_row, skip);
}

}  // namespace

// Default for shapes that fill a tile (CRUX_GEMM_TC5=0 selects the FFMA tile kernel: read per call, A/B runs and tests switch it).
// Measured on B200 (scripts/gemm_tc5_bench.py, scripts/sac_launches.py --time): 2048 x 256 x 2048 forward 57 us against 130 us (FFMA tiles),
// 16384 x 256 x 256 31 against 70 us (two 2-stage CTAs per SM: 69 TFLOP/s of fp32-accurate GEMM); SAC 376/17/256-256 at B = 2048: 0.81 ms per update against 1.17 ms.  The first version of this kernel
// (both operands hi | lo through shared memory, lock-step phases) was SLOWER than the FFMA tiles (1.60 ms): profiles/r2_notes.md.
bool gemm_tc5_eligible(int64_t M, int64_t N, int64_t K) {
  const char *on = getenv("CRUX_GEMM_TC5");
  if (on && on[0] == '0') return false;
  if (on && on[0] == '1') return M >= 64 && N >= 16 && K >= 16;
  return M >= 128 && N >= 32 && K >= 32;
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
