// value(π, s) over a whole rollout column on the 5th-generation tensor cores (tcgen05 + TMEM), 3xTF32 split accumulation.
//
// A forward pass only has row GEMMs (contraction over features), i.e. every operand is K-major -- the form tcgen05 accepts for tf32
// without swizzled transposes (experiments/tcgen05_layouts_test.cu).  Per 128-row tile (one CTA per SM, persistent):
//
//   gather   : the tile's rows arrive through cp.async one tile ahead, are split x = hi + lo (hi = rna_tf32(x)) and stored as two
//              canonical no-swizzle K-major planes (core matrix = 8 rows x 16 B, cores adjacent along K)
//   layer 1  : ONE thread issues 3 x 3 tcgen05.mma (M = 128, N = 64, K = 8 each; lo*hi + hi*lo + hi*hi) into TMEM columns [0, 64)
//   epilogue : all 8 warps read their TMEM lanes (tcgen05.ld 32x32b: lane = row), add the bias, tanh, split, and store h1 hi/lo as
//              the next layer's A operand -- the activation never exists anywhere else
//   layer 2  : 3 x 8 tcgen05.mma into TMEM columns [64, 128), same epilogue into the h2 planes
//   output   : 3 x 8 tcgen05.mma with N = 16 (outputs >= O are zero columns) into TMEM columns [128, 144); warps 0..3 read one
//              column and write V(s)
// MMA completion is tracked with tcgen05.commit on one mbarrier (phase parity alternates per GEMM); generic-proxy stores are made
// visible to the tensor cores with fence.proxy.async before each hand-off.  Weights are converted once per CTA into canonical hi/lo
// planes: W1 as [64 n][24 k] (k >= I zero), W2 as [64 n][64 k], W3 as [16 n][64 k].
//
// Included by ppo_fused.cu (inside its anonymous namespace: H, NT, MAX_I, NetDesc, stage_params, tanh_fast, smem_u32, off_* are
// defined there) and by nothing else.
#pragma once

namespace tc5 {

constexpr int TR = 128;                 // rows per tile = TMEM lanes
constexpr int KX = 24;                  // layer-1 K (input width padded to a multiple of 8; MAX_I for this path is 24)
constexpr int NOUT = 16;                // N of the output GEMM (M = 128 needs N % 16 == 0)
constexpr int TMEM_COLS = 256;          // z1 [0,64) | z2 [64,128) | out [128,144)

__device__ __forceinline__ int canon(int row, int k, int K) { return (row >> 3) * (32 * K) + (k >> 2) * 128 + (row & 7) * 16 + (k & 3) * 4; }  // bytes
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // sm_100 descriptor version; no swizzle
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
__device__ __forceinline__ bool elect_one() {   // one lane of a converged warp
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void split(float x, float &hi, float &lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
  lo = x - hi;
}

// shared-memory map (bytes).  The raw parameter vector is staged into the activation planes (dead during the prologue).
struct Map {
  static constexpr int W1 = 0;                                  // hi | lo planes, [64][24] canonical: 2 x 6144
  static constexpr int W2 = W1 + 2 * 64 * KX * 4;               // 2 x 16384
  static constexpr int W3 = W2 + 2 * 64 * 64 * 4;               // [16][64]: 2 x 4096
  static constexpr int XA = W3 + 2 * NOUT * 64 * 4;             // [128][24]: 2 x 12288
  static constexpr int HA = XA + 2 * TR * KX * 4;               // [128][64]: 2 x 32768   (prologue: raw parameters)
  static constexpr int HB = HA + 2 * TR * 64 * 4;               // 2 x 32768
  static constexpr int ST = HB + 2 * TR * 64 * 4;               // gather staging: [128][I] raw floats (<= 128 x 24 x 4)
  static constexpr int BIAS = ST + TR * KX * 4;                 // b1[64] b2[64] b3[16]
  static constexpr int BAR = BIAS + (64 + 64 + 16) * 4;         // mbarrier (MMA completion), mbarrier (parameter TMA), tmem base
  static constexpr int TOTAL = BAR + 32;
  static_assert(HA % 1024 == 0 || HA % 16 == 0, "alignment");
};

struct Args {
  NetDesc net; const float *x; int64_t B; float *y;
  // crux_value_next (forward_kernel_tmem only): tiles whose rows equal x_alt bit for bit copy y_alt instead of running the network
  const float *x_alt; const float *y_alt; int64_t alt_rows;
};

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

// one GEMM = 3 passes x ksteps MMAs; a/b: shared addresses of the hi planes, *_lo = offset of the lo plane
__device__ __forceinline__ void issue_gemm(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo_off, uint32_t a_sbo, uint32_t b_hi, uint32_t b_lo_off, uint32_t b_sbo,
                                           int ksteps, uint32_t idesc, uint32_t bar) {
  int first = 1;
#pragma unroll 1
  for (int p = 0; p < 3; ++p) {   // small terms first: lo*hi, hi*lo, hi*hi
    const uint32_t a = a_hi + (p == 0 ? a_lo_off : 0), b = b_hi + (p == 1 ? b_lo_off : 0);
#pragma unroll 1
    for (int ks = 0; ks < ksteps; ++ks) {
      mma_tf32_ss(d_tmem, make_desc(a + ks * 256, 128, a_sbo), make_desc(b + ks * 256, 128, b_sbo), idesc, first ? 0u : 1u);
      first = 0;
    }
  }
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// epilogue of a hidden layer: this thread's row, 32 columns starting at c0: act(z + b) -> hi/lo planes of the next A operand
__device__ __forceinline__ void hidden_epilogue(uint32_t taddr, const float *__restrict__ bias, int act, unsigned char *__restrict__ plane_hi, int row, int c0) {
  uint32_t v[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
      "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
        "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
        "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  unsigned char *lo_plane = plane_hi + TR * 64 * 4;
#pragma unroll
  for (int q = 0; q < 8; ++q) {   // 4 consecutive features = one 16-byte chunk of a core-matrix row
    float4 hi, lo;
    split(act_fused(act, __uint_as_float(v[4 * q + 0]) + bias[c0 + 4 * q + 0]), hi.x, lo.x);
    split(act_fused(act, __uint_as_float(v[4 * q + 1]) + bias[c0 + 4 * q + 1]), hi.y, lo.y);
    split(act_fused(act, __uint_as_float(v[4 * q + 2]) + bias[c0 + 4 * q + 2]), hi.z, lo.z);
    split(act_fused(act, __uint_as_float(v[4 * q + 3]) + bias[c0 + 4 * q + 3]), hi.w, lo.w);
    const int off = canon(row, c0 + 4 * q, 64);
    *reinterpret_cast<float4 *>(plane_hi + off) = hi;
    *reinterpret_cast<float4 *>(lo_plane + off) = lo;
  }
}

__global__ void __launch_bounds__(NT, 1) forward_kernel(Args a) {
  extern __shared__ __align__(1024) unsigned char smb[];
  const NetDesc nd = a.net;
  const int I = nd.I, O = nd.O, act = nd.act;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const uint32_t bar_mma = smem_u32(smb + Map::BAR), bar_par = smem_u32(smb + Map::BAR + 8);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smb + Map::BAR + 16);
  float *bias = reinterpret_cast<float *>(smb + Map::BIAS);
  const int64_t n_tiles = (a.B + TR - 1) / TR;
  const int chunks = TR * I / 4;   // 16-byte chunks of one tile (128 * I is a multiple of 4)
  auto issue_tile = [&](int64_t tile) {
    const int64_t f0 = tile * TR * I, f_end = a.B * I;
    for (int c = t; c < chunks; c += NT) {
      const int64_t f = f0 + 4 * c, left = f_end - f;
      const int nbytes = left >= 4 ? 16 : (left > 0 ? (int)left * 4 : 0);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smb + Map::ST + 16 * c)), "l"(a.x + (nbytes ? f : 0)), "r"(nbytes) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  issue_tile(blockIdx.x);
  // ---- prologue: raw parameters -> HA region (TMA bulk), TMEM allocation, zero the x planes (their padding columns stay zero)
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_mma), "r"(1) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_par), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  for (int e = t; e < 2 * TR * KX; e += NT) reinterpret_cast<float *>(smb + Map::XA)[e] = 0.f;
  __syncthreads();
  if (t == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_par), "r"(nd.bytes16) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smb + Map::HA)), "l"(nd.params),
                 "r"(nd.bytes16), "r"(bar_par)
                 : "memory");
  }
  if (w == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  mbar_wait(bar_par, 0);
  {  // canonical hi/lo weight planes + biases
    const float *P = reinterpret_cast<const float *>(smb + Map::HA);
    const float *W1 = P, *b1 = P + off_b1(I), *W2 = P + off_W2(I), *b2 = P + off_b2(I), *W3 = P + off_W3(I), *b3 = P + off_b3(I, O);
    for (int e = t; e < 64 * KX; e += NT) {       // (n = j, k = i)
      const int j = e / KX, i = e - j * KX;
      float hi, lo;
      split(i < I ? W1[i * H + j] : 0.f, hi, lo);
      const int off = canon(j, i, KX);
      *reinterpret_cast<float *>(smb + Map::W1 + off) = hi;
      *reinterpret_cast<float *>(smb + Map::W1 + 64 * KX * 4 + off) = lo;
    }
    for (int e = t; e < 64 * 64; e += NT) {       // (n = j, k): B[n][k] = W2[k][j]; e = k * 64 + j reads W2 coalesced
      const int k = e >> 6, j = e & 63;
      float hi, lo;
      split(W2[e], hi, lo);
      const int off = canon(j, k, 64);
      *reinterpret_cast<float *>(smb + Map::W2 + off) = hi;
      *reinterpret_cast<float *>(smb + Map::W2 + 64 * 64 * 4 + off) = lo;
    }
    for (int e = t; e < NOUT * 64; e += NT) {     // (n = o, k): B[o][k] = W3[k][o]
      const int o = e >> 6, k = e & 63;
      float hi, lo;
      split(o < O ? W3[k * O + o] : 0.f, hi, lo);
      const int off = canon(o, k, 64);
      *reinterpret_cast<float *>(smb + Map::W3 + off) = hi;
      *reinterpret_cast<float *>(smb + Map::W3 + NOUT * 64 * 4 + off) = lo;
    }
    if (t < 64) { bias[t] = b1[t]; bias[64 + t] = b2[t]; }
    if (t < NOUT) bias[128 + t] = t < O ? b3[t] : 0.f;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  const uint32_t idesc64 = make_idesc(TR, 64), idesc16 = make_idesc(TR, NOUT);
  const uint32_t sW1 = smem_u32(smb + Map::W1), sW2 = smem_u32(smb + Map::W2), sW3 = smem_u32(smb + Map::W3);
  const uint32_t sXA = smem_u32(smb + Map::XA), sHA = smem_u32(smb + Map::HA), sHB = smem_u32(smb + Map::HB);
  const int row = 32 * (w & 3) + lane, c0 = 32 * (w >> 2);                       // epilogue coordinates: TMEM lane = row, column half
  const uint32_t lane_addr = tmem + ((uint32_t)(32 * (w & 3)) << 16);
  const uint32_t inv_I = (65536u + (uint32_t)I - 1u) / (uint32_t)I;
  uint32_t ph = 0;   // parity of the next MMA completion

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();   // the staged rows are visible; every MMA of the previous tile has completed (its last wait is behind us)
    for (int e = t; e < TR * I; e += NT) {
      const int r = (int)(((uint32_t)e * inv_I) >> 16), i = e - r * I;
      float hi, lo;
      split(reinterpret_cast<const float *>(smb + Map::ST)[e], hi, lo);
      const int off = canon(r, i, KX);
      *reinterpret_cast<float *>(smb + Map::XA + off) = hi;
      *reinterpret_cast<float *>(smb + Map::XA + TR * KX * 4 + off) = lo;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tile + gridDim.x < n_tiles) issue_tile(tile + gridDim.x);   // the staging area has been consumed
    // ---------------- layer 1
    if (w == 0) {   // warp-uniform branch + elect.sync: under `if (t == 0)` nvcc wraps every MMA in an election loop (96 instead of 24-48 cycles each)
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) issue_gemm(tmem + 0, sXA, TR * KX * 4, 32 * KX, sW1, 64 * KX * 4, 32 * KX, KX / 8, idesc64, bar_mma);
      __syncwarp();
    }
    mbar_wait(bar_mma, ph); ph ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    hidden_epilogue(lane_addr + 0 + c0, bias, act, smb + Map::HA, row, c0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    // ---------------- layer 2
    if (w == 0) {   // warp-uniform branch + elect.sync: under `if (t == 0)` nvcc wraps every MMA in an election loop (96 instead of 24-48 cycles each)
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) issue_gemm(tmem + 64, sHA, TR * 64 * 4, 32 * 64, sW2, 64 * 64 * 4, 32 * 64, 8, idesc64, bar_mma);
      __syncwarp();
    }
    mbar_wait(bar_mma, ph); ph ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    hidden_epilogue(lane_addr + 64 + c0, bias + 64, act, smb + Map::HB, row, c0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    // ---------------- output layer
    if (w == 0) {   // warp-uniform branch + elect.sync: under `if (t == 0)` nvcc wraps every MMA in an election loop (96 instead of 24-48 cycles each)
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) issue_gemm(tmem + 128, sHB, TR * 64 * 4, 32 * 64, sW3, NOUT * 64 * 4, 32 * 64, 8, idesc16, bar_mma);
      __syncwarp();
    }
    mbar_wait(bar_mma, ph); ph ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (w < 4) {
      uint32_t v[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(lane_addr + 128)
                   : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const int64_t grow = tile * TR + row;
      if (grow < a.B)
        for (int o = 0; o < O; ++o) a.y[grow * O + o] = __uint_as_float(v[o]) + bias[128 + o];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
}


// ===================================================================================================================================
// Variant 2: the A operand lives in TENSOR MEMORY.  The epilogue thread that owns a row (TMEM lane) writes the split activation
// straight back into TMEM (tcgen05.st) and the next layer's MMA reads it from there (tcgen05.mma with [a_tmem]): activations never
// touch shared memory, no canonical-layout stores, no proxy fences.  Shared memory only holds the weights (52 KB) and the gather
// staging, so TWO CTAs share an SM (256 TMEM columns each) and overlap each other's MMA / epilogue phases.
//   TMEM columns of a CTA:  [0,24) x hi | [24,48) x lo | C1 [64,128) z1 -> h1 hi, later the output accumulator | C2 [128,192) h1 lo -> h2 lo
//                           | C3 [192,256) z2 -> h2 hi
struct Map2 {
  static constexpr int W1 = 0;
  static constexpr int W2 = W1 + 2 * 64 * KX * 4;
  static constexpr int W3 = W2 + 2 * 64 * 64 * 4;
  static constexpr int ST = W3 + 2 * NOUT * 64 * 4;             // gather staging [128][I] raw floats; prologue: raw parameters (27.2 KB max)
  static constexpr int ST_BYTES = 28 * 1024;
  static constexpr int BIAS = ST + ST_BYTES;
  static constexpr int BAR = BIAS + (64 + 64 + 16) * 4;
  static constexpr int TOTAL = BAR + 32;
};
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem), "r"(a_tmem),
               "l"(db), "r"(idesc), "r"(acc)
               : "memory");
}
// 3 passes x KS MMAs, A = TMEM columns (8 per k-step), B = canonical K-major shared-memory planes.  Fully unrolled: the descriptor of
// k-step ks is the base descriptor + 16 * ks (the 14-bit start-address field counts 16-byte units; no carry for < 256 KB of smem), so
// the single issuing thread spends ~3 instructions per MMA instead of rebuilding 64-bit descriptors.
template <int KS>
__device__ __forceinline__ void issue_gemm_ts(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo_off, uint32_t b_sbo, uint32_t idesc,
                                              uint32_t bar) {
  const uint64_t dbh = make_desc(b_hi, 128, b_sbo), dbl = make_desc(b_hi + b_lo_off, 128, b_sbo);
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) mma_tf32_ts(d_tmem, a_lo + 8 * ks, dbh + 16 * ks, idesc, ks ? 1u : 0u);   // lo * hi
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) mma_tf32_ts(d_tmem, a_hi + 8 * ks, dbl + 16 * ks, idesc, 1u);             // hi * lo
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) mma_tf32_ts(d_tmem, a_hi + 8 * ks, dbh + 16 * ks, idesc, 1u);             // hi * hi
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
#define TC5_LD32(v, taddr)                                                                                                                             \
  asm volatile(                                                                                                                                          \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, " \
      "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                                                                                                  \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),     \
        "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),    \
        "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                              \
      : "r"(taddr)                                                                                                                                       \
      : "memory")
#define TC5_ST32(taddr, v)                                                                                                                             \
  asm volatile(                                                                                                                                          \
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, " \
      "%24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),                                                                                      \
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]),      \
      "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]),    \
      "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])                                                               \
      : "memory")
#define TC5_ST8(taddr, v)                                                                                                                              \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),  \
               "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])                                                                                             \
               : "memory")

// z (32 accumulator columns of this thread's row, at acc_addr) -> act(z + b): hi overwrites the accumulator columns, lo goes to lo_addr
__device__ __forceinline__ void hidden_epilogue_tmem(uint32_t acc_addr, uint32_t lo_addr, const float *__restrict__ bias, int act) {
  uint32_t v[32], l[32];
  TC5_LD32(v, acc_addr);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    float hi, lo;
    split(act_fused(act, __uint_as_float(v[j]) + bias[j]), hi, lo);
    v[j] = __float_as_uint(hi); l[j] = __float_as_uint(lo);
  }
  TC5_ST32(acc_addr, v);
  TC5_ST32(lo_addr, l);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(NT, 2) forward_kernel_tmem(Args a) {
  extern __shared__ __align__(1024) unsigned char smb[];
  const NetDesc nd = a.net;
  const int I = nd.I, O = nd.O, act = nd.act;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const uint32_t bar_mma = smem_u32(smb + Map2::BAR), bar_par = smem_u32(smb + Map2::BAR + 8);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smb + Map2::BAR + 16);
  float *bias = reinterpret_cast<float *>(smb + Map2::BIAS);
  const int64_t n_tiles = (a.B + TR - 1) / TR;
  const int chunks = TR * I / 4;
  auto has_alt = [&](int64_t tile) { return a.x_alt != nullptr && (tile + 1) * TR <= a.alt_rows; };   // every row of the tile has a successor row
  constexpr int ST2 = TR * KX * 4;   // second staging tile (the x_alt rows)
  auto issue_tile = [&](int64_t tile) {
    const int64_t f0 = tile * TR * I, f_end = a.B * I;
    const bool alt = has_alt(tile);
    for (int c = t; c < chunks; c += NT) {
      const int64_t f = f0 + 4 * c, left = f_end - f;
      const int nbytes = left >= 4 ? 16 : (left > 0 ? (int)left * 4 : 0);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smb + Map2::ST + 16 * c)), "l"(a.x + (nbytes ? f : 0)), "r"(nbytes) : "memory");
      if (alt) asm volatile("cp.async.cg.shared.global [%0], [%1], 16, 16;" ::"r"(smem_u32(smb + Map2::ST + ST2 + 16 * c)), "l"(a.x_alt + f) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // ---- prologue: raw parameters -> staging area (TMA bulk), TMEM allocation, canonical hi/lo weight planes
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_mma), "r"(1) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_par), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (t == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_par), "r"(nd.bytes16) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smb + Map2::ST)), "l"(nd.params),
                 "r"(nd.bytes16), "r"(bar_par)
                 : "memory");
  }
  if (w == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  mbar_wait(bar_par, 0);
  {
    const float *P = reinterpret_cast<const float *>(smb + Map2::ST);
    const float *W1 = P, *b1 = P + off_b1(I), *W2 = P + off_W2(I), *b2 = P + off_b2(I), *W3 = P + off_W3(I), *b3 = P + off_b3(I, O);
    for (int e = t; e < 64 * KX; e += NT) {
      const int j = e / KX, i = e - j * KX;
      float hi, lo;
      split(i < I ? W1[i * H + j] : 0.f, hi, lo);
      const int off = canon(j, i, KX);
      *reinterpret_cast<float *>(smb + Map2::W1 + off) = hi;
      *reinterpret_cast<float *>(smb + Map2::W1 + 64 * KX * 4 + off) = lo;
    }
    for (int e = t; e < 64 * 64; e += NT) {
      const int k = e >> 6, j = e & 63;
      float hi, lo;
      split(W2[e], hi, lo);
      const int off = canon(j, k, 64);
      *reinterpret_cast<float *>(smb + Map2::W2 + off) = hi;
      *reinterpret_cast<float *>(smb + Map2::W2 + 64 * 64 * 4 + off) = lo;
    }
    for (int e = t; e < NOUT * 64; e += NT) {
      const int o = e >> 6, k = e & 63;
      float hi, lo;
      split(o < O ? W3[k * O + o] : 0.f, hi, lo);
      const int off = canon(o, k, 64);
      *reinterpret_cast<float *>(smb + Map2::W3 + off) = hi;
      *reinterpret_cast<float *>(smb + Map2::W3 + NOUT * 64 * 4 + off) = lo;
    }
    if (t < 64) { bias[t] = b1[t]; bias[64 + t] = b2[t]; }
    if (t < NOUT) bias[128 + t] = t < O ? b3[t] : 0.f;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();   // the raw parameters have been consumed: the staging area is free for the gather
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  issue_tile(blockIdx.x);
  const uint32_t tmem = *tmem_slot;
  const uint32_t idesc64 = make_idesc(TR, 64), idesc16 = make_idesc(TR, NOUT);
  const uint32_t sW1 = smem_u32(smb + Map2::W1), sW2 = smem_u32(smb + Map2::W2), sW3 = smem_u32(smb + Map2::W3);
  const int row = 32 * (w & 3) + lane, c0 = 32 * (w >> 2);
  const uint32_t lane_addr = tmem + ((uint32_t)(32 * (w & 3)) << 16);
  constexpr uint32_t XH = 0, XL = 24, C1 = 64, C2 = 128, C3 = 192;
  uint32_t ph = 0;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();   // staged rows visible; the previous tile's last MMA has completed
    if (has_alt(tile)) {   // bitwise comparison of the staged tile with the staged x_alt tile (uniform outcome for the CTA)
      int diff = 0;
      const int *p0 = reinterpret_cast<const int *>(smb + Map2::ST), *p1 = reinterpret_cast<const int *>(smb + Map2::ST + ST2);
      for (int e = t; e < TR * I; e += NT) diff |= p0[e] != p1[e];
      if (__syncthreads_or(diff) == 0) {
        if (t < TR) a.y[tile * TR + t] = a.y_alt[tile * TR + t];   // one output on this path (checked by the launcher)
        __syncthreads();
        if (tile + gridDim.x < n_tiles) issue_tile(tile + gridDim.x);
        continue;
      }
    }
    if (w < 4) {       // thread = row: the row's I inputs (zero-padded to 24) -> x hi / x lo columns
      const float *xr = reinterpret_cast<const float *>(smb + Map2::ST) + row * I;
#pragma unroll
      for (int q = 0; q < KX / 8; ++q) {
        uint32_t h8[8], l8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int i = 8 * q + j;
          float hi, lo;
          split(i < I ? xr[i] : 0.f, hi, lo);
          h8[j] = __float_as_uint(hi); l8[j] = __float_as_uint(lo);
        }
        TC5_ST8(lane_addr + XH + 8 * q, h8);
        TC5_ST8(lane_addr + XL + 8 * q, l8);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tile + gridDim.x < n_tiles) issue_tile(tile + gridDim.x);
    // ---------------- layer 1: C1 = x W1
    if (w == 0) {   // warp-uniform branch + elect.sync: under `if (t == 0)` nvcc wraps every MMA in an election loop (96 instead of 24-48 cycles each)
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) issue_gemm_ts<KX / 8>(tmem + C1, tmem + XH, tmem + XL, sW1, 64 * KX * 4, 32 * KX, idesc64, bar_mma);
      __syncwarp();
    }
    mbar_wait(bar_mma, ph); ph ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    hidden_epilogue_tmem(lane_addr + C1 + c0, lane_addr + C2 + c0, bias + c0, act);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    // ---------------- layer 2: C3 = h1 W2
    if (w == 0) {   // warp-uniform branch + elect.sync: under `if (t == 0)` nvcc wraps every MMA in an election loop (96 instead of 24-48 cycles each)
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) issue_gemm_ts<8>(tmem + C3, tmem + C1, tmem + C2, sW2, 64 * 64 * 4, 32 * 64, idesc64, bar_mma);
      __syncwarp();
    }
    mbar_wait(bar_mma, ph); ph ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    hidden_epilogue_tmem(lane_addr + C3 + c0, lane_addr + C2 + c0, bias + 64 + c0, act);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    // ---------------- output layer: C1[0,16) = h2 W3
    if (w == 0) {   // warp-uniform branch + elect.sync: under `if (t == 0)` nvcc wraps every MMA in an election loop (96 instead of 24-48 cycles each)
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) issue_gemm_ts<8>(tmem + C1, tmem + C3, tmem + C2, sW3, NOUT * 64 * 4, 32 * 64, idesc16, bar_mma);
      __syncwarp();
    }
    mbar_wait(bar_mma, ph); ph ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (w < 4) {
      uint32_t v[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(lane_addr + C1)
                   : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const int64_t grow = tile * TR + row;
      if (grow < a.B)
        for (int o = 0; o < O; ++o) a.y[grow * O + o] = __uint_as_float(v[o]) + bias[128 + o];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
}

}  // namespace tc5
