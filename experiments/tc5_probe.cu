// Round-2 probe for the transposed ("features on TMEM lanes") formulation of the PPO minibatch kernel:
//   D^T[M = 64 features][N rows] = A[M][K] * B[N][K]^T, tf32, K-major canonical no-swizzle operands.
// Checks (against a CPU reference) and times (clock64 around a chain of R accumulating MMAs, single issuing thread):
//   * M = 64 accumulators at data-path offset 0 and 16 (two interleaved 16-lane atoms per 32-lane quarter)
//   * A from shared memory (SS) and A from tensor memory (TS) with the M = 64 lane layout lane = 32*(m/16) + dp + m%16
//   * padded K-major layout (LBO = 144 B between K-adjacent core matrices) for conflict-free transposed stores
//   * per-instruction cost of tcgen05.mma for M in {64, 128} x N in {8 .. 256}
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o experiments/tc5_probe experiments/tc5_probe.cu && ./experiments/tc5_probe
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(da), "l"(db),
               "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(db),
               "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ float tf32_rn(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }

struct Case {
  int M, N, K;      // MMA shape; K = total contraction (multiple of 8, <= 64)
  int a_tmem;       // A operand from tensor memory
  int dp;           // data-path offset of the M = 64 atom (0 or 16)
  int lbo;          // bytes between K-adjacent core matrices (128 = dense, 144 = padded)
  int reps;         // timing: number of extra accumulating MMAs of the first k-step issued back to back (0 = correctness only)
};

// A [M][K], B [N][K] plain row-major fp32 in global memory (already tf32-representable); D out [128 lanes][N]
__global__ void __launch_bounds__(128) probe(const float *__restrict__ A, const float *__restrict__ B, float *__restrict__ D, long long *cycles, Case c) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base;
  const int t = threadIdx.x, w = t >> 5, l = t & 31;
  const int sboA = (c.K / 4) * c.lbo, sboB = (c.K / 4) * c.lbo;
  unsigned char *sA = smem, *sB = smem + (c.M / 8) * sboA + 1024;
  auto off = [&](int row, int k, int sbo) { return (row >> 3) * sbo + (k >> 2) * c.lbo + (row & 7) * 16 + (k & 3) * 4; };
  for (int e = t; e < c.M * c.K; e += 128) *reinterpret_cast<float *>(sA + off(e / c.K, e % c.K, sboA)) = A[e];
  for (int e = t; e < c.N * c.K; e += 128) *reinterpret_cast<float *>(sB + off(e / c.K, e % c.K, sboB)) = B[e];
  if (t == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (w == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_base, lane_addr = tm + ((uint32_t)(32 * w) << 16);
  constexpr uint32_t ACOL = 256;
  // marker in the accumulator columns; A operand (TS): this thread's lane holds feature m of the atom at data-path offset dp
  for (int c0 = 0; c0 < c.N; c0 += 8) {
    const uint32_t mk = __float_as_uint(-12345.0f);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(lane_addr + c0), "r"(mk) : "memory");
  }
  if (c.a_tmem) {
    int m = -1;
    if (c.M == 128) m = 32 * w + l;
    else if (l >= c.dp && l < c.dp + 16) m = 16 * w + (l - c.dp);
    for (int k0 = 0; k0 < c.K; k0 += 8) {
      uint32_t v[8];
      for (int j = 0; j < 8; ++j) v[j] = __float_as_uint(m >= 0 ? A[m * c.K + k0 + j] : 777.f);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(lane_addr + ACOL + k0), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                   "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                   : "memory");
    }
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (t == 0) {
    const uint32_t idesc = make_idesc(c.M, c.N);
    const uint32_t d_addr = tm + ((uint32_t)c.dp << 16), a_addr = tm + ((uint32_t)c.dp << 16) + ACOL;
    const uint64_t da = make_desc(smem_u32(sA), c.lbo, sboA), db = make_desc(smem_u32(sB), c.lbo, sboB);
    const uint32_t adv = (2 * c.lbo) >> 4;   // one MMA consumes K = 8 = two core matrices along K
    const long long t0 = clock64();
    for (int ks = 0; ks < c.K / 8; ++ks) {
      if (c.a_tmem) mma_ts(d_addr, a_addr + 8 * ks, db + (uint64_t)adv * ks, idesc, ks ? 1u : 0u);
      else mma_ss(d_addr, da + (uint64_t)adv * ks, db + (uint64_t)adv * ks, idesc, ks ? 1u : 0u);
    }
    // timing chain: scratch accumulator in columns [256 + 64, ...) would collide with A; use a second D region at column N (<= 256 total)
    const uint32_t d2 = d_addr + (c.reps ? (uint32_t)c.N : 0u);
    for (int r = 0; r < c.reps; ++r) {
      if (c.a_tmem) mma_ts(d2, a_addr, db, idesc, 1u);
      else mma_ss(d2, da, db, idesc, 1u);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    uint32_t done = 0;
    while (!done)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0) : "memory");
    *cycles = clock64() - t0;
  }
  __syncthreads();
  uint32_t done = 0;
  while (!done)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < c.N; c0 += 8) {
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(lane_addr + c0)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) D[(32 * w + l) * c.N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}

static float tf32_host(float x) {
  uint32_t u; memcpy(&u, &x, 4); u = (u + 0x1000u) & 0xFFFFE000u; memcpy(&x, &u, 4); return x;
}

static int run(const char *name, Case c) {
  std::vector<float> hA(c.M * c.K), hB(c.N * c.K), hD(128 * c.N, 0.f);
  srand(7);
  for (auto &x : hA) x = tf32_host((float)rand() / RAND_MAX * 2.f - 1.f);
  for (auto &x : hB) x = tf32_host((float)rand() / RAND_MAX * 2.f - 1.f);
  float *dA, *dB, *dD; long long *dC, hC = 0;
  cudaMalloc(&dA, hA.size() * 4); cudaMalloc(&dB, hB.size() * 4); cudaMalloc(&dD, hD.size() * 4); cudaMalloc(&dC, 8);
  cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = (size_t)(c.M / 8 + c.N / 8 + 2) * (c.K / 4) * c.lbo + 4096;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  double best_cyc = 1e30;
  for (int it = 0; it < 3; ++it) {
    probe<<<1, 128, smem>>>(dA, dB, dD, dC, c);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-44s CUDA error %s\n", name, cudaGetErrorString(e)); return 2; }
    cudaMemcpy(&hC, dC, 8, cudaMemcpyDeviceToHost);
    best_cyc = fmin(best_cyc, (double)hC);
  }
  cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
  double max_err = 0, max_ref = 0;
  int untouched_ok = 1;
  for (int m = 0; m < c.M; ++m) {
    const int lane = c.M == 128 ? m : 32 * (m / 16) + c.dp + m % 16;
    for (int n = 0; n < c.N; ++n) {
      double ref = 0;
      for (int k = 0; k < c.K; ++k) ref += (double)hA[m * c.K + k] * (double)hB[n * c.K + k];
      max_err = fmax(max_err, fabs(ref - hD[lane * c.N + n]));
      max_ref = fmax(max_ref, fabs(ref));
    }
  }
  if (c.M == 64)   // the other 16 lanes of every quarter must still hold the marker
    for (int q = 0; q < 4; ++q)
      for (int i = 0; i < 16; ++i)
        if (hD[(32 * q + (16 - c.dp) + i) * c.N] != -12345.0f) untouched_ok = 0;
  const double rel = max_err / fmax(max_ref, 1e-30);
  printf("%-44s rel err %.2e  other-atom lanes untouched: %s", name, rel, c.M == 64 ? (untouched_ok ? "yes" : "NO") : "-");
  if (c.reps) printf("   %.1f cycles / MMA (chain of %d, total %.0f)", best_cyc / (c.reps + c.K / 8), c.reps, best_cyc);
  printf("\n");
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC);
  return (rel < 1e-5 && untouched_ok) ? 0 : 1;
}

int main() {
  int rc = 0;
  //                                                        M    N   K  a_tmem dp  lbo reps
  rc |= run("M128 N64  K64 SS", Case{128, 64, 64, 0, 0, 128, 0});
  rc |= run("M64  N64  K64 SS dp0", Case{64, 64, 64, 0, 0, 128, 0});
  rc |= run("M64  N64  K64 SS dp16", Case{64, 64, 64, 0, 16, 128, 0});
  rc |= run("M64  N32  K24 SS dp16", Case{64, 32, 24, 0, 16, 128, 0});
  rc |= run("M64  N128 K64 SS dp0", Case{64, 128, 64, 0, 0, 128, 0});
  rc |= run("M64  N64  K64 TS dp0", Case{64, 64, 64, 1, 0, 128, 0});
  rc |= run("M64  N64  K64 TS dp16", Case{64, 64, 64, 1, 16, 128, 0});
  rc |= run("M64  N8   K32 TS dp16", Case{64, 8, 32, 1, 16, 128, 0});
  rc |= run("M128 N64  K64 TS", Case{128, 64, 64, 1, 0, 128, 0});
  rc |= run("M64  N64  K64 SS dp0 LBO144", Case{64, 64, 64, 0, 0, 144, 0});
  rc |= run("M64  N32  K64 TS dp16 LBO144", Case{64, 32, 64, 1, 16, 144, 0});
  printf("---- timing (single issuing thread, accumulating chain; includes one commit + mbarrier round trip)\n");
  const int Ns[] = {8, 16, 32, 64, 128, 256};
  for (int M : {64, 128})
    for (int N : Ns) {
      if (M == 128 && N % 16) continue;
      if (2 * N > 256 && N != 256) {}
      char nm[64];
      for (int ts = 0; ts < 2; ++ts) {
        if (N == 256) continue;   // the scratch accumulator sits at column N: keep 2N <= 256
        snprintf(nm, sizeof nm, "M%-3d N%-3d K8 %s x512", M, N, ts ? "TS" : "SS");
        run(nm, Case{M, N, 8, ts, 0, 128, 512});
      }
    }
  printf(rc ? "FAIL\n" : "PASS\n");
  return rc;
}
