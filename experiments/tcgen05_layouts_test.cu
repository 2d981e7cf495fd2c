// Bring-up test for every tcgen05 operand form the fused PPO minibatch kernel relies on (3xTF32, no swizzle):
//   T1  K-major A and B with K = 24 (3 MMAs per pass), M = 128, N = 64                       (layer-1 forward)
//   T2  the SAME K-major-stored activation tiles used as MN-major A and B (A^T B over K = 128 rows):
//         M = 64 (where do the 64 rows land in TMEM?) and M = 128 with the M-side over-reading its buffer (weight gradients)
//   T3  N = 16 with an MN-major B that is a K-major-stored [64][8] tile (output layer), and K-major use of the same tile with K = 8
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o experiments/tcgen05_layouts_test experiments/tcgen05_layouts_test.cu
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
__device__ __forceinline__ float tf32_hi(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// canonical K-major no-swizzle tile [rows][K]: core = 8 rows x 16 B; cores adjacent along K contiguous (128 B), 8-row groups at 128*K/4
__device__ __forceinline__ int canon_off(int row, int k, int K) { return (row >> 3) * (32 * K) + (k >> 2) * 128 + (row & 7) * 16 + (k & 3) * 4; }
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(da),
               "l"(db), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct Case {
  int M, N, K;          // MMA shape (K = total contraction length)
  int a_mn, b_mn;       // operand major-ness
  int a_rows, a_K;      // stored A tile [a_rows][a_K] (K-major canonical)
  int b_rows, b_K;      // stored B tile
};

// generic runner: A tile and B tile are given as plain row-major [rows][K] fp32 in global memory
__global__ void __launch_bounds__(128) run_case(const float *__restrict__ A, const float *__restrict__ B, float *__restrict__ Dout /* [128 lanes][N] */, Case c) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base;
  const int t = threadIdx.x, warp = t >> 5;
  const int a_bytes = c.a_rows * c.a_K * 4, b_bytes = c.b_rows * c.b_K * 4;
  unsigned char *sAhi = smem, *sAlo = smem + a_bytes, *sBhi = smem + 2 * a_bytes, *sBlo = sBhi + b_bytes;
  for (int e = t; e < c.a_rows * c.a_K; e += 128) {
    const int r = e / c.a_K, k = e % c.a_K;
    const float v = A[e], hi = tf32_hi(v);
    *reinterpret_cast<float *>(sAhi + canon_off(r, k, c.a_K)) = hi;
    *reinterpret_cast<float *>(sAlo + canon_off(r, k, c.a_K)) = v - hi;
  }
  for (int e = t; e < c.b_rows * c.b_K; e += 128) {
    const int r = e / c.b_K, k = e % c.b_K;
    const float v = B[e], hi = tf32_hi(v);
    *reinterpret_cast<float *>(sBhi + canon_off(r, k, c.b_K)) = hi;
    *reinterpret_cast<float *>(sBlo + canon_off(r, k, c.b_K)) = v - hi;
  }
  if (t == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base;
  // clear the accumulator columns first so that lanes the MMA does not write read back as a marker
  {
    const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < 64; c0 += 8) {
      const uint32_t mk = __float_as_uint(-12345.0f);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr + c0), "r"(mk) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  if (t == 0) {
    const uint32_t idesc = make_idesc(c.M, c.N, c.a_mn, c.b_mn);
    // K-major use  : LBO = 128 (K-adjacent cores), SBO = 32*K_stored (8-row groups), advance 256 B per MMA (8 k = 2 cores)
    // MN-major use : SBO field = 128 (next 4 features), LBO field = 32*K_stored (next 8 rows = next 8 k), advance 32*K_stored per MMA
    const uint32_t a_lbo = c.a_mn ? 32u * c.a_K : 128u, a_sbo = c.a_mn ? 128u : 32u * c.a_K, a_adv = c.a_mn ? 32u * c.a_K : 256u;
    const uint32_t b_lbo = c.b_mn ? 32u * c.b_K : 128u, b_sbo = c.b_mn ? 128u : 32u * c.b_K, b_adv = c.b_mn ? 32u * c.b_K : 256u;
    int first = 1;
    for (int p = 0; p < 3; ++p) {
      const unsigned char *a = (p == 2) ? sAlo : sAhi;   // hi*hi, hi*lo, lo*hi
      const unsigned char *b = (p == 1) ? sBlo : sBhi;
      for (int kk = 0; kk < c.K / 8; ++kk) {
        mma_tf32(tmem_d, make_desc(smem_u32(a) + kk * a_adv, a_lbo, a_sbo), make_desc(smem_u32(b) + kk * b_adv, b_lbo, b_sbo), idesc, first ? 0u : 1u);
        first = 0;
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
  }
  uint32_t done = 0;
  while (!done)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int lane_row = warp * 32 + (t & 31);
  for (int c0 = 0; c0 < c.N; c0 += 8) {
    uint32_t v[8];
    const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) Dout[lane_row * c.N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(64) : "memory");
}

static std::vector<float> rnd(int n, unsigned seed) {
  std::vector<float> v(n);
  srand(seed);
  for (auto &x : v) x = (float)rand() / RAND_MAX * 2.f - 1.f;
  return v;
}

// reference D[m][n] for logical A(m,k), B(n,k) given storage conventions
static int run(const char *name, Case c, int m_valid) {
  std::vector<float> hA = rnd(c.a_rows * c.a_K, 1), hB = rnd(c.b_rows * c.b_K, 2), hD(128 * c.N, 0.f);
  float *dA, *dB, *dD;
  cudaMalloc(&dA, hA.size() * 4); cudaMalloc(&dB, hB.size() * 4); cudaMalloc(&dD, hD.size() * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = 2 * (size_t)(c.a_rows * c.a_K + c.b_rows * c.b_K) * 4 + 8192;   // slack: the over-reading cases stay inside the allocation
  cudaFuncSetAttribute(run_case, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  run_case<<<1, 128, smem>>>(dA, dB, dD, c);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); return 2; }
  cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
  // logical element getters: K-major use -> (row = m, col = k); MN-major use -> (row = k, col = m)
  auto Aat = [&](int m, int k) { return c.a_mn ? hA[k * c.a_K + m] : hA[m * c.a_K + k]; };
  auto Bat = [&](int n, int k) { return c.b_mn ? hB[k * c.b_K + n] : hB[n * c.b_K + k]; };
  // which TMEM lane holds logical row m?  try the identity and the 16-rows-per-quadrant map, report the better one
  double best = 1e30; const char *best_name = "";
  for (int map = 0; map < 2; ++map) {
    double max_err = 0, max_ref = 0;
    for (int m = 0; m < m_valid; ++m) {
      const int lane = map == 0 ? m : (m / 16) * 32 + (m % 16);
      if (lane >= 128) { max_err = 1e30; break; }
      for (int n = 0; n < c.N; ++n) {
        const bool b_ok = c.b_mn ? n < c.b_K : n < c.b_rows;
        if (!b_ok) continue;   // over-read columns are garbage by construction
        double ref = 0;
        for (int k = 0; k < c.K; ++k) ref += (double)Aat(m, k) * (double)Bat(n, k);
        max_err = fmax(max_err, fabs(ref - hD[lane * c.N + n]));
        max_ref = fmax(max_ref, fabs(ref));
      }
    }
    if (max_err / fmax(max_ref, 1e-30) < best) { best = max_err / fmax(max_ref, 1e-30); best_name = map == 0 ? "lane = m" : "lane = 32*(m/16) + m%16"; }
  }
  printf("%-58s rel err %.3e  (%s)   lanes 0,16,32,64 col0: %g %g %g %g\n", name, best, best_name, hD[0], hD[16 * c.N], hD[32 * c.N], hD[64 * c.N]);
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return best < 2e-5 ? 0 : 1;
}

int main() {
  int rc = 0;
  //                                                       M    N   K  a_mn b_mn a_rows a_K b_rows b_K
  rc |= run("T1 K-major A,B  M128 N64 K24", Case{128, 64, 24, 0, 0, 128, 24, 64, 24}, 128);
  rc |= run("T1b K-major A,B M128 N64 K64", Case{128, 64, 64, 0, 0, 128, 64, 64, 64}, 128);
  rc |= run("T2 MN-major A,B (A^T B) M64 N64 K128", Case{64, 64, 128, 1, 1, 128, 64, 128, 64}, 64);
  rc |= run("T2b MN-major A,B M128(over-read) N64 K128", Case{128, 64, 128, 1, 1, 128, 64, 128, 64}, 64);
  rc |= run("T2c MN-major A (24 wide, over-read) M64 N64 K128", Case{64, 64, 128, 1, 1, 128, 24, 128, 64}, 24);
  rc |= run("T2d MN-major A (64) B (8 wide) M64 N8 K128", Case{64, 8, 128, 1, 1, 128, 64, 128, 8}, 64);
  rc |= run("T2e MN-major A (64) B (8 wide) M128 N16(over-read) K128", Case{128, 16, 128, 1, 1, 128, 64, 128, 8}, 64);
  rc |= run("T3 K-major A, MN-major B (W3c [64][8]) M128 N16 K64", Case{128, 16, 64, 0, 1, 128, 64, 64, 8}, 128);
  rc |= run("T3b K-major A (K8), K-major B ([64][8]) M128 N64 K8", Case{128, 64, 8, 0, 0, 128, 8, 64, 8}, 128);
  rc |= run("T3c K-major A, MN-major B (W2c [64][64]) M128 N64 K64", Case{128, 64, 64, 0, 1, 128, 64, 64, 64}, 128);
  printf(rc ? "FAIL\n" : "PASS\n");
  return rc;
}
