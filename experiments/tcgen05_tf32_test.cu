// Stand-alone bring-up test for the 3xTF32 tcgen05 path planned for the 64x64 Dense layer of the fused kernels:
//   D[128][64] = A[128][64] * B[64][64]^T   (A, B row-major with K contiguous = "K-major"), fp32 in/out,
// one CTA, operands in shared memory in the canonical no-swizzle core-matrix layout, accumulator in TMEM.
// Modes: 1xTF32 (expect ~1e-3 relative error) and 3xTF32 (a_hi*b_hi + a_hi*b_lo + a_lo*b_hi, expect ~1e-6).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o experiments/tcgen05_tf32_test experiments/tcgen05_tf32_test.cu && ./experiments/tcgen05_tf32_test
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

constexpr int M = 128, N = 64, K = 64;
// core matrix = 8 rows x 16 bytes (4 tf32).  Element (row, k): core (row/8, k/4); cores of one 8-row group are contiguous along K.
constexpr int CORE_BYTES = 128;
constexpr int LBO = CORE_BYTES;                 // byte offset between cores adjacent along K
constexpr int SBO = CORE_BYTES * (K / 4);       // byte offset between 8-row groups

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);            // start address, 16-byte units
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;        // leading-dimension byte offset
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;        // stride-dimension byte offset
  d |= (uint64_t)1 << 46;                            // descriptor version (sm_100)
  // base offset 0, lbo mode 0, layout type SWIZZLE_NONE (0) in bits [61,64)
  return d;
}

__device__ __forceinline__ float tf32_hi(float x) {  // round-to-nearest tf32 (cvt.rna.tf32.f32)
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__device__ __forceinline__ void store_core(float *base, int row, int k, float v) {
  const int off = (row >> 3) * SBO + (k >> 2) * LBO + (row & 7) * 16 + (k & 3) * 4;
  *reinterpret_cast<float *>(reinterpret_cast<char *>(base) + off) = v;
}

__global__ void __launch_bounds__(128) gemm_kernel(const float *__restrict__ A, const float *__restrict__ B, float *__restrict__ D, int split) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float *sAhi = reinterpret_cast<float *>(smem_raw);                 // M*K*4 = 32 KB
  float *sAlo = sAhi + M * K;                                        // 32 KB
  float *sBhi = sAlo + M * K;                                        // N*K*4 = 16 KB
  float *sBlo = sBhi + N * K;                                        // 16 KB
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base;
  const int t = threadIdx.x, warp = t >> 5;

  for (int e = t; e < M * K; e += 128) {
    const int r = e / K, k = e % K;
    const float v = A[e], hi = split ? tf32_hi(v) : v;
    store_core(sAhi, r, k, hi);
    store_core(sAlo, r, k, v - hi);
  }
  for (int e = t; e < N * K; e += 128) {
    const int r = e / K, k = e % K;
    const float v = B[e], hi = split ? tf32_hi(v) : v;
    store_core(sBhi, r, k, hi);
    store_core(sBlo, r, k, v - hi);
  }
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1) : "memory");
  }
  // make the generic-proxy smem writes visible to the async (tensor core) proxy
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (warp == 0) {  // one warp allocates 64 TMEM columns (power of two >= 32)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base;

  // instruction descriptor: D=F32 (1<<4), A=B=TF32 (2<<7, 2<<10), K-major both, N>>3 at bit 17, M>>4 at bit 24
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  if (t == 0) {
    const int passes = split ? 3 : 1;
    int first = 1;
    for (int p = 0; p < passes; ++p) {
      const float *a = (p == 2) ? sAlo : sAhi;   // hi*hi, hi*lo, lo*hi
      const float *b = (p == 1) ? sBlo : sBhi;
      for (int kk = 0; kk < K / 8; ++kk) {       // one MMA consumes K = 8 tf32 = two cores along K
        const uint64_t da = make_desc(smem_u32(a) + kk * 2 * LBO, LBO, SBO);
        const uint64_t db = make_desc(smem_u32(b) + kk * 2 * LBO, LBO, SBO);
        const uint32_t acc = first ? 0u : 1u;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
            "l"(da), "l"(db), "r"(idesc), "r"(acc)
            : "memory");
        first = 0;
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
  }
  // everybody waits for the MMAs
  uint32_t done = 0;
  while (!done) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // epilogue: warp w reads TMEM lanes 32w..32w+31 (= rows), 64 columns, 8 columns at a time
  const int row = warp * 32 + (t & 31);
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t v[8];
    const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) D[row * N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(64) : "memory");
}

int main() {
  float *hA = (float *)malloc(M * K * 4), *hB = (float *)malloc(N * K * 4), *hD = (float *)malloc(M * N * 4);
  srand(1);
  for (int i = 0; i < M * K; ++i) hA[i] = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (int i = 0; i < N * K; ++i) hB[i] = (float)rand() / RAND_MAX * 2.f - 1.f;
  float *dA, *dB, *dD;
  cudaMalloc(&dA, M * K * 4); cudaMalloc(&dB, N * K * 4); cudaMalloc(&dD, M * N * 4);
  cudaMemcpy(dA, hA, M * K * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, N * K * 4, cudaMemcpyHostToDevice);
  const size_t smem = (size_t)(2 * M * K + 2 * N * K) * 4 + 1024;
  cudaFuncSetAttribute(gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int rc = 0;
  for (int split = 0; split < 2; ++split) {
    cudaMemset(dD, 0, M * N * 4);
    gemm_kernel<<<1, 128, smem>>>(dA, dB, dD, split);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("split=%d CUDA error: %s\n", split, cudaGetErrorString(e)); return 2; }
    cudaMemcpy(hD, dD, M * N * 4, cudaMemcpyDeviceToHost);
    double max_err = 0, max_ref = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int k = 0; k < K; ++k) ref += (double)hA[m * K + k] * (double)hB[n * K + k];
        max_err = fmax(max_err, fabs(ref - hD[m * N + n]));
        max_ref = fmax(max_ref, fabs(ref));
      }
    printf("%s: max |err| = %.3e (max |ref| = %.3f, rel %.3e)  D[0][0..3] = %f %f %f %f\n", split ? "3xTF32" : "1xTF32", max_err, max_ref, max_err / max_ref,
           hD[0], hD[1], hD[2], hD[3]);
    if (split && max_err / max_ref > 2e-5) rc = 1;
    if (!split && max_err / max_ref > 5e-3) rc = 1;
  }
  printf(rc ? "FAIL\n" : "PASS\n");
  return rc;
}
