// Probe 3 (= probe 2 issued from a warp-uniform branch with elect.sync instead of `if (thread == 0)`): is the ~96 cycles per small tcgen05.mma (tc5_probe.cu) a dependent-accumulator latency, a per-thread issue cost, or a
// tensor-pipe floor?  Chains of R MMAs (K = 8, SS operands) round-robin over n_acc independent accumulators, issued by n_iss threads
// (one per warp, each with its own accumulators and its own commit barrier).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o experiments/tc5_probe5 experiments/tc5_probe5.cu && ./experiments/tc5_probe5
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(da), "l"(db),
               "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(db),
               "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
struct Case { int M, N, n_acc, n_iss, reps, kind; };   // kind 0 = tf32 SS, 2 = tf32 TS (A from tensor memory, columns [384, 392))   // kind 0 = tf32 (K = 8), 1 = f16/bf16 (K = 16)
__global__ void __launch_bounds__(128) probe(long long *cycles, Case c) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t mbar[4];
  __shared__ uint32_t tmem_base;
  const int t = threadIdx.x, w = t >> 5;
  for (int e = t; e < 16384; e += 128) reinterpret_cast<float *>(smem)[e] = 0.f;
  if (t == 0) for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar[i])), "r"(1) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (w == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_base;
  if (w < c.n_iss) { if (elect_one()) {
    const uint32_t idesc = c.kind != 1 ? ((1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(c.M >> 4) << 24))
                                       : ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(c.M >> 4) << 24));
    const uint64_t da = make_desc(smem_u32(smem), 128, 256), db = make_desc(smem_u32(smem) + 16384, 128, 256);
    const uint32_t base = tm + (uint32_t)(w * c.n_acc * c.N);
    const long long t0 = clock64();
    if (c.kind == 2) { for (int r = 0; r < c.reps; ++r) mma_ts(base + (uint32_t)((r & (c.n_acc - 1)) * c.N), tm + 384u, db, idesc, 1u); }
    else for (int r = 0; r < c.reps; ++r) mma_ss(base + (uint32_t)((r & (c.n_acc - 1)) * c.N), da, db, idesc, 1u);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar[w])) : "memory");
    const long long t_issue = clock64() - t0;
    uint32_t done = 0;
    while (!done)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&mbar[w])), "r"(0) : "memory");
    cycles[2 * w] = clock64() - t0;
    cycles[2 * w + 1] = t_issue; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}
int main() {
  long long *d, h[8];
  cudaMalloc(&d, 64);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  const Case cases[] = {
      {64, 8, 1, 1, 512, 2}, {64, 16, 1, 1, 512, 2}, {64, 24, 1, 1, 512, 2}, {64, 32, 1, 1, 512, 2}, {64, 64, 1, 1, 512, 2}, {64, 128, 1, 1, 512, 2},
      {128, 16, 1, 1, 512, 2}, {128, 64, 1, 1, 512, 2}, {128, 128, 1, 1, 512, 2},
      {64, 32, 1, 1, 512, 0}, {64, 64, 1, 1, 512, 0}, {64, 32, 1, 1, 24, 2}, {64, 32, 1, 1, 24, 0}, {64, 32, 1, 2, 512, 2},
  };


  for (const Case &c : cases) {
    double best = 1e30, best_issue = 0;
    for (int it = 0; it < 3; ++it) {
      cudaMemset(d, 0, 64);
      probe<<<1, 128, 96 * 1024>>>(d, c);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error\n"); return 2; }
      cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
      double mx = 0, mi = 0;
      for (int w = 0; w < c.n_iss; ++w) { if (h[2 * w] > mx) mx = (double)h[2 * w]; if (h[2 * w + 1] > mi) mi = (double)h[2 * w + 1]; }
      if (mx < best) { best = mx; best_issue = mi; }
    }
    printf("%s M%-3d N%-3d  accumulators %d  issuing threads %d  chain %3d : %7.0f cycles total = %6.1f / MMA per thread, %6.1f / MMA overall (issue loop alone %6.0f)\n",
           c.kind == 1 ? "bf16 K16" : (c.kind == 2 ? "tf32 TS " : "tf32 SS "), c.M, c.N, c.n_acc, c.n_iss, c.reps, best, best / c.reps, best / (c.reps * c.n_iss), best_issue);
  }
  return 0;
}
