// Legacy warp-level tensor-core rate on sm_100a: mma.sync.aligned.m16n8k8 tf32 (1024 MAC per instruction), register operands only.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o experiments/mma_sync_tf32_rate experiments/mma_sync_tf32_rate.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
__global__ void __launch_bounds__(256) k(float *out, int iters) {
  float c[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 7, a3 = 9, b0 = threadIdx.x ^ 5, b1 = 11;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0; for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  for (int ctas = 1; ctas <= 4; ctas *= 2) {
    const int iters = 20000;
    k<<<148 * ctas, 256>>>(d, 100); cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a); k<<<148 * ctas, 256>>>(d, iters); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double n_mma = (double)148 * ctas * 8 * iters * 8;  // warps x iterations x 8
    printf("%d CTA/SM x 8 warps: %.3f ms, %.1f TFLOP/s tf32 dense (2*MAC), %.2f mma/cycle/SM @1.9GHz\n", ctas, ms, n_mma * 2048 / ms / 1e9, n_mma / 148 / (ms * 1e-3 * 1.9e9));
  }
  return 0;
}
