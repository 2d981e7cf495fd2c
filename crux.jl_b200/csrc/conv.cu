// Pixel-DQN trunk (examples/rl/atari.jl:8): Chain(x -> x ./ 255f0, Conv((8,8), 4 => 16, relu, stride = 4),
// Conv((4,4), 16 => 32, relu, stride = 2), flatten, Dense(F, 256, relu), Dense(256, nA)) forward, Zygote-equivalent backward and the
// Flux.Optimiser(ClipValue(1f0), Adam(1f-3)) step (atari.jl:10).
//
// Every convolution is an IMPLICIT GEMM: C[m][co] = sum_k A(m, k) W(k, co) with m = (sample, output pixel), k = (ci, kh, kw) in the
// memory order of the Flux weight array [kw, kh, ci, co] -- the A operand is never materialised (no im2col buffer: at B = 512 that would
// be 210 MB for the first layer); its elements are fetched by an index functor while the tile is staged in shared memory, and the
// u8 -> float32 conversion with the `./ 255f0` of the example is fused into that fetch (the observations stay u8 in HBM: 28 224 B per
// sample instead of 112 896 B).  Flux's Conv is a true convolution (NNlib.conv flips the kernel): y[ow, oh] uses x[s*ow + K-1-kw, ...].
// The three GEMMs of a layer share one kernel with different functors:
//     forward          A = input patches,                  B = W(k, co),            epilogue relu(acc + b) -> NHWC (conv1) / NCHW = flatten (conv2)
//     weight gradient  A = input patches TRANSPOSED,       B = dY(row, co),         split over rows into partial slabs, reduced in a fixed order
//     data gradient    A = dY gathered per INPUT pixel,    B = W(co,kh,kw ; ci),    epilogue * relu'(y1)            (conv2 -> conv1 only)
// Layouts: observations [b][c][h][w] (w fastest: Flux WHCN memory), conv1 output NHWC [b][pixel][c1] (the natural GEMM output), conv2
// output written straight into the flattened NCHW head input [b][c2*P2] (Flux.flatten order), so no transpose kernel exists.
// Bound: fp32 FFMA (SIMT), 64 x BN x 16 tiles.  The tcgen05 version (3xTF32 like mb_t5.cuh) is the next step; this one is the parity
// anchor (1e-5 against the torch-CPU conv2d autograd restatement in oracle/conv_oracle.py).
#include "conv.cuh"

namespace {

constexpr int BM = 64, BK = 16;

// ---- operand functors ------------------------------------------------------------------------------------------------------------------
// A patch operand is SEPARABLE: its address is rp(m) + cp(k) (row part from the output pixel, column part from the kernel tap), so the
// kernel computes the part that is fixed for a thread once and the other part once per k-step of the tile (a 16-entry shared table)
// instead of ~8 integer divisions per fetched element.
template <class T>
struct PatchNCHW {   // A(m, k) of a convolution over an NCHW input (u8 scaled by 1/255, or float32 as is)
  static constexpr bool separable = true;
  const T *x; ConvGeom g; int scale255;
  __device__ __forceinline__ int64_t rp(int m) const {
    const int P = g.OH * g.OW, b = m / P, p = m - b * P, oh = p / g.OW, ow = p - oh * g.OW;
    return ((int64_t)b * g.C * g.H + oh * g.S) * g.W + ow * g.S;
  }
  __device__ __forceinline__ int64_t cp(int k) const {
    const int kw = k % g.K, t = k / g.K, kh = t % g.K, ci = t / g.K;
    return ((int64_t)ci * g.H + (g.K - 1 - kh)) * g.W + (g.K - 1 - kw);
  }
  __device__ __forceinline__ float ld(int64_t i) const {
    const float v = (float)x[i];
    return scale255 ? __fdiv_rn(v, 255.0f) : v;
  }
  __device__ __forceinline__ float operator()(int m, int k) const { return ld(rp(m) + cp(k)); }
};
struct PatchNHWC {   // the same over an NHWC float32 input (conv1's output)
  static constexpr bool separable = true;
  const float *x; ConvGeom g;
  __device__ __forceinline__ int64_t rp(int m) const {
    const int P = g.OH * g.OW, b = m / P, p = m - b * P, oh = p / g.OW, ow = p - oh * g.OW;
    return ((int64_t)b * g.H * g.W + (int64_t)oh * g.S * g.W + ow * g.S) * g.C;
  }
  __device__ __forceinline__ int64_t cp(int k) const {
    const int kw = k % g.K, t = k / g.K, kh = t % g.K, ci = t / g.K;
    return ((int64_t)(g.K - 1 - kh) * g.W + (g.K - 1 - kw)) * g.C + ci;
  }
  __device__ __forceinline__ float ld(int64_t i) const { return x[i]; }
  __device__ __forceinline__ float operator()(int m, int k) const { return ld(rp(m) + cp(k)); }
};
template <class AF>
struct Transposed {
  static constexpr bool separable = AF::separable;
  AF a;
  __device__ __forceinline__ int64_t rp(int m) const { return a.cp(m); }
  __device__ __forceinline__ int64_t cp(int k) const { return a.rp(k); }
  __device__ __forceinline__ float ld(int64_t i) const { return a.ld(i); }
  __device__ __forceinline__ float operator()(int m, int k) const { return a(k, m); }
};
struct WeightKxCo {  // B(k, co) = W[co][k]  (Flux memory: k = kw + K (kh + K ci) fastest, co slowest)
  const float *w; int KK;
  __device__ __forceinline__ float operator()(int k, int n) const { return w[(int64_t)n * KK + k]; }
};
struct RowMajorB { const float *b; int ld; __device__ __forceinline__ float operator()(int k, int n) const { return b[(int64_t)k * ld + n]; } };
struct DY2 {         // gradient wrt conv2's pre-activation, read from the head's input gradient dF [b][c2*P2] masked by relu'(F)
  const float *dF, *F; int P, CO;
  __device__ __forceinline__ float at(int row, int co) const {
    const int b = row / P, p = row - b * P;
    const int64_t i = (int64_t)b * P * CO + (int64_t)co * P + p;
    return F[i] > 0.0f ? dF[i] : 0.0f;
  }
  __device__ __forceinline__ float operator()(int k, int n) const { return at(k, n); }   // as the B operand (k = row)
};
struct DY2Gather {   // A(m1, kidx) of the data-gradient GEMM: m1 = (b, ih, iw) of conv2's INPUT, kidx = co + CO (kw + K kh)
  static constexpr bool separable = false;
  DY2 dy; ConvGeom g;
  __device__ __forceinline__ int64_t rp(int) const { return 0; }
  __device__ __forceinline__ int64_t cp(int) const { return 0; }
  __device__ __forceinline__ float ld(int64_t) const { return 0.f; }
  __device__ __forceinline__ float operator()(int m, int k) const {
    const int P1 = g.H * g.W, b = m / P1, p = m - b * P1, ih = p / g.W, iw = p - ih * g.W;
    const int co = k % g.CO, t = k / g.CO, kw = t % g.K, kh = t / g.K;
    const int th = ih - (g.K - 1 - kh), tw = iw - (g.K - 1 - kw);
    if (th < 0 || tw < 0 || th % g.S || tw % g.S) return 0.0f;
    const int oh = th / g.S, ow = tw / g.S;
    if (oh >= g.OH || ow >= g.OW) return 0.0f;
    return dy.at(b * g.OH * g.OW + oh * g.OW + ow, co);
  }
};
struct WeightGatherB {   // B(kidx, ci) = W2[co][ci][kh][kw]
  const float *w; ConvGeom g;
  __device__ __forceinline__ float operator()(int k, int n) const {
    const int co = k % g.CO, t = k / g.CO, kw = t % g.K, kh = t / g.K;
    return w[kw + g.K * (kh + g.K * (n + g.C * co))];
  }
};
// ---- epilogues -------------------------------------------------------------------------------------------------------------------------
struct EpiReluNHWC {
  float *out; const float *bias; int N;
  __device__ __forceinline__ void operator()(int m, int n, float acc, int) const { out[(int64_t)m * N + n] = fmaxf(acc + bias[n], 0.0f); }
};
struct EpiReluNCHW {   // flatten order: [b][co * P + p]
  float *out; const float *bias; int N, P;
  __device__ __forceinline__ void operator()(int m, int n, float acc, int) const {
    const int b = m / P, p = m - b * P;
    out[(int64_t)b * P * N + (int64_t)n * P + p] = fmaxf(acc + bias[n], 0.0f);
  }
};
struct EpiPartial {    // slab z of the weight-gradient partials: [z][M][N]
  float *out; int M, N;
  __device__ __forceinline__ void operator()(int m, int n, float acc, int z) const { out[((int64_t)z * M + m) * N + n] = acc; }
};
struct EpiMaskRelu {   // dY1 = acc * relu'(y1)
  float *out; const float *y; int N;
  __device__ __forceinline__ void operator()(int m, int n, float acc, int) const {
    const int64_t i = (int64_t)m * N + n;
    out[i] = y[i] > 0.0f ? acc : 0.0f;
  }
};

// C(m, n) = sum_{k in slab} A(m, k) B(k, n): 64 x BN x 16 tiles, 256 threads, thread = 4 rows x BN/16 columns
template <int BN, class AF, class BF, class EF>
__global__ void __launch_bounds__(256) igemm_kernel(AF A, BF Bf, EF E, int M, int N, int K, int k_per_slab) {
  constexpr int TN = BN / 16;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
  const int k_begin = blockIdx.z * k_per_slab, k_end = min(K, k_begin + k_per_slab);
  __shared__ int64_t ktab[BK];
  const int am = tid & (BM - 1), ak = tid / BM;   // this thread stages rows am of the A tile, k-steps ak, ak + 4, ak + 8, ak + 12
  int64_t rbase = 0;
  if (AF::separable && m0 + am < M) rbase = A.rp(m0 + am);
  float acc[4][TN];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  for (int k0 = k_begin; k0 < k_end; k0 += BK) {
    if (AF::separable) {
      if (tid < BK && k0 + tid < k_end) ktab[tid] = A.cp(k0 + tid);
      __syncthreads();   // (the previous tile's compute phase ended with a barrier: ktab is free)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = ak + 4 * i;
        As[k][am] = (m0 + am < M && k0 + k < k_end) ? A.ld(rbase + ktab[k]) : 0.f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = tid + i * 256, m = idx & (BM - 1), k = idx / BM;
        const int gm = m0 + m, gk = k0 + k;
        As[k][m] = (gm < M && gk < k_end) ? A(gm, gk) : 0.f;
      }
    }
    for (int idx = tid; idx < BK * BN; idx += 256) {
      const int n = idx % BN, k = idx / BN;
      const int gn = n0 + n, gk = k0 + k;
      Bs[k][n] = (gn < N && gk < k_end) ? Bf(gk, gn) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      float bv[TN];
#pragma unroll
      for (int j = 0; j < TN; ++j) bv[j] = Bs[k][tx * TN + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int gn = n0 + tx * TN + j;
      if (gn < N) E(gm, gn, acc[i][j], (int)blockIdx.z);
    }
  }
}

// grads_W[co][k] = sum_z partial[z][k][co] (fixed order: bit-reproducible); grads_b[co] = sum_rows dY(row, co) (one block per co)
__global__ void conv_reduce_kernel(const float *__restrict__ part, int slabs, int KK, int CO, float *__restrict__ gW) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (int64_t)KK * CO; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i / CO), co = (int)(i - (int64_t)k * CO);
    float s = 0.f;
    for (int z = 0; z < slabs; ++z) s += part[(int64_t)z * KK * CO + i];
    gW[(int64_t)co * KK + k] = s;
  }
}
__device__ __forceinline__ double block_sum_d(double v, double *sh /*32*/) {   // valid on thread 0
  v = warp_sum_d(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x < 32) {
    r = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
    r = warp_sum_d(r);
  }
  return r;
}
template <class DYF>
__global__ void conv_bias_grad_kernel(DYF dy, int64_t rows, float *__restrict__ gb) {
  __shared__ double sh[32];
  const int co = blockIdx.x;
  double s = 0.0;
  for (int64_t r = threadIdx.x; r < rows; r += blockDim.x) s += (double)dy((int)r, co);
  s = block_sum_d(s, sh);
  if (threadIdx.x == 0) gb[co] = (float)s;
}

int ensure_ws(crux_convq *net, int64_t B) {
  crux_ctx *ctx = net->ctx;
  if (B <= net->cap) return CRUX_OK;
  CRUX_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const int64_t cap = B + B / 8;
  if (net->y1) cudaFree(net->y1);
  if (net->f) cudaFree(net->f);
  if (net->dy1) cudaFree(net->dy1);
  net->y1 = net->f = net->dy1 = nullptr;
  const size_t n1 = (size_t)cap * net->g1.OH * net->g1.OW * net->g1.CO * sizeof(float);
  if (cudaMalloc((void **)&net->y1, n1) != cudaSuccess || cudaMalloc((void **)&net->dy1, n1) != cudaSuccess ||
      cudaMalloc((void **)&net->f, (size_t)cap * net->F * sizeof(float)) != cudaSuccess)
    return crux_set_err(ctx, CRUX_ERR_OOM, "convq workspace for batch %lld", (long long)cap);
  net->cap = cap;
  return CRUX_OK;
}

template <class T>
int forward_t(crux_convq *net, const T *s, int scale, int64_t B) {
  crux_ctx *ctx = net->ctx;
  const ConvGeom &g1 = net->g1, &g2 = net->g2;
  const int KK1 = g1.K * g1.K * g1.C, KK2 = g2.K * g2.K * g2.C;
  const int64_t M1 = B * g1.OH * g1.OW, M2 = B * g2.OH * g2.OW;
  CRUX_REQUIRE(ctx, M1 < (int64_t)1 << 31, "convq: batch too large for 32-bit row indices");
  {
    PatchNCHW<T> A{s, g1, scale};
    WeightKxCo W{net->params, KK1};
    EpiReluNHWC E{net->y1, net->params + net->off_b1, g1.CO};
    dim3 grid((unsigned)cdiv(g1.CO, 16), (unsigned)cdiv(M1, BM), 1);
    if (g1.CO <= 16) igemm_kernel<16><<<grid, 256, 0, ctx->stream>>>(A, W, E, (int)M1, g1.CO, KK1, KK1);
    else { grid.x = (unsigned)cdiv(g1.CO, 32); igemm_kernel<32><<<grid, 256, 0, ctx->stream>>>(A, W, E, (int)M1, g1.CO, KK1, KK1); }
    CRUX_LAUNCHED(ctx);
  }
  {
    PatchNHWC A{net->y1, g2};
    WeightKxCo W{net->params + net->off_w2, KK2};
    EpiReluNCHW E{net->f, net->params + net->off_b2, g2.CO, g2.OH * g2.OW};
    dim3 grid((unsigned)cdiv(g2.CO, 32), (unsigned)cdiv(M2, BM), 1);
    igemm_kernel<32><<<grid, 256, 0, ctx->stream>>>(A, W, E, (int)M2, g2.CO, KK2, KK2);
    CRUX_LAUNCHED(ctx);
  }
  return mlp_forward_keep(net->head, net->f, B, nullptr);
}

// split of `rows` into slabs of a multiple of BK rows so that ~2 CTAs per SM are in flight
void slab_plan(crux_ctx *ctx, int64_t rows, int64_t tiles, int &slabs, int &per) {
  int64_t S = cdiv(2 * (int64_t)ctx->num_sms, tiles);
  S = i64max(1, i64min(S, cdiv(rows, 256)));
  int64_t p = cdiv(cdiv(rows, S), BK) * BK;
  slabs = (int)cdiv(rows, p); per = (int)p;
}

template <class T>
int backward_t(crux_convq *net, const T *s, int scale, int64_t B, float *dq) {
  crux_ctx *ctx = net->ctx;
  const ConvGeom &g1 = net->g1, &g2 = net->g2;
  const int KK1 = g1.K * g1.K * g1.C, KK2 = g2.K * g2.K * g2.C, P1 = g1.OH * g1.OW, P2 = g2.OH * g2.OW;
  const int64_t M1 = B * P1, M2 = B * P2;
  // head: dF = head->dz[0] (gradient wrt the flattened input; relu'(F) is applied by the DY2 functor)
  int rc = mlp_backward(net->head, net->f, B, dq, true, false, true, nullptr);
  if (rc) return rc;
  DY2 dy2{net->head->dz[0], net->f, P2, g2.CO};
  int s1, p1, s2, p2;
  slab_plan(ctx, M1, cdiv(KK1, BM) * cdiv(g1.CO, 16), s1, p1);
  slab_plan(ctx, M2, cdiv(KK2, BM) * cdiv(g2.CO, 32), s2, p2);
  const size_t need = ((size_t)s1 * KK1 * g1.CO + (size_t)s2 * KK2 * g2.CO) * sizeof(float);
  if (need > net->partials_bytes) {
    CRUX_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (net->partials) cudaFree(net->partials);
    net->partials = nullptr; net->partials_bytes = 0;
    if (cudaMalloc((void **)&net->partials, need + need / 4) != cudaSuccess) return crux_set_err(ctx, CRUX_ERR_OOM, "convq partials %zu B", need);
    net->partials_bytes = need + need / 4;
  }
  float *part1 = net->partials, *part2 = net->partials + (size_t)s1 * KK1 * g1.CO;
  {   // conv2: dW2[k][co] = sum_rows patch(row, k) dY2(row, co)
    Transposed<PatchNHWC> A{PatchNHWC{net->y1, g2}};
    EpiPartial E{part2, KK2, g2.CO};
    dim3 grid((unsigned)cdiv(g2.CO, 32), (unsigned)cdiv(KK2, BM), (unsigned)s2);
    igemm_kernel<32><<<grid, 256, 0, ctx->stream>>>(A, dy2, E, KK2, g2.CO, (int)M2, p2);
    CRUX_LAUNCHED(ctx);
    conv_reduce_kernel<<<(unsigned)cdiv((int64_t)KK2 * g2.CO, 256), 256, 0, ctx->stream>>>(part2, s2, KK2, g2.CO, net->grads + net->off_w2);
    CRUX_LAUNCHED(ctx);
    conv_bias_grad_kernel<<<g2.CO, 256, 0, ctx->stream>>>(dy2, M2, net->grads + net->off_b2);
    CRUX_LAUNCHED(ctx);
  }
  {   // dY1 = (dY2 gathered per input pixel) x W2, masked by relu'(y1)
    ConvGeom gg = g2;   // g2.C == g1.CO, g2.H x g2.W == conv1's output grid
    DY2Gather A{dy2, gg};
    WeightGatherB W{net->params + net->off_w2, gg};
    EpiMaskRelu E{net->dy1, net->y1, g1.CO};
    const int KD = g2.K * g2.K * g2.CO;
    dim3 grid((unsigned)cdiv(g1.CO, 16), (unsigned)cdiv(M1, BM), 1);
    if (g1.CO <= 16) igemm_kernel<16><<<grid, 256, 0, ctx->stream>>>(A, W, E, (int)M1, g1.CO, KD, KD);
    else { grid.x = (unsigned)cdiv(g1.CO, 32); igemm_kernel<32><<<grid, 256, 0, ctx->stream>>>(A, W, E, (int)M1, g1.CO, KD, KD); }
    CRUX_LAUNCHED(ctx);
  }
  {   // conv1: dW1[k][co] = sum_rows patch(row, k) dY1(row, co)
    Transposed<PatchNCHW<T>> A{PatchNCHW<T>{s, g1, scale}};
    RowMajorB Bf{net->dy1, g1.CO};
    EpiPartial E{part1, KK1, g1.CO};
    dim3 grid((unsigned)cdiv(g1.CO, 16), (unsigned)cdiv(KK1, BM), (unsigned)s1);
    if (g1.CO <= 16) igemm_kernel<16><<<grid, 256, 0, ctx->stream>>>(A, Bf, E, KK1, g1.CO, (int)M1, p1);
    else { grid.x = (unsigned)cdiv(g1.CO, 32); igemm_kernel<32><<<grid, 256, 0, ctx->stream>>>(A, Bf, E, KK1, g1.CO, (int)M1, p1); }
    CRUX_LAUNCHED(ctx);
    conv_reduce_kernel<<<(unsigned)cdiv((int64_t)KK1 * g1.CO, 256), 256, 0, ctx->stream>>>(part1, s1, KK1, g1.CO, net->grads);
    CRUX_LAUNCHED(ctx);
    conv_bias_grad_kernel<<<g1.CO, 256, 0, ctx->stream>>>(Bf, M1, net->grads + net->off_b1);
    CRUX_LAUNCHED(ctx);
  }
  return CRUX_OK;
}

__global__ void polyak2_kernel(float *__restrict__ to, const float *__restrict__ from, int64_t n, float tau) {
  const float omt = 1.0f - tau;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) to[i] = tau * from[i] + omt * to[i];
}

}  // namespace

int convq_forward_keep(crux_convq *net, const void *s, int s_is_u8, int64_t B) {
  int rc = ensure_ws(net, B);
  if (rc) return rc;
  return s_is_u8 ? forward_t<uint8_t>(net, (const uint8_t *)s, net->scale255, B) : forward_t<float>(net, (const float *)s, net->scale255, B);
}
int convq_backward(crux_convq *net, const void *s, int s_is_u8, int64_t B, float *dq) {
  return s_is_u8 ? backward_t<uint8_t>(net, (const uint8_t *)s, net->scale255, B, dq) : backward_t<float>(net, (const float *)s, net->scale255, B, dq);
}
int convq_adam_step(crux_convq *net, float *gnorm_out_dev) {
  crux_mlp *h = net->head;
  AdamSegs segs;
  segs.n = 2;
  segs.clip = net->clip;
  segs.s[0] = AdamSeg{net->params, net->grads, net->m, net->v, net->n_conv};
  segs.s[1] = AdamSeg{h->params, h->grads, h->m, h->v, h->n_params};
  return adam_step_segments(net->ctx, segs, h->eta, h->beta1, h->beta2, h->eps, h->step_dev, gnorm_out_dev, nullptr, h->norm_part);
}

extern "C" {

int32_t crux_convq_create(crux_ctx *ctx, int32_t C, int32_t H, int32_t W, int32_t scale255, int32_t k1, int32_t s1, int32_t c1, int32_t k2, int32_t s2,
                          int32_t c2, int32_t hidden, int32_t nA, crux_convq **out) {
  if (!ctx || !out) return CRUX_ERR_INVALID;
  CRUX_REQUIRE(ctx, C >= 1 && H >= k1 && W >= k1 && k1 >= 1 && s1 >= 1 && c1 >= 1 && c1 <= 32 && k2 >= 1 && s2 >= 1 && c2 >= 1 && c2 <= 32 && hidden >= 1 && nA >= 1,
               "crux_convq_create: bad shape (channels per conv layer <= 32)");
  crux_convq *net = new crux_convq();
  net->ctx = ctx;
  net->scale255 = scale255 ? 1 : 0;
  net->g1 = ConvGeom{C, H, W, k1, s1, (H - k1) / s1 + 1, (W - k1) / s1 + 1, c1};
  CRUX_REQUIRE(ctx, net->g1.OH >= k2 && net->g1.OW >= k2, "crux_convq_create: second kernel larger than the first layer's output");
  net->g2 = ConvGeom{c1, net->g1.OH, net->g1.OW, k2, s2, (net->g1.OH - k2) / s2 + 1, (net->g1.OW - k2) / s2 + 1, c2};
  net->F = c2 * net->g2.OH * net->g2.OW;
  const int64_t nw1 = (int64_t)c1 * C * k1 * k1, nw2 = (int64_t)c2 * c1 * k2 * k2;
  net->off_b1 = nw1; net->off_w2 = nw1 + c1; net->off_b2 = net->off_w2 + nw2; net->n_conv = net->off_b2 + c2;
  const size_t bytes = (size_t)net->n_conv * sizeof(float);
  if (cudaMalloc((void **)&net->params, bytes) != cudaSuccess || cudaMalloc((void **)&net->grads, bytes) != cudaSuccess ||
      cudaMalloc((void **)&net->m, bytes) != cudaSuccess || cudaMalloc((void **)&net->v, bytes) != cudaSuccess) {
    crux_convq_destroy(net);
    return crux_set_err(ctx, CRUX_ERR_OOM, "crux_convq_create: cudaMalloc");
  }
  cudaMemsetAsync(net->params, 0, bytes, ctx->stream); cudaMemsetAsync(net->grads, 0, bytes, ctx->stream);
  cudaMemsetAsync(net->m, 0, bytes, ctx->stream); cudaMemsetAsync(net->v, 0, bytes, ctx->stream);
  const int32_t dims[3] = {net->F, hidden, nA}, acts[2] = {CRUX_ACT_RELU, CRUX_ACT_IDENTITY};
  int rc = crux_mlp_create(ctx, 2, dims, acts, &net->head);
  if (rc) { crux_convq_destroy(net); return rc; }
  *out = net;
  return CRUX_OK;
}
int32_t crux_convq_destroy(crux_convq *net) {
  if (!net) return CRUX_OK;
  cudaStreamSynchronize(net->ctx->stream);
  for (float *p : {net->params, net->grads, net->m, net->v, net->y1, net->f, net->dy1, net->partials})
    if (p) cudaFree(p);
  if (net->head) crux_mlp_destroy(net->head);
  delete net;
  return CRUX_OK;
}
int32_t crux_convq_num_params(crux_convq *net, int64_t *out) {
  if (!net || !out) return CRUX_ERR_INVALID;
  *out = net->n_conv + net->head->n_params;
  return CRUX_OK;
}
int32_t crux_convq_shape(crux_convq *net, int32_t *flatten_out, int32_t *oh1, int32_t *ow1, int32_t *oh2, int32_t *ow2) {
  if (!net) return CRUX_ERR_INVALID;
  if (flatten_out) *flatten_out = net->F;
  if (oh1) *oh1 = net->g1.OH;
  if (ow1) *ow1 = net->g1.OW;
  if (oh2) *oh2 = net->g2.OH;
  if (ow2) *ow2 = net->g2.OW;
  return CRUX_OK;
}
int32_t crux_convq_set_params(crux_convq *net, const float *flat_host) {
  if (!net || !flat_host) return CRUX_ERR_INVALID;
  crux_ctx *ctx = net->ctx;
  CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(net->params, flat_host, (size_t)net->n_conv * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  CRUX_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return crux_mlp_set_params(net->head, flat_host + net->n_conv);
}
int32_t crux_convq_get_params(crux_convq *net, float *flat_host) {
  if (!net || !flat_host) return CRUX_ERR_INVALID;
  crux_ctx *ctx = net->ctx;
  CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(flat_host, net->params, (size_t)net->n_conv * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  CRUX_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return crux_mlp_get_params(net->head, flat_host + net->n_conv);
}
int32_t crux_convq_grads(crux_convq *net, float *flat_host) {
  if (!net || !flat_host) return CRUX_ERR_INVALID;
  crux_ctx *ctx = net->ctx;
  CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(flat_host, net->grads, (size_t)net->n_conv * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(flat_host + net->n_conv, net->head->grads, (size_t)net->head->n_params * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  CRUX_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return CRUX_OK;
}
int32_t crux_convq_set_adam(crux_convq *net, double eta, double beta1, double beta2, double eps, float clip_value) {
  if (!net) return CRUX_ERR_INVALID;
  crux_ctx *ctx = net->ctx;
  CRUX_REQUIRE(ctx, clip_value >= 0.f, "crux_convq_set_adam: clip_value must be >= 0 (0 = no ClipValue)");
  net->clip = clip_value;
  const size_t bytes = (size_t)net->n_conv * sizeof(float);
  CRUX_CHECK_CUDA(ctx, cudaMemsetAsync(net->m, 0, bytes, ctx->stream));
  CRUX_CHECK_CUDA(ctx, cudaMemsetAsync(net->v, 0, bytes, ctx->stream));
  return crux_mlp_set_adam(net->head, eta, beta1, beta2, eps);
}
int32_t crux_convq_forward(crux_convq *net, const void *s, int32_t s_is_u8, int64_t B, float *q_out) {
  if (!net) return CRUX_ERR_INVALID;
  crux_ctx *ctx = net->ctx;
  CRUX_REQUIRE(ctx, s && q_out && B >= 1, "crux_convq_forward: bad arguments");
  int rc = convq_forward_keep(net, s, s_is_u8, B);
  if (rc) return rc;
  const crux_mlp *h = net->head;
  CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(q_out, h->act[h->n_layers], (size_t)B * h->dims[h->n_layers] * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
  return CRUX_OK;
}
int32_t crux_convq_copy(crux_convq *to, crux_convq *from) {
  if (!to || !from) return CRUX_ERR_INVALID;
  crux_ctx *ctx = to->ctx;
  CRUX_REQUIRE(ctx, to->n_conv == from->n_conv, "crux_convq_copy: shape mismatch");
  CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(to->params, from->params, (size_t)to->n_conv * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
  return crux_mlp_copy(to->head, from->head);
}
int32_t crux_convq_polyak(crux_convq *to, crux_convq *from, float tau) {
  if (!to || !from) return CRUX_ERR_INVALID;
  crux_ctx *ctx = to->ctx;
  CRUX_REQUIRE(ctx, to->n_conv == from->n_conv, "crux_convq_polyak: shape mismatch");
  polyak2_kernel<<<(unsigned)i64min(cdiv(to->n_conv, 256), 1024), 256, 0, ctx->stream>>>(to->params, from->params, to->n_conv, tau);
  CRUX_LAUNCHED(ctx);
  return crux_mlp_polyak(to->head, from->head, tau);
}

}  // extern "C"
