// Generic Dense/Chain engine: forward, backward and the Flux Adam step for arbitrary layer widths.
// This is the layer-by-layer path every solver can fall back on (SAC 256-wide nets, DQN 2-8-4, ...);
// the PPO headline shapes additionally have a fused kernel (ppo_fused.cu).
//
// Reference semantics: Dense = act.(W*x .+ b) with W [out,in] column-major (policies.jl:94-96,120 call
// Flux Chains); gradients are what Zygote's pullback returns (training.jl:16-18); the optimiser is Flux's
// Adam with Float64 scalars and float32 moments (training.jl:3,21).
#include "mlp.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16;
enum { EPI_FWD = 0, EPI_BWD_DATA = 1, EPI_PARTIAL = 2 };

// C[M][N] = sum_k A(m,k) B(k,n);  A(m,k) = TA ? A[k*lda+m] : A[m*lda+k];  B(k,n) = TB ? B[n*ldb+k] : B[k*ldb+n]
template <bool TA, bool TB, int EPI>
__global__ void __launch_bounds__(256)
sgemm_kernel(const float *__restrict__ A, int lda, const float *__restrict__ Bm, int ldb, float *__restrict__ C,
             int ldc, int M, int N, int K, const float *__restrict__ bias, int act,
             const float *__restrict__ yprev, int prev_act, int k_per_slab, int bias_row,
             const int *__restrict__ skip) {
  if (skip && *skip) return;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
  int k_begin = 0, k_end = K;
  if (EPI == EPI_PARTIAL) { k_begin = blockIdx.z * k_per_slab; k_end = min(K, k_begin + k_per_slab); }
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
  const bool do_bias = (EPI == EPI_PARTIAL) && bias_row && ty == 0 && blockIdx.y == 0;

  for (int k0 = k_begin; k0 < k_end; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      int m, k;
      if (TA) { m = idx & (BM - 1); k = idx / BM; } else { k = idx & (BK - 1); m = idx / BK; }
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < k_end) v = TA ? A[(int64_t)gk * lda + gm] : A[(int64_t)gm * lda + gk];
      As[k][m] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      int n, k;
      if (TB) { k = idx & (BK - 1); n = idx / BK; } else { n = idx & (BN - 1); k = idx / BN; }
      const int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < N && gk < k_end) v = TB ? Bm[(int64_t)gn * ldb + gk] : Bm[(int64_t)gk * ldb + gn];
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      if (do_bias) {
#pragma unroll
        for (int j = 0; j < 4; ++j) bsum[j] += bv[j];
      }
    }
    __syncthreads();
  }

  float *Cz = C;
  if (EPI == EPI_PARTIAL) Cz = C + (int64_t)blockIdx.z * (int64_t)(M + (bias_row ? 1 : 0)) * ldc;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (EPI == EPI_FWD) v = act_fwd_rt(act, v + bias[gn]);
      if (EPI == EPI_BWD_DATA && yprev) v *= act_bwd_from_out(prev_act, yprev[(int64_t)gm * ldc + gn]);
      Cz[(int64_t)gm * ldc + gn] = v;
    }
  }
  if (do_bias) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn < N) Cz[(int64_t)M * ldc + gn] = bsum[j];
    }
  }
}

__global__ void act_bwd_kernel(float *__restrict__ dy, const float *__restrict__ y, int64_t n, int act,
                               const int *__restrict__ skip) {
  if (skip && *skip) return;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dy[i] *= act_bwd_from_out(act, y[i]);
}

struct PartialTable {
  const float *src[CRUX_MAX_LAYERS];
  int slabs[CRUX_MAX_LAYERS];
  int64_t count[CRUX_MAX_LAYERS];   // (in+1)*out elements of layer l
  int64_t dst_off[CRUX_MAX_LAYERS];
  int n;
};
__global__ void reduce_partials_kernel(PartialTable t, float *__restrict__ grads, int accumulate,
                                       const int *__restrict__ skip) {
  if (skip && *skip) return;
  const int l = blockIdx.y;
  if (l >= t.n) return;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < t.count[l]; i += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    const float *p = t.src[l] + i;
    const int nz = t.slabs[l];
    const int64_t cnt = t.count[l];
    int z = 0;
    for (; z + 8 <= nz; z += 8) {   // eight loads in flight, added in slab order (the sum is that of the plain loop, bit for bit)
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldcs(p + (int64_t)(z + u) * cnt);
#pragma unroll
      for (int u = 0; u < 8; ++u) s += v[u];
    }
    for (; z < nz; ++z) s += __ldcs(p + (int64_t)z * cnt);
    float *d = grads + t.dst_off[l] + i;
    *d = accumulate ? (*d + s) : s;
  }
}

// ---- Adam -----------------------------------------------------------------------------------------
__global__ void gradnorm_kernel(AdamSegs segs, double *__restrict__ part, int *__restrict__ step_dev,
                                const int *__restrict__ skip) {
  if (skip && *skip) return;
  double s = 0.0;
  for (int q = 0; q < segs.n; ++q) {
    const float *g = segs.s[q].g;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < segs.s[q].n; i += (int64_t)gridDim.x * blockDim.x) {
      const double v = (double)g[i];
      s += v * v;
    }
  }
  __shared__ double sh[32];
  s = warp_sum_d(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
    s = warp_sum_d(s);
    if (threadIdx.x == 0) part[blockIdx.x] = s;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *step_dev += 1;
}

__global__ void adam_kernel(AdamSegs segs, const double *__restrict__ part, int nparts, double eta, double beta1,
                            double beta2, double eps, const int *__restrict__ step_dev, float *__restrict__ gnorm_out,
                            unsigned int *__restrict__ err_flags, const int *__restrict__ skip) {
  if (skip && *skip) return;
  __shared__ double s_norm2;
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < nparts; ++i) s += part[i];
    s_norm2 = s;
  }
  __syncthreads();
  const double n2 = s_norm2;
  if (blockIdx.x == 0 && threadIdx.x == 0 && gnorm_out) *gnorm_out = (float)sqrt(n2);
  if (isnan(n2)) {  // training.jl:20: error before Flux.update!
    if (threadIdx.x == 0) atomicOr(err_flags, CRUX_FLAG_NAN);
    return;
  }
  const int t = *step_dev;  // already counts this step
  const double c1 = 1.0 - pow(beta1, (double)t), c2 = 1.0 - pow(beta2, (double)t);
  for (int q = 0; q < segs.n; ++q) {
    const AdamSeg sg = segs.s[q];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < sg.n; i += (int64_t)gridDim.x * blockDim.x) {
      double g = (double)sg.g[i];
      if (segs.clip > 0.f) g = fmin(fmax(g, -(double)segs.clip), (double)segs.clip);   // ClipValue (examples/rl/atari.jl:10)
      const float mt = (float)(beta1 * (double)sg.m[i] + (1.0 - beta1) * g);
      const float vt = (float)(beta2 * (double)sg.v[i] + (1.0 - beta2) * g * g);
      sg.m[i] = mt; sg.v[i] = vt;
      const float delta = (float)((double)mt / c1 / (sqrt((double)vt / c2) + eps) * eta);
      sg.p[i] = sg.p[i] - delta;
    }
  }
}

__global__ void polyak_kernel(float *__restrict__ to, const float *__restrict__ from, int64_t n, float tau) {
  const float omt = 1.0f - tau;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    to[i] = tau * from[i] + omt * to[i];  // policies.jl:54
}

__global__ void concat2_kernel(const float *__restrict__ s, int sd, const float *__restrict__ a, int ad, int64_t B,
                               float *__restrict__ out) {
  const int d = sd + ad;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B * d; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / d; const int c = (int)(i % d);
    out[i] = c < sd ? s[b * sd + c] : a[b * ad + (c - sd)];
  }
}

__global__ void mse_head_kernel(const float *__restrict__ pred, const float *__restrict__ y, int64_t n, float inv_n,
                                float *__restrict__ dpred, double *__restrict__ part) {
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = pred[i] - y[i];
    s += (double)(d * d);
    dpred[i] = 2.0f * d * inv_n;
  }
  __shared__ double sh[32];
  s = warp_sum_d(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
    s = warp_sum_d(s);
    if (threadIdx.x == 0) part[blockIdx.x] = s;
  }
}
__global__ void sum_parts_kernel(const double *__restrict__ part, int n, double scale, float *__restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += part[i];
    *out = (float)(s * scale);
  }
}

}  // namespace

// gemm_tc5.cu: the same three GEMMs on tcgen05 (3xTF32) for shapes that fill a 128 x 64 tile; the FFMA kernel above serves the rest
bool gemm_tc5_eligible(int64_t M, int64_t N, int64_t K);
int gemm_tc5_fwd(crux_ctx *ctx, const float *x, int K, const float *W, int N, float *y, int64_t B, const float *b, int act, const int *skip);
int gemm_tc5_bwd_data(crux_ctx *ctx, const float *dz, int N, const float *W, float *dx, int K, int64_t B, const float *yprev, int prev_act, const int *skip);
int gemm_tc5_wgrad(crux_ctx *ctx, const float *x, int K, const float *dz, int N, float *part, int64_t B, int slabs, int rows_per_slab, const int *skip);

// ------------------------------------------------------------------------------------------------
int mlp_ensure_workspace(crux_mlp *mlp, int64_t B) {
  crux_ctx *ctx = mlp->ctx;
  if (B <= mlp->cap) return CRUX_OK;
  CRUX_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const int64_t cap = B + B / 8;
  for (int l = 0; l <= mlp->n_layers; ++l) {
    if (mlp->act[l]) cudaFree(mlp->act[l]);
    if (mlp->dz[l]) cudaFree(mlp->dz[l]);
    mlp->act[l] = mlp->dz[l] = nullptr;
    const size_t bytes = (size_t)cap * mlp->dims[l] * sizeof(float);
    if (l >= 1) {
      if (cudaMalloc((void **)&mlp->act[l], bytes) != cudaSuccess) return crux_set_err(ctx, CRUX_ERR_OOM, "mlp workspace act[%d] %zu B", l, bytes);
    }
    if (cudaMalloc((void **)&mlp->dz[l], bytes) != cudaSuccess) return crux_set_err(ctx, CRUX_ERR_OOM, "mlp workspace dz[%d] %zu B", l, bytes);
  }
  mlp->cap = cap;
  return CRUX_OK;
}

// Dense forward with a handful of outputs (the 256 -> 1 head of a critic, a 6-action head): y[b][n] = act(sum_k x[b][k] W[k][n] + bias[n]), N <= 8.
// One warp per row: lanes stride k (coalesced reads of the row), a shuffle tree per output.  The 64 x 64 tile kernel spends 26 us on the
// 2048 x 256 x 1 head of the SAC critics (32 CTAs, 63 of 64 tile columns idle); this takes the time of reading x once.
template <int NO>
__global__ void __launch_bounds__(256) skinny_fwd_kernel(const float *__restrict__ x, int K, const float *__restrict__ W, const float *__restrict__ bias,
                                                         float *__restrict__ y, int64_t B, int act, const int *__restrict__ skip) {
  if (skip && *skip) return;
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= B) return;
  const float *xr = x + row * K;
  float acc[NO];
#pragma unroll
  for (int n = 0; n < NO; ++n) acc[n] = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float xv = xr[k];
#pragma unroll
    for (int n = 0; n < NO; ++n) acc[n] = fmaf(xv, __ldg(W + (int64_t)k * NO + n), acc[n]);
  }
#pragma unroll
  for (int n = 0; n < NO; ++n) {
    float v = acc[n];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) y[row * NO + n] = act_fwd_rt(act, v + bias[n]);
  }
}

static int launch_fwd_layer(crux_mlp *mlp, int l /*1-based*/, const float *x, int64_t B, float *y, const int *skip) {
  crux_ctx *ctx = mlp->ctx;
  const int K = mlp->dims[l - 1], N = mlp->dims[l];
  const float *W = mlp->params + mlp->w_off[l - 1];
  const float *b = W + (int64_t)K * N;
  if (gemm_tc5_eligible(B, N, K)) return gemm_tc5_fwd(ctx, x, K, W, N, y, B, b, mlp->acts[l - 1], skip);
  if (N <= 8 && K >= 64 && B >= 256 && !getenv("CRUX_NO_SKINNY")) {   // (small problems stay on the tile kernel: same launch, nothing to gain)
    const unsigned g = (unsigned)cdiv(B, 8);
    switch (N) {
      case 1: skinny_fwd_kernel<1><<<g, 256, 0, ctx->stream>>>(x, K, W, b, y, B, mlp->acts[l - 1], skip); break;
      case 2: skinny_fwd_kernel<2><<<g, 256, 0, ctx->stream>>>(x, K, W, b, y, B, mlp->acts[l - 1], skip); break;
      case 3: skinny_fwd_kernel<3><<<g, 256, 0, ctx->stream>>>(x, K, W, b, y, B, mlp->acts[l - 1], skip); break;
      case 4: skinny_fwd_kernel<4><<<g, 256, 0, ctx->stream>>>(x, K, W, b, y, B, mlp->acts[l - 1], skip); break;
      case 5: skinny_fwd_kernel<5><<<g, 256, 0, ctx->stream>>>(x, K, W, b, y, B, mlp->acts[l - 1], skip); break;
      case 6: skinny_fwd_kernel<6><<<g, 256, 0, ctx->stream>>>(x, K, W, b, y, B, mlp->acts[l - 1], skip); break;
      case 7: skinny_fwd_kernel<7><<<g, 256, 0, ctx->stream>>>(x, K, W, b, y, B, mlp->acts[l - 1], skip); break;
      default: skinny_fwd_kernel<8><<<g, 256, 0, ctx->stream>>>(x, K, W, b, y, B, mlp->acts[l - 1], skip); break;
    }
    CRUX_LAUNCHED(ctx);
    return CRUX_OK;
  }
  dim3 grid((unsigned)cdiv(N, BN), (unsigned)cdiv(B, BM), 1);
  sgemm_kernel<false, false, EPI_FWD><<<grid, 256, 0, ctx->stream>>>(x, K, W, N, y, N, (int)B, N, K, b, mlp->acts[l - 1],
                                                                    nullptr, 0, 0, 0, skip);
  CRUX_LAUNCHED(ctx);
  return CRUX_OK;
}

int mlp_forward_keep(crux_mlp *mlp, const float *x, int64_t B, const int *skip) {
  int rc = mlp_ensure_workspace(mlp, B);
  if (rc) return rc;
  const float *in = x;
  for (int l = 1; l <= mlp->n_layers; ++l) {
    rc = launch_fwd_layer(mlp, l, in, B, mlp->act[l], skip);
    if (rc) return rc;
    in = mlp->act[l];
  }
  return CRUX_OK;
}

int mlp_forward_out(crux_mlp *mlp, const float *x, int64_t B, float *y) {
  int rc = mlp_ensure_workspace(mlp, B);
  if (rc) return rc;
  const float *in = x;
  for (int l = 1; l <= mlp->n_layers; ++l) {
    float *out = (l == mlp->n_layers) ? y : mlp->act[l];
    rc = launch_fwd_layer(mlp, l, in, B, out, nullptr);
    if (rc) return rc;
    in = out;
  }
  return CRUX_OK;
}

int mlp_backward(crux_mlp *mlp, const float *x, int64_t B, float *dY, bool need_dx, bool accumulate,
                 bool params_grad, const int *skip) {
  crux_ctx *ctx = mlp->ctx;
  const int L = mlp->n_layers;
  CRUX_REQUIRE(ctx, B <= mlp->cap, "mlp_backward: forward_keep must run first");
  // dz[L] = dY * act'(y_L)
  float *dz_cur = dY;
  if (mlp->acts[L - 1] != CRUX_ACT_IDENTITY) {
    const int64_t n = B * mlp->dims[L];
    act_bwd_kernel<<<(unsigned)i64min(cdiv(n, 256), (int64_t)ctx->num_sms * 8), 256, 0, ctx->stream>>>(
        dY, mlp->act[L], n, mlp->acts[L - 1], skip);
    CRUX_LAUNCHED(ctx);
  }
  PartialTable tab;
  tab.n = 0;
  if (params_grad) {
    // size the partial buffer
    size_t need = 0;
    int slabs[CRUX_MAX_LAYERS];
    int kps[CRUX_MAX_LAYERS];
    for (int l = 1; l <= L; ++l) {
      const int K = mlp->dims[l - 1], N = mlp->dims[l];
      const bool tc = gemm_tc5_eligible(K, N, B);                       // 128 x 64 tiles, 32-row k-tiles
      const int64_t tiles = tc ? cdiv(K, 128) * cdiv(N, 64) : cdiv(K, BM) * cdiv(N, BN);
      int64_t S = cdiv(2 * (int64_t)ctx->num_sms, tiles);
      S = i64max(1, i64min(S, cdiv(B, tc ? 128 : 64)));
      int64_t per = cdiv(cdiv(B, S), 32) * 32;
      S = cdiv(B, per);
      slabs[l - 1] = (int)S; kps[l - 1] = (int)per;
      need += (size_t)S * (K + 1) * N * sizeof(float);
    }
    if (need > mlp->partials_bytes) {
      CRUX_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      if (mlp->partials) cudaFree(mlp->partials);
      mlp->partials = nullptr; mlp->partials_bytes = 0;
      if (cudaMalloc((void **)&mlp->partials, need + need / 4) != cudaSuccess)
        return crux_set_err(ctx, CRUX_ERR_OOM, "mlp partials %zu B", need);
      mlp->partials_bytes = need + need / 4;
    }
    float *pp = mlp->partials;
    for (int l = 1; l <= L; ++l) {
      const int K = mlp->dims[l - 1], N = mlp->dims[l];
      tab.src[l - 1] = pp; tab.slabs[l - 1] = slabs[l - 1];
      tab.count[l - 1] = (int64_t)(K + 1) * N; tab.dst_off[l - 1] = mlp->w_off[l - 1];
      pp += (int64_t)slabs[l - 1] * (K + 1) * N;
      tab.n = l;
      (void)kps;
    }
    // launches below need kps: recompute inline
    for (int l = L; l >= 1; --l) {
      const int K = mlp->dims[l - 1], N = mlp->dims[l];
      const float *xin = (l == 1) ? x : mlp->act[l - 1];
      const float *dzl = (l == L) ? dz_cur : mlp->dz[l];
      if (gemm_tc5_eligible(K, N, B)) {
        int rc2 = gemm_tc5_wgrad(ctx, xin, K, dzl, N, (float *)tab.src[l - 1], B, slabs[l - 1], kps[l - 1], skip);
        if (rc2) return rc2;
      } else {
        dim3 grid((unsigned)cdiv(N, BN), (unsigned)cdiv(K, BM), (unsigned)slabs[l - 1]);
        sgemm_kernel<true, false, EPI_PARTIAL><<<grid, 256, 0, ctx->stream>>>(
            xin, K, dzl, N, (float *)tab.src[l - 1], N, K, N, (int)B, nullptr, 0, nullptr, 0, kps[l - 1], 1, skip);
        CRUX_LAUNCHED(ctx);
      }
      if (l > 1 || need_dx) {
        const float *W = mlp->params + mlp->w_off[l - 1];
        const float *yp = (l > 1) ? mlp->act[l - 1] : nullptr;
        const int pa = (l > 1) ? mlp->acts[l - 2] : 0;
        if (gemm_tc5_eligible(B, K, N)) {
          int rc2 = gemm_tc5_bwd_data(ctx, dzl, N, W, mlp->dz[l - 1], K, B, yp, pa, skip);
          if (rc2) return rc2;
        } else {
          dim3 g2((unsigned)cdiv(K, BN), (unsigned)cdiv(B, BM), 1);
          sgemm_kernel<false, true, EPI_BWD_DATA><<<g2, 256, 0, ctx->stream>>>(dzl, N, W, N, mlp->dz[l - 1], K, (int)B, K, N, nullptr, 0, yp, pa, 0, 0, skip);
          CRUX_LAUNCHED(ctx);
        }
      }
    }
    int64_t maxc = 0;
    for (int l = 0; l < L; ++l) maxc = i64max(maxc, tab.count[l]);
    dim3 rg((unsigned)i64min(cdiv(maxc, 256), 256), (unsigned)L);
    reduce_partials_kernel<<<rg, 256, 0, ctx->stream>>>(tab, mlp->grads, accumulate ? 1 : 0, skip);
    CRUX_LAUNCHED(ctx);
  } else {
    for (int l = L; l >= 1; --l) {
      const int K = mlp->dims[l - 1], N = mlp->dims[l];
      const float *dzl = (l == L) ? dz_cur : mlp->dz[l];
      if (l > 1 || need_dx) {
        const float *W = mlp->params + mlp->w_off[l - 1];
        const float *yp = (l > 1) ? mlp->act[l - 1] : nullptr;
        const int pa = (l > 1) ? mlp->acts[l - 2] : 0;
        if (gemm_tc5_eligible(B, K, N)) {
          int rc2 = gemm_tc5_bwd_data(ctx, dzl, N, W, mlp->dz[l - 1], K, B, yp, pa, skip);
          if (rc2) return rc2;
        } else {
          dim3 g2((unsigned)cdiv(K, BN), (unsigned)cdiv(B, BM), 1);
          sgemm_kernel<false, true, EPI_BWD_DATA><<<g2, 256, 0, ctx->stream>>>(dzl, N, W, N, mlp->dz[l - 1], K, (int)B, K, N, nullptr, 0, yp, pa, 0, 0, skip);
          CRUX_LAUNCHED(ctx);
        }
      }
    }
  }
  return CRUX_OK;
}

int adam_step_segments(crux_ctx *ctx, const AdamSegs &segs, double eta, double beta1, double beta2, double eps,
                       int *step_dev, float *gnorm_out_dev, const int *skip_dev, double *norm_part) {
  int64_t total = 0;
  for (int q = 0; q < segs.n; ++q) total += segs.s[q].n;
  const int blocks = (int)i64max(1, i64min(cdiv(total, 256 * 4), 512));
  gradnorm_kernel<<<blocks, 256, 0, ctx->stream>>>(segs, norm_part, step_dev, skip_dev);
  CRUX_LAUNCHED(ctx);
  adam_kernel<<<blocks, 256, 0, ctx->stream>>>(segs, norm_part, blocks, eta, beta1, beta2, eps, step_dev, gnorm_out_dev,
                                               ctx->flags_dev, skip_dev);
  CRUX_LAUNCHED(ctx);
  return CRUX_OK;
}

int mlp_adam_step(crux_mlp *mlp, float *gnorm_out_dev, const int *skip_dev) {
  AdamSegs segs;
  segs.n = 1;
  segs.s[0] = AdamSeg{mlp->params, mlp->grads, mlp->m, mlp->v, mlp->n_params};
  return adam_step_segments(mlp->ctx, segs, mlp->eta, mlp->beta1, mlp->beta2, mlp->eps, mlp->step_dev, gnorm_out_dev,
                            skip_dev, mlp->norm_part);
}

extern "C" int32_t crux_nccl_allreduce_f32(crux_ctx *ctx, float *buf, int64_t n);
int grads_allreduce(crux_ctx *ctx, float *g, int64_t n) {
  if (ctx->world <= 1) return CRUX_OK;
  return crux_nccl_allreduce_f32(ctx, g, n);
}

// ------------------------------------------------------------------------------------------------ ABI
extern "C" {

int32_t crux_mlp_create(crux_ctx *ctx, int32_t n_layers, const int32_t *dims, const int32_t *acts, crux_mlp **out) {
  if (!ctx) return CRUX_ERR_INVALID;
  CRUX_REQUIRE(ctx, out && dims && acts, "crux_mlp_create: NULL argument");
  CRUX_REQUIRE(ctx, n_layers >= 1 && n_layers <= CRUX_MAX_LAYERS, "crux_mlp_create: 1..8 layers supported");
  crux_mlp *m = new crux_mlp();
  m->ctx = ctx;
  m->n_layers = n_layers;
  int64_t off = 0;
  for (int l = 0; l <= n_layers; ++l) {
    if (dims[l] < 1) { delete m; return crux_set_err(ctx, CRUX_ERR_INVALID, "crux_mlp_create: dims[%d] < 1", l); }
    m->dims[l] = dims[l];
  }
  for (int l = 0; l < n_layers; ++l) {
    if (acts[l] < 0 || acts[l] > CRUX_ACT_RELU) { delete m; return crux_set_err(ctx, CRUX_ERR_INVALID, "crux_mlp_create: unknown activation %d", acts[l]); }
    m->acts[l] = acts[l];
    m->w_off[l] = off;
    off += (int64_t)(dims[l] + 1) * dims[l + 1];
  }
  m->n_params = off;
  const size_t bytes = (size_t)off * sizeof(float);
  const size_t gbytes = bytes + 128 * sizeof(float);  // tail: logΣ gradient + info sums ride the same all-reduce
  // params is padded: the fused kernels stage it with a 16-byte-granular TMA bulk copy
  if (cudaMalloc((void **)&m->params, bytes + 64) != cudaSuccess || cudaMalloc((void **)&m->grads, gbytes) != cudaSuccess ||
      cudaMalloc((void **)&m->m, bytes) != cudaSuccess || cudaMalloc((void **)&m->v, bytes) != cudaSuccess ||
      cudaMalloc((void **)&m->step_dev, 4 * sizeof(int)) != cudaSuccess ||   // [0] Adam steps, [1] ticket, [2] NaN seen (ppo_fused.cu reduce_adam_kernel)
      cudaMalloc((void **)&m->norm_part, 1024 * sizeof(double)) != cudaSuccess) {
    crux_mlp_destroy(m);
    return crux_set_err(ctx, CRUX_ERR_OOM, "crux_mlp_create: cudaMalloc failed");
  }
  cudaMemsetAsync(m->params, 0, bytes + 64, ctx->stream);
  cudaMemsetAsync(m->grads, 0, gbytes, ctx->stream);
  cudaMemsetAsync(m->m, 0, bytes, ctx->stream);
  cudaMemsetAsync(m->v, 0, bytes, ctx->stream);
  cudaMemsetAsync(m->step_dev, 0, 4 * sizeof(int), ctx->stream);
  cudaMemsetAsync(m->norm_part, 0, 1024 * sizeof(double), ctx->stream);   // incl. the β-power cache of the fused PPO tails (ppo_fused.cu)
  *out = m;
  return CRUX_OK;
}

int32_t crux_mlp_destroy(crux_mlp *m) {
  if (!m) return CRUX_OK;
  cudaStreamSynchronize(m->ctx->stream);
  if (m->frag) cudaFree(m->frag);
  cudaFree(m->params); cudaFree(m->grads); cudaFree(m->m); cudaFree(m->v); cudaFree(m->step_dev); cudaFree(m->norm_part);
  for (int l = 0; l <= CRUX_MAX_LAYERS; ++l) { if (m->act[l]) cudaFree(m->act[l]); if (m->dz[l]) cudaFree(m->dz[l]); }
  if (m->partials) cudaFree(m->partials);
  delete m;
  return CRUX_OK;
}

int32_t crux_mlp_num_params(crux_mlp *m, int64_t *out) {
  if (!m || !out) return CRUX_ERR_INVALID;
  *out = m->n_params;
  return CRUX_OK;
}
int32_t crux_mlp_set_params(crux_mlp *m, const float *flat_host) {
  if (!m || !flat_host) return CRUX_ERR_INVALID;
  CRUX_CHECK_CUDA(m->ctx, cudaMemcpyAsync(m->params, flat_host, (size_t)m->n_params * sizeof(float), cudaMemcpyHostToDevice, m->ctx->stream));
  CRUX_CHECK_CUDA(m->ctx, cudaStreamSynchronize(m->ctx->stream));
  return CRUX_OK;
}
int32_t crux_mlp_get_params(crux_mlp *m, float *flat_host) {
  if (!m || !flat_host) return CRUX_ERR_INVALID;
  CRUX_CHECK_CUDA(m->ctx, cudaMemcpyAsync(flat_host, m->params, (size_t)m->n_params * sizeof(float), cudaMemcpyDeviceToHost, m->ctx->stream));
  CRUX_CHECK_CUDA(m->ctx, cudaStreamSynchronize(m->ctx->stream));
  return CRUX_OK;
}
int32_t crux_mlp_params_ptr(crux_mlp *m, float **dev_out) {
  if (!m || !dev_out) return CRUX_ERR_INVALID;
  *dev_out = m->params;
  return CRUX_OK;
}
int32_t crux_mlp_grads_ptr(crux_mlp *m, float **dev_out) {
  if (!m || !dev_out) return CRUX_ERR_INVALID;
  *dev_out = m->grads;
  return CRUX_OK;
}
int32_t crux_mlp_set_adam(crux_mlp *m, double eta, double beta1, double beta2, double eps) {
  if (!m) return CRUX_ERR_INVALID;
  m->eta = eta; m->beta1 = beta1; m->beta2 = beta2; m->eps = eps;
  const size_t bytes = (size_t)m->n_params * sizeof(float);
  CRUX_CHECK_CUDA(m->ctx, cudaMemsetAsync(m->m, 0, bytes, m->ctx->stream));
  CRUX_CHECK_CUDA(m->ctx, cudaMemsetAsync(m->v, 0, bytes, m->ctx->stream));
  CRUX_CHECK_CUDA(m->ctx, cudaMemsetAsync(m->step_dev, 0, 4 * sizeof(int), m->ctx->stream));
  return CRUX_OK;
}

int32_t crux_mlp_forward(crux_mlp *m, const float *x, int64_t B, float *y) {
  if (!m) return CRUX_ERR_INVALID;
  CRUX_REQUIRE(m->ctx, B >= 0, "crux_mlp_forward: negative batch");
  if (B == 0) return CRUX_OK;
  CRUX_REQUIRE(m->ctx, x && y, "crux_mlp_forward: NULL pointer");
  int handled = 0;
  const int rc = mlp_forward_fused(m, x, B, y, &handled);
  if (rc || handled) return rc;
  return mlp_forward_out(m, x, B, y);
}

int32_t crux_value_next(crux_mlp *m, const float *sp, const float *s, const float *v_s, int64_t T, int64_t N, float *v_sp) {
  if (!m) return CRUX_ERR_INVALID;
  CRUX_REQUIRE(m->ctx, T >= 0 && N >= 0, "crux_value_next: negative shape");
  if (T == 0 || N == 0) return CRUX_OK;
  CRUX_REQUIRE(m->ctx, sp && s && v_s && v_sp, "crux_value_next: NULL pointer");
  CRUX_REQUIRE(m->ctx, m->dims[m->n_layers] == 1, "crux_value_next: needs a value network (one output)");
  int handled = 0;
  const int rc = mlp_value_next_fused(m, sp, s, v_s, T, N, v_sp, &handled);
  if (rc || handled) return rc;
  return mlp_forward_out(m, sp, T * N, v_sp);
}

int32_t crux_mlp_forward_sa(crux_mlp *m, const float *s, int32_t sdim, const float *a, int32_t adim, int64_t B, float *y) {
  if (!m) return CRUX_ERR_INVALID;
  crux_ctx *ctx = m->ctx;
  CRUX_REQUIRE(ctx, sdim + adim == m->dims[0], "crux_mlp_forward_sa: sdim + adim != input width (policies.jl:96 vcat)");
  if (B <= 0) return CRUX_OK;
  float *cat = (float *)crux_scratch(ctx, 2, (size_t)B * (sdim + adim) * sizeof(float));
  if (!cat) return CRUX_ERR_OOM;
  const int64_t n = B * (sdim + adim);
  concat2_kernel<<<(unsigned)i64min(cdiv(n, 256), (int64_t)ctx->num_sms * 8), 256, 0, ctx->stream>>>(s, sdim, a, adim, B, cat);
  CRUX_LAUNCHED(ctx);
  return mlp_forward_out(m, cat, B, y);
}

int32_t crux_mlp_copy(crux_mlp *to, crux_mlp *from) {
  if (!to || !from) return CRUX_ERR_INVALID;
  CRUX_REQUIRE(to->ctx, to->n_params == from->n_params, "crux_mlp_copy: parameter count mismatch");
  CRUX_CHECK_CUDA(to->ctx, cudaMemcpyAsync(to->params, from->params, (size_t)to->n_params * sizeof(float), cudaMemcpyDeviceToDevice, to->ctx->stream));
  return CRUX_OK;
}

int32_t crux_mlp_polyak(crux_mlp *to, crux_mlp *from, float tau) {
  if (!to || !from) return CRUX_ERR_INVALID;
  crux_ctx *ctx = to->ctx;
  CRUX_REQUIRE(ctx, to->n_params == from->n_params, "crux_mlp_polyak: parameter count mismatch");
  polyak_kernel<<<(unsigned)i64min(cdiv(to->n_params, 256), (int64_t)ctx->num_sms * 4), 256, 0, ctx->stream>>>(
      to->params, from->params, to->n_params, tau);
  CRUX_LAUNCHED(ctx);
  return CRUX_OK;
}

int32_t crux_mlp_train_mse(crux_mlp *m, const float *x, const float *y, int64_t B, float *info_out_host) {
  if (!m) return CRUX_ERR_INVALID;
  crux_ctx *ctx = m->ctx;
  CRUX_REQUIRE(ctx, B > 0 && x && y, "crux_mlp_train_mse: bad arguments");
  const int L = m->n_layers;
  const int64_t n = B * m->dims[L];
  int rc = mlp_forward_keep(m, x, B, nullptr);
  if (rc) return rc;
  const int blocks = (int)i64min(cdiv(n, 256), 512);
  char *sc = (char *)crux_scratch(ctx, 3, 512 * sizeof(double) + 64);
  if (!sc) return CRUX_ERR_OOM;
  float *info_dev = (float *)sc;
  double *part = (double *)(sc + 64);
  const int64_t nglob = n * ctx->world;
  mse_head_kernel<<<blocks, 256, 0, ctx->stream>>>(m->act[L], y, n, 1.0f / (float)nglob, m->dz[L], part);
  CRUX_LAUNCHED(ctx);
  sum_parts_kernel<<<1, 32, 0, ctx->stream>>>(part, blocks, 1.0 / (double)nglob, info_dev);
  CRUX_LAUNCHED(ctx);
  rc = mlp_backward(m, x, B, m->dz[L], false, false, true, nullptr);
  if (rc) return rc;
  if (ctx->world > 1) {
    rc = grads_allreduce(ctx, m->grads, m->n_params); if (rc) return rc;
    rc = grads_allreduce(ctx, info_dev, 1); if (rc) return rc;
  }
  rc = mlp_adam_step(m, info_dev + 1, nullptr);
  if (rc) return rc;
  if (info_out_host) {
    CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(info_out_host, info_dev, 2 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    rc = crux_ctx_check(ctx);
    if (rc) return rc;
  }
  return CRUX_OK;
}

}  // extern "C"
