// Device-resident ExperienceBuffer (src/experience_buffer.jl): structure-of-arrays ring buffer in HBM,
// ring push (mod1 index arithmetic, bit-exact), uniform / prioritized sampling and priority updates.
// The reference keeps priorities on the CPU as a flat Vector{Float32} + cumsum + searchsortedfirst
// (experience_buffer.jl:38-50,290-349); the same algorithm runs here as a device prefix scan, a batched
// binary search with Float64 thresholds, and an atomic max/min -- integer results are exact given the
// same prefix array.
#include "common.cuh"
#include <vector>

struct BufCol {
  int id, dtype;
  int64_t rowlen;
  size_t rowbytes;
  double init;
  void *data;
};

struct crux_buffer {
  crux_ctx *ctx = nullptr;
  int64_t capacity = 0, elements = 0, next_ind = 0, total_count = 0;
  std::vector<BufCol> cols;
  int col_index[64];
  bool prioritized = false;
  float alpha = 0.6f;
  float *prs = nullptr, *cumsum = nullptr;
  bool cumsum_valid = false;
  float *pstate = nullptr;        // device {max_priority, min_priority}
  float *pstate_pinned = nullptr;
  int32_t *indices = nullptr;     // last sample's indices (device)
  int64_t n_indices = 0, indices_cap = 0;
  int32_t *owner = nullptr;       // duplicate resolution for update_priorities!
  float *block_sums = nullptr;    // scan scratch
  int64_t block_cap = 0;
};

namespace {

size_t dtype_size(int dt) { return dt == CRUX_U8 ? 1 : dt == CRUX_I64 ? 8 : 4; }

__global__ void fill_kernel(void *p, int dtype, int64_t n, double v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (dtype == CRUX_U8) ((uint8_t *)p)[i] = (uint8_t)v;
    else if (dtype == CRUX_F32) ((float *)p)[i] = (float)v;
    else if (dtype == CRUX_I32) ((int32_t *)p)[i] = (int32_t)v;
    else ((int64_t *)p)[i] = (int64_t)v;
  }
}

// dst row ((start + j) % C) <- src row (ids ? ids[j] : j), j in [j0, n).  W = bytes per vector element.
template <typename V>
__global__ void scatter_rows_kernel(V *__restrict__ dst, const V *__restrict__ src, const int32_t *__restrict__ ids,
                                    int64_t j0, int64_t n, int64_t start, int64_t C, int64_t per_row) {
  const int64_t total = (n - j0) * per_row;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = j0 + i / per_row, o = i % per_row;
    const int64_t srow = ids ? (int64_t)ids[j] : j;
    const int64_t drow = (start + j) % C;  // mod1.(next_ind:next_ind+N-1, C), 0-based
    dst[drow * per_row + o] = src[srow * per_row + o];
  }
}
template <typename V>
__global__ void gather_rows_kernel(V *__restrict__ dst, const V *__restrict__ src, const int32_t *__restrict__ ids, int64_t n,
                                   int64_t per_row) {
  const int64_t total = n * per_row;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = i / per_row, o = i % per_row;
    dst[i] = src[(int64_t)ids[j] * per_row + o];
  }
}

// new rows get max_priority (experience_buffer.jl:254): v = max_priority*ones(N) is Float64 in the reference
__global__ void push_priorities_kernel(float *__restrict__ prs, float *__restrict__ pstate, int64_t j0, int64_t n, int64_t start,
                                       int64_t C, float alpha) {
  const double val = (double)pstate[0] + (double)1.1920929e-07f;  // v[i] + eps(Float32)
  const float pr = (float)pow(val, (double)alpha);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n - j0; i += (int64_t)gridDim.x * blockDim.x)
    prs[(start + j0 + i) % C] = pr;
}
__global__ void push_pstate_kernel(float *__restrict__ pstate) {
  const double val = (double)pstate[0] + (double)1.1920929e-07f;
  pstate[0] = (float)fmax(val, (double)pstate[0]);
  pstate[1] = (float)fmin(val, (double)pstate[1]);
}

__global__ void owner_mark_kernel(int32_t *__restrict__ owner, const int32_t *__restrict__ idx, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicMax(owner + idx[i], (int32_t)i);
}
__global__ void update_priorities_kernel(float *__restrict__ prs, float *__restrict__ pstate, int32_t *__restrict__ owner,
                                         const int32_t *__restrict__ idx, const float *__restrict__ v, int64_t n, float alpha) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float val = 0.f;
  bool live = i < n;
  if (live) {
    val = v[i] + 1.1920929e-07f;  // experience_buffer.jl:293 (Float32 + eps(Float32))
    if (owner[idx[i]] == (int32_t)i) prs[idx[i]] = (float)pow((double)val, (double)alpha);  // last write wins
  }
  // max / min over the UN-exponentiated values (:297-298); positive floats order like their bit patterns
  float mx = live ? val : 0.f, mn = live ? val : INFINITY;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  }
  if ((threadIdx.x & 31) == 0) {
    if (mx > 0.f) atomicMax((int *)pstate, __float_as_int(mx));
    if (mn > 0.f) atomicMin((int *)pstate + 1, __float_as_int(mn));
  }
}
__global__ void owner_reset_kernel(int32_t *__restrict__ owner, const int32_t *__restrict__ idx, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) owner[idx[i]] = -1;
}

// ---- inclusive scan, 3 phases, blocks of 2048 elements (256 threads x 8) -------------------------
constexpr int SCAN_T = 256, SCAN_E = 8, SCAN_B = SCAN_T * SCAN_E;
__device__ __forceinline__ float block_scan_excl(float v, float *total) {  // exclusive scan of one value per thread
  __shared__ float wsum[SCAN_T / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const float t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  if (w == 0) {
    float x = lane < SCAN_T / 32 ? wsum[lane] : 0.f;
#pragma unroll
    for (int o = 1; o < SCAN_T / 32; o <<= 1) { const float t = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += t; }
    if (lane < SCAN_T / 32) wsum[lane] = x;
  }
  __syncthreads();
  const float woff = w ? wsum[w - 1] : 0.f;
  if (total) *total = wsum[SCAN_T / 32 - 1];
  __syncthreads();
  return woff + inc - v;
}
__global__ void scan_block_sums_kernel(const float *__restrict__ x, int64_t n, float *__restrict__ bsum) {
  const int64_t base = (int64_t)blockIdx.x * SCAN_B + (int64_t)threadIdx.x * SCAN_E;
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < SCAN_E; ++k) if (base + k < n) s += x[base + k];
  float tot;
  block_scan_excl(s, &tot);
  if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}
__global__ void scan_top_kernel(float *__restrict__ bsum, int nb) {  // single block, sequential chunks of SCAN_T
  __shared__ float carry;
  if (threadIdx.x == 0) carry = 0.f;
  __syncthreads();
  for (int b0 = 0; b0 < nb; b0 += SCAN_T) {
    const int i = b0 + threadIdx.x;
    const float v = i < nb ? bsum[i] : 0.f;
    float tot;
    const float ex = block_scan_excl(v, &tot);
    const float c = carry;
    if (i < nb) bsum[i] = c + ex;  // exclusive prefix of block sums
    __syncthreads();
    if (threadIdx.x == 0) carry = c + tot;
    __syncthreads();
  }
}
__global__ void scan_apply_kernel(const float *__restrict__ x, int64_t n, const float *__restrict__ bsum, float *__restrict__ out) {
  const int64_t base = (int64_t)blockIdx.x * SCAN_B + (int64_t)threadIdx.x * SCAN_E;
  float v[SCAN_E];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < SCAN_E; ++k) { v[k] = base + k < n ? x[base + k] : 0.f; s += v[k]; }
  const float ex = block_scan_excl(s, nullptr) + bsum[blockIdx.x];
  float run = ex;
#pragma unroll
  for (int k = 0; k < SCAN_E; ++k) { run += v[k]; if (base + k < n) out[base + k] = run; }
}

__global__ void uniform_ids_kernel(int32_t *__restrict__ ids, int64_t B, int64_t len, uint64_t seed, uint64_t ctr) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const Philox4 r = philox4x32_10(seed, ctr, (uint64_t)i);
  ids[i] = (int32_t)min((int64_t)(u64_to_unit(r.x, r.y) * (double)len), len - 1);  // rand(1:length(source)) 0-based
}

// prioritized_sample! (experience_buffer.jl:333-348)
__global__ void per_sample_kernel(const float *__restrict__ cumsum, const float *__restrict__ prs, const float *__restrict__ pstate,
                                  int64_t N, int64_t B, float beta, const double *__restrict__ u_in, uint64_t seed, uint64_t ctr,
                                  int32_t *__restrict__ ids, float *__restrict__ weight_col, int64_t weight_rowlen) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= B) return;
  const float ptot = cumsum[N - 1];
  const float dp = ptot / (float)B;  // Float32 / Int
  double u;
  if (u_in) u = u_in[j];
  else { const Philox4 r = philox4x32_10(seed, ctr, (uint64_t)j); u = u64_to_unit(r.x, r.y); }
  const double x = ((double)(j + 1) + u - 1.0) * (double)dp;
  int64_t lo = 0, hi = N;  // searchsortedfirst: first i with cumsum[i] >= x
  while (lo < hi) {
    const int64_t mid = lo + ((hi - lo) >> 1);
    if ((double)cumsum[mid] < x) lo = mid + 1; else hi = mid;
  }
  if (lo >= N) lo = N - 1;  // the reference would throw a BoundsError here; clamp
  ids[j] = (int32_t)lo;
  const float pmin = pstate[1] / ptot;
  const float max_w = (float)pow((double)(pmin * (float)N), -(double)beta);
  const float w = (float)pow((double)(((float)N * prs[lo]) / ptot), (double)beta);
  weight_col[lo * weight_rowlen] = w / max_w;
}

int launch_copy_rows(crux_ctx *ctx, void *dst, const void *src, const int32_t *ids, int64_t j0, int64_t n, int64_t start,
                     int64_t C, size_t rowbytes, bool scatter) {
  if (n - j0 <= 0) return CRUX_OK;
  int w = 1;
  if (rowbytes % 16 == 0 && ((uintptr_t)dst % 16 == 0) && ((uintptr_t)src % 16 == 0)) w = 16;
  else if (rowbytes % 4 == 0 && ((uintptr_t)dst % 4 == 0) && ((uintptr_t)src % 4 == 0)) w = 4;
  const int64_t per = (int64_t)(rowbytes / w);
  const int64_t total = (n - j0) * per;
  const unsigned blocks = (unsigned)i64max(1, i64min(cdiv(total, 256), (int64_t)ctx->num_sms * 16));
  if (scatter) {
    if (w == 16) scatter_rows_kernel<uint4><<<blocks, 256, 0, ctx->stream>>>((uint4 *)dst, (const uint4 *)src, ids, j0, n, start, C, per);
    else if (w == 4) scatter_rows_kernel<uint32_t><<<blocks, 256, 0, ctx->stream>>>((uint32_t *)dst, (const uint32_t *)src, ids, j0, n, start, C, per);
    else scatter_rows_kernel<uint8_t><<<blocks, 256, 0, ctx->stream>>>((uint8_t *)dst, (const uint8_t *)src, ids, j0, n, start, C, per);
  } else {
    if (w == 16) gather_rows_kernel<uint4><<<blocks, 256, 0, ctx->stream>>>((uint4 *)dst, (const uint4 *)src, ids, n, per);
    else if (w == 4) gather_rows_kernel<uint32_t><<<blocks, 256, 0, ctx->stream>>>((uint32_t *)dst, (const uint32_t *)src, ids, n, per);
    else gather_rows_kernel<uint8_t><<<blocks, 256, 0, ctx->stream>>>((uint8_t *)dst, (const uint8_t *)src, ids, n, per);
  }
  CRUX_LAUNCHED(ctx);
  return CRUX_OK;
}

int push_bookkeeping(crux_buffer *b, int64_t n, int64_t j0, int64_t start) {
  crux_ctx *ctx = b->ctx;
  if (b->prioritized && n > 0) {
    const unsigned blocks = (unsigned)i64max(1, i64min(cdiv(n - j0, 256), 1024));
    push_priorities_kernel<<<blocks, 256, 0, ctx->stream>>>(b->prs, b->pstate, j0, n, start, b->capacity, b->alpha);
    CRUX_LAUNCHED(ctx);
    // the reference's loop applies max/min once per pushed row; the value is identical every time
    push_pstate_kernel<<<1, 1, 0, ctx->stream>>>(b->pstate);
    CRUX_LAUNCHED(ctx);
    b->cumsum_valid = false;
  }
  b->total_count += n;
  b->elements = i64min(b->capacity, b->elements + n);
  b->next_ind = (b->next_ind + n) % b->capacity;
  return CRUX_OK;
}

int ensure_indices(crux_buffer *b, int64_t B) {
  if (b->indices_cap >= B) return CRUX_OK;
  crux_ctx *ctx = b->ctx;
  CRUX_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (b->indices) cudaFree(b->indices);
  b->indices = nullptr; b->indices_cap = 0;
  if (cudaMalloc((void **)&b->indices, (size_t)(B + 64) * sizeof(int32_t)) != cudaSuccess) return crux_set_err(ctx, CRUX_ERR_OOM, "indices");
  b->indices_cap = B + 64;
  return CRUX_OK;
}

}  // namespace

extern "C" {

int32_t crux_buffer_create(crux_ctx *ctx, int64_t capacity, int32_t n_cols, const crux_col_desc *cols, int32_t prioritized,
                           float alpha, crux_buffer **out) {
  if (!ctx) return CRUX_ERR_INVALID;
  CRUX_REQUIRE(ctx, out && cols && n_cols >= 1 && n_cols <= 64, "crux_buffer_create: bad columns");
  CRUX_REQUIRE(ctx, capacity >= 1 && capacity < (1ll << 31), "crux_buffer_create: capacity must be 1..2^31-1");
  crux_buffer *b = new crux_buffer();
  b->ctx = ctx; b->capacity = capacity; b->prioritized = prioritized != 0; b->alpha = alpha;
  for (int i = 0; i < 64; ++i) b->col_index[i] = -1;
  for (int i = 0; i < n_cols; ++i) {
    const crux_col_desc &d = cols[i];
    if (d.id < 0 || d.id >= 64 || b->col_index[d.id] >= 0 || d.rowlen < 0 || d.dtype < 0 || d.dtype > CRUX_I64) {
      crux_buffer_destroy(b);
      return crux_set_err(ctx, CRUX_ERR_INVALID, "crux_buffer_create: bad column descriptor %d", i);
    }
    BufCol c;
    c.id = d.id; c.dtype = d.dtype; c.rowlen = d.rowlen; c.rowbytes = (size_t)d.rowlen * dtype_size(d.dtype); c.init = d.init;
    c.data = nullptr;
    const size_t bytes = (size_t)capacity * c.rowbytes;
    cudaError_t e = cudaMalloc(&c.data, bytes ? bytes : 16);
    if (e != cudaSuccess) {
      crux_buffer_destroy(b);
      return crux_set_err(ctx, CRUX_ERR_OOM, "crux_buffer_create: column %d needs %zu bytes: %s", d.id, bytes, cudaGetErrorString(e));
    }
    b->col_index[d.id] = (int)b->cols.size();
    b->cols.push_back(c);
    const int64_t nel = capacity * d.rowlen;
    if (nel > 0) {
      if (d.init == 0.0) cudaMemsetAsync(c.data, 0, bytes, ctx->stream);
      else {
        fill_kernel<<<(unsigned)i64min(cdiv(nel, 256), (int64_t)ctx->num_sms * 8), 256, 0, ctx->stream>>>(c.data, d.dtype, nel, d.init);
        ctx->launches++;
      }
    }
  }
  if (b->prioritized) {
    if (cudaMalloc((void **)&b->prs, capacity * sizeof(float)) != cudaSuccess || cudaMalloc((void **)&b->cumsum, capacity * sizeof(float)) != cudaSuccess ||
        cudaMalloc((void **)&b->pstate, 2 * sizeof(float)) != cudaSuccess || cudaMalloc((void **)&b->owner, capacity * sizeof(int32_t)) != cudaSuccess ||
        cudaMallocHost((void **)&b->pstate_pinned, 2 * sizeof(float)) != cudaSuccess) {
      crux_buffer_destroy(b);
      return crux_set_err(ctx, CRUX_ERR_OOM, "crux_buffer_create: priority arrays");
    }
    b->block_cap = cdiv(capacity, SCAN_B) + 1;
    if (cudaMalloc((void **)&b->block_sums, b->block_cap * sizeof(float)) != cudaSuccess) { crux_buffer_destroy(b); return crux_set_err(ctx, CRUX_ERR_OOM, "scan scratch"); }
    cudaMemsetAsync(b->prs, 0, capacity * sizeof(float), ctx->stream);
    cudaMemsetAsync(b->owner, 0xFF, capacity * sizeof(int32_t), ctx->stream);
    const float init[2] = {1.0f, INFINITY};  // PriorityParams defaults (experience_buffer.jl:44-45)
    cudaMemcpyAsync(b->pstate, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
  }
  *out = b;
  return CRUX_OK;
}

int32_t crux_buffer_destroy(crux_buffer *b) {
  if (!b) return CRUX_OK;
  cudaStreamSynchronize(b->ctx->stream);
  for (auto &c : b->cols) if (c.data) cudaFree(c.data);
  if (b->prs) cudaFree(b->prs);
  if (b->cumsum) cudaFree(b->cumsum);
  if (b->pstate) cudaFree(b->pstate);
  if (b->pstate_pinned) cudaFreeHost(b->pstate_pinned);
  if (b->owner) cudaFree(b->owner);
  if (b->block_sums) cudaFree(b->block_sums);
  if (b->indices) cudaFree(b->indices);
  delete b;
  return CRUX_OK;
}

int32_t crux_buffer_col(crux_buffer *b, int32_t col_id, void **dev_ptr, int64_t *rowlen, int32_t *dtype) {
  if (!b) return CRUX_ERR_INVALID;
  CRUX_REQUIRE(b->ctx, col_id >= 0 && col_id < 64 && b->col_index[col_id] >= 0, "crux_buffer_col: no such column (KeyError)");
  const BufCol &c = b->cols[b->col_index[col_id]];
  if (dev_ptr) *dev_ptr = c.data;
  if (rowlen) *rowlen = c.rowlen;
  if (dtype) *dtype = c.dtype;
  return CRUX_OK;
}

int32_t crux_buffer_state(crux_buffer *b, int64_t *elements, int64_t *next_ind, int64_t *total_count, int64_t *capacity) {
  if (!b) return CRUX_ERR_INVALID;
  if (elements) *elements = b->elements;
  if (next_ind) *next_ind = b->next_ind;
  if (total_count) *total_count = b->total_count;
  if (capacity) *capacity = b->capacity;
  return CRUX_OK;
}

int32_t crux_buffer_clear(crux_buffer *b) {
  if (!b) return CRUX_ERR_INVALID;
  crux_ctx *ctx = b->ctx;
  b->elements = 0; b->next_ind = 0; b->total_count = 0; b->n_indices = 0;
  if (b->prioritized) {  // PriorityParams(capacity, old): zero priorities, keep α/β/max_priority, min_priority = Inf
    CRUX_CHECK_CUDA(ctx, cudaMemsetAsync(b->prs, 0, b->capacity * sizeof(float), ctx->stream));
    const float inf = INFINITY;
    CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(b->pstate + 1, &inf, sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CRUX_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    b->cumsum_valid = false;
  }
  return CRUX_OK;
}

int32_t crux_buffer_push(crux_buffer *b, int64_t n_rows, int32_t n_src_cols, const int32_t *col_ids, const void *const *col_ptrs,
                         int32_t src_on_host, const int32_t *ids, int64_t *first_index_out) {
  if (!b) return CRUX_ERR_INVALID;
  crux_ctx *ctx = b->ctx;
  CRUX_REQUIRE(ctx, n_rows >= 0 && n_src_cols >= 0, "crux_buffer_push: negative count");
  if (first_index_out) *first_index_out = b->next_ind;
  if (n_rows == 0) return CRUX_OK;
  CRUX_REQUIRE(ctx, n_src_cols == 0 || (col_ids && col_ptrs), "crux_buffer_push: NULL column table");
  const int64_t C = b->capacity, start = b->next_ind;
  const int64_t j0 = n_rows > C ? n_rows - C : 0;  // N > capacity: the later rows win (experience_buffer.jl:236,249-252)
  const int32_t *ids_dev = ids;
  if (ids && src_on_host) {
    int32_t *tmp = (int32_t *)crux_scratch(ctx, 5, (size_t)n_rows * sizeof(int32_t));
    if (!tmp) return CRUX_ERR_OOM;
    CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(tmp, ids, (size_t)n_rows * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    CRUX_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // pageable source
    ids_dev = tmp;
  }
  for (int k = 0; k < n_src_cols; ++k) {
    const int id = col_ids[k];
    if (id < 0 || id >= 64 || b->col_index[id] < 0) continue;  // keys(b) drives the loop; extra source keys are ignored
    const BufCol &c = b->cols[b->col_index[id]];
    if (c.rowbytes == 0) continue;
    const void *src = col_ptrs[k];
    CRUX_REQUIRE(ctx, src, "crux_buffer_push: NULL source column");
    if (src_on_host) {
      if (!ids) {
        // contiguous source rows j0..n-1 -> at most two ring segments, copied straight into place
        const int64_t first = (start + j0) % C, cnt = n_rows - j0;
        const int64_t seg1 = i64min(cnt, C - first);
        CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync((char *)c.data + first * c.rowbytes, (const char *)src + j0 * c.rowbytes, seg1 * c.rowbytes,
                                             cudaMemcpyHostToDevice, ctx->stream));
        if (cnt > seg1)
          CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(c.data, (const char *)src + (j0 + seg1) * c.rowbytes, (cnt - seg1) * c.rowbytes,
                                               cudaMemcpyHostToDevice, ctx->stream));
      } else {
        // gather on the host side of the bus is the caller's job only in the reference; here stage then scatter
        int64_t max_id = 0;
        for (int64_t j = 0; j < n_rows; ++j) max_id = i64max(max_id, ids[j]);
        void *stage = crux_scratch(ctx, 6, (size_t)(max_id + 1) * c.rowbytes);
        if (!stage) return CRUX_ERR_OOM;
        CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(stage, src, (size_t)(max_id + 1) * c.rowbytes, cudaMemcpyHostToDevice, ctx->stream));
        int rc = launch_copy_rows(ctx, c.data, stage, ids_dev, j0, n_rows, start, C, c.rowbytes, true);
        if (rc) return rc;
        CRUX_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // stage is reused by the next column
      }
    } else {
      int rc = launch_copy_rows(ctx, c.data, src, ids_dev, j0, n_rows, start, C, c.rowbytes, true);
      if (rc) return rc;
    }
  }
  return push_bookkeeping(b, n_rows, j0, start);
}

int32_t crux_buffer_push_from(crux_buffer *t, crux_buffer *s, int64_t n_rows, const int32_t *ids_dev) {
  if (!t || !s) return CRUX_ERR_INVALID;
  crux_ctx *ctx = t->ctx;
  CRUX_REQUIRE(ctx, n_rows >= 0, "crux_buffer_push_from: negative count");
  if (n_rows == 0) return CRUX_OK;
  if (!ids_dev) CRUX_REQUIRE(ctx, n_rows <= s->elements, "crux_buffer_push_from: more rows than the source holds");
  const int64_t C = t->capacity, start = t->next_ind;
  const int64_t j0 = n_rows > C ? n_rows - C : 0;
  for (auto &c : t->cols) {
    if (s->col_index[c.id] < 0) continue;
    const BufCol &sc = s->cols[s->col_index[c.id]];
    CRUX_REQUIRE(ctx, sc.rowbytes == c.rowbytes, "crux_buffer_push_from: row shape mismatch (experience_buffer.jl:251 @assert)");
    if (c.rowbytes == 0) continue;
    int rc = launch_copy_rows(ctx, c.data, sc.data, ids_dev, j0, n_rows, start, C, c.rowbytes, true);
    if (rc) return rc;
  }
  return push_bookkeeping(t, n_rows, j0, start);
}

int32_t crux_buffer_last_n_indices(crux_buffer *b, int64_t N, int64_t *out_host, int64_t *n_out) {
  if (!b || !out_host || !n_out) return CRUX_ERR_INVALID;
  N = i64min(b->elements, i64max(N, 0));
  const int64_t C = b->capacity;
  const int64_t start = ((b->next_ind - N) % C + C) % C;  // mod1(next_ind - N, C) 0-based
  for (int64_t j = 0; j < N; ++j) out_host[j] = (start + j) % C;
  *n_out = N;
  return CRUX_OK;
}

int32_t crux_buffer_sample_uniform(crux_buffer *t, crux_buffer *s, int64_t B, const int32_t *ids_in_host, uint64_t seed, uint64_t ctr) {
  if (!t || !s) return CRUX_ERR_INVALID;
  crux_ctx *ctx = t->ctx;
  CRUX_REQUIRE(ctx, B >= 0, "crux_buffer_sample_uniform: negative B");
  if (B == 0) return CRUX_OK;
  CRUX_REQUIRE(ctx, s->elements > 0, "crux_buffer_sample_uniform: empty source");
  int rc = ensure_indices(t, B); if (rc) return rc;
  if (ids_in_host) {
    for (int64_t j = 0; j < B; ++j) CRUX_REQUIRE(ctx, ids_in_host[j] >= 0 && ids_in_host[j] < s->elements, "crux_buffer_sample_uniform: id out of range");
    CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(t->indices, ids_in_host, (size_t)B * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    CRUX_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  } else {
    uniform_ids_kernel<<<(unsigned)cdiv(B, 256), 256, 0, ctx->stream>>>(t->indices, B, s->elements, seed, ctr);
    CRUX_LAUNCHED(ctx);
  }
  t->n_indices = B;
  return crux_buffer_push_from(t, s, B, t->indices);
}

static int ensure_cumsum(crux_buffer *s) {
  crux_ctx *ctx = s->ctx;
  if (s->cumsum_valid) return CRUX_OK;
  const int64_t N = s->elements;
  const int nb = (int)cdiv(N, SCAN_B);
  scan_block_sums_kernel<<<nb, SCAN_T, 0, ctx->stream>>>(s->prs, N, s->block_sums);
  CRUX_LAUNCHED(ctx);
  scan_top_kernel<<<1, SCAN_T, 0, ctx->stream>>>(s->block_sums, nb);
  CRUX_LAUNCHED(ctx);
  scan_apply_kernel<<<nb, SCAN_T, 0, ctx->stream>>>(s->prs, N, s->block_sums, s->cumsum);
  CRUX_LAUNCHED(ctx);
  s->cumsum_valid = true;
  return CRUX_OK;
}

int32_t crux_buffer_sample_prioritized(crux_buffer *t, crux_buffer *s, int64_t B, float beta, int32_t weight_col,
                                       const double *u_in_host, uint64_t seed, uint64_t ctr) {
  if (!t || !s) return CRUX_ERR_INVALID;
  crux_ctx *ctx = t->ctx;
  CRUX_REQUIRE(ctx, s->prioritized, "crux_buffer_sample_prioritized: source is not prioritized");
  CRUX_REQUIRE(ctx, weight_col >= 0 && weight_col < 64 && s->col_index[weight_col] >= 0, "crux_buffer_sample_prioritized: source needs a :weight column (experience_buffer.jl:325)");
  CRUX_REQUIRE(ctx, B >= 0, "crux_buffer_sample_prioritized: negative B");
  if (B == 0) return CRUX_OK;
  CRUX_REQUIRE(ctx, s->elements > 0, "crux_buffer_sample_prioritized: empty source");
  const BufCol &wc = s->cols[s->col_index[weight_col]];
  CRUX_REQUIRE(ctx, wc.dtype == CRUX_F32 && wc.rowlen >= 1, "crux_buffer_sample_prioritized: weight column must be float32");
  int rc = ensure_indices(t, B); if (rc) return rc;
  rc = ensure_cumsum(s); if (rc) return rc;
  const double *u_dev = nullptr;
  if (u_in_host) {
    double *tmp = (double *)crux_scratch(ctx, 5, (size_t)B * sizeof(double));
    if (!tmp) return CRUX_ERR_OOM;
    CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(tmp, u_in_host, (size_t)B * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CRUX_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    u_dev = tmp;
  }
  per_sample_kernel<<<(unsigned)cdiv(B, 128), 128, 0, ctx->stream>>>(s->cumsum, s->prs, s->pstate, s->elements, B, beta, u_dev, seed, ctr,
                                                                     t->indices, (float *)wc.data, wc.rowlen);
  CRUX_LAUNCHED(ctx);
  t->n_indices = B;
  return crux_buffer_push_from(t, s, B, t->indices);
}

int32_t crux_buffer_indices(crux_buffer *b, int32_t **idx_dev, int64_t *n) {
  if (!b) return CRUX_ERR_INVALID;
  if (idx_dev) *idx_dev = b->indices;
  if (n) *n = b->n_indices;
  return CRUX_OK;
}

int32_t crux_buffer_update_priorities(crux_buffer *b, const int32_t *idx_dev, const float *v_dev, int64_t n) {
  if (!b) return CRUX_ERR_INVALID;
  crux_ctx *ctx = b->ctx;
  CRUX_REQUIRE(ctx, b->prioritized, "crux_buffer_update_priorities: buffer is not prioritized");
  CRUX_REQUIRE(ctx, n >= 0, "crux_buffer_update_priorities: negative n");
  if (n == 0) return CRUX_OK;
  CRUX_REQUIRE(ctx, idx_dev && v_dev, "crux_buffer_update_priorities: NULL pointer");
  const unsigned blocks = (unsigned)cdiv(n, 256);
  owner_mark_kernel<<<blocks, 256, 0, ctx->stream>>>(b->owner, idx_dev, n);
  CRUX_LAUNCHED(ctx);
  update_priorities_kernel<<<blocks, 256, 0, ctx->stream>>>(b->prs, b->pstate, b->owner, idx_dev, v_dev, n, b->alpha);
  CRUX_LAUNCHED(ctx);
  owner_reset_kernel<<<blocks, 256, 0, ctx->stream>>>(b->owner, idx_dev, n);
  CRUX_LAUNCHED(ctx);
  b->cumsum_valid = false;
  return CRUX_OK;
}

int32_t crux_buffer_priorities(crux_buffer *b, float **prs_dev, float **cumsum_dev, float *max_p, float *min_p) {
  if (!b) return CRUX_ERR_INVALID;
  crux_ctx *ctx = b->ctx;
  CRUX_REQUIRE(ctx, b->prioritized, "crux_buffer_priorities: buffer is not prioritized");
  if (cumsum_dev && b->elements > 0) { int rc = ensure_cumsum(b); if (rc) return rc; }
  if (prs_dev) *prs_dev = b->prs;
  if (cumsum_dev) *cumsum_dev = b->cumsum;
  if (max_p || min_p) {
    CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(b->pstate_pinned, b->pstate, 2 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CRUX_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (max_p) *max_p = b->pstate_pinned[0];
    if (min_p) *min_p = b->pstate_pinned[1];
  }
  return CRUX_OK;
}

int32_t crux_split_batches(int64_t N, const double *fracs, int32_t n_fracs, int64_t *out) {
  if (!fracs || !out || n_fracs < 1) return CRUX_ERR_INVALID;
  double s = 0.0;
  for (int i = 0; i < n_fracs; ++i) s += fracs[i];
  if (fabs(s - 1.0) > 1.4901161193847656e-08 * fmax(fabs(s), 1.0)) return CRUX_ERR_INVALID;  // @assert sum(fracs) ≈ 1
  int64_t tot = 0;
  for (int i = 0; i < n_fracs; ++i) { out[i] = (int64_t)floor((double)N * fracs[i]); tot += out[i]; }
  out[0] += N - tot;
  return CRUX_OK;
}

int32_t crux_gather_rows(crux_ctx *ctx, void *dst, const void *src, const int32_t *idx_dev, int64_t n, int64_t rowbytes) {
  if (!ctx) return CRUX_ERR_INVALID;
  CRUX_REQUIRE(ctx, n >= 0 && rowbytes >= 0, "crux_gather_rows: negative size");
  if (n == 0 || rowbytes == 0) return CRUX_OK;
  CRUX_REQUIRE(ctx, dst && src && idx_dev, "crux_gather_rows: NULL pointer");
  return launch_copy_rows(ctx, dst, src, idx_dev, 0, n, 0, 1, (size_t)rowbytes, false);
}

}  // extern "C"
