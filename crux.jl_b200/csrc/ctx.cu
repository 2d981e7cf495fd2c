// Context, memory and CUDA-graph plumbing of libcrux_cuda.so.
// Replaces src/devices.jl:1-21 of the reference: nothing hops between host and device per call.
#include "common.cuh"
#include <stdarg.h>

static thread_local std::string g_create_err;

int crux_set_err(crux_ctx *ctx, int code, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf; else g_create_err = buf;
  return code;
}

void *crux_scratch(crux_ctx *ctx, int slot, size_t bytes) {
  if (bytes == 0) bytes = 16;
  if (ctx->scratch_bytes[slot] >= bytes) return ctx->scratch[slot];
  // growing: the old block may still be in use by queued kernels -> stream-ordered free
  if (ctx->scratch[slot]) {
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->scratch[slot]);
    ctx->scratch[slot] = nullptr;
    ctx->scratch_bytes[slot] = 0;
  }
  size_t want = bytes + bytes / 4 + 256;
  void *p = nullptr;
  cudaError_t e = cudaMalloc(&p, want);
  if (e != cudaSuccess) {
    crux_set_err(ctx, CRUX_ERR_OOM, "scratch slot %d: cudaMalloc(%zu): %s", slot, want, cudaGetErrorString(e));
    return nullptr;
  }
  ctx->scratch[slot] = p;
  ctx->scratch_bytes[slot] = want;
  return p;
}

extern "C" {

int32_t crux_abi_version(void) { return CRUX_ABI_VERSION; }

const char *crux_last_error(crux_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int32_t crux_ctx_create(int32_t device, void *stream, crux_ctx **out) {
  if (!out) return crux_set_err(nullptr, CRUX_ERR_INVALID, "crux_ctx_create: out is NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return crux_set_err(nullptr, CRUX_ERR_CUDA, "crux_ctx_create: no CUDA device (%s)",
                        e == cudaSuccess ? "count is 0" : cudaGetErrorString(e));
  if (device < 0 || device >= n) return crux_set_err(nullptr, CRUX_ERR_INVALID, "crux_ctx_create: bad device %d", device);
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return crux_set_err(nullptr, CRUX_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major < 10)
    return crux_set_err(nullptr, CRUX_ERR_CUDA, "crux_ctx_create: device sm_%d%d is not Blackwell (built for sm_100a)",
                        prop.major, prop.minor);
  crux_ctx *ctx = new crux_ctx();
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  if (stream) {
    ctx->stream = (cudaStream_t)stream;
  } else {
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete ctx; return crux_set_err(nullptr, CRUX_ERR_CUDA, "stream: %s", cudaGetErrorString(e)); }
    ctx->own_stream = true;
  }
  cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&ctx->side_fork, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->side_done, cudaEventDisableTiming);
  cudaMalloc((void **)&ctx->flags_dev, sizeof(unsigned int) * 4);
  cudaMemset(ctx->flags_dev, 0, sizeof(unsigned int) * 4);
  cudaMallocHost((void **)&ctx->flags_pinned, sizeof(unsigned int) * 4);
  *out = ctx;
  return CRUX_OK;
}

int32_t crux_ctx_destroy(crux_ctx *ctx) {
  if (!ctx) return CRUX_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (int i = 0; i < 8; ++i) if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
  if (ctx->flags_dev) cudaFree(ctx->flags_dev);
  if (ctx->flags_pinned) cudaFreeHost(ctx->flags_pinned);
  if (ctx->peer_recv) cudaFree(ctx->peer_recv);
  if (ctx->peer_flags) cudaFree(ctx->peer_flags);
  if (ctx->side_stream) { cudaStreamSynchronize(ctx->side_stream); cudaStreamDestroy(ctx->side_stream); }
  if (ctx->side_fork) cudaEventDestroy(ctx->side_fork);
  if (ctx->side_done) cudaEventDestroy(ctx->side_done);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return CRUX_OK;
}

int32_t crux_ctx_set_stream(crux_ctx *ctx, void *stream) {
  if (!ctx) return CRUX_ERR_INVALID;
  if (ctx->own_stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); ctx->own_stream = false; }
  ctx->stream = (cudaStream_t)stream;
  return CRUX_OK;
}

int32_t crux_ctx_stream(crux_ctx *ctx, void **stream_out) {
  if (!ctx || !stream_out) return CRUX_ERR_INVALID;
  *stream_out = (void *)ctx->stream;
  return CRUX_OK;
}

int32_t crux_ctx_sync(crux_ctx *ctx) {
  if (!ctx) return CRUX_ERR_INVALID;
  CRUX_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return CRUX_OK;
}

int32_t crux_ctx_launch_count(crux_ctx *ctx, int64_t *out) {
  if (!ctx || !out) return CRUX_ERR_INVALID;
  *out = ctx->launches;
  return CRUX_OK;
}

int32_t crux_ctx_timing_begin(crux_ctx *ctx) {
  if (!ctx) return CRUX_ERR_INVALID;
  for (auto e : ctx->t_start) cudaEventDestroy(e);
  for (auto e : ctx->t_stop) cudaEventDestroy(e);
  ctx->t_start.clear(); ctx->t_stop.clear(); ctx->t_family.clear();
  ctx->timing = true;
  return CRUX_OK;
}

int32_t crux_ctx_timing_end(crux_ctx *ctx, float *ms_out_host, int32_t *count_out_host) {
  if (!ctx || !ms_out_host || !count_out_host) return CRUX_ERR_INVALID;
  ctx->timing = false;
  CRUX_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int f = 0; f < CRUX_T_FAMILIES; ++f) { ms_out_host[f] = 0.f; count_out_host[f] = 0; }
  for (size_t i = 0; i < ctx->t_start.size(); ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->t_start[i], ctx->t_stop[i]) == cudaSuccess) {
      ms_out_host[ctx->t_family[i]] += ms; count_out_host[ctx->t_family[i]] += 1;
    }
    cudaEventDestroy(ctx->t_start[i]); cudaEventDestroy(ctx->t_stop[i]);
  }
  ctx->t_start.clear(); ctx->t_stop.clear(); ctx->t_family.clear();
  return CRUX_OK;
}

int32_t crux_ctx_check(crux_ctx *ctx) {
  if (!ctx) return CRUX_ERR_INVALID;
  CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(ctx->flags_pinned, ctx->flags_dev, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
  CRUX_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->flags_pinned[0] & CRUX_FLAG_NAN) {
    cudaMemsetAsync(ctx->flags_dev, 0, sizeof(unsigned int), ctx->stream);
    return crux_set_err(ctx, CRUX_ERR_NAN, "NaN detected! (gradient norm or advantage; training.jl:20 / sampler.jl:270)");
  }
  return CRUX_OK;
}

int32_t crux_dev_alloc(crux_ctx *ctx, size_t bytes, void **out) {
  if (!ctx || !out) return CRUX_ERR_INVALID;
  cudaError_t e = cudaMalloc(out, bytes ? bytes : 16);
  if (e != cudaSuccess) return crux_set_err(ctx, CRUX_ERR_OOM, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
  return CRUX_OK;
}
int32_t crux_dev_free(crux_ctx *ctx, void *ptr) {
  if (!ctx) return CRUX_ERR_INVALID;
  if (ptr) { cudaStreamSynchronize(ctx->stream); CRUX_CHECK_CUDA(ctx, cudaFree(ptr)); }
  return CRUX_OK;
}
int32_t crux_pinned_alloc(crux_ctx *ctx, size_t bytes, void **out_host) {
  if (!ctx || !out_host) return CRUX_ERR_INVALID;
  cudaError_t e = cudaMallocHost(out_host, bytes ? bytes : 16);
  if (e != cudaSuccess) return crux_set_err(ctx, CRUX_ERR_OOM, "cudaMallocHost(%zu): %s", bytes, cudaGetErrorString(e));
  return CRUX_OK;
}
int32_t crux_pinned_free(crux_ctx *ctx, void *ptr_host) {
  if (!ctx) return CRUX_ERR_INVALID;
  if (ptr_host) CRUX_CHECK_CUDA(ctx, cudaFreeHost(ptr_host));
  return CRUX_OK;
}
int32_t crux_memcpy_h2d(crux_ctx *ctx, void *dst, const void *src_host, size_t bytes) {
  if (!ctx) return CRUX_ERR_INVALID;
  if (bytes) CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(dst, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return CRUX_OK;
}
int32_t crux_memcpy_d2h(crux_ctx *ctx, void *dst_host, const void *src, size_t bytes) {
  if (!ctx) return CRUX_ERR_INVALID;
  if (bytes) CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(dst_host, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return CRUX_OK;
}
int32_t crux_memcpy_d2d(crux_ctx *ctx, void *dst, const void *src, size_t bytes) {
  if (!ctx) return CRUX_ERR_INVALID;
  if (bytes) CRUX_CHECK_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  return CRUX_OK;
}
int32_t crux_memset(crux_ctx *ctx, void *dst, int32_t byte, size_t bytes) {
  if (!ctx) return CRUX_ERR_INVALID;
  if (bytes) CRUX_CHECK_CUDA(ctx, cudaMemsetAsync(dst, byte, bytes, ctx->stream));
  return CRUX_OK;
}

int32_t crux_graph_begin(crux_ctx *ctx) {
  if (!ctx) return CRUX_ERR_INVALID;
  CRUX_CHECK_CUDA(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
  return CRUX_OK;
}
int32_t crux_graph_end(crux_ctx *ctx, void **graph_exec_out) {
  if (!ctx || !graph_exec_out) return CRUX_ERR_INVALID;
  cudaGraph_t g = nullptr;
  CRUX_CHECK_CUDA(ctx, cudaStreamEndCapture(ctx->stream, &g));
  cudaGraphExec_t ge = nullptr;
  cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) return crux_set_err(ctx, CRUX_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
  *graph_exec_out = (void *)ge;
  return CRUX_OK;
}
int32_t crux_graph_launch(crux_ctx *ctx, void *graph_exec) {
  if (!ctx || !graph_exec) return CRUX_ERR_INVALID;
  CRUX_CHECK_CUDA(ctx, cudaGraphLaunch((cudaGraphExec_t)graph_exec, ctx->stream));
  return CRUX_OK;
}
int32_t crux_graph_destroy(crux_ctx *ctx, void *graph_exec) {
  if (!ctx) return CRUX_ERR_INVALID;
  if (graph_exec) CRUX_CHECK_CUDA(ctx, cudaGraphExecDestroy((cudaGraphExec_t)graph_exec));
  return CRUX_OK;
}

}  // extern "C"
